"""Minimal reader for R's XDR serialisation (.rda / .RData, RDX2 / RDX3).

Test / fixture tooling only: it is used by tools/make_golden.py (run in the build
container, where /root/reference is mounted) to turn the reference's bundled data
objects (data/*.rda, simulation_studies/*/data/*.RData) into small numpy fixtures
under tests/golden/.  Handles SEXP types 2 (pairlist), 9 (CHARSXP), 10 (LGLSXP),
13 (INTSXP), 14 (REALSXP), 16 (STRSXP), 19 (VECSXP), 1 (SYMSXP), 22 (EXTPTR-free
subset), 238 (ALTREP compact seq / wrap), 254 (NILVALUE), 255 (REFSXP), attributes.
"""
import bz2
import gzip
import lzma
import struct

import numpy as np


def _decompress(raw: bytes) -> bytes:
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.p = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def nbytes(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def length(self):
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == 254 or t == 253 or t == 242:  # NILVALUE, GLOBALENV, EMPTYENV
            return None
        if t == 255:  # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 2 or t == 6:  # LISTSXP / LANGSXP -> python list of (tag, value)
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                del attr
                flags = self.i32()
                t2 = flags & 0xFF
                if t2 == 254:
                    break
                if t2 != 2 and t2 != 6:
                    raise ValueError(f"unexpected cdr type {t2}")
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return out
        if t == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            return self.nbytes(n).decode("utf-8", "replace")
        if t == 238:  # ALTREP
            info = self.item()
            state = self.item()
            attr = self.item()
            cls = info[0][1] if isinstance(info, list) else None
            val = self._altrep(cls, state)
            return self._with_attr(val, attr)
        if t in (10, 13):
            n = self.length()
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.p).astype(np.int32)
            self.p += 4 * n
            val = v
        elif t == 14:
            n = self.length()
            v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)
            self.p += 8 * n
            val = v
        elif t == 16:
            n = self.length()
            val = [self.item() for _ in range(n)]
        elif t == 19 or t == 20:
            n = self.length()
            val = [self.item() for _ in range(n)]
        elif t == 22:  # EXTPTRSXP: prot, tag
            self.refs.append(None)
            self.item(); self.item()
            val = None
        else:
            raise ValueError(f"unsupported SEXP type {t} at {self.p}")
        attr = self.item() if has_attr else None
        return self._with_attr(val, attr)

    @staticmethod
    def _altrep(cls, state):
        if cls == "compact_intseq":
            n, start, inc = (int(x) for x in state)
            return (start + inc * np.arange(n)).astype(np.int32)
        if cls == "compact_realseq":
            n, start, inc = state
            return start + inc * np.arange(int(n), dtype=np.float64)
        if cls in ("wrap_real", "wrap_integer", "wrap_string", "wrap_logical"):
            return state[0][1] if isinstance(state, list) and isinstance(state[0], tuple) else state[0]
        if cls == "deferred_string":
            v = state[0][1]
            base = v.value if isinstance(v, RObj) else v
            return [str(int(x)) if float(x).is_integer() else repr(float(x)) for x in base]
        raise ValueError(f"unsupported ALTREP class {cls}")

    @staticmethod
    def _with_attr(val, attr):
        if attr:
            return RObj(val, {k: v for k, v in attr})
        return val


class RObj:
    """A value with R attributes (names, dim, dimnames, class, row.names, levels...)."""

    def __init__(self, value, attrs):
        self.value = value
        self.attrs = attrs

    def __repr__(self):
        return f"RObj({type(self.value).__name__}, attrs={list(self.attrs)})"


def _plain(x):
    return x.value if isinstance(x, RObj) else x


def load_rdata(path):
    """Return {name: object}.  data.frames -> dict(column -> ndarray/list);
    matrices -> 2-D ndarray (column-major reshaped) with .dimnames kept in RObj."""
    buf = _decompress(open(path, "rb").read())
    if buf[:5] not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError("not an RDX2/RDX3 file")
    r = _Reader(buf)
    r.p = 5
    fmt = r.nbytes(2)
    if fmt != b"X\n":
        raise ValueError("only XDR format supported")
    version = r.i32(); r.i32(); r.i32()
    if version == 3:
        n = r.i32(); r.nbytes(n)
    top = r.item()
    return {tag: convert(val) for tag, val in top}


def convert(x):
    if not isinstance(x, RObj):
        return x
    a = x.attrs
    cls = _plain(a.get("class"))
    if cls and "data.frame" in cls:
        names = _plain(a["names"])
        cols = {}
        for nm, col in zip(names, x.value):
            if isinstance(col, RObj) and "levels" in col.attrs:
                lv = _plain(col.attrs["levels"])
                cols[nm] = [lv[i - 1] for i in col.value]
            else:
                cols[nm] = _plain(col)
        return cols
    if "dim" in a:
        dim = tuple(int(d) for d in _plain(a["dim"]))
        arr = np.asarray(x.value).reshape(dim, order="F")
        out = RObj(arr, {})
        if "dimnames" in a:
            out.attrs["dimnames"] = [_plain(d) for d in _plain(a["dimnames"])]
        return out
    if isinstance(x.value, list) and "names" in a:
        return {k: convert(v) for k, v in zip(_plain(a["names"]), x.value)}
    return x.value
