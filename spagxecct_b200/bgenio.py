"""BGEN reader for the GenoFile argument (the GRAB::GRAB.ReadGeno step of R/SPAGxECCT.R:484-487 for `.bgen` input).

Host-side data format code, not the hot path: a .bgen file holds compressed genotype probabilities, so it is decoded
on the host into the N x M dosage matrix the reference's loop would receive, and that matrix goes through the C ABI
(hard-call-valued dosages take the 2-bit pipeline, non-integer ones the dense dosage kernel).

Supported: BGEN v1.1 (layout 1: zlib, 16-bit probability triples) and v1.2 (layout 2: uncompressed or zlib, 1..32 bits
per probability, phased or unphased, biallelic, ploidy <= 2); zstd-compressed files are refused (no decoder in the
image).  Dosage convention of GRAB for BGEN input: AlleleOrder = "ref-first" (the default) counts the SECOND allele,
"alt-first" the first; samples with a missing call are NaN.
"""
import os
import struct
import zlib

import numpy as np


def _rd(f, fmt):
    return struct.unpack(fmt, f.read(struct.calcsize(fmt)))


def _rd_str(f, fmt):
    (ln,) = _rd(f, fmt)
    return f.read(ln).decode()


class BgenFile:
    def __init__(self, path, sample_path=None):
        self.path = path
        with open(path, "rb") as f:
            (self.offset,) = _rd(f, "<I")
            hl, self.m, self.n = _rd(f, "<III")
            if f.read(4) not in (b"bgen", b"\0\0\0\0"):
                raise ValueError(f"{path}: not a BGEN file")
            f.read(hl - 20)
            (flags,) = _rd(f, "<I")
            self.compression = flags & 3
            self.layout = (flags >> 2) & 15
            has_ids = bool(flags >> 31)
            if self.layout not in (1, 2):
                raise NotImplementedError(f"{path}: BGEN layout {self.layout}")
            if self.compression == 2:
                raise NotImplementedError(f"{path}: zstd-compressed BGEN (no zstd decoder in this environment)")
            self.sample_ids = None
            if has_ids:
                _, ns = _rd(f, "<II")
                self.sample_ids = [_rd_str(f, "<H") for _ in range(ns)]
            # variant index: one sequential scan over the identifying blocks (the .bgi file is not needed)
            f.seek(4 + self.offset)
            self.snp_ids, self.rsids, self.chrom, self.pos, self.alleles, self._blocks = [], [], [], [], [], []
            for _ in range(self.m):
                if self.layout == 1:
                    _rd(f, "<I")
                vid = _rd_str(f, "<H"); rsid = _rd_str(f, "<H"); chrom = _rd_str(f, "<H")
                (pos,) = _rd(f, "<I")
                k = _rd(f, "<H")[0] if self.layout == 2 else 2
                al = [_rd_str(f, "<I") for _ in range(k)]
                if self.layout == 1:
                    ln = _rd(f, "<I")[0] if self.compression else 6 * self.n
                else:
                    (ln,) = _rd(f, "<I")
                self._blocks.append((f.tell(), ln))
                f.seek(ln, os.SEEK_CUR)
                self.snp_ids.append(rsid if rsid else vid); self.rsids.append(rsid)
                self.chrom.append(chrom); self.pos.append(pos); self.alleles.append(al)
        if sample_path is not None and os.path.exists(sample_path):
            with open(sample_path) as sf:
                lines = [l.split() for l in sf if l.strip()]
            self.sample_ids = [l[1] if len(l) > 1 else l[0] for l in lines[2:]]     # Oxford .sample: two header lines, ID_2
        if self.sample_ids is None:
            self.sample_ids = [f"sample_{i + 1}" for i in range(self.n)]

    def bim_chrom_pos(self):
        return np.asarray(self.chrom), np.asarray(self.pos, dtype=np.int64)

    # ---- probability blocks -> dosage of the second allele ----
    def _dosage_layout1(self, raw):
        p = np.frombuffer(raw, dtype="<u2", count=3 * self.n).reshape(self.n, 3).astype(np.float64) / 32768.0
        d = p[:, 1] + 2 * p[:, 2]
        d[p.sum(1) == 0] = np.nan
        return d

    def _dosage_layout2(self, raw):
        n, k, pmin, pmax = struct.unpack("<IHBB", raw[:8])
        if n != self.n:
            raise ValueError(f"{self.path}: block with {n} samples, header says {self.n}")
        if k != 2:
            raise NotImplementedError(f"{self.path}: multi-allelic variant ({k} alleles)")
        pl = np.frombuffer(raw, dtype=np.uint8, count=n, offset=8)
        missing = (pl & 0x80) != 0
        ploidy = (pl & 0x3F).astype(np.int64)
        phased, bits = raw[8 + n], raw[9 + n]
        if pmax > 2:
            raise NotImplementedError(f"{self.path}: ploidy {pmax}")
        nval = ploidy.copy()                     # stored values per sample: unphased diploid 2 (AA, AB), haploid 1; phased: ploidy
        body = np.frombuffer(raw, dtype=np.uint8, offset=10 + n)
        tot = int(nval.sum())
        if bits == 8:
            vals = body[:tot].astype(np.float64)
        elif bits == 16:
            vals = np.frombuffer(body[:2 * tot].tobytes(), dtype="<u2").astype(np.float64)
        else:                                     # general width: little-endian bit stream
            bitarr = np.unpackbits(body, bitorder="little")[: tot * bits].reshape(tot, bits)
            vals = (bitarr.astype(np.float64) * (2.0 ** np.arange(bits))).sum(1)
        vals /= float((1 << bits) - 1)
        start = np.concatenate([[0], np.cumsum(nval)[:-1]])
        d = np.full(n, np.nan)
        if phased:
            # per haplotype P(allele 1); dosage of allele 2 = sum over haplotypes of (1 - P(allele 1))
            h1 = vals[start]
            two = ploidy == 2
            d[:] = 1 - h1
            d[two] = (1 - h1[two]) + (1 - vals[start[two] + 1])
        else:
            two = ploidy == 2; one = ploidy == 1
            paa = vals[start[two]]; pab = vals[start[two] + 1]
            d[two] = pab + 2 * (1 - paa - pab)
            d[one] = 1 - vals[start[one]]
        d[missing | (ploidy == 0)] = np.nan
        return d

    def read_dosages(self, sel=None, allele_order="ref-first"):
        """(N, M') float64, column-major: dosage of the second allele ("ref-first", GRAB's default for BGEN) or of the
        first ("alt-first") of the selected variants; NaN = missing."""
        idx = range(self.m) if sel is None else list(sel)
        out = np.empty((self.n, len(idx)), dtype=np.float64, order="F")
        with open(self.path, "rb") as f:
            for c, j in enumerate(idx):
                off, ln = self._blocks[j]
                f.seek(off)
                if self.layout == 1:
                    raw = f.read(ln)
                    if self.compression:
                        raw = zlib.decompress(raw)
                    d = self._dosage_layout1(raw)
                else:
                    if self.compression:
                        (_,) = _rd(f, "<I")
                        raw = zlib.decompress(f.read(ln - 4))
                    else:
                        raw = f.read(ln)
                    d = self._dosage_layout2(raw)
                out[:, c] = (2 - d) if allele_order == "alt-first" else d
        return out


def write_bgen(path, dosage_probs, sample_ids, snp_ids, alleles=("G", "A"), bits=8, compress=True):
    """Minimal BGEN v1.2 writer for tests: dosage_probs (N, M, 2) = P(AA), P(AB) per sample (NaN row = missing),
    unphased diploid biallelic."""
    P = np.asarray(dosage_probs, dtype=np.float64)
    n, m, _ = P.shape
    ids_block = b"".join(struct.pack("<H", len(s)) + s.encode() for s in sample_ids)
    ids_block = struct.pack("<II", 8 + len(ids_block), n) + ids_block
    flags = (1 if compress else 0) | (2 << 2) | (1 << 31)
    header = struct.pack("<III", 20, m, n) + b"bgen" + struct.pack("<I", flags)
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(header) + len(ids_block)))
        f.write(header); f.write(ids_block)
        scale = (1 << bits) - 1
        for j in range(m):
            sid = snp_ids[j].encode()
            f.write(struct.pack("<H", 0) + struct.pack("<H", len(sid)) + sid + struct.pack("<H", 1) + b"1" + struct.pack("<I", j + 1))
            f.write(struct.pack("<H", 2))
            for a in alleles:
                f.write(struct.pack("<I", len(a)) + a.encode())
            miss = np.isnan(P[:, j, 0])
            q = np.rint(np.nan_to_num(P[:, j, :]) * scale).astype(np.uint64)
            pl = np.where(miss, 0x82, 0x02).astype(np.uint8)
            if bits == 8:
                body = q.astype(np.uint8).tobytes()
            elif bits == 16:
                body = q.astype("<u2").tobytes()
            else:
                bitarr = ((q.reshape(-1, 1) >> np.arange(bits, dtype=np.uint64)) & 1).astype(np.uint8).ravel()
                body = np.packbits(bitarr, bitorder="little").tobytes()
            raw = struct.pack("<IHBB", n, 2, 2, 2) + pl.tobytes() + bytes([0, bits]) + body
            if compress:
                z = zlib.compress(raw)
                f.write(struct.pack("<II", len(z) + 4, len(raw)) + z)
            else:
                f.write(struct.pack("<I", len(raw)) + raw)
