"""PLINK .bed/.bim/.fam access for the GenoFile path (what the reference delegates to
GRAB::GRAB.ReadGeno at R/SPAGxECCT.R:484-487).  Only the container format is handled here on the
host (memory-map, header check, sample/marker ids); the 2-bit decode itself runs on the GPU."""
import os

import numpy as np

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])  # variant-major


class BedFile:
    def __init__(self, bed_path, bim_path=None, fam_path=None):
        stem = bed_path[:-4] if bed_path.endswith(".bed") else bed_path
        self.bed_path = stem + ".bed"
        self.bim_path = bim_path or stem + ".bim"
        self.fam_path = fam_path or stem + ".fam"
        self.sample_ids = [l.split()[1] for l in open(self.fam_path)]
        bim = [l.split() for l in open(self.bim_path)]
        self._bim = bim
        self.snp_ids = [b[1] for b in bim]
        self.n = len(self.sample_ids)
        self.m = len(self.snp_ids)
        self.row_bytes = (self.n + 3) // 4
        size = os.path.getsize(self.bed_path)
        if size != 3 + self.row_bytes * self.m:
            raise ValueError(f"{self.bed_path}: size {size} does not match {self.m} x ceil({self.n}/4) + 3")
        mm = np.memmap(self.bed_path, dtype=np.uint8, mode="r")
        if bytes(mm[:3]) != BED_MAGIC:
            raise ValueError(f"{self.bed_path}: not a variant-major PLINK .bed file")
        self.rows = mm[3:]  # M * row_bytes packed bytes, zero-copy

    def bim_chrom_pos(self):
        """(chromosome strings, base-pair positions) of the markers, for range filters."""
        return np.asarray([b[0] for b in self._bim]), np.asarray([int(b[3]) for b in self._bim], dtype=np.int64)

    def rows_array(self):
        return np.ascontiguousarray(self.rows)


def from_bytes(bed_bytes, n):
    """(packed rows, M) from an in-memory .bed image (with the 3-byte magic)."""
    b = np.frombuffer(bytes(bed_bytes[:3]), dtype=np.uint8)
    if bytes(b) != BED_MAGIC:
        raise ValueError("not a variant-major PLINK .bed image")
    rb = (n + 3) // 4
    rows = np.ascontiguousarray(np.asarray(bed_bytes, dtype=np.uint8)[3:])
    if rows.size % rb:
        raise ValueError("size mismatch")
    return rows, rows.size // rb


def pack_rows(G):
    """Host-side packer for tests/fixtures: (N, M) array with NaN/-1 = missing -> packed rows.
    A1-count convention: 2 -> 00, NA -> 01, 1 -> 10, 0 -> 11; padding bits are 00 like PLINK."""
    G = np.asarray(G)
    n, m = G.shape
    code = np.full((n, m), 1, dtype=np.uint8)
    with np.errstate(invalid="ignore"):
        code[G == 2] = 0
        code[G == 1] = 2
        code[G == 0] = 3
    rb = (n + 3) // 4
    pad = np.zeros((rb * 4 - n, m), dtype=np.uint8)
    c = np.vstack([code, pad]).reshape(rb, 4, m)
    packed = (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)
    return np.ascontiguousarray(packed.T).ravel()  # variant-major


def reorder_samples(rows, n_file, m, index):
    """Host gather used only when Phen.mtx$ID is not in .fam order (GRAB's SampleIDs reorder).
    index[i] = position in the file of output sample i."""
    rb = (n_file + 3) // 4
    r = np.asarray(rows, dtype=np.uint8).reshape(m, rb)
    idx = np.asarray(index, dtype=np.int64)
    code = (r[:, idx >> 2] >> (2 * (idx & 3)).astype(np.uint8)) & 3  # (m, n_out)
    n_out = idx.size
    rbo = (n_out + 3) // 4
    pad = np.zeros((m, rbo * 4 - n_out), dtype=np.uint8)
    c = np.hstack([code, pad]).reshape(m, rbo, 4)
    out = (c[:, :, 0] | (c[:, :, 1] << 2) | (c[:, :, 2] << 4) | (c[:, :, 3] << 6)).astype(np.uint8)
    return np.ascontiguousarray(out).ravel()
