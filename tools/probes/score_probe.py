#!/usr/bin/env python
"""Tuning probe: time ONLY the score pass (spagxe_score_stats) of the probe build for a list of kernel variants on one
1.024 GB batch (N = 500,000 x 8,192 SNPs, the bench recipe).  Several variants give wrong sums by design, so nothing
downstream of the score pass runs here.  Never a reported number.
    usage: python tools/probes/score_probe.py 0 51 52 53 54"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import torch

import bench
from spagxecct_b200 import Context, _lib


def main():
    cfgs = [int(a) for a in sys.argv[1:]] or [0]
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libspagxe_cuda_probes.so")
    n, m = 500000, 8192
    dev = torch.device("cuda:0")
    pitch = ((n + 3) // 4 + 127) // 128 * 128
    bed = bench.make_batch_device(n, m, 0, dev, pitch)
    g = torch.Generator(device=dev); g.manual_seed(7)
    R = torch.randn(n, generator=g, device=dev, dtype=torch.float64) * 0.1
    E = torch.randn(n, generator=g, device=dev, dtype=torch.float64)
    ctx = Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_vectors_device(n, R.data_ptr(), E.data_ptr())
    geno = ctx.geno_desc(_lib.GENO_BED, bed.data_ptr(), n, m, pitch, _lib.LOC_DEVICE)
    ref = None
    for c in cfgs:
        ctx.set_option(100, c)
        for _ in range(3):
            st = ctx.score_stats(geno)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            st = ctx.score_stats(geno)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        if ref is None:
            ref = st
        same = all(np.array_equal(st[k], ref[k]) for k in ref)
        print(f"cfg {c:3d}: {ms:.4f} ms per 1.024 GB call (score kernel + reduce + 200 KB D2H), "
              f"{1.024 / ms:.3f} TB/s, sums {'== cfg ' + str(cfgs[0]) if same else 'DIFFER (expected for timing probes)'}", flush=True)


if __name__ == "__main__":
    main()
