#!/usr/bin/env python
"""bench.py -- SNPs/sec of the SPAGxE_CCT Step-2 loop at N = 500,000 samples (BASELINE.json configs[2]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port, all host threads)

A "step" is one pass of the hot path (2-bit unpack + score + routing + SPA + genotype-adjusted
branch) over one batch of synthetic UK-Biobank-shaped SNPs that is already resident in HBM.  The
batch (default 16,384 SNPs x 500,000 samples = 2.05 GB packed) is far larger than the 126 MB L2, so
no flush is needed between iterations.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES = 500_000
PREVALENCE = 0.01
SEED_PHENO = 12345
SEED_GENO = 20251019


def make_pheno(n=N_SAMPLES):
    """Recipe of simulation_studies/SPAGxECCT/code/SPAGxECCT_binary_typeIerror_simulation_normal_envi.R:30-40
    at 1% prevalence; Step-1 residuals from the harness-side logistic fit (SPA_G_Get_Resid stays in R)."""
    from spagxecct_b200.harness import get_resid
    rng = np.random.default_rng(SEED_PHENO)
    X1 = rng.standard_normal(n); X2 = rng.binomial(1, 0.5, n).astype(np.float64); E = rng.standard_normal(n)
    eta = 0.5 * X1 + 0.5 * X2 + 0.5 * E
    b0 = np.log(PREVALENCE / (1 - PREVALENCE)) - 0.5
    Y = rng.binomial(1, 1 / (1 + np.exp(-(b0 + eta)))).astype(np.float64)
    R = get_resid("binary", Y, [X1, X2, E])
    return dict(Y=Y, E=E, Cova=np.column_stack([X1, X2]), R=R)


def make_batch_device(n, m, snp0, device, pitch):
    """Packed PLINK rows generated on the device: p_j ~ U(0.001, 0.5), g ~ Binomial(2, p_j), counted allele
    A1 for even j / A2 for odd j (exercises the MAF > 0.5 flip), missing rate 0 for half of the SNPs and
    U(0, 0.05) for the rest.  Returns a uint8 tensor [m, pitch] (padding = 0xFF)."""
    import torch
    g = torch.Generator(device=device)
    out = torch.full((m, pitch), 0xFF, dtype=torch.uint8, device=device)
    n4 = (n + 3) // 4
    step = 128
    for j0 in range(0, m, step):
        mc = min(step, m - j0)
        g.manual_seed(SEED_GENO + snp0 + j0)
        p = torch.rand((mc, 1), generator=g, device=device) * 0.499 + 0.001
        jj = torch.arange(snp0 + j0, snp0 + j0 + mc, device=device).view(mc, 1)
        has_miss = torch.rand((mc, 1), generator=g, device=device) < 0.5
        mrate = torch.rand((mc, 1), generator=g, device=device) * 0.05 * has_miss
        u = torch.rand((mc, n4 * 4), generator=g, device=device)
        dos = (u < p).to(torch.uint8)
        u = torch.rand((mc, n4 * 4), generator=g, device=device)
        dos += (u < p).to(torch.uint8)                        # minor-allele count 0/1/2
        dos = torch.where((jj % 2) == 1, 2 - dos, dos)        # odd SNPs count the major allele
        code = torch.where(dos == 2, 0, torch.where(dos == 1, 2, 3)).to(torch.uint8)  # A1 count -> PLINK code
        u = torch.rand((mc, n4 * 4), generator=g, device=device)
        code = torch.where(u < mrate, torch.tensor(1, dtype=torch.uint8, device=device), code)
        if n4 * 4 > n:
            code[:, n:] = 3
        c = code.view(mc, n4, 4)
        out[j0:j0 + mc, :n4] = c[:, :, 0] | (c[:, :, 1] << 2) | (c[:, :, 2] << 4) | (c[:, :, 3] << 6)
        del u, dos, code, c
    return out


class ClockSampler:
    def __init__(self, device_index):
        self.idx = device_index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def hbm_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def oracle_sample_rate(ph, rows_host, n, nthreads):
    """reference CPU path (oracle port of SPAGxE_CCT_one_SNP) on a bounded sample of the same batch."""
    from oracle import oracle as O
    O.build()
    t0 = time.perf_counter()
    out, route, iters, rc = O.cct_bed("binary", rows_host, n, ph["R"], ph["E"], ph["Y"], ph["Cova"], nthreads=nthreads)
    dt = time.perf_counter() - t0
    m = rows_host.size // ((n + 3) // 4)
    return m / dt, dt, m


def workload_config(n, m):
    rb = (n + 3) // 4
    return {"workload": f"C3 slice: SPAGxE_CCT binary trait 1% prevalence, N={n}, {m} SNPs/step/GPU "
                        f"({m * rb / 1e9:.2f} GB packed .bed rows resident in HBM, > L2 so no flush), MAF~U(0.001,0.5), "
                        "50% of SNPs with U(0,5%) missing, odd SNPs count the major allele",
            "n_samples": n, "snps_per_step_per_gpu": m, "l2_policy": "inputs larger than L2",
            "sharding": "SNP ranges per rank, one NCCL broadcast of R/E, no other collective"}


def run_reference(args):
    """The reference's CPU path for the same metric: the C restatement of SPAGxE_CCT_one_SNP (R is not installed, and
    the reference has no native code to compile) on all host cores, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.n_samples
    ph = make_pheno(n)
    cores = os.cpu_count() or 1
    per_step = args.ref_snps or 8 * cores
    rng = np.random.default_rng(SEED_GENO)
    from spagxecct_b200.bedio import pack_rows
    p = rng.uniform(0.001, 0.5, per_step)
    G = (rng.random((n, per_step)) < p).astype(np.int8) + (rng.random((n, per_step)) < p).astype(np.int8)
    G[:, 1::2] = 2 - G[:, 1::2]
    mr = rng.uniform(0, 0.05, per_step) * (rng.random(per_step) < 0.5)
    G[rng.random((n, per_step)) < mr] = -1
    rows = pack_rows(G)
    del G
    times, total = [], 0
    for it in range(args.warmup + args.steps):
        rate, dt, m = oracle_sample_rate(ph, rows, n, cores)
        if it >= args.warmup:
            times.append(dt); total += m
    T = sum(times)
    value = total / T
    cfg = workload_config(n, args.batch_snps)
    cfg["reference_sample"] = f"{per_step} SNPs of the same recipe per step (bounded sample of the {args.batch_snps}-SNP batch)"
    line = {
        "impl": "reference", "metric": "SNPs/sec at N=500k (SPAGxE_CCT Step-2 loop, device-timed)", "value": value,
        "unit": "SNPs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * T / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "SNPs/s", "cores": cores, "kind": "port",
                         "sample": f"{total} SNP evaluations ({per_step} distinct SNPs) of the C3 recipe at N={n}; restated "
                                   "reference (C oracle, -O2, one thread per core), not the R interpreter"},
        "e2e": {"value": value, "unit": "SNPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-samples", type=int, default=N_SAMPLES)
    ap.add_argument("--batch-snps", type=int, default=16384)
    ap.add_argument("--ref-snps", type=int, default=0)
    ap.add_argument("--cpu-sample-snps", type=int, default=1536)   # ~15 s of one host core at N = 500k
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from spagxecct_b200 import Context, _lib

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (spagxecct_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, m = args.n_samples, args.batch_snps
    W, K = max(3, args.warmup), args.steps

    # ---- inputs: rank 0 owns Step-1 output; ONE NCCL broadcast distributes R and E (north_star) ----
    from spagxecct_b200.sharded import broadcast_inputs
    vec = torch.empty((2, n), dtype=torch.float64, device=dev)
    ycov = torch.empty((3, n), dtype=torch.float64, device=dev)   # Wald data (Y, Cova) rides along
    ph = None
    if rank == 0:
        ph = make_pheno(n)
        vec[0] = torch.from_numpy(ph["R"]).to(dev); vec[1] = torch.from_numpy(ph["E"]).to(dev)
        ycov[0] = torch.from_numpy(ph["Y"]).to(dev); ycov[1:] = torch.from_numpy(ph["Cova"].T.copy()).to(dev)
    if world > 1:
        vec, ycov = broadcast_inputs(dist, [vec, ycov])           # the ONE collective of the run (NCCL)
        vec = vec.contiguous(); ycov = ycov.contiguous()
    Y = ycov[0].cpu().numpy(); cova = ycov[1:].cpu().numpy().T.copy()
    ctx = Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_vectors_device(n, vec[0].data_ptr(), vec[1].data_ptr())
    ctx.set_wald(_lib.TRAIT_BINARY, Y, cova)

    rb = (n + 3) // 4
    pitch = (rb + 127) // 128 * 128
    bed = make_batch_device(n, m, rank * m, dev, pitch)       # each rank tests its own SNP shard
    out_dev = torch.empty((m, 10), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    geno_dev = ctx.geno_desc(_lib.GENO_BED, bed.data_ptr(), n, m, pitch, _lib.LOC_DEVICE)
    prm_dev = _lib.make_params(out_location=_lib.LOC_DEVICE)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) ----
    diags = []
    sampler = ClockSampler(local); sampler.start()
    for _ in range(W):
        ctx.run_cct(geno_dev, prm_dev, out_ptr=out_dev.data_ptr())
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        _, dg = ctx.run_cct(geno_dev, prm_dev, out_ptr=out_dev.data_ptr())
        diags.append(dg)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * m * K / (ms_max * 1e-3)

    # ---- end-to-end through the C ABI with HOST buffers (pinned), H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        host_rows = torch.empty((m, rb), dtype=torch.uint8).pin_memory()
        host_rows.copy_(bed[:, :rb])
        host_out = torch.empty((10, m), dtype=torch.float64).pin_memory()
        torch.cuda.synchronize()
        geno_host = ctx.geno_desc(_lib.GENO_BED, host_rows.data_ptr(), n, m, rb, _lib.LOC_HOST)
        prm_host = _lib.make_params()
        ctx.run_cct(geno_host, prm_host, out_ptr=host_out.data_ptr())
        barrier()
        e2 = torch.cuda.Event(enable_timing=True); e3 = torch.cuda.Event(enable_timing=True)
        e2.record()
        ke = max(2, min(K, 3))
        for _ in range(ke):
            _, dge = ctx.run_cct(geno_host, prm_host, out_ptr=host_out.data_ptr())
        e3.record()
        barrier()
        ms_e = e2.elapsed_time(e3)
        te = torch.tensor([ms_e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * m * ke / (float(te.item()) * 1e-3), "unit": "SNPs/s",
               "h2d_bytes_per_step": int(dge["h2d_bytes"]), "d2h_bytes_per_step": int(dge["d2h_bytes"]), "steps": ke}
        # e2e output must equal the device-resident output
        same = np.array_equal(np.nan_to_num(host_out.numpy().T, nan=-1), np.nan_to_num(out_dev.cpu().numpy().reshape(10, m).T, nan=-1))
        e2e["matches_device_path"] = bool(same)

    if rank == 0:
        peak, peak_src = hbm_peak()
        ms_score = float(np.mean([d["ms_score"] for d in diags]))
        score_bytes = float(diags[-1]["score_bytes"])
        achieved = score_bytes / (ms_score * 1e-3) / 1e9
        # one score launch per ~1 GiB chunk of packed rows (csrc/spagxe_cuda.cu make_plan)
        chunk_rows = max(128, ((1 << 30) // pitch) // 128 * 128)
        n_launch = -(-m // min(chunk_rows, 65536))
        # DRAM traffic of the same kernel from the committed `ncu --set full` capture (profiles/r01_ncu_k_score_tc3_final.txt:
        # 1.070415 GB read + 7.95 MB written for a launch of 1.024 GB of packed rows), scaled to this launch size
        traffic = (1.070415e9 + 7.950592e6) / 1.024e9 * (score_bytes / n_launch)
        dg = diags[-1]
        # bounded CPU sample of the very same batch (first rows), single thread
        cpu = None
        if args.cpu_sample_snps > 0:
            ns = min(args.cpu_sample_snps, m)
            rows_host = bed[:ns, :rb].contiguous().cpu().numpy().ravel()
            rate, dt, mm = oracle_sample_rate(ph, rows_host, n, 1)
            from oracle import oracle as O
            ref_out, *_ = O.cct_bed("binary", rows_host[: 8 * rb], n, ph["R"], ph["E"], ph["Y"], ph["Cova"])
            got = out_dev.cpu().numpy().reshape(10, m).T[:8]
            pc = [2, 3, 4, 5, 6]
            with np.errstate(all="ignore"):
                chk = float(np.nanmax(np.abs(np.log10(got[:, pc]) - np.log10(ref_out[:, pc])) / np.maximum(np.abs(np.log10(ref_out[:, pc])), 1e-3)))
            cpu = {"value": rate, "unit": "SNPs/s", "cores": 1, "kind": "port",
                   "sample": f"first {mm} SNPs of the timed batch at N={n} ({dt:.1f} s); restated reference (C oracle, -O2), not the R interpreter",
                   "max_rel_dlog10p_vs_gpu_first8": chk}
        try:
            fp64 = ctx.fp64_peak_tflops()
        except Exception:
            fp64 = None
        line = {
            "metric": "SNPs/sec at N=500k (SPAGxE_CCT Step-2 loop, device-timed)", "value": value, "unit": "SNPs/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n, m),
            "e2e": e2e, "gpu_launches": int(sum(d["kernel_launches"] for d in diags)),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_score_tc3 (2-bit unpack + score pass, tcgen05 kind::i8)",
                         "peak_source": peak_src, "launches_per_step": n_launch,
                         "algorithmic_bytes_per_launch": score_bytes / n_launch, "ms_per_launch": ms_score / n_launch,
                         "traffic_source": "ncu --set full of one 1.024 GB launch (profiles/r01_ncu_k_score_tc3_final.txt), scaled by launch size",
                         "share_of_step": ms_score / (ms_max / K)},
            "cpu_baseline": cpu, "clocks": clocks,
            "stages_ms": {"score": ms_score, "tests(finalize+SPA+branchB)": float(np.mean([d["ms_tests"] for d in diags])),
                          "spa(finalize+quadrature SPA)": float(np.mean([d["ms_spa"] for d in diags])),
                          "branch_b(solver farm)": float(np.mean([d["ms_branch_b"] for d in diags])),
                          "total_device": float(np.mean([d["ms_total"] for d in diags]))},
            "farm_work": {k: int(dg[k]) for k in ["bb_irls_sample_passes", "bb_tail_sample_passes", "bb_prep_sample_passes",
                                                  "bb_iterations", "spa_nodes"]},
            "routing": {k: int(dg[k]) for k in ["n_filtered", "n_branch_a", "n_branch_b", "n_spa", "newton_iters"]},
            "fp64_peak_tflops_measured": fp64,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
