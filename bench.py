#!/usr/bin/env python
"""bench.py -- SNPs/sec of the SPAGxE_CCT Step-2 loop at N = 500,000 samples (BASELINE.json configs[2]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port, all host threads)

    python bench.py --workload mix|plus|mixplus ...                  # configs 4 / 5 (SPAGxEmix_CCT, SPAGxE_Plus), same JSON contract

A "step" is one pass of the hot path (2-bit unpack + score + routing + SPA + genotype-adjusted
branch) over one batch of synthetic UK-Biobank-shaped SNPs that is already resident in HBM.  The
batch (default 65,536 SNPs x 500,000 samples = 8.19 GB packed, eight ~1 GB chunks of the Step-2 pipeline; rounds 1 and
early 2 used 16,384 SNPs, `--batch-snps 16384` reproduces those lines) is far larger than the 126 MB L2, so no flush
is needed between iterations.  One call of the north-star job holds 125,000 SNPs per GPU: the per-call costs (one
branch-B solver-farm launch, result copy) amortise as there.  One JSON line is printed by rank 0.
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
DEFAULT_BATCH_SNPS = 65536
PREVALENCE = 0.01
SEED_PHENO = 12345
SEED_GENO = 20251019
# DRAM bytes per algorithmic byte of the score kernel: mean of the two 1.024 GB launches of the committed ncu --set full
# capture (profiles/r02_ncu_k_score_unpack_mma.txt): 1.008 / 1.106 GB read, 8.0 / 8.8 MB written
SCORE_TRAFFIC_PER_ALGORITHMIC_BYTE = ((1.008232e9 + 1.105881e9) / 2 + (8.013056e6 + 8.835072e6) / 2) / 1.024e9


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


# fp64-pipe instructions per unit of work, counted in the SASS of the shipped kernels (ncu source page of
# profiles/r02_ncu_k_bb_farm.txt: DFMA + DMUL + DADD + DSETP of one loop trip / samples per trip); DESIGN.md section 4.
FP64_INSTR = {"irls_pass": 70.5, "irls_pass0": 46.5, "tail_eval": 44.0, "prep": 8.0, "quad_eval": 46.0}


def fp64_stage(instr, ms, peak_tflops):
    """fp64 'roofline' of a stage: issued fp64 instructions x 2 flop (the DFMA-equivalent the peak loop counts) over time."""
    if not ms or not peak_tflops:
        return None
    ach = instr * 2.0 / (ms * 1e-3) / 1e12
    return {"bound": "fp64", "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s (fp64 instructions x 2)", "frac": ach / peak_tflops,
            "fp64_instructions": instr, "ms": ms}


def pin_to_gpu_numa(local):
    """Bind this rank (and the pinned buffers it allocates afterwards) to the CPUs of its GPU's NUMA node."""
    info = {"numa_node": None, "cpus": None}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(base + "/numa_node").read().strip())
        cpulist = open(base + "/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        if cpus:
            os.sched_setaffinity(0, cpus)
        info = {"numa_node": node, "cpus": len(cpus), "pci": bdf}
    except Exception as ex:   # affinity is best effort
        info["error"] = str(ex)[:80]
    return info


def r_reference_rate(ph, rows_host, n, m_sample):
    """When an Rscript and the reference sources are reachable, time the UNMODIFIED R SPAGxE_CCT on a slice of the
    batch (tools/run_reference.R); None otherwise (the GPU box has neither)."""
    try:
        from tests import _rscript
        if not _rscript.available():
            return None
        from oracle import oracle as O
        rb = (n + 3) // 4
        G = O.bed_decode(rows_host[: m_sample * rb], n)
        out, secs = _rscript.run("cct", "binary", G, ph["R"], ph["E"], ph["Y"], ph["Cova"])
        return {"value": m_sample / secs, "unit": "SNPs/s", "cores": 1, "kind": "R",
                "sample": f"{m_sample} SNPs of the timed batch at N={n} through the reference's own SPAGxE_CCT ({secs:.1f} s in R)"}
    except Exception:
        return None


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


def workload_config_other(kind, n, m, extra):
    rb = (n + 3) // 4
    base = {"mix": "C4: SPAGxEmix_CCT, two-way admixed population, 10 PCs, binary trait with ancestry-specific prevalence 0.1 / 0.01",
            "plus": "C5: SPAGxE_Plus, 20,000 four-member families + 20,000 singletons, sparse GRM (200,000 entries), Cox martingale residuals",
            "mixplus": "C4 data through SPAGxEmix_CCT_Plus: two-way admixed population, 10 PCs, sparse GRM = diagonal + 50,000 pairs at 0.5 "
                       "(branch-B Wald = caller's glmer, not timed)"}[kind]
    return {"workload": f"{base}, N={n}, {m} SNPs/step/GPU ({m * rb / 1e6:.0f} MB packed .bed rows resident in HBM)", "n_samples": n,
            "snps_per_step_per_gpu": m, "l2_policy": "L2 flushed by a 256 MB write between steps" if m * rb < 3e8 else "inputs larger than L2",
            "sharding": "SNP ranges per rank, no data-path collective", **extra}


def run_other(args, kind):
    """configs 4 and 5 (SPAGxEmix_CCT / SPAGxE_Plus): same contract, device-resident rows; not the headline."""
    import torch
    from spagxecct_b200 import Context, _lib
    from tests._configs import c4_admixed, c5_families
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    W, K = max(3, args.warmup), args.steps
    m = args.batch_snps if args.batch_snps != DEFAULT_BATCH_SNPS else {"mix": 4096, "mixplus": 1024}.get(kind, 8192)
    x = c4_admixed(m=m, device=dev) if kind in ("mix", "mixplus") else c5_families(m=m, device=dev)
    n = x["n"]
    rb = (n + 3) // 4
    ctx = Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_vectors(x["R"], x["E"])
    if kind == "mix":
        ctx.set_wald(_lib.TRAIT_BINARY, x["Y"], x["Cova"]); ctx.set_pcs(x["PCs"])
        if args.mix_cluster:
            ctx.set_option(_lib.OPT_MIX_CLUSTER, args.mix_cluster)
        run, ncol, prm_kw = ctx.run_mix, 12, dict(min_maf=1e-5)
        metric = "SNPs/sec at N=200k, 10 PCs (SPAGxEmix_CCT Step-2 loop, device-timed)"
    elif kind == "mixplus":
        npair = 50_000
        x["grm_i"] = np.r_[np.arange(x["n"]), np.arange(1, 2 * npair, 2)].astype(np.int32)
        x["grm_j"] = np.r_[np.arange(x["n"]), np.arange(0, 2 * npair, 2)].astype(np.int32)
        x["grm_v"] = np.r_[np.ones(x["n"]), np.full(npair, 0.5)]
        ctx.set_pcs(x["PCs"]); ctx.set_sparse_grm(x["grm_i"], x["grm_j"], x["grm_v"])
        if args.mix_cluster:
            ctx.set_option(_lib.OPT_MIX_CLUSTER, args.mix_cluster)
        run, ncol, prm_kw = ctx.run_mix_plus, 12, dict(min_maf=1e-5)
        metric = "SNPs/sec at N=200k, 10 PCs, sparse GRM (SPAGxEmix_CCT_Plus Step-2 loop, device-timed)"
    else:
        a, b, Rnew, c = ctx.plus_nullmodel(x["grm_i"], x["grm_j"], x["grm_v"], x["R"], x["E"])
        ctx.set_grm(x["grm_i"], x["grm_j"], x["grm_v"], a, b, c, Rnew)
        run, ncol, prm_kw = ctx.run_plus, 8, dict()
        metric = "SNPs/sec at N=100k with a sparse GRM (SPAGxE_Plus Step-2 loop, device-timed)"
    pitch = (rb + 127) // 128 * 128
    bed = torch.full((m, pitch), 0xFF, dtype=torch.uint8, device=dev)
    bed[:, :rb] = torch.from_numpy(x["rows"].reshape(m, rb)).to(dev)
    out_dev = torch.empty((m, ncol), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    geno = ctx.geno_desc(_lib.GENO_BED, bed.data_ptr(), n, m, pitch, _lib.LOC_DEVICE)
    prm = _lib.make_params(out_location=_lib.LOC_DEVICE, **prm_kw)
    sampler = ClockSampler(local); sampler.start()
    for _ in range(W):
        run(geno, prm, out_ptr=out_dev.data_ptr())
    torch.cuda.synchronize()
    ms_sum, diags = 0.0, []
    for _ in range(K):
        flush.fill_(1)                                   # 256 MB > 126 MB L2
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _, dg = run(geno, prm, out_ptr=out_dev.data_ptr())
        e1.record(); torch.cuda.synchronize()
        ms_sum += e0.elapsed_time(e1); diags.append(dg)
    clocks = sampler.stop()
    # end to end: pinned host rows in, host matrix out
    host_rows = torch.from_numpy(x["rows"].reshape(m, rb)).pin_memory()
    host_out = torch.empty((ncol, m), dtype=torch.float64).pin_memory()
    geno_h = ctx.geno_desc(_lib.GENO_BED, host_rows.data_ptr(), n, m, rb, _lib.LOC_HOST)
    prm_h = _lib.make_params(**prm_kw)
    run(geno_h, prm_h, out_ptr=host_out.data_ptr())
    ke = max(3, min(K, 10))
    e2 = torch.cuda.Event(enable_timing=True); e3 = torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(ke):
        _, dge = run(geno_h, prm_h, out_ptr=host_out.data_ptr())
    e3.record(); torch.cuda.synchronize()
    ms_e = e2.elapsed_time(e3)
    peak, peak_src = hbm_peak()
    try:
        fp64 = ctx.fp64_peak_tflops()
    except Exception:
        fp64 = None
    ms_score = float(np.mean([d["ms_score"] for d in diags]))
    ms_tests = float(np.mean([d["ms_tests"] for d in diags]))
    dg = diags[-1]
    k1 = x["PCs"].shape[1] + 1 if kind in ("mix", "mixplus") else 0
    line = {"metric": metric, "value": m * K / (ms_sum * 1e-3), "unit": "SNPs/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_sum / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config_other(kind, n, m, dict({"n_pcs": k1 - 1} if k1 else {}, **({"grm_nnz": int(x["grm_i"].size)} if kind != "mix" else {}))),
            "e2e": {"value": m * ke / (ms_e * 1e-3), "unit": "SNPs/s", "h2d_bytes_per_step": int(dge["h2d_bytes"]),
                    "d2h_bytes_per_step": int(dge["d2h_bytes"]), "steps": ke},
            "gpu_launches": int(sum(d["kernel_launches"] for d in diags)),
            "roofline": ({"bound": "tensor", "achieved": 2.0 * n * k1 * 2 * m / (ms_tests * 1e-3) / 1e12, "peak": fp64, "unit": "TFLOP/s",
                          "frac": (2.0 * n * k1 * 2 * m / (ms_tests * 1e-3) / 1e12 / fp64) if fp64 else None, "traffic": None,
                          "kernel": ("k_mix_tile + k_mix_finish (allele-frequency projection, fp64 CUDA cores; 'tensor' = compute bound)" if kind == "mix"
                                     else "k_mixplus_snp (per-SNP cluster kernel: allele-frequency model, GRM forms, saddlepoint; 'tensor' = compute bound)"),
                          "flop_model": "2 passes x 2 N (k+1) flop per SNP (X'g and x_i'beta)", "peak_source": "DFMA loop measured in this run"}
                         if k1 else
                         {"bound": "hbm", "achieved": float(dg["score_bytes"]) / (ms_score * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": float(dg["score_bytes"]) / (ms_score * 1e-3) / 1e9 / peak, "traffic": None,
                          "kernel": "k_score_unpack_mma", "peak_source": peak_src}),
            "cpu_baseline": None, "clocks": clocks,
            "stages_ms": {"score": ms_score, "tests": ms_tests, "total_device": float(np.mean([d["ms_total"] for d in diags]))},
            "routing": {k: int(dg[k]) for k in ["n_filtered", "n_branch_a", "n_branch_b", "n_spa", "n_mix_logistic", "newton_iters"]},
            "fp64_peak_tflops_measured": fp64}
    if args.cpu_sample_snps > 0:
        from oracle import oracle as O
        O.build()
        ns = min(64 if k1 else 256, m)
        t0 = time.perf_counter()
        if kind == "mix":
            ref, *_ = O.mix_bed("binary", x["rows"][: ns * rb], n, x["R"], x["E"], x["Y"], x["Cova"], x["PCs"], min_maf=1e-5)
        elif kind == "mixplus":
            ref, *_ = O.mixplus_bed(x["rows"][: ns * rb], n, x["R"], x["E"], x["PCs"], x["grm_i"], x["grm_j"], x["grm_v"], min_maf=1e-5)
        else:
            ref, *_ = O.plus_bed(x["rows"][: ns * rb], n, x["R"], x["E"], Rnew, a, b, c, x["grm_i"], x["grm_j"], x["grm_v"])
        dt = time.perf_counter() - t0
        got = out_dev.cpu().numpy().reshape(ncol, m).T[:ns]
        pc = [2, 3, 4, 5, 6] if kind == "mix" else ([2, 5, 6] if kind == "mixplus" else [2, 3, 4])
        with np.errstate(all="ignore"):
            chk = float(np.nanmax(np.abs(np.log10(got[:, pc]) - np.log10(ref[:, pc])) / np.maximum(np.abs(np.log10(ref[:, pc])), 1e-3)))
        line["cpu_baseline"] = {"value": ns / dt, "unit": "SNPs/s", "cores": 1, "kind": "port",
                                "sample": f"first {ns} SNPs of the timed batch ({dt:.1f} s); restated reference (C oracle, -O2)",
                                "max_rel_dlog10p_vs_gpu": chk}
    if rank == 0:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cct", choices=["cct", "mix", "plus", "mixplus"])
    ap.add_argument("--n-samples", type=int, default=N_SAMPLES)
    ap.add_argument("--batch-snps", type=int, default=DEFAULT_BATCH_SNPS)
    ap.add_argument("--ref-snps", type=int, default=0)
    ap.add_argument("--cpu-sample-snps", type=int, default=1536)   # ~15 s of one host core at N = 500k
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-affinity", action="store_true")
    ap.add_argument("--mix-cluster", type=int, default=0)           # tuning probe for --workload mix
    ap.add_argument("--trait", default="binary", choices=["binary", "survival"])   # survival: a variant run (Cox Wald test in branch B), single GPU
    ap.add_argument("--probe-cfg", type=int, default=0)             # tools only: score-kernel variant of libspagxe_cuda_probes.so
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload != "cct":
        return run_other(args, args.workload)

    import torch
    import torch.distributed as dist
    from spagxecct_b200 import Context, _lib
    if args.probe_cfg:      # tuning runs against the probe build (make -C spagxecct_b200/csrc probes); never a reported number
        _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libspagxe_cuda_probes.so")

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (spagxecct_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = {"numa_node": None} if args.no_affinity else pin_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, m = args.n_samples, args.batch_snps
    W, K = max(3, args.warmup), args.steps

    # ---- inputs: rank 0 owns Step-1 output; ONE NCCL broadcast distributes R, E and the Wald data (north_star) ----
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
    ctx = Context(local)
    if args.probe_cfg:
        ctx.set_option(100, args.probe_cfg)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_vectors_device(n, vec[0].data_ptr(), vec[1].data_ptr())
    ctx.set_wald_device(_lib.TRAIT_BINARY, ycov[0].data_ptr(), ycov[1].data_ptr(), 2)
    if args.trait == "survival":   # same cohort read as time-to-event: the cases are the events (1 %), exponential times
        assert world == 1, "--trait survival is a single-GPU variant run"
        rs = np.random.default_rng(777)
        ctx.set_wald_survival(rs.exponential(1.0, n) * np.where(ph["Y"] > 0, 0.5, 1.0), ph["Y"], ph["Cova"])

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

    def gather_floats(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world == 1:
            return [float(v)]
        outl = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outl, t)
        return [float(o.item()) for o in outl]

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
    ms_all = gather_floats(e0.elapsed_time(e1))
    ms_max = max(ms_all)
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
        ke = max(3, min(K, args.e2e_steps))
        for _ in range(ke):
            _, dge = ctx.run_cct(geno_host, prm_host, out_ptr=host_out.data_ptr())
        e3.record()
        barrier()
        ms_e_all = gather_floats(e2.elapsed_time(e3))
        h2d = int(dge["h2d_bytes"])
        e2e = {"value": world * m * ke / (max(ms_e_all) * 1e-3), "unit": "SNPs/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(dge["d2h_bytes"]), "steps": ke,
               "per_rank_h2d_gbs": [h2d * ke / (t * 1e-3) / 1e9 for t in ms_e_all],
               "aggregate_h2d_gbs": world * h2d * ke / (max(ms_e_all) * 1e-3) / 1e9,
               "device_ms_per_step_inside_e2e": float(dge["ms_total"]),
               "copy_compute_overlap": "chunks of 1 GiB double-buffered: H2D of chunk k+1 on the copy stream while chunk k is tested",
               "numa": numa}
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
        # DRAM traffic of the same kernel from the committed `ncu --set full` capture (profiles/r02_ncu_k_score_unpack_mma.txt),
        # scaled to this launch size
        traffic = SCORE_TRAFFIC_PER_ALGORITHMIC_BYTE * (score_bytes / n_launch)
        dg = diags[-1]
        # bounded CPU sample of the very same batch (first rows), single thread
        cpu = None
        if args.cpu_sample_snps > 0 and world == 1:     # the CPU baseline is timed on rank 0 at N = 1 only
            ns = min(args.cpu_sample_snps, m)
            rows_host = bed[:ns, :rb].contiguous().cpu().numpy().ravel()
            rate, dt, mm = oracle_sample_rate(ph, rows_host, n, 1)
            from oracle import oracle as O
            nchk = min(ns, 256)
            ref_out, route, *_ = O.cct_bed("binary", rows_host[: nchk * rb], n, ph["R"], ph["E"], ph["Y"], ph["Cova"], nthreads=os.cpu_count())
            got = out_dev.cpu().numpy().reshape(10, m).T[:nchk]
            pc = [2, 3, 4, 5, 6]
            with np.errstate(all="ignore"):
                chk = float(np.nanmax(np.abs(np.log10(got[:, pc]) - np.log10(ref_out[:, pc])) / np.maximum(np.abs(np.log10(ref_out[:, pc])), 1e-3)))
            cpu = {"value": rate, "unit": "SNPs/s", "cores": 1, "kind": "port",
                   "sample": f"first {mm} SNPs of the timed batch at N={n} ({dt:.1f} s); restated reference (C oracle, -O2), not the R interpreter",
                   "parity_first_snps": {"snps": nchk, "max_rel_dlog10p_vs_gpu": chk, "n_spa": int(((route & 4) > 0).sum()),
                                         "n_branch_b": int(((route & 2) > 0).sum())}}
            # ... and of the batch's genotype-adjusted SNPs (branch B: adjusted residual SPA + glm.fit Wald + CCT), up to 16
            full = out_dev.cpu().numpy().reshape(10, m).T
            bsel = np.flatnonzero(full[:, 6] <= 0.001)[:16]
            if bsel.size:
                rows_b = np.ascontiguousarray(bed[torch.as_tensor(bsel, device=bed.device), :rb].cpu().numpy()).ravel()
                ref_b, route_b, *_ = O.cct_bed("binary", rows_b, n, ph["R"], ph["E"], ph["Y"], ph["Cova"], nthreads=os.cpu_count())
                with np.errstate(all="ignore"):
                    chk_b = float(np.nanmax(np.abs(np.log10(full[bsel][:, pc]) - np.log10(ref_b[:, pc])) / np.maximum(np.abs(np.log10(ref_b[:, pc])), 1e-3)))
                cpu["parity_branch_b_snps"] = {"snps": int(bsel.size), "max_rel_dlog10p_vs_gpu": chk_b,
                                               "n_branch_b": int(((route_b & 2) > 0).sum()), "n_spa": int(((route_b & 4) > 0).sum())}
            rr = r_reference_rate(ph, rows_host, n, min(32, ns))
            if rr:
                cpu["r_interpreter"] = rr
        try:
            fp64 = ctx.fp64_peak_tflops()
        except Exception:
            fp64 = None
        ms_spa = float(np.mean([d["ms_spa"] for d in diags])); ms_bb = float(np.mean([d["ms_branch_b"] for d in diags]))
        n_items = max(1, int(dg["n_branch_b"]))
        pass0 = float(n_items) * n                                   # one mustart pass per work item
        irls = max(0.0, float(dg["bb_irls_sample_passes"]) - pass0)
        bb_instr = (irls * FP64_INSTR["irls_pass"] + pass0 * FP64_INSTR["irls_pass0"] +
                    float(dg["bb_tail_sample_passes"]) * FP64_INSTR["tail_eval"] + float(dg["bb_prep_sample_passes"]) * FP64_INSTR["prep"])
        farm_tail_iters = float(dg["bb_tail_sample_passes"]) / n
        quad_evals = max(0.0, float(dg["newton_iters"]) + 2.0 * float(dg["n_spa"]) - farm_tail_iters) * float(dg["spa_nodes"])
        line = {
            "metric": "SNPs/sec at N=500k (SPAGxE_CCT Step-2 loop, device-timed)", "value": value, "unit": "SNPs/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(n, m), **({"variant": "traits=survival: branch-B Wald test = Cox model (k_cox_wald), cpu_baseline skipped"} if args.trait == "survival" else {})),
            "e2e": e2e, "gpu_launches": int(sum(d["kernel_launches"] for d in diags)),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_score_unpack_mma (2-bit unpack + score pass, tcgen05 kind::i8)",
                         "peak_source": peak_src, "launches_per_step": n_launch,
                         "algorithmic_bytes_per_launch": score_bytes / n_launch, "ms_per_launch": ms_score / n_launch,
                         "traffic_source": "ncu --set full of one 1.024 GB launch (profiles/r02_ncu_k_score_unpack_mma.txt), scaled by launch size",
                         "share_of_step": ms_score / (ms_max / K)},
            "roofline_stages": {
                "spa_quadrature(k_spa_quad)": fp64_stage(quad_evals * FP64_INSTR["quad_eval"], ms_spa, fp64),
                "branch_b(k_bb_farm)": fp64_stage(bb_instr, ms_bb, fp64),
                "model": "fp64 instructions per unit counted in the SASS loops (bench.py FP64_INSTR) x units counted on the device",
            },
            "cpu_baseline": cpu, "clocks": clocks,
            "stages_ms": {"score": ms_score, "tests(finalize+SPA+branchB)": float(np.mean([d["ms_tests"] for d in diags])),
                          "spa(finalize+quadrature SPA)": ms_spa, "branch_b(solver farm)": ms_bb,
                          "total_device": float(np.mean([d["ms_total"] for d in diags]))},
            "farm_work": {k: int(dg[k]) for k in ["cox_evals", "bb_irls_sample_passes", "bb_tail_sample_passes", "bb_prep_sample_passes",
                                                  "bb_iterations", "spa_nodes"]},
            "routing": {k: int(dg[k]) for k in ["n_filtered", "n_branch_a", "n_branch_b", "n_spa", "newton_iters"]},
            "per_rank_ms": ms_all,
            "fp64_peak_tflops_measured": fp64,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
