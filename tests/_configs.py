"""Full-size synthetic inputs of BASELINE.json's configs 3, 4 and 5, shared by the GPU parity tests
and by bench.py's --workload arms.  Everything is seeded; genotype blocks are drawn on the GPU with torch (the tests
that use them are -m gpu) and handed to the CPU oracle as packed PLINK rows, so both sides see identical bytes.

    C3  bench.make_pheno / bench.make_batch_device          (N = 500,000, binary 1 % prevalence)
    C4  c4_admixed(n=200_000, k=10)                          two-way admixture, 10 PCs, ancestry-specific prevalence
    C5  c5_families(n_fam=20_000, n_single=20_000)           4-member families + singletons, sparse GRM, Cox residuals
"""
import os

import numpy as np

from spagxecct_b200.harness import cox_martingale_resid, get_resid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pack_codes_torch(code):
    """(m, n) uint8 PLINK codes (torch, any device) -> (m, ceil(n/4)) packed bytes."""
    import torch
    m, n = code.shape
    n4 = (n + 3) // 4
    if n4 * 4 > n:
        pad = torch.zeros((m, n4 * 4 - n), dtype=torch.uint8, device=code.device)   # PLINK pads with 00
        code = torch.cat([code, pad], dim=1)
    c = code.view(m, n4, 4)
    return (c[:, :, 0] | (c[:, :, 1] << 2) | (c[:, :, 2] << 4) | (c[:, :, 3] << 6)).contiguous()


def dosage_to_code_torch(dos, missing=None):
    """A1 dosage 0/1/2 (uint8) -> PLINK code (2 -> 00, 1 -> 10, 0 -> 11, missing -> 01)."""
    import torch
    code = torch.where(dos == 2, 0, torch.where(dos == 1, 2, 3)).to(torch.uint8)
    if missing is not None:
        code = torch.where(missing, torch.tensor(1, dtype=torch.uint8, device=dos.device), code)
    return code


def c4_admixed(n=200_000, m=512, k=10, seed=20251104, device="cuda", golden_dir=None, n_pc_snps=2000):
    """config 4: ancestry proportions = the reference's a6 vector tiled to n with jitter; per-SNP ancestry-specific allele
    frequencies from the simulation script's grid plus rare / monomorphic-in-one-ancestry cases; PCs = top-k left singular
    vectors of a standardised n x n_pc_snps genotype block drawn the same way
    (simulation_studies/SPAGxEmixCCT_local/code/generate_PCs_forLocalanceAnalysis.R:53-64); binary trait with
    ancestry-specific prevalence 0.1 / 0.01 (SPAGxEmixCCT_AdmixedPopulationAnalysis_binary...R:32-33)."""
    import torch
    golden_dir = golden_dir or os.path.join(ROOT, "tests", "golden")
    a6 = np.load(os.path.join(golden_dir, "mix.npz"))["a6"][:, 0]
    rng = np.random.default_rng(seed)
    reps = -(-n // a6.size)
    alpha = np.clip(np.tile(a6, reps)[:n] + rng.normal(0, 0.01, n), 0.0, 1.0)
    g = torch.Generator(device=device); g.manual_seed(seed)
    al = torch.from_numpy(alpha).to(device)

    def draw(p1, p2, mm):   # (mm, n) uint8 dosages, per-sample frequency alpha p1 + (1 - alpha) p2
        p = al[None, :] * p1[:, None] + (1 - al[None, :]) * p2[:, None]
        d = (torch.rand((mm, n), generator=g, device=device, dtype=torch.float64) < p).to(torch.uint8)
        d += (torch.rand((mm, n), generator=g, device=device, dtype=torch.float64) < p).to(torch.uint8)
        return d

    # ---- PCs from an independent genotype block ----
    pe = torch.from_numpy(rng.uniform(0.05, 0.5, n_pc_snps)).to(device)
    pa = torch.clamp(pe + torch.from_numpy(rng.normal(0, 0.15, n_pc_snps)).to(device), 0.02, 0.6)
    gram = torch.zeros((n_pc_snps, n_pc_snps), dtype=torch.float64, device=device)
    Z = torch.empty((n, n_pc_snps), dtype=torch.float64, device=device)
    for j0 in range(0, n_pc_snps, 250):
        d = draw(pe[j0:j0 + 250], pa[j0:j0 + 250], min(250, n_pc_snps - j0)).to(torch.float64)
        ph = d.mean(dim=1, keepdim=True) / 2
        Z[:, j0:j0 + d.shape[0]] = ((d - 2 * ph) / torch.sqrt(2 * ph * (1 - ph))).T
    gram = Z.T @ Z
    evals, evecs = torch.linalg.eigh(gram)
    V = evecs[:, -k:].flip(1)
    U = Z @ V
    U = U / torch.linalg.norm(U, dim=0, keepdim=True)          # unit-norm loadings like princomp(covmat = GRM)
    PCs = U.cpu().numpy()
    del Z, gram, U

    # ---- test SNPs ----
    pairs = [(0.3, 0.01), (0.1, 0.01), (0.05, 0.4), (0.01, 0.001), (0.2, 0.0), (0.0, 0.15), (0.0012, 0.0003),
             (0.25, 0.25), (0.00004, 0.00002), (0.002, 0.0), (0.00008, 0.00008)]
    p1 = torch.tensor([pairs[j % len(pairs)][0] for j in range(m)], dtype=torch.float64, device=device)
    p2 = torch.tensor([pairs[j % len(pairs)][1] for j in range(m)], dtype=torch.float64, device=device)
    dos = draw(p1, p2, m)
    jj = torch.arange(m, device=device)[:, None]
    dos = torch.where((jj % 5) == 0, 2 - dos, dos)              # every fifth SNP counts the other allele
    miss = (torch.rand((m, n), generator=g, device=device) < 0.02) & ((jj % 3) == 0)
    rows = pack_codes_torch(dosage_to_code_torch(dos, miss)).cpu().numpy()
    del dos, miss

    X1 = rng.standard_normal(n); X2 = rng.binomial(1, 0.5, n).astype(np.float64); E = rng.standard_normal(n)
    b0 = alpha * np.log(0.1 / 0.9) + (1 - alpha) * np.log(0.01 / 0.99)
    eta = b0 + 0.1 * X1 + 0.1 * X2 + 0.1 * E
    Y = rng.binomial(1, 1 / (1 + np.exp(-eta))).astype(np.float64)
    cova = np.column_stack([X1, X2, PCs[:, 0], PCs[:, 1], PCs[:, 2], PCs[:, 3]])
    R = get_resid("binary", Y, [X1, X2, PCs[:, 0], PCs[:, 1], PCs[:, 2], PCs[:, 3], E])
    return dict(n=n, m=m, rows=rows.ravel(), R=R, E=E, Y=Y, Cova=cova, PCs=PCs, alpha=alpha)


def c5_families(n_fam=20_000, n_single=20_000, m=512, seed=20251105, device="cuda", event_rate=0.01):
    """config 5: n_fam four-member families (2 founders + 2 offspring, gene dropping) + singletons, mirroring
    data/sparseGRM.SPAGxEPlus.rda (lower triangle incl. diagonal; diag 1, parent-offspring / sibs 0.5);
    Weibull time-to-event per simulation_studies/SPAGxECCT/code/SPAGxECCT_survival_typeIerror...R:33-47 with the
    scale tuned to the event rate; Step-1 residuals = martingale residuals of a harness-side Cox fit."""
    import torch
    n = 4 * n_fam + n_single
    rng = np.random.default_rng(seed)
    g = torch.Generator(device=device); g.manual_seed(seed)
    maf = torch.from_numpy(rng.uniform(0.005, 0.5, m)).to(device)

    def hap(cnt):   # (m, cnt) alleles 0/1
        return (torch.rand((m, cnt), generator=g, device=device, dtype=torch.float64) < maf[:, None]).to(torch.uint8)

    def transmit(h1, h2):
        pick = torch.rand(h1.shape, generator=g, device=device) < 0.5
        return torch.where(pick, h1, h2)

    fa1, fa2, mo1, mo2 = hap(n_fam), hap(n_fam), hap(n_fam), hap(n_fam)
    kids = [transmit(fa1, fa2) + transmit(mo1, mo2) for _ in range(2)]
    dos = torch.empty((m, n), dtype=torch.uint8, device=device)
    # family f occupies samples 4f .. 4f+3: father, mother, child 1, child 2
    dos[:, 0:4 * n_fam:4] = fa1 + fa2
    dos[:, 1:4 * n_fam:4] = mo1 + mo2
    dos[:, 2:4 * n_fam:4] = kids[0]
    dos[:, 3:4 * n_fam:4] = kids[1]
    dos[:, 4 * n_fam:] = hap(n_single) + hap(n_single)
    jj = torch.arange(m, device=device)[:, None]
    dos = torch.where((jj % 4) == 1, 2 - dos, dos)
    miss = (torch.rand((m, n), generator=g, device=device) < 0.01) & ((jj % 2) == 0)
    rows = pack_codes_torch(dosage_to_code_torch(dos, miss)).cpu().numpy()
    del dos, miss
    f = np.arange(n_fam) * 4
    gi = np.r_[np.arange(n), f + 2, f + 2, f + 3, f + 3, f + 3].astype(np.int32)   # lower triangle: ID1 > ID2
    gj = np.r_[np.arange(n), f + 0, f + 1, f + 0, f + 1, f + 2].astype(np.int32)
    gv = np.r_[np.ones(n), np.full(5 * n_fam, 0.5)]

    X1 = rng.standard_normal(n); X2 = rng.binomial(1, 0.5, n).astype(np.float64); E = rng.standard_normal(n)
    fam_eff = np.r_[np.repeat(rng.normal(0, 0.3, n_fam), 4), np.zeros(n_single)]   # shared familial component
    cens = rng.weibull(1.0, n) * 0.15
    eps = rng.random(n)
    base = (-np.log(eps) * np.exp(-0.1 * X1 - 0.1 * X2 - 0.1 * E - fam_eff)) ** 0.5
    lo, hi = 1e-3, 1e3
    for _ in range(60):                                    # scale of the event time giving the wanted event rate
        mid = np.sqrt(lo * hi)
        if np.mean(base * mid < cens) > event_rate:
            lo = mid
        else:
            hi = mid
    t_event = base * np.sqrt(lo * hi)
    time = np.minimum(t_event, cens); event = (t_event < cens).astype(np.float64)
    R = cox_martingale_resid(time, event, np.column_stack([X1, X2, E]))
    return dict(n=n, m=m, rows=rows.ravel(), R=R, E=E, Y=event, time=time, Cova=np.column_stack([X1, X2]),
                grm_i=gi, grm_j=gj, grm_v=gv)
