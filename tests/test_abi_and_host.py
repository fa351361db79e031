"""CPU tests: the C-ABI library loads and exports every symbol include/spagxe.h declares, fails
loudly without a GPU, and the host-side mirror of the reference interface behaves like the R code."""
import ctypes
import os
import re

import numpy as np
import pytest

from spagxecct_b200 import _lib, api, bedio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "spagxe.h")).read()
    return sorted(set(re.findall(r"\b(spagxe_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/spagxe.h but not exported"
    assert sorted(_lib.EXPORTS) == syms
    assert b"sm_100a" in _lib.load().spagxe_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.SpagxeError) as ei:
        _lib.Context(0)
    assert ei.value.code == _lib.ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "spagxecct_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".R")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("CPU oracle", "").replace("the oracle", "") or f == "harness.py", (dirpath, f)


def test_check_input_resid_messages():
    ids = [f"IID-{i}" for i in range(5)]
    assert api.check_input_Resid(ids, ids, np.zeros(5)) is None
    assert list(api.check_input_Resid(ids, ids[::-1], np.zeros(5))) == [4, 3, 2, 1, 0]
    with pytest.raises(api.RStop, match="unique\\(pIDs\\) should be the same"):
        api.check_input_Resid(ids, ids[:4] + ["x"], np.zeros(5))
    with pytest.raises(api.RStop, match="should not have a duplicated element"):
        api.check_input_Resid(ids, ids + ids[:1], np.zeros(5))
    with pytest.raises(api.RStop, match="same length as input data"):
        api.check_input_Resid(ids, ids, np.zeros(4))
    # the bundled Pheno.mtx has IID, not ID: Phen.mtx$ID is NULL.  `sort(unique(NULL)) != ...` is logical(0), so
    # any() is FALSE and the stop() the reference hits is the length check (SURVEY.md section 4, R/SPAGxECCT.R:1906)
    with pytest.raises(api.RStop, match="same length as input data"):
        api.check_input_Resid(None, ids, np.zeros(5))
    # rownames(Geno.mtx) NULL: every comparison is against logical(0) and R passes silently
    assert api.check_input_Resid(ids, None, np.zeros(5)) is None
    with pytest.raises(api.RStop, match="are required"):
        api.check_input_Resid(None, None, np.zeros(5))


def test_pack_rows_roundtrip_and_reorder(oracle):
    rng = np.random.default_rng(0)
    n, m = 1003, 7
    G = rng.integers(0, 3, (n, m)).astype(float)
    G[rng.random((n, m)) < 0.05] = np.nan
    rows = bedio.pack_rows(np.where(np.isnan(G), -1, G))
    assert rows.size == m * ((n + 3) // 4)
    back = oracle.bed_decode(rows, n)
    assert np.array_equal(np.nan_to_num(back, nan=-1), np.nan_to_num(G, nan=-1))
    perm = rng.permutation(n)[:600]
    rows2 = bedio.reorder_samples(rows, n, m, perm)
    back2 = oracle.bed_decode(rows2, 600)
    assert np.array_equal(np.nan_to_num(back2, nan=-1), np.nan_to_num(G[perm], nan=-1))


def test_bedfile_reader(tmp_path, golden_dir, oracle):
    d = np.load(f"{golden_dir}/bed_SPAGxE.npz")
    n = int(d["n"])
    stem = tmp_path / "toy"
    d["bed"].tofile(str(stem) + ".bed")
    with open(str(stem) + ".fam", "w") as f:
        for i in range(n):
            f.write(f"IID-{i + 1} IID-{i + 1} 0 0 1 -9\n")
    with open(str(stem) + ".bim", "w") as f:
        for j, s in enumerate(d["snp_ids"]):
            f.write(f"1\t{s}\t0\t{j + 1}\tA\tG\n")
    bf = bedio.BedFile(str(stem) + ".bed")
    assert (bf.n, bf.m) == (n, 100) and bf.snp_ids[0] == "SNP-1" and bf.sample_ids[1] == "IID-2"
    assert np.array_equal(bf.rows_array(), d["bed"][3:])
    with open(str(stem) + ".bed", "r+b") as f:
        f.write(b"\x6c\x1b\x00")
    with pytest.raises(ValueError, match="variant-major"):
        bedio.BedFile(str(stem) + ".bed")


def test_output_column_names_match_reference():
    assert api.COLS_CCT == ["MAF", "missing.rate", "p.value.spaGxE", "p.value.spaGxE.Wald", "p.value.spaGxE.CCT.Wald",
                            "p.value.normGxE", "p.value.betaG", "Stat.betaG", "Var.betaG", "z.betaG"]  # R/SPAGxECCT.R:499-501
    assert api.COLS_MIX[10:] == ["MAF.est.negative.num", "MAC"]                                      # :963-966
    assert api.COLS_PLUS[2:4] == ["p.value.spaGxE.plus", "p.value.normGxE.plus"]                    # Plus :203-205


def test_bench_reference_arm_contract_on_cpu():
    """`bench.py --impl reference` (the reference's CPU path: the oracle port on the host cores) prints the driver's JSON
    line with impl/cpu_baseline/e2e and never touches a GPU; run here at a reduced N so that the contract is checked
    without one.  Non-zero ranks print nothing."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--n-samples", "3000", "--ref-snps", "8",
           "--steps", "2", "--warmup", "1"]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "SNPs/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 2 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "SNPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
    out1 = subprocess.run(cmd, capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=300)
    assert out1.returncode == 0 and out1.stdout.strip() == ""


def test_r_shim_compiles_against_the_c_abi():
    """R is not installed here, so the .Call shim cannot be built with R CMD SHLIB; at least every call it makes into
    include/spagxe.h and every R API use must type-check (gcc -fsyntax-only against minimal R declarations), and the
    .Call names used by the R wrappers must all be registered by the shim with the right arity."""
    import subprocess
    shim = os.path.join(ROOT, "spagxecct_b200", "R", "spagxe_rshim.c")
    cp = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "r_stub"), shim],
                        capture_output=True, text=True)
    assert cp.returncode == 0, cp.stderr
    src = open(shim).read()
    registered = dict(re.findall(r'\{"(spagxe_r_\w+)", \(DL_FUNC\)&\w+, (\d+)\}', src))
    rsrc = open(os.path.join(ROOT, "spagxecct_b200", "R", "SPAGxECCT_cuda.R")).read()
    calls = re.findall(r'\.Call\("(spagxe_r_\w+)"((?:[^()]|\([^()]*(?:\([^()]*\)[^()]*)*\))*)\)', rsrc)
    assert len(calls) >= 20
    for name, args in calls:
        assert name in registered, name
        depth, nargs = 0, 0
        for ch in args:
            depth += ch in "(["; depth -= ch in ")]"
            nargs += (ch == "," and depth == 0)
        assert nargs == int(registered[name]), (name, nargs, registered[name])
    # every exported Step-2 function of the reference's NAMESPACE has a wrapper
    for fn in ["SPAGxE_CCT", "SPAGxE_CCT_one_SNP", "SPAGxEmix_CCT", "SPAGxEmix_CCT_one_SNP", "SPAGxE_Plus", "SPAGxE_Plus_one_SNP"]:
        assert re.search(rf"^{fn} <- function\(", rsrc, re.M), fn
