#!/usr/bin/env python
"""Turn an .ncu-rep into the small text summary committed under profiles/ (key raw metrics, stall mix,
hottest SASS lines).  Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, dst):
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    lines = []
    for r in raw[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append(f"== kernel: {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"{k:84s} {r[i]:>16s} {units[i]}")
    src = page(rep, "source")
    # the source page holds one block per kernel: a "Kernel Name" row, a header row, then one row per SASS line
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1] if len(r) > 1 else "?", "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    for blk in blocks:
        h = blk["hdr"]
        if not h or "Source" not in h or "# Samples" not in h:
            continue
        si, ci = h.index("Source"), h.index("# Samples")
        stall = [(i, x) for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        tot, agg, rows = 0.0, {x: 0.0 for _, x in stall}, []
        for k, r in enumerate(blk["rows"]):
            if len(r) <= max(ci, si):
                continue
            try:
                v = float(r[ci])
            except ValueError:
                continue
            tot += v
            rows.append((v, k, r[si][:100]))
            for i, x in stall:
                try:
                    agg[x] += float(r[i])
                except (ValueError, IndexError):
                    pass
        lines.append(f"== warp-state samples of {blk['name'][:70]}: {tot:.0f}")
        lines.append("   " + ", ".join(f"{k[6:]}={v / max(tot, 1):.1%}" for k, v in sorted(agg.items(), key=lambda z: -z[1])[:8]))
        lines.append("== hottest SASS lines")
        for v, k, s in sorted(rows, reverse=True)[:12]:
            lines.append(f"{v / max(tot, 1):7.2%}  line {k:5d}  {s}")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:30]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
