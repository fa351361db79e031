#!/bin/bash
# Round evidence on one B200: GPU tests, the three bench workloads, the reference arm, the ncu launch list of the
# default bench command and one `ncu --set full` capture per hot kernel.  Everything lands in gpurun_out/; the
# summaries that are judged are copied into profiles/ by hand (tools/ncu_summary.py).
#   usage: gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh r02'
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "ref rc=$?"
python bench.py --workload mix --steps 5 --warmup 3 > $out/${tag}_bench_mix.json 2> $out/${tag}_bench_mix.err; echo "mix rc=$?"
python bench.py --workload plus --steps 10 --warmup 3 > $out/${tag}_bench_plus.json 2> $out/${tag}_bench_plus.err; echo "plus rc=$?"
python bench.py --workload mixplus --steps 3 --warmup 3 > $out/${tag}_bench_mixplus.json 2> $out/${tag}_bench_mixplus.err; echo "mixplus rc=$?"
short="python bench.py --steps 2 --warmup 3 --no-e2e --cpu-sample-snps 0"
$short > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|spagxe' -c 400 --csv \
    --log-file $out/${tag}_launches_bench_cct.csv $short > $out/${tag}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$short > $out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_bb_farm|k_spa_quad|k_finalize_cct|k_score_unpack_mma' -s 8 -c 8 \
    -o /tmp/${tag}_prof_cct $short > $out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"
# the .ncu-rep stays on the box (gpurun copies back at most 64 MiB): only its text summary travels
python tools/ncu_summary.py /tmp/${tag}_prof_cct.ncu-rep $out/${tag}_ncu_cct_kernels.txt > /dev/null 2>&1; echo "ncu summary rc=$?"
shortmix="python bench.py --workload mix --steps 1 --warmup 2 --no-e2e --cpu-sample-snps 0"
$shortmix > $out/${tag}_plain_mix.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|spagxe' -c 400 --csv \
    --log-file $out/${tag}_launches_bench_mix.csv $shortmix > $out/${tag}_ncu_launches_mix.log 2>&1
echo "mix launch list rc=$?"
for f in default reference mix plus; do tail -c 400 $out/${tag}_bench_$f.json | head -c 400; echo; done
