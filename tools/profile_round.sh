#!/bin/bash
# Evidence for one round, run on the GPU box:  gpurun -- tools/profile_round.sh <tag>
# 1. bench.py (plain)  2. ncu launch list of a short run of the same program  3. ncu --set full of the fused kernels on the
# config-2 workload for the three fragment sets.  Every ncu run follows a plain run of the same command.  The reports are
# summarised ON the box (gpurun_out/ may carry 64 MiB back): text summaries, per-line shares, DRAM traffic per launch.
tag=${1:-rXX}
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-config-legs --no-renderer-legs > gpurun_out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"shade_|finalize_|seed_" -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-config-legs --no-renderer-legs > gpurun_out/ncu_launch_${tag}.log 2>&1
cp profiles/traffic.json gpurun_out/traffic_${tag}.json
for kind in rasterised realistic dense; do
  python tools/prof_driver.py $kind 3 > gpurun_out/plain_${kind}_${tag}.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:shade_ -s 5 -c 5 -f -o /tmp/prof_${tag}_${kind} \
      python tools/prof_driver.py $kind 3 > gpurun_out/ncu_${kind}_${tag}.log 2>&1
  python profiles/ncu_summary.py /tmp/prof_${tag}_${kind}.ncu-rep ${kind}:8x256x50x64 > gpurun_out/${tag}_ncu_full_${kind}.txt 2>&1
  python profiles/ncu_lines.py /tmp/prof_${tag}_${kind}.ncu-rep shade_ 45 > gpurun_out/${tag}_hot_lines_${kind}.txt 2>&1
done
cp profiles/traffic.json gpurun_out/traffic_${tag}_new.json
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -12
