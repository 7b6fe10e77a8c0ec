#!/bin/bash
# Evidence for one round, run on the GPU box:  gpurun -- tools/profile_round.sh <tag>
# 1. bench.py (plain)  2. ncu launch list of the same command  3. ncu --set full of the fused kernels on the
# config-2 workload (realistic and dense fragments).  Every ncu run follows a plain run of the same command.
tag=${1:-rXX}
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_${tag}.log 2>&1
for kind in realistic dense; do
  python tools/prof_driver.py $kind 3 > gpurun_out/plain_${kind}_${tag}.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:shade_ -s 4 -c 4 -f -o gpurun_out/prof_${tag}_${kind} \
      python tools/prof_driver.py $kind 3 > gpurun_out/ncu_${kind}_${tag}.log 2>&1
done
ls -la gpurun_out | tail -20
