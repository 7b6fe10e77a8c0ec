#!/bin/bash
# Evidence for one round, run on the GPU box:  gpurun -- tools/profile_round.sh <tag>
# 1. bench.py (plain)  2. ncu launch list of the same command  3. ncu --set full of the fused kernels on the
# config-2 workload (realistic and dense fragments), of the Phong kernels and of the rasteriser.  Every ncu run follows
# a plain run of the same command.
tag=${1:-rXX}
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"shade_|phong_|rasterize_|finalize_" -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_${tag}.log 2>&1
for kind in realistic dense; do
  python tools/prof_driver.py $kind 3 > gpurun_out/plain_${kind}_${tag}.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:shade_ -s 4 -c 4 -f -o gpurun_out/prof_${tag}_${kind} \
      python tools/prof_driver.py $kind 3 > gpurun_out/ncu_${kind}_${tag}.log 2>&1
done
python tools/prof_phong.py realistic 3 > gpurun_out/plain_phong_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:phong_ -s 2 -c 2 -f -o gpurun_out/prof_${tag}_phong \
    python tools/prof_phong.py realistic 3 > gpurun_out/ncu_phong_${tag}.log 2>&1
python tools/time_raster.py > gpurun_out/plain_raster_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rasterize_ -s 1 -c 2 -f -o gpurun_out/prof_${tag}_raster \
    python tools/time_raster.py > gpurun_out/ncu_raster_${tag}.log 2>&1
ls -la gpurun_out | tail -20
