#!/bin/bash
# usage: tools/quick_bench.sh tag [extra bench args]  (run on the GPU box through gpurun)
tag=$1; shift
python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python tools/bench_summary.py gpurun_out/bench_$tag.json
tail -3 gpurun_out/bench_$tag.err
