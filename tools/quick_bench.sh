#!/bin/bash
# usage: tools/quick_bench.sh tag   (run on the GPU box through gpurun)
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$1.json"))
r=d["roofline"]; o=d["also"]["roofline"]
print("realistic ms/step %.4f fwd %.4f (%.3f) bwd %.4f (%.3f)" % (d["ms_per_step"], r["fwd"]["ms"], r["fwd"]["frac"], r["bwd"]["ms"], r["bwd"]["frac"]))
print("dense     ms/step %.4f fwd %.4f (%.3f) bwd %.4f (%.3f)" % (d["also"]["ms_per_step"], o["fwd"]["ms"], o["fwd"]["frac"], o["bwd"]["ms"], o["bwd"]["frac"]))
PY
tail -3 gpurun_out/bench_$1.err
