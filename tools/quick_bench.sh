#!/bin/bash
# usage: tools/quick_bench.sh tag   (run on the GPU box through gpurun)
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$1.json"))
r=d["roofline"]; o=d["also"]["fragments_dense"]["roofline"]; ps=d["also"]["per_sample_noise"]
print("realistic ms/step %.4f fwd %.4f (%.3f) bwd %.4f (%.3f)" % (d["ms_per_step"], r["fwd"]["ms"], r["fwd"]["frac"], r["bwd"]["ms"], r["bwd"]["frac"]))
print("dense     ms/step %.4f fwd %.4f (%.3f) bwd %.4f (%.3f)" % (d["also"]["fragments_dense"]["ms_per_step"], o["fwd"]["ms"], o["fwd"]["frac"], o["bwd"]["ms"], o["bwd"]["frac"]))
print("persample ms/step %.4f bwd %.4f (%.3f)" % (ps["ms_per_step"], ps["roofline"]["bwd"]["ms"], ps["roofline"]["bwd"]["frac"]))
fc=d["also"]["face_colour_gather"]
print("facecol   ms/step %.4f fwd %.4f (%.3f) bwd %.4f (%.3f)" % (fc["ms_per_step"], fc["roofline"]["fwd"]["ms"], fc["roofline"]["fwd"]["frac"], fc["roofline"]["bwd"]["ms"], fc["roofline"]["bwd"]["frac"]))
sf=d["also"]["softras_pair"]
print("softras   ms/step %.4f fwd %.4f (%.3f) bwd %.4f (%.3f)" % (sf["ms_per_step"], sf["roofline"]["fwd"]["ms"], sf["roofline"]["fwd"]["frac"], sf["roofline"]["bwd"]["ms"], sf["roofline"]["bwd"]["frac"]))
ph=d["also"]["random_phong_shader"]
print("phong     ms/step %.4f phong_fwd %.4f (%.3f) shade %.4f + %.4f phong_bwd %.4f (%.3f)" % (ph["ms_per_step"], ph["phong_fwd"]["ms"], ph["phong_fwd"]["frac"], ph["shade_fwd_ms"], ph["shade_bwd_ms"], ph["phong_bwd"]["ms"], ph["phong_bwd"]["frac"]))
rz=d["also"]["fragments_rasterised"]; rn=d["also"]["renderer"]
print("rasterised ms/step %.4f fwd %.4f bwd %.4f" % (rz["ms_per_step"], rz["roofline"]["fwd"]["ms"], rz["roofline"]["bwd"]["ms"]))
print("renderer  ms/step %.4f raster %.4f shade_fwd %.4f backward %.4f valid/px %.1f" % (rn["ms_per_step"], rn["rasterize_ms"], rn["shade_fwd_ms"], rn["backward_ms"], rn["valid_per_covered_pixel"]))
print("clocks", d["clocks"])
PY
tail -3 gpurun_out/bench_$1.err
