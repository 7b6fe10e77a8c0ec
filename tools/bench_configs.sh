#!/bin/bash
# The five BASELINE configs on one GPU (per-GPU share for the sharded ones); one JSON line each.
run() { echo "== $1"; shift; python bench.py --no-e2e --no-cpu-baseline --no-renderer-legs --steps 5 --warmup 3 "$@" 2>gpurun_out/cfg.err | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print('value %.4g units/s  ms/step %.4f  fwd %.4f ms (%.3f)  bwd %.4f ms (%.3f)  fwd+bwd frac %.3f' % (d['value'], d['ms_per_step'], r['fwd']['ms'], r['fwd']['frac'], r['bwd']['ms'], r['bwd']['frac'], r['fwd_bwd_frac']))
for k,v in d['also'].items(): print('   also %-20s ms/step %.4f  frac %.3f' % (k, v['ms_per_step'], v['roofline']['fwd_bwd_frac']))
"; tail -2 gpurun_out/cfg.err; }
run "cfg1 1x64^2 K50 S16" --views 1 --image-size 64 --faces-per-pixel 50 --nb-samples 16
run "cfg2 8x256^2 K50 S64" --views 8 --image-size 256 --faces-per-pixel 50 --nb-samples 64
run "cfg3 per-GPU share 8x512^2 K100 S256" --views 8 --image-size 512 --faces-per-pixel 100 --nb-samples 256
run "cfg4 1x128^2 K50 S4096 (one GPU, all samples)" --views 1 --image-size 128 --faces-per-pixel 50 --nb-samples 4096
run "cfg5 16x1024^2 K50 S32" --views 16 --image-size 1024 --faces-per-pixel 50 --nb-samples 32
