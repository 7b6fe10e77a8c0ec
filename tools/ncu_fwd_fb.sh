#!/bin/bash
# ncu --set full of the forward fallback launches on the rasterised set:  tools/ncu_fwd_fb.sh <tag> [flags] [env...]
tag=$1; flags=${2:-0}
python tools/prof_driver.py rasterised 2 $flags > gpurun_out/plain_fb_${tag}.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:shade_fwd_fallback -s 2 -c 2 -f -o /tmp/fb_${tag} \
    python tools/prof_driver.py rasterised 2 $flags > gpurun_out/ncu_fb_${tag}.log 2>&1
python profiles/ncu_summary.py /tmp/fb_${tag}.ncu-rep > gpurun_out/fb_${tag}_summary.txt 2>&1
python profiles/ncu_lines.py /tmp/fb_${tag}.ncu-rep shade_ 30 > gpurun_out/fb_${tag}_lines.txt 2>&1
