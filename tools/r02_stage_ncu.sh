#!/bin/bash
# Per-stage ncu of the staged kernels (IDCT 9P, colour 7P, fused IDCT+colour 10P) at 1080p x 1000 frames.
# usage: tools/r02_stage_ncu.sh <tag>
TAG=${1:-r02a}
python tools/profile_run.py --frames 1000 --staged 2 --passes 2 --profile 1 > gpurun_out/staged2_$TAG.log 2>&1; tail -2 gpurun_out/staged2_$TAG.log | cut -c1-400
python tools/profile_run.py --frames 1000 --staged 1 --passes 2 --profile 1 > gpurun_out/staged1_$TAG.log 2>&1; tail -2 gpurun_out/staged1_$TAG.log | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:'k_idct|k_colour' -s 10 -c 2 -o gpurun_out/prof_staged2_$TAG -f \
    python tools/profile_run.py --frames 1000 --staged 2 --passes 2 --profile 0 > gpurun_out/ncu_staged2_$TAG.log 2>&1
tail -2 gpurun_out/ncu_staged2_$TAG.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:k_idct_colour -s 5 -c 1 -o gpurun_out/prof_staged1_$TAG -f \
    python tools/profile_run.py --frames 1000 --staged 1 --passes 2 --profile 0 > gpurun_out/ncu_staged1_$TAG.log 2>&1
tail -2 gpurun_out/ncu_staged1_$TAG.log | cut -c1-200
