#!/bin/bash
# one ncu --set full capture of the second pass of tools/profile_run.py
# usage: tools/ncu_gpu.sh <tag> [frames] [kernel regex] [skip] [count]
TAG=${1:-x}; FR=${2:-256}; K=${3:-k_}; SKIP=${4:-6}; CNT=${5:-5}
python tools/profile_run.py --frames $FR --passes 2 --profile 0 > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -o gpurun_out/prof_$TAG -f \
    python tools/profile_run.py --frames $FR --passes 2 --profile 0 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
