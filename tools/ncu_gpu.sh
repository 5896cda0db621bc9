#!/bin/bash
# one ncu --set full capture of the second pass of tools/profile_run.py (every kernel of the pipeline once)
# usage: tools/ncu_gpu.sh <tag> [frames] [kernel regex]
TAG=${1:-x}; FR=${2:-256}; K=${3:-k_}
python tools/profile_run.py --frames $FR --passes 2 --profile 0 > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
NL=$(python - <<PY
print(5)
PY
)
ncu --set full --clock-control none --import-source on -k regex:$K -s 5 -c 5 -o gpurun_out/prof_$TAG -f \
    python tools/profile_run.py --frames $FR --passes 2 --profile 0 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
