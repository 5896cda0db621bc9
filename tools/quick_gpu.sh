#!/bin/bash
# quick GPU check: parity tests + per-stage timing of resident decodes (no ncu)
# usage: tools/quick_gpu.sh [frames-1080p] [more workloads as name:frames ...]     (SKIP_TESTS=1 skips pytest)
[ -n "$SKIP_TESTS" ] || python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/profile_run.py --frames ${1:-2000} --passes 3 2>&1 | tail -2 | cut -c1-150
shift
for wf in "$@"; do
  echo "$wf: $(python tools/profile_run.py --workload ${wf%%:*} --frames ${wf##*:} --passes 3 2>&1 | tail -2 | cut -c1-150)"
done
