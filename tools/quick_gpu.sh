#!/bin/bash
# quick GPU check: parity tests + per-stage timing of a resident 1080p decode (no ncu)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/profile_run.py --frames ${1:-2000} --passes 3 2>&1 | tail -1 | cut -c1-150
