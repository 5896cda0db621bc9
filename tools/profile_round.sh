#!/bin/bash
# Round-end evidence: bench lines for every workload, the reference arm, the ncu launch list of the bench command
# and one `ncu --set full` capture of every pipeline kernel.  usage: tools/profile_round.sh <tag>
TAG=${1:-r01x}
tools/bench_all.sh 1080p 4k 480p 4k-q1
python bench.py --impl reference > gpurun_out/bench_reference.log 2>gpurun_out/bench_reference.err; tail -c 400 gpurun_out/bench_reference.log
python bench.py --workload enc-1080p > gpurun_out/bench_enc-1080p.log 2>gpurun_out/bench_enc.err
python bench.py --workload enc-1080p --impl reference --steps 2 > gpurun_out/bench_enc_reference.log 2>gpurun_out/bench_enc_reference.err; tail -c 300 gpurun_out/bench_enc_reference.log; tail -2 gpurun_out/bench_enc_reference.err
# launch list of the bench command (times under ncu are cold-cache and serialised: only the SHARES are meaningful)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
tools/ncu_gpu.sh $TAG 1000 k_ 6 5
