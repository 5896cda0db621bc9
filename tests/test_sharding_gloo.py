"""Multi-rank host logic on CPU: two gloo processes shard a stream by frame range exactly as bench.py does
on N GPUs (contiguous ranges, every rank writes its own slice, no data-path collective), and the MAX / SUM
reductions bench.py uses for its timing and counts.  The per-rank decode here is the CPU oracle; the GPU
decode of a sub-range is covered by tests/test_gpu_parity.py."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["REPO_ROOT"])
import torch, torch.distributed as dist
import bench
import mjpeg423_b200
from mjpeg423_b200 import api, synth
from oracle import oracle

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
total = 11
mpg = synth.synth_mpg(64, 48, total, 0, 16, 0, nthreads=1)
lo, hi = bench.shard_range(total, rank, world)
mine = oracle.port().decode_mpg(mpg, lo, hi - lo)                 # this rank's own output slice
h = torch.zeros(total, dtype=torch.int64)
h[lo:hi] = torch.from_numpy(api.frame_hash_host(mine).view(np.int64))
dist.all_reduce(h)                                                # verification only: ranges are disjoint
# the optional single contiguous output (bench.py --gather): uneven slices, point-to-point into rank 0's buffer
fb = mine[0].nbytes
full = bench.gather_frames(dist, torch.from_numpy(np.ascontiguousarray(mine).reshape(-1)), [bench.shard_range(total, r, world)[1] - bench.shard_range(total, r, world)[0] for r in range(world)], fb)
if rank == 0:
    assert np.array_equal(full.numpy().reshape(total, 48, 64, 4), oracle.port().decode_mpg(mpg)), "gathered output differs"
(tmax,), (frames, launches) = bench.reduce_over_ranks(dist, [10.0 * (rank + 1)], [hi - lo, 4 * (hi - lo)], "cpu")
if rank == 0:
    full = api.frame_hash_host(oracle.port().decode_mpg(mpg)).view(np.int64)
    assert np.array_equal(h.numpy(), full), "sharded decode differs from the single-rank decode"
    assert tmax == 10.0 * world and frames == total and launches == 4 * total
    print("OK", lo, hi, tmax, frames)
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_frame_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, REPO_ROOT=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "OK 0 5 20.0 11.0" in res.stdout


def test_shard_ranges_partition():
    sys.path.insert(0, ROOT)
    import bench
    for total in (0, 1, 7, 8192):
        for world in (1, 2, 3, 8):
            r = [bench.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
