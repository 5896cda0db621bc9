"""Host-link ceiling of the box: plain device -> pinned-host copies on EVERY GPU at once, no kernels (what bounds the
end-to-end number of bench.py when several ranks read frames back).  Also tries write-combined pinned memory.

    python tools/pcie_probe_multi.py [n_gpus]          -> one JSON line
"""
import ctypes as C
import json
import sys
import time

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
size = 1 << 30
rt = C.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaFreeHost.argtypes = [C.c_void_p]
out = {"gpus": n, "bytes_per_copy": size}
for name, flags in (("pinned", 0), ("pinned_write_combined", 4)):
    devs, hosts, streams = [], [], []
    for g in range(n):
        torch.cuda.set_device(g)
        devs.append(torch.empty(size, dtype=torch.uint8, device=f"cuda:{g}"))
        p = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(p), size, flags) == 0
        hosts.append(p)
        streams.append(torch.cuda.Stream(device=g))
    res = {}
    for active in sorted({1, n}):
        for rep in range(3):
            for g in range(active):
                torch.cuda.synchronize(g)
            t0 = time.perf_counter()
            for g in range(active):
                torch.cuda.set_device(g)
                for _ in range(4):
                    assert rt.cudaMemcpyAsync(hosts[g], C.c_void_p(devs[g].data_ptr()), size, 2, C.c_void_p(streams[g].cuda_stream)) == 0
            for g in range(active):
                torch.cuda.synchronize(g)
            dt = time.perf_counter() - t0
        res[f"{active}_gpus_aggregate_GBs"] = round(active * 4 * size / dt / 1e9, 2)
    out[name] = res
    for p in hosts:
        rt.cudaFreeHost(p)
    del devs
print(json.dumps(out))
