"""The C-ABI library loads and exports every symbol include/mjpeg423_b200.h declares; host-only entry
points (container probe) work; compute entry points fail loudly without a GPU.  CPU only."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mjpeg423_b200
from mjpeg423_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return api.load_library(build_if_missing=True)


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "mjpeg423_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", hdr)) - {"defined", "sizeof", "void"}   # "void (" = a function-pointer parameter
    declared |= {"Yquant", "Cquant", "zigzag_table"}
    declared = {d for d in declared if not d.isupper() and not d.endswith("_t")}
    assert declared == set(api.EXPORTS), declared ^ set(api.EXPORTS)
    for name in api.EXPORTS:
        assert hasattr(lib, name), name


def test_exported_tables_match_reference_values(lib):
    yq = np.ctypeslib.as_array((C.c_int16 * 64).in_dll(lib, "Yquant"))
    cq = np.ctypeslib.as_array((C.c_int16 * 64).in_dll(lib, "Cquant"))
    zz = np.ctypeslib.as_array((C.c_int * 64).in_dll(lib, "zigzag_table"))
    assert np.array_equal(yq, api.YQUANT.ravel()) and np.array_equal(cq, api.CQUANT.ravel())
    assert np.array_equal(zz, api.ZIGZAG) and sorted(zz) == list(range(64))
    assert yq[0] == 16 and cq[0] == 17 and yq[63] == 99


def test_probe_parses_container():
    fr = np.stack([synth.synth_frame(32, 16, i, 16) for i in range(5)])
    mpg = synth.encode_mpg(fr, gop=3)
    info = mjpeg423_b200.probe(mpg)
    assert (info.num_frames, info.w_size, info.h_size) == (5, 32, 16)
    assert info.num_iframes == 2 and info.num_pframes == 3 and info.frame_bytes == 32 * 16 * 4


def test_probe_rejects_bad_input():
    fr = synth.synth_frame(32, 16, 0, 16)[None]
    mpg = synth.encode_mpg(fr)
    with pytest.raises(RuntimeError):
        mjpeg423_b200.probe(mpg[:10])                 # shorter than the file header
    with pytest.raises(RuntimeError):
        mjpeg423_b200.probe(mpg[:40])                 # truncated inside frame 0
    bad = mpg.copy()
    bad[4:8] = np.frombuffer(np.uint32(30).tobytes(), np.uint8)   # width not a multiple of 8
    with pytest.raises(RuntimeError):
        mjpeg423_b200.probe(bad)
    bad = mpg.copy()
    bad[28:32] = 255                                   # Ysize larger than the frame
    with pytest.raises(RuntimeError):
        mjpeg423_b200.probe(bad)


def test_no_gpu_fails_loudly(lib):
    if lib.mjpeg423_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA device"):
        mjpeg423_b200.Decoder()
    with pytest.raises(RuntimeError):
        mjpeg423_b200.idct(np.zeros((1, 8, 8), np.int16))
    h = C.c_void_p()
    assert lib.mjpeg423_b200_create(C.byref(h), 0) == api.E_CUDA
    assert b"no CUDA device" in lib.mjpeg423_b200_last_error()


def test_product_package_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "mjpeg423-video-decoder-software_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)


def test_frame_hash_host_is_position_sensitive():
    a = np.arange(2 * 64, dtype=np.uint8).reshape(2, 64)
    b = a.copy()
    b[0, :8], b[0, 8:16] = a[0, 8:16].copy(), a[0, :8].copy()   # swap two words
    ha, hb = api.frame_hash_host(a), api.frame_hash_host(b)
    assert ha[0] != hb[0] and ha[1] == hb[1]


def test_header_is_c99_and_example_links(tmp_path, lib):
    """include/mjpeg423_b200.h must be plain C (the reference is C99): the sample caller compiles with gcc -std=c99
    -pedantic-errors and links against the shared library (no GPU needed to link)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = tmp_path / "decode_file"
    pkg = os.path.join(ROOT, "mjpeg423-video-decoder-software_b200")
    for src in ("decode_file.c", "seek_and_shard.c"):
        cmd = ["gcc", "-std=c99", "-pedantic-errors", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
               os.path.join(ROOT, "examples", src), "-L", pkg, "-lmjpeg423_b200", f"-Wl,-rpath,{pkg}", "-o", str(exe)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        out = subprocess.run([str(exe)], capture_output=True, text=True)
        assert out.returncode == 2 and "usage" in out.stderr
