"""Round-2 GPU parity cases: the BASELINE configurations that had no test (4K dense with the default tables, 1080p with
many distinct frames through both decode paths), the sharded multi-device entry, STAGED=2 with several chunks in flight,
the remaining accelerator-seam symbol and the length-less lossless_decode() on an exact-size buffer.  Needs a B200."""
import ctypes as C
import mmap
import os

import numpy as np
import pytest

import mjpeg423_b200
from mjpeg423_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dec():
    d = mjpeg423_b200.Decoder(0)
    yield d
    d.close()


def _resident(dec, mpg, shape):
    dec.upload(mpg)
    nbytes = int(np.prod(shape))
    d_out = dec.device_alloc(nbytes)
    try:
        dec.decode_resident(d_out)
        return dec.to_host(d_out, nbytes).reshape(shape)
    finally:
        dec.device_free(d_out)


def test_stream_4k_dense_default_tables(checker, dec):
    """BASELINE configs[3]: 3840x2160, full-range noise, default tables (entropy-bound, ~11.5 bit/px)."""
    mpg = synth.synth_mpg(3840, 2160, 4, 0, 256, 0)
    assert mpg.size / 4 > 10e6
    want = checker.decode_mpg(mpg, nthreads=min(4, os.cpu_count() or 1))
    assert np.array_equal(dec.decode_frames(mpg), want)
    assert np.array_equal(_resident(dec, mpg, want.shape), want)


def test_stream_1080p_64_distinct_frames_both_paths(checker, dec):
    """BASELINE configs[2] at test scale: 64 DISTINCT 1080p frames through the host-buffer call and the resident path."""
    mpg = synth.synth_mpg(1920, 1080, 64, 0, 16, 0)
    want = checker.decode_mpg(mpg, nthreads=os.cpu_count() or 1)
    assert len({want[i].tobytes()[:4096] for i in range(64)}) == 64
    out = dec.pinned(want.nbytes)
    try:
        assert np.array_equal(dec.decode_frames(mpg, out=out), want)
    finally:
        out.free()
    dec.set_option(api.OPT_CHUNK_FRAMES, 24)                 # three chunks: two buffers in flight
    try:
        assert np.array_equal(_resident(dec, mpg, want.shape), want)
    finally:
        dec.set_option(api.OPT_CHUNK_FRAMES, 0)


@pytest.mark.parametrize("chunk", [1, 3, 7])
@pytest.mark.parametrize("gop", [1, 6])
def test_staged2_chunked_resident(checker, dec, chunk, gop):
    """STAGED=2 with several chunks in flight shared one sample buffer between the two streams."""
    W, H, n = 320, 240, 29
    fr = np.stack([synth.synth_frame(W, H, i, 24) for i in range(n)])
    mpg = synth.encode_mpg(fr, gop=gop)
    want = checker.decode_mpg(mpg)
    dec.set_option(api.OPT_STAGED, 2)
    dec.set_option(api.OPT_CHUNK_FRAMES, chunk)
    try:
        for _ in range(3):                                   # the race was nondeterministic
            assert np.array_equal(_resident(dec, mpg, want.shape), want)
    finally:
        dec.set_option(api.OPT_STAGED, 0)
        dec.set_option(api.OPT_CHUNK_FRAMES, 0)


def test_sharded_decode_matches_single_device(checker, dec):
    """SURVEY 8e through the C-ABI: pieces cut on I frames, one host thread per device slot, every slot writing its own
    slice.  On a one-GPU box both slots are device 0 (two contexts, two threads); with two GPUs the second slot is GPU 1."""
    W, H, n = 320, 240, 61
    fr = np.stack([synth.synth_frame(W, H, i, 24) for i in range(n)])
    mpg = synth.encode_mpg(fr, gop=8)                        # I frames every 8: cuts must move to them
    want = dec.decode_frames(mpg)
    assert np.array_equal(want, checker.decode_mpg(mpg))
    ndev = api.load_library().mjpeg423_b200_device_count()
    for devices in ([0], [0, 0], [0, 1 % ndev, 0], [0] * 5):
        got, cuts = mjpeg423_b200.decode_frames_multi(mpg, devices)
        assert np.array_equal(got, want), devices
        assert cuts[0] == 0 and cuts[-1] == n and all(int(c) % 8 == 0 or int(c) == n for c in cuts)
    # a frame range that starts on a later I frame; a P-frame start is refused
    got, cuts = mjpeg423_b200.decode_frames_multi(mpg, [0, 0], first=16, n=30)
    assert np.array_equal(got, want[16:46]) and int(cuts[1]) in (32, 40)
    with pytest.raises(RuntimeError, match="P frame"):
        mjpeg423_b200.decode_frames_multi(mpg, [0, 0], first=3, n=10)
    # a SET of files as one logical stream (the container's offsets are 32-bit: long streams come in pieces)
    a, b = synth.encode_mpg(fr[:24], gop=8), synth.encode_mpg(fr[24:], gop=8)
    got, cuts = mjpeg423_b200.decode_frames_multi([a, b], [0, 0, 0])
    assert np.array_equal(got, want)                         # (frame 24 is an I frame in both encodings)
    got, _ = mjpeg423_b200.decode_frames_multi([a, b], [0, 0], first=16, n=20)
    assert np.array_equal(got, want[16:36])


def test_accel_calculate_buffer(checker):
    """C0/idct_ycbcr_to_rgb_accel.h:19-20: block planes -> a raster wider than the converted area."""
    lib = api.load_library()
    assert lib.init_idct_ycbcr_to_rgb_accel() == 1
    rng = np.random.default_rng(9)
    hb, wb, w_size = 3, 5, 64
    Y, Cb, Cr = (rng.integers(0, 256, size=(hb * wb, 8, 8), dtype=np.uint8) for _ in range(3))
    out = np.full((hb * 8, w_size, 4), 0x5A, dtype=np.uint8)
    lib.ycbcr_to_rgb_accel_calculate_buffer(Y.ctypes.data, Cr.ctypes.data, Cb.ctypes.data, out.ctypes.data, hb, wb, w_size)
    lib.wait_for_ycbcr_to_rgb_finsh()
    want = checker.ycbcr_to_rgb(Y, Cb, Cr, wb * 8, hb * 8)
    assert np.array_equal(out[:, :wb * 8], want)
    assert np.all(out[:, wb * 8:] == 0x5A)                    # nothing outside the converted area is touched


def test_lossless_decode_reads_nothing_past_the_stream(checker):
    """the reference signature has no length, and the shim used to read num_blocks*152+8 bytes.  Here the
    stream ends on the last byte of a page whose successor is unmapped-for-access: an over-read would fault."""
    rng = np.random.default_rng(21)
    nb = 40
    coef = np.zeros((nb, 64), dtype=np.int64)
    coef[:, 0] = rng.integers(-30, 30, nb)
    for b in range(nb):
        for k in rng.choice(np.arange(1, 64), size=5, replace=False):
            coef[b, k] = rng.integers(-9, 10) or 1
    bits = []
    def put(v, n):
        bits.extend((v >> i) & 1 for i in range(n - 1, -1, -1))
    def vli(v):
        s = int(abs(v)).bit_length()
        return s, (v if v > 0 else v - 1) & ((1 << s) - 1)
    prev = 0
    for b in range(nb):
        s, a = vli(int(coef[b, 0]) - prev); prev = int(coef[b, 0])
        put(s, 4); put(a, s)
        run = 0
        for k in range(1, 64):
            if coef[b, k] == 0:
                run += 1
                continue
            while run > 15:
                put(0xF0, 8); run -= 16
            s, a = vli(int(coef[b, k]))
            put(run, 4); put(s, 4); put(a, s); run = 0
        if coef[b, 63] == 0:
            put(0, 8)
    while len(bits) % 8:
        bits.append(0)
    stream = np.packbits(np.array(bits, dtype=np.uint8))
    want = checker.lossless_decode(nb, stream.tobytes() + b"\0" * 8, api.YQUANT, 0)
    page = mmap.PAGESIZE
    libc = C.CDLL(None, use_errno=True)
    libc.mmap.restype = C.c_void_p
    libc.mmap.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_long]
    libc.mprotect.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
    libc.munmap.argtypes = [C.c_void_p, C.c_size_t]
    base = libc.mmap(None, 2 * page, mmap.PROT_READ | mmap.PROT_WRITE, mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS, -1, 0)
    assert base not in (None, C.c_void_p(-1).value)
    try:
        assert libc.mprotect(base + page, page, 0) == 0
        start = base + page - stream.size
        C.memmove(start, stream.ctypes.data, stream.size)
        out = np.zeros((nb, 8, 8), dtype=np.int16)
        q = np.ascontiguousarray(api.YQUANT)
        api.load_library().lossless_decode(nb, start, out.ctypes.data, q.ctypes.data, 0)
        assert np.array_equal(out, want)
    finally:
        libc.munmap(base, 2 * page)
