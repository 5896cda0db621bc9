"""Bit-exact parity of the CUDA path (through the C-ABI) against the CPU oracle.  Needs a B200.

The checker is the compiled reference (oracle/_ref) when it travelled with the snapshot, else the C
restatement; nothing here reads /root/reference."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import mjpeg423_b200
from mjpeg423_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "golden_small.npz"))


@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def dec():
    d = mjpeg423_b200.Decoder(0)
    yield d
    d.close()


def _quant(name):
    return api.YQUANT if name == "Y" else api.CQUANT


# ---- known answers --------------------------------------------------------------
def test_kat_entropy(kat):
    for k in kat["entropy"]:
        c = mjpeg423_b200.lossless_decode(k["nb"], bytes.fromhex(k["hex"]), None, _quant(k["quant"]), k["P"]).ravel()
        want = np.zeros(k["nb"] * 64, dtype=np.int16)
        for i, v in k["expect"].items():
            want[int(i)] = v
        assert np.array_equal(c, want), k["name"]
    k = kat["entropy_P"]
    buf = mjpeg423_b200.lossless_decode(1, bytes.fromhex(k["hex"]), None, api.YQUANT, 0)
    buf = mjpeg423_b200.lossless_decode(1, bytes.fromhex(k["hex"]), buf, api.YQUANT, 1).ravel()
    assert buf[0] == 64 and buf[1] == -154 and np.count_nonzero(buf) == 2


def test_kat_idct_and_colour(kat):
    for k in kat["idct"]:
        c = np.zeros(64, dtype=np.int16)
        for i, v in k["coef"].items():
            c[int(i)] = v
        assert np.array_equal(mjpeg423_b200.idct(c.reshape(8, 8)), np.array(k["rows"], dtype=np.uint8)), k["name"]
    r, kk = np.meshgrid(np.arange(8), np.arange(8), indexing="ij")
    Y, Cb, Cr = (32 * r + 3 * kk).astype(np.uint8), (255 - 30 * r).astype(np.uint8), (36 * kk).astype(np.uint8)
    raster = np.full((16, 24, 4), 0xAB, dtype=np.uint8)
    mjpeg423_b200.ycbcr_to_rgb(8, 16, 24, Y, Cb, Cr, raster)        # tile placed at row 8, column 16
    words = raster[8:, 16:].copy().view("<u4")[..., 0]
    assert [f"{w:08x}" for w in words[0]] == kat["colour"]["row0"]
    assert [f"{w:08x}" for w in words[7]] == kat["colour"]["row7"]
    assert np.all(raster[:8] == 0xAB) and np.all(raster[8:, :16] == 0xAB)   # nothing else touched


def test_reference_symbol_lossless_decode_without_length():
    """The exact reference signature (no length): the shim reads up to the configured limit."""
    lib = api.load_library()
    stream = np.zeros(400, dtype=np.uint8)
    stream[:3] = [0x28, 0x0C, 0x00]
    out = np.zeros((1, 8, 8), dtype=np.int16)
    q = np.ascontiguousarray(api.YQUANT)
    lib.lossless_decode(1, stream.ctypes.data, out.ctypes.data, q.ctypes.data, 0)
    assert out.ravel()[0] == 32 and out.ravel()[1] == -77
    lib.mjpeg423_b200_set_read_limit(16)
    assert lib.mjpeg423_b200_get_read_limit() == 16
    out[:] = 7
    lib.lossless_decode(1, stream.ctypes.data, out.ctypes.data, q.ctypes.data, 0)
    assert out.ravel()[0] == 32 and out.ravel()[2] == 0
    lib.mjpeg423_b200_set_read_limit(0)


# ---- committed reference fixtures -----------------------------------------------------------------------
def test_golden_stage_functions(gold):
    assert np.array_equal(mjpeg423_b200.idct(gold["coef"]), gold["samp"])
    ycc = gold["ycc"]
    assert np.array_equal(api.ycbcr_to_rgb_frame(ycc[0], ycc[1], ycc[2], 128, 128), gold["bgra"])
    got = mjpeg423_b200.lossless_decode(300, gold["ystream"].tobytes(), None, api.YQUANT, 0)
    assert np.array_equal(got, gold["ystream_coef"])


@pytest.mark.parametrize("staged", [0, 1, 2])
def test_golden_streams(gold, dec, staged):
    dec.set_option(api.OPT_STAGED, staged)
    try:
        assert np.array_equal(dec.decode_frames(gold["mpg_flat"]), gold["frames_flat"])
        ones = np.ones(64, dtype=np.int16)
        dec.set_quant(ones, ones)
        assert np.array_equal(dec.decode_frames(gold["mpg_dense"]), gold["frames_dense"])
        dec.set_quant(None, None)
        assert mjpeg423_b200.probe(gold["mpg_ip"]).num_pframes == 3          # I P P I P
        assert np.array_equal(dec.decode_frames(gold["mpg_ip"]), gold["frames_ip"])
        assert np.array_equal(dec.decode_frames(gold["mpg_ip"], first=3, n=2), gold["frames_ip"][3:])
        with pytest.raises(RuntimeError, match="P frame"):
            dec.decode_frames(gold["mpg_ip"], first=1, n=2)                   # a range must start on an I frame
    finally:
        dec.set_quant(None, None)
        dec.set_option(api.OPT_STAGED, 0)


# ---- randomised property tests against the oracle -------------------------------------------------------
def test_idct_random_blocks(checker):
    rng = np.random.default_rng(11)
    for lo, hi, n in ((-32768, 32768, 5000), (-2048, 2048, 5000), (-64, 64, 3001)):
        coef = rng.integers(lo, hi, size=(n, 8, 8)).astype(np.int16)
        coef[rng.random(n) < 0.3, 1:, :] = 0          # DC-only-column shortcut territory
        assert np.array_equal(mjpeg423_b200.idct(coef), checker.idct(coef))
    edge = np.zeros((4, 8, 8), np.int16)
    edge[0] = 32767; edge[1] = -32768; edge[2, 0, 0] = 32767; edge[3, 7, 7] = -32768
    assert np.array_equal(mjpeg423_b200.idct(edge), checker.idct(edge))
    assert mjpeg423_b200.idct(np.zeros((0, 8, 8), np.int16)).size == 0      # empty input


def test_colour_exhaustive_cube(port):
    """Every (Y, Cb, Cr) triple: 2^24 pixels = 262144 blocks, compared with the oracle."""
    g = np.arange(256, dtype=np.uint8)
    Y = np.broadcast_to(g[:, None, None], (256, 256, 256)).reshape(-1, 8, 8)
    Cb = np.broadcast_to(g[None, :, None], (256, 256, 256)).reshape(-1, 8, 8)
    Cr = np.broadcast_to(g[None, None, :], (256, 256, 256)).reshape(-1, 8, 8)
    W, H = 4096, 4096                                   # 512 x 512 blocks = 262144
    got = api.ycbcr_to_rgb_frame(Y, Cb, Cr, W, H)
    # oracle on the same blocks, vectorised restatement checked against the C oracle on a slice
    y, cb, cr = (a.astype(np.int32) for a in (Y, Cb, Cr))
    cb -= 128; cr -= 128
    def sat(t):
        return np.where(t < 0, 0, np.minimum(t >> 14, 255)).astype(np.uint8)
    yy = y << 14
    want_blocks = np.stack([sat(yy + 29032 * cb), sat(yy - 5638 * cb - 11700 * cr), sat(yy + 22970 * cr),
                            np.zeros_like(Y)], axis=-1)           # (nb, 8, 8, 4) BGRA
    want = want_blocks.reshape(512, 512, 8, 8, 4).transpose(0, 2, 1, 3, 4).reshape(H, W, 4)
    sl = slice(1000, 1512)
    chk = port.ycbcr_to_rgb(Y[sl], Cb[sl], Cr[sl], 8 * 512, 8)
    assert np.array_equal(chk, want_blocks[sl].reshape(1, 512, 8, 8, 4).transpose(0, 2, 1, 3, 4).reshape(8, 4096, 4))
    assert np.array_equal(got, want)


def _random_levels(rng, nb, density, amp, dcamp):
    lv = (rng.integers(-amp, amp + 1, size=(nb, 64)) * (rng.random((nb, 64)) < density)).astype(np.int16)
    lv[:, 0] = rng.integers(-dcamp, dcamp + 1, size=nb)
    return lv


def _encode_levels(levels):
    """From-spec entropy coder in Python for small random cases."""
    bits = []
    def put(v, n):
        for i in range(n - 1, -1, -1):
            bits.append((v >> i) & 1)
    def vli(x):
        s = int(abs(int(x))).bit_length()
        return s, (int(x) if x > 0 else int(x) - 1) & ((1 << s) - 1)
    for blk in levels:
        s, a = vli(blk[0]); put(s, 4); put(a, s)
        z = [int(blk[n]) for n in api.ZIGZAG]
        last = 63
        while last > 0 and z[last] == 0:
            last -= 1
        run = 0
        for k in range(1, last + 1):
            if z[k] == 0:
                run += 1
                if run == 16:
                    put(0xF0, 8); run = 0
                continue
            s, a = vli(z[k]); put(run, 4); put(s, 4); put(a, s); run = 0
        if last < 63:
            put(0, 8)
    while len(bits) % 8:
        bits.append(0)
    return np.packbits(np.array(bits, dtype=np.uint8)).tobytes()


@pytest.mark.parametrize("nb,density,amp,dcamp,seed", [
    (1, 0.0, 1, 0, 0), (1, 1.0, 2047, 2047, 1), (37, 0.05, 20, 100, 2), (500, 0.15, 40, 300, 3),
    (3000, 0.02, 5, 10, 4), (2000, 0.9, 2047, 2047, 5), (4800, 0.1, 30, 200, 6), (171, 0.0, 1, 0, 7),
])
def test_entropy_random_streams(checker, nb, density, amp, dcamp, seed):
    rng = np.random.default_rng(seed)
    lv = _random_levels(rng, nb, density, amp, dcamp)
    wire = lv.copy()
    wire[1:, 0] = lv[1:, 0] - lv[:-1, 0]
    wire[:, 0] = np.clip(wire[:, 0], -2047, 2047)
    stream = _encode_levels(wire)
    for q in (api.YQUANT, api.CQUANT, np.ones(64, np.int16)):
        want = checker.lossless_decode(nb, stream, q, 0)
        got = mjpeg423_b200.lossless_decode(nb, stream, None, q, 0)
        assert np.array_equal(got, want)
    # P frame: deltas accumulate onto an arbitrary previous state, int16 wrap included
    prev = rng.integers(-32768, 32768, size=(nb, 8, 8)).astype(np.int16)
    want = checker.lossless_decode(nb, stream, api.YQUANT, 1, DCACq=prev.copy())
    got = mjpeg423_b200.lossless_decode(nb, stream, prev.copy(), api.YQUANT, 1)
    assert np.array_equal(got, want)


def test_entropy_size15_amplitudes_and_wrap(checker):
    """Sizes 12..15 never come out of the reference encoder but its decoder accepts them (int16 wrap)."""
    bits = []
    def put(v, n):
        for i in range(n - 1, -1, -1):
            bits.append((v >> i) & 1)
    for s, a in ((15, 0x7FFF), (15, 0x0000), (12, 0xABC), (13, 0x1), (14, 0x3FFF)):
        put(s, 4); put(a, s)            # DC
        put(2, 4); put(s, 4); put(a ^ 0x155, s)    # one AC at run 2
        put(0, 8)                        # END
    while len(bits) % 8:
        bits.append(0)
    stream = np.packbits(np.array(bits, dtype=np.uint8)).tobytes()
    want = checker.lossless_decode(5, stream, api.YQUANT, 0)
    got = mjpeg423_b200.lossless_decode(5, stream, None, api.YQUANT, 0)
    assert np.array_equal(got, want)


# ---- whole streams ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("W,H,n,amp,flat", [(64, 48, 3, 16, 0), (640, 480, 4, 16, 0), (640, 480, 2, 256, 0),
                                            (640, 480, 2, 16, 480), (640, 480, 2, 16, 200), (8, 8, 5, 64, 0),
                                            (1920, 1080, 2, 16, 0)])
def test_stream_bit_exact(checker, dec, W, H, n, amp, flat):
    mpg = synth.synth_mpg(W, H, n, 0, amp, flat)
    want = checker.decode_mpg(mpg)
    got = dec.decode_frames(mpg)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    st = dec.stats()
    assert st["frames"] == n and st["kernel_launches"] >= 4


def test_stream_4k_dense_all_ones_quant(checker, dec):
    """BASELINE configs[3] extreme: 3840x2160, full-range noise, all-ones quantisation tables passed through the
    `quant` parameter on both sides (dense coefficients, amplitudes up to 11 bits, ~25 bit/px)."""
    ones = np.ones(64, np.int16)
    mpg = synth.synth_mpg(3840, 2160, 2, 0, 256, 0, ones, ones)
    assert mpg.size / 2 > 20e6                               # really is the dense regime
    want = checker.decode_mpg(mpg, yq=ones, cq=ones, nthreads=2)
    dec.set_quant(ones, ones)
    try:
        got = dec.decode_frames(mpg)
        assert np.array_equal(got, want)
        st = dec.stats()
        assert st["list_entries"] > 0.9 * 2 * 3 * (3840 * 2160 // 64) * 63   # nearly every coefficient is coded
    finally:
        dec.set_quant(None, None)


def test_stream_640x480_300_frames_bit_exact(checker, dec):
    """BASELINE configs[1]: the 640x480 x 300 stream, every frame compared byte for byte."""
    mpg = synth.synth_mpg(640, 480, 300, 0, 16, 0)
    want = checker.decode_mpg(mpg, nthreads=os.cpu_count() or 1)
    out = dec.pinned(300 * 640 * 480 * 4)
    got = dec.decode_frames(mpg, out=out)
    assert np.array_equal(got, want)
    # frame sub-ranges and the device-resident path give the same bytes
    assert np.array_equal(dec.decode_frames(mpg, first=17, n=5), want[17:22])
    dec.upload(mpg, 100, 64)
    d_out = dec.device_alloc(64 * 640 * 480 * 4)
    dec.decode_resident(d_out)
    res = dec.to_host(d_out, 64 * 640 * 480 * 4).reshape(64, 480, 640, 4)
    assert np.array_equal(res, want[100:164])
    h = dec.hash_frames(d_out, 640 * 480 * 4, 64)
    assert np.array_equal(h, api.frame_hash_host(want[100:164]))
    dec.device_free(d_out)
    out.free()


def test_chunked_pipeline_matches_single_chunk(checker, dec):
    mpg = synth.synth_mpg(320, 240, 23, 5, 32, 0)
    want = checker.decode_mpg(mpg)
    for k in (1, 4, 23, 100):
        dec.set_option(api.OPT_CHUNK_FRAMES, k)
        assert np.array_equal(dec.decode_frames(mpg), want), k
        dec.upload(mpg)
        d_out = dec.device_alloc(want.nbytes)
        dec.decode_resident(d_out)
        assert np.array_equal(dec.to_host(d_out, want.nbytes).reshape(want.shape), want), k
        dec.device_free(d_out)
    dec.set_option(api.OPT_CHUNK_FRAMES, 0)


@pytest.mark.parametrize("gop,chunk", [(2, 0), (5, 3), (24, 7), (24, 1)])
def test_pframe_streams(checker, dec, gop, chunk):
    """P frames: coefficient-domain accumulation over the GOP (LIB/decoder/lossless_decode.c:90-92,121-123),
    with chunk boundaries forced through the middle of GOPs (the pipeline must not split them)."""
    W, H, n = 160, 96, 30
    fr = np.stack([synth.synth_frame(W, H, i, 24) for i in range(n)])
    mpg = synth.encode_mpg(fr, gop=gop)
    want = checker.decode_mpg(mpg)
    dec.set_option(api.OPT_CHUNK_FRAMES, chunk)
    try:
        assert np.array_equal(dec.decode_frames(mpg), want)
        dec.upload(mpg)
        d_out = dec.device_alloc(want.nbytes)
        dec.decode_resident(d_out)
        assert np.array_equal(dec.to_host(d_out, want.nbytes).reshape(want.shape), want)
        dec.device_free(d_out)
    finally:
        dec.set_option(api.OPT_CHUNK_FRAMES, 0)


@pytest.mark.parametrize("staged,chunk", [(0, 0), (0, 5), (1, 0)])
def test_pframe_static_picture(checker, dec, staged, chunk):
    """A static picture with a small moving patch: its P frames are long runs of unchanged blocks (twelve zero bits
    each), which never self-synchronise -- the chain kernel finds their block phase (zero-run shortcut) -- and most
    planes carry coefficient state through the GOP (k_decode_fused<true>: parked slots, DC and masks in registers)."""
    W, H, n = 640, 480, 27
    rng = np.random.default_rng(5)
    fr = np.repeat(synth.synth_frame(W, H, 0, 24)[None], n, 0).copy()
    for f in range(n):
        x0 = (f * 24) % (W - 64)
        fr[f, 200:264, x0:x0 + 64, :3] = rng.integers(0, 256, size=(64, 64, 3), dtype=np.uint8)
    mpg = synth.encode_mpg(fr, gop=12)
    assert mjpeg423_b200.probe(mpg).num_pframes >= 20
    want = checker.decode_mpg(mpg)
    dec.set_option(api.OPT_STAGED, staged)
    dec.set_option(api.OPT_CHUNK_FRAMES, chunk)
    try:
        assert np.array_equal(dec.decode_frames(mpg), want)
        dec.upload(mpg)
        d_out = dec.device_alloc(want.nbytes)
        dec.decode_resident(d_out)
        assert np.array_equal(dec.to_host(d_out, want.nbytes).reshape(want.shape), want)
        assert dec.stats()["fixups"] > 0          # the zero runs went through the chain kernel
        dec.device_free(d_out)
    finally:
        dec.set_option(api.OPT_STAGED, 0)
        dec.set_option(api.OPT_CHUNK_FRAMES, 0)


def test_stage_entry_points(checker, dec):
    """Per-stage device entry points reproduce the reference's intermediate buffers."""
    W, H, n = 320, 240, 3
    nb = (W // 8) * (H // 8)
    mpg = synth.synth_mpg(W, H, n, 0, 24, 0)
    dec.upload(mpg)
    d_coef = dec.device_alloc(n * 3 * nb * 128)
    d_samp = dec.device_alloc(n * 3 * nb * 64)
    d_out = dec.device_alloc(n * W * H * 4)
    dec.resident_entropy(d_coef)
    coef = dec.to_host(d_coef, n * 3 * nb * 128, np.int16).reshape(n, 3, nb, 8, 8)
    off = 20
    for f in range(n):
        fsz, typ, ys, cbs = (int(x) for x in mpg[off:off + 16].view("<u4"))
        pay = mpg[off + 16:off + fsz]
        for p, (a, b, q) in enumerate(((0, ys, api.YQUANT), (ys, ys + cbs, api.CQUANT), (ys + cbs, fsz - 16, api.CQUANT))):
            want = checker.lossless_decode(nb, pay[a:b].tobytes(), q, 0)
            assert np.array_equal(coef[f, p], want), (f, p)
        off += fsz
    dec.resident_idct(d_coef, d_samp)
    samp = dec.to_host(d_samp, n * 3 * nb * 64).reshape(n, 3, nb, 8, 8)
    assert np.array_equal(samp, checker.idct(coef.reshape(-1, 8, 8)).reshape(samp.shape))
    want = checker.decode_mpg(mpg)
    dec.resident_colour(d_samp, d_out)
    assert np.array_equal(dec.to_host(d_out, want.nbytes).reshape(want.shape), want)
    dec.to_device(d_out, np.zeros(want.nbytes, np.uint8))
    dec.resident_idct_colour(d_coef, d_out)
    assert np.array_equal(dec.to_host(d_out, want.nbytes).reshape(want.shape), want)
    for p in (d_coef, d_samp, d_out):
        dec.device_free(p)


def test_accelerator_seam(checker):
    """C0/idct_ycbcr_to_rgb_accel.h call sequence of C0/playback.c:53-121 on a 640x480 frame."""
    lib = api.load_library()
    W, H = 640, 480
    nb = (W // 8) * (H // 8)
    mpg = synth.synth_mpg(W, H, 1, 0, 16, 0)
    fsz, typ, ys, cbs = (int(x) for x in mpg[20:36].view("<u4"))
    pay = mpg[36:20 + fsz]
    Yc = checker.lossless_decode(nb, pay[:ys].tobytes(), api.YQUANT, 0)
    Cbc = checker.lossless_decode(nb, pay[ys:ys + cbs].tobytes(), api.CQUANT, 0)
    Crc = checker.lossless_decode(nb, pay[ys + cbs:].tobytes(), api.CQUANT, 0)
    assert lib.init_idct_ycbcr_to_rgb_accel() == 1
    out = np.zeros((H, W, 4), dtype=np.uint8)
    lib.idct_accel_calculate_buffer_cb(Cbc.ctypes.data, nb * 128)
    lib.idct_accel_calculate_buffer_cr(Crc.ctypes.data, nb * 128)
    lib.idct_accel_calculate_buffer_y(Yc.ctypes.data, nb * 128)
    lib.ycbcr_to_rgb_accel_get_results(out.ctypes.data, W * H * 4)
    lib.wait_for_idct_y_finsh()
    lib.wait_for_ycbcr_to_rgb_finsh()
    assert np.array_equal(out, checker.decode_mpg(mpg)[0])


def test_file_level_decode(tmp_path, checker):
    """mjpeg423_decode(): .mpg file in, name0000.bmp ... out (pixels compared; BMP rows are bottom-up)."""
    W, H, n = 64, 48, 3
    mpg = synth.synth_mpg(W, H, n, 0, 16, 0)
    src = tmp_path / "in.mpg"
    src.write_bytes(mpg.tobytes())
    mjpeg423_b200.mjpeg423_decode(str(src), str(tmp_path / "out0000.bmp"))
    want = checker.decode_mpg(mpg)
    for f in range(n):
        raw = (tmp_path / f"out{f:04d}.bmp").read_bytes()
        assert raw[:2] == b"BM" and len(raw) == 54 + W * H * 4
        px = np.frombuffer(raw[54:], np.uint8).reshape(H, W, 4)[::-1]
        assert np.array_equal(px, want[f])


def test_error_paths(dec):
    mpg = synth.synth_mpg(64, 48, 2, 0, 16, 0)
    with pytest.raises(RuntimeError):
        dec.decode_frames(mpg[:100])                    # truncated container
    with pytest.raises(RuntimeError):
        dec.decode_frames(mpg, first=1, n=5)             # range beyond num_frames
    bad = mpg.copy()
    fsz, typ, ys, cbs = (int(x) for x in bad[20:36].view("<u4"))
    bad[36 + ys // 2:36 + ys] = 0                        # Y stream now ends early: fewer than nb blocks
    bad[36 + ys // 2 - 1] = 0
    try:
        dec.decode_frames(bad)
    except RuntimeError as e:
        assert "blocks" in str(e)
    assert np.array_equal(dec.decode_frames(mpg, 0, 0).shape, (0, 48, 64, 4))   # empty range


def test_garbage_streams_are_survived(checker):
    """Non-conforming input (random bytes, all-ZRL, all-ones, truncated): the reference has undefined behaviour there
    (it indexes zigzag_table unchecked, LIB/decoder/lossless_decode.c:101-125), so nothing is compared with it --
    the library must neither fault nor hang, must be deterministic, and must decode a good stream afterwards."""
    rng = np.random.default_rng(99)
    cases = [rng.integers(0, 256, size=n, dtype=np.uint8).tobytes() for n in (1, 7, 600, 5000, 70000)]
    cases += [b"\xF0" * 9000, b"\xFF" * 9000, b"\x00" * 3, b"\x0F" * 4000]
    for nb in (1, 50, 3000):
        for raw in cases:
            a = mjpeg423_b200.lossless_decode(nb, raw, None, api.YQUANT, 0)
            b = mjpeg423_b200.lossless_decode(nb, raw, None, api.YQUANT, 0)
            assert a.shape == (nb, 8, 8) and np.array_equal(a, b)
    lv = _random_levels(rng, 200, 0.1, 30, 100)
    wire = lv.copy()
    wire[1:, 0] = lv[1:, 0] - lv[:-1, 0]
    stream = _encode_levels(wire)
    assert np.array_equal(mjpeg423_b200.lossless_decode(200, stream, None, api.YQUANT, 0),
                          checker.lossless_decode(200, stream, api.YQUANT, 0))


@pytest.mark.parametrize("staged", [0, 1])
def test_garbage_containers_are_survived(checker, dec, staged):
    """Well-formed containers whose plane streams are garbage (random bytes, zeros, all-ones; I and P frame types):
    the whole pipeline -- with P frames the GOP-walking fused kernel and the chain kernel's zero-run shortcut -- must
    neither fault nor hang, and must decode a good stream afterwards.  (Nothing is compared: the reference has
    undefined behaviour on such input.)"""
    W, H, n = 160, 96, 7
    rng = np.random.default_rng(17)
    fr = np.repeat(synth.synth_frame(W, H, 0, 24)[None], n, 0).copy()
    for f in range(n):
        fr[f, 8:40, 16 * f:16 * f + 32, :3] = rng.integers(0, 256, size=(32, 32, 3), dtype=np.uint8)
    good = synth.encode_mpg(fr, gop=3)
    assert mjpeg423_b200.probe(good).num_pframes >= 3
    want = checker.decode_mpg(good)
    dec.set_option(api.OPT_STAGED, staged)
    dec.set_option(api.OPT_VALIDATE, 0)
    try:
        for fill in ("random", "zeros", "ones"):
            bad = good.copy()
            off = 20
            for f in range(n):
                fsz = int(bad[off:off + 4].view("<u4")[0])
                body = bad[off + 16:off + fsz]
                body[:] = {"random": rng.integers(0, 256, size=body.size, dtype=np.uint8), "zeros": 0, "ones": 255}[fill]
                off += fsz
            try:
                out = dec.decode_frames(bad)
                assert out.shape == want.shape
            except RuntimeError:
                pass                                    # a reported stream error is fine as well
            assert np.array_equal(dec.decode_frames(good), want)
    finally:
        dec.set_option(api.OPT_STAGED, 0)
        dec.set_option(api.OPT_VALIDATE, 1)
