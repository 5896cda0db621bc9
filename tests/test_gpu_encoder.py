"""GPU parity of the encoder: the .mpg produced by mjpeg423_b200_encode_frames must be
byte-identical to the reference encoder's -- the compiled reference's frame loop where oracle/_ref is present (its
driver is pinned against the real mjpeg423_encode() through BMP files in tests/test_oracle.py), else the restatement."""
import numpy as np
import pytest

import mjpeg423_b200
from oracle import oracle

pytestmark = pytest.mark.gpu


def make_frames(n, H, W, amp, seed=0, drift=3, flat=False):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W]
    out = np.zeros((n, H, W, 4), np.uint8)
    for f in range(n):
        if flat:
            base = np.full((H, W, 3), (f * 37) & 255)
        else:
            base = np.stack([(255 * x // W + f * drift) & 255, (255 * y // H + 2 * f) & 255, (255 * (x + y) // (W + H)) & 255], -1)
        noise = rng.integers(0, amp, size=(H, W, 3)) if amp else 0
        out[f, ..., :3] = (base + noise) & 255
        out[f, ..., 3] = rng.integers(0, 256, size=(H, W))          # alpha is ignored by the encoder
    return out


@pytest.fixture(scope="module")
def dec():
    d = mjpeg423_b200.Decoder(0)
    yield d
    d.close()


@pytest.mark.parametrize("n,H,W,amp,max_i", [
    (1, 8, 8, 0, 1), (3, 16, 24, 16, 1), (6, 32, 48, 8, 24), (5, 24, 16, 256, 3), (8, 40, 64, 4, 4),
    (4, 480, 640, 16, 24), (3, 1080, 1920, 16, 2), (2, 64, 64, 0, 1),
])
def test_encode_bit_exact(dec, n, H, W, amp, max_i):
    chk = oracle.best()
    fr = make_frames(n, H, W, amp, seed=n * 7 + H)
    want = chk.encode_mpg(fr, max_i)
    got = dec.encode_frames(fr, max_i)
    assert got.size == want.size
    assert np.array_equal(got, want)


def test_encode_grey_ramp_truncation(dec):
    """R = G = B = v: 0.299 v + 0.587 v + 0.114 v is v only up to double rounding -- the uint8 truncation must follow
    the reference's IEEE arithmetic for all 256 greys (LIB/encoder/rgb_to_ycbcr.c:64)."""
    fr = np.zeros((1, 16, 16, 4), np.uint8)
    fr[0, ..., :3] = np.arange(256, dtype=np.uint8).reshape(16, 16, 1)
    assert np.array_equal(dec.encode_frames(fr, 1), oracle.best().encode_mpg(fr, 1))


def test_encode_all_colours_sample(dec):
    """A dense sample of the RGB cube (every 5th level per channel + random triples) through the colour stage."""
    rng = np.random.default_rng(5)
    v = np.arange(0, 256, 5, dtype=np.uint8)
    cube = np.stack(np.meshgrid(v, v, v, indexing="ij"), -1).reshape(-1, 3)
    rnd = rng.integers(0, 256, size=(64 * 64 * 40 - cube.shape[0] % (64 * 64), 3), dtype=np.uint8)
    px = np.concatenate([cube, rnd])[: (cube.shape[0] + rnd.shape[0]) // (64 * 64) * 64 * 64]
    fr = np.zeros((px.shape[0] // (64 * 64), 64, 64, 4), np.uint8)
    fr[..., :3] = px.reshape(-1, 64, 64, 3)
    assert np.array_equal(dec.encode_frames(fr, 1), oracle.best().encode_mpg(fr, 1))


def test_encode_pframes_and_chunks(dec):
    """P frames across pipeline chunks (the previous chunk's last frame is the P reference) and fix_tail."""
    fr = make_frames(12, 48, 64, 6, seed=3, drift=1)
    want = oracle.best().encode_mpg(fr, 5)
    for chunk in (0, 1, 5, 7):
        dec.set_option(mjpeg423_b200.api.OPT_CHUNK_FRAMES, chunk)
        got = dec.encode_frames(fr, 5)
        assert np.array_equal(got, want), f"chunk_frames={chunk}"
    dec.set_option(mjpeg423_b200.api.OPT_CHUNK_FRAMES, 0)
    assert mjpeg423_b200.probe(want).num_pframes > 0
    fixed = dec.encode_frames(fr, 5, fix_tail=True)
    assert np.array_equal(fixed, oracle.port().encode_mpg(fr, 5, fix_tail=True))


def test_encode_custom_quant_and_roundtrip(dec):
    """Custom tables go through set_quant; decode(encode(x)) with the fixed tail equals the oracle's decode of the
    same file, and the GPU decoder reads the GPU encoder's output."""
    fr = make_frames(4, 64, 96, 32, seed=9)
    yq = np.full(64, 7, np.int16)
    cq = np.full(64, 13, np.int16)
    chk = oracle.best()
    dec.set_quant(yq, cq)
    try:
        got = dec.encode_frames(fr, 3)
        assert np.array_equal(got, chk.encode_mpg(fr, 3, yq, cq))
        fixed = dec.encode_frames(fr, 3, fix_tail=True)
        assert np.array_equal(dec.decode_frames(fixed), chk.decode_mpg(fixed, yq=yq, cq=cq))
    finally:
        dec.set_quant(None, None)


def test_encode_from_device_frames(dec):
    fr = make_frames(3, 32, 32, 16, seed=11)
    d = dec.device_alloc(fr.nbytes)
    try:
        dec.to_device(d, fr)
        got = dec.encode_frames(None, 2, d_frames=d, shape=(3, 32, 32))
        assert np.array_equal(got, oracle.best().encode_mpg(fr, 2))
    finally:
        dec.device_free(d)


def test_encode_errors(dec):
    with pytest.raises(RuntimeError):
        dec.encode_frames(np.zeros((1, 12, 16, 4), np.uint8))        # H not a multiple of 8


def test_encode_matches_golden_reference_files(dec):
    """The committed .mpg files written by the reference's real mjpeg423_encode() (tests/golden/golden_encoder.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_encoder.npz"))
    for gop in (24, 1, 3):
        got = dec.encode_frames(g["enc_frames"], gop)
        assert np.array_equal(got[:-512], g[f"enc_mpg_gop{gop}"]), f"gop {gop}"
