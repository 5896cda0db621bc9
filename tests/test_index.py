"""Container index / seek API and container hardening: host only, no GPU needed.

The trailers in tests/golden/golden_encoder.npz were written by the reference's own mjpeg423_encode()
(tests/golden/make_golden_encoder.py), so they pin the trailer layout the reader must accept."""
import os

import numpy as np
import pytest

import mjpeg423_b200
from mjpeg423_b200 import api, synth


@pytest.fixture(scope="module")
def enc(golden_dir):
    return np.load(os.path.join(golden_dir, "golden_encoder.npz"))


def _walk(mpg):
    """Independent header walk: [(frame_index, offset)] of the I frames."""
    a = np.frombuffer(bytes(mpg), dtype=np.uint8)
    nf = int(a[:4].view("<u4")[0])
    off, out = 20, []
    for f in range(nf):
        size, typ = (int(x) for x in a[off:off + 8].view("<u4"))
        if typ == 0:
            out.append((f, off))
        off += size
    return out


@pytest.mark.parametrize("key", ["enc_mpg_gop24", "enc_mpg_gop1", "enc_mpg_gop3"])
def test_index_matches_reference_trailer(enc, key):
    mpg = enc[key]
    info = mjpeg423_b200.probe(mpg)
    ix = mjpeg423_b200.IFrameIndex(mpg)
    assert ix.trailer_ok                                    # the real encoder's trailer agrees with the header walk
    assert len(ix) == info.num_iframes
    assert [tuple(int(v) for v in e) for e in ix.entries] == _walk(mpg)
    # the trailer bytes themselves: num_iframes x {frame_index, frame_position} at 20 + payload_size
    t = np.frombuffer(bytes(mpg), dtype=np.uint8)[20 + info.payload_size:20 + info.payload_size + 8 * len(ix)].view("<u4")
    assert np.array_equal(t.reshape(-1, 2), ix.entries)


def test_index_with_damaged_or_missing_trailer(enc):
    mpg = np.array(enc["enc_mpg_gop3"], copy=True)
    info = mjpeg423_b200.probe(mpg)
    good = mjpeg423_b200.IFrameIndex(mpg).entries
    mpg[20 + info.payload_size + 4] ^= 0x10                 # a wrong frame_position
    ix = mjpeg423_b200.IFrameIndex(mpg)
    assert not ix.trailer_ok and np.array_equal(ix.entries, good)     # the index is rebuilt from the walk
    cut = mpg[:20 + info.payload_size]                      # file without trailer
    ix = mjpeg423_b200.IFrameIndex(cut)
    assert not ix.trailer_ok and np.array_equal(ix.entries, good)


def test_seek_rules():
    fr = np.stack([synth.synth_frame(32, 16, i, 8) for i in range(300)])
    mpg = synth.encode_mpg(fr, gop=24)
    ix = mjpeg423_b200.IFrameIndex(mpg)
    idx = [int(e[0]) for e in ix.entries]
    assert idx == list(range(0, 300, 24)) and ix.trailer_ok
    # the I frame a decode of frame f starts from / the next one a player can jump to
    for f in (0, 1, 23, 24, 25, 299):
        assert idx[ix.seek(f)] == f // 24 * 24
        nxt = ix.seek(f, +1)
        assert (nxt == -1) if f > idx[-1] else idx[nxt] == -(-f // 24) * 24
    # C0/playback.c:157-194: first I frame at least 108 frames ahead; nothing with fewer than 120 frames left
    assert idx[ix.fast_forward(300, 0)] == 120
    assert idx[ix.fast_forward(300, 13)] == 144             # 13 + 108 = 121 -> next I frame
    assert idx[ix.fast_forward(300, 180)] == 288
    assert ix.fast_forward(300, 181) == -1
    # :196-227: last I frame at least 108 frames back; the start when less than 120 frames in
    assert idx[ix.rewind(119)] == 0
    assert idx[ix.rewind(120)] == 0
    assert idx[ix.rewind(250)] == 120                       # 250 - 108 = 142 -> I frame 120
    assert idx[ix.rewind(299)] == 168


def test_hostile_headers_do_not_crash():
    """num_frames = 0xFFFFFFFF asked for a 128 GB reserve and ended the process in std::terminate."""
    hdr = np.zeros(84, dtype=np.uint8)
    hdr[:20].view("<u4")[:] = [0xFFFFFFFF, 64, 48, 1, 64]
    with pytest.raises(RuntimeError, match="truncated|inconsistent"):
        mjpeg423_b200.probe(hdr)
    with pytest.raises(RuntimeError):
        mjpeg423_b200.IFrameIndex(hdr)
    # (W/8)*(H/8) used to wrap in 32 bits: 524288 x 524288 gave nb = 0
    hdr[:20].view("<u4")[:] = [1, 524288, 524288, 1, 64]
    with pytest.raises(RuntimeError, match="too large"):
        mjpeg423_b200.probe(hdr)
    hdr[:20].view("<u4")[:] = [1, 0xFFFFFFF8, 8, 1, 64]
    with pytest.raises(RuntimeError, match="too large"):
        mjpeg423_b200.probe(hdr)
