"""Display side: the frame ring of C0/libs/ece423_vid_ctl and the BMP dump -- host logic,
no GPU needed; the playback loop itself is in the gpu-marked tests below."""
import ctypes as C
import os

import numpy as np
import pytest

import mjpeg423_b200
from oracle import oracle


def test_ring_state_machine():
    """The reference's sequence (ece423_vid_ctl.c:80,125-224): written starts one ahead of displayed; the producer may
    fill num_buffers-1 frames before the ring is full; switch_frames fails when nothing new was registered."""
    d = mjpeg423_b200.Display(16, 8, 4)
    try:
        assert d.num_buffers == 4
        assert d.switch_frames() == -1                       # nothing written yet: stay on the black frame
        assert not d.get_displayed_buffer().any()            # init clears every buffer (:104-107)
        filled = 0
        while d.buffer_is_available() == 0:
            d.get_buffer()[...] = filled + 1
            d.register_written_buffer()
            filled += 1
            assert filled <= 4
        assert filled == 3                                   # written has caught up with displayed
        for k in range(3):
            assert d.switch_frames() == 0
            assert int(d.get_displayed_buffer()[0, 0, 0]) == k + 1     # display order = write order
        assert d.switch_frames() == -1
        assert d.buffer_is_available() == 0
        d.clear_screen(0x7F)
        assert (d.get_buffer() == 0x7F).all()
    finally:
        d.close()


@pytest.mark.parametrize("asked,got", [(1, 2), (2, 2), (3, 2), (4, 4), (7, 4), (8, 8), (25, 16), (99, 16)])
def test_ring_buffer_count(asked, got):
    """Clamped to [2, 25] like ece423_video_display_init (:55-61) and rounded down to a power of two (the indices are
    masked; COMMON/config.h:27 'MUST BE A POWER OF 2')."""
    d = mjpeg423_b200.Display(8, 8, asked)
    try:
        assert d.num_buffers == got
    finally:
        d.close()


def test_bmp_matches_reference_encode_bmp(tmp_path):
    r = oracle.ref()
    rng = np.random.default_rng(4)
    fr = rng.integers(0, 256, size=(24, 40, 4), dtype=np.uint8)
    mine = tmp_path / "mine.bmp"
    mjpeg423_b200.write_bmp(str(mine), fr)
    raw = mine.read_bytes()
    assert len(raw) == 54 + fr.nbytes
    px = np.frombuffer(raw[54:], np.uint8).reshape(24, 40, 4)
    assert np.array_equal(px[::-1], fr)                     # bottom-up, B G R A as stored
    if r is None:
        pytest.skip("oracle/_ref not built: header compared only where the reference's libbmp is available")
    r.lib.encode_bmp.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_char_p]
    ref = tmp_path / "ref.bmp"
    r.lib.encode_bmp(fr.ctypes.data, 40, 24, os.fsencode(str(ref)))
    assert raw == ref.read_bytes()


@pytest.mark.gpu
def test_play_no_timer_shows_every_frame():
    """noTimer mode (C0/playback.c:130-133): flip after every frame; the displayed frames are the decoded frames."""
    from mjpeg423_b200 import synth
    mpg = synth.synth_mpg(64, 48, 9, 0, 16, 0)
    want = oracle.best().decode_mpg(mpg)
    dec = mjpeg423_b200.Decoder(0)
    disp = mjpeg423_b200.Display(64, 48, 4)
    try:
        seen = []
        shown, dropped = mjpeg423_b200.play(dec, mpg, disp, on_display=lambda i, f: seen.append((i, f.copy())))
        assert (shown, dropped) == (9, 0)
        assert [i for i, _ in seen] == list(range(9))
        assert all(np.array_equal(f, want[i]) for i, f in seen)
        seen.clear()
        shown, _ = mjpeg423_b200.play(dec, mpg, disp, first=3, n=4, on_display=lambda i, f: seen.append((i, f.copy())))
        assert shown == 4 and [i for i, _ in seen] == [3, 4, 5, 6]
        assert all(np.array_equal(f, want[i]) for i, f in seen)
    finally:
        disp.close()
        dec.close()


@pytest.mark.gpu
def test_play_paced_with_pframes():
    """Timer mode with a short period, a stream with P frames (batches must start on I frames) and a 2-deep ring."""
    import time
    dec = mjpeg423_b200.Decoder(0)
    rng = np.random.default_rng(3)
    y, x = np.mgrid[0:32, 0:48]
    fr = np.zeros((150, 32, 48, 4), np.uint8)
    for f in range(150):
        base = np.stack([(5 * x + f) & 255, (7 * y + 2 * f) & 255, (3 * (x + y)) & 255], -1)
        fr[f, ..., :3] = (base + rng.integers(0, 6, size=(32, 48, 3))) & 255
    mpg = dec.encode_frames(fr, 24, fix_tail=True)
    assert mjpeg423_b200.probe(mpg).num_pframes > 0
    want = oracle.best().decode_mpg(mpg)
    disp = mjpeg423_b200.Display(48, 32, 2)
    try:
        seen = []
        t0 = time.perf_counter()
        shown, dropped = mjpeg423_b200.play(dec, mpg, disp, frame_period_us=2000, on_display=lambda i, f: seen.append((i, f.copy())))
        dt = time.perf_counter() - t0
        assert shown == 150 and [i for i, _ in seen] == list(range(150))
        assert all(np.array_equal(f, want[i]) for i, f in seen)
        assert dt >= 150 * 0.002 * 0.95                     # paced: never faster than the timer
    finally:
        disp.close()
        dec.close()
