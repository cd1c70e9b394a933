"""K13 parity: the changed-frame prefilter on the GPU (slideo_b200_mark_changed_bgr8) == the oracle MarkSimilarIter
(oracle/area_oracle.c, pinned against cv2.resize(INTER_AREA) / cv2.norm): identical similarities (bitwise) and flags."""
import numpy as np
import pytest

import oracle
import synth

pytestmark = pytest.mark.gpu


def _sequence():
    pages = [synth.make_page(p) for p in range(3)]
    base = [synth.make_frame(f, 3, pages) for f in range(5)]
    rng = np.random.default_rng(8)
    seq = []
    for i, f in enumerate(base):
        seq.append(f)
        g = f.copy()                                   # a near-identical repeat: must be "unchanged"
        ys, xs = rng.integers(0, 1080, 200), rng.integers(0, 1920, 200)
        g[ys, xs] = 255 - g[ys, xs]
        seq.append(g)
        if i == 2:
            seq.append(g.copy())                       # exact repeat: similarity 1.0
    return np.stack(seq)


def test_mark_changed_equals_oracle(ctx):
    frames = _sequence()
    want_c, want_s = oracle.mark_similar(frames)
    got_c, got_s = ctx.mark_changed_bgr8(frames, reset=True)
    assert np.array_equal(got_s.view(np.uint32), want_s.view(np.uint32))
    assert np.array_equal(got_c, want_c)
    assert got_c[0] and not got_c[1] and got_s[0] == 0.0 and (got_s == 1.0).any()
    # the chain continues across calls, in any batching, and the device entry point agrees
    a_c, a_s = ctx.mark_changed_bgr8(frames[:4], reset=True)
    b_c, b_s = ctx.mark_changed_bgr8(frames[4:], reset=False)
    assert np.array_equal(np.concatenate([a_s, b_s]).view(np.uint32), want_s.view(np.uint32))
    assert np.array_equal(np.concatenate([a_c, b_c]), want_c)
    import torch
    d = torch.from_numpy(frames).cuda()
    d_c, d_s = ctx.mark_changed_bgr8_device(d.data_ptr(), len(frames), 1920, 1080, reset=True)
    assert np.array_equal(d_s.view(np.uint32), want_s.view(np.uint32)) and np.array_equal(d_c, want_c)


def test_mark_changed_other_geometry_and_small_batches():
    import slideo_b200
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 256, (7, 480, 854, 3), dtype=np.uint8)
    frames[3] = frames[2]
    want_c, want_s = oracle.mark_similar(frames)
    with slideo_b200.Context(slideo_b200.default_config(max_batch=2)) as c:
        got_c, got_s = c.mark_changed_bgr8(frames, reset=True)
    assert np.array_equal(got_s.view(np.uint32), want_s.view(np.uint32)) and np.array_equal(got_c, want_c)
    assert not got_c[3] and got_c[4]
