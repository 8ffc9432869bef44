"""Property tests (hypothesis) of the host-side logic that sits either side of the network call: alignment
padding (eval/interpolator.py:30-63), the tile split / stitch of the block path (eval/interpolator.py:66-126),
the recursion schedule (eval/util.py:85-118) and the multi-GPU partition helpers. No GPU, no oracle."""
import numpy as np
from hypothesis import given, settings, strategies as st

from frame_interpolation_b200 import eval_util, parallel, spec
from frame_interpolation_b200.interpolator import image_to_patches, patches_to_image


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 5000), st.integers(1, 5000), st.sampled_from([None, 1, 2, 16, 64, 128]))
def test_padded_shape_properties(h, w, align):
    ph, pw, oy, ox = spec.padded_shape(h, w, align)
    if not align:
        assert (ph, pw, oy, ox) == (h, w, 0, 0)
        return
    assert ph % align == 0 and pw % align == 0                     # aligned
    assert 0 <= ph - h < align and 0 <= pw - w < align             # minimal
    assert oy == (ph - h) // 2 and ox == (pw - w) // 2             # centred, floor offset (pad // 2 first)
    assert oy + h <= ph and ox + w <= pw                           # the crop window fits


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 4), st.integers(1, 4), st.integers(1, 6), st.integers(1, 6), st.integers(1, 3))
def test_patches_round_trip_and_row_major_tile_order(bh, bw, ph, pw, c):
    img = np.arange(bh * ph * bw * pw * c, dtype=np.float32).reshape(1, bh * ph, bw * pw, c)
    patches = image_to_patches(img, [bh, bw])
    assert patches.shape == (bh * bw, ph, pw, c)
    for t in range(bh * bw):                                       # tile t = row-major (block row, block col)
        r, q = divmod(t, bw)
        np.testing.assert_array_equal(patches[t], img[0, r * ph:(r + 1) * ph, q * pw:(q + 1) * pw])
    np.testing.assert_array_equal(patches_to_image(patches, [bh, bw]), img)


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 1000), st.integers(1, 16))
def test_block_partition_is_a_contiguous_balanced_cover(n, world):
    spans = [parallel.block_partition(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 1000), st.integers(1, 16))
def test_round_robin_is_a_disjoint_cover(n, world):
    seen = []
    for r in range(world):
        seen += list(parallel.round_robin(n, world, r))
    assert sorted(seen) == list(range(n))


@settings(max_examples=20, deadline=None)
@given(st.integers(0, 5))
def test_recursion_yields_2_pow_n_plus_1_frames_in_time_order(n):
    """With a linear stand-in, frame k of 2^n + 1 must be the k / 2^n blend: checks count and ordering."""
    a = np.zeros((4, 4, 3), np.float32)
    b = np.ones((4, 4, 3), np.float32)

    def mid(x0, x1, dt):
        return 0.5 * (x0 + x1)

    frames = list(eval_util.interpolate_recursively_from_memory([a, b], n, mid))
    assert len(frames) == 2 ** n + 1
    for k, f in enumerate(frames):
        np.testing.assert_allclose(f, k / 2 ** n, atol=1e-6)


@given(st.integers(min_value=1, max_value=1 << 19), st.integers(min_value=0, max_value=(1 << 21) - 1))
@settings(max_examples=400, deadline=None)
def test_multiply_shift_tile_decode_is_exact(d, x):
    """The device-side tile decode (csrc/film_tc_ptx.cuh `FastDiv`) replaces `x / d` by (x * ceil(2^40 / d)) >> 40 when
    x * d < 2^40 and falls back to the plain division otherwise; the multiply-shift form must be exact on its domain
    (tile indices / tile counts of every frame the engine accepts are far inside it)."""
    if x * d >= 1 << 40:
        return
    mul = ((1 << 40) + d - 1) // d
    q = (x * mul) >> 40
    assert q == x // d and x - q * d == x % d
