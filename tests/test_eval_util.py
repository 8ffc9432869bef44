"""Frame scheduling + image I/O + CLI plumbing (eval_util.py, interpolator_cli.py). CPU only
(a stand-in interpolator replaces the engine)."""
import os

import numpy as np
import pytest

from frame_interpolation_b200 import eval_util, interpolator_cli


def fake_interp(x0, x1, dt):
    assert x0.shape[0] == 1 and dt.shape == (1,)
    return 0.5 * (x0 + x1) + 0.125


def reference_order(f1, f2, n):
    """eval/util.py:62-91 semantics: depth-first, first frame included, last excluded."""
    if n == 0:
        return [f1]
    mid = fake_interp(f1[None], f2[None], np.full((1,), 0.5, np.float32))[0]
    return reference_order(f1, mid, n - 1) + reference_order(mid, f2, n - 1)


def test_recursive_schedule_matches_reference_order():
    rng = np.random.default_rng(0)
    frames = [rng.random((6, 8, 3), dtype=np.float32) for _ in range(3)]
    got = list(eval_util.interpolate_recursively_from_memory(frames, 3, fake_interp))
    want = reference_order(frames[0], frames[1], 3) + reference_order(frames[1], frames[2], 3) + [frames[2]]
    assert len(got) == 2 * 8 + 1
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    assert [f.shape for f in eval_util.interpolate_recursively_from_memory(frames, 0, fake_interp)] == [(6, 8, 3)] * 3


def test_write_read_round_trip_and_quantisation(tmp_path):
    img = np.zeros((4, 5, 3), np.float32)
    img[..., 0] = 1.0          # pure red: catches a BGR/RGB swap
    img[0, 0] = [0.5, 0.25, 2.0]   # 2.0 clips to 255
    img[0, 1] = [-1.0, 0.0019, 0.0020]  # -1 clips to 0 ; 0.0019*255+.5 = 0.98 -> 0 ; 0.0020*255+.5 = 1.01 -> 1
    p = str(tmp_path / "a.png")
    eval_util.write_image(p, img)
    back = eval_util.read_image(p)
    assert back.dtype == np.float32 and back.shape == (4, 5, 3)
    np.testing.assert_array_equal(eval_util.to_uint8(img)[0, 0], [128, 64, 255])
    np.testing.assert_array_equal(eval_util.to_uint8(img)[0, 1], [0, 0, 1])
    np.testing.assert_array_equal((back * 255 + 0.5).astype(np.uint8), eval_util.to_uint8(img))
    assert back[1, 1, 0] == 1.0 and back[1, 1, 2] == 0.0
    eval_util.write_image(str(tmp_path / "b.jpg"), img)
    assert eval_util.read_image(str(tmp_path / "b.jpg")).shape == (4, 5, 3)
    with pytest.raises(FileNotFoundError):
        eval_util.read_image(str(tmp_path / "missing.png"))


def test_natural_sort():
    assert eval_util.natural_sorted(["f10.png", "f2.png", "f1.png"]) == ["f1.png", "f2.png", "f10.png"]


def test_cli_flags_match_the_reference():
    with pytest.raises(SystemExit):                      # like the reference, --model_path is required:
        interpolator_cli.build_parser().parse_args(["--pattern", "x/*"])   # no silent random weights
    a = interpolator_cli.build_parser().parse_args(["--pattern", "x/*", "--model_path", "synthetic"])
    assert (a.times_to_interpolate, a.fps, a.align, a.block_height, a.block_width, a.output_video) == (5, 30, 64, 1, 1, False)
    a = interpolator_cli.build_parser().parse_args(
        ["--pattern", "x", "--model_path", "m", "--times_to_interpolate", "2", "--block_height", "2",
         "--block_width", "4", "--align", "32", "--output_video", "--fps", "60"])
    assert (a.model_path, a.times_to_interpolate, a.block_height, a.block_width, a.align, a.output_video, a.fps) == \
        ("m", 2, 2, 4, 32, True, 60)


def test_process_directory_with_stand_in(tmp_path):
    d = tmp_path / "clip"
    d.mkdir()
    rng = np.random.default_rng(1)
    for i in (1, 2, 10):
        eval_util.write_image(str(d / f"im{i}.png"), rng.random((8, 8, 3)).astype(np.float32))
    (d / "interpolated_frames").mkdir()
    (d / "interpolated_frames" / "frame_999.png").write_bytes(b"stale")
    n = interpolator_cli.process_directory(str(d), fake_interp, times=2, fps=30, video=False)
    assert n == 2 * 4 + 1
    out = sorted(os.listdir(d / "interpolated_frames"))
    assert out == [f"frame_{i:03d}.png" for i in range(9)]            # stale frame removed, natural input order
    first = eval_util.read_image(str(d / "interpolated_frames" / "frame_000.png"))
    np.testing.assert_array_equal(first, eval_util.read_image(str(d / "im1.png")))
