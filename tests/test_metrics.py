"""PSNR / SSIM definitions (metrics.py). CPU only."""
import numpy as np
import pytest

from frame_interpolation_b200 import metrics


def test_psnr_closed_form():
    a = np.zeros((1, 8, 8, 3), np.float32)
    b = np.full((1, 8, 8, 3), 0.1, np.float32)
    assert abs(metrics.psnr(a, b) - 20.0) < 1e-4          # mse = 0.01 -> 10*log10(1/0.01)
    assert metrics.psnr(a, a) == float("inf")


def test_ssim_properties():
    rng = np.random.default_rng(0)
    a = rng.random((2, 32, 40, 3))
    assert abs(metrics.ssim(a, a) - 1.0) < 1e-12
    noisy = np.clip(a + 0.1 * rng.standard_normal(a.shape), 0, 1)
    s = metrics.ssim(a, noisy)
    assert 0.0 < s < 1.0 and abs(s - metrics.ssim(noisy, a)) < 1e-12
    const = np.full((1, 16, 16, 1), 0.5)
    # two constant images: contrast/structure term = 1, luminance term = (2ab + c1) / (a^2 + b^2 + c1)
    want = (2 * 0.5 * 0.25 + 1e-4) / (0.25 + 0.0625 + 1e-4)
    assert abs(metrics.ssim(const, const * 0.5) - want) < 1e-12
    with pytest.raises(ValueError):
        metrics.ssim(np.zeros((1, 8, 8, 3)), np.zeros((1, 8, 8, 3)))
    g = metrics._gauss_window()
    assert abs(g.sum() - 1) < 1e-15 and g.argmax() == 5


def test_benchmark_evaluator_on_triplet_folders(tmp_path):
    """eval_cli (TF-free counterpart of eval/eval_cli.py:88-178): triplet discovery, per-example rows, mean row, clipping."""
    import os
    import numpy as np
    from frame_interpolation_b200 import eval_cli, eval_util, metrics
    rng = np.random.default_rng(3)
    keys = []
    for name in ("00001/0001", "00001/0002", "extra"):
        d = tmp_path / "data" / name
        d.mkdir(parents=True)
        base = rng.random((24, 32, 3)).astype(np.float32)
        for i, f in enumerate(("im1.png", "im2.png", "im3.png")):
            eval_util.write_image(str(d / f), np.clip(base + 0.05 * i, 0, 1))
        keys.append(name.replace("/", "_"))
    (tmp_path / "data" / "not_a_triplet").mkdir()
    eval_util.write_image(str(tmp_path / "data" / "not_a_triplet" / "a.png"), rng.random((8, 8, 3)).astype(np.float32))
    trip = eval_cli.find_triplets(str(tmp_path / "data"))
    assert [k for k, _ in trip] == sorted(keys)

    def mean_interp(x0, x1, dt):                      # stand-in: the average frame, pushed out of range to test the clip
        return 0.5 * (x0 + x1) + 2.0 * (x0 > 0.99)

    totals = eval_cli.run_evaluation(mean_interp, trip, str(tmp_path / "out"), metrics=["l1", "l2", "ssim", "psnr"],
                                     output_frames=True)
    rows = [r.strip().split(", ") for r in open(tmp_path / "out" / "results.csv")]
    assert rows[0] == ["key", "l1", "l2", "ssim", "psnr"] and rows[-1][0] == "mean" and len(rows) == 5
    y = eval_util.read_image(trip[0][1][1])
    pred = np.clip(mean_interp(eval_util.read_image(trip[0][1][0])[None], eval_util.read_image(trip[0][1][2])[None], None)[0], 0, 1)
    assert abs(float(rows[1][4]) - metrics.psnr(pred[None], y[None])) < 1e-9
    assert abs(float(rows[-1][1]) - np.mean([float(r[1]) for r in rows[1:-1]])) < 1e-12 and abs(totals["l1"] - float(rows[-1][1])) < 1e-12
    assert os.path.exists(tmp_path / "out" / f"{trip[0][0]}_image.png") and os.path.exists(tmp_path / "out" / "readme.txt")
    assert eval_cli.run_evaluation(mean_interp, trip, str(tmp_path / "o2"), max_examples=1)["psnr"] == float(rows[1][4])
