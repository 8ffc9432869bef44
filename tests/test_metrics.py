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
