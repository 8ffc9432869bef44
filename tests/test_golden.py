"""Oracle vs the committed golden vectors (tests/golden/make_golden.py). CPU only.
The vectors were produced by the fp64 oracle in the build container; see the header of
oracle/film_oracle.py for why no reference-generated vectors exist (parity unpinned)."""
import ast
import glob
import os

import numpy as np
import pytest
import torch

from frame_interpolation_b200 import synthetic, weights
from oracle.film_oracle import OracleInterpolator

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(HERE, "*.npz")))


def load_case(path):
    z = np.load(path)
    c = dict(h=int(str(z["h"])), w=int(str(z["w"])), seed=int(str(z["seed"])), align=int(str(z["align"])),
             block=ast.literal_eval(str(z["block"])))
    return z, c


def test_golden_files_present():
    assert len(CASES) == 3


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_oracle_reproduces_golden(path):
    z, c = load_case(path)
    torch.set_num_threads(4)
    w = weights.synthetic_weights(1234)
    assert weights.digest(w) == str(z["weights_sha256"])
    x0, x1 = synthetic.frame_pair(c["h"], c["w"], seed=c["seed"], n_waves=6)
    assert abs(float(x0.sum()) - float(z["x0_sum"])) < 1e-2
    dt = np.full((1,), 0.5, np.float32)
    orc = OracleInterpolator(w, align=c["align"], block_shape=c["block"])
    out = orc(x0, x1, dt)
    assert out.shape == z["image"].shape == (1, c["h"], c["w"], 3)
    assert np.abs(out - z["image"]).max() < 1e-5
    if c["block"] is None:
        aux = {}
        orc.interpolate(x0, x1, dt, aux)
        fwd = aux["forward_flow_pyramid"][0][0].permute(1, 2, 0).numpy()
        assert np.abs(fwd - z["flow_fwd_l0"]).max() < 1e-4


def test_time_is_ignored():
    w = weights.synthetic_weights(1234)
    x0, x1 = synthetic.frame_pair(64, 64, seed=1, n_waves=4)
    orc = OracleInterpolator(w, align=64)
    a = orc(x0, x1, np.full((1,), 0.5, np.float32))
    b = orc(x0, x1, np.full((1,), 0.9, np.float32))
    np.testing.assert_array_equal(a, b)
