"""One mid-frame between two image files -- the reference's `eval/interpolator_test.py:39-109` flags.

    python -m frame_interpolation_b200.interpolator_test --frame1 photos/one.png --frame2 photos/two.png \
        --model_path synthetic [--output_frame out.png] [--align 64] [--block_height 1 --block_width 1]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from . import eval_util
from .interpolator import Interpolator


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--frame1", required=True)
    p.add_argument("--frame2", required=True)
    p.add_argument("--model_path", default="synthetic")
    p.add_argument("--output_frame", default=None)
    p.add_argument("--align", type=int, default=64)
    p.add_argument("--block_height", type=int, default=1)
    p.add_argument("--block_width", type=int, default=1)
    a = p.parse_args(argv)
    interpolator = Interpolator(a.model_path, a.align, [a.block_height, a.block_width])
    x0 = eval_util.read_image(a.frame1)[np.newaxis]
    x1 = eval_util.read_image(a.frame2)[np.newaxis]
    mid = interpolator(x0, x1, np.full((1,), 0.5, np.float32))[0]
    out = a.output_frame or os.path.join(os.path.dirname(a.frame1), "output_frame.png")
    eval_util.write_image(out, mid)
    print(f"[film_b200] wrote {out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
