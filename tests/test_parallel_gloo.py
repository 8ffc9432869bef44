"""Host-side sharding logic (frame_interpolation_b200/parallel.py) with world_size 2 on
the gloo backend. The engine is replaced by a deterministic stand-in: what is tested is the
partitioning, the single all-gather and the reassembly order, not the network."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from frame_interpolation_b200 import parallel


def fake_engine(x0, x1, dt):
    # position-dependent so that any stitching / ordering mistake changes the result
    ramp = np.arange(x0.shape[1] * x0.shape[2], dtype=np.float32).reshape(1, x0.shape[1], x0.shape[2], 1)
    return 0.5 * (x0 + x1) + 1e-3 * ramp


def fake_engine_dev(x0, x1, out):
    """Tensor-view form of `fake_engine` (the engine_dev interface of the device-resident path)."""
    h, w, _ = x0.shape
    ramp = torch.arange(h * w, dtype=torch.float32).view(h, w, 1)
    out.copy_(0.5 * (x0 + x1) + 1e-3 * ramp)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    res = {}
    a = rng.random((5, 8, 12, 3), dtype=np.float32)
    b = rng.random((5, 8, 12, 3), dtype=np.float32)
    res["pairs"] = parallel.interpolate_pairs(fake_engine, a, b)
    res["pairs_local"] = parallel.interpolate_pairs(fake_engine, a, b, gather=False)
    big0 = rng.random((1, 12, 18, 3), dtype=np.float32)
    big1 = rng.random((1, 12, 18, 3), dtype=np.float32)
    res["tiled"] = parallel.interpolate_tiled(fake_engine, big0, big1, [3, 3])
    res["rec"] = np.stack(parallel.interpolate_recursively(fake_engine, a[0], b[0], 3))
    # device-resident path (tensor views, in-place all-gather of the slot buffer), here on CPU tensors
    ta, tb, t0, t1 = (torch.from_numpy(x) for x in (a, b, big0, big1))
    res["pairs_dev"] = parallel.interpolate_pairs_device(fake_engine_dev, ta, tb).numpy()
    res["tiled_dev"] = parallel.interpolate_tiled_device(fake_engine_dev, t0, t1, [3, 3]).numpy()
    res["tiled_dev_4x4"] = parallel.interpolate_tiled_device(fake_engine_dev, ta[:1], tb[:1], [4, 4]).numpy()
    res["rec_dev"] = parallel.interpolate_recursively_device(fake_engine_dev, ta[0], tb[0], 3).numpy()
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def _serial_reference():
    rng = np.random.default_rng(0)
    a = rng.random((5, 8, 12, 3), dtype=np.float32)
    b = rng.random((5, 8, 12, 3), dtype=np.float32)
    big0 = rng.random((1, 12, 18, 3), dtype=np.float32)
    big1 = rng.random((1, 12, 18, 3), dtype=np.float32)
    dt = np.full((1,), 0.5, np.float32)
    pairs = np.concatenate([fake_engine(a[i:i + 1], b[i:i + 1], dt) for i in range(5)])
    from frame_interpolation_b200.interpolator import image_to_patches, patches_to_image
    p0, p1 = image_to_patches(big0, [3, 3]), image_to_patches(big1, [3, 3])
    tiled = patches_to_image(np.concatenate([fake_engine(p0[t][None], p1[t][None], dt) for t in range(9)]), [3, 3])

    def rec(f1, f2, n):        # eval/util.py:62-91 order
        if n == 0:
            return [f1]
        mid = fake_engine(f1[None], f2[None], dt)[0]
        return rec(f1, mid, n - 1) + rec(mid, f2, n - 1)
    seq = np.stack(rec(a[0], b[0], 3) + [b[0]])
    return pairs, tiled, seq


def test_partition_helpers():
    assert [parallel.block_partition(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [parallel.block_partition(2, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert parallel.round_robin(9, 4, 1) == [1, 5]
    covered = sorted(t for r in range(8) for t in parallel.round_robin(16, 8, r))
    assert covered == list(range(16))


def test_single_process_paths_match_serial():
    pairs, tiled, seq = _serial_reference()
    rng = np.random.default_rng(0)
    a = rng.random((5, 8, 12, 3), dtype=np.float32)
    b = rng.random((5, 8, 12, 3), dtype=np.float32)
    big0 = rng.random((1, 12, 18, 3), dtype=np.float32)
    big1 = rng.random((1, 12, 18, 3), dtype=np.float32)
    np.testing.assert_array_equal(parallel.interpolate_pairs(fake_engine, a, b), pairs)
    np.testing.assert_array_equal(parallel.interpolate_tiled(fake_engine, big0, big1, [3, 3]), tiled)
    np.testing.assert_array_equal(np.stack(parallel.interpolate_recursively(fake_engine, a[0], b[0], 3)), seq)
    assert seq.shape[0] == 2 ** 3 + 1
    ta, tb, t0, t1 = (torch.from_numpy(x) for x in (a, b, big0, big1))
    np.testing.assert_array_equal(parallel.interpolate_pairs_device(fake_engine_dev, ta, tb).numpy(), pairs)
    np.testing.assert_array_equal(parallel.interpolate_tiled_device(fake_engine_dev, t0, t1, [3, 3]).numpy(), tiled)
    np.testing.assert_array_equal(parallel.interpolate_recursively_device(fake_engine_dev, ta[0], tb[0], 3).numpy(), seq)
    # strided tile views: a tile is a row-pitched window of the frame, never a copy
    v = parallel.tile_view(t0, [3, 3], 4)
    assert v.shape == (4, 6, 3) and v.stride(0) == 18 * 3 and v.data_ptr() == t0[0, 4, 6].data_ptr()


@pytest.mark.timeout(240)
@pytest.mark.parametrize("world", [2, 3])
def test_world_size_n_gloo_bitwise_equals_serial(world):
    """world 3 exercises the uneven shares: 5 pairs -> 2/2/1, 9 tiles -> 3/3/3, 16 tiles -> 6/5/5, recursion levels of 1/2/4
    calls over 3 ranks (ranks with nothing to do still take part in the in-place all-gather)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pairs, tiled, seq = _serial_reference()
    rng = np.random.default_rng(0)
    a = rng.random((5, 8, 12, 3), dtype=np.float32)
    b = rng.random((5, 8, 12, 3), dtype=np.float32)
    tiled44 = parallel.interpolate_tiled(fake_engine, a[:1], b[:1], [4, 4])    # world 1: the serial tiled path
    for r in range(world):
        np.testing.assert_array_equal(got[r]["pairs"], pairs)
        np.testing.assert_array_equal(got[r]["tiled"], tiled)
        np.testing.assert_array_equal(got[r]["rec"], seq)
        np.testing.assert_array_equal(got[r]["pairs_dev"], pairs)
        np.testing.assert_array_equal(got[r]["tiled_dev"], tiled)
        np.testing.assert_array_equal(got[r]["rec_dev"], seq)
        np.testing.assert_array_equal(got[r]["tiled_dev_4x4"], tiled44)
    # local shares are the contiguous block partition
    for r in range(world):
        lo, hi = parallel.block_partition(5, world, r)
        np.testing.assert_array_equal(got[r]["pairs_local"], pairs[lo:hi])
