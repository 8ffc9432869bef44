"""Run under torchrun on N GPUs: every sharded path (frame pairs, tiles, recursion; host-staged AND device-resident)
with the REAL engine and NCCL, compared bit for bit with the single-GPU result computed by the same rank alone.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py            # prints one MULTIGPU line; exit code 0 iff every check is bit-exact
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from frame_interpolation_b200 import parallel, synthetic
from frame_interpolation_b200.interpolator import Interpolator

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
eng = Interpolator("synthetic", align=64, device=local)
edev = parallel.device_engine(eng)
dt = np.full((1,), 0.5, np.float32)
checks = {}

# ---- host-staged helpers (numpy in / numpy out) ------------------------------------------
pairs = [synthetic.frame_pair(180, 320, seed=s, n_waves=6) for s in range(5)]
x0 = np.concatenate([a for a, _ in pairs]); x1 = np.concatenate([b for _, b in pairs])
ref_pairs = np.concatenate([eng(x0[i:i+1], x1[i:i+1], dt) for i in range(5)])
checks["pairs_host"] = np.array_equal(parallel.interpolate_pairs(eng, x0, x1, device=dev), ref_pairs)
b0, b1 = synthetic.frame_pair(360, 640, seed=11, n_waves=6)
ref_tiled = {}
ok = True
for bs in ([2, 2], [4, 4]):
    single = Interpolator("synthetic", align=64, block_shape=bs, device=local)
    ref_tiled[tuple(bs)] = single(b0, b1, dt).copy()
    single.close()
    ok = ok and np.array_equal(parallel.interpolate_tiled(eng, b0, b1, bs, device=dev), ref_tiled[tuple(bs)])
checks["tiled_host"] = ok
def rec(a, b, n):
    if n == 0: return [a]
    m = eng(a[None], b[None], dt)[0]
    return rec(a, m, n - 1) + rec(m, b, n - 1)
serial = np.stack(rec(pairs[0][0][0], pairs[0][1][0], 3) + [pairs[0][1][0]])
f = parallel.interpolate_recursively(eng, pairs[0][0][0], pairs[0][1][0], 3, device=dev)
checks["recursive_host"] = len(f) == 9 and all(np.array_equal(u, v) for u, v in zip(f, serial))

# ---- device-resident path: tensor views in, NCCL in-place all-gather, nothing through host ---
t0, t1 = torch.from_numpy(x0).to(dev), torch.from_numpy(x1).to(dev)
got = parallel.interpolate_pairs_device(edev, t0, t1)
torch.cuda.synchronize()
checks["pairs_device"] = np.array_equal(got.cpu().numpy(), ref_pairs)
tb0, tb1 = torch.from_numpy(b0).to(dev), torch.from_numpy(b1).to(dev)
ok = True
for bs in ([2, 2], [4, 4]):
    out = parallel.interpolate_tiled_device(edev, tb0, tb1, bs)
    torch.cuda.synchronize()
    ok = ok and np.array_equal(out.cpu().numpy(), ref_tiled[tuple(bs)])
checks["tiled_device"] = ok
seq = parallel.interpolate_recursively_device(edev, t0[0], t1[0], 3)
torch.cuda.synchronize()
checks["recursive_device"] = np.array_equal(seq.cpu().numpy(), serial)

names = sorted(checks)
res = torch.tensor([int(checks[n]) for n in names], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU world", world, " ".join(f"{n}={bool(v)}" for n, v in zip(names, res.tolist())), flush=True)
dist.destroy_process_group()
sys.exit(0 if int(res.min()) == 1 else 1)
