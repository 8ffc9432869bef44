"""Run under torchrun on N GPUs: sharded paths (frame pairs, tiles, recursion) with the REAL engine,
NCCL all-gather reassembly, compared bit-for-bit with the single-GPU result computed on rank 0's GPU."""
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
dt = np.full((1,), 0.5, np.float32)

# frame pairs
pairs = [synthetic.frame_pair(180, 320, seed=s, n_waves=6) for s in range(5)]
x0 = np.concatenate([a for a, _ in pairs]); x1 = np.concatenate([b for _, b in pairs])
got = parallel.interpolate_pairs(eng, x0, x1, device=dev)
ref = np.concatenate([eng(x0[i:i+1], x1[i:i+1], dt) for i in range(5)])
ok1 = np.array_equal(got, ref)
# tiles (4K-like 2x2 at reduced size) + 4x4
b0, b1 = synthetic.frame_pair(360, 640, seed=11, n_waves=6)
ok2 = True
for bs in ([2, 2], [4, 4]):
    t = parallel.interpolate_tiled(eng, b0, b1, bs, device=dev)
    single = Interpolator("synthetic", align=64, block_shape=bs, device=local)
    ok2 = ok2 and np.array_equal(t, single(b0, b1, dt))
    single.close()
# recursion
f = parallel.interpolate_recursively(eng, pairs[0][0][0], pairs[0][1][0], 3, device=dev)
def rec(a, b, n):
    if n == 0: return [a]
    m = eng(a[None], b[None], dt)[0]
    return rec(a, m, n - 1) + rec(m, b, n - 1)
serial = rec(pairs[0][0][0], pairs[0][1][0], 3) + [pairs[0][1][0]]
ok3 = len(f) == 9 and all(np.array_equal(u, v) for u, v in zip(f, serial))
res = torch.tensor([int(ok1), int(ok2), int(ok3)], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU world", world, "pairs", bool(res[0]), "tiled", bool(res[1]), "recursive", bool(res[2]), flush=True)
dist.destroy_process_group()
sys.exit(0 if int(res.min()) == 1 else 1)
