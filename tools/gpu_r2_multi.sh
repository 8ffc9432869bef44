#!/bin/bash
# multi-GPU call (gpurun --gpus N): bit-exactness of every sharded path + the bench line with its sharded workloads
N=${1:-2}
mkdir -p gpurun_out
T=r2m$N
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/${T}_multigpu_check.log 2>&1; echo "check rc=$?" | tee -a gpurun_out/${T}_multigpu_check.log
grep MULTIGPU gpurun_out/${T}_multigpu_check.log; tail -3 gpurun_out/${T}_multigpu_check.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
