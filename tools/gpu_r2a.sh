#!/bin/bash
# round 2, first GPU call: sanity matrix of the new single-pass kernels, the parity suite, the precision study, a bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt
timeout 600 python tools/gpu_sanity.py > gpurun_out/r2a_sanity.log 2>&1; echo "sanity rc=$?" | tee -a gpurun_out/r2a_sanity.log
tail -30 gpurun_out/r2a_sanity.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
timeout 900 python tools/precision_study.py --height 1080 --width 1920 --seeds 0 --out gpurun_out/r2a_precision_1080p > gpurun_out/r2a_precision.log 2>&1; echo "study rc=$?"
tail -5 gpurun_out/r2a_precision.log
timeout 600 python bench.py --steps 20 --warmup 3 --op-table gpurun_out/r2a_ops.csv > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
cat gpurun_out/r2a_bench.json | head -c 3000
