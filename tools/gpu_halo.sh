#!/bin/bash
# Validation of the wide-halo conv mode: full GPU suite with the default (halo on), then the bench with halo on / pair-only / off.
mkdir -p gpurun_out
TAG=${1:-r1v}
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py --steps 20 --warmup 3 --op-table gpurun_out/ops_$TAG.csv > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 2500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
FILM_HALO=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --op-table gpurun_out/ops_${TAG}_halo0.csv > gpurun_out/bench_${TAG}_halo0.json 2> gpurun_out/bench_${TAG}_halo0.err
head -c 400 gpurun_out/bench_${TAG}_halo0.json; echo; tail -3 gpurun_out/bench_${TAG}_halo0.err
FILM_HALO=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --op-table gpurun_out/ops_${TAG}_halo1.csv > gpurun_out/bench_${TAG}_halo1.json 2> gpurun_out/bench_${TAG}_halo1.err
head -c 400 gpurun_out/bench_${TAG}_halo1.json; echo; tail -3 gpurun_out/bench_${TAG}_halo1.err
