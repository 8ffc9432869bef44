#!/bin/bash
# Final check of a build: quick oracle comparison of the default and the all-off configuration, the GPU parity suite,
# smoke(), the bench line (both arms), the per-launch ncu counters of one call.
mkdir -p gpurun_out
T=${1:-r2f}
L=gpurun_out/${T}_quick.log; : > $L
OFF="FILM_STRAIGHT=0 FILM_RGB_FUSE=0 FILM_FE0_TC=1 FILM_PLANE_SKIP=0 FILM_ARENA_REUSE=0 FILM_FLOW_HEAD_FUSE=0 FILM_DUAL=0 FILM_HALO=2"
env QUICK_BIG=1 timeout 300 python tools/gpu_quick.py default 2>&1 | grep -E "^QUICK" | tee -a $L
env QUICK_BIG=1 $OFF timeout 300 python tools/gpu_quick.py all_off 2>&1 | grep -E "^QUICK" | tee -a $L
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 --op-table gpurun_out/${T}_ops.csv > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; echo "reference rc=$?"
cut -c1-400 gpurun_out/${T}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 200 --csv --log-file gpurun_out/${T}_counters.csv python tools/profile_step.py 0 > gpurun_out/${T}_ncu_counters.log 2>&1
ls -la gpurun_out | grep ${T}
