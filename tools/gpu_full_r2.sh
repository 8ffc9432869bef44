#!/bin/bash
# round 2, GPU call B: full parity suite, kernel-option A/B benches on one box, precision study (2 seeds),
# the complete bench line (with the 4K / 8K / 720p workloads), ncu evidence.
mkdir -p gpurun_out
T=${1:-r2c}
export BISECT_TAG=$T
bash tools/gpu_bisect.sh
source gpurun_out/good_env.sh
cat gpurun_out/good_env.sh
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -25 gpurun_out/${T}_pytest.log
for cfg in "default" "FILM_HALO=2" "FILM_HALO=0" "FILM_2CTA=2" "FILM_2CTA=0" "FILM_FE0_TC=1" "FILM_RGB_FUSE=0" "FILM_DUAL=0" "FILM_STRAIGHT=0" "FILM_PLANE_SKIP=0" "FILM_ARENA_REUSE=0" "FILM_FLOW_HEAD_FUSE=0" "FILM_FLOW_HEAD_FUSE=2"; do
  name=$(echo $cfg | tr '=' '_')
  if [ "$cfg" = "default" ]; then envs=""; else envs="$cfg"; fi
  env $envs timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-workloads --op-table gpurun_out/${T}_ops_$name.csv > gpurun_out/${T}_bench_$name.json 2> gpurun_out/${T}_bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_$name.json"))
    print("$cfg", "ms/step %.3f" % d["ms_per_step"], "fps %.2f" % d["value"], "e2e %.2f" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"], "conv ms %.3f" % d["roofline"]["conv_kernel_ms_per_step"], "gather GB/s %.0f" % d["gather"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$cfg failed", e)
PY
done
timeout 1200 python tools/precision_study.py --height 1080 --width 1920 --seeds 0,1 --out gpurun_out/${T}_precision_1080p > gpurun_out/${T}_precision.log 2>&1; echo "study rc=$?"
tail -3 gpurun_out/${T}_precision.log
timeout 900 python bench.py --steps 20 --warmup 3 --op-table gpurun_out/${T}_ops.csv > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
# ncu: launch list of one eager call (second call); light sections for EVERY kernel of a call (DRAM bytes, L2->SM bytes, tensor
# pipe, issue slots per launch); full set + source for the first launches of each tensor-core kernel and the gathers.
# gpurun brings back at most 64 MiB: keep the reports small and drop the biggest ones if the total still exceeds it.
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_step.py 1 > gpurun_out/${T}_ncu_list.log 2>&1
# every launch of ONE call with the counters the roofline discussion needs (explicit metrics: small CSV)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 200 --csv --log-file gpurun_out/${T}_counters.csv python tools/profile_step.py 0 > gpurun_out/${T}_ncu_counters.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3_tc2 -c 4 -f -o gpurun_out/${T}_pair python tools/profile_step.py 0 > gpurun_out/${T}_ncu_pair.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_conv3x3_tc<" -c 4 -f -o gpurun_out/${T}_single python tools/profile_step.py 0 > gpurun_out/${T}_ncu_single.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_flow_warp|k_fusion_warp" -c 4 -f -o gpurun_out/${T}_gather python tools/profile_step.py 0 > gpurun_out/${T}_ncu_gather.log 2>&1
while [ $(du -sm gpurun_out | cut -f1) -gt 58 ]; do big=$(ls -S gpurun_out/*.ncu-rep | head -1); echo "dropping $big"; rm -f $big; done
# memcheck of one small network call (arena reuse, lo-plane skipping, fused epilogues): any out-of-bounds access shows up here
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/${T}_memcheck.log
tail -4 gpurun_out/${T}_memcheck.log
ls -la gpurun_out | grep ${T}
