#!/bin/bash
# Bisects the round-2 optimisations on the GPU: every configuration in its own process.
# Leaves the maximal working environment in gpurun_out/good_env.sh (sourced by the follow-up steps).
mkdir -p gpurun_out
L=gpurun_out/${BISECT_TAG:-r2c}_bisect.log
: > $L
OFF="FILM_STRAIGHT=0 FILM_RGB_FUSE=0 FILM_FE0_TC=1 FILM_PLANE_SKIP=0 FILM_ARENA_REUSE=0 FILM_FLOW_HEAD_FUSE=0"
run() { tag=$1; shift; env "$@" timeout 300 python tools/gpu_quick.py $tag 2>&1 | grep -E "^QUICK" >> $L || echo "QUICK $tag FAIL (no output / timeout)" >> $L; tail -1 $L; }
run all_off $OFF
run only_straight $OFF FILM_STRAIGHT=1
run only_rgb_fuse $OFF FILM_RGB_FUSE=1
run only_fe0_simt $OFF FILM_FE0_TC=0
run only_plane_skip $OFF FILM_PLANE_SKIP=1
run only_arena_reuse $OFF FILM_ARENA_REUSE=1
run only_flow_head_fuse $OFF FILM_FLOW_HEAD_FUSE=1
run all_off_no_pair $OFF FILM_2CTA=0
run all_on FILM_STRAIGHT=1
# maximal working set
GOOD=""
for f in STRAIGHT:only_straight:0 RGB_FUSE:only_rgb_fuse:0 PLANE_SKIP:only_plane_skip:0 ARENA_REUSE:only_arena_reuse:0 FLOW_HEAD_FUSE:only_flow_head_fuse:0; do
  IFS=: read var tag offval <<< "$f"
  if ! grep -q "QUICK $tag ok" $L; then GOOD="$GOOD FILM_$var=$offval"; fi
done
if ! grep -q "QUICK only_fe0_simt ok" $L; then GOOD="$GOOD FILM_FE0_TC=1"; fi
echo "export $GOOD" > gpurun_out/good_env.sh
[ -z "$GOOD" ] && echo "true" > gpurun_out/good_env.sh
echo "good env: $GOOD" | tee -a $L
env $GOOD QUICK_BIG=1 timeout 300 python tools/gpu_quick.py good_set 2>&1 | grep -E "^QUICK" | tee -a $L
env $GOOD QUICK_BIG=1 FILM_DUAL=0 timeout 300 python tools/gpu_quick.py good_set_nodual 2>&1 | grep -E "^QUICK" | tee -a $L
env $GOOD QUICK_BIG=1 FILM_HALO=2 timeout 300 python tools/gpu_quick.py good_set_halo2 2>&1 | grep -E "^QUICK" | tee -a $L
