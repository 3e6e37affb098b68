#!/bin/bash
# Profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then ONE --set full capture of the
# dominant kernels.  Outputs go to gpurun_out/ (scratch); summaries are copied into profiles/ by scripts/summarise_ncu.py.
# usage: scripts/profile.sh <tag> [bench args...]
set -u
TAG=${1:-r01}; shift || true
OUT=gpurun_out
mkdir -p $OUT
# warm-up 2: the first sweep on each of the two layouts builds its guiding cache (10 probe passes of bwd_kernel); steady state after
CMD="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e $*"
echo "== plain: $CMD"
$CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/plain_$TAG.log; exit 1; }
tail -1 $OUT/plain_$TAG.log | cut -c1-400
echo "== launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "rc=$?"; tail -3 $OUT/ncu_launch_$TAG.log | cut -c1-300
echo "== full capture of the sweep kernels (draw, invsolve_ll, bwd)"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:fwd_kernel|bwd_kernel|bwd_coop_kernel|cache_apply_kernel' -s ${NCU_SKIP:-28} -c ${NCU_COUNT:-6} -o $OUT/prof_$TAG -f $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "rc=$?"; tail -3 $OUT/ncu_full_$TAG.log | cut -c1-300
# gpurun copies back at most 64 MiB: keep the exported pages, drop the report unless KEEP_REP=1
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
[ "${KEEP_REP:-0}" = 1 ] || rm -f $OUT/prof_$TAG.ncu-rep
ls -la $OUT
