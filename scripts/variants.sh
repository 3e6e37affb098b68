#!/bin/bash
# Run bench.py once per tuning variant of libdmt (built by `build.py --tag <t> -D...`) and print the per-kernel times.
# usage: scripts/variants.sh "<bench args>" tag1 tag2 ...   ("base" = the shipped libdmt.so)
ARGS=$1; shift
mkdir -p gpurun_out
for t in "$@"; do
  if [ "$t" = base ]; then LIB=diffusionmcmctools.jl_b200/libdmt.so; else LIB=diffusionmcmctools.jl_b200/libdmt_$t.so; fi
  DMT_LIB=$PWD/$LIB timeout 300 python bench.py --no-cpu-baseline --no-e2e $ARGS > gpurun_out/var_$t.log 2>&1
  python - "$t" <<'EOF'
import json, sys
t = sys.argv[1]
try:
    j = json.loads(open("gpurun_out/var_%s.log" % t).read().strip().splitlines()[-1])
    print("%-12s ms/step %7.3f  value %.3e  draw_frac %.3f  " % (t, j["ms_per_step"], j["value"], j["roofline"]["frac"]) +
          " ".join("%s=%.3f" % (k, v) for k, v in j["kernel_ms"].items()))
except Exception as e:
    print(t, "FAILED", e); print(open("gpurun_out/var_%s.log" % t).read()[-800:])
EOF
done
