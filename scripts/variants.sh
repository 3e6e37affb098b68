#!/bin/bash
# Run bench.py once per tuning variant of libdmt (built by `build.py --tag <t> -D...`) and print the per-kernel times.
# usage: scripts/variants.sh "<bench args>" spec1 spec2 ...   spec = tag[:ENV=VAL[:ENV=VAL...]]  ("base" = the shipped libdmt.so)
ARGS=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=':' read -r -a parts <<< "$spec"
  t=${parts[0]}
  envs=("${parts[@]:1}")
  if [ "$t" = base ]; then LIB=diffusionmcmctools.jl_b200/libdmt.so; else LIB=diffusionmcmctools.jl_b200/libdmt_$t.so; fi
  name=$(echo "$spec" | tr ':=' '__')
  env DMT_LIB=$PWD/$LIB "${envs[@]}" timeout 300 python bench.py --no-cpu-baseline --no-e2e $ARGS > gpurun_out/var_$name.log 2>&1
  python - "$name" <<'EOF'
import json, sys
t = sys.argv[1]
try:
    j = json.loads(open("gpurun_out/var_%s.log" % t).read().strip().splitlines()[-1])
    print("%-26s ms/step %7.3f  value %.3e  draw_frac %.3f  " % (t, j["ms_per_step"], j["value"], j["roofline"]["frac"]) +
          " ".join("%s=%.3f" % (k, v) for k, v in j["kernel_ms"].items()) + "  | by layout: " +
          " ".join("%s=%s" % (k, "/".join("%.2f" % x for x in v)) for k, v in j.get("kernel_ms_by_layout", {}).items()))
except Exception as e:
    print(t, "FAILED", e); print(open("gpurun_out/var_%s.log" % t).read()[-800:])
EOF
done
