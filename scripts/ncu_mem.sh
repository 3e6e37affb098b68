#!/bin/bash
# Memory-traffic metrics of the sweep kernels for one or more libdmt variants (cheap metric set, few replays).
# usage: scripts/ncu_mem.sh spec1 spec2 ...   spec = tag[:ENV=VAL...]
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_lookup_miss.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum"
for spec in "$@"; do
  IFS=':' read -r -a parts <<< "$spec"
  t=${parts[0]}; envs=("${parts[@]:1}")
  if [ "$t" = base ]; then LIB=diffusionmcmctools.jl_b200/libdmt.so; else LIB=diffusionmcmctools.jl_b200/libdmt_$t.so; fi
  name=$(echo "$spec" | tr ':=' '__')
  CMD="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e"
  env DMT_LIB=$PWD/$LIB "${envs[@]}" $CMD > gpurun_out/mem_plain_$name.log 2>&1 || { echo "$name plain run failed"; continue; }
  env DMT_LIB=$PWD/$LIB "${envs[@]}" ncu --metrics $M --clock-control none --kernel-name-base demangled -k 'regex:fwd_kernel|bwd_kernel' -s 5 -c 6 \
      --csv --log-file gpurun_out/mem_$name.csv $CMD > gpurun_out/mem_ncu_$name.log 2>&1
  python - "$name" <<'EOF'
import csv, sys, collections
name = sys.argv[1]
rows = [r for r in csv.reader(open("gpurun_out/mem_%s.csv" % name)) if len(r) > 10]
hdr = rows[0]; d = collections.OrderedDict()
for r in rows[1:]:
    x = dict(zip(hdr, r))
    d.setdefault((x["ID"], x["Kernel Name"].split("(")[0][-40:]), {})[x["Metric Name"]] = (x["Metric Value"], x["Metric Unit"])
print("==", name)
for (i, k), m in d.items():
    g = lambda n: float(m[n][0].replace(",", "")) if n in m else float("nan")
    print("  %-38s %7.3f ms  dramR %6.2f %s dramW %6.2f  texR %.3e sect  texW %.3e  L2hit %5.1f%%  dram%% %5.1f" % (
        k, g("gpu__time_duration.sum") / (1e6 if m["gpu__time_duration.sum"][1] == "ns" else 1), g("dram__bytes_read.sum"), m["dram__bytes_read.sum"][1],
        g("dram__bytes_write.sum"), g("lts__t_sectors_srcunit_tex_op_read.sum"), g("lts__t_sectors_srcunit_tex_op_write.sum"),
        g("lts__t_sector_hit_rate.pct"), g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
EOF
done
