"""Summarise an ncu report's source page: stall samples per SASS line region.  usage: ncu_src_summary.py rep [kernel_index] [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ki = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = starts[ki]; e = starts[ki + 1] if ki + 1 < len(starts) else len(rows)
print(rows[s][1][:150])
hdr = rows[s + 1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[s + 2:e] if len(r) >= len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("instructions", len(body), "samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in body) for h in stalls}
print({k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
rank = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
for i in sorted(rank):
    r = body[i]
    st = {h[6:]: int(r[ix[h]] or 0) for h in stalls if int(r[ix[h]] or 0)}
    print(i, r[ix["Source"]][:60].ljust(60), r[ix["# Samples"]], r[ix["Instructions Executed"]], dict(sorted(st.items(), key=lambda kv: -kv[1])[:3]))
