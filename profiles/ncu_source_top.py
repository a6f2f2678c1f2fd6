"""Top stall-sample instructions of one kernel in an ncu report (--set full --import-source on):
    python profiles/ncu_source_top.py <rep> <kernel-id, e.g. ::fprop_ring64_kernel:1> [N]"""
import csv
import io
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[h]
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
ix = {k: i for i, k in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[ix[k]]) for r in data) for k in stall_cols}
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:n]
for i in sorted(top):
    r = data[i]
    st = {k: int(r[ix[k]]) for k in stall_cols if int(r[ix[k]]) > 0}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{i:6d} {int(r[ix['# Samples']]):6d} {100 * int(r[ix['# Samples']]) / tot:5.1f}%  {r[ix['Source']].strip()[:86]:86s} {main}")
