"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv`."""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[idx["# Samples"]])
    except ValueError:
        continue
    data.append((s, r))
total = sum(s for s, _ in data)
print(f"total samples {total}")
for s, r in sorted(data, key=lambda x: -x[0])[:n]:
    st = {c: int(r[idx[c]]) for c in stall_cols if r[idx[c]].isdigit() and int(r[idx[c]]) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{100 * s / max(total, 1):5.1f}%  {r[idx['Source']].strip()[:70]:70s} {top}")
