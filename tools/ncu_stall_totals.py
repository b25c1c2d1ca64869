"""Totals of warp-stall samples by reason from an ncu report's source page, with the spin-wait
instructions (mbarrier try_wait loops, the exit barrier) listed separately."""
import collections
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot, spin = collections.Counter(), collections.Counter()
n_inst = 0
for r in rows[2:]:
    if len(r) < len(hdr) or not r[idx["# Samples"]].isdigit():
        continue
    s = int(r[idx["# Samples"]])
    src = r[idx["Source"]]
    n_inst += 1
    big_wait = s > 300 and ("BRA" in src or "S2R" in src or "BAR" in src)
    for c in stall:
        if r[idx[c]].isdigit():
            (spin if big_wait else tot)[c] += int(r[idx[c]])
print(n_inst, "SASS instructions")
print("spin/wait instructions:", sum(spin.values()), dict(spin.most_common(4)))
print("everything else       :", sum(tot.values()), dict(tot.most_common(10)))
