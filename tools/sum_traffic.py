"""Sum dram bytes per kernel family from an ncu CSV with dram__bytes_read.sum / dram__bytes_write.sum."""
import csv
import json
import sys
from collections import defaultdict

path, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if not l.startswith("==")]
agg = defaultdict(lambda: [0, 0.0])
ids = defaultdict(set)
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in csv.DictReader(lines):
    if not r.get("Metric Name", "").startswith("dram__bytes"):
        continue
    name = r["Kernel Name"].split("(")[0].replace("void ", "")
    gemm = any(k in name for k in ("vb_gemm_kernel", "vb_gemm2_kernel", "vb_conv_halo_kernel", "vb_chain_kernel"))
    fam = "vb_gemm_kernel" if gemm else ("attn_fused_kernel" if "attn_fused" in name else name)  # all tcgen05 GEMM kernels
    agg[fam][1] += float(r["Metric Value"].replace(",", "")) * mult.get(r.get("Metric Unit", "byte"), 1)
    ids[fam].add(r["ID"])
res = {k: {"launches": len(ids[k]), "dram_bytes": v[1]} for k, v in agg.items()}
json.dump(res, open(out, "w"), indent=1, sort_keys=True)
for k, v in sorted(res.items(), key=lambda kv: -kv[1]["dram_bytes"])[:8]:
    print(f"{v['dram_bytes'] / 1e6:10.1f} MB  {v['launches']:5d} launches  {k}")
