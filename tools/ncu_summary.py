"""Key numbers of one kernel from an `ncu --set full` report: `python tools/ncu_summary.py X.ncu-rep`."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_op_tcgen05_mma.sum" if False else "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__cycles_active.avg",
]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print(f"== {rep}")
    for k in want:
        if k in d and d[k] != "":
            print(f"{k:72s} {d[k]:>18s} {u.get(k, '')}")
    tens = [k for k in hdr if "tensor" in k and "pct" in k]
    for k in tens[:6]:
        print(f"{k:72s} {d[k]:>18s} {u.get(k, '')}")
    break
