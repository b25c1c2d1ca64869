#!/bin/bash
# Round-end profile collection on one B200 (run under gpurun): launch lists (cold = ncu default cache
# control, warm = --cache-control none), per-launch DRAM traffic, and `--set full` captures of the top
# kernels. Outputs land in gpurun_out/prof2/ ; summaries are made from them with tools/summarize_launches.py,
# tools/sum_traffic.py and tools/ncu_summary.py and committed under profiles/ (r2_*).
set -u
O=gpurun_out/prof2
mkdir -p $O
NCU="timeout 420 ncu --clock-control none"
$NCU --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --csv --log-file $O/launches_warm.csv python tools/profile_step.py > $O/warm.log 2>&1
$NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/launches_cold.csv python tools/profile_step.py > $O/cold.log 2>&1
$NCU --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file $O/traffic.csv python tools/profile_step.py > $O/traffic.log 2>&1
for w in wo_h w13_h; do
  $NCU --set full --import-source on -k regex:"vb_gemm|vb_conv" -s 4 -c 1 -o $O/full_$w -f python tools/profile_gemm_one.py $w > $O/full_$w.log 2>&1
done
$NCU --set full --import-source on -k regex:attn_fused -s 3 -c 1 -o $O/full_attn -f python tools/profile_attn.py > $O/full_attn.log 2>&1
$NCU --set full --import-source on -k regex:resblock_pair -s 4 -c 1 -o $O/full_pair32 -f python tools/profile_pair.py 32 7 3 > $O/full_pair32.log 2>&1
$NCU --set full --import-source on -k regex:resblock_pair -s 4 -c 1 -o $O/full_pair64 -f python tools/profile_pair.py 64 3 1 > $O/full_pair64.log 2>&1
ls -la $O | tail -20
