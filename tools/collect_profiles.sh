#!/bin/bash
# Round-end profile collection on one B200 (run under gpurun): launch lists (cold = ncu default cache
# control, warm = --cache-control none), per-launch DRAM traffic, and `--set full` captures of the top
# kernels. Outputs land in gpurun_out/prof/ ; summaries are made from them with tools/summarize_launches.py,
# tools/sum_traffic.py and tools/ncu_summary.py and committed under profiles/.
set -u
O=gpurun_out/prof
mkdir -p $O
NCU="timeout 300 ncu --clock-control none"
$NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/launches_cold.csv python tools/profile_step.py > $O/cold.log 2>&1
$NCU --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --csv --log-file $O/launches_warm.csv python tools/profile_step.py > $O/warm.log 2>&1
$NCU --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file $O/traffic.csv python tools/profile_step.py > $O/traffic.log 2>&1
for w in wo w13 voc64; do
  $NCU --set full --import-source on -k regex:"vb_gemm|vb_conv" -s 4 -c 1 -o $O/full_$w -f python tools/profile_gemm_one.py $w > $O/full_$w.log 2>&1
done
$NCU --set full --import-source on -k regex:attn_fused -s 3 -c 1 -o $O/full_attn -f python tools/profile_attn.py > $O/full_attn.log 2>&1
ls -la $O | tail -20
