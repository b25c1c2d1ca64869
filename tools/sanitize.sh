#!/bin/bash
# compute-sanitizer passes over the tcgen05/TMA/mbarrier kernels (run under gpurun): memcheck on the op-level
# tests of every hand-rolled pipeline, racecheck (shared-memory hazards) on the small cases. Logs -> gpurun_out/sanitize/
set -u
O=gpurun_out/sanitize
mkdir -p $O
CS="timeout 900 compute-sanitizer --print-limit 20"
T="tests/test_gemm_gpu.py tests/test_gemm2_gpu.py tests/test_attn_gpu.py tests/test_conv_halo_gpu.py tests/test_chain_gpu.py tests/test_resblock_gpu.py tests/test_route_gpu.py tests/test_ops_gpu.py"
$CS --tool memcheck python -m pytest $T -x -q -k "not 60000" > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log
$CS --tool racecheck --racecheck-report all python -m pytest tests/test_resblock_gpu.py tests/test_conv_halo_gpu.py tests/test_gemm2_gpu.py tests/test_attn_gpu.py -x -q -k "not 60000 and not 4111" > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck.log
$CS --tool synccheck python -m pytest tests/test_resblock_gpu.py tests/test_chain_gpu.py tests/test_attn_gpu.py -x -q -k "not 60000" > $O/synccheck.log 2>&1; echo "synccheck rc=$?" >> $O/synccheck.log
tail -5 $O/memcheck.log $O/racecheck.log $O/synccheck.log
