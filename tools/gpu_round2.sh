#!/bin/bash
# One GPU-box visit for the round-2 evidence: bench lines, timeline, ncu launch list, ncu full captures (GEMM + fused
# attention), memcheck of the fused attention kernels.   usage: bash tools/gpu_round2.sh <tag>   -> gpurun_out/<tag>_*
T=${1:-r2}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_smi.txt
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 300 gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
timeout 300 python tools/timeline.py > gpurun_out/${T}_timeline.txt 2>&1
rm -f gpurun_out/trace.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/${T}_launches_eager.csv \
  python bench.py --steps 1 --warmup 1 --no-graph --quick > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:yv_gemm -c 8 -o gpurun_out/${T}_gemm_full -f \
  python tools/ncu_gemm.py > gpurun_out/${T}_ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:yv_attn -c 16 -o gpurun_out/${T}_attn_full -f \
  python tools/ncu_attn.py > gpurun_out/${T}_ncu_attn.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --report-api-errors no --log-file gpurun_out/${T}_memcheck_attn.log \
  python -m pytest tests/test_attn_gpu.py -q -x -k "80-80 or 130-70 or 36-20" > gpurun_out/${T}_memcheck_attn.out 2>&1
grep -B2 -A12 "Invalid\|out of bounds\|misaligned\|Error" gpurun_out/${T}_memcheck_attn.log | head -60
# condense the ncu reports on the box (they are too large to travel back) and drop them
YV_PROFILE_OUT=gpurun_out python tools/ncu_summarise.py ${T}
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out
