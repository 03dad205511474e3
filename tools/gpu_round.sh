#!/bin/bash
# One GPU-box visit: parity tests, bench line, graph timeline, ncu launch list, ncu full capture of the GEMM.
# usage: bash tools/gpu_round.sh <tag>      (outputs land in gpurun_out/<tag>_*)
T=${1:-r1}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_smi.txt
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 400 gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
timeout 300 python tools/timeline.py > gpurun_out/${T}_timeline.txt 2>&1
rm -f gpurun_out/trace.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_eager.csv \
  python bench.py --steps 1 --warmup 1 --no-graph --quick > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:yv_gemm -c 12 -o gpurun_out/${T}_gemm_full -f \
  python tools/ncu_gemm.py > gpurun_out/${T}_ncu_gemm.log 2>&1
ls -la gpurun_out
