#!/bin/bash
# One GPU-box visit: parity tests, bench line, graph timeline, ncu launch list, ncu full capture of the GEMM.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 600 gpurun_out/bench.err
timeout 300 python tools/timeline.py > gpurun_out/timeline.txt 2>&1
rm -f gpurun_out/trace.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file gpurun_out/launches_eager.csv \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:yv_gemm -c 12 -o gpurun_out/gemm_full -f \
  python tools/ncu_gemm.py > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
