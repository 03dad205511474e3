#!/bin/bash
# Gradient-exchange transports side by side on N GPUs of one box:  bash tools/exchange_ab.sh <N> [transports...]
# -> gpurun_out/exchange_ab_<N>gpu.txt (bench.py --quick lines: ms per step, rank mismatch, error vs all-gather mean)
N=${1:-2}; shift
T=${@:-nccl ce nvls}
mkdir -p gpurun_out
OUT=gpurun_out/exchange_ab_${N}gpu.txt
: > $OUT
PORT=29600
for t in $T; do
  PORT=$((PORT + 1))
  echo "== $t" | tee -a $OUT
  YVB200_EXCHANGE=$t timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $PORT bench.py --gpus $N --quick --steps 20 --warmup 5 > gpurun_out/exchange_ab_$t.log 2>&1
  echo "rc=$?" | tee -a $OUT
  grep '"quick"' gpurun_out/exchange_ab_$t.log | tee -a $OUT
  grep -i "warn\|error\|Traceback" gpurun_out/exchange_ab_$t.log | head -8 | tee -a $OUT
  tail -4 gpurun_out/exchange_ab_$t.log | cut -c1-400 >> $OUT
done
