#!/bin/bash
# Multi-GPU validation on one box: parity of the peer-memory exchange, the weak-scaling bench line (with the config-5
# strong-scaling leg and its gather parity), and -- at N = 2 -- a single-pass ncu capture of the exchange kernels
# (no replay: the peers' flags would not repeat) with the NVLink byte counters.
# usage: tools/multi_gpu.sh <N> <tag>
N=${1:-2}; TAG=${2:-r02}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 tools/exchange_check.py > gpurun_out/${TAG}_xchk_n$N.log 2>&1
timeout 900 $RUN --master-port 29512 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "rc=$?" >> gpurun_out/${TAG}_bench_n$N.err
if [ "$N" = 2 ]; then
  timeout 600 ncu --target-processes all --replay-mode kernel --clock-control none -k regex:exchange \
      --metrics gpu__time_duration.sum,smsp__inst_executed.sum,nvltx__bytes.sum,nvlrx__bytes.sum,lts__t_sectors_op_write.sum \
      -c 40 --csv --log-file gpurun_out/${TAG}_ncu_exchange_n2_%p.csv \
      $RUN --master-port 29513 tools/exchange_check.py > gpurun_out/${TAG}_ncu_exchange_n2.log 2>&1
  echo "ncu rc=$?" >> gpurun_out/${TAG}_ncu_exchange_n2.log
fi
