#!/bin/bash
# Multi-GPU validation on one box: parity of the peer-memory exchange, the weak-scaling bench line (with the config-5
# strong-scaling leg and its gather parity) bracketed by the NVLink byte counters of GPU 0, and -- with `ncu` as the
# third argument -- ncu captures of the exchange kernels in APPLICATION replay mode (kernel replay fails with UnknownError on
# the first kernel that touches a CUDA-IPC mapping: ncu cannot save / restore peer memory).
# usage: tools/multi_gpu.sh <N> <tag> [ncu]
N=${1:-2}; TAG=${2:-r02}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 tools/exchange_check.py > gpurun_out/${TAG}_xchk_n$N.log 2>&1
nvidia-smi nvlink -gt d -i 0 > gpurun_out/${TAG}_nvlink_n$N.before.txt 2>&1
timeout 900 $RUN --master-port 29512 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "rc=$?" >> gpurun_out/${TAG}_bench_n$N.err
nvidia-smi nvlink -gt d -i 0 > gpurun_out/${TAG}_nvlink_n$N.after.txt 2>&1
if [ "$3" = ncu ]; then
  # shard shape of the headline step (32 images per rank, T = 200 rows)
  for M in gpu__time_duration.sum,smsp__inst_executed.sum nvltx__bytes.sum,nvlrx__bytes.sum; do
    T=$(echo $M | cut -c1-5)
    XCHK_BATCH=32 XCHK_T=200 timeout 600 ncu --target-processes all --replay-mode application --clock-control none -k regex:exchange \
        --metrics $M -c 60 --csv --log-file gpurun_out/${TAG}_ncu_exchange_n${N}_${T}_%p.csv \
        $RUN --master-port 29513 tools/exchange_check.py > gpurun_out/${TAG}_ncu_exchange_n${N}_$T.log 2>&1
    echo "ncu rc=$?" >> gpurun_out/${TAG}_ncu_exchange_n${N}_$T.log
  done
  python tools/ncu_exchange_summary.py gpurun_out/${TAG}_ncu_exchange_n${N}_ > gpurun_out/${TAG}_ncu_exchange_n${N}.md 2>&1
fi
