#!/bin/bash
# A/B of NCCL settings for the data-parallel step at N GPUs (measurement helper, not product code).
# usage: tools/nccl_sweep.sh N   -> one JSON line per setting in gpurun_out/r2_nccl_sweep_nN.jsonl
N=${1:-8}
OUT=gpurun_out/r2_nccl_sweep_n$N.jsonl
: > $OUT
run() {
  env "$@" timeout -k 10 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 30 --warmup 3 --only-device >> $OUT 2>> gpurun_out/r2_nccl_sweep_n$N.err
  echo "rc=$? $*"
}
run NCCL_DEBUG=WARN
run NCCL_MAX_CTAS=8
run NCCL_MAX_CTAS=16 NCCL_MIN_CTAS=16
run NCCL_MAX_CTAS=4
run NCCL_ALGO=Ring
cat $OUT
