#!/bin/bash
# First GPU call of the next round: validate the peer-memory collectives on N GPUs of one box,
# then compare the sharded bench over NCCL and over peer memory.
#   gpurun --gpus 2 --timeout 420 -- 'bash profiles/run_peer_validation.sh 2'
# Every step is bounded by `timeout`; the kernels' own waits are bounded too (~2 s).
N="${1:-2}"
OUT=gpurun_out/peer_validation_n${N}.txt
mkdir -p gpurun_out
{
  echo "== peer collectives test (2 ranks)"
  CDR_TEST_PEER=1 timeout 300 python -m pytest tests/test_distributed.py -q -m gpu -k peer 2>&1 | tail -15
  for WL in gpnh aa; do
    for PEER in 0 1; do
      echo "== bench --gpus $N --workload $WL CDR_PEER_COLLECTIVES=$PEER"
      CDR_PEER_COLLECTIVES=$PEER timeout 240 python -m torch.distributed.run --nnodes=1 \
        --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
        bench.py --gpus "$N" --steps 40 --warmup 5 --workload $WL --cpu-steps 0 2>&1 | tail -1 | \
        python -c "import sys, json; d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'final_cost') if k in d})"
    done
  done
} 2>&1 | tee "$OUT"
