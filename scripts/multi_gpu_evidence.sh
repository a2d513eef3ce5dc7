#!/bin/bash
# The multi-GPU part of the evidence (separate gpurun calls: `gpurun --gpus N -- bash scripts/multi_gpu_evidence.sh N`):
# the strong-scaling bench line at N GPUs and, at N = 2, the NCCL sharded-stream test. scripts/collect_profiles.py picks the files up.
N=${1:-2}
O=gpurun_out/ev_r2
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 > ${O}_pytest_multi.log
  cat ${O}_pytest_multi.log
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > ${O}_bench_n$N.json 2> ${O}_bench_n$N.err
tail -c 600 ${O}_bench_n$N.json
