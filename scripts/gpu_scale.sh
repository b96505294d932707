#!/bin/bash
# Run on a multi-GPU box (gpurun --gpus 8): the bench at N = 2, 4, 8 ranks, one JSON line each under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for n in ${1:-2 4 8}; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 50 --warmup 3 --e2e-steps 2 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  echo "N=$n rc=$?"; tail -c 3000 gpurun_out/scale_n$n.json; tail -3 gpurun_out/scale_n$n.err
done
