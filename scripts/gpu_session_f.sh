#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-8}
CC_MULTICAST=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/gpu_gather_mc.py 2>&1 | grep -E "^\{|Error|error" | cut -c1-1200
CC_MULTICAST=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 scripts/gpu_gather_mc.py 2>&1 | grep -E "^\{|Error|error" | cut -c1-1200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 scripts/gpu_scale_profile.py > gpurun_out/r02_scale_profile_n$N.log 2>&1; echo "prof=$?"; tail -6 gpurun_out/r02_scale_profile_n$N.log | cut -c1-500
