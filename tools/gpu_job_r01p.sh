#!/bin/bash
# Round 1p GPU job (2 GPUs): multi-GPU parity (ICP with speculative LM trials, Path B, normals) + the 2-GPU bench line.
mkdir -p gpurun_out
O=gpurun_out/r01p
T0=$(date +%s)
timeout 280 python -m pytest tests/test_gpu_reg_dist.py -q -m gpu > ${O}_dist.log 2>&1; echo "dist rc=$? t=$(( $(date +%s) - T0 ))s"; tail -12 ${O}_dist.log | cut -c1-400
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > ${O}_bench2.json 2> ${O}_bench2.err; echo "bench2 rc=$? t=$(( $(date +%s) - T0 ))s"
tail -c 1200 ${O}_bench2.json | cut -c1-1200; tail -3 ${O}_bench2.err
