#!/bin/bash
# Round 1n GPU job: all 15 camera models (camera eval parity, Path B with every distortion family) + Path B regression suite.
mkdir -p gpurun_out
O=gpurun_out/r01n
T0=$(date +%s)
timeout 400 python -m pytest tests/test_gpu_reg_models.py -q -m gpu -x > ${O}_models.log 2>&1; echo "models rc=$? t=$(( $(date +%s) - T0 ))s"; tail -40 ${O}_models.log | cut -c1-300
timeout 500 python -m pytest tests/test_gpu_reg.py tests/test_gpu_reg_camera.py tests/test_gpu_reg_rig.py tests/test_gpu_ground_truth.py tests/test_gpu_multiscale.py -q -m gpu > ${O}_reg.log 2>&1; echo "reg rc=$? t=$(( $(date +%s) - T0 ))s"; tail -15 ${O}_reg.log | cut -c1-300
