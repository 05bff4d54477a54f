#!/bin/bash
# Round 1m GPU job: heap k-best lists (k up to 800), cleaner bench, normals regression check.
mkdir -p gpurun_out
O=gpurun_out/r01m
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_cleaner.py tests/test_gpu_normals.py tests/test_gpu_multiscale.py -q -m gpu > ${O}_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s"; tail -8 ${O}_pytest.log
timeout 300 python bench_normals.py --no-cpu-baseline > ${O}_normals.json 2> ${O}_normals.err; echo "normals rc=$? t=$(( $(date +%s) - T0 ))s"; cut -c1-300 ${O}_normals.json
timeout 600 python bench_cleaner.py > ${O}_cleaner.json 2> ${O}_cleaner.err; echo "cleaner rc=$? t=$(( $(date +%s) - T0 ))s"; cut -c1-900 ${O}_cleaner.json; tail -3 ${O}_cleaner.err
