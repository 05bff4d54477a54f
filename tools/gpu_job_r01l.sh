#!/bin/bash
# Round 1l GPU job: point-cloud tools (LSOR, mesh distance, splats) parity + the full GPU suite after the K7 statistic modes.
mkdir -p gpurun_out
O=gpurun_out/r01l
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_cleaner.py -q -m gpu > ${O}_cleaner.log 2>&1; echo "cleaner rc=$? t=$(( $(date +%s) - T0 ))s"; tail -30 ${O}_cleaner.log
timeout 540 python -m pytest tests -q -m gpu --deselect tests/test_gpu_cleaner.py > ${O}_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s"; tail -5 ${O}_pytest.log
