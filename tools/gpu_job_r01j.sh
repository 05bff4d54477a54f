#!/bin/bash
# Round 1j GPU job: full GPU suite (incl. the GroundTruthCreator kernels) and the contract bench line.
mkdir -p gpurun_out
O=gpurun_out/r01j
T0=$(date +%s)
timeout 540 python -m pytest tests -q -m gpu --durations=8 > ${O}_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s"; tail -14 ${O}_pytest.log
timeout 300 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$? t=$(( $(date +%s) - T0 ))s"; tail -c 1500 ${O}_bench.json
