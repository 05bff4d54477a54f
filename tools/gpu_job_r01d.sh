#!/bin/bash
# Round 1d GPU job: multi-resolution point-cloud parity, full GPU suite after the pool / BVH refactor, ICP bench (pooled allocations), K7 capture.
mkdir -p gpurun_out
O=gpurun_out/r01d
T0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_multiscale.py -x -q --durations=8 > ${O}_pytest_ms.log 2>&1; echo "pytest ms rc=$? t=$(( $(date +%s) - T0 ))s"; tail -15 ${O}_pytest_ms.log
timeout 700 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_multiscale.py > ${O}_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s"; tail -3 ${O}_pytest.log
timeout 600 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$? t=$(( $(date +%s) - T0 ))s"
B2_POOL=0 timeout 400 python bench.py --no-cpu-baseline --steps 3 > ${O}_bench_nopool.json 2> ${O}_bench_nopool.err; echo "bench nopool rc=$? t=$(( $(date +%s) - T0 ))s"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'kn_knn_normals' -c 1 -f -o ${O}_k7 python bench_normals.py --scan-w 5000 --scan-h 2500 --profile > ${O}_ncu_k7.log 2>&1; echo "ncu k7 rc=$? t=$(( $(date +%s) - T0 ))s"
[ -f ${O}_k7.ncu-rep ] && ncu -i ${O}_k7.ncu-rep --page raw --csv > ${O}_k7.raw.csv 2>/dev/null
cat ${O}_bench.json | head -c 4000; echo
cat ${O}_bench_nopool.json | head -c 1500; echo
tail -5 ${O}_bench.err
