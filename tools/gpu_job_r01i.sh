#!/bin/bash
# Round 1i GPU job: GroundTruthCreator kernels, pipelined neighbour search, final Path B captures.
mkdir -p gpurun_out
O=gpurun_out/r01i
T0=$(date +%s)
timeout 900 python -m pytest tests -q -m gpu --durations=6 > ${O}_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s"; tail -12 ${O}_pytest.log
B2_MS_TRACE=1 timeout 600 python bench_multiscale.py > ${O}_ms.json 2> ${O}_ms.err; echo "ms rc=$? t=$(( $(date +%s) - T0 ))s"
NCU="ncu --set full --clock-control none --profile-from-start off"
timeout 300 $NCU --import-source on -k regex:'kr_jacobians|kr_residual_weights|kr_accumulate' -c 3 -f -o ${O}_reg_pinhole_k11_k12 python bench_reg.py --images 1 --profile > ${O}_ncu1.log 2>&1; echo "ncu1 rc=$? t=$(( $(date +%s) - T0 ))s"
[ -f ${O}_reg_pinhole_k11_k12.ncu-rep ] && ncu -i ${O}_reg_pinhole_k11_k12.ncu-rep --page raw --csv > ${O}_reg_pinhole_k11_k12.raw.csv 2>/dev/null
timeout 300 $NCU -k regex:'km_|kr_min_max|kr_undist' -c 40 -f -o ${O}_ms python bench_multiscale.py --scans 1 --scan-w 3000 --scan-h 1200 --steps 1 --warmup 0 --no-cpu-baseline --profile > ${O}_ncu2.log 2>&1; echo "ncu2 rc=$? t=$(( $(date +%s) - T0 ))s"
[ -f ${O}_ms.ncu-rep ] && ncu -i ${O}_ms.ncu-rep --page raw --csv > ${O}_ms.raw.csv 2>/dev/null
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r01i_ms.json").read().strip().splitlines()[-1]); print(d["value"], d["seconds"], d["merge"], d.get("cpu_baseline",{}).get("value"))
except Exception as e: print(e)
PY
grep "b2_ms_point" gpurun_out/r01i_ms.err | awk '{n+=$4; k+=$7; s+=$11} END {print "neighbour trace: points", n, "kNN ms", k, "shuffle ms", s}'
tail -3 ${O}_ncu2.log
