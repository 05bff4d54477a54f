#!/bin/bash
# Round 1h GPU job: full GPU suite, Path B benches with the parallel partial reduction, multi-resolution bench with the pinned kNN lists.
mkdir -p gpurun_out
O=gpurun_out/r01h
T0=$(date +%s)
timeout 900 python -m pytest tests -q -m gpu --durations=8 > ${O}_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s"; tail -14 ${O}_pytest.log
timeout 300 python bench_reg.py --steps 3 --warmup 2 > ${O}_reg_pinhole.json 2> ${O}_reg_pinhole.err; echo "reg pinhole rc=$? t=$(( $(date +%s) - T0 ))s"
B2_K12=thread timeout 300 python bench_reg.py --steps 3 --warmup 2 --no-cpu-baseline > ${O}_reg_pinhole_thread.json 2> ${O}_reg_pinhole_thread.err; echo "reg pinhole thread rc=$? t=$(( $(date +%s) - T0 ))s"
timeout 300 python bench_reg.py --camera benchmark --steps 3 --warmup 2 > ${O}_reg_fisheye.json 2> ${O}_reg_fisheye.err; echo "reg fisheye rc=$? t=$(( $(date +%s) - T0 ))s"
B2_MS_TRACE=1 timeout 600 python bench_multiscale.py > ${O}_ms.json 2> ${O}_ms.err; echo "ms rc=$? t=$(( $(date +%s) - T0 ))s"
python - <<'PY'
import json
for f in ["reg_pinhole","reg_pinhole_thread","reg_fisheye"]:
    try:
        d=json.loads(open("gpurun_out/r01h_%s.json"%f).read().strip().splitlines()[-1]); p=d["per_scale"]["0"]
        print(f, "%.3g evals/s"%d["value"], "acc %.2f ms jac %.2f ms total %.2f ms"%(p["ms_accumulate_kernels"],p["ms_jacobian_kernels"],1e3*p["s_per_accumulate"]), d["lm_apply"])
    except Exception as e: print(f, e)
try:
    d=json.loads(open("gpurun_out/r01h_ms.json").read().strip().splitlines()[-1]); print(d["value"], d["seconds"], d["merge"], d.get("cpu_baseline",{}).get("value"))
except Exception as e: print(e)
PY
grep "b2_ms_point" gpurun_out/r01h_ms.err | awk '{n+=$4; k+=$7; s+=$(NF-1)} END {print "neighbour trace: points", n, "kNN ms", k, "shuffle ms", s}'
