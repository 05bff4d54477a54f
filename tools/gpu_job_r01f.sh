#!/bin/bash
# Round 1f GPU job: pinhole K12 on top of the pre-pass (A/B), multi-resolution bench at full size, e2e phase breakdown.
mkdir -p gpurun_out
O=gpurun_out/r01f
T0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_reg.py tests/test_gpu_reg_camera.py tests/test_gpu_reg_rig.py -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s"; tail -3 ${O}_pytest.log
timeout 300 python bench_reg.py --steps 3 --warmup 2 --no-cpu-baseline > ${O}_reg_pinhole.json 2> ${O}_reg_pinhole.err; echo "reg pinhole rc=$? t=$(( $(date +%s) - T0 ))s"
B2_K12=thread timeout 300 python bench_reg.py --steps 3 --warmup 2 --no-cpu-baseline > ${O}_reg_pinhole_thread.json 2> ${O}_reg_pinhole_thread.err; echo "reg pinhole thread rc=$? t=$(( $(date +%s) - T0 ))s"
timeout 600 python bench_multiscale.py > ${O}_ms.json 2> ${O}_ms.err; echo "ms rc=$? t=$(( $(date +%s) - T0 ))s"; tail -3 ${O}_ms.err
timeout 400 python bench.py --no-cpu-baseline --steps 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$? t=$(( $(date +%s) - T0 ))s"
python - <<'PY'
import json
for f in ["reg_pinhole","reg_pinhole_thread"]:
    try:
        d=json.loads(open("gpurun_out/r01f_%s.json"%f).read().strip().splitlines()[-1]); p=d["per_scale"]["0"]
        print(f, "%.3g evals/s"%d["value"], "acc %.2f ms jac %.2f ms total %.2f ms"%(p["ms_accumulate_kernels"],p["ms_jacobian_kernels"],1e3*p["s_per_accumulate"]))
    except Exception as e: print(f, e)
try:
    d=json.loads(open("gpurun_out/r01f_bench.json").read().strip().splitlines()[-1]); print(d["value"], d["e2e"])
except Exception as e: print(e)
print(open("gpurun_out/r01f_ms.json").read()[:3000])
PY
