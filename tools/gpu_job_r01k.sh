#!/bin/bash
# Round 1k GPU job: speculative LM trials in K5 (B2_LM_SPEC) — parity tests and A/B bench.
mkdir -p gpurun_out
O=gpurun_out/r01k
T0=$(date +%s)
for spec in "1,3" "3,3" "0,0"; do
  B2_LM_SPEC=$spec timeout 300 python -m pytest tests/test_gpu_icp.py -q -m gpu -x > ${O}_pytest_${spec/,/_}.log 2>&1; echo "pytest spec=$spec rc=$? t=$(( $(date +%s) - T0 ))s"; tail -3 ${O}_pytest_${spec/,/_}.log
done
for spec in "0,0" "1,3" "2,3" "3,3" "1,2" "0,3"; do
  B2_LM_SPEC=$spec timeout 200 python bench.py --no-cpu-baseline --no-e2e > ${O}_bench_${spec/,/_}.json 2> ${O}_bench_${spec/,/_}.err; echo "bench spec=$spec rc=$? t=$(( $(date +%s) - T0 ))s"
  python - <<PY
import json
try:
    d=json.loads(open("${O}_bench_${spec/,/_}.json").read().strip().splitlines()[-1]); c=d["config"]
    print("  it/s %.3f ms %.2f passes %.1f tries %.1f inner %s k5 avg ms %s" % (d["value"], d["ms_per_step"], c["passes_per_step"], c["lm_tries"], c["ms_breakdown"], d.get("roofline_second_kernel",{}).get("avg_launch_ms")))
except Exception as e: print("  ", e)
PY
done
