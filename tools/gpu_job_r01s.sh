#!/bin/bash
# Round 1s GPU job: index build / search overlap (B2_ICP_OVERLAP) — ICP parity + A/B bench.
mkdir -p gpurun_out
O=gpurun_out/r01s
T0=$(date +%s)
timeout 200 python -m pytest tests/test_gpu_icp.py -q -m gpu > ${O}_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s"; tail -3 ${O}_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for ov in 1 0; do
  B2_ICP_OVERLAP=$ov timeout 150 python bench.py --no-cpu-baseline --no-e2e > ${O}_bench_$ov.json 2> ${O}_bench_$ov.err; echo "bench overlap=$ov rc=$? t=$(( $(date +%s) - T0 ))s"
  python - <<PY
import json
d=json.loads(open("${O}_bench_$ov.json").read().strip().splitlines()[-1]); c=d["config"]
print("  it/s %.3f ms %.2f passes %.1f breakdown %s" % (d["value"], d["ms_per_step"], c["passes_per_step"], c["ms_breakdown"]))
PY
done
