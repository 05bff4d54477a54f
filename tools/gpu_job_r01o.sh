#!/bin/bash
# Round 1o GPU job: K3 with per-lane neighbour work masks — correspondence parity tests + bench.
mkdir -p gpurun_out
O=gpurun_out/r01o
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_icp.py -q -m gpu > ${O}_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s"; tail -4 ${O}_pytest.log
timeout 200 python bench.py --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$? t=$(( $(date +%s) - T0 ))s"
python - <<PY
import json
d=json.loads(open("${O}_bench.json").read().strip().splitlines()[-1]); c=d["config"]
print("it/s %.3f ms %.2f passes %.1f breakdown %s e2e %.3f k3 avg ms %s frac %.3f" % (d["value"], d["ms_per_step"], c["passes_per_step"], c["ms_breakdown"], d["e2e"]["value"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"]))
PY
