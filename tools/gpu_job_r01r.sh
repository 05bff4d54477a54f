#!/bin/bash
# Round 1r GPU job: final state — full GPU suite, smoke(), contract bench (with CPU baseline), cleaner and normals benches.
mkdir -p gpurun_out
O=gpurun_out/r01r
T0=$(date +%s)
timeout 400 python -m pytest tests -q -m gpu --durations=5 > ${O}_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s"; tail -12 ${O}_pytest.log | cut -c1-250
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - T0 ))s"; tail -2 ${O}_smoke.log
timeout 250 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$? t=$(( $(date +%s) - T0 ))s"
python - <<PY
import json
d=json.loads(open("${O}_bench.json").read().strip().splitlines()[-1]); c=d["config"]
print("it/s %.3f ms %.2f passes %.1f breakdown %s e2e %.3f cpu %.5f (%d cores) roofline %s %.3f | %s %.3f" % (d["value"], d["ms_per_step"], c["passes_per_step"], c["ms_breakdown"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["roofline"]["kernel"][:12], d["roofline"]["frac"], d["roofline_second_kernel"]["kernel"][:16], d["roofline_second_kernel"]["frac"]))
PY
timeout 200 python bench_cleaner.py > ${O}_cleaner.json 2> ${O}_cleaner.err; echo "cleaner rc=$? t=$(( $(date +%s) - T0 ))s"; cut -c1-330 ${O}_cleaner.json; tail -2 ${O}_cleaner.err
timeout 100 python bench_normals.py --no-cpu-baseline > ${O}_normals.json 2> ${O}_normals.err; echo "normals rc=$? t=$(( $(date +%s) - T0 ))s"; cut -c1-200 ${O}_normals.json
