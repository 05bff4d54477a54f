#!/bin/bash
# Round 1t GPU job: source-level ncu capture of one warm K3 launch (for the next round's work on the search kernel).
mkdir -p gpurun_out
O=gpurun_out/r01t
T0=$(date +%s)
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_nn_radius1' --launch-skip 3 -c 1 -f -o ${O}_k3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > ${O}_ncu.log 2>&1; echo "ncu rc=$? t=$(( $(date +%s) - T0 ))s"
[ -f ${O}_k3.ncu-rep ] && ncu -i ${O}_k3.ncu-rep --page source --csv > ${O}_k3.source.csv 2>/dev/null
[ -f ${O}_k3.ncu-rep ] && ncu -i ${O}_k3.ncu-rep --page raw --csv > ${O}_k3.raw.csv 2>/dev/null
ls -la ${O}_*; rm -f ${O}_k3.ncu-rep
