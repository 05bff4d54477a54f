#!/bin/bash
# Round 1q GPU job: ncu launch list of the contract bench + full captures of K5 (speculative variants), K3 and K7w.
mkdir -p gpurun_out
O=gpurun_out/r01q
T0=$(date +%s)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > ${O}_launches.log 2>&1; echo "launches rc=$? t=$(( $(date +%s) - T0 ))s"
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:'k_accumulate_tma' -c 5 -f -o ${O}_k5 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > ${O}_ncu_k5.log 2>&1; echo "ncu k5 rc=$? t=$(( $(date +%s) - T0 ))s"
timeout 300 $NCU -k regex:'k_nn_radius1' -c 2 -f -o ${O}_k3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > ${O}_ncu_k3.log 2>&1; echo "ncu k3 rc=$? t=$(( $(date +%s) - T0 ))s"
timeout 300 $NCU -k regex:'kn_knn_wide' -c 1 -f -o ${O}_k7w python tools/knn_wide_probe.py 271 > ${O}_ncu_k7w.log 2>&1; echo "ncu k7w rc=$? t=$(( $(date +%s) - T0 ))s"
for f in ${O}_k5 ${O}_k3 ${O}_k7w; do
  [ -f $f.ncu-rep ] && ncu -i $f.ncu-rep --page raw --csv > $f.raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
rm -f ${O}_k5.ncu-rep ${O}_k3.ncu-rep     # keep the raw CSVs (the reports with sources are large); K7w's report is kept for the source page
