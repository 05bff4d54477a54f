#!/bin/bash
# Round 1c GPU job: parity tests, Path B benches after the K12 rewrite, ncu captures of the Path B and normals kernels.
# Run through gpurun from the repo root; everything lands in gpurun_out/r01c_*.
mkdir -p gpurun_out
O=gpurun_out/r01c
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv > ${O}_gpu.txt 2>&1
nproc >> ${O}_gpu.txt
T0=$(date +%s)
timeout 700 python -m pytest tests -x -q -m gpu --durations=12 > ${O}_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s"
tail -3 ${O}_pytest.log
timeout 300 python bench_reg.py --steps 3 --warmup 2 > ${O}_reg_pinhole.json 2> ${O}_reg_pinhole.err; echo "reg pinhole rc=$? t=$(( $(date +%s) - T0 ))s"
timeout 300 python bench_reg.py --camera benchmark --steps 3 --warmup 2 > ${O}_reg_fisheye.json 2> ${O}_reg_fisheye.err; echo "reg fisheye rc=$? t=$(( $(date +%s) - T0 ))s"
NCU="ncu --set full --clock-control none --profile-from-start off"
timeout 300 $NCU --import-source on -k regex:'kr_jacobians|kr_accumulate' -c 2 -f -o ${O}_reg_pinhole_k11_k12 python bench_reg.py --images 1 --profile > ${O}_ncu1.log 2>&1; echo "ncu1 rc=$? t=$(( $(date +%s) - T0 ))s"
timeout 300 $NCU --import-source on -k regex:'kr_jacobians|kr_residual_weights|kr_accumulate_blocks' -c 3 -f -o ${O}_reg_fisheye_k11_k12b python bench_reg.py --images 1 --camera benchmark --profile > ${O}_ncu2.log 2>&1; echo "ncu2 rc=$? t=$(( $(date +%s) - T0 ))s"
timeout 300 $NCU -k regex:'kr_visibility|kr_compact|kr_neighbors_observed|kr_intensity|kr_color_accumulate|kr_residual_sums' -c 40 -f -o ${O}_reg_pinhole_other python bench_reg.py --images 1 --profile > ${O}_ncu3.log 2>&1; echo "ncu3 rc=$? t=$(( $(date +%s) - T0 ))s"
timeout 300 $NCU --import-source on -k regex:'kn_' -c 8 -f -o ${O}_normals python bench_normals.py --scan-w 5000 --scan-h 2500 --profile > ${O}_ncu4.log 2>&1; echo "ncu4 rc=$? t=$(( $(date +%s) - T0 ))s"
for f in ${O}_reg_pinhole_k11_k12 ${O}_reg_fisheye_k11_k12b ${O}_reg_pinhole_other ${O}_normals; do
  [ -f $f.ncu-rep ] && ncu -i $f.ncu-rep --page raw --csv > $f.raw.csv 2>/dev/null
done
ls -la gpurun_out | head -40
cat ${O}_reg_pinhole.json | head -c 3000; echo
cat ${O}_reg_fisheye.json | head -c 3000; echo
