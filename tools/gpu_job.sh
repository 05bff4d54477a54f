#!/bin/bash
# GPU job runner for `gpurun`: named stages, every stage under its own timeout, logs into gpurun_out/<tag>_*.
#   tools/gpu_job.sh <tag> <stage> [<stage> ...]
# stages: icp_tests | all_tests | bench | bench_n <N> | smoke | ncu_launches | ncu_k3 | ncu_k5 | sass
set -u
TAG=$1; shift
mkdir -p gpurun_out
O=gpurun_out/$TAG
T0=$(date +%s)
stamp() { echo "[$1] rc=$2 t=$(( $(date +%s) - T0 ))s"; }
while [ $# -gt 0 ]; do
  S=$1; shift
  case $S in
    icp_tests) timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_icp_dense.py -x -q -m gpu > ${O}_icp_tests.log 2>&1; stamp $S $?; tail -5 ${O}_icp_tests.log ;;
    all_tests) timeout 2400 python -m pytest tests -x -q -m gpu > ${O}_all_tests.log 2>&1; stamp $S $?; tail -5 ${O}_all_tests.log ;;
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; stamp $S $?; tail -2 ${O}_smoke.log ;;
    bench) timeout 1200 python bench.py ${BENCH_ARGS:---steps 5 --warmup 3} > ${O}_bench.json 2> ${O}_bench.err; stamp $S $?; tail -c 1500 ${O}_bench.json; tail -3 ${O}_bench.err ;;
    bench_ref) timeout 1700 python bench.py --impl reference ${BENCH_ARGS:---steps 1 --warmup 0} > ${O}_bench_ref.json 2> ${O}_bench_ref.err; stamp $S $?; tail -c 1500 ${O}_bench_ref.json; tail -3 ${O}_bench_ref.err ;;
    bench_n) N=$1; shift
      timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N ${BENCH_ARGS:---steps 5 --warmup 3} > ${O}_bench_n$N.json 2> ${O}_bench_n$N.err; stamp "$S $N" $?; tail -c 1200 ${O}_bench_n$N.json; tail -3 ${O}_bench_n$N.err ;;
    ncu_launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > ${O}_launches.log 2>&1; stamp $S $? ;;
    ncu_k3) timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_nn_tiles' --launch-skip 60 -c 2 -f -o ${O}_k3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-extra --no-secondary > ${O}_ncu_k3.log 2>&1; stamp $S $?
      [ -f ${O}_k3.ncu-rep ] && ncu -i ${O}_k3.ncu-rep --page raw --csv > ${O}_k3.raw.csv 2>/dev/null
      [ -f ${O}_k3.ncu-rep ] && ncu -i ${O}_k3.ncu-rep --page source --csv > ${O}_k3.source.csv 2>/dev/null; rm -f ${O}_k3.ncu-rep ;;
    ncu_k5) timeout 400 ncu --set full --clock-control none -k regex:'k_accumulate_tma' --launch-skip 5 -c 5 -f -o ${O}_k5 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-extra --no-secondary > ${O}_ncu_k5.log 2>&1; stamp $S $?
      [ -f ${O}_k5.ncu-rep ] && ncu -i ${O}_k5.ncu-rep --page raw --csv > ${O}_k5.raw.csv 2>/dev/null; rm -f ${O}_k5.ncu-rep ;;   # (gpurun_out is capped at 64 MiB)
    ncu_reg_k11) B2_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'kr_jacobians|kr_accumulate_weighted|kr_residual_weights' -c 60 -f -o ${O}_regk python bench_reg.py --images 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > ${O}_ncu_regk.log 2>&1; stamp $S $?
      [ -f ${O}_regk.ncu-rep ] && ncu -i ${O}_regk.ncu-rep --page raw --csv > ${O}_regk.raw.csv 2>/dev/null; rm -f ${O}_regk.ncu-rep ;;
    k3_variants) V=$1; shift; bash tools/k3_job.sh $(echo $V | tr , ' ') > ${O}_k3_variants.txt 2>&1; stamp $S $?; cat ${O}_k3_variants.txt ;;
    dmin_parity) B2_LIB_PATH=$PWD/dataset_pipeline_b200/_build/variants/libeth3d_b200_dmin.so timeout 600 python -m pytest tests/test_gpu_icp_dense.py tests/test_gpu_icp.py -x -q -m gpu > ${O}_dmin_parity.log 2>&1; stamp $S $?; tail -3 ${O}_dmin_parity.log ;;
    k3_dual_ab) (timeout 200 python tools/k3_bench.py 2>/dev/null | tail -1; B2_K3_DUAL=0 timeout 200 python tools/k3_bench.py 2>/dev/null | tail -1) > ${O}_k3_dual_ab.txt; stamp $S $?; cat ${O}_k3_dual_ab.txt ;;
    k3_persist_ab) (for v in 1 2 4 0; do B2_K3_PERSIST=$v timeout 200 python tools/k3_bench.py 2>/dev/null | tail -1; done; B2_K3_PERSIST=1 B2_K3_STREAMS=1 timeout 200 python tools/k3_bench.py 2>/dev/null | tail -1) > ${O}_k3_persist_ab.txt; stamp $S $?; cat ${O}_k3_persist_ab.txt ;;
    micro) ./dataset_pipeline_b200/_build/micro/ffma2_bench > ${O}_micro.txt 2>&1; stamp $S $?; cat ${O}_micro.txt ;;
    ncu_reg_launches) B2_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file ${O}_reg_launches.csv python bench_reg.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > ${O}_reg_launches.log 2>&1; stamp $S $? ;;
    ncu_reg_full) B2_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'kr_jacobians|kr_accumulate_weighted|kr_residual_weights|kr_visibility|kr_raster_small|kr_mask_edges' -c 12 -f -o ${O}_reg python bench_reg.py --images 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > ${O}_ncu_reg.log 2>&1; stamp $S $?
      [ -f ${O}_reg.ncu-rep ] && ncu -i ${O}_reg.ncu-rep --page raw --csv > ${O}_reg.raw.csv 2>/dev/null ;;
    *) echo "unknown stage $S" ;;
  esac
done
ls -la ${O}_* 2>/dev/null | awk '{print $5, $9}'
