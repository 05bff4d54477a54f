#!/bin/bash
# A/B builds of K3: tools/k3_variants.sh name "-DB2_K3_MINB=8 ..."  ->  dataset_pipeline_b200/_build/variants/libeth3d_b200_<name>.so
set -e
cd "$(dirname "$0")/../dataset_pipeline_b200/csrc"
mkdir -p ../_build/variants
NAME=$1; shift
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-O2,-ffp-contract=off --expt-relaxed-constexpr -Wno-deprecated-gpu-targets "$@" -c -o ../_build/variants/b2_icp_$NAME.o b2_icp.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../_build/variants/libeth3d_b200_$NAME.so ../_build/variants/b2_icp_$NAME.o $(ls ../_build/obj/*.o | grep -v b2_icp.o) -lcudart -ldl
rm -f ../_build/variants/b2_icp_$NAME.o
