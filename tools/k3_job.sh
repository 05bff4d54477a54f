for v in base minb8 minb7 tile128 inl32 inl128 coop128 coop1024; do
  B2_LIB_PATH=$PWD/dataset_pipeline_b200/_build/variants/libeth3d_b200_$v.so timeout 120 python tools/k3_bench.py 2>/dev/null | tail -1
done
B2_K3_ORDER=grid timeout 120 python tools/k3_bench.py 2>/dev/null | tail -1
B2_K3_STREAMS=1 timeout 120 python tools/k3_bench.py 2>/dev/null | tail -1
B2_K3_STREAMS=2 timeout 120 python tools/k3_bench.py 2>/dev/null | tail -1
B2_K3_STREAMS=8 timeout 120 python tools/k3_bench.py 2>/dev/null | tail -1
