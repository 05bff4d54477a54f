# K3 A/B runs on one box: tools/k3_job.sh <variant> ...   (variants built by tools/k3_variants.sh; "default" = the in-tree library)
for v in "$@"; do
  if [ "$v" = default ]; then timeout 160 python tools/k3_bench.py 2>/dev/null | tail -1
  else B2_LIB_PATH=$PWD/dataset_pipeline_b200/_build/variants/libeth3d_b200_$v.so timeout 160 python tools/k3_bench.py 2>/dev/null | tail -1; fi
done
