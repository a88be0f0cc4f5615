#!/bin/bash
# op bench under A/B builds of libinerf_b200 (csrc/Makefile VARIANT=...): gpurun -- bash scripts/gpu_ab_ops.sh <tag> v1 v2 ...
tag=$1; shift
out=gpurun_out; mkdir -p $out
for v in "$@"; do
  INERF_B200_LIB=$PWD/instance_nerf_b200/libinerf_b200_$v.so timeout -k 10 300 python tests/dev_op_bench.py > $out/${tag}_opbench_$v.log 2>&1
  cp $out/op_bench.json $out/${tag}_op_bench_$v.json 2>/dev/null
  grep -o '"op": "field_backward[^}]*' $out/${tag}_opbench_$v.log | cut -c1-200
done
