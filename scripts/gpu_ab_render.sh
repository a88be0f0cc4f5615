#!/bin/bash
# A/B of render-kernel variants on one box: parity tests of the fused renderer, then the render-only bench line, per variant.
# usage: gpurun -- bash scripts/gpu_ab_render.sh <tag> default v1 v2 ...   (variants: csrc/Makefile VARIANT=_v1)
tag=$1; shift
out=gpurun_out; mkdir -p $out
for v in "$@"; do
  lib=$PWD/instance_nerf_b200/libinerf_b200_${v}.so
  [ "$v" = "default" ] && lib=$PWD/instance_nerf_b200/libinerf_b200.so
  echo "== variant $v"
  if [ "${v#nog}" = "$v" ]; then
    INERF_B200_LIB=$lib timeout -k 10 300 python -m pytest tests/test_field_gpu.py -m gpu -q -x --timeout=200 -k "render_fused or field_fused" 2>&1 | tail -3 > $out/${tag}_pytest_$v.log
    tail -1 $out/${tag}_pytest_$v.log
  fi
  INERF_B200_LIB=$lib timeout -k 10 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-train > $out/${tag}_ab_${v}.json 2> $out/${tag}_ab_${v}.err || tail -5 $out/${tag}_ab_${v}.err
  python - <<PY
import json
try:
    j=json.load(open("$out/${tag}_ab_${v}.json"))
    spr=j["config"]["samples_per_ray"]; k=j["roofline"]["kernel_ms"]
    print("   fill %.3f ms/frame %.3f  kernel_ms %.3f  samples/ray %.1f  Mrays/s %.2f  e2e %.2f" % (j["config"].get("tile_fill",0), j["ms_per_step"], k, spr, j["value"], j["e2e"]["value"]))
except Exception as e:
    print("   no bench line:", e)
PY
done
