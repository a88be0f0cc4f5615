# A/B of library variants on the same box: usage gpu_ab.sh <variant-suffix>...   ("default" = the shipped build)
mkdir -p gpurun_out
for v in "$@"; do
  lib=$PWD/instance_nerf_b200/libinerf_b200${v}.so
  [ "$v" = "default" ] && lib=$PWD/instance_nerf_b200/libinerf_b200.so
  echo "== variant $v"
  INERF_B200_LIB=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${v}.json 2> gpurun_out/ab_${v}.err || tail -5 gpurun_out/ab_${v}.err
  python - <<PY
import json
j=json.load(open("gpurun_out/ab_${v}.json"))
spr=j["config"]["samples_per_ray"]; k=j["roofline"]["kernel_ms"]
print("   fill %.3f ms/frame %.3f  kernel_ms %.3f  samples/ray %.1f  ns/ksample %.2f  Mrays/s %.2f  frac %.3f  train_ms %.3f (median %.3f)" % (j["config"].get("tile_fill",0), j["ms_per_step"], k, spr, k*1e6/(spr*307200)*1e3/1e3, j["value"], j["roofline"]["frac"], j["train_step"]["ms"], j["train_step"]["ms_median"]))
PY
done
