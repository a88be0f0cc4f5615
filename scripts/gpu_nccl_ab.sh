mkdir -p gpurun_out
for c in default 2 8; do
  if [ "$c" = "default" ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$c; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/nccl_$c.json 2> gpurun_out/nccl_$c.err
  python - <<PY
import json
j=json.load(open("gpurun_out/nccl_$c.json"))
print("NCCL_MAX_CTAS=$c", "Mrays/s %.2f ms/step %.3f kernel_ms %.3f e2e %.2f spr %.1f" % (j["value"], j["ms_per_step"], j["roofline"]["kernel_ms"], j["e2e"]["value"], j["config"]["samples_per_ray"]))
PY
done
