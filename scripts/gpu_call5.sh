mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_field_gpu.py -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
lib=$PWD/instance_nerf_b200/libinerf_b200_dbg.so
INERF_B200_LIB=$lib timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | grep bubbles | tail -2
bash scripts/gpu_ab.sh default _nofh _r12
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_first_hit|k_render_fused" -c 12 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train 2>&1 | grep -A3 "k_first_hit\|k_render_fused" | grep -v "^--" | tail -24
