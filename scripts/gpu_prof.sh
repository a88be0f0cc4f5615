mkdir -p gpurun_out
TAG=${1:-x}
timeout 300 python -m pytest tests/test_field_gpu.py -x -q > gpurun_out/pytest_field.log 2>&1; echo "pytest field rc=$?"; tail -5 gpurun_out/pytest_field.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_fused -s 3 -c 1 -o gpurun_out/prof_render_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
