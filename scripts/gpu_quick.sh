mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_field_gpu.py -x -q > gpurun_out/pytest_field.log 2>&1; echo "pytest field rc=$?"; tail -15 gpurun_out/pytest_field.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
