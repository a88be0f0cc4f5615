mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2>/dev/null; echo "ref arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_fused -s 3 -c 1 -o gpurun_out/prof_render -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
timeout 600 python tests/dev_op_bench.py > gpurun_out/op_bench.log 2>&1; echo "op bench rc=$?"
timeout 600 python tests/dev_ref_gpu_path.py > gpurun_out/ref_gpu_path.log 2>&1; echo "ref path rc=$?"; tail -3 gpurun_out/ref_gpu_path.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_field_forward_ws|k_field_backward_mask" -s 12 -c 2 -o gpurun_out/prof_train -f python bench.py --workload train --steps 4 --warmup 3 > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"; tail -2 gpurun_out/ncu_train.log
