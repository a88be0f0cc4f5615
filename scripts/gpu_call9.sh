mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_vs_ref_gpu.py tests/test_field_gpu.py tests/test_stage1_gpu.py tests/test_backend_shim_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python tests/dev_op_bench.py > gpurun_out/op_bench.log 2>&1; tail -15 gpurun_out/op_bench.log
