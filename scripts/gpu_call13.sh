mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_vs_ref_gpu.py tests/test_train_gpu.py tests/test_backend_shim_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python tests/dev_op_bench.py 2>&1 | grep "march" | cut -c1-330
