mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_vs_ref_gpu.py tests/test_train_gpu.py tests/test_backend_shim_gpu.py -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scripts/gpu_ab.sh default
