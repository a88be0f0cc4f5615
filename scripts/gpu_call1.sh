mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests/test_field_gpu.py tests/test_ops_vs_ref_gpu.py -m gpu -x -q > gpurun_out/pytest_call1.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_call1.log
bash scripts/gpu_ab.sh default _nopatch _fmul2 _ring16
