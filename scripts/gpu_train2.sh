mkdir -p gpurun_out
TAG=${1:-x}
timeout 600 python -m pytest tests/test_train_gpu.py -x -q > gpurun_out/pytest_train.log 2>&1; echo "pytest train rc=$?"; tail -30 gpurun_out/pytest_train.log
timeout 300 python bench.py --workload train --steps 30 --warmup 5 > gpurun_out/train_$TAG.json 2> gpurun_out/train_$TAG.err; echo "train rc=$?"; cat gpurun_out/train_$TAG.json; tail -5 gpurun_out/train_$TAG.err
