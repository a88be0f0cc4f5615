mkdir -p gpurun_out
TAG=${1:-x}
timeout 300 python bench.py --workload train --steps 20 --warmup 5 > gpurun_out/train_$TAG.json 2> gpurun_out/train_$TAG.err; echo "train rc=$?"; cat gpurun_out/train_$TAG.json; tail -5 gpurun_out/train_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 600 --csv --log-file gpurun_out/launches_train_$TAG.csv python bench.py --workload train --steps 3 --warmup 5 > gpurun_out/train_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python -m pytest tests/test_backend_shim_gpu.py -x -q 2>&1 | tail -15
