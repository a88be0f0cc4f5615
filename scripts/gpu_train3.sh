mkdir -p gpurun_out
TAG=${1:-x}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches_train_$TAG.csv python bench.py --workload train --steps 3 --warmup 5 > gpurun_out/train_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python scripts/dev_train_profile.py 2>&1 | cut -c1-180 | head -12
