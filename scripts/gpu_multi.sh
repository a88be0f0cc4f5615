#!/bin/bash
# Multi-GPU session: gpurun --gpus N -- bash scripts/gpu_multi.sh N <tag> [all|dp]
N=${1:-2}
tag=${2:-m}
what=${3:-all}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/${tag}_smi.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -k 10 400 $TR --master-port 29513 bench.py --gpus $N --workload train --rays 65536 --steps 30 --warmup 5 > $out/${tag}_train_c5_n$N.json 2> $out/${tag}_train_c5_n$N.err
if [ "$what" = "all" ] || [ "$what" = "main" ]; then
  timeout -k 10 600 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $out/${tag}_bench_n$N.json 2> $out/${tag}_bench_n$N.err
  timeout -k 10 400 $TR --master-port 29512 bench.py --gpus $N --workload c4 > $out/${tag}_c4_n$N.json 2> $out/${tag}_c4_n$N.err
fi
if [ "$what" = "all" ]; then
  INERF_NO_GRAPH=1 timeout -k 10 400 $TR --master-port 29514 bench.py --gpus $N --workload train --rays 65536 --steps 30 --warmup 5 > $out/${tag}_train_c5_eager_n$N.json 2> $out/${tag}_train_c5_eager_n$N.err
fi
if [ "$N" = "2" ]; then
  timeout -k 10 450 python -m pytest tests/test_dp_gpu.py -m gpu -q --timeout=420 2>&1 | tail -40 > $out/${tag}_dp_pytest.log
  tail -3 $out/${tag}_dp_pytest.log
fi
head -c 600 $out/${tag}_train_c5_n$N.json; echo
[ "$what" != "dp" ] && (head -c 600 $out/${tag}_bench_n$N.json; echo; head -c 400 $out/${tag}_c4_n$N.json; echo)
tail -3 $out/${tag}_train_c5_n$N.err
