mkdir -p gpurun_out
N=${1:-2}
TAG=${2:-x}
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "render N=$N rc=$?"; cat gpurun_out/bench_n${N}_$TAG.json; tail -3 gpurun_out/bench_n${N}_$TAG.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload train --rays 65536 --steps 10 --warmup 5 > gpurun_out/train_n${N}_$TAG.json 2> gpurun_out/train_n${N}_$TAG.err; echo "train N=$N rc=$?"; cat gpurun_out/train_n${N}_$TAG.json; tail -3 gpurun_out/train_n${N}_$TAG.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 | tail -1 | cut -c1-300
