mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -25
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; tail -3 gpurun_out/bench_graph.err; python -c "
import json; j=json.load(open('gpurun_out/bench_graph.json')); print(j['train_step'])"
INERF_NO_GRAPH=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; python -c "
import json; j=json.load(open('gpurun_out/bench_nograph.json')); print(j['train_step'])"
