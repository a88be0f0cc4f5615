#!/bin/bash
# One GPU session: tests, probes, bench, op bench, backward-kernel A/B.  Usage: gpurun -- bash scripts/gpu_call.sh <tag>
tag=${1:-c}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/${tag}_smi.txt 2>&1
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -150 > $out/${tag}_pytest.log
timeout -k 10 300 python tests/dev_l2_peak.py > $out/${tag}_l2.log 2>&1
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout -k 10 400 python tests/dev_op_bench.py > $out/${tag}_opbench.log 2>&1
cp $out/op_bench.json $out/${tag}_op_bench.json 2>/dev/null
# A/B builds of the render kernel: the bench's render figure only
for v in nowd; do
  if [ -f instance_nerf_b200/libinerf_b200_$v.so ]; then
    INERF_B200_LIB=$PWD/instance_nerf_b200/libinerf_b200_$v.so timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline --no-ref-cuda \
        > $out/${tag}_bench_$v.json 2> $out/${tag}_bench_$v.err
  fi
done
timeout -k 10 300 python bench.py --workload c4 > $out/${tag}_c4_n1.json 2> $out/${tag}_c4_n1.err
timeout -k 10 300 python bench.py --workload train --rays 65536 --steps 30 --warmup 5 > $out/${tag}_train_c5_n1.json 2> $out/${tag}_train_c5_n1.err
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
tail -5 $out/${tag}_pytest.log
head -c 1500 $out/${tag}_bench.json
