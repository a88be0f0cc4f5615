#!/bin/bash
# quick session: backward-kernel parity tests + op bench on the default build and on A/B variants.  gpurun -- bash scripts/gpu_quick.sh <tag> [variants...]
tag=$1; shift
out=gpurun_out; mkdir -p $out
timeout -k 10 400 python -m pytest tests/test_train_gpu.py tests/test_occupancy_gpu.py -m gpu -q --timeout=300 2>&1 | tail -15 > $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout -k 10 300 python tests/dev_op_bench.py > $out/${tag}_opbench.log 2>&1
cp $out/op_bench.json $out/${tag}_op_bench.json 2>/dev/null
grep -o '"op": "field_backward[^}]*' $out/${tag}_opbench.log | cut -c1-200
bash scripts/gpu_ab_ops.sh $tag "$@"
grep update_extra_state_ms $out/test_metrics.jsonl | tail -1
