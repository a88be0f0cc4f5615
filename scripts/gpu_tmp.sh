bash scripts/gpu_ab.sh default _noloads _nomlp _neither
timeout 600 python -m pytest tests/test_ops_vs_ref_gpu.py -x -q -k "composite" 2>&1 | tail -5
timeout 600 python tests/dev_op_bench.py 2>&1 | grep -i "composite\|march\|Error\|error" | cut -c1-400
