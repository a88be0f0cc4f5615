#!/bin/bash
# Profiling session (1 GPU): launch lists of the render frame and of the c3 / c5 training steps, ncu --set full of the op-level
# kernels north_star names (hash encode fwd/bwd, compositing fwd/bwd), the fused field kernels and the render kernel.
# Reports are summarised ON the GPU box (scripts/ncu_summary.py) and deleted: gpurun_out/ must stay under 64 MiB.
# Usage: gpurun -- bash scripts/gpu_profile.sh <tag>
tag=${1:-p}
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"
# (1) launch lists (device time per launch; cold-cache, serialised: compare SHARES): the NVTX range bench.py opens around its timed region
$NCU --metrics gpu__time_duration.sum --nvtx --nvtx-include "inerf_timed/" --csv --log-file $out/${tag}_launches_render.csv \
    python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-ref-cuda > $out/${tag}_l1.log 2>&1
INERF_NO_GRAPH=1 $NCU --metrics gpu__time_duration.sum --nvtx --nvtx-include "inerf_timed/" --csv --log-file $out/${tag}_launches_train_c3.csv \
    python bench.py --workload train --rays 4096 --steps 3 --warmup 3 > $out/${tag}_l2.log 2>&1
INERF_NO_GRAPH=1 $NCU --metrics gpu__time_duration.sum --nvtx --nvtx-include "inerf_timed/" --csv --log-file $out/${tag}_launches_train_c5.csv \
    python bench.py --workload train --rays 65536 --steps 3 --warmup 3 > $out/${tag}_l3.log 2>&1
# (2) --set full: op-level kernels on the c5-sized stream (two launches per kernel in the script)
$NCU --set full --import-source on -k regex:"k_composite_train_(fwd|bwd)_scan|k_grid_(fwd|bwd)3x2|k_field_backward_mask|k_field_forward_ws|k_march_expand|k_march_count|k_occ_ema|k_occ_pack" \
    -c 40 -o /tmp/${tag}_ops python tests/dev_op_bench.py --iters 1 --warm 1 > $out/${tag}_ops.log 2>&1
python scripts/ncu_summary.py /tmp/${tag}_ops.ncu-rep > $out/${tag}_ops_ncu_full.txt 2>&1
# (3) --set full: the render kernel (one frame, kept: source-level stalls are read offline) and the occupancy sweep
$NCU --set full --import-source on -k regex:"k_render_fused" -s 3 -c 1 -o $out/${tag}_render \
    python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --no-ref-cuda > $out/${tag}_r.log 2>&1
python scripts/ncu_summary.py $out/${tag}_render.ncu-rep > $out/${tag}_render_ncu_full.txt 2>&1
$NCU --set full -k regex:"k_occupancy_density|k_mark_untrained|k_occ_pick|k_occ_compact" -c 6 -o /tmp/${tag}_occ \
    python -m pytest tests/test_occupancy_gpu.py -m gpu -q -k "full_size" > $out/${tag}_o.log 2>&1
python scripts/ncu_summary.py /tmp/${tag}_occ.ncu-rep > $out/${tag}_occ_ncu_full.txt 2>&1
du -sh $out; ls -la $out | grep ${tag}_
