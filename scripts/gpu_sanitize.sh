#!/bin/bash
# compute-sanitizer pass (SURVEY.md section 5 "race detection" row): memcheck on smoke() + the small op / occupancy / loss tests,
# racecheck + synccheck on smoke() (the three fused persistent kernels on a 48x64 frame).  Summaries -> gpurun_out/<tag>_sanitizer_*.txt
tag=${1:-s}
out=gpurun_out
mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout -k 10 500 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_sanitizer_memcheck_smoke.txt 2>&1; echo "exit $?" >> $out/${tag}_sanitizer_memcheck_smoke.txt
timeout -k 10 700 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_occupancy_gpu.py tests/test_get_rays_gpu.py tests/test_projection.py tests/test_evaluate_gpu.py tests/test_provider.py -m gpu -q -x --timeout=600 -k "not full_size and not 480" > $out/${tag}_sanitizer_memcheck_tests.txt 2>&1; echo "exit $?" >> $out/${tag}_sanitizer_memcheck_tests.txt
timeout -k 10 500 $CS --tool racecheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_sanitizer_racecheck_smoke.txt 2>&1; echo "exit $?" >> $out/${tag}_sanitizer_racecheck_smoke.txt
timeout -k 10 500 $CS --tool synccheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_sanitizer_synccheck_smoke.txt 2>&1; echo "exit $?" >> $out/${tag}_sanitizer_synccheck_smoke.txt
for f in $out/${tag}_sanitizer_*.txt; do echo "== $f"; tail -4 $f; done
