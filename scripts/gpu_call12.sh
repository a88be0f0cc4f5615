mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_field_gpu.py -m gpu -x -q -k "extra_state" 2>&1 | tail -12
python - <<'PY'
import sys, time, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from test_field_gpu import build_model
m, _ = build_model(torch.device("cuda:0"), 32)
m.train()
for it in (0, 20):
    m.iter_density = it
    for _ in range(2):
        with torch.autocast("cuda", dtype=torch.float16):
            m.update_extra_state()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        m.iter_density = it
        with torch.autocast("cuda", dtype=torch.float16):
            m.update_extra_state()
    torch.cuda.synchronize()
    print(f"update_extra_state iter_density={it}: {(time.perf_counter()-t0)/3*1e3:.2f} ms")
PY
