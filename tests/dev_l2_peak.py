"""MEASUREMENT INFRASTRUCTURE: the L2-resident read bandwidth and the random-gather rate of the box ("the
builder must additionally measure an L2-resident read peak on the same box"), the denominators of the gather roofline.
Sweeps the buffer size across the L2 capacity and writes gpurun_out/l2_peak.json (copied to profiles/ by hand).

    python tests/dev_l2_peak.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from instance_nerf_b200 import probe  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    out = {"gpu": torch.cuda.get_device_name(0), "sweep": []}
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks):
        out["measured_peaks_hbm_gbs"] = json.load(open(peaks)).get("hbm_gbs")
    for mb in (8, 16, 32, 53.3, 64, 96, 128, 256):
        r = probe.measure_l2_peaks(dev, table_bytes=int(mb * 1e6), hbm_bytes=0)
        r["buffer_mb"] = mb
        out["sweep"].append(r)
        print(json.dumps(r), flush=True)
    out["table"] = probe.measure_l2_peaks(dev)     # the interleaved fp16 table's size (6 664 784 x 8 B) + an HBM-sized stream
    print(json.dumps(out["table"]), flush=True)
    out["red"] = probe.measure_red_peak(dev)       # scatter rate into the fp32 table gradient (hash-grid backward)
    print(json.dumps(out["red"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "l2_peak.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
