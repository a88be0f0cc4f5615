"""Bandwidth probes for the roofline denominators (measurement helper, not on the product path).

`MEASURED_PEAKS.json` (driver-written) holds the HBM copy bandwidth only; the hash tables of the instance field are
L2-resident by design (53 MB interleaved fp16 table in a 126 MB L2), so the bound of the gather kernels is an L2 / L1TEX
figure that has to be measured on the box.  `measure_l2_peaks` times, with CUDA events on the
current stream, the two kernels of csrc/probe_bw.cu (libinerf_probe.so):

* `l2_stream_gbs`  -- coalesced 16-byte loads over an L2-resident buffer of `table_bytes` (default = the interleaved table);
* `gather_gps`     -- random 8-byte `ld.global.nc` gathers from a table of that size at full occupancy (gathers / s) and
  `gather_gbs` = 8 B x that rate: the rate for the access shape of the hashed levels (one sector per lane, no reuse);
* `hbm_stream_gbs` -- the same streaming kernel over a 2 GiB buffer (sanity check against MEASURED_PEAKS.json).
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
PROBE_PATH = os.path.join(_HERE, "libinerf_probe.so")
_lib = None


def _probe():
    global _lib
    if _lib is None:
        if not os.path.exists(PROBE_PATH):
            raise RuntimeError(f"{PROBE_PATH} is missing: run `make -C instance_nerf_b200/csrc`")
        L = ctypes.CDLL(PROBE_PATH)
        L.inerf_probe_stream.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
        L.inerf_probe_gather.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
        L.inerf_probe_red.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p]
        L.inerf_probe_stream.restype = L.inerf_probe_gather.restype = L.inerf_probe_red.restype = ctypes.c_int
        _lib = L
    return _lib


def _best_ms(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def measure_l2_peaks(device=None, table_bytes: int = 6664784 * 8, hbm_bytes: int = 2 << 30) -> dict:
    dev = torch.device(device if device is not None else "cuda")
    L = _probe()
    st = torch.cuda.current_stream(dev).cuda_stream
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    sink = torch.zeros(4, dtype=torch.int32, device=dev)
    table_bytes -= table_bytes % 16
    buf = torch.randint(0, 2 ** 31 - 1, (table_bytes // 4,), dtype=torch.int32, device=dev)
    out = {"table_bytes": table_bytes, "sms": sms}

    def check(rc):
        if rc != 0:
            raise RuntimeError(f"probe kernel failed: cudaError {rc}")

    # L2-resident streaming read: 20 passes inside one launch, 2 CTAs of 1024 threads per SM
    reps = 20
    ms = _best_ms(lambda: check(L.inerf_probe_stream(buf.data_ptr(), table_bytes, reps, 2 * sms, sink.data_ptr(), st)))
    out["l2_stream_gbs"] = table_bytes * reps / (ms * 1e-3) / 1e9
    # random 8-byte gathers, 2048 threads per SM, 8 loads in flight per thread
    per_thread = 2048
    blocks = 2 * sms
    ms = _best_ms(lambda: check(L.inerf_probe_gather(buf.data_ptr(), table_bytes // 8, per_thread, blocks, sink.data_ptr(), st)))
    gps = blocks * 1024 * per_thread / (ms * 1e-3)
    out["gather_gps"] = gps
    out["gather_gbs"] = gps * 8 / 1e9
    out["gather_sector_gbs"] = gps * 32 / 1e9      # what L2 -> L1 actually moves (one 32-byte sector per gather)
    if hbm_bytes:
        big = torch.empty(hbm_bytes // 4, dtype=torch.int32, device=dev)
        big.fill_(1)
        ms = _best_ms(lambda: check(L.inerf_probe_stream(big.data_ptr(), hbm_bytes, 1, 8 * sms, sink.data_ptr(), st)), reps=3)
        out["hbm_stream_gbs"] = hbm_bytes / (ms * 1e-3) / 1e9
        del big
    return out


def measure_red_peak(device=None, n_entries: int = 6664784) -> dict:
    """Random 8-byte (float2) and 16-byte (float4) REDs into an fp32 [n_entries, 2] table (the hash-table gradient, 53 MB at the
    instance field's size) at full occupancy -> REDs / s: the scatter rate that bounds the hash-grid backward."""
    dev = torch.device(device if device is not None else "cuda")
    L = _probe()
    st = torch.cuda.current_stream(dev).cuda_stream
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    table = torch.zeros(n_entries, 2, dtype=torch.float32, device=dev)
    out = {"n_entries": n_entries, "table_bytes": n_entries * 8}
    per_thread, blocks = 256, 2 * sms
    for vec, key in ((2, "red8_gps"), (4, "red16_gps")):
        def run():
            rc = L.inerf_probe_red(table.data_ptr(), n_entries, per_thread, blocks, vec, st)
            if rc != 0:
                raise RuntimeError(f"probe kernel failed: cudaError {rc}")
        out[key] = blocks * 1024 * per_thread / (_best_ms(run) * 1e-3)
    return out
