"""FP64 ceilings of the box (MEASURED_PEAKS.json has only HBM and bf16): cuBLAS DGEMM yard-stick, raw
DMMA.8x8x4 issue rate and raw DFMA rate.  Writes gpurun_out/fp64_peaks.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpexp_b200._lib import check, lib  # noqa: E402
from gpexp_b200.device import Device, ptr  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


def main():
    dev = Device.get(0)
    out = {"gpu": torch.cuda.get_device_name(0), "sm_count": torch.cuda.get_device_properties(0).multi_processor_count}
    n = 8192
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    t = timed(lambda: torch.matmul(A, B), reps=5)
    out["cublas_dgemm_8192_tflops"] = 2 * n ** 3 / t / 1e12
    # sustained: back to back for ~3 s
    reps = max(3, int(3.0 / t))
    ts = timed(lambda: [torch.matmul(A, B) for _ in range(reps)], reps=1)
    out["cublas_dgemm_8192_tflops_sustained"] = reps * 2 * n ** 3 / ts / 1e12
    del A, B
    sink = dev.zeros(4)
    iters = 20000
    sms = out["sm_count"]
    t = timed(lambda: check(lib.gpx_bench_dmma(dev.h, iters, ptr(sink), dev.stream)))
    out["dmma_tflops"] = sms * 4 * 8 * iters * 16 * 512 / t / 1e12
    t = timed(lambda: check(lib.gpx_bench_dfma(dev.h, iters, ptr(sink), dev.stream)))
    out["dfma_tflops"] = sms * 4 * 256 * iters * 16 * 2 / t / 1e12
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/fp64_peaks.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
