"""TF32 (and bf16 cross-check) dense matmul peak of this GPU, measured the way MEASURED_PEAKS.json measured bf16:
torch.matmul 8192^3 (2*N^3 FLOP), best of 10 (burst) and back to back for 4 s (sustained), CUDA events.
    python tools/measure_peaks.py > profiles/<tag>_tf32_peak.json
bench.py reads profiles/tf32_peak.json (copy the result there) to quote TF32-kernel fractions against a TF32 roof."""
import json
import time

import torch


def measure(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    flop = 2.0 * n ** 3
    for _ in range(3):
        a @ b
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record()
        torch.cuda.synchronize()
        best = max(best, flop / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); reps = 0
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            a @ b
        reps += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return round(best, 1), round(flop * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12, 1)


if __name__ == "__main__":
    tf_b, tf_s = measure(torch.float32, True)
    bf_b, bf_s = measure(torch.bfloat16, False)
    f32_b, f32_s = measure(torch.float32, False)
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "tf32_tflops": tf_b, "tf32_tflops_sustained": tf_s, "bf16_tflops": bf_b,
                      "bf16_tflops_sustained": bf_s, "fp32_simt_tflops": f32_b, "fp32_simt_tflops_sustained": f32_s,
                      "how": "torch.matmul 8192^3, best of 10 (burst) / back to back 4 s (sustained), CUDA events; TF32 = allow_tf32 cuBLAS"}))
