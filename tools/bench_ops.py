"""Micro-benchmarks of single operators through the C ABI (CUDA events, L2 flushed between reps by cycling buffers).
   python tools/bench_ops.py gemm|conv|attn"""
import math
import sys
import torch
sys.path.insert(0, ".")
from divergen_b200 import ops


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def gemm():
    for M, K, N in [(32768, 2880, 320), (32768, 320, 320), (32768, 320, 960), (8192, 5760, 640), (2048, 11520, 1280),
                    (512, 11520, 1280), (32768, 2880, 1280), (16384, 4096, 4096)]:
        x = torch.randn(M, K, device="cuda").half()
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).half()
        us = timeit(lambda: ops.linear(x, w))
        print(f"gemm M={M} K={K} N={N}: {us:8.1f} us  {2.0 * M * K * N / us / 1e6:8.1f} TFLOP/s", flush=True)


def conv():
    for B, H, C, N in [(8, 64, 320, 320), (8, 32, 640, 640), (8, 16, 1280, 1280), (8, 8, 1280, 1280)]:
        x = torch.randn(B, H, H, C, device="cuda").half()
        w = (torch.randn(N, C, 3, 3, device="cuda") / math.sqrt(9 * C)).half()
        b = torch.zeros(N, device="cuda").half()
        us = timeit(lambda: ops.conv3x3_nhwc(x, w, b))
        print(f"conv B={B} {H}x{H} C={C}->{N} (incl. weight pack): {us:8.1f} us", flush=True)




def attn():
    for B, heads, S, Sk, d in [(8, 8, 4096, 4096, 40), (8, 8, 1024, 1024, 80), (8, 8, 256, 256, 160), (8, 8, 4096, 77, 40), (4, 5, 9216, 9216, 64)]:
        C = heads * d
        q = torch.randn(B, S, C, device="cuda").half()
        k = torch.randn(B, Sk, C, device="cuda").half()
        v = torch.randn(B, Sk, C, device="cuda").half()
        us = timeit(lambda: ops.attention(q, k, v, heads), reps=10)
        print(f"attn B={B} h={heads} S={S}x{Sk} d={d}: {us:8.1f} us  {4.0 * B * heads * S * Sk * d / us / 1e6:7.1f} TFLOP/s  "
              f"{B * heads * S * Sk / us / 1e3 / 148:6.2f} exp/ns/SM", flush=True)


if __name__ == "__main__":
    {"gemm": gemm, "conv": conv, "attn": attn}[sys.argv[1]]()
