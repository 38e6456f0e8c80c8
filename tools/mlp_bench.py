"""Fused MLP sublayer kernel alone at the cfg2 shape (B=4 estimators x 2300 rows x T=27 tokens):
CUDA-event timing, algorithmic TFLOP/s and GB/s, next to the two-GEMM form."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodalpfn_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 2300 * 27
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(M, 192, device=dev, generator=g)
xb = x.to(torch.bfloat16)
w1 = (torch.randn(768, 192, device=dev, generator=g) / 192 ** 0.5).to(torch.bfloat16)
w2 = (torch.randn(192, 768, device=dev, generator=g) / 768 ** 0.5).to(torch.bfloat16)
hid = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream


def fused():
    _lib.check(lib.mmpfn_mlp_bf16(x.data_ptr(), xb.data_ptr(), w1.data_ptr(), w2.data_ptr(), M, st), "mlp")


def up_only():
    _lib.check(lib.mmpfn_linear_bf16(xb.data_ptr(), w1.data_ptr(), M, 768, 192, 1, hid.data_ptr(), st), "up")


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


ms = timeit(fused)
fl = 2.0 * M * 192 * 768 * 2
by = M * 192 * (2 + 4 + 4 + 2)
print(f"fused mlp M={M}: {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.0f} GB/s (algorithmic)")
ms2 = timeit(up_only)
print(f"up-projection + exact-erf GELU alone (two-GEMM form, first half): {ms2:.4f} ms  {fl / 2 / ms2 / 1e9:.1f} TFLOP/s")
