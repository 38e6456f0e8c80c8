"""clock64 timeline of the MMA-issuing thread of CTA 0 of the fused MLP kernel (MMPFN_MLP_DBG=1)."""
import ctypes as C
import os
import sys

os.environ["MMPFN_MLP_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multimodalpfn_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda")
M = 4 * 2000 * 27
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(M, 192, device=dev, generator=g)
xb = x.to(torch.bfloat16)
w1 = (torch.randn(768, 192, device=dev, generator=g) / 192 ** 0.5).to(torch.bfloat16)
w2 = (torch.randn(192, 768, device=dev, generator=g) / 768 ** 0.5).to(torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(lib.mmpfn_mlp_bf16(x.data_ptr(), xb.data_ptr(), w1.data_ptr(), w2.data_ptr(), M, st), "mlp")
torch.cuda.synchronize()
buf = np.zeros(4096, dtype=np.int64)
fn = lib.mmpfn_debug_mlp_trace
fn.argtypes = [C.c_void_p, C.c_int]
fn.restype = C.c_int
assert fn(buf.ctypes.data, 4096) == 0
t0 = buf[0]
print("chunk g: gemm1(g) [enter, w1_full ok, h_free ok, issued]  gemm2(g) [enter, w2_full ok, hs_full ok, issued]   waits: w1 h_free w2 hs")
for g_ in range(0, 60):
    s = buf[g_ * 8:g_ * 8 + 8] - t0
    print(g_, [int(v) for v in s], " waits", int(s[1] - s[0]), int(s[2] - s[1]), int(s[5] - s[4]), int(s[6] - s[5]))
