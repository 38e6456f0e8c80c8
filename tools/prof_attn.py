"""Run the item-attention kernel and one LN-epilogue GEMM a few times at the cfg2 shape (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalpfn_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda")
B, T, n = 4, 27, 2000
pad = (n + 63) // 64 * 64
planes = B * T * 6
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn(planes, pad, 32, device=dev, generator=g).to(torch.bfloat16)
k = torch.randn(planes, pad, 32, device=dev, generator=g).to(torch.bfloat16)
vt = torch.randn(planes, 32, pad, device=dev, generator=g).to(torch.bfloat16)
out = torch.empty(B, n, T, 192, device=dev, dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(lib.mmpfn_item_attention_bf16(q.data_ptr(), k.data_ptr(), vt.data_ptr(), B, T, n, pad, n, pad, 0, out.data_ptr(), st), "attn")
M = B * 2300 * T
A = torch.randn(M, 192, device=dev, generator=g).to(torch.bfloat16)
W = torch.randn(768, 192, device=dev, generator=g).to(torch.bfloat16)
O = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    _lib.check(lib.mmpfn_linear_bf16(A.data_ptr(), W.data_ptr(), M, 768, 192, 1, O.data_ptr(), st), "gemm")
torch.cuda.synchronize()
print("ok")
