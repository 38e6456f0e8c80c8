"""One launch of each hot kernel at the cfg2 shape (B=4 estimators, 2000 train rows, T=27), for ncu:
QKV projection, fused feature QKV + attention, item QKV projection + scatter, output projection + LayerNorm, fused MLP,
item attention (5 + 1 launches per round, two rounds)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodalpfn_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda")
B, T, n, E, HID = 4, 27, 2000, 192, 768
M = B * n * T
g = torch.Generator(device=dev).manual_seed(0)
st = torch.cuda.current_stream().cuda_stream
A = torch.randn(M, E, device=dev, generator=g).to(torch.bfloat16)
W = (torch.randn(3 * E, E, device=dev, generator=g) / E ** 0.5).to(torch.bfloat16)
O = torch.empty(M, 3 * E, device=dev, dtype=torch.bfloat16)
x = torch.randn(M, E, device=dev, generator=g)
xb = x.to(torch.bfloat16)
w1 = (torch.randn(HID, E, device=dev, generator=g) / E ** 0.5).to(torch.bfloat16)
w2 = (torch.randn(E, HID, device=dev, generator=g) / HID ** 0.5).to(torch.bfloat16)
pad = (n + 63) // 64 * 64
planes = B * T * 6
q = torch.randn(planes, pad, 32, device=dev, generator=g).to(torch.bfloat16)
k = torch.randn(planes, pad, 32, device=dev, generator=g).to(torch.bfloat16)
vt = torch.randn(planes, 32, pad, device=dev, generator=g).to(torch.bfloat16)
out = torch.empty(B, n, T, E, device=dev, dtype=torch.bfloat16)
kp = torch.empty_like(q)
ctx = [torch.empty(B * T * pad * 32, device=dev, dtype=torch.bfloat16) for _ in range(2)]
for _ in range(2):
    _lib.check(lib.mmpfn_linear_bf16(A.data_ptr(), W.data_ptr(), M, 3 * E, E, 0, O.data_ptr(), st), "qkv")
    _lib.check(lib.mmpfn_feature_qkv_attention_bf16(A.data_ptr(), W.data_ptr(), B * n, T, xb.data_ptr(), st), "feat_fused")
    _lib.check(lib.mmpfn_item_qkv_bf16(A.data_ptr(), W.data_ptr(), B, n, T, pad, 3, q.data_ptr(), kp.data_ptr(),
                                       vt.data_ptr(), ctx[0].data_ptr(), ctx[1].data_ptr(), st), "item_qkv")
    _lib.check(lib.mmpfn_linear_ln_bf16(A.data_ptr(), W.data_ptr(), M, x.data_ptr(), xb.data_ptr(), st), "out_ln")
    _lib.check(lib.mmpfn_mlp_bf16(x.data_ptr(), xb.data_ptr(), w1.data_ptr(), w2.data_ptr(), M, st), "mlp")
    _lib.check(lib.mmpfn_item_attention_bf16(q.data_ptr(), k.data_ptr(), vt.data_ptr(), B, T, n, pad, n, pad, 0,
                                             out.data_ptr(), st), "attn")
torch.cuda.synchronize()
print("ok")
