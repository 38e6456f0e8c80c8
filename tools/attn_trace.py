"""clock64 timeline of one CTA of the item-attention kernel — tuning build only:
    MMPFN_DEBUG_LIB=1 [MMPFN_ATTN_PP=n] python tools/attn_trace.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MMPFN_DEBUG_LIB"] = "1"
import numpy as np
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
out = torch.zeros(B, n, T, 192, device=dev, dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(lib.mmpfn_item_attention_bf16(q.data_ptr(), k.data_ptr(), vt.data_ptr(), B, T, n, pad, n, pad, 0,
                                             out.data_ptr(), st), "attn")
torch.cuda.synchronize()
buf = np.zeros(8192, dtype=np.int64)
fn = C.CDLL(_lib.lib_path()).mmpfn_debug_attn_trace
fn.argtypes = [C.c_void_p, C.c_int]
fn.restype = C.c_int
assert fn(buf.ctypes.data, 8192) == 0
t0 = buf[0]
nkt = (n + 47) // 48
print("softmax thread 0 per tile: deltas [wait done, S in regs, sweep, P st issued, st waited, arrived, vote] | tile total")
tot = []
for j in range(nkt):
    s = buf[j * 8:j * 8 + 8].copy()
    nxt = buf[(j + 1) * 8] if j + 1 < nkt else s[7]
    d = [int(s[i + 1] - s[i]) for i in range(7)]
    tot.append(int(nxt - s[0]))
    if j < 12 or j >= nkt - 3:
        print(j, int(s[0] - t0), d, "| tile", int(nxt - s[0]))
print("median tile", int(np.median(tot[2:-2])), "cycles")
print("mma thread per tile: [P ready at, issue+commit took] and lag from softmax arrive to P-ready")
for j in list(range(8)) + [nkt - 2, nkt - 1]:
    a, b = buf[4096 + j * 2], buf[4096 + j * 2 + 1]
    print(j, int(a - t0), int(b - a), "lag from the softmax arrive", int(a - buf[j * 8 + 6]))
