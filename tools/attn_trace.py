"""clock64 timeline of one CTA of the item-attention kernel (MMPFN_ATTN_DBG=64 build variant)."""
import ctypes as C
import os
import sys

os.environ["MMPFN_ATTN_DBG"] = "64"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
buf = np.zeros(4096, dtype=np.int64)
fn = lib.mmpfn_debug_attn_trace
fn.argtypes = [C.c_void_p, C.c_int]
fn.restype = C.c_int
assert fn(buf.ctypes.data, 4096) == 0
t0 = buf[0]
BK = int(os.environ.get("MMPFN_ATTN_BK", "48"))
nkt = min((n + BK - 1) // BK, 40)
print("softmax thread 0: per tile [loop top, s_full ok, S in regs, first-tile max done, sweep + P stored, -, rescale done, p_full arrived] (cycles from the CTA's first stamp; slot 5 is unused)")
for j in range(nkt):
    print(j, [int(x - t0) for x in buf[j * 8:j * 8 + 8]], "tile", int(buf[j * 8 + 7] - buf[j * 8]))
print("mma thread: per tile [-, -, p_full ok, PV (+ S(j+2)) issued and committed]  (slots 0-1 only with P through shared memory)")
for j in range(nkt):
    print(j, [int(x - t0) if x else None for x in buf[2048 + j * 4:2048 + j * 4 + 4]])
