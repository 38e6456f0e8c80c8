"""Item-attention kernel alone at the cfg2 train shape: accuracy against a torch fp32 reference on
a few planes, then CUDA-event timing.  Kernel variants exist only in the tuning build
(MMPFN_DEBUG_LIB=1 MMPFN_ATTN_PP=<polynomial pairs of 24>), one process per variant:
    python tools/attn_bench.py [n_q n_kv shared_kv scale B T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodalpfn_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda")
B = int(sys.argv[5]) if len(sys.argv) > 5 else 4
T = int(sys.argv[6]) if len(sys.argv) > 6 else 27
n_q = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_kv = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
shared = int(sys.argv[3]) if len(sys.argv) > 3 else 0
scale = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
qpad, kpad = (n_q + 63) // 64 * 64, (n_kv + 63) // 64 * 64
planes = B * T * 6
kv_planes = B * T if shared else planes
g = torch.Generator(device=dev).manual_seed(0)
q = (torch.randn(planes, qpad, 32, device=dev, generator=g) * scale).to(torch.bfloat16)
k = (torch.randn(kv_planes, kpad, 32, device=dev, generator=g) * scale).to(torch.bfloat16)
vt = torch.randn(kv_planes, 32, kpad, device=dev, generator=g).to(torch.bfloat16)
out = torch.zeros(B, n_q, T, 192, device=dev, dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream


def run():
    _lib.check(lib.mmpfn_item_attention_bf16(q.data_ptr(), k.data_ptr(), vt.data_ptr(), B, T, n_q, qpad, n_kv, kpad,
                                             shared, out.data_ptr(), st), "attn")


run()
torch.cuda.synchronize()
# reference on a few planes: plane = (b*T + t)*6 + h
err = 0.0
for plane in (0, 7, planes // 2 + 1, planes - 1):
    bt, h = divmod(plane, 6)
    b, t = divmod(bt, T)
    kp = bt if shared else plane
    qq = q[plane, :n_q].float()
    kk = k[kp, :n_kv].float()
    vv = vt[kp, :, :n_kv].float().T
    ref = torch.softmax(qq @ kk.T / 32 ** 0.5, dim=-1) @ vv
    got = out[b, :, t, h * 32:(h + 1) * 32].float()
    err = max(err, float((got - ref).abs().max()))
for _ in range(3):
    run()
torch.cuda.synchronize()
a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    run()
bb.record()
torch.cuda.synchronize()
ms = a.elapsed_time(bb) / 20
fl = 4.0 * planes * n_q * n_kv * 32
print(f"pp={os.environ.get('MMPFN_ATTN_PP', 'default')} B={B} T={T} n_q={n_q} n_kv={n_kv} shared={shared} scale={scale} "
      f"max|err|={err:.3e}  {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
