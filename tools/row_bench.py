"""Row-wise kernels of one layer at the cfg2 shape, timed alone: persistent projections (with and without
the LayerNorm epilogue), the fused MLP; MMPFN_ROWGEMM=0 selects the old one-tile-per-CTA GEMM."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from multimodalpfn_b200 import _lib

dev = torch.device("cuda")
peaks = bench.load_peaks()
for (S, T) in ((2000, 27), (2300, 27), (2000, 20)):
    r = bench.kernel_breakdown(torch, _lib, dev, S, T, 4, peaks)
    for k, v in r.items():
        if isinstance(v, dict):
            print(f"ROWGEMM={os.environ.get('MMPFN_ROWGEMM', '1')} S={S} T={T} {k:30s} {v['ms'] * 1e3:8.1f} us  "
                  f"{v['gbs']:7.0f} GB/s ({100 * v['frac_hbm']:.0f}% of HBM)" + (f"  {v['tflops']:6.0f} TFLOP/s" if 'tflops' in v else ""))

# fused QKV projection + feature attention (kernels_featfused.cu) against the two kernels it replaces
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
for (S, T) in ((2000, 27), (2300, 27), (2000, 20)):
    rows = 4 * S
    M = rows * T
    x = torch.randn(M, 192, device=dev).to(torch.bfloat16)
    w = (torch.randn(576, 192, device=dev) * 0.1).to(torch.bfloat16)
    att = torch.empty(M, 192, dtype=torch.bfloat16, device=dev)
    ms = bench._time_kernel(torch, lambda: _lib.check(lib.mmpfn_feature_qkv_attention_bf16(
        x.data_ptr(), w.data_ptr(), rows, T, att.data_ptr(), st), "fused"))
    nbytes = M * 192 * 2 * 2
    print(f"S={S} T={T} feature_qkv_attention_fused {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.0f} GB/s algorithmic "
          f"({nbytes / ms / 1e6 / peaks['hbm'] * 100:.0f}% of HBM)", flush=True)
