"""Image/text stem (MGM + CAP, transformer.py:33-88) alone at the head counts the authors swept
(mmpfn/charts/pad_ufes_20.csv: up to mgm_heads=256, cap_heads=24): ms per call for the 300 test rows of a bench
step and for the 2000 train rows projected once in fit.  usage: python tools/stem_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodalpfn_b200.model import B200PerFeatureTransformer
from multimodalpfn_b200.synth import Geometry, make_state_dict

for mgm, cap in ((8, 8), (64, 24), (256, 24)):
    geom = Geometry(mgm_heads=mgm, cap_heads=cap, nlayers=1)
    sd = make_state_dict(geom, seed=1)
    model = B200PerFeatureTransformer(sd, geom, precision="bf16", seed=0)
    del sd
    for rows in (300, 2000):
        img = torch.randn(rows, 1, 768, device="cuda")
        for _ in range(2):
            model.stem_image(img)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            model.stem_image(img)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        fl = rows * mgm * (2 * 768 * 768 + 2 * 384 * 192) + rows * (4 * mgm * 192 * 192 + 10 * cap * 192 * 192)
        print(f"mgm_heads={mgm:3d} cap_heads={cap:2d} rows={rows:4d}: {ms:8.3f} ms  ({fl / ms / 1e9:6.1f} TFLOP/s; MGM gated projection on tcgen05 from 32 heads, the rest fp32 FFMA)", flush=True)
    del model
    torch.cuda.empty_cache()
