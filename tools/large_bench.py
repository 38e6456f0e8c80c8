"""BASELINE.json configs[2] / configs[3] on one GPU, ONE estimator, bf16: context build + test pass timed
with CUDA events (device-resident inputs), against the algorithmic FLOPs of bench.flops_estimator.
usage: python tools/large_bench.py [img_text_10k|large_ctx_50k ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from multimodalpfn_b200.model import B200PerFeatureTransformer
from multimodalpfn_b200.synth import DATASETS, Geometry, make_dataset, make_state_dict

names = sys.argv[1:] or ["img_text_10k", "large_ctx_50k"]
geom = Geometry(mgm_heads=8, cap_heads=8)
sd = make_state_dict(geom, seed=1)
model = B200PerFeatureTransformer(sd, geom, precision="bf16", seed=0)
for name in names:
    d = make_dataset(name, 0)
    has_img = d["img_train"] is not None
    Xtr, Xte = torch.as_tensor(d["X_train"]).cuda(), torch.as_tensor(d["X_test"]).cuda()
    itr = torch.as_tensor(d["img_train"]).cuda() if has_img else None
    ite = torch.as_tensor(d["img_test"]).cuda() if has_img else None
    y = torch.as_tensor(d["y_train"].astype(np.float32)).cuda()
    n_tr, n_te = Xtr.shape[0], Xte.shape[0]
    for rep in range(2):
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        ctx = model.fit_context(Xtr, itr, y, check=False)
        e[1].record()
        lg = model.predict_with_context(ctx, Xte, ite, check=False)
        e[2].record()
        torch.cuda.synchronize()
        t_ctx, t_te = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        T = ctx.T
        del ctx
    fl = bench.flops_estimator(n_tr, n_te, T)
    print(f"{name}: {n_tr} train / {n_te} test rows, T={T}, 1 estimator bf16: context {t_ctx:.1f} ms + test {t_te:.1f} ms = "
          f"{t_ctx + t_te:.1f} ms -> {n_te / (t_ctx + t_te) * 1e3:.0f} test rows/s rebuilt, {n_te / t_te * 1e3:.0f} rows/s cached; "
          f"{fl / 1e12:.1f} TFLOP algorithmic -> {fl / (t_ctx + t_te) / 1e9:.0f} TFLOP/s; "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del lg
    torch.cuda.empty_cache()
