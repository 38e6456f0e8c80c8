"""SURVEY.md section 7 "to probe first on the GPU box": six facts about the UNMODIFIED reference (oracle/_ref
snapshot) on a B200, printed as JSON lines.  Run: python tools/ref_probe.py > gpurun_out/ref_probe.jsonl"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multimodalpfn_b200.synth import Geometry, make_checkpoint_config, make_dataset, make_state_dict
from oracle import ref_compat

dev = torch.device("cuda")


def out(**kw):
    print(json.dumps(kw), flush=True)


# (1) does a 16-bit randn on the CUDA generator follow the fp32 stream?
res = {}
for dt in (torch.bfloat16, torch.float16):
    g = torch.Generator(device=dev); g.manual_seed(7)
    a = torch.randn((26, 48), generator=g, device=dev, dtype=torch.float32)
    g = torch.Generator(device=dev); g.manual_seed(7)
    b = torch.randn((26, 48), generator=g, device=dev, dtype=dt)
    res[str(dt)] = float((a - b.float()).abs().max())
g = torch.Generator(device="cpu"); g.manual_seed(7)
c = torch.randn((26, 48), generator=g, dtype=torch.float32)
g = torch.Generator(device=dev); g.manual_seed(7)
a = torch.randn((26, 48), generator=g, device=dev, dtype=torch.float32)
out(probe=1, what="max |randn fp32 - randn 16-bit| on the CUDA generator, same seed (0 = same stream, rounded)", **res,
    cpu_vs_cuda_fp32=float((a.cpu() - c).abs().max()))

# (4) the string comparison that decides GQA in the reference (multi_head_attention.py:603-609)
cap = torch.cuda.get_device_capability(0)
s = f"{cap[0]}.{cap[1]}"
out(probe=4, capability=s, string_compare_ge_8=(s >= "8"),
    what="USE_TORCH_2_GQA = (capability string >= '8') and torch supports enable_gqa: on sm_100 the string compare is "
         "False, so the reference expands the shared K/V head 6x before SDPA")

# model-level probes on the PAD-UFES shape, T=27 estimator inputs (raw table: F=21 -> use 35 columns like the recipe)
geom = Geometry(mgm_heads=8, cap_heads=8)
sd = make_state_dict(geom, seed=1)
model, _ = ref_compat.load_reference_model(sd, make_checkpoint_config(geom), mgm_heads=8, cap_heads=8)
d = make_dataset("pad_ufes", 0)
X = np.concatenate([d["X_train"], d["X_test"]])
img = np.concatenate([d["img_train"], d["img_test"]])
y = d["y_train"].astype(np.float32)
n_tr, n_cls = len(y), d["n_classes"]


def proba(z):
    z = z[:, :n_cls] / 0.9
    e = np.exp(z - z.max(1, keepdims=True))
    return e / e.sum(1, keepdims=True)


def call(m, device, autocast=None, fp32_noise=False):
    xs = torch.as_tensor(X)[:, None].to(device)
    im = torch.as_tensor(img).to(device)
    ys = torch.as_tensor(y).to(device)
    orig = torch.randn
    if fp32_noise:
        def randn(*a, dtype=None, **k):
            t = orig(*a, dtype=torch.float32, **k)
            return t if dtype is None else t.to(dtype)
        torch.randn = randn
    try:
        with torch.inference_mode(), torch.autocast(device, enabled=autocast is not None, dtype=autocast):
            o = m(None, xs, im, ys, only_return_standard_out=True, categorical_inds=[], single_eval_pos=n_tr)
    finally:
        torch.randn = orig
    return o.squeeze(1).float().cpu().numpy()


with ref_compat.without_diagnostic_loop():
    cpu32 = call(model, "cpu")
    model.to(dev)
    gpu32 = call(model, "cuda")
    # (2) reference CUDA fp32 vs CPU fp32 — NOTE the positional noise comes from different generators; inject the CPU draw
    orig = torch.randn

    def cpu_noise(*a, **k):      # the draw the CPU reference makes (fresh default-seeded CPU generator, fp32)
        t = orig(*a, generator=torch.Generator(device="cpu"), dtype=torch.float32)
        return t.to(k.get("device", "cpu")).to(k.get("dtype") or torch.float32)
    torch.randn = cpu_noise
    try:
        gpu32_same_noise = call(model, "cuda")
        # (5) reference autocast fp16 / bf16 on CUDA vs its own fp32, same positional noise
        gpu16 = call(model, "cuda", autocast=torch.float16)
        gpubf = call(model, "cuda", autocast=torch.bfloat16)
    finally:
        torch.randn = orig
out(probe=2, what="reference fp32 on CUDA vs on CPU, PAD-UFES shape T=19+8 (raw 21-column table), one estimator",
    dp_own_generators=float(np.abs(proba(gpu32) - proba(cpu32)).max()),
    dp_same_positional_noise=float(np.abs(proba(gpu32_same_noise) - proba(cpu32)).max()),
    dlogit_same_positional_noise=float(np.abs(gpu32_same_noise - cpu32).max()))
out(probe=5, what="reference CUDA autocast vs its own CUDA fp32, same positional noise",
    fp16_dp=float(np.abs(proba(gpu16) - proba(gpu32_same_noise)).max()),
    fp16_argmax_agree=float((proba(gpu16).argmax(1) == proba(gpu32_same_noise).argmax(1)).mean()),
    bf16_dp=float(np.abs(proba(gpubf) - proba(gpu32_same_noise)).max()),
    bf16_argmax_agree=float((proba(gpubf).argmax(1) == proba(gpu32_same_noise).argmax(1)).mean()))

# (3) which kernels the reference's item attention launches on sm_100 ([27,6,2000,32] SDPA, fp16 and fp32)
from torch.profiler import ProfilerActivity, profile
for dt in (torch.float16, torch.float32):
    q = torch.randn(27, 6, 2000, 32, device=dev, dtype=dt)
    k = torch.randn(27, 6, 2000, 32, device=dev, dtype=dt)
    v = torch.randn(27, 6, 2000, 32, device=dev, dtype=dt)
    for _ in range(3):
        torch.nn.functional.scaled_dot_product_attention(q, k, v)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            torch.nn.functional.scaled_dot_product_attention(q, k, v)
        torch.cuda.synchronize()
    ks = sorted(((e.key, e.device_time_total / max(e.count, 1)) for e in prof.key_averages() if e.device_time_total > 0),
                key=lambda x: -x[1])
    us = sum(t for _, t in ks)
    fl = 4.0 * 27 * 6 * 2000 * 2000 * 32
    out(probe=3, dtype=str(dt), kernels=[(n[:90], round(t, 1)) for n, t in ks[:4]], us_per_call=round(us, 1),
        tflops=round(fl / us / 1e6, 1), what="F.scaled_dot_product_attention [27,6,2000,32], the reference's item attention at T=27")

# (6) reference GPU predict_proba, with and without the diagnostic loop: measured by bench.py (gpu_reference leg)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
r = bench.gpu_reference_leg(steps=3)
out(probe=6, **r)
