"""Do the layer kernels stay correct when two of them run at the same time on different streams (different buffers)?
Every pair of {feature QKV+attention fused, QKV projection, item QKV scatter, item attention (train / test shape),
out-projection + LayerNorm, fused MLP}: outputs of a concurrent run against the outputs of each kernel run alone."""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodalpfn_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda")
E, HID = 192, 768


def make_ops(B, T, n, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    M = B * n * T
    pad = (n + 63) // 64 * 64
    planes = B * T * 6
    A = torch.randn(M, E, device=dev, generator=g).to(torch.bfloat16)
    W3 = (torch.randn(3 * E, E, device=dev, generator=g) / E ** 0.5).to(torch.bfloat16)
    Wo = (torch.randn(E, E, device=dev, generator=g) / E ** 0.5).to(torch.bfloat16)
    w1 = (torch.randn(HID, E, device=dev, generator=g) / E ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(E, HID, device=dev, generator=g) / HID ** 0.5).to(torch.bfloat16)
    x0 = torch.randn(M, E, device=dev, generator=g)
    q = torch.randn(planes, pad, 32, device=dev, generator=g).to(torch.bfloat16)
    k = torch.randn(planes, pad, 32, device=dev, generator=g).to(torch.bfloat16)
    vt = torch.randn(planes, 32, pad, device=dev, generator=g).to(torch.bfloat16)
    nq = 300
    qpad = (nq + 63) // 64 * 64
    qs = torch.randn(planes, qpad, 32, device=dev, generator=g).to(torch.bfloat16)
    ops = {}

    def op(name, outs, fn, reset=None):
        ops[name] = (outs, fn, reset or (lambda: None))
    O = torch.empty(M, 3 * E, device=dev, dtype=torch.bfloat16)
    op("qkv_proj", [O], lambda s: lib.mmpfn_linear_bf16(A.data_ptr(), W3.data_ptr(), M, 3 * E, E, 0, O.data_ptr(), s))
    att = torch.empty(M, E, device=dev, dtype=torch.bfloat16)
    op("feat_fused", [att], lambda s: lib.mmpfn_feature_qkv_attention_bf16(A.data_ptr(), W3.data_ptr(), B * n, T, att.data_ptr(), s))
    qo, ko, vo = (torch.zeros(planes * pad * 32, device=dev, dtype=torch.bfloat16) for _ in range(3))
    c0, c1 = (torch.zeros(B * T * pad * 32, device=dev, dtype=torch.bfloat16) for _ in range(2))
    op("item_qkv", [qo, ko, vo, c0, c1], lambda s: lib.mmpfn_item_qkv_bf16(
        A.data_ptr(), W3.data_ptr(), B, n, T, pad, 3, qo.data_ptr(), ko.data_ptr(), vo.data_ptr(), c0.data_ptr(), c1.data_ptr(), s))
    out = torch.empty(B, n, T, E, device=dev, dtype=torch.bfloat16)
    op("item_attn_train", [out], lambda s: lib.mmpfn_item_attention_bf16(
        q.data_ptr(), k.data_ptr(), vt.data_ptr(), B, T, n, pad, n, pad, 0, out.data_ptr(), s))
    outs = torch.empty(B, nq, T, E, device=dev, dtype=torch.bfloat16)
    op("item_attn_test", [outs], lambda s: lib.mmpfn_item_attention_bf16(
        qs.data_ptr(), k.data_ptr(), vt.data_ptr(), B, T, nq, qpad, n, pad, 1, outs.data_ptr(), s))
    x = x0.clone()
    xb = x0.to(torch.bfloat16)

    def reset_x():
        x.copy_(x0)
        xb.copy_(x0)
    op("out_proj_ln", [x, xb], lambda s: lib.mmpfn_linear_ln_bf16(A.data_ptr(), Wo.data_ptr(), M, x.data_ptr(), xb.data_ptr(), s), reset_x)
    y = x0.clone()
    yb = x0.to(torch.bfloat16)

    def reset_y():
        y.copy_(x0)
        yb.copy_(x0)
    op("mlp", [y, yb], lambda s: lib.mmpfn_mlp_bf16(y.data_ptr(), yb.data_ptr(), w1.data_ptr(), w2.data_ptr(), M, s), reset_y)
    return ops


a_ops = make_ops(4, 27, 2000, 1)        # the build's shapes
b_ops = make_ops(8, 20, 300, 2)         # the test pass's shapes (row-wise kernels); attention keys from its own planes
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def alone(ops, name):
    outs, fn, reset = ops[name]
    reset()
    torch.cuda.synchronize()
    _lib.check(fn(torch.cuda.current_stream().cuda_stream), name)
    torch.cuda.synchronize()
    return [o.clone() for o in outs]


ref_a = {n: alone(a_ops, n) for n in a_ops}
ref_b = {n: alone(b_ops, n) for n in b_ops}
bad = []
for na, nb in itertools.product(a_ops, b_ops):
    worst = 0.0
    for rep in range(6):
        (oa, fa, ra), (ob, fb, rb) = a_ops[na], b_ops[nb]
        ra()
        rb()
        torch.cuda.synchronize()
        first, second = ((fa, s1, na), (fb, s2, nb)) if rep % 2 == 0 else ((fb, s2, nb), (fa, s1, na))
        for _ in range(1 if "ln" in na + nb or "mlp" in na + nb else 5):     # (in-place kernels run once per reset)
            _lib.check(first[0](first[1].cuda_stream), first[2])
            _lib.check(second[0](second[1].cuda_stream), second[2])
        torch.cuda.synchronize()
        for got, ref in zip(oa, ref_a[na]):
            worst = max(worst, float((got.float() - ref.float()).abs().max()))
        for got, ref in zip(ob, ref_b[nb]):
            worst = max(worst, float((got.float() - ref.float()).abs().max()))
    flag = "" if worst == 0.0 else "   <-- DIFFERS"
    if worst != 0.0:
        bad.append((na, nb, worst))
    print(f"{na:16s} || {nb:16s}: max |concurrent - alone| = {worst:.3e}{flag}", flush=True)
print("pairs that differ:", bad)

# ---- chains: every stream runs a LAYER-like sequence of launches (many cross-stream alternations per SM) --------------
chain = ["feat_fused", "out_proj_ln", "item_qkv", "item_attn_train", "mlp", "qkv_proj"]
chain_b = ["feat_fused", "out_proj_ln", "item_qkv", "item_attn_test", "mlp", "qkv_proj"]
worst = {}
for rep in range(8):
    for ops, names in ((a_ops, chain), (b_ops, chain_b)):
        for nme in names:
            ops[nme][2]()
    torch.cuda.synchronize()
    for i in range(len(chain)):
        _lib.check(a_ops[chain[i]][1](s1.cuda_stream), chain[i])
        _lib.check(b_ops[chain_b[i]][1](s2.cuda_stream), chain_b[i])
        if rep % 2:      # a second, third round of the non-in-place kernels keeps both streams busy
            for nme, ops, st_ in ((chain[i], a_ops, s1), (chain_b[i], b_ops, s2)):
                if nme not in ("out_proj_ln", "mlp"):
                    _lib.check(ops[nme][1](st_.cuda_stream), nme)
    torch.cuda.synchronize()
    for ops, names, refs, tag in ((a_ops, chain, ref_a, "A"), (b_ops, chain_b, ref_b, "B")):
        for nme in names:
            for got, ref in zip(ops[nme][0], refs[nme]):
                d = float((got.float() - ref.float()).abs().max())
                worst[(tag, nme)] = max(worst.get((tag, nme), 0.0), d)
print("chains of six launches per stream, 8 repetitions: max |concurrent - alone| per kernel:")
for k, v in worst.items():
    print("  ", k, f"{v:.3e}", "" if v == 0.0 else "  <-- DIFFERS")

# ---- is it concurrency or the predecessor on the stream?  chain A alone on one stream ---------------------------------
for nme in chain:
    a_ops[nme][2]()
torch.cuda.synchronize()
for nme in chain:
    _lib.check(a_ops[nme][1](s1.cuda_stream), nme)
torch.cuda.synchronize()
for nme in chain:
    d = max(float((got.float() - ref.float()).abs().max()) for got, ref in zip(a_ops[nme][0], ref_a[nme]))
    print(f"chain A alone on one stream: {nme:16s} max |chain - alone| = {d:.3e}")
# which rows of the MLP output differ under concurrency
for rep in range(3):
    for ops, names in ((a_ops, chain), (b_ops, chain_b)):
        for nme in names:
            ops[nme][2]()
    torch.cuda.synchronize()
    for i in range(len(chain)):
        _lib.check(a_ops[chain[i]][1](s1.cuda_stream), chain[i])
        _lib.check(b_ops[chain_b[i]][1](s2.cuda_stream), chain_b[i])
    torch.cuda.synchronize()
    for tag, ops, refs in (("A", a_ops, ref_a), ("B", b_ops, ref_b)):
        got, ref = ops["mlp"][0][0], refs["mlp"][0]
        badrows = ((got - ref).abs().amax(1) > 0).nonzero().flatten()
        tiles = torch.unique(badrows // 128)
        print(f"rep {rep} stream {tag}: MLP rows that differ: {badrows.numel()} of {got.shape[0]}; tiles (of 128 rows): {tiles[:12].tolist()} ... "
              f"{tiles.numel()} tiles; tile index mod 148: {torch.unique(tiles % 148)[:10].tolist()}")
        gotb, refb = ops["mlp"][0][1], refs["mlp"][1]
        for tl in tiles[:3].tolist():
            rows = badrows[(badrows // 128) == tl]
            r0 = int(rows[0])
            cols32 = ((got[r0] - ref[r0]).abs() > 0).nonzero().flatten()
            cols16 = ((gotb[r0].float() - refb[r0].float()).abs() > 0).nonzero().flatten()
            print(f"     tile {tl}: rows in tile {[int(r) % 128 for r in rows]}; first bad row: {cols32.numel()} of 192 fp32 columns differ "
                  f"(chunks of 32: {sorted(set((cols32 // 32).tolist()))}), {cols16.numel()} bf16 columns differ; "
                  f"row mean/std got {float(got[r0].mean()):.3f}/{float(got[r0].std()):.3f} ref {float(ref[r0].mean()):.3f}/{float(ref[r0].std()):.3f}; "
                  f"max|d| {float((got[r0]-ref[r0]).abs().max()):.3f}")
