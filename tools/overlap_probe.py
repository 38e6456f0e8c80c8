"""Do the two estimator groups of the PAD-UFES step overlap when each runs its layer chain on its own stream?
(the item attention leaves DRAM ~94 % idle, the row-wise sublayers leave the MUFU/tensor pipes idle).

    python tools/overlap_probe.py

Times, with CUDA events around the eager device-resident step: (a) the engine's one-pass-over-all-groups path,
(b) group after group on one stream, (c) the groups on two streams (two model objects: private scratch each),
(d) as (c) with the narrower group on a high-priority stream.  (b)-(d) use the same model objects (each draws its own
positional noise, so they are compared with each other, not with (a)).

Measured (B200, profiles/r02_overlap_probe.txt): 31.88 / 33.35 / 31.38 / 32.12 ms — both kernels families fill all
148 SMs, so there is nothing to overlap; the one-pass path stays."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multimodalpfn_b200.classifier import MMPFNClassifier
from multimodalpfn_b200.model import B200PerFeatureTransformer
from multimodalpfn_b200.preprocessing import transform_all
from multimodalpfn_b200.synth import Geometry, make_dataset, make_state_dict

dev = torch.device("cuda", 0)
geom = Geometry(mgm_heads=8, cap_heads=8)
sd = make_state_dict(geom, seed=1)
d = make_dataset(sys.argv[1] if len(sys.argv) > 1 else "pad_ufes", 0)
clf = MMPFNClassifier(mixer_type="MGM+CAP", mgm_heads=8, cap_heads=8, features_per_group=2, n_estimators=8,
                      model_path=(sd, geom), device="cuda:0", inference_precision="bf16", ignore_pretraining_limits=True,
                      random_state=0).fit(d["X_train"], d["img_train"], d["y_train"])
eng = clf.executor_
staged = eng.stage(transform_all(clf.members_, d["X_test"]), d["img_test"])
models = [eng.model] + [B200PerFeatureTransformer(sd, geom, precision="bf16", seed=0) for _ in eng.groups[1:]]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def group_pass(m, g, Xte, tok_tr, tok_te):
    ctx = m.fit_context(g["X_train"], None, g["y_train"], X_all=torch.cat([g["X_train"], Xte], dim=1), img_tok_train=tok_tr,
                        check=False, label_stats=g["label_stats"], nan_flag=eng.nan_flag)
    return m.predict_with_context(ctx, Xte, None, img_tok_test=tok_te, check=False, nan_flag=eng.nan_flag)


def serial():
    tok_tr, tok_te = eng.train_image_tokens(), eng.model.stem_image(staged["img_test"])
    return [group_pass(m, g, Xte, tok_tr, tok_te) for m, g, Xte in zip(models, eng.groups, staged["X_test"])]


def streams(prio):
    main = torch.cuda.current_stream(dev)
    tok_tr, tok_te = eng.train_image_tokens(), eng.model.stem_image(staged["img_test"])
    ev = torch.cuda.Event()
    ev.record(main)
    outs = []
    for i, (m, g, Xte) in enumerate(zip(models, eng.groups, staged["X_test"])):
        s = side[prio][i]
        s.wait_event(ev)
        with torch.cuda.stream(s):
            outs.append(group_pass(m, g, Xte, tok_tr, tok_te))
            e = torch.cuda.Event()
            e.record(s)
        main.wait_event(e)
    return outs


side = {False: [torch.cuda.Stream(dev) for _ in eng.groups],
        True: [torch.cuda.Stream(dev, priority=-1 if i == 0 else 0) for i in range(len(eng.groups))]}


def timed(fn, n=8):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), out


print("groups:", [(len(g["idx"]), g["F"]) for g in eng.groups])
t_multi, ref = timed(lambda: eng.logits_staged(staged))
order = [i for g in eng.groups for i in g["idx"]]
t_serial, o1 = timed(serial)
t_par, o2 = timed(lambda: streams(False))
t_prio, o3 = timed(lambda: streams(True))
for name, o in (("two streams", o2), ("two streams, priority", o3)):
    print(name, "bit-identical to group after group:", bool(torch.equal(torch.cat(o), torch.cat(o1))))
print(f"one pass over all groups {t_multi:.2f} ms | group after group {t_serial:.2f} ms | two streams {t_par:.2f} ms | "
      f"two streams, narrow group high priority {t_prio:.2f} ms")
