"""torchrun --nproc-per-node N tools/check_dist.py — sharded engine (per-layer K/V all-gather over NCCL, test pass
reading the gather buffer in place) vs the unsharded engine on every rank's own test chunk: bit-identical logits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from multimodalpfn_b200.classifier import MMPFNClassifier
from multimodalpfn_b200.dist import ShardedEngine
from multimodalpfn_b200.engine import proba_from_logits
from multimodalpfn_b200.synth import Geometry, make_dataset, make_state_dict

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
geom = Geometry(mgm_heads=2, cap_heads=4)
sd = make_state_dict(geom, seed=1)
d = make_dataset("pad_ufes_small", 0)
rng = np.random.default_rng(5 + rank)
Xte = d["X_test"][rng.permutation(len(d["X_test"]))]
img = rng.standard_normal(d["img_test"].shape).astype(np.float32)
for precision in ("bf16", "fp32"):
    clf = MMPFNClassifier(mixer_type="MGM+CAP", mgm_heads=2, cap_heads=4, n_estimators=8, model_path=(sd, geom),
                          device=f"cuda:{local}", inference_precision=precision, ignore_pretraining_limits=True,
                          random_state=0).fit(d["X_train"], d["img_train"], d["y_train"])
    X_tests = [m.transform(Xte) for m in clf.members_]
    ref = clf.executor_.logits(X_tests, img, graph=False).clone()
    sh = ShardedEngine(clf.executor_, rank, world)
    got = sh.logits(X_tests, img).clone()
    got2 = sh.logits(X_tests, img)                    # second call: the gather buffer is reused
    err = (got - ref).abs().max().item()
    perms = [m.class_perm for m in clf.members_]
    p = torch.as_tensor(proba_from_logits(got, perms, n_classes=clf.n_classes_)).cuda()
    allp = sh.proba_gathered(got, perms, n_classes=clf.n_classes_)
    ok = allp.shape[0] == world * p.shape[0] and torch.allclose(allp[rank * p.shape[0]:(rank + 1) * p.shape[0]], p, atol=1e-6)
    print(f"rank {rank}/{world} {precision} mode={sh.plan.mode}: max |sharded - unsharded| logits = {err:.3e}; "
          f"repeat equal = {bool(torch.equal(got, got2))}; gather ok = {bool(ok)}; exchange = {sh.exchange}", flush=True)
    assert err == 0.0 if sh.plan.mode != "broadcast" and precision == "bf16" else err < 1e-4, err
    assert ok and torch.equal(got, got2)

# ---- train rows sharded over the ranks (shard="rows"): bit-identical to the unsharded engine as well ----------------
name, n_est = ("pad_ufes", 8) if world <= 4 else ("img_text_10k", 2)
d = make_dataset(name, 0)
n_te = 120
Xte = d["X_test"][rng.permutation(len(d["X_test"]))[:n_te]]
img = rng.standard_normal((n_te,) + d["img_test"].shape[1:]).astype(np.float32)
for n_members in (n_est, 1):
    clf = MMPFNClassifier(mixer_type="MGM+CAP", mgm_heads=2, cap_heads=4, n_estimators=n_members, model_path=(sd, geom),
                          device=f"cuda:{local}", inference_precision="bf16", ignore_pretraining_limits=True,
                          random_state=0).fit(d["X_train"], d["img_train"], d["y_train"])
    X_tests = [m.transform(Xte) for m in clf.members_]
    ref = clf.executor_.logits(X_tests, img, graph=False).clone()
    sh = ShardedEngine(clf.executor_, rank, world, shard="rows")
    got = sh.logits(X_tests, img).clone()
    got2 = sh.logits(X_tests, img)
    err = (got - ref).abs().max().item()
    print(f"rank {rank}/{world} rows mode, {name}, {n_members} estimator(s), {sh.seg_rows} rows per rank: "
          f"max |sharded - unsharded| logits = {err:.3e}; repeat equal = {bool(torch.equal(got, got2))}; "
          f"exchange = {sh.exchange}", flush=True)
    assert err == 0.0 and torch.equal(got, got2), err
    del clf, sh
    torch.cuda.empty_cache()
dist.barrier()
dist.destroy_process_group()
