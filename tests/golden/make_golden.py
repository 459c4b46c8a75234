"""Generates the committed golden fixtures by running the REFERENCE's own Python
(/root/reference, imported unmodified through oracle/ref_harness.py) on top of the C restatement of
its kernels (oracle/pn2_oracle.c registered as pointnet2_ops._ext).  Container-only:

    python tests/golden/make_golden.py

Outputs (tests/golden/):
  state_dict_keys.json   key -> shape of the reference SGPNModelWrapper state_dict (no_gt config) and
                         the extra keys of the image config that belong to the hot path
  ops_small.npz          reference pointnet2_utils operator outputs on adversarial small clouds
  sa_msg.npz             one reference PointnetSAModuleMSG forward/backward (train-mode BN)
  model_cfg1.npz         reference SGPNModelWrapper forward+loss+backward on a BASELINE config-1
                         shaped scene (4 objects, 6 edges, 2048 points), train-mode BN, dropout off,
                         plus an eval-mode forward
  fp_module.npz          reference three_nn / three_interpolate (pointnet2_utils.py:104-195) and one reference
                         PointnetFPModule forward/backward (pointnet2_modules.py:149-209); `--only fp` regenerates
                         just this file
Weights are never stored: both sides rebuild them with oracle/weights.synth_state_dict.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pn2_ext_cpu as ext, ref_harness as rh, weights  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def adversarial_clouds(seed, b, n):
    """Random clouds with the cases the data pipeline produces: duplicates, zero rows, tiny norms."""
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(b, n, 3, generator=g) * 2 - 1
    xyz[0, n // 2:] = xyz[0, : n - n // 2]                  # exact duplicates -> FPS ties
    if b > 1:
        xyz[1, torch.randperm(n, generator=g)[: n // 8]] = 0.0   # zero rows -> 1e-3 skip rule
        xyz[1, 0] = 0.0                                          # the seed point itself skipped
    if b > 2:
        xyz[2] *= 0.02                                      # every point inside every ball; most |p|^2 < 1e-3
    if b > 3:
        xyz[3] = (xyz[3] * 4).round() / 4                   # lattice: masses of equal distances
    return xyz.contiguous()


def fp_fixture(U, M):
    """feature propagation: the reference's own Python wrappers + FP module on top of the C restatement"""
    g = torch.Generator().manual_seed(31)
    unknown, known = torch.rand(3, 300, 3, generator=g), torch.rand(3, 70, 3, generator=g)
    known[0, 9] = known[0, 4]                                # exact distance ties
    dist, idx = U.three_nn(unknown, known)
    rec = 1.0 / (dist + 1e-8)
    weight = rec / torch.sum(rec, dim=2, keepdim=True)
    kf = torch.randn(3, 6, 70, generator=g).requires_grad_(True)
    interp = U.three_interpolate(kf, idx, weight)
    wi = torch.randn(interp.shape, generator=g)
    (interp * wi).sum().backward()
    fix = {"unknown": unknown, "known": known, "dist": dist, "idx": idx, "known_feats": kf.detach(), "weight": weight,
           "interp": interp.detach(), "w_interp": wi, "d_known_feats": kf.grad.clone()}
    torch.manual_seed(7)
    fp = M.PointnetFPModule(mlp=[6 + 4, 16, 8])
    shapes = {k: list(v.shape) for k, v in fp.state_dict().items()}
    fp.load_state_dict(weights.synth_state_dict(shapes, seed=9))
    fp.train()
    uf = torch.randn(3, 4, 300, generator=g)
    kf2 = kf.detach().clone().requires_grad_(True)
    out = fp(unknown, known, uf, kf2)
    wo = torch.randn(out.shape, generator=g)
    (out * wo).sum().backward()
    fix.update({"fp_shapes": json.dumps(shapes), "unknown_feats": uf, "fp_out": out.detach(), "w_out": wo,
                "fp_d_known_feats": kf2.grad})
    for k, p in fp.named_parameters():
        fix["fp_grad." + k] = p.grad
    np.savez_compressed(os.path.join(OUT, "fp_module.npz"), **{k: np.asarray(v) for k, v in fix.items()})


def main():
    rh.install(ext)
    U, M = rh.ref_utils(), rh.ref_modules()
    fp_fixture(U, M)
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "fp":
        print("fp_module.npz written to", OUT)
        return

    # ---- state_dict layout
    model = rh.build_ref_model()
    shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
    image_extra = {"full_image_feature_reduction.weight": [128, 2048], "full_image_feature_reduction.bias": [128],
                   "rel_predictor.fc3.weight": [15, 1036]}
    with open(os.path.join(OUT, "state_dict_keys.json"), "w") as f:
        json.dump({"no_gt": shapes, "image_extra": image_extra}, f, indent=0)

    # ---- operator level (reference pointnet2_utils.py on the C kernels)
    ops = {}
    for tag, (b, n, m, r, ns) in {"a": (4, 700, 96, 0.35, 16), "b": (3, 64, 16, 0.5, 8), "c": (2, 2048, 512, 0.1, 16)}.items():
        xyz = adversarial_clouds(11 + len(tag) + n, b, n)
        fps = U.furthest_point_sample(xyz, m)
        new_xyz = U.gather_operation(xyz.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
        bq = U.ball_query(r, ns, xyz, new_xyz)
        feats = torch.randn(b, 5, n, generator=torch.Generator().manual_seed(n)).requires_grad_(True)
        qg = U.QueryAndGroup(r, ns)(xyz, new_xyz, feats)
        w = torch.randn(qg.shape, generator=torch.Generator().manual_seed(n + 1))
        (qg * w).sum().backward()
        ops.update({f"{tag}_xyz": xyz, f"{tag}_m": np.int64(m), f"{tag}_r": np.float32(r), f"{tag}_ns": np.int64(ns),
                    f"{tag}_fps": fps, f"{tag}_new_xyz": new_xyz, f"{tag}_ball": bq, f"{tag}_feats": feats.detach(),
                    f"{tag}_qg": qg.detach(), f"{tag}_w": w, f"{tag}_dfeats": feats.grad})
    np.savez_compressed(os.path.join(OUT, "ops_small.npz"), **{k: np.asarray(v) for k, v in ops.items()})

    # ---- one MSG set-abstraction module
    torch.manual_seed(3)
    sa = M.PointnetSAModuleMSG(npoint=64, radii=[0.25, 0.5], nsamples=[8, 16], mlps=[[5, 16, 24], [5, 16, 32]])
    sa_shapes = {k: list(v.shape) for k, v in sa.state_dict().items()}
    sa.load_state_dict(weights.synth_state_dict(sa_shapes, seed=5))
    sa.train()
    xyz = adversarial_clouds(21, 4, 500)
    feats = torch.randn(4, 5, 500, generator=torch.Generator().manual_seed(22)).requires_grad_(True)
    new_xyz, out = sa(xyz, feats)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(23))
    (out * w).sum().backward()
    fix = {"xyz": xyz, "feats": feats.detach(), "new_xyz": new_xyz, "out": out.detach(), "w": w, "dfeats": feats.grad,
           "shapes": json.dumps(sa_shapes)}
    for k, p in sa.named_parameters():
        fix["grad." + k] = p.grad
    for k, v in sa.state_dict().items():
        if "running" in k or "num_batches" in k:
            fix["after." + k] = v
    np.savez_compressed(os.path.join(OUT, "sa_msg.npz"), **{k: np.asarray(v) for k, v in fix.items()})

    # ---- whole model, BASELINE config 1 shape
    sys.path.insert(0, ROOT)
    import sg4d.synthetic as syn
    batch = syn.make_scene(0, n_obj=4, n_points_obj=2048, n_points_rel=2048)
    model.load_state_dict(weights.synth_state_dict(shapes, seed=0))
    model.train()
    model.obj_predictor.dropout.eval()   # train-mode BatchNorm, dropout off (comparable run to run)
    model.rel_predictor.dropout.eval()
    wo, wr = torch.linspace(0.5, 1.5, 12), torch.linspace(0.2, 2.0, 15)
    model.weights_obj, model.weights_rel = wo, wr
    model.mconfig["lambda_o"] = 0.1      # the reference's 1e-6 would hide the object branch's gradients
    outs = model(batch, return_meta_data=True)
    loss = model.mconfig["lambda_o"] * torch.nn.functional.nll_loss(outs[0], batch["gt_class"], weight=wo) + \
        torch.nn.functional.nll_loss(outs[1], batch["gt_rels"], weight=wr)
    loss.backward()
    fix = {"loss": loss.detach(), "lambda_o": np.float32(0.1), "w_obj": wo, "w_rel": wr}
    for name, t in zip(("obj_cls", "rel_cls", "obj_feature", "rel_feature", "gcn_obj", "gcn_rel"), outs):
        fix[name] = t.detach()
    gnorm = {}
    for k, p in model.named_parameters():
        gnorm[k] = None if p.grad is None else float(p.grad.double().norm())
    fix["grad_norms"] = json.dumps(gnorm)
    for k in ("obj_encoder.backbone.SA_modules.0.mlps.0.0.weight", "rel_encoder.backbone.SA_modules.1.mlps.1.3.weight",
              "obj_encoder.backbone.SA_modules.2.mlps.0.4.bias", "gcn.gconvs.0.nn1.1.weight", "obj_predictor.fc3.weight",
              "rel_predictor.fc3.weight"):
        fix["grad." + k] = dict(model.named_parameters())[k].grad
    for k, v in model.state_dict().items():
        if k.endswith("SA_modules.0.mlps.1.4.running_var") or k.endswith("SA_modules.2.mlps.0.1.running_mean") \
                or k.endswith("SA_modules.1.mlps.0.1.num_batches_tracked"):
            fix["after." + k] = v
    model.eval()
    with torch.no_grad():
        eo = model(batch, return_meta_data=True)
    fix["eval_obj_cls"], fix["eval_rel_cls"] = eo[0], eo[1]
    np.savez_compressed(os.path.join(OUT, "model_cfg1.npz"), **{k: np.asarray(v) for k, v in fix.items()})
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
