"""bf16 compute mode (BASELINE.json configs[3]; sg4d.precision("bf16") -> sg4d_set_compute_precision(1)): every tensor-core
operand rounded to bf16, one product per k-step, fp32 accumulation and statistics.  Checked against (a) a PyTorch emulation of
exactly that arithmetic (operands rounded to bf16, fp32/fp64 product) at 1e-5 and (b) the fp32-level path / the fp32 oracle at
the looser bound SURVEY.md 8(c) states for this configuration (2e-2 relative)."""
import json
import os

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_linear_bf16_matches_emulation(cuda):
    import sg4d
    from sg4d import dense
    torch.manual_seed(0)
    lin = nn.Linear(512, 256).to(cuda)
    x = torch.randn(700, 512, device=cuda, requires_grad=True)
    with sg4d.precision("bf16"):
        y = dense.linear(x, lin)
        (y * y).sum().backward()
    xb, wb = x.detach().bfloat16().double(), lin.weight.detach().bfloat16().double()
    want = xb @ wb.t() + lin.bias.detach().double()
    assert _rel(y.detach(), want) < 1e-5                      # exactly the bf16-operand / fp32-accumulate product
    full = x.detach().double() @ lin.weight.detach().double().t() + lin.bias.detach().double()
    assert 1e-4 < _rel(y.detach(), full) < 2e-2               # ... and visibly not the fp32-level one
    dy = (2 * y.detach()).bfloat16().double()
    assert _rel(x.grad, dy @ wb) < 1e-5
    assert _rel(lin.weight.grad, dy.t() @ xb) < 1e-5
    assert sg4d._lib.load().sg4d_get_compute_precision() == 0  # the context manager restores the default


def test_model_bf16_close_to_fp32(cuda):
    import sg4d
    from oracle import model_ref, weights
    from sg4d import synthetic
    from sg4d.model import SGPNModelWrapper
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = json.load(open(os.path.join(root, "tests", "golden", "no_gt.json")))
    cfg["MODEL"]["lambda_o"] = 0.1
    sd = weights.synth_state_dict(seed=0)
    batch = synthetic.make_batch(20, 2, n_obj=4, n_points_obj=2048, n_points_rel=2048)
    db = synthetic.to_device(batch, cuda)

    def run(mode):
        m = SGPNModelWrapper(cfg, 12, 15, torch.ones(12), torch.ones(15), [f"r{i}" for i in range(14)] + ["none"])
        m.load_state_dict(sd)
        m.to(cuda).train()
        m.obj_predictor.dropout.eval()
        m.rel_predictor.dropout.eval()
        with sg4d.precision(mode):
            outs = m(db, return_meta_data=True)
            loss = m.loss(outs[0], outs[1], db)
            # gradients through the encoders only (the GCN's BatchNorm1d over 8 / 24 rows amplifies any perturbation by
            # orders of magnitude, tests/test_gpu_model.py::_noise_floor -- it would say nothing about the kernels)
            (outs[2].pow(2).sum() + outs[3].pow(2).sum()).backward()
        g = [m.rel_encoder.backbone.SA_modules[i].mlps[0][0].weight.grad.clone() for i in range(3)]
        return [o.detach() for o in outs[:6]], float(loss), g

    o32, l32, g32 = run("fp32")
    o16, l16, g16 = run("bf16")
    with torch.no_grad():
        want = model_ref.forward(model_ref.clone_state(sd), batch, training=True, dropout=False)
    # encoder features vs the fp32 oracle: fp32 path at 1e-4 (the strict bound), bf16 path at the stated 2e-2 relative
    for i in (2, 3):
        assert float((o32[i].cpu() - want[i]).abs().max()) <= 1e-4
        assert _rel(o16[i].cpu(), want[i]) <= 2e-2, (i, _rel(o16[i].cpu(), want[i]))
        assert _rel(o16[i].cpu(), want[i]) > 1e-5              # it really is the reduced-precision path
    assert abs(l16 - l32) <= 2e-2 * max(1.0, abs(l32))
    # Encoder weight gradients are cancellation-heavy sums over 10^5 grouped rows: at 3xTF32 (2^-22 per operand) they carry a
    # relative error of ~1e-4 (tests/test_gpu_full_size.py); bf16 operands (2^-9) scale that by 2^13, i.e. O(0.1 .. 1) noise
    # on top of the same direction -- what a bf16 tensor-core backward pass produces anywhere.  Checked: cosine >= 0.9.
    for a, b in zip(g16, g32):
        cos = float((a.double() * b.double()).sum() / (a.double().norm() * b.double().norm()))
        assert cos >= 0.9, cos
