"""GPU parity tests of the kernels, called through the C ABI (ctypes -> libsg4d.so), against the CPU
oracle (oracle/pn2_oracle.c) on identical seeded inputs and against the golden fixtures produced by the
reference's own Python.  Index outputs: bit-exact.  Copies: bit-exact.  Sums: <= 1e-5 absolute (the
reference's own atomicAdd order is arbitrary)."""
import os

import numpy as np
import pytest
import torch

from oracle import pn2_ext_cpu as ora

pytestmark = pytest.mark.gpu


def _clouds(seed, b, n, kind="mixed"):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(b, n, 3, generator=g) * 2 - 1
    if kind == "mixed":
        xyz[0, n // 2:] = xyz[0, : n - n // 2].clone()                          # duplicates -> exact FPS ties
        if b > 1:
            xyz[1, torch.randperm(n, generator=g)[: max(1, n // 8)]] = 0.0   # zero rows -> skip rule
            xyz[1, 0] = 0.0
        if b > 2:
            xyz[2] = (xyz[2] * 4).round() / 4                           # lattice: many equal distances
        if b > 3:
            xyz[3] *= 0.02                                              # almost everything skipped
    elif kind == "gauss":
        xyz = torch.randn(b, n, 3, generator=g) * 0.3
        xyz /= xyz.pow(2).sum(2).sqrt().amax(1, keepdim=True).unsqueeze(-1)
    return xyz.contiguous()


def _ext():
    from sg4d.pointnet2_ops import _ext
    return _ext


@pytest.mark.parametrize("b,n,m", [(3, 5, 5), (4, 64, 16), (4, 700, 96), (2, 2048, 512), (4, 4000, 512),
                                   (3, 8000, 512), (2, 13000, 128), (2, 30000, 96), (2, 80000, 512),
                                   (1, 131072, 64), (1, 200001, 24)])
def test_fps_bit_exact(cuda, b, n, m):
    xyz = _clouds(100 + n, b, n)
    want = ora.furthest_point_sampling(xyz, m)
    got = _ext().furthest_point_sampling(xyz.to(cuda), m)
    assert got.dtype == torch.int32 and got.shape == (b, m)
    np.testing.assert_array_equal(got.cpu().numpy(), want.numpy())


def test_fps_all_points_skipped_returns_zeros(cuda):
    xyz = torch.rand(2, 300, 3) * 0.01          # |p|^2 <= 1e-3 everywhere: the reference emits index 0
    got = _ext().furthest_point_sampling(xyz.to(cuda), 7).cpu()
    assert torch.equal(got, ora.furthest_point_sampling(xyz, 7)) and int(got.abs().sum()) == 0


@pytest.mark.parametrize("b,n,m,r,ns", [(4, 700, 96, 0.35, 16), (3, 64, 16, 0.5, 8), (2, 2048, 512, 0.1, 16),
                                        (2, 2048, 512, 0.2, 32), (3, 512, 128, 0.4, 64), (2, 9000, 40, 0.05, 32),
                                        (1, 80000, 512, 0.1, 16)])
def test_ball_query_bit_exact(cuda, b, n, m, r, ns):
    xyz = _clouds(200 + n + ns, b, n, "mixed" if n < 80000 else "gauss")
    fps = ora.furthest_point_sampling(xyz, m)
    new_xyz = ora.gather_points(xyz.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
    want = ora.ball_query(new_xyz, xyz, r, ns)
    got = _ext().ball_query(new_xyz.to(cuda), xyz.to(cuda), r, ns)
    np.testing.assert_array_equal(got.cpu().numpy(), want.numpy())


def test_ball_query_rows_without_hits_stay_zero(cuda):
    xyz = torch.rand(1, 100, 3)
    far = torch.full((1, 4, 3), 9.0)
    got = _ext().ball_query(far.to(cuda), xyz.to(cuda), 0.1, 8).cpu()
    assert int(got.abs().sum()) == 0 and torch.equal(got, ora.ball_query(far, xyz, 0.1, 8))


def test_gather_and_group_match_oracle(cuda):
    g = torch.Generator().manual_seed(5)
    b, c, n, m, ns = 3, 7, 300, 40, 12
    pts = torch.randn(b, c, n, generator=g)
    idx1 = torch.randint(0, n, (b, m), generator=g, dtype=torch.int32)
    idx2 = torch.randint(0, n, (b, m, ns), generator=g, dtype=torch.int32)
    e = _ext()
    assert torch.equal(e.gather_points(pts.to(cuda), idx1.to(cuda)).cpu(), ora.gather_points(pts, idx1))
    assert torch.equal(e.group_points(pts.to(cuda), idx2.to(cuda)).cpu(), ora.group_points(pts, idx2))
    go1 = torch.randn(b, c, m, generator=g)
    go2 = torch.randn(b, c, m, ns, generator=g)
    torch.testing.assert_close(e.gather_points_grad(go1.to(cuda), idx1.to(cuda), n).cpu(),
                               ora.gather_points_grad(go1, idx1, n), rtol=0, atol=1e-5)
    torch.testing.assert_close(e.group_points_grad(go2.to(cuda), idx2.to(cuda), n).cpu(),
                               ora.group_points_grad(go2, idx2, n), rtol=0, atol=1e-5)


def test_operator_api_against_reference_fixture(cuda, golden_dir):
    """pointnet2_utils (autograd wrappers, QueryAndGroup) vs outputs of the REFERENCE's pointnet2_utils"""
    from sg4d.pointnet2_ops import pointnet2_utils as U
    fx = np.load(os.path.join(golden_dir, "ops_small.npz"))
    for tag in "abc":
        xyz = torch.from_numpy(fx[f"{tag}_xyz"]).to(cuda)
        m, r, ns = int(fx[f"{tag}_m"]), float(fx[f"{tag}_r"]), int(fx[f"{tag}_ns"])
        fps = U.furthest_point_sample(xyz, m)
        np.testing.assert_array_equal(fps.cpu().numpy(), fx[f"{tag}_fps"])
        new_xyz = U.gather_operation(xyz.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
        np.testing.assert_array_equal(new_xyz.cpu().numpy(), fx[f"{tag}_new_xyz"])
        np.testing.assert_array_equal(U.ball_query(r, ns, xyz, new_xyz).cpu().numpy(), fx[f"{tag}_ball"])
        feats = torch.from_numpy(fx[f"{tag}_feats"]).to(cuda).requires_grad_(True)
        qg = U.QueryAndGroup(r, ns)(xyz, new_xyz, feats)
        np.testing.assert_array_equal(qg.detach().cpu().numpy(), fx[f"{tag}_qg"])
        (qg * torch.from_numpy(fx[f"{tag}_w"]).to(cuda)).sum().backward()
        np.testing.assert_allclose(feats.grad.cpu().numpy(), fx[f"{tag}_dfeats"], rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------ point-major fused operators

@pytest.mark.parametrize("stride,n,m", [(6, 2048, 512), (7, 5000, 128), (3, 512, 128), (7, 80000, 512)])
def test_fps_rows_equals_fps_on_xyz(cuda, stride, n, m):
    from sg4d import rows
    b = 2
    g = torch.Generator().manual_seed(stride + n)
    pts = torch.rand(b, n, stride, generator=g)
    pts[:, :, :3] = _clouds(n + 1, b, n)
    want = ora.furthest_point_sampling(pts[:, :, :3].contiguous(), m)
    idx, new_xyz = rows.fps_rows(pts.to(cuda), m)
    np.testing.assert_array_equal(idx.cpu().numpy(), want.numpy())
    picked = torch.gather(pts[:, :, :3], 1, want.long().unsqueeze(-1).expand(-1, -1, 3))
    assert torch.equal(new_xyz.cpu(), picked)


@pytest.mark.parametrize("n,m,radii,nss,stride", [(2048, 512, [0.1, 0.2], [16, 32], 6), (512, 128, [0.2, 0.4], [32, 64], 3),
                                                 (3000, 64, [0.3], [8], 7), (1000, 50, [0.1, 0.2, 0.4], [4, 8, 16], 6)])
def test_ball_query_rows_multi_radius(cuda, n, m, radii, nss, stride):
    from sg4d import rows
    b = 3
    pts = torch.rand(b, n, stride, generator=torch.Generator().manual_seed(n))
    pts[:, :, :3] = _clouds(n + 7, b, n)
    xyz = pts[:, :, :3].contiguous()
    fps = ora.furthest_point_sampling(xyz, m)
    new_xyz = ora.gather_points(xyz.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
    idx, cnt = rows.ball_query_rows(new_xyz.to(cuda), pts.to(cuda), radii, nss)
    for s, (r, ns) in enumerate(zip(radii, nss)):
        want = ora.ball_query(new_xyz, xyz, r, ns)
        np.testing.assert_array_equal(idx[s].cpu().numpy(), want.numpy())
        # cnt = number of distinct hits: slots >= cnt repeat slot 0
        w = want.numpy()
        assert (cnt[s].cpu().numpy() <= ns).all() and (cnt[s].cpu().numpy() >= 1).all()
        c = cnt[s].cpu().numpy()
        for bi in range(b):
            for j in range(0, m, 7):
                row = w[bi, j]
                assert (np.diff(row[: c[bi, j]]) > 0).all() and (row[c[bi, j]:] == row[0]).all()


# ------------------------------------------------------------------ spatial index (csrc/spatial.cu)

def _bench_like_clouds(seed, b, n, stride):
    """the benchmark's own cloud generator (Gaussian mixture on the unit sphere, duplicates, zero rows)"""
    from sg4d import synthetic
    g = torch.Generator().manual_seed(seed)
    return torch.stack([synthetic.make_cloud(g, n, stride) for _ in range(b)])


@pytest.mark.parametrize("stride,n,m,kind", [(3, 1024, 128, "mixed"), (6, 1500, 200, "mixed"), (7, 4000, 512, "mixed"),
                                             (6, 8191, 512, "mixed"), (7, 20000, 512, "bench"), (6, 80000, 512, "bench"),
                                             (7, 80000, 512, "mixed"), (3, 200001, 64, "mixed")])
def test_fps_indexed_bit_exact(cuda, stride, n, m, kind):
    """bucket-pruned FPS == the oracle's literal simulation of the reference kernel (ties, skip rule, padding)"""
    from sg4d import rows
    b = 4 if n <= 20000 else 2
    if kind == "bench":
        pts = _bench_like_clouds(n + stride, b, n, stride)
    else:
        pts = torch.rand(b, n, stride, generator=torch.Generator().manual_seed(n))
        pts[:, :, :3] = _clouds(n + 3, b, n)
    want = ora.furthest_point_sampling(pts[:, :, :3].contiguous(), m)
    dpts = pts.to(cuda)
    idx, new_xyz = rows.fps_rows(dpts, m, rows.SpatialIndex(dpts))
    np.testing.assert_array_equal(idx.cpu().numpy(), want.numpy())
    picked = torch.gather(pts[:, :, :3], 1, want.long().unsqueeze(-1).expand(-1, -1, 3))
    assert torch.equal(new_xyz.cpu(), picked)
    idx2, _ = rows.fps_rows(dpts, m, rows.SpatialIndex(dpts))       # the bucket layout depends on atomics; the picks do not
    assert torch.equal(idx, idx2)


def test_fps_indexed_degenerate_clouds(cuda):
    from sg4d import rows
    n, m = 2048, 40
    pts = torch.rand(5, n, 3, generator=torch.Generator().manual_seed(9))
    pts[0] *= 0.01                          # everything skipped by the |p|^2 <= 1e-3 rule: all picks are index 0
    pts[1] = pts[1, :1]                     # one point repeated: all distances tie at 0 after the first pick
    pts[2, :, 1:] = 0.25                    # collinear
    pts[3, 5:] = 0.0                        # five candidates, the rest skipped
    pts[4, 100] = float("nan")              # a NaN point keeps temp = 1e10 forever in the reference
    want = ora.furthest_point_sampling(pts, m)
    got, _ = rows.fps_rows(pts.to(cuda), m, rows.SpatialIndex(pts.to(cuda)))
    np.testing.assert_array_equal(got.cpu().numpy(), want.numpy())


@pytest.mark.parametrize("n,m,radii,nss,stride,prefix,kind", [
    (1024, 128, [0.2, 0.4], [32, 64], 3, 0, "mixed"), (4000, 512, [0.1, 0.2], [16, 32], 6, 0, "mixed"),
    (4000, 512, [0.1, 0.2], [16, 32], 7, 512, "mixed"), (20000, 512, [0.1, 0.2], [16, 32], 7, 4096, "bench"),
    (20000, 200, [0.5], [64], 6, 0, "bench"), (80000, 512, [0.1, 0.2], [16, 32], 7, 4096, "bench"),
    (80000, 512, [0.1, 0.2], [16, 32], 6, 0, "bench"), (30000, 64, [0.05, 0.1, 0.3], [4, 8, 16], 6, 1000, "mixed")])
def test_ball_query_indexed_bit_exact(cuda, n, m, radii, nss, stride, prefix, kind):
    """prefix scan + spatial index == the oracle's serial scan (first nsample hits in index order, first-hit fill,
    zero rows), including dense balls with thousands of hits (buffer compaction) and prefix = 0 (index only)"""
    from sg4d import rows
    b = 2
    if kind == "bench":
        pts = _bench_like_clouds(n + stride + prefix, b, n, stride)
    else:
        pts = torch.rand(b, n, stride, generator=torch.Generator().manual_seed(n))
        pts[:, :, :3] = _clouds(n + 11, b, n)
    xyz = pts[:, :, :3].contiguous()
    fps = ora.furthest_point_sampling(xyz, m)
    new_xyz = ora.gather_points(xyz.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
    new_xyz[0, -1] = 7.0                   # a centre without any neighbour: the row stays zero
    dpts = pts.to(cuda)
    idx, cnt = rows.ball_query_rows(new_xyz.to(cuda), dpts, radii, nss, rows.SpatialIndex(dpts), prefix=prefix)
    for s, (r, ns) in enumerate(zip(radii, nss)):
        want = ora.ball_query(new_xyz, xyz, r, ns)
        np.testing.assert_array_equal(idx[s].cpu().numpy(), want.numpy())
        w = want.numpy()
        distinct = np.array([[len(set(row.tolist())) for row in cl] for cl in w])
        got_cnt = cnt[s].cpu().numpy()
        nohit = (new_xyz.numpy()[:, :, 0] == 7.0)
        np.testing.assert_array_equal(got_cnt[~nohit], distinct[~nohit])
        assert (got_cnt[nohit] == 0).all()


@pytest.mark.parametrize("c,feat_stride,feat_off,n,m,ns", [(3, 6, 3, 2048, 512, 16), (4, 7, 3, 2048, 512, 32),
                                                           (192, 192, 0, 512, 128, 64), (5, 5, 0, 300, 20, 8)])
def test_group_rows_forward_backward(cuda, c, feat_stride, feat_off, n, m, ns):
    from sg4d import rows
    b = 2
    g = torch.Generator().manual_seed(c + n)
    xyz = _clouds(c + n, b, n, "gauss")
    if feat_off:                                  # SA1 style: features live in the same rows as xyz
        pts = torch.rand(b, n, feat_stride, generator=g)
        pts[:, :, :3] = xyz
        feats = pts
    else:                                         # SA2 style: separate dense feature rows
        pts = xyz
        feats = torch.randn(b, n, feat_stride, generator=g)
    fps = ora.furthest_point_sampling(xyz, m)
    new_xyz = ora.gather_points(xyz.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
    r = 0.25
    idx_o = ora.ball_query(new_xyz, xyz, r, ns)
    idx, cnt = rows.ball_query_rows(new_xyz.to(cuda), pts.to(cuda), [r], [ns])
    assert torch.equal(idx[0].cpu(), idx_o)
    stride = 8 if 3 + c <= 8 else (3 + c + 3) // 4 * 4
    fd = feats.to(cuda).requires_grad_(feat_off == 0)
    out = rows.group_rows(pts.to(cuda) if feat_off == 0 else fd, fd, new_xyz.to(cuda), idx[0], cnt[0], c, feat_off, stride)
    # oracle composition in the reference layout
    gx = ora.group_points(xyz.transpose(1, 2).contiguous(), idx_o) - new_xyz.transpose(1, 2).unsqueeze(-1)
    fsel = feats[:, :, feat_off:feat_off + c].transpose(1, 2).contiguous()
    want = torch.cat([gx, ora.group_points(fsel, idx_o)], 1).permute(0, 2, 3, 1)       # (b,m,ns,3+c)
    got = out.detach().cpu()
    assert got.shape == (b, m, ns, stride)
    assert torch.equal(got[..., : 3 + c], want) and float(got[..., 3 + c:].abs().sum()) == 0.0
    if feat_off == 0:
        w = torch.randn(b, m, ns, stride, generator=g)
        (out * w.to(cuda)).sum().backward()
        want_g = ora.group_points_grad(w[..., 3:3 + c].permute(0, 3, 1, 2).contiguous(), idx_o, n).transpose(1, 2)
        torch.testing.assert_close(fd.grad.cpu(), want_g, rtol=0, atol=2e-5)
        # deterministic: a second backward gives the same bits
        fd.grad = None
        out2 = rows.group_rows(pts.to(cuda), fd, new_xyz.to(cuda), idx[0], cnt[0], c, feat_off, stride)
        (out2 * w.to(cuda)).sum().backward()
        g1 = fd.grad.clone()
        fd.grad = None
        out3 = rows.group_rows(pts.to(cuda), fd, new_xyz.to(cuda), idx[0], cnt[0], c, feat_off, stride)
        (out3 * w.to(cuda)).sum().backward()
        assert torch.equal(g1, fd.grad)


@pytest.mark.parametrize("c,n,m,ns", [(192, 512, 128, 64), (64, 300, 33, 16), (8, 1000, 50, 32)])
def test_group_rows_feature_first_layout(cuda, c, n, m, ns):
    """[feats | xyz - centre | 0-pad] rows (float4 kernel) hold the same values as the reference order, permuted;
    backward sums the feature columns back deterministically"""
    from sg4d import rows
    b = 2
    g = torch.Generator().manual_seed(c + n)
    xyz = _clouds(c + n + 1, b, n, "gauss")
    feats = torch.randn(b, n, c, generator=g)
    fps = ora.furthest_point_sampling(xyz, m)
    new_xyz = ora.gather_points(xyz.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
    idx, cnt = rows.ball_query_rows(new_xyz.to(cuda), xyz.to(cuda), [0.3], [ns])
    stride = (3 + c + 3) // 4 * 4
    fd = feats.to(cuda).requires_grad_(True)
    ref = rows.group_rows(xyz.to(cuda), fd, new_xyz.to(cuda), idx[0], cnt[0], c, 0, stride, xyz_last=False)
    out = rows.group_rows(xyz.to(cuda), fd, new_xyz.to(cuda), idx[0], cnt[0], c, 0, stride, xyz_last=True)
    assert torch.equal(out[..., :c], ref[..., 3:3 + c]) and torch.equal(out[..., c:c + 3], ref[..., :3])
    assert float(out[..., c + 3:].abs().sum()) == 0.0
    w = torch.randn(b, m, ns, stride, generator=g).to(cuda)
    (out * w).sum().backward()
    g1 = fd.grad.clone()
    fd.grad = None
    w_ref = torch.cat([w[..., c:c + 3], w[..., :c], w[..., c + 3:]], dim=-1)
    (ref * w_ref).sum().backward()
    assert torch.equal(g1, fd.grad)


@pytest.mark.parametrize("b,n,m,c", [(3, 700, 150, 5), (2, 2048, 512, 64), (2, 64, 2, 3), (2, 300, 1, 4), (1, 9000, 2500, 8)])
def test_three_nn_and_interpolate_match_oracle(cuda, b, n, m, c):
    """feature-propagation ops through the C ABI: index / distance / interpolation
    outputs bit-exact vs the oracle, the atomics-based gradient within 2e-5"""
    e = _ext()
    g = torch.Generator().manual_seed(n + m)
    unknown, known = torch.rand(b, n, 3, generator=g), torch.rand(b, m, 3, generator=g)
    if m > 8:
        known[0, 7] = known[0, 3]
        known[0, 1] = float("nan")                              # never selected
    od, oi = ora.three_nn(unknown, known)
    sd, si = e.three_nn(unknown.to(cuda), known.to(cuda))
    np.testing.assert_array_equal(si.cpu().numpy(), oi.numpy())
    np.testing.assert_array_equal(sd.cpu().numpy(), od.numpy())
    feats = torch.randn(b, c, m, generator=g)
    w = torch.rand(b, n, 3, generator=g)
    np.testing.assert_array_equal(e.three_interpolate(feats.to(cuda), si, w.to(cuda)).cpu().numpy(),
                                  ora.three_interpolate(feats, oi, w).numpy())
    go = torch.randn(b, c, n, generator=g)
    torch.testing.assert_close(e.three_interpolate_grad(go.to(cuda), si, w.to(cuda), m).cpu(),
                               ora.three_interpolate_grad(go, oi, w, m), rtol=1e-5, atol=2e-5)


def test_fp_module_against_reference_fixture(cuda, golden_dir):
    """three_nn / three_interpolate wrappers and PointnetFPModule vs outputs of the REFERENCE's own Python
    (tests/golden/fp_module.npz, generated by tests/golden/make_golden.py through oracle/ref_harness.py)"""
    import json
    from oracle import weights
    from sg4d.pointnet2_ops import pointnet2_utils as U
    from sg4d.pointnet2_ops.pointnet2_modules import PointnetFPModule
    fx = np.load(os.path.join(golden_dir, "fp_module.npz"))
    unknown, known = torch.from_numpy(fx["unknown"]).to(cuda), torch.from_numpy(fx["known"]).to(cuda)
    dist, idx = U.three_nn(unknown, known)
    np.testing.assert_array_equal(idx.cpu().numpy(), fx["idx"])
    np.testing.assert_allclose(dist.cpu().numpy(), fx["dist"], rtol=1e-6, atol=0)       # sqrt on the GPU vs the CPU
    kf = torch.from_numpy(fx["known_feats"]).to(cuda).requires_grad_(True)
    interp = U.three_interpolate(kf, idx, torch.from_numpy(fx["weight"]).to(cuda))
    np.testing.assert_array_equal(interp.detach().cpu().numpy(), fx["interp"])
    (interp * torch.from_numpy(fx["w_interp"]).to(cuda)).sum().backward()
    np.testing.assert_allclose(kf.grad.cpu().numpy(), fx["d_known_feats"], rtol=1e-5, atol=2e-5)
    fp = PointnetFPModule(mlp=[6 + 4, 16, 8])
    fp.load_state_dict(weights.synth_state_dict(json.loads(str(fx["fp_shapes"])), seed=9))
    fp.to(cuda).train()
    kf2 = torch.from_numpy(fx["known_feats"]).to(cuda).requires_grad_(True)
    out = fp(unknown, known, torch.from_numpy(fx["unknown_feats"]).to(cuda), kf2)
    np.testing.assert_allclose(out.detach().cpu().numpy(), fx["fp_out"], rtol=0, atol=1e-4)
    (out * torch.from_numpy(fx["w_out"]).to(cuda)).sum().backward()
    np.testing.assert_allclose(kf2.grad.cpu().numpy(), fx["fp_d_known_feats"], rtol=1e-3, atol=1e-4)
    for k, p in fp.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), fx["fp_grad." + k], rtol=1e-3, atol=2e-4, err_msg=k)


def test_fp_module_forward_backward(cuda):
    """PointnetFPModule (OPS/pointnet2_modules.py:149-209) on the sg4d ops vs the same module arithmetic in torch"""
    from sg4d.pointnet2_ops.pointnet2_modules import PointnetFPModule
    torch.manual_seed(3)
    fp = PointnetFPModule(mlp=[6 + 4, 16, 8]).to(cuda).train()
    g = torch.Generator().manual_seed(4)
    unknown, known = torch.rand(2, 200, 3, generator=g).to(cuda), torch.rand(2, 40, 3, generator=g).to(cuda)
    uf = torch.randn(2, 4, 200, generator=g).to(cuda)
    kf = torch.randn(2, 6, 40, generator=g).to(cuda).requires_grad_(True)
    out = fp(unknown, known, uf, kf)
    out.sum().backward()
    g1 = kf.grad.clone()
    # the same maths with dense torch ops
    kf2 = kf.detach().clone().requires_grad_(True)
    d = ((unknown[:, :, None, :] - known[:, None, :, :]) ** 2).sum(-1).sqrt()
    dist, idx = torch.sort(d, dim=2, stable=True)
    dist, idx = dist[:, :, :3], idx[:, :, :3]
    rec = 1.0 / (dist + 1e-8)
    wgt = rec / rec.sum(2, keepdim=True)
    gathered = torch.gather(kf2[:, :, None, :].expand(-1, -1, 200, -1), 3, idx[:, None].expand(-1, 6, -1, -1))
    interp = (gathered * wgt[:, None]).sum(-1)
    out2 = fp.mlp(torch.cat([interp, uf], dim=1).unsqueeze(-1)).squeeze(-1)
    out2.sum().backward()
    torch.testing.assert_close(out, out2, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(g1, kf2.grad, rtol=1e-3, atol=1e-4)


def test_gnn_gather_and_scatter(cuda):
    from sg4d import rows
    g = torch.Generator().manual_seed(9)
    n_nodes, d, de, dh = 24, 256, 256, 512
    ei = torch.stack([torch.randint(0, n_nodes, (132,), generator=g), torch.randint(0, n_nodes, (132,), generator=g)])
    x = torch.randn(n_nodes, d, generator=g)
    e = torch.randn(132, de, generator=g)
    csr = rows.EdgeCSR(ei.to(cuda), n_nodes)
    xd, ed = x.to(cuda).requires_grad_(True), e.to(cuda).requires_grad_(True)
    out = rows.triplet_gather(xd, ed, csr)
    xr, er = x.clone().requires_grad_(True), e.clone().requires_grad_(True)
    want = torch.cat([xr.index_select(0, ei[1]), er, xr.index_select(0, ei[0])], 1)
    assert torch.equal(out.detach().cpu(), want.detach())
    w = torch.randn(132, 2 * d + de, generator=g)
    (out * w.to(cuda)).sum().backward()
    (want * w).sum().backward()
    torch.testing.assert_close(xd.grad.cpu(), xr.grad, rtol=1e-5, atol=1e-5)
    assert torch.equal(ed.grad.cpu(), er.grad)

    h = torch.randn(132, 2 * dh + de, generator=g)
    hd, hr = h.to(cuda).requires_grad_(True), h.clone().requires_grad_(True)
    m = rows.message_aggregate(hd, dh, de, csr)
    msg = hr[:, :dh] + hr[:, dh + de:]
    want_m = torch.zeros(n_nodes, dh).index_add_(0, ei[1], msg)
    torch.testing.assert_close(m.detach().cpu(), want_m.detach(), rtol=1e-5, atol=1e-5)
    w2 = torch.randn(n_nodes, dh, generator=g)
    (m * w2.to(cuda)).sum().backward()
    (want_m * w2).sum().backward()
    assert torch.equal(hd.grad.cpu(), hr.grad)


# ------------------------------------------------------------------ full-size properties (80 000 points)

def test_full_size_properties(cuda):
    """BASELINE shapes: size-independent properties instead of a (slow) CPU replay."""
    from sg4d import rows, synthetic
    gen = torch.Generator().manual_seed(77)
    pts = torch.stack([synthetic.make_cloud(gen, 80000, 7) for _ in range(6)]).to(cuda)
    idx, new_xyz = rows.fps_rows(pts, 512)
    idx_l = idx.long()
    xyz = pts[:, :, :3]
    assert (idx_l[:, 0] == 0).all() and int(idx_l.min()) >= 0 and int(idx_l.max()) < 80000
    picked = torch.gather(xyz, 1, idx_l.unsqueeze(-1).expand(-1, -1, 3))
    assert torch.equal(picked, new_xyz)
    # FPS property: the distance of each pick to the previously picked set never increases
    for bi in range(pts.shape[0]):
        d = torch.cdist(picked[bi].double(), picked[bi].double())
        mins = torch.stack([d[j, :j].min() for j in range(1, 512)])
        valid = xyz[bi].pow(2).sum(1)[idx_l[bi, 1:]] > 1e-3
        mv = mins[valid]
        assert (mv[1:] <= mv[:-1] + 1e-6).all()
    # repeatable
    idx2, _ = rows.fps_rows(pts, 512)
    assert torch.equal(idx, idx2)
    # ball query: hits inside the radius, ascending, padded with the first hit
    (i1, i2), (c1, c2) = rows.ball_query_rows(new_xyz, pts, [0.1, 0.2], [16, 32])
    for ii, cc, r, ns in ((i1, c1, 0.1, 16), (i2, c2, 0.2, 32)):
        nb = torch.gather(xyz.unsqueeze(1).expand(-1, 512, -1, -1), 2, ii.long().unsqueeze(-1).expand(-1, -1, -1, 3))
        d2 = (nb - new_xyz.unsqueeze(2)).pow(2).sum(-1)
        assert float(d2.max()) < r * r * (1 + 1e-5)
        k = torch.arange(ns, device=cuda)
        in_prefix = k.view(1, 1, -1) < cc.unsqueeze(-1)
        asc = (ii[:, :, 1:] > ii[:, :, :-1]) | ~in_prefix[:, :, 1:]
        assert asc.all() and ((ii == ii[:, :, :1]) | in_prefix).all() and (cc >= 1).all()
    # one radius answered alone gives the same rows as the fused two-radius pass
    (j1,), _ = rows.ball_query_rows(new_xyz, pts, [0.1], [16])
    assert torch.equal(j1, i1)


def test_invalid_arguments_raise(cuda):
    from sg4d import _lib
    x = torch.zeros(1, 8, 3, device=cuda)
    with pytest.raises(RuntimeError, match="invalid argument"):
        _lib.call("sg4d_furthest_point_sampling", x, 1, 0, 4, x.data_ptr(), 0, x.data_ptr())
    with pytest.raises(RuntimeError, match="invalid argument"):
        _lib.call("sg4d_ball_query", x, 1, 8, 4, 0.1, 0, x.data_ptr(), x.data_ptr(), x.data_ptr())
