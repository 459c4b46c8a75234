"""Pins the oracle to the REFERENCE ITSELF: the reference's own CUDA kernels (its _ext-src sources
compiled unmodified for sm_100a into oracle/_ref/ by oracle/build_ref_ext.py) are run on the B200 and
compared bit for bit with oracle/pn2_oracle.c (the CPU restatement) and with libsg4d.so."""
import numpy as np
import pytest
import torch

from oracle import build_ref_ext, pn2_ext_cpu as ora

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    mod = build_ref_ext.load_module()
    if mod is None:
        pytest.skip("oracle/_ref/pn2_ref_ext.so was not built (container-only recipe)")
    return mod


def _clouds(seed, b, n):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(b, n, 3, generator=g) * 2 - 1
    xyz[0, n // 2:] = xyz[0, : n - n // 2].clone()
    if b > 1:
        xyz[1, torch.randperm(n, generator=g)[: max(1, n // 8)]] = 0.0
        xyz[1, 0] = 0.0
    if b > 2:
        xyz[2] = (xyz[2] * 4).round() / 4
    if b > 3:
        xyz[3] = torch.randn(n, 3, generator=g) * 0.05
    return xyz.contiguous()


@pytest.mark.parametrize("b,n,m", [(4, 37, 20), (4, 300, 64), (4, 700, 96), (4, 2048, 512), (4, 8000, 512), (2, 80000, 512)])
def test_reference_fps_kernel_equals_oracle_and_sg4d(cuda, ref, b, n, m):
    from sg4d.pointnet2_ops import _ext
    xyz = _clouds(31 + n, b, n)
    r = ref.furthest_point_sampling(xyz.to(cuda), m).cpu()
    np.testing.assert_array_equal(ora.furthest_point_sampling(xyz, m).numpy(), r.numpy())
    np.testing.assert_array_equal(_ext.furthest_point_sampling(xyz.to(cuda), m).cpu().numpy(), r.numpy())


@pytest.mark.parametrize("b,n,m,r,ns", [(4, 700, 96, 0.35, 16), (4, 2048, 512, 0.1, 16), (4, 2048, 512, 0.2, 32),
                                        (3, 512, 128, 0.4, 64), (2, 20000, 256, 0.05, 32)])
def test_reference_ball_query_kernel_equals_oracle_and_sg4d(cuda, ref, b, n, m, r, ns):
    from sg4d.pointnet2_ops import _ext
    xyz = _clouds(41 + n + ns, b, n)
    fps = ora.furthest_point_sampling(xyz, m)
    new_xyz = ora.gather_points(xyz.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
    want = ref.ball_query(new_xyz.to(cuda), xyz.to(cuda), r, ns).cpu()
    np.testing.assert_array_equal(ora.ball_query(new_xyz, xyz, r, ns).numpy(), want.numpy())
    np.testing.assert_array_equal(_ext.ball_query(new_xyz.to(cuda), xyz.to(cuda), r, ns).cpu().numpy(), want.numpy())


def test_reference_gather_group_kernels_equal_oracle_and_sg4d(cuda, ref):
    from sg4d.pointnet2_ops import _ext
    g = torch.Generator().manual_seed(3)
    b, c, n, m, ns = 3, 9, 400, 50, 16
    pts = torch.randn(b, c, n, generator=g)
    i1 = torch.randint(0, n, (b, m), generator=g, dtype=torch.int32)
    i2 = torch.randint(0, n, (b, m, ns), generator=g, dtype=torch.int32)
    d = lambda t: t.to(cuda)
    assert torch.equal(ref.gather_points(d(pts), d(i1)).cpu(), ora.gather_points(pts, i1))
    assert torch.equal(ref.group_points(d(pts), d(i2)).cpu(), ora.group_points(pts, i2))
    assert torch.equal(_ext.gather_points(d(pts), d(i1)).cpu(), ref.gather_points(d(pts), d(i1)).cpu())
    assert torch.equal(_ext.group_points(d(pts), d(i2)).cpu(), ref.group_points(d(pts), d(i2)).cpu())
    go = torch.randn(b, c, m, ns, generator=g)
    torch.testing.assert_close(ref.group_points_grad(d(go), d(i2), n).cpu(), ora.group_points_grad(go, i2, n),
                               rtol=0, atol=1e-5)
    torch.testing.assert_close(_ext.group_points_grad(d(go), d(i2), n).cpu(), ref.group_points_grad(d(go), d(i2), n).cpu(),
                               rtol=0, atol=1e-5)


@pytest.mark.parametrize("b,n,m,c", [(3, 700, 150, 5), (2, 2048, 512, 64), (2, 64, 2, 3), (1, 5000, 1300, 16)])
def test_reference_interpolate_kernels_equal_oracle_and_sg4d(cuda, ref, b, n, m, c):
    """three_nn / three_interpolate (+grad): the reference's kernels vs the C oracle vs libsg4d.so"""
    from sg4d.pointnet2_ops import _ext
    g = torch.Generator().manual_seed(b + n + m)
    unknown, known = torch.rand(b, n, 3, generator=g), torch.rand(b, m, 3, generator=g)
    if m > 8:
        known[0, 7] = known[0, 3]                               # exact distance ties
    rd, ri = ref.three_nn(unknown.to(cuda), known.to(cuda))
    od, oi = ora.three_nn(unknown, known)
    sd, si = _ext.three_nn(unknown.to(cuda), known.to(cuda))
    np.testing.assert_array_equal(ri.cpu().numpy(), oi.numpy())
    np.testing.assert_array_equal(si.cpu().numpy(), oi.numpy())
    np.testing.assert_array_equal(rd.cpu().numpy(), od.numpy())
    np.testing.assert_array_equal(sd.cpu().numpy(), od.numpy())
    feats = torch.randn(b, c, m, generator=g)
    w = torch.rand(b, n, 3, generator=g)
    r_out = ref.three_interpolate(feats.to(cuda), ri, w.to(cuda)).cpu()
    np.testing.assert_array_equal(ora.three_interpolate(feats, oi, w).numpy(), r_out.numpy())
    np.testing.assert_array_equal(_ext.three_interpolate(feats.to(cuda), si, w.to(cuda)).cpu().numpy(), r_out.numpy())
    go = torch.randn(b, c, n, generator=g)
    r_g = ref.three_interpolate_grad(go.to(cuda), ri, w.to(cuda), m).cpu()
    torch.testing.assert_close(ora.three_interpolate_grad(go, oi, w, m), r_g, rtol=1e-5, atol=2e-5)   # atomics order
    torch.testing.assert_close(_ext.three_interpolate_grad(go.to(cuda), si, w.to(cuda), m).cpu(), r_g, rtol=1e-5, atol=2e-5)
