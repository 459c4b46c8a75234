"""Set-abstraction modules with the constructor signatures, attribute names (``npoint``, ``groupers``,
``mlps``) and ``state_dict`` layout of the reference's ``pointnet2_ops.pointnet2_modules``
(OPS/pointnet2_modules.py:9-146).

``forward(xyz, features)`` keeps the reference contract (xyz (B,N,3), features (B,C,N) ->
new_xyz (B,npoint,3), new_features (B,sum C_out,npoint)).  Internally everything runs point-major
through ``forward_rows``: one FPS launch that also emits the picked centres, one ball-query launch
for all radii of the level, then per scale the fused tensor-core kernels
that go from the ball-query indices to the pooled features (``mlp.fused_sa_scale``; other shapes: grouped rows +
``dense.pooled_shared_mlp``).  The shared MLP is
still ``[1x1 conv -> BatchNorm2d -> ReLU] x L`` (modules.py:9-19) evaluated on the same values --
a 1x1 convolution over (B,C,npoint,nsample) IS a matrix product over rows -- so parameters,
running statistics and results match the reference module.
"""
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import dense
from .. import mlp as fused
from .. import rows
from . import pointnet2_utils


def build_shared_mlp(mlp_spec: List[int], bn: bool = True):
    layers = []
    for c_in, c_out in zip(mlp_spec[:-1], mlp_spec[1:]):
        layers.append(nn.Conv2d(c_in, c_out, kernel_size=1, bias=not bn))
        if bn:
            layers.append(nn.BatchNorm2d(c_out))
        layers.append(nn.ReLU(True))
    return nn.Sequential(*layers)


def _pad4(k):
    return (k + 3) // 4 * 4


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None

    def forward_rows(self, pts: torch.Tensor, feats: Optional[torch.Tensor], feat_offset: int,
                     c: int) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
        """Point-major SA level.

        pts (B,n,S): xyz in columns 0..2.  feats (B,n,Sf) or None: the C feature channels are
        columns feat_offset..feat_offset+c-1 (feats may be the same tensor as pts).
        Returns (new_xyz (B,npoint,3) or None, new_features (B,npoint,sum C_out) point-major).
        """
        b, n = pts.shape[0], pts.shape[1]
        outs = []
        if self.npoint is None:  # GroupAll (utils.py:353-383): xyz not recentred, one group of n points
            if not self.groupers[0].use_xyz and feats is None:
                raise RuntimeError("GroupAll without xyz needs features")
            needs_dx = feats is not None and feats.requires_grad and torch.is_grad_enabled()
            if n not in (1, 2, 4, 8, 16, 32, 64, 128):
                raise NotImplementedError(f"GroupAll over {n} points: the fused max-pool needs a power of two <= 128")
            f = feats[:, :, feat_offset:feat_offset + c] if feats is not None and c > 0 else None
            use_xyz = self.groupers[0].use_xyz or f is None
            k = (3 if use_xyz else 0) + (c if f is not None else 0)
            pad = pts.new_zeros(b, n, _pad4(k) - k)
            # feature-first columns when a gradient flows back (16-byte aligned dX); the weight columns are permuted to match
            xyz_last = needs_dx and use_xyz
            cols = ([f] if f is not None else []) + ([pts[:, :, :3]] if use_xyz else []) if xyz_last else \
                   ([pts[:, :, :3]] if use_xyz else []) + ([f] if f is not None else [])
            x = torch.cat(cols + [pad], dim=2).reshape(b * n, _pad4(k))
            for mlp in self.mlps:
                if not use_xyz:
                    raise NotImplementedError("GroupAll with use_xyz=False")
                outs.append(dense.pooled_shared_mlp(x, k, n, mlp, xyz_last).view(b, 1, -1))
            return None, torch.cat(outs, dim=2) if len(outs) > 1 else outs[0]

        # large clouds: one spatial index (Morton-sorted copy + bucket boxes) serves both FPS and the ball query
        index = rows.SpatialIndex(pts) if rows.wants_index(n) else None
        _, new_xyz = rows.fps_rows(pts, self.npoint, index)
        radii = [g.radius for g in self.groupers]
        nsamples = [g.nsample for g in self.groupers]
        if index is not None and max(nsamples) > 64:
            raise NotImplementedError("ball query through the spatial index keeps up to 64 samples per centre")
        idx, cnt = rows.ball_query_rows(new_xyz, pts, radii, nsamples, index)
        needs_dx = feats is not None and feats.requires_grad and torch.is_grad_enabled()
        for s, mlp in enumerate(self.mlps):
            assert self.groupers[s].use_xyz, "the hot path always groups xyz (use_xyz=True)"
            k = 3 + c
            fsrc = feats if feats is not None else pts
            kind = fused.sa_scale_kind(mlp, c, nsamples[s], needs_dx, fsrc.shape[2], feat_offset)
            if kind is not None:
                # ball-query indices -> pooled features in the tensor-core kernels; the grouped tensor never exists
                outs.append(fused.fused_sa_scale(kind, pts, feats, feat_offset, c, new_xyz, idx[s], cnt[s], mlp)
                            .view(b, self.npoint, -1))
                continue
            # any other shape: materialised grouped rows + the generic tensor-core MLP (zero-padded widths); no PyTorch path
            stride = 8 if k <= 8 else _pad4(k)
            xyz_last = needs_dx                 # feature-first columns when a gradient flows back (aligned dX)
            x = rows.group_rows(pts, feats, new_xyz, idx[s], cnt[s], c, feat_offset, stride, xyz_last)
            outs.append(dense.pooled_shared_mlp(x.view(-1, stride), k, nsamples[s], mlp, xyz_last).view(b, self.npoint, -1))
        return new_xyz, torch.cat(outs, dim=2) if len(outs) > 1 else outs[0]

    def forward(self, xyz: torch.Tensor, features: Optional[torch.Tensor]
                ) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
        feats = features.transpose(1, 2).contiguous() if features is not None else None
        c = feats.shape[2] if feats is not None else 0
        new_xyz, out = self.forward_rows(xyz.contiguous(), feats, 0, c)
        return new_xyz, out.transpose(1, 2).contiguous()


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Multi-scale-grouping set abstraction (OPS/pointnet2_modules.py:77-115)."""

    def __init__(self, npoint, radii, nsamples, mlps, bn=True, use_xyz=True):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for radius, nsample, spec in zip(radii, nsamples, mlps):
            self.groupers.append(pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz)
                                 if npoint is not None else pointnet2_utils.GroupAll(use_xyz))
            if use_xyz:
                spec[0] += 3  # in place, like the reference (modules.py:112-113)
            self.mlps.append(build_shared_mlp(spec, bn))


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction (OPS/pointnet2_modules.py:118-146)."""

    def __init__(self, mlp, npoint=None, radius=None, nsample=None, bn=True, use_xyz=True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn,
                         use_xyz=use_xyz)


class PointnetFPModule(nn.Module):
    """Feature propagation (OPS/pointnet2_modules.py:149-209) on sg4d's three_nn / three_interpolate kernels
    (csrc/interpolate.cu).  Not used by the scene-graph encoder; provided so that the package is a complete
    drop-in for ``pointnet2_ops``."""

    def __init__(self, mlp, bn=True):
        super().__init__()
        self.mlp = build_shared_mlp(mlp, bn=bn)

    def forward(self, unknown, known, unknow_feats, known_feats):
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
            interpolated = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated = known_feats.expand(*(list(known_feats.size()[0:2]) + [unknown.size(1)]))
        x = interpolated if unknow_feats is None else torch.cat([interpolated, unknow_feats], dim=1)
        return self.mlp(x.unsqueeze(-1)).squeeze(-1)
