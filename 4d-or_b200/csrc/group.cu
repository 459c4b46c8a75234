// group.cu -- gather / group kernels for sm_100a.
//
// One-to-one replacements of the reference kernels
//   gather_points(_grad)   EXT/src/sampling_gpu.cu:8-57
//   group_points(_grad)    EXT/src/group_points_gpu.cu:8-75
// (channel-major tensors, the operator-API contract) plus the point-major fused variants used by the
// model path (group + recentre + concat in one pass; deterministic gather-style backward).
//
// The reference runs ONE CTA per cloud and lets each thread walk `nsample` outputs (writes strided by
// nsample*4 B across a warp).  Here every kernel is a flat grid-stride over OUTPUT elements with the
// fastest-varying output index on threadIdx.x, so stores are fully coalesced and the grid fills all
// 148 SMs regardless of the number of clouds.
#include "common.cuh"

namespace sg4d {

// ---------------------------------------------------------------- channel-major (reference layout)

__global__ void __launch_bounds__(256)
gather_points_kernel(long long total, int c, int n, int m, const float *__restrict__ points,
                     const int32_t *__restrict__ idx, float *__restrict__ out) {
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int j = (int)(e % m);
        const long long bc = e / m;  // b*c + l
        const long long bi = bc / c;
        out[e] = __ldg(points + bc * n + __ldg(idx + bi * m + j));
    }
}

__global__ void __launch_bounds__(256)
gather_points_grad_kernel(long long total, int c, int n, int m, const float *__restrict__ grad_out,
                          const int32_t *__restrict__ idx, float *__restrict__ grad_points) {
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int j = (int)(e % m);
        const long long bc = e / m;
        const long long bi = bc / c;
        atomicAdd(grad_points + bc * n + __ldg(idx + bi * m + j), __ldg(grad_out + e));
    }
}

// out[b,l,j,k] = points[b,l,idx[b,j,k]]; e enumerates (b,l,j,k) with k fastest
__global__ void __launch_bounds__(256)
group_points_kernel(long long total, int c, int n, int npoints, int nsample,
                    const float *__restrict__ points, const int32_t *__restrict__ idx,
                    float *__restrict__ out) {
    const long long jk = (long long)npoints * nsample;
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long bc = e / jk;
        const long long r = e - bc * jk;  // j*nsample + k
        const long long bi = bc / c;
        out[e] = __ldg(points + bc * n + __ldg(idx + bi * jk + r));
    }
}

__global__ void __launch_bounds__(256)
group_points_grad_kernel(long long total, int c, int n, int npoints, int nsample,
                         const float *__restrict__ grad_out, const int32_t *__restrict__ idx,
                         float *__restrict__ grad_points) {
    const long long jk = (long long)npoints * nsample;
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long bc = e / jk;
        const long long r = e - bc * jk;
        const long long bi = bc / c;
        atomicAdd(grad_points + bc * n + __ldg(idx + bi * jk + r), __ldg(grad_out + e));
    }
}

static unsigned flat_grid(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = (long long)SG4D_NUM_SMS * 32;  // grid-stride beyond 32 CTAs per SM
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

// ---------------------------------------------------------------- point-major fused grouping

// narrow rows (3 + c <= 16, SA1): one thread per output row, vectorised 16-byte stores
template <int STRIDE>
__global__ void __launch_bounds__(256)
group_rows_narrow_kernel(long long rows, int n, int m, int nsample, int c, int pts_stride,
                         int feat_stride, int feat_offset, const float *__restrict__ pts,
                         const float *__restrict__ feats, const float *__restrict__ centers,
                         const int32_t *__restrict__ idx, float *__restrict__ out) {
    const long long per_cloud = (long long)m * nsample;
    for (long long r = blockIdx.x * 256LL + threadIdx.x; r < rows; r += (long long)gridDim.x * 256) {
        const long long bi = r / per_cloud;
        const long long j = (r - bi * per_cloud) / nsample;
        const int i = __ldg(idx + r);
        const float *p = pts + (bi * n + i) * pts_stride;
        const float *f = feats + (bi * n + i) * feat_stride + feat_offset;
        const float *q = centers + (bi * m + j) * 3;
        float v[STRIDE];
        v[0] = __ldg(p) - __ldg(q);
        v[1] = __ldg(p + 1) - __ldg(q + 1);
        v[2] = __ldg(p + 2) - __ldg(q + 2);
#pragma unroll
        for (int ch = 0; ch < STRIDE - 3; ++ch) v[3 + ch] = ch < c ? __ldg(f + ch) : 0.f;
        float4 *o = reinterpret_cast<float4 *>(out + r * STRIDE);
#pragma unroll
        for (int u = 0; u < STRIDE / 4; ++u) o[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
    }
}

// wide rows (SA2 / anything): one warp per output row, lanes stride over the columns
__global__ void __launch_bounds__(256)
group_rows_wide_kernel(long long rows, int n, int m, int nsample, int c, int pts_stride, int feat_stride,
                       int feat_offset, int out_stride, int xyz_col0, const float *__restrict__ pts,
                       const float *__restrict__ feats, const float *__restrict__ centers,
                       const int32_t *__restrict__ idx, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long per_cloud = (long long)m * nsample;
    const long long warp0 = (blockIdx.x * 256LL + threadIdx.x) >> 5, nwarp = ((long long)gridDim.x * 256) >> 5;
    for (long long r = warp0; r < rows; r += nwarp) {
        const long long bi = r / per_cloud;
        const long long j = (r - bi * per_cloud) / nsample;
        const int i = __ldg(idx + r);
        const float *f = feats + (bi * n + i) * feat_stride + feat_offset;
        float *o = out + r * out_stride;
        const int fcol0 = xyz_col0 == 0 ? 3 : 0;   // features first (xyz_col0 == c) or xyz first (xyz_col0 == 0)
        for (int col = lane; col < out_stride; col += 32) {
            float v = 0.f;
            if (col >= xyz_col0 && col < xyz_col0 + 3)
                v = __ldg(pts + (bi * n + i) * pts_stride + (col - xyz_col0)) - __ldg(centers + (bi * m + j) * 3 + (col - xyz_col0));
            else if (col >= fcol0 && col < fcol0 + c)
                v = __ldg(f + col - fcol0);
            o[col] = v;
        }
    }
}

// wide rows, feature-first layout (xyz_col0 == c, c % 4 == 0, 16-byte aligned rows): lanes move float4 columns, two
// output rows per warp in flight
__global__ void __launch_bounds__(256)
group_rows_wide4_kernel(long long rows, int n, int m, int nsample, int c, int pts_stride, int feat_stride,
                        int feat_offset, int out_stride, const float *__restrict__ pts,
                        const float *__restrict__ feats, const float *__restrict__ centers,
                        const int32_t *__restrict__ idx, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long per_cloud = (long long)m * nsample;
    const long long warp0 = (blockIdx.x * 256LL + threadIdx.x) >> 5, nwarp = ((long long)gridDim.x * 256) >> 5;
    const int c4 = c >> 2, o4 = out_stride >> 2;   // float4 columns: [0, c4) features, c4 = {xyz - centre, 0}, then zeros
    for (long long r0 = 2 * warp0; r0 < rows; r0 += 2 * nwarp) {
        long long bi[2], j[2];
        int i[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long r = r0 + u;
            ok[u] = r < rows;
            bi[u] = ok[u] ? r / per_cloud : 0;
            j[u] = ok[u] ? (r - bi[u] * per_cloud) / nsample : 0;
            i[u] = ok[u] ? __ldg(idx + r) : 0;
        }
        for (int col = lane; col < o4; col += 32) {
            float4 v[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok[u]) {
                    if (col < c4) {
                        v[u] = __ldg(reinterpret_cast<const float4 *>(feats + (bi[u] * n + i[u]) * feat_stride + feat_offset) + col);
                    } else if (col == c4) {
                        const float *p = pts + (bi[u] * n + i[u]) * pts_stride, *q = centers + (bi[u] * m + j[u]) * 3;
                        v[u] = make_float4(__ldg(p) - __ldg(q), __ldg(p + 1) - __ldg(q + 1), __ldg(p + 2) - __ldg(q + 2), 0.f);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (ok[u]) reinterpret_cast<float4 *>(out + (r0 + u) * out_stride)[col] = v[u];
        }
    }
}

// Backward of the feature columns, as a GATHER (no atomics, fixed summation order).
// One CTA per (cloud, slice of source points); one warp per source point i; the lanes first search
// the 128-ish centre rows in parallel (each row of a ball query is an ascending run of `cnt` distinct
// indices followed by repeats of the first), then walk the hits in (j, k) order while each lane
// accumulates c/32 channels with coalesced 128-byte reads of grad_out.
constexpr int kGgWarps = 16;
// V4: the feature columns are 16-byte aligned (gcol0, out_stride, c multiples of 4): lanes accumulate float4 columns
// DY: the summed rows are not stored anywhere -- they are the first layer's dY1 = p .* dz1 - (q .* y1 + u), evaluated from dz1
//     (= grad_out) and y1 while they are summed (c = C1 <= 128 channels, V4 layout).  Because dX = dY1 W1 is linear, the
//     feature gradient of a source point is (sum of its rows' dY1) W1: one small GEMM over B*n rows afterwards replaces the
//     dX GEMM over all B*m*nsample grouped rows and the (rows x 196) dX tensor it wrote for this kernel to read back.
template <bool V4, bool DY = false>
__global__ void __launch_bounds__(kGgWarps * 32)
group_rows_grad_kernel(int n, int m, int nsample, int c, int out_stride, int gcol0, int slices, int accumulate,
                       const float *__restrict__ grad_out, const int32_t *__restrict__ idx,
                       const int32_t *__restrict__ cnt, float *__restrict__ grad_feats, const float *__restrict__ y1 = nullptr,
                       const float *__restrict__ p1 = nullptr, const float *__restrict__ q1 = nullptr,
                       const float *__restrict__ u1 = nullptr) {
    extern __shared__ int32_t s_idx[];  // (m, nsample) indices of this cloud, then (m) counts
    int32_t *s_cnt = s_idx + (size_t)m * nsample;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x / slices, slice = blockIdx.x % slices;
    idx += (size_t)cloud * m * nsample;
    cnt += (size_t)cloud * m;
    grad_out += (size_t)cloud * m * nsample * out_stride;
    grad_feats += (size_t)cloud * n * c;
    float4 pv = make_float4(0.f, 0.f, 0.f, 0.f), qv = pv, uv = pv;
    if (DY) {
        y1 += (size_t)cloud * m * nsample * out_stride;
        if (lane < (c >> 2)) {
            pv = __ldg(reinterpret_cast<const float4 *>(p1) + lane), qv = __ldg(reinterpret_cast<const float4 *>(q1) + lane);
            uv = __ldg(reinterpret_cast<const float4 *>(u1) + lane);
        }
    }
    for (int e = tid; e < m * nsample; e += kGgWarps * 32) s_idx[e] = __ldg(idx + e);
    for (int e = tid; e < m; e += kGgWarps * 32) s_cnt[e] = __ldg(cnt + e);
    __syncthreads();

    constexpr int kMaxChunk = 8;  // c <= 256
    const int per_slice = (n + slices - 1) / slices;
    const int i_end = min(n, (slice + 1) * per_slice);
    const int c4 = c >> 2;
    for (int i = slice * per_slice + warp; i < i_end; i += kGgWarps) {
        float acc[kMaxChunk];
#pragma unroll
        for (int u = 0; u < kMaxChunk; ++u) acc[u] = 0.f;   // V4: acc[4v .. 4v+3] = float4 column lane + 32 v
        for (int j0 = 0; j0 < m; j0 += 32) {
            const int j = j0 + lane;
            int pos = -1, cj = 0;
            if (j < m) {
                cj = s_cnt[j];
                if (cj == 0) {  // a row without hits is all zeros: it groups point 0 nsample times
                    if (i == 0) pos = 0;
                    cj = 1;
                } else {
                    const int32_t *row = s_idx + (size_t)j * nsample;
                    int lo = 0, hi = cj;  // lower_bound over the ascending prefix [0, cj)
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (row[mid] < i) lo = mid + 1; else hi = mid;
                    }
                    if (lo < cj && row[lo] == i) pos = lo;
                }
            }
            unsigned mask = __ballot_sync(0xffffffffu, pos >= 0);
            while (mask) {
                const int l = __ffs(mask) - 1;
                mask &= mask - 1;
                const int jj = j0 + l;
                const int pp = __shfl_sync(0xffffffffu, pos, l);
                const int cc = __shfl_sync(0xffffffffu, cj, l);
                if (DY && pp != 0 && mask) {
                    // two plain matches (one row each) per trip: four loads in flight, added in the same order as one by one
                    const int l2 = __ffs(mask) - 1;
                    const int pp2 = __shfl_sync(0xffffffffu, pos, l2);
                    if (pp2 != 0) {
                        mask &= mask - 1;
                        if (lane < c4) {
                            const size_t ra = ((size_t)jj * nsample + pp) * out_stride, rb = ((size_t)(j0 + l2) * nsample + pp2) * out_stride;
                            const float4 dza = __ldg(reinterpret_cast<const float4 *>(grad_out + ra + gcol0) + lane);
                            const float4 ya = __ldg(reinterpret_cast<const float4 *>(y1 + ra) + lane);
                            const float4 dzb = __ldg(reinterpret_cast<const float4 *>(grad_out + rb + gcol0) + lane);
                            const float4 yb = __ldg(reinterpret_cast<const float4 *>(y1 + rb) + lane);
                            acc[0] += fmaf(dza.x, pv.x, -fmaf(ya.x, qv.x, uv.x)), acc[1] += fmaf(dza.y, pv.y, -fmaf(ya.y, qv.y, uv.y));
                            acc[2] += fmaf(dza.z, pv.z, -fmaf(ya.z, qv.z, uv.z)), acc[3] += fmaf(dza.w, pv.w, -fmaf(ya.w, qv.w, uv.w));
                            acc[0] += fmaf(dzb.x, pv.x, -fmaf(yb.x, qv.x, uv.x)), acc[1] += fmaf(dzb.y, pv.y, -fmaf(yb.y, qv.y, uv.y));
                            acc[2] += fmaf(dzb.z, pv.z, -fmaf(yb.z, qv.z, uv.z)), acc[3] += fmaf(dzb.w, pv.w, -fmaf(yb.w, qv.w, uv.w));
                        }
                        continue;
                    }
                }
                const float *g = grad_out + ((size_t)jj * nsample) * out_stride + gcol0;
                // slot pp, then (when i is the first hit) the padding slots cc..nsample-1
                int k = pp;
                while (k < nsample) {
                    const float *gr = g + (size_t)k * out_stride;
                    if (DY) {      // c <= 128: one float4 column group per lane
                        if (lane < c4) {
                            const float4 dz = __ldg(reinterpret_cast<const float4 *>(gr) + lane);
                            const float4 yy = __ldg(reinterpret_cast<const float4 *>(y1 + ((size_t)jj * nsample + k) * out_stride) + lane);
                            acc[0] += fmaf(dz.x, pv.x, -fmaf(yy.x, qv.x, uv.x)), acc[1] += fmaf(dz.y, pv.y, -fmaf(yy.y, qv.y, uv.y));
                            acc[2] += fmaf(dz.z, pv.z, -fmaf(yy.z, qv.z, uv.z)), acc[3] += fmaf(dz.w, pv.w, -fmaf(yy.w, qv.w, uv.w));
                        }
                    } else if (V4) {
#pragma unroll
                        for (int v = 0; v < kMaxChunk / 4; ++v) {
                            const int q4 = lane + 32 * v;
                            if (q4 < c4) {
                                const float4 t = __ldg(reinterpret_cast<const float4 *>(gr) + q4);
                                acc[4 * v] += t.x, acc[4 * v + 1] += t.y, acc[4 * v + 2] += t.z, acc[4 * v + 3] += t.w;
                            }
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < kMaxChunk; ++u) {
                            const int ch = lane + 32 * u;
                            if (ch < c) acc[u] += __ldg(gr + ch);
                        }
                    }
                    if (pp != 0) break;
                    k = (k == pp) ? cc : k + 1;
                }
            }
        }
        float *o = grad_feats + (size_t)i * c;
        if (V4) {
#pragma unroll
            for (int v = 0; v < kMaxChunk / 4; ++v) {
                const int q4 = lane + 32 * v;
                if (q4 < c4) {
                    float4 t = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
                    float4 *dst = reinterpret_cast<float4 *>(o) + q4;
                    if (accumulate) {
                        const float4 old = *dst;
                        t.x += old.x, t.y += old.y, t.z += old.z, t.w += old.w;
                    }
                    *dst = t;
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < kMaxChunk; ++u) {
                const int ch = lane + 32 * u;
                if (ch < c) o[ch] = accumulate ? o[ch] + acc[u] : acc[u];
            }
        }
    }
}

}  // namespace sg4d

using namespace sg4d;

extern "C" int sg4d_gather_points(int b, int c, int n, int npoints, const float *points,
                                  const int32_t *idx, float *out, sg4d_stream_t stream) {
    if (b < 0 || c < 0 || n <= 0 || npoints < 0 || !points || !idx || !out) return SG4D_EINVAL;
    const long long total = (long long)b * c * npoints;
    if (total == 0) return SG4D_OK;
    gather_points_kernel<<<flat_grid(total), 256, 0, (cudaStream_t)stream>>>(total, c, n, npoints, points, idx, out);
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out,
                                       const int32_t *idx, float *grad_points, sg4d_stream_t stream) {
    if (b < 0 || c < 0 || n <= 0 || npoints < 0 || !grad_out || !idx || !grad_points) return SG4D_EINVAL;
    const long long total = (long long)b * c * npoints;
    if (total == 0) return SG4D_OK;
    gather_points_grad_kernel<<<flat_grid(total), 256, 0, (cudaStream_t)stream>>>(total, c, n, npoints, grad_out,
                                                                                 idx, grad_points);
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                                 const int32_t *idx, float *out, sg4d_stream_t stream) {
    if (b < 0 || c < 0 || n <= 0 || npoints < 0 || nsample < 0 || !points || !idx || !out) return SG4D_EINVAL;
    const long long total = (long long)b * c * npoints * nsample;
    if (total == 0) return SG4D_OK;
    group_points_kernel<<<flat_grid(total), 256, 0, (cudaStream_t)stream>>>(total, c, n, npoints, nsample, points,
                                                                           idx, out);
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                      const int32_t *idx, float *grad_points, sg4d_stream_t stream) {
    if (b < 0 || c < 0 || n <= 0 || npoints < 0 || nsample < 0 || !grad_out || !idx || !grad_points)
        return SG4D_EINVAL;
    const long long total = (long long)b * c * npoints * nsample;
    if (total == 0) return SG4D_OK;
    group_points_grad_kernel<<<flat_grid(total), 256, 0, (cudaStream_t)stream>>>(total, c, n, npoints, nsample,
                                                                                grad_out, idx, grad_points);
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_group_rows(int b, int n, int m, int nsample, int c, int pts_stride, int feat_stride,
                               int feat_offset, int out_stride, int xyz_col0, const float *pts, const float *feats,
                               const float *centers, const int32_t *idx, float *out, sg4d_stream_t stream) {
    if (b < 0 || n <= 0 || m < 0 || nsample < 0 || c < 0 || pts_stride < 3 || out_stride < 3 + c || !pts ||
        (xyz_col0 != 0 && xyz_col0 != c) ||
        (!feats && c > 0) || !centers || !idx || !out)
        return SG4D_EINVAL;
    const long long rows = (long long)b * m * nsample;
    if (rows == 0) return SG4D_OK;
    if (!feats) feats = pts;
    cudaStream_t st = (cudaStream_t)stream;
    if (out_stride == 8 && xyz_col0 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        group_rows_narrow_kernel<8><<<flat_grid(rows), 256, 0, st>>>(rows, n, m, nsample, c, pts_stride, feat_stride,
                                                                    feat_offset, pts, feats, centers, idx, out);
    } else if (xyz_col0 == c && c > 0 && !(c & 3) && !(out_stride & 3) && !(feat_stride & 3) && !(feat_offset & 3) &&
               !((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(feats)) & 15)) {
        group_rows_wide4_kernel<<<flat_grid(rows * 16), 256, 0, st>>>(rows, n, m, nsample, c, pts_stride, feat_stride,
                                                                     feat_offset, out_stride, pts, feats, centers, idx, out);
    } else {
        group_rows_wide_kernel<<<flat_grid(rows * 32), 256, 0, st>>>(rows, n, m, nsample, c, pts_stride, feat_stride,
                                                                    feat_offset, out_stride, xyz_col0, pts, feats,
                                                                    centers, idx, out);
    }
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_group_rows_grad(int b, int n, int m, int nsample, int c, int out_stride, int gcol0, int accumulate,
                                    const float *grad_out, const int32_t *idx, const int32_t *cnt,
                                    float *grad_feats, sg4d_stream_t stream) {
    if (b < 0 || n <= 0 || m <= 0 || nsample <= 0 || c <= 0 || c > 256 || gcol0 < 0 || out_stride < gcol0 + c || !grad_out ||
        !idx || !cnt || !grad_feats)
        return SG4D_EINVAL;
    if (b == 0) return SG4D_OK;
    const size_t smem = ((size_t)m * nsample + m) * sizeof(int32_t);
    if (smem > 200 * 1024) return SG4D_EINVAL;
    const bool v4 = !((c | out_stride | gcol0) & 3) && !((reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(grad_feats)) & 15);
    cudaError_t e = cudaFuncSetAttribute(v4 ? group_rows_grad_kernel<true> : group_rows_grad_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return status_of(e);
    // enough CTAs to cover the chip a few times even for a handful of clouds
    int slices = 1;
    while ((long long)b * slices < 4LL * SG4D_NUM_SMS && slices * kGgWarps * 2 <= n) slices *= 2;
    if (v4)
        group_rows_grad_kernel<true><<<(unsigned)(b * slices), kGgWarps * 32, smem, (cudaStream_t)stream>>>(
            n, m, nsample, c, out_stride, gcol0, slices, accumulate, grad_out, idx, cnt, grad_feats);
    else
        group_rows_grad_kernel<false><<<(unsigned)(b * slices), kGgWarps * 32, smem, (cudaStream_t)stream>>>(
            n, m, nsample, c, out_stride, gcol0, slices, accumulate, grad_out, idx, cnt, grad_feats);
    return SG4D_LAUNCH_CHECK();
}

// G (b, n, c1) = per source point, the sum of dY1 = p1 .* dz1 - (q1 .* y1 + u1) over the grouped rows that reference it
// (deterministic gather, fixed order).  dz1 / y1 (b*m*nsample, c1), c1 in {64, 128}.  The feature gradient is G * W1[:, feats].
extern "C" int sg4d_group_rows_grad_dy(int b, int n, int m, int nsample, int c1, const float *y1, const float *dz1,
                                       const float *p1, const float *q1, const float *u1, const int32_t *idx,
                                       const int32_t *cnt, float *g_out, sg4d_stream_t stream) {
    if (b < 0 || n <= 0 || m <= 0 || nsample <= 0 || (c1 != 64 && c1 != 128) || !y1 || !dz1 || !p1 || !q1 || !u1 || !idx || !cnt ||
        !g_out || ((reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(dz1) | reinterpret_cast<uintptr_t>(g_out)) & 15))
        return SG4D_EINVAL;
    if (b == 0) return SG4D_OK;
    const size_t smem = ((size_t)m * nsample + m) * sizeof(int32_t);
    if (smem > 200 * 1024) return SG4D_EINVAL;
    cudaError_t e = cudaFuncSetAttribute(group_rows_grad_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return status_of(e);
    int slices = 1;
    while ((long long)b * slices < 4LL * SG4D_NUM_SMS && slices * kGgWarps * 2 <= n) slices *= 2;
    group_rows_grad_kernel<true, true><<<(unsigned)(b * slices), kGgWarps * 32, smem, (cudaStream_t)stream>>>(
        n, m, nsample, c1, c1, 0, slices, 0, dz1, idx, cnt, g_out, y1, p1, q1, u1);
    return SG4D_LAUNCH_CHECK();
}

// ================================================================================================
// SA2 through its linearity.  The first layer of a scale is linear in the grouped row x = [feats(i) | xyz(i) - centre_j], so
//     y1[r] = W1 x[r] = Z[i] - Cc[j],      Z = [feats | xyz] W1^T  per SOURCE POINT (b*n rows),  Cc = centre W1x^T per centre:
// one small GEMM over the b*n source points replaces the GEMM over all b*m*nsample grouped rows (24x fewer rows at SA2), and the
// forward pass of the layer becomes the gather below.  The backward pass uses the same linearity: with G[i] = sum of dY1 over the
// rows that reference point i (sg4d_group_rows_grad_dy) and H[j] = sum of dY1 over the rows of centre j (group_sum_dy_kernel),
//     dFeats = G W1f,    dW1 = G^T [feats | xyz] - H^T [0 | centre].
namespace sg4d {

// y1[r, :] = Z[cloud * n + idx[r], :] - Cc[r / nsample, :]; per-channel sum / sum of squares as fp64 partial pairs in the layout
// sg4d_bn_finalize reads (pair index % C1 = channel).  blockDim = 256 = (C1 / 4 column groups) x (row lanes); fixed grid.
__global__ void __launch_bounds__(256)
gather_y1_kernel(long long rows, int n, int m, int nsample, int c1, const float *__restrict__ Z, const float *__restrict__ Cc,
                 const int32_t *__restrict__ idx, float *__restrict__ y1, double *__restrict__ partial) {
    const int vec = c1 >> 2, lanes = 256 / vec;
    const int cg = threadIdx.x % vec, rl = threadIdx.x / vec;
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    double ds[4] = {0.0, 0.0, 0.0, 0.0}, dq[4] = {0.0, 0.0, 0.0, 0.0};
    int since = 0;
    // (group, cloud) with 32-bit divisions (groups < 2^31, checked by the launcher; a 64-bit division costs ~100 instructions);
    // the neighbour index of a thread's NEXT row is loaded one iteration ahead of the gather that depends on it
    // two rows per thread and iteration (r and r + stride): twice the bytes in flight per warp
    const long long stride = (long long)gridDim.x * lanes;
    long long r = (long long)blockIdx.x * lanes + rl;
    int ix0 = r < rows ? __ldg(idx + r) : 0, ix1 = r + stride < rows ? __ldg(idx + r + stride) : 0;
    const bool small = rows <= 0xffffffffLL;
    for (; r < rows; r += 2 * stride) {
        const long long r1 = r + stride, rn0 = r + 2 * stride, rn1 = r + 3 * stride;
        const bool two = r1 < rows;
        const int ixn0 = rn0 < rows ? __ldg(idx + rn0) : 0, ixn1 = rn1 < rows ? __ldg(idx + rn1) : 0;
        const uint32_t g0 = small ? (uint32_t)r / (uint32_t)nsample : (uint32_t)(r / nsample);
        const uint32_t g1 = two ? (small ? (uint32_t)r1 / (uint32_t)nsample : (uint32_t)(r1 / nsample)) : g0;
        const uint32_t cl0 = g0 / (uint32_t)m, cl1 = g1 / (uint32_t)m;
        const float4 z0 = __ldg(reinterpret_cast<const float4 *>(Z + ((size_t)cl0 * n + ix0) * c1) + cg);
        const float4 z1 = __ldg(reinterpret_cast<const float4 *>(Z + ((size_t)cl1 * n + ix1) * c1) + cg);
        const float4 c0 = __ldg(reinterpret_cast<const float4 *>(Cc + (size_t)g0 * c1) + cg);
        const float4 cc1 = __ldg(reinterpret_cast<const float4 *>(Cc + (size_t)g1 * c1) + cg);
        const float4 y = make_float4(z0.x - c0.x, z0.y - c0.y, z0.z - c0.z, z0.w - c0.w);
        reinterpret_cast<float4 *>(y1 + r * c1)[cg] = y;
        s[0] += y.x, s[1] += y.y, s[2] += y.z, s[3] += y.w;
        q[0] = fmaf(y.x, y.x, q[0]), q[1] = fmaf(y.y, y.y, q[1]), q[2] = fmaf(y.z, y.z, q[2]), q[3] = fmaf(y.w, y.w, q[3]);
        if (two) {
            const float4 w = make_float4(z1.x - cc1.x, z1.y - cc1.y, z1.z - cc1.z, z1.w - cc1.w);
            reinterpret_cast<float4 *>(y1 + r1 * c1)[cg] = w;
            s[0] += w.x, s[1] += w.y, s[2] += w.z, s[3] += w.w;
            q[0] = fmaf(w.x, w.x, q[0]), q[1] = fmaf(w.y, w.y, q[1]), q[2] = fmaf(w.z, w.z, q[2]), q[3] = fmaf(w.w, w.w, q[3]);
        }
        ix0 = ixn0, ix1 = ixn1;
        if (++since == 32) {   // fp32 over short runs (64 rows), fp64 across them
#pragma unroll
            for (int u = 0; u < 4; ++u) ds[u] += (double)s[u], dq[u] += (double)q[u], s[u] = 0.f, q[u] = 0.f;
            since = 0;
        }
    }
    // one pair per (block, channel): the row lanes of a block are folded in a fixed order through shared memory
    __shared__ double red[256][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) red[threadIdx.x][2 * u] = ds[u] + (double)s[u], red[threadIdx.x][2 * u + 1] = dq[u] + (double)q[u];
    __syncthreads();
    if (rl == 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double a = 0.0, b = 0.0;
            for (int l = 0; l < lanes; ++l) a += red[l * vec + cg][2 * u], b += red[l * vec + cg][2 * u + 1];
            const size_t i = (size_t)blockIdx.x * c1 + 4 * cg + u;
            partial[2 * i] = a;
            partial[2 * i + 1] = b;
        }
    }
}

// H[j, :] = sum over the nsample rows of centre j of dY1 = p .* dz1 - (q .* y1 + u).  One warp per (group, 128-column chunk).
__global__ void __launch_bounds__(256)
group_sum_dy_kernel(long long groups, int nsample, int c1, const float *__restrict__ y1, const float *__restrict__ dz1,
                    const float *__restrict__ p1, const float *__restrict__ q1, const float *__restrict__ u1, float *__restrict__ H) {
    const int lane = threadIdx.x & 31, vec = c1 >> 2;
    if (lane >= vec) return;
    const float4 pv = __ldg(reinterpret_cast<const float4 *>(p1) + lane), qv = __ldg(reinterpret_cast<const float4 *>(q1) + lane);
    const float4 uv = __ldg(reinterpret_cast<const float4 *>(u1) + lane);
    const long long nwarp = ((long long)gridDim.x * 256) >> 5;
    for (long long g = (blockIdx.x * 256LL + threadIdx.x) >> 5; g < groups; g += nwarp) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float *dz = dz1 + g * nsample * c1, *yy = y1 + g * nsample * c1;
        for (int k = 0; k < nsample; k += 2) {     // two rows in flight
            const float4 d0 = __ldg(reinterpret_cast<const float4 *>(dz + (size_t)k * c1) + lane);
            const float4 y0 = __ldg(reinterpret_cast<const float4 *>(yy + (size_t)k * c1) + lane);
            float4 d1 = make_float4(0.f, 0.f, 0.f, 0.f), y1v = d1;
            const bool two = k + 1 < nsample;
            if (two) {
                d1 = __ldg(reinterpret_cast<const float4 *>(dz + (size_t)(k + 1) * c1) + lane);
                y1v = __ldg(reinterpret_cast<const float4 *>(yy + (size_t)(k + 1) * c1) + lane);
            }
            acc.x += fmaf(d0.x, pv.x, -fmaf(y0.x, qv.x, uv.x)), acc.y += fmaf(d0.y, pv.y, -fmaf(y0.y, qv.y, uv.y));
            acc.z += fmaf(d0.z, pv.z, -fmaf(y0.z, qv.z, uv.z)), acc.w += fmaf(d0.w, pv.w, -fmaf(y0.w, qv.w, uv.w));
            if (two) {
                acc.x += fmaf(d1.x, pv.x, -fmaf(y1v.x, qv.x, uv.x)), acc.y += fmaf(d1.y, pv.y, -fmaf(y1v.y, qv.y, uv.y));
                acc.z += fmaf(d1.z, pv.z, -fmaf(y1v.z, qv.z, uv.z)), acc.w += fmaf(d1.w, pv.w, -fmaf(y1v.w, qv.w, uv.w));
            }
        }
        reinterpret_cast<float4 *>(H + g * c1)[lane] = acc;
    }
}

}  // namespace sg4d

extern "C" int sg4d_gather_y1_parts(int c1) { return SG4D_NUM_SMS * 8 * c1; }   // fp64 PAIRS: one per (block, channel)

extern "C" int sg4d_gather_y1(long long rows, int n, int m, int nsample, int c1, const float *z, const float *cc,
                              const int32_t *idx, float *y1, double *partial, sg4d_stream_t stream) {
    if (rows <= 0 || n <= 0 || m <= 0 || nsample <= 0 || (c1 != 64 && c1 != 128) || rows % ((long long)m * nsample) ||
        rows / nsample > 0x7fffffffLL || !z || !cc ||
        !idx || !y1 || !partial || ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(cc) | reinterpret_cast<uintptr_t>(y1)) & 15))
        return SG4D_EINVAL;
    gather_y1_kernel<<<SG4D_NUM_SMS * 8, 256, 0, (cudaStream_t)stream>>>(rows, n, m, nsample, c1, z, cc, idx, y1, partial);
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_group_sum_dy(long long groups, int nsample, int c1, const float *y1, const float *dz1, const float *p1,
                                 const float *q1, const float *u1, float *h, sg4d_stream_t stream) {
    if (groups <= 0 || nsample <= 0 || (c1 != 64 && c1 != 128) || !y1 || !dz1 || !p1 || !q1 || !u1 || !h ||
        ((reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(dz1) | reinterpret_cast<uintptr_t>(h)) & 15))
        return SG4D_EINVAL;
    long long blocks = (groups + 7) / 8;
    if (blocks > SG4D_NUM_SMS * 32) blocks = SG4D_NUM_SMS * 32;
    group_sum_dy_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(groups, nsample, c1, y1, dz1, p1, q1, u1, h);
    return SG4D_LAUNCH_CHECK();
}
