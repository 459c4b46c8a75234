// interpolate.cu -- three_nn / three_interpolate (+grad) for sm_100a: the feature-propagation ops that complete the
// reference's nine-function FFI.  They are NOT on the scene-graph hot path
// (PointNet2ClassificationMSG has no FP modules); PointnetFPModule users (semantic segmentation, the Group-Free-3D
// backbone) get them through the same C ABI.
//
// Replaces  three_nn_kernel               EXT/src/interpolate_gpu.cu:9-59   (one CTA per cloud, each thread re-reads
//                                                                           the whole `known` cloud from global memory)
//           three_interpolate_kernel      EXT/src/interpolate_gpu.cu:72-101
//           three_interpolate_grad_kernel EXT/src/interpolate_gpu.cu:116-143
// Here: grids sized by the work (not by the number of clouds), `known` staged through shared memory in coalesced
// tiles and read as broadcasts, outputs written with unit stride.
#include <math_constants.h>

#include "common.cuh"

namespace sg4d {

constexpr int kNnThreads = 256;
constexpr int kNnTile = 1024;   // known points per shared-memory tile (12 KB as SoA)

// The reference scans k = 0..m-1 with strict '<' against (best1, best2, best3) kept as doubles initialised to 1e40
// (interpolate_gpu.cu:28-50): the result is the three smallest (d, k) pairs in lexicographic order; a distance that
// is NaN or +inf never enters (inf < 1e40 is false), and missing entries keep index 0 / distance (float)1e40 = +inf.
// Distances use the SASS order of the reference build: t = dy*dy; t = fma(dx,dx,t); d = fma(dz,dz,t)  (sqdist3).
__global__ void __launch_bounds__(kNnThreads)
three_nn_kernel(int n, int m, const float *__restrict__ unknown, const float *__restrict__ known,
                float *__restrict__ dist2, int32_t *__restrict__ idx) {
    __shared__ float sx[kNnTile], sy[kNnTile], sz[kNnTile];
    const int cloud = blockIdx.y;
    const int j = blockIdx.x * kNnThreads + threadIdx.x;
    unknown += (size_t)cloud * n * 3;
    known += (size_t)cloud * m * 3;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (j < n) ux = __ldg(unknown + 3 * j), uy = __ldg(unknown + 3 * j + 1), uz = __ldg(unknown + 3 * j + 2);
    float b1 = CUDART_INF_F, b2 = CUDART_INF_F, b3 = CUDART_INF_F;   // +inf plays the role of 1e40: nothing finite is >= it
    int i1 = 0, i2 = 0, i3 = 0;
    for (int base = 0; base < m; base += kNnTile) {
        const int tn = min(kNnTile, m - base);
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * tn; i += kNnThreads) {   // coalesced copy, de-interleaved into SoA
            const float v = __ldg(known + (size_t)3 * base + i);
            const int p = i / 3, a = i - 3 * p;
            (a == 0 ? sx : (a == 1 ? sy : sz))[p] = v;
        }
        __syncthreads();
        for (int k = 0; k < tn; ++k) {
            const float d = sqdist3(ux - sx[k], uy - sy[k], uz - sz[k]);
            if (d < b1) {
                b3 = b2, i3 = i2, b2 = b1, i2 = i1, b1 = d, i1 = base + k;
            } else if (d < b2) {
                b3 = b2, i3 = i2, b2 = d, i2 = base + k;
            } else if (d < b3) {
                b3 = d, i3 = base + k;
            }
        }
    }
    if (j < n) {
        float *o = dist2 + ((size_t)cloud * n + j) * 3;
        int32_t *oi = idx + ((size_t)cloud * n + j) * 3;
        o[0] = b1, o[1] = b2, o[2] = b3;
        oi[0] = i1, oi[1] = i2, oi[2] = i3;
    }
}

// out[b,l,j] = p[b,l,i1]*w1 + p[b,l,i2]*w2 + p[b,l,i3]*w3 in the reference's contraction order
// (SASS of interpolate_gpu.cu:96-97: t = p2*w2; t = fma(p1,w1,t); out = fma(p3,w3,t))
__global__ void __launch_bounds__(256)
three_interpolate_kernel(long long total, int c, int m, int n, const float *__restrict__ points,
                         const int32_t *__restrict__ idx, const float *__restrict__ weight, float *__restrict__ out) {
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int j = (int)(e % n);
        const long long bl = e / n;          // b*c + l
        const long long bi = bl / c;
        const int32_t *ix = idx + (bi * n + j) * 3;
        const float *w = weight + (bi * n + j) * 3;
        const float *p = points + bl * m;
        float t = __fmul_rn(__ldg(p + __ldg(ix + 1)), __ldg(w + 1));
        t = __fmaf_rn(__ldg(p + __ldg(ix)), __ldg(w), t);
        out[e] = __fmaf_rn(__ldg(p + __ldg(ix + 2)), __ldg(w + 2), t);
    }
}

__global__ void __launch_bounds__(256)
three_interpolate_grad_kernel(long long total, int c, int n, int m, const float *__restrict__ grad_out,
                              const int32_t *__restrict__ idx, const float *__restrict__ weight,
                              float *__restrict__ grad_points) {
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int j = (int)(e % n);
        const long long bl = e / n;
        const long long bi = bl / c;
        const int32_t *ix = idx + (bi * n + j) * 3;
        const float *w = weight + (bi * n + j) * 3;
        const float g = __ldg(grad_out + e);
        float *gp = grad_points + bl * m;
        atomicAdd(gp + __ldg(ix), __fmul_rn(g, __ldg(w)));
        atomicAdd(gp + __ldg(ix + 1), __fmul_rn(g, __ldg(w + 1)));
        atomicAdd(gp + __ldg(ix + 2), __fmul_rn(g, __ldg(w + 2)));
    }
}

static unsigned interp_grid(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = (long long)SG4D_NUM_SMS * 32;
    return (unsigned)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace sg4d

using namespace sg4d;

extern "C" int sg4d_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int32_t *idx,
                             sg4d_stream_t stream) {
    if (b < 0 || n < 0 || m < 0 || b > 65535 || !unknown || !known || !dist2 || !idx) return SG4D_EINVAL;
    if (b == 0 || n == 0) return SG4D_OK;
    three_nn_kernel<<<dim3((n + kNnThreads - 1) / kNnThreads, b), kNnThreads, 0, (cudaStream_t)stream>>>(n, m, unknown, known,
                                                                                                        dist2, idx);
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_three_interpolate(int b, int c, int m, int n, const float *points, const int32_t *idx,
                                      const float *weight, float *out, sg4d_stream_t stream) {
    if (b < 0 || c < 0 || m <= 0 || n < 0 || !points || !idx || !weight || !out) return SG4D_EINVAL;
    const long long total = (long long)b * c * n;
    if (total == 0) return SG4D_OK;
    three_interpolate_kernel<<<interp_grid(total), 256, 0, (cudaStream_t)stream>>>(total, c, m, n, points, idx, weight, out);
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int32_t *idx,
                                           const float *weight, float *grad_points, sg4d_stream_t stream) {
    if (b < 0 || c < 0 || m <= 0 || n < 0 || !grad_out || !idx || !weight || !grad_points) return SG4D_EINVAL;
    const long long total = (long long)b * c * n;
    if (total == 0) return SG4D_OK;
    three_interpolate_grad_kernel<<<interp_grid(total), 256, 0, (cudaStream_t)stream>>>(total, c, n, m, grad_out, idx, weight,
                                                                                       grad_points);
    return SG4D_LAUNCH_CHECK();
}
