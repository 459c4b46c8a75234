// gnn.cu -- TripletGCN message gather and deterministic segmented scatter-add for sm_100a.
//
// Replaces, for SGH/model/gcns/network_TripletGCN.py:
//   * PyG MessagePassing.__collect__ (x_i = x[edge_index[1]], x_j = x[edge_index[0]]) followed by
//     torch.cat([x_i, edge_feature, x_j], dim=1)                       (:45-46)
//   * torch_scatter.scatter(x, index, dim=0, dim_size, reduce='add')   (:54-58), which lowers to
//     Tensor.scatter_add_ = fp32 atomics in arbitrary order.
// Here the scatter is a GATHER over destination-sorted edge lists (CSR built once per batch on the
// host side of the API): one thread per (node, column), edges of a node summed in ascending edge id,
// so results are bit-reproducible run to run.
#include "common.cuh"

namespace sg4d {

// out[e] = [ x[dst[e]] | edge_feat[e] | x[src[e]] ], float4 granularity (d, de multiples of 4)
__global__ void __launch_bounds__(256)
triplet_gather_kernel(long long total4, int d4, int de4, const float4 *__restrict__ x,
                      const float4 *__restrict__ ef, const int64_t *__restrict__ src,
                      const int64_t *__restrict__ dst, float4 *__restrict__ out) {
    const int w4 = 2 * d4 + de4;
    for (long long t = blockIdx.x * 256LL + threadIdx.x; t < total4; t += (long long)gridDim.x * 256) {
        const long long e = t / w4;
        const int col = (int)(t - e * w4);
        float4 v;
        if (col < d4)
            v = __ldg(x + __ldg(dst + e) * d4 + col);
        else if (col < d4 + de4)
            v = __ldg(ef + e * de4 + (col - d4));
        else
            v = __ldg(x + __ldg(src + e) * d4 + (col - d4 - de4));
        out[t] = v;
    }
}

// out[v, col] = sum_{p in [seg_ptr[v], seg_ptr[v+1])} ( src[order[p]*stride + col0 + col]
//                                                     (+ src[order[p]*stride + col1 + col]) )
// Warp-segmented: one warp per (node, 128-column chunk); the lanes own one float4 column group each, so every edge row is
// one coalesced 512-byte read, and the node's (destination-sorted) edge segment is walked in its fixed order -- the sum is
// deterministic, unlike torch_scatter's atomics (network_TripletGCN.py:57).  Two edges are in flight per step.
__global__ void __launch_bounds__(256)
segment_sum_kernel(int n_nodes, int d, long long stride, int col0, int col1, int has_second,
                   const float *__restrict__ src, const int32_t *__restrict__ order,
                   const int32_t *__restrict__ seg_ptr, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int chunks = (d + 127) / 128;
    const long long nwarp = ((long long)gridDim.x * 256) >> 5;
    for (long long w = (blockIdx.x * 256LL + threadIdx.x) >> 5; w < (long long)n_nodes * chunks; w += nwarp) {
        const int v = (int)(w / chunks), col = (int)(w % chunks) * 128 + lane * 4;
        if (col >= d) continue;
        const int p0 = __ldg(seg_ptr + v), p1 = __ldg(seg_ptr + v + 1);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        auto row_val = [&](int p) {
            const float *row = src + (long long)__ldg(order + p) * stride;
            float4 a = __ldg(reinterpret_cast<const float4 *>(row + col0 + col));
            if (has_second) {   // new_x_i + new_x_j (:48-51)
                const float4 b = __ldg(reinterpret_cast<const float4 *>(row + col1 + col));
                a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
            }
            return a;
        };
        int p = p0;
        for (; p + 1 < p1; p += 2) {
            const float4 a = row_val(p), b = row_val(p + 1);     // both loads issued before either is consumed
            acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
            acc.x += b.x, acc.y += b.y, acc.z += b.z, acc.w += b.w;
        }
        if (p < p1) {
            const float4 a = row_val(p);
            acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
        }
        *reinterpret_cast<float4 *>(out + (long long)v * d + col) = acc;
    }
}

static unsigned flat_grid(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = (long long)SG4D_NUM_SMS * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace sg4d

using namespace sg4d;

extern "C" int sg4d_triplet_gather(int64_t n_edges, int d, int de, const float *x, const float *edge_feat,
                                   const int64_t *src, const int64_t *dst, float *out,
                                   sg4d_stream_t stream) {
    if (n_edges < 0 || d <= 0 || de < 0 || (d & 3) || (de & 3) || !x || (!edge_feat && de) || !src || !dst || !out)
        return SG4D_EINVAL;
    if (n_edges == 0) return SG4D_OK;
    const long long total4 = (long long)n_edges * (2 * d + de) / 4;
    triplet_gather_kernel<<<flat_grid(total4), 256, 0, (cudaStream_t)stream>>>(
        total4, d / 4, de / 4, reinterpret_cast<const float4 *>(x), reinterpret_cast<const float4 *>(edge_feat),
        src, dst, reinterpret_cast<float4 *>(out));
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_segment_sum(int n_nodes, int d, int64_t src_stride, int col0, int col1, const float *src,
                                int has_second, const int32_t *order, const int32_t *seg_ptr, float *out,
                                sg4d_stream_t stream) {
    if (n_nodes < 0 || d <= 0 || (d & 3) || src_stride < d || (src_stride & 3) || col0 < 0 || col1 < 0 || (col0 & 3) || (col1 & 3) ||
        !src || !order || !seg_ptr || !out || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
        return SG4D_EINVAL;
    if (n_nodes == 0) return SG4D_OK;
    const long long warps = (long long)n_nodes * ((d + 127) / 128);
    segment_sum_kernel<<<flat_grid(warps * 32), 256, 0, (cudaStream_t)stream>>>(n_nodes, d, src_stride, col0, col1,
                                                                               has_second, src, order, seg_ptr, out);
    return SG4D_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------

extern "C" int sg4d_abi_version(void) { return SG4D_ABI_VERSION; }

extern "C" const char *sg4d_error_string(int status) {
    switch (status) {
        case SG4D_OK: return "ok";
        case SG4D_EINVAL: return "sg4d: invalid argument (shape, null pointer or unsupported size)";
        case SG4D_ENODEV: return "sg4d: current device is not an sm_100 (B200) GPU";
        default: return cudaGetErrorString(static_cast<cudaError_t>(status));
    }
}

extern "C" int sg4d_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    return major == 10 ? SG4D_OK : SG4D_ENODEV;
}
