// frontend.cu -- GPU crop / sample front-end: scene points + per-point object masks ->
// the per-object and per-edge (union bounding box) point clouds the encoders consume.
//
// Replaces the per-scene numpy / torch-CPU loop of
//   SGH/dataset/data_preparation_utils.py:104-125  object crops: members of mask i+1, bounding box +- padding, down-sample
//                                       :178-218  edge crops: points strictly inside the union of the two padded boxes, a 4th
//                                                 feature = 1 / 2 for points of the subject / object instance, down-sample
//                                       :12-18    zero_mean: centroid to the origin, scale to the unit sphere
//                                       :37-39    calculate_downsample_indices, the `replace=True` draw
// so that one scene (~5 MB) crosses PCIe instead of its 78 crops (171 MB).  The random draws are an INPUT (uniforms u in [0,1)):
// index = candidates[min(floor(u * len), len - 1)], candidates in ascending original point order (np.where), which makes the
// selection reproducible bit for bit by the CPU restatement oracle/frontend_ref.py.  The reference's other down-sampling branch
// (len >= target: open3d voxel trace + a draw without replacement, :41-49) depends on open3d, which is not available offline; this
// front-end draws with replacement there too (scope).
//
// Stable compaction = one warp per chunk of 1024 points walking 32 points at a time in order: lane o carries the running count of
// category o, __match_any_sync gives the in-group rank.  Two passes (count, scatter) around a per-category scan over chunks.
#include "common.cuh"

namespace sg4d {

constexpr int kFeChunk = 1024;     // points per warp-chunk
constexpr int kFeMaxCat = 32;      // categories (objects + 1) per pass: one lane each

// category of point i: objects pass -> masks[i] (0 = none); edge pass -> 1 if strictly inside the edge's union box else 0
struct FeSrc {
    const float *pts;      // (P, stride), xyz in columns 0..2
    const int32_t *masks;  // (P)
    int P, stride;
    const float *box;      // edge pass: (E, 6) union box {min xyz, max xyz}; objects pass: nullptr
};

__device__ __forceinline__ int fe_category(const FeSrc &s, int edge, int i) {
    if (!s.box) return s.masks[i];
    const float *b = s.box + (size_t)edge * 6;
    const float x = s.pts[(size_t)i * s.stride], y = s.pts[(size_t)i * s.stride + 1], z = s.pts[(size_t)i * s.stride + 2];
    return (x > b[0] && x < b[3] && y > b[1] && y < b[4] && z > b[2] && z < b[5]) ? 1 : 0;   // strict, like :197-199
}

// grid = (chunks, E or 1).  SCATTER = false: counts[(y * chunks + chunk) * ncat + cat] = members of cat in the chunk.
// SCATTER = true: list[(y * ncat_lists + cat - 1) * P + base + rank] = i for cat >= 1, base from the scanned counts.
template <bool SCATTER>
__global__ void __launch_bounds__(128)
fe_compact_kernel(FeSrc s, int ncat, int chunks, int *__restrict__ counts, int32_t *__restrict__ list) {
    const int lane = threadIdx.x & 31;
    const int chunk = blockIdx.x * 4 + (threadIdx.x >> 5), y = blockIdx.y;
    if (chunk >= chunks) return;
    const int i0 = chunk * kFeChunk, i1 = min(s.P, i0 + kFeChunk);
    int *cnt = counts + ((size_t)y * chunks + chunk) * ncat;
    int run = SCATTER ? (lane < ncat ? cnt[lane] : 0) : 0;    // scatter: exclusive base of my category in this chunk
    for (int i = i0 + lane; i - lane < i1; i += 32) {
        const int cat = i < i1 ? fe_category(s, y, i) : -1;
        const unsigned same = __match_any_sync(0xffffffffu, cat);
        if (SCATTER) {
            const int base = __shfl_sync(0xffffffffu, run, cat < 0 ? 0 : cat);
            if (cat >= 1) list[((size_t)y * (ncat - 1) + cat - 1) * s.P + base + __popc(same & ((1u << lane) - 1u))] = i;
        }
        // lane o adds the number of points of category o in this group
        for (int o = 0; o < ncat; ++o) {
            const unsigned b = __ballot_sync(0xffffffffu, cat == o);
            if (lane == o) run += __popc(b);
        }
    }
    if (!SCATTER && lane < ncat) cnt[lane] = run;
}

// exclusive scan over the chunks of every (y, category); totals[(y * ncat) + cat] = members.  One warp per (y, cat).
__global__ void fe_scan_kernel(int ncat, int chunks, int *__restrict__ counts, int *__restrict__ totals) {
    const int y = blockIdx.y, cat = blockIdx.x, lane = threadIdx.x;
    int carry = 0;
    for (int c0 = 0; c0 < chunks; c0 += 32) {
        const int c = c0 + lane;
        int v = c < chunks ? counts[((size_t)y * chunks + c) * ncat + cat] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (c < chunks) counts[((size_t)y * chunks + c) * ncat + cat] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) totals[y * ncat + cat] = carry;
}

// per object: bounding box of its members +- padding (:107-109).  One block per object; box (nobj, 6).
__global__ void __launch_bounds__(256)
fe_bbox_kernel(FeSrc s, const int32_t *__restrict__ list, const int *__restrict__ totals, int ncat, float padding,
               float *__restrict__ box) {
    const int o = blockIdx.x, n = totals[o + 1];
    const int32_t *mem = list + (size_t)o * s.P;
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int k = threadIdx.x; k < n; k += 256) {
        const float *p = s.pts + (size_t)mem[k] * s.stride;
#pragma unroll
        for (int a = 0; a < 3; ++a) lo[a] = fminf(lo[a], p[a]), hi[a] = fmaxf(hi[a], p[a]);
    }
    __shared__ float sl[8][3], sh[8][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], off));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], off));
        }
        if ((threadIdx.x & 31) == 0) sl[threadIdx.x >> 5][a] = lo[a], sh[threadIdx.x >> 5][a] = hi[a];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float l = sl[0][threadIdx.x], h = sh[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) l = fminf(l, sl[w][threadIdx.x]), h = fmaxf(h, sh[w][threadIdx.x]);
        box[o * 6 + threadIdx.x] = l - padding;          // one fp32 subtract / add, like np.min(...) - padding
        box[o * 6 + 3 + threadIdx.x] = h + padding;
    }
}

// union boxes of the edges (:192-196): element-wise min / max of the two padded object boxes
__global__ void fe_union_kernel(int E, const int64_t *__restrict__ edges, const float *__restrict__ obox, float *__restrict__ ebox) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int a = (int)edges[e], b = (int)edges[E + e];
    for (int k = 0; k < 3; ++k) {
        ebox[e * 6 + k] = fminf(obox[a * 6 + k], obox[b * 6 + k]);
        ebox[e * 6 + 3 + k] = fmaxf(obox[a * 6 + 3 + k], obox[b * 6 + 3 + k]);
    }
}

// the draw: index of sample j of cloud c in the scene, or -1 for an empty candidate list
__device__ __forceinline__ int fe_pick(const int32_t *__restrict__ list, int P, int len, const float *__restrict__ u, int c, int N, int j) {
    if (len <= 0) return -1;
    int k = (int)(u[(size_t)c * N + j] * (float)len);      // fp32 product, truncated: the restatement does the same
    k = k < len - 1 ? k : len - 1;
    return list[(size_t)c * P + k];
}

// Three passes over the DRAWS (the scene itself is a few MB and stays in L2; the 171 MB of crops are written exactly once):
//   PASS 0  per-block fp64 partial sums of the drawn xyz                      -> centroid (fe_mean_kernel)
//   PASS 1  per-cloud max squared norm of (xyz - centroid)                    -> unit-sphere scale
//   PASS 2  write [ (xyz - centroid) / dist | features | edge mask ] and the picked indices
// grid = (ceil(N / 256), clouds).
template <int PASS>
__global__ void __launch_bounds__(256)
fe_sample_kernel(FeSrc s, const int32_t *__restrict__ list, const int *__restrict__ totals, int tot_stride, const float *__restrict__ u,
                 int N, int fout, const int64_t *__restrict__ edges, int E, const float *__restrict__ mean,
                 unsigned *__restrict__ maxn2, double *__restrict__ part, float *__restrict__ out, int32_t *__restrict__ picked,
                 float *__restrict__ dist) {
    const int c = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
    const int len = totals[c * tot_stride + 1];
    // the draw (a dependent gather through the candidate list) is made once, in pass 0; passes 1 and 2 read it back coalesced
    int i = -1;
    if (j < N) {
        if (PASS == 0) picked[(size_t)c * N + j] = i = fe_pick(list, s.P, len, u, c, N, j);
        else i = picked[(size_t)c * N + j];
    }
    const float *p = s.pts + (size_t)(i < 0 ? 0 : i) * s.stride;
    if (PASS == 0) {
        double sx = i >= 0 ? (double)p[0] : 0.0, sy = i >= 0 ? (double)p[1] : 0.0, sz = i >= 0 ? (double)p[2] : 0.0;
        __shared__ double sh[8][3];
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, off), sy += __shfl_xor_sync(0xffffffffu, sy, off), sz += __shfl_xor_sync(0xffffffffu, sz, off);
        }
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][0] = sx, sh[threadIdx.x >> 5][1] = sy, sh[threadIdx.x >> 5][2] = sz;
        __syncthreads();
        if (threadIdx.x < 3) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
            part[((size_t)c * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = t;
        }
        return;
    }
    const float m0 = mean[c * 3], m1 = mean[c * 3 + 1], m2 = mean[c * 3 + 2];
    const float x = i >= 0 ? p[0] - m0 : 0.f, y = i >= 0 ? p[1] - m1 : 0.f, z = i >= 0 ? p[2] - m2 : 0.f;   // point -= mean (:14)
    if (PASS == 1) {
        float n2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));    // pow(2).sum(1), left to right
#pragma unroll
        for (int off = 16; off; off >>= 1) n2 = fmaxf(n2, __shfl_xor_sync(0xffffffffu, n2, off));
        if ((threadIdx.x & 31) == 0) atomicMax(maxn2 + c, __float_as_uint(n2));   // non-negative floats order like their bits
        return;
    }
    // sqrt is monotone: max over rows of sqrt(n2) = sqrt(max n2) (:16)
    const float d = sqrtf(__uint_as_float(maxn2[c]));
    if (blockIdx.x == 0 && threadIdx.x == 0) dist[c] = d;
    // rows of fout <= 8 floats are staged in shared memory and leave as consecutive 4-byte stores of the block's contiguous
    // output range (a thread writing its own 6 / 7 floats would touch every 32-byte sector seven times)
    __shared__ float so[256 * 8];
    const bool staged = fout <= 8;
    float *o = staged ? so + threadIdx.x * fout : out + ((size_t)c * N + j) * fout;
    if (j < N) {
        if (i < 0) {
            for (int a = 0; a < fout; ++a) o[a] = 0.f;
        } else {
            o[0] = __fdiv_rn(x, d), o[1] = __fdiv_rn(y, d), o[2] = __fdiv_rn(z, d);             // point /= furthest_distance (:17)
            for (int a = 3; a < s.stride; ++a) o[a] = p[a];
            if (edges) {      // 4th feature: 1 = point of the subject instance, 2 = of the object instance (:188-190)
                const int m = s.masks[i];
                o[s.stride] = m == (int)edges[c] + 1 ? 1.f : (m == (int)edges[E + c] + 1 ? 2.f : 0.f);
            }
        }
    }
    if (staged) {
        __syncthreads();
        const int j0 = blockIdx.x * 256, nvalid = min(256, N - j0) * fout;
        float *dst = out + ((size_t)c * N + j0) * fout;
        for (int t = threadIdx.x; t < nvalid; t += 256) dst[t] = so[t];
    }
}

// centroid of every cloud: fp64 sum of the per-block partials in a fixed order (one warp per cloud), rounded once to fp32
__global__ void fe_mean_kernel(int N, int nparts, const double *__restrict__ part, float *__restrict__ mean) {
    const int c = blockIdx.x, lane = threadIdx.x;
    double t[3] = {0.0, 0.0, 0.0};
    for (int q = lane; q < nparts; q += 32)
#pragma unroll
        for (int a = 0; a < 3; ++a) t[a] += part[((size_t)c * nparts + q) * 3 + a];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int off = 16; off; off >>= 1) t[a] += __shfl_xor_sync(0xffffffffu, t[a], off);
        if (lane == 0) mean[c * 3 + a] = (float)(t[a] / (double)N);
    }
}

}  // namespace sg4d

using namespace sg4d;

extern "C" long long sg4d_frontend_workspace_bytes(int P, int nobj, int E) {
    const long long chunks = (P + kFeChunk - 1) / kFeChunk;
    const long long a = chunks * (nobj + 1), b = chunks * 2LL * E;
    return (a > b ? a : b) * 4 + 256;      // per-chunk category counts (objects pass, then reused by the edges pass)
}

// Stage 1 (objects): member lists (ascending point index) and padded boxes.  ws layout is private to the library;
// totals_out (nobj + 1) ints: [0] = unlabelled points, [i + 1] = members of object i.
extern "C" int sg4d_frontend_objects(int P, int stride, int nobj, const float *pts, const int32_t *masks, float padding, void *ws,
                                     int32_t *obj_list, int *totals_out, float *obj_box, sg4d_stream_t stream) {
    if (P <= 0 || stride < 3 || nobj <= 0 || nobj + 1 > kFeMaxCat || !pts || !masks || !ws || !obj_list || !totals_out || !obj_box)
        return SG4D_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = (P + kFeChunk - 1) / kFeChunk, ncat = nobj + 1;
    int *counts = reinterpret_cast<int *>(ws);
    FeSrc s{pts, masks, P, stride, nullptr};
    fe_compact_kernel<false><<<dim3((chunks + 3) / 4, 1), 128, 0, st>>>(s, ncat, chunks, counts, nullptr);
    fe_scan_kernel<<<dim3(ncat, 1), 32, 0, st>>>(ncat, chunks, counts, totals_out);
    fe_compact_kernel<true><<<dim3((chunks + 3) / 4, 1), 128, 0, st>>>(s, ncat, chunks, counts, obj_list);
    fe_bbox_kernel<<<nobj, 256, 0, st>>>(s, obj_list, totals_out, ncat, padding, obj_box);
    return SG4D_LAUNCH_CHECK();
}

// Stage 2 (edges): union boxes and the lists of points strictly inside them.  edges (2, E) int64 object indices;
// edge_totals (E, 2) ints: [e][1] = points inside the box of edge e.
extern "C" int sg4d_frontend_edges(int P, int stride, int E, const float *pts, const int32_t *masks, const int64_t *edges,
                                   const float *obj_box, void *ws, int32_t *edge_list, int *edge_totals, float *edge_box,
                                   sg4d_stream_t stream) {
    if (P <= 0 || stride < 3 || E <= 0 || !pts || !masks || !edges || !obj_box || !ws || !edge_list || !edge_totals || !edge_box)
        return SG4D_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = (P + kFeChunk - 1) / kFeChunk;
    int *counts = reinterpret_cast<int *>(ws);
    fe_union_kernel<<<(E + 63) / 64, 64, 0, st>>>(E, edges, obj_box, edge_box);
    FeSrc s{pts, masks, P, stride, edge_box};
    fe_compact_kernel<false><<<dim3((chunks + 3) / 4, E), 128, 0, st>>>(s, 2, chunks, counts, nullptr);
    fe_scan_kernel<<<dim3(2, E), 32, 0, st>>>(2, chunks, counts, edge_totals);
    fe_compact_kernel<true><<<dim3((chunks + 3) / 4, E), 128, 0, st>>>(s, 2, chunks, counts, edge_list);
    return SG4D_LAUNCH_CHECK();
}

// Stage 3: draw n points per cloud from its list (u: (clouds, n) uniforms in [0,1)), gather [xyz | features | edge mask],
// zero_mean.  edges == NULL: object clouds (fout = stride, totals = (nobj + 1), entry c + 1); else edge clouds
// (fout = stride + 1, totals = (E, 2), entry [c][1]).  out (clouds, n, fout); picked (clouds, n) original indices or NULL;
// mean (clouds, 3), dist (clouds); scratch: clouds * (ceil(n / 256) * 3 * 8 + 4 + 4 n) bytes.
extern "C" int sg4d_frontend_sample(int P, int stride, int clouds, int n, const float *pts, const int32_t *masks,
                                    const int32_t *list, const int *totals, const int64_t *edges, const float *u, float *out,
                                    int32_t *picked, float *mean, float *dist, void *scratch, sg4d_stream_t stream) {
    if (P <= 0 || stride < 3 || clouds <= 0 || n <= 0 || !pts || !masks || !list || !totals || !u || !out || !mean || !dist || !scratch)
        return SG4D_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = (n + 255) / 256, fout = stride + (edges ? 1 : 0);
    double *part = reinterpret_cast<double *>(scratch);
    unsigned *maxn2 = reinterpret_cast<unsigned *>(part + (size_t)clouds * nb * 3);
    if (!picked) picked = reinterpret_cast<int32_t *>(maxn2 + clouds);     // the draws live in the scratch area
    cudaError_t e = cudaMemsetAsync(maxn2, 0, (size_t)clouds * 4, st);
    if (e != cudaSuccess) return status_of(e);
    FeSrc s{pts, masks, P, stride, nullptr};
    const int ts = edges ? 2 : 1;
    const int *tot = edges ? totals : totals;      // objects: entry c + 1 = totals[c * 1 + 1]; edges: totals[c * 2 + 1]
    fe_sample_kernel<0><<<dim3(nb, clouds), 256, 0, st>>>(s, list, tot, ts, u, n, fout, edges, clouds, mean, maxn2, part, out, picked, dist);
    fe_mean_kernel<<<clouds, 32, 0, st>>>(n, nb, part, mean);
    fe_sample_kernel<1><<<dim3(nb, clouds), 256, 0, st>>>(s, list, tot, ts, u, n, fout, edges, clouds, mean, maxn2, part, out, picked, dist);
    fe_sample_kernel<2><<<dim3(nb, clouds), 256, 0, st>>>(s, list, tot, ts, u, n, fout, edges, clouds, mean, maxn2, part, out, picked, dist);
    return SG4D_LAUNCH_CHECK();
}
