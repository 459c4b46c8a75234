// spatial.cu -- a per-cloud spatial index, and furthest point sampling / ball query on top of it (sm_100a).
//
// Replaces, for LARGE clouds (n >~ 10^4 points, the BASELINE shape n = 80 000), the same reference kernels as
// fps.cu and ball_query.cu:
//   furthest_point_sampling_kernel   EXT/src/sampling_gpu.cu:69-173   (m-1 rounds x a full pass over the cloud)
//   query_ball_point_kernel          EXT/src/ball_query_gpu.cu:9-44   (m centres x a full pass over the cloud)
// Both are brute force in the reference: 512 x 80 000 distance evaluations per cloud and per op.  The results
// they define are, however, LOCAL: a new FPS pick only lowers the running minimum distance of points that are
// closer to it than to every earlier pick, and a ball only contains nearby points.  This file keeps the results
// bit-identical and skips the work that provably cannot change them.
//
//   sg4d_spatial_index_build   bounding box (streaming pre-pass), then one CTA per cloud: 15-bit Morton cell per point
//                              -> counting sort in shared memory (32 768 bins) -> points rewritten cell by cell as
//                              float4 {x, y, z, original index k} (one 16-byte scattered store per point) plus the FPS
//                              running minimum t -> consecutive runs of 64 points form BUCKETS, each with an exact
//                              fp32 bounding box.
//   sg4d_fps_indexed           one CTA per cloud, all buckets' boxes + (max t, arg-max slot) in shared memory.
//                              Round: (A) lower bound of the distance from the new pick to every box, computed with
//                              the SAME rounded operations as the point distance (fp32 rounding is monotone, so
//                              bound <= d for every point of the box, bit-exactly); a bucket with bound >= max t
//                              cannot change and is skipped; (B) one warp per surviving bucket updates its 64
//                              points and its (max, arg-max); (C) block arg-max over the bucket records.
//                              On the benchmark clouds a round touches a median of 27 of 1250 buckets
//                              (tools/fps_prune_sim.py: 21 N point visits instead of 511 N).
//   sg4d_ball_query_indexed    warp per centre: box-sphere test against the largest radius, distance tests only in
//                              the surviving buckets, and a warp-level selection of the nsample SMALLEST original
//                              indices (the reference's "first nsample hits in index order").
//
// Tie-breaks are the reference's: FPS prefers, among equal maxima, the point minimising
// (bitreverse_L(k mod T), k div T) (see fps.cu); that pair is recomputed from k whenever two candidates are equal.
#include <math_constants.h>

#include "common.cuh"

namespace sg4d {

constexpr int kBucket = 64;               // points per bucket (2 per lane)
constexpr int kCellBits = 5;              // per axis
constexpr int kBins = 1 << (3 * kCellBits);
constexpr int kBuildThreads = 1024;
constexpr int kFpsThreads = 256;
constexpr int kFpsWarps = kFpsThreads / 32;
constexpr int kMaxBuckets = 6144;         // 32 B of shared memory per bucket in the FPS kernel
constexpr int kTodoCap = 1024;            // centres per cloud the ball query's todo list can hold (larger m: no list)

// ---- layout of one cloud's index inside the workspace (all arrays 128-byte aligned) ----
constexpr int kBoxSlices = 16;            // CTAs per cloud of the bounding-box pre-pass
struct IndexLayout {
    int nb;           // buckets
    long long np;     // padded points = nb * 64
    long long p4;     // float4 {x, y, z, original index k (bits)} per sorted point      (4 words each)
    long long t;      // FPS running minimum distance per sorted point
    long long lox, loy, loz, hix, hiy, hiz;    // bucket boxes
    long long ub0, slot0;                      // initial (max t, arg-max slot) per bucket
    long long cbox;                            // kBoxSlices x 6 partial cloud boxes (build scratch)
    long long words;                           // total 4-byte words per cloud
};

__host__ __device__ inline IndexLayout index_layout(int n) {
    IndexLayout L;
    L.nb = (n + kBucket - 1) / kBucket;
    L.np = (long long)L.nb * kBucket;
    const long long nbp = (L.nb + 31) / 32 * 32;
    long long o = 0;
    L.p4 = o, o += 4 * L.np;
    L.t = o, o += L.np;
    L.lox = o, o += nbp;
    L.loy = o, o += nbp;
    L.loz = o, o += nbp;
    L.hix = o, o += nbp;
    L.hiy = o, o += nbp;
    L.hiz = o, o += nbp;
    L.ub0 = o, o += nbp;
    L.slot0 = o, o += nbp;
    L.cbox = o, o += 32 * ((kBoxSlices * 6 + 31) / 32);
    L.words = o;
    return L;
}

// FPS priority of original index k for T = 2^L reference threads (smaller = preferred), 26 bits for n < 2^26
__device__ __forceinline__ unsigned fps_prio(unsigned k, int L) {
    return ((__brev(k & ((1u << L) - 1u)) >> (32 - L)) << 17) | (k >> L);
}

__device__ __forceinline__ unsigned spread5(unsigned v) {   // abcde -> a00b00c00d00e
    v = (v | (v << 8)) & 0x100fu;
    v = (v | (v << 4)) & 0x10c3u;
    v = (v | (v << 2)) & 0x1249u;
    return v;
}

__device__ __forceinline__ int cell_of(float v, float lo, float scale) {
    // NaN -> 0 (cvt.rzi of NaN is 0), +-inf clamp; any cell is valid -- the bucket boxes are computed from the points
    int q = __float2int_rd((v - lo) * scale);
    return min((1 << kCellBits) - 1, max(0, q));
}

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------ build
// pre-pass: partial bounding boxes, kBoxSlices CTAs per cloud (a plain streaming read at full memory parallelism)
__global__ void __launch_bounds__(256)
cloud_box_kernel(int n, int row_stride, const float *__restrict__ pts, float *__restrict__ ws, long long ws_words) {
    __shared__ float s_red[6][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x / kBoxSlices, slice = blockIdx.x % kBoxSlices;
    const IndexLayout lay = index_layout(n);
    pts += (size_t)cloud * n * row_stride;
    const int per = (n + kBoxSlices - 1) / kBoxSlices;
    const int k1 = min(n, (slice + 1) * per);
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int k = slice * per + tid; k < k1; k += 256) {   // fminf / fmaxf ignore NaNs
        const float *r = pts + (size_t)k * row_stride;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(r + a);
            lo[a] = fminf(lo[a], v), hi[a] = fmaxf(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = warp_min_f(lo[a]), hi[a] = warp_max_f(hi[a]);
        if (lane == 0) s_red[a][warp] = lo[a], s_red[3 + a][warp] = hi[a];
    }
    __syncthreads();
    if (tid < 6) {
        float v = s_red[tid][0];
        for (int w = 1; w < 8; ++w) v = tid < 3 ? fminf(v, s_red[tid][w]) : fmaxf(v, s_red[tid][w]);
        ws[(size_t)cloud * ws_words + lay.cbox + slice * 6 + tid] = v;
    }
}

__global__ void __launch_bounds__(kBuildThreads, 1)
spatial_build_kernel(int n, int row_stride, int L, const float *__restrict__ pts, float *__restrict__ ws,
                     long long ws_words) {
    extern __shared__ uint32_t s_bin[];   // kBins counters -> exclusive offsets -> scatter cursors
    __shared__ uint32_t s_scan[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const IndexLayout lay = index_layout(n);
    pts += (size_t)blockIdx.x * n * row_stride;
    ws += (size_t)blockIdx.x * ws_words;
    float4 *P4 = reinterpret_cast<float4 *>(ws + lay.p4);
    float *T = ws + lay.t;

    for (int i = tid; i < kBins; i += kBuildThreads) s_bin[i] = 0u;
    float lo[3], scale[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float l = CUDART_INF_F, h = -CUDART_INF_F;
        for (int sl = 0; sl < kBoxSlices; ++sl) {
            l = fminf(l, ws[lay.cbox + sl * 6 + a]);
            h = fmaxf(h, ws[lay.cbox + sl * 6 + 3 + a]);
        }
        const float ext = h - l;
        lo[a] = l;
        scale[a] = (ext > 0.f && ext < CUDART_INF_F) ? (float)(1 << kCellBits) / ext : 0.f;
    }
    auto bin_of = [&](float x, float y, float z) -> unsigned {
        return spread5((unsigned)cell_of(x, lo[0], scale[0])) | (spread5((unsigned)cell_of(y, lo[1], scale[1])) << 1) |
               (spread5((unsigned)cell_of(z, lo[2], scale[2])) << 2);
    };
    __syncthreads();

    // ---- pass 1: histogram of the Morton cells (4 points per thread in flight)
    constexpr int U = 4;
    for (int k0 = tid; k0 < n; k0 += U * kBuildThreads) {
        float x[U], y[U], z[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = min(k0 + u * kBuildThreads, n - 1);
            const float *r = pts + (size_t)k * row_stride;
            x[u] = __ldg(r), y[u] = __ldg(r + 1), z[u] = __ldg(r + 2);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (k0 + u * kBuildThreads < n) atomicAdd(&s_bin[bin_of(x[u], y[u], z[u])], 1u);
    }
    __syncthreads();
    // ---- exclusive scan over the bins: every thread owns kBins / 1024 = 32 consecutive bins
    {
        constexpr int per = kBins / kBuildThreads;
        uint32_t sum = 0;
        for (int i = 0; i < per; ++i) sum += s_bin[tid * per + i];
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        uint32_t wsum = s_scan[lane], winc = wsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += v;
        }
        const uint32_t wbase = __shfl_sync(0xffffffffu, winc - wsum, warp);
        uint32_t run = wbase + inc - sum;
        for (int i = 0; i < per; ++i) {
            const uint32_t c = s_bin[tid * per + i];
            s_bin[tid * per + i] = run;
            run += c;
        }
    }
    __syncthreads();
    // ---- pass 2: scatter into cell order, ONE 16-byte store per point (order inside a cell is arbitrary: no
    //      result depends on it)
    for (int k0 = tid; k0 < n; k0 += U * kBuildThreads) {
        float x[U], y[U], z[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = min(k0 + u * kBuildThreads, n - 1);
            const float *r = pts + (size_t)k * row_stride;
            x[u] = __ldg(r), y[u] = __ldg(r + 1), z[u] = __ldg(r + 2);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = k0 + u * kBuildThreads;
            if (k < n) {
                const uint32_t pos = atomicAdd(&s_bin[bin_of(x[u], y[u], z[u])], 1u);
                P4[pos] = make_float4(x[u], y[u], z[u], __int_as_float(k));
            }
        }
    }
    for (long long p = n + tid; p < lay.np; p += kBuildThreads)   // padding of the last bucket: never hit, never picked
        P4[p] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, __int_as_float(0));
    __syncthreads();   // global writes of this CTA are visible to the whole CTA after the barrier

    // ---- buckets: exact box of the real points, running minima, initial (max t, arg-max slot)
    float *LOX = ws + lay.lox, *LOY = ws + lay.loy, *LOZ = ws + lay.loz;
    float *HIX = ws + lay.hix, *HIY = ws + lay.hiy, *HIZ = ws + lay.hiz;
    float *UB0 = ws + lay.ub0;
    uint32_t *SLOT0 = reinterpret_cast<uint32_t *>(ws + lay.slot0);
    for (int b = warp; b < lay.nb; b += kBuildThreads / 32) {
        const long long p0 = (long long)b * kBucket + lane, p1 = p0 + 32;
        const bool v0 = p0 < n, v1 = p1 < n;
        const float4 a0 = P4[p0], a1 = P4[p1];
        const float qn = CUDART_NAN_F;
        const float x0 = v0 ? a0.x : qn, x1 = v1 ? a1.x : qn, y0 = v0 ? a0.y : qn, y1 = v1 ? a1.y : qn;
        const float z0 = v0 ? a0.z : qn, z1 = v1 ? a1.z : qn;
        const float blx = warp_min_f(fminf(x0, x1)), bly = warp_min_f(fminf(y0, y1)), blz = warp_min_f(fminf(z0, z1));
        const float bhx = warp_max_f(fmaxf(x0, x1)), bhy = warp_max_f(fmaxf(y0, y1)), bhz = warp_max_f(fmaxf(z0, z1));
        // sampling_gpu.cu:100-101: points with |p|^2 <= 1e-3 (compared in fp64) never take part
        const bool c0 = v0 && !((double)sqdist3(a0.x, a0.y, a0.z) <= 1e-3), c1 = v1 && !((double)sqdist3(a1.x, a1.y, a1.z) <= 1e-3);
        T[p0] = c0 ? 1e10f : -1.0f;
        T[p1] = c1 ? 1e10f : -1.0f;
        unsigned key = 0xffffffffu;
        if (c0) key = (fps_prio((unsigned)__float_as_int(a0.w), L) << 6) | (unsigned)lane;
        if (c1) key = min(key, (fps_prio((unsigned)__float_as_int(a1.w), L) << 6) | (unsigned)(32 + lane));
        key = redux_min_u32(key);
        if (lane == 0) {
            // a bucket whose points are all NaN keeps (+inf, -inf): its bound is +inf and it is always skipped
            LOX[b] = blx, LOY[b] = bly, LOZ[b] = blz, HIX[b] = bhx, HIY[b] = bhy, HIZ[b] = bhz;
            UB0[b] = key == 0xffffffffu ? -1.0f : 1e10f;
            SLOT0[b] = key & 63u;
        }
    }
}

// lower bound of sqdist3(p - c) over every point p of the box [lo, hi], evaluated with the same rounded
// operations: |fl(p - c)| >= e component-wise (rounding is monotone and symmetric), and fl(a*a), fl(fma(a,a,t))
// are monotone in |a| and t.
__device__ __forceinline__ float box_bound(float cx, float cy, float cz, float lx, float ly, float lz, float hx,
                                           float hy, float hz) {
    const float ex = fmaxf(0.f, fmaxf(lx - cx, cx - hx));
    const float ey = fmaxf(0.f, fmaxf(ly - cy, cy - hy));
    const float ez = fmaxf(0.f, fmaxf(lz - cz, cz - hz));
    return sqdist3(ex, ey, ez);
}

// ------------------------------------------------------------------------------------------------ FPS
struct FpsBucketOut {
    int maxbits;
    unsigned slot;
};

__global__ void __launch_bounds__(1024)
fps_indexed_kernel(int n, int m, int row_stride, int L, const float *__restrict__ pts, float *__restrict__ ws,
                   long long ws_words, int32_t *__restrict__ idxs, float *__restrict__ new_xyz) {
    extern __shared__ float s_dyn[];
    __shared__ int s_nact;
    __shared__ int s_wbest[32];
    __shared__ unsigned s_wmin[32], s_wmax[32];
    __shared__ unsigned s_tie_pr[32];
    __shared__ int s_tie_bucket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kFpsThreads = blockDim.x, kFpsWarps = kFpsThreads >> 5;   // 256 .. 1024 threads: more per cloud when clouds are few
    const unsigned lt = (1u << lane) - 1u;
    const IndexLayout lay = index_layout(n);
    const int nb = lay.nb, nbp = (nb + 31) / 32 * 32;
    pts += (size_t)blockIdx.x * n * row_stride;
    ws += (size_t)blockIdx.x * ws_words;
    idxs += (size_t)blockIdx.x * m;
    if (new_xyz) new_xyz += (size_t)blockIdx.x * m * 3;
    const float4 *P4 = reinterpret_cast<const float4 *>(ws + lay.p4);
    float *T = ws + lay.t;

    float *s_lox = s_dyn, *s_loy = s_lox + nbp, *s_loz = s_loy + nbp;
    float *s_hix = s_loz + nbp, *s_hiy = s_hix + nbp, *s_hiz = s_hiy + nbp;
    int *s_ub = reinterpret_cast<int *>(s_hiz + nbp);            // float bits of the bucket's max t (-1.0f: no candidate)
    unsigned *s_slot = reinterpret_cast<unsigned *>(s_ub + nbp);
    uint16_t *s_list = reinterpret_cast<uint16_t *>(s_slot + nbp);
    for (int i = tid; i < 8 * nbp; i += kFpsThreads) s_dyn[i] = ws[lay.lox + i];   // the 8 bucket arrays are contiguous
    if (tid == 0) s_nact = 0;

    const float p0x = __ldg(pts), p0y = __ldg(pts + 1), p0z = __ldg(pts + 2);
    float cx = p0x, cy = p0y, cz = p0z;
    if (tid == 0) {
        idxs[0] = 0;
        if (new_xyz) new_xyz[0] = p0x, new_xyz[1] = p0y, new_xyz[2] = p0z;
    }
    __syncthreads();

    // one bucket: update the 64 running minima against the new pick, return (max bits, arg-max slot)
    struct Bk {
        float4 a, b;
        float ta, tb;
    };
    auto load_bucket = [&](int b, Bk &v) {
        const long long p = (long long)b * kBucket + lane;
        v.a = P4[p], v.b = P4[p + 32], v.ta = T[p], v.tb = T[p + 32];
    };
    auto update_bucket = [&](int b, const Bk &v) {
        const long long p = (long long)b * kBucket + lane;
        const float n0 = fminf(sqdist3(v.a.x - cx, v.a.y - cy, v.a.z - cz), v.ta);
        const float n1 = fminf(sqdist3(v.b.x - cx, v.b.y - cy, v.b.z - cz), v.tb);
        if (n0 != v.ta) T[p] = n0;
        if (n1 != v.tb) T[p + 32] = n1;
        const int i0 = __float_as_int(n0), i1 = __float_as_int(n1);
        const int wmax = redux_max_s32(max(i0, i1));
        const unsigned m0 = __ballot_sync(0xffffffffu, i0 == wmax), m1 = __ballot_sync(0xffffffffu, i1 == wmax);
        unsigned slot;
        if (__popc(m0) + __popc(m1) == 1 || wmax < 0) {
            slot = m0 ? (unsigned)(__ffs(m0) - 1) : (unsigned)(31 + __ffs(m1));
        } else {   // equal maxima inside the bucket (duplicate points, lattices): the reference's priority decides
            unsigned key = 0xffffffffu;
            if (i0 == wmax) key = (fps_prio((unsigned)__float_as_int(v.a.w), L) << 6) | (unsigned)lane;
            if (i1 == wmax) key = min(key, (fps_prio((unsigned)__float_as_int(v.b.w), L) << 6) | (unsigned)(32 + lane));
            slot = redux_min_u32(key) & 63u;
        }
        if (lane == 0) s_ub[b] = wmax, s_slot[b] = slot;
    };

    bool dead = false;   // no candidate anywhere: the reference keeps emitting index 0
    for (int j = 1; j < m; ++j) {
        if (!dead) {
            // ---- (A) which buckets can change?
            for (int b0 = warp * 32; b0 < nbp; b0 += kFpsThreads) {
                const int b = b0 + lane;
                bool act = false;
                if (b < nb) {
                    const float bound = box_bound(cx, cy, cz, s_lox[b], s_loy[b], s_loz[b], s_hix[b], s_hiy[b], s_hiz[b]);
                    act = !(bound >= __int_as_float(s_ub[b]));   // NaN bound: not skipped
                }
                const unsigned mask = __ballot_sync(0xffffffffu, act);
                if (mask) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_nact, __popc(mask));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (act) s_list[base + __popc(mask & lt)] = (uint16_t)b;
                }
            }
        }
        __syncthreads();
        // ---- (B) update the surviving buckets, two per warp in flight
        const int nact = dead ? 0 : s_nact;
        for (int a = warp; a < nact; a += 2 * kFpsWarps) {
            const int b_a = s_list[a];
            const bool two = a + kFpsWarps < nact;
            const int b_b = two ? s_list[a + kFpsWarps] : b_a;
            Bk va, vb;
            load_bucket(b_a, va);
            if (two) load_bucket(b_b, vb);
            update_bucket(b_a, va);
            if (two) update_bucket(b_b, vb);
        }
        __syncthreads();
        if (tid == 0) s_nact = 0;
        // ---- (C) block arg-max over the bucket records; ties are detected as "more than one bucket at the maximum"
        int best = (int)0x80000000;
        unsigned bmin = 0xffffffffu, bmax = 0u;
        for (int b = tid; b < nb; b += kFpsThreads) {
            const int v = s_ub[b];
            if (v > best) best = v, bmin = (unsigned)b, bmax = (unsigned)b;
            else if (v == best) bmax = (unsigned)b;
        }
        {
            const int wb = redux_max_s32(best);
            const unsigned mn = redux_min_u32(best == wb ? bmin : 0xffffffffu), mx = redux_max_u32(best == wb ? bmax : 0u);
            if (lane == 0) s_wbest[warp] = wb, s_wmin[warp] = mn, s_wmax[warp] = mx;
        }
        __syncthreads();
        int M;
        unsigned Bmin, Bmax;
        {
            const int v = lane < kFpsWarps ? s_wbest[lane] : (int)0x80000000;
            M = redux_max_s32(v);
            const bool mine = lane < kFpsWarps && v == M;
            Bmin = redux_min_u32(mine ? s_wmin[lane] : 0xffffffffu);
            Bmax = redux_max_u32(mine ? s_wmax[lane] : 0u);
        }
        int k = 0;
        if (M < 0) {   // nothing qualifies: index 0 (sampling_gpu.cu: best = -1, besti = 0)
            dead = true;
            cx = p0x, cy = p0y, cz = p0z;
        } else {
            unsigned wbucket = Bmin;
            if (Bmin != Bmax) {   // several buckets share the maximum: compare the reference priorities of their winners
                unsigned pr = 0xffffffffu;
                int pb = -1;
                for (int b = tid; b < nb; b += kFpsThreads) {
                    if (s_ub[b] == M) {
                        const unsigned q = fps_prio((unsigned)__float_as_int(P4[(long long)b * kBucket + s_slot[b]].w), L);
                        if (q < pr) pr = q, pb = b;
                    }
                }
                const unsigned wpr = redux_min_u32(pr);
                if (lane == 0) s_tie_pr[warp] = wpr;
                __syncthreads();
                const unsigned gpr = redux_min_u32(lane < kFpsWarps ? s_tie_pr[lane] : 0xffffffffu);
                if (pr == gpr && pb >= 0) s_tie_bucket = pb;   // priorities are unique per point: exactly one writer
                __syncthreads();
                wbucket = (unsigned)s_tie_bucket;
            }
            const long long p = (long long)wbucket * kBucket + s_slot[wbucket];
            const float4 w = P4[p];
            cx = w.x, cy = w.y, cz = w.z;
            k = __float_as_int(w.w);
        }
        if (tid == 0) {
            idxs[j] = k;
            if (new_xyz) new_xyz[3 * j] = cx, new_xyz[3 * j + 1] = cy, new_xyz[3 * j + 2] = cz;
        }
    }
}

// ------------------------------------------------------------------------------------------------ ball query
constexpr int kBqiWarps = 8;
constexpr int kBqiCap = 128;     // per-warp, per-scale candidate buffer (>= 2 * max nsample and >= nsample + 64)
constexpr int kBqiMaxNs = 64;

struct BqiScale {
    float r2;
    int ns;
    int32_t *idx;
    int32_t *cnt;
};

// ascending bitonic sort of 128 uint32 in shared memory by one warp
__device__ __forceinline__ void warp_sort128(uint32_t *buf, int lane) {
#pragma unroll 1
    for (int k = 2; k <= kBqiCap; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int h = 0; h < kBqiCap / 64; ++h) {
                const int t = lane + 32 * h;                 // pair index 0..63
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool up = (i & k) == 0;
                const uint32_t a = buf[i], b = buf[l];
                if ((a > b) == up) buf[i] = b, buf[l] = a;
            }
            __syncwarp();
        }
    }
}

template <int NSC>
__global__ void __launch_bounds__(kBqiWarps * 32)
ball_query_indexed_kernel(int n, int m, int ctr_stride, int ctas_per_cloud, const float *__restrict__ centers,
                          const float *__restrict__ ws, long long ws_words, const int *__restrict__ todo_cnt,
                          const int *__restrict__ todo_list, int todo_cap, BqiScale s0, BqiScale s1) {
    extern __shared__ float s_box[];                        // 6 x nbp bucket boxes of this cloud
    __shared__ uint32_t s_buf[kBqiWarps][2][kBqiCap];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int cloud = blockIdx.x / ctas_per_cloud;
    const int slot = (blockIdx.x % ctas_per_cloud) * kBqiWarps + warp;
    const IndexLayout lay = index_layout(n);
    const int nb = lay.nb, nbp = (nb + 31) / 32 * 32;
    ws += (size_t)cloud * ws_words;
    const float4 *P4 = reinterpret_cast<const float4 *>(ws + lay.p4);
    BqiScale sc[2] = {s0, s1};

    // with a todo list (written by the prefix pass) the CTAs of a cloud are filled densely with the centres that
    // still need an answer; the remaining CTAs exit before touching the index
    int j = slot;
    bool todo = slot < m;
    if (todo_cnt) {
        todo = slot < __ldg(todo_cnt + cloud);
        j = todo ? __ldg(todo_list + (size_t)cloud * todo_cap + slot) : 0;
    }
    if (!__syncthreads_or(todo)) return;
    for (int i = tid; i < 6 * nbp; i += kBqiWarps * 32) s_box[i] = __ldg(ws + lay.lox + i);
    __syncthreads();
    if (!todo) return;

    const float *c = centers + ((size_t)cloud * m + j) * ctr_stride;
    const float qx = __ldg(c), qy = __ldg(c + 1), qz = __ldg(c + 2);
    const float *s_lox = s_box, *s_loy = s_lox + nbp, *s_loz = s_loy + nbp;
    const float *s_hix = s_loz + nbp, *s_hiy = s_hix + nbp, *s_hiz = s_hiy + nbp;
    float rmax = sc[0].r2;
    if (NSC == 2) rmax = fmaxf(rmax, sc[1].r2);

    int total[2] = {0, 0};                       // hits seen so far (all of them)
    int held[2] = {0, 0};                        // entries in the buffer
    uint32_t thr[2] = {0xffffffffu, 0xffffffffu};   // only indices below thr can still be among the nsample smallest
    uint32_t *buf[2] = {s_buf[warp][0], s_buf[warp][1]};

    auto compact = [&](int s) {   // keep the ns smallest of the buffer
        for (int i = held[s] + lane; i < kBqiCap; i += 32) buf[s][i] = 0xffffffffu;
        __syncwarp();
        warp_sort128(buf[s], lane);
        if (held[s] > sc[s].ns) {
            held[s] = sc[s].ns;
            thr[s] = buf[s][sc[s].ns - 1];
        }
        __syncwarp();
    };

    auto scan_bucket = [&](const float4 &a0, const float4 &a1) {
        const float d0 = sqdist3(qx - a0.x, qy - a0.y, qz - a0.z);
        const float d1 = sqdist3(qx - a1.x, qy - a1.y, qz - a1.z);
        const uint32_t k0 = (uint32_t)__float_as_int(a0.w), k1 = (uint32_t)__float_as_int(a1.w);
#pragma unroll
        for (int s = 0; s < NSC; ++s) {
            const bool h0 = d0 < sc[s].r2, h1 = d1 < sc[s].r2;   // ordered compares: NaN never hits
            const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
            if ((m0 | m1) == 0u) continue;
            total[s] += __popc(m0) + __popc(m1);
            const bool w0 = h0 && k0 < thr[s], w1 = h1 && k1 < thr[s];
            const unsigned e0 = __ballot_sync(0xffffffffu, w0), e1 = __ballot_sync(0xffffffffu, w1);
            if (w0) buf[s][held[s] + __popc(e0 & lt)] = k0;
            if (w1) buf[s][held[s] + __popc(e0) + __popc(e1 & lt)] = k1;
            held[s] += __popc(e0) + __popc(e1);
            __syncwarp();
            if (held[s] > kBqiCap - kBucket) compact(s);
        }
    };

    for (int b0 = 0; b0 < nbp; b0 += 32) {
        const int b = b0 + lane;
        bool near = false;
        if (b < nb) {
            // the query evaluates d = sqdist3(centre - p): same monotone-rounding argument as in the FPS kernel
            const float bound = box_bound(qx, qy, qz, s_lox[b], s_loy[b], s_loz[b], s_hix[b], s_hiy[b], s_hiz[b]);
            near = bound < rmax;
        }
        unsigned mask = __ballot_sync(0xffffffffu, near);
        while (mask) {   // two candidate buckets per trip: the second one's loads overlap the first one's selection
            const int b1 = b0 + __ffs(mask) - 1;
            mask &= mask - 1;
            const bool two = mask != 0u;
            const int b2 = two ? b0 + __ffs(mask) - 1 : b1;
            mask &= mask - 1;
            const long long p1 = (long long)b1 * kBucket + lane, p2 = (long long)b2 * kBucket + lane;
            const float4 a0 = __ldg(P4 + p1), a1 = __ldg(P4 + p1 + 32);
            float4 c0 = a0, c1 = a1;
            if (two) c0 = __ldg(P4 + p2), c1 = __ldg(P4 + p2 + 32);
            scan_bucket(a0, a1);
            if (two) scan_bucket(c0, c1);
        }
    }
#pragma unroll
    for (int s = 0; s < NSC; ++s) {
        compact(s);
        const int ns = sc[s].ns;
        const int cnt = min(total[s], ns);
        int32_t *row = sc[s].idx + ((size_t)cloud * m + j) * ns;
        // ascending hits, then repeats of the first one; a row without hits is all zeros (ball_query.cpp:19-21)
        const int32_t first = cnt > 0 ? (int32_t)buf[s][0] : 0;
        for (int i = lane; i < ns; i += 32) row[i] = i < cnt ? (int32_t)buf[s][i] : first;
        if (sc[s].cnt && lane == 0) sc[s].cnt[(size_t)cloud * m + j] = cnt;
    }
}

static int floor_log2_i(int v) {
    int l = 0;
    while ((2 << l) <= v) ++l;
    return l;
}

}  // namespace sg4d

using namespace sg4d;

// workspace = b per-cloud indices, then the ball query's todo lists: b counters followed by b x kTodoCap centre ids
extern "C" long long sg4d_spatial_index_bytes(int b, int n) {
    if (b < 0 || n <= 0) return 0;
    const IndexLayout lay = index_layout(n);
    return (long long)b * ((lay.words + 31) / 32 * 32) * 4 + ((long long)b * (1 + kTodoCap) + 32) * 4;
}

static long long ws_words_of(int n) { return (index_layout(n).words + 31) / 32 * 32; }

extern "C" int sg4d_spatial_index_supported(int n) {
    return n >= 1024 && index_layout(n).nb <= kMaxBuckets ? 1 : 0;
}

extern "C" int sg4d_spatial_index_build(int b, int n, int row_stride, const float *pts, void *index,
                                        sg4d_stream_t stream) {
    if (b < 0 || n <= 0 || row_stride < 3 || !pts || !index || !sg4d_spatial_index_supported(n)) return SG4D_EINVAL;
    if (b == 0) return SG4D_OK;
    int L = floor_log2_i(n);
    if (L > 9) L = 9;
    const size_t smem = (size_t)kBins * sizeof(uint32_t);
    cudaError_t e = cudaFuncSetAttribute(spatial_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return status_of(e);
    cloud_box_kernel<<<b * kBoxSlices, 256, 0, (cudaStream_t)stream>>>(n, row_stride, pts, (float *)index, ws_words_of(n));
    spatial_build_kernel<<<b, kBuildThreads, smem, (cudaStream_t)stream>>>(n, row_stride, L, pts, (float *)index,
                                                                          ws_words_of(n));
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_fps_indexed(int b, int n, int m, int row_stride, const float *pts, void *index, int32_t *idxs,
                                float *new_xyz, sg4d_stream_t stream) {
    if (b < 0 || n <= 0 || m < 0 || row_stride < 3 || !pts || !index || (!idxs && m > 0) ||
        !sg4d_spatial_index_supported(n))
        return SG4D_EINVAL;
    if (b == 0 || m == 0) return SG4D_OK;
    int L = floor_log2_i(n);
    if (L > 9) L = 9;
    const int nbp = (index_layout(n).nb + 31) / 32 * 32;
    const size_t smem = (size_t)nbp * (8 * 4 + 2);
    cudaError_t e = cudaFuncSetAttribute(fps_indexed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return status_of(e);
    // one CTA per cloud; with few clouds a CTA gets more warps (the early rounds touch every bucket)
    const int threads = b <= 160 ? 1024 : (b <= 320 ? 512 : kFpsThreads);
    fps_indexed_kernel<<<b, threads, smem, (cudaStream_t)stream>>>(n, m, row_stride, L, pts, (float *)index,
                                                                      ws_words_of(n), idxs, new_xyz);
    return SG4D_LAUNCH_CHECK();
}

// one pass (<= 2 scales) answered from the index; todo_cnt != nullptr: only the centres queued by the prefix pass
static int ball_query_indexed_pass(int b, int n, int m, int center_stride, int k, const float *radius,
                                   const int *nsample, const float *centers, const void *index, int32_t *const *idx,
                                   int32_t *const *cnt, const int *todo_cnt, const int *todo_list, sg4d_stream_t stream) {
    const int nbp = (index_layout(n).nb + 31) / 32 * 32;
    const size_t smem = (size_t)nbp * 6 * 4;
    const int cpc = (m + kBqiWarps - 1) / kBqiWarps;
    const long long grid = (long long)b * cpc;
    if (grid > 0x7fffffffLL) return SG4D_EINVAL;
    {
        const int s = 0;
        BqiScale sc[2] = {{0.f, 0, nullptr, nullptr}, {0.f, 0, nullptr, nullptr}};
        for (int u = 0; u < k; ++u) {
            if (nsample[s + u] <= 0 || nsample[s + u] > kBqiMaxNs || !idx[s + u]) return SG4D_EINVAL;
            const float r = radius[s + u];
            sc[u].r2 = r * r;
            sc[u].ns = nsample[s + u];
            sc[u].idx = idx[s + u];
            sc[u].cnt = cnt ? cnt[s + u] : nullptr;
        }
        cudaError_t e;
        if (k == 1) {
            e = cudaFuncSetAttribute(ball_query_indexed_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return status_of(e);
            ball_query_indexed_kernel<1><<<(unsigned)grid, kBqiWarps * 32, smem, (cudaStream_t)stream>>>(
                n, m, center_stride, cpc, centers, (const float *)index, ws_words_of(n), todo_cnt, todo_list, kTodoCap, sc[0], sc[1]);
        } else {
            e = cudaFuncSetAttribute(ball_query_indexed_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return status_of(e);
            ball_query_indexed_kernel<2><<<(unsigned)grid, kBqiWarps * 32, smem, (cudaStream_t)stream>>>(
                n, m, center_stride, cpc, centers, (const float *)index, ws_words_of(n), todo_cnt, todo_list, kTodoCap, sc[0], sc[1]);
        }
        const int st = SG4D_LAUNCH_CHECK();
        if (st != SG4D_OK) return st;
    }
    return SG4D_OK;
}

extern "C" int sg4d_ball_query_rows_indexed(int b, int n, int m, int row_stride, int center_stride, int nscales,
                                            const float *radius, const int *nsample, const float *centers,
                                            const float *pts, const void *index, int prefix, int32_t *const *idx,
                                            int32_t *const *cnt, sg4d_stream_t stream) {
    if (b < 0 || n <= 0 || m < 0 || row_stride < 3 || center_stride < 3 || nscales < 1 || nscales > SG4D_MAX_SCALES ||
        !radius || !nsample || !centers || !pts || !index || !idx || prefix < 0 || (prefix > 0 && !cnt) ||
        !sg4d_spatial_index_supported(n))
        return SG4D_EINVAL;
    if (b == 0 || m == 0) return SG4D_OK;
    if (prefix > n) prefix = n;
    if (m > kTodoCap) prefix = 0;   // no room for the todo list: answer every centre from the index
    int *todo_cnt = reinterpret_cast<int *>(const_cast<char *>(static_cast<const char *>(index)) + (size_t)b * ws_words_of(n) * 4);
    int *todo_list = todo_cnt + (b + 31) / 32 * 32;
    for (int s = 0; s < nscales; s += 2) {   // two radii per pass over the cloud
        const int k = nscales - s >= 2 ? 2 : 1;
        if (prefix > 0) {
            for (int u = 0; u < k; ++u)
                if (!cnt[s + u]) return SG4D_EINVAL;
            // pass 1: brute force over the first `prefix` points with early exit -- centres in dense regions find
            // their nsample lowest-index neighbours within a few hundred points and never need the index; the others
            // are queued per cloud
            cudaError_t e = cudaMemsetAsync(todo_cnt, 0, (size_t)b * sizeof(int), (cudaStream_t)stream);
            if (e != cudaSuccess) return status_of(e);
            int st = bq_launch(b, n, prefix, m, row_stride, center_stride, k, radius + s, nsample + s, centers, pts, idx + s,
                               cnt + s, (cudaStream_t)stream, todo_cnt, todo_list, kTodoCap);
            if (st != SG4D_OK) return st;
            if (prefix == n) continue;
            // pass 2: the queued centres are answered exactly from the spatial index
            st = ball_query_indexed_pass(b, n, m, center_stride, k, radius + s, nsample + s, centers, index, idx + s, cnt + s,
                                         todo_cnt, todo_list, stream);
            if (st != SG4D_OK) return st;
        } else {
            const int st = ball_query_indexed_pass(b, n, m, center_stride, k, radius + s, nsample + s, centers, index, idx + s,
                                                   cnt ? cnt + s : nullptr, nullptr, nullptr, stream);
            if (st != SG4D_OK) return st;
        }
    }
    return SG4D_OK;
}
