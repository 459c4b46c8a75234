// fps.cu -- furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel (+ the gather of the picked xyz) of the reference:
//   EXT/src/sampling_gpu.cu:69-173 (kernel), :175-229 (launch), EXT/src/sampling.cpp:66-87 (host).
//
// Design (B200-first, not a translation):
//  * The reference keeps the running min-distance array `temp` in GLOBAL memory and re-streams
//    xyz + temp from L2/HBM in each of the m-1 dependent rounds (1.28 MB per round per cloud at
//    n = 80 000).  Here a cloud lives ON CHIP for the whole call: every thread owns PPT points
//    (x, y, z, temp in registers), a CTA owns NT*PPT points and a thread-block CLUSTER of up to
//    16 CTAs owns one cloud (80 000 points = 8 CTAs x 1024 threads x 10 points).  HBM traffic is
//    the algorithmic minimum: read xyz once, write idx (+ the picked xyz) once.
//  * One round = per-thread scan of its registers, a 2-instruction `redux.sync` warp arg-max on a
//    packed (distance bits, priority) key, one __syncthreads, and -- for clusters -- one DSMEM
//    exchange of the per-CTA winners (key + winner coordinates, so no second trip is needed)
//    followed by one barrier.cluster.
//  * Bit-exactness.  Distances use the reference's FMA order (common.cuh: sqdist3) and the same
//    fp64 `|p|^2 <= 1e-3` skip rule.  The reference's winner among EQUAL maxima is decided by its
//    strided per-thread scan (strict '>') and its shared-memory tree that keeps the lower slot:
//    the point k minimising (bitreverse_L(k mod T), k div T) with T = opt_n_threads(n) = 2^L
//    threads (cuda_utils.h:15-19).  We encode exactly that pair as a 32-bit priority, so the
//    arg-max is order-independent and any thread/CTA layout reproduces the reference's pick.
#include "common.cuh"

namespace sg4d {

struct __align__(16) FpsEntry {  // one candidate: packed key + its coordinates
    int hi;       // float bits of the min-distance (as signed int: -1.0f < every valid value)
    unsigned lo;  // ~priority  (larger = preferred on ties)
    float x, y;
    float z;
    unsigned pad0, pad1, pad2;
};

__device__ __forceinline__ void entry_store(FpsEntry *e, int hi, unsigned lo, float x, float y,
                                            float z) {
    uint4 a = make_uint4((unsigned)hi, lo, __float_as_uint(x), __float_as_uint(y));
    uint4 b = make_uint4(__float_as_uint(z), 0u, 0u, 0u);
    reinterpret_cast<uint4 *>(e)[0] = a;
    reinterpret_cast<uint4 *>(e)[1] = b;
}

// Reduce one entry per lane (lanes >= count hold "nothing") to the warp-wide best lane.
// Returns the lane id that holds the best entry.
__device__ __forceinline__ int warp_best_lane(int hi, unsigned lo) {
    const int whi = redux_max_s32(hi);
    const unsigned mylo = (hi == whi) ? lo : 0u;
    const unsigned wlo = redux_max_u32(mylo);
    const unsigned ball = __ballot_sync(0xffffffffu, hi == whi && mylo == wlo);
    return __ffs(ball) - 1;
}

constexpr int kMaxWarps = 32;
constexpr int kMaxCluster = 16;

// dynamic smem layout: [3][PPT*NT] floats (copy of the owned xyz, for the winner look-up)
template <int PPT>
__global__ void __launch_bounds__(1024, 1)
fps_onchip_kernel(int n, int m, int row_stride, int L, const float *__restrict__ pts,
                  int32_t *__restrict__ idxs, float *__restrict__ new_xyz) {
    extern __shared__ float s_xyz[];
    __shared__ FpsEntry s_warp[2][kMaxWarps];      // per-warp winners, double-buffered by round parity
    __shared__ FpsEntry s_cta[2][kMaxCluster];     // per-CTA winners of the cluster (written via DSMEM)

    const int NT = blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    const unsigned CS = cluster_nctarank(), rank = cluster_ctarank();
    const int cloud = blockIdx.x / CS;
    const int G = NT * (int)CS;              // threads cooperating on this cloud
    const int g = (int)rank * NT + tid;      // my id among them
    const unsigned T1 = (1u << L) - 1u;      // T - 1  (T = reference thread count)

    pts += (size_t)cloud * n * row_stride;
    idxs += (size_t)cloud * m;
    if (new_xyz) new_xyz += (size_t)cloud * m * 3;

    // point k = g + G*p  ->  k mod T is the same for all my points (G is a multiple of T) and
    // k div T grows with p, so scanning p upwards with strict '>' visits my points in priority order.
    const unsigned brev_r = __brev((unsigned)g & T1);
    const unsigned q0 = (unsigned)g >> L, qstep = (unsigned)G >> L;

    float x[PPT], y[PPT], z[PPT], t[PPT];
    float *sx = s_xyz, *sy = s_xyz + PPT * NT, *sz = s_xyz + 2 * PPT * NT;
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const int k = g + G * p;
        float px = 0.f, py = 0.f, pz = 0.f, tt = -1.0f;  // -1: never a candidate, never updated
        if (k < n) {
            const float *r = pts + (size_t)k * row_stride;
            px = __ldg(r), py = __ldg(r + 1), pz = __ldg(r + 2);
            const float mag = sqdist3(px, py, pz);
            if (!((double)mag <= 1e-3)) tt = 1e10f;  // sampling_gpu.cu:100-101 (NaN magnitudes are kept)
        }
        x[p] = px, y[p] = py, z[p] = pz, t[p] = tt;
        sx[p * NT + tid] = px, sy[p * NT + tid] = py, sz[p * NT + tid] = pz;
    }
    const float p0x = __ldg(pts), p0y = __ldg(pts + 1), p0z = __ldg(pts + 2);
    float cx = p0x, cy = p0y, cz = p0z;  // coordinates of the last pick
    if (g == 0) {
        idxs[0] = 0;
        if (new_xyz) new_xyz[0] = p0x, new_xyz[1] = p0y, new_xyz[2] = p0z;
    }
    if (CS > 1) {  // all CTAs of the cluster must have started before anyone writes their smem
        cluster_arrive_release();
        cluster_wait_acquire();
    }

    for (int j = 1; j < m; ++j) {
        const int par = j & 1;
        float best = -1.0f;
        int bp = 0;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const float d = sqdist3(x[p] - cx, y[p] - cy, z[p] - cz);
            const float d2 = fminf(d, t[p]);
            t[p] = d2;
            if (d2 > best) best = d2, bp = p;
        }
        {
            const int hi = __float_as_int(best);
            const unsigned lo = ~(brev_r | (q0 + qstep * (unsigned)bp));
            const int wl = warp_best_lane(hi, lo);
            if (lane == wl) {
                const int s = bp * NT + tid;
                entry_store(&s_warp[par][warp], hi, lo, sx[s], sy[s], sz[s]);
            }
        }
        __syncthreads();

        int whi;
        unsigned wlo;
        if (CS == 1) {
            // every warp reduces the per-warp winners redundantly: no second barrier needed
            int hi = __float_as_int(-1.0f);
            unsigned lo = 0u;
            if (lane < nwarps) hi = s_warp[par][lane].hi, lo = s_warp[par][lane].lo;
            const int wl = warp_best_lane(hi, lo);
            const FpsEntry *e = &s_warp[par][wl];
            whi = e->hi, wlo = e->lo, cx = e->x, cy = e->y, cz = e->z;
        } else {
            if (warp < (int)CS) {  // warp w forwards this CTA's winner to CTA w of the cluster
                int hi = __float_as_int(-1.0f);
                unsigned lo = 0u;
                if (lane < nwarps) hi = s_warp[par][lane].hi, lo = s_warp[par][lane].lo;
                const int wl = warp_best_lane(hi, lo);
                if (lane < 2) {
                    const uint4 v = reinterpret_cast<const uint4 *>(&s_warp[par][wl])[lane];
                    const uint32_t dst = mapa_shared(smem_u32(&s_cta[par][rank]), (unsigned)warp) + 16u * lane;
                    st_cluster_v4(dst, v.x, v.y, v.z, v.w);
                }
            }
            cluster_arrive_release();
            cluster_wait_acquire();
            int hi = __float_as_int(-1.0f);
            unsigned lo = 0u;
            if (lane < (int)CS) hi = s_cta[par][lane].hi, lo = s_cta[par][lane].lo;
            const int wl = warp_best_lane(hi, lo);
            const FpsEntry *e = &s_cta[par][wl];
            whi = e->hi, wlo = e->lo, cx = e->x, cy = e->y, cz = e->z;
        }

        int k = 0;
        if (whi >= 0) {  // a valid candidate exists; otherwise the reference falls back to index 0
            const unsigned prio = ~wlo;
            const unsigned topmask = L ? (0xffffffffu << (32 - L)) : 0u;
            k = (int)(((prio & ~topmask) << L) | __brev(prio & topmask));
        } else {
            cx = p0x, cy = p0y, cz = p0z;
        }
        if (g == 0) {
            idxs[j] = k;
            if (new_xyz) new_xyz[3 * j] = cx, new_xyz[3 * j + 1] = cy, new_xyz[3 * j + 2] = cz;
        }
    }
    if (CS > 1) {  // no CTA may exit while a peer can still write into its shared memory
        cluster_arrive_release();
        cluster_wait_acquire();
    }
}

// Fallback for clouds that do not fit on chip (n > 16 x 1024 x 12): one CTA per cloud, temp in global
// memory like the reference, same packed-key arg-max.
__global__ void __launch_bounds__(1024, 1)
fps_stream_kernel(int n, int m, int row_stride, int L, const float *__restrict__ pts,
                  float *__restrict__ temp, int32_t *__restrict__ idxs, float *__restrict__ new_xyz) {
    __shared__ FpsEntry s_warp[2][kMaxWarps];
    const int NT = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    const int cloud = blockIdx.x;
    const unsigned T1 = (1u << L) - 1u;
    pts += (size_t)cloud * n * row_stride;
    temp += (size_t)cloud * n;
    idxs += (size_t)cloud * m;
    if (new_xyz) new_xyz += (size_t)cloud * m * 3;
    const unsigned brev_r = __brev((unsigned)tid & T1);
    const float p0x = __ldg(pts), p0y = __ldg(pts + 1), p0z = __ldg(pts + 2);
    float cx = p0x, cy = p0y, cz = p0z;
    if (tid == 0) {
        idxs[0] = 0;
        if (new_xyz) new_xyz[0] = p0x, new_xyz[1] = p0y, new_xyz[2] = p0z;
    }
    for (int k = tid; k < n; k += NT) temp[k] = 1e10f;  // sampling.cpp:74-76 (each k is owned by one thread)
    for (int j = 1; j < m; ++j) {
        const int par = j & 1;
        float best = -1.0f, bx = 0.f, by = 0.f, bz = 0.f;
        int bk = 0;
        for (int k = tid; k < n; k += NT) {  // NT is a multiple of T: same residue, ascending k div T
            const float *r = pts + (size_t)k * row_stride;
            const float px = __ldg(r), py = __ldg(r + 1), pz = __ldg(r + 2);
            const float mag = sqdist3(px, py, pz);
            if ((double)mag <= 1e-3) continue;
            const float d = sqdist3(px - cx, py - cy, pz - cz);
            const float d2 = fminf(d, temp[k]);
            temp[k] = d2;
            if (d2 > best) best = d2, bk = k, bx = px, by = py, bz = pz;
        }
        const int hi = __float_as_int(best);
        const unsigned lo = ~(brev_r | ((unsigned)bk >> L));
        int wl = warp_best_lane(hi, lo);
        if (lane == wl) entry_store(&s_warp[par][warp], hi, lo, bx, by, bz);
        __syncthreads();
        int h2 = __float_as_int(-1.0f);
        unsigned l2 = 0u;
        if (lane < nwarps) h2 = s_warp[par][lane].hi, l2 = s_warp[par][lane].lo;
        wl = warp_best_lane(h2, l2);
        const FpsEntry *e = &s_warp[par][wl];
        int k = 0;
        if (e->hi >= 0) {
            const unsigned prio = ~e->lo;
            const unsigned topmask = L ? (0xffffffffu << (32 - L)) : 0u;
            k = (int)(((prio & ~topmask) << L) | __brev(prio & topmask));
            cx = e->x, cy = e->y, cz = e->z;
        } else {
            cx = p0x, cy = p0y, cz = p0z;
        }
        if (tid == 0) {
            idxs[j] = k;
            if (new_xyz) new_xyz[3 * j] = cx, new_xyz[3 * j + 1] = cy, new_xyz[3 * j + 2] = cz;
        }
    }
}

static int floor_log2(int v) {
    int l = 0;
    while ((2 << l) <= v) ++l;
    return l;
}

template <int PPT>
static int launch_onchip(int b, int n, int m, int row_stride, int L, int NT, int CS, const float *pts,
                         int32_t *idxs, float *new_xyz, cudaStream_t stream) {
    auto kern = fps_onchip_kernel<PPT>;
    const size_t smem = (size_t)3 * PPT * NT * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return status_of(e);
    if (CS > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return status_of(e);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)b * CS);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, n, m, row_stride, L, pts, idxs, new_xyz);
    return status_of(e);
}

static int fps_dispatch(int b, int n, int m, int row_stride, const float *pts, float *temp,
                        int32_t *idxs, float *new_xyz, cudaStream_t stream) {
    if (b < 0 || n <= 0 || m < 0 || row_stride < 3 || !pts || (!idxs && m > 0)) return SG4D_EINVAL;
    if (b == 0 || m == 0) return SG4D_OK;
    // reference thread count T = opt_n_threads(n) = min(2^floor(log2 n), 512)  (cuda_utils.h:15-19;
    // its int(log(n)/log(2.0)) equals the exact floor for every n <= 200000, tests/test_oracle.py)
    int L = floor_log2(n);
    if (L > 9) L = 9;
    const int T = 1 << L;
    int NT = T < 32 ? 32 : T;  // always a multiple of T
    int CS = 1;
    if (n > 512 * 12) NT = 1024;
    while ((long long)NT * CS * 12 < n && CS < 16) CS *= 2;
    if ((long long)NT * CS * 12 < n) {  // does not fit on chip
        if (!temp) return SG4D_EINVAL;
        fps_stream_kernel<<<b, 1024, 0, stream>>>(n, m, row_stride, L, pts, temp, idxs, new_xyz);
        return SG4D_LAUNCH_CHECK();
    }
    const int ppt = (n + NT * CS - 1) / (NT * CS);
#define SG4D_FPS_CASE(P) \
    if (ppt <= P) return launch_onchip<P>(b, n, m, row_stride, L, NT, CS, pts, idxs, new_xyz, stream);
    SG4D_FPS_CASE(1)
    SG4D_FPS_CASE(2)
    SG4D_FPS_CASE(4)
    SG4D_FPS_CASE(6)
    SG4D_FPS_CASE(8)
    SG4D_FPS_CASE(10)
    SG4D_FPS_CASE(12)
#undef SG4D_FPS_CASE
    return SG4D_EINVAL;
}

}  // namespace sg4d

extern "C" int sg4d_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp,
                                            int32_t *idxs, sg4d_stream_t stream) {
    return sg4d::fps_dispatch(b, n, m, 3, dataset, temp, idxs, nullptr, (cudaStream_t)stream);
}

extern "C" int sg4d_fps_rows(int b, int n, int m, int row_stride, const float *pts, float *temp,
                             int32_t *idxs, float *new_xyz, sg4d_stream_t stream) {
    return sg4d::fps_dispatch(b, n, m, row_stride, pts, temp, idxs, new_xyz, (cudaStream_t)stream);
}
