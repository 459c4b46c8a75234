// ball_query.cu -- radius neighbour search for sm_100a.
//
// Replaces query_ball_point_kernel of the reference (EXT/src/ball_query_gpu.cu:9-44, launch :46-54,
// host EXT/src/ball_query.cpp:8-32), where ONE THREAD per centre walks the cloud serially and the
// cloud is re-scanned once per radius.
//
// Here: one WARP per centre, 32 points tested per step, `__ballot_sync` + prefix-popcount to append
// hits in ascending index order (so the "first nsample hits in index order, remaining slots = first
// hit" contract of the reference is reproduced bit for bit), up to two radii answered in the same
// pass, early exit as soon as every radius has its nsample hits.  The 16 warps of a CTA share the
// point tiles, which are staged once into shared memory as SoA so that both the global reads
// (coalesced) and the per-lane shared reads (conflict free) are unit stride.
#include <math_constants.h>

#include "common.cuh"

namespace sg4d {

constexpr int kBqWarps = 16;     // centres per CTA
constexpr int kBqTile = 2048;    // points per staged tile (24 KB as SoA)

struct BqScale {
    float r2;       // radius*radius, rounded to fp32 like the reference (ball_query_gpu.cu:22)
    int ns;         // nsample
    int32_t *idx;   // (b, m, ns)
    int32_t *cnt;   // (b, m) or nullptr
};

template <int NSC>
__global__ void __launch_bounds__(kBqWarps * 32)
ball_query_kernel(int n, int n_scan, int m, int pts_stride, int ctr_stride, int ctas_per_cloud,
                  const float *__restrict__ centers, const float *__restrict__ pts, BqScale s0,
                  BqScale s1, int *__restrict__ todo_cnt, int *__restrict__ todo_list, int todo_cap) {
    __shared__ float sx[kBqTile], sy[kBqTile], sz[kBqTile];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x / ctas_per_cloud;
    const int j = (blockIdx.x % ctas_per_cloud) * kBqWarps + warp;  // my centre
    const bool active = j < m;
    pts += (size_t)cloud * n * pts_stride;

    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float *c = centers + ((size_t)cloud * m + j) * ctr_stride;
        qx = __ldg(c), qy = __ldg(c + 1), qz = __ldg(c + 2);
    }
    BqScale sc[2] = {s0, s1};
    int cnt[2] = {0, 0}, first[2] = {0, 0};
    int32_t *row[2];
#pragma unroll
    for (int s = 0; s < NSC; ++s) row[s] = sc[s].idx + ((size_t)cloud * m + (active ? j : 0)) * sc[s].ns;

    bool done = !active;
    const unsigned lt = (1u << lane) - 1u;
    for (int base = 0; base < n_scan; base += kBqTile) {   // n_scan < n: prefix pass, completed by spatial.cu
        const int tn = min(kBqTile, n_scan - base);
        for (int i = tid; i < tn; i += kBqWarps * 32) {
            const float *r = pts + (size_t)(base + i) * pts_stride;
            sx[i] = __ldg(r), sy[i] = __ldg(r + 1), sz[i] = __ldg(r + 2);
        }
        __syncthreads();
        if (!done) {
            for (int i = 0; i < tn; i += 32) {
                const int k = i + lane;
                float d2 = CUDART_INF_F;
                if (k < tn) d2 = sqdist3(qx - sx[k], qy - sy[k], qz - sz[k]);
                bool all_full = true;
#pragma unroll
                for (int s = 0; s < NSC; ++s) {
                    if (cnt[s] < sc[s].ns) {
                        const bool hit = d2 < sc[s].r2;  // ordered compare: NaN never hits
                        const unsigned mask = __ballot_sync(0xffffffffu, hit);
                        if (mask) {
                            if (cnt[s] == 0) first[s] = base + i + __ffs(mask) - 1;
                            const int pos = cnt[s] + __popc(mask & lt);
                            if (hit && pos < sc[s].ns) row[s][pos] = base + k;
                            cnt[s] = min(sc[s].ns, cnt[s] + __popc(mask));
                        }
                        all_full = all_full && (cnt[s] >= sc[s].ns);
                    }
                }
                if (all_full) {
                    done = true;
                    break;
                }
            }
        }
        if (__syncthreads_and(done)) break;  // also guards the tile buffers before the next refill
    }
    if (active) {
#pragma unroll
        for (int s = 0; s < NSC; ++s) {
            // slots cnt..ns-1 keep the first hit (ball_query_gpu.cu:34-38); a row without hits is
            // all zeros (ball_query.cpp:19-21 zero-initialises the output)
            for (int p = cnt[s] + lane; p < sc[s].ns; p += 32) row[s][p] = first[s];
            if (sc[s].cnt && lane == 0) sc[s].cnt[(size_t)cloud * m + j] = cnt[s];
        }
        // prefix pass (n_scan < n): centres still short of nsample hits are queued for the spatial-index pass
        if (todo_cnt && lane == 0) {
            bool shortfall = false;
#pragma unroll
            for (int s = 0; s < NSC; ++s) shortfall = shortfall || (cnt[s] < sc[s].ns);
            if (shortfall) todo_list[(size_t)cloud * todo_cap + atomicAdd(todo_cnt + cloud, 1)] = j;
        }
    }
}

int bq_launch(int b, int n, int n_scan, int m, int pts_stride, int ctr_stride, int nsc, const float *radius,
                     const int *nsample, const float *centers, const float *pts, int32_t *const *idx,
                     int32_t *const *cnt, cudaStream_t stream, int *todo_cnt, int *todo_list, int todo_cap) {
    BqScale sc[2] = {{0.f, 0, nullptr, nullptr}, {0.f, 0, nullptr, nullptr}};
    for (int s = 0; s < nsc; ++s) {
        if (nsample[s] <= 0 || !idx[s]) return SG4D_EINVAL;
        const float r = radius[s];
        sc[s].r2 = r * r;
        sc[s].ns = nsample[s];
        sc[s].idx = idx[s];
        sc[s].cnt = cnt ? cnt[s] : nullptr;
    }
    const int cpc = (m + kBqWarps - 1) / kBqWarps;
    const long long grid = (long long)b * cpc;
    if (grid > 0x7fffffffLL) return SG4D_EINVAL;
    if (nsc == 1)
        ball_query_kernel<1><<<(unsigned)grid, kBqWarps * 32, 0, stream>>>(n, n_scan, m, pts_stride, ctr_stride, cpc,
                                                                          centers, pts, sc[0], sc[1], todo_cnt, todo_list, todo_cap);
    else
        ball_query_kernel<2><<<(unsigned)grid, kBqWarps * 32, 0, stream>>>(n, n_scan, m, pts_stride, ctr_stride, cpc,
                                                                          centers, pts, sc[0], sc[1], todo_cnt, todo_list, todo_cap);
    return SG4D_LAUNCH_CHECK();
}

}  // namespace sg4d

extern "C" int sg4d_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                               const float *xyz, int32_t *idx, sg4d_stream_t stream) {
    if (b < 0 || n <= 0 || m < 0 || nsample <= 0 || !new_xyz || !xyz || !idx) return SG4D_EINVAL;
    if (b == 0 || m == 0) return SG4D_OK;
    int32_t *idxs[1] = {idx};
    return sg4d::bq_launch(b, n, n, m, 3, 3, 1, &radius, &nsample, new_xyz, xyz, idxs, nullptr,
                           (cudaStream_t)stream);
}

extern "C" int sg4d_ball_query_rows(int b, int n, int m, int row_stride, int center_stride, int nscales,
                                    const float *radius, const int *nsample, const float *centers,
                                    const float *pts, int32_t *const *idx, int32_t *const *cnt,
                                    sg4d_stream_t stream) {
    if (b < 0 || n <= 0 || m < 0 || row_stride < 3 || center_stride < 3 || nscales < 1 ||
        nscales > SG4D_MAX_SCALES || !radius || !nsample || !centers || !pts || !idx)
        return SG4D_EINVAL;
    if (b == 0 || m == 0) return SG4D_OK;
    for (int s = 0; s < nscales; s += 2) {  // two radii per pass over the cloud
        const int k = nscales - s >= 2 ? 2 : 1;
        const int st = sg4d::bq_launch(b, n, n, m, row_stride, center_stride, k, radius + s, nsample + s,
                                       centers, pts, idx + s, cnt ? cnt + s : nullptr, (cudaStream_t)stream);
        if (st != SG4D_OK) return st;
    }
    return SG4D_OK;
}
