// common.cuh -- shared device helpers for libsg4d (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sg4d.h"

#define SG4D_NUM_SMS 148  // B200: 2 dies x 74 SMs

namespace sg4d {

// Squared distance in the exact operation order the reference kernels compile to
// (t = dy*dy; t = fma(dx,dx,t); d = fma(dz,dz,t) -- SASS of ball_query_gpu.cu:31-32 and
// sampling_gpu.cu:100-104 built with nvcc 12.9 for sm_100a).  Explicit intrinsics so that no
// compiler flag can re-associate it.
__device__ __forceinline__ float sqdist3(float dx, float dy, float dz) {
    float t = __fmul_rn(dy, dy);
    t = __fmaf_rn(dx, dx, t);
    return __fmaf_rn(dz, dz, t);
}

__device__ __forceinline__ int redux_max_s32(int v) {
    int r;
    asm volatile("redux.sync.max.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ unsigned redux_max_u32(unsigned v) {
    unsigned r;
    asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ unsigned redux_min_u32(unsigned v) {
    unsigned r;
    asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ int redux_add_s32(int v) {
    int r;
    asm volatile("redux.sync.add.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}

// ---- thread-block cluster primitives (raw PTX; no cooperative_groups dependency) ----
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_id_x() {
    unsigned r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive_release() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// map a local shared-memory address to the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_smem_addr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                              uint32_t d) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b),
                 "r"(c), "r"(d)
                 : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t a) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ball_query.cu: brute-force scan of the first n_scan points of every cloud (n_scan == n: the complete query)
int bq_launch(int b, int n, int n_scan, int m, int pts_stride, int ctr_stride, int nsc, const float *radius,
              const int *nsample, const float *centers, const float *pts, int32_t *const *idx, int32_t *const *cnt,
              cudaStream_t stream, int *todo_cnt = nullptr, int *todo_list = nullptr, int todo_cap = 0);

inline int status_of(cudaError_t e) { return e == cudaSuccess ? SG4D_OK : static_cast<int>(e); }

}  // namespace sg4d

#define SG4D_LAUNCH_CHECK() ::sg4d::status_of(cudaGetLastError())
