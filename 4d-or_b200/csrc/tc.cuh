// tc.cuh -- raw-PTX wrappers for the Blackwell tensor-core path (tcgen05 / TMEM / mbarrier / bulk-TMA),
// sm_100a only.  No CUTLASS dependency; descriptor layouts follow cute/arch/mma_sm100_desc.hpp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sg4d {
namespace tc {

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies read smem through it)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------ bulk TMA (1-D, no tensor map)
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void *gmem_src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gmem_src), "r"(bytes), "r"(bar)
                 : "memory");
}

// ------------------------------------------------------------------ TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets row (lane base + i), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ UMMA descriptors
// K-major operand tile in shared memory, 128-byte swizzle: rows of 32 fp32 (128 B); 8-row atoms of 1024 B;
// inside an atom the 16-byte chunk j of row r sits at chunk position j ^ (r & 7).  Tile base 1024-B aligned.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);   // start address, 16-byte units      bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: next 8-row atom   bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version 1 (sm_100)
    d |= (uint64_t)2 << 61;                          // layout type: SWIZZLE_128B
    return d;
}
// byte offset of element (row r, fp32 column c in [0,32)) inside such a tile
__device__ __forceinline__ uint32_t sw128_offset(int r, int c) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 2) ^ (r & 7)) << 4) | ((c & 3) << 2)));
}
// instruction descriptor: kind::tf32, fp32 accumulate, A and B K-major, shape M x N x 8
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4)                      // D format  : F32
           | (2u << 7) | (2u << 10)       // A, B format: TF32
           | ((uint32_t)(N >> 3) << 17)   // N >> 3
           | ((uint32_t)(M >> 4) << 24);  // M >> 4
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// round-to-nearest split of an fp32 value into a TF32-exact high part and the (exact) remainder
__device__ __forceinline__ void split_tf32(float a, float &hi, float &lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(a));
    hi = __uint_as_float(h);
    const float r = a - hi;  // exact remainder, |r| <= 2^-11 |a|
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(r));
    lo = __uint_as_float(h);  // |a - hi - lo| <= 2^-22 |a|, unbiased
}

// Same contract in 3 instructions instead of 9 (cvt.rna.tf32 compiles to add / isfinite / select / mask on sm_100):
// hi = a with the 13 low mantissa bits cleared (what the tensor core would read anyway), r = a - hi exact with
// |r| < 2^-10 |a|, lo = r rounded to nearest by adding half an ulp of tf32 before the hardware truncates it
// (r is tiny, so the add cannot overflow the exponent).  |a - hi - lo| <= 2^-21 |a|, unbiased.
__device__ __forceinline__ void split_tf32_fast(float a, float &hi, float &lo) {
    hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
    const float r = a - hi;
    lo = __uint_as_float(__float_as_uint(r) + 0x1000u);
}

// Round-to-nearest variant for the weight-gradient kernels (one more integer add): hi = a rounded to tf32 by adding half
// an ulp to the magnitude before masking, so |r| <= 2^-11 |a| instead of 2^-10 |a| -- the residual |a - hi - lo| and
// the lo*lo product are 2x / 4x smaller.  Weight gradients are sums with heavy cancellation (|sum| << sum |terms|), where
// the per-product error of the truncating split (about 10x an fp32 FMA) showed up as 1e-4..3e-4 of the result at
// benchmark size (tests/test_gpu_full_size.py).
__device__ __forceinline__ void split_tf32_rn(float a, float &hi, float &lo) {
    hi = __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u);
    const float r = a - hi;
    lo = __uint_as_float(__float_as_uint(r) + 0x1000u);
}

}  // namespace tc
}  // namespace sg4d
