// mlp.cu -- the shared point-MLP of a set-abstraction level on the 5th-generation tensor cores.
//
// Replaces, for one scale of build_shared_mlp (OPS/pointnet2_modules.py:9-19,66-70), the chain
//     [Conv2d(1x1, bias=False) -> BatchNorm2d (batch statistics) -> ReLU(inplace)] x 2 -> max_pool2d over nsample
// forward AND backward, which the reference runs as separate cuDNN / ATen launches with
// (B,C,npoint,nsample) activations round-tripping HBM between every op.
//
// A 1x1 convolution over (B,C,npoint,nsample) is Y[R,N] = A[R,K] * W[N,K]^T over the R = B*npoint*nsample rows
// of the point-major grouped tensor.  Two persistent, warp-specialised kernel families share one engine:
//
//   row_gemm_kernel  (T1)   D[128-row tile, N] = P(tile)[128, K] * W^T      forward layers, dA and dX
//   wgrad_kernel     (T2)   D[M, N] += P(tile)^T[M, 128] * Q(tile)[128, N]  weight gradients, K = rows
//
//   producers (4 warps)  build the operand tiles: coalesced 16-byte loads from HBM, an elementwise PROLOGUE
//                        applied in registers (BatchNorm scale/shift + ReLU of the previous layer in the forward
//                        pass; the BatchNorm / ReLU / max-pool backward formulas in the backward pass -- so
//                        neither normalised activations nor dY tensors ever exist in memory), a round-to-nearest
//                        split into TF32 hi + lo parts, and stores into 128-byte-swizzled shared-memory tiles.
//                        Weight k-blocks (pre-split, pre-swizzled image) arrive by one bulk-TMA copy per stage.
//   MMA (1 thread)       tcgen05.mma kind::tf32, three products per k-step (lo*hi + hi*lo + hi*hi: "3xTF32",
//                        fp32-level accuracy; plain TF32 would miss the 1e-4 parity bound), fp32 accumulators in
//                        TMEM (double buffered across tiles in T1, resident for the whole kernel in T2).
//   epilogue (4 warps)   tcgen05.ld the accumulator, stage it in shared memory, store coalesced, and in the same
//                        pass produce the per-channel reductions BatchNorm needs (fp64 across tiles) and, for the
//                        last layer of a scale, the per-group (nsample rows) max or min pre-activation and its
//                        row.  BatchNorm+ReLU are monotone per channel, so max_k relu(bn(y_k)) = relu(bn(max_k y_k))
//                        for a non-negative scale and relu(bn(min_k y_k)) otherwise.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc.cuh"

namespace sg4d {

// Process-wide compute precision of the tensor-core layers (sg4d_set_compute_precision):
//   0  fp32-level: 3xTF32 (4 products in the weight-gradient kernels), the default and the headline configuration
//   1  bf16: every MMA operand is rounded to bf16 (round to nearest even) when it is staged, ONE product per k-step, fp32
//      accumulation in TMEM, fp32 BatchNorm statistics / activations in HBM -- BASELINE.json configs[3].  A bf16 value is a
//      tf32 value, so the same kind::tf32 instruction computes exactly what a bf16 MMA would; what changes is the traffic:
//      no lo tiles are written or read (shared-memory bytes per k-block drop from 208 KB to 112 KB).
static int g_precision = 0;

__device__ __forceinline__ float bf16_round(float a) {
    const uint32_t u = __float_as_uint(a);
    return __uint_as_float((u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u);      // RNE to 8 exponent + 7 mantissa bits
}
__device__ __forceinline__ float4 bf16_round4(const float4 &v) {
    return make_float4(bf16_round(v.x), bf16_round(v.y), bf16_round(v.z), bf16_round(v.w));
}

constexpr int kTileM = 128;       // rows per tile = UMMA M
constexpr int kKB = 32;           // fp32 per k-block (one 128-byte swizzle row)
constexpr int kStages = 2;
// T1 (row_gemm): warps 0-7 producers, 8 MMA, 9-12 epilogue (kEpiThreads = 256 -- two warps per TMEM lane quarter --
// is supported by the code below but measured slower: the register cap drops to 96 and the producers spill).
constexpr int kProdThreads = 256, kEpiThreads = 128;
constexpr int kProdWarps = kProdThreads / 32;
constexpr int kProdRows = kTileM * 8 / kProdThreads;          // float4 items per producer thread per T1 k-block
constexpr int kMlpThreads = kProdThreads + 32 + kEpiThreads;
// T2 (wgrad): no per-tile epilogue; the operand transform bounds it, hence 16 producer warps (+1 MMA, +4 epilogue)
constexpr int kProdThreadsT2 = 512;
constexpr int kMlpThreadsT2 = kProdThreadsT2 + 32 + 128;

// ------------------------------------------------------------------------------------------------
// Operand generator: element (row, col) of the logical operand matrix P, computed from arrays in HBM.
//   mode 0   P = A
//   mode 1   P = relu(A * s + t)                              BatchNorm + ReLU of the previous layer
//   mode 2   P = dsel[g] * [row % S == garg[g]] - (A * s + t)  dY of a pooled layer (A = Y; g = row / S)
//   mode 3   P = A2 * p - (A * s + t)                          dY of an inner layer (A = Y, A2 = dZ)
//   mode 4   P = relu(W1s * x(row) + t1)   the first layer of an SA1 scale RECOMPUTED from the gathered neighbour
//            (x = [xyz - centre | feats | 0.. | 1], K <= 7): neither the grouped tensor nor y1 exist in memory
//   mode 5   P = grouped row gathered on the fly: [feats(idx) (c) | xyz(idx) - centre (3) | 0-pad]   (SA2)
// Grouped-row source of one SA scale (replaces the materialised output of sg4d_group_rows):
// row r = (cloud * m + centre) * ns + slot, neighbour point = idx[r].
struct GroupSrc {
    const float *pts, *feats, *centers;   // pts (B,n,pstride) xyz in 0..2; feats (B,n,fstride); centers (B,m,3)
    const int32_t *idx;                   // (B,m,ns)
    int n, m, logns, pstride, fstride, foff, c;
    const float *w1s, *t1;                // mode 4: (8, 64) BatchNorm-scaled first-layer weights (input-major), (64) shifts
};

struct Operand {
    const float *A;
    int lda;
    const float *A2;
    int lda2;
    const float *s, *t, *p;     // per-column constants (length >= ncols)
    const float *dsel;          // (rows / S, ldsel)
    const uint8_t *garg;
    int S, logS, ldsel;         // S = 1 << logS rows per group
    int ncols;                  // valid columns; beyond -> 0
    GroupSrc g;                 // modes 4 / 5
};

struct RawVec {                 // what one thread fetches for 4 consecutive columns of one row
    float4 a, a2;
    uint32_t arg;
};

template <int MODE>
__device__ __forceinline__ void op_load(const Operand &o, long long row, long long nrows, int col, RawVec &r) {
    const bool ok = row < nrows && col < o.ncols;
    r.a = ok ? __ldg(reinterpret_cast<const float4 *>(o.A + row * o.lda + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == 3)
        r.a2 = ok ? __ldg(reinterpret_cast<const float4 *>(o.A2 + row * o.lda2 + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == 2) {
        const long long g = row >> o.logS;
        r.a2 = ok ? __ldg(reinterpret_cast<const float4 *>(o.dsel + g * o.ldsel + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
        r.arg = ok ? __ldg(reinterpret_cast<const uint32_t *>(o.garg + g * o.ldsel + col)) : 0u;
    }
}

template <int MODE>
__device__ __forceinline__ float4 op_apply(const Operand &o, const RawVec &r, long long row, long long nrows, int col,
                                           const float *s_s, const float *s_t, const float *s_p) {
    if (MODE == 0) return r.a;
    const bool ok = row < nrows && col < o.ncols;
    if (!ok) return make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 s = *reinterpret_cast<const float4 *>(s_s + col), t = *reinterpret_cast<const float4 *>(s_t + col);
    float4 v;
    v.x = fmaf(r.a.x, s.x, t.x), v.y = fmaf(r.a.y, s.y, t.y), v.z = fmaf(r.a.z, s.z, t.z), v.w = fmaf(r.a.w, s.w, t.w);
    if (MODE == 1) {
        v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
    } else if (MODE == 2) {
        const uint32_t k = (uint32_t)row & (uint32_t)(o.S - 1);
        v.x = (((r.arg) & 0xffu) == k ? r.a2.x : 0.f) - v.x;
        v.y = (((r.arg >> 8) & 0xffu) == k ? r.a2.y : 0.f) - v.y;
        v.z = (((r.arg >> 16) & 0xffu) == k ? r.a2.z : 0.f) - v.z;
        v.w = (((r.arg >> 24) & 0xffu) == k ? r.a2.w : 0.f) - v.w;
    } else {  // MODE == 3
        const float4 p = *reinterpret_cast<const float4 *>(s_p + col);
        v.x = fmaf(r.a2.x, p.x, -v.x), v.y = fmaf(r.a2.y, p.y, -v.y), v.z = fmaf(r.a2.z, p.z, -v.z),
        v.w = fmaf(r.a2.w, p.w, -v.w);
    }
    return v;
}

// ---- gather modes -----------------------------------------------------------------------------
// x = [xyz(idx) - centre | feats(idx) (c <= 4) | 0.. | 1]: the reference's grouped row (grouping_operation on xyz,
// `grouped_xyz -= new_xyz`, cat with the grouped features: OPS/pointnet2_utils.py:319-328) plus a constant 1 in the
// last slot, which turns the per-channel sums of the backward pass into one more column of the same product.
// (cloud, centre-group) of a grouped row with 32-bit arithmetic (groups < 2^31; a 64-bit division costs ~100 instructions)
__device__ __forceinline__ void src_locate(const GroupSrc &g, long long row, uint32_t &gi, uint32_t &cloud) {
    gi = (uint32_t)(row >> g.logns);
    cloud = gi / (uint32_t)g.m;
}
__device__ __forceinline__ void sa1_gather_row(const GroupSrc &g, long long row, bool ok, int i, float (&x)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = 0.f;
    if (!ok) return;
    uint32_t gi, cloud;
    src_locate(g, row, gi, cloud);
    const size_t pi = (size_t)cloud * (uint32_t)g.n + (uint32_t)i;
    const float *pp = g.pts + pi * (uint32_t)g.pstride;
    const float *ff = g.feats + pi * (uint32_t)g.fstride + g.foff;
    const float *qq = g.centers + (size_t)gi * 3;
    x[0] = __ldg(pp) - __ldg(qq);
    x[1] = __ldg(pp + 1) - __ldg(qq + 1);
    x[2] = __ldg(pp + 2) - __ldg(qq + 2);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
        if (ch < g.c) x[3 + ch] = __ldg(ff + ch);
    x[7] = 1.f;
}
// BatchNorm-folded first layer for 4 consecutive output channels: ((t + x0 w0) + x1 w1) + ... in THIS order everywhere
// (forward producer, backward epilogue, weight-gradient producer), so that the recomputed activation and its ReLU
// mask are bit-identical to what the forward pass fed into the second layer.
__device__ __forceinline__ float4 sa1_y1bn(const float (&x)[8], const float4 (&w)[8], const float4 &t) {
    float4 v = t;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v.x = __fmaf_rn(x[j], w[j].x, v.x), v.y = __fmaf_rn(x[j], w[j].y, v.y);
        v.z = __fmaf_rn(x[j], w[j].z, v.z), v.w = __fmaf_rn(x[j], w[j].w, v.w);
    }
    return v;
}
// same arithmetic, same order, with the weights read from shared memory one input at a time (w: 8 rows of 64 floats, input-major):
// 4 live weight registers instead of 32 -- the register-starved backward kernels use this form
__device__ __forceinline__ float4 sa1_y1bn_smem(const float (&x)[8], const float *w, int col, const float4 &t) {
    float4 v = t;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 wj = *reinterpret_cast<const float4 *>(w + j * 64 + col);
        v.x = __fmaf_rn(x[j], wj.x, v.x), v.y = __fmaf_rn(x[j], wj.y, v.y);
        v.z = __fmaf_rn(x[j], wj.z, v.z), v.w = __fmaf_rn(x[j], wj.w, v.w);
    }
    return v;
}
// mode 5: 4 consecutive columns of the grouped row [feats (c, multiple of 4) | xyz - centre | 0-pad]
__device__ __forceinline__ float4 sa2_gather4(const GroupSrc &g, long long row, bool ok, int col, int i) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!ok || col > g.c) return v;
    uint32_t gi, cloud;
    src_locate(g, row, gi, cloud);
    const size_t pi = (size_t)cloud * (uint32_t)g.n + (uint32_t)i;
    if (col < g.c) return __ldg(reinterpret_cast<const float4 *>(g.feats + pi * (uint32_t)g.fstride + g.foff + col));
    const float *pp = g.pts + pi * (uint32_t)g.pstride;
    const float *qq = g.centers + (size_t)gi * 3;
    v.x = __ldg(pp) - __ldg(qq), v.y = __ldg(pp + 1) - __ldg(qq + 1), v.z = __ldg(pp + 2) - __ldg(qq + 2);
    return v;
}

// One lane polls the mbarrier, the rest of the warp waits at the warp barrier: 32x fewer try_wait requests hit
// the barrier unit (with every thread polling, the waits themselves were the top stall in the first ncu capture).
// SLEEP_NS > 0: back off between polls.  The waiting warps share the four issue ports with the producers, and the
// ncu instruction counts showed ~45 % of all issued instructions to be try_wait / branch / yield of such loops.
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait_warp(int lane, uint32_t bar, uint32_t parity) {
    if (lane == 0) {
        while (!tc::mbar_try_wait(bar, parity)) {
            if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void split4(const float4 &v, float4 &hi, float4 &lo) {
    tc::split_tf32_fast(v.x, hi.x, lo.x);
    tc::split_tf32_fast(v.y, hi.y, lo.y);
    tc::split_tf32_fast(v.z, hi.z, lo.z);
    tc::split_tf32_fast(v.w, hi.w, lo.w);
}

__device__ __forceinline__ void split4_rn(const float4 &v, float4 &hi, float4 &lo) {
    tc::split_tf32_rn(v.x, hi.x, lo.x);
    tc::split_tf32_rn(v.y, hi.y, lo.y);
    tc::split_tf32_rn(v.z, hi.z, lo.z);
    tc::split_tf32_rn(v.w, hi.w, lo.w);
}

// ================================================================================================
// T1: row-tile GEMM
//   EMODE 0  store D; stats (sum d, sum d^2); optional group max/min + arg            forward layers
//   EMODE 1  v = d * [E*es + et > 0]; store v; stats (sum v, sum v * (E*ei + em))      dZ of the inner layer
//   EMODE 2  store D into Y(:, ycol0 : ycol0+N) with row stride ldy; no stats           dX
//   EMODE 3  SA1 single-pass backward (N = 64): dz1 = D * [y1bn(x) > 0] with the first layer recomputed from the
//            gathered neighbour x; accumulates S1 = dz1^T [x | 1] (64 x 8, fp64 across tiles).  dz1 is never stored:
//            d_beta1, d_gamma1 and dW1 are all linear in S1 and in the forward moments sum x x^T (sa1_bwd_finalize).
struct RowGemmArgs {
    Operand op;
    long long R;
    const float *wimg;       // packed weight image: [nkb][hi|lo][N rows x 128 B swizzled]
    float *Y;                // output rows (or nullptr)
    int ldy, ycol0;
    double *partial;         // (gridDim.x, kEpiThreads, 2) fp64 partial sums (EMODE 0/1)
    int S, logS;             // EMODE 0: rows per group (power of two) for the max/min reduction, 0 = none
    const float *gamma;      // (N): sign picks max (>= 0) or min (< 0) per channel
    float *gsel;             // (R/S, N)
    uint8_t *garg;           // (R/S, N)
    const float *E;          // EMODE 1: (R, N) pre-activation of the layer whose ReLU is differentiated
    const float *es, *et, *ei, *em;   // (N) each
    double *s1part;          // EMODE 3: (gridDim.x, 64, 8) per-CTA partial of S1
    // column panels (gridDim.y > 1): CTA (x, y) computes output columns [y*N, (y+1)*N) of a wider layer
    long long wimg_panel;    // floats between the weight images of two panels
    long long partial_panel; // doubles between the statistics buffers of two panels
    int ldg, lde;            // row strides of gsel / garg and of E (0 = N)
    const float *bias;       // EMODE 2: added to every output row (N per panel) or nullptr
    int dbg_no_tma, dbg_no_mma, dbg_no_load, dbg_no_epi;   // ablation switches, honoured only in SG4D_DEBUG builds
    int lp;                  // 1: bf16 operands, one product per k-step (g_precision)
};

template <int N>
struct RowSmem {
    static constexpr int kStg = (N == 64) ? 3 : 2;          // smem stages: as many as fit beside the C tile
    static constexpr int kABytes = kTileM * 128;            // one hi or lo A tile
    static constexpr int kWBytes = N * 128;                 // one hi or lo W tile
    static constexpr int kStageBytes = 2 * kABytes + 2 * kWBytes;
    static constexpr int kCStride = N + 4;                  // padded fp32 row stride of the staged C tile
    static constexpr int kCBytes = kTileM * kCStride * 4;
    static constexpr int kCF = 1280;                        // capacity (floats) of each prologue constant vector
    static constexpr int kConst = 3 * kCF * 4;              // prologue constants s, t, p
    static constexpr int stages(int pmode) { return pmode == 2 ? 2 : kStg; }   // PMODE 2 gathers through L1: keep it large
    static constexpr int kXaBytes = kTileM * 8 * 4;         // EMODE 3: the tile's gathered rows [x | 1]
    static constexpr int kSaccBytes = 32 * kEpiThreads * 8; // EMODE 3: fp64 accumulators, 32 per epilogue thread
    static constexpr int extra(int emode) { return emode == 3 ? kXaBytes + kSaccBytes : 0; }
    static constexpr int total(int pmode, int emode) { return 1024 + stages(pmode) * kStageBytes + kCBytes + kConst + extra(emode) + 128; }
};

// PT = producer threads: 256, or 512 for layers with many k-blocks per tile (K >= 128), which are producer-bound
#ifdef SG4D_DEBUG
#define SG4D_DBG(x) (x)
// timeline tracing of CTA 0 (tools/trace_mlp.py): one clock64() stamp per (role, event, index)
__device__ unsigned long long *g_trace = nullptr;
#define SG4D_TRACE_INIT() unsigned long long *trc = (blockIdx.x == 0 && blockIdx.y == 0) ? g_trace : nullptr
#define SG4D_TRACE(cond, slot)                                                 \
    do {                                                                       \
        if (trc && (cond) && (slot) < 8192) trc[(slot)] = clock64();           \
    } while (0)
#else
#define SG4D_DBG(x) 0
#define SG4D_TRACE_INIT() do { } while (0)
#define SG4D_TRACE(cond, slot) do { } while (0)
#endif

template <int N, int PMODE, int EMODE, int PT>
__global__ void __launch_bounds__(PT + 32 + kEpiThreads, 1) row_gemm_kernel(RowGemmArgs p) {
    constexpr int kProdThreads = PT, kProdWarps = PT / 32, kProdRows = kTileM * 8 / PT, kMlpThreads = PT + 32 + kEpiThreads;
    using SM = RowSmem<N>;
    constexpr int kStages = SM::stages(PMODE);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // 1024-byte aligned (swizzle atoms)
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
    float *Cs = reinterpret_cast<float *>(smem + kStages * SM::kStageBytes);
    float *s_s = reinterpret_cast<float *>(smem + kStages * SM::kStageBytes + SM::kCBytes);
    float *s_t = s_s + SM::kCF, *s_p = s_t + SM::kCF;
    __shared__ __align__(8) uint64_t s_bar[2 * 4 + 4];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    SG4D_TRACE_INIT();
    const int K = p.op.ncols;
    const int nkb = (K + kKB - 1) / kKB;
    const long long ntiles = (p.R + kTileM - 1) / kTileM;
    // column panel of a wider layer (gridDim.y > 1): everything that is indexed by output column is shifted -- in LOCAL
    // copies; writing to the parameter struct would move all of it to local memory
    const int c0 = blockIdx.y * N;
    const float *q_wimg = p.wimg + (size_t)blockIdx.y * p.wimg_panel;
    const int q_ycol0 = p.ycol0 + c0;
    double *q_partial = p.partial ? p.partial + (size_t)blockIdx.y * p.partial_panel : nullptr;
    const float *q_gamma = p.gamma ? p.gamma + c0 : nullptr;
    float *q_gsel = p.gsel ? p.gsel + c0 : nullptr;
    uint8_t *q_garg = p.garg ? p.garg + c0 : nullptr;
    const float *q_E = p.E ? p.E + c0 : nullptr;
    const float *q_es = p.es ? p.es + c0 : nullptr, *q_et = p.et ? p.et + c0 : nullptr;
    const float *q_ei = p.ei ? p.ei + c0 : nullptr, *q_em = p.em ? p.em + c0 : nullptr;
    const float *q_bias = p.bias ? p.bias + c0 : nullptr;
    const int ldg = p.ldg ? p.ldg : N, lde = p.lde ? p.lde : N;
    const uint32_t bar_full = smem_u32(&s_bar[0]), bar_empty = smem_u32(&s_bar[kStages]);
    const uint32_t bar_tfull = smem_u32(&s_bar[2 * kStages]), bar_tempty = smem_u32(&s_bar[2 * kStages + 2]);

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            tc::mbar_init(bar_full + 8 * s, kProdWarps);
            tc::mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(bar_tfull + 8 * a, 1);
            tc::mbar_init(bar_tempty + 8 * a, kEpiThreads / 32);
        }
        tc::mbar_fence_init();
    }
    if (warp == kProdWarps) tc::tmem_alloc(smem_u32(&s_tmem), 2 * N);
    if (EMODE == 3)              // s_p[0..511] = W1s for the epilogue's recomputation (PMODE 2 does not use s_p)
        for (int k = tid; k < 512; k += kMlpThreads) s_p[k] = p.op.g.w1s[k];
    if (PMODE == 4) {            // s_s[0..511] = W1s (input-major, 8 x 64), s_p[0..63] = t1
        for (int k = tid; k < 512; k += kMlpThreads) s_s[k] = p.op.g.w1s[k];
        for (int k = tid; k < 64; k += kMlpThreads) s_p[k] = p.op.g.t1[k];
    } else if (PMODE != 0 && PMODE != 5)
        for (int k = tid; k < ((K + 31) & ~31); k += kMlpThreads) {
            s_s[k] = k < K ? p.op.s[k] : 0.f;
            s_t[k] = k < K ? p.op.t[k] : 0.f;
            if (EMODE != 3) s_p[k] = (PMODE == 3 && k < K) ? p.op.p[k] : 0.f;     // EMODE 3 keeps W1s there
        }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = s_tmem;

    if (warp < kProdWarps) {
        // =============================================================== producers
        const int c = tid & 7;        // my 16-byte chunk (4 fp32 columns) of every k-block
        const int r0 = tid >> 3;      // my rows: r0 + 32*i
        // one k-block of the A tile is complete: weights by bulk TMA (one thread), my stores visible to the async proxy
        auto weights_tma = [&](int stage, int kb) {
            if (tid == 0 && !SG4D_DBG(p.dbg_no_tma)) {   // weight k-block: one bulk-TMA copy (hi tile followed by lo tile)
                uint8_t *st = smem + stage * SM::kStageBytes;
                const uint32_t wbytes = p.lp ? (uint32_t)SM::kWBytes : 2u * SM::kWBytes;     // bf16 mode: the hi tile only
                asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar_full + 8 * stage), "r"(wbytes)
                             : "memory");
                tc::bulk_g2s(smem_u32(st + 2 * SM::kABytes), q_wimg + (size_t)kb * (2 * SM::kWBytes / 4), wbytes, bar_full + 8 * stage);
            }
        };
        auto put = [&](uint8_t *st, int r, const float4 &v) {
            const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
            if (p.lp) {   // bf16 operands: the rounded value IS the operand, no lo tile
                *reinterpret_cast<float4 *>(st + off) = bf16_round4(v);
                return;
            }
            float4 hi, lo;
            split4(v, hi, lo);
            *reinterpret_cast<float4 *>(st + off) = hi;
            *reinterpret_cast<float4 *>(st + SM::kABytes + off) = lo;
        };
        if constexpr (PMODE == 4) {
            // SA1: the operand is the first layer's activation, recomputed from the gathered neighbour of every row
            // (two dependent loads: idx two tiles ahead, the point one tile ahead).  K = 64 -> two k-blocks per tile.
            const GroupSrc &g = p.op.g;
            float xc[kProdRows][8], xn[kProdRows][8];
            int ib[kProdRows];
            auto load_idx = [&](long long tile, int (&dst)[kProdRows]) {
#pragma unroll
                for (int i = 0; i < kProdRows; ++i) {
                    const long long row = tile * kTileM + r0 + (PT / 8) * i;
                    dst[i] = (tile < ntiles && row < p.R) ? __ldg(g.idx + row) : 0;
                }
            };
            auto gather = [&](long long tile, const int (&ix)[kProdRows], float (&x)[kProdRows][8]) {
#pragma unroll
                for (int i = 0; i < kProdRows; ++i) {
                    const long long row = tile * kTileM + r0 + (PT / 8) * i;
                    sa1_gather_row(g, row, tile < ntiles && row < p.R, ix[i], x[i]);
                }
            };
            long long tile = blockIdx.x;
            load_idx(tile, ib);
            gather(tile, ib, xc);
            load_idx(tile + gridDim.x, ib);
            uint32_t it = 0;
            for (; tile < ntiles; tile += gridDim.x) {
                gather(tile + gridDim.x, ib, xn);
                load_idx(tile + 2LL * gridDim.x, ib);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int stage = (int)(it % kStages);
                    SG4D_TRACE(tid == 0 && it < 400, it * 5);
                    mbar_wait_warp<40>(lane, bar_empty + 8 * stage, ((it / kStages) & 1u) ^ 1u);
                    SG4D_TRACE(tid == 0 && it < 400, it * 5 + 1);
                    uint8_t *st = smem + stage * SM::kStageBytes;
                    weights_tma(stage, kb);
                    const int col = kb * kKB + 4 * c;
                    float4 w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[j] = *reinterpret_cast<const float4 *>(s_s + j * 64 + col);
                    const float4 t = *reinterpret_cast<const float4 *>(s_p + col);
#pragma unroll
                    for (int i = 0; i < kProdRows; ++i) {
                        const int r = r0 + (PT / 8) * i;
                        float4 v = sa1_y1bn(xc[i], w, t);
                        const bool ok = tile * kTileM + r < p.R;
                        v.x = ok ? fmaxf(v.x, 0.f) : 0.f, v.y = ok ? fmaxf(v.y, 0.f) : 0.f;
                        v.z = ok ? fmaxf(v.z, 0.f) : 0.f, v.w = ok ? fmaxf(v.w, 0.f) : 0.f;
                        put(st, r, v);
                    }
                    SG4D_TRACE(tid == 0 && it < 400, it * 5 + 2);
                    tc::fence_proxy_async_smem();
                    __syncwarp();
                    SG4D_TRACE(tid == 0 && it < 400, it * 5 + 3);
                    if (lane == 0) tc::mbar_arrive(bar_full + 8 * stage);
                    SG4D_TRACE(tid == 0 && it < 400, it * 5 + 4);
                }
#pragma unroll
                for (int i = 0; i < kProdRows; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) xc[i][j] = xn[i][j];
            }
        } else {
        // Flat sequence of chunks q = (my tile index, k-block).  A static ring of PD register buffers keeps PD chunks
        // of loads in flight per thread (memory-level parallelism: 128 threads x PD x 8 x 16 B per SM), so the HBM
        // latency of one chunk is hidden behind the transform + MMA of the previous ones.
        constexpr int PD = (PMODE <= 1 || PMODE == 5) ? 4 : 2;      // modes 2/3 fetch two arrays per item: keep the ring within the register file
        // (tile, k-block) positions of the transform and of the loads PD chunks ahead of it are advanced incrementally
        // (the first version divided a 64-bit chunk counter by nkb four times per chunk: ~100 instructions each)
        RawVec buf[PD][kProdRows];
        long long tile_t = blockIdx.x, tile_i = blockIdx.x;
        int kb_t = 0, kb_i = 0;
        // Everything that depends only on (thread, tile) is computed once per tile, not once per 16-byte item: the row pointers of
        // my rows, their validity bits, the shared-memory offsets of my stores (the first version spent ~95 instructions per
        // float4, two thirds of them 64-bit address arithmetic and bounds predicates; the SM was issue-bound -- DESIGN.md).
        uint32_t soff[kProdRows];
#pragma unroll
        for (int i = 0; i < kProdRows; ++i) {
            const int r = r0 + (PT / 8) * i;
            soff[i] = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
        }
        const float *pa[kProdRows], *pa2[kProdRows];
        const uint8_t *pg[kProdRows];
        unsigned ok_i = 0, ok_t = 0;
        auto row_mask = [&](long long tile) {
            unsigned m = 0;
#pragma unroll
            for (int i = 0; i < kProdRows; ++i) m |= (unsigned)(tile < ntiles && tile * kTileM + r0 + (PT / 8) * i < p.R) << i;
            return m;
        };
        auto set_tile_i = [&]() {
            ok_i = row_mask(tile_i);
            if constexpr (PMODE != 5) {
#pragma unroll
                for (int i = 0; i < kProdRows; ++i) {
                    const long long row = tile_i * kTileM + r0 + (PT / 8) * i;
                    pa[i] = p.op.A + row * p.op.lda + 4 * c;
                    if (PMODE == 3) pa2[i] = p.op.A2 + row * p.op.lda2 + 4 * c;
                    if (PMODE == 2) {
                        const long long g = row >> p.op.logS;
                        pa2[i] = p.op.dsel + g * p.op.ldsel + 4 * c;
                        pg[i] = p.op.garg + g * p.op.ldsel + 4 * c;
                    }
                }
            }
        };
        // mode 5 (rows gathered on the fly): the neighbour indices of my rows, for the tile being loaded and the next one
        int ix_cur[kProdRows], ix_nxt[kProdRows];
        auto load_idx = [&](long long tile, int (&dst)[kProdRows]) {
#pragma unroll
            for (int i = 0; i < kProdRows; ++i) {
                const long long row = tile * kTileM + r0 + (PT / 8) * i;
                dst[i] = (tile < ntiles && row < p.R) ? __ldg(p.op.g.idx + row) : 0;
            }
        };
        if (PMODE == 5) load_idx(tile_i, ix_cur), load_idx(tile_i + gridDim.x, ix_nxt);
        set_tile_i();
        ok_t = ok_i;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        auto issue = [&](RawVec (&dst)[kProdRows]) {
            if (ok_i && !SG4D_DBG(p.dbg_no_load)) {
                const int off = kb_i * kKB;
                const bool colv = off + 4 * c < K;
#pragma unroll
                for (int i = 0; i < kProdRows; ++i) {
                    const bool ok = ((ok_i >> i) & 1u) && colv;
                    if constexpr (PMODE == 5) {
                        dst[i].a = sa2_gather4(p.op.g, tile_i * kTileM + r0 + (PT / 8) * i, ok, off + 4 * c, ix_cur[i]);
                    } else {
                        dst[i].a = ok ? __ldg(reinterpret_cast<const float4 *>(pa[i] + off)) : zero4;
                        if (PMODE == 3 || PMODE == 2) dst[i].a2 = ok ? __ldg(reinterpret_cast<const float4 *>(pa2[i] + off)) : zero4;
                        if (PMODE == 2) dst[i].arg = ok ? __ldg(reinterpret_cast<const uint32_t *>(pg[i] + off)) : 0u;
                    }
                }
            }
            if (++kb_i == nkb) {
                kb_i = 0, tile_i += gridDim.x;
                if (PMODE == 5) {
#pragma unroll
                    for (int i = 0; i < kProdRows; ++i) ix_cur[i] = ix_nxt[i];
                    load_idx(tile_i + gridDim.x, ix_nxt);
                }
                set_tile_i();
            }
        };
        // element transform of one 16-byte item (the prologue of the layer), validity already resolved
        auto transform = [&](const RawVec &r, bool ok, int rr, int col) -> float4 {
            if (PMODE == 0 || PMODE == 5) return r.a;                      // invalid items were loaded as zeros
            if (!ok) return zero4;
            const float4 sv = *reinterpret_cast<const float4 *>(s_s + col), tv = *reinterpret_cast<const float4 *>(s_t + col);
            float4 v;
            v.x = fmaf(r.a.x, sv.x, tv.x), v.y = fmaf(r.a.y, sv.y, tv.y), v.z = fmaf(r.a.z, sv.z, tv.z), v.w = fmaf(r.a.w, sv.w, tv.w);
            if (PMODE == 1) {
                v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
            } else if (PMODE == 2) {
                const uint32_t k = (uint32_t)rr & (uint32_t)(p.op.S - 1);   // 128 % S == 0: the row's slot in its group
                v.x = (((r.arg) & 0xffu) == k ? r.a2.x : 0.f) - v.x;
                v.y = (((r.arg >> 8) & 0xffu) == k ? r.a2.y : 0.f) - v.y;
                v.z = (((r.arg >> 16) & 0xffu) == k ? r.a2.z : 0.f) - v.z;
                v.w = (((r.arg >> 24) & 0xffu) == k ? r.a2.w : 0.f) - v.w;
            } else {   // PMODE == 3
                const float4 pv = *reinterpret_cast<const float4 *>(s_p + col);
                v.x = fmaf(r.a2.x, pv.x, -v.x), v.y = fmaf(r.a2.y, pv.y, -v.y), v.z = fmaf(r.a2.z, pv.z, -v.z), v.w = fmaf(r.a2.w, pv.w, -v.w);
            }
            return v;
        };
#pragma unroll
        for (int j = 0; j < PD; ++j) issue(buf[j]);
        uint32_t it = 0;
        while (tile_t < ntiles) {
#pragma unroll
            for (int j = 0; j < PD; ++j) {
                if (tile_t < ntiles) {
                    const int kb = kb_t;
                    const int stage = (int)(it % kStages);
                    SG4D_TRACE(tid == 0 && it < 400, it * 5);
                    mbar_wait_warp<40>(lane, bar_empty + 8 * stage, ((it / kStages) & 1u) ^ 1u);
                    SG4D_TRACE(tid == 0 && it < 400, it * 5 + 1);
                    uint8_t *st = smem + stage * SM::kStageBytes;
                    weights_tma(stage, kb);
                    const int col = kb * kKB + 4 * c;
                    const bool colv = col < K;
#pragma unroll
                    for (int i = 0; i < kProdRows; ++i) {
                        const float4 v = transform(buf[j][i], ((ok_t >> i) & 1u) && colv, r0 + (PT / 8) * i, col);
                        if (p.lp) {   // bf16 operands: the rounded value IS the operand, no lo tile
                            *reinterpret_cast<float4 *>(st + soff[i]) = bf16_round4(v);
                        } else {
                            float4 hi, lo;
                            split4(v, hi, lo);
                            *reinterpret_cast<float4 *>(st + soff[i]) = hi;
                            *reinterpret_cast<float4 *>(st + SM::kABytes + soff[i]) = lo;
                        }
                    }
                    SG4D_TRACE(tid == 0 && it < 400, it * 5 + 2);
                    tc::fence_proxy_async_smem();   // my smem writes -> visible to the tensor core (async proxy)
                    __syncwarp();
                    SG4D_TRACE(tid == 0 && it < 400, it * 5 + 3);
                    if (lane == 0) tc::mbar_arrive(bar_full + 8 * stage);
                    SG4D_TRACE(tid == 0 && it < 400, it * 5 + 4);
                    issue(buf[j]);   // refill this ring slot
                    if (++kb_t == nkb) {
                        kb_t = 0, tile_t += gridDim.x;
                        ok_t = row_mask(tile_t);
                    }
                    ++it;
                }
            }
        }
        }
    } else if (warp == kProdWarps) {
        // =============================================================== MMA issuer
        constexpr uint32_t idesc = tc::umma_idesc_tf32(kTileM, N);
        uint32_t it = 0;
        long long ti = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
            const int acc = (int)(ti & 1);
            SG4D_TRACE(lane == 0 && ti < 256, 4096 + (int)ti * 2);
            mbar_wait_warp<32>(lane, bar_tempty + 8 * acc, (uint32_t)(((ti >> 1) & 1) ^ 1));
            SG4D_TRACE(lane == 0 && ti < 256, 4096 + (int)ti * 2 + 1);
            tc::tc_fence_after_sync();
            const uint32_t d_tmem = tmem + (uint32_t)(acc * N);
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int stage = (int)(it % kStages);
                SG4D_TRACE(lane == 0 && it < 400, 2048 + it * 3);
                mbar_wait_warp<40>(lane, bar_full + 8 * stage, (uint32_t)((it / kStages) & 1));
                SG4D_TRACE(lane == 0 && it < 400, 2048 + it * 3 + 1);
                tc::tc_fence_after_sync();
                if (lane == 0) {
                    const uint32_t a_hi = smem_base + stage * SM::kStageBytes, a_lo = a_hi + SM::kABytes;
                    const uint32_t w_hi = a_hi + 2 * SM::kABytes, w_lo = w_hi + SM::kWBytes;
#pragma unroll
                    for (int ks = 0; ks < kKB / 8 && !SG4D_DBG(p.dbg_no_mma); ++ks) {
                        const uint64_t dah = tc::umma_desc_k_sw128(a_hi + ks * 32), dal = tc::umma_desc_k_sw128(a_lo + ks * 32);
                        const uint64_t dwh = tc::umma_desc_k_sw128(w_hi + ks * 32), dwl = tc::umma_desc_k_sw128(w_lo + ks * 32);
                        if (p.lp) {
                            tc::umma_tf32(d_tmem, dah, dwh, idesc, (kb | ks) != 0);
                        } else {
                            tc::umma_tf32(d_tmem, dal, dwh, idesc, (kb | ks) != 0);   // small terms first
                            tc::umma_tf32(d_tmem, dah, dwl, idesc, 1u);
                            tc::umma_tf32(d_tmem, dah, dwh, idesc, 1u);
                        }
                    }
                    tc::umma_commit(bar_empty + 8 * stage);                      // smem stage free when these finish
                    if (kb == nkb - 1) tc::umma_commit(bar_tfull + 8 * acc);     // accumulator ready
                }
                SG4D_TRACE(lane == 0 && it < 400, 2048 + it * 3 + 2);
                __syncwarp();
            }
        }
    } else {
        // =============================================================== epilogue
        const int e = tid - (kProdThreads + 32);   // 0..255
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int half = e >> 7;                   // which half of the 32-column chunks this warp moves
        const int row = q * 32 + lane;             // my accumulator row
        // column-owner mapping for the statistics / group pass: N = 128 -> 2 threads per column (64 rows each),
        // N = 64 -> 4 threads per column (32 rows each); a max-pool group never straddles two owners
        const int col = e & (N - 1);
        const int rbeg = (e / N) * (kTileM / (kEpiThreads / N));
        const int rcnt = kTileM / (kEpiThreads / N);
        bool want_max = true;
        if (EMODE == 0 && p.S > 0) want_max = __ldg(q_gamma + col) >= 0.f;
        double dacc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        // EMODE 3: thread (cg, rs) owns channels 4cg..4cg+3 and rows 16rs..16rs+15 of every tile
        float *xa_s = reinterpret_cast<float *>(smem + kStages * SM::kStageBytes + SM::kCBytes + SM::kConst);
        double *sacc = reinterpret_cast<double *>(smem + kStages * SM::kStageBytes + SM::kCBytes + SM::kConst + SM::kXaBytes);
        const int cg = e & 15, rs = e >> 4;
        float4 t1v = make_float4(0.f, 0.f, 0.f, 0.f);
        float sa[32];
        int since = 0, ixn = 0;
        if constexpr (EMODE == 3) {   // W1s was staged into s_p (free in PMODE 2) before the role split
            t1v = __ldg(reinterpret_cast<const float4 *>(p.op.g.t1 + 4 * cg));
#pragma unroll
            for (int u = 0; u < 32; ++u) sa[u] = 0.f, sacc[u * kEpiThreads + e] = 0.0;
            const long long row = (long long)blockIdx.x * kTileM + e;
            ixn = row < p.R ? __ldg(p.op.g.idx + row) : 0;
        }
        long long ti = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
            const int acc = (int)(ti & 1);
            float xrow[8];
            if constexpr (EMODE == 3) {   // my row of this tile (one row per epilogue thread), issued before the wait
                const long long row = tile * kTileM + e;
                sa1_gather_row(p.op.g, row, row < p.R, ixn, xrow);
                const long long rown = row + (long long)gridDim.x * kTileM;
                ixn = rown < p.R ? __ldg(p.op.g.idx + rown) : 0;
            }
            SG4D_TRACE(e == 0 && ti < 256, 5120 + (int)ti * 5);
            mbar_wait_warp<128>(lane, bar_tfull + 8 * acc, (uint32_t)((ti >> 1) & 1));
            SG4D_TRACE(e == 0 && ti < 256, 5120 + (int)ti * 5 + 1);
            tc::tc_fence_after_sync();
            // two 32-column chunks per wait: the second tcgen05.ld overlaps the first one's latency
            if constexpr (EMODE == 3) {
#pragma unroll
                for (int ch = 0; ch < N / 32; ++ch) {
                    uint32_t v[32];
                    tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * N + ch * 32), v);
                    tc::tmem_ld_wait();
                    float4 *dst = reinterpret_cast<float4 *>(Cs + row * SM::kCStride + ch * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                             __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
            } else
#pragma unroll
            for (int c2 = 0; c2 < N / 32 / (kEpiThreads / 128); c2 += 2) {
                const int ch = (kEpiThreads / 128) * c2 + half, ch2 = ch + (kEpiThreads / 128);
                uint32_t v[32], w[32];
                tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * N + ch * 32), v);
                tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * N + ch2 * 32), w);
                tc::tmem_ld_wait();
                float4 *dst = reinterpret_cast<float4 *>(Cs + row * SM::kCStride + ch * 32);
                float4 *dst2 = reinterpret_cast<float4 *>(Cs + row * SM::kCStride + ch2 * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    dst2[j] = make_float4(__uint_as_float(w[4 * j]), __uint_as_float(w[4 * j + 1]),
                                          __uint_as_float(w[4 * j + 2]), __uint_as_float(w[4 * j + 3]));
            }
            tc::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(bar_tempty + 8 * acc);   // accumulator may be overwritten
            if constexpr (EMODE == 3) {
                *reinterpret_cast<float4 *>(xa_s + e * 8) = make_float4(xrow[0], xrow[1], xrow[2], xrow[3]);
                *reinterpret_cast<float4 *>(xa_s + e * 8 + 4) = make_float4(xrow[4], xrow[5], xrow[6], xrow[7]);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");

            SG4D_TRACE(e == 0 && ti < 256, 5120 + (int)ti * 5 + 2);
            const long long row0 = tile * kTileM;
            const int nvalid = SG4D_DBG(p.dbg_no_epi) ? 0 : (int)min((long long)kTileM, p.R - row0);
            constexpr int kVecPerRow = N / 4;
            constexpr int kRowStep = kEpiThreads / kVecPerRow;   // rows between two float4 items of one thread
            if constexpr (EMODE == 3) {
                const int rend = min(16 * rs + 16, nvalid);
                for (int r = 16 * rs; r < rend; ++r) {
                    const float4 xa0 = *reinterpret_cast<const float4 *>(xa_s + r * 8), xa1 = *reinterpret_cast<const float4 *>(xa_s + r * 8 + 4);
                    const float xa[8] = {xa0.x, xa0.y, xa0.z, xa0.w, xa1.x, xa1.y, xa1.z, xa1.w};
                    float4 d = *reinterpret_cast<const float4 *>(Cs + r * SM::kCStride + 4 * cg);
                    const float4 y = sa1_y1bn_smem(xa, s_p, 4 * cg, t1v);
                    d.x = y.x > 0.f ? d.x : 0.f, d.y = y.y > 0.f ? d.y : 0.f, d.z = y.z > 0.f ? d.z : 0.f, d.w = y.w > 0.f ? d.w : 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        sa[4 * j + 0] = fmaf(d.x, xa[j], sa[4 * j + 0]), sa[4 * j + 1] = fmaf(d.y, xa[j], sa[4 * j + 1]);
                        sa[4 * j + 2] = fmaf(d.z, xa[j], sa[4 * j + 2]), sa[4 * j + 3] = fmaf(d.w, xa[j], sa[4 * j + 3]);
                    }
                }
                if (++since == 16) {   // fp32 over 256 rows per thread, fp64 across
                    since = 0;
#pragma unroll
                    for (int u = 0; u < 32; ++u) sacc[u * kEpiThreads + e] += (double)sa[u], sa[u] = 0.f;
                }
            } else if (EMODE == 1) {
                // ReLU mask of the inner layer from its pre-activation E (one coalesced read of E), store of dz,
                // and the two BatchNorm-backward reductions -- every thread owns 4 fixed columns in this pass
                const int cc = (e % kVecPerRow) * 4;
                const float4 sv = __ldg(reinterpret_cast<const float4 *>(q_es + cc)), tv = __ldg(reinterpret_cast<const float4 *>(q_et + cc));
                const float4 iv = __ldg(reinterpret_cast<const float4 *>(q_ei + cc)), mv = __ldg(reinterpret_cast<const float4 *>(q_em + cc));
                float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
                for (int rb = e / kVecPerRow; rb < nvalid; rb += 4 * kRowStep) {
                    float4 ev[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = rb + u * kRowStep;
                        ev[u] = r < nvalid ? __ldg(reinterpret_cast<const float4 *>(q_E + (row0 + r) * lde + cc)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = rb + u * kRowStep;
                        if (r < nvalid) {
                            float4 dv = *reinterpret_cast<const float4 *>(Cs + r * SM::kCStride + cc);
                            dv.x = fmaf(ev[u].x, sv.x, tv.x) > 0.f ? dv.x : 0.f;
                            dv.y = fmaf(ev[u].y, sv.y, tv.y) > 0.f ? dv.y : 0.f;
                            dv.z = fmaf(ev[u].z, sv.z, tv.z) > 0.f ? dv.z : 0.f;
                            dv.w = fmaf(ev[u].w, sv.w, tv.w) > 0.f ? dv.w : 0.f;
                            *reinterpret_cast<float4 *>(p.Y + (row0 + r) * p.ldy + q_ycol0 + cc) = dv;
                            a0.x += dv.x, a0.y += dv.y, a0.z += dv.z, a0.w += dv.w;
                            a1.x = fmaf(dv.x, fmaf(ev[u].x, iv.x, mv.x), a1.x);
                            a1.y = fmaf(dv.y, fmaf(ev[u].y, iv.y, mv.y), a1.y);
                            a1.z = fmaf(dv.z, fmaf(ev[u].z, iv.z, mv.z), a1.z);
                            a1.w = fmaf(dv.w, fmaf(ev[u].w, iv.w, mv.w), a1.w);
                        }
                    }
                }
                dacc[0] += a0.x, dacc[1] += a0.y, dacc[2] += a0.z, dacc[3] += a0.w;
                dacc[4] += a1.x, dacc[5] += a1.y, dacc[6] += a1.z, dacc[7] += a1.w;
            } else if (p.Y) {   // coalesced store of the tile: 16 bytes per thread per step
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (EMODE == 2 && q_bias) bv = __ldg(reinterpret_cast<const float4 *>(q_bias + (e % kVecPerRow) * 4));   // kEpiThreads % kVecPerRow == 0: my columns are fixed
                // my column group is fixed (kEpiThreads % kVecPerRow == 0); rows rb, rb + kRowStep, ...  Eight shared-memory
                // loads are issued before the first store (the one-at-a-time loop spent 2.7 k cycles per tile on LDS latency)
                const int cc = (e % kVecPerRow) * 4, rb = e / kVecPerRow;
                const float *cs = Cs + rb * SM::kCStride + cc;
                float *yp = p.Y + (row0 + rb) * p.ldy + q_ycol0 + cc;
                const long long ystep = (long long)kRowStep * p.ldy;
                constexpr int kPerThread = kTileM / kRowStep;
#pragma unroll
                for (int b8 = 0; b8 < kPerThread; b8 += 8) {
                    float4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4 *>(cs + (b8 + u) * kRowStep * SM::kCStride);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (EMODE == 2) v[u].x += bv.x, v[u].y += bv.y, v[u].z += bv.z, v[u].w += bv.w;
                        if (rb + (b8 + u) * kRowStep < nvalid) *reinterpret_cast<float4 *>(yp + (b8 + u) * ystep) = v[u];
                    }
                }
            }
            SG4D_TRACE(e == 0 && ti < 256, 5120 + (int)ti * 5 + 3);
            if (EMODE == 0) {   // per-channel statistics (+ group max/min) by the column owners
                const int smask = p.S > 0 ? p.S - 1 : 0;
                float s = 0.f, sq = 0.f;
                if (nvalid == kTileM && (p.S == 0 || p.S >= 8)) {
                    // fast path (full tile): 8 rows per batch, branch-free, independent chains inside a batch
                    const float sgn = want_max ? 1.f : -1.f;    // arg-min = arg-max of the negated values
                    float s1 = 0.f, q1 = 0.f, best = 0.f;
                    int bi = 0;
                    for (int r = rbeg; r < rbeg + rcnt; r += 8) {
                        float y[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) y[u] = Cs[(r + u) * SM::kCStride + col];
                        s += (y[0] + y[1]) + (y[2] + y[3]);
                        s1 += (y[4] + y[5]) + (y[6] + y[7]);
                        sq = fmaf(y[0], y[0], fmaf(y[1], y[1], fmaf(y[2], y[2], fmaf(y[3], y[3], sq))));
                        q1 = fmaf(y[4], y[4], fmaf(y[5], y[5], fmaf(y[6], y[6], fmaf(y[7], y[7], q1))));
                        if (p.S > 0) {
                            // tournament over the batch; ties keep the lower row (first occurrence, like max_pool2d)
                            float c[8];
                            int ix[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) c[u] = y[u] * sgn, ix[u] = u;
#pragma unroll
                            for (int w = 1; w < 8; w <<= 1)
#pragma unroll
                                for (int u = 0; u < 8; u += 2 * w) {
                                    const bool gt = c[u + w] > c[u];
                                    c[u] = gt ? c[u + w] : c[u];
                                    ix[u] = gt ? ix[u + w] : ix[u];
                                }
                            const int k0 = r & smask;                 // position of this batch inside its group
                            const bool take = (k0 == 0) || (c[0] > best);
                            best = take ? c[0] : best;
                            bi = take ? k0 + ix[0] : bi;
                            if (((r + 8) & smask) == 0) {             // last batch of the group (warp-uniform)
                                const long long g = (row0 + r) >> p.logS;
                                q_gsel[g * ldg + col] = best * sgn;
                                q_garg[g * ldg + col] = (uint8_t)bi;
                            }
                        }
                    }
                    s += s1, sq += q1;
                } else {   // partial last tile or tiny groups: simple row loop
                    const int rend = min(rbeg + rcnt, nvalid);
                    float best = 0.f;
                    int bi = 0;
                    for (int r = rbeg; r < rend; ++r) {
                        const float y = Cs[r * SM::kCStride + col];
                        s += y, sq = fmaf(y, y, sq);
                        if (p.S > 0) {
                            const int k = r & smask;
                            const bool better = (k == 0) || (want_max ? (y > best) : (y < best));
                            best = better ? y : best;
                            bi = better ? k : bi;
                            if (k == smask) {
                                const long long g = (row0 + r) >> p.logS;
                                q_gsel[g * ldg + col] = best;
                                q_garg[g * ldg + col] = (uint8_t)bi;
                            }
                        }
                    }
                }
                dacc[0] += (double)s, dacc[1] += (double)sq;
            }
            SG4D_TRACE(e == 0 && ti < 256, 5120 + (int)ti * 5 + 4);
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");   // Cs is reused by the next tile
        }
        if constexpr (EMODE == 3) {
            // S1 partial of this CTA: fold the 8 row slices in a fixed order -> (64 channels, 8 inputs) fp64
#pragma unroll
            for (int u = 0; u < 32; ++u) sacc[u * kEpiThreads + e] += (double)sa[u];
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            if (rs == 0) {
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    double t = 0.0;
                    for (int q2 = 0; q2 < 8; ++q2) t += sacc[u * kEpiThreads + q2 * 16 + cg];
                    p.s1part[((size_t)blockIdx.x * 64 + 4 * cg + (u & 3)) * 8 + (u >> 2)] = t;
                }
            }
        } else if (EMODE == 0) {
            q_partial[((size_t)blockIdx.x * kEpiThreads + e) * 2 + 0] = dacc[0];
            q_partial[((size_t)blockIdx.x * kEpiThreads + e) * 2 + 1] = dacc[1];
        } else if (EMODE == 1) {
            // fold the per-thread (4 columns x 2) fp64 sums into one pair per column, in a fixed order
            constexpr int kVecPerRow = N / 4;
            double *sd = reinterpret_cast<double *>(Cs);
#pragma unroll
            for (int j = 0; j < 8; ++j) sd[e * 8 + j] = dacc[j];
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            double t0 = 0.0, t1 = 0.0;
            if (e < N) {
                for (int u = 0; u < kEpiThreads / kVecPerRow; ++u) {
                    const int t = (e >> 2) + kVecPerRow * u;
                    t0 += sd[t * 8 + (e & 3)], t1 += sd[t * 8 + 4 + (e & 3)];
                }
            }
            q_partial[((size_t)blockIdx.x * kEpiThreads + e) * 2 + 0] = t0;
            q_partial[((size_t)blockIdx.x * kEpiThreads + e) * 2 + 1] = t1;
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == kProdWarps) tc::tmem_dealloc(tmem, 2 * N);
}

// ================================================================================================
// T2: weight gradient  D[M=128, N] = sum over row tiles of P(tile)^T * Q(tile),  K = rows.
// Both operands are staged ROW-major exactly like a T1 A tile (coalesced loads, 16-byte swizzled stores) and
// handed to the tensor core as MN-major operands (the 128-byte lines run along M / N, the 8-line atoms along K).
// Each CTA owns a contiguous range of row tiles, keeps D in TMEM for the whole kernel and finally writes its
// partial (M x N) to HBM; wgrad_reduce_kernel sums the partials in a fixed order (deterministic).
struct WgradArgs {
    Operand P, Q;            // P: (R, 128) -> M = 128 channels (zero beyond P.ncols); Q: (R, N)
    long long R;
    float *partial;          // (gridDim.x, 128, N)
    uint32_t d_lbo, d_sbo, d_type, d_kstep;   // MN-major descriptor fields (bytes / layout type)
    // block grid (gridDim.y = mblocks * nnb): CTA (x, y) accumulates the (128 x N) block (y / nnb, y % nnb) of a larger dW
    int nnb, mtot, ktot;
    int dbg_no_mma, dbg_no_load;
    int lp;                  // 1: bf16 operands, one product per k-step
};

__device__ __forceinline__ void shift_operand(Operand &o, int c0, int total, int width) {
    o.A += c0;
    if (o.A2) o.A2 += c0;
    if (o.s) o.s += c0;
    if (o.t) o.t += c0;
    if (o.p) o.p += c0;
    if (o.dsel) o.dsel += c0, o.garg += c0;
    o.ncols = max(0, min(width, total - c0));
}

__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t type) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;   // next 32-element chunk along M/N
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;   // next 8-line atom along K
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)type << 61;
    return d;
}

// MN-major operand tile for 32-bit elements: the tensor core only accepts the "128-byte swizzle with 32-byte
// atomicity" layout here (layout type 1; with the 16-byte-atom types 0/2 a tf32 MN-major MMA yields zeros --
// measured, profiles/README.md).  Element (k-line r in [0,32), float4 column c4 along M/N):
//   32-channel block (c4 / 8) * 4096  +  line r * 128  +  32-byte chunk ((c4 % 8) / 2) ^ (r % 4)  +  (c4 % 2) * 16
__device__ __forceinline__ uint32_t mn_b32_offset(int r, int c4) {
    return (uint32_t)((c4 >> 3) * 4096 + r * 128 + (((((c4 & 7) >> 1) ^ (r & 3)) << 5) | ((c4 & 1) << 4)));
}

template <int N>
struct WgSmem {
    static constexpr int kStg = N <= 32 ? 5 : (N <= 64 ? 4 : (N <= 128 ? 3 : 2));   // smem stages that fit in 227 KB
    static constexpr int kPBytes = 4 * 4096;                // 4 chunks of 32 channels x (32 k-lines x 128 B)
    static constexpr int kQBytes = (N / 32) * 4096;
    static constexpr int kStageBytes = 2 * kPBytes + 2 * kQBytes;
    static constexpr int kConst = 6 * 256 * 4;
    // pool_bwd_dw (PMODE 2) gathers dsel / garg through L1: it keeps 2 stages so that the L1 carve-out stays large
    static constexpr int stages(int pmode) { return pmode == 2 ? 2 : kStg; }
    static constexpr int total(int pmode) { return 1024 + stages(pmode) * kStageBytes + kConst + 128; }
};

// MV = valid channels of P (64 or 128): with 64 the upper half of the P tiles is zeroed once and never rewritten, so
// the producers only transform the 16 valid float4 columns of each row.
// PMODE 6 ("Gram", N = MV = 64): P IS Q -- the producers write the Q operand into the P tiles and the tensor core reads the
// same tiles as both operands, so D[0..63] = Q^T Q; channel 64 of P is a constant 1, so D[64] = the column sums of Q.
template <int N, int PMODE, int QMODE, int MV, bool LP = false>
__global__ void __launch_bounds__(kMlpThreadsT2, 1) wgrad_kernel(WgradArgs p) {
    // (registers are allocated per 4 warps: 21 warps count as 24, which caps this kernel at 80 registers per thread -- the
    //  bf16 variant is a separate instantiation so that its branches do not add to the pressure of the fp32 one)
    constexpr int kProdThreads = kProdThreadsT2, kProdWarps = kProdThreads / 32, kMlpThreads = kMlpThreadsT2;
    using SM = WgSmem<N>;
    constexpr int kStages = SM::stages(PMODE);
    // D lives in TMEM in RUNS of kRunKb k-blocks (512 rows); two accumulator buffers alternate, and the epilogue warps
    // fold a finished run into the CTA's fp32 partial in global memory (round-to-nearest adds) while the next run
    // accumulates.  The tensor core's own fp32 accumulation loses up to an ulp per step in one direction: over the
    // ~7000 rows a CTA owns at benchmark size that bias reached 1.7e-4 of the result (tests/test_gpu_full_size.py).
    constexpr int kTmemCols = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
    constexpr int kRunKb = 16;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
    float *cst = reinterpret_cast<float *>(smem + kStages * SM::kStageBytes);
    float *ps = cst, *pt = cst + 256, *pp = cst + 512, *qs = cst + 768, *qt = cst + 1024, *qp = cst + 1280;
    __shared__ __align__(8) uint64_t s_bar[2 * 5 + 4];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // (block of a larger dW: operands shifted in LOCAL copies -- writing to the parameter struct would move it to local memory
    //  and turn every p.* access of the hot loops into a local load)
    Operand OP = p.P, OQ = p.Q;
    float *partial_out = p.partial;
    if (gridDim.y > 1) {
        const int mb = blockIdx.y / p.nnb, nb = blockIdx.y % p.nnb;
        shift_operand(OP, mb * kTileM, p.mtot, kTileM);
        shift_operand(OQ, nb * N, p.ktot, N);
        partial_out += (size_t)blockIdx.y * gridDim.x * kTileM * N;
    }
    const long long ntiles = (p.R + kTileM - 1) / kTileM;
    const long long per = (ntiles + gridDim.x - 1) / gridDim.x;
    const long long t_beg = min(ntiles, (long long)blockIdx.x * per), t_end = min(ntiles, t_beg + per);
    const long long nkb_total = (t_end - t_beg) * (kTileM / kKB);   // k-blocks of 32 rows
    const uint32_t bar_full = smem_u32(&s_bar[0]), bar_empty = smem_u32(&s_bar[kStages]);
    const uint32_t bar_afull = smem_u32(&s_bar[2 * kStages]), bar_aempty = smem_u32(&s_bar[2 * kStages + 2]);

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            tc::mbar_init(bar_full + 8 * s, kProdWarps);
            tc::mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(bar_afull + 8 * a, 1);
            tc::mbar_init(bar_aempty + 8 * a, 4);
        }
        tc::mbar_fence_init();
    }
    if (warp == kProdWarps) tc::tmem_alloc(smem_u32(&s_tmem), 2 * kTmemCols);
    for (int k = tid; k < 256; k += kMlpThreads) {
        ps[k] = (PMODE != 0 && PMODE != 6 && k < OP.ncols) ? OP.s[k] : 0.f;
        pt[k] = (PMODE != 0 && PMODE != 6 && k < OP.ncols) ? OP.t[k] : 0.f;
        pp[k] = (PMODE == 3 && k < OP.ncols) ? OP.p[k] : 0.f;
        if (QMODE == 4) {      // qs[0..511] (= qs | qt) = W1s (input-major, 8 x 64), qp[0..63] = t1
            qs[k] = OQ.g.w1s[k], qt[k] = OQ.g.w1s[256 + k];
            qp[k] = k < 64 ? OQ.g.t1[k] : 0.f;
        } else {
            qs[k] = (QMODE == 1 && k < OQ.ncols) ? OQ.s[k] : 0.f;
            qt[k] = (QMODE == 1 && k < OQ.ncols) ? OQ.t[k] : 0.f;
            qp[k] = 0.f;
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = s_tmem;

    if (warp < kProdWarps) {
        // producers: k-block = 32 rows.  P slab: 32 rows x 128 channels = 1024 float4 (8 per thread);
        // Q slab: 32 rows x N channels = 8N float4.  Two k-blocks of loads are kept in flight per thread.
        constexpr int kQVec = N / 4;                    // float4 per Q row
        constexpr int kQItems = (32 * kQVec + kProdThreads - 1) / kProdThreads;
        constexpr int PD = (N > 128) ? 1 : 2;           // k-blocks of loads in flight per thread (register ring, no spills)
        constexpr int kPVec = MV / 4;                   // valid float4 per P row
        constexpr bool kGram = PMODE == 6;
        static_assert(!kGram || (N == 64 && MV == 64), "Gram mode: 64 x 64");
        constexpr int kPItems = kGram ? 1 : 32 * kPVec / kProdThreads;   // float4 of the P slab per thread (Gram: none, ring unused)
        RawVec pv[PD][kPItems], qv[PD][kQItems];
        if (MV < 128) {   // channels MV..127 of every P tile (hi and lo, all stages) stay zero for the whole kernel
            constexpr int kZero = (4 - MV / 32) * 4096;
            for (int s2 = 0; s2 < kStages; ++s2)
                for (int t = 0; t < 2; ++t) {
                    uint8_t *zb = smem + s2 * SM::kStageBytes + t * SM::kPBytes + (MV / 32) * 4096;
                    for (int o = tid * 16; o < kZero; o += kProdThreads * 16) {
                        // Gram: channel 64 (first float of 32-byte chunk (line & 3) of every line of chunk 2) of the hi tile = 1
                        const bool one = kGram && t == 0 && o < 4096 && (o & 127) == (((o >> 7) & 3) << 5);
                        *reinterpret_cast<float4 *>(zb + o) = make_float4(one ? 1.f : 0.f, 0.f, 0.f, 0.f);
                    }
                }
        }
        // Q gathered on the fly (modes 4 / 5): qd[i].arg carries the neighbour index of the row this ring slot gathers
        // next; it is loaded one ring round (PD k-blocks) before the gather that depends on it
        auto q_row_idx = [&](long long kbk, int i) -> uint32_t {
            const int item = tid + kProdThreads * i;
            const long long row = t_beg * kTileM + kbk * kKB + item / kQVec;
            return (kbk < nkb_total && item < 32 * kQVec && row < p.R) ? (uint32_t)__ldg(OQ.g.idx + row) : 0u;
        };
        auto issue = [&](long long kbk, RawVec (&pd)[kPItems], RawVec (&qd)[kQItems]) {
            if (kbk < nkb_total && !SG4D_DBG(p.dbg_no_load)) {
                const long long row_base = t_beg * kTileM + kbk * kKB;
                if constexpr (!kGram) {
#pragma unroll
                    for (int i = 0; i < kPItems; ++i) {
                        const int item = tid + kProdThreads * i;
                        op_load<PMODE>(OP, row_base + item / kPVec, p.R, 4 * (item % kPVec), pd[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < kQItems; ++i) {
                    const int item = tid + kProdThreads * i;
                    const int r = item / kQVec, c4 = item % kQVec;
                    if (item < 32 * kQVec) {
                        if constexpr (QMODE == 4) {
                            float x[8];
                            sa1_gather_row(OQ.g, row_base + r, row_base + r < p.R, (int)qd[i].arg, x);
                            qd[i].a = make_float4(x[0], x[1], x[2], x[3]), qd[i].a2 = make_float4(x[4], x[5], x[6], x[7]);
                        } else if constexpr (QMODE == 5) {
                            qd[i].a = sa2_gather4(OQ.g, row_base + r, row_base + r < p.R, 4 * c4, (int)qd[i].arg);
                        } else {
                            op_load<QMODE>(OQ, row_base + r, p.R, 4 * c4, qd[i]);
                        }
                    }
                }
            }
            if constexpr (QMODE == 4 || QMODE == 5) {
#pragma unroll
                for (int i = 0; i < kQItems; ++i) qd[i].arg = q_row_idx(kbk + PD, i);
            }
        };
        if constexpr (QMODE == 4 || QMODE == 5) {
#pragma unroll
            for (int j = 0; j < PD; ++j)
#pragma unroll
                for (int i = 0; i < kQItems; ++i) qv[j][i].arg = q_row_idx(j, i);
        }
#pragma unroll
        for (int j = 0; j < PD; ++j) issue(j, pv[j], qv[j]);
        for (long long k0 = 0; k0 < nkb_total; k0 += PD) {
#pragma unroll
            for (int j = 0; j < PD; ++j) {
                const long long kbk = k0 + j;
                if (kbk < nkb_total) {
                    const long long row_base = t_beg * kTileM + kbk * kKB;
                    const int stage = (int)(kbk % kStages);
                    mbar_wait_warp<40>(lane, bar_empty + 8 * stage, (uint32_t)(((kbk / kStages) & 1) ^ 1));
                    uint8_t *st = smem + stage * SM::kStageBytes;
                    uint8_t *stq = kGram ? st : st + 2 * SM::kPBytes;                       // where the Q operand goes
                    constexpr int kQLo = kGram ? SM::kPBytes : SM::kQBytes;                 // hi -> lo tile distance
#pragma unroll
                    for (int i = 0; i < (kGram ? 0 : kPItems); ++i) {
                        const int item = tid + kProdThreads * i;
                        const int r = item / kPVec, pc4 = item % kPVec;
                        const float4 v = op_apply<(kGram ? 0 : PMODE)>(OP, pv[j][i], row_base + r, p.R, 4 * pc4, ps, pt, pp);
                        const uint32_t off = mn_b32_offset(r, pc4);
                        if constexpr (LP) {
                            *reinterpret_cast<float4 *>(st + off) = bf16_round4(v);
                        } else {
                            float4 hi, lo;
                            split4_rn(v, hi, lo);
                            *reinterpret_cast<float4 *>(st + off) = hi;
                            *reinterpret_cast<float4 *>(st + SM::kPBytes + off) = lo;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < kQItems; ++i) {
                        const int item = tid + kProdThreads * i;
                        if (item < 32 * kQVec) {
                            const int r = item / kQVec, c4 = item % kQVec;
                            float4 v;
                            if constexpr (QMODE == 4) {     // relu(W1s x + t1), the second layer's input, recomputed
                                const float x[8] = {qv[j][i].a.x, qv[j][i].a.y, qv[j][i].a.z, qv[j][i].a.w,
                                                    qv[j][i].a2.x, qv[j][i].a2.y, qv[j][i].a2.z, qv[j][i].a2.w};
                                v = sa1_y1bn_smem(x, qs, 4 * c4, *reinterpret_cast<const float4 *>(qp + 4 * c4));
                                const bool ok = row_base + r < p.R;
                                v.x = ok ? fmaxf(v.x, 0.f) : 0.f, v.y = ok ? fmaxf(v.y, 0.f) : 0.f;
                                v.z = ok ? fmaxf(v.z, 0.f) : 0.f, v.w = ok ? fmaxf(v.w, 0.f) : 0.f;
                            } else {
                                v = op_apply<(QMODE == 5 ? 0 : QMODE)>(OQ, qv[j][i], row_base + r, p.R, 4 * c4, qs, qt, qp);
                            }
                            const uint32_t off = mn_b32_offset(r, c4);
                            if constexpr (LP) {
                                *reinterpret_cast<float4 *>(stq + off) = bf16_round4(v);
                            } else {
                                float4 hi, lo;
                                split4_rn(v, hi, lo);
                                *reinterpret_cast<float4 *>(stq + off) = hi;
                                *reinterpret_cast<float4 *>(stq + kQLo + off) = lo;
                            }
                        }
                    }
                    tc::fence_proxy_async_smem();   // my smem writes -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(bar_full + 8 * stage);
                    issue(kbk + PD, pv[j], qv[j]);
                }
            }
        }
    } else if (warp == kProdWarps) {
        // MMA issuer: A = P^T (M = 128 channels, MN-major), B = Q^T (N channels, MN-major), K = 8 rows per MMA
        constexpr uint32_t idesc = tc::umma_idesc_tf32(kTileM, N) | (1u << 15) | (1u << 16);
        for (long long kbk = 0; kbk < nkb_total; ++kbk) {
            const int stage = (int)(kbk % kStages);
            const long long run = kbk / kRunKb;
            const int pos = (int)(kbk % kRunKb), buf = (int)(run & 1);
            if (pos == 0 && run >= 2) {   // the epilogue must have drained this buffer's previous run
                mbar_wait_warp<40>(lane, bar_aempty + 8 * buf, (uint32_t)(((run >> 1) & 1) ^ 1));
                tc::tc_fence_after_sync();
            }
            mbar_wait_warp<40>(lane, bar_full + 8 * stage, (uint32_t)((kbk / kStages) & 1));
            tc::tc_fence_after_sync();
            if (lane == 0) {
                const uint32_t d_tmem = tmem + (uint32_t)(buf * kTmemCols);
                const uint32_t p_hi = smem_base + stage * SM::kStageBytes, p_lo = p_hi + SM::kPBytes;
                const uint32_t q_hi = PMODE == 6 ? p_hi : p_hi + 2 * SM::kPBytes, q_lo = PMODE == 6 ? p_lo : q_hi + SM::kQBytes;
#pragma unroll
                for (int ks = 0; ks < kKB / 8 && !SG4D_DBG(p.dbg_no_mma); ++ks) {   // 8 rows = one 1024-byte atom per chunk
                    const uint32_t ko = ks * p.d_kstep;
                    const uint64_t dph = umma_desc_mn(p_hi + ko, p.d_lbo, p.d_sbo, p.d_type), dpl = umma_desc_mn(p_lo + ko, p.d_lbo, p.d_sbo, p.d_type);
                    const uint64_t dqh = umma_desc_mn(q_hi + ko, p.d_lbo, p.d_sbo, p.d_type), dql = umma_desc_mn(q_lo + ko, p.d_lbo, p.d_sbo, p.d_type);
                    if constexpr (LP) {
                        tc::umma_tf32(d_tmem, dph, dqh, idesc, (pos | ks) != 0);
                    } else {
                        tc::umma_tf32(d_tmem, dpl, dql, idesc, (pos | ks) != 0);   // 4 products: the tensor pipe has the time,
                        tc::umma_tf32(d_tmem, dpl, dqh, idesc, 1u);               // and lo*lo is the largest error term left
                        tc::umma_tf32(d_tmem, dph, dql, idesc, 1u);
                        tc::umma_tf32(d_tmem, dph, dqh, idesc, 1u);
                    }
                }
                tc::umma_commit(bar_empty + 8 * stage);
                if (pos == kRunKb - 1 || kbk == nkb_total - 1) tc::umma_commit(bar_afull + 8 * buf);
            }
            __syncwarp();
        }
    } else {
        // epilogue: after every run, partial (128 x N) of this CTA (+)= the run's accumulator (zeros when it had no tile)
        const int q = warp & 3, row = q * 32 + lane;
        float *out = partial_out + ((size_t)blockIdx.x * kTileM + row) * N;
        const long long nruns = (nkb_total + kRunKb - 1) / kRunKb;
        for (long long run = 0; run < nruns; ++run) {
            const int buf = (int)(run & 1);
            mbar_wait_warp<500>(lane, bar_afull + 8 * buf, (uint32_t)((run >> 1) & 1));
            tc::tc_fence_after_sync();
#pragma unroll
            for (int ch = 0; ch < N / 32; ++ch) {
                uint32_t v[32];
                tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kTmemCols + ch * 32), v);
                tc::tmem_ld_wait();
                float4 *o4 = reinterpret_cast<float4 *>(out + ch * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 a = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                           __uint_as_float(v[4 * j + 3]));
                    if (run > 0) {
                        const float4 o = o4[j];
                        a.x += o.x, a.y += o.y, a.z += o.z, a.w += o.w;
                    }
                    o4[j] = a;
                }
            }
            tc::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(bar_aempty + 8 * buf);
        }
        if (nruns == 0)
            for (int j = 0; j < N / 4; ++j) reinterpret_cast<float4 *>(out)[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == kProdWarps) tc::tmem_dealloc(tmem, 2 * kTmemCols);
}

// dW[m, n] = sum_cta partial[cta, m, n] for m < M, n < Nv  (fixed order -> deterministic)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(int nparts, int M, int Nv, int N, const float *__restrict__ partial, float *__restrict__ dw, int lddw) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= M * Nv) return;
    const int m = t / Nv, n = t % Nv;
    double acc = 0.0;
    for (int c = 0; c < nparts; ++c) acc += (double)partial[((size_t)c * kTileM + m) * N + n];
    dw[(size_t)m * lddw + n] = (float)acc;
}

// ------------------------------------------------------------------------------------------------
// weight image: W (N, K) fp32 -> [nkb][hi|lo][N rows x 128 B, 128-byte swizzle], zero padded to nkb*32 columns
__global__ void __launch_bounds__(256)
pack_weight_kernel(int N, int K, int ldw, const float *__restrict__ W, float *__restrict__ img, int lp) {
    const int nkb = (K + kKB - 1) / kKB;
    const int total = nkb * N * kKB;
    for (int t = blockIdx.x * 256 + threadIdx.x; t < total; t += gridDim.x * 256) {
        const int kb = t / (N * kKB), rem = t - kb * N * kKB, n = rem / kKB, c = rem % kKB;
        const int k = kb * kKB + c;
        const float w = k < K ? W[(size_t)n * ldw + k] : 0.f;
        float hi, lo;
        if (lp) hi = bf16_round(w), lo = 0.f;
        else tc::split_tf32(w, hi, lo);
        const size_t base = (size_t)kb * (2 * N * kKB);
        const uint32_t off = tc::sw128_offset(n, c) / 4;
        img[base + off] = hi;
        img[base + (size_t)N * kKB + off] = lo;
    }
}

// BatchNorm batch statistics from the per-thread fp64 partials -> scale/shift (+ saved mean / invstd, running
// statistics update with momentum and the unbiased variance, nn.BatchNorm2d semantics).
// one warp per channel: lanes stride over the partials, fixed-order shuffle tree (deterministic)
__device__ __forceinline__ void warp_sum_pairs(int c, int N, int nparts, const double *__restrict__ partial, double &s, double &q) {
    const int lane = threadIdx.x & 31;
    s = 0.0, q = 0.0;
    for (int i = c + lane * N; i < nparts; i += 32 * N) s += partial[2 * (size_t)i], q += partial[2 * (size_t)i + 1];   // threads e with e % N == c
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o), q += __shfl_xor_sync(0xffffffffu, q, o);
}

__global__ void bn_finalize_kernel(int N, int nparts, long long R, const double *__restrict__ partial,
                                   const float *__restrict__ gamma, const float *__restrict__ beta, float eps,
                                   float momentum, float *running_mean, float *running_var, float *scale, float *shift,
                                   float *save_mean, float *save_invstd) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= N) return;
    double s, q;
    warp_sum_pairs(c, N, nparts, partial, s, q);
    if (threadIdx.x & 31) return;
    const double mean = s / (double)R;
    double var = q / (double)R - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    save_mean[c] = (float)mean;
    save_invstd[c] = invstd;
    if (running_mean) {
        const double unbiased = R > 1 ? var * (double)R / (double)(R - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// sums of the fp64 partial pairs per channel: out[0][c] = sum of first components, out[1][c] = second
__global__ void partial_sum_kernel(int N, int nparts, const double *__restrict__ partial, float *__restrict__ out) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= N) return;
    double s, q;
    warp_sum_pairs(c, N, nparts, partial, s, q);
    if (threadIdx.x & 31) return;
    out[c] = (float)s;
    out[N + c] = (float)q;
}

// Prologue of the pooled-layer backward (replaces six ATen elementwise / reduction launches over (G, N) tensors):
//   dz = d_out * [out > 0];  dsel = dz * s2;  partial sums of  dz  and  dz * (gsel - m2) * i2  per channel.
// blockDim = 256 = (N / 4 float4 columns) x (1024 / N rows per pass); fixed grid -> deterministic sums.
__global__ void __launch_bounds__(256)
pool_bwd_prologue_kernel(long long G, int N, int ldd, const float *__restrict__ d_out, const float *__restrict__ out,
                         const float *__restrict__ gsel, const float *__restrict__ s2, const float *__restrict__ m2,
                         const float *__restrict__ i2, float *__restrict__ dsel, double *__restrict__ partial) {
    __shared__ double s_acc[256][8];
    const int vec = N >> 2, rows_per_pass = 256 / vec;
    const int c4 = threadIdx.x % vec, rl = threadIdx.x / vec;
    const float4 sv = __ldg(reinterpret_cast<const float4 *>(s2) + c4), mv = __ldg(reinterpret_cast<const float4 *>(m2) + c4);
    const float4 iv = __ldg(reinterpret_cast<const float4 *>(i2) + c4);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int since = 0;
    for (long long g = (long long)blockIdx.x * rows_per_pass + rl; g < G; g += (long long)gridDim.x * rows_per_pass) {
        const float4 d = __ldg(reinterpret_cast<const float4 *>(d_out + g * ldd) + c4);
        const float4 o = __ldg(reinterpret_cast<const float4 *>(out + g * N) + c4);
        const float4 y = __ldg(reinterpret_cast<const float4 *>(gsel + g * N) + c4);
        float4 z;
        z.x = o.x > 0.f ? d.x : 0.f, z.y = o.y > 0.f ? d.y : 0.f, z.z = o.z > 0.f ? d.z : 0.f, z.w = o.w > 0.f ? d.w : 0.f;
        reinterpret_cast<float4 *>(dsel + g * N)[c4] = make_float4(z.x * sv.x, z.y * sv.y, z.z * sv.z, z.w * sv.w);
        a[0] += z.x, a[1] += z.y, a[2] += z.z, a[3] += z.w;
        a[4] = fmaf(z.x, (y.x - mv.x) * iv.x, a[4]), a[5] = fmaf(z.y, (y.y - mv.y) * iv.y, a[5]);
        a[6] = fmaf(z.z, (y.z - mv.z) * iv.z, a[6]), a[7] = fmaf(z.w, (y.w - mv.w) * iv.w, a[7]);
        if (++since == 64) {   // fp32 over short runs, fp64 across them
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] += (double)a[u], a[u] = 0.f;
            since = 0;
        }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) s_acc[threadIdx.x][u] = acc[u] + (double)a[u];
    __syncthreads();
    if (threadIdx.x < N) {   // column c = threadIdx.x: fold the row lanes in a fixed order
        const int c = threadIdx.x, q = c >> 2, w = c & 3;
        double t0 = 0.0, t1 = 0.0;
        for (int r = 0; r < rows_per_pass; ++r) t0 += s_acc[r * vec + q][w], t1 += s_acc[r * vec + q][4 + w];
        partial[((size_t)blockIdx.x * N + c) * 2 + 0] = t0;
        partial[((size_t)blockIdx.x * N + c) * 2 + 1] = t1;
    }
}

template <int N, int PM, int EM, int PT>
static int launch_row(const RowGemmArgs &a0, int grid, cudaStream_t stream, int panels = 1) {
    RowGemmArgs a = a0;
#ifdef SG4D_DEBUG   // ablation switches exist only in debug builds (a stray variable must never change production numerics)
    static const bool no_mma = getenv("SG4D_DBG_NOMMA") != nullptr, no_load = getenv("SG4D_DBG_NOLOAD") != nullptr;
    static const bool no_epi = getenv("SG4D_DBG_NOEPI") != nullptr, no_tma = getenv("SG4D_DBG_NOTMA") != nullptr;
    a.dbg_no_mma = no_mma, a.dbg_no_load = no_load, a.dbg_no_epi = no_epi, a.dbg_no_tma = no_tma;
#endif
    a.lp = g_precision;
    auto kern = row_gemm_kernel<N, PM, EM, PT>;
    const int smem = RowSmem<N>::total(PM, EM);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return status_of(e);
    kern<<<dim3(grid, panels), PT + 32 + kEpiThreads, smem, stream>>>(a);
    return SG4D_LAUNCH_CHECK();
}

template <int PM, int EM>
static int launch_row_n(int n, const RowGemmArgs &a, int grid, cudaStream_t stream, int panels = 1) {
    // (measured in round 1: 16 producer warps do not pay for the row-tile kernels -- 80-register cap, epilogue-bound)
    return n == 128 ? launch_row<128, PM, EM, 256>(a, grid, stream, panels) : launch_row<64, PM, EM, 256>(a, grid, stream, panels);
}

template <int N, int PM, int QM>
static int launch_wgrad(const WgradArgs &a0, int grid, cudaStream_t stream, int blocks = 1) {
    WgradArgs a = a0;
#ifdef SG4D_DEBUG
    static const bool no_mma = getenv("SG4D_DBG_NOMMA") != nullptr, no_load = getenv("SG4D_DBG_NOLOAD") != nullptr;
    a.dbg_no_mma = no_mma, a.dbg_no_load = no_load;
#endif
    a.lp = g_precision;
    constexpr int kWide = PM == 6 ? 64 : 128;   // Gram mode exists for 64 channels only
    auto kern = (a.P.ncols <= 64 && blocks == 1) ? (a.lp ? wgrad_kernel<N, PM, QM, 64, true> : wgrad_kernel<N, PM, QM, 64, false>)
                                                  : (a.lp ? wgrad_kernel<N, PM, QM, kWide, true> : wgrad_kernel<N, PM, QM, kWide, false>);
    const int smem = WgSmem<N>::total(PM);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return status_of(e);
    kern<<<dim3(grid, blocks), kMlpThreadsT2, smem, stream>>>(a);
    return SG4D_LAUNCH_CHECK();
}

template <int PM, int QM>
static int launch_wgrad_n(int n, const WgradArgs &a, int grid, cudaStream_t stream) {
    switch (n) {
        case 32: return launch_wgrad<32, PM, QM>(a, grid, stream);
        case 64: return launch_wgrad<64, PM, QM>(a, grid, stream);
        case 128: return launch_wgrad<128, PM, QM>(a, grid, stream);
        case 224: return launch_wgrad<224, PM, QM>(a, grid, stream);
        default: return SG4D_EINVAL;
    }
}

static int mlp_grid(long long rows) {
    const long long ntiles = (rows + kTileM - 1) / kTileM;
    return (int)(ntiles < SG4D_NUM_SMS ? (ntiles < 1 ? 1 : ntiles) : SG4D_NUM_SMS);
}

}  // namespace sg4d

using namespace sg4d;

extern "C" int sg4d_mlp_grid(long long rows) { return mlp_grid(rows); }
extern "C" long long sg4d_mlp_partial_doubles(long long rows) { return (long long)mlp_grid(rows) * kEpiThreads * 2; }

extern "C" long long sg4d_weight_image_floats(int n, int k) {
    return (long long)((k + kKB - 1) / kKB) * 2 * n * kKB;
}

extern "C" int sg4d_pack_weight(int n, int k, int ldw, const float *w, float *img, sg4d_stream_t stream) {
    if (n <= 0 || k <= 0 || ldw < k || (n & 7) || !w || !img) return SG4D_EINVAL;
    const int total = ((k + kKB - 1) / kKB) * n * kKB;
    pack_weight_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, k, ldw, w, img, g_precision);
    return SG4D_LAUNCH_CHECK();
}

static int ilog2(int v) {
    int l = 0;
    while ((1 << (l + 1)) <= v) ++l;
    return l;
}
static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static bool row_common_ok(long long rows, int k, int lda, int n) {
    return rows > 0 && k > 0 && k <= 256 && !(k & 3) && lda >= k && !(lda & 3) && (n == 64 || n == 128);
}

extern "C" int sg4d_linear_fwd(long long rows, int k, int lda, int n, int group, const float *a, const float *scale,
                               const float *shift, const float *wimg, float *y, double *partial, const float *gamma,
                               float *gsel, uint8_t *garg, sg4d_stream_t stream) {
    if (!row_common_ok(rows, k, lda, n) || !a || !wimg || !partial || (scale && !shift)) return SG4D_EINVAL;
    if (group != 0 && (group < 1 || (128 % group) || group > kTileM / (kEpiThreads / n) || rows % group || !gamma || !gsel ||
                       !garg))
        return SG4D_EINVAL;
    RowGemmArgs args{};
    args.op.A = a, args.op.lda = lda, args.op.s = scale, args.op.t = shift, args.op.ncols = k;
    args.R = rows, args.wimg = wimg, args.Y = y, args.ldy = n, args.ycol0 = 0, args.partial = partial;
    args.S = group, args.logS = ilog2(group), args.gamma = gamma, args.gsel = gsel, args.garg = garg;
    const int grid = mlp_grid(rows);
    return scale ? launch_row_n<1, 0>(n, args, grid, (cudaStream_t)stream)
                 : launch_row_n<0, 0>(n, args, grid, (cudaStream_t)stream);
}

extern "C" int sg4d_pool_bwd_da(long long rows, int k, int n, int group, const float *y2, const float *a2, const float *b2,
                                const float *dsel, const uint8_t *garg, const float *wimg_t, const float *y1,
                                const float *es, const float *et, const float *ei, const float *em, float *dz1,
                                double *partial, sg4d_stream_t stream) {
    // k = C2 (<= 1280), n = C1: 64, 128 or a multiple of 128 (column panels; wimg_t then is the dense image, and partial
    // holds sg4d_dense_partial_doubles(rows, n) doubles)
    if (rows <= 0 || k <= 0 || (k & 3) || k > RowSmem<128>::kCF || (n != 64 && (n & 127)) || !pow2(group) || group > 128 ||
        rows % group || !y2 || !a2 || !b2 || !dsel || !garg || !wimg_t || !y1 || !es || !et || !ei || !em || !dz1 || !partial)
        return SG4D_EINVAL;
    const int np = n == 64 ? 64 : 128;
    RowGemmArgs args{};
    args.op.A = y2, args.op.lda = k, args.op.s = a2, args.op.t = b2, args.op.dsel = dsel, args.op.garg = garg;
    args.op.S = group, args.op.logS = ilog2(group), args.op.ldsel = k, args.op.ncols = k;
    args.R = rows, args.wimg = wimg_t, args.Y = dz1, args.ldy = n, args.ycol0 = 0, args.partial = partial;
    args.wimg_panel = sg4d_weight_image_floats(np, k), args.partial_panel = sg4d_mlp_partial_doubles(rows);
    args.E = y1, args.lde = n, args.es = es, args.et = et, args.ei = ei, args.em = em;
    return launch_row_n<2, 1>(np, args, mlp_grid(rows), (cudaStream_t)stream, n / np);
}

extern "C" int sg4d_inner_bwd_dx(long long rows, int k, int n, const float *y1, const float *dz1, const float *p1,
                                 const float *q1, const float *u1, const float *wimg_t, float *dx, int lddx, int col0,
                                 sg4d_stream_t stream) {
    if (!row_common_ok(rows, k, k, n) || !y1 || !dz1 || !p1 || !q1 || !u1 || !wimg_t || !dx || (lddx & 3) || (col0 & 3) ||
        col0 < 0 || col0 + n > lddx)
        return SG4D_EINVAL;
    RowGemmArgs args{};
    args.op.A = y1, args.op.lda = k, args.op.A2 = dz1, args.op.lda2 = k, args.op.s = q1, args.op.t = u1, args.op.p = p1;
    args.op.ncols = k;
    args.R = rows, args.wimg = wimg_t, args.Y = dx, args.ldy = lddx, args.ycol0 = col0;
    return launch_row_n<3, 2>(n, args, mlp_grid(rows), (cudaStream_t)stream);
}

static int wgrad_common(long long rows, int m, int nq, int npad, float *partial, float *dw, int lddw, WgradArgs &args,
                        int pmode, int qmode, cudaStream_t stream) {
    const int grid = mlp_grid(rows);
    args.R = rows, args.partial = partial;
    // LBO = next 32-channel block, SBO = next 4-line swizzle atom, layout type 1 = SWIZZLE_128B_BASE32B, 8 k-lines per MMA
    args.d_lbo = 4096, args.d_sbo = 512, args.d_type = 1, args.d_kstep = 1024;
    int st;
    if (pmode == 2 && qmode == 1) st = launch_wgrad_n<2, 1>(npad, args, grid, stream);
    else if (pmode == 3 && qmode == 0) st = launch_wgrad_n<3, 0>(npad, args, grid, stream);
    else st = SG4D_EINVAL;
    if (st != SG4D_OK) return st;
    wgrad_reduce_kernel<<<(m * nq + 255) / 256, 256, 0, stream>>>(grid, m, nq, npad, partial, dw, lddw);
    return SG4D_LAUNCH_CHECK();
}

extern "C" long long sg4d_wgrad_partial_floats(long long rows, int npad) { return (long long)mlp_grid(rows) * kTileM * npad; }

extern "C" int sg4d_pool_bwd_dw(long long rows, int m, int n, int group, const float *y2, const float *a2, const float *b2,
                                const float *dsel, const uint8_t *garg, const float *y1, const float *s1, const float *t1,
                                float *partial, float *dw, sg4d_stream_t stream) {
    if (rows <= 0 || (m != 64 && m != 128) || (n != 64 && n != 128) || !pow2(group) || group > 128 || rows % group || !y2 || !a2 ||
        !b2 || !dsel || !garg || !y1 || !s1 || !t1 || !partial || !dw)
        return SG4D_EINVAL;
    WgradArgs args{};
    args.P.A = y2, args.P.lda = m, args.P.s = a2, args.P.t = b2, args.P.dsel = dsel, args.P.garg = garg, args.P.S = group, args.P.logS = ilog2(group);
    args.P.ldsel = m, args.P.ncols = m;
    args.Q.A = y1, args.Q.lda = n, args.Q.s = s1, args.Q.t = t1, args.Q.ncols = n;
    return wgrad_common(rows, m, n, n, partial, dw, n, args, 2, 1, (cudaStream_t)stream);
}

extern "C" int sg4d_inner_bwd_dw(long long rows, int m, int k, int ldx, const float *y1, const float *dz1, const float *p1,
                                 const float *q1, const float *u1, const float *x, float *partial, float *dw,
                                 sg4d_stream_t stream) {
    if (rows <= 0 || (m != 64 && m != 128) || k <= 0 || k > 224 || ldx < k || (ldx & 3) || !y1 || !dz1 || !p1 || !q1 || !u1 ||
        !x || !partial || !dw)
        return SG4D_EINVAL;
    const int npad = k <= 32 ? 32 : (k <= 64 ? 64 : (k <= 128 ? 128 : 224));
    WgradArgs args{};
    args.P.A = y1, args.P.lda = m, args.P.A2 = dz1, args.P.lda2 = m, args.P.s = q1, args.P.t = u1, args.P.p = p1;
    args.P.ncols = m;
    args.Q.A = x, args.Q.lda = ldx, args.Q.ncols = (k + 3) & ~3;
    return wgrad_common(rows, m, k, npad, partial, dw, k, args, 3, 0, (cudaStream_t)stream);
}

extern "C" int sg4d_bn_finalize(int n, int nparts, long long rows, const double *partial, const float *gamma,
                                const float *beta, float eps, float momentum, float *running_mean, float *running_var,
                                float *scale, float *shift, float *save_mean, float *save_invstd, sg4d_stream_t stream) {
    if (n <= 0 || nparts <= 0 || rows <= 0 || !partial || !gamma || !beta || !scale || !shift || !save_mean || !save_invstd)
        return SG4D_EINVAL;
    bn_finalize_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(n, nparts, rows, partial, gamma, beta, eps, momentum,
                                                                       running_mean, running_var, scale, shift, save_mean,
                                                                       save_invstd);
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_pool_bwd_prologue_parts(void) { return SG4D_NUM_SMS * 4; }

extern "C" int sg4d_pool_bwd_prologue(long long groups, int n, int ldd, const float *d_out, const float *out,
                                      const float *gsel, const float *s2, const float *m2, const float *i2, float *dsel,
                                      double *partial, sg4d_stream_t stream) {
    if (groups <= 0 || (n != 64 && n != 128 && n != 256) || ldd < n || (ldd & 3) || !d_out || !out || !gsel || !s2 || !m2 || !i2 || !dsel ||
        !partial || (reinterpret_cast<uintptr_t>(d_out) & 15))
        return SG4D_EINVAL;
    pool_bwd_prologue_kernel<<<sg4d_pool_bwd_prologue_parts(), 256, 0, (cudaStream_t)stream>>>(groups, n, ldd, d_out, out, gsel,
                                                                                              s2, m2, i2, dsel, partial);
    return SG4D_LAUNCH_CHECK();
}

// BatchNorm-backward coefficients of one layer from its two per-channel sums (d_beta = sum dz, d_gamma = sum dz xhat):
//   coef[0] = q = scale d_gamma invstd / rows,  coef[1] = u = scale d_beta / rows - q mean   (both 0 without batch statistics),
//   coef[2] = -mean invstd.   dY = p dz - (q y + u) with p = scale.  One launch instead of a dozen per-channel ATen kernels.
__global__ void bn_bwd_coeffs_kernel(int n, long long rows, int batch_stats, const float *__restrict__ d_beta,
                                     const float *__restrict__ d_gamma, const float *__restrict__ scale,
                                     const float *__restrict__ mean, const float *__restrict__ invstd, float *__restrict__ coef) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const float inv_r = 1.0f / (float)rows;
    float q = 0.f, u = 0.f;
    if (batch_stats) {
        q = scale[c] * d_gamma[c] * invstd[c] * inv_r;
        u = scale[c] * d_beta[c] * inv_r - q * mean[c];
    }
    coef[c] = q, coef[n + c] = u, coef[2 * n + c] = -mean[c] * invstd[c];
}

extern "C" int sg4d_bn_bwd_coeffs(int n, long long rows, int batch_stats, const float *d_beta, const float *d_gamma,
                                  const float *scale, const float *mean, const float *invstd, float *coef, sg4d_stream_t stream) {
    if (n <= 0 || rows <= 0 || !d_beta || !d_gamma || !scale || !mean || !invstd || !coef) return SG4D_EINVAL;
    bn_bwd_coeffs_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, rows, batch_stats, d_beta, d_gamma, scale, mean, invstd, coef);
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_partial_sums(int n, int nparts, const double *partial, float *out, sg4d_stream_t stream) {
    if (n <= 0 || nparts <= 0 || !partial || !out) return SG4D_EINVAL;
    partial_sum_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(n, nparts, partial, out);
    return SG4D_LAUNCH_CHECK();
}

// ================================================================================================
// Fused set-abstraction scales: the grouped tensor is never materialised.
//
// SA1 (K = 3 + c <= 7 input channels, 64 first-layer channels, no gradient into the points):
//   forward   sa_moments_kernel  -> M = sum over grouped rows of [x | 1][x | 1]^T (8 x 8, fp64).  The first layer is
//                                   linear in x, so BatchNorm1's batch statistics follow from M alone:
//                                   mean = W1 mu, E[y^2] = w^T (M / R) w  (sa1_bn1_kernel; also emits the
//                                   BatchNorm-scaled weights W1s and shifts t1)
//             row_gemm<N2, 4, 0> -> y2 = relu(W1s x + t1) W2^T with the first layer recomputed in the producers
//                                   (8 FMAs per activation instead of a 256-byte HBM round trip per row),
//                                   BatchNorm2 statistics + group max/min in the epilogue
//   backward  row_gemm<64, 2, 3> -> dz1 = (dY2 W2) * [y1bn > 0] consumed in the epilogue: S1 = dz1^T [x | 1]
//             wgrad<64, 2, 4>    -> dW2 = dY2^T relu(W1s x + t1), the second operand recomputed from the gather
//             sa1_bwd_finalize   -> d_beta1 = S1[:,7], d_gamma1 = i1 (sum_j W1 S1 - m1 d_beta1),
//                                   dW1 = p1 S1 - q1 (W1 M) - u1 (1^T x): every term is linear in S1 and M
// SA2 (K = 195): row_gemm<128, 5, 0> and wgrad<224, 3, 5> gather the rows straight into the swizzled operand tiles.
namespace sg4d {

constexpr int kMomentParts = SG4D_NUM_SMS * 4;

__global__ void __launch_bounds__(256) sa_moments_kernel(GroupSrc g, long long R, double *__restrict__ part) {
    double acc[36];
#pragma unroll
    for (int u = 0; u < 36; ++u) acc[u] = 0.0;
    for (long long row = blockIdx.x * 256LL + threadIdx.x; row < R; row += (long long)gridDim.x * 256) {
        float x[8];
        sa1_gather_row(g, row, true, __ldg(g.idx + row), x);
        int u = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = i; j < 8; ++j, ++u) acc[u] = fma((double)x[i], (double)x[j], acc[u]);   // products are exact in fp64
    }
    __shared__ double s_part[8][36];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int u = 0; u < 36; ++u) {
        double v = acc[u];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp][u] = v;
    }
    __syncthreads();
    if (threadIdx.x < 36) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += s_part[w][threadIdx.x];
        part[(size_t)blockIdx.x * 36 + threadIdx.x] = v;
    }
}

// one block of 64 threads.  moments (8 x 8, full symmetric) is kept for the backward pass.
__global__ void sa1_bn1_kernel(int k, int nparts, const double *__restrict__ part, const float *__restrict__ w1, int ldw,
                               const float *__restrict__ gamma, const float *__restrict__ beta, float eps, float momentum,
                               float *running_mean, float *running_var, int use_running, double *__restrict__ moments,
                               float *__restrict__ stats, float *__restrict__ w1s) {
    __shared__ double M[8][8];
    const int t = threadIdx.x;
    if (t < 36) {
        double v = 0.0;
        for (int q = 0; q < nparts; ++q) v += part[(size_t)q * 36 + t];
        int i = 0, rem = t;
        while (rem >= 8 - i) rem -= 8 - i, ++i;
        const int j = i + rem;
        M[i][j] = v, M[j][i] = v;
    }
    __syncthreads();
    moments[t] = M[t >> 3][t & 7];
    const double R = M[7][7];
    double mean = 0.0, ey2 = 0.0;
    for (int i = 0; i < k; ++i) {
        const double wi = (double)w1[(size_t)t * ldw + i];
        mean += wi * M[i][7];
        for (int j = 0; j < k; ++j) ey2 += wi * (double)w1[(size_t)t * ldw + j] * M[i][j];
    }
    mean /= R, ey2 /= R;
    double var = ey2 - mean * mean;
    if (var < 0.0) var = 0.0;
    float invstd, meanf;
    if (use_running) {
        meanf = running_mean[t];
        invstd = (float)(1.0 / sqrt((double)running_var[t] + (double)eps));
    } else {
        meanf = (float)mean;
        invstd = (float)(1.0 / sqrt(var + (double)eps));
        if (running_mean) {
            const double unbiased = R > 1.0 ? var * R / (R - 1.0) : var;
            running_mean[t] = (1.f - momentum) * running_mean[t] + momentum * meanf;
            running_var[t] = (1.f - momentum) * running_var[t] + momentum * (float)unbiased;
        }
    }
    const float sc = gamma[t] * invstd;
    stats[t] = sc, stats[64 + t] = beta[t] - meanf * sc, stats[128 + t] = meanf, stats[192 + t] = invstd;
    for (int j = 0; j < 8; ++j) w1s[j * 64 + t] = j < k ? __fmul_rn(sc, w1[(size_t)t * ldw + j]) : 0.f;
}

// one block of 512 threads: thread (c = t / 8, j = t % 8)
__global__ void __launch_bounds__(512)
sa1_bwd_finalize_kernel(int k, int nparts, const double *__restrict__ s1part, const double *__restrict__ moments,
                        const float *__restrict__ w1, int ldw, const float *__restrict__ stats, int batch_stats,
                        float *__restrict__ d_w1, int lddw, float *__restrict__ d_g1, float *__restrict__ d_be1) {
    __shared__ double S[64][8];
    const int t = threadIdx.x, c = t >> 3, j = t & 7;
    double v = 0.0;
    for (int q = 0; q < nparts; ++q) v += s1part[(size_t)q * 512 + t];
    S[c][j] = v;
    __syncthreads();
    const double R = moments[63];
    const double s1 = stats[c], m1 = stats[128 + c], i1 = stats[192 + c];
    const double dbe = S[c][7];
    double sdy = 0.0;
    for (int i = 0; i < k; ++i) sdy += (double)w1[(size_t)c * ldw + i] * S[c][i];
    const double dg = i1 * (sdy - m1 * dbe);
    const double q1 = batch_stats ? s1 * dg * i1 / R : 0.0;
    const double u1 = batch_stats ? s1 * dbe / R - q1 * m1 : 0.0;
    if (j < k) {
        double wm = 0.0;
        for (int i = 0; i < k; ++i) wm += (double)w1[(size_t)c * ldw + i] * moments[i * 8 + j];
        d_w1[(size_t)c * lddw + j] = (float)(s1 * S[c][j] - q1 * wm - u1 * moments[j * 8 + 7]);
    }
    if (j == 0) d_g1[c] = (float)dg, d_be1[c] = (float)dbe;
}

static bool fill_src(GroupSrc &g, long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c,
                     const float *pts, const float *feats, const float *centers, const int32_t *idx) {
    if (rows <= 0 || n <= 0 || m <= 0 || !pow2(ns) || ns > 128 || rows % ((long long)m * ns) || pstride < 3 || c < 0 || !pts ||
        !centers || !idx || (c > 0 && (!feats || fstride < foff + c || foff < 0)))
        return false;
    g.pts = pts, g.feats = feats ? feats : pts, g.centers = centers, g.idx = idx;
    g.n = n, g.m = m, g.logns = ilog2(ns), g.pstride = pstride, g.fstride = feats ? fstride : pstride, g.foff = foff, g.c = c;
    g.w1s = nullptr, g.t1 = nullptr;
    return true;
}

}  // namespace sg4d

extern "C" int sg4d_sa_moments_parts(void) { return kMomentParts; }

extern "C" int sg4d_sa_moments(long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c,
                               const float *pts, const float *feats, const float *centers, const int32_t *idx,
                               double *part, sg4d_stream_t stream) {
    GroupSrc g;
    if (!fill_src(g, rows, n, m, ns, pstride, fstride, foff, c, pts, feats, centers, idx) || c > 4 || !part) return SG4D_EINVAL;
    sa_moments_kernel<<<kMomentParts, 256, 0, (cudaStream_t)stream>>>(g, rows, part);
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_sa1_bn1(int k, int nparts, const double *part, const float *w1, int ldw, const float *gamma,
                            const float *beta, float eps, float momentum, float *running_mean, float *running_var,
                            int use_running, double *moments, float *stats, float *w1s, sg4d_stream_t stream) {
    if (k < 3 || k > 7 || nparts <= 0 || !part || !w1 || ldw < k || !gamma || !beta || !moments || !stats || !w1s ||
        (use_running && (!running_mean || !running_var)))
        return SG4D_EINVAL;
    sa1_bn1_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(k, nparts, part, w1, ldw, gamma, beta, eps, momentum, running_mean,
                                                      running_var, use_running, moments, stats, w1s);
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_sa1_fwd(long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c, const float *pts,
                            const float *feats, const float *centers, const int32_t *idx, const float *w1s, const float *t1,
                            int n2, const float *wimg2, float *y2, double *partial, const float *gamma2, float *gsel,
                            uint8_t *garg, sg4d_stream_t stream) {
    RowGemmArgs args{};
    if (!fill_src(args.op.g, rows, n, m, ns, pstride, fstride, foff, c, pts, feats, centers, idx) || c > 4 || !w1s || !t1 ||
        (n2 != 64 && n2 != 128) || !wimg2 || !y2 || !partial || !gamma2 || !gsel || !garg || ns < 8 ||
        ns > kTileM / (kEpiThreads / n2))
        return SG4D_EINVAL;
    args.op.g.w1s = w1s, args.op.g.t1 = t1, args.op.ncols = 64;
    args.R = rows, args.wimg = wimg2, args.Y = y2, args.ldy = n2, args.ycol0 = 0, args.partial = partial;
    args.S = ns, args.logS = ilog2(ns), args.gamma = gamma2, args.gsel = gsel, args.garg = garg;
    return launch_row_n<4, 0>(n2, args, mlp_grid(rows), (cudaStream_t)stream);
}

extern "C" long long sg4d_sa1_s1part_doubles(long long rows) { return (long long)mlp_grid(rows) * 512; }

extern "C" int sg4d_sa1_bwd_da(long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c,
                               const float *pts, const float *feats, const float *centers, const int32_t *idx,
                               const float *w1s, const float *t1, int n2, const float *y2, const float *a2, const float *b2,
                               const float *dsel, const uint8_t *garg, const float *wimg2_t, double *s1part,
                               sg4d_stream_t stream) {
    RowGemmArgs args{};
    if (!fill_src(args.op.g, rows, n, m, ns, pstride, fstride, foff, c, pts, feats, centers, idx) || c > 4 || !w1s || !t1 ||
        (n2 != 64 && n2 != 128) || !y2 || !a2 || !b2 || !dsel || !garg || !wimg2_t || !s1part)
        return SG4D_EINVAL;
    args.op.g.w1s = w1s, args.op.g.t1 = t1;
    args.op.A = y2, args.op.lda = n2, args.op.s = a2, args.op.t = b2, args.op.dsel = dsel, args.op.garg = garg;
    args.op.S = ns, args.op.logS = ilog2(ns), args.op.ldsel = n2, args.op.ncols = n2;
    args.R = rows, args.wimg = wimg2_t, args.s1part = s1part;
    return launch_row<64, 2, 3, 256>(args, mlp_grid(rows), (cudaStream_t)stream);
}

extern "C" int sg4d_sa1_bwd_dw2(long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c,
                                const float *pts, const float *feats, const float *centers, const int32_t *idx,
                                const float *w1s, const float *t1, int n2, const float *y2, const float *a2, const float *b2,
                                const float *dsel, const uint8_t *garg, float *partial, float *dw2, sg4d_stream_t stream) {
    WgradArgs args{};
    if (!fill_src(args.Q.g, rows, n, m, ns, pstride, fstride, foff, c, pts, feats, centers, idx) || c > 4 || !w1s || !t1 ||
        (n2 != 64 && n2 != 128) || !y2 || !a2 || !b2 || !dsel || !garg || !partial || !dw2)
        return SG4D_EINVAL;
    args.Q.g.w1s = w1s, args.Q.g.t1 = t1, args.Q.ncols = 64;
    args.P.A = y2, args.P.lda = n2, args.P.s = a2, args.P.t = b2, args.P.dsel = dsel, args.P.garg = garg, args.P.S = ns;
    args.P.logS = ilog2(ns), args.P.ldsel = n2, args.P.ncols = n2;
    const int grid = mlp_grid(rows);
    args.R = rows, args.partial = partial;
    args.d_lbo = 4096, args.d_sbo = 512, args.d_type = 1, args.d_kstep = 1024;
    const int st = launch_wgrad<64, 2, 4>(args, grid, (cudaStream_t)stream);
    if (st != SG4D_OK) return st;
    wgrad_reduce_kernel<<<(n2 * 64 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(grid, n2, 64, 64, partial, dw2, 64);
    return SG4D_LAUNCH_CHECK();
}

// ---- SA1 second-layer weight gradient without the dY2 operand -----------------------------------------------------------
// dY2[r, c] = [garg(g, c) == slot(r)] dsel(g, c) - (a2[c] y2[r, c] + b2[c])  and  y2 = W2 h1  (h1 = relu(W1s x + t1)), hence
//     dW2 = dY2^T h1 = T1 - diag(a2) W2 M - b2 (x) s,      M = h1^T h1 (64 x 64),  s = column sums of h1,
//     T1[c, :] = sum over groups g of dsel(g, c) h1[row of g that won channel c, :].
// M and s come from ONE 64-wide operand (wgrad_kernel in Gram mode: a third of the operand work of the dY2^T h1 form and no
// read of y2); T1 touches one row per (group, channel) and runs on the CUDA cores (sa1_t1_kernel); sa1_dw2_combine_kernel
// adds the three terms in fp64.
namespace sg4d {

constexpr int kT1Grid = SG4D_NUM_SMS * 2;

template <int N2>
__global__ void __launch_bounds__(256, 2)
sa1_t1_kernel(GroupSrc g, long long R, const float *__restrict__ dsel, const uint8_t *__restrict__ garg, float *__restrict__ partial) {
    constexpr int kHS = 68;                         // row stride of the h1 tile (floats): 16-byte aligned, spreads the banks
    constexpr int kParts = 256 / N2, kW = 64 / kParts;   // thread (c, part) owns T1[c, part * kW .. + kW)
    __shared__ __align__(16) float w_s[8 * 64];
    __shared__ __align__(16) float t_s[64];
    __shared__ __align__(16) float xs[kTileM][8];
    __shared__ __align__(16) float hs[kTileM * kHS];
    const int tid = threadIdx.x;
    for (int k = tid; k < 512; k += 256) w_s[k] = g.w1s[k];
    if (tid < 64) t_s[tid] = g.t1[tid];
    const int ns = 1 << g.logns, gpt = kTileM >> g.logns;     // groups per tile
    const long long ntiles = (R + kTileM - 1) / kTileM, groups = R >> g.logns;
    const int c = tid % N2, kbase = (tid / N2) * kW;
    float acc[kW];
#pragma unroll
    for (int j = 0; j < kW; ++j) acc[j] = 0.f;
    int ixn = 0;                                    // neighbour index of my row of the NEXT tile (gather threads)
    if (tid < kTileM) {
        const long long row = (long long)blockIdx.x * kTileM + tid;
        ixn = row < R ? __ldg(g.idx + row) : 0;
    }
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();                            // the previous tile's hs / xs are no longer read (and w_s / t_s are written)
        if (tid < kTileM) {
            const long long row = tile * kTileM + tid;
            float x[8];
            sa1_gather_row(g, row, row < R, ixn, x);
            const long long rown = row + (long long)gridDim.x * kTileM;
            ixn = rown < R ? __ldg(g.idx + rown) : 0;
            *reinterpret_cast<float4 *>(&xs[tid][0]) = make_float4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<float4 *>(&xs[tid][4]) = make_float4(x[4], x[5], x[6], x[7]);
        }
        // the first 8 groups' (dsel, winner) of my channel: in flight while h1 is computed
        float dv[8];
        int av[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const long long gi = tile * gpt + q;
            const bool ok = q < gpt && gi < groups;
            dv[q] = ok ? __ldg(dsel + gi * N2 + c) : 0.f;
            av[q] = ok ? (int)__ldg(garg + gi * N2 + c) : 0;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kTileM * 16 / 256; ++i) {
            const int item = tid + 256 * i, r = item >> 4, c4 = item & 15;
            const float4 xa = *reinterpret_cast<const float4 *>(&xs[r][0]), xb = *reinterpret_cast<const float4 *>(&xs[r][4]);
            const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            float4 v = sa1_y1bn_smem(x, w_s, 4 * c4, *reinterpret_cast<const float4 *>(t_s + 4 * c4));
            v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
            *reinterpret_cast<float4 *>(hs + r * kHS + 4 * c4) = v;     // (rows >= R are never selected below)
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (q < gpt) {                          // (dv = 0 beyond the last group)
                const float *h = hs + (q * ns + av[q]) * kHS + kbase;
#pragma unroll
                for (int j = 0; j < kW; j += 4) {
                    const float4 hv = *reinterpret_cast<const float4 *>(h + j);
                    acc[j] = fmaf(dv[q], hv.x, acc[j]), acc[j + 1] = fmaf(dv[q], hv.y, acc[j + 1]);
                    acc[j + 2] = fmaf(dv[q], hv.z, acc[j + 2]), acc[j + 3] = fmaf(dv[q], hv.w, acc[j + 3]);
                }
            }
        }
        for (int gl = 8; gl < gpt; ++gl) {          // nsample 8: groups 8..15 of the tile
            const long long gi = tile * gpt + gl;
            if (gi >= groups) break;
            const float d = __ldg(dsel + gi * N2 + c);
            const float *h = hs + (gl * ns + (int)__ldg(garg + gi * N2 + c)) * kHS + kbase;
#pragma unroll
            for (int j = 0; j < kW; j += 4) {
                const float4 hv = *reinterpret_cast<const float4 *>(h + j);
                acc[j] = fmaf(d, hv.x, acc[j]), acc[j + 1] = fmaf(d, hv.y, acc[j + 1]);
                acc[j + 2] = fmaf(d, hv.z, acc[j + 2]), acc[j + 3] = fmaf(d, hv.w, acc[j + 3]);
            }
        }
    }
    float *out = partial + ((size_t)blockIdx.x * N2 + c) * 64 + kbase;
#pragma unroll
    for (int j = 0; j < kW; j += 4) *reinterpret_cast<float4 *>(out + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
}

// dW2[c, k] = sum_parts T1 - a2[c] sum_j W2[c, j] M[j, k] - b2[c] s[k];  mh (65, 64): rows 0..63 = M, row 64 = s
__global__ void __launch_bounds__(256)
sa1_dw2_combine_kernel(int n2, int nparts, const float *__restrict__ t1part, const float *__restrict__ mh,
                       const float *__restrict__ w2, const float *__restrict__ a2, const float *__restrict__ b2,
                       float *__restrict__ dw2) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= n2 * 64) return;
    const int c = t >> 6, k = t & 63;
    double acc = 0.0;
    for (int q = 0; q < nparts; ++q) acc += (double)t1part[((size_t)q * n2 + c) * 64 + k];
    double wm = 0.0;
    for (int j = 0; j < 64; ++j) wm += (double)w2[c * 64 + j] * (double)mh[j * 64 + k];
    dw2[t] = (float)(acc - (double)a2[c] * wm - (double)b2[c] * (double)mh[64 * 64 + k]);
}

}  // namespace sg4d

extern "C" long long sg4d_sa1_bwd_dw2_gram_ws_floats(long long rows, int n2) {
    return (long long)mlp_grid(rows) * kTileM * 64 + 65 * 64 + (long long)kT1Grid * n2 * 64;
}

extern "C" int sg4d_sa1_bwd_dw2_gram(long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c,
                                     const float *pts, const float *feats, const float *centers, const int32_t *idx,
                                     const float *w1s, const float *t1, int n2, const float *w2, const float *a2, const float *b2,
                                     const float *dsel, const uint8_t *garg, float *ws, float *dw2, sg4d_stream_t stream) {
    WgradArgs args{};
    if (!fill_src(args.Q.g, rows, n, m, ns, pstride, fstride, foff, c, pts, feats, centers, idx) || c > 4 || !w1s || !t1 ||
        (n2 != 64 && n2 != 128) || !w2 || !a2 || !b2 || !dsel || !garg || !ws || !dw2 || ns < 8)
        return SG4D_EINVAL;
    args.Q.g.w1s = w1s, args.Q.g.t1 = t1, args.Q.ncols = 64, args.P.ncols = 64;
    const int grid = mlp_grid(rows);
    float *gram_part = ws, *mh = ws + (size_t)grid * kTileM * 64, *t1part = mh + 65 * 64;
    args.R = rows, args.partial = gram_part;
    args.d_lbo = 4096, args.d_sbo = 512, args.d_type = 1, args.d_kstep = 1024;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_wgrad<64, 6, 4>(args, grid, st);
    if (rc != SG4D_OK) return rc;
    wgrad_reduce_kernel<<<(65 * 64 + 255) / 256, 256, 0, st>>>(grid, 65, 64, 64, gram_part, mh, 64);
    if (n2 == 64) sa1_t1_kernel<64><<<kT1Grid, 256, 0, st>>>(args.Q.g, rows, dsel, garg, t1part);
    else sa1_t1_kernel<128><<<kT1Grid, 256, 0, st>>>(args.Q.g, rows, dsel, garg, t1part);
    sa1_dw2_combine_kernel<<<(n2 * 64 + 255) / 256, 256, 0, st>>>(n2, kT1Grid, t1part, mh, w2, a2, b2, dw2);
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_sa1_bwd_finalize(int k, long long rows, const double *s1part, const double *moments, const float *w1,
                                     int ldw, const float *stats, int batch_stats, float *d_w1, int lddw, float *d_g1,
                                     float *d_be1, sg4d_stream_t stream) {
    if (k < 3 || k > 7 || rows <= 0 || !s1part || !moments || !w1 || ldw < k || !stats || !d_w1 || lddw < k || !d_g1 || !d_be1)
        return SG4D_EINVAL;
    sa1_bwd_finalize_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(k, mlp_grid(rows), s1part, moments, w1, ldw, stats, batch_stats,
                                                                 d_w1, lddw, d_g1, d_be1);
    return SG4D_LAUNCH_CHECK();
}

// SA2: first layer with the grouped rows [feats(idx) | xyz(idx) - centre | 0] gathered by the producers
extern "C" int sg4d_linear_fwd_grouped(long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c,
                                       const float *pts, const float *feats, const float *centers, const int32_t *idx,
                                       int nout, const float *wimg, float *y, double *partial, sg4d_stream_t stream) {
    RowGemmArgs args{};
    if (!fill_src(args.op.g, rows, n, m, ns, pstride, fstride, foff, c, pts, feats, centers, idx) || c <= 0 || (c & 3) ||
        (fstride & 3) || (foff & 3) || (reinterpret_cast<uintptr_t>(feats) & 15) || c + 4 > 256 || (nout != 64 && nout != 128) ||
        !wimg || !y || !partial)
        return SG4D_EINVAL;
    args.op.ncols = c + 4;
    args.R = rows, args.wimg = wimg, args.Y = y, args.ldy = nout, args.ycol0 = 0, args.partial = partial;
    return launch_row_n<5, 0>(nout, args, mlp_grid(rows), (cudaStream_t)stream);
}

// SA2: dW1 (mout x (c + 3)) = dY1^T x with x gathered on the fly; columns in the grouped order [feats | xyz]
extern "C" int sg4d_inner_bwd_dw_grouped(long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c,
                                         const float *pts, const float *feats, const float *centers, const int32_t *idx,
                                         int mout, const float *y1, const float *dz1, const float *p1, const float *q1,
                                         const float *u1, float *partial, float *dw, int lddw, sg4d_stream_t stream) {
    WgradArgs args{};
    const int k = c + 3;
    if (!fill_src(args.Q.g, rows, n, m, ns, pstride, fstride, foff, c, pts, feats, centers, idx) || c <= 0 || (c & 3) ||
        (fstride & 3) || (foff & 3) || (reinterpret_cast<uintptr_t>(feats) & 15) || k > 224 || k <= 128 || (mout != 64 && mout != 128) ||
        !y1 || !dz1 || !p1 || !q1 || !u1 || !partial || !dw || lddw < k)
        return SG4D_EINVAL;
    args.P.A = y1, args.P.lda = mout, args.P.A2 = dz1, args.P.lda2 = mout, args.P.s = q1, args.P.t = u1, args.P.p = p1;
    args.P.ncols = mout;
    args.Q.ncols = c + 4;
    const int grid = mlp_grid(rows);
    args.R = rows, args.partial = partial;
    args.d_lbo = 4096, args.d_sbo = 512, args.d_type = 1, args.d_kstep = 1024;
    const int st = launch_wgrad<224, 3, 5>(args, grid, (cudaStream_t)stream);
    if (st != SG4D_OK) return st;
    wgrad_reduce_kernel<<<(mout * k + 255) / 256, 256, 0, (cudaStream_t)stream>>>(grid, mout, k, 224, partial, dw, lddw);
    return SG4D_LAUNCH_CHECK();
}

// ================================================================================================
// Dense layers on the same engine (the GroupAll level SA3, the TripletGCN MLPs, the classifier heads):
//   Y (rows, n) = act(A) W^T [+ bias], n a multiple of 64, computed as column panels of 128 (or 64) by
//   row_gemm_kernel (gridDim.y = panels); dW as (128 x 128) blocks by wgrad_kernel (gridDim.y = blocks).
// Replaces nn.Linear / Conv2d(1x1) + BatchNorm + ReLU of OPS/pointnet2_modules.py:130-146,
// SGH/model/gcns/network_TripletGCN.py:11-58 and SGH/model/pointnets/network_PointNet.py:188-271, which round 1 left
// on cuBLAS / ATen.
namespace sg4d {

static int dense_panel(int n) { return (n % 128 == 0) ? 128 : 64; }

// out = relu(y * s + t), float4
__global__ void __launch_bounds__(256)
bn_relu_apply_kernel(long long rows, int n4, const float *__restrict__ y, int ldy, const float *__restrict__ s,
                     const float *__restrict__ t, float *__restrict__ out, int ldo) {
    const long long total = rows * n4;
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long r = e / n4;
        const int c = (int)(e - r * n4) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(y + r * ldy + c));
        const float4 sv = __ldg(reinterpret_cast<const float4 *>(s + c)), tv = __ldg(reinterpret_cast<const float4 *>(t + c));
        float4 o;
        o.x = fmaxf(fmaf(v.x, sv.x, tv.x), 0.f), o.y = fmaxf(fmaf(v.y, sv.y, tv.y), 0.f);
        o.z = fmaxf(fmaf(v.z, sv.z, tv.z), 0.f), o.w = fmaxf(fmaf(v.w, sv.w, tv.w), 0.f);
        *reinterpret_cast<float4 *>(out + r * ldo + c) = o;
    }
}

// Backward prologue of [BatchNorm -> ReLU]: dz = dh .* [h > 0];  dzs = dz .* s;  per-column fp64 sums of dz and
// dz .* (y - m) .* i.  One column group of 4 per thread, blockDim = (32 column groups, 8 row lanes); grid = (column
// chunks of 128, row slices); every (row slice, column) partial is written once -> fixed-order final sum.
__global__ void __launch_bounds__(256)
bn_relu_bwd_kernel(long long rows, int n, const float *__restrict__ dh, int lddh, const float *__restrict__ h, int ldh,
                   const float *__restrict__ y, int ldy, const float *__restrict__ s, const float *__restrict__ m,
                   const float *__restrict__ iv, float *__restrict__ dzs, int lddz, double *__restrict__ part) {
    const int c = blockIdx.x * 128 + (threadIdx.x & 31) * 4, rl = threadIdx.x >> 5;
    double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (c < n) {
        const float4 sv = __ldg(reinterpret_cast<const float4 *>(s + c)), mv = __ldg(reinterpret_cast<const float4 *>(m + c));
        const float4 ii = __ldg(reinterpret_cast<const float4 *>(iv + c));
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int since = 0;
        for (long long r = (long long)blockIdx.y * 8 + rl; r < rows; r += (long long)gridDim.y * 8) {
            const float4 d = __ldg(reinterpret_cast<const float4 *>(dh + r * lddh + c));
            const float4 o = __ldg(reinterpret_cast<const float4 *>(h + r * ldh + c));
            const float4 yy = __ldg(reinterpret_cast<const float4 *>(y + r * ldy + c));
            float4 z;
            z.x = o.x > 0.f ? d.x : 0.f, z.y = o.y > 0.f ? d.y : 0.f, z.z = o.z > 0.f ? d.z : 0.f, z.w = o.w > 0.f ? d.w : 0.f;
            *reinterpret_cast<float4 *>(dzs + r * lddz + c) = make_float4(z.x * sv.x, z.y * sv.y, z.z * sv.z, z.w * sv.w);
            a[0] += z.x, a[1] += z.y, a[2] += z.z, a[3] += z.w;
            a[4] = fmaf(z.x, (yy.x - mv.x) * ii.x, a[4]), a[5] = fmaf(z.y, (yy.y - mv.y) * ii.y, a[5]);
            a[6] = fmaf(z.z, (yy.z - mv.z) * ii.z, a[6]), a[7] = fmaf(z.w, (yy.w - mv.w) * ii.w, a[7]);
            if (++since == 64) {
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[u] += (double)a[u], a[u] = 0.f;
                since = 0;
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += (double)a[u];
    }
    __shared__ double sh[8][32][8];
#pragma unroll
    for (int u = 0; u < 8; ++u) sh[rl][threadIdx.x & 31][u] = acc[u];
    __syncthreads();
    if (rl == 0 && c < n) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double t = 0.0;
            for (int q = 0; q < 8; ++q) t += sh[q][threadIdx.x & 31][u];
            // part layout: (2, gridDim.y, n): [0] = sum dz, [1] = sum dz * xhat
            part[((size_t)(u >> 2) * gridDim.y + blockIdx.y) * n + c + (u & 3)] = t;
        }
    }
}

// out[j][c] = sum over slices of part[j][slice][c]   (j = 0, 1)
__global__ void colsum_final_kernel(int n, int slices, const double *__restrict__ part, float *__restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n) return;
    const int j = t / n, c = t % n;
    double a = 0.0;
    for (int q = 0; q < slices; ++q) a += part[((size_t)j * slices + q) * n + c];
    out[t] = (float)a;
}

// column sums of a (rows, n) matrix (bias gradients): same two-stage scheme, single quantity
__global__ void __launch_bounds__(256)
colsum_kernel(long long rows, int n, const float *__restrict__ a, int lda, double *__restrict__ part) {
    const int c = blockIdx.x * 128 + (threadIdx.x & 31) * 4, rl = threadIdx.x >> 5;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (c < n)
        for (long long r = (long long)blockIdx.y * 8 + rl; r < rows; r += (long long)gridDim.y * 8) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(a + r * lda + c));
            acc[0] += v.x, acc[1] += v.y, acc[2] += v.z, acc[3] += v.w;
        }
    __shared__ double sh[8][32][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) sh[rl][threadIdx.x & 31][u] = acc[u];
    __syncthreads();
    if (rl == 0 && c < n)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double t = 0.0;
            for (int q = 0; q < 8; ++q) t += sh[q][threadIdx.x & 31][u];
            part[(size_t)blockIdx.y * n + c + u] = t;
            part[((size_t)gridDim.y + blockIdx.y) * n + c + u] = 0.0;
        }
}

// BatchNorm finalize over column panels: channel c lives in panel c / np, thread slot c % np of every epilogue thread group
__global__ void dense_bn_finalize_kernel(int n, int np, int pairs_per_panel, long long R, const double *__restrict__ partial,
                                         const float *__restrict__ gamma, const float *__restrict__ beta, float eps,
                                         float momentum, float *running_mean, float *running_var, float *__restrict__ stats) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= n) return;
    double s, q;
    warp_sum_pairs(c % np, np, pairs_per_panel, partial + (size_t)(c / np) * pairs_per_panel * 2, s, q);
    if (threadIdx.x & 31) return;
    const double mean = s / (double)R;
    double var = q / (double)R - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    stats[c] = sc, stats[n + c] = beta[c] - (float)mean * sc, stats[2 * n + c] = (float)mean, stats[3 * n + c] = invstd;
    if (running_mean) {
        const double unbiased = R > 1 ? var * (double)R / (double)(R - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// dW[m, k] from the per-(block, CTA) partials of a block-grid wgrad launch
__global__ void __launch_bounds__(256)
dense_wgrad_reduce_kernel(int nparts, int mtot, int ktot, int nb_width, int nnb, const float *__restrict__ partial,
                          float *__restrict__ dw, int lddw) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= mtot * ktot) return;
    const int m = t / ktot, k = t % ktot;
    const int blk = (m / kTileM) * nnb + k / nb_width;
    const float *src = partial + ((size_t)blk * nparts * kTileM + (m % kTileM)) * nb_width + (k % nb_width);
    double acc = 0.0;
    for (int c = 0; c < nparts; ++c) acc += (double)src[(size_t)c * kTileM * nb_width];
    dw[(size_t)m * lddw + k] = (float)acc;
}

template <int PM, int QM>
static int launch_dense_wgrad(const WgradArgs &a0, int grid, int blocks, cudaStream_t stream) {
    WgradArgs a = a0;
    a.lp = g_precision;
    auto kern = a.lp ? wgrad_kernel<128, PM, QM, 128, true> : wgrad_kernel<128, PM, QM, 128, false>;
    const int smem = WgSmem<128>::total(PM);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return status_of(e);
    kern<<<dim3(grid, blocks), kMlpThreadsT2, smem, stream>>>(a);
    return SG4D_LAUNCH_CHECK();
}

}  // namespace sg4d

extern "C" long long sg4d_dense_weight_floats(int n, int k) {
    const int np = dense_panel(n);
    return (long long)(n / np) * sg4d_weight_image_floats(np, k);
}

extern "C" int sg4d_dense_pack_weight(int n, int k, int ldw, const float *w, float *img, sg4d_stream_t stream) {
    if (n <= 0 || (n & 63) || k <= 0 || ldw < k || !w || !img) return SG4D_EINVAL;
    const int np = dense_panel(n);
    const long long per = sg4d_weight_image_floats(np, k);
    for (int pnl = 0; pnl < n / np; ++pnl) {
        const int st = sg4d_pack_weight(np, k, ldw, w + (size_t)pnl * np * ldw, img + (size_t)pnl * per, stream);
        if (st != SG4D_OK) return st;
    }
    return SG4D_OK;
}

extern "C" long long sg4d_dense_partial_doubles(long long rows, int n) {
    return (long long)(n / dense_panel(n)) * sg4d_mlp_partial_doubles(rows);
}

static bool dense_ok(long long rows, int k, int lda, int n, const void *a, int ldy) {
    return rows > 0 && k > 0 && !(k & 3) && lda >= k && !(lda & 3) && n > 0 && !(n & 63) && a && ldy >= n && !(ldy & 3) &&
           !(reinterpret_cast<uintptr_t>(a) & 15);
}

extern "C" int sg4d_dense_fwd(long long rows, int k, int lda, int n, const float *a, const float *scale, const float *shift,
                              const float *wimg, const float *bias, float *y, int ldy, double *partial, int group,
                              const float *gamma, float *gsel, uint8_t *garg, sg4d_stream_t stream) {
    if (!dense_ok(rows, k, lda, n, a, ldy) || !wimg || !y || (scale && (!shift || k > RowSmem<128>::kCF)) || (partial && bias))
        return SG4D_EINVAL;
    const int np = dense_panel(n);
    if (group != 0 && (!partial || group < 1 || (128 % group) || group > kTileM / (kEpiThreads / np) || rows % group || !gamma || !gsel || !garg))
        return SG4D_EINVAL;
    RowGemmArgs args{};
    args.op.A = a, args.op.lda = lda, args.op.s = scale, args.op.t = shift, args.op.ncols = k;
    args.R = rows, args.wimg = wimg, args.Y = y, args.ldy = ldy, args.ycol0 = 0, args.partial = partial;
    args.wimg_panel = sg4d_weight_image_floats(np, k), args.partial_panel = sg4d_mlp_partial_doubles(rows);
    args.S = group, args.logS = group ? ilog2(group) : 0, args.gamma = gamma, args.gsel = gsel, args.garg = garg, args.ldg = n;
    args.bias = bias;
    const int grid = mlp_grid(rows), panels = n / np;
    cudaStream_t st = (cudaStream_t)stream;
    if (partial) return scale ? launch_row_n<1, 0>(np, args, grid, st, panels) : launch_row_n<0, 0>(np, args, grid, st, panels);
    return scale ? launch_row_n<1, 2>(np, args, grid, st, panels) : launch_row_n<0, 2>(np, args, grid, st, panels);
}

extern "C" int sg4d_dense_bn_finalize(int n, long long rows, const double *partial, const float *gamma, const float *beta,
                                      float eps, float momentum, float *running_mean, float *running_var, float *stats,
                                      sg4d_stream_t stream) {
    if (n <= 0 || (n & 63) || rows <= 0 || !partial || !gamma || !beta || !stats) return SG4D_EINVAL;
    dense_bn_finalize_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(n, dense_panel(n), mlp_grid(rows) * kEpiThreads, rows,
                                                                             partial, gamma, beta, eps, momentum, running_mean,
                                                                             running_var, stats);
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_bn_relu_apply(long long rows, int n, const float *y, int ldy, const float *scale, const float *shift,
                                  float *out, int ldo, sg4d_stream_t stream) {
    if (rows <= 0 || n <= 0 || (n & 3) || (ldy & 3) || (ldo & 3) || !y || !scale || !shift || !out) return SG4D_EINVAL;
    const long long total = rows * (n / 4);
    long long g = (total + 255) / 256;
    if (g > SG4D_NUM_SMS * 16) g = SG4D_NUM_SMS * 16;
    bn_relu_apply_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(rows, n / 4, y, ldy, scale, shift, out, ldo);
    return SG4D_LAUNCH_CHECK();
}

static int colsum_slices(long long rows) {
    long long s = (rows + 255) / 256;
    return (int)(s < 1 ? 1 : (s > 64 ? 64 : s));
}
extern "C" long long sg4d_colsum_part_doubles(long long rows, int n) { return 2LL * colsum_slices(rows) * n; }

extern "C" int sg4d_bn_relu_bwd(long long rows, int n, const float *dh, int lddh, const float *h, int ldh, const float *y,
                                int ldy, const float *stats, float *dzs, int lddz, double *part, float *sums,
                                sg4d_stream_t stream) {
    if (rows <= 0 || n <= 0 || (n & 3) || (lddh & 3) || (ldh & 3) || (ldy & 3) || (lddz & 3) || !dh || !h || !y || !stats || !dzs ||
        !part || !sums)
        return SG4D_EINVAL;
    const int slices = colsum_slices(rows);
    bn_relu_bwd_kernel<<<dim3((n + 127) / 128, slices), 256, 0, (cudaStream_t)stream>>>(rows, n, dh, lddh, h, ldh, y, ldy, stats,
                                                                                         stats + 2 * n, stats + 3 * n, dzs, lddz, part);
    colsum_final_kernel<<<(2 * n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, slices, part, sums);
    return SG4D_LAUNCH_CHECK();
}

extern "C" int sg4d_colsum(long long rows, int n, const float *a, int lda, double *part, float *out, sg4d_stream_t stream) {
    if (rows <= 0 || n <= 0 || (n & 3) || (lda & 3) || !a || !part || !out) return SG4D_EINVAL;
    const int slices = colsum_slices(rows);
    colsum_kernel<<<dim3((n + 127) / 128, slices), 256, 0, (cudaStream_t)stream>>>(rows, n, a, lda, part);
    colsum_final_kernel<<<(2 * n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, slices, part, out);   // out: (2, n), second row zeros
    return SG4D_LAUNCH_CHECK();
}

// dX (rows, nout) = dY * W  [.* [E*es + et > 0]]  with dY = a (mode 0) or a2 .* p - (a .* q + u) (mode 3); kk = columns
// of dY (<= 1280 for mode 3), wimg_t = dense image of W^T (nout x kk).
extern "C" int sg4d_dense_bwd_dx(long long rows, int kk, int lda, int nout, int mode, const float *a, const float *a2,
                                 const float *p1, const float *q1, const float *u1, const float *wimg_t, const float *e,
                                 int lde, const float *es, const float *et, float *dx, int lddx, double *partial,
                                 sg4d_stream_t stream) {
    if (!dense_ok(rows, kk, lda, nout, a, lddx) || !wimg_t || !dx || (mode != 0 && mode != 3) ||
        (mode == 3 && (!a2 || !p1 || !q1 || !u1 || kk > RowSmem<128>::kCF)) || (e && (!es || !et || !partial || lde < nout)))
        return SG4D_EINVAL;
    const int np = dense_panel(nout);
    RowGemmArgs args{};
    args.op.A = a, args.op.lda = lda, args.op.A2 = a2, args.op.lda2 = lda, args.op.s = q1, args.op.t = u1, args.op.p = p1;
    args.op.ncols = kk;
    args.R = rows, args.wimg = wimg_t, args.Y = dx, args.ldy = lddx, args.ycol0 = 0, args.partial = partial;
    args.wimg_panel = sg4d_weight_image_floats(np, kk), args.partial_panel = sg4d_mlp_partial_doubles(rows);
    args.E = e, args.lde = lde, args.es = es, args.et = et, args.ei = es, args.em = et;
    const int grid = mlp_grid(rows), panels = nout / np;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) return e ? launch_row_n<0, 1>(np, args, grid, st, panels) : launch_row_n<0, 2>(np, args, grid, st, panels);
    return e ? launch_row_n<3, 1>(np, args, grid, st, panels) : launch_row_n<3, 2>(np, args, grid, st, panels);
}

extern "C" long long sg4d_dense_wgrad_partial_floats(long long rows, int m, int k) {
    return (long long)((m + 127) / 128) * ((k + 127) / 128) * mlp_grid(rows) * kTileM * 128;
}

// dW (m x k, row stride lddw) = dY^T * act(x);  dY (rows, m) as in sg4d_dense_bwd_dx;  act(x) = x or relu(x .* xs + xt)
extern "C" int sg4d_dense_bwd_dw(long long rows, int m, int lda, int k, int mode, const float *a, const float *a2,
                                 const float *p1, const float *q1, const float *u1, const float *x, int ldx, const float *xs,
                                 const float *xt, float *partial, float *dw, int lddw, sg4d_stream_t stream) {
    if (rows <= 0 || m <= 0 || (m & 3) || k <= 0 || (k & 3) || lda < m || (lda & 3) || ldx < k || (ldx & 3) || !a || !x || !partial ||
        !dw || lddw < k || (mode != 0 && mode != 3) || (mode == 3 && (!a2 || !p1 || !q1 || !u1)) || (xs && !xt))
        return SG4D_EINVAL;
    WgradArgs args{};
    args.P.A = a, args.P.lda = lda, args.P.A2 = a2, args.P.lda2 = lda, args.P.s = q1, args.P.t = u1, args.P.p = p1, args.P.ncols = m;
    args.Q.A = x, args.Q.lda = ldx, args.Q.s = xs, args.Q.t = xt, args.Q.ncols = k;
    const int grid = mlp_grid(rows), mb = (m + 127) / 128, nnb = (k + 127) / 128;
    args.R = rows, args.partial = partial, args.nnb = nnb, args.mtot = m, args.ktot = k;
    args.d_lbo = 4096, args.d_sbo = 512, args.d_type = 1, args.d_kstep = 1024;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (mb * nnb == 1) {   // single block: the kernel's un-shifted path needs the clamps applied here
        args.P.ncols = m < 128 ? m : 128, args.Q.ncols = k < 128 ? k : 128;
    }
    if (mode == 0) rc = xs ? launch_dense_wgrad<0, 1>(args, grid, mb * nnb, st) : launch_dense_wgrad<0, 0>(args, grid, mb * nnb, st);
    else rc = xs ? launch_dense_wgrad<3, 1>(args, grid, mb * nnb, st) : launch_dense_wgrad<3, 0>(args, grid, mb * nnb, st);
    if (rc != SG4D_OK) return rc;
    dense_wgrad_reduce_kernel<<<(m * k + 255) / 256, 256, 0, st>>>(grid, m, k, 128, nnb, partial, dw, lddw);
    return SG4D_LAUNCH_CHECK();
}

// dW2 (m x n) = dY2^T * relu(y1 .* s1 + t1) for wide layers (m, n multiples of 4; 128 x 128 blocks, one launch);
// dY2 as in sg4d_pool_bwd_da.  partial: sg4d_dense_wgrad_partial_floats(rows, m, n).
extern "C" int sg4d_dense_pool_bwd_dw(long long rows, int m, int n, int group, const float *y2, const float *a2,
                                      const float *b2, const float *dsel, const uint8_t *garg, const float *y1,
                                      const float *s1, const float *t1, float *partial, float *dw, sg4d_stream_t stream) {
    if (rows <= 0 || m <= 0 || (m & 3) || n <= 0 || (n & 3) || !pow2(group) || group > 128 || rows % group || !y2 || !a2 || !b2 ||
        !dsel || !garg || !y1 || !s1 || !t1 || !partial || !dw)
        return SG4D_EINVAL;
    WgradArgs args{};
    args.P.A = y2, args.P.lda = m, args.P.s = a2, args.P.t = b2, args.P.dsel = dsel, args.P.garg = garg, args.P.S = group;
    args.P.logS = ilog2(group), args.P.ldsel = m, args.P.ncols = m < 128 ? m : 128;
    args.Q.A = y1, args.Q.lda = n, args.Q.s = s1, args.Q.t = t1, args.Q.ncols = n < 128 ? n : 128;
    const int grid = mlp_grid(rows), mb = (m + 127) / 128, nnb = (n + 127) / 128;
    args.R = rows, args.partial = partial, args.nnb = nnb, args.mtot = m, args.ktot = n;
    args.d_lbo = 4096, args.d_sbo = 512, args.d_type = 1, args.d_kstep = 1024;
    const int rc = launch_dense_wgrad<2, 1>(args, grid, mb * nnb, (cudaStream_t)stream);
    if (rc != SG4D_OK) return rc;
    dense_wgrad_reduce_kernel<<<(m * n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(grid, m, n, 128, nnb, partial, dw, n);
    return SG4D_LAUNCH_CHECK();
}

#ifdef SG4D_DEBUG
extern "C" int sg4d_debug_set_trace(unsigned long long *buf) {
    return status_of(cudaMemcpyToSymbol(sg4d::g_trace, &buf, sizeof(buf)));
}
#endif

// Process-wide precision of the tensor-core layers: 0 = fp32-level (3xTF32, default), 1 = bf16 operands / fp32 accumulation.
extern "C" int sg4d_set_compute_precision(int mode) {
    if (mode != 0 && mode != 1) return SG4D_EINVAL;
    sg4d::g_precision = mode;
    return SG4D_OK;
}
extern "C" int sg4d_get_compute_precision(void) { return sg4d::g_precision; }
