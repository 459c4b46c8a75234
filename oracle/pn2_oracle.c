/*
 * oracle/pn2_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the five PointNet++ index/gather ops on 4D-OR's scene-graph hot path.
 * The reference has NO CPU path (every host wrapper ends in AT_ASSERT(false, "CPU not supported"),
 * e.g. _ext-src/src/sampling.cpp:83), so this file restates the reference *CUDA kernels* as plain C
 * loops.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it; the product package never does.
 *
 * Citations are relative to
 *   /root/reference/scene_graph_prediction/pointnet2_dir/pointnet2_ops_lib/pointnet2_ops/_ext-src/
 *
 * Parity pin: this restatement is checked bit-for-bit against the reference's own kernels
 * (the src/ *_gpu.cu files compiled unmodified for sm_100a into oracle/_ref/, run on the B200 box) by
 * tests/test_gpu_ref_ext.py, and against fixtures in tests/golden/ produced by running the
 * reference's own Python (pointnet2_utils.py / pointnet2_modules.py / model code) on top of it.
 *
 * Rounding: nvcc (default --fmad=true) contracts the reference's distance expressions as
 *     t = dy*dy;  t = fma(dx,dx,t);  d = fma(dz,dz,t)
 * (read from the SASS of the reference sources built with nvcc 12.9 for sm_100a; see
 * oracle/README.md).  This file is compiled with -ffp-contract=off and spells the same order with
 * explicit fmaf so that gcc cannot pick another one.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* squared distance in the contraction order the reference's SASS uses */
static inline float sqdist3(float dx, float dy, float dz) {
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

/* include/cuda_utils.h:15-19 -- opt_n_threads: largest power of two <= work_size, capped at 512.
 * The reference evaluates it as int(log(double)/log(2.0)); we keep that expression. */
int oracle_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 512) t = 512;
    if (t < 1) t = 1;
    return t;
}

/*
 * Furthest point sampling.  Kernel src/sampling_gpu.cu:69-173, launch :175-229 (one CTA of
 * T = opt_n_threads(n) threads per cloud), host src/sampling.cpp:66-87 (temp filled with 1e10).
 *
 * Literal simulation of the kernel: T virtual threads each scan k = tid, tid+T, ... keeping a
 * strict-'>' running best (:95-110), then the shared-memory tree of :115-168 where __update
 * (:59-65) keeps the LOWER slot on ties.  dataset (b,n,3) fp32, temp (b,n) fp32 scratch
 * (overwritten), idxs (b,m) int32.
 */
void oracle_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp,
                                    int32_t *idxs) {
    if (m <= 0) return;
    const int T = oracle_opt_n_threads(n);
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < b; ++bi) {
        const float *pts = dataset + (size_t)bi * n * 3;
        float *tmp = temp + (size_t)bi * n;
        int32_t *out = idxs + (size_t)bi * m;
        float *dists = (float *)malloc(sizeof(float) * (size_t)T);
        int *dists_i = (int *)malloc(sizeof(int) * (size_t)T);
        for (int k = 0; k < n; ++k) tmp[k] = 1e10f; /* sampling.cpp:74-76 */
        int old = 0;
        out[0] = old;
        for (int j = 1; j < m; ++j) {
            const float x1 = pts[old * 3 + 0], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
            for (int tid = 0; tid < T; ++tid) {
                int besti = 0;
                float best = -1.0f;
                for (int k = tid; k < n; k += T) {
                    const float x2 = pts[k * 3 + 0], y2 = pts[k * 3 + 1], z2 = pts[k * 3 + 2];
                    const float mag = sqdist3(x2, y2, z2);
                    if ((double)mag <= 1e-3) continue; /* :100-101, fp64 compare */
                    const float d = sqdist3(x2 - x1, y2 - y1, z2 - z1);
                    const float d2 = fminf(d, tmp[k]);
                    tmp[k] = d2;
                    besti = d2 > best ? k : besti;
                    best = d2 > best ? d2 : best;
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int s = T >> 1; s >= 1; s >>= 1) { /* the block_size >= 2s stages */
                for (int tid = 0; tid < s; ++tid) {
                    const float v1 = dists[tid], v2 = dists[tid + s];
                    const int i1 = dists_i[tid], i2 = dists_i[tid + s];
                    dists[tid] = fmaxf(v1, v2);
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            out[j] = old;
        }
        free(dists);
        free(dists_i);
    }
}

/* src/sampling_gpu.cu:8-20 -- out[b,c,j] = points[b,c,idx[b,j]] */
void oracle_gather_points(int b, int c, int n, int m, const float *points, const int32_t *idx,
                          float *out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < b; ++i)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < m; ++j) {
                const int a = idx[(size_t)i * m + j];
                out[((size_t)i * c + l) * m + j] = points[((size_t)i * c + l) * n + a];
            }
}

/* src/sampling_gpu.cu:34-47 -- grad_points[b,c,idx[b,j]] += grad_out[b,c,j]  (zeroed target;
 * the reference uses atomicAdd so its summation order is arbitrary: here ascending j) */
void oracle_gather_points_grad(int b, int c, int n, int m, const float *grad_out,
                               const int32_t *idx, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < b; ++i)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < m; ++j) {
                const int a = idx[(size_t)i * m + j];
                grad_points[((size_t)i * c + l) * n + a] += grad_out[((size_t)i * c + l) * m + j];
            }
}

/*
 * Ball query.  Kernel src/ball_query_gpu.cu:9-44; host src/ball_query.cpp:8-32 (idx = zeros).
 * new_xyz (b,m,3), xyz (b,n,3) fp32 -> idx (b,m,nsample) int32.  radius arrives as a C float.
 */
void oracle_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                       const float *xyz, int32_t *idx) {
    const float radius2 = radius * radius;
    memset(idx, 0, sizeof(int32_t) * (size_t)b * m * nsample);
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < b; ++bi) {
        const float *p = xyz + (size_t)bi * n * 3;
        const float *q = new_xyz + (size_t)bi * m * 3;
        int32_t *o = idx + (size_t)bi * m * nsample;
        for (int j = 0; j < m; ++j) {
            const float nx = q[j * 3 + 0], ny = q[j * 3 + 1], nz = q[j * 3 + 2];
            for (int k = 0, cnt = 0; k < n && cnt < nsample; ++k) {
                const float d2 = sqdist3(nx - p[k * 3 + 0], ny - p[k * 3 + 1], nz - p[k * 3 + 2]);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[(size_t)j * nsample + l] = k;
                    o[(size_t)j * nsample + cnt] = k;
                    ++cnt;
                }
            }
        }
    }
}

/* src/group_points_gpu.cu:8-28 -- out[b,c,j,k] = points[b,c,idx[b,j,k]] */
void oracle_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                         const int32_t *idx, float *out) {
#pragma omp parallel for schedule(static)
    for (int bi = 0; bi < b; ++bi) {
        const float *p = points + (size_t)bi * n * c;
        const int32_t *ix = idx + (size_t)bi * npoints * nsample;
        float *o = out + (size_t)bi * npoints * nsample * c;
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < npoints; ++j)
                for (int k = 0; k < nsample; ++k)
                    o[((size_t)l * npoints + j) * nsample + k] =
                        p[(size_t)l * n + ix[(size_t)j * nsample + k]];
    }
}

/* src/group_points_gpu.cu:43-64 -- grad_points[b,c,idx[b,j,k]] += grad_out[b,c,j,k]
 * (atomicAdd in the reference; here the fixed order j ascending, k ascending) */
void oracle_group_points_grad(int b, int c, int n, int npoints, int nsample,
                              const float *grad_out, const int32_t *idx, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
#pragma omp parallel for schedule(static)
    for (int bi = 0; bi < b; ++bi) {
        const float *g = grad_out + (size_t)bi * npoints * nsample * c;
        const int32_t *ix = idx + (size_t)bi * npoints * nsample;
        float *gp = grad_points + (size_t)bi * n * c;
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < npoints; ++j)
                for (int k = 0; k < nsample; ++k)
                    gp[(size_t)l * n + ix[(size_t)j * nsample + k]] +=
                        g[((size_t)l * npoints + j) * nsample + k];
    }
}

/* src/interpolate_gpu.cu:9-59 -- three nearest known points of every unknown point.  Literal restatement: k ascending,
 * strict '<' against doubles initialised to 1e40 (so +inf and NaN distances never enter), distances in the SASS order
 * of the reference build (t = dy*dy; t = fma(dx,dx,t); d = fma(dz,dz,t)); outputs stored as float (1e40 -> +inf). */
void oracle_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int32_t *idx) {
#pragma omp parallel for schedule(static)
    for (int bi = 0; bi < b; ++bi) {
        const float *u = unknown + (size_t)bi * n * 3, *kn = known + (size_t)bi * m * 3;
        for (int j = 0; j < n; ++j) {
            const float ux = u[3 * j], uy = u[3 * j + 1], uz = u[3 * j + 2];
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int besti1 = 0, besti2 = 0, besti3 = 0;
            for (int k = 0; k < m; ++k) {
                const float d = sqdist3(ux - kn[3 * k], uy - kn[3 * k + 1], uz - kn[3 * k + 2]);
                if (d < best1) {
                    best3 = best2, besti3 = besti2, best2 = best1, besti2 = besti1, best1 = d, besti1 = k;
                } else if (d < best2) {
                    best3 = best2, besti3 = besti2, best2 = d, besti2 = k;
                } else if (d < best3) {
                    best3 = d, besti3 = k;
                }
            }
            float *o = dist2 + ((size_t)bi * n + j) * 3;
            int32_t *oi = idx + ((size_t)bi * n + j) * 3;
            o[0] = (float)best1, o[1] = (float)best2, o[2] = (float)best3;
            oi[0] = besti1, oi[1] = besti2, oi[2] = besti3;
        }
    }
}

/* src/interpolate_gpu.cu:72-101 -- out[b,l,j] = p[i1]*w1 + p[i2]*w2 + p[i3]*w3, contracted by nvcc as
 * t = p[i2]*w2; t = fma(p[i1],w1,t); out = fma(p[i3],w3,t)  (read from the SASS of the reference build) */
void oracle_three_interpolate(int b, int c, int m, int n, const float *points, const int32_t *idx,
                              const float *weight, float *out) {
#pragma omp parallel for schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l) {
            const float *p = points + ((size_t)bi * c + l) * m;
            for (int j = 0; j < n; ++j) {
                const int32_t *ix = idx + ((size_t)bi * n + j) * 3;
                const float *w = weight + ((size_t)bi * n + j) * 3;
                float t = p[ix[1]] * w[1];
                t = fmaf(p[ix[0]], w[0], t);
                out[((size_t)bi * c + l) * n + j] = fmaf(p[ix[2]], w[2], t);
            }
        }
}

/* src/interpolate_gpu.cu:116-143 -- grad_points[b,l,i_t] += grad_out[b,l,j] * w_t (atomicAdd in the reference; here
 * the fixed order j ascending, t = 1, 2, 3) */
void oracle_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int32_t *idx,
                                   const float *weight, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * c * m);
#pragma omp parallel for schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l) {
            float *gp = grad_points + ((size_t)bi * c + l) * m;
            for (int j = 0; j < n; ++j) {
                const int32_t *ix = idx + ((size_t)bi * n + j) * 3;
                const float *w = weight + ((size_t)bi * n + j) * 3;
                const float g = grad_out[((size_t)bi * c + l) * n + j];
                gp[ix[0]] += g * w[0];
                gp[ix[1]] += g * w[1];
                gp[ix[2]] += g * w[2];
            }
        }
}
