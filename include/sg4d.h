/*
 * sg4d.h -- C ABI of libsg4d.so, the B200 (sm_100a) implementation of 4D-OR's scene-graph
 * prediction hot path.
 *
 * This is the drop-in boundary.  Every entry point takes raw DEVICE pointers, plain ints/floats and
 * an explicit cudaStream_t (passed as void*), launches asynchronously on that stream, never
 * synchronises, keeps no host state and returns an int status (0 = ok, otherwise a cudaError_t
 * value or one of the SG4D_E* codes below; sg4d_error_string() explains it).  The reference's
 * launchers instead read the stream from ATen and call exit(-1) on a launch failure
 * (_ext-src/include/cuda_utils.h:30-39).
 *
 * Reference paths are relative to
 *   scene_graph_prediction/pointnet2_dir/pointnet2_ops_lib/pointnet2_ops/_ext-src/   ("EXT/")
 *
 * Section 1 mirrors, one to one, the launcher prototypes the reference's host wrappers bind
 * (EXT/src/sampling.cpp:4-13, ball_query.cpp:4-6, group_points.cpp:4-10): same argument order and
 * meaning, same tensor layouts, same zero/1e10 initialisation contract -- plus stream and status.
 * Section 2 holds the fused entry points the host-side model uses (no reference counterpart as a
 * single call; each names the reference call sequence it replaces).
 *
 * All offsets are 64-bit inside the kernels (the reference's 32-bit ints overflow above 1365 clouds
 * per call, group_points_gpu.cu:14-16).
 */
#ifndef SG4D_H_
#define SG4D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG4D_ABI_VERSION 1

/* status codes beyond cudaError_t (which occupies small positive ints) */
#define SG4D_OK 0
#define SG4D_EINVAL 10001  /* bad shape / null pointer / unsupported size          */
#define SG4D_ENODEV 10002  /* device is not compute capability 10.x                */

typedef void *sg4d_stream_t; /* a cudaStream_t */

int sg4d_abi_version(void);
const char *sg4d_error_string(int status);
/* 0 when the current device can run this library (sm_100), SG4D_ENODEV otherwise */
int sg4d_check_device(void);

/* ------------------------------------------------------------------------------------------------
 * Section 1 -- one-to-one replacements of the reference launchers
 * ---------------------------------------------------------------------------------------------- */

/* replaces furthest_point_sampling_kernel_wrapper (EXT/src/sampling.cpp:11-13,
 * sampling_gpu.cu:175-229).  dataset (b,n,3) fp32; temp (b,n) fp32 scratch -- accepted for ABI
 * compatibility; the kernel keeps the running minimum distances on chip and only uses temp when a
 * cloud is too large for that (it then expects it pre-filled with 1e10 like the reference,
 * sampling.cpp:74-76); may be NULL otherwise.  idxs (b,m) int32.  Bit-identical selection
 * order to the reference kernel launched with opt_n_threads(n) threads (tie-break included). */
int sg4d_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp,
                                 int32_t *idxs, sg4d_stream_t stream);

/* replaces gather_points_kernel_wrapper (sampling.cpp:4-6, sampling_gpu.cu:8-30).
 * points (b,c,n) fp32, idx (b,npoints) int32 -> out (b,c,npoints) fp32 */
int sg4d_gather_points(int b, int c, int n, int npoints, const float *points, const int32_t *idx,
                       float *out, sg4d_stream_t stream);

/* replaces gather_points_grad_kernel_wrapper (sampling.cpp:7-9, sampling_gpu.cu:34-57).
 * grad_out (b,c,npoints), idx (b,npoints) -> grad_points (b,c,n), which must be ZERO on entry
 * (the reference host wrapper allocates it with torch::zeros, sampling.cpp:50-52). */
int sg4d_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out,
                            const int32_t *idx, float *grad_points, sg4d_stream_t stream);

/* replaces query_ball_point_kernel_wrapper (ball_query.cpp:4-6, ball_query_gpu.cu:46-54).
 * new_xyz (b,m,3), xyz (b,n,3) fp32 -> idx (b,m,nsample) int32.  Every slot is written (rows
 * without any hit are set to 0, which is what the reference's zero-initialised output holds). */
int sg4d_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                    const float *xyz, int32_t *idx, sg4d_stream_t stream);

/* replaces group_points_kernel_wrapper (group_points.cpp:4-6, group_points_gpu.cu:30-39).
 * points (b,c,n) fp32, idx (b,npoints,nsample) int32 -> out (b,c,npoints,nsample) fp32 */
int sg4d_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                      const int32_t *idx, float *out, sg4d_stream_t stream);

/* replaces group_points_grad_kernel_wrapper (group_points.cpp:8-10, group_points_gpu.cu:66-75).
 * grad_out (b,c,npoints,nsample), idx (b,npoints,nsample) -> grad_points (b,c,n), ZERO on entry
 * (group_points.cpp:49-51). */
int sg4d_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                           const int32_t *idx, float *grad_points, sg4d_stream_t stream);

/* replaces three_nn_kernel_wrapper (interpolate.cpp:4-5, interpolate_gpu.cu:9-70).
 * unknown (b,n,3), known (b,m,3) fp32 -> dist2 (b,n,3) fp32 = SQUARED distances to the three nearest known points
 * (ascending; ties keep the lower index), idx (b,n,3) int32.  With m < 3 the missing slots hold index 0 and +inf
 * (the reference stores (float)1e40). */
int sg4d_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int32_t *idx,
                  sg4d_stream_t stream);

/* replaces three_interpolate_kernel_wrapper (interpolate.cpp:6-8, interpolate_gpu.cu:72-114).
 * points (b,c,m), idx (b,n,3), weight (b,n,3) -> out (b,c,n) */
int sg4d_three_interpolate(int b, int c, int m, int n, const float *points, const int32_t *idx, const float *weight,
                           float *out, sg4d_stream_t stream);

/* replaces three_interpolate_grad_kernel_wrapper (interpolate.cpp:9-12, interpolate_gpu.cu:116-154).
 * grad_out (b,c,n), idx, weight (b,n,3) -> grad_points (b,c,m), ZERO on entry (interpolate.cpp:85-87). */
int sg4d_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int32_t *idx,
                                const float *weight, float *grad_points, sg4d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Section 2 -- fused / point-major entry points used by the host-side model
 *
 * "Point-major" = one row per point, channels contiguous: a cloud is (n, row_stride) fp32 with
 * xyz in columns 0..2 of each row and the feature channels after them.  This is the layout the
 * reference's collate produces before its permute (or_dataset.py:66) and what
 * PointNetfeat.forward recovers with transpose(1,2) (network_PointNet2.py:23), so no
 * _break_up_pc / transpose().contiguous() copies are needed (pointnet2_ssg_cls.py:98-102,
 * pointnet2_modules.py:50-56).
 * ---------------------------------------------------------------------------------------------- */

/* FPS on strided rows + gather of the picked rows' xyz in one launch.
 * Replaces furthest_point_sample + gather_operation + 2 transposes (pointnet2_modules.py:50-59).
 * pts (b,n,row_stride) fp32 with xyz in columns 0..2; idxs (b,m) int32; new_xyz (b,m,3) fp32
 * (may be NULL).  temp as in sg4d_furthest_point_sampling. */
int sg4d_fps_rows(int b, int n, int m, int row_stride, const float *pts, float *temp,
                  int32_t *idxs, float *new_xyz, sg4d_stream_t stream);

/* Ball query for up to SG4D_MAX_SCALES radii over the same centres in ONE scan of the cloud, on
 * strided rows.  Replaces one ball_query launch per scale (pointnet2_utils.py:318 via
 * pointnet2_modules.py:61-64).  For scale s: idx[s] (b,m,nsample[s]) int32 and cnt[s] (b,m) int32 =
 * number of distinct hits found (<= nsample[s]; slots cnt..nsample-1 repeat slot 0, or the whole row
 * is 0 when cnt == 0).  cnt[s] may be NULL. */
#define SG4D_MAX_SCALES 4
int sg4d_ball_query_rows(int b, int n, int m, int row_stride, int center_stride, int nscales,
                         const float *radius, const int *nsample, const float *centers,
                         const float *pts, int32_t *const *idx, int32_t *const *cnt,
                         sg4d_stream_t stream);

/* Grouping into point-major rows, recentring and concatenation fused:
 *   out[b,j,k, 0:3]  = xyz(pts[b, idx[b,j,k]]) - centers[b,j]            (pointnet2_utils.py:319-321)
 *   out[b,j,k, 3:3+c] = feats[b, idx[b,j,k], 0:c]                         (pointnet2_utils.py:324-328)
 * pts (b,n,pts_stride) supplies xyz (columns 0..2); feats (b,n,feat_stride) supplies c channels
 * starting at column feat_offset (feats may alias pts: SA1 reads rgb/mask from the raw rows).
 * out (b,m,nsample,out_stride) with out_stride >= 3+c; columns 3+c..out_stride-1 are zeroed.
 * xyz_col0 = 0 gives the reference's channel order (xyz first); xyz_col0 = c puts the features in columns
 * 0..c-1 and xyz in c..c+2 (16-byte aligned feature columns for the tensor-core backward; the host side
 * permutes the first conv's weight columns accordingly).
 * Replaces 2 x group_points + the in-place subtract + torch.cat. */
int sg4d_group_rows(int b, int n, int m, int nsample, int c, int pts_stride, int feat_stride,
                    int feat_offset, int out_stride, int xyz_col0, const float *pts, const float *feats,
                    const float *centers, const int32_t *idx, float *out, sg4d_stream_t stream);

/* Backward of the feature part of sg4d_group_rows (replaces group_points_grad,
 * pointnet2_utils.py:236-241): grad_feats[b,i,0:c] = sum over (j,k) with idx[b,j,k]==i of
 * grad_out[b,j,k,3:3+c], accumulated in the FIXED order j ascending then k ascending
 * (deterministic; the reference's atomicAdd order is arbitrary).  Requires every idx row to be the
 * output of a ball query (ascending distinct hits followed by repeats of the first hit) and cnt
 * (b,m) from sg4d_ball_query_rows; the feature gradients are columns gcol0..gcol0+c-1 of grad_out.
 * grad_feats (b,n,c) is fully overwritten when accumulate == 0
 * and added to (the second scale of an MSG level) when accumulate != 0. */
/* Same gather over rows that are never stored: g_out (b, n, c1) = per source point, the sum of the first layer's
 * dY1 = p1 .* dz1 - (q1 .* y1 + u1) over the grouped rows that reference it (y1, dz1: (b*m*nsample, c1), c1 in {64, 128}).
 * dX = dY1 W1 is linear, so the feature gradient of group_points_grad is then ONE small GEMM g_out * W1[:, feats] over b*n
 * rows (sg4d_dense_bwd_dx) instead of the dX GEMM over all grouped rows + its (rows x K) tensor. */
int sg4d_group_rows_grad_dy(int b, int n, int m, int nsample, int c1, const float *y1, const float *dz1, const float *p1,
                            const float *q1, const float *u1, const int32_t *idx, const int32_t *cnt, float *g_out,
                            sg4d_stream_t stream);
/* SA2 through the first layer's linearity: y1[r] = W1 [feats(i) | xyz(i) - centre_j] = z[cloud*n + idx[r]] - cc[r / nsample], with
 * z = [feats | xyz] W1^T per SOURCE POINT (b*n rows, one small GEMM: sg4d_dense_fwd) and cc = centre W1x^T per centre.  Writes
 * y1 (rows, c1) and the per-channel (sum, sum of squares) fp64 partial pairs (sg4d_gather_y1_parts(c1) of them) for
 * sg4d_bn_finalize.  Replaces the first layer's GEMM over all b*m*nsample grouped rows. */
int sg4d_gather_y1_parts(int c1);
int sg4d_gather_y1(long long rows, int n, int m, int nsample, int c1, const float *z, const float *cc, const int32_t *idx,
                   float *y1, double *partial, sg4d_stream_t stream);
/* h (groups, c1) = per centre, the sum of dY1 = p1 .* dz1 - (q1 .* y1 + u1) over its nsample rows.  With g_out of
 * sg4d_group_rows_grad_dy:  dW1 = g_out^T [feats | xyz] - h^T [0 | centre]  (two small GEMMs, sg4d_dense_bwd_dw). */
int sg4d_group_sum_dy(long long groups, int nsample, int c1, const float *y1, const float *dz1, const float *p1,
                      const float *q1, const float *u1, float *h, sg4d_stream_t stream);
int sg4d_group_rows_grad(int b, int n, int m, int nsample, int c, int out_stride, int gcol0, int accumulate,
                         const float *grad_out, const int32_t *idx, const int32_t *cnt,
                         float *grad_feats, sg4d_stream_t stream);

/* ---- spatial index: exact acceleration of FPS and ball query on large clouds (csrc/spatial.cu) ----
 * index = caller-provided device workspace of sg4d_spatial_index_bytes(b, n) bytes.  Build sorts every cloud into
 * Morton-cell order (float4 {x, y, z, original index} + the FPS running minimum per point) and forms buckets of 64
 * consecutive points with exact bounding boxes.  Supported for 1024 <= n <= 393216 (sg4d_spatial_index_supported). */
long long sg4d_spatial_index_bytes(int b, int n);
int sg4d_spatial_index_supported(int n);
int sg4d_spatial_index_build(int b, int n, int row_stride, const float *pts, void *index, sg4d_stream_t stream);
/* Same result as sg4d_fps_rows, bit for bit (same arithmetic, same tie-break); buckets whose bounding-box distance
 * to the new pick is >= their largest running minimum are skipped.  Consumes the index's running minima: build the
 * index again before another FPS call (the ball query below does not depend on them). */
int sg4d_fps_indexed(int b, int n, int m, int row_stride, const float *pts, void *index, int32_t *idxs,
                     float *new_xyz, sg4d_stream_t stream);
/* Same result as sg4d_ball_query_rows, bit for bit.  prefix > 0: the first `prefix` points of every cloud are
 * scanned by brute force with early exit (cheap for centres in dense regions) and only the centres still short of
 * nsample hits are answered from the index; prefix == 0: index only.  cnt must be non-NULL when prefix > 0. */
int sg4d_ball_query_rows_indexed(int b, int n, int m, int row_stride, int center_stride, int nscales,
                                 const float *radius, const int *nsample, const float *centers, const float *pts,
                                 const void *index, int prefix, int32_t *const *idx, int32_t *const *cnt,
                                 sg4d_stream_t stream);

/* TripletGCN message input (network_TripletGCN.py:45-46 + PyG __collect__):
 *   out[e, 0:d]      = x[dst[e]]      (x_i, edge_index[1])
 *   out[e, d:d+de]   = edge_feat[e]
 *   out[e, d+de:2d+de] = x[src[e]]    (x_j, edge_index[0])
 * x (n_nodes,d), edge_feat (n_edges,de) fp32; src/dst int64 (the reference's edge_index rows). */
int sg4d_triplet_gather(int64_t n_edges, int d, int de, const float *x, const float *edge_feat,
                        const int64_t *src, const int64_t *dst, float *out, sg4d_stream_t stream);

/* Deterministic segmented scatter-add (replaces torch_scatter.scatter(reduce='add'),
 * network_TripletGCN.py:57 and the index_select backward):
 *   out[v, 0:d] = sum over edges e in order[seg_ptr[v] .. seg_ptr[v+1]) of src[e*src_stride + col0 + 0:d]
 * and, when src2 != NULL, the same edges' src2[e*src_stride + col1 + 0:d] is added too (the
 * reference's new_x_i + new_x_j split of the message, network_TripletGCN.py:48-51).
 * order (n_edges) int32 = edge ids sorted by destination (stable), seg_ptr (n_nodes+1) int32. */
int sg4d_segment_sum(int n_nodes, int d, int64_t src_stride, int col0, int col1, const float *src,
                     int has_second, const int32_t *order, const int32_t *seg_ptr, float *out,
                     sg4d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Section 3 -- the shared point-MLP on the tensor cores (tcgen05, 3xTF32, fp32 accumulate in TMEM)
 *
 * One layer of build_shared_mlp (OPS/pointnet2_modules.py:9-19) = Conv2d(1x1, no bias) -> BatchNorm2d ->
 * ReLU, applied to the (rows, channels) point-major matrix; the last layer of a scale is followed by
 * max_pool2d over nsample (modules.py:67-70).  sg4d_linear_fwd computes Y = act(A) * W^T for one layer,
 * where act() is the PREVIOUS layer's BatchNorm scale/shift + ReLU applied while the operand is staged
 * (so normalised activations never touch HBM), and emits what this layer's BatchNorm (and the max-pool)
 * need: per-channel sum / sum of squares and, optionally, the per-group extreme pre-activation.
 * ---------------------------------------------------------------------------------------------- */

/* number of CTAs sg4d_linear_fwd launches for `rows` rows */
int sg4d_mlp_grid(long long rows);
/* size (in doubles) of the `partial` statistics buffer of sg4d_linear_fwd / sg4d_pool_bwd_da for `rows` rows */
long long sg4d_mlp_partial_doubles(long long rows);
/* size (in floats) of the packed image of an (n, k) weight */
long long sg4d_weight_image_floats(int n, int k);
/* w (n, k) fp32, row stride ldw  ->  img: per 32-column k-block, the TF32 hi and lo parts as 128-byte
 * swizzled K-major tiles (the exact shared-memory image the MMA reads; loaded by one bulk-TMA copy) */
int sg4d_pack_weight(int n, int k, int ldw, const float *w, float *img, sg4d_stream_t stream);
/* a (rows, lda) fp32, k valid columns (k % 4 == 0, k <= 256); scale/shift (k) or NULL (identity, no ReLU);
 * n in {64, 128}; y (rows, n) or NULL; partial: sg4d_mlp_partial_doubles(rows) fp64 sums for
 * sg4d_bn_finalize; group = nsample for the fused max-pool (0 = none; a power of two dividing rows; <= 128 for n = 128, <= 64 for n = 64):
 * gsel (rows/group, n) = max (gamma[c] >= 0) or min (gamma[c] < 0) of y over the group, garg = its row. */
int sg4d_linear_fwd(long long rows, int k, int lda, int n, int group, const float *a, const float *scale,
                    const float *shift, const float *wimg, float *y, double *partial, const float *gamma,
                    float *gsel, uint8_t *garg, sg4d_stream_t stream);
/* BatchNorm2d batch statistics (nn.BatchNorm2d semantics: biased variance for normalisation, unbiased
 * for the running estimate, momentum update; running_* may be NULL):
 *   scale = gamma / sqrt(var + eps), shift = beta - mean * scale, save_mean, save_invstd. */
int sg4d_bn_finalize(int n, int nparts, long long rows, const double *partial, const float *gamma,
                     const float *beta, float eps, float momentum, float *running_mean, float *running_var,
                     float *scale, float *shift, float *save_mean, float *save_invstd, sg4d_stream_t stream);

/* ---- backward of one scale (replaces the autograd graph of conv/BN/ReLU/max_pool2d, modules.py:66-70) ----
 * Notation: layer 1 = K0 -> C1 (pre-activation y1), layer 2 = C1 -> C2 (pre-activation y2), group = nsample.
 * dY tensors are never materialised: the operand stagers evaluate the BatchNorm / ReLU / max-pool backward
 * formulas on the fly from y1, y2 and per-channel constants prepared on the host side (mlp.py).          */

/* Prologue of the pooled-layer backward, one pass over the (groups, n) pooled tensors (n in {64, 128}):
 *   dz = d_out .* [out > 0];  dsel = dz .* s2;  partial <- per-CTA fp64 sums of (dz, dz .* (gsel - m2) .* i2)
 * d_out has row stride ldd (a slice of the concatenated MSG gradient), out / gsel / dsel are dense.  partial holds
 * sg4d_pool_bwd_prologue_parts() * n pairs; sg4d_partial_sums(n, parts * n, ...) folds them into (d_beta2, d_gamma2). */
int sg4d_pool_bwd_prologue_parts(void);
int sg4d_pool_bwd_prologue(long long groups, int n, int ldd, const float *d_out, const float *out, const float *gsel,
                           const float *s2, const float *m2, const float *i2, float *dsel, double *partial,
                           sg4d_stream_t stream);

/* dz1 = ((dY2) * W2) .* [y1*es + et > 0]   with  dY2[r,c] = dsel[r/group,c]*[r%group == garg[r/group,c]]
 *                                                          - (a2[c]*y2[r,c] + b2[c])
 * k = C2, n = C1; wimg_t = packed image of W2^T (n x k); also emits partial sums (sum dz1, sum dz1*yhat1),
 * yhat1 = y1*ei + em, for sg4d_partial_sums. */
int sg4d_pool_bwd_da(long long rows, int k, int n, int group, const float *y2, const float *a2, const float *b2,
                     const float *dsel, const uint8_t *garg, const float *wimg_t, const float *y1,
                     const float *es, const float *et, const float *ei, const float *em, float *dz1,
                     double *partial, sg4d_stream_t stream);
/* dW2 (m x n) = dY2^T * relu(y1*s1 + t1);  m = C2, n = C1;  partial: sg4d_wgrad_partial_floats(rows, n) floats */
int sg4d_pool_bwd_dw(long long rows, int m, int n, int group, const float *y2, const float *a2, const float *b2,
                     const float *dsel, const uint8_t *garg, const float *y1, const float *s1, const float *t1,
                     float *partial, float *dw, sg4d_stream_t stream);
/* dX(:, col0:col0+n) = dY1 * W1(:, cols)   with  dY1[r,c] = p1[c]*dz1[r,c] - (q1[c]*y1[r,c] + u1[c]);
 * k = C1; wimg_t = packed image of the (n x k) slice of W1^T */
int sg4d_inner_bwd_dx(long long rows, int k, int n, const float *y1, const float *dz1, const float *p1,
                      const float *q1, const float *u1, const float *wimg_t, float *dx, int lddx, int col0,
                      sg4d_stream_t stream);
/* dW1 (m x k) = dY1^T * x(:, 0:k);  m = C1;  partial: sg4d_wgrad_partial_floats(rows, pad(k)) floats,
 * pad(k) = 32 / 64 / 128 / 224 */
int sg4d_inner_bwd_dw(long long rows, int m, int k, int ldx, const float *y1, const float *dz1, const float *p1,
                      const float *q1, const float *u1, const float *x, float *partial, float *dw,
                      sg4d_stream_t stream);
long long sg4d_wgrad_partial_floats(long long rows, int npad);
/* BatchNorm-backward coefficients of one layer from its per-channel sums d_beta = sum dz, d_gamma = sum dz xhat:
 * coef (3, n): q = scale d_gamma invstd / rows, u = scale d_beta / rows - q mean (both 0 unless batch_stats), -mean invstd. */
int sg4d_bn_bwd_coeffs(int n, long long rows, int batch_stats, const float *d_beta, const float *d_gamma, const float *scale,
                       const float *mean, const float *invstd, float *coef, sg4d_stream_t stream);
/* out[0:n] = sum of the first, out[n:2n] = sum of the second component of the fp64 partial pairs */
int sg4d_partial_sums(int n, int nparts, const double *partial, float *out, sg4d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Section 4 -- fused set-abstraction scales: ball-query indices -> shared MLP, the grouped tensor
 * (QueryAndGroup's output, OPS/pointnet2_utils.py:300-337) is never materialised.
 *
 * Grouped-row source ("SRC" below), common to every entry point of this section:
 *   rows = b * m * ns grouped rows; row r = (cloud * m + centre) * ns + slot; neighbour i = idx[r]
 *   pts (b, n, pstride): xyz in columns 0..2;  feats (b, n, fstride): the c feature channels are columns
 *   foff..foff+c-1 (feats may be pts itself, or NULL when c == 0);  centers (b, m, 3);  ns a power of two.
 * The grouped row is x = [xyz(i) - centre | feats(i)] for SA1 (the reference's channel order,
 * utils.py:326-328) and [feats(i) | xyz(i) - centre | 0-pad] for the *_grouped entry points (feature columns
 * 16-byte aligned; the caller permutes the first layer's weight columns accordingly).
 * ---------------------------------------------------------------------------------------------- */

/* SA1 (k = 3 + c <= 7): per-CTA partials (sg4d_sa_moments_parts() x 36 doubles, upper triangle) of
 * M = sum over rows of [x | 1][x | 1]^T.  The first layer is linear in x, so BatchNorm1's batch statistics are a
 * function of M: replaces the statistics pass of nn.BatchNorm2d over the (B, 64, npoint, nsample) tensor. */
int sg4d_sa_moments_parts(void);
int sg4d_sa_moments(long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c, const float *pts,
                    const float *feats, const float *centers, const int32_t *idx, double *part, sg4d_stream_t stream);
/* w1 (64, k) row stride ldw.  Folds the partials into moments (8 x 8 fp64, kept for the backward pass) and emits
 * stats (4, 64) = BatchNorm1 scale, shift, mean, invstd (batch statistics; running_* updated like nn.BatchNorm2d
 * in training mode, or USED instead when use_running != 0) and w1s (8, 64) = scale .* W1, input-major, zero rows >= k. */
int sg4d_sa1_bn1(int k, int nparts, const double *part, const float *w1, int ldw, const float *gamma, const float *beta,
                 float eps, float momentum, float *running_mean, float *running_var, int use_running, double *moments,
                 float *stats, float *w1s, sg4d_stream_t stream);
/* y2 (rows, n2) = relu(W1s x + t1) * W2^T with the first layer recomputed from the gathered neighbour in the operand
 * stagers (8 FMAs per activation; y1 never exists).  t1 = stats + 64.  n2 in {64, 128}; wimg2 = sg4d_pack_weight of
 * W2 (n2, 64); partial / gamma2 / gsel / garg exactly as sg4d_linear_fwd with group = ns (8 <= ns <= 64, 128 for n2 = 128). */
int sg4d_sa1_fwd(long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c, const float *pts,
                 const float *feats, const float *centers, const int32_t *idx, const float *w1s, const float *t1, int n2,
                 const float *wimg2, float *y2, double *partial, const float *gamma2, float *gsel, uint8_t *garg,
                 sg4d_stream_t stream);
/* Single-pass backward of the first layer: dz1 = ((dY2) * W2) .* [W1s x + t1 > 0] (dY2 as in sg4d_pool_bwd_da) is
 * consumed where it is produced; s1part (sg4d_sa1_s1part_doubles(rows)) receives per-CTA partials of
 * S1 = dz1^T [x | 1] (64 x 8).  wimg2_t = packed image of W2^T (64, n2). */
long long sg4d_sa1_s1part_doubles(long long rows);
int sg4d_sa1_bwd_da(long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c, const float *pts,
                    const float *feats, const float *centers, const int32_t *idx, const float *w1s, const float *t1,
                    int n2, const float *y2, const float *a2, const float *b2, const float *dsel, const uint8_t *garg,
                    const float *wimg2_t, double *s1part, sg4d_stream_t stream);
/* dW2 (n2 x 64) = dY2^T * relu(W1s x + t1), second operand recomputed; partial: sg4d_wgrad_partial_floats(rows, 64) */
int sg4d_sa1_bwd_dw2(long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c, const float *pts,
                     const float *feats, const float *centers, const int32_t *idx, const float *w1s, const float *t1,
                     int n2, const float *y2, const float *a2, const float *b2, const float *dsel, const uint8_t *garg,
                     float *partial, float *dw2, sg4d_stream_t stream);
/* Same result as sg4d_sa1_bwd_dw2 without reading y2: dY2 = [winner] dsel - (a2 y2 + b2) and y2 = W2 h1 give
 *   dW2 = T1 - diag(a2) W2 (h1^T h1) - b2 (x) colsum(h1),   T1[c, :] = sum_g dsel(g, c) h1[winning row of (g, c), :];
 * the Gram matrix comes from one 64-wide tensor-core operand, T1 from one row per (group, channel).  w2 (n2, 64) row-major;
 * ws: sg4d_sa1_bwd_dw2_gram_ws_floats(rows, n2) floats of scratch. */
long long sg4d_sa1_bwd_dw2_gram_ws_floats(long long rows, int n2);
int sg4d_sa1_bwd_dw2_gram(long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c,
                          const float *pts, const float *feats, const float *centers, const int32_t *idx,
                          const float *w1s, const float *t1, int n2, const float *w2, const float *a2, const float *b2,
                          const float *dsel, const uint8_t *garg, float *ws, float *dw2, sg4d_stream_t stream);
/* d_beta1 = S1[:, 7];  d_gamma1 = i1 .* (sum_j W1 .* S1 - m1 .* d_beta1);  dW1 (64 x k, row stride lddw) =
 * p1 .* S1 - q1 .* (W1 M) - u1 (x) M[:, 7]  (BatchNorm backward, linear in S1 and M; q1 = u1 = 0 unless batch_stats) */
int sg4d_sa1_bwd_finalize(int k, long long rows, const double *s1part, const double *moments, const float *w1, int ldw,
                          const float *stats, int batch_stats, float *d_w1, int lddw, float *d_g1, float *d_be1,
                          sg4d_stream_t stream);

/* SA2-style first layer (c % 4 == 0, c + 4 <= 256): y (rows, nout) = x * W^T with x gathered by the operand stagers;
 * wimg = sg4d_pack_weight of the (nout, c + 4) weight in grouped column order; partial as sg4d_linear_fwd. */
int sg4d_linear_fwd_grouped(long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c,
                            const float *pts, const float *feats, const float *centers, const int32_t *idx, int nout,
                            const float *wimg, float *y, double *partial, sg4d_stream_t stream);
/* dW1 (mout x (c + 3), row stride lddw, grouped column order) = dY1^T * x, dY1 as in sg4d_inner_bwd_dx;
 * 128 < c + 3 <= 224;  partial: sg4d_wgrad_partial_floats(rows, 224) */
int sg4d_inner_bwd_dw_grouped(long long rows, int n, int m, int ns, int pstride, int fstride, int foff, int c,
                              const float *pts, const float *feats, const float *centers, const int32_t *idx, int mout,
                              const float *y1, const float *dz1, const float *p1, const float *q1, const float *u1,
                              float *partial, float *dw, int lddw, sg4d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Section 5 -- dense layers on the same tensor-core engine: the GroupAll level SA3 (OPS/pointnet2_modules.py:130-146),
 * the TripletGCN MLPs (SGH/model/gcns/network_TripletGCN.py:11-58) and the classifier heads
 * (SGH/model/pointnets/network_PointNet.py:188-271).  n = output channels, a multiple of 64 (the host pads narrower
 * layers); wide layers run as column panels of 128 (64 if n % 128 != 0) in ONE launch.
 * ---------------------------------------------------------------------------------------------- */
long long sg4d_dense_weight_floats(int n, int k);
int sg4d_dense_pack_weight(int n, int k, int ldw, const float *w, float *img, sg4d_stream_t stream);
long long sg4d_dense_partial_doubles(long long rows, int n);
/* y (rows, n; row stride ldy) = act(a) * W^T [+ bias];  act = identity (scale NULL) or relu(a .* scale + shift) (k <= 1280).
 * partial != NULL: also the per-channel sum / sum of squares for sg4d_dense_bn_finalize (bias must be NULL: a bias in
 * front of a BatchNorm cancels) and, with group > 0, the fused max/min over groups of rows exactly like sg4d_linear_fwd
 * (gsel / garg: (rows / group, n)). */
int sg4d_dense_fwd(long long rows, int k, int lda, int n, const float *a, const float *scale, const float *shift,
                   const float *wimg, const float *bias, float *y, int ldy, double *partial, int group, const float *gamma,
                   float *gsel, uint8_t *garg, sg4d_stream_t stream);
/* stats (4, n) = scale, shift, mean, invstd (nn.BatchNorm1d / BatchNorm2d batch statistics; running_* updated when given) */
int sg4d_dense_bn_finalize(int n, long long rows, const double *partial, const float *gamma, const float *beta, float eps,
                           float momentum, float *running_mean, float *running_var, float *stats, sg4d_stream_t stream);
/* out = relu(y .* scale + shift) */
int sg4d_bn_relu_apply(long long rows, int n, const float *y, int ldy, const float *scale, const float *shift, float *out,
                       int ldo, sg4d_stream_t stream);
/* Backward prologue of [BatchNorm -> ReLU]: dz = dh .* [h > 0]; dzs = dz .* scale; sums (2, n) = (sum dz, sum dz .* xhat)
 * = (d_beta, d_gamma).  part: sg4d_colsum_part_doubles(rows, n) doubles of scratch. */
long long sg4d_colsum_part_doubles(long long rows, int n);
int sg4d_bn_relu_bwd(long long rows, int n, const float *dh, int lddh, const float *h, int ldh, const float *y, int ldy,
                     const float *stats, float *dzs, int lddz, double *part, float *sums, sg4d_stream_t stream);
/* out (2, n): row 0 = column sums of a (bias gradients), row 1 = 0 */
int sg4d_colsum(long long rows, int n, const float *a, int lda, double *part, float *out, sg4d_stream_t stream);
/* dX (rows, nout) = dY * W [.* [e .* es + et > 0]];  dY (rows, kk) = a (mode 0) or a2 .* p1 - (a .* q1 + u1) (mode 3, the
 * BatchNorm backward evaluated in the operand stagers);  wimg_t = sg4d_dense_pack_weight of W^T (nout x kk);
 * partial (sg4d_dense_partial_doubles(rows, nout)) is scratch, required with e. */
int sg4d_dense_bwd_dx(long long rows, int kk, int lda, int nout, int mode, const float *a, const float *a2, const float *p1,
                      const float *q1, const float *u1, const float *wimg_t, const float *e, int lde, const float *es,
                      const float *et, float *dx, int lddx, double *partial, sg4d_stream_t stream);
/* dW (m x k, row stride lddw) = dY^T * act(x);  act(x) = x (xs NULL) or relu(x .* xs + xt);  computed as 128 x 128 blocks in
 * one launch;  partial: sg4d_dense_wgrad_partial_floats(rows, m, k) floats of scratch. */
long long sg4d_dense_wgrad_partial_floats(long long rows, int m, int k);
int sg4d_dense_bwd_dw(long long rows, int m, int lda, int k, int mode, const float *a, const float *a2, const float *p1,
                      const float *q1, const float *u1, const float *x, int ldx, const float *xs, const float *xt,
                      float *partial, float *dw, int lddw, sg4d_stream_t stream);
/* dW2 (m x n) = dY2^T * relu(y1 .* s1 + t1) for wide pooled layers (dY2 as in sg4d_pool_bwd_da; m = C2, n = C1);
 * partial: sg4d_dense_wgrad_partial_floats(rows, m, n).  sg4d_pool_bwd_da itself accepts n = C1 that is a multiple of 128
 * (column panels, wimg_t = the dense image of W2^T, partial = sg4d_dense_partial_doubles(rows, n)) and k = C2 <= 1280. */
int sg4d_dense_pool_bwd_dw(long long rows, int m, int n, int group, const float *y2, const float *a2, const float *b2,
                           const float *dsel, const uint8_t *garg, const float *y1, const float *s1, const float *t1,
                           float *partial, float *dw, sg4d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Section 6 -- crop / sample front-end: one scene (points (P, stride) fp32 with xyz first, per-point
 * object masks: 0 = none, i + 1 = object i) -> the per-object and per-edge clouds the encoders consume.  Replaces the numpy /
 * open3d loop of SGH/dataset/data_preparation_utils.py:104-125 (objects), :178-218 (edges), :12-18 (zero_mean), :37-39 (the
 * replace=True draw).  Draws are an input: index = candidates[min(floor(u * len), len - 1)], candidates in ascending point order.
 * ---------------------------------------------------------------------------------------------- */
long long sg4d_frontend_workspace_bytes(int P, int nobj, int E);
/* obj_list (nobj, P) int32: members of every object in ascending point index; totals (nobj + 1): [0] unlabelled, [i + 1] members
 * of object i; obj_box (nobj, 6): {min xyz - padding, max xyz + padding}.  nobj <= 31. */
int sg4d_frontend_objects(int P, int stride, int nobj, const float *pts, const int32_t *masks, float padding, void *ws,
                          int32_t *obj_list, int *totals, float *obj_box, sg4d_stream_t stream);
/* edges (2, E) int64 (subject, object) object indices; edge_box (E, 6) = union of the two padded boxes; edge_list (E, P): points
 * STRICTLY inside it, ascending; edge_totals (E, 2): [e][1] = their number. */
int sg4d_frontend_edges(int P, int stride, int E, const float *pts, const int32_t *masks, const int64_t *edges,
                        const float *obj_box, void *ws, int32_t *edge_list, int *edge_totals, float *edge_box,
                        sg4d_stream_t stream);
/* out (clouds, n, fout): drawn points [xyz | features | (edges only) 1 / 2 for points of the subject / object instance], xyz
 * centred and scaled to the unit sphere; fout = stride (+ 1 with edges).  list / totals as written by the two calls above
 * (edges == NULL: object clouds).  picked (clouds, n) original point indices (may be NULL); mean (clouds, 3); dist (clouds);
 * scratch: clouds * (ceil(n / 256) * 24 + 4 + 4 n) bytes (partial sums, max norms and -- when picked is NULL -- the draws). */
int sg4d_frontend_sample(int P, int stride, int clouds, int n, const float *pts, const int32_t *masks, const int32_t *list,
                         const int *totals, const int64_t *edges, const float *u, float *out, int32_t *picked, float *mean,
                         float *dist, void *scratch, sg4d_stream_t stream);

/* Process-wide compute precision of every tensor-core layer of sections 3-5 (not thread-safe; set it before launching):
 *   0  fp32-level accuracy: 3xTF32 (default; what the parity bounds and the headline benchmark use)
 *   1  bf16: operands rounded to bf16 as they are staged, one product per k-step, fp32 accumulation, fp32 statistics and
 *      activations in memory -- BASELINE.json configs[3]; the reference's counterpart is `precision=16` autocast
 *      (scene_graph_prediction/main.py:62-64). */
int sg4d_set_compute_precision(int mode);
int sg4d_get_compute_precision(void);

#ifdef __cplusplus
}
#endif
#endif /* SG4D_H_ */
