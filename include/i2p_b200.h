/*
 * i2p_b200.h -- C ABI of libi2p_b200.so: the I2PNet hot-path operators as hand-written
 * sm_100a CUDA kernels.
 *
 * Conventions (they mirror the reference's pybind layer):
 *   - every pointer is a DEVICE pointer into memory the caller owns; outputs are written
 *     in place and the library keeps no reference to them;
 *   - tensors are dense, row-major ("contiguous"), float = IEEE f32, indices int32 unless
 *     the reference's binding uses int64 (the fused select outputs, kNN);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream), never synchronises, never allocates, and is therefore
 *     capturable into a CUDA graph;
 *   - return value: 0 on success, I2P_ERR_* otherwise.  Nothing calls exit(); the
 *     reference's launchers do (e.g. pointnet2/src/group_points_gpu.cu:81-85).
 *   - `citations` are relative to the reference repository IRMVLab/I2PNet.
 */
#ifndef I2P_B200_H
#define I2P_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define I2P_OK 0
#define I2P_ERR_INVALID_ARGUMENT 1 /* shape / size outside what the operator defines      */
#define I2P_ERR_UNSUPPORTED 2      /* defined by the reference but beyond this build       */
#define I2P_ERR_CUDA 3             /* a CUDA runtime error; see i2p_last_error()           */

#define I2P_FLAG_COPY 1  /* fused_conv_select_k.py:6 */
#define I2P_FLAG_SHIFT 2 /* fused_conv_select_k.py:7 */

/* Human-readable description of the last error raised on the calling thread. */
const char *i2p_last_error(void);
/* ABI version of this header (bumped when a signature changes). */
int i2p_abi_version(void);
/* Number of kernel launches issued through this library since load (all threads). */
uint64_t i2p_launch_count(void);

/* ---- PointNet++ operators: replace pointnet2/src/pointnet2_api.cpp:10-24 ------------- */

/* furthest_point_sampling_wrapper  (pointnet2/src/sampling.cpp:40-51, sampling_gpu.cu:93-253)
 * dataset (B,N,3); temp (B,N) scratch pre-filled by the caller with 1e10 (on return it
 * holds each point's squared distance to the sampled set, as the reference leaves it);
 * idxs (B,M).  Starts at index 0; ties resolve exactly as the reference's tree reduction. */
int i2p_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp, int32_t *idxs,
                                void *stream);

/* gather_points_wrapper / _grad_wrapper  (sampling.cpp:11-37, sampling_gpu.cu:8-79)
 * points (B,C,N), idx (B,M) -> out (B,C,M);  grad_out (B,C,M) -> grad_points (B,C,N) += */
int i2p_gather_points(int b, int c, int n, int npoints, const float *points, const int32_t *idx, float *out,
                      void *stream);
int i2p_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int32_t *idx,
                           float *grad_points, void *stream);

/* ball_query_wrapper  (ball_query.cpp:14-31, ball_query_gpu.cu:9-66)
 * new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample), pre-zeroed by the caller. */
int i2p_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                   int32_t *idx, void *stream);

/* group_points_wrapper / _grad_wrapper  (group_points.cpp:10-44, group_points_gpu.cu:8-86)
 * points (B,C,N), idx (B,npoints,nsample) -> out (B,C,npoints,nsample); grad: += */
int i2p_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int32_t *idx,
                     float *out, void *stream);
int i2p_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                          const int32_t *idx, float *grad_points, void *stream);

/* three_nn_wrapper  (interpolate.cpp:14-27, interpolate_gpu.cu:9-73)
 * unknown (B,N,3), known (B,M,3) -> dist2 (B,N,3) SQUARED distances, idx (B,N,3) */
int i2p_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int32_t *idx,
                 void *stream);

/* three_interpolate_wrapper / _grad_wrapper  (interpolate.cpp:30-60, interpolate_gpu.cu:77-160)
 * points (B,C,M), idx/weight (B,N,3) -> out (B,C,N);  grad_out (B,C,N) -> grad_points (B,C,M) += */
int i2p_three_interpolate(int b, int c, int m, int n, const float *points, const int32_t *idx,
                          const float *weight, float *out, void *stream);
int i2p_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int32_t *idx,
                               const float *weight, float *grad_points, void *stream);

/* knn_wrapper -- called by pointnet2/pointnet2_utils.py:32 but never bound by the reference.
 * unknown (B,N,3), known (B,M,3) -> dist2 (B,N,k) squared distances ascending, idx (B,N,k). */
int i2p_knn(int b, int n, int m, int k, const float *unknown, const float *known, float *dist2, int32_t *idx,
            void *stream);

/* ---- projection-aware operators: replace src/projectPN ------------------------------ */

/* fused_conv_select_k  (fused_conv_select/fused_conv_g.cpp:15-67, fused_conv_go.cu:11-264)
 * xyz1 (B,H,W,3), xyz2 (B,small_h,small_w,3), idx_n2 (B,npoints,2) centre (h,w) in xyz1,
 * random_hw (kH*kW) window visiting order -> selected_{b,h,w}_idx (B,npoints,K) int64 and
 * selected_mask (B,npoints,K) f32, all pre-zeroed by the caller and only partially written,
 * exactly as the reference does.  The reference's valid_idx / valid_in_dis_idx arguments are
 * never written by it (fused_conv_go.cu:150,167) and are therefore not part of this ABI. */
int i2p_fused_conv_select_k(int batch, int H, int W, int npoints, int kH, int kW, int K, int flag,
                            float distance, int stride_h, int stride_w, const float *xyz1, const float *xyz2,
                            const int32_t *idx_n2, const int32_t *random_hw, int64_t *selected_b_idx,
                            int64_t *selected_h_idx, int64_t *selected_w_idx, float *selected_mask,
                            int small_h, int small_w, void *stream);

/* Same selection, compact output for the fused consumers below: flat = h*small_w + w as int32
 * (B,npoints,K) and mask (B,npoints,K); both FULLY written (0 where the reference leaves its
 * pre-zeroed buffers untouched).  Centres are the regular grid (h*stride_ch, w*stride_cw),
 * out_h x out_w of them, as built by get_stride_idx_cuda (src/projectPN/utils.py:28-33) when
 * idx_n2 == NULL. */
int i2p_select_k_flat(int batch, int H, int W, int npoints, int kH, int kW, int K, int flag, float distance,
                      int stride_h, int stride_w, const float *xyz1, const float *xyz2, const int32_t *idx_n2,
                      int out_w, int stride_ch, int stride_cw, int32_t *flat_idx, float *mask, int small_h,
                      int small_w, void *stream);

/* gather_torch  (src/projectPN/utils.py:36-60): feature (B,HW,C) channels-last, flat index
 * (B,M) int32 -> out (B,M,C).  Backward: grad_feature (B,HW,C) += scatter of grad_out. */
int i2p_gather_rows(int b, int hw, int c, int m, const float *feature, const int32_t *flat_idx, float *out,
                    void *stream);
/* First-level set-abstraction operand (ProjectPointNet.forward_center, src/projectPN/PPBackbone_center.py:150-178) in one
 * pass: out (b, n, k, 10) = [ g - c | centre | g | |g - c| ], g = src (b,hw,3)[idx (b,n,k)], c = ctr (b,n,3), centre = cen (b,n,3).
 * Forward only (the level-1 coordinates are data). */
int i2p_sa_geometry(int b, int hw, int n, int k, const float *src, const float *ctr, const float *cen, const int32_t *flat_idx,
                    float *out, void *stream);
int i2p_gather_rows_grad(int b, int hw, int c, int m, const float *grad_out, const int32_t *flat_idx,
                         float *grad_feature, void *stream);

/* knn_point  (src/projectPN/utils.py:344-380; twins src/modules/point_utils.py:114-177,
 * pointnet_util.py:14-57): the nsample nearest of xyz (B,N,3) for each new_xyz (B,S,3) under
 * dist = -2 q.x + |q|^2 + |x|^2 (f32), returned sorted by (dist, index) as int64 (B,S,nsample).
 * The reference's order is unspecified (topk sorted=False).  dist_out may be NULL. */
int i2p_knn_point(int b, int n, int s, int nsample, const float *xyz, const float *new_xyz, int64_t *group_idx,
                  float *dist_out, void *stream);

/* project_seq with rank=False  (src/projectPN/utils.py:111-187): spherical range-image
 * scatter.  xyz (B,N,3) -> cell (row,col) by the reference's truncating formulas; each of the
 * nfeat feature arrays feat[j] (B,N,fdim[j]) and xyz itself are written to (B,H,W,.) images,
 * zero where no point lands.  Duplicate cells: the highest point index wins (the reference's
 * index_put_ is last-writer-wins on CPU and unordered on CUDA).  owner (B,H,W) int32 scratch.
 * feats / outs: DEVICE-resident float pointers passed by value in host arrays. */
int i2p_project_seq(int b, int n, int H, int W, float fup_deg, float fdown_deg, const float *xyz, int nfeat,
                    const float *const *feats, const int *fdims, float *xyz_proj, float *const *feat_projs,
                    int32_t *owner, void *stream);

/* ---- shared per-point MLP: replaces the Conv2d block of src/projectPN/PPBackbone_center.py:10-46 ----
 * (permute -> 1x1 nn.Conv2d -> BatchNorm2d on batch statistics -> (Leaky)ReLU -> permute), forward and
 * backward, on channels-last (rows, C) tensors.  Only the raw pre-normalisation outputs y are
 * materialised; normalise + activation are folded into the next kernel's operand load.
 * Activation: act(z) = z > 0 ? z : slope * z  (slope 0 = ReLU, 0.1 = LeakyReLU(0.1), 1 = identity). */

/* Which GEMMs of the shared MLP run on the tensor cores (tcgen05.mma kind::tf32 with a 3-term hi/lo split,
 * f32-accurate, accumulator in TMEM), as a bit mask: 1 forward, 2 dX, 4 dW (the i2p_pw_*_tc entry points
 * below; the host consults this mask), 8 = the first-generation forward kernel inside i2p_pw_linear_fwd.
 * 16 = SWIZZLE_128B operand tiles + bulk-copied (TMA engine) weights in the forward / dX kernels.
 * 32 = the host also sends max-over-k gradient sources (dout + arg-max) to the tensor-core dX / dW kernels.
 * 64 = (with 16) two MMAs per k-step instead of three in the forward / dX kernels: A_hi x [B_hi | B_lo] as one MMA of
 *      N = 2 BN, the correction terms in their own accumulator (one thread issues ~120 cycles per MMA whatever its N).
 * Default: environment variable I2P_MLP_TC, else 119 (= 7 | 16 | 32 | 64).  0 = f32 FMA kernels everywhere. */
void i2p_set_mlp_tensor_cores(int mask);
int i2p_get_mlp_tensor_cores(void);
/* Number of 128-row tiles (rows of the tile_stats buffer). */
int i2p_pw_num_tiles(int rows);
/* y = f(x) W^T + bias, f(x) = act(x * in_scale + in_shift) or x when in_scale == NULL.
 * x (rows,cin), w (cout,cin), y (rows,cout); tile_stats (ntiles,cout,2) per-tile (mean, M2) or NULL. */
int i2p_pw_linear_fwd(int rows, int cin, int cout, const float *x, const float *in_scale, const float *in_shift,
                      float in_slope, const float *w, const float *bias, float *y, float *tile_stats, void *stream);
/* Exact (f64, Chan) merge of the tile statistics -> mean, rstd = 1/sqrt(var_biased + eps),
 * scale = gamma * rstd, shift = beta - mean * scale.  All (cout). */
int i2p_bn_finalize(int rows, int cout, const float *tile_stats, const float *gamma, const float *beta, float eps,
                    float *mean, float *rstd, float *scale, float *shift, void *stream);
/* out = act(y * scale + shift), (rows, c), c % 4 == 0 */
int i2p_bn_act(long long rows, int c, const float *y, const float *scale, const float *shift, float slope, float *out,
               void *stream);
/* out[g,c] = max_k act(y[g,k,c] * scale + shift), arg[g,c] = first k attaining it; y (groups,k,c) */
int i2p_bn_act_maxk(long long groups, int k, int c, const float *y, const float *scale, const float *shift,
                    float slope, float *out, int32_t *arg, void *stream);
/* Batch-norm backward sums of a layer: s12[0,c] = sum_r dz, s12[1,c] = sum_r dz * yhat (f64, pre-zeroed),
 * dz = g * act'(y*scale+shift).  g is either dense (rows,c) or, when g_dense == NULL, the gradient
 * dout (rows/k, c) of a max-over-k output routed through arg. */
int i2p_bn_bwd_reduce(long long rows, int c, const float *g_dense, const float *dout, const int32_t *arg, int k,
                      const float *y, const float *mean, const float *rstd, const float *scale, const float *shift,
                      float slope, double *s12, void *stream);
/* dx (rows,cin) = dY W with dY = scale * (dz - S1/n - yhat * S2/n) rebuilt on the fly; when prev_y != NULL
 * (the input was itself a normalised layer, raw output prev_y (rows,cin)) the epilogue accumulates that
 * layer's s12 into prev_s12 (pre-zeroed). */
int i2p_pw_linear_bwd_dx(int rows, int cin, int cout, const float *g_dense, const float *dout, const int32_t *arg,
                         int k, const float *y, const float *mean, const float *rstd, const float *scale,
                         const float *shift, float slope, const double *s12, const float *w, float *dx,
                         const float *prev_y, const float *prev_mean, const float *prev_rstd, const float *prev_scale,
                         const float *prev_shift, float prev_slope, double *prev_s12, void *stream);
/* dw (cout,cin) += dY^T A, A = act(x * prev_scale + prev_shift) or x when prev_scale == NULL; dw pre-zeroed. */
int i2p_pw_linear_bwd_dw(int rows, int cin, int cout, const float *g_dense, const float *dout, const int32_t *arg,
                         int k, const float *y, const float *mean, const float *rstd, const float *scale,
                         const float *shift, float slope, const double *s12, const float *x, const float *prev_scale,
                         const float *prev_shift, float prev_slope, float *dw, void *stream);

/* ---- tensor-core variants of the three shared-MLP GEMMs (csrc/mlp_tc.cu) --------------------------
 * Same algebra and tensors as i2p_pw_linear_fwd / _bwd_dx / _bwd_dw; the weight operand comes from a
 * per-layer pack (tf32 hi/lo halves in the UMMA shared-memory layout) written by i2p_pw_pack_weights.
 * Covered shapes: i2p_pw_tc_supported(kind, ...) with kind 0 forward, 1 dX, 2 dW (cout % 32 == 0,
 * cout >= 64; backward: cout <= 256, cin >= 32; gradient source dense or max-over-k as in i2p_bn_bwd_reduce).
 * Activation slopes in [0, 1].
 * All float tensors must be 16-byte aligned except x, which may be 4-byte aligned. */
int i2p_pw_tc_supported(int kind, int rows, int cin, int cout);
long long i2p_pw_pack_floats(int cin, int cout);
int i2p_pw_pack_weights(int cin, int cout, const float *w, float *pack, void *stream);
/* The same for n layers in one launch: table (n, 3) int64 on the device, row l = { cin | cout << 32, weight pointer,
 * pack pointer } (the pointers as integers).  A step engine packs every layer once per optimiser step this way. */
int i2p_pw_pack_weights_multi(int n, const long long *table, void *stream);
int i2p_pw_linear_fwd_tc(int rows, int cin, int cout, const float *x, const float *in_scale, const float *in_shift,
                         float in_slope, const float *wpack, const float *bias, float *y, float *tile_stats, void *stream);
int i2p_pw_linear_bwd_dx_tc(int rows, int cin, int cout, const float *g_dense, const float *dout, const int32_t *arg, int k,
                            const float *y, const float *mean, const float *rstd, const float *scale, const float *shift,
                            float slope, const double *s12, const float *wpack, float *dx, const float *prev_y,
                            const float *prev_mean, const float *prev_rstd, const float *prev_scale, const float *prev_shift,
                            float prev_slope, double *prev_s12, void *stream);
int i2p_pw_linear_bwd_dw_tc(int rows, int cin, int cout, const float *g_dense, const float *dout, const int32_t *arg, int k,
                            const float *y, const float *mean, const float *rstd, const float *scale, const float *shift,
                            float slope, const double *s12, const float *x, const float *prev_scale, const float *prev_shift,
                            float prev_slope, float *dw, void *stream);

/* ---- RGB feature-pyramid block tail: replaces BatchNorm2d -> LeakyReLU(0.1) -> MaxPool2d(3, stride, 1)
 * of src/modules/basicConv.py:11-17 on the NCHW f32 output y (B,C,H,W) of the block's convolution.
 * Neither the normalised nor the activated tensor is materialised; backward gathers through the
 * pooling windows (no atomics).  stats (4,C) = mean, rstd, scale = gamma*rstd, beta; z = (y-mean)*scale+beta. */

/* chunks of one (sample, channel) plane the statistics pass splits into; pooled extent (n+2-3)/stride+1 */
int i2p_rgb_num_chunks(int hw);
int i2p_rgb_pool_out(int n, int stride);
/* tile_stats (C, B*chunks, 3) = (count, mean, M2) per chunk */
int i2p_rgb_bn_stats(int B, int C, int H, int W, const float *y, float *tile_stats, void *stream);
/* Chan merge -> stats; running_mean/var (may be NULL) updated with `momentum` and the unbiased variance,
 * *num_batches_tracked += 1 (nn.BatchNorm2d in training); s12 (i2p_rgb_s12_slots(),C) f64 (may be NULL) zeroed. */
int i2p_rgb_bn_finalize(int C, int ntiles, const float *tile_stats, const float *gamma, const float *beta, float eps,
                        float momentum, float *running_mean, float *running_var, long long *num_batches_tracked,
                        float *stats, double *s12, void *stream);
/* eval mode: stats from the running statistics */
int i2p_rgb_bn_from_running(int C, const float *gamma, const float *beta, float eps, const float *running_mean,
                            const float *running_var, float *stats, double *s12, void *stream);
/* out (B,C,Ho,Wo) = maxpool3x3(leaky((y-mean)*scale+beta)); ties: first maximum in scan order, like ATen */
int i2p_rgb_bn_act_pool_fwd(int B, int C, int H, int W, int stride, const float *y, const float *stats, float slope,
                            float *out, void *stream);
/* number of (C)-rows of the f64 s12 buffer the finalize kernels zero and the backward accumulates into */
int i2p_rgb_s12_slots(void);
/* dout (B,C,Ho,Wo) -> dy (B,C,H,W), dgamma (C), dbeta (C); the pooling arg-max is re-derived from y (no saved
 * indices); s12 (i2p_rgb_s12_slots(), C) f64 zeroed by the forward finalize.
 * batch_stats = 1: full batch-norm backward; 0: running statistics (dy = scale * dz). */
int i2p_rgb_bn_act_pool_bwd(int B, int C, int H, int W, int stride, int batch_stats, const float *y, const float *stats,
                            float slope, const float *dout, double *s12, float *dy, float *dgamma, float *dbeta,
                            void *stream);

/* ---- cost-volume glue: replaces the repeat / multiply / mask / max / concatenate and the softmax-weighted
 * sums of CostVolume.forward (src/projectPN/PPBackbone_center.py:366-420, 470-488) ------------------------
 * cv_build: X (B,N,K,Cx) = [xyz1[b,n] (3) | xyz2[b,j] (3) | pi[b,n,:]*qi[b,j,:] (C) | maxc[b,j,:] (C, if maxc)],
 * j = k when idx == NULL (then K == N2) else idx[b,n,k]; xyz6 (B,N,K,6) = the coordinate channels alone.
 * xyz1 (B,N,3), xyz2 (B,N2,3), pi (B,N,C), qi (B,N2,C), maxc (B,N2,C) or NULL, idx (B,N,K) int32 or NULL. */
int i2p_cv_build(int B, int N, int K, int N2, int C, const float *xyz1, const float *xyz2, const float *pi, const float *qi,
                 const float *maxc, const int32_t *idx, float *X, float *xyz6, void *stream);
/* dX (B,N,K,Cx), dxyz6 (B,N,K,6) or NULL -> dxyz1 (B,N,3), dpi (B,N,C), dxyz2 (B,N2,3), dqi (B,N2,C), dmaxc (B,N2,C),
 * all ACCUMULATED with atomics: zero them first. */
int i2p_cv_build_bwd(int B, int N, int K, int N2, int C, int has_max, const float *dX, const float *dxyz6, const float *pi,
                     const float *qi, const int32_t *idx, float *dxyz1, float *dxyz2, float *dpi, float *dqi, float *dmaxc,
                     void *stream);
/* out (G,C) = sum_k softmax_k(l) v, l = logit (G,K,C) [* mask + -1e10 (1 - mask), mask (G,K) or NULL], v = value (G,K,C);
 * stat (G,C,2) or NULL receives (max_k l, 1 / sum_k exp(l - max)) for the backward pass (needs C % 8 == 0) */
int i2p_softmax_wsum(long long groups, int K, int C, const float *logit, const float *value, const float *mask, float *out,
                     float *stat, void *stream);
/* gout (G,C) -> dlogit, dvalue (G,K,C); with stat from the forward an element-wise pass, without it the softmax is
 * recomputed from logit */
int i2p_softmax_wsum_bwd(long long groups, int K, int C, const float *logit, const float *value, const float *mask,
                         const float *out, const float *gout, const float *stat, float *dlogit, float *dvalue, void *stream);

/* Operand preparation of the cost volume (src/projectPN/PPBackbone_center.py:379-397; replaces ~40 element-wise /
 * reduction launches per direction): xyz (B,N,3) = uv * z; pi (B,N,C), qi (B,N2,C) = the point / pixel features
 * standardised over the channel axis, (x - mean) / max(unbiased std, 1e-12), with the clipped denominators in den_p (B,N),
 * den_q (B,N2); has_max: hi / lo (B,C) = largest / smallest pi[:,c] over the valid points (xyz != 0), arg_hi / arg_lo (B,C)
 * the first point attaining it (0 / -1 for a cloud without valid points), maxc (B,N2,C) = qi * (qi > 0 ? hi : lo)
 * (-1e10 without valid points) = max over the valid points of pi[n,c] qi[k,c], the backward-validation channel. */
int i2p_cv_prep_fwd(int B, int N, int N2, int C, int has_max, const float *uv, const float *z, const float *pf, const float *qf,
                    float *xyz, float *pi, float *qi, float *den_p, float *den_q, float *maxc, float *hi, float *lo, int32_t *arg_hi,
                    int32_t *arg_lo, float *scratch, void *stream);
/* has_max: `scratch` = i2p_cv_prep_scratch_floats() 4-byte words whose LAST B words (the clouds' tickets) are zero before the
 * first launch; the kernel leaves them zero, so one buffer zeroed once serves every later launch of these sizes on a stream. */
int i2p_cv_prep_scratch_floats(int B, int N, int N2, int C);
/* d_xyz (B,N,3), d_pi (B,N,C), d_qi, d_maxc (B,N2,C), each or NULL -> d_uv (B,N,3), d_z (B,N), d_pf (B,N,C), d_qf (B,N2,C) */
int i2p_cv_prep_bwd(int B, int N, int N2, int C, int has_max, const float *uv, const float *z, const float *pi, const float *qi,
                    const float *den_p, const float *den_q, const float *hi, const float *lo, const int32_t *arg_hi,
                    const int32_t *arg_lo, const float *d_xyz, const float *d_pi, const float *d_qi, const float *d_maxc, float *d_uv,
                    float *d_z, float *d_pf, float *d_qf, void *stream);
/* rays (B, h*w, 3) = K'^-1 [u, v, 1]: pixel centres of an (h, w) feature map on the normalised camera plane, K' = intrinsic
 * (B,3,3) with row 0 scaled by sx and row 1 by sy (src/modellearn_proj_center.py:275-287: change_intrinsic, the host
 * round trip through torch.inverse, set_id_grid and the batched product). */
int i2p_pixel_rays(int B, int h, int w, float sx, float sy, const float *intrinsic, float *rays, void *stream);

/* ---- quaternion product: replaces mul_q of src/modules/warp_utils.py:25-60 (one launch instead of ~30) ----
 * out (B,N,4) = A (x) B with A = a (B,na,4), B = b (B,nb,4), na / nb in {1, N} (broadcast over points),
 * either operand optionally conjugated (the backward pass is da = dc (x) conj(b), db = conj(a) (x) dc). */
int i2p_quat_mul(int B, int N, int na, int nb, int conj_a, int conj_b, const float *a, const float *b, float *out,
                 void *stream);
/* Rigid warp of a point set by a pose, warp_quat_xyz (src/modules/warp_utils.py:78-94) and the pose composition
 * t = R(q3) t_in + t3 (src/modellearn_proj_center.py:414-421) as one launch per direction:
 * out (B,N,3) = (q (x) [0, p] (x) q^-1)[1:4] + t, q^-1 = conj(q) / (|q|^2 + 1e-10); p (B,N,3), q (B,4), t (B,3).
 * mask_invalid: all-zero points stay zero (the reference multiplies by check_valid afterwards).
 * Backward: g (B,N,3) -> dp (B,N,3), dq (B,4), dt (B,3), each or NULL. */
int i2p_quat_warp_fwd(int B, int N, int mask_invalid, const float *p, const float *q, const float *t, float *out, void *stream);
int i2p_quat_warp_bwd(int B, int N, int mask_invalid, const float *p, const float *q, const float *g, float *dp, float *dq,
                      float *dt, void *stream);

/* ---- 3x3 convolutions of the RGB feature pyramid (src/modules/basicConv.py:6-20; replaces the library convolution
 * behind nn.Conv2d(k = 3, stride 1, padding 1) forward, its data gradient and its weight gradient) -------------------
 * NCHW f32.  i2p_conv3x3_pack lays the (cout, cin, 3, 3) weights out once per step as tf32 (hi | lo) halves in the
 * tensor-core operand layout (dgrad = 1: transposed and flipped, for the data gradient); i2p_conv3x3_tc is the
 * implicit GEMM on tcgen05 (3xTF32, f32-accurate): y (B, no, H, W) = conv(x (B, ki, H, W)) [+ bias], and, when
 * tile_stats (no, i2p_conv3x3_stat_slots(B, no, H, W), 3) is given, one (count, mean, M2) of y per persistent CTA and
 * output channel (the input of i2p_rgb_bn_finalize with ntiles = that slot count).  Data gradient: the same call on dy with the dgrad pack
 * (ki = cout, no = cin).  i2p_conv3x3_wgrad ADDS the weight gradient into dw (cout, cin, 3, 3) (f32 FMA, atomics). */
long long i2p_conv3x3_pack_floats(int cin, int cout, int dgrad);
int i2p_conv3x3_tiles(int H, int W);
int i2p_conv3x3_stat_slots(int B, int no, int H, int W);
int i2p_conv3x3_pack(int cin, int cout, int dgrad, const float *w, float *pack, void *stream);
/* n packs in one launch: table (n, 4) int64 on the device, row = { cin | cout << 32, dgrad, weight pointer, pack pointer } */
int i2p_conv3x3_pack_multi(int n, const long long *table, void *stream);
int i2p_conv3x3_tc(int B, int ki, int no, int H, int W, const float *x, const float *wpack, const float *bias, float *y,
                   float *tile_stats, void *stream);
int i2p_conv3x3_wgrad(int B, int cin, int cout, int H, int W, const float *x, const float *dy, float *dw, void *stream);

/* ---- pose head and pose loss (src/projectPN/PPBackbone_center.py:503-560 PoseHead; compute_loss.py:102-133 Get_loss) ----
 * i2p_pose_head_fwd: mask_p = softmax over the N points of mask (B, N, C) per channel; pooled = sum_n pred * mask_p;
 * hidden = (W1 pooled + b1) * drop (drop (B, Hd): dropout multipliers or NULL); q_raw = Wq hidden + bq, t = Wt hidden + bt,
 * q = q_raw / (|q_raw| + 1e-10).  One CTA per sample.  i2p_pose_head_bwd: gradients of all of it from dq (B, 4), dt (B, 3)
 * (either may be NULL); parameter gradients are ADDED (atomics over the batch) into dw1 (Hd, C), db1, dwq (4, Hd), dbq,
 * dwt (3, Hd), dbt.  i2p_pose_loss_fwd -> loss3 = (total, rotation part, translation part); l1: L1 translation term. */
int i2p_pose_head_fwd(int B, int N, int C, int Hd, const float *pred, const float *mask, const float *w1, const float *b1,
                      const float *wq, const float *bq, const float *wt, const float *bt, const float *drop, float *mask_p,
                      float *pooled, float *hidden, float *q_raw, float *q, float *t, void *stream);
int i2p_pose_head_bwd(int B, int N, int C, int Hd, const float *pred, const float *mask_p, const float *pooled,
                      const float *hidden, const float *q_raw, const float *drop, const float *w1, const float *wq,
                      const float *wt, const float *dq, const float *dt, float *dpred, float *dmask, float *dw1, float *db1,
                      float *dwq, float *dbq, float *dwt, float *dbt, void *stream);
int i2p_pose_loss_fwd(int B, int l1, const float *out3, const float *out4, const float *q_gt, const float *t_gt,
                      const float *sx, const float *sq, float *loss3, void *stream);
int i2p_pose_loss_bwd(int B, int l1, const float *out3, const float *out4, const float *q_gt, const float *t_gt,
                      const float *sx, const float *sq, const float *dloss, float *dout3, float *dout4, float *dsx,
                      float *dsq, void *stream);

/* ---- optimiser step on flat buffers: replaces torch.nn.utils.clip_grad_norm_(parameters, max_norm) followed by
 * torch.optim.Adam(lr, betas, eps, weight_decay).step() (train20v2learn_wandb_proj.py:198-205, 481-483) -----------
 * param, grad, exp_avg, exp_avg_sq: n f32 each (grad 16-byte aligned).  grad holds the SUM of the ranks' gradients
 * (world = 1: the gradient); it is averaged, clipped to total norm max_norm (max_norm <= 0: no clipping) and given
 * L2 weight decay inside the update and left untouched in memory.  state: i2p_optim_state_bytes() bytes, zero-filled
 * once by the caller; it carries the step count across calls.  Two launches.
 * lr > 0: that learning rate; lr <= 0: the f32 at byte i2p_optim_lr_offset() of `state` (so that a step captured into
 * a CUDA graph follows the trainer's per-epoch ExponentialLR decay, train20v2learn_wandb_proj.py:205, 524). */
int i2p_optim_state_bytes(void);
int i2p_optim_lr_offset(void);
int i2p_clip_adam_step(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, void *state,
                       float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm, int world,
                       void *stream);

#ifdef __cplusplus
}
#endif
#endif /* I2P_B200_H */
