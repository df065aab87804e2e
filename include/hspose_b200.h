/*
 * hspose_b200.h — C ABI of libhspose_b200.so (sm_100a only).
 *
 * Drop-in boundary for the HS-Pose hybrid-scope feature extractor.  The
 * reference (Lynne-Zheng-Linfang/HS-Pose) has no FFI of its own: the hot path
 * is the Python module network/fs_net_repo/gcn3d.py.  Every entry point below
 * therefore cites the gcn3d.py function (file:line) whose arithmetic it
 * replaces; INTEGRATION.md shows the ctypes stub a maintainer adds to gcn3d.py.
 *
 * Conventions
 *   - all pointers are DEVICE pointers, contiguous row-major, borrowed;
 *     outputs are caller-allocated; the library never allocates user-visible
 *     memory.  Scratch comes from a caller-provided workspace whose size is
 *     returned by the matching *_workspace_bytes() query.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *     re-entrant, and never synchronises the device.
 *   - return value: 0 (HSP_OK) or a negative HSP_E* code; hsp_strerror()
 *     gives the text.  No exceptions, no printf.
 *   - float = IEEE fp32.  Neighbour indices are int32 inside the library;
 *     int64 copies (the dtype torch.topk returns) are optional outputs.
 */
#ifndef HSPOSE_B200_H_
#define HSPOSE_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSP_OK          0
#define HSP_EINVAL     -1   /* bad argument (null pointer, k > n, unsupported size) */
#define HSP_ELAUNCH    -2   /* kernel launch failed (cudaGetLastError != success)  */
#define HSP_EDEVICE    -3   /* device is not sm_100 (Blackwell B200)               */
#define HSP_EWORKSPACE -4   /* workspace too small / missing                       */

/* Storage types of the dense (M,C) activation matrices (arithmetic is fp32). */
#define HSP_DTYPE_F32   0
#define HSP_DTYPE_BF16  1

/* Distance formulas (evaluation order matters: KNN indices are bit-exact). */
#define HSP_DIST_NEIGHBOR 0 /* ((-2*inner) + q[j]) + q[i]     gcn3d.py:19-21 */
#define HSP_DIST_NEAREST  1 /* (s[j] + t[i]) - (2*inner)      gcn3d.py:31-34 */

int         hsp_version(void);
const char* hsp_strerror(int code);
/* 0 if the current device is compute capability 10.x, else HSP_EDEVICE. */
int         hsp_device_check(void);

/* ------------------------------------------------------------------ K1 ---
 * Fused 3-D pairwise distance + top-k.  Replaces
 *   get_neighbor_index(vertices, k)   gcn3d.py:15-24  (formula NEIGHBOR, drop_first=1)
 *   get_nearest_index(target, source) gcn3d.py:27-36  (formula NEAREST, k=1, drop_first=0)
 * query (B,M,3), cand (B,N,3) (may alias).  For every query the k+drop_first
 * smallest distances are selected (ascending; ties broken by lower index) and
 * the first `drop_first` are discarded, exactly as `topk(k+1)[..., 1:]`.
 * idx64 (B,M,k) and/or idx32 (B,M,k) receive the indices (either may be NULL).
 * Never materialises the MxN matrix.  Requires k+drop_first <= min(N,64),
 * N <= 8192.                                                              */
int hsp_knn3(const float* query, const float* cand, int B, int M, int N, int k,
             int drop_first, int formula, int64_t* idx64, int32_t* idx32,
             void* stream);

/* ------------------------------------------------------------------ K2 ---
 * Fused D-dimensional (feature-space, "RF-F") distance + top-k. Replaces
 *   get_neighbor_index(feature_map, k) gcn3d.py:15-24 with D in {128,256,...}
 * feat (B,N,D), D % 32 == 0.  inner[i,j] is accumulated as ONE sequential FP32
 * FMA chain over d = 0..D-1; |f|^2 as rounded squares added left to right;
 * formula NEIGHBOR.  workspace: B*N floats (row norms).                      */
size_t hsp_knn_feat_workspace_bytes(int B, int N);
int hsp_knn_feat(const float* feat, int B, int N, int D, int k, int drop_first,
                 int64_t* idx64, int32_t* idx32, void* workspace,
                 size_t workspace_bytes, void* stream);

/* get_neighbor_direction_norm (gcn3d.py:49-59) materialised (API parity only;
 * K3/K4 compute it in-kernel): out_norm (B,N,k,3) = F.normalize(nbr - centre),
 * out_raw (B,N,k,3) optional un-normalised differences.                     */
int hsp_neighbor_direction_norm(const float* xyz, const int32_t* idx, int B, int N,
                                int k, float* out_norm, float* out_raw, void* stream);

/* ------------------------------------------------------------------ K3 ---
 * Surface graph-convolution (HSlayer_surface.graph_conv, gcn3d.py:92-107):
 *   out[b,i,c] = 1/S * sum_s max_n relu( rhat[b,i,n] . dirn[:, s*C+c] )
 * rhat = normalize(xyz[idx[b,i,n]] - xyz[b,i])  (gcn3d.py:49-59), computed
 * in-kernel.  xyz (B,N,3); idx (B,N,k) int32; dirn (3,S*C) already
 * column-normalised (F.normalize(directions, dim=0)); out (B,N,C).         */
int hsp_surface_conv_fwd(const float* xyz, const int32_t* idx, const float* dirn,
                         int B, int N, int k, int S, int C, float* out, uint8_t* argmax,
                         void* stream);
/* argmax (B,N,S*C) uint8, optional (NULL when no gradient is needed): the
 * neighbour slot that won max_n, 255 when ReLU zeroed every neighbour.
 * grad wrt dirn: gdirn (3,S*C) (overwritten) replays only the winners
 * (N*S*C work instead of N*k*S*C).  Deterministic two-stage reduction through
 * `workspace`.                                                             */
size_t hsp_surface_conv_bwd_workspace_bytes(int B, int N, int k, int S, int C);
int hsp_surface_conv_bwd(const float* xyz, const int32_t* idx, const uint8_t* argmax,
                         const float* gout, int B, int N, int k, int S, int C,
                         float* gdirn, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ------------------------------------------------------------------ K4 ---
 * HS graph-convolution (HS_layer.graph_conv, gcn3d.py:158-181):
 *   out[b,i,c] = P[b,i,c] + 1/S * sum_s max_n( relu(rhat[b,i,n].dirn[:,s*C+c])
 *                                              * P[b, idx[b,i,n], C + s*C + c] )
 * P (B,N,(S+1)*C) = feature_map @ weights + bias (dense GEMM done by the
 * caller); idx = feature-space neighbours; rhat from xyz of those neighbours.
 * argmax (B,N,S*C) uint8 (optional, may be NULL) receives the winning n for
 * the backward pass.  P is fp32 or bf16 (p_dtype = HSP_DTYPE_*; bf16 halves the
 * gather traffic in the mixed-precision train step); gP is always fp32.      */
int hsp_graph_conv_fwd(const float* xyz, const int32_t* idx, const float* dirn,
                       const void* P, int p_dtype, int B, int N, int k, int S, int C,
                       float* out, uint8_t* argmax, void* stream);
/* Backward: gP (B,N,(S+1)*C) (overwritten: centre part = gout, support part
 * = scatter of gout/S*theta to the winning neighbour rows), gdirn (3,S*C), and
 * optionally gbias ((S+1)*C, may be NULL) = column sums of gP — the gradient of
 * HS_layer.bias (gcn3d.py:170) — accumulated in registers alongside gdirn, so no
 * separate reduction pass over gP is needed.                                 */
size_t hsp_graph_conv_bwd_workspace_bytes(int B, int N, int k, int S, int C);
int hsp_graph_conv_bwd(const float* xyz, const int32_t* idx, const float* dirn,
                       const void* P, int p_dtype, const uint8_t* argmax, const float* gout,
                       int B, int N, int k, int S, int C, float* gP, float* gdirn,
                       float* gbias, void* workspace, size_t workspace_bytes, void* stream);

/* K4b, object-resident variant: one CTA owns the gradient slab of (object, support, 32 channels) in shared memory, so
 * gP is written exactly once, in `gp_dtype` (HSP_DTYPE_F32 / HSP_DTYPE_BF16), without a memset, global float atomics or
 * a cast pass; gdirn / gbias as hsp_graph_conv_bwd (bit-reproducible).  Same reference semantics (gcn3d.py:158-181
 * backward through index_put_(accumulate=True), gcn3d.py:39-47).  Requires S == 7, C % 32 == 0 and a slab that fits
 * shared memory: ask hsp_graph_conv_bwd_obj_supported first (HSP_EINVAL otherwise).  Workspace: the _obj_ query. */
int hsp_graph_conv_bwd_obj_supported(int N, int k, int C);
size_t hsp_graph_conv_bwd_obj_workspace_bytes(int B, int N, int k, int S, int C);
int hsp_graph_conv_bwd_obj(const float* xyz, const int32_t* idx, const float* dirn, const void* P, int p_dtype,
                           const uint8_t* argmax, const float* gout, int B, int N, int k, int S, int C, void* gP,
                           int gp_dtype, float* gdirn, float* gbias, void* workspace, size_t workspace_bytes,
                           void* stream);

/* ------------------------------------------------------------------ K5 ---
 * Row gather + max over neighbours, evaluated at selected rows only:
 *   out[b,r,c] = max_{n<kuse} feat[b, idx[b, rows[r], n], c]
 * rows == NULL means rows[r] = r (R must equal N).  idx has row stride
 * kstride >= kuse (a k=20 neighbour table serves the k=4 pooling).  Replaces
 *   Pool_layer.forward       gcn3d.py:234-246  (rows = randperm sample)
 *   get_ORL_global max part  gcn3d.py:213-216
 * argmax (B,R,C) uint8 optional.                                            */
int hsp_gather_max_fwd(const float* feat, const int32_t* idx, const int32_t* rows,
                       int B, int N, int C, int R, int kuse, int kstride,
                       float* out, uint8_t* argmax, void* stream);
/* gfeat (B,N,C) = scatter-add of gout through the saved arg-max (autograd of index + max).  gfeat is WRITTEN IN
 * FULL (rows nobody selected are zero).  C % 4 == 0, C <= 512: gathered per source row in a fixed order —
 * deterministic, no float atomics; other shapes: memset + float atomics.                                      */
int hsp_gather_max_bwd(const float* gout, const int32_t* idx, const int32_t* rows,
                       const uint8_t* argmax, int B, int N, int C, int R, int kuse,
                       int kstride, float* gfeat, void* stream);

/* ORL global descriptor (get_ORL_global, gcn3d.py:211-218):
 *   G[b,c] = 1/N * sum_i max_n feat[b, idx[b,i,n], c]          G (B,C)
 * Deterministic: per-CTA partial sums go through `workspace`.               */
size_t hsp_orl_global_workspace_bytes(int B, int N, int C);
int hsp_orl_global_fwd(const float* feat, const int32_t* idx, int B, int N, int C,
                       int k, float* G, uint8_t* argmax, void* workspace,
                       size_t workspace_bytes, void* stream);
/* gfeat (B,N,C) += scatter(gG[b,c]/N) to the winning rows.                  */
int hsp_orl_global_bwd(const float* gG, const int32_t* idx, const uint8_t* argmax,
                       int B, int N, int C, int k, float* gfeat, void* stream);

/* Nearest-neighbour up-sampling (FaceRecon.py:100-104):
 *   out[b,i, col0:col0+C] = feat[b, nn[b,i], :]     out row stride = ldo
 * writes straight into the (B,M,ldo) concat buffer.  nn == NULL selects the
 * identity (Nsrc == M: row i) or broadcast (Nsrc == 1: row 0) copy, so the
 * whole torch.cat of FaceRecon.py:107 is served by this one entry point.    */
int hsp_upsample_rows_fwd(const float* feat, const int32_t* nn, int B, int Nsrc,
                          int M, int C, void* out, int ldo, int col0, int out_dtype,
                          void* stream);
/* gfeat (B,Nsrc,C) = sum over i with nn[b,i]==r of gout[b,i,col0:col0+C]  (nn != NULL; written in full, rows
 * without a target are zero; 16-byte aligned C % 8 == 0 slices are gathered per source row in ascending target
 * order — deterministic, no float atomics; other shapes: memset + float atomics), or
 * gfeat = gout[..., col0:col0+C] (nn == NULL, Nsrc == M).  out / gout are fp32 or bf16 (HSP_DTYPE_*): the
 * (B,M,ldo) concat buffer feeds the bf16 tensor-core MLPs directly.                                           */
int hsp_upsample_rows_bwd(const void* gout, const int32_t* nn, int B, int Nsrc,
                          int M, int C, int ldo, int col0, int gout_dtype, float* gfeat,
                          void* stream);

/* ----------------------------------------------------------------- K5d ---
 * The residual sum that closes every HS layer (gcn3d.py:109-113 / :183-187 and
 * the `+ f_STE` of :90 / :156), one pass:
 *   out[b,i,c] = feature[b,i,c] + lin[b,i,c] + gproj[b,c] + ste[b,i,c]
 * lin = feature @ W2[:, :C]^T and ste = STE(layer input) are (B,N,C) fp32 or bf16
 * (HSP_DTYPE_*), gproj = G @ W2[:, C:]^T is (B,C) fp32; lin, gproj, ste may each
 * be NULL.  The surface layer's STE is a 3 -> C linear on the coordinates
 * (HSlayer_surface.STE_layer, gcn3d.py:71,86): pass xyz (B,N,3) and wxyz (C,3)
 * instead of ste and it is evaluated in the same pass.  C % 4 == 0.
 * bwd: g = d out (B,N,C) fp32.  g_bf16 (B,N,C) bf16 (optional) = cast(g) — the
 * gradient of a bf16 lin / ste; g_gproj (B,C) (optional) = sum_i g[b,i,:];
 * g_wxyz_partial (B,C,3) (optional, needs xyz) = sum_i g[b,i,c]*xyz[b,i,d], the
 * per-object part of d wxyz (the caller adds the B slices).  d feature = g
 * needs no kernel.                                                          */
int hsp_residual_sum_fwd(const float* feature, const void* lin, int lin_dtype,
                         const float* gproj, const void* ste, int ste_dtype,
                         const float* xyz, const float* wxyz, int B, int N, int C,
                         float* out, void* stream);
int hsp_residual_sum_bwd(const float* g, const float* xyz, int B, int N, int C, void* g_bf16,
                         float* g_gproj, float* g_wxyz_partial, void* stream);

/* Max over the points of every object (the torch.max(x, 2) in front of the last
 * block of each pose head, PoseR.py:30, PoseTs.py:35) on a (B,N,C) activation in
 * point-major layout: out (B,C) (same dtype as x, HSP_DTYPE_*), arg (B,C) int32 =
 * the winning point (ties -> lowest index).  C % 8 == 0.                      */
int hsp_colmax_fwd(const void* x, int dtype, int B, int N, int C, void* out, int32_t* arg,
                   void* stream);

/* ------------------------------------------------------------------ K7 ---
 * Chamfer distance (tools/pyTorchChamferDistance/chamfer_distance.cu:6-187):
 * for each point of a (B,N,3) the squared distance to / index of the nearest
 * point of b (B,M,3) and vice versa.  Direct (x-y)^2 sums as in the reference
 * kernel.                                                                   */
int hsp_chamfer_fwd(const float* a, const float* b, int B, int N, int M,
                    float* dist_a, int32_t* idx_a, float* dist_b, int32_t* idx_b,
                    void* stream);
/* ga (B,N,3), gb (B,M,3) overwritten.  The scatter to the matched point uses
 * float atomics, like the reference's ChamferDistanceGradKernel (:158-187).  */
int hsp_chamfer_bwd(const float* a, const float* b, const int32_t* idx_a,
                    const int32_t* idx_b, const float* gdist_a, const float* gdist_b,
                    int B, int N, int M, float* ga, float* gb, void* stream);

/* ----------------------------------------------------------------- K6b ---
 * Batch-statistics BatchNorm1d (+ optional ReLU) over the channel axis of a
 * row-major (M,C) activation matrix: the BatchNorm1d -> ReLU pairs of the dense
 * per-point MLPs (FaceRecon.py:38-68,89-95; PoseR.py:22-29; PoseTs.py:24-34) in
 * training mode.  x / y / dy / dx may be column slices of wider matrices (row
 * strides ld*, in elements; 16-byte aligned; C % 8 == 0).  dtype = HSP_DTYPE_*.
 *   fwd: mean, invstd (C each) and scale_shift (2C) are outputs kept for the
 *        backward; running_mean / running_var (may be NULL) are updated with
 *        `momentum` (unbiased variance), as nn.BatchNorm1d does.
 *   bwd: dgamma, dbeta (C each), dx.  The ReLU mask is recomputed from x.  dx_colsum
 *        (C, optional) receives sum_rows dx as stored — the bias gradient of the Linear /
 *        Conv1d(k=1) that produced x, so that layer needs no separate reduction pass.
 * Deterministic (fixed-order partial sums through `workspace`).             */
size_t hsp_bn_workspace_bytes(int M, int C);
int hsp_bn_relu_fwd(const void* x, int ldx, int M, int C, int dtype, const float* gamma,
                    const float* beta, float eps, float momentum, int relu,
                    float* running_mean, float* running_var, float* mean, float* invstd,
                    float* scale_shift, void* y, int ldy, void* workspace,
                    size_t workspace_bytes, void* stream);
int hsp_bn_relu_bwd(const void* x, int ldx, const void* dy, int lddy, int M, int C, int dtype,
                    const float* gamma, const float* beta, const float* mean,
                    const float* invstd, int relu, float* dgamma, float* dbeta, void* dx,
                    int lddx, float* dx_colsum, void* workspace, size_t workspace_bytes,
                    void* stream);

/* Forward with the statistics partials already produced by hsp_gemm_bf16's epilogue: `partials`
 * is (nblocks, 2, C) column sums / sums of squares over disjoint row blocks covering the M rows,
 * row pitch ldp >= C floats (a column slice of a wider partial matrix).                           */
int hsp_bn_apply_fwd(const void* x, int ldx, int M, int C, int dtype, const float* partials,
                     int nblocks, int ldp, const float* gamma, const float* beta, float eps, float momentum,
                     int relu, float* running_mean, float* running_var, float* mean, float* invstd,
                     float* scale_shift, void* y, int ldy, void* stream);

/* ------------------------------------------------------------------ K6 ---
 * Dense per-point MLP contraction on the tensor cores (tcgen05.mma, TMA tensor loads,
 * TMEM accumulators):   out[M,N] (+ bias[N]) = A[M,K] . B[N,K]^T,  bf16 operands, fp32
 * accumulation, bf16 or fp32 output.  Replaces the library GEMM behind every
 * Conv1d(kernel_size=1) / `@` of the reference's dense stages, forward and backward:
 *   heads conv1..2           PoseR.py:27-34, PoseTs.py:32-39
 *   conv1d_block / recon / face stacks   FaceRecon.py:38-68,114-124
 *   feature_map @ weights + bias         gcn3d.py:171
 *   STE_layer / conv2 (1x1)              gcn3d.py:85,112,149,186
 * Operand majors: *_mn_major = 0 -> the matrix is stored (rows = M|N, cols = K) with leading
 * dimension ld (K contiguous: activations as A, nn.Linear weights (out,in) as B);
 * = 1 -> stored (rows = K, cols = M|N) (the M|N index contiguous: the transposed operands of
 * dgrad / wgrad, and gcn3d's (in,out) `weights`), so nothing is transposed in memory.
 * ld* in elements; pointers and row pitches 16-byte aligned.
 *   splits > 1 (fp32 output only): split-K; plane s of `out` (M*ldo elements each) receives the
 *     partial product of k-slice s — the caller adds the planes in order (deterministic).
 *     hsp_gemm_bf16_splits() returns the factor the library would choose.
 *   stats (optional, bf16 output): (ceil(M/128), 2, N) floats — per 128-row block the column
 *     sums and sums of squares of the values AS STORED; this is the partial layout
 *     hsp_bn_apply_fwd() consumes, so BatchNorm needs no statistics pass of its own.
 *   bias_rows (optional, (ceil(M/rows_per_group), N) floats, N % 8 == 0): a second bias shared by groups of
 *     rows_per_group consecutive rows — the per-object `f_global` block of face_head[0]
 *     (FaceRecon.py:118-121: cat[f_global repeated over the points, ...]) as a broadcast term.
 *   relu != 0: max(0, .) after the biases (the ReLU of a folded Conv1d -> BatchNorm(eval) -> ReLU block).
 *   tile_n in {64,128,256} and ctas in {1,2} pick the tile (0 = library default).            */
int hsp_gemm_bf16_splits(int M, int N, int K, int out_f32);
/* Diagnostics for profiling (results are WRONG while non-zero): bit 0 skips the epilogue's staging
 * writes and stores, bit 1 the MMAs, bit 2 the TMA loads.  Returns the previous value.          */
int hsp_gemm_debug(int flags);
int hsp_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major,
                  int M, int N, int K, const float* bias, const float* bias_rows, int rows_per_group,
                  int relu, void* out, int ldo, int out_f32, int splits, float* stats, int tile_n,
                  int ctas, void* stream);
/* Same, with a residual: out = c_in + A . B^T (+ biases), c_in (M, ldc) fp32 or NULL (fp32 output, splits == 1,
 * N % 4 == 0, 16-byte aligned rows).  c_in may not alias out.  A dgrad GEMM whose result joins a gradient that
 * already exists (the pass-through term of a residual connection) needs no separate element-wise add.        */
int hsp_gemm_bf16_acc(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major,
                      int M, int N, int K, const float* bias, const float* bias_rows, int rows_per_group,
                      int relu, const float* c_in, int ldc, void* out, int ldo, int out_f32, int splits,
                      float* stats, int tile_n, int ctas, void* stream);

/* ------------------------------------------------------------------ K8 ---
 * The 19-term loss graph of training stage 'PoseNet_only' (L1 loss type), forward and backward:
 *   losses/fs_net_loss.py:31-76 (Rot1, Rot1_cos, Rot2, Rot2_cos, Rot_r_a, Tran, Size, R_con)
 *   losses/recon_loss.py:464-649 (recon_per_p, recon_p_f, recon_point_vote/_r/_t/_s/_self)
 *   losses/geometry_loss.py:123-150 (geo_point), losses/prop_loss.py:156-277 (Prop_pm, Prop_sym_recon/_rt)
 * incl. the face normalisation / sigmoid of PoseNet9D.py:28-33 and tools/plane_utils.py:24-48.
 *   face (B,N,30) raw face-head output, recon (B,N,3), PC (B,N,3);
 *   pred (B,14) = [p_green 3 | p_red 3 | f_green | f_red | Pred_T 3 | Pred_s 3];
 *   gt   (B,23) = [gt_R 9 row-major | gt_t 3 | gt_s 3 | mean_shape 3 | sym 4 | obj_id];
 *   weights: HOST array of 17 floats (rot_1_w, rot_2_w, rot_regular, tran_w, size_w, r_con_w, recon_n_w,
 *            recon_d_w, recon_f_w, recon_v_w, recon_bb_r_w, recon_bb_t_w, recon_bb_s_w, recon_bb_self_w,
 *            geo_p_w, prop_pm_w, prop_sym_w of config/config.py).
 * fwd: sums (B, hsp_losses_num_sums()) per-object sums kept for the backward; pieces (B, 19) per-object
 *      contributions — term k = sum_b pieces[b,k] (order: the list above).
 * bwd: gterm (19) upstream gradient per term -> gface (B,N,30), grecon (B,N,3), gpred (B,14);
 *      gmom_ws: B*54 floats of scratch.  Deterministic.                                          */
int hsp_losses_num_terms(void);
int hsp_losses_num_sums(void);
int hsp_losses_fwd(const float* face, const float* recon, const float* PC, const float* pred,
                   const float* gt, const float* weights, int B, int N, float* sums, float* pieces,
                   void* stream);
int hsp_losses_bwd(const float* face, const float* recon, const float* PC, const float* pred,
                   const float* gt, const float* weights, const float* sums, const float* gterm,
                   int B, int N, float* gface, float* grecon, float* gpred, float* gmom_ws,
                   void* stream);

/* ------------------------------------------------------------------ K9 ---
 * clip_grad_norm_ + optimiser step over flat fp32 buffers of n elements, three launches.
 *   kind 1 = Ranger (RAdam + Lookahead + gradient centralisation): tools/torch_utils/solver/ranger2020.py:135-246
 *            (the reference's optimiser, tools/solver_utils.py:49-50), gc_loc = True, gc_conv_only = False;
 *   kind 0 = Adam (torch.optim.Adam semantics; BASELINE.json configs[2] names Adam).
 * Clipping as engine/train.py:107 (`clip_grad_norm_(.., 5)`): coef = min(1, clip_max_norm / (||g|| + 1e-6)),
 * folded into the update (grad is not modified); clip_max_norm <= 0 disables it.
 * Segments describe the flat buffer: seg_len > 0 = one row of a >= 2-D parameter (its mean is subtracted
 * from the gradient: gradient centralisation), seg_len < 0 = |seg_len| elements without centralisation.
 * `step` (device int32) is incremented on the device and `lr` is read from device memory, so the call can
 * live inside a CUDA graph.  slow: Lookahead weights (Ranger; initialised by the caller to the parameters).
 * grad_norm_out (device, optional) receives ||g|| before clipping.  Deterministic.                     */
size_t hsp_optim_workspace_bytes(void);
int hsp_optim_step(int kind, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                   float* slow, long n, const int* seg_off, const int* seg_len, int nseg, int* step,
                   const float* lr, float beta1, float beta2, float eps, float weight_decay,
                   float clip_max_norm, float alpha, int k, int nsma_threshold, float* grad_norm_out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ K10 --
 * HSPose.data_augment (network/HSPose.py:185-256 over datasets/data_augmentation.py:70-79,106-127,134-140,
 * 183-190) as one launch: Bernoulli-gated box scaling, rigid perturbation, bowl/mug taper, point noise.
 * Random inputs (uniform [0,1), drawn by the caller in the reference's order): gates (B,4) = [bb, rt, bc, pc],
 * ey (B,2) = [up, down], defor (B,N,3).  p_* = FLAGS.aug_*_pro, pc_r = FLAGS.aug_pc_r.
 * R (B,3,3) row-major, model_point (B,Nm,3).  Outputs: PC_out (B,N,3), R_out, t_out, s_out (B,3).      */
int hsp_augment(const float* PC, const float* R, const float* t, const float* s, const float* mean_shape,
                const float* sym, const float* aug_bb, const float* aug_rt_t, const float* aug_rt_r,
                const float* model_point, const float* nocs_scale, const float* obj_id,
                const float* gates, const float* ey, const float* defor, float p_bb, float p_rt,
                float p_bc, float p_pc, float pc_r, int B, int N, int Nm, float* PC_out, float* R_out,
                float* t_out, float* s_out, void* stream);

/* Unit support directions: F.normalize(directions, dim=0) (gcn3d.py:95, :162) and its backward, one launch
 * each.  d, out, g, gd (3, n) row-major fp32; nrm (n) = ||d[:, j]|| kept for the backward; eps = 1e-12.    */
int hsp_normalize_cols_fwd(const float* d, int n, float eps, float* out, float* nrm, void* stream);
int hsp_normalize_cols_bwd(const float* g, const float* out, const float* nrm, int n, float eps, float* gd,
                           void* stream);

/* ------------------------------------------------------------------ K11 --
 * Input pre-stage: depth ROI -> camera-frame point cloud -> n_pts sampled points.
 * hsp_depth_to_cloud replaces datasets/load_data.py:322-333 `_depth_to_pcl` followed by `/ 1000.0` (:277)
 *   [camK_is_f64 = 1: numpy's float64 arithmetic, camK (B,3,3) double] and the back-projection of
 *   network/point_sample/pc_sample.py:24-54 `PC_sample` [camK_is_f64 = 0: torch float32 arithmetic, camK float].
 *   depth, mask (B,H,W) fp32; xymap (B,2,H,W) fp32 (x map, y map); a pixel is kept when depth > 0 and mask > 0.
 *   cloud (B,H*W,3): the kept pixels in raster order, metres ((x - cx) * d / fx, (y - cy) * d / fy, d) / 1000;
 *   rows >= count[b] are not written.  count (B) int32.
 * hsp_sample_points replaces load_data.py:307-320 `_sample_points` / pc_sample.py:57-75:
 *   choose (B,n_pts) int32 != NULL: out[b,i] = cloud[b, choose[b,i]] (the caller's numpy draw: exact parity);
 *   choose == NULL: count <= n_pts -> out[b,i] = cloud[b, i mod count] (the reference's tile rule);
 *                   count >  n_pts -> a uniformly random subset without replacement (counter-hash keys from
 *                   `seed`, n_pts smallest), in raster order.
 *   status (device int32, optional, OR-ed): bit 0 = an object without valid pixels (its rows are zeros),
 *   bit 1 = a choose index outside [0, count) (clamped).                                                  */
int hsp_depth_to_cloud(const float* depth, const float* mask, const float* xymap, const void* camK,
                       int camK_is_f64, int B, int H, int W, float* cloud, int* count, void* stream);
int hsp_sample_points(const float* cloud, const int* count, const int* choose, unsigned long long seed,
                      int B, int cap, int n_pts, float* out, int* status, void* stream);

/* fp32 -> bf16 multi-term split feeding hsp_gemm_bf16 for fp32-accurate contractions (the fp32 evaluation
 * forward, BASELINE configs[1]): out (M, nterms*Kpad) bf16, term t = component comp[t] (0: bf16(x),
 * 1: bf16(x - x1), 2: bf16(x - x1 - x2)) of x (M,K | ld), zero-padded to Kpad (multiple of 8) columns.
 * comp is a HOST array.  With A' = split(x,[0,1,2,0,1,0]) and B' = split(w,[2,1,0,1,0,0]) (smallest cross
 * terms first, so they are accumulated before the large one) the product A'.B'^T equals x.w^T to fp32
 * accuracy (three dropped terms <= 2^-24 relative).                                                     */
int hsp_split_bf16(const float* x, int ld, int M, int K, int Kpad, int nterms, const int* comp, void* out,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HSPOSE_B200_H_ */
