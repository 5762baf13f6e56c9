/* vistracker_b200 -- C ABI of libvistracker_sm100a.so
 *
 * Drop-in boundary for the VisTracker per-frame hot path on NVIDIA B200 (sm_100a).  The reference (xiexh20/VisTracker)
 * is pure Python over PyTorch; the operators below are what its Python call sites would bind (ctypes stub in
 * INTEGRATION.md).  Each entry names the reference interface it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless it says "host"; fp32 unless stated; tensors are NHWC, channel stride
 *     ("ld*", in elements) given explicitly so channel slices of a wider tensor can be read / written in place;
 *   - the library never allocates, frees or synchronises device memory: the caller (PyTorch) owns inputs, outputs and
 *     scratch; kernels are enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - return value 0 = enqueued, < 0 = rejected (nothing enqueued); vt_last_error() gives the reason.  No exceptions cross
 *     the ABI.  Re-entrant; no global mutable state besides the per-thread error string;
 *   - `stats` arguments: double[n_img][ld_stats][2] per-channel (sum, sum of squares) of the tensor the call WRITES,
 *     accumulated with atomics (caller zeroes the slot first).  vt_gn_finalize turns them into the GroupNorm affine.
 */
#ifndef VISTRACKER_B200_H
#define VISTRACKER_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ---------------------------------------------------------------------------------------------------- */
const char* vt_last_error(void);          /* host string, valid until the next failing call on this thread */
int vt_version(void);                     /* ABI version, currently 1 */
int vt_compiled_arch(void);               /* 100 -> built with -gencode arch=compute_100a,code=sm_100a */

/* ---- one-time re-packing of a reference checkpoint (HOST pointers in and out; csrc/pack.cu) -- SURVEY.md 8(b) `vt_pack_weights_<op>` ----
 * The reference keeps `CHORETriplaneVisibility.state_dict()` tensors in torch layouts (model/HGFilters.py: Conv2d [Cout][Cin][k][k];
 * model/chore.py:113-126: Conv1d(k=1) decoders [out][in][1]) and the SMPL-H buffers of lib_smpl/smplpytorch/smplpytorch/pytorch/smpl_layer.py:30-71;
 * these functions write the layouts the kernels below read.  The caller allocates the outputs and uploads them. */

/* input channels of the fp16 planes, padded to the K chunk of the tensor-core convolution (64) */
int vt_conv_cin_pad(int cin);
/* Conv2d weight w[Cout][Cin][ks][ks] -> ffma fp32 [ks*ks][Cin][Cout] (vt_conv_ffma) and fp16 planes hi / lo [ks*ks][Cout][cin_pad] with
 * w ~= hi + lo / 2048 (vt_conv_mma).  ffma or the plane pair may be NULL.  -4: a weight exceeds the fp16 range. */
int vt_pack_weights_conv(const float* w, int cout, int cin, int ks, float* ffma, void* hi, void* lo);
/* Conv2d(cin, cout, 7, stride 2) weight w[Cout][Cin][7][7] -> fp32 [49*cin][cout] (vt_stem_conv7x7s2) */
int vt_pack_weights_stem(const float* w, int cout, int cin, float* out);
/* The five decoders: w[h*4 + l] / b[h*4 + l] = Conv1d(k=1) weight [out][in] / bias [out] of head h (df, pca_predictor, part_predictor,
 * center_predictor, visib_predictor) layer l (in = 611, 128, 128, 128).  wpack: vt_query_wpack_floats() floats (vt_query_fwd and the fp32
 * part of the tensor-core kernels), wpack_bwd: vt_query_wpack_bwd_floats() floats (vt_query_bwd); either may be NULL. */
int vt_pack_weights_decoders(const float* const* w, const float* const* b, float* wpack, float* wpack_bwd);
/* fp16 hi / lo planes of the tcgen05 decoder kernels (x ~= hi + lo): w1 [5*128][640], w23 [2*5*128][128] (vt_query_fwd_tc) and the
 * transposes w23t [2*5*128][128], w1t [5*640][128] (vt_query_bwd_tc, vt_query_project_step_tc, vt_query_losses_*_tc; may be NULL together). */
int vt_pack_weights_decoders_tc(const float* const* w, void* w1_hi, void* w1_lo, void* w23_hi, void* w23_lo, void* w23t_hi, void* w23t_lo,
                                void* w1t_hi, void* w1t_lo);
/* SMPL-H model buffers -> the arrays of SmplModel below.  Sizes: kd = 9 (J - 1) + n_betas, kdp / nv3p = kd / 3 V rounded up to 4,
 * nnz = most non-zero skinning weights of a vertex.  v_template [V][3], shapedirs [V][3][n_betas], posedirs [V][3][9 (J - 1)],
 * J_regressor [J][V] in float64 (as the model pickles hold them), weights [V][J] fp32, parents [J] (root: any value <= 0).
 * Outputs: templ [3 V], dirs [kdp][nv3p], dirsT [nv3p][kdp], j_templ [J][3], j_dirs [J][3][n_betas], parents_out [J], skin_idx / skin_w [V][nnz]. */
int vt_smpl_pack_dims(int V, int J, int n_betas, const float* weights, int* kd, int* kdp, int* nv3p, int* nnz);
int vt_pack_weights_smpl(const double* v_template, const double* shapedirs, const double* posedirs, const double* J_regressor,
                         const float* weights, const int* parents, int V, int J, int n_betas, float* templ, float* dirs, float* dirsT,
                         float* j_templ, float* j_dirs, int* parents_out, int* skin_idx, float* skin_w);
/* Caller-owned workspaces (SURVEY.md 8(b) `vt_workspace_bytes_<op>`): only vt_raster_* (optional culling boxes) and vt_procrustes take
 * one; every other entry point works in the buffers named in its signature and needs no scratch memory. */
long long vt_workspace_bytes_raster_cull(int B, int F);
long long vt_workspace_bytes_procrustes(int B);

/* ---- stacked-hourglass encoder pieces: model/HGFilters.py:162-203 (HGFilter.forward), :26-50 (HourGlass._forward),
 *      model/net_util.py:374-396 (ConvBlock.forward) ------------------------------------------------------------------ */

/* conv1 of HGFilter: nn.Conv2d(cin, cout, 7, stride 2, padding 3, bias) (HGFilters.py:120,167) read directly from the
 * NCHW frame tensor images[B][Ctot][Hin][Win].  Encoder image n = view*B + b uses channels c_off + view*cin + [0,cin).
 * w: [49*cin][cout] (tap-major, then input channel), out: NHWC [B*n_views][Hin/2][Win/2][cout]. */
int vt_stem_conv7x7s2(const float* images, int B, int Ctot, int Hin, int Win, int c_off, int cin, int n_views, const float* w,
                      const float* bias, int cout, float* out, double* stats, int ld_stats, void* stream);

/* nn.GroupNorm(groups, C) statistics -> per-(image, channel) scale/shift so that GN(x) = x*scale + shift
 * (net_util.py:358-362; eps inside the sqrt, biased variance).  count_per_channel = H*W. */
int vt_gn_finalize(const double* stats, int ld_stats, const float* gamma, const float* beta, int n_img, int C, int groups,
                   long long count_per_channel, float eps, float* scale, float* shift, void* stream);

/* out = relu?(x*scale + shift) in fp32 (F.relu(self.bn1(self.conv1(x)), True), HGFilters.py:167).  scale may be NULL. */
int vt_affine_act(const float* x, int ldx, const float* scale, const float* shift, int relu, int n_img, int HW, int C,
                  float* out, int ldo, double* stats, int ld_stats, void* stream);

/* GroupNorm-apply + ReLU + fp16 (hi, lo*2^11) split into zero-bordered planes [n][H+2pad][W+2pad][Cpad] (fp16) that
 * vt_conv_mma consumes.  *overflow is incremented if any |value| exceeded the fp16 range (the caller must then fail). */
int vt_prep_split(const float* x, int ldx, const float* scale, const float* shift, int relu, int n_img, int H, int W, int C,
                  int Cpad, int pad, void* hi, void* lo, int* overflow, void* stream);

/* vt_gn_finalize + vt_prep_split in one launch: the GroupNorm(groups, C) affine (net_util.py:358-362, 376-388: F.relu(bnK(x)) feeding
 * convK) is derived inside the kernel from the producer-accumulated statistics (same fp64 arithmetic as vt_gn_finalize -> identical
 * planes), so a normalised conv operand costs one pass and no scale/shift buffers.  C <= 1024. */
int vt_prep_split_gn(const float* x, int ldx, const double* stats, int ld_stats, const float* gamma, const float* beta, int groups,
                     long long count_per_channel, float eps, int relu, int n_img, int H, int W, int C, int Cpad, int pad, void* hi,
                     void* lo, int* overflow, void* stream);

/* conv3x3 (padding 1, net_util.py:213-216) / 1x1 conv on the tcgen05 tensor cores.  a_hi/a_lo from vt_prep_split,
 * w_hi/w_lo: fp16 planes [ks*ks][Cout][Cin_pad].  out[pix][0..Cout) = conv + bias + res[pix][0..Cout); res may alias out.
 * Needs W in {8,16,32,64} or a multiple of 128, and H a multiple of 128/min(W,128). */
int vt_conv_mma(const void* a_hi, const void* a_lo, int n_img, int H, int W, int Cin_pad, int pad, int ks, const void* w_hi,
                const void* w_lo, int Cout, const float* bias, const float* res, int ldr, float* out, int ldo, double* stats,
                int ld_stats, void* stream);

/* vt_conv_mma plus a second output written by the same epilogue: out2 = out + res2 with its own statistics.  Used by ConvBlocks with
 * an identity residual: `out` keeps the raw conv slice for the next GroupNorm, `out2` is the block output slice
 * (torch.cat((out1, out2, out3), 1) + residual, net_util.py:389-394) -- no separate add pass over HBM. */
int vt_conv_mma_dual(const void* a_hi, const void* a_lo, int n_img, int H, int W, int Cin_pad, int pad, int ks, const void* w_hi,
                     const void* w_lo, int Cout, const float* bias, const float* res, int ldr, float* out, int ldo, double* stats,
                     int ld_stats, float* out2, int ldo2, const float* res2, int ldr2, double* stats2, int ld_stats2, void* stream);

/* Same contract on the fp32 CUDA cores with the GroupNorm affine + ReLU fused into the load: any H, W; w: fp32
 * [ks*ks][Cin][Cout].  Used where the 128-pixel tensor-core tile does not fit, and as the on-device cross-check. */
int vt_conv_ffma(const float* x, int ldx, const float* scale, const float* shift, int relu, int n_img, int H, int W, int Cin,
                 int ks, const float* w, int Cout, const float* bias, const float* res, int ldr, float* out, int ldo,
                 double* stats, int ld_stats, void* stream);

/* out = a + b (ConvBlock residual, net_util.py:391-394). */
int vt_add(const float* a, int lda, const float* b, int ldb, int n_img, int HW, int C, float* out, int ldo, double* stats,
           int ld_stats, void* stream);

/* F.avg_pool2d(x, 2, stride=2) (HGFilters.py:32,170). */
int vt_avgpool2(const float* x, int n_img, int H, int W, int C, float* out, double* stats, int ld_stats, void* stream);

/* out = up1 + F.interpolate(low, scale_factor=2, mode='bicubic', align_corners=True) (HGFilters.py:47-50). */
int vt_upsample2x_add(const float* low, const float* up1, int n_img, int Hl, int Wl, int C, float* out, double* stats,
                      int ld_stats, void* stream);

/* ---- point query: CHORETriplane.query / query_features + CHORETriplaneVisibility.decode
 *      (model/chore_triplane.py:97-205, model/chore_tri_vis.py:31-50, model/geometry.py:4-14, model/camera.py:45-89) ------ */

/* number of floats of the packed decoder weights (layout in csrc/query.cu, packer in vistracker_b200/weights.py) */
long long vt_query_wpack_floats(void);

/* points[B][N][3], crop_center[B][2], body_center[B][3]; maps NHWC: im_feat[B][Hf][Wf][c_im], tmpx[B][Ht][Wt][c_tmpx],
 * tri_tmpx[3][B][Ht][Wt][c_tt], tri_feat[3][B][Hf][Wf][c_tf] (views right, back, top);
 * cam7 (host) = {fx_px, fy_px, cx_px, cy_px, crop_size, z0, out_dist}.
 * out[B][29][N] = df 2 | pca 9 | parts 14 | centers 3 | visibility 1 (sigmoid applied, df = out_dist outside the image);
 * feat_out[B][611][N] (reference channel order) and xy_out[B][2][N] are optional (NULL to skip). */
int vt_query_fwd(const float* points, const float* crop_center, const float* body_center, int B, int N,
                 const float* im_feat, const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht,
                 int Wt, int c_im, int c_tmpx, int c_tt, int c_tf, const float* cam7, const float* wpack, float* out,
                 float* feat_out, float* xy_out, void* stream);

/* vt_query_fwd with the five decoder MLPs on the tcgen05 tensor cores (fp16 hi/lo split, fp32 TMEM accumulators; 128 points per
 * CTA).  w1_hi/lo: fp16 [5*128][640], w23_hi/lo: fp16 [2*5*128][128] (vistracker_b200/weights.py: pack_decoders_tc); `wpack` is the
 * fp32 pack of vt_query_fwd (biases and the last layer are read from it).  *overflow counts values beyond the fp16 range.
 * Channel counts are fixed to the tri-vis layout (256 / 64 / 32 / 64). */
int vt_query_fwd_tc(const float* points, const float* crop_center, const float* body_center, int B, int N, const float* im_feat,
                    const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht, int Wt, const float* cam7,
                    const float* wpack, const void* w1_hi, const void* w1_lo, const void* w23_hi, const void* w23_lo, float* out,
                    float* xy_out, int* overflow, void* stream);

/* vt_query_fwd_tc restricted to the heads in head_mask (bit 0 df, 1 pca, 2 parts, 3 centers, 4 visibility): rows of `out` that belong
 * to other heads are left untouched.  One feature-gather pass serves up to four heads, so e.g. {df, centers} costs ~1/3 of all five. */
int vt_query_fwd_tc_heads(const float* points, const float* crop_center, const float* body_center, int B, int N, const float* im_feat,
                          const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht, int Wt, const float* cam7,
                          const float* wpack, const void* w1_hi, const void* w1_lo, const void* w23_hi, const void* w23_lo, int head_mask,
                          float* out, float* xy_out, int* overflow, void* stream);

/* Gradient of sum(g_out * out) w.r.t. the points -- what autograd computes for `df.sum().backward()` in
 * Generator.approx_surface (recon/gen/generator.py:86-96) and for the df / part / centre losses of the fitters
 * (recon/recon_fit_behave.py:467-513, recon/recon_fit_trivis_full.py:193-270).  The forward is recomputed on chip.
 * g_out[B][29][N] (same packing as `out`), g_points[B][N][3].  wpack_bwd: vt_query_wpack_bwd_floats() floats. */
long long vt_query_wpack_bwd_floats(void);
int vt_query_bwd(const float* points, const float* crop_center, const float* body_center, int B, int N,
                 const float* im_feat, const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht,
                 int Wt, int c_im, int c_tmpx, int c_tt, int c_tf, const float* cam7, const float* wpack, const float* wpack_bwd,
                 const float* g_out, float* g_points, void* stream);

/* vt_query_bwd restricted to the heads whose bit is set in head_mask (bit 0 df, 1 pca, 2 parts, 3 centers, 4 visibility): heads
 * with an identically-zero cotangent cost nothing (the fitters differentiate through one or two heads only). */
int vt_query_bwd_heads(const float* points, const float* crop_center, const float* body_center, int B, int N,
                       const float* im_feat, const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht,
                       int Wt, int c_im, int c_tmpx, int c_tt, int c_tf, const float* cam7, const float* wpack, const float* wpack_bwd,
                       const float* g_out, int head_mask, float* g_points, void* stream);

/* One step of Generator.approx_surface (recon/gen/generator.py:72-104) fused into one launch: predictions at `points`
 * (written to out[B][29][N] when not NULL), d sum(clamp(df[df_idx], max=threshold)) / d points (g_points, optional) and
 * points_out = points - normalize(grad, eps 1e-12) * clamp(df[df_idx], max=threshold). */
int vt_query_project_step(const float* points, const float* crop_center, const float* body_center, int B, int N,
                          const float* im_feat, const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht,
                          int Wt, int c_im, int c_tmpx, int c_tt, int c_tf, const float* cam7, const float* wpack, const float* wpack_bwd,
                          int df_idx, float threshold, float* points_out, float* out, float* g_points, void* stream);

/* vt_query_bwd_heads / vt_query_project_step with the decoder MLPs on the tcgen05 tensor cores (csrc/query_bwd_tc.cu): 128 points per
 * CTA, forward recompute + analytic backward + second tap gather in one launch.  w1/w23 planes as for vt_query_fwd_tc; w23t_hi/lo:
 * fp16 [2*5*128][128] = W2^T | W3^T, w1t_hi/lo: fp16 [5*640][128] = W1^T in the kernel's feature order
 * (vistracker_b200/weights.py: pack_decoders_tc_bwd).  g_out[B][29][N], g_points[B][N][3]; *overflow counts fp16 range overflows.
 * The projection step does not return the predictions (call vt_query_fwd_tc at the same points for those). */
int vt_query_bwd_tc(const float* points, const float* crop_center, const float* body_center, int B, int N, const float* im_feat,
                    const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht, int Wt, const float* cam7,
                    const float* wpack, const void* w1_hi, const void* w1_lo, const void* w23_hi, const void* w23_lo, const void* w23t_hi,
                    const void* w23t_lo, const void* w1t_hi, const void* w1t_lo, const float* g_out, int head_mask, float* g_points,
                    int* overflow, void* stream);
int vt_query_project_step_tc(const float* points, const float* crop_center, const float* body_center, int B, int N, const float* im_feat,
                             const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht, int Wt,
                             const float* cam7, const float* wpack, const void* w1_hi, const void* w1_lo, const void* w23_hi,
                             const void* w23_lo, const void* w23t_hi, const void* w23t_lo, const void* w1t_hi, const void* w1t_lo,
                             int df_idx, float threshold, float* points_out, float* g_points, int* overflow, void* stream);

/* The two query-dependent loss terms of the fitters, fused: values AND point gradients in ONE launch (no separate forward, no [B,29,N]
 * round trip).  vals_df[B][N] = clamp(df[df_idx], max=clamp_max) -- `torch.clamp(df_pred[:, 0:1], max=0.1)` of forward_smpl
 * (recon/recon_fit_behave.py:471) and `torch.clamp(df_pred[:, 1], max=0.8)` of forward_step (recon/recon_fit_trivis_full.py:235) --
 * with g_df[B][N][3] = d vals_df / d point; when part_labels[B][N] (int64) is given, vals_ce[B][N] = F.cross_entropy(parts, labels,
 * reduction='none') (recon_fit_behave.py:476) with g_ce[B][N][3] = d vals_ce / d point.  Reductions and loss weights are linear in these
 * and stay with the caller.  fwd_mask: further heads to evaluate forward-only in the same launch (they share the feature gather), written
 * into the packed prediction buffer out_fwd[B][29][N] (e.g. bit 3: the centre head forward_step reports, recon_fit_behave.py:370-380). */
int vt_query_losses_tc(const float* points, const float* crop_center, const float* body_center, int B, int N, const float* im_feat,
                       const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht, int Wt, const float* cam7,
                       const float* wpack, const void* w1_hi, const void* w1_lo, const void* w23_hi, const void* w23_lo, const void* w23t_hi,
                       const void* w23t_lo, const void* w1t_hi, const void* w1t_lo, int df_idx, float clamp_max, const long long* part_labels,
                       float* vals_df, float* g_df, float* vals_ce, float* g_ce, int fwd_mask, float* out_fwd, int* overflow, void* stream);

/* vt_query_losses_tc with the two heads MERGED (the SMPL refinement step, recon/recon_fit_behave.py:471-476): the caller's loss weights
 * w_df = *w_df * w_df_mul and w_ce = *w_ce * w_ce_mul (device scalars times host factors, e.g. the schedule word and 1 / (B N)) are folded into
 * the cotangents, both heads' hidden-layer gradients stay in tensor memory and accumulate into one feature-gradient tile, and the second
 * gather (the contraction with d feature / d point) runs once: g_points[B][N][3] = w_df d clamp(df) / d point + w_ce d CE / d point.
 * vals_df / vals_ce as in vt_query_losses_tc. */
int vt_query_losses_merged_tc(const float* points, const float* crop_center, const float* body_center, int B, int N, const float* im_feat,
                              const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht, int Wt, const float* cam7,
                              const float* wpack, const void* w1_hi, const void* w1_lo, const void* w23_hi, const void* w23_lo, const void* w23t_hi,
                              const void* w23t_lo, const void* w1t_hi, const void* w1t_lo, int df_idx, float clamp_max, const long long* part_labels,
                              const float* w_df, float w_df_mul, const float* w_ce, float w_ce_mul, float* vals_df, float* vals_ce, float* g_points,
                              int* overflow, void* stream);

/* ---- SMPL-H layer: SMPL_Layer.forward (lib_smpl/smplpytorch/smplpytorch/pytorch/smpl_layer.py:73-176) and its gradient
 *      w.r.t. pose / betas / trans; landmark regressors (lib_smpl/torch_functions.py:52-76, wrapper_pytorch.py:187-203) ---- */

/* Body model re-packed for the kernels (device pointers; built once by vistracker_b200/smpl.py from the th_* buffers). */
typedef struct vt_smpl_model {
  int V, J, n_betas;        /* 6890, 52, 10 for SMPL-H */
  int kd, kdp;              /* 9*(J-1) + n_betas blend coefficients, padded to a multiple of 4 */
  int nv3p;                 /* 3*V padded to a multiple of 4 */
  int nnz;                  /* skinning weights per vertex (ELL width; 4 in SMPL) */
  const float* templ;       /* [3V]            v_template */
  const float* dirs;        /* [kdp][nv3p]     rows: posedirs[:, :, k] (k < 9(J-1)), then shapedirs[:, :, k] */
  const float* dirsT;       /* [nv3p][kdp]     transpose, for the backward GEMM */
  const float* j_templ;     /* [J][3]          J_regressor * v_template */
  const float* j_dirs;      /* [J][3][n_betas] J_regressor * shapedirs */
  const int* parents;       /* [J]             kintree_table[0]; parents[0] is ignored */
  const int* skin_idx;      /* [V][nnz] */
  const float* skin_w;      /* [V][nnz] */
} vt_smpl_model;

/* pose[B][3J] axis-angle, betas[B][n_betas], trans[B][3], offsets[B][V][3] or NULL -> verts[B][V][3], jtr[B][J][3],
 * naked[B][V][3], v_posed[B][V][3] (only written when offsets != NULL; otherwise v_posed == naked).
 * coef[B][kdp], R[B][J][9], J[B][J][3], G[B][J][12], A[B][J][12] are outputs the backward call needs again. */
int vt_smpl_fwd(const vt_smpl_model* model, const float* pose, const float* betas, const float* trans, const float* offsets,
                float scale, int B, float* coef, float* R, float* J, float* G, float* A, float* naked, float* v_posed,
                float* verts, float* jtr, void* stream);

/* g_verts[B][V][3] and / or g_jtr[B][J][3] (NULL = zero) -> g_pose[B][3J], g_betas[B][n_betas], g_trans[B][3].
 * Scratch (caller-owned, overwritten): g_vposed[B][nv3p] (= d/d offsets on return), gA[B][J][12], g_coef[B][kdp],
 * g_trans_skin[B][3]. */
int vt_smpl_bwd(const vt_smpl_model* model, const float* pose, const float* R, const float* J, const float* G, const float* A,
                const float* v_posed, const float* g_verts, const float* g_jtr, float scale, int B, float* g_vposed, float* gA,
                float* g_coef, float* g_trans_skin, float* g_pose, float* g_betas, float* g_trans, void* stream);

/* out[B][L][3] = regressor^T verts for a sparse [V x L] regressor given as CSR over landmarks (rowptr[L+1], col = vertex). */
int vt_landmarks_fwd(const float* verts, int B, int V, const int* rowptr, const int* col, const float* val, int L, float* out,
                     void* stream);
/* g_verts[B][V][3] += regressor g_out (atomics; the caller initialises g_verts). */
int vt_landmarks_bwd(const float* g_out, int B, int V, const int* rowptr, const int* col, const float* val, int L, float* g_verts,
                     void* stream);

/* ---- SMPL-T keypoint pre-fit objective and optimiser: SMPLHFitter30fps.compute_loss (preprocess/fit_SMPLH_30fps.py:153-200),
 *      BaseFitter.sum_dict / fit_one_batch / init_*_optimizer (preprocess/fit_SMPLH_kpts.py:67-75,114-190), priors
 *      (lib_smpl/th_smpl_prior.py:25-39, lib_smpl/th_hand_prior.py:46-72).  All state lives on the device:
 *        ctrl  float[vt_fit_ctrl_words()]: [0..5] weights of kpts,temp,ptemp,pose,hand,pinit ALREADY divided by (1+decay);
 *              [8] Adam lr; [9] phase (0: trans+global_pose+top_betas, 1: + body_pose + other_betas); [10] Adam step count;
 *              [11] history row to write next.  The host only rewrites it when the schedule changes.
 *        acc   double[8] per-step accumulators of the unweighted term sums (zeroed by vt_fit_begin_step)
 *        hist  double[max_hist][8]: six term means, weighted total, Adam step -- one row per optimisation step ------------- */
int vt_fit_ctrl_words(void);
int vt_fit_begin_step(double* acc, void* stream);
/* J[B][25][3] body-25 landmarks, kpts[B][25][3] = (x, y, confidence); cam4 (host) = {fx, fy, cx, cy};
 * gJ[B][25][3] = weight * d(mean reprojection error)/dJ. */
int vt_fit_kpts(const float* J, const float* kpts, int B, int L, const float* cam4, const float* ctrl, float* gJ, double* acc, void* stream);
/* g_verts[B][n_coords] = weight * d mse(v[1:-1]-v[:-2], v[2:]-v[1:-1]) / d verts  (every element is written). */
int vt_fit_temporal_verts(const float* verts, int B, int n_coords, const float* ctrl, float* g_verts, double* acc, void* stream);
/* pose[B][156]: Mahalanobis body prior (mean[63], precision[63][63]), GRAB hand prior value (2 x mean[45], precision[45][45]),
 * weighted pose second differences (joint_w[66]) and stay-near-init; g_pose[B][156] is overwritten. */
int vt_fit_pose_terms(const float* pose, const float* pose_init, int B, const float* body_mean, const float* body_prec,
                      const float* lh_mean, const float* lh_prec, const float* rh_mean, const float* rh_prec, const float* joint_w,
                      const float* ctrl, float* g_pose, double* acc, void* stream);
/* torch.optim.Adam (default betas / eps) on the entries the current phase optimises; gradient of pose = g_pose_a + g_pose_b;
 * m, v: float[B][169] moment buffers (zero them when a new optimiser starts). */
int vt_fit_adam(float* pose, float* betas, float* trans, const float* g_pose_a, const float* g_pose_b, const float* g_betas,
                const float* g_trans, float* m, float* v, int B, const float* ctrl, void* stream);
int vt_fit_end_step(const double* acc, int B, int n_coords, float* ctrl, double* hist, int max_hist, void* stream);

/* ---- joint optimisation helpers ------------------------------------------------------------------------------------------- */

/* ReconFitterBase.project_so3 (recon/recon_fit_base.py:178-199): R[B][3][3] = U diag(1,1,det(UV^T)) V^T of M[B][3][3], and the
 * vector-Jacobian product gM = (dR/dM)^T gR (what autograd derives through torch.svd / det / matmul there). */
int vt_so3_project_fwd(const float* M, int B, float* R, void* stream);
int vt_so3_project_bwd(const float* M, const float* gR, int B, float* gM, void* stream);

/* init_object_orientation (recon/recon_fit_base.py:202-216; recon/pca_util.py:59-72): R[b] = project_so3((S^T S)^-1 S^T T[b] + 1e-4 noise[b])
 * with T = tgt_axis[B][3][3] (the predicted PCA axes), S = src_axis (the template's; [B][3][3] when src_per_frame != 0, else one [3][3]
 * shared by all frames), noise[B][3][3] the U(0,1) draws of decopose_axis or NULL (PCAUtil's variant adds none). */
int vt_pca_orientation(const float* tgt_axis, const float* src_axis, int src_per_frame, const float* noise, int B, float* R, void* stream);

/* pytorch3d.loss.chamfer_distance(Pointclouds(xs), Pointclouds(ys)) with default reductions, as called by compute_contact_loss
 * (recon/recon_fit_trivis_full.py:452-456): N ragged cloud pairs, x[sum_n][3] with x_off[N+1] row offsets (same for y).
 * loss[1] = mean_n ( mean_i min_j |x_i-y_j|^2 + mean_j min_i |y_j-x_i|^2 ); nn_x / nn_y receive the arg-min rows for backward.
 * vt_chamfer_bwd ACCUMULATES into gx / gy (caller zeroes them). */
int vt_chamfer_fwd(const float* x, const int* x_off, const float* y, const int* y_off, int N, int* nn_x, int* nn_y, float* loss,
                   void* stream);
int vt_chamfer_bwd(const float* x, const int* x_off, const float* y, const int* y_off, int N, const int* nn_x, const int* nn_y,
                   const float* g_loss, float* gx, float* gy, void* stream);

/* ---- joint optimisation steps as fixed kernel sequences: ReconFitterBehave.optimize_smpl / forward_smpl (recon/recon_fit_behave.py:393-513)
 *      and ReconFitterTriVisFull.optimize_smpl_object / forward_step (recon/recon_fit_trivis_full.py:124-391).  State on the device:
 *        ctrl  float[vt_recon_ctrl_words()]: [0..15] per-term weights ALREADY divided by (1 + decay) (0 = the term is not in this phase's
 *              loss_dict); [16] lr of the first parameter group, [17] lr of the second, [18] phase, [19] early-stop tolerance, [20] early-stop
 *              window open (0/1), [21] temporal multiplier (10 in the joint phase, recon_fit_trivis_full.py:386-388), [22] noise seed (bits);
 *              [32] Adam step count, [33] next history row, [34] stopped (0/1), [35] previous total loss (fp32), [36] noise draw counter.
 *              The host rewrites words [0..31] when the schedule changes and zeroes [32..] when a loop / a new optimiser starts.
 *        acc   double[8]  per-step sums of the unweighted terms (zeroed by vt_recon_end_step)
 *        hist  double[max_hist][vt_recon_hist_ld()]: term means (NaN where the weight is 0), [14] weighted total, [15] Adam step.
 *      Term slots, SMPL refinement: 0 df_h, 1 pose, 2 hand, 3 part, 4 pinit, 5 j2d, 6 stemp; object phases: 0 otemp, 1 ovtemp, 2 mask,
 *      3 scale, 4 trans, 5 object, 6 contact.  Once the early stop has fired (ctrl[34]) the Adam and end-step kernels leave all state
 *      untouched, so replays the host had already queued are no-ops. ---- */
int vt_recon_ctrl_words(void);
int vt_recon_hist_ld(void);
/* cudaMemsetAsync(p, 0, bytes) on the caller's stream. */
int vt_zero(void* p, long long bytes, void* stream);
/* x[B][n_points][3]: second / first difference smoothness over the batch axis (temporal_loss_smpl, recon_fit_trivis_full.py:170-177;
 * temporal_loss_joint, :379-391; skipped when B < 4) with weights ctrl[iw2] / ctrl[iw1] (-1 = absent; use_k: times ctrl[21]) accumulated in
 * acc[slot2] / acc[slot1], plus up to two per-point terms from vt_query_losses_tc: g[B][n][3] (overwritten) = temporal gradients +
 * ctrl[iwA] frameA[b] gA / denA + ctrl[iwB] gB / denB, acc[slotA] += frameA[b] valsA[b][n] (frameA NULL = 1), acc[slotB] += valsB[b][n]. */
int vt_recon_point_terms(const float* x, int B, int n_points, int iw2, int iw1, int slot2, int slot1, int use_k, const float* valsA,
                         const float* gA, const float* frameA, int iwA, int slotA, float denA, const float* valsB, const float* gB, int iwB,
                         int slotB, float denB, const float* ctrl, float* g, double* acc, void* stream);
/* projection_loss (recon/recon_fit_base.py:787-802): J[B][25][3] into the network-input crop, kpts[B][25][3] = (x, y, confidence);
 * cam6 (host) = {fx_px, fy_px, cx_px, cy_px, crop_size, net_input_size}; gJ = ctrl[5] * d j2d / dJ; acc[5] += sum. */
int vt_recon_kpts(const float* J, const float* kpts, const float* crop_center, int B, int L, const float* cam6, const float* ctrl, float* gJ,
                  double* acc, void* stream);
/* compute_prior_loss (recon_fit_base.py:625-638) + the pinit term (recon_fit_behave.py:486-487): pose[B][156], pose_init[B][69];
 * g_pose[B][156] overwritten with the gradient of ctrl[1] 'pose' + ctrl[4] 'pinit' (the hand prior is value-only). */
int vt_recon_pose_terms(const float* pose, const float* pose_init, int B, const float* body_mean, const float* body_prec, const float* lh_mean,
                        const float* lh_prec, const float* rh_mean, const float* rh_prec, const float* ctrl, float* g_pose, double* acc,
                        void* stream);
/* torch.optim.Adam (defaults) on the parameters of the current phase: 0 = [top_betas, trans], 1 = [trans, global_pose, body_pose, top_betas,
 * other_betas] (recon_fit_behave.py:402,426-432); m, v: float[B][169]. */
int vt_recon_adam_smpl(float* pose, float* betas, float* trans, const float* g_pose_a, const float* g_pose_b, const float* g_betas,
                       const float* g_trans, float* m, float* v, int B, const float* ctrl, void* stream);
/* Adam on obj_R[B][9] (lr ctrl[16]; skipped in phase 2 = 'joint') and obj_t[B][3] (lr ctrl[17]) (recon_fit_trivis_full.py:300-309,339,347);
 * m, v: float[B][12]. */
int vt_recon_adam_obj(float* obj_R, float* obj_t, const float* g_R, const float* g_t, float* m, float* v, int B, const float* ctrl, void* stream);
/* Close a step: term k = acc[k] / div[k] (host array of n_terms doubles; the term in contact_slot is read from *contact_val instead, -1 = none),
 * fp32 weighted total, history row, the early-stop predicate `abs(prev - loss) / prev < prev * tol` on fp32 values (recon_fit_behave.py:452,
 * recon_fit_trivis_full.py:371), counters, acc zeroed. */
int vt_recon_end_step(double* acc, int n_terms, const double* div, int contact_slot, const float* contact_val, float* ctrl, double* hist,
                      int max_hist, void* stream);
/* decopose_axis (recon_fit_base.py:461-469) input: M = obj_R + 1e-4 * U(0,1); noise[B][9] replays given draws, NULL draws them on the device
 * (Philox4x32-10 keyed on ctrl[22], counter ctrl[36]); noise_out (optional) receives the draws used. */
int vt_recon_obj_noise(const float* obj_R, const float* noise, int B, const float* ctrl, float* M, float* noise_out, void* stream);
/* transform_obj_verts (recon_fit_base.py:455-459), row vectors: out[b][n] = (P[n] R[b] + t[b]) * s[b]; P is [N][3] (per_frame 0) or [B][N][3];
 * and its backward: gR[B][9] (+)= s P^T g, gt[B][3] (+)= s sum_n g. */
int vt_recon_obj_transform(const float* P, int per_frame, const float* R, const float* t, const float* s, int B, int N, float* out, void* stream);
int vt_recon_obj_transform_bwd(const float* P, int per_frame, const float* g, const float* s, int B, int N, int accumulate, float* gR, float* gt,
                               void* stream);
/* SilLossROI.forward + compute_mask_loss (recon/obj_pose_roi.py:183-202, recon_fit_trivis_full.py:179-184) on a rendered alpha[B][S][S]:
 * acc[2] += occ[b] sum_px (keep alpha - ref)^2, g_alpha = ctrl[2] occ[b] / B * 2 (keep alpha - ref) keep. */
int vt_recon_sil_loss(const float* alpha, const float* keep, const float* ref, const float* occ, int B, int image_size, const float* ctrl,
                      float* g_alpha, double* acc, void* stream);
/* 'scale' = mean (obj_s - s0)^2 (value only) and, with_trans, 'trans' = mean (obj_t - t_init)^2 whose gradient is ADDED to gt[B][3]. */
int vt_recon_obj_small_terms(const float* obj_t, const float* t_init, const float* obj_s, float s0, int B, int with_trans, const float* ctrl,
                             float* gt, double* acc, void* stream);
/* dst[n][3] = src[idx[n]][3]; dst[idx[n]][3] += src[n][3] (the contact sets of compute_contact_loss, recon_fit_trivis_full.py:405-449). */
int vt_recon_gather_rows(const float* src, const long long* idx, int n, float* dst, void* stream);
int vt_recon_scatter_add_rows(const float* src, const long long* idx, int n, float* dst, void* stream);

/* ---- rasteriser with neural_renderer semantics (third-party, un-vendored: restated from the upstream algorithm, PARITY
 *      UNPINNED): silhouettes for SilLossROI (recon/obj_pose_roi.py:87-94,183-202) and orthographic depth / occupancy for the
 *      triplane renderings (render/render_triplane_nr.py:25-30,88-110).  fill_back = True, near 0.1, far 100, no anti-aliasing. ---- */

/* verts[B][V][3] camera-space, faces[F][3] (shared); mode 0: nr.projection with K4[B][4] = (fx, fy, cx, cy) normalised to the ROI
 * (orig_size 1), mode 1: orthographic (x, y in [-1, 1] rasterised directly).  Outputs: faces_ndc[B][2F][9] and face_index[B][S][S]
 * (kept for backward), alpha[B][S][S] and / or depth[B][S][S] (either may be NULL). */
/* cull_ws: optional workspace of vt_raster_cull_floats(B, F) floats (16-byte aligned) for per-face / per-256-face-chunk bounding boxes; with it
 * a pixel tile skips the chunks that miss it (same image, the face list is just scanned sparsely).  NULL scans every face per tile. */
long long vt_raster_cull_floats(int B, int F);
int vt_raster_fwd(const float* verts, const int* faces, int B, int V, int F, int mode, const float* K4, int image_size,
                  float* faces_ndc, int* face_index, float* alpha, float* depth, float* cull_ws, void* stream);
/* NMR pseudo-gradient: g_alpha[B][S][S] -> g_verts[B][V][3] (overwritten); g_faces[B][2F][9] is scratch. */
int vt_raster_bwd(const float* verts, const int* faces, int B, int V, int F, int mode, const float* K4, int image_size,
                  const float* faces_ndc, const int* face_index, const float* alpha, const float* g_alpha, float* g_faces,
                  float* g_verts, void* stream);
/* vt_raster_bwd with a caller-owned workspace of vt_workspace_bytes_raster_bwd(B, image_size) bytes (2-byte aligned): prefix counts of the
 * background pixels that can contribute (alpha == 0 and g_alpha < 0) along every row and column, so that the walks from a visible edge to the
 * image border -- the bulk of the work -- are skipped where they cannot contribute.  Same result bit for bit; skip_ws == NULL = vt_raster_bwd. */
long long vt_workspace_bytes_raster_bwd(int B, int image_size);
int vt_raster_bwd_ws(const float* verts, const int* faces, int B, int V, int F, int mode, const float* K4, int image_size,
                     const float* faces_ndc, const int* face_index, const float* alpha, const float* g_alpha, float* g_faces,
                     float* g_verts, void* skip_ws, void* stream);

/* Evaluation Chamfer (the "Chamfer vs ref" half of the metric): per-point Euclidean nearest-neighbour distances between
 * dense clouds, both directions -- recon/eval/chamfer_distance.py:10-52 is mean(dist_x) + mean(dist_y) per frame (sklearn kd-tree there).
 * x[B][nx][3], y[B][ny][3] -> dist_x[B][nx] (every x_i to its nearest y), dist_y[B][ny]. */
int vt_nn_dist(const float* x, int nx, const float* y, int ny, int B, float* dist_x, float* dist_y, void* stream);

/* compute_transform (recon/eval/pose_utils.py:153-198): the similarity transform (scale, R, t) that takes cloud S1 closest to S2 in the
 * least-squares sense (orthogonal Procrustes with det R = +1), per pair of clouds S1/S2[B][N][3].  workspace: 16 doubles per pair (zeroed
 * by the call).  R[B][9] row-major, t[B][3], scale[B].  vt_similarity_apply: out = scale * R p + t (pose_utils.py:31). */
int vt_procrustes(const float* S1, const float* S2, int N, int B, double* workspace, float* R, float* t, float* scale, void* stream);
int vt_similarity_apply(const float* points, int n, int B, const float* R, const float* t, const float* scale, float* out, void* stream);

/* ---- SmoothNet stage: smoothnet/smooth_smplt.py, smooth_objrot.py, smooth_base.py, models/smoothnet*.py,
 *      utils/utils.py:63-103, utils/geometry_utils.py -- the trajectory stays in device memory between the fitting stages ---- */

/* floats of one packed SmoothNet (k-major): We[64][512] be[512] { W1[512][16] b1[16] W2[16][512] b2[512] } x n_blocks Wd[512][64] bd[64] */
long long vt_smoothnet_pack_floats(int n_blocks);

/* SMPLTSmoother.preprocess_input (smooth_smplt.py:73-101) without the windowing: poses[L][72|156] axis-angle (SMPL-H reduced to the 24
 * SMPL joints) -> 6-D rotations, seq[L][157] = [pose6d 144 | betas 10 | trans 3]. */
int vt_smooth_pack_smplt(const float* poses, int pose_dim, const float* betas, const float* trans, int L, float* seq, void* stream);

/* SmoothNet.forward (models/smoothnet.py:125-141, eval mode) on every (sliding window, channel) row of channels [c0, c0+nC) of seq[L][D]:
 * clips[b][t][c] for b in [0, L-window], window step 1 (smooth_base.py:45-73).  relative != 0: the window's first frame is subtracted from
 * the input and added back to the output (SMPL translation, smooth_smplt.py:88-91 and 39-42).  Built for window 64 / hidden 512 / residual
 * 16 (smoothnet/configs/pw3d_spin_3D.yaml).  clips: [L-window+1][window][D], only the selected channels are written. */
int vt_smoothnet_clips(const float* seq, int L, int D, int c0, int nC, int relative, int window, int hidden, int res_hidden, int n_blocks,
                       const float* wpack, float* clips, void* stream);

/* slide_window_to_sequence / clips2seq_fast for step 1 (utils/utils.py:63-103): out[l][c] = mean of clips[b][l-b][c] over the windows that
 * contain frame l; channels [pass0, pass0+passN) are taken from seq instead (betas are not smoothed, models/smoothnet_smpl.py:38-45). */
int vt_smooth_window_mean(const float* clips, const float* seq, int L, int D, int window, int pass0, int passN, float* out, void* stream);

/* SMPLTSmoother.post_processing (smooth_smplt.py:43-47): rot6D_to_axis on the 24 joints (Gram-Schmidt, kornia rotation-matrix ->
 * quaternion -> angle-axis, NaN -> 0), betas and translation split out.  seq[L][157] -> poses[L][72], betas[L][10], trans[L][3]. */
int vt_smooth_unpack_smplt(const float* seq, int L, float* poses, float* betas, float* trans, void* stream);

/* rot6d_to_rotmat (geometry_utils.py:63-77) on rot6d[L][6]; transposed != 0 writes R^T as ObjrotSmoother.post_processing stores
 * `obj_angles` (smooth_objrot.py:104-112). */
int vt_smooth_rot6d_to_rotmat(const float* rot6d, int L, int transposed, float* out, void* stream);

/* ---- HVOP-Net: model/infill/mfiller_cond.py:17-104 (ConditionalMInfiller), model/transformers/former_deci.py:31-175
 * (pre-norm encoder layers, nn.MultiheadAttention with key_padding_mask), posi_embed.py:35-66, and the autoregressive clip loop of
 * interp/test_infill_autoreg.py:34-174 / test_cinfill_autoreg.py:32-51.  All tensors fp32 on the device; tokens are rows [clip][t].
 * A layer pack holds, k-major ([in][out]):  ln1.w ln1.b | Wqkv^T [D][3D] bqkv [3D] | Wo^T [D][D] bo [D] | ln2.w ln2.b | W1^T [D][F] b1 [F] |
 * W2^T [F][D] b2 [D]. ---- */
long long vt_infill_layer_pack_floats(int D, int F);

/* Start of an encoder's first layer: optional feature projection x = in P + b (in[n_tok][in_ld], proj = P^T [in_dim][D] | b [D]; in == NULL:
 * x already holds the stream), then forward_pre up to the attention (former_deci.py:83-84): h = LayerNorm1(x), q = (h + pos) Wq / sqrt(D /
 * heads), k = (h + pos) Wk, v = h Wv -> qkv[n_tok][3D].  pos[T][D] is PositionEmbeddingSine_1D(B, T) (identical for every clip). */
int vt_infill_head(const float* in, int in_ld, int in_dim, const float* proj, float* x, int x_ld, int n_tok, int T, int D, int F, int heads,
                   const float* layer, const float* pos, float* qkv, void* stream);

/* softmax(q k^T + key_padding_mask) v per head (former_deci.py:84-88); key_mask[n_clips][T] bytes, non-zero = key ignored, NULL = none.
 * attn[n_tok][D] is the concatenation of the heads BEFORE the output projection. */
int vt_infill_attn(const float* qkv, const unsigned char* key_mask, int n_clips, int T, int D, int heads, float* attn, void* stream);

/* Rest of the layer (former_deci.py:89-93): x += attn Wo + bo; x += W2 act(W1 LayerNorm2(x)); activation 0 gelu / 1 relu / 2 leaky_relu.
 * final_ln != NULL applies the encoder's closing LayerNorm (former_deci.py:126-127, only built when the pre_norm option is set).  The stream
 * is written to y[n_tok][y_ld] (y may be x; a wider y_ld concatenates two encoders' features, mfiller_cond.py:95).  next_layer != NULL also
 * starts the following layer exactly as vt_infill_head does. */
int vt_infill_tail(float* x, int x_ld, const float* attn, int n_tok, int T, int D, int F, int heads, int activation, const float* layer,
                   const float* final_ln, float* y, int y_ld, const float* next_layer, int next_F, const float* pos, float* qkv, void* stream);

/* make_predictor (mfiller_cond.py:57-73): n_layers Linear layers with nn.LeakyReLU() between; dims[n_layers + 1] (host array);
 * pack = per layer W^T [in][out] | b [out]. */
int vt_infill_mlp(const float* x, int x_ld, int n_tok, int n_layers, const int* dims, const float* pack, float* out, int out_ld, void* stream);

/* One clip of the autoregressive loop with obj_dim 6 (test_infill_autoreg.py:93-105 and 116-153, test_cinfill_autoreg.py:43-49): frames
 * [start, start + T) of rot6d_smpl[L][144] | trans_smpl[L][3] -> data_smpl[T][147]; data_obj[T][6] takes rot6d_out (the running prediction)
 * for the first n_ctx frames and rot6d_obj for the others, rows with mask[t] != 0 multiplied by zero. */
int vt_infill_pack_clip(const float* rot6d_smpl, const float* trans_smpl, const float* rot6d_obj, const float* rot6d_out, const unsigned char* mask,
                        int L, int start, int T, int n_ctx, float* data_smpl, float* data_obj, void* stream);

/* rot6d_out[start + t] = pred[t] for t in [t0, T) (test_infill_autoreg.py:110 and 160). */
int vt_infill_commit_clip(const float* pred, int L, int start, int t0, int T, float* rot6d_out, void* stream);

/* ---- Test-time frame preparation: data/testdata_triplane.py:42-74, data/train_data.py:143-162,
 * data/base_data.py:139-171, 204-265.  uint8 frames already decoded into device memory. ---- */

/* BaseDataset.masks2bbox + center_from_masks: bbox[B][4] = (xmin, ymin, xmax, ymax) (max exclusive) of the pixels where the uint8 sum of
 * person[B][H][W] and obj[B][H][W] (wrapping, as numpy's +=) exceeds thres (127); center[B][2] = (min + max) // 2 as floats (may be NULL).
 * A frame without foreground keeps the reference's sentinels (50000, 50000, -100, -100). */
int vt_mask_bbox(const unsigned char* person, const unsigned char* obj, int B, int H, int W, int thres, int* bbox, float* center, void* stream);

/* prepare_image_crop + compose_images (+ the triplane channels of TestDataTriplane.get_item): rgb[B][H][W][3], person / obj[B][H][W],
 * triplane[B][S][S][3] or NULL (uint8, already in (right, back, top) order) -> images[B][channels][S][S] float32 in [0, 1]: channels 0-2 RGB
 * zeroed where neither mask exceeds 0.5, 3 person, 4 object, 5-7 triplane.  crop_size^2 pixels around crop_center[B][2] (zero padded,
 * base_data.py:204-232) are resized to S = net_size with OpenCV's 8-bit INTER_LINEAR arithmetic; resize_tab[3][S] (host-built, see
 * vistracker_b200/frameio.py) = source index, 11-bit weight of it, 11-bit weight of its right / lower neighbour. */
int vt_prepare_image_crop(const unsigned char* rgb, const unsigned char* person, const unsigned char* obj, const unsigned char* triplane, int B, int H, int W,
                          const float* crop_center, int crop_size, int net_size, const int* resize_tab, float* images, int channels, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VISTRACKER_B200_H */
