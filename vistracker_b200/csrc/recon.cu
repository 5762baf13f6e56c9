// Joint human-object optimisation steps as fixed kernel sequences (captured into CUDA graphs by vistracker_b200/recon_fit.py):
//   ReconFitterBehave.optimize_smpl / forward_smpl            (recon/recon_fit_behave.py:393-513)
//   ReconFitterTriVisFull.optimize_smpl_object / forward_step (recon/recon_fit_trivis_full.py:124-391)
// The reference assembles every step from a few hundred tiny torch ops + autograd + three torch.optim.Adam objects and formats a progress
// string with 8-10 .item() synchronisations per step.  Here the loss terms, their analytic gradients, the masked Adam updates, the loss
// history and the early-stop predicate are kernels over static buffers; per-phase weights / learning rates live in a device control block
// that the host rewrites only when the schedule changes.  The heavy operators of a step (SMPL-H layer, fused SIF-Net query losses, SO(3)
// projection, ragged Chamfer, rasteriser) are the kernels of smpl.cu / query_bwd_tc.cu / geom.cu / raster.cu, launched between these.
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

// ---- control block (float words).  [0..15] per-term weights ALREADY divided by (1 + decay); [16..31] schedule words written by the host;
//      [32..] counters owned by the device (the host zeroes them when a loop starts).
enum { RC_LR0 = 16, RC_LR1 = 17, RC_PHASE = 18, RC_TOL = 19, RC_ESTOP = 20, RC_TEMP_K = 21, RC_SEED = 22,
       RC_STEP = 32, RC_HIST = 33, RC_STOP = 34, RC_PREV = 35, RC_DRAW = 36, RC_WORDS = 48 };
// SMPL refinement terms (forward_smpl, in the order the reference's loss_dict is filled)
enum { T_DFH = 0, T_POSE = 1, T_HAND = 2, T_PART = 3, T_PINIT = 4, T_J2D = 5, T_STEMP = 6, T_N = 7 };
// object / joint terms (forward_step)
enum { O_OTEMP = 0, O_OVTEMP = 1, O_MASK = 2, O_SCALE = 3, O_TRANS = 4, O_OBJECT = 5, O_CONTACT = 6, O_N = 7 };
constexpr int HIST_LD = 16;       // history row: [0..13] term means (NaN = not part of this phase's loss), [14] weighted total, [15] Adam step

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------ per-point terms
// One thread per coordinate i of a [B][n] point tensor x (n = 3 * points), walking a chunk of frames with a sliding window:
//   second differences  L2 = mean_{t in [1,B-2], i} (x[t-1] - 2 x[t] + x[t+1])^2      (temporal_loss_smpl / temporal_loss_joint 'otemp')
//   first differences   L1 = mean_{t in [1,B-1], i} (x[t] - x[t-1])^2                  ('ovtemp')
// plus up to two per-point terms whose values / point gradients come from the fused query launch (vt_query_losses_tc):
//   g[s][i] = w2 dL2 + w1 dL1 + wA fA[s] gA[s][i] / denA + wB gB[s][i] / denB,   acc[slotA] += fA[s] valsA[s][p], acc[slotB] += valsB[s][p]
struct PointTerms {
  const float* x; int B; int n;
  int iw2, iw1, slot2, slot1;            // ctrl indices of the weights (-1 = term absent) and accumulator slots
  int use_k;                             // multiply w2, w1 (and the accumulated values) by ctrl[RC_TEMP_K]
  const float* valsA; const float* gA; const float* frameA; int iwA, slotA; float denA;
  const float* valsB; const float* gB; int iwB, slotB; float denB;
  float* g;
};
constexpr int PT_CHUNK = 16;
__global__ void recon_point_terms_kernel(PointTerms p, const float* __restrict__ ctrl, double* __restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int B = p.B, n = p.n;
  const int s0 = blockIdx.y * PT_CHUNK, s1 = min(B, s0 + PT_CHUNK);
  const bool temporal = B >= 4;                                        // `if verts.shape[0] < 4: return`
  const float kmul = p.use_k ? ctrl[RC_TEMP_K] : 1.f;
  const float w2 = (p.iw2 >= 0 && temporal) ? ctrl[p.iw2] * kmul : 0.f, w1 = (p.iw1 >= 0 && temporal) ? ctrl[p.iw1] * kmul : 0.f;
  // weight index -2: the gradient tensor already carries its loss weight (merged query-loss launch) -> coefficient 1 / den
  const float wA = p.iwA >= 0 ? ctrl[p.iwA] / p.denA : (p.iwA == -2 ? 1.f / p.denA : 0.f);
  const float wB = p.iwB >= 0 ? ctrl[p.iwB] / p.denB : (p.iwB == -2 ? 1.f / p.denB : 0.f);
  float l2 = 0.f, l1 = 0.f, lA = 0.f, lB = 0.f;
  if (i < n) {
    const float k2 = temporal ? w2 * 2.f / ((float)(B - 2) * (float)n) : 0.f, k1 = temporal ? w1 * 2.f / ((float)(B - 1) * (float)n) : 0.f;
    auto X = [&](int t) { return (t >= 0 && t < B) ? p.x[(size_t)t * n + i] : 0.f; };
    auto A = [&](float vm, float v0, float vp, int t) { return (t >= 1 && t <= B - 2) ? (vm - 2.f * v0 + vp) : 0.f; };
    float v0 = X(s0 - 2), v1 = X(s0 - 1), v2 = X(s0), v3 = X(s0 + 1), v4;
    for (int s = s0; s < s1; ++s) {
      v4 = X(s + 2);
      float g = 0.f;
      if (temporal) {
        const float am = A(v0, v1, v2, s - 1), a0 = A(v1, v2, v3, s), ap = A(v2, v3, v4, s + 1);
        const float dm = s >= 1 ? v2 - v1 : 0.f, dp = s <= B - 2 ? v3 - v2 : 0.f;      // x[s] - x[s-1], x[s+1] - x[s]
        g = k2 * (am - 2.f * a0 + ap) + k1 * (dm - dp);
        l2 += a0 * a0;
        l1 += dm * dm;
      }
      const size_t e = (size_t)s * n + i;
      if (p.gA || p.valsA) {
        const float f = p.frameA ? p.frameA[s] : 1.f;
        if (p.gA) g += wA * f * p.gA[e];
        if (p.valsA && i % 3 == 0) lA += f * p.valsA[(size_t)s * (n / 3) + i / 3];
      }
      if (p.gB) g += wB * p.gB[e];
      if (p.valsB && i % 3 == 0) lB += p.valsB[(size_t)s * (n / 3) + i / 3];
      p.g[e] = g;
      v0 = v1; v1 = v2; v2 = v3; v3 = v4;
    }
  }
  l2 = warp_sum(l2); l1 = warp_sum(l1); lA = warp_sum(lA); lB = warp_sum(lB);
  if ((threadIdx.x & 31) == 0) {
    if (p.iw2 >= 0 && temporal) atomicAdd(acc + p.slot2, (double)l2 * kmul);
    if (p.iw1 >= 0 && temporal) atomicAdd(acc + p.slot1, (double)l1 * kmul);
    if (p.valsA) atomicAdd(acc + p.slotA, (double)lA);
    if (p.valsB) atomicAdd(acc + p.slotB, (double)lB);
  }
}

// ------------------------------------------------------------------------------------------------ 2-D key points in the network-input crop
// projection_loss (recon/recon_fit_base.py:787-802): proj = (crop/2 + K x/z + c - crop_center) * net_in / crop;
// j2d = mean_{b,l} ((px - kx)^2 + (py - ky)^2) * conf
__global__ void recon_kpts_kernel(const float* __restrict__ J, const float* __restrict__ kpts, const float* __restrict__ crop_center, int B, int L,
                                  float fx, float fy, float cx, float cy, float crop, float net_in, const float* __restrict__ ctrl,
                                  float* __restrict__ gJ, double* __restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float e = 0.f;
  if (i < B * L) {
    const int b = i / L;
    const float x = J[i * 3], y = J[i * 3 + 1], z = J[i * 3 + 2];
    const float kx = kpts[i * 3], ky = kpts[i * 3 + 1], c = kpts[i * 3 + 2];
    const float sc = net_in / crop;
    const float px = (crop / 2 + (fx * x / z + cx) - crop_center[b * 2]) * sc, py = (crop / 2 + (fy * y / z + cy) - crop_center[b * 2 + 1]) * sc;
    const float dx = px - kx, dy = py - ky;
    e = (dx * dx + dy * dy) * c;
    const float k = ctrl[T_J2D] * 2.f * c * sc / (float)(B * L);
    gJ[i * 3] = k * dx * fx / z;
    gJ[i * 3 + 1] = k * dy * fy / z;
    gJ[i * 3 + 2] = -k * (dx * x * fx + dy * y * fy) / (z * z);
  }
  e = warp_sum(e);
  if ((threadIdx.x & 31) == 0 && e != 0.f) atomicAdd(acc + T_J2D, (double)e);
}

// ------------------------------------------------------------------------------------------------ pose-space terms, one CTA per frame
// compute_prior_loss (recon_fit_base.py:625-638): 'pose' = mean_b |(pose[3:66] - mu) P|^2 (lib_smpl/th_smpl_prior.py:25-48),
// 'hand' = sum over both hands and frames of ((pose[66:] - mu_h) P_h)^2 / 45 (th_hand_prior.py:62-72; value only, the hand pose is not optimised),
// 'pinit' = mean_b sum_k (pose[3:72] - pose_init)^2 (recon_fit_behave.py:486-487).
struct ReconPriors { const float* body_mean; const float* body_prec; const float* lh_mean; const float* lh_prec; const float* rh_mean; const float* rh_prec; };
__global__ void __launch_bounds__(128) recon_pose_terms_kernel(const float* __restrict__ pose, const float* __restrict__ pose_init, int B, ReconPriors pr,
                                                               const float* __restrict__ ctrl, float* __restrict__ g_pose, double* __restrict__ acc) {
  __shared__ float t[96], y[96], red[3][4];
  const int b = blockIdx.x, k = threadIdx.x;
  const float* p = pose + (size_t)b * 156;
  float l_pose = 0.f, l_hand = 0.f, l_pinit = 0.f, g = 0.f;
  if (k < 63) t[k] = p[3 + k] - pr.body_mean[k];
  __syncthreads();
  if (k < 63) {
    float s = 0.f;
    for (int j = 0; j < 63; ++j) s = fmaf(t[j], pr.body_prec[j * 63 + k], s);
    y[k] = s; l_pose = s * s;
  }
  __syncthreads();
  if (k < 63) {
    float s = 0.f;
    for (int j = 0; j < 63; ++j) s = fmaf(y[j], pr.body_prec[k * 63 + j], s);
    g = ctrl[T_POSE] * 2.f * s / (float)B;
  }
  if (k < 69) {                                                                         // pose[3:72] against the mocap initialisation
    const float d = p[3 + k] - pose_init[(size_t)b * 69 + k];
    l_pinit = d * d;
    g += ctrl[T_PINIT] * 2.f * d / (float)B;
  }
  __syncthreads();
  if (k < 90) t[k] = p[66 + k] - (k < 45 ? pr.lh_mean[k] : pr.rh_mean[k - 45]);
  __syncthreads();
  if (k < 90) {
    const float* P = k < 45 ? pr.lh_prec : pr.rh_prec;
    const int kk = k < 45 ? k : k - 45, o = k < 45 ? 0 : 45;
    float s = 0.f;
    for (int j = 0; j < 45; ++j) s = fmaf(t[o + j], P[j * 45 + kk], s);
    l_hand = s * s;
  }
  for (int e = k; e < 156; e += 128)
    if (e < 3 || e >= 72) g_pose[(size_t)b * 156 + e] = 0.f;
  if (k < 69) g_pose[(size_t)b * 156 + 3 + k] = g;
  float v[3] = {l_pose, l_hand, l_pinit};
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    v[q] = warp_sum(v[q]);
    if ((k & 31) == 0) red[q][k >> 5] = v[q];
  }
  __syncthreads();
  if (k < 3) {
    const double s = (double)red[k][0] + red[k][1] + red[k][2] + red[k][3];
    atomicAdd(acc + (k == 0 ? T_POSE : k == 1 ? T_HAND : T_PINIT), s);
  }
}

// ------------------------------------------------------------------------------------------------ masked Adam (torch.optim.Adam defaults)
__device__ __forceinline__ void adam_update(float* param, float g, float* m, float* v, float step, float lr) {
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  const float mm = *m = b1 * *m + (1.f - b1) * g;
  const float vv = *v = b2 * *v + (1.f - b2) * g * g;
  const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  const float denom = sqrtf(vv) / sqrtf(bc2) + eps;
  *param -= (lr / bc1) * (mm / denom);
}

// phase 0 ('global'): [top_betas, trans]; phase 1 ('smpl all pose' / 'kpts'): [trans, global_pose, body_pose, top_betas, other_betas]
// (recon_fit_behave.py:402,426-432).  The hand pose is never optimised.  Nothing moves once the early stop has fired.
__global__ void recon_adam_smpl_kernel(float* __restrict__ pose, float* __restrict__ betas, float* __restrict__ trans, const float* __restrict__ g_pose_a,
                                       const float* __restrict__ g_pose_b, const float* __restrict__ g_betas, const float* __restrict__ g_trans,
                                       float* __restrict__ m, float* __restrict__ v, int B, const float* __restrict__ ctrl) {
  const int per = 156 + 10 + 3;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * per || ctrl[RC_STOP] != 0.f) return;
  const int b = idx / per, e = idx % per;
  const int phase = (int)ctrl[RC_PHASE];
  float* param; float g;
  if (e < 156) {
    if (!(phase == 1 && e < 66)) return;
    param = pose + (size_t)b * 156 + e; g = g_pose_a[(size_t)b * 156 + e] + g_pose_b[(size_t)b * 156 + e];
  } else if (e < 166) {
    const int k = e - 156;
    if (!(k < 2 || phase == 1)) return;
    param = betas + (size_t)b * 10 + k; g = g_betas[(size_t)b * 10 + k];
  } else {
    param = trans + (size_t)b * 3 + (e - 166); g = g_trans[(size_t)b * 3 + (e - 166)];
  }
  adam_update(param, g, m + idx, v + idx, ctrl[RC_STEP] + 1.f, ctrl[RC_LR0]);
}

// obj_R [B][9] with lr ctrl[RC_LR0] (skipped in phase 2 = 'joint': Adam([obj_t])), obj_t [B][3] with lr ctrl[RC_LR1]
// (recon_fit_trivis_full.py:300-309,339,347)
__global__ void recon_adam_obj_kernel(float* __restrict__ obj_R, float* __restrict__ obj_t, const float* __restrict__ g_R, const float* __restrict__ g_t,
                                      float* __restrict__ m, float* __restrict__ v, int B, const float* __restrict__ ctrl) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 12 || ctrl[RC_STOP] != 0.f) return;
  const int b = idx / 12, e = idx % 12;
  const float step = ctrl[RC_STEP] + 1.f;
  if (e < 9) {
    if ((int)ctrl[RC_PHASE] == 2) return;
    adam_update(obj_R + b * 9 + e, g_R[b * 9 + e], m + idx, v + idx, step, ctrl[RC_LR0]);
  } else {
    adam_update(obj_t + b * 3 + e - 9, g_t[b * 3 + e - 9], m + idx, v + idx, step, ctrl[RC_LR1]);
  }
}

// ------------------------------------------------------------------------------------------------ end of a step
// unweighted sums -> means, fp32 weighted total in the reference's term order, history row, early-stop predicate
// `abs(prev_loss - loss) / prev_loss < prev_loss * tol` evaluated on fp32 values as the reference's tensors are
// (recon_fit_behave.py:452, recon_fit_trivis_full.py:371; `and it > ...` is the host's RC_ESTOP window), counters.
struct EndSpec { int n_terms; double div[8]; int contact_slot; const float* contact_val; };
__global__ void recon_end_step_kernel(double* __restrict__ acc, EndSpec sp, float* __restrict__ ctrl, double* __restrict__ hist, int max_hist) {
  if (threadIdx.x != 0) return;
  if (ctrl[RC_STOP] == 0.f) {
    float total = 0.f;
    const int h = (int)ctrl[RC_HIST];
    for (int k = 0; k < sp.n_terms; ++k) {
      const float w = ctrl[k];
      double t = (k == sp.contact_slot && sp.contact_val) ? (double)*sp.contact_val : acc[k] / sp.div[k];
      if (w == 0.f) t = nan("");                                       // not in this phase's loss_dict
      else total += w * (float)t;
      if (h < max_hist) hist[(size_t)h * HIST_LD + k] = t;
    }
    if (h < max_hist) {
      hist[(size_t)h * HIST_LD + 14] = (double)total;
      hist[(size_t)h * HIST_LD + 15] = (double)ctrl[RC_STEP] + 1.0;
    }
    const float prev = ctrl[RC_PREV];
    if (ctrl[RC_ESTOP] != 0.f && fabsf(prev - total) / prev < prev * ctrl[RC_TOL]) ctrl[RC_STOP] = 1.f;
    ctrl[RC_PREV] = total;
    ctrl[RC_HIST] = (float)(h + 1);
    ctrl[RC_STEP] = ctrl[RC_STEP] + 1.f;
    ctrl[RC_DRAW] = ctrl[RC_DRAW] + 1.f;
  }
  for (int k = 0; k < 8; ++k) acc[k] = 0.0;
}

// ------------------------------------------------------------------------------------------------ object pose: noise, rigid transform
// Philox4x32-10 (counter = (element, draw), key = seed): the U[0,1) draws of decopose_axis (recon_fit_base.py:461-469) without a host RNG
__device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
__device__ inline float philox_uniform(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
  uint32_t c[4] = {c0, c1, 0u, 0u};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return (float)(c[0] >> 8) * (1.0f / 16777216.0f);
}

// M = obj_R + 1e-4 * noise; noise [B][9] from the caller (parity runs replay the reference's draws) or, when NULL, Philox keyed on
// (ctrl[RC_SEED], draw counter ctrl[RC_DRAW]) and written to noise_out for inspection
__global__ void recon_obj_noise_kernel(const float* __restrict__ obj_R, const float* __restrict__ noise, int B, const float* __restrict__ ctrl,
                                       float* __restrict__ M, float* __restrict__ noise_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 9) return;
  float u;
  if (noise) u = noise[i];
  else u = philox_uniform((uint32_t)i, (uint32_t)ctrl[RC_DRAW], __float_as_uint(ctrl[RC_SEED]), 0x5eedu);
  if (noise_out) noise_out[i] = u;
  M[i] = obj_R[i] + 1e-4f * u;
}

// transform_obj_verts (recon_fit_base.py:455-459), row-vector convention: out[b][n] = (P[n] R[b] + t[b]) * s[b]; P is [N][3] shared by the
// batch (per_frame = 0) or [B][N][3]
__global__ void recon_obj_transform_kernel(const float* __restrict__ P, int per_frame, const float* __restrict__ R, const float* __restrict__ t,
                                           const float* __restrict__ s, int B, int N, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int b = i / N, n = i % N;
  const float* p = P + ((size_t)(per_frame ? b : 0) * N + n) * 3;
  const float* r = R + b * 9;
  const float x = p[0], y = p[1], z = p[2], sc = s[b];
  out[(size_t)i * 3 + 0] = (x * r[0] + y * r[3] + z * r[6] + t[b * 3 + 0]) * sc;
  out[(size_t)i * 3 + 1] = (x * r[1] + y * r[4] + z * r[7] + t[b * 3 + 1]) * sc;
  out[(size_t)i * 3 + 2] = (x * r[2] + y * r[5] + z * r[8] + t[b * 3 + 2]) * sc;
}

// backward of the transform, one CTA per frame: gR[b] (+)= s P^T g[b], gt[b] (+)= s sum_n g[b][n]
__global__ void __launch_bounds__(256) recon_obj_transform_bwd_kernel(const float* __restrict__ P, int per_frame, const float* __restrict__ g,
                                                                      const float* __restrict__ s, int N, int accumulate, float* __restrict__ gR,
                                                                      float* __restrict__ gt) {
  __shared__ float red[12][8];
  const int b = blockIdx.x;
  float a[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) a[k] = 0.f;
  const float* Pb = P + (size_t)(per_frame ? b : 0) * N * 3;
  const float* gb = g + (size_t)b * N * 3;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float x = Pb[n * 3], y = Pb[n * 3 + 1], z = Pb[n * 3 + 2];
    const float g0 = gb[n * 3], g1 = gb[n * 3 + 1], g2 = gb[n * 3 + 2];
    a[0] += x * g0; a[1] += x * g1; a[2] += x * g2;
    a[3] += y * g0; a[4] += y * g1; a[5] += y * g2;
    a[6] += z * g0; a[7] += z * g1; a[8] += z * g2;
    a[9] += g0; a[10] += g1; a[11] += g2;
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    a[k] = warp_sum(a[k]);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = a[k];
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    v *= s[b];
    float* dst = threadIdx.x < 9 ? gR + b * 9 + threadIdx.x : gt + b * 3 + threadIdx.x - 9;
    *dst = accumulate ? *dst + v : v;
  }
}

// ------------------------------------------------------------------------------------------------ silhouette loss (SilLossROI.forward)
// image = keep * alpha; per_frame = sum_px (image - ref)^2; 'mask' = mean_b per_frame[b] * occ[b] (recon_fit_trivis_full.py:179-184);
// g_alpha = w_mask occ[b] / B * 2 (image - ref) keep.  grid (ceil(S*S / 256), B)
__global__ void recon_sil_loss_kernel(const float* __restrict__ alpha, const float* __restrict__ keep, const float* __restrict__ ref,
                                      const float* __restrict__ occ, int B, int npx, const float* __restrict__ ctrl, float* __restrict__ g_alpha,
                                      double* __restrict__ acc) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  float e = 0.f;
  if (i < npx) {
    const size_t o = (size_t)b * npx + i;
    const float k = keep[o], d = k * alpha[o] - ref[o];
    e = d * d * occ[b];
    g_alpha[o] = ctrl[O_MASK] * occ[b] / (float)B * 2.f * d * k;
  }
  e = warp_sum(e);
  if ((threadIdx.x & 31) == 0 && e != 0.f) atomicAdd(acc + O_MASK, (double)e);
}

// 'scale' = mean (s - s0)^2 (value only: the scale is not optimised), 'trans' = mean (t - t_init)^2 with its gradient added to gt
__global__ void recon_obj_small_terms_kernel(const float* __restrict__ obj_t, const float* __restrict__ t_init, const float* __restrict__ obj_s,
                                             float s0, int B, int with_trans, const float* __restrict__ ctrl, float* __restrict__ gt,
                                             double* __restrict__ acc) {
  float ls = 0.f, lt = 0.f;
  for (int i = threadIdx.x; i < B * 3; i += blockDim.x) {
    if (i < B) { const float d = obj_s[i] - s0; ls += d * d; }
    if (with_trans) {
      const float d = obj_t[i] - t_init[i];
      lt += d * d;
      gt[i] += ctrl[O_TRANS] * 2.f * d / (float)(B * 3);
    }
  }
  ls = warp_sum(ls); lt = warp_sum(lt);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(acc + O_SCALE, (double)ls);
    if (with_trans) atomicAdd(acc + O_TRANS, (double)lt);
  }
}

// rows of a [M][3] array by index (contact sets of compute_contact_loss, recon_fit_trivis_full.py:405-449) and the scatter-add back
__global__ void recon_gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx, int n, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  dst[i] = src[idx[i / 3] * 3 + i % 3];
}
__global__ void recon_scatter_add_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx, int n, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  atomicAdd(dst + idx[i / 3] * 3 + i % 3, src[i]);
}

}  // namespace vt

using namespace vt;

extern "C" {

int vt_recon_ctrl_words(void) { return RC_WORDS; }
int vt_recon_hist_ld(void) { return HIST_LD; }

int vt_recon_point_terms(const float* x, int B, int n_points, int iw2, int iw1, int slot2, int slot1, int use_k, const float* valsA,
                         const float* gA, const float* frameA, int iwA, int slotA, float denA, const float* valsB, const float* gB, int iwB,
                         int slotB, float denB, const float* ctrl, float* g, double* acc, void* stream) {
  VT_CHECK_ARG(B >= 1 && n_points >= 1, "vt_recon_point_terms: empty input (B=%d, points=%d)", B, n_points);
  PointTerms p{x, B, 3 * n_points, iw2, iw1, slot2, slot1, use_k, valsA, gA, frameA, iwA, slotA, denA, valsB, gB, iwB, slotB, denB, g};
  dim3 grid(ceil_div(3 * n_points, 256), ceil_div(B, PT_CHUNK));
  recon_point_terms_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, ctrl, acc);
  VT_CHECK_LAUNCH("vt_recon_point_terms");
  return 0;
}

int vt_recon_kpts(const float* J, const float* kpts, const float* crop_center, int B, int L, const float* cam6, const float* ctrl, float* gJ,
                  double* acc, void* stream) {
  if (B <= 0) return 0;
  recon_kpts_kernel<<<ceil_div(B * L, 128), 128, 0, (cudaStream_t)stream>>>(J, kpts, crop_center, B, L, cam6[0], cam6[1], cam6[2], cam6[3], cam6[4],
                                                                           cam6[5], ctrl, gJ, acc);
  VT_CHECK_LAUNCH("vt_recon_kpts");
  return 0;
}

int vt_recon_pose_terms(const float* pose, const float* pose_init, int B, const float* body_mean, const float* body_prec, const float* lh_mean,
                        const float* lh_prec, const float* rh_mean, const float* rh_prec, const float* ctrl, float* g_pose, double* acc,
                        void* stream) {
  if (B <= 0) return 0;
  ReconPriors pr{body_mean, body_prec, lh_mean, lh_prec, rh_mean, rh_prec};
  recon_pose_terms_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(pose, pose_init, B, pr, ctrl, g_pose, acc);
  VT_CHECK_LAUNCH("vt_recon_pose_terms");
  return 0;
}

int vt_recon_adam_smpl(float* pose, float* betas, float* trans, const float* g_pose_a, const float* g_pose_b, const float* g_betas,
                       const float* g_trans, float* m, float* v, int B, const float* ctrl, void* stream) {
  if (B <= 0) return 0;
  recon_adam_smpl_kernel<<<ceil_div(B * 169, 256), 256, 0, (cudaStream_t)stream>>>(pose, betas, trans, g_pose_a, g_pose_b, g_betas, g_trans, m, v, B, ctrl);
  VT_CHECK_LAUNCH("vt_recon_adam_smpl");
  return 0;
}

int vt_recon_adam_obj(float* obj_R, float* obj_t, const float* g_R, const float* g_t, float* m, float* v, int B, const float* ctrl, void* stream) {
  if (B <= 0) return 0;
  recon_adam_obj_kernel<<<ceil_div(B * 12, 128), 128, 0, (cudaStream_t)stream>>>(obj_R, obj_t, g_R, g_t, m, v, B, ctrl);
  VT_CHECK_LAUNCH("vt_recon_adam_obj");
  return 0;
}

int vt_recon_end_step(double* acc, int n_terms, const double* div, int contact_slot, const float* contact_val, float* ctrl, double* hist,
                      int max_hist, void* stream) {
  VT_CHECK_ARG(n_terms >= 1 && n_terms <= 8, "vt_recon_end_step: 1..8 terms (got %d)", n_terms);
  EndSpec sp;
  sp.n_terms = n_terms;
  for (int k = 0; k < 8; ++k) sp.div[k] = k < n_terms ? div[k] : 1.0;
  sp.contact_slot = contact_slot; sp.contact_val = contact_val;
  recon_end_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc, sp, ctrl, hist, max_hist);
  VT_CHECK_LAUNCH("vt_recon_end_step");
  return 0;
}

int vt_recon_obj_noise(const float* obj_R, const float* noise, int B, const float* ctrl, float* M, float* noise_out, void* stream) {
  if (B <= 0) return 0;
  recon_obj_noise_kernel<<<ceil_div(B * 9, 128), 128, 0, (cudaStream_t)stream>>>(obj_R, noise, B, ctrl, M, noise_out);
  VT_CHECK_LAUNCH("vt_recon_obj_noise");
  return 0;
}

int vt_recon_obj_transform(const float* P, int per_frame, const float* R, const float* t, const float* s, int B, int N, float* out, void* stream) {
  if (B <= 0 || N <= 0) return 0;
  recon_obj_transform_kernel<<<ceil_div(B * N, 256), 256, 0, (cudaStream_t)stream>>>(P, per_frame, R, t, s, B, N, out);
  VT_CHECK_LAUNCH("vt_recon_obj_transform");
  return 0;
}

int vt_recon_obj_transform_bwd(const float* P, int per_frame, const float* g, const float* s, int B, int N, int accumulate, float* gR, float* gt,
                               void* stream) {
  if (B <= 0) return 0;
  recon_obj_transform_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(P, per_frame, g, s, N, accumulate, gR, gt);
  VT_CHECK_LAUNCH("vt_recon_obj_transform_bwd");
  return 0;
}

int vt_recon_sil_loss(const float* alpha, const float* keep, const float* ref, const float* occ, int B, int image_size, const float* ctrl,
                      float* g_alpha, double* acc, void* stream) {
  if (B <= 0) return 0;
  const int npx = image_size * image_size;
  dim3 grid(ceil_div(npx, 256), B);
  recon_sil_loss_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(alpha, keep, ref, occ, B, npx, ctrl, g_alpha, acc);
  VT_CHECK_LAUNCH("vt_recon_sil_loss");
  return 0;
}

int vt_recon_obj_small_terms(const float* obj_t, const float* t_init, const float* obj_s, float s0, int B, int with_trans, const float* ctrl,
                             float* gt, double* acc, void* stream) {
  if (B <= 0) return 0;
  recon_obj_small_terms_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(obj_t, t_init, obj_s, s0, B, with_trans, ctrl, gt, acc);
  VT_CHECK_LAUNCH("vt_recon_obj_small_terms");
  return 0;
}

int vt_recon_gather_rows(const float* src, const long long* idx, int n, float* dst, void* stream) {
  if (n <= 0) return 0;
  recon_gather_rows_kernel<<<ceil_div(n * 3, 256), 256, 0, (cudaStream_t)stream>>>(src, idx, n, dst);
  VT_CHECK_LAUNCH("vt_recon_gather_rows");
  return 0;
}

int vt_recon_scatter_add_rows(const float* src, const long long* idx, int n, float* dst, void* stream) {
  if (n <= 0) return 0;
  recon_scatter_add_rows_kernel<<<ceil_div(n * 3, 256), 256, 0, (cudaStream_t)stream>>>(src, idx, n, dst);
  VT_CHECK_LAUNCH("vt_recon_scatter_add_rows");
  return 0;
}

}  // extern "C"
