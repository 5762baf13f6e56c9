// SMPL-T keypoint pre-fit objective (preprocess/fit_SMPLH_30fps.py:153-200, preprocess/fit_SMPLH_kpts.py:67-75,114-190,
// 306-310; priors: lib_smpl/th_smpl_prior.py:25-39, lib_smpl/th_hand_prior.py:46-72) as loss+gradient kernels and a fused
// masked Adam.  The reference builds these terms from ~40 tiny torch ops per step, re-reads three prior pickles from disk
// every step and synchronises 7 times per step for its progress string; here one optimisation step is a fixed sequence of
// ~17 launches with no host involvement (captured into a CUDA graph by vistracker_b200/fit_smplt.py): the per-term
// weights, the Adam step count / learning rate / phase live in a small device control block.
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

// device control block (float words): [0..7] term weights already divided by (1 + decay)
//   0 kpts, 1 temp, 2 ptemp, 3 pose, 4 hand, 5 pinit ; [8] lr ; [9] phase (0 global, 1 all pose) ; [10] adam step ; [11] history index
enum { W_KPTS = 0, W_TEMP = 1, W_PTEMP = 2, W_POSE = 3, W_HAND = 4, W_PINIT = 5, C_LR = 8, C_PHASE = 9, C_STEP = 10, C_HIST = 11, C_WORDS = 16 };
constexpr int N_TERMS = 6;

__global__ void fit_begin_step_kernel(double* acc) { if (threadIdx.x < 8) acc[threadIdx.x] = 0.0; }

// 2-D reprojection of the body-25 joints: err = (proj - k2d)^2 * conf, loss = mean over B*25*2 (fit_SMPLH_30fps.py:166-168)
__global__ void fit_kpts_kernel(const float* __restrict__ J, const float* __restrict__ kpts, int B, int L, float fx, float fy,
                                float cx, float cy, const float* __restrict__ ctrl, float* __restrict__ gJ, double* __restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float e = 0.f;
  if (i < B * L) {
    const float x = J[i * 3], y = J[i * 3 + 1], z = J[i * 3 + 2];
    const float kx = kpts[i * 3], ky = kpts[i * 3 + 1], c = kpts[i * 3 + 2];
    const float px = x * fx / z + cx, py = y * fy / z + cy;
    const float dx = px - kx, dy = py - ky;
    e = (dx * dx + dy * dy) * c;
    const float k = ctrl[W_KPTS] * 2.f * c / (float)(B * L * 2);
    gJ[i * 3] = k * dx * fx / z;
    gJ[i * 3 + 1] = k * dy * fy / z;
    gJ[i * 3 + 2] = -k * (dx * x * fx + dy * y * fy) / (z * z);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if ((threadIdx.x & 31) == 0 && e != 0.f) atomicAdd(acc + W_KPTS, (double)e);
}

// vertex second-difference smoothness: loss = mean_t,i (v[t-1] - 2 v[t] + v[t+1])^2 over t in [1, B-2] (fit_SMPLH_30fps.py:196-200)
// g_v[s] = w * 2/N * (a[s-1] - 2 a[s] + a[s+1]).  Thread = one coordinate, walks a chunk of frames with a 5-frame window.
constexpr int TV_CHUNK = 16;
__global__ void fit_temporal_verts_kernel(const float* __restrict__ verts, int B, int n, const float* __restrict__ ctrl,
                                          float* __restrict__ g_verts, double* __restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int s0 = blockIdx.y * TV_CHUNK, s1 = min(B, s0 + TV_CHUNK);
  float loss = 0.f;
  if (i < n) {
    const float k = ctrl[W_TEMP] * 2.f / ((float)(B - 2) * (float)n);
    auto V = [&](int t) { return (t >= 0 && t < B) ? verts[(size_t)t * n + i] : 0.f; };
    auto A = [&](float vm, float v0, float vp, int t) { return (t >= 1 && t <= B - 2) ? (vm - 2.f * v0 + vp) : 0.f; };
    float v0 = V(s0 - 2), v1 = V(s0 - 1), v2 = V(s0), v3 = V(s0 + 1), v4;
    for (int s = s0; s < s1; ++s) {
      v4 = V(s + 2);
      const float am = A(v0, v1, v2, s - 1), a0 = A(v1, v2, v3, s), ap = A(v2, v3, v4, s + 1);
      g_verts[(size_t)s * n + i] = k * (am - 2.f * a0 + ap);
      loss += a0 * a0;
      v0 = v1; v1 = v2; v2 = v3; v3 = v4;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc + W_TEMP, (double)loss);
}

// pose-space terms, one CTA (128 threads) per frame: Mahalanobis body prior, GRAB hand prior, pose second differences, stay-near-init
struct PosePriors { const float* body_mean; const float* body_prec; const float* lh_mean; const float* lh_prec; const float* rh_mean;
                    const float* rh_prec; const float* joint_w; };
__global__ void __launch_bounds__(128) fit_pose_terms_kernel(const float* __restrict__ pose, const float* __restrict__ pose_init, int B,
                                                             PosePriors pr, const float* __restrict__ ctrl, float* __restrict__ g_pose,
                                                             double* __restrict__ acc) {
  __shared__ float t[96], y[96], gsh[66], red[4][4];
  const int b = blockIdx.x, k = threadIdx.x;
  const float* p = pose + (size_t)b * 156;
  float l_pose = 0.f, l_hand = 0.f, l_ptemp = 0.f, l_pinit = 0.f;
  if (k < 66) gsh[k] = 0.f;
  // body prior: y = (pose[3:66] - mean) P ; loss = mean_b sum y^2
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
    float g = ctrl[W_POSE] * 2.f * s / (float)B;
    const float d = p[3 + k] - pose_init[(size_t)b * 156 + 3 + k];                       // pinit: mean over B*63
    l_pinit = d * d;
    g += ctrl[W_PINIT] * 2.f * d / (float)(B * 63);
    gsh[3 + k] = g;
  }
  __syncthreads();
  // hand prior value (the hand pose is never optimised, so no gradient): sum over both hands / 45 (th_hand_prior.py:62-72)
  if (k < 90) t[k] = p[66 + k] - (k < 45 ? pr.lh_mean[k] : pr.rh_mean[k - 45]);
  __syncthreads();
  if (k < 90) {
    const float* P = k < 45 ? pr.lh_prec : pr.rh_prec;
    const int kk = k < 45 ? k : k - 45, o = k < 45 ? 0 : 45;
    float s = 0.f;
    for (int j = 0; j < 45; ++j) s = fmaf(t[o + j], P[j * 45 + kk], s);
    l_hand = s * s;
  }
  // pose second differences on the first 66 parameters with per-joint weights (fit_SMPLH_30fps.py:189-194)
  float gp = 0.f;
  if (k < 66) {
    auto Pz = [&](int tt) { return (tt >= 0 && tt < B) ? pose[(size_t)tt * 156 + k] : 0.f; };
    auto A = [&](int tt) { return (tt >= 1 && tt <= B - 2) ? (Pz(tt - 1) - 2.f * Pz(tt) + Pz(tt + 1)) : 0.f; };
    const float a0 = A(b), jw = pr.joint_w[k];
    l_ptemp = a0 * a0 * jw;
    gp = ctrl[W_PTEMP] * 2.f * jw * (A(b - 1) - 2.f * a0 + A(b + 1)) / ((float)(B - 2) * 66.f) + gsh[k];
  }
  if (k < 156) g_pose[(size_t)b * 156 + k] = gp;
  // block reduction of the four loss terms
  float v[4] = {l_pose, l_hand, l_ptemp, l_pinit};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    if ((k & 31) == 0) red[q][k >> 5] = v[q];
  }
  __syncthreads();
  if (k < 4) {
    const double s = (double)red[k][0] + red[k][1] + red[k][2] + red[k][3];
    const int slot = k == 0 ? W_POSE : k == 1 ? W_HAND : k == 2 ? W_PTEMP : W_PINIT;
    atomicAdd(acc + slot, s);
  }
}

// fused masked Adam over [pose | betas | trans] (torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8, no weight decay)
// phase 0 optimises trans, global_pose (pose[:3]), top_betas (betas[:2]); phase 1 adds body_pose (pose[3:66]) and other_betas
// (fit_SMPLH_kpts.py:182-190).  The hand pose is never touched.
__global__ void fit_adam_kernel(float* __restrict__ pose, float* __restrict__ betas, float* __restrict__ trans, const float* __restrict__ g_pose_a,
                                const float* __restrict__ g_pose_b, const float* __restrict__ g_betas, const float* __restrict__ g_trans,
                                float* __restrict__ m, float* __restrict__ v, int B, const float* __restrict__ ctrl) {
  const int per = 156 + 10 + 3;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * per) return;
  const int b = idx / per, e = idx % per;
  const int phase = (int)ctrl[C_PHASE];
  float* param; float g;
  if (e < 156) {
    if (!(e < 3 || (phase == 1 && e < 66))) return;
    param = pose + (size_t)b * 156 + e; g = g_pose_a[(size_t)b * 156 + e] + g_pose_b[(size_t)b * 156 + e];
  } else if (e < 166) {
    const int k = e - 156;
    if (!(k < 2 || phase == 1)) return;
    param = betas + (size_t)b * 10 + k; g = g_betas[(size_t)b * 10 + k];
  } else {
    param = trans + (size_t)b * 3 + (e - 166); g = g_trans[(size_t)b * 3 + (e - 166)];
  }
  const float step = ctrl[C_STEP] + 1.f, lr = ctrl[C_LR];
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  const float mm = m[idx] = b1 * m[idx] + (1.f - b1) * g;          // exp_avg.lerp_(grad, 1 - beta1)
  const float vv = v[idx] = b2 * v[idx] + (1.f - b2) * g * g;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  const float denom = sqrtf(vv) / sqrtf(bc2) + eps;
  *param -= (lr / bc1) * (mm / denom);
}

// close the step: unweighted terms -> means, weighted total, history row, Adam step counter
__global__ void fit_end_step_kernel(const double* __restrict__ acc, int B, int n_coords, float* __restrict__ ctrl, double* __restrict__ hist,
                                    int max_hist) {
  if (threadIdx.x != 0) return;
  double t[N_TERMS];
  t[W_KPTS] = acc[W_KPTS] / ((double)B * 25 * 2);
  t[W_TEMP] = acc[W_TEMP] / ((double)(B - 2) * n_coords);
  t[W_PTEMP] = acc[W_PTEMP] / ((double)(B - 2) * 66);
  t[W_POSE] = acc[W_POSE] / (double)B;
  t[W_HAND] = acc[W_HAND] / 45.0;
  t[W_PINIT] = acc[W_PINIT] / ((double)B * 63);
  double total = 0;
  for (int k = 0; k < N_TERMS; ++k) total += (double)ctrl[k] * t[k];
  const int h = (int)ctrl[C_HIST];
  if (h < max_hist) {
    for (int k = 0; k < N_TERMS; ++k) hist[(size_t)h * 8 + k] = t[k];
    hist[(size_t)h * 8 + 6] = total;
    hist[(size_t)h * 8 + 7] = (double)ctrl[C_STEP] + 1.0;
  }
  ctrl[C_HIST] = (float)(h + 1);
  ctrl[C_STEP] = ctrl[C_STEP] + 1.f;
}

}  // namespace vt

using namespace vt;

extern "C" {

int vt_fit_ctrl_words(void) { return C_WORDS; }

int vt_fit_begin_step(double* acc, void* stream) {
  fit_begin_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc);
  VT_CHECK_LAUNCH("vt_fit_begin_step");
  return 0;
}

int vt_fit_kpts(const float* J, const float* kpts, int B, int L, const float* cam4, const float* ctrl, float* gJ, double* acc, void* stream) {
  VT_CHECK_ARG(L == 25, "vt_fit_kpts: the SMPL-T objective uses the 25 body landmarks (got %d)", L);
  if (B <= 0) return 0;
  fit_kpts_kernel<<<ceil_div(B * L, 128), 128, 0, (cudaStream_t)stream>>>(J, kpts, B, L, cam4[0], cam4[1], cam4[2], cam4[3], ctrl, gJ, acc);
  VT_CHECK_LAUNCH("vt_fit_kpts");
  return 0;
}

int vt_fit_temporal_verts(const float* verts, int B, int n_coords, const float* ctrl, float* g_verts, double* acc, void* stream) {
  VT_CHECK_ARG(B >= 3, "vt_fit_temporal_verts: second differences need at least 3 frames (got %d)", B);
  dim3 grid(ceil_div(n_coords, 256), ceil_div(B, TV_CHUNK));
  fit_temporal_verts_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(verts, B, n_coords, ctrl, g_verts, acc);
  VT_CHECK_LAUNCH("vt_fit_temporal_verts");
  return 0;
}

int vt_fit_pose_terms(const float* pose, const float* pose_init, int B, const float* body_mean, const float* body_prec,
                      const float* lh_mean, const float* lh_prec, const float* rh_mean, const float* rh_prec, const float* joint_w,
                      const float* ctrl, float* g_pose, double* acc, void* stream) {
  VT_CHECK_ARG(B >= 3, "vt_fit_pose_terms: second differences need at least 3 frames (got %d)", B);
  PosePriors pr{body_mean, body_prec, lh_mean, lh_prec, rh_mean, rh_prec, joint_w};
  fit_pose_terms_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(pose, pose_init, B, pr, ctrl, g_pose, acc);
  VT_CHECK_LAUNCH("vt_fit_pose_terms");
  return 0;
}

int vt_fit_adam(float* pose, float* betas, float* trans, const float* g_pose_a, const float* g_pose_b, const float* g_betas,
                const float* g_trans, float* m, float* v, int B, const float* ctrl, void* stream) {
  if (B <= 0) return 0;
  fit_adam_kernel<<<ceil_div(B * 169, 256), 256, 0, (cudaStream_t)stream>>>(pose, betas, trans, g_pose_a, g_pose_b, g_betas, g_trans, m, v, B, ctrl);
  VT_CHECK_LAUNCH("vt_fit_adam");
  return 0;
}

int vt_fit_end_step(const double* acc, int B, int n_coords, float* ctrl, double* hist, int max_hist, void* stream) {
  fit_end_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc, B, n_coords, ctrl, hist, max_hist);
  VT_CHECK_LAUNCH("vt_fit_end_step");
  return 0;
}

}  // extern "C"
