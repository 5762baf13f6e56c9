// SMPL-H linear-blend-skinning layer, forward and analytic backward
// (lib_smpl/smplpytorch/smplpytorch/pytorch/smpl_layer.py:73-176, rodrigues_layer.py:13-52, tensutils.py:6-53) and the
// sparse landmark regressors (lib_smpl/torch_functions.py:52-76, lib_smpl/wrapper_pytorch.py:187-203).
//
// The reference issues ~1.5k tiny torch kernels per forward (52 Rodrigues x ~10 ops, a 51-step Python chain, a 52-step
// fill loop, 3*B sparse.mm) and the same again in backward.  Here a forward is 3 launches and a backward 3:
//   pose_fwd   (1 CTA / frame)   Rodrigues, blend coefficients [R-I | betas], joints, kinematic chain, A_j = G_j - [0 | G_j J_j]
//   sgemm      (tiled FFMA GEMM) v_posed = template + coef[B,469] x dirs[469, 3V]      (pose + shape blend shapes in ONE GEMM)
//   skin_fwd   (1 thread / vertex) verts = (sum_j w_vj A_j) [v_posed; 1] * scale + trans   (ELL skinning weights, 4 nnz in SMPL)
//   skin_bwd   g_v_posed, g_A (shared-memory atomics -> global), g_trans
//   sgemm      g_coef[B,469] += g_v_posed[B,3V] x dirs^T   (split-K)
//   pose_bwd   reverse chain, joint-regressor and Rodrigues backward (forward-mode dual numbers) -> g_pose, g_betas, g_trans
// Joints use J = Jreg*template + (Jreg*shapedirs) betas, precomputed in fp64 at load (exactly the reference's
// Jreg (template + shapedirs betas), re-associated).
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

constexpr int SM_MAXJ = 64;

// ------------------------------------------------------------------------------------------------ Rodrigues
// batch_rodrigues + quat2mat: angle = |theta + 1e-8|, axis = theta / angle, q = (cos a/2, sin a/2 * axis), re-normalised.
template <typename T>
__device__ __forceinline__ void quat_to_rot(T w, T x, T y, T z, T (&R)[9]) {
  T w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
  T wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  R[0] = w2 + x2 - y2 - z2; R[1] = xy * 2.f - wz * 2.f;   R[2] = wy * 2.f + xz * 2.f;
  R[3] = wz * 2.f + xy * 2.f; R[4] = w2 - x2 + y2 - z2;   R[5] = yz * 2.f - wx * 2.f;
  R[6] = xz * 2.f - wy * 2.f; R[7] = wx * 2.f + yz * 2.f; R[8] = w2 - x2 - y2 + z2;
}

// forward-mode dual number with three tangents (d/d theta_0..2)
struct D3 {
  float v, d[3];
  __device__ D3() {}
  __device__ D3(float a) : v(a) { d[0] = d[1] = d[2] = 0.f; }
};
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { D3 r; r.v = a.v + b.v; for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { D3 r; r.v = a.v - b.v; for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
__device__ __forceinline__ D3 operator*(D3 a, D3 b) { D3 r; r.v = a.v * b.v; for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
__device__ __forceinline__ D3 operator*(D3 a, float s) { D3 r; r.v = a.v * s; for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] * s; return r; }
__device__ __forceinline__ D3 operator/(D3 a, D3 b) {
  D3 r; float inv = 1.f / b.v; r.v = a.v * inv;
  for (int i = 0; i < 3; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
__device__ __forceinline__ D3 dsqrt(D3 a) { D3 r; r.v = sqrtf(a.v); float k = 0.5f / r.v; for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] * k; return r; }
__device__ __forceinline__ D3 dsin(D3 a) { D3 r; float c = cosf(a.v); r.v = sinf(a.v); for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] * c; return r; }
__device__ __forceinline__ D3 dcos(D3 a) { D3 r; float s = -sinf(a.v); r.v = cosf(a.v); for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] * s; return r; }

__device__ __forceinline__ void rodrigues(const float* th, float (&R)[9]) {
  float t0 = th[0] + 1e-8f, t1 = th[1] + 1e-8f, t2 = th[2] + 1e-8f;
  float a = sqrtf(t0 * t0 + t1 * t1 + t2 * t2);
  float n0 = th[0] / a, n1 = th[1] / a, n2 = th[2] / a;
  float h = a * 0.5f, c = cosf(h), s = sinf(h);
  float qx = s * n0, qy = s * n1, qz = s * n2;
  float qn = sqrtf(c * c + qx * qx + qy * qy + qz * qz);
  quat_to_rot<float>(c / qn, qx / qn, qy / qn, qz / qn, R);
}

// g_theta[k] = sum_e gR[e] * dR[e]/dtheta_k
__device__ __forceinline__ void rodrigues_bwd(const float* th, const float* gR, float* g_theta) {
  D3 x[3];
  for (int i = 0; i < 3; ++i) { x[i] = D3(th[i]); x[i].d[i] = 1.f; }
  D3 t0 = x[0] + D3(1e-8f), t1 = x[1] + D3(1e-8f), t2 = x[2] + D3(1e-8f);
  D3 a = dsqrt(t0 * t0 + t1 * t1 + t2 * t2);
  D3 n0 = x[0] / a, n1 = x[1] / a, n2 = x[2] / a;
  D3 h = a * 0.5f, c = dcos(h), s = dsin(h);
  D3 qx = s * n0, qy = s * n1, qz = s * n2;
  D3 qn = dsqrt(c * c + qx * qx + qy * qy + qz * qz);
  D3 R[9];
  quat_to_rot<D3>(c / qn, qx / qn, qy / qn, qz / qn, R);
  for (int k = 0; k < 3; ++k) {
    float g = 0.f;
    for (int e = 0; e < 9; ++e) g += gR[e] * R[e].d[k];
    g_theta[k] = g;
  }
}

// ------------------------------------------------------------------------------------------------ pose forward
struct SmplModel {
  int V, J, n_betas, kd, kdp, nv3p, nnz;
  const float* templ;      // [3V]
  const float* dirs;       // [kdp][nv3p]   rows: 9(J-1) pose directions, then n_betas shape directions
  const float* dirsT;      // [nv3p][kdp]
  const float* j_templ;    // [J][3]
  const float* j_dirs;     // [J][3][n_betas]
  const int* parents;      // [J]
  const int* skin_idx;     // [V][nnz]
  const float* skin_w;     // [V][nnz]
};

__global__ void __launch_bounds__(64) smpl_pose_fwd_kernel(SmplModel m, const float* __restrict__ pose, const float* __restrict__ betas,
                                                           const float* __restrict__ trans, float scale, float* __restrict__ coef,
                                                           float* __restrict__ Rout, float* __restrict__ Jout, float* __restrict__ Gout,
                                                           float* __restrict__ Aout, float* __restrict__ jtr) {
  __shared__ float sR[SM_MAXJ][9], sJ[SM_MAXJ][3], sG[SM_MAXJ][12];
  const int b = blockIdx.x, t = threadIdx.x, J = m.J;
  if (t < J) {
    float R[9];
    rodrigues(pose + ((size_t)b * J + t) * 3, R);
#pragma unroll
    for (int e = 0; e < 9; ++e) { sR[t][e] = R[e]; Rout[((size_t)b * J + t) * 9 + e] = R[e]; }
    if (t >= 1) {
#pragma unroll
      for (int e = 0; e < 9; ++e) coef[(size_t)b * m.kdp + (t - 1) * 9 + e] = R[e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
    }
    for (int c = 0; c < 3; ++c) {
      float v = m.j_templ[t * 3 + c];
      for (int k = 0; k < m.n_betas; ++k) v = fmaf(m.j_dirs[(t * 3 + c) * m.n_betas + k], betas[(size_t)b * m.n_betas + k], v);
      sJ[t][c] = v; Jout[((size_t)b * J + t) * 3 + c] = v;
    }
  }
  for (int k = t; k < m.kdp - 9 * (J - 1); k += 64)
    coef[(size_t)b * m.kdp + 9 * (J - 1) + k] = k < m.n_betas ? betas[(size_t)b * m.n_betas + k] : 0.f;
  __syncthreads();
  if (t == 0) {   // kinematic chain (smpl_layer.py:111-123): parents precede children
    for (int e = 0; e < 9; ++e) sG[0][(e / 3) * 4 + e % 3] = sR[0][e];
    for (int r = 0; r < 3; ++r) sG[0][r * 4 + 3] = sJ[0][r];
    for (int i = 1; i < J; ++i) {
      const int p = m.parents[i];
      float d[3] = {sJ[i][0] - sJ[p][0], sJ[i][1] - sJ[p][1], sJ[i][2] - sJ[p][2]};
      for (int r = 0; r < 3; ++r) {
        const float g0 = sG[p][r * 4], g1 = sG[p][r * 4 + 1], g2 = sG[p][r * 4 + 2];
        for (int c = 0; c < 3; ++c) sG[i][r * 4 + c] = g0 * sR[i][c] + g1 * sR[i][3 + c] + g2 * sR[i][6 + c];
        sG[i][r * 4 + 3] = g0 * d[0] + g1 * d[1] + g2 * d[2] + sG[p][r * 4 + 3];
      }
    }
  }
  __syncthreads();
  if (t < J) {
    float* G = Gout + ((size_t)b * J + t) * 12;
    float* A = Aout + ((size_t)b * J + t) * 12;
    for (int r = 0; r < 3; ++r) {
      const float g0 = sG[t][r * 4], g1 = sG[t][r * 4 + 1], g2 = sG[t][r * 4 + 2], g3 = sG[t][r * 4 + 3];
      G[r * 4] = g0; G[r * 4 + 1] = g1; G[r * 4 + 2] = g2; G[r * 4 + 3] = g3;
      A[r * 4] = g0; A[r * 4 + 1] = g1; A[r * 4 + 2] = g2;
      A[r * 4 + 3] = g3 - (g0 * sJ[t][0] + g1 * sJ[t][1] + g2 * sJ[t][2]);        // th_results - th_pack(G [j; 0]), :129-137
      jtr[((size_t)b * J + t) * 3 + r] = g3 * scale + trans[(size_t)b * 3 + r];
    }
  }
}

// ------------------------------------------------------------------------------------------------ pose backward
__global__ void __launch_bounds__(64) smpl_pose_bwd_kernel(SmplModel m, const float* __restrict__ pose, const float* __restrict__ Rin,
                                                           const float* __restrict__ Jin, const float* __restrict__ Gin,
                                                           const float* __restrict__ gA, const float* __restrict__ g_jtr,
                                                           const float* __restrict__ g_coef, const float* __restrict__ g_trans_skin,
                                                           float scale, float* __restrict__ g_pose, float* __restrict__ g_betas,
                                                           float* __restrict__ g_trans) {
  __shared__ float sR[SM_MAXJ][9], sJ[SM_MAXJ][3], sG[SM_MAXJ][12];
  __shared__ float gG[SM_MAXJ][12], gR[SM_MAXJ][9], gJ[SM_MAXJ][3];
  const int b = blockIdx.x, t = threadIdx.x, J = m.J;
  if (t < J) {
    for (int e = 0; e < 9; ++e) sR[t][e] = Rin[((size_t)b * J + t) * 9 + e];
    for (int e = 0; e < 3; ++e) sJ[t][e] = Jin[((size_t)b * J + t) * 3 + e];
    for (int e = 0; e < 12; ++e) sG[t][e] = Gin[((size_t)b * J + t) * 12 + e];
    // A.R = G.R ; A.t = G.t - G.R J ; jtr = G.t * scale + trans
    const float* a = gA + ((size_t)b * J + t) * 12;
    float gt[3] = {a[3], a[7], a[11]};
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) gG[t][r * 4 + c] = a[r * 4 + c] - gt[r] * sJ[t][c];
      gG[t][r * 4 + 3] = gt[r] + (g_jtr ? scale * g_jtr[((size_t)b * J + t) * 3 + r] : 0.f);
    }
    for (int c = 0; c < 3; ++c) gJ[t][c] = -(sG[t][c] * gt[0] + sG[t][4 + c] * gt[1] + sG[t][8 + c] * gt[2]);
  }
  __syncthreads();
  if (t == 0) {   // reverse kinematic chain: children before parents
    for (int i = J - 1; i >= 1; --i) {
      const int p = m.parents[i];
      float d[3] = {sJ[i][0] - sJ[p][0], sJ[i][1] - sJ[p][1], sJ[i][2] - sJ[p][2]};
      // G_i.R = G_p.R R_i : gR_i = G_p.R^T gG_i.R ; gG_p.R += gG_i.R R_i^T
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          gR[i][r * 3 + c] = sG[p][r] * gG[i][c] + sG[p][4 + r] * gG[i][4 + c] + sG[p][8 + r] * gG[i][8 + c];
          gG[p][r * 4 + c] += gG[i][r * 4] * sR[i][c * 3] + gG[i][r * 4 + 1] * sR[i][c * 3 + 1] + gG[i][r * 4 + 2] * sR[i][c * 3 + 2];
        }
      // G_i.t = G_p.R d + G_p.t
      for (int r = 0; r < 3; ++r) {
        const float g = gG[i][r * 4 + 3];
        for (int c = 0; c < 3; ++c) gG[p][r * 4 + c] += g * d[c];
        gG[p][r * 4 + 3] += g;
      }
      for (int c = 0; c < 3; ++c) {
        const float gd = sG[p][c] * gG[i][3] + sG[p][4 + c] * gG[i][7] + sG[p][8 + c] * gG[i][11];
        gJ[i][c] += gd; gJ[p][c] -= gd;
      }
    }
    for (int e = 0; e < 9; ++e) gR[0][e] = gG[0][(e / 3) * 4 + e % 3];
    for (int c = 0; c < 3; ++c) gJ[0][c] += gG[0][c * 4 + 3];
  }
  __syncthreads();
  if (t < J) {
    float g[9];
    for (int e = 0; e < 9; ++e) g[e] = gR[t][e] + ((t >= 1 && g_coef) ? g_coef[(size_t)b * m.kdp + (t - 1) * 9 + e] : 0.f);
    float gth[3];
    rodrigues_bwd(pose + ((size_t)b * J + t) * 3, g, gth);
    for (int k = 0; k < 3; ++k) g_pose[((size_t)b * J + t) * 3 + k] = gth[k];
  }
  if (t < m.n_betas) {
    float g = g_coef ? g_coef[(size_t)b * m.kdp + 9 * (J - 1) + t] : 0.f;
    for (int j = 0; j < J; ++j)
      for (int c = 0; c < 3; ++c) g = fmaf(m.j_dirs[(j * 3 + c) * m.n_betas + t], gJ[j][c], g);
    g_betas[(size_t)b * m.n_betas + t] = g;
  }
  if (t < 3) {
    float g = g_trans_skin ? g_trans_skin[(size_t)b * 3 + t] : 0.f;
    if (g_jtr) for (int j = 0; j < J; ++j) g += g_jtr[((size_t)b * J + j) * 3 + t];
    g_trans[(size_t)b * 3 + t] = g;
  }
}

// ------------------------------------------------------------------------------------------------ tiled FFMA GEMM
// C[M,N] = A[M,K] Bm[K,N] (+ bias[N]) (+ add[M,N]); gridDim.z > 1 splits K and accumulates with atomics into a zeroed C.
// M is the number of frames (tens to a few hundred), N / K the 20 670 vertex coordinates / 472 blend coefficients: the tile height GM is picked
// per launch (sgemm_gm; thread = TM x 4 outputs, TM = GM / 8), and the next K step's operands are fetched into registers while the current one is
// multiplied (the first version waited for every 16-deep step's global loads in turn: 35 TFLOP/s fp32).
constexpr int GN = 128, GK = 16;
template <int GM>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
                                                    float* __restrict__ C, int ldc, int M, int N, int K, int k_per_split,
                                                    const float* __restrict__ bias, const float* __restrict__ add, int ldadd) {
  constexpr int TM = GM / 8;                                // rows per thread
  constexpr int NA = (GM * 4 + 255) / 256;                  // float4 loads of the A tile per thread
  static_assert(GM % 8 == 0 && TM % 2 == 0, "tile height");
  __shared__ __align__(16) float sA[GK][GM + 4];
  __shared__ __align__(16) float sB[GK][GN];
  const int tid = threadIdx.x, tx = tid % 32, ty = tid / 32;
  const int n0 = blockIdx.x * GN, m0 = blockIdx.y * GM;
  const int kb = blockIdx.z * k_per_split, ke = min(K, kb + k_per_split);
  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float4 av[NA], bv[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < NA; ++r) {
      const int idx = tid + r * 256, a_row = idx / 4, a_k4 = idx % 4;
      av[r] = make_float4(0, 0, 0, 0);
      if (idx < GM * 4 && m0 + a_row < M && k0 + a_k4 * 4 < ke) av[r] = ld4(A + (size_t)(m0 + a_row) * lda + k0 + a_k4 * 4);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = tid / 32 + r * 8, c = n0 + tx * 4;
      bv[r] = make_float4(0, 0, 0, 0);
      if (k0 + row < ke && c < N) bv[r] = ld4(Bm + (size_t)(k0 + row) * ldb + c);
    }
  };
  fetch(kb);
  for (int k0 = kb; k0 < ke; k0 += GK) {
    __syncthreads();                                        // the previous step's reads of sA / sB are done
#pragma unroll
    for (int r = 0; r < NA; ++r) {
      const int idx = tid + r * 256, a_row = idx / 4, a_k4 = idx % 4;
      if (idx < GM * 4) { sA[a_k4 * 4 + 0][a_row] = av[r].x; sA[a_k4 * 4 + 1][a_row] = av[r].y; sA[a_k4 * 4 + 2][a_row] = av[r].z; sA[a_k4 * 4 + 3][a_row] = av[r].w; }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) *reinterpret_cast<float4*>(&sB[tid / 32 + r * 8][tx * 4]) = bv[r];
    __syncthreads();
    if (k0 + GK < ke) fetch(k0 + GK);                       // in flight during the products below
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const float4 b4 = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
      float a[TM];
#pragma unroll
      for (int i = 0; i < TM; i += 2) {
        const float2 a2 = *reinterpret_cast<const float2*>(&sA[kk][ty * TM + i]);
        a[i] = a2.x; a[i + 1] = a2.y;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        acc[i][0] = fmaf(a[i], b4.x, acc[i][0]); acc[i][1] = fmaf(a[i], b4.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], b4.z, acc[i][2]); acc[i][3] = fmaf(a[i], b4.w, acc[i][3]);
      }
    }
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int mrow = m0 + ty * TM + i;
    if (mrow >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (!split || blockIdx.z == 0) {
        if (bias) v += bias[n];
        if (add) v += add[(size_t)mrow * ldadd + n];
      }
      if (split) atomicAdd(C + (size_t)mrow * ldc + n, v); else C[(size_t)mrow * ldc + n] = v;
    }
  }
}

// tile height for M rows: the one with the shortest makespan -- waves of CTAs over the SMs x rows per tile (a 96-row tile would read the Bm panel
// once, but 96 frames x 162 column tiles are 162 CTAs = two rounds on 148 SMs, while 486 tiles of 32 rows are four rounds of a third of the work);
// ties go to the taller tile (fewer panel reads)
static int sgemm_gm(int M, int ctas_per_row_tile) {
  static int sms = 0;
  if (!sms) { cudaDeviceProp p; int dev = 0; cudaGetDevice(&dev); sms = (cudaGetDeviceProperties(&p, dev) == cudaSuccess && p.multiProcessorCount > 0) ? p.multiProcessorCount : 148; }
  const int cand[2] = {48, 32};                            // (64- and 96-row tiles never win at SMPL-H's 162 column tiles: not instantiated)
  int best = 32;
  long long best_cost = -1;
  for (int gm : cand) {
    const long long ctas = (long long)ctas_per_row_tile * ceil_div(M, gm);
    const long long cost = ((ctas + sms - 1) / sms) * gm;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = gm; }
  }
  return best;
}
static void sgemm_launch(int n_tiles, int k_splits, cudaStream_t s, const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int M, int N,
                         int K, int k_per_split, const float* bias, const float* add, int ldadd) {
  const int gm = sgemm_gm(M, n_tiles * k_splits);
  const dim3 grid(n_tiles, ceil_div(M, gm), k_splits);
  if (gm == 32) sgemm_kernel<32><<<grid, 256, 0, s>>>(A, lda, Bm, ldb, C, ldc, M, N, K, k_per_split, bias, add, ldadd);
  else sgemm_kernel<48><<<grid, 256, 0, s>>>(A, lda, Bm, ldb, C, ldc, M, N, K, k_per_split, bias, add, ldadd);
}

__global__ void vec_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ c, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) c[i] = a[i] + b[i];
}

// ------------------------------------------------------------------------------------------------ skinning
__global__ void __launch_bounds__(256) smpl_skin_fwd_kernel(SmplModel m, const float* __restrict__ A, const float* __restrict__ v_posed,
                                                            const float* __restrict__ trans, float scale, float* __restrict__ verts) {
  __shared__ float sA[SM_MAXJ * 12];
  const int b = blockIdx.y, v = blockIdx.x * 256 + threadIdx.x;
  for (int i = threadIdx.x; i < m.J * 12; i += 256) sA[i] = A[(size_t)b * m.J * 12 + i];
  __syncthreads();
  if (v >= m.V) return;
  float T[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) T[e] = 0.f;
  for (int k = 0; k < m.nnz; ++k) {
    const float w = m.skin_w[(size_t)v * m.nnz + k];
    const float* a = sA + m.skin_idx[(size_t)v * m.nnz + k] * 12;
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = fmaf(w, a[e], T[e]);
  }
  const float* p = v_posed + ((size_t)b * m.V + v) * 3;
  const float x = p[0], y = p[1], z = p[2];
  float* o = verts + ((size_t)b * m.V + v) * 3;
#pragma unroll
  for (int r = 0; r < 3; ++r) o[r] = (T[r * 4] * x + T[r * 4 + 1] * y + T[r * 4 + 2] * z + T[r * 4 + 3]) * scale + trans[(size_t)b * 3 + r];
}

__global__ void __launch_bounds__(256) smpl_skin_bwd_kernel(SmplModel m, const float* __restrict__ A, const float* __restrict__ v_posed,
                                                            const float* __restrict__ g_verts, float scale, float* __restrict__ g_vposed,
                                                            int ld_gvp, float* __restrict__ gA, float* __restrict__ g_trans) {
  __shared__ float sA[SM_MAXJ * 12], sgA[SM_MAXJ * 12], sgt[3];
  const int b = blockIdx.y, v = blockIdx.x * 256 + threadIdx.x;
  for (int i = threadIdx.x; i < m.J * 12; i += 256) { sA[i] = A[(size_t)b * m.J * 12 + i]; sgA[i] = 0.f; }
  if (threadIdx.x < 3) sgt[threadIdx.x] = 0.f;
  __syncthreads();
  float gx = 0.f, gy = 0.f, gz = 0.f;
  float sx = 0.f, sy = 0.f, sz = 0.f, ph[4] = {0.f, 0.f, 0.f, 1.f};
  float T[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) T[e] = 0.f;
  if (v < m.V) {
    const float* g = g_verts + ((size_t)b * m.V + v) * 3;
    gx = g[0]; gy = g[1]; gz = g[2];
    sx = gx * scale; sy = gy * scale; sz = gz * scale;
    const float* p = v_posed + ((size_t)b * m.V + v) * 3;
    ph[0] = p[0]; ph[1] = p[1]; ph[2] = p[2];
  }
  // d/dA[j] += w (s (x) ph): neighbouring vertices of a real mesh are skinned to the same joints, so the lanes of a warp mostly hit the
  // same 12 shared-memory words -- fp32 shared atomics are compare-and-swap loops, 32 lanes on one address take 32 rounds each (this kernel
  // took 0.14 ms for 96 frames with random skinning and 0.47 ms with a body whose skinning follows the surface).  Lanes are grouped by joint
  // (match.any); a group of 8 or more is summed with a butterfly over the whole warp (the other lanes add zeros) and ONE lane does the 12
  // atomics, smaller groups keep their direct atomics.
  const int lane = threadIdx.x & 31;
  for (int k = 0; k < m.nnz; ++k) {
    float w = 0.f;
    int j = -1 - lane;                                     // inactive lanes: a private key, no group
    if (v < m.V) {
      w = m.skin_w[(size_t)v * m.nnz + k];
      const int jj = m.skin_idx[(size_t)v * m.nnz + k];
      const float* a = sA + jj * 12;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) T[r * 3 + c] = fmaf(w, a[r * 4 + c], T[r * 3 + c]);
      if (w != 0.f) j = jj;
    }
    const unsigned grp = __match_any_sync(0xffffffffu, j);
    const bool big = j >= 0 && __popc(grp) >= 8;
    unsigned todo = __ballot_sync(0xffffffffu, big && lane == __ffs(grp) - 1);
    while (todo) {                                         // (warp-uniform)
      const int leader = __ffs(todo) - 1;
      todo &= todo - 1;
      const int jl = __shfl_sync(0xffffffffu, j, leader);
      const float wl = (j == jl) ? w : 0.f;
      float* ga = sgA + jl * 12;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float ws = wl * (r == 0 ? sx : r == 1 ? sy : sz);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t = ws * ph[c];
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0) atomicAdd(ga + r * 4 + c, t);
        }
      }
    }
    if (j >= 0 && !big) {
      float* ga = sgA + j * 12;
#pragma unroll
      for (int c = 0; c < 4; ++c) { atomicAdd(ga + c, w * sx * ph[c]); atomicAdd(ga + 4 + c, w * sy * ph[c]); atomicAdd(ga + 8 + c, w * sz * ph[c]); }
    }
  }
  if (v < m.V) {
    float* o = g_vposed + (size_t)b * ld_gvp + (size_t)v * 3;
    o[0] = T[0] * sx + T[3] * sy + T[6] * sz;
    o[1] = T[1] * sx + T[4] * sy + T[7] * sz;
    o[2] = T[2] * sx + T[5] * sy + T[8] * sz;
  }
  // sum of g_verts over the block -> g_trans
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) { gx += __shfl_xor_sync(0xffffffffu, gx, o); gy += __shfl_xor_sync(0xffffffffu, gy, o); gz += __shfl_xor_sync(0xffffffffu, gz, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&sgt[0], gx); atomicAdd(&sgt[1], gy); atomicAdd(&sgt[2], gz); }
  __syncthreads();
  for (int i = threadIdx.x; i < m.J * 12; i += 256) if (sgA[i] != 0.f) atomicAdd(gA + (size_t)b * m.J * 12 + i, sgA[i]);
  if (threadIdx.x < 3) atomicAdd(g_trans + (size_t)b * 3 + threadIdx.x, sgt[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------ landmark regressors
// out[b][l] = sum_{(v,w) in row l} w * verts[b][v]   (CSR by landmark; one warp per (frame, landmark))
__global__ void __launch_bounds__(128) landmarks_fwd_kernel(const float* __restrict__ verts, int V, const int* __restrict__ rowptr,
                                                            const int* __restrict__ col, const float* __restrict__ val, int L,
                                                            float* __restrict__ out) {
  const int b = blockIdx.y, l = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (l >= L) return;
  float x = 0.f, y = 0.f, z = 0.f;
  for (int i = rowptr[l] + lane; i < rowptr[l + 1]; i += 32) {
    const float w = val[i];
    const float* p = verts + ((size_t)b * V + col[i]) * 3;
    x = fmaf(w, p[0], x); y = fmaf(w, p[1], y); z = fmaf(w, p[2], z);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) { x += __shfl_xor_sync(0xffffffffu, x, o); y += __shfl_xor_sync(0xffffffffu, y, o); z += __shfl_xor_sync(0xffffffffu, z, o); }
  if (lane == 0) { float* o = out + ((size_t)b * L + l) * 3; o[0] = x; o[1] = y; o[2] = z; }
}

// g_verts[b][v] += w * g_out[b][l]  (scatter; ~8.5k atomics per frame for body-25)
__global__ void __launch_bounds__(128) landmarks_bwd_kernel(const float* __restrict__ g_out, int V, const int* __restrict__ rowptr,
                                                            const int* __restrict__ col, const float* __restrict__ val, int L,
                                                            float* __restrict__ g_verts) {
  const int b = blockIdx.y, l = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (l >= L) return;
  const float* g = g_out + ((size_t)b * L + l) * 3;
  const float gx = g[0], gy = g[1], gz = g[2];
  for (int i = rowptr[l] + lane; i < rowptr[l + 1]; i += 32) {
    const float w = val[i];
    float* p = g_verts + ((size_t)b * V + col[i]) * 3;
    atomicAdd(p, w * gx); atomicAdd(p + 1, w * gy); atomicAdd(p + 2, w * gz);
  }
}

static int check_model(const vt_smpl_model* m, const char* who) {
  VT_CHECK_ARG(m && m->J >= 1 && m->J <= SM_MAXJ, "%s: joint count %d exceeds %d", who, m ? m->J : -1, SM_MAXJ);
  VT_CHECK_ARG(m->kd == 9 * (m->J - 1) + m->n_betas && m->kdp >= m->kd && m->kdp % 4 == 0 && m->nv3p >= 3 * m->V && m->nv3p % 4 == 0,
               "%s: inconsistent blend-shape dimensions", who);
  return 0;
}

static SmplModel to_dev(const vt_smpl_model* m) {
  SmplModel d;
  d.V = m->V; d.J = m->J; d.n_betas = m->n_betas; d.kd = m->kd; d.kdp = m->kdp; d.nv3p = m->nv3p; d.nnz = m->nnz;
  d.templ = m->templ; d.dirs = m->dirs; d.dirsT = m->dirsT; d.j_templ = m->j_templ; d.j_dirs = m->j_dirs;
  d.parents = m->parents; d.skin_idx = m->skin_idx; d.skin_w = m->skin_w;
  return d;
}

}  // namespace vt

using namespace vt;

extern "C" {

int vt_smpl_fwd(const vt_smpl_model* model, const float* pose, const float* betas, const float* trans, const float* offsets,
                float scale, int B, float* coef, float* R, float* J, float* G, float* A, float* naked, float* v_posed,
                float* verts, float* jtr, void* stream) {
  if (int rc = check_model(model, "vt_smpl_fwd")) return rc;
  if (B <= 0) return 0;
  SmplModel m = to_dev(model);
  cudaStream_t s = (cudaStream_t)stream;
  smpl_pose_fwd_kernel<<<B, 64, 0, s>>>(m, pose, betas, trans, scale, coef, R, J, G, A, jtr);
  VT_CHECK_LAUNCH("vt_smpl_fwd(pose)");
  const int N = 3 * m.V;
  // naked = template + coef x dirs ; v_posed = naked + offsets (smpl_layer.py:98-106)
  sgemm_launch(ceil_div(N, GN), 1, s, coef, m.kdp, m.dirs, m.nv3p, naked, N, B, N, m.kdp, m.kdp, m.templ, nullptr, 0);
  VT_CHECK_LAUNCH("vt_smpl_fwd(blend)");
  if (offsets) {
    const size_t total = (size_t)B * N;
    vec_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(naked, offsets, v_posed, total);
    VT_CHECK_LAUNCH("vt_smpl_fwd(offsets)");
  }
  dim3 g2(ceil_div(m.V, 256), B);
  smpl_skin_fwd_kernel<<<g2, 256, 0, s>>>(m, A, offsets ? v_posed : naked, trans, scale, verts);
  VT_CHECK_LAUNCH("vt_smpl_fwd(skin)");
  return 0;
}

int vt_smpl_bwd(const vt_smpl_model* model, const float* pose, const float* R, const float* J, const float* G, const float* A,
                const float* v_posed, const float* g_verts, const float* g_jtr, float scale, int B, float* g_vposed, float* gA,
                float* g_coef, float* g_trans_skin, float* g_pose, float* g_betas, float* g_trans, void* stream) {
  if (int rc = check_model(model, "vt_smpl_bwd")) return rc;
  if (B <= 0) return 0;
  SmplModel m = to_dev(model);
  cudaStream_t s = (cudaStream_t)stream;
  // scratch the caller provides: g_vposed [B][nv3p], gA [B][J][12], g_coef [B][kdp], g_trans_skin [B][3]  (zeroed here)
  cudaError_t e = cudaMemsetAsync(gA, 0, (size_t)B * m.J * 12 * sizeof(float), s);
  if (e == cudaSuccess) e = cudaMemsetAsync(g_coef, 0, (size_t)B * m.kdp * sizeof(float), s);
  if (e == cudaSuccess) e = cudaMemsetAsync(g_trans_skin, 0, (size_t)B * 3 * sizeof(float), s);
  if (e == cudaSuccess && g_verts) e = cudaMemsetAsync(g_vposed, 0, (size_t)B * m.nv3p * sizeof(float), s);
  if (e != cudaSuccess) return cuda_fail(e, "vt_smpl_bwd memset");
  if (g_verts) {
    dim3 g2(ceil_div(m.V, 256), B);
    smpl_skin_bwd_kernel<<<g2, 256, 0, s>>>(m, A, v_posed, g_verts, scale, g_vposed, m.nv3p, gA, g_trans_skin);
    VT_CHECK_LAUNCH("vt_smpl_bwd(skin)");
    const int ksplit = 512;
    sgemm_launch(ceil_div(m.kdp, GN), ceil_div(m.nv3p, ksplit), s, g_vposed, m.nv3p, m.dirsT, m.kdp, g_coef, m.kdp, B, m.kdp, m.nv3p, ksplit, nullptr, nullptr, 0);
    VT_CHECK_LAUNCH("vt_smpl_bwd(blend)");
  }
  smpl_pose_bwd_kernel<<<B, 64, 0, s>>>(m, pose, R, J, G, gA, g_jtr, g_verts ? g_coef : nullptr, g_trans_skin, scale, g_pose, g_betas, g_trans);
  VT_CHECK_LAUNCH("vt_smpl_bwd(pose)");
  return 0;
}

int vt_landmarks_fwd(const float* verts, int B, int V, const int* rowptr, const int* col, const float* val, int L, float* out,
                     void* stream) {
  if (B <= 0 || L <= 0) return 0;
  dim3 grid(ceil_div(L, 4), B);
  landmarks_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(verts, V, rowptr, col, val, L, out);
  VT_CHECK_LAUNCH("vt_landmarks_fwd");
  return 0;
}

int vt_landmarks_bwd(const float* g_out, int B, int V, const int* rowptr, const int* col, const float* val, int L, float* g_verts,
                     void* stream) {
  if (B <= 0 || L <= 0) return 0;
  dim3 grid(ceil_div(L, 4), B);
  landmarks_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(g_out, V, rowptr, col, val, L, g_verts);
  VT_CHECK_LAUNCH("vt_landmarks_bwd");
  return 0;
}

}  // extern "C"
