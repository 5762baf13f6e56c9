// Small geometric operators of the joint human-object optimisation:
//  * SO(3) projection of a free 3x3 matrix, R = U diag(1, 1, det(U V^T)) V^T (recon/recon_fit_base.py:178-199), forward and
//    analytic backward, one thread per matrix (the reference calls torch.svd -> cuSOLVER plus det / cat / matmul, 1550x per batch);
//  * bidirectional squared-L2 Chamfer distance between ragged lists of small point clouds, as
//    pytorch3d.loss.chamfer_distance(Pointclouds, Pointclouds) with its defaults computes it (point mean, batch mean, sum of the
//    two directions) -- recon/recon_fit_trivis_full.py:452-456.  Brute-force nearest neighbour: the contact sets are tens to a few
//    hundred points per (frame, body part) pair.
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

// ------------------------------------------------------------------------------------------------ SO(3) projection
// The maximiser of tr(R^T M) over SO(3) equals U diag(1,1,det(UV^T)) V^T.  It is found as the dominant eigenvector (a unit
// quaternion) of Horn's symmetric 4x4 matrix N(M), by cyclic Jacobi sweeps in double precision.
__device__ void jacobi_eig4(double (&A)[4][4], double (&V)[4][4]) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 24; ++sweep) {
    double off = 0;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) off += A[p][q] * A[p][q];
    if (off < 1e-30) break;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
        for (int k = 0; k < 4; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < 4; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
      }
  }
}

__device__ void so3_project(const float* M, double (&R)[9]) {
  const double m00 = M[0], m01 = M[1], m02 = M[2], m10 = M[3], m11 = M[4], m12 = M[5], m20 = M[6], m21 = M[7], m22 = M[8];
  // tr(R(q)^T M) = q^T N q
  double N[4][4] = {{m00 + m11 + m22, m21 - m12, m02 - m20, m10 - m01},
                    {m21 - m12, m00 - m11 - m22, m01 + m10, m02 + m20},
                    {m02 - m20, m01 + m10, -m00 + m11 - m22, m12 + m21},
                    {m10 - m01, m02 + m20, m12 + m21, -m00 - m11 + m22}};
  double V[4][4];
  jacobi_eig4(N, V);
  int best = 0;
  for (int i = 1; i < 4; ++i) if (N[i][i] > N[best][best]) best = i;
  double w = V[0][best], x = V[1][best], y = V[2][best], z = V[3][best];
  const double n = sqrt(w * w + x * x + y * y + z * z);
  w /= n; x /= n; y /= n; z /= n;
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z);           R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z);           R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y);           R[7] = 2 * (y * z + w * x);           R[8] = w * w - x * x - y * y + z * z;
}

__global__ void so3_fwd_kernel(const float* __restrict__ M, int B, float* __restrict__ Rout) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double R[9];
  so3_project(M + (size_t)b * 9, R);
  for (int e = 0; e < 9; ++e) Rout[(size_t)b * 9 + e] = (float)R[e];
}

// init_object_orientation (recon/recon_fit_base.py:202-216, recon/pca_util.py:59-72): rot = (S^T S)^-1 S^T T, then the SO(3) projection of
// rot (+ 1e-4 noise when given: decopose_axis).  src_stride 0 repeats one template axis set for every frame.
__global__ void pca_orientation_kernel(const float* __restrict__ tgt, const float* __restrict__ src, int src_stride, const float* __restrict__ noise, int B,
                                       float* __restrict__ Rout) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* S = src + (size_t)b * src_stride;
  const float* T = tgt + (size_t)b * 9;
  double G[9], Gi[9], P[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double g = 0;
      for (int k = 0; k < 3; ++k) g += (double)S[k * 3 + i] * (double)S[k * 3 + j];
      G[i * 3 + j] = g;
    }
  const double det = G[0] * (G[4] * G[8] - G[5] * G[7]) - G[1] * (G[3] * G[8] - G[5] * G[6]) + G[2] * (G[3] * G[7] - G[4] * G[6]);
  const double id = 1.0 / det;                       // singular template axes: inf / NaN, as torch.inverse would raise
  Gi[0] = (G[4] * G[8] - G[5] * G[7]) * id; Gi[1] = (G[2] * G[7] - G[1] * G[8]) * id; Gi[2] = (G[1] * G[5] - G[2] * G[4]) * id;
  Gi[3] = (G[5] * G[6] - G[3] * G[8]) * id; Gi[4] = (G[0] * G[8] - G[2] * G[6]) * id; Gi[5] = (G[2] * G[3] - G[0] * G[5]) * id;
  Gi[6] = (G[3] * G[7] - G[4] * G[6]) * id; Gi[7] = (G[1] * G[6] - G[0] * G[7]) * id; Gi[8] = (G[0] * G[4] - G[1] * G[3]) * id;
  for (int i = 0; i < 3; ++i)                        // pseudo-inverse P = G^-1 S^T
    for (int j = 0; j < 3; ++j) {
      double v = 0;
      for (int k = 0; k < 3; ++k) v += Gi[i * 3 + k] * (double)S[j * 3 + k];
      P[i * 3 + j] = v;
    }
  float M[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double v = 0;
      for (int k = 0; k < 3; ++k) v += P[i * 3 + k] * (double)T[k * 3 + j];
      M[i * 3 + j] = (float)v + (noise ? 1e-4f * noise[(size_t)b * 9 + i * 3 + j] : 0.f);
    }
  double R[9];
  so3_project(M, R);
  for (int e = 0; e < 9; ++e) Rout[(size_t)b * 9 + e] = (float)R[e];
}

// dL/dM = R [c]x with c = (tr(P) I - P)^-1 b, P = sym(R^T M), b = axial(R^T G - G^T R)
__global__ void so3_bwd_kernel(const float* __restrict__ M, const float* __restrict__ G, int B, float* __restrict__ gM) {
  const int bi = blockIdx.x * blockDim.x + threadIdx.x;
  if (bi >= B) return;
  double R[9], P[9], A[9];
  const float* m = M + (size_t)bi * 9;
  const float* g = G + (size_t)bi * 9;
  so3_project(m, R);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double p = 0, a = 0;
      for (int k = 0; k < 3; ++k) { p += R[k * 3 + i] * (double)m[k * 3 + j]; a += R[k * 3 + i] * (double)g[k * 3 + j]; }
      P[i * 3 + j] = p; A[i * 3 + j] = a;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j) { const double s = 0.5 * (P[i * 3 + j] + P[j * 3 + i]); P[i * 3 + j] = P[j * 3 + i] = s; }
  const double b[3] = {A[7] - A[5], A[2] - A[6], A[3] - A[1]};
  const double tr = P[0] + P[4] + P[8];
  double K[9];
  for (int i = 0; i < 9; ++i) K[i] = -P[i];
  K[0] += tr; K[4] += tr; K[8] += tr;
  // c = K^-1 b by cofactors (K is symmetric)
  const double c00 = K[4] * K[8] - K[5] * K[7], c01 = K[5] * K[6] - K[3] * K[8], c02 = K[3] * K[7] - K[4] * K[6];
  const double det = K[0] * c00 + K[1] * c01 + K[2] * c02;
  const double c11 = K[0] * K[8] - K[2] * K[6], c12 = K[1] * K[6] - K[0] * K[7], c22 = K[0] * K[4] - K[1] * K[3];
  const double inv = 1.0 / det;
  const double c[3] = {(c00 * b[0] + c01 * b[1] + c02 * b[2]) * inv, (c01 * b[0] + c11 * b[1] + c12 * b[2]) * inv,
                       (c02 * b[0] + c12 * b[1] + c22 * b[2]) * inv};
  const double C[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += R[i * 3 + k] * C[k * 3 + j];
      gM[(size_t)bi * 9 + i * 3 + j] = (float)s;
    }
}

// ------------------------------------------------------------------------------------------------ ragged Chamfer
// x points of pair n are rows [xo[n], xo[n+1]) of the packed [sum, 3] array.
// loss += (1/N) * ( mean_i min_j |x_i - y_j|^2 + mean_j min_i |y_j - x_i|^2 )
// Grid (N, CH_SPLIT): the query points of a pair -- both directions, chunks of 128 -- are dealt round-robin to the CH_SPLIT CTAs of the pair (one CTA per
// pair left 52 SMs idle and 4 warps on the others: 0.48 ms for 96 pairs of 296 x 3000 points).  The other cloud is staged through shared memory in
// tiles of CH_TILE points (one broadcast LDS.128 per candidate); every thread scans the candidates in increasing order with a strict `<`, i.e. the
// first nearest neighbour, as before.
constexpr int CH_SPLIT = 8, CH_TILE = 1024;
__global__ void __launch_bounds__(128) chamfer_fwd_kernel(const float* __restrict__ x, const int* __restrict__ xo, const float* __restrict__ y,
                                                          const int* __restrict__ yo, int N, int* __restrict__ nn_x, int* __restrict__ nn_y,
                                                          float* __restrict__ loss) {
  __shared__ float4 tile[CH_TILE];
  const int n = blockIdx.x;
  const int x0 = xo[n], x1 = xo[n + 1], y0 = yo[n], y1 = yo[n + 1];
  const int cx = (x1 - x0 + 127) >> 7, cy = (y1 - y0 + 127) >> 7;       // chunks of 128 query points per direction
  float part = 0.f;
  for (int ch = blockIdx.y; ch < cx + cy; ch += gridDim.y) {
    const int dir = ch < cx ? 0 : 1;
    const float* a = dir == 0 ? x : y;  const float* b = dir == 0 ? y : x;
    const int a0 = dir == 0 ? x0 : y0, a1 = dir == 0 ? x1 : y1, b0 = dir == 0 ? y0 : x0, b1 = dir == 0 ? y1 : x1;
    int* nn = dir == 0 ? nn_x : nn_y;
    const float scale = (a1 > a0) ? 1.f / (float)(a1 - a0) : 0.f;
    const int i = a0 + (dir == 0 ? ch : ch - cx) * 128 + threadIdx.x;
    const bool live = i < a1;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (live) { px = a[i * 3]; py = a[i * 3 + 1]; pz = a[i * 3 + 2]; }
    float best = 3.4e38f; int bj = -1;
    for (int t0 = b0; t0 < b1; t0 += CH_TILE) {
      const int nt = min(CH_TILE, b1 - t0);
      __syncthreads();                                     // the previous tile (or chunk) has been read by every thread
      for (int j = threadIdx.x; j < nt; j += 128) tile[j] = make_float4(b[(t0 + j) * 3], b[(t0 + j) * 3 + 1], b[(t0 + j) * 3 + 2], 0.f);
      __syncthreads();
      if (live) {
#pragma unroll 4
        for (int j = 0; j < nt; ++j) {
          const float4 q = tile[j];
          const float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
          const float d = dx * dx + dy * dy + dz * dz;
          if (d < best) { best = d; bj = t0 + j; }
        }
      }
    }
    if (live) {
      nn[i] = bj;
      if (bj >= 0) part += best * scale;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0 && part != 0.f) atomicAdd(loss, part / (float)N);
}

__global__ void __launch_bounds__(128) chamfer_bwd_kernel(const float* __restrict__ x, const int* __restrict__ xo, const float* __restrict__ y,
                                                          const int* __restrict__ yo, int N, const int* __restrict__ nn_x,
                                                          const int* __restrict__ nn_y, const float* __restrict__ g_loss,
                                                          float* __restrict__ gx, float* __restrict__ gy) {
  const int n = blockIdx.x;
  const float g = g_loss[0] / (float)N;
  for (int dir = 0; dir < 2; ++dir) {
    const float* a = dir == 0 ? x : y;  const float* b = dir == 0 ? y : x;
    float* ga = dir == 0 ? gx : gy;  float* gb = dir == 0 ? gy : gx;
    const int a0 = dir == 0 ? xo[n] : yo[n], a1 = dir == 0 ? xo[n + 1] : yo[n + 1];
    const int* nn = dir == 0 ? nn_x : nn_y;
    const float scale = (a1 > a0) ? 2.f * g / (float)(a1 - a0) : 0.f;
    for (int i = a0 + threadIdx.x; i < a1; i += blockDim.x) {
      const int j = nn[i];
      if (j < 0) continue;
      for (int c = 0; c < 3; ++c) {
        const float d = (a[i * 3 + c] - b[j * 3 + c]) * scale;
        atomicAdd(ga + i * 3 + c, d);
        atomicAdd(gb + j * 3 + c, -d);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ evaluation Chamfer (dense clouds)
// recon/eval/chamfer_distance.py:10-52: mean Euclidean (NOT squared) nearest-neighbour distance, evaluated on 10 000 surface samples per
// mesh (recon/eval/evaluate.py:43,126-154).  One thread per query point, the other cloud streamed through shared memory in 256-point tiles;
// blockIdx.z = direction (0: every x_i against y, 1: every y_j against x).  Writes the per-point distances; the means stay with the caller.
__global__ void __launch_bounds__(256) nn_dist_kernel(const float* __restrict__ x, int nx, const float* __restrict__ y, int ny,
                                                      float* __restrict__ dx_out, float* __restrict__ dy_out) {
  __shared__ float sp[256 * 3];
  const int b = blockIdx.y, dir = blockIdx.z;
  const float* q = (dir == 0 ? x + (size_t)b * nx * 3 : y + (size_t)b * ny * 3);
  const float* r = (dir == 0 ? y + (size_t)b * ny * 3 : x + (size_t)b * nx * 3);
  const int nq = dir == 0 ? nx : ny, nr = dir == 0 ? ny : nx;
  float* out = dir == 0 ? dx_out + (size_t)b * nx : dy_out + (size_t)b * ny;
  if ((int)(blockIdx.x * 256) >= nq) return;                 // whole block beyond this direction's cloud
  const int i = blockIdx.x * 256 + threadIdx.x;
  const bool live = i < nq;
  const float px = live ? q[i * 3] : 0.f, py = live ? q[i * 3 + 1] : 0.f, pz = live ? q[i * 3 + 2] : 0.f;
  float best = 3.4e38f;
  for (int j0 = 0; j0 < nr; j0 += 256) {
    const int cnt = min(256, nr - j0);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * 3; t += 256) sp[t] = r[(size_t)j0 * 3 + t];
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const float ddx = px - sp[j * 3], ddy = py - sp[j * 3 + 1], ddz = pz - sp[j * 3 + 2];
      best = fminf(best, ddx * ddx + ddy * ddy + ddz * ddz);
    }
  }
  if (live) out[i] = sqrtf(best);
}

// ------------------------------------------------------------------------------------------------ Procrustes (similarity) alignment
// compute_transform (recon/eval/pose_utils.py:153-198): scale, rotation and translation that take point set S1 closest to S2 -- the
// alignment ProcrusteAlign / VideoPackedEvaluator compute once per window of frames from the combined SMPL + object vertices
// (pose_utils.py:22-69, evalvideo_packed.py:108-131).  The reference centres, forms K = X1 X2^T, takes a 3x3 SVD and fixes the
// determinant; R = V Z U^T is the maximiser of tr(R K) over SO(3), i.e. so3_project(K^T) above.  Raw moments are accumulated in
// double precision (one pass), the 3x3 problem is solved by one thread per pair of clouds.
__global__ void __launch_bounds__(256) procrustes_moments_kernel(const float* __restrict__ S1, const float* __restrict__ S2, int N,
                                                                 double* __restrict__ acc /*[B][16]*/) {
  const int b = blockIdx.y;
  const float* p1 = S1 + (size_t)b * N * 3;
  const float* p2 = S2 + (size_t)b * N * 3;
  double m[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) m[i] = 0.0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256) {
    const double x[3] = {p1[i * 3], p1[i * 3 + 1], p1[i * 3 + 2]}, y[3] = {p2[i * 3], p2[i * 3 + 1], p2[i * 3 + 2]};
#pragma unroll
    for (int k = 0; k < 3; ++k) { m[k] += x[k]; m[3 + k] += y[k]; m[6] += x[k] * x[k]; }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) m[7 + r * 3 + c] += x[r] * y[c];
  }
  __shared__ double red[8][16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    double v = m[i];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    atomicAdd(acc + (size_t)b * 16 + threadIdx.x, v);
  }
}

__global__ void procrustes_solve_kernel(const double* __restrict__ acc, int N, int B, float* __restrict__ Rout, float* __restrict__ tout,
                                        float* __restrict__ sout) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* m = acc + (size_t)b * 16;
  const double n = (double)N;
  const double mu1[3] = {m[0] / n, m[1] / n, m[2] / n}, mu2[3] = {m[3] / n, m[4] / n, m[5] / n};
  const double var1 = m[6] - n * (mu1[0] * mu1[0] + mu1[1] * mu1[1] + mu1[2] * mu1[2]);
  double K[9];                                             // K = sum (x - mu1)(y - mu2)^T
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) K[r * 3 + c] = m[7 + r * 3 + c] - n * mu1[r] * mu2[c];
  float Kt[9];
  double nrm = 0;
  for (int i = 0; i < 9; ++i) nrm = fmax(nrm, fabs(K[i]));
  const double inv = nrm > 0 ? 1.0 / nrm : 1.0;            // so3_project takes floats: normalise first (the maximiser is scale-free)
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Kt[r * 3 + c] = (float)(K[c * 3 + r] * inv);
  double R[9];
  so3_project(Kt, R);
  double tr = 0;                                           // trace(R K)
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) tr += R[i * 3 + k] * K[k * 3 + i];
  const double scale = tr / var1;
  for (int i = 0; i < 9; ++i) Rout[(size_t)b * 9 + i] = (float)R[i];
  for (int i = 0; i < 3; ++i)
    tout[(size_t)b * 3 + i] = (float)(mu2[i] - scale * (R[i * 3] * mu1[0] + R[i * 3 + 1] * mu1[1] + R[i * 3 + 2] * mu1[2]));
  sout[b] = (float)scale;
}

// out = scale * R p + t for every point of cloud b (ProcrusteAlign.align_meshes, pose_utils.py:31; evalvideo_packed.py:131)
__global__ void similarity_apply_kernel(const float* __restrict__ pts, int n, const float* __restrict__ R, const float* __restrict__ t,
                                        const float* __restrict__ scale, float* __restrict__ out) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* r = R + (size_t)b * 9;
  const float s = scale[b];
  const float* p = pts + ((size_t)b * n + i) * 3;
  float* o = out + ((size_t)b * n + i) * 3;
  const float x = p[0], y = p[1], z = p[2];
  o[0] = s * (r[0] * x + r[1] * y + r[2] * z) + t[b * 3];
  o[1] = s * (r[3] * x + r[4] * y + r[5] * z) + t[b * 3 + 1];
  o[2] = s * (r[6] * x + r[7] * y + r[8] * z) + t[b * 3 + 2];
}

}  // namespace vt

using namespace vt;

extern "C" {

int vt_so3_project_fwd(const float* M, int B, float* R, void* stream) {
  if (B <= 0) return 0;
  so3_fwd_kernel<<<ceil_div(B, 64), 64, 0, (cudaStream_t)stream>>>(M, B, R);
  VT_CHECK_LAUNCH("vt_so3_project_fwd");
  return 0;
}

int vt_pca_orientation(const float* tgt_axis, const float* src_axis, int src_per_frame, const float* noise, int B, float* R, void* stream) {
  if (B <= 0) return 0;
  VT_CHECK_ARG(tgt_axis && src_axis && R, "vt_pca_orientation: null pointer");
  pca_orientation_kernel<<<ceil_div(B, 64), 64, 0, (cudaStream_t)stream>>>(tgt_axis, src_axis, src_per_frame ? 9 : 0, noise, B, R);
  VT_CHECK_LAUNCH("vt_pca_orientation");
  return 0;
}

int vt_so3_project_bwd(const float* M, const float* gR, int B, float* gM, void* stream) {
  if (B <= 0) return 0;
  so3_bwd_kernel<<<ceil_div(B, 64), 64, 0, (cudaStream_t)stream>>>(M, gR, B, gM);
  VT_CHECK_LAUNCH("vt_so3_project_bwd");
  return 0;
}

int vt_chamfer_fwd(const float* x, const int* x_off, const float* y, const int* y_off, int N, int* nn_x, int* nn_y, float* loss,
                   void* stream) {
  if (N <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "vt_chamfer_fwd memset");
  chamfer_fwd_kernel<<<dim3(N, CH_SPLIT), 128, 0, (cudaStream_t)stream>>>(x, x_off, y, y_off, N, nn_x, nn_y, loss);
  VT_CHECK_LAUNCH("vt_chamfer_fwd");
  return 0;
}

int vt_chamfer_bwd(const float* x, const int* x_off, const float* y, const int* y_off, int N, const int* nn_x, const int* nn_y,
                   const float* g_loss, float* gx, float* gy, void* stream) {
  if (N <= 0) return 0;
  chamfer_bwd_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(x, x_off, y, y_off, N, nn_x, nn_y, g_loss, gx, gy);
  VT_CHECK_LAUNCH("vt_chamfer_bwd");
  return 0;
}

int vt_nn_dist(const float* x, int nx, const float* y, int ny, int B, float* dist_x, float* dist_y, void* stream) {
  VT_CHECK_ARG(nx > 0 && ny > 0 && B > 0, "vt_nn_dist: empty clouds (%d x %d points, batch %d)", nx, ny, B);
  VT_CHECK_ARG(dist_x != nullptr && dist_y != nullptr, "vt_nn_dist: both outputs are required");
  const int n = nx > ny ? nx : ny;
  dim3 grid(ceil_div(n, 256), B, 2);
  nn_dist_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, nx, y, ny, dist_x, dist_y);
  VT_CHECK_LAUNCH("vt_nn_dist");
  return 0;
}

int vt_procrustes(const float* S1, const float* S2, int N, int B, double* workspace, float* R, float* t, float* scale, void* stream) {
  VT_CHECK_ARG(N >= 3 && B > 0, "vt_procrustes: %d points, batch %d", N, B);
  VT_CHECK_ARG(workspace != nullptr, "vt_procrustes: a workspace of 16 doubles per pair is required");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)B * 16 * sizeof(double), s);
  if (e != cudaSuccess) return cuda_fail(e, "vt_procrustes memset");
  int blocks = ceil_div(N, 256 * 8);
  if (blocks > 148 * 4) blocks = 148 * 4;
  procrustes_moments_kernel<<<dim3(blocks, B), 256, 0, s>>>(S1, S2, N, workspace);
  VT_CHECK_LAUNCH("vt_procrustes(moments)");
  procrustes_solve_kernel<<<ceil_div(B, 32), 32, 0, s>>>(workspace, N, B, R, t, scale);
  VT_CHECK_LAUNCH("vt_procrustes(solve)");
  return 0;
}

int vt_similarity_apply(const float* points, int n, int B, const float* R, const float* t, const float* scale, float* out, void* stream) {
  if (n <= 0 || B <= 0) return 0;
  similarity_apply_kernel<<<dim3(ceil_div(n, 256), B), 256, 0, (cudaStream_t)stream>>>(points, n, R, t, scale, out);
  VT_CHECK_LAUNCH("vt_similarity_apply");
  return 0;
}

}  // extern "C"
