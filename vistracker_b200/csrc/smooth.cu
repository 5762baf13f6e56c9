// SmoothNet stage: SMPL-T and object-rotation smoothing of a whole trajectory that already sits in
// device memory (the all-gathered [T, 169] block of vistracker_b200/parallel.py) instead of the reference's per-frame pickle ->
// joblib round trip (smoothnet/smooth_smplt.py, smoothnet/smooth_objrot.py, smoothnet/smooth_base.py).
//
//   vt_smooth_pack_smplt       axis-angle -> 6-D rotation (numpy_axis_to_rot6D, geometry_utils.py:285-346) + [pose6d | betas | trans] rows
//   vt_smoothnet_clips         SmoothNet (smoothnet/models/smoothnet.py:125-141) on every sliding window x channel row: gather of the
//                              window (step 1, smooth_base.py:45-73), optional "relative to the window's first frame" translation
//                              (smooth_smplt.py:88-91, 39-42), encoder 64->512 + LeakyReLU(.1), residual blocks 512->16->512 with
//                              LeakyReLU(.2), decoder 512->64, all in shared memory: nothing but the trajectory is read from HBM
//   vt_smooth_window_mean      clips2seq_fast (smoothnet/utils/utils.py:63-86): mean over the windows that contain a frame
//   vt_smooth_unpack_smplt     rot6D_to_axis (geometry_utils.py:63-77, 93-247, 279-282): Gram-Schmidt, kornia's branchy
//                              rotation-matrix -> quaternion -> angle-axis
//   vt_smooth_rot6d_to_rotmat  rot6d_to_rotmat for the object rotations (smooth_objrot.py:104-105), written transposed as the
//                              reference stores `obj_angles`
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

constexpr int SN_W = 64;          // window length
constexpr int SN_H = 512;         // hidden width
constexpr int SN_R = 16;          // residual-block width
constexpr int SN_P = 32;          // (window, channel) rows per CTA, one per lane
constexpr int SN_LD = SN_P + 1;
constexpr int SN_KC = 32;         // weight rows staged per step
constexpr int SN_TN = 128;        // output columns per pass (8 warps x 16)

__host__ __device__ constexpr int sn_block_floats() { return SN_H * SN_R + SN_R + SN_R * SN_H + SN_H; }
__host__ __device__ constexpr int sn_pack_floats(int n_blocks) { return SN_W * SN_H + SN_H + n_blocks * sn_block_floats() + SN_H * SN_W + SN_W; }

__device__ __forceinline__ float lrelu(float x, float a) { return x > 0.f ? x : a * x; }

// acc[i] = bias[col0 + 16 warp + i] + sum_k inT[k][lane] * Wg[k][col0 + 16 warp + i] for the warps whose columns exist (ncols <= 128);
// weight rows are staged through shared memory SN_KC at a time with cp.async, double-buffered (same scheme as query.cu: dense16)
__device__ __forceinline__ void sn_stage(float* dst, const float* __restrict__ Wg, int ldw, int col0, int ncols, int k0, int kc) {
  const int c4n = ncols / 4;
  for (int i = threadIdx.x; i < kc * c4n; i += 256) {
    const int r = i / c4n, c4 = i % c4n;
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + (size_t)r * SN_TN + c4 * 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(Wg + (size_t)(k0 + r) * ldw + col0 + c4 * 4) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void sn_dense16(const float* inT, int K, const float* __restrict__ Wg, int ldw, int col0, int ncols,
                                           const float* __restrict__ bias, float* sW /*[2][SN_KC][SN_TN]*/, float (&acc)[16]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool active = warp * 16 < ncols;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = active ? bias[col0 + warp * 16 + i] : 0.f;
  const int nchunks = (K + SN_KC - 1) / SN_KC;
  sn_stage(sW, Wg, ldw, col0, ncols, 0, min(SN_KC, K));
  for (int c = 0; c < nchunks; ++c) {
    const int k0 = c * SN_KC, kc = min(SN_KC, K - k0);
    if (c + 1 < nchunks) {
      sn_stage(sW + ((c + 1) & 1) * SN_KC * SN_TN, Wg, ldw, col0, ncols, k0 + SN_KC, min(SN_KC, K - k0 - SN_KC));
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (active) {
      const float* w = sW + (c & 1) * SN_KC * SN_TN;
#pragma unroll 4
      for (int kk = 0; kk < kc; ++kk) {
        const float a = inT[(size_t)(k0 + kk) * SN_LD + lane];
        const float4* wr = reinterpret_cast<const float4*>(w + kk * SN_TN + warp * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 ww = wr[j];
          acc[j * 4 + 0] = fmaf(a, ww.x, acc[j * 4 + 0]); acc[j * 4 + 1] = fmaf(a, ww.y, acc[j * 4 + 1]);
          acc[j * 4 + 2] = fmaf(a, ww.z, acc[j * 4 + 2]); acc[j * 4 + 3] = fmaf(a, ww.w, acc[j * 4 + 3]);
        }
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256, 1) smoothnet_clips_kernel(const float* __restrict__ seq, int L, int D, int c0, int nC, int relative,
                                                                 int n_blocks, const float* __restrict__ wpack, float* __restrict__ clips) {
  extern __shared__ float smem[];
  float* xT = smem;                          // [64][33]   the window of each row (time-major), later the decoder output
  float* hT = xT + SN_W * SN_LD;             // [512][33]  hidden state
  float* rT = hT + SN_H * SN_LD;             // [16][33]
  float* sW = rT + SN_R * SN_LD;             // [2][32][128]
  __shared__ float s_init[SN_P];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = L - SN_W + 1;
  const long long row = (long long)blockIdx.x * SN_P + lane, n_rows = (long long)B * nC;
  const bool valid = row < n_rows;
  const int b = valid ? (int)(row / nC) : 0, c = valid ? (int)(row % nC) : 0;
  const float* We = wpack;
  const float* be = We + SN_W * SN_H;
  const float* blk = be + SN_H;
  const float* Wd = blk + (size_t)n_blocks * sn_block_floats();
  const float* bd = Wd + SN_H * SN_W;

  // window gather: lanes = consecutive channels of (mostly) one window -> coalesced rows of the trajectory
  const float init = (valid && relative) ? seq[(size_t)b * D + c0 + c] : 0.f;
  if (warp == 0) s_init[lane] = init;
  for (int t = warp; t < SN_W; t += 8) xT[t * SN_LD + lane] = valid ? seq[(size_t)(b + t) * D + c0 + c] - init : 0.f;
  __syncthreads();

  // encoder: Linear(64 -> 512) + LeakyReLU(0.1)
  for (int col0 = 0; col0 < SN_H; col0 += SN_TN) {
    float acc[16];
    sn_dense16(xT, SN_W, We, SN_H, col0, SN_TN, be, sW, acc);
#pragma unroll
    for (int i = 0; i < 16; ++i) hT[(size_t)(col0 + warp * 16 + i) * SN_LD + lane] = lrelu(acc[i], 0.1f);
  }
  __syncthreads();
  // residual blocks: x + lrelu(W2 lrelu(W1 x + b1) + b2), LeakyReLU(0.2), dropout is the identity in eval mode
  for (int blkI = 0; blkI < n_blocks; ++blkI) {
    const float* W1 = blk + (size_t)blkI * sn_block_floats();      // [512][16]
    const float* b1 = W1 + SN_H * SN_R;
    const float* W2 = b1 + SN_R;                                   // [16][512]
    const float* b2 = W2 + SN_R * SN_H;
    {   // 512 -> 16: warp w owns outputs 2w, 2w+1
      float a0 = b1[2 * warp], a1 = b1[2 * warp + 1];
#pragma unroll 8
      for (int k = 0; k < SN_H; ++k) {
        const float h = hT[(size_t)k * SN_LD + lane];
        const float2 w = __ldg(reinterpret_cast<const float2*>(W1 + k * SN_R + 2 * warp));
        a0 = fmaf(h, w.x, a0); a1 = fmaf(h, w.y, a1);
      }
      rT[(2 * warp) * SN_LD + lane] = lrelu(a0, 0.2f);
      rT[(2 * warp + 1) * SN_LD + lane] = lrelu(a1, 0.2f);
    }
    __syncthreads();
    for (int col0 = 0; col0 < SN_H; col0 += SN_TN) {
      float acc[16];
      sn_dense16(rT, SN_R, W2, SN_H, col0, SN_TN, b2, sW, acc);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float* h = &hT[(size_t)(col0 + warp * 16 + i) * SN_LD + lane];
        *h = lrelu(acc[i], 0.2f) + *h;
      }
    }
    __syncthreads();
  }
  // decoder: Linear(512 -> 64); warps 0-3 own the 64 output frames
  {
    float acc[16];
    sn_dense16(hT, SN_H, Wd, SN_W, 0, SN_W, bd, sW, acc);
    if (warp < 4 && valid) {
      const float add = s_init[lane];
#pragma unroll
      for (int i = 0; i < 16; ++i) clips[((size_t)b * SN_W + warp * 16 + i) * D + c0 + c] = acc[i] + add;
    }
  }
}

// out[l][c] = mean over the windows b in [max(0, l-W+1), min(l, B-1)] of clips[b][l-b][c]; channels in [pass0, pass0+passN) were not
// smoothed (betas): every window holds seq[l][c] there (smoothnet_smpl.py:38-45), summed and divided the same way
__global__ void smooth_window_mean_kernel(const float* __restrict__ clips, const float* __restrict__ seq, int L, int D, int W, int pass0,
                                          int passN, float* __restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)L * D) return;
  const int l = (int)(idx / D), c = (int)(idx % D);
  const int B = L - W + 1, b0 = max(0, l - W + 1), b1 = min(l, B - 1);
  const bool pass = c >= pass0 && c < pass0 + passN;
  float s = 0.f;
  for (int b = b1; b >= b0; --b) s += pass ? seq[(size_t)l * D + c] : clips[((size_t)b * W + (l - b)) * D + c];   // t = W-1 .. 0 maps to b descending
  out[idx] = s / (float)(b1 - b0 + 1);
}

// ------------------------------------------------------------------------------------------------ rotation conversions
// numpy_axis_to_rot6D in fp32: quaternion route with the reference's two epsilons
__device__ __forceinline__ void axis_to_rot6d(const float* ax, float* r6) {
  const float ex = __fadd_rn(ax[0], 1e-8f), ey = __fadd_rn(ax[1], 1e-8f), ez = __fadd_rn(ax[2], 1e-8f);
  const float l = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
  const float nx = ax[0] / l, ny = ax[1] / l, nz = ax[2] / l;
  const float half = l * 0.5f, cw = cosf(half), sw = sinf(half);
  float q[4] = {cw, sw * nx, sw * ny, sw * nz};
  const float e0 = __fadd_rn(q[0], 1e-8f), e1 = __fadd_rn(q[1], 1e-8f), e2 = __fadd_rn(q[2], 1e-8f), e3 = __fadd_rn(q[3], 1e-8f);
  const float qn = sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1)), __fmul_rn(e2, e2)), __fmul_rn(e3, e3)));
  const float w = q[0] / qn, x = q[1] / qn, y = q[2] / qn, z = q[3] / qn;
  const float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z, wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  // R = [[w2+x2-y2-z2, 2xy-2wz, .], [2wz+2xy, w2-x2+y2-z2, .], [2xz-2wy, 2wx+2yz, .]]; rot6d = first two columns, row-major
  r6[0] = __fadd_rn(__fadd_rn(__fadd_rn(w2, x2), -y2), -z2); r6[1] = __fadd_rn(2.f * xy, -(2.f * wz));
  r6[2] = __fadd_rn(2.f * wz, 2.f * xy);                      r6[3] = __fadd_rn(__fadd_rn(__fadd_rn(w2, -x2), y2), -z2);
  r6[4] = __fadd_rn(2.f * xz, -(2.f * wy));                   r6[5] = __fadd_rn(2.f * wx, 2.f * yz);
}

// rot6d_to_rotmat: columns b1, b2, b3 (F.normalize eps 1e-12)
__device__ __forceinline__ void rot6d_to_rotmat(const float* r6, float (&R)[3][3]) {
  const float a1[3] = {r6[0], r6[2], r6[4]}, a2[3] = {r6[1], r6[3], r6[5]};
  const float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
  const float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
  const float d = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
  const float u[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
  const float n2 = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
  const float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
  const float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
#pragma unroll
  for (int i = 0; i < 3; ++i) { R[i][0] = b1[i]; R[i][1] = b2[i]; R[i][2] = b3[i]; }
}

// rotation_matrix_to_quaternion (kornia's four-branch form, eps 1e-6) + quaternion_to_angle_axis, NaN -> 0
__device__ __forceinline__ void rotmat_to_axis(const float (&R)[3][3], float* aa) {
  // rmat_t = R^T: rmat_t[i][j] = R[j][i]
  const float t00 = R[0][0], t11 = R[1][1], t22 = R[2][2];
  const float t01 = R[1][0], t10 = R[0][1], t02 = R[2][0], t20 = R[0][2], t12 = R[2][1], t21 = R[1][2];
  const bool d2 = t22 < 1e-6f, d01 = t00 > t11, d0n1 = t00 < -t11;
  float q[4], t;
  if (d2 && d01)        { t = 1.f + t00 - t11 - t22; q[0] = t12 - t21; q[1] = t; q[2] = t01 + t10; q[3] = t20 + t02; }
  else if (d2 && !d01)  { t = 1.f - t00 + t11 - t22; q[0] = t20 - t02; q[1] = t01 + t10; q[2] = t; q[3] = t12 + t21; }
  else if (!d2 && d0n1) { t = 1.f - t00 - t11 + t22; q[0] = t01 - t10; q[1] = t20 + t02; q[2] = t12 + t21; q[3] = t; }
  else                  { t = 1.f + t00 + t11 + t22; q[0] = t; q[1] = t12 - t21; q[2] = t20 - t02; q[3] = t01 - t10; }
  const float rt = sqrtf(t);                                 // q /= sqrt(t); q *= 0.5 (same two roundings as the reference)
  const float qw = q[0] / rt * 0.5f, q1 = q[1] / rt * 0.5f, q2 = q[2] / rt * 0.5f, q3 = q[3] / rt * 0.5f;
  const float s2 = q1 * q1 + q2 * q2 + q3 * q3, s = sqrtf(s2);
  const float two_theta = 2.f * (qw < 0.f ? atan2f(-s, -qw) : atan2f(s, qw));
  const float k = s2 > 0.f ? two_theta / s : 2.f;
  aa[0] = q1 * k; aa[1] = q2 * k; aa[2] = q3 * k;
#pragma unroll
  for (int i = 0; i < 3; ++i) if (aa[i] != aa[i]) aa[i] = 0.f;
}

__global__ void smooth_pack_smplt_kernel(const float* __restrict__ poses, int pose_dim, const float* __restrict__ betas,
                                         const float* __restrict__ trans, int L, float* __restrict__ seq) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L * 25) return;
  const int l = idx / 25, j = idx % 25;
  float* row = seq + (size_t)l * 157;
  if (j == 24) {
    for (int i = 0; i < 10; ++i) row[144 + i] = betas[(size_t)l * 10 + i];
    for (int i = 0; i < 3; ++i) row[154 + i] = trans[(size_t)l * 3 + i];
    return;
  }
  // SMPL-H (156) -> SMPL (72): joints 0..22 as they are, joint 23 = SMPL-H joint 37 (smooth_smplt.py:103-113)
  const int src = (pose_dim == 156 && j == 23) ? 111 : j * 3;
  axis_to_rot6d(poses + (size_t)l * pose_dim + src, row + j * 6);
}

__global__ void smooth_unpack_smplt_kernel(const float* __restrict__ seq, int L, float* __restrict__ poses, float* __restrict__ betas,
                                           float* __restrict__ trans) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L * 25) return;
  const int l = idx / 25, j = idx % 25;
  const float* row = seq + (size_t)l * 157;
  if (j == 24) {
    for (int i = 0; i < 10; ++i) betas[(size_t)l * 10 + i] = row[144 + i];
    for (int i = 0; i < 3; ++i) trans[(size_t)l * 3 + i] = row[154 + i];
    return;
  }
  float R[3][3];
  rot6d_to_rotmat(row + j * 6, R);
  rotmat_to_axis(R, poses + (size_t)l * 72 + j * 3);
}

__global__ void smooth_rot6d_to_rotmat_kernel(const float* __restrict__ r6, int L, int transposed, float* __restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  float R[3][3];
  rot6d_to_rotmat(r6 + (size_t)l * 6, R);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out[(size_t)l * 9 + i * 3 + j] = transposed ? R[j][i] : R[i][j];
}

}  // namespace vt

using namespace vt;

extern "C" {

long long vt_smoothnet_pack_floats(int n_blocks) { return n_blocks < 0 ? -1 : (long long)sn_pack_floats(n_blocks); }

int vt_smooth_pack_smplt(const float* poses, int pose_dim, const float* betas, const float* trans, int L, float* seq, void* stream) {
  VT_CHECK_ARG(pose_dim == 72 || pose_dim == 156, "vt_smooth_pack_smplt: pose_dim %d (72 or 156)", pose_dim);
  if (L <= 0) return 0;
  smooth_pack_smplt_kernel<<<ceil_div(L * 25, 128), 128, 0, (cudaStream_t)stream>>>(poses, pose_dim, betas, trans, L, seq);
  VT_CHECK_LAUNCH("vt_smooth_pack_smplt");
  return 0;
}

int vt_smoothnet_clips(const float* seq, int L, int D, int c0, int nC, int relative, int window, int hidden, int res_hidden, int n_blocks,
                       const float* wpack, float* clips, void* stream) {
  VT_CHECK_ARG(window == SN_W && hidden == SN_H && res_hidden == SN_R, "vt_smoothnet_clips: built for window %d, hidden %d, residual %d (got %d / %d / %d)",
               SN_W, SN_H, SN_R, window, hidden, res_hidden);
  VT_CHECK_ARG(n_blocks >= 0 && n_blocks <= 8, "vt_smoothnet_clips: %d residual blocks", n_blocks);
  VT_CHECK_ARG(c0 >= 0 && nC > 0 && c0 + nC <= D, "vt_smoothnet_clips: channels [%d, %d) outside the %d-wide rows", c0, c0 + nC, D);
  VT_CHECK_ARG(L >= window, "vt_smoothnet_clips: %d frames are fewer than one window of %d", L, window);
  const long long rows = (long long)(L - window + 1) * nC;
  const size_t smem = (size_t)(SN_W * SN_LD + SN_H * SN_LD + SN_R * SN_LD + 2 * SN_KC * SN_TN) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(smoothnet_clips_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "vt_smoothnet_clips smem attr");
  smoothnet_clips_kernel<<<(unsigned)((rows + SN_P - 1) / SN_P), 256, smem, (cudaStream_t)stream>>>(seq, L, D, c0, nC, relative, n_blocks, wpack, clips);
  VT_CHECK_LAUNCH("vt_smoothnet_clips");
  return 0;
}

int vt_smooth_window_mean(const float* clips, const float* seq, int L, int D, int window, int pass0, int passN, float* out, void* stream) {
  VT_CHECK_ARG(L >= window && window > 0, "vt_smooth_window_mean: %d frames, window %d", L, window);
  const size_t n = (size_t)L * D;
  smooth_window_mean_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(clips, seq, L, D, window, pass0, passN, out);
  VT_CHECK_LAUNCH("vt_smooth_window_mean");
  return 0;
}

int vt_smooth_unpack_smplt(const float* seq, int L, float* poses, float* betas, float* trans, void* stream) {
  if (L <= 0) return 0;
  smooth_unpack_smplt_kernel<<<ceil_div(L * 25, 128), 128, 0, (cudaStream_t)stream>>>(seq, L, poses, betas, trans);
  VT_CHECK_LAUNCH("vt_smooth_unpack_smplt");
  return 0;
}

int vt_smooth_rot6d_to_rotmat(const float* rot6d, int L, int transposed, float* out, void* stream) {
  if (L <= 0) return 0;
  smooth_rot6d_to_rotmat_kernel<<<ceil_div(L, 128), 128, 0, (cudaStream_t)stream>>>(rot6d, L, transposed, out);
  VT_CHECK_LAUNCH("vt_smooth_rot6d_to_rotmat");
  return 0;
}

}  // extern "C"
