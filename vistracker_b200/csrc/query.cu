// SIF-Net point query (model/chore_triplane.py:97-164, model/chore_tri_vis.py:31-50, model/geometry.py:4-14,
// model/camera.py:45-89) as ONE kernel per direction: perspective + three orthographic projections, eight bilinear
// gathers from NHWC feature maps, the 611-d feature vector kept in shared memory, and the five 611->128->128->128->out
// decoders.  The reference runs 8 grid_sample launches on NCHW maps, ~10 cat/transpose copies of a [B,611,N] tensor and
// 20 Conv1d launches for the same work.
//
// Internal feature order (weights are re-packed to it, see vistracker_b200/weights.py):
//   [ im_feat 256 | tmpx 64 | tri_tmpx right,back,top 3x32 | tri_feat right,back,top 3x64 | x, y, z-2.2 | 5 zero pad ] = 616
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

constexpr int QP = 32;            // points per CTA (one per lane)
constexpr int QLD = QP + 1;       // padded leading dimension of the transposed shared tiles
constexpr int QK = 616;           // padded feature length
constexpr int QH = 128;           // hidden width
constexpr int QKC = 32;           // weight rows staged per step
constexpr int Q_NHEAD = 5;
constexpr int Q_NOUT = 29;

struct QueryMaps {
  const float* im_feat; const float* tmpx; const float* tri_tmpx; const float* tri_feat;   // NHWC; tri_*: [3*B, ...] view-major
  int Hf, Wf, Ht, Wt;         // spatial size of im_feat / tri_feat and of tmpx / tri_tmpx
  int c_im, c_tmpx, c_tt, c_tf;
};

struct QueryCam { float fx, fy, cx, cy, crop, z0, out_dist; };

// packed decoder weights (floats): per head  W1[616][128] b1[128] W2[128][128] b2[128] W3[128][128] b3[128] W4[128][16] b4[16]
__host__ __device__ constexpr int q_head_stride() { return QK * QH + QH + 2 * (QH * QH + QH) + QH * 16 + 16; }

__device__ __constant__ int c_head_nout[Q_NHEAD] = {2, 9, 14, 3, 1};
__device__ __constant__ int c_head_off[Q_NHEAD] = {0, 2, 11, 25, 28};

// F.grid_sample(bilinear, zeros padding, align_corners=True) of C channels at (u, v) in [-1, 1]; one warp, lanes over
// float4 channel groups; result written transposed into featT[(dst + c)][p].
__device__ __forceinline__ void gather_bilinear(const float* __restrict__ map, int H, int W, int C, float u, float v,
                                                float* featT, int dst, int p, int lane) {
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(u, 1.f), 0.5f), (float)(W - 1));
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.f), 0.5f), (float)(H - 1));
  float fx0 = floorf(ix), fy0 = floorf(iy);
  float tx = ix - fx0, ty = iy - fy0;
  // out-of-range coordinates (incl. inf/nan from z <= 0) contribute zeros
  bool finite = (fabsf(ix) < 1e9f) && (fabsf(iy) < 1e9f);
  int x0 = finite ? (int)fx0 : -10, y0 = finite ? (int)fy0 : -10;
  float w00 = (1.f - tx) * (1.f - ty), w01 = tx * (1.f - ty), w10 = (1.f - tx) * ty, w11 = tx * ty;
  bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
  const float* b00 = map + ((size_t)y0 * W + x0) * C;
  for (int c4 = lane; c4 < C / 4; c4 += 32) {
    float4 a = make_float4(0, 0, 0, 0);
    if (vy0 && vx0) { float4 t = ld4(b00 + c4 * 4); a.x += t.x * w00; a.y += t.y * w00; a.z += t.z * w00; a.w += t.w * w00; }
    if (vy0 && vx1) { float4 t = ld4(b00 + C + c4 * 4); a.x += t.x * w01; a.y += t.y * w01; a.z += t.z * w01; a.w += t.w * w01; }
    if (vy1 && vx0) { float4 t = ld4(b00 + (size_t)W * C + c4 * 4); a.x += t.x * w10; a.y += t.y * w10; a.z += t.z * w10; a.w += t.w * w10; }
    if (vy1 && vx1) { float4 t = ld4(b00 + (size_t)W * C + C + c4 * 4); a.x += t.x * w11; a.y += t.y * w11; a.z += t.z * w11; a.w += t.w * w11; }
    float* d = featT + (size_t)(dst + c4 * 4) * QLD + p;
    d[0] = a.x; d[QLD] = a.y; d[2 * QLD] = a.z; d[3 * QLD] = a.w;
  }
}

// outT[c][p] = act(b[c] + sum_k inT[k][p] * W[k][c]) for c in [0,128): warp w owns channels 16w..16w+15, lane = point.
__device__ __forceinline__ void dense128(const float* inT, int K, const float* __restrict__ Wg, const float* __restrict__ bg,
                                         float* outT, float* sW, bool relu) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = bg[warp * 16 + i];
  for (int k0 = 0; k0 < K; k0 += QKC) {
    const int kc = min(QKC, K - k0);
    __syncthreads();                                    // previous chunk fully consumed
    for (int i = tid; i < kc * (QH / 4); i += 256)
      reinterpret_cast<float4*>(sW)[i] = reinterpret_cast<const float4*>(Wg + (size_t)k0 * QH)[i];
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < kc; ++kk) {
      float a = inT[(size_t)(k0 + kk) * QLD + lane];
      const float4* wr = reinterpret_cast<const float4*>(sW + kk * QH + warp * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 ww = wr[j];
        acc[j * 4 + 0] = fmaf(a, ww.x, acc[j * 4 + 0]); acc[j * 4 + 1] = fmaf(a, ww.y, acc[j * 4 + 1]);
        acc[j * 4 + 2] = fmaf(a, ww.z, acc[j * 4 + 2]); acc[j * 4 + 3] = fmaf(a, ww.w, acc[j * 4 + 3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) outT[(size_t)(warp * 16 + i) * QLD + lane] = relu ? fmaxf(acc[i], 0.f) : acc[i];
}

__global__ void __launch_bounds__(256, 1) query_fwd_kernel(const float* __restrict__ points, const float* __restrict__ crop_center,
                                                           const float* __restrict__ body_center, int B, int N, QueryMaps m,
                                                           QueryCam cam, const float* __restrict__ wpack,
                                                           float* __restrict__ out /*[B][29][N]*/,
                                                           float* __restrict__ feat_out /*[B][611][N] or null*/,
                                                           float* __restrict__ xy_out /*[B][2][N] or null*/) {
  extern __shared__ float smem[];
  float* featT = smem;                       // [616][33]
  float* hA = featT + QK * QLD;              // [128][33]
  float* hB = hA + QH * QLD;                 // [128][33]
  float* sW = hB + QH * QLD;                 // [32][128]
  __shared__ int s_in_img[QP];
  const int b = blockIdx.y, n0 = blockIdx.x * QP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- phase 1: projections + gathers; warp w handles points w, w+8, w+16, w+24 of the tile
  for (int pp = warp; pp < QP; pp += 8) {
    const int n = n0 + pp;
    if (n >= N) {        // tail: zero features so the MLP runs on defined data
      for (int k = lane; k < QK; k += 32) featT[(size_t)k * QLD + pp] = 0.f;
      if (lane == 0) s_in_img[pp] = 1;
      continue;
    }
    const float* pt = points + ((size_t)b * N + n) * 3;
    const float x = pt[0], y = pt[1], z = pt[2];
    // KinectColorCamera.project_points, model/camera.py:45-82 -- same operation order, no FMA contraction
    float px = __fadd_rn(__fdiv_rn(__fmul_rn(cam.fx, x), z), cam.cx);
    float py = __fadd_rn(__fdiv_rn(__fmul_rn(cam.fy, y), z), cam.cy);
    px = __fadd_rn(__fadd_rn(cam.crop * 0.5f, px), -crop_center[b * 2 + 0]);
    py = __fadd_rn(__fadd_rn(cam.crop * 0.5f, py), -crop_center[b * 2 + 1]);
    const float nx = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, px), cam.crop), -1.f);
    const float ny = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, py), cam.crop), -1.f);
    const bool in_img = nx >= -1.f && nx <= 1.f && ny >= -1.f && ny <= 1.f;
    // triplane_project, model/chore_triplane.py:220-251
    const float cxr = __fadd_rn(x, -body_center[b * 3 + 0]), cyr = __fadd_rn(y, -body_center[b * 3 + 1]),
                czr = __fadd_rn(z, -body_center[b * 3 + 2]);
    const float tu[3] = {czr, -cxr, cxr}, tv[3] = {cyr, cyr, -czr};
    gather_bilinear(m.im_feat + (size_t)b * m.Hf * m.Wf * m.c_im, m.Hf, m.Wf, m.c_im, nx, ny, featT, 0, pp, lane);
    gather_bilinear(m.tmpx + (size_t)b * m.Ht * m.Wt * m.c_tmpx, m.Ht, m.Wt, m.c_tmpx, nx, ny, featT, m.c_im, pp, lane);
    int dst = m.c_im + m.c_tmpx;
#pragma unroll
    for (int v = 0; v < 3; ++v)
      gather_bilinear(m.tri_tmpx + ((size_t)v * B + b) * m.Ht * m.Wt * m.c_tt, m.Ht, m.Wt, m.c_tt, tu[v], tv[v], featT,
                      dst + v * m.c_tt, pp, lane);
    dst += 3 * m.c_tt;
#pragma unroll
    for (int v = 0; v < 3; ++v)
      gather_bilinear(m.tri_feat + ((size_t)v * B + b) * m.Hf * m.Wf * m.c_tf, m.Hf, m.Wf, m.c_tf, tu[v], tv[v], featT,
                      dst + v * m.c_tf, pp, lane);
    dst += 3 * m.c_tf;
    if (lane < 8) {
      float zf = lane == 0 ? x : lane == 1 ? y : lane == 2 ? __fadd_rn(z, -cam.z0) : 0.f;   // get_zfeat, :207-218
      featT[(size_t)(dst + lane) * QLD + pp] = zf;
    }
    if (lane == 0) {
      s_in_img[pp] = in_img ? 1 : 0;
      if (xy_out) { xy_out[((size_t)b * 2 + 0) * N + n] = nx; xy_out[((size_t)b * 2 + 1) * N + n] = ny; }
    }
  }
  __syncthreads();

  // optional [B, 611, N] feature dump in the reference's channel order (CHORETriplane.query_features)
  if (feat_out && n0 + lane < N) {
    const int n_im = m.c_im, n_rest = m.c_tmpx + 3 * m.c_tt + 3 * m.c_tf;      // 256, 352
    for (int k = warp; k < n_im + 3 + n_rest; k += 8) {
      int src = k < n_im ? k : (k < n_im + 3 ? n_im + n_rest + (k - n_im) : k - 3);
      feat_out[((size_t)b * (n_im + 3 + n_rest) + k) * N + n0 + lane] = featT[(size_t)src * QLD + lane];
    }
  }

  // ---- phase 2: the five decoders
  for (int h = 0; h < Q_NHEAD; ++h) {
    const float* W1 = wpack + (size_t)h * q_head_stride();
    const float* b1 = W1 + QK * QH;
    const float* W2 = b1 + QH;  const float* b2 = W2 + QH * QH;
    const float* W3 = b2 + QH;  const float* b3 = W3 + QH * QH;
    const float* W4 = b3 + QH;  const float* b4 = W4 + QH * 16;
    dense128(featT, QK, W1, b1, hA, sW, true);
    dense128(hA, QH, W2, b2, hB, sW, true);      // (entry barrier of dense128 orders the hA writes before these reads)
    dense128(hB, QH, W3, b3, hA, sW, true);
    __syncthreads();
    const int nout = c_head_nout[h], off = c_head_off[h];
    for (int c = warp; c < nout; c += 8) {
      float acc = b4[c];
#pragma unroll 8
      for (int k = 0; k < QH; ++k) acc = fmaf(hA[(size_t)k * QLD + lane], W4[k * 16 + c], acc);
      if (h == 4) acc = 1.f / (1.f + expf(-acc));                       // nn.Sigmoid on the visibility head
      if (h == 0 && !s_in_img[lane]) acc = cam.out_dist;                // chore_triplane.py:156-159
      if (n0 + lane < N) out[((size_t)b * Q_NOUT + off + c) * N + n0 + lane] = acc;
    }
  }
}

}  // namespace vt

using namespace vt;

extern "C" {

long long vt_query_wpack_floats(void) { return (long long)Q_NHEAD * q_head_stride(); }

int vt_query_fwd(const float* points, const float* crop_center, const float* body_center, int B, int N,
                 const float* im_feat, const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht,
                 int Wt, int c_im, int c_tmpx, int c_tt, int c_tf, const float* cam7, const float* wpack, float* out,
                 float* feat_out, float* xy_out, void* stream) {
  VT_CHECK_ARG(c_im + c_tmpx + 3 * c_tt + 3 * c_tf + 3 <= QK && c_im % 4 == 0 && c_tmpx % 4 == 0 && c_tt % 4 == 0 && c_tf % 4 == 0,
               "vt_query_fwd: feature layout %d/%d/%d/%d does not fit the packed length %d", c_im, c_tmpx, c_tt, c_tf, QK);
  VT_CHECK_ARG(c_im + c_tmpx + 3 * c_tt + 3 * c_tf + 3 == 611, "vt_query_fwd: only the 611-feature tri-vis layout is built");
  if (B <= 0 || N <= 0) return 0;
  QueryMaps m{im_feat, tmpx, tri_tmpx, tri_feat, Hf, Wf, Ht, Wt, c_im, c_tmpx, c_tt, c_tf};
  QueryCam cam{cam7[0], cam7[1], cam7[2], cam7[3], cam7[4], cam7[5], cam7[6]};
  size_t smem = (size_t)(QK * QLD + 2 * QH * QLD + QKC * QH) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(query_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "vt_query_fwd smem attr");
  dim3 grid(ceil_div(N, QP), B);
  query_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(points, crop_center, body_center, B, N, m, cam, wpack, out, feat_out, xy_out);
  VT_CHECK_LAUNCH("vt_query_fwd");
  return 0;
}

}  // extern "C"
