// SIF-Net point query (model/chore_triplane.py:97-164, model/chore_tri_vis.py:31-50, model/geometry.py:4-14,
// model/camera.py:45-89) as ONE kernel per direction: perspective + three orthographic projections, eight bilinear
// gathers from NHWC feature maps, the 611-d feature vector kept in shared memory, and the five 611->128->128->128->out
// decoders.  The reference runs 8 grid_sample launches on NCHW maps, ~10 cat/transpose copies of a [B,611,N] tensor and
// 20 Conv1d launches for the same work, and builds an autograd graph over all of it for the gradient w.r.t. the points
// (recon/gen/generator.py:86-96, recon/recon_fit_behave.py:467-513); query_bwd_kernel recomputes the forward in shared
// memory and back-propagates analytically, so nothing but points-in / gradients-out touches HBM.
//
// Internal feature order (weights are re-packed to it, see vistracker_b200/weights.py):
//   [ im_feat 256 | tmpx 64 | tri_tmpx right,back,top 3x32 | tri_feat right,back,top 3x64 | x, y, z-2.2 | 5 zero pad ] = 616
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

constexpr int QP = 32;            // points per CTA (one per lane)
constexpr int QLD = QP + 1;       // padded leading dimension of the transposed shared tiles
constexpr int QK = 616;           // padded feature length
constexpr int QKB = 640;          // feature length padded to whole 128-column chunks (backward W1 layout)
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

// packed decoder weights (floats)
//   forward : per head  W1[616][128] b1[128] W2[128][128] b2[128] W3[128][128] b3[128] W4[128][16] b4[16]   (k-major)
//   backward: per head  W1b[128][640] W2b[128][128] W3b[128][128] W4b[16][128]                              (out-major = torch layout)
__host__ __device__ constexpr int q_head_stride() { return QK * QH + QH + 2 * (QH * QH + QH) + QH * 16 + 16; }
__host__ __device__ constexpr int q_head_stride_bwd() { return QH * QKB + 2 * QH * QH + 16 * QH; }

__device__ __constant__ int c_head_nout[Q_NHEAD] = {2, 9, 14, 3, 1};
__device__ __constant__ int c_head_off[Q_NHEAD] = {0, 2, 11, 25, 28};

struct Bilinear {
  int x0, y0; float tx, ty; bool vx0, vx1, vy0, vy1;
};

// F.grid_sample(bilinear, zeros padding, align_corners=True) coordinates for (u, v) in [-1, 1]
__device__ __forceinline__ Bilinear bilinear_setup(int H, int W, float u, float v) {
  Bilinear s;
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(u, 1.f), 0.5f), (float)(W - 1));
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.f), 0.5f), (float)(H - 1));
  float fx0 = floorf(ix), fy0 = floorf(iy);
  s.tx = ix - fx0; s.ty = iy - fy0;
  bool finite = (fabsf(ix) < 1e9f) && (fabsf(iy) < 1e9f);      // inf/nan (z <= 0) -> every tap out of range -> zeros
  s.x0 = finite ? (int)fx0 : -10; s.y0 = finite ? (int)fy0 : -10;
  s.vx0 = s.x0 >= 0 && s.x0 < W; s.vx1 = s.x0 + 1 >= 0 && s.x0 + 1 < W;
  s.vy0 = s.y0 >= 0 && s.y0 < H; s.vy1 = s.y0 + 1 >= 0 && s.y0 + 1 < H;
  return s;
}

// one warp, lanes over float4 channel groups; result written transposed into featT[(dst + c)][p]
__device__ __forceinline__ void gather_bilinear(const float* __restrict__ map, int H, int W, int C, float u, float v,
                                                float* featT, int dst, int p, int lane) {
  const Bilinear s = bilinear_setup(H, W, u, v);
  const float w00 = (1.f - s.tx) * (1.f - s.ty), w01 = s.tx * (1.f - s.ty), w10 = (1.f - s.tx) * s.ty, w11 = s.tx * s.ty;
  const float* b00 = map + ((long long)s.y0 * W + s.x0) * C;
  for (int c4 = lane; c4 < C / 4; c4 += 32) {
    float4 a = make_float4(0, 0, 0, 0);
    if (s.vy0 && s.vx0) { float4 t = ld4(b00 + c4 * 4); a.x += t.x * w00; a.y += t.y * w00; a.z += t.z * w00; a.w += t.w * w00; }
    if (s.vy0 && s.vx1) { float4 t = ld4(b00 + C + c4 * 4); a.x += t.x * w01; a.y += t.y * w01; a.z += t.z * w01; a.w += t.w * w01; }
    if (s.vy1 && s.vx0) { float4 t = ld4(b00 + (long long)W * C + c4 * 4); a.x += t.x * w10; a.y += t.y * w10; a.z += t.z * w10; a.w += t.w * w10; }
    if (s.vy1 && s.vx1) { float4 t = ld4(b00 + (long long)W * C + C + c4 * 4); a.x += t.x * w11; a.y += t.y * w11; a.z += t.z * w11; a.w += t.w * w11; }
    float* d = featT + (size_t)(dst + c4 * 4) * QLD + p;
    d[0] = a.x; d[QLD] = a.y; d[2 * QLD] = a.z; d[3 * QLD] = a.w;
  }
}

// d(sum_c g[c] * sample_c)/d(u, v): per-lane partial sums (caller warp-reduces).  gT[(src + c)][p] is the feature gradient.
__device__ __forceinline__ void gather_bilinear_grad(const float* __restrict__ map, int H, int W, int C, float u, float v,
                                                     const float* gT, int src, int p, int lane, float& gu, float& gv) {
  const Bilinear s = bilinear_setup(H, W, u, v);
  const float* b00 = map + ((long long)s.y0 * W + s.x0) * C;
  float dix = 0.f, diy = 0.f;
  for (int c4 = lane; c4 < C / 4; c4 += 32) {
    float4 z = make_float4(0, 0, 0, 0), v00 = z, v01 = z, v10 = z, v11 = z;
    if (s.vy0 && s.vx0) v00 = ld4(b00 + c4 * 4);
    if (s.vy0 && s.vx1) v01 = ld4(b00 + C + c4 * 4);
    if (s.vy1 && s.vx0) v10 = ld4(b00 + (long long)W * C + c4 * 4);
    if (s.vy1 && s.vx1) v11 = ld4(b00 + (long long)W * C + C + c4 * 4);
    const float* g = gT + (size_t)(src + c4 * 4) * QLD + p;
    const float g0 = g[0], g1 = g[QLD], g2 = g[2 * QLD], g3 = g[3 * QLD];
    dix += g0 * ((v01.x - v00.x) * (1.f - s.ty) + (v11.x - v10.x) * s.ty) + g1 * ((v01.y - v00.y) * (1.f - s.ty) + (v11.y - v10.y) * s.ty) +
           g2 * ((v01.z - v00.z) * (1.f - s.ty) + (v11.z - v10.z) * s.ty) + g3 * ((v01.w - v00.w) * (1.f - s.ty) + (v11.w - v10.w) * s.ty);
    diy += g0 * ((v10.x - v00.x) * (1.f - s.tx) + (v11.x - v01.x) * s.tx) + g1 * ((v10.y - v00.y) * (1.f - s.tx) + (v11.y - v01.y) * s.tx) +
           g2 * ((v10.z - v00.z) * (1.f - s.tx) + (v11.z - v01.z) * s.tx) + g3 * ((v10.w - v00.w) * (1.f - s.tx) + (v11.w - v01.w) * s.tx);
  }
  gu += dix * 0.5f * (float)(W - 1);
  gv += diy * 0.5f * (float)(H - 1);
}

struct PointProj { float x, y, z, nx, ny; float tu[3], tv[3]; bool in_img; };

// KinectColorCamera.project_points (model/camera.py:45-82) + triplane_project (model/chore_triplane.py:220-251);
// same operation order as the reference, no FMA contraction, so xy (and the in-image test) is bit-identical.
__device__ __forceinline__ PointProj project_point(const float* pt, const float* cc, const float* bc, const QueryCam& cam) {
  PointProj q;
  q.x = pt[0]; q.y = pt[1]; q.z = pt[2];
  float px = __fadd_rn(__fdiv_rn(__fmul_rn(cam.fx, q.x), q.z), cam.cx);
  float py = __fadd_rn(__fdiv_rn(__fmul_rn(cam.fy, q.y), q.z), cam.cy);
  px = __fadd_rn(__fadd_rn(cam.crop * 0.5f, px), -cc[0]);
  py = __fadd_rn(__fadd_rn(cam.crop * 0.5f, py), -cc[1]);
  q.nx = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, px), cam.crop), -1.f);
  q.ny = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, py), cam.crop), -1.f);
  q.in_img = q.nx >= -1.f && q.nx <= 1.f && q.ny >= -1.f && q.ny <= 1.f;
  const float cx = __fadd_rn(q.x, -bc[0]), cy = __fadd_rn(q.y, -bc[1]), cz = __fadd_rn(q.z, -bc[2]);
  q.tu[0] = cz; q.tu[1] = -cx; q.tu[2] = cx;       // right (z, y) | back (-x, y) | top (x, -z)
  q.tv[0] = cy; q.tv[1] = cy;  q.tv[2] = -cz;
  return q;
}

// phase 1 of both kernels: features of the tile's 32 points into featT; warp w handles points w, w+8, w+16, w+24
__device__ __forceinline__ void gather_tile(const float* __restrict__ points, const float* __restrict__ crop_center,
                                            const float* __restrict__ body_center, int B, int N, int b, int n0,
                                            const QueryMaps& m, const QueryCam& cam, float* featT, int* s_in_img,
                                            float* __restrict__ xy_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int pp = warp; pp < QP; pp += 8) {
    const int n = n0 + pp;
    if (n >= N) {        // tail: zero features so the MLP runs on defined data
      for (int k = lane; k < QK; k += 32) featT[(size_t)k * QLD + pp] = 0.f;
      if (lane == 0) s_in_img[pp] = 1;
      continue;
    }
    const PointProj q = project_point(points + ((size_t)b * N + n) * 3, crop_center + b * 2, body_center + b * 3, cam);
    gather_bilinear(m.im_feat + (size_t)b * m.Hf * m.Wf * m.c_im, m.Hf, m.Wf, m.c_im, q.nx, q.ny, featT, 0, pp, lane);
    gather_bilinear(m.tmpx + (size_t)b * m.Ht * m.Wt * m.c_tmpx, m.Ht, m.Wt, m.c_tmpx, q.nx, q.ny, featT, m.c_im, pp, lane);
    int dst = m.c_im + m.c_tmpx;
#pragma unroll
    for (int v = 0; v < 3; ++v)
      gather_bilinear(m.tri_tmpx + ((size_t)v * B + b) * m.Ht * m.Wt * m.c_tt, m.Ht, m.Wt, m.c_tt, q.tu[v], q.tv[v], featT,
                      dst + v * m.c_tt, pp, lane);
    dst += 3 * m.c_tt;
#pragma unroll
    for (int v = 0; v < 3; ++v)
      gather_bilinear(m.tri_feat + ((size_t)v * B + b) * m.Hf * m.Wf * m.c_tf, m.Hf, m.Wf, m.c_tf, q.tu[v], q.tv[v], featT,
                      dst + v * m.c_tf, pp, lane);
    dst += 3 * m.c_tf;
    if (lane < 8) {
      float zf = lane == 0 ? q.x : lane == 1 ? q.y : lane == 2 ? __fadd_rn(q.z, -cam.z0) : 0.f;   // get_zfeat, :207-218
      featT[(size_t)(dst + lane) * QLD + pp] = zf;
    }
    if (lane == 0) {
      s_in_img[pp] = q.in_img ? 1 : 0;
      if (xy_out) { xy_out[((size_t)b * 2 + 0) * N + n] = q.nx; xy_out[((size_t)b * 2 + 1) * N + n] = q.ny; }
    }
  }
}

// acc[i] = bias[col0+16w+i] + sum_k inT[k][lane] * W[k][col0 + 16w + i]: warp w owns 16 output columns, lane = point.
// W rows are staged through shared memory QKC at a time (broadcast reads), double-buffered with cp.async so the L2 latency of
// chunk c+1 hides behind the FMAs of chunk c (one CTA per SM: there is no other CTA to hide it); inT is read conflict-free.
__device__ __forceinline__ void stage_weights(float* dst, const float* __restrict__ Wg, int ldw, int col0, int k0, int kc) {
  for (int i = threadIdx.x; i < kc * (QH / 4); i += 256) {
    const int r = i / (QH / 4), c4 = i % (QH / 4);
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + (size_t)i * 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(Wg + (size_t)(k0 + r) * ldw + col0 + c4 * 4) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void dense16(const float* inT, int K, const float* __restrict__ Wg, int ldw, int col0,
                                        const float* __restrict__ bias, float* sW /*[2][QKC][QH]*/, float (&acc)[16]) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = bias ? bias[col0 + warp * 16 + i] : 0.f;
  const int nchunks = (K + QKC - 1) / QKC;
  stage_weights(sW, Wg, ldw, col0, 0, min(QKC, K));
  for (int c = 0; c < nchunks; ++c) {
    const int k0 = c * QKC, kc = min(QKC, K - k0);
    if (c + 1 < nchunks) {
      stage_weights(sW + ((c + 1) & 1) * QKC * QH, Wg, ldw, col0, k0 + QKC, min(QKC, K - k0 - QKC));
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();                                    // chunk c landed for every thread (also orders earlier tile writes)
    const float* w = sW + (c & 1) * QKC * QH;
#pragma unroll 4
    for (int kk = 0; kk < kc; ++kk) {
      const float a = inT[(size_t)(k0 + kk) * QLD + lane];
      const float4* wr = reinterpret_cast<const float4*>(w + kk * QH + warp * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 ww = wr[j];
        acc[j * 4 + 0] = fmaf(a, ww.x, acc[j * 4 + 0]); acc[j * 4 + 1] = fmaf(a, ww.y, acc[j * 4 + 1]);
        acc[j * 4 + 2] = fmaf(a, ww.z, acc[j * 4 + 2]); acc[j * 4 + 3] = fmaf(a, ww.w, acc[j * 4 + 3]);
      }
    }
    __syncthreads();                                    // buffer (c & 1) may be refilled by the next iteration's prefetch
  }
}

// hidden layer: outT = relu(dense); returns the 16-bit mask of active units owned by this thread
__device__ __forceinline__ unsigned hidden_layer(const float* inT, int K, const float* W, const float* b, float* outT, float* sW) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[16];
  dense16(inT, K, W, QH, 0, b, sW, acc);
  unsigned mask = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool on = acc[i] > 0.f;
    mask |= (on ? 1u : 0u) << i;
    outT[(size_t)(warp * 16 + i) * QLD + lane] = on ? acc[i] : 0.f;
  }
  return mask;
}

struct HeadW { const float *W1, *b1, *W2, *b2, *W3, *b3, *W4, *b4; };
__device__ __forceinline__ HeadW head_weights(const float* wpack, int h) {
  HeadW w;
  w.W1 = wpack + (size_t)h * q_head_stride();
  w.b1 = w.W1 + QK * QH;
  w.W2 = w.b1 + QH;  w.b2 = w.W2 + QH * QH;
  w.W3 = w.b2 + QH;  w.b3 = w.W3 + QH * QH;
  w.W4 = w.b3 + QH;  w.b4 = w.W4 + QH * 16;
  return w;
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
  float* sW = hB + QH * QLD;                 // [2][32][128]
  __shared__ int s_in_img[QP];
  const int b = blockIdx.y, n0 = blockIdx.x * QP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  gather_tile(points, crop_center, body_center, B, N, b, n0, m, cam, featT, s_in_img, xy_out);
  __syncthreads();

  // optional [B, 611, N] feature dump in the reference's channel order (CHORETriplane.query_features)
  if (feat_out && n0 + lane < N) {
    const int n_im = m.c_im, n_rest = m.c_tmpx + 3 * m.c_tt + 3 * m.c_tf;      // 256, 352
    for (int k = warp; k < n_im + 3 + n_rest; k += 8) {
      int src = k < n_im ? k : (k < n_im + 3 ? n_im + n_rest + (k - n_im) : k - 3);
      feat_out[((size_t)b * (n_im + 3 + n_rest) + k) * N + n0 + lane] = featT[(size_t)src * QLD + lane];
    }
  }
  if (!out) return;

  for (int h = 0; h < Q_NHEAD; ++h) {
    const HeadW w = head_weights(wpack, h);
    hidden_layer(featT, QK, w.W1, w.b1, hA, sW);
    hidden_layer(hA, QH, w.W2, w.b2, hB, sW);
    hidden_layer(hB, QH, w.W3, w.b3, hA, sW);
    __syncthreads();
    const int nout = c_head_nout[h], off = c_head_off[h];
    for (int c = warp; c < nout; c += 8) {
      float acc = w.b4[c];
#pragma unroll 8
      for (int k = 0; k < QH; ++k) acc = fmaf(hA[(size_t)k * QLD + lane], w.W4[k * 16 + c], acc);
      if (h == 4) acc = 1.f / (1.f + expf(-acc));                       // nn.Sigmoid on the visibility head
      if (h == 0 && !s_in_img[lane]) acc = cam.out_dist;                // chore_triplane.py:156-159
      if (n0 + lane < N) out[((size_t)b * Q_NOUT + off + c) * N + n0 + lane] = acc;
    }
  }
}

// d(sum g_out * out)/d(points).  Forward activations are recomputed in shared memory (ReLU masks kept in registers).
__global__ void __launch_bounds__(256, 1) query_bwd_kernel(const float* __restrict__ points, const float* __restrict__ crop_center,
                                                           const float* __restrict__ body_center, int B, int N, QueryMaps m,
                                                           QueryCam cam, const float* __restrict__ wpack,
                                                           const float* __restrict__ wpack_bwd,
                                                           const float* __restrict__ g_out /*[B][29][N]*/,
                                                           float* __restrict__ g_points /*[B][N][3]*/,
                                                           int mode, int df_idx, float threshold, int head_mask,
                                                           float* __restrict__ out /*[B][29][N] or null*/,
                                                           float* __restrict__ points_out /*[B][N][3] (mode 1)*/) {
  // mode 0: generic vector-Jacobian product with the cotangent g_out.
  // mode 1: one step of Generator.approx_surface (recon/gen/generator.py:86-98): cotangent = d sum(clamp(df[df_idx], max=thr)),
  //         points_out = p - normalize(grad) * clamp(df, max=thr); `out` (optional) receives the predictions at p.
  extern __shared__ float smem[];
  float* featT = smem;                       // [616][33]
  float* gfeatT = featT + QK * QLD;          // [616][33]   gradient w.r.t. the features, summed over the heads
  float* hA = gfeatT + QK * QLD;             // [128][33]
  float* hB = hA + QH * QLD;                 // [128][33]
  float* sW = hB + QH * QLD;                 // [2][32][128]
  float* g4s = sW + 2 * QKC * QH;            // [16][33]
  __shared__ int s_in_img[QP];
  __shared__ float s_df[QP];
  const int b = blockIdx.y, n0 = blockIdx.x * QP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  gather_tile(points, crop_center, body_center, B, N, b, n0, m, cam, featT, s_in_img, nullptr);
  for (int i = tid; i < QK * QLD; i += 256) gfeatT[i] = 0.f;
  __syncthreads();

  for (int h = 0; h < Q_NHEAD; ++h) {
    if (mode == 1 && h > 0 && !out) break;            // only the distance head carries gradient in a projection step
    if (mode == 0 && !((head_mask >> h) & 1)) continue;       // heads whose cotangent is identically zero are skipped
    const HeadW w = head_weights(wpack, h);
    const float* W1b = wpack_bwd + (size_t)h * q_head_stride_bwd();
    const float* W2b = W1b + QH * QKB;
    const float* W3b = W2b + QH * QH;
    const float* W4b = W3b + QH * QH;
    const unsigned m1 = hidden_layer(featT, QK, w.W1, w.b1, hA, sW);
    const unsigned m2 = hidden_layer(hA, QH, w.W2, w.b2, hB, sW);
    const unsigned m3 = hidden_layer(hB, QH, w.W3, w.b3, hA, sW);
    __syncthreads();
    // gradient at the head outputs (sigmoid derivative for visibility; zero for out-of-image df and tail points)
    const int nout = c_head_nout[h], off = c_head_off[h];
    for (int c = warp; c < 16; c += 8) {
      float g = 0.f;
      if (c < nout && n0 + lane < N) {
        float val = 0.f;
        if (h == 4 || mode == 1) {                    // the head output itself is needed
          float acc = w.b4[c];
#pragma unroll 8
          for (int k = 0; k < QH; ++k) acc = fmaf(hA[(size_t)k * QLD + lane], w.W4[k * 16 + c], acc);
          val = h == 4 ? 1.f / (1.f + expf(-acc)) : acc;
          if (h == 0 && !s_in_img[lane]) val = cam.out_dist;
        }
        if (mode == 0) {
          g = g_out[((size_t)b * Q_NOUT + off + c) * N + n0 + lane];
          if (h == 4) g *= val * (1.f - val);
        } else {
          if (out) out[((size_t)b * Q_NOUT + off + c) * N + n0 + lane] = val;
          if (h == 0 && c == df_idx) { g = val <= threshold ? 1.f : 0.f; s_df[lane] = fminf(val, threshold); }
        }
        if (h == 0 && !s_in_img[lane]) g = 0.f;
      }
      g4s[c * QLD + lane] = g;
    }
    __syncthreads();
    if (mode == 1 && h > 0) continue;                 // forward-only heads (predictions requested)
    // layer 4 backward: gh3[k] = relu'(h3[k]) * sum_c g4[c] * W4b[c][k]
    {
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      for (int c = 0; c < nout; ++c) {
        const float g = g4s[c * QLD + lane];
        const float4* wr = reinterpret_cast<const float4*>(W4b + c * QH + warp * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 ww = wr[j];
          acc[j * 4 + 0] = fmaf(g, ww.x, acc[j * 4 + 0]); acc[j * 4 + 1] = fmaf(g, ww.y, acc[j * 4 + 1]);
          acc[j * 4 + 2] = fmaf(g, ww.z, acc[j * 4 + 2]); acc[j * 4 + 3] = fmaf(g, ww.w, acc[j * 4 + 3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) hB[(size_t)(warp * 16 + i) * QLD + lane] = ((m3 >> i) & 1u) ? acc[i] : 0.f;
    }
    {   // layer 3 backward -> hA
      float acc[16];
      dense16(hB, QH, W3b, QH, 0, nullptr, sW, acc);
#pragma unroll
      for (int i = 0; i < 16; ++i) hA[(size_t)(warp * 16 + i) * QLD + lane] = ((m2 >> i) & 1u) ? acc[i] : 0.f;
    }
    {   // layer 2 backward -> hB
      float acc[16];
      dense16(hA, QH, W2b, QH, 0, nullptr, sW, acc);
#pragma unroll
      for (int i = 0; i < 16; ++i) hB[(size_t)(warp * 16 + i) * QLD + lane] = ((m1 >> i) & 1u) ? acc[i] : 0.f;
    }
    // layer 1 backward: gfeat[k] += sum_c gh1[c] * W1b[c][k], 128 feature columns at a time
    for (int col0 = 0; col0 < QKB; col0 += QH) {
      float acc[16];
      dense16(hB, QH, W1b, QKB, col0, nullptr, sW, acc);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = col0 + warp * 16 + i;
        if (k < QK) gfeatT[(size_t)k * QLD + lane] += acc[i];
      }
    }
  }
  __syncthreads();

  // features -> points: bilinear-sample derivatives, projection Jacobians, the direct (x, y, z - z0) inputs
  for (int pp = warp; pp < QP; pp += 8) {
    const int n = n0 + pp;
    if (n >= N) continue;
    const PointProj q = project_point(points + ((size_t)b * N + n) * 3, crop_center + b * 2, body_center + b * 3, cam);
    float gu = 0.f, gv = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
    gather_bilinear_grad(m.im_feat + (size_t)b * m.Hf * m.Wf * m.c_im, m.Hf, m.Wf, m.c_im, q.nx, q.ny, gfeatT, 0, pp, lane, gu, gv);
    gather_bilinear_grad(m.tmpx + (size_t)b * m.Ht * m.Wt * m.c_tmpx, m.Ht, m.Wt, m.c_tmpx, q.nx, q.ny, gfeatT, m.c_im, pp, lane, gu, gv);
    {   // nx = 2 (crop/2 + fx x / z + cx - ccx) / crop - 1
      const float k = 2.f / cam.crop, iz = 1.f / q.z;
      gx += gu * k * cam.fx * iz;  gz += -gu * k * cam.fx * q.x * iz * iz;
      gy += gv * k * cam.fy * iz;  gz += -gv * k * cam.fy * q.y * iz * iz;
    }
    int src = m.c_im + m.c_tmpx;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const float* base = pass == 0 ? m.tri_tmpx : m.tri_feat;
      const int H = pass == 0 ? m.Ht : m.Hf, W = pass == 0 ? m.Wt : m.Wf, C = pass == 0 ? m.c_tt : m.c_tf;
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        float tu = 0.f, tv = 0.f;
        gather_bilinear_grad(base + ((size_t)v * B + b) * H * W * C, H, W, C, q.tu[v], q.tv[v], gfeatT, src + v * C, pp, lane, tu, tv);
        if (v == 0) { gz += tu; gy += tv; }          // right: (z, y)
        else if (v == 1) { gx -= tu; gy += tv; }     // back : (-x, y)
        else { gx += tu; gz -= tv; }                 // top  : (x, -z)
      }
      src += 3 * C;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      gx += __shfl_xor_sync(0xffffffffu, gx, o); gy += __shfl_xor_sync(0xffffffffu, gy, o); gz += __shfl_xor_sync(0xffffffffu, gz, o);
    }
    if (lane == 0) {
      gx += gfeatT[(size_t)(src + 0) * QLD + pp]; gy += gfeatT[(size_t)(src + 1) * QLD + pp]; gz += gfeatT[(size_t)(src + 2) * QLD + pp];
      if (g_points) { float* gp = g_points + ((size_t)b * N + n) * 3; gp[0] = gx; gp[1] = gy; gp[2] = gz; }
      if (mode == 1) {     // samples - F.normalize(gradient, dim=2) * df_target   (eps 1e-12, generator.py:96)
        const float inv = 1.f / fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-12f), d = s_df[pp];
        float* po = points_out + ((size_t)b * N + n) * 3;
        po[0] = q.x - gx * inv * d; po[1] = q.y - gy * inv * d; po[2] = q.z - gz * inv * d;
      }
    }
  }
}

}  // namespace vt

using namespace vt;

static int check_layout(const char* who, int c_im, int c_tmpx, int c_tt, int c_tf) {
  VT_CHECK_ARG(c_im % 4 == 0 && c_tmpx % 4 == 0 && c_tt % 4 == 0 && c_tf % 4 == 0 && c_im + c_tmpx + 3 * c_tt + 3 * c_tf + 3 == 611,
               "%s: only the 611-feature tri-vis layout is built (got %d/%d/%d/%d)", who, c_im, c_tmpx, c_tt, c_tf);
  return 0;
}

extern "C" {

long long vt_query_wpack_floats(void) { return (long long)Q_NHEAD * q_head_stride(); }
long long vt_query_wpack_bwd_floats(void) { return (long long)Q_NHEAD * q_head_stride_bwd(); }

int vt_query_fwd(const float* points, const float* crop_center, const float* body_center, int B, int N,
                 const float* im_feat, const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht,
                 int Wt, int c_im, int c_tmpx, int c_tt, int c_tf, const float* cam7, const float* wpack, float* out,
                 float* feat_out, float* xy_out, void* stream) {
  if (int rc = check_layout("vt_query_fwd", c_im, c_tmpx, c_tt, c_tf)) return rc;
  if (B <= 0 || N <= 0) return 0;
  QueryMaps m{im_feat, tmpx, tri_tmpx, tri_feat, Hf, Wf, Ht, Wt, c_im, c_tmpx, c_tt, c_tf};
  QueryCam cam{cam7[0], cam7[1], cam7[2], cam7[3], cam7[4], cam7[5], cam7[6]};
  size_t smem = (size_t)(QK * QLD + 2 * QH * QLD + 2 * QKC * QH) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(query_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "vt_query_fwd smem attr");
  dim3 grid(ceil_div(N, QP), B);
  query_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(points, crop_center, body_center, B, N, m, cam, wpack, out, feat_out, xy_out);
  VT_CHECK_LAUNCH("vt_query_fwd");
  return 0;
}

int vt_query_bwd(const float* points, const float* crop_center, const float* body_center, int B, int N,
                 const float* im_feat, const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht,
                 int Wt, int c_im, int c_tmpx, int c_tt, int c_tf, const float* cam7, const float* wpack, const float* wpack_bwd,
                 const float* g_out, float* g_points, void* stream) {
  if (int rc = check_layout("vt_query_bwd", c_im, c_tmpx, c_tt, c_tf)) return rc;
  if (B <= 0 || N <= 0) return 0;
  QueryMaps m{im_feat, tmpx, tri_tmpx, tri_feat, Hf, Wf, Ht, Wt, c_im, c_tmpx, c_tt, c_tf};
  QueryCam cam{cam7[0], cam7[1], cam7[2], cam7[3], cam7[4], cam7[5], cam7[6]};
  size_t smem = (size_t)(2 * QK * QLD + 2 * QH * QLD + 2 * QKC * QH + 16 * QLD) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(query_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "vt_query_bwd smem attr");
  dim3 grid(ceil_div(N, QP), B);
  query_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(points, crop_center, body_center, B, N, m, cam, wpack, wpack_bwd, g_out, g_points,
                                                              0, 0, 0.f, 31, nullptr, nullptr);
  VT_CHECK_LAUNCH("vt_query_bwd");
  return 0;
}

int vt_query_bwd_heads(const float* points, const float* crop_center, const float* body_center, int B, int N,
                       const float* im_feat, const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht,
                       int Wt, int c_im, int c_tmpx, int c_tt, int c_tf, const float* cam7, const float* wpack, const float* wpack_bwd,
                       const float* g_out, int head_mask, float* g_points, void* stream) {
  if (int rc = check_layout("vt_query_bwd_heads", c_im, c_tmpx, c_tt, c_tf)) return rc;
  VT_CHECK_ARG(head_mask >= 0 && head_mask < 32, "vt_query_bwd_heads: head mask %d", head_mask);
  if (B <= 0 || N <= 0) return 0;
  QueryMaps m{im_feat, tmpx, tri_tmpx, tri_feat, Hf, Wf, Ht, Wt, c_im, c_tmpx, c_tt, c_tf};
  QueryCam cam{cam7[0], cam7[1], cam7[2], cam7[3], cam7[4], cam7[5], cam7[6]};
  size_t smem = (size_t)(2 * QK * QLD + 2 * QH * QLD + 2 * QKC * QH + 16 * QLD) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(query_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "vt_query_bwd_heads smem attr");
  dim3 grid(ceil_div(N, QP), B);
  query_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(points, crop_center, body_center, B, N, m, cam, wpack, wpack_bwd, g_out, g_points,
                                                              0, 0, 0.f, head_mask, nullptr, nullptr);
  VT_CHECK_LAUNCH("vt_query_bwd_heads");
  return 0;
}

int vt_query_project_step(const float* points, const float* crop_center, const float* body_center, int B, int N,
                          const float* im_feat, const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht,
                          int Wt, int c_im, int c_tmpx, int c_tt, int c_tf, const float* cam7, const float* wpack, const float* wpack_bwd,
                          int df_idx, float threshold, float* points_out, float* out, float* g_points, void* stream) {
  if (int rc = check_layout("vt_query_project_step", c_im, c_tmpx, c_tt, c_tf)) return rc;
  VT_CHECK_ARG(df_idx == 0 || df_idx == 1, "vt_query_project_step: df_idx %d (0 human, 1 object)", df_idx);
  VT_CHECK_ARG(points_out != nullptr, "vt_query_project_step: points_out is required");
  if (B <= 0 || N <= 0) return 0;
  QueryMaps m{im_feat, tmpx, tri_tmpx, tri_feat, Hf, Wf, Ht, Wt, c_im, c_tmpx, c_tt, c_tf};
  QueryCam cam{cam7[0], cam7[1], cam7[2], cam7[3], cam7[4], cam7[5], cam7[6]};
  size_t smem = (size_t)(2 * QK * QLD + 2 * QH * QLD + 2 * QKC * QH + 16 * QLD) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(query_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "vt_query_project_step smem attr");
  dim3 grid(ceil_div(N, QP), B);
  query_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(points, crop_center, body_center, B, N, m, cam, wpack, wpack_bwd, nullptr, g_points,
                                                              1, df_idx, threshold, 31, out, points_out);
  VT_CHECK_LAUNCH("vt_query_project_step");
  return 0;
}

}  // extern "C"
