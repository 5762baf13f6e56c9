// HBM-bound NHWC helper kernels of the stacked-hourglass encoder (model/HGFilters.py, model/net_util.py):
// GroupNorm finalisation, GN-apply(+ReLU), fp16 hi/lo operand preparation for the tcgen05 convolution,
// residual add, 2x2 average pool and bicubic x2 upsample + add.  Every kernel that produces a tensor a later
// GroupNorm reads also accumulates that tensor's per-channel (sum, sum^2) so no separate statistics pass
// over HBM is needed.
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

constexpr int EW_THREADS = 256;
constexpr int EW_MAX_C = 1024;     // widest tensor vt_prep_split_gn normalises

// Block-level commit of per-thread float4 partial sums.  Thread t owns channel group (t % lanes) where
// lanes = C/4; threads with equal (t % lanes) are summed through shared memory and one thread per channel
// group adds into the fp64 statistics slot of this image.
__device__ __forceinline__ void commit_stats4(float4 s, float4 q, int lanes, int cg_valid, double* st /*[C][2] of this image*/) {
  __shared__ float4 red_s[EW_THREADS];
  __shared__ float4 red_q[EW_THREADS];
  const int t = threadIdx.x;
  red_s[t] = s; red_q[t] = q;
  __syncthreads();
  if (t < lanes && t < cg_valid) {
    float4 a = red_s[t], b = red_q[t];
    for (int r = t + lanes; r < EW_THREADS; r += lanes) {
      float4 x = red_s[r], y = red_q[r];
      a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
      b.x += y.x; b.y += y.y; b.z += y.z; b.w += y.w;
    }
    double* p = st + (size_t)t * 8;
    atomicAdd(p + 0, (double)a.x); atomicAdd(p + 1, (double)b.x);
    atomicAdd(p + 2, (double)a.y); atomicAdd(p + 3, (double)b.y);
    atomicAdd(p + 4, (double)a.z); atomicAdd(p + 5, (double)b.z);
    atomicAdd(p + 6, (double)a.w); atomicAdd(p + 7, (double)b.w);
  }
}

__device__ __forceinline__ void acc4(float4& s, float4& q, float4 v) {
  s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
}

// ---------------------------------------------------------------------------------------------------
// GroupNorm(groups, C) statistics -> per-(image, channel) affine:  y = x * scale + shift
// (nn.GroupNorm: biased variance, eps inside the sqrt; model/net_util.py:358-362)
__global__ void gn_finalize_kernel(const double* __restrict__ stats, int ld_stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int n_img, int C, int groups, double count, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * groups) return;
  int img = idx / groups, g = idx % groups, cpg = C / groups;
  const double* st = stats + ((size_t)img * ld_stats + (size_t)g * cpg) * 2;
  double s = 0, q = 0;
  for (int c = 0; c < cpg; ++c) { s += st[2 * c]; q += st[2 * c + 1]; }
  double mean = s / count;
  double var = q / count - mean * mean;
  if (var < 0) var = 0;
  double rstd = 1.0 / sqrt(var + (double)eps);
  for (int c = 0; c < cpg; ++c) {
    int ch = g * cpg + c;
    double sc = rstd * (double)gamma[ch];
    scale[(size_t)img * C + ch] = (float)sc;
    shift[(size_t)img * C + ch] = (float)((double)beta[ch] - mean * sc);
  }
}

// ---------------------------------------------------------------------------------------------------
// out = relu?(x * scale[img, c] + shift[img, c]) in fp32, with statistics of `out` (the encoder's `tmpx`,
// model/HGFilters.py:167-168, is both returned and normalised again by conv2.bn1).
__global__ void affine_act_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ scale,
                                  const float* __restrict__ shift, int relu, int HW, int C, int px_per_cta,
                                  float* __restrict__ out, int ldo, double* __restrict__ stats, int ld_stats) {
  const int img = blockIdx.y, lanes = C / 4, cg = threadIdx.x % lanes, row = threadIdx.x / lanes, rows = EW_THREADS / lanes;
  const int p0 = blockIdx.x * px_per_cta, p1 = min(p0 + px_per_cta, HW);
  float4 sc = make_float4(1, 1, 1, 1), sh = make_float4(0, 0, 0, 0);
  if (scale) { sc = ld4(scale + (size_t)img * C + cg * 4); sh = ld4(shift + (size_t)img * C + cg * 4); }
  float4 s = make_float4(0, 0, 0, 0), q = s;
  for (int p = p0 + row; p < p1; p += rows) {
    size_t pix = (size_t)img * HW + p;
    float4 v = ld4(x + pix * ldx + cg * 4);
    v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    st4(out + pix * ldo + cg * 4, v);
    acc4(s, q, v);
  }
  if (stats) commit_stats4(s, q, lanes, lanes, stats + (size_t)img * ld_stats * 2);
}

// ---------------------------------------------------------------------------------------------------
// Operand preparation for the tcgen05 convolution: y = relu?(x*scale+shift) split into fp16 (hi, lo*2^11)
// and written into a zero-bordered NHWC buffer [n, H+2p, W+2p, Cpad].  The kernel walks the PADDED domain
// and writes every byte (borders and channel padding as zeros), so recycled buffers need no memset.
struct GnArgs { const double* stats; int ld_stats; const float* gamma; const float* beta; int groups; double count; float eps; };

__global__ void __launch_bounds__(EW_THREADS, 3) prep_split_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ scale,
                                  const float* __restrict__ shift, const GnArgs gn, int relu, int H, int W, int C, int Cpad, int pad,
                                  __half* __restrict__ hi, __half* __restrict__ lo, int* __restrict__ overflow) {
  const int img = blockIdx.y, lanes = Cpad / 8, cg = threadIdx.x % lanes, row = threadIdx.x / lanes, rows = EW_THREADS / lanes;
  const int Wp = W + 2 * pad, Hp = H + 2 * pad, HWp = Hp * Wp;
  const int px_per_cta = ceil_div(HWp, gridDim.x);
  const int p0 = blockIdx.x * px_per_cta, p1 = min(p0 + px_per_cta, HWp);
  const bool ch_valid = cg * 8 < C;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = 1.f; sh[i] = 0.f; }
  if (scale && ch_valid) {
    float4 a = ld4(scale + (size_t)img * C + cg * 8), b = ld4(scale + (size_t)img * C + cg * 8 + 4);
    float4 c = ld4(shift + (size_t)img * C + cg * 8), d = ld4(shift + (size_t)img * C + cg * 8 + 4);
    sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w; sc[4] = b.x; sc[5] = b.y; sc[6] = b.z; sc[7] = b.w;
    sh[0] = c.x; sh[1] = c.y; sh[2] = c.z; sh[3] = c.w; sh[4] = d.x; sh[5] = d.y; sh[6] = d.z; sh[7] = d.w;
  } else if (gn.stats) {
    // GroupNorm finalisation in place (same fp64 arithmetic and summation order as gn_finalize_kernel), once per CTA: thread t
    // derives the affine of channel t into shared memory (the statistics are a few KB, always L2 hits; one latency chain per CTA)
    __shared__ float s_sc[EW_MAX_C], s_sh[EW_MAX_C];
    const int cpg = C / gn.groups;
    for (int ch = threadIdx.x; ch < C; ch += EW_THREADS) {
      const double* st = gn.stats + ((size_t)img * gn.ld_stats + (size_t)(ch / cpg) * cpg) * 2;
      double s = 0, q = 0;
      for (int c = 0; c < cpg; ++c) { s += st[2 * c]; q += st[2 * c + 1]; }
      const double mean = s / gn.count;
      double var = q / gn.count - mean * mean;
      if (var < 0) var = 0;
      const double scd = (1.0 / sqrt(var + (double)gn.eps)) * (double)gn.gamma[ch];
      s_sc[ch] = (float)scd;
      s_sh[ch] = (float)((double)gn.beta[ch] - mean * scd);
    }
    __syncthreads();
    if (ch_valid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { sc[i] = s_sc[cg * 8 + i]; sh[i] = s_sh[cg * 8 + i]; }
    }
  }
  int sat = 0;
  constexpr int U = 4;            // pixels in flight per thread: 4 x 32 B of loads before the first dependent instruction
  for (int pb = p0 + row; pb < p1; pb += U * rows) {
    float4 a[U], b[U];
    bool inside[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + u * rows;
      const int yy = p / Wp, xx = p - yy * Wp;
      const int y = yy - pad, xq = xx - pad;
      inside[u] = p < p1 && ch_valid && y >= 0 && y < H && xq >= 0 && xq < W;
      if (inside[u]) {
        const float* src = x + ((size_t)img * H * W + (size_t)y * W + xq) * ldx + cg * 8;
        a[u] = ld4(src); b[u] = ld4(src + 4);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + u * rows;
      if (p >= p1) break;
      __align__(16) __half h[8];
      __align__(16) __half l[8];
      if (inside[u]) {
        float v[8] = {a[u].x, a[u].y, a[u].z, a[u].w, b[u].x, b[u].y, b[u].z, b[u].w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float t = fmaf(v[i], sc[i], sh[i]);
          if (relu) t = fmaxf(t, 0.f);
          split_f16(t, h[i], l[i], sat);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { h[i] = __float2half_rn(0.f); l[i] = h[i]; }
      }
      size_t o = ((size_t)img * HWp + p) * Cpad + cg * 8;
      *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(h);
      *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(l);
    }
  }
  if (sat) atomicAdd(overflow, 1);
}

// ---------------------------------------------------------------------------------------------------
// out = a + b (ConvBlock identity residual, model/net_util.py:391-394) with statistics of the sum.
__global__ void add_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb, int HW, int C,
                           int px_per_cta, float* __restrict__ out, int ldo, double* __restrict__ stats, int ld_stats) {
  const int img = blockIdx.y, lanes = C / 4, cg = threadIdx.x % lanes, row = threadIdx.x / lanes, rows = EW_THREADS / lanes;
  const int p0 = blockIdx.x * px_per_cta, p1 = min(p0 + px_per_cta, HW);
  float4 s = make_float4(0, 0, 0, 0), q = s;
  for (int p = p0 + row; p < p1; p += rows) {
    size_t pix = (size_t)img * HW + p;
    float4 u = ld4(a + pix * lda + cg * 4), v = ld4(b + pix * ldb + cg * 4);
    u.x += v.x; u.y += v.y; u.z += v.z; u.w += v.w;
    st4(out + pix * ldo + cg * 4, u);
    acc4(s, q, u);
  }
  if (stats) commit_stats4(s, q, lanes, lanes, stats + (size_t)img * ld_stats * 2);
}

// ---------------------------------------------------------------------------------------------------
// F.avg_pool2d(x, 2, stride=2) (model/HGFilters.py:32,170) on NHWC, with statistics of the pooled tensor.
__global__ void avgpool2_kernel(const float* __restrict__ x, int H, int W, int C, int px_per_cta,
                                float* __restrict__ out, double* __restrict__ stats, int ld_stats) {
  const int img = blockIdx.y, lanes = C / 4, cg = threadIdx.x % lanes, row = threadIdx.x / lanes, rows = EW_THREADS / lanes;
  const int Ho = H / 2, Wo = W / 2, HWo = Ho * Wo;
  const int p0 = blockIdx.x * px_per_cta, p1 = min(p0 + px_per_cta, HWo);
  float4 s = make_float4(0, 0, 0, 0), q = s;
  for (int p = p0 + row; p < p1; p += rows) {
    int yo = p / Wo, xo = p % Wo;
    const float* src = x + (((size_t)img * H + 2 * yo) * W + 2 * xo) * C + cg * 4;
    float4 a = ld4(src), b = ld4(src + C), c = ld4(src + (size_t)W * C), d = ld4(src + (size_t)W * C + C);
    float4 v;
    v.x = (a.x + b.x + c.x + d.x) * 0.25f; v.y = (a.y + b.y + c.y + d.y) * 0.25f;
    v.z = (a.z + b.z + c.z + d.z) * 0.25f; v.w = (a.w + b.w + c.w + d.w) * 0.25f;
    st4(out + ((size_t)img * HWo + p) * C + cg * 4, v);
    acc4(s, q, v);
  }
  if (stats) commit_stats4(s, q, lanes, lanes, stats + (size_t)img * ld_stats * 2);
}

// ---------------------------------------------------------------------------------------------------
// out = up1 + F.interpolate(low, scale_factor=2, mode='bicubic', align_corners=True) (model/HGFilters.py:47-50).
// ATen's upsample_bicubic2d: source coordinate = dst * (in-1)/(out-1), Keys cubic kernel with A = -0.75,
// the four taps index-clamped to the border.
__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  float x = t + 1.f;  w[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
  x = t;              w[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 1.f - t;        w[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 2.f - t;        w[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}

// One thread = one 2x2 block of output pixels x 4 channels: the four outputs read a shared 5x5 source window (25 16-byte loads
// instead of 4 x 16), the horizontal pass is shared by the two output rows, and every output keeps ATen's arithmetic order
// (horizontal taps first, then vertical).
__device__ __forceinline__ float4 sel4(bool c, float4 a, float4 b) { return c ? a : b; }

__global__ void __launch_bounds__(EW_THREADS) upsample2x_add_kernel(const float* __restrict__ low, const float* __restrict__ up1, int Hl, int Wl, int C,
                                                                     int blk_per_cta, float* __restrict__ out, double* __restrict__ stats, int ld_stats) {
  const int img = blockIdx.y, lanes = C / 4, cg = threadIdx.x % lanes, row = threadIdx.x / lanes, rows = EW_THREADS / lanes;
  const int Ho = 2 * Hl, Wo = 2 * Wl, nblk = Hl * Wl;
  const int q0 = blockIdx.x * blk_per_cta, q1 = min(q0 + blk_per_cta, nblk);
  const float sy = Ho > 1 ? (float)(Hl - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(Wl - 1) / (float)(Wo - 1) : 0.f;
  float4 s = make_float4(0, 0, 0, 0), q = s;
  for (int blk = q0 + row; blk < q1; blk += rows) {
    const int Y = blk / Wl, X = blk % Wl;
    float wy[2][4], wx[2][4];
    int iy[2], ix[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const float ry = sy * (float)(2 * Y + a), rx = sx * (float)(2 * X + a);
      iy[a] = (int)floorf(ry); ix[a] = (int)floorf(rx);
      cubic_coeffs(ry - iy[a], wy[a]);
      cubic_coeffs(rx - ix[a], wx[a]);
    }
    const bool dy1 = iy[1] != iy[0], dx1 = ix[1] != ix[0];          // the second row / column of the block starts one source texel later
    const float* base = low + (size_t)img * Hl * Wl * C + cg * 4;
    float4 v[5][5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int yy = min(max(iy[0] - 1 + j, 0), Hl - 1);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int xx = min(max(ix[0] - 1 + i, 0), Wl - 1);
        v[j][i] = ld4(base + ((size_t)yy * Wl + xx) * C);
      }
    }
    float4 h[5][2];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        float4 r = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = b == 0 ? v[j][i] : sel4(dx1, v[j][i + 1], v[j][i]);
          r.x += t.x * wx[b][i]; r.y += t.y * wx[b][i]; r.z += t.z * wx[b][i]; r.w += t.w * wx[b][i];
        }
        h[j][b] = r;
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 r = a == 0 ? h[j][b] : sel4(dy1, h[j + 1][b], h[j][b]);
          acc.x += r.x * wy[a][j]; acc.y += r.y * wy[a][j]; acc.z += r.z * wy[a][j]; acc.w += r.w * wy[a][j];
        }
        const size_t o = (((size_t)img * Ho + 2 * Y + a) * Wo + 2 * X + b) * C + cg * 4;
        float4 u = ld4(up1 + o);
        u.x += acc.x; u.y += acc.y; u.z += acc.z; u.w += acc.w;
        st4(out + o, u);
        acc4(s, q, u);
      }
    }
  }
  if (stats) commit_stats4(s, q, lanes, lanes, stats + (size_t)img * ld_stats * 2);
}

static inline int pick_px_per_cta(int HW, int n_img, int lanes) {
  // aim for >= 4 waves of 148 SMs x 8 resident CTAs while keeping >= 2 rows of work per thread
  int rows = EW_THREADS / lanes;
  int px = 256;
  while (px > rows * 2 && (long)ceil_div(HW, px) * n_img < 148L * 8 * 2) px >>= 1;
  return px < rows ? rows : px;
}

}  // namespace vt

using namespace vt;

extern "C" {

int vt_gn_finalize(const double* stats, int ld_stats, const float* gamma, const float* beta, int n_img, int C, int groups,
                   long long count_per_channel, float eps, float* scale, float* shift, void* stream) {
  VT_CHECK_ARG(C % groups == 0, "vt_gn_finalize: C=%d not divisible by groups=%d", C, groups);
  int total = n_img * groups;
  double count = (double)count_per_channel * (C / groups);
  gn_finalize_kernel<<<ceil_div(total, 128), 128, 0, (cudaStream_t)stream>>>(stats, ld_stats, gamma, beta, n_img, C, groups,
                                                                               count, eps, scale, shift);
  VT_CHECK_LAUNCH("vt_gn_finalize");
  return 0;
}

int vt_affine_act(const float* x, int ldx, const float* scale, const float* shift, int relu, int n_img, int HW, int C,
                  float* out, int ldo, double* stats, int ld_stats, void* stream) {
  VT_CHECK_ARG(C % 4 == 0 && C <= 1024 && EW_THREADS % (C / 4) == 0, "vt_affine_act: unsupported C=%d", C);
  int px = pick_px_per_cta(HW, n_img, C / 4);
  dim3 grid(ceil_div(HW, px), n_img);
  affine_act_kernel<<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, scale, shift, relu, HW, C, px, out, ldo, stats, ld_stats);
  VT_CHECK_LAUNCH("vt_affine_act");
  return 0;
}

int vt_prep_split(const float* x, int ldx, const float* scale, const float* shift, int relu, int n_img, int H, int W, int C,
                  int Cpad, int pad, void* hi, void* lo, int* overflow, void* stream) {
  VT_CHECK_ARG(C % 8 == 0 && Cpad % 8 == 0 && Cpad >= C && EW_THREADS % (Cpad / 8) == 0, "vt_prep_split: unsupported C=%d Cpad=%d", C, Cpad);
  int HWp = (H + 2 * pad) * (W + 2 * pad);
  int px = pick_px_per_cta(HWp, n_img, Cpad / 8);
  dim3 grid(ceil_div(HWp, px), n_img);
  GnArgs gn{nullptr, 0, nullptr, nullptr, 1, 1.0, 0.f};
  prep_split_kernel<<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, scale, shift, gn, relu, H, W, C, Cpad, pad,
                                                                   (__half*)hi, (__half*)lo, overflow);
  VT_CHECK_LAUNCH("vt_prep_split");
  return 0;
}

int vt_prep_split_gn(const float* x, int ldx, const double* stats, int ld_stats, const float* gamma, const float* beta, int groups,
                     long long count_per_channel, float eps, int relu, int n_img, int H, int W, int C, int Cpad, int pad, void* hi,
                     void* lo, int* overflow, void* stream) {
  VT_CHECK_ARG(C % 8 == 0 && Cpad % 8 == 0 && Cpad >= C && EW_THREADS % (Cpad / 8) == 0, "vt_prep_split_gn: unsupported C=%d Cpad=%d", C, Cpad);
  VT_CHECK_ARG(stats && gamma && beta && groups > 0 && C % groups == 0, "vt_prep_split_gn: C=%d not divisible by groups=%d", C, groups);
  const int cpg = C / groups;
  VT_CHECK_ARG(C <= EW_MAX_C, "vt_prep_split_gn: C=%d exceeds %d", C, EW_MAX_C);
  int HWp = (H + 2 * pad) * (W + 2 * pad);
  int px = pick_px_per_cta(HWp, n_img, Cpad / 8);
  dim3 grid(ceil_div(HWp, px), n_img);
  GnArgs gn{stats, ld_stats, gamma, beta, groups, (double)count_per_channel * cpg, eps};
  prep_split_kernel<<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, nullptr, nullptr, gn, relu, H, W, C, Cpad, pad,
                                                                   (__half*)hi, (__half*)lo, overflow);
  VT_CHECK_LAUNCH("vt_prep_split_gn");
  return 0;
}

int vt_add(const float* a, int lda, const float* b, int ldb, int n_img, int HW, int C, float* out, int ldo, double* stats,
           int ld_stats, void* stream) {
  VT_CHECK_ARG(C % 4 == 0 && EW_THREADS % (C / 4) == 0, "vt_add: unsupported C=%d", C);
  int px = pick_px_per_cta(HW, n_img, C / 4);
  dim3 grid(ceil_div(HW, px), n_img);
  add_kernel<<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, HW, C, px, out, ldo, stats, ld_stats);
  VT_CHECK_LAUNCH("vt_add");
  return 0;
}

int vt_avgpool2(const float* x, int n_img, int H, int W, int C, float* out, double* stats, int ld_stats, void* stream) {
  VT_CHECK_ARG(C % 4 == 0 && EW_THREADS % (C / 4) == 0 && H % 2 == 0 && W % 2 == 0, "vt_avgpool2: unsupported shape %dx%dx%d", H, W, C);
  int HWo = (H / 2) * (W / 2);
  int px = pick_px_per_cta(HWo, n_img, C / 4);
  dim3 grid(ceil_div(HWo, px), n_img);
  avgpool2_kernel<<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, H, W, C, px, out, stats, ld_stats);
  VT_CHECK_LAUNCH("vt_avgpool2");
  return 0;
}

int vt_upsample2x_add(const float* low, const float* up1, int n_img, int Hl, int Wl, int C, float* out, double* stats,
                      int ld_stats, void* stream) {
  VT_CHECK_ARG(C % 4 == 0 && EW_THREADS % (C / 4) == 0, "vt_upsample2x_add: unsupported C=%d", C);
  int nblk = Hl * Wl;                                        // 2x2 output blocks
  int rows = EW_THREADS / (C / 4);
  int per = rows;                                            // one block per thread: 25 independent loads each already fill the pipes
  while ((long)ceil_div(nblk, per) * n_img > 148L * 16 && per < 8 * rows) per += rows;
  dim3 grid(ceil_div(nblk, per), n_img);
  upsample2x_add_kernel<<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(low, up1, Hl, Wl, C, per, out, stats, ld_stats);
  VT_CHECK_LAUNCH("vt_upsample2x_add");
  return 0;
}

}  // extern "C"
