// CUDA-core convolutions of the encoder:
//  * stem_conv_kernel  -- the 7x7 stride-2 stem (model/HGFilters.py:120,167) read straight from the NCHW input frames.
//    Cin is 5 (RGB + 2 masks) or 1 (one triplane view): K = 245 / 49 is too thin and too ragged for the tensor
//    pipe, and the layer is 0.3 % of the encoder FLOPs, so it stays on FFMA.
//  * conv_ffma_kernel  -- generic 1x1 / 3x3 stride-1 NHWC implicit GEMM in fp32 with the GroupNorm affine + ReLU of
//    the input fused into the tile load.  It serves feature maps too small for the 128-pixel tcgen05 tiles and is
//    the on-device cross-check for the tensor-core kernel (tests/test_gpu_conv.py).
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

// ------------------------------------------------------------------------------------------------ stem
constexpr int ST_TILE = 16;                       // 16x16 output pixels per CTA
constexpr int ST_PATCH = ST_TILE * 2 + 5;         // 37x37 input pixels

template <int COUT>
__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ images, int B, int Ctot, int Hin, int Win,
                                                        int c_off, int cin, const float* __restrict__ w /*[49*cin][COUT]*/,
                                                        const float* __restrict__ bias, float* __restrict__ out,
                                                        double* __restrict__ stats, int ld_stats) {
  extern __shared__ float smem[];
  float* sw = smem;                                // [49*cin][COUT]
  float* sp = smem + 49 * cin * COUT;              // [cin][37][37+1]
  __shared__ float ssum[COUT], ssq[COUT];
  const int n = blockIdx.z, b = n % B, view = n / B;
  const int Ho = Hin / 2, Wo = Win / 2;
  const int oy0 = blockIdx.y * ST_TILE, ox0 = blockIdx.x * ST_TILE;
  const int tid = threadIdx.x;
  for (int i = tid; i < 49 * cin * COUT; i += 256) sw[i] = w[i];
  if (tid < COUT) { ssum[tid] = 0.f; ssq[tid] = 0.f; }
  const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
  for (int i = tid; i < cin * ST_PATCH * ST_PATCH; i += 256) {
    int ci = i / (ST_PATCH * ST_PATCH), r = i % (ST_PATCH * ST_PATCH);
    int py = r / ST_PATCH, px = r % ST_PATCH;
    int iy = iy0 + py, ix = ix0 + px;
    float v = 0.f;
    if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win)
      v = images[(((size_t)b * Ctot + c_off + view * cin + ci) * Hin + iy) * Win + ix];
    sp[(ci * ST_PATCH + py) * (ST_PATCH + 1) + px] = v;
  }
  __syncthreads();
  const int ty = tid / ST_TILE, tx = tid % ST_TILE;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = bias ? bias[c] : 0.f;
  for (int ci = 0; ci < cin; ++ci) {
    for (int ky = 0; ky < 7; ++ky) {
      const float* prow = sp + (ci * ST_PATCH + ty * 2 + ky) * (ST_PATCH + 1) + tx * 2;
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        float v = prow[kx];
        const float4* wr = reinterpret_cast<const float4*>(sw + ((ky * 7 + kx) * cin + ci) * COUT);
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; ++c4) {
          float4 ww = wr[c4];
          acc[c4 * 4 + 0] = fmaf(v, ww.x, acc[c4 * 4 + 0]); acc[c4 * 4 + 1] = fmaf(v, ww.y, acc[c4 * 4 + 1]);
          acc[c4 * 4 + 2] = fmaf(v, ww.z, acc[c4 * 4 + 2]); acc[c4 * 4 + 3] = fmaf(v, ww.w, acc[c4 * 4 + 3]);
        }
      }
    }
  }
  const int oy = oy0 + ty, ox = ox0 + tx;
  const bool valid = oy < Ho && ox < Wo;
  if (valid) {
    float* o = out + (((size_t)n * Ho + oy) * Wo + ox) * COUT;
#pragma unroll
    for (int c4 = 0; c4 < COUT / 4; ++c4)
      st4(o + c4 * 4, make_float4(acc[c4 * 4], acc[c4 * 4 + 1], acc[c4 * 4 + 2], acc[c4 * 4 + 3]));
  }
  if (stats) {
#pragma unroll
    for (int ch = 0; ch < COUT / 32; ++ch) {
      float v[32], q[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) { float a = valid ? acc[ch * 32 + i] : 0.f; v[i] = a; q[i] = a * a; }
      float s = warp_transpose_reduce32(v), qq = warp_transpose_reduce32(q);
      atomicAdd(&ssum[ch * 32 + (tid & 31)], s);
      atomicAdd(&ssq[ch * 32 + (tid & 31)], qq);
    }
    __syncthreads();
    if (tid < COUT) {
      double* p = stats + ((size_t)n * ld_stats + tid) * 2;
      atomicAdd(p, (double)ssum[tid]);
      atomicAdd(p + 1, (double)ssq[tid]);
    }
  }
}

// ------------------------------------------------------------------------------------- generic FFMA conv
constexpr int FB_M = 64, FB_K = 16;

template <int BN>
__global__ void __launch_bounds__(256) conv_ffma_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, int relu, int H, int W, int Cin, int ks,
                                                        const float* __restrict__ w /*[ks*ks][Cin][Cout]*/, int Cout,
                                                        const float* __restrict__ bias, const float* __restrict__ res, int ldr,
                                                        float* __restrict__ out, int ldo, double* __restrict__ stats, int ld_stats) {
  constexpr int TX = BN / 4;            // threads along channels
  constexpr int TY = 256 / TX;          // threads along pixels
  constexpr int PM = FB_M / TY;         // pixels per thread (BN=64: 4, BN=32: 2)
  __shared__ __align__(16) float sA[FB_K][FB_M + 4];
  __shared__ __align__(16) float sB[FB_K][BN];
  __shared__ float ssum[BN], ssq[BN];
  const int img = blockIdx.y, HW = H * W, p0 = blockIdx.x * FB_M, n0 = blockIdx.z * BN;
  const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
  const int pad = ks / 2;
  float acc[PM][4];
#pragma unroll
  for (int i = 0; i < PM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  if (tid < BN) { ssum[tid] = 0.f; ssq[tid] = 0.f; }

  // A-load role: one float4 (4 input channels of one pixel) per thread
  const int a_px = tid / 4, a_c4 = tid % 4;
  const int ap = p0 + a_px;
  const int ay = ap / W, ax = ap % W;
  // B-load role: FB_K x BN floats = FB_K*BN/4 float4
  const int b_r = tid / (BN / 4), b_c4 = tid % (BN / 4);

  for (int tap = 0; tap < ks * ks; ++tap) {
    const int dy = tap / ks - pad, dx = tap % ks - pad;
    const int iy = ay + dy, ix = ax + dx;
    const bool a_ok = ap < HW && iy >= 0 && iy < H && ix >= 0 && ix < W;
    const float* xrow = x + ((size_t)img * HW + (size_t)iy * W + ix) * ldx;
    for (int c0 = 0; c0 < Cin; c0 += FB_K) {
      float4 v = make_float4(0, 0, 0, 0);
      const int c = c0 + a_c4 * 4;
      if (a_ok && c < Cin) {
        v = ld4(xrow + c);
        if (scale) {
          float4 sc = ld4(scale + (size_t)img * Cin + c), sh = ld4(shift + (size_t)img * Cin + c);
          v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      }
      float4 wv = make_float4(0, 0, 0, 0);
      if (b_r < FB_K && c0 + b_r < Cin) wv = ld4(w + ((size_t)tap * Cin + c0 + b_r) * Cout + n0 + b_c4 * 4);
      __syncthreads();
      sA[a_c4 * 4 + 0][a_px] = v.x; sA[a_c4 * 4 + 1][a_px] = v.y; sA[a_c4 * 4 + 2][a_px] = v.z; sA[a_c4 * 4 + 3][a_px] = v.w;
      if (b_r < FB_K) *reinterpret_cast<float4*>(&sB[b_r][b_c4 * 4]) = wv;
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < FB_K; ++kk) {
        float a[PM];
#pragma unroll
        for (int i = 0; i < PM; ++i) a[i] = sA[kk][ty * PM + i];
        float4 bv = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
#pragma unroll
        for (int i = 0; i < PM; ++i) {
          acc[i][0] = fmaf(a[i], bv.x, acc[i][0]); acc[i][1] = fmaf(a[i], bv.y, acc[i][1]);
          acc[i][2] = fmaf(a[i], bv.z, acc[i][2]); acc[i][3] = fmaf(a[i], bv.w, acc[i][3]);
        }
      }
    }
  }
  // epilogue: + bias + residual, store, statistics of the stored value
  float4 bz = make_float4(0, 0, 0, 0);
  if (bias) bz = ld4(bias + n0 + tx * 4);
  float s4[4] = {0, 0, 0, 0}, q4[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < PM; ++i) {
    int p = p0 + ty * PM + i;
    if (p < HW) {
      size_t pix = (size_t)img * HW + p;
      float4 v = make_float4(acc[i][0] + bz.x, acc[i][1] + bz.y, acc[i][2] + bz.z, acc[i][3] + bz.w);
      if (res) { float4 r = ld4(res + pix * ldr + n0 + tx * 4); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
      st4(out + pix * ldo + n0 + tx * 4, v);
      s4[0] += v.x; s4[1] += v.y; s4[2] += v.z; s4[3] += v.w;
      q4[0] += v.x * v.x; q4[1] += v.y * v.y; q4[2] += v.z * v.z; q4[3] += v.w * v.w;
    }
  }
  if (stats) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) { atomicAdd(&ssum[tx * 4 + j], s4[j]); atomicAdd(&ssq[tx * 4 + j], q4[j]); }
    __syncthreads();
    if (tid < BN) {
      double* p = stats + ((size_t)img * ld_stats + n0 + tid) * 2;
      atomicAdd(p, (double)ssum[tid]);
      atomicAdd(p + 1, (double)ssq[tid]);
    }
  }
}

}  // namespace vt

using namespace vt;

extern "C" {

int vt_stem_conv7x7s2(const float* images, int B, int Ctot, int Hin, int Win, int c_off, int cin, int n_views, const float* w,
                      const float* bias, int cout, float* out, double* stats, int ld_stats, void* stream) {
  VT_CHECK_ARG(cout == 64 || cout == 32, "vt_stem_conv7x7s2: cout=%d (only 32 / 64 are built)", cout);
  VT_CHECK_ARG(Hin % 2 == 0 && Win % 2 == 0, "vt_stem_conv7x7s2: odd input size %dx%d", Hin, Win);
  VT_CHECK_ARG(c_off + n_views * cin <= Ctot, "vt_stem_conv7x7s2: channel range exceeds the frame tensor");
  size_t smem = (size_t)(49 * cin * cout + cin * ST_PATCH * (ST_PATCH + 1)) * sizeof(float);
  VT_CHECK_ARG(smem <= 200 * 1024, "vt_stem_conv7x7s2: cin=%d needs %zu B of shared memory", cin, smem);
  dim3 grid(ceil_div(Win / 2, ST_TILE), ceil_div(Hin / 2, ST_TILE), B * n_views);
  cudaError_t e;
  if (cout == 64) {
    e = cudaFuncSetAttribute(stem_conv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "stem smem attr");
    stem_conv_kernel<64><<<grid, 256, smem, (cudaStream_t)stream>>>(images, B, Ctot, Hin, Win, c_off, cin, w, bias, out, stats, ld_stats);
  } else {
    e = cudaFuncSetAttribute(stem_conv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "stem smem attr");
    stem_conv_kernel<32><<<grid, 256, smem, (cudaStream_t)stream>>>(images, B, Ctot, Hin, Win, c_off, cin, w, bias, out, stats, ld_stats);
  }
  VT_CHECK_LAUNCH("vt_stem_conv7x7s2");
  return 0;
}

int vt_conv_ffma(const float* x, int ldx, const float* scale, const float* shift, int relu, int n_img, int H, int W, int Cin,
                 int ks, const float* w, int Cout, const float* bias, const float* res, int ldr, float* out, int ldo,
                 double* stats, int ld_stats, void* stream) {
  VT_CHECK_ARG(ks == 1 || ks == 3, "vt_conv_ffma: kernel size %d", ks);
  VT_CHECK_ARG(Cin % 4 == 0 && Cout % 32 == 0, "vt_conv_ffma: Cin=%d Cout=%d", Cin, Cout);
  int HW = H * W;
  if (Cout % 64 == 0) {
    dim3 grid(ceil_div(HW, FB_M), n_img, Cout / 64);
    conv_ffma_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, scale, shift, relu, H, W, Cin, ks, w, Cout, bias, res, ldr,
                                                                 out, ldo, stats, ld_stats);
  } else {
    dim3 grid(ceil_div(HW, FB_M), n_img, Cout / 32);
    conv_ffma_kernel<32><<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, scale, shift, relu, H, W, Cin, ks, w, Cout, bias, res, ldr,
                                                                 out, ldo, stats, ld_stats);
  }
  VT_CHECK_LAUNCH("vt_conv_ffma");
  return 0;
}

}  // extern "C"
