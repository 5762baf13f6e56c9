// Shared device helpers of the tcgen05 point-query kernels (query_tc.cu: forward, query_bwd_tc.cu: gradient w.r.t. the points).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace vt {

constexpr int TQ_M = 128;                 // points per CTA
constexpr int TQ_KC = 64;                 // features per chunk
constexpr int TQ_NCHUNK = 10;             // 640 padded features
constexpr int TQ_H = 128;                 // hidden width
constexpr int TQ_THREADS = 448;
constexpr int TQ_GATHER_WARPS = 8;
constexpr int TQ_PLANE = TQ_M * TQ_KC * 2;      // 16 KB: one fp16 plane of a [128 x 64] operand tile
constexpr int TQ_SLOT = 2 * TQ_PLANE;           // hi + lo
constexpr int TQ_NF = 2, TQ_NW = 2;             // feature-ring and weight-ring depths
constexpr int TQ_SMEM = TQ_NF * TQ_SLOT + TQ_NW * TQ_SLOT + 2 * TQ_SLOT + 1024;   // rings + activation buffer (two K chunks)
constexpr uint32_t TQ_IDESC = (1u << 4) | ((uint32_t)(TQ_H >> 3) << 17) | ((uint32_t)(TQ_M >> 4) << 24);

struct TqMaps {
  const float* im_feat; const float* tmpx; const float* tri_tmpx; const float* tri_feat;
  int Hf, Wf, Ht, Wt;
};
struct TqCam { float fx, fy, cx, cy, crop, z0, out_dist; };

__device__ __forceinline__ uint32_t tq_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tq_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void tq_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tq_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool tq_mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// spin on try_wait (which itself suspends the thread for a hardware-bounded time); the deadlock guard counts polls instead of reading the
// clock: the waiting warps share their scheduler's issue slots with the working ones, so the loop is kept to try_wait + add + branch
// (out of line: inlined, the printf set-up of ~150 wait sites was a fifth of the backward kernel's code)
static __device__ __noinline__ void tq_mbar_timeout() {
  printf("vt query_tc: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void tq_mbar_wait(uint32_t bar, uint32_t parity) {
  unsigned polls = 0;
  while (!tq_mbar_try_wait(bar, parity)) {
    if (++polls > (1u << 27)) tq_mbar_timeout();
  }
}
__device__ __forceinline__ void tq_tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// one lane of a converged warp; tcgen05.mma / commit issued under this predicate in warp-uniform control flow compile to straight-line
// UTCHMMA (under a plain `lane == 0` test every one sits in an ELECT / BRA.U.ANY loop over the active lanes)
__device__ __forceinline__ bool tq_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tq_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tq_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tq_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tq_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tq_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(TQ_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint64_t tq_desc(uint32_t saddr) {      // K-major, SWIZZLE_128B, SBO 1024 B, version 1
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tq_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// byte offset of element (row, k) inside one [128 x 64] fp16 plane in the K-major 128-byte-swizzled layout
__device__ __forceinline__ uint32_t tq_sw_off(int row, int k) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}
// x ~= hi + lo (lo not rescaled), two values per instruction (cvt.rn.f16x2.f32): ~4 issue slots per value instead of ~10.  Values
// beyond the fp16 range become inf/nan; `amax` records the largest magnitude so the kernel can raise the overflow flag.
__device__ __forceinline__ void tq_split2(float a, float b, uint32_t& hi, uint32_t& lo, float& amax) {
  amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)));
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

struct TqProj { float nx, ny, tu0, tv0, tu1, tv1, tu2, tv2; };

// bilinear sample of 2 consecutive channels at (u, v) in [-1, 1] -- grid_sample(align_corners=True, zeros padding) -- split into an
// address/weight set-up and the loads, so that the taps of several points can be in flight together (the gather is latency-bound).
struct TqTap { const float* p; float w00, w01, w10, w11, tx, ty; int rowstride; int C; unsigned valid; };

__device__ __forceinline__ TqTap tq_tap_setup(const float* __restrict__ map, int H, int W, int C, int c, float u, float v) {
  TqTap t;
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(u, 1.f), 0.5f), (float)(W - 1));
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.f), 0.5f), (float)(H - 1));
  float fx0 = floorf(ix), fy0 = floorf(iy);
  float tx = ix - fx0, ty = iy - fy0;
  bool finite = (fabsf(ix) < 1e9f) && (fabsf(iy) < 1e9f);
  int x0 = finite ? (int)fx0 : -10, y0 = finite ? (int)fy0 : -10;
  bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
  t.valid = (vy0 && vx0 ? 1u : 0u) | (vy0 && vx1 ? 2u : 0u) | (vy1 && vx0 ? 4u : 0u) | (vy1 && vx1 ? 8u : 0u);
  // clamp the base so that even masked-out taps would be in-bounds addresses (they are not dereferenced)
  t.p = map + ((long long)y0 * W + x0) * C + c;
  t.w00 = (1.f - tx) * (1.f - ty); t.w01 = tx * (1.f - ty); t.w10 = (1.f - tx) * ty; t.w11 = tx * ty;
  t.rowstride = W * C; t.C = C; t.tx = tx; t.ty = ty;
  return t;
}

// ---- per-point tap table: the bilinear footprint of a point depends on the (map size, projection) pair only -- 8 combinations:
// im_feat, tmpx (perspective xy) and tri_feat / tri_tmpx for the three views -- not on the feature chunk, so it is computed once per
// point in the prologue (12 KB of shared memory per 128 points) instead of once per point AND 64-feature chunk in the gather loops.
constexpr int TQ_NCOMBO = 8;
struct TqTapTable {
  uint32_t offv[TQ_NCOMBO][TQ_M];        // (y0 * W + x0 + W + 1) | valid-mask << 28
  float tx[TQ_NCOMBO][TQ_M], ty[TQ_NCOMBO][TQ_M];
};

__device__ __forceinline__ void tq_tap_coords(int H, int W, float u, float v, uint32_t& offv, float& tx, float& ty) {
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(u, 1.f), 0.5f), (float)(W - 1));
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.f), 0.5f), (float)(H - 1));
  float fx0 = floorf(ix), fy0 = floorf(iy);
  tx = ix - fx0; ty = iy - fy0;
  bool finite = (fabsf(ix) < 1e9f) && (fabsf(iy) < 1e9f);
  int x0 = finite ? (int)fx0 : -10, y0 = finite ? (int)fy0 : -10;
  bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
  const uint32_t valid = (vy0 && vx0 ? 1u : 0u) | (vy0 && vx1 ? 2u : 0u) | (vy1 && vx0 ? 4u : 0u) | (vy1 && vx1 ? 8u : 0u);
  const int off = valid ? y0 * W + x0 + W + 1 : 0;          // >= 0 whenever a tap is valid (x0, y0 >= -1)
  offv = (uint32_t)off | (valid << 28);
}

__device__ __forceinline__ void tq_tap_fill(TqTapTable& T, int pp, const TqProj& q, const TqMaps& m) {
  tq_tap_coords(m.Hf, m.Wf, q.nx, q.ny, T.offv[0][pp], T.tx[0][pp], T.ty[0][pp]);
  tq_tap_coords(m.Ht, m.Wt, q.nx, q.ny, T.offv[1][pp], T.tx[1][pp], T.ty[1][pp]);
  tq_tap_coords(m.Hf, m.Wf, q.tu0, q.tv0, T.offv[2][pp], T.tx[2][pp], T.ty[2][pp]);
  tq_tap_coords(m.Hf, m.Wf, q.tu1, q.tv1, T.offv[3][pp], T.tx[3][pp], T.ty[3][pp]);
  tq_tap_coords(m.Hf, m.Wf, q.tu2, q.tv2, T.offv[4][pp], T.tx[4][pp], T.ty[4][pp]);
  tq_tap_coords(m.Ht, m.Wt, q.tu0, q.tv0, T.offv[5][pp], T.tx[5][pp], T.ty[5][pp]);
  tq_tap_coords(m.Ht, m.Wt, q.tu1, q.tv1, T.offv[6][pp], T.tx[6][pp], T.ty[6][pp]);
  tq_tap_coords(m.Ht, m.Wt, q.tu2, q.tv2, T.offv[7][pp], T.tx[7][pp], T.ty[7][pp]);
}

// where feature k (multiple of 4) of chunk c comes from -- depends on (chunk, lane) only, hoisted out of the point loops.
// view: -1 perspective xy, 0 / 1 / 2 triplane right / back / top, 3 none (padding or the direct x, y, z - z0 lane)
struct TqChunkSrc { const float* base; int combo, W, C, ch, view; bool sampled, direct; };

__device__ __forceinline__ TqChunkSrc tq_chunk_src(int c, int k, const TqMaps& m, int b, int B) {
  TqChunkSrc s;
  s.sampled = true; s.direct = false;
  if (c < 4)       { s.combo = 0; s.base = m.im_feat + (size_t)b * m.Hf * m.Wf * 256; s.W = m.Wf; s.C = 256; s.ch = c * 64 + k; s.view = -1; }
  else if (c == 4) { s.combo = 1; s.base = m.tmpx + (size_t)b * m.Ht * m.Wt * 64;    s.W = m.Wt; s.C = 64;  s.ch = k;          s.view = -1; }
  else if (c < 8)  { const int v = c - 5; s.combo = 2 + v; s.base = m.tri_feat + ((size_t)v * B + b) * m.Hf * m.Wf * 64; s.W = m.Wf; s.C = 64; s.ch = k; s.view = v; }
  else if (c == 8) { const int v = k >> 5; s.combo = 5 + v; s.base = m.tri_tmpx + ((size_t)v * B + b) * m.Ht * m.Wt * 32; s.W = m.Wt; s.C = 32; s.ch = k & 31; s.view = v; }
  else {
    s.combo = 7; s.base = m.tri_tmpx + ((size_t)2 * B + b) * m.Ht * m.Wt * 32; s.W = m.Wt; s.C = 32; s.ch = k & 31; s.view = 2;
    if (k >= 32) { s.sampled = false; s.direct = (k == 32); s.view = 3; }
  }
  return s;
}

__device__ __forceinline__ TqTap tq_tap_get(const TqTapTable& T, const TqChunkSrc& s, int pp) {
  TqTap t;
  const uint32_t offv = T.offv[s.combo][pp];
  const float tx = T.tx[s.combo][pp], ty = T.ty[s.combo][pp];
  t.valid = s.sampled ? (offv >> 28) : 0u;
  const int off = (int)(offv & 0x0FFFFFFFu) - (s.W + 1);
  t.p = s.base + ((long long)off * s.C + s.ch);
  t.w00 = (1.f - tx) * (1.f - ty); t.w01 = tx * (1.f - ty); t.w10 = (1.f - tx) * ty; t.w11 = tx * ty;
  t.rowstride = s.W * s.C; t.C = s.C; t.tx = tx; t.ty = ty;
  return t;
}

// ---- host side: 2-D tensor map over an fp16 [rows][k_cols] weight plane, box {64, 128}, 128-byte swizzle
typedef CUresult (*TqEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline int tq_make_map(CUtensorMap* m, const void* base, int k_cols, int rows) {
  static TqEncodeFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (TqEncodeFn)sym;
  }
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return -2; }
  cuuint64_t dims[2] = {(cuuint64_t)k_cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)k_cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)TQ_KC, (cuuint32_t)TQ_H}, estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return -2; }
  return 0;
}

}  // namespace vt
