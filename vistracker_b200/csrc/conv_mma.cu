// tcgen05 implicit-GEMM convolution (3x3 / 1x1, stride 1) for the stacked-hourglass encoder
// (model/net_util.py:374-396, model/HGFilters.py:26-50,186-201) -- the dense contraction that is 99.7 % of `filter`.
//
// Precision scheme ("fp16x2, 3 MMA"): the reference computes in fp32 and north_star asks for 1e-4 relative parity, which
// a single fp16/bf16/tf32 pass cannot give (10-bit mantissas -> ~1e-3).  Both operands are therefore pre-split into two
// fp16 planes, x = hi + lo * 2^-11 with lo scaled back into the normal fp16 range, and each K-step issues three
// kind::f16 MMAs into two fp32 TMEM accumulators:
//        acc0 += A_hi * B_hi          acc1 += A_hi * B_lo + A_lo * B_hi          D = acc0 + acc1 * 2^-11
// (the dropped lo*lo term is 2^-22 relative).  That is 22 mantissa bits at 3/2 the cost of one tf32 pass and half its
// operand bytes.
//
// Data layout: activations live in zero-bordered NHWC fp16 planes [n, H+2p, W+2p, Cpad] written by prep_split_kernel
// (GroupNorm-apply + ReLU + split fused there).  An output tile is 128 consecutive pixels (bw = min(W,128) columns x
// bh = 128/bw rows), so for filter tap (dy, dx) the A tile is one contiguous TMA box {64 ch, bw, bh} at (x0+dx, y0+dy)
// of the padded plane -- no im2col materialisation, the halo is the zero border.  Weights are [tap][Cout][Cin_pad]
// fp16 planes, B tile = TMA box {64, BN}.  Both land in shared memory in the 128-byte-swizzled K-major layout the UMMA
// descriptors expect.
//
// Roles (192 threads): warps 0-3 epilogue (one TMEM lane = one output pixel each), warp 4 TMA producer, warp 5 MMA
// issuer.  One output tile per CTA; STAGES-deep mbarrier ring between TMA and MMA.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

constexpr int MM_M = 128;       // pixels per tile == UMMA_M == TMEM lanes
constexpr int MM_KC = 64;       // fp16 channels per K chunk = one 128-byte swizzle row
constexpr int MM_THREADS = 192;

struct ConvMmaParams {
  int H, W, pad, ks, kchunks;   // kchunks = Cin_pad / 64
  int prefetch;                 // persistent kernel: tiles of A-operand L2 prefetch distance (0 = off)
  int dbg_nostore;              // debug (VT_CONV_DEBUG_NOSTORE=1): skip the global stores of the persistent epilogue (timing experiments only)
  long long* trace;             // debug (VT_CONV_TRACE=1): clock64 stamps of CTA 0: per tile [wait start, accumulator ready, tile drained]
  int bw, bh, tiles_x, bw_shift;   // bw is a power of two: pixel r of a tile sits at (y0 + (r >> bw_shift), x0 + (r & (bw-1)))
  int cout_total;               // rows per tap in the packed weight planes
  const float* bias; const float* res; int ldr;
  float* out; int ldo;
  double* stats; int ld_stats;
  // optional second output: out2 = conv + bias (+ res) + res2, with its own statistics (ConvBlock identity residual fused into
  // the conv that produces the slice: the raw slice feeds the next GroupNorm, the summed slice is the block output)
  float* out2; int ldo2; const float* res2; int ldr2; double* stats2; int ld_stats2;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (-> cudaErrorLaunchFailure reported to the caller), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) { printf("vt conv_mma: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Same, for waits that are expected to be long (a whole mainloop or epilogue): back off with nanosleep so that the spinning thread
// does not take issue slots from the warp that shares its scheduler (warps 0/1 share SMSPs with the TMA / MMA warps).
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  unsigned ns = 32;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(ns);
    if (ns < 256) ns <<= 1;
    if (clock64() - t0 > 4000000000LL) { printf("vt conv_mma: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// L2 prefetch of a TMA box: no shared memory, no barrier -- HBM requests in flight are no longer bounded by the ring depth
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// The MMA-issuer WARP walks its (warp-uniform) control flow as a whole and polls the barriers; tcgen05.mma / commit are issued by one lane
// elected here.  Under a plain `lane == 0` branch the compiler wraps every tcgen05.mma in an ELECT / BRA.U.ANY loop over the active lanes
// (~10 instructions per MMA); with elect.sync in converged code it emits straight-line UTCHMMA.
__device__ __forceinline__ bool tc_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  if (tc_elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16, M=128, single CTA
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (tc_elect_one())
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
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

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp): start address
// >> 4 in bits [0,14), leading byte offset (unused for swizzled K-major, canonical value 1) in [16,30), stride byte offset =
// 1024 B (one 8-row swizzle atom) >> 4 in [32,46), descriptor version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Epilogue shared by both kernels (warps 0-3).  Every TMA load has landed and every MMA has retired (bar_accum), so the pipeline
// buffers are free: the fp32 tile is staged there so that global stores / residual loads are whole, coalesced pixel rows.
template <int BN>
__device__ __forceinline__ void conv_epilogue(const ConvMmaParams& p, uint32_t bar_accum_addr, uint32_t tmem_base, float* stile,
                                              float* s_sum, float* s_sq, float* s_sum2, float* s_sq2, int img, int n0, int y0, int x0) {
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  mbar_wait(bar_accum_addr, 0);
  tc_fence_after();
  constexpr int LDT = BN + 4;
  {
    const int r = warp * 32 + lane;                             // tile row == TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
      float v[32], w[32];
      tc_ld32(lane_base + ch * 32, v);
      tc_ld32(lane_base + BN + ch * 32, w);
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(&stile[r * LDT + ch * 32 + i]) =
            make_float4(fmaf(w[i], kLoInv, v[i]), fmaf(w[i + 1], kLoInv, v[i + 1]), fmaf(w[i + 2], kLoInv, v[i + 2]), fmaf(w[i + 3], kLoInv, v[i + 3]));
    }
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");                // the four epilogue warps only
  constexpr int LPR = BN / 4;                                   // lanes per pixel row (float4 each)
  constexpr int RPI = 32 / LPR;                                 // pixel rows per warp iteration
  const int cl = lane % LPR, rsub = lane / LPR;
  float4 bz = make_float4(0, 0, 0, 0);
  if (p.bias) bz = ld4(p.bias + n0 + cl * 4);
  float4 s4 = make_float4(0, 0, 0, 0), q4 = s4, t4 = s4, u4 = s4;
#pragma unroll 4
  for (int rr = warp * 32 + rsub; rr < warp * 32 + 32; rr += RPI) {
    const int y = y0 + rr / p.bw, x = x0 + rr % p.bw;
    const size_t pix = ((size_t)img * p.H + y) * p.W + x;
    float4 v = *reinterpret_cast<const float4*>(&stile[rr * LDT + cl * 4]);
    v.x += bz.x; v.y += bz.y; v.z += bz.z; v.w += bz.w;
    if (p.res) { const float4 rv = ld4(p.res + pix * p.ldr + n0 + cl * 4); v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w; }
    st4(p.out + pix * p.ldo + n0 + cl * 4, v);
    s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
    q4.x += v.x * v.x; q4.y += v.y * v.y; q4.z += v.z * v.z; q4.w += v.w * v.w;
    if (p.out2) {
      const float4 r2 = ld4(p.res2 + pix * p.ldr2 + n0 + cl * 4);
      v.x += r2.x; v.y += r2.y; v.z += r2.z; v.w += r2.w;
      st4(p.out2 + pix * p.ldo2 + n0 + cl * 4, v);
      t4.x += v.x; t4.y += v.y; t4.z += v.z; t4.w += v.w;
      u4.x += v.x * v.x; u4.y += v.y * v.y; u4.z += v.z * v.z; u4.w += v.w * v.w;
    }
  }
  if (p.out2 && p.stats2) {
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
      t4.x += __shfl_xor_sync(0xffffffffu, t4.x, o); t4.y += __shfl_xor_sync(0xffffffffu, t4.y, o);
      t4.z += __shfl_xor_sync(0xffffffffu, t4.z, o); t4.w += __shfl_xor_sync(0xffffffffu, t4.w, o);
      u4.x += __shfl_xor_sync(0xffffffffu, u4.x, o); u4.y += __shfl_xor_sync(0xffffffffu, u4.y, o);
      u4.z += __shfl_xor_sync(0xffffffffu, u4.z, o); u4.w += __shfl_xor_sync(0xffffffffu, u4.w, o);
    }
    if (rsub == 0) {
      atomicAdd(&s_sum2[cl * 4 + 0], t4.x); atomicAdd(&s_sum2[cl * 4 + 1], t4.y); atomicAdd(&s_sum2[cl * 4 + 2], t4.z); atomicAdd(&s_sum2[cl * 4 + 3], t4.w);
      atomicAdd(&s_sq2[cl * 4 + 0], u4.x); atomicAdd(&s_sq2[cl * 4 + 1], u4.y); atomicAdd(&s_sq2[cl * 4 + 2], u4.z); atomicAdd(&s_sq2[cl * 4 + 3], u4.w);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x < BN) {
      double* st = p.stats2 + ((size_t)img * p.ld_stats2 + n0 + threadIdx.x) * 2;
      atomicAdd(st, (double)s_sum2[threadIdx.x]);
      atomicAdd(st + 1, (double)s_sq2[threadIdx.x]);
    }
  }
  if (p.stats) {
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {                        // fold the RPI row groups of the warp
      s4.x += __shfl_xor_sync(0xffffffffu, s4.x, o); s4.y += __shfl_xor_sync(0xffffffffu, s4.y, o);
      s4.z += __shfl_xor_sync(0xffffffffu, s4.z, o); s4.w += __shfl_xor_sync(0xffffffffu, s4.w, o);
      q4.x += __shfl_xor_sync(0xffffffffu, q4.x, o); q4.y += __shfl_xor_sync(0xffffffffu, q4.y, o);
      q4.z += __shfl_xor_sync(0xffffffffu, q4.z, o); q4.w += __shfl_xor_sync(0xffffffffu, q4.w, o);
    }
    if (rsub == 0) {
      atomicAdd(&s_sum[cl * 4 + 0], s4.x); atomicAdd(&s_sum[cl * 4 + 1], s4.y); atomicAdd(&s_sum[cl * 4 + 2], s4.z); atomicAdd(&s_sum[cl * 4 + 3], s4.w);
      atomicAdd(&s_sq[cl * 4 + 0], q4.x); atomicAdd(&s_sq[cl * 4 + 1], q4.y); atomicAdd(&s_sq[cl * 4 + 2], q4.z); atomicAdd(&s_sq[cl * 4 + 3], q4.w);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x < BN) {
      double* st = p.stats + ((size_t)img * p.ld_stats + n0 + threadIdx.x) * 2;
      atomicAdd(st, (double)s_sum[threadIdx.x]);
      atomicAdd(st + 1, (double)s_sq[threadIdx.x]);
    }
  }
}

template <int BN>
struct MmaCfg {
  static constexpr int STAGES = BN == 128 ? 3 : 2;                 // BN <= 64: two CTAs per SM share the shared memory
  static constexpr int MIN_CTAS = BN == 128 ? 1 : 2;
  static constexpr int A_BYTES = MM_M * MM_KC * 2;                 // 16 KB per plane
  static constexpr int B_BYTES = BN * MM_KC * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;   // + alignment slack
  static constexpr int TMEM_COLS = 2 * BN;                         // acc0 | acc1, power of two >= 64
  // cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format F16 (0) @7/@10, K-major both, N>>3 @17, M>>4 @24
  static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(MM_M >> 4) << 24);
};

template <int BN>
__global__ void __launch_bounds__(MM_THREADS, MmaCfg<BN>::MIN_CTAS)
conv_mma_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, const ConvMmaParams p) {
  using Cfg = MmaCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[Cfg::STAGES], bar_empty[Cfg::STAGES], bar_accum;
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_sum[BN], s_sq[BN], s_sum2[BN], s_sq2[BN];

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int img = blockIdx.y, n0 = blockIdx.z * BN;
  const int tile_y = blockIdx.x / p.tiles_x, tile_x = blockIdx.x % p.tiles_x;
  const int y0 = tile_y * p.bh, x0 = tile_x * p.bw;
  const int n_iter = p.ks * p.ks * p.kchunks;

  if (threadIdx.x < BN) { s_sum[threadIdx.x] = 0.f; s_sq[threadIdx.x] = 0.f; s_sum2[threadIdx.x] = 0.f; s_sq2[threadIdx.x] = 0.f; }
  if (warp == 5 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
    mbar_init(smem_u32(&bar_accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_lo) : "memory");
  }
  if (warp == 0) {   // TMEM allocation is warp-wide; the same warp frees it at the end
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int row0 = img * (p.H + 2 * p.pad) + y0;
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&bar_full[s]);
        mbar_expect_tx(full, Cfg::STAGE_BYTES);
        const int tap = it / p.kchunks, kc = it % p.kchunks;
        const int dy = tap / p.ks, dx = tap % p.ks;
        const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES;
        tma_load_3d(sa, &tm_a_hi, full, kc * MM_KC, x0 + dx, row0 + dy);
        tma_load_3d(sa + Cfg::A_BYTES, &tm_a_lo, full, kc * MM_KC, x0 + dx, row0 + dy);
        tma_load_2d(sa + 2 * Cfg::A_BYTES, &tm_b_hi, full, kc * MM_KC, tap * p.cout_total + n0);
        tma_load_2d(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &tm_b_lo, full, kc * MM_KC, tap * p.cout_total + n0);
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer (whole warp; one elected lane issues, see tc_elect_one)
    {
      const uint32_t acc0 = tmem_base, acc1 = tmem_base + BN;
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
        mbar_wait(smem_u32(&bar_full[s]), ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES;
        const uint64_t a_hi = make_kmajor_sw128_desc(sa), a_lo = make_kmajor_sw128_desc(sa + Cfg::A_BYTES);
        const uint64_t b_hi = make_kmajor_sw128_desc(sa + 2 * Cfg::A_BYTES), b_lo = make_kmajor_sw128_desc(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
        for (int k = 0; k < MM_KC / 16; ++k) {
          const uint64_t adv = (uint64_t)(k * 32 >> 4);          // 16 fp16 = 32 bytes along K inside the swizzle row
          const uint32_t accum = (it | k) != 0;
          tc_mma_f16(acc0, a_hi + adv, b_hi + adv, Cfg::IDESC, accum);
          tc_mma_f16(acc1, a_hi + adv, b_lo + adv, Cfg::IDESC, accum);
          tc_mma_f16(acc1, a_lo + adv, b_hi + adv, Cfg::IDESC, 1u);
        }
        tc_commit(smem_u32(&bar_empty[s]));       // frees the stage once these MMAs have read it
      }
      tc_commit(smem_u32(&bar_accum));            // accumulators complete
    }
  } else {
    conv_epilogue<BN>(p, smem_u32(&bar_accum), tmem_base, reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw))), s_sum, s_sq,
                      s_sum2, s_sq2, img, n0, y0, x0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ persistent variant (default)
// One CTA per SM walks a contiguous range of tiles (pixel tiles of an image in raster order, then the next Cout tile, then the next
// image), so halo rows are re-read from L2 and consecutive tiles share (image, Cout tile): the per-channel statistics are kept in
// shared memory and flushed to the fp64 slots only when that pair changes -- ~10x fewer same-address L2 atomics.  Two TMEM accumulator sets (2 x [acc0 | acc1] = 4*BN columns) let the epilogue warps drain tile i while the MMA warp
// already accumulates tile i+1, and the TMA ring runs ahead across tile boundaries; TMEM allocation, barrier init and tensor-map
// prefetch are paid once per CTA instead of once per tile.
// The three MMAs of the fp16x2 scheme are issued as two: the weight planes B_hi | B_lo are adjacent in the stage and acc0 | acc1 are
// adjacent in TMEM, so   [acc0 | acc1] += A_hi * [B_hi | B_lo]   is ONE N = 2*BN instruction, followed by   acc1 += A_lo * B_hi.
// Same products, same accumulation order, but A_hi is read from shared memory once instead of twice (24 -> 20 KB per K-step at
// BN = 128, 18 -> 14 KB at BN = 64: the shared-memory port is what bounds the N <= 64 tiles).
// The epilogue stages one 32-row x 32-channel block per warp in a private 4.5 KB buffer, so every global access is a whole 128-byte
// pixel-channel run and no tile-sized staging buffer has to be carved out of the pipeline.
// RESB ("resident B", 1x1 convolutions with Cin <= 256): the whole weight panel of one Cout tile (<= 4 K-chunks x hi/lo) stays in
// shared memory while the CTA walks pixel tiles of that (image, Cout tile) run; the ring then carries A stages only.  A 1x1 conv
// has as many weight bytes as activation bytes per tile, so this halves its L2 -> SM traffic and gives the HBM-facing A stream the
// whole ring.
// STRIP (3x3 convolutions on maps at least 128 wide, one image row per tile): every 3x3 shape of the step turned out to be paced by
// L2 -> shared-memory delivery (~60-70 B/clk per SM, VT_CONV_TRACE), not by the tensor pipe, because the ring fetches a fresh
// 128-pixel A tile for each of the 9 taps.  Here one 130-pixel strip per (K-chunk, dy) serves the three dx taps through UMMA
// descriptor row offsets (as in conv_mma_strip_kernel below), with separate A-strip and weight rings: 1.5x (BN = 128) to 1.8x
// (BN = 64) fewer bytes per MMA.
// E8 (1x1 convolutions, opt-in experiment): a SECOND group of four epilogue warps (warps 6-9, TMEM lane quadrant = warp % 4) that drains
// the upper half of the tile's channel chunks, paid for with a 2-stage ring.  The trace says the 1x1 tiles wait on the epilogue, not on
// the MMAs -- but neither removing the global stores (VT_CONV_DEBUG_NOSTORE) nor doubling the epilogue warps changes the time: the
// shared-memory port is shared by the MMA operand reads, the TMA fills and the epilogue staging, and that sum is what paces the tile.
template <int BN, bool RESB = false, bool STRIP = false, bool E8 = false>
struct PersistCfg {
  static constexpr int A_BYTES = MM_M * MM_KC * 2;
  static constexpr int B_BYTES = BN * MM_KC * 2;
  static constexpr int PANEL_CHUNKS = 4;
  static constexpr int PANEL_BYTES = RESB ? PANEL_CHUNKS * 2 * B_BYTES : 0;
  static constexpr int STAGES = E8 ? (BN == 128 ? 2 : 3) : RESB ? (BN == 128 ? 2 : (BN == 64 ? 4 : 5)) : (BN == 128 ? 3 : (BN == 64 ? 4 : 5));
  static constexpr int EPI_WARPS = E8 ? 8 : 4;
  static constexpr int THREADS = E8 ? 320 : MM_THREADS;
  static constexpr int STAGE_BYTES = RESB ? 2 * A_BYTES : 2 * A_BYTES + 2 * B_BYTES;
  // strip mode: separate rings
  static constexpr int STRIP_BYTES = (MM_M + 2) * 128;               // one plane of a 130-pixel strip, what TMA writes
  static constexpr int STRIP_PLANE = 17408;                          // padded to a multiple of 1024
  static constexpr int STRIP_SLOT = 2 * STRIP_PLANE;
  static constexpr int B_SLOT = 2 * B_BYTES;
  static constexpr int NA = STRIP ? (BN == 128 ? 2 : 3) : 1;
  static constexpr int NB = STRIP ? (BN == 128 ? 4 : (BN == 64 ? 6 : 9)) : 1;
  static constexpr int RING_BYTES = STRIP ? NA * STRIP_SLOT + NB * B_SLOT : STAGES * STAGE_BYTES + PANEL_BYTES;
  static constexpr int EPI_LD = 36;                                  // floats per staged row: 32 + 4 keeps float4 rows conflict-free
  static constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_LD * 4;
  static constexpr int SMEM_BYTES = RING_BYTES + EPI_BYTES + 1024;
  static constexpr int TMEM_COLS = 4 * BN;
  static constexpr uint32_t IDESC_WIDE = (1u << 4) | ((uint32_t)(2 * BN >> 3) << 17) | ((uint32_t)(MM_M >> 4) << 24);
  static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(MM_M >> 4) << 24);
  static_assert(!(RESB && STRIP) && !(E8 && (RESB || STRIP)) && !(E8 && BN < 64), "resident panel / second epilogue group are for 1x1, strips for 3x3");
  static_assert(SMEM_BYTES + 16 * BN * 4 + 1024 <= 227 * 1024 && TMEM_COLS <= 512, "persistent configuration (dynamic + static shared memory) exceeds the SM");
};

struct TileCoord { int img, n0, y0, x0; };
// tiles [first, last) of CTA b out of g: balanced contiguous split
__device__ __forceinline__ void tile_range(int total, int& first, int& last) {
  const int base = total / (int)gridDim.x, rem = total % (int)gridDim.x, b = (int)blockIdx.x;
  first = b * base + min(b, rem);
  last = first + base + (b < rem ? 1 : 0);
}
__device__ __forceinline__ TileCoord decode_tile(const ConvMmaParams& p, int t, int n_tiles_n, int tiles_per_img, int BN) {
  TileCoord c;                                   // order: image, then Cout tile, then pixel tile (raster)
  const int per_img = tiles_per_img * n_tiles_n;
  c.img = t / per_img;
  const int rem = t - c.img * per_img;
  const int nt = rem / tiles_per_img, tl = rem - nt * tiles_per_img;
  c.n0 = nt * BN;
  c.y0 = (tl / p.tiles_x) * p.bh;
  c.x0 = (tl % p.tiles_x) * p.bw;
  return c;
}

__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

template <int BN, bool RESB, bool STRIP, bool E8>
__global__ void __launch_bounds__(E8 ? 320 : MM_THREADS, 1)
conv_mma_persist_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                        const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, const ConvMmaParams p,
                        int tiles_per_img, int n_tiles_n, int total_tiles) {
  using Cfg = PersistCfg<BN, RESB, STRIP, E8>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[Cfg::STAGES], bar_empty[Cfg::STAGES], acc_full[2], acc_empty[2], panel_full, panel_empty;
  __shared__ __align__(8) uint64_t a_full[Cfg::NA], a_empty[Cfg::NA], b_full[Cfg::NB], b_empty[Cfg::NB];      // strip mode rings
  __shared__ uint32_t s_tmem_base;
  // per-warp partial statistics [warp][channel]: exactly one lane owns a slot, so the running sums are plain read-modify-writes (a
  // shared-memory fp32 atomicAdd is a CAS loop, and four warps contending on it cost more than the rest of the chunk)
  __shared__ __align__(16) float s_sum[4][BN], s_sq[4][BN], s_sum2[4][BN], s_sq2[4][BN];

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int n_iter = p.ks * p.ks * p.kchunks;

  for (int i = threadIdx.x; i < 4 * BN; i += Cfg::THREADS) { (&s_sum[0][0])[i] = 0.f; (&s_sq[0][0])[i] = 0.f; (&s_sum2[0][0])[i] = 0.f; (&s_sq2[0][0])[i] = 0.f; }
  if (warp == 5 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&acc_full[s]), 1); mbar_init(smem_u32(&acc_empty[s]), Cfg::EPI_WARPS); }
    mbar_init(smem_u32(&panel_full), 1); mbar_init(smem_u32(&panel_empty), 1);
    for (int s = 0; s < Cfg::NA; ++s) { mbar_init(smem_u32(&a_full[s]), 1); mbar_init(smem_u32(&a_empty[s]), 1); }
    for (int s = 0; s < Cfg::NB; ++s) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_lo) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  int t_first, t_last;
  tile_range(total_tiles, t_first, t_last);

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer: the ring runs across tile boundaries
    if (lane == 0) {
      if constexpr (STRIP) {
        uint32_t ga = 0, gb = 0;
        const uint32_t b_base = smem_base + Cfg::NA * Cfg::STRIP_SLOT;
        const int n_strips = 3 * p.kchunks;
        for (int t = t_first; t < t_last; ++t) {
          const TileCoord c = decode_tile(p, t, n_tiles_n, tiles_per_img, BN);
          for (int is = 0; is < n_strips; ++is, ++ga) {
            const int kc = is / 3, dy = is % 3;
            const int sa = ga % Cfg::NA;
            mbar_wait_relaxed(smem_u32(&a_empty[sa]), ((ga / Cfg::NA) & 1u) ^ 1u);
            const uint32_t af = smem_u32(&a_full[sa]);
            mbar_expect_tx(af, 2 * Cfg::STRIP_BYTES);
            const uint32_t adst = smem_base + sa * Cfg::STRIP_SLOT;
            const int row = c.img * (p.H + 2) + c.y0 + dy;
            tma_load_3d(adst, &tm_a_hi, af, kc * MM_KC, c.x0, row);
            tma_load_3d(adst + Cfg::STRIP_PLANE, &tm_a_lo, af, kc * MM_KC, c.x0, row);
            for (int dx = 0; dx < 3; ++dx, ++gb) {
              const int sb = gb % Cfg::NB;
              mbar_wait_relaxed(smem_u32(&b_empty[sb]), ((gb / Cfg::NB) & 1u) ^ 1u);
              const uint32_t bf = smem_u32(&b_full[sb]);
              mbar_expect_tx(bf, Cfg::B_SLOT);
              const uint32_t bdst = b_base + sb * Cfg::B_SLOT;
              const int tap = dy * 3 + dx;
              tma_load_2d(bdst, &tm_b_hi, bf, kc * MM_KC, tap * p.cout_total + c.n0);
              tma_load_2d(bdst + Cfg::B_BYTES, &tm_b_lo, bf, kc * MM_KC, tap * p.cout_total + c.n0);
            }
          }
        }
      } else {
        uint32_t g = 0, pg = 0;
        int panel_n0 = -1;
        const uint32_t panel_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
        for (int t = t_first; t < t_last; ++t) {
          const TileCoord c = decode_tile(p, t, n_tiles_n, tiles_per_img, BN);
          const int row0 = c.img * (p.H + 2 * p.pad) + c.y0;
          // HBM-streaming shapes (1x1 convs, Cin <= 64): the ring alone keeps ~100 KB in flight per SM, not enough to cover DRAM
          // latency at full bandwidth -- pull the A boxes of the tile `prefetch` tiles ahead into L2 now
          if (p.prefetch > 0) {
            for (int tp = (t == t_first ? t + 1 : t + p.prefetch); tp <= t + p.prefetch && tp < t_last; ++tp) {
              const TileCoord cp = decode_tile(p, tp, n_tiles_n, tiles_per_img, BN);
              const int rowp = cp.img * (p.H + 2 * p.pad) + cp.y0;
              for (int it = 0; it < n_iter; ++it) {
                const int tap = it / p.kchunks, kc = it % p.kchunks;
                const int dy = tap / p.ks, dx = tap % p.ks;
                if (p.ks == 3 && dx == 1) continue;          // the dx = 0 and dx = 2 boxes already cover the 130-pixel row
                tma_prefetch_3d(&tm_a_hi, kc * MM_KC, cp.x0 + dx, rowp + dy);
                tma_prefetch_3d(&tm_a_lo, kc * MM_KC, cp.x0 + dx, rowp + dy);
              }
            }
          }
          if (RESB && c.n0 != panel_n0) {                    // (re)load the weight panel of this Cout tile once the old one is unused
            if (pg > 0) mbar_wait_relaxed(smem_u32(&panel_empty), (pg - 1) & 1u);
            const uint32_t pf = smem_u32(&panel_full);
            mbar_expect_tx(pf, (uint32_t)p.kchunks * 2 * Cfg::B_BYTES);
            for (int kc = 0; kc < p.kchunks; ++kc) {
              tma_load_2d(panel_base + kc * 2 * Cfg::B_BYTES, &tm_b_hi, pf, kc * MM_KC, c.n0);
              tma_load_2d(panel_base + kc * 2 * Cfg::B_BYTES + Cfg::B_BYTES, &tm_b_lo, pf, kc * MM_KC, c.n0);
            }
            panel_n0 = c.n0; ++pg;
          }
          for (int it = 0; it < n_iter; ++it, ++g) {
            const int s = g % Cfg::STAGES;
            const uint32_t ph = (g / Cfg::STAGES) & 1u;
            mbar_wait_relaxed(smem_u32(&bar_empty[s]), ph ^ 1u);
            const uint32_t full = smem_u32(&bar_full[s]);
            mbar_expect_tx(full, Cfg::STAGE_BYTES);
            const int tap = it / p.kchunks, kc = it % p.kchunks;
            const int dy = tap / p.ks, dx = tap % p.ks;
            const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES;
            tma_load_3d(sa, &tm_a_hi, full, kc * MM_KC, c.x0 + dx, row0 + dy);
            tma_load_3d(sa + Cfg::A_BYTES, &tm_a_lo, full, kc * MM_KC, c.x0 + dx, row0 + dy);
            if (!RESB) {
              tma_load_2d(sa + 2 * Cfg::A_BYTES, &tm_b_hi, full, kc * MM_KC, tap * p.cout_total + c.n0);
              tma_load_2d(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &tm_b_lo, full, kc * MM_KC, tap * p.cout_total + c.n0);
            }
          }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer (whole warp; one elected lane issues, see tc_elect_one)
    {
      if constexpr (STRIP) {
        uint32_t ga = 0, gb = 0, i = 0;
        const uint32_t b_base = smem_base + Cfg::NA * Cfg::STRIP_SLOT;
        const int n_strips = 3 * p.kchunks;
        for (int t = t_first; t < t_last; ++t, ++i) {
          const uint32_t set = i & 1u;
          mbar_wait_relaxed(smem_u32(&acc_empty[set]), ((i >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t acc0 = tmem_base + set * 2 * BN, acc1 = acc0 + BN;
          uint32_t started = 0;
          for (int is = 0; is < n_strips; ++is, ++ga) {
            const int sa = ga % Cfg::NA;
            mbar_wait(smem_u32(&a_full[sa]), (ga / Cfg::NA) & 1u);
            const uint32_t abase = smem_base + sa * Cfg::STRIP_SLOT;
            for (int dx = 0; dx < 3; ++dx, ++gb) {
              const int sb = gb % Cfg::NB;
              mbar_wait(smem_u32(&b_full[sb]), (gb / Cfg::NB) & 1u);
              tc_fence_after();
              // tap dx = the same strip read from row dx on (128 B per pixel row; the matrix-base-offset field stays 0, see below)
              const uint64_t a_hi = make_kmajor_sw128_desc(abase + dx * 128), a_lo = make_kmajor_sw128_desc(abase + Cfg::STRIP_PLANE + dx * 128);
              const uint64_t b_hi = make_kmajor_sw128_desc(b_base + sb * Cfg::B_SLOT);     // rows [0, BN) = B_hi, [BN, 2 BN) = B_lo
#pragma unroll
              for (int k = 0; k < MM_KC / 16; ++k) {
                const uint64_t adv = (uint64_t)(k * 32 >> 4);
                tc_mma_f16(acc0, a_hi + adv, b_hi + adv, Cfg::IDESC_WIDE, started);
                tc_mma_f16(acc1, a_lo + adv, b_hi + adv, Cfg::IDESC, 1u);
                started = 1u;
              }
              tc_commit(smem_u32(&b_empty[sb]));
            }
            tc_commit(smem_u32(&a_empty[sa]));
          }
          tc_commit(smem_u32(&acc_full[set]));
        }
      } else {
        uint32_t g = 0, i = 0, pg = 0;
        int panel_n0 = -1;
        const uint32_t panel_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
        for (int t = t_first; t < t_last; ++t, ++i) {
          const uint32_t set = i & 1u;
          mbar_wait_relaxed(smem_u32(&acc_empty[set]), ((i >> 1) & 1u) ^ 1u);      // the epilogue has drained this accumulator set
          tc_fence_after();
          int n0 = 0, n0_next = -1;
          if (RESB) {
            n0 = decode_tile(p, t, n_tiles_n, tiles_per_img, BN).n0;
            if (t + 1 < t_last) n0_next = decode_tile(p, t + 1, n_tiles_n, tiles_per_img, BN).n0;
            if (n0 != panel_n0) { mbar_wait(smem_u32(&panel_full), pg & 1u); tc_fence_after(); panel_n0 = n0; ++pg; }
          }
          const uint32_t acc0 = tmem_base + set * 2 * BN, acc1 = acc0 + BN;
          for (int it = 0; it < n_iter; ++it, ++g) {
            const int s = g % Cfg::STAGES;
            const uint32_t ph = (g / Cfg::STAGES) & 1u;
            mbar_wait(smem_u32(&bar_full[s]), ph);
            tc_fence_after();
            const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES;
            const uint64_t a_hi = make_kmajor_sw128_desc(sa), a_lo = make_kmajor_sw128_desc(sa + Cfg::A_BYTES);
            // rows [0, BN) = B_hi, rows [BN, 2 BN) = B_lo; 1x1: it == K-chunk index into the resident panel
            const uint64_t b_hi = make_kmajor_sw128_desc(RESB ? panel_base + it * 2 * Cfg::B_BYTES : sa + 2 * Cfg::A_BYTES);
  #pragma unroll
            for (int k = 0; k < MM_KC / 16; ++k) {
              const uint64_t adv = (uint64_t)(k * 32 >> 4);
              tc_mma_f16(acc0, a_hi + adv, b_hi + adv, Cfg::IDESC_WIDE, (it | k) != 0);
              tc_mma_f16(acc1, a_lo + adv, b_hi + adv, Cfg::IDESC, 1u);
            }
            tc_commit(smem_u32(&bar_empty[s]));
          }
          tc_commit(smem_u32(&acc_full[set]));
          if (RESB && n0_next != n0) tc_commit(smem_u32(&panel_empty));   // last tile of this panel (the producer may overwrite it)
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 0-3 (+ 6-9 with E8): TMEM lanes 32*(warp % 4) .. +31
    const int team = warp >= 6 ? 1 : 0, q = warp & 3, ew = team * 4 + q;      // q: TMEM lane quadrant this warp may read
    float* wstage = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + Cfg::RING_BYTES) + ew * 32 * Cfg::EPI_LD;
    const int cl = lane & 7, rsub = lane >> 3;
    const bool st1 = p.stats != nullptr, st2 = p.out2 != nullptr && p.stats2 != nullptr;
    constexpr int NCH = BN / 32;
    constexpr int CH_PER_TEAM = E8 ? NCH / 2 : NCH;                           // E8: team 0 drains the lower channel chunks, team 1 the upper
    const int ch_first = team * CH_PER_TEAM;
    // pixel offsets of this lane's 8 rows inside a tile (tile-invariant)
    int roff[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { const int r = q * 32 + rsub + 4 * j; roff[j] = (r >> p.bw_shift) * p.W + (r & (p.bw - 1)); }
    uint32_t i = 0;
    int stats_img = -1, stats_n0 = 0;
    auto flush_stats = [&]() {       // block partials (fp32, <= a few thousand values per channel) -> fp64 slots of image stats_img
      if (E8) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x < BN) {
        if (st1) {
          double* st = p.stats + ((size_t)stats_img * p.ld_stats + stats_n0 + threadIdx.x) * 2;
          const int c = threadIdx.x;
          atomicAdd(st, (double)((s_sum[0][c] + s_sum[1][c]) + (s_sum[2][c] + s_sum[3][c])));
          atomicAdd(st + 1, (double)((s_sq[0][c] + s_sq[1][c]) + (s_sq[2][c] + s_sq[3][c])));
#pragma unroll
          for (int w = 0; w < 4; ++w) { s_sum[w][c] = 0.f; s_sq[w][c] = 0.f; }
        }
        if (st2) {
          double* st = p.stats2 + ((size_t)stats_img * p.ld_stats2 + stats_n0 + threadIdx.x) * 2;
          const int c = threadIdx.x;
          atomicAdd(st, (double)((s_sum2[0][c] + s_sum2[1][c]) + (s_sum2[2][c] + s_sum2[3][c])));
          atomicAdd(st + 1, (double)((s_sq2[0][c] + s_sq2[1][c]) + (s_sq2[2][c] + s_sq2[3][c])));
#pragma unroll
          for (int w = 0; w < 4; ++w) { s_sum2[w][c] = 0.f; s_sq2[w][c] = 0.f; }
        }
      }
      if (E8) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 1, 128;" ::: "memory");
    };
    for (int t = t_first; t < t_last; ++t, ++i) {
      const TileCoord c = decode_tile(p, t, n_tiles_n, tiles_per_img, BN);
      if ((st1 || st2) && (c.img != stats_img || c.n0 != stats_n0)) {
        if (stats_img >= 0) flush_stats();
        stats_img = c.img; stats_n0 = c.n0;
      }
      const size_t pix0 = ((size_t)c.img * p.H + c.y0) * p.W + c.x0;
      const int cbase = c.n0 + ch_first * 32 + cl * 4;
      // residual operands are fetched one 32-channel chunk ahead (the first one before the accumulator is even ready): the loads
      // must not sit between dependent stores, or every row pays a full HBM round trip (out may alias res, so the compiler keeps
      // the program order)
      float4 rv[8], rw[8];
      if (p.res) {
#pragma unroll
        for (int j = 0; j < 8; ++j) rv[j] = ld4(p.res + (pix0 + roff[j]) * p.ldr + cbase);
      }
      if (p.out2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) rw[j] = ld4(p.res2 + (pix0 + roff[j]) * p.ldr2 + cbase);
      }
      // ... and the NEXT tile's residual lines are pulled into L2 now, a whole mainloop ahead of their use (one lane per 128-byte run)
      if ((p.res || p.out2) && cl == 0 && t + 1 < t_last) {
        const TileCoord cn = decode_tile(p, t + 1, n_tiles_n, tiles_per_img, BN);
        const size_t pixn = ((size_t)cn.img * p.H + cn.y0) * p.W + cn.x0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int lc = 0; lc < CH_PER_TEAM; ++lc) {
            const int ch = ch_first + lc;
            if (p.res) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res + (pixn + roff[j]) * p.ldr + cn.n0 + ch * 32));
            if (p.out2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res2 + (pixn + roff[j]) * p.ldr2 + cn.n0 + ch * 32));
          }
        }
      }
      const uint32_t set = i & 1u;
      const bool tracing = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && i < 10;
      if (tracing) p.trace[3 * i] = clock64();
      mbar_wait_relaxed(smem_u32(&acc_full[set]), (i >> 1) & 1u);
      tc_fence_after();
      if (tracing) p.trace[3 * i + 1] = clock64();
      const uint32_t lane_base = tmem_base + set * 2 * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int lc = 0; lc < CH_PER_TEAM; ++lc) {
        const int ch = ch_first + lc;
        const bool micro = tracing && i == 2 && ch == 0;
        if (micro) p.trace[32] = clock64();
        {
          uint32_t v[32], w[32];
          tc_ld32_issue(lane_base + ch * 32, v);
          tc_ld32_issue(lane_base + BN + ch * 32, w);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (micro) p.trace[33] = clock64();
          if (lc == CH_PER_TEAM - 1) {                                   // this warp's last TMEM read of the tile: hand the set back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[set]));
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(&wstage[lane * Cfg::EPI_LD + j]) =
                make_float4(fmaf(__uint_as_float(w[j]), kLoInv, __uint_as_float(v[j])), fmaf(__uint_as_float(w[j + 1]), kLoInv, __uint_as_float(v[j + 1])),
                            fmaf(__uint_as_float(w[j + 2]), kLoInv, __uint_as_float(v[j + 2])), fmaf(__uint_as_float(w[j + 3]), kLoInv, __uint_as_float(v[j + 3])));
        }
        __syncwarp();
        if (micro) p.trace[34] = clock64();
        const int cc = ch * 32 + cl * 4;                                 // channel inside the tile
        float4 bz = make_float4(0, 0, 0, 0);
        if (p.bias) bz = ld4(p.bias + c.n0 + cc);
        float4 val[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v = *reinterpret_cast<const float4*>(&wstage[(rsub + 4 * j) * Cfg::EPI_LD + cl * 4]);
          v.x += bz.x; v.y += bz.y; v.z += bz.z; v.w += bz.w;
          if (p.res) { v.x += rv[j].x; v.y += rv[j].y; v.z += rv[j].z; v.w += rv[j].w; }
          val[j] = v;
        }
        __syncwarp();                                                    // wstage is rewritten by the next chunk
        if (micro) p.trace[35] = clock64();
        float4 nv[8], nw[8];
        if (lc + 1 < CH_PER_TEAM) {                                      // next chunk's residuals: in flight during this chunk's stores
          if (p.res) {
#pragma unroll
            for (int j = 0; j < 8; ++j) nv[j] = ld4(p.res + (pix0 + roff[j]) * p.ldr + cbase + (lc + 1) * 32);
          }
          if (p.out2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) nw[j] = ld4(p.res2 + (pix0 + roff[j]) * p.ldr2 + cbase + (lc + 1) * 32);
          }
        }
        if (micro) p.trace[36] = clock64();
        float4 s4 = make_float4(0, 0, 0, 0), q4 = s4, t4 = s4, u4 = s4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v = val[j];
          if (!p.dbg_nostore) st4(p.out + (pix0 + roff[j]) * p.ldo + c.n0 + cc, v);
          s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
          q4.x += v.x * v.x; q4.y += v.y * v.y; q4.z += v.z * v.z; q4.w += v.w * v.w;
          if (p.out2) {
            v.x += rw[j].x; v.y += rw[j].y; v.z += rw[j].z; v.w += rw[j].w;
            st4(p.out2 + (pix0 + roff[j]) * p.ldo2 + c.n0 + cc, v);
            t4.x += v.x; t4.y += v.y; t4.z += v.z; t4.w += v.w;
            u4.x += v.x * v.x; u4.y += v.y * v.y; u4.z += v.z * v.z; u4.w += v.w * v.w;
          }
        }
        if (lc + 1 < CH_PER_TEAM) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { rv[j] = nv[j]; rw[j] = nw[j]; }
        }
        if (micro) p.trace[37] = clock64();
        if (st1) {
#pragma unroll
          for (int o = 8; o < 32; o <<= 1) {
            s4.x += __shfl_xor_sync(0xffffffffu, s4.x, o); s4.y += __shfl_xor_sync(0xffffffffu, s4.y, o);
            s4.z += __shfl_xor_sync(0xffffffffu, s4.z, o); s4.w += __shfl_xor_sync(0xffffffffu, s4.w, o);
            q4.x += __shfl_xor_sync(0xffffffffu, q4.x, o); q4.y += __shfl_xor_sync(0xffffffffu, q4.y, o);
            q4.z += __shfl_xor_sync(0xffffffffu, q4.z, o); q4.w += __shfl_xor_sync(0xffffffffu, q4.w, o);
          }
          if (rsub == 0) {
            float4 a = *reinterpret_cast<float4*>(&s_sum[q][cc]), bq = *reinterpret_cast<float4*>(&s_sq[q][cc]);
            a.x += s4.x; a.y += s4.y; a.z += s4.z; a.w += s4.w;
            bq.x += q4.x; bq.y += q4.y; bq.z += q4.z; bq.w += q4.w;
            *reinterpret_cast<float4*>(&s_sum[q][cc]) = a; *reinterpret_cast<float4*>(&s_sq[q][cc]) = bq;
          }
        }
        if (st2) {
#pragma unroll
          for (int o = 8; o < 32; o <<= 1) {
            t4.x += __shfl_xor_sync(0xffffffffu, t4.x, o); t4.y += __shfl_xor_sync(0xffffffffu, t4.y, o);
            t4.z += __shfl_xor_sync(0xffffffffu, t4.z, o); t4.w += __shfl_xor_sync(0xffffffffu, t4.w, o);
            u4.x += __shfl_xor_sync(0xffffffffu, u4.x, o); u4.y += __shfl_xor_sync(0xffffffffu, u4.y, o);
            u4.z += __shfl_xor_sync(0xffffffffu, u4.z, o); u4.w += __shfl_xor_sync(0xffffffffu, u4.w, o);
          }
          if (rsub == 0) {
            float4 a = *reinterpret_cast<float4*>(&s_sum2[q][cc]), bq = *reinterpret_cast<float4*>(&s_sq2[q][cc]);
            a.x += t4.x; a.y += t4.y; a.z += t4.z; a.w += t4.w;
            bq.x += u4.x; bq.y += u4.y; bq.z += u4.z; bq.w += u4.w;
            *reinterpret_cast<float4*>(&s_sum2[q][cc]) = a; *reinterpret_cast<float4*>(&s_sq2[q][cc]) = bq;
          }
        }
        if (micro) p.trace[38] = clock64();
      }
      if (tracing) p.trace[3 * i + 2] = clock64();
    }
    if ((st1 || st2) && stats_img >= 0) flush_stats();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ A-strip variant (3x3, W >= 128)
// The plain kernel is L2->SMEM bound: it fetches a fresh 128-pixel A tile for each of the 9 taps.  Here one (bw+2)-pixel strip
// per (dy, K-chunk) serves the three dx taps: tap dx is the same shared-memory strip read through a UMMA descriptor whose start
// address is advanced by dx rows (128 B each).  Measured on B200: the 128-byte swizzle is a function of the ABSOLUTE shared-memory
// address bits, so the descriptor's matrix-base-offset field must stay 0 even though the start is not 1024-byte aligned (setting
// it to (start >> 7) & 7 gives wrong results; tests/test_gpu_conv_mma.py).  A bytes per tile drop 2.76x; weights keep their own,
// deeper ring.
// T = 2 ("pair"): one CTA also owns the SAME tile of the next image, so every weight tile feeds two M tiles (four TMEM accumulators
// = 4 * BN columns).  Per (dy, K-chunk) that is 2 strips + 3 weight tiles for 72 MMAs: ~36 B/clk (BN=128) instead of 85 B/clk for
// the plain kernel -- below what the latency-bound TMA ring can deliver (~50 B/clk with ~130 KB in flight).
template <int BN, int T>
struct StripCfg {
  static constexpr int STRIP_ROWS = MM_M + 2;                         // 130 pixels
  static constexpr int STRIP_BYTES = STRIP_ROWS * 128;                // per plane, what TMA writes
  static constexpr int STRIP_SLOT = 2 * 17408;                        // two planes, each padded to a 1024-byte multiple (17 KB)
  static constexpr int A_SLOT = T * STRIP_SLOT;
  static constexpr int B_PLANE = BN * MM_KC * 2;
  static constexpr int B_SLOT = 2 * B_PLANE;
  static constexpr int NA = T == 2 ? 2 : (BN == 128 ? 2 : 3);
  static constexpr int NB = T == 2 ? (BN == 128 ? 2 : 4) : (BN == 128 ? 4 : 6);
  static constexpr int SMEM_BYTES = NA * A_SLOT + NB * B_SLOT + 1024;
  static constexpr int TMEM_COLS = 2 * BN * T;
  static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(MM_M >> 4) << 24);
};

template <int BN, int T>
__global__ void __launch_bounds__(MM_THREADS, 1)
conv_mma_strip_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                      const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, const ConvMmaParams p) {
  using Cfg = StripCfg<BN, T>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[Cfg::NA], a_empty[Cfg::NA], b_full[Cfg::NB], b_empty[Cfg::NB], bar_accum;
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_sum[T][BN], s_sq[T][BN], s_sum2[T][BN], s_sq2[T][BN];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base + Cfg::NA * Cfg::A_SLOT;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int img0 = blockIdx.y * T, n0 = blockIdx.z * BN;
  const int y0 = blockIdx.x / p.tiles_x, x0 = (blockIdx.x % p.tiles_x) * MM_M;     // bh == 1: one image row per tile
  const int n_strips = 3 * p.kchunks;

  for (int i = threadIdx.x; i < T * BN; i += MM_THREADS) { (&s_sum[0][0])[i] = 0.f; (&s_sq[0][0])[i] = 0.f; (&s_sum2[0][0])[i] = 0.f; (&s_sq2[0][0])[i] = 0.f; }
  if (warp == 5 && lane == 0) {
    for (int s = 0; s < Cfg::NA; ++s) { mbar_init(smem_u32(&a_full[s]), 1); mbar_init(smem_u32(&a_empty[s]), 1); }
    for (int s = 0; s < Cfg::NB; ++s) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), 1); }
    mbar_init(smem_u32(&bar_accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_lo) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 4) {
    if (lane == 0) {     // ---------------- TMA producer: the T strips of (kc, dy), then the three weight tiles they share
      int ib = 0;
      for (int is = 0; is < n_strips; ++is) {
        const int kc = is / 3, dy = is % 3;
        const int sa = is % Cfg::NA;
        mbar_wait(smem_u32(&a_empty[sa]), ((uint32_t)(is / Cfg::NA) & 1u) ^ 1u);
        const uint32_t af = smem_u32(&a_full[sa]);
        mbar_expect_tx(af, T * 2 * Cfg::STRIP_BYTES);
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const uint32_t adst = smem_base + sa * Cfg::A_SLOT + t * Cfg::STRIP_SLOT;
          const int row = (img0 + t) * (p.H + 2) + y0 + dy;
          tma_load_3d(adst, &tm_a_hi, af, kc * MM_KC, x0, row);
          tma_load_3d(adst + Cfg::STRIP_SLOT / 2, &tm_a_lo, af, kc * MM_KC, x0, row);
        }
        for (int dx = 0; dx < 3; ++dx, ++ib) {
          const int sb = ib % Cfg::NB;
          mbar_wait(smem_u32(&b_empty[sb]), ((uint32_t)(ib / Cfg::NB) & 1u) ^ 1u);
          const uint32_t bf = smem_u32(&b_full[sb]);
          mbar_expect_tx(bf, Cfg::B_SLOT);
          const uint32_t bdst = b_base + sb * Cfg::B_SLOT;
          const int tap = dy * 3 + dx;
          tma_load_2d(bdst, &tm_b_hi, bf, kc * MM_KC, tap * p.cout_total + n0);
          tma_load_2d(bdst + Cfg::B_PLANE, &tm_b_lo, bf, kc * MM_KC, tap * p.cout_total + n0);
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer (whole warp; one elected lane issues, see tc_elect_one)
    {
      int ib = 0;
      for (int is = 0; is < n_strips; ++is) {
        const int sa = is % Cfg::NA;
        mbar_wait(smem_u32(&a_full[sa]), (uint32_t)(is / Cfg::NA) & 1u);
        for (int dx = 0; dx < 3; ++dx, ++ib) {
          const int sb = ib % Cfg::NB;
          mbar_wait(smem_u32(&b_full[sb]), (uint32_t)(ib / Cfg::NB) & 1u);
          tc_fence_after();
          const uint32_t bbase = b_base + sb * Cfg::B_SLOT;
          const uint64_t b_hi = make_kmajor_sw128_desc(bbase), b_lo = make_kmajor_sw128_desc(bbase + Cfg::B_PLANE);
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const uint32_t abase = smem_base + sa * Cfg::A_SLOT + t * Cfg::STRIP_SLOT;
            const uint64_t a_hi = make_kmajor_sw128_desc(abase + dx * 128);
            const uint64_t a_lo = make_kmajor_sw128_desc(abase + Cfg::STRIP_SLOT / 2 + dx * 128);
            const uint32_t acc0 = tmem_base + t * 2 * BN, acc1 = acc0 + BN;
#pragma unroll
            for (int k = 0; k < MM_KC / 16; ++k) {
              const uint64_t adv = (uint64_t)(k * 32 >> 4);
              const uint32_t accum = (ib | k) != 0;
              tc_mma_f16(acc0, a_hi + adv, b_hi + adv, Cfg::IDESC, accum);
              tc_mma_f16(acc1, a_hi + adv, b_lo + adv, Cfg::IDESC, accum);
              tc_mma_f16(acc1, a_lo + adv, b_hi + adv, Cfg::IDESC, 1u);
            }
          }
          tc_commit(smem_u32(&b_empty[sb]));
        }
        tc_commit(smem_u32(&a_empty[sa]));
      }
      tc_commit(smem_u32(&bar_accum));
    }
  } else {
    // each warp stages and drains only its own 32 rows, so the two tiles can reuse the staging buffer back to back
#pragma unroll
    for (int t = 0; t < T; ++t)
      conv_epilogue<BN>(p, smem_u32(&bar_accum), tmem_base + t * 2 * BN, reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw))),
                        s_sum[t], s_sq[t], s_sum2[t], s_sq2[t], img0 + t, n0, y0, x0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

static int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return -2; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return -2; }
  return 0;
}

template <int BN, int T>
static int launch_conv_strip(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                             const ConvMmaParams& p, dim3 grid, cudaStream_t stream) {
  using Cfg = StripCfg<BN, T>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024 && Cfg::TMEM_COLS <= 512, "strip configuration exceeds the SM");
  cudaError_t e = cudaFuncSetAttribute(conv_mma_strip_kernel<BN, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  if (e != cudaSuccess) return cuda_fail(e, "conv_mma_strip smem attr");
  grid.y /= T;
  conv_mma_strip_kernel<BN, T><<<grid, MM_THREADS, Cfg::SMEM_BYTES, stream>>>(a_hi, a_lo, b_hi, b_lo, p);
  VT_CHECK_LAUNCH("vt_conv_mma(strip)");
  return 0;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
  }
  return n;
}

template <int BN, bool RESB, bool STRIP = false, bool E8 = false>
static int launch_conv_persist(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                               const ConvMmaParams& p, dim3 grid, cudaStream_t stream) {
  using Cfg = PersistCfg<BN, RESB, STRIP, E8>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_mma_persist_kernel<BN, RESB, STRIP, E8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "conv_mma_persist smem attr");
    attr_set = true;
  }
  const int tiles_per_img = (int)grid.x, n_tiles_n = (int)grid.z;
  const int total = tiles_per_img * (int)grid.y * n_tiles_n;
  // VT_CONV_MAX_CTAS caps the persistent grid (experiments with a second stream: leave SMs to concurrent HBM-bound kernels)
  static int max_ctas = -1;
  if (max_ctas < 0) { const char* e = getenv("VT_CONV_MAX_CTAS"); max_ctas = e && atoi(e) > 0 ? atoi(e) : num_sms(); }
  const int cap = max_ctas < num_sms() ? max_ctas : num_sms();
  const int ctas = total < cap ? total : cap;
  ConvMmaParams p2 = p;
  const bool trace = getenv("VT_CONV_TRACE") != nullptr;         // debug only: synchronises and prints the epilogue stamps of CTA 0
  if (trace) { cudaMalloc(&p2.trace, 48 * sizeof(long long)); cudaMemset(p2.trace, 0, 48 * sizeof(long long)); }
  conv_mma_persist_kernel<BN, RESB, STRIP, E8><<<ctas, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(a_hi, a_lo, b_hi, b_lo, p2, tiles_per_img, n_tiles_n, total);
  VT_CHECK_LAUNCH("vt_conv_mma(persistent)");
  if (trace) {
    long long h[48];
    cudaDeviceSynchronize();
    cudaMemcpy(h, p2.trace, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(p2.trace);
    fprintf(stderr, "[conv_mma_persist<%d,%d,%d,%d> trace: per tile wait / drain cycles]", BN, (int)RESB, (int)STRIP, (int)E8);
    for (int i = 0; i < 10 && h[3 * i + 2]; ++i) fprintf(stderr, " %lld/%lld", h[3 * i + 1] - h[3 * i], h[3 * i + 2] - h[3 * i + 1]);
    fprintf(stderr, " | chunk 0 of tile 2: tmem-ld %lld, combine+sts %lld, lds+bias+res %lld, prefetch-issue %lld, stg+stats-acc %lld, stats-reduce %lld",
            h[33] - h[32], h[34] - h[33], h[35] - h[34], h[36] - h[35], h[37] - h[36], h[38] - h[37]);
    fprintf(stderr, "\n");
  }
  return 0;
}

template <int BN>
static int launch_conv_mma(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                           const ConvMmaParams& p, dim3 grid, cudaStream_t stream) {
  using Cfg = MmaCfg<BN>;
  cudaError_t e = cudaFuncSetAttribute(conv_mma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  if (e != cudaSuccess) return cuda_fail(e, "conv_mma smem attr");
  conv_mma_kernel<BN><<<grid, MM_THREADS, Cfg::SMEM_BYTES, stream>>>(a_hi, a_lo, b_hi, b_lo, p);
  VT_CHECK_LAUNCH("vt_conv_mma");
  return 0;
}

}  // namespace vt

using namespace vt;

extern "C" {

int vt_conv_mma(const void* a_hi, const void* a_lo, int n_img, int H, int W, int Cin_pad, int pad, int ks, const void* w_hi,
                const void* w_lo, int Cout, const float* bias, const float* res, int ldr, float* out, int ldo, double* stats,
                int ld_stats, void* stream) {
  return vt_conv_mma_dual(a_hi, a_lo, n_img, H, W, Cin_pad, pad, ks, w_hi, w_lo, Cout, bias, res, ldr, out, ldo, stats, ld_stats,
                          nullptr, 0, nullptr, 0, nullptr, 0, stream);
}

int vt_conv_mma_dual(const void* a_hi, const void* a_lo, int n_img, int H, int W, int Cin_pad, int pad, int ks, const void* w_hi,
                     const void* w_lo, int Cout, const float* bias, const float* res, int ldr, float* out, int ldo, double* stats,
                     int ld_stats, float* out2, int ldo2, const float* res2, int ldr2, double* stats2, int ld_stats2, void* stream) {
  VT_CHECK_ARG(out2 == nullptr || (res2 != nullptr && ldo2 % 4 == 0 && ldr2 % 4 == 0), "vt_conv_mma_dual: out2 needs res2 and 4-aligned strides");
  VT_CHECK_ARG(ks == 1 || ks == 3, "vt_conv_mma: kernel size %d", ks);
  VT_CHECK_ARG(pad == ks / 2, "vt_conv_mma: operand planes must carry a border of %d (got %d)", ks / 2, pad);
  VT_CHECK_ARG(Cin_pad % MM_KC == 0, "vt_conv_mma: Cin_pad=%d is not a multiple of %d", Cin_pad, MM_KC);
  VT_CHECK_ARG(Cout == 32 || Cout == 64 || Cout % 128 == 0, "vt_conv_mma: Cout=%d (32, 64 or a multiple of 128)", Cout);
  int bw = W >= 128 ? 128 : W;
  VT_CHECK_ARG(W % bw == 0 && 128 % bw == 0 && bw >= 8, "vt_conv_mma: W=%d cannot be tiled into 128-pixel strips", W);
  int bh = 128 / bw;
  VT_CHECK_ARG(H % bh == 0, "vt_conv_mma: H=%d is not a multiple of the tile height %d", H, bh);
  VT_CHECK_ARG(ldo % 4 == 0 && (res == nullptr || ldr % 4 == 0), "vt_conv_mma: channel strides must be multiples of 4");
  const int BN = Cout >= 128 ? 128 : Cout;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;

  // A-strip variants (3x3 on maps at least 128 wide) are opt-in: VT_CONV_STRIP=1 (one strip serves the three dx taps) or =2 (strip +
  // two images per CTA sharing the weight tiles).  They cut TMA bytes per MMA 1.5x / 2.4x and are bit-identical, but measured no
  // faster than the plain kernel on B200 (profiles/r01d_conv_strip_vs_plain.txt): at M=128 x N=128 the three MMAs of a K-step already
  // read 24 KB of operands from shared memory per 192 tensor cycles (= the 128 B/clk shared-memory port) and the kernel runs at
  // ~1.2 PFLOP/s executed = 0.72 of the measured cuBLAS bf16 burst peak / 0.85 of the sustained one; fewer global->shared bytes do
  // not relieve that.  Kept (and tested) as the basis for round 2's A-from-TMEM / 2-CTA work.
  const char* strip_e = getenv("VT_CONV_STRIP");
  const int strip_mode = strip_e ? atoi(strip_e) : 0;
  const char* persist_e = getenv("VT_CONV_PERSIST");
  const int persist_mode = persist_e ? atoi(persist_e) : 1;         // 1 = persistent (default), 0 = one tile per CTA, 2 / 3 = persistent
                                                                    // without the resident 1x1 panel / without the 3x3 strips
  // persistent + strip: the default for 3x3 convolutions with Cout >= 128 on maps at least 128 wide (+3 % there; for BN <= 64 it is
  // 1-3 % slower than the plain ring, profiles/r01l_*; VT_CONV_PERSIST=4 forces it for every BN)
  const bool pstrip = persist_mode != 0 && persist_mode != 3 && strip_mode == 0 && ks == 3 && bh == 1 && (BN == 128 || persist_mode == 4);
  const bool strip = pstrip || (ks == 3 && bh == 1 && strip_mode != 0 && BN >= 64 && !(strip_mode == 2 && (n_img % 2)));
  const bool pair = !pstrip && strip && strip_mode == 2;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  {
    cuuint64_t dims[3] = {(cuuint64_t)Cin_pad, (cuuint64_t)Wp, (cuuint64_t)n_img * Hp};
    cuuint64_t strides[2] = {(cuuint64_t)Cin_pad * 2, (cuuint64_t)Wp * Cin_pad * 2};
    cuuint32_t box[3] = {(cuuint32_t)MM_KC, (cuuint32_t)(strip ? bw + 2 : bw), (cuuint32_t)bh};
    int rc = make_map(&ma_hi, a_hi, 3, dims, strides, box); if (rc) return rc;
    rc = make_map(&ma_lo, a_lo, 3, dims, strides, box); if (rc) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cin_pad, (cuuint64_t)ks * ks * Cout};
    cuuint64_t strides[1] = {(cuuint64_t)Cin_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)MM_KC, (cuuint32_t)BN};
    int rc = make_map(&mb_hi, w_hi, 2, dims, strides, box); if (rc) return rc;
    rc = make_map(&mb_lo, w_lo, 2, dims, strides, box); if (rc) return rc;
  }
  ConvMmaParams p;
  p.H = H; p.W = W; p.pad = pad; p.ks = ks; p.kchunks = Cin_pad / MM_KC;
  p.bw = bw; p.bh = bh; p.tiles_x = W / bw; p.cout_total = Cout;
  p.bw_shift = 0; while ((1 << p.bw_shift) < bw) ++p.bw_shift;
  p.trace = nullptr;
  p.dbg_nostore = getenv("VT_CONV_DEBUG_NOSTORE") != nullptr;
  {
    const char* pf = getenv("VT_CONV_PREFETCH");       // tiles of L2 prefetch distance for the HBM-streaming shapes; default off: measured 5-25 % SLOWER on B200 (profiles/r01k_*)
    p.prefetch = (ks == 1 || p.kchunks == 1) ? (pf ? atoi(pf) : 0) : 0;
  }
  p.bias = bias; p.res = res; p.ldr = ldr; p.out = out; p.ldo = ldo; p.stats = stats; p.ld_stats = ld_stats;
  p.out2 = out2; p.ldo2 = ldo2; p.res2 = res2; p.ldr2 = ldr2; p.stats2 = stats2; p.ld_stats2 = ld_stats2;
  dim3 grid((H / bh) * (W / bw), n_img, Cout / BN);
  cudaStream_t s = (cudaStream_t)stream;
  if (pair) {
    if (BN == 128) return launch_conv_strip<128, 2>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
    return launch_conv_strip<64, 2>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
  }
  if (strip && !pstrip) {
    if (BN == 128) return launch_conv_strip<128, 1>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
    return launch_conv_strip<64, 1>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
  }
  // default: persistent kernel (overlapped epilogue, merged N = 2*BN MMA); VT_CONV_PERSIST=0 selects the one-tile-per-CTA kernel
  if (pstrip) {
    if (BN == 128) return launch_conv_persist<128, false, true>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
    if (BN == 64) return launch_conv_persist<64, false, true>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
    return launch_conv_persist<32, false, true>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
  }
  if (persist_mode != 0) {
    // 1x1 convolutions with Cout >= 64, two epilogue warp groups: opt-in (VT_CONV_PERSIST=6) -- measured no faster without a residual and
    // 25-40 % slower with one (profiles/r01n_*): the 1x1 tiles are not bound by epilogue issue slots but, like the 3x3 ones, by the
    // shared-memory port (per 128x128x256 tile: 320 KB of MMA operand reads + 128-256 KB of TMA fills + 128 KB of epilogue staging)
    if (ks == 1 && BN >= 64 && persist_mode == 6) {
      if (BN == 128) return launch_conv_persist<128, false, false, true>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
      return launch_conv_persist<64, false, false, true>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
    }
    // 1x1 convolutions with Cin <= 256 can keep the weight panel of a Cout tile resident in shared memory
    if (ks == 1 && p.kchunks <= 4 && persist_mode != 2) {
      if (BN == 128) return launch_conv_persist<128, true>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
      if (BN == 64) return launch_conv_persist<64, true>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
      return launch_conv_persist<32, true>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
    }
    if (BN == 128) return launch_conv_persist<128, false>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
    if (BN == 64) return launch_conv_persist<64, false>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
    return launch_conv_persist<32, false>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
  }
  if (BN == 128) return launch_conv_mma<128>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
  if (BN == 64) return launch_conv_mma<64>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
  return launch_conv_mma<32>(ma_hi, ma_lo, mb_hi, mb_lo, p, grid, s);
}

}  // extern "C"
