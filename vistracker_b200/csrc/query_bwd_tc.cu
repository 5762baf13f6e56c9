// SIF-Net point query, gradient w.r.t. the points, with the decoder MLPs on the tcgen05 tensor cores -- the operator behind
// `df.sum().backward()` in Generator.approx_surface (recon/gen/generator.py:86-98) and behind the df / part / centre losses of the
// fitters (recon/recon_fit_behave.py:467-513, recon/recon_fit_trivis_full.py:193-270).  Same contract as query_bwd_kernel in
// query.cu (which stays as the CUDA-core cross-check); the reference builds an autograd graph over 8 grid_sample + ~10 cat + 20
// Conv1d launches for the same result.
//
// One CTA = 128 query points (M of every MMA), one decoder head at a time:
//   forward   F1  acc  = feat[128 x 640] * W1^T      (10 gathered feature chunks, fp16 hi/lo split, 3 MMAs per K-step)
//             F2/F3    = relu(.)*W2^T, relu(.)*W3^T  (ReLU masks stay in the epilogue threads' registers: 3 x 128 bits per point)
//   cotangent g4       = g_out (x sigmoid' for visibility) or d clamp(df, thr) for a projection step; g3 = mask3 . (W4^T g4) on CUDA cores
//   backward  B3/B2    = g3 * W3, g2 * W2            (transposed weight planes; each vector is renormalised per point by a power of
//                                                     two before the fp16 split, so tiny loss scalings cannot underflow; the
//                                                     exponent travels with the point and is re-applied at the end -- the map is linear)
//             B1  gf   = g1[128 x 128] * W1          (640 feature-gradient columns, five N=128 groups through a 3-slot TMEM ring)
//   gather-dot         the epilogue drains gf into a shared fp32 staging ring (the idle feature ring), the gather warps re-read the
//                      four bilinear taps of every feature and contract gf with d(feature)/d(u, v), then with the projection Jacobians.
// Warp roles (448 threads) as in query_tc.cu: warps 0-3 epilogue (thread = point = TMEM lane), warp 4 TMA weight producer, warp 5
// MMA issuer, warps 6-13 gather (half a warp per point, 16-byte tap loads, four point pairs in flight).
#include "query_tc_common.cuh"
#include <stdlib.h>
#include <type_traits>
#include "vt_internal.h"

namespace vt {

constexpr int TB_GATHER_WARPS = 16;                          // gather warps (8 points each); the forward kernel keeps TQ_GATHER_WARPS = 8
constexpr int TB_THREADS = (6 + TB_GATHER_WARPS) * 32;
constexpr int TB_NGF = 2;                                  // TMEM slots of 128 columns for gf, after the two 128-column head accumulators
constexpr int TB_SMEM = 2 * TQ_SLOT /*features | gf staging*/ + TQ_NW * TQ_SLOT /*weights*/ + 2 * TQ_SLOT /*activations*/ + 1024;

struct TbParams {
  const float* g_out;      // [B][29][N] cotangent (mode 0)
  float* g_points;         // [B][N][3] (optional in mode 1; mode 2: d clamp(df) / d point)
  float* points_out;       // [B][N][3] (mode 1)
  int mode, df_idx, head_mask;
  float threshold;         // clamp maximum of the distance term (modes 1 and 2)
  // mode 2 (fused fitting losses, recon_fit_behave.py:471-476 / recon_fit_trivis_full.py:226-235): per-point loss values and the
  // gradient of each term w.r.t. its point with a unit cotangent; the caller applies the reduction weights (the map is linear)
  const long long* labels; // [B][N] part labels or null (no cross-entropy term)
  float* vals_df;          // [B][N] clamp(df[df_idx], max=threshold)
  float* vals_ce;          // [B][N] cross-entropy of the 14 part logits against labels
  float* g_points2;        // [B][N][3] d CE / d point
  int fwd_mask;            // mode 2: heads evaluated forward-only (they share the gather; no backward) ...
  float* out_fwd;          // ... into the packed [B][29][N] prediction buffer
  long long* trace;        // debug (VT_QUERY_TRACE=1): clock64 stamps of CTA (0,0): [0..63] epilogue thread 0, [64..127] gather warp 0
  // mode 2, merged heads (both pointers set, labels given): the caller's loss weights w_df = *w_df_ptr * w_df_mul, w_ce = *w_ce_ptr * w_ce_mul are
  // folded into the cotangents at the head outputs, both heads' g1 stay in tensor memory and accumulate into ONE feature-gradient tile, so the
  // backward gather-dot runs once per tile: g_points = w_df d clamp(df) / d point + w_ce d CE / d point (g_points2 is not written)
  const float* w_df_ptr; const float* w_ce_ptr; float w_df_mul, w_ce_mul;
};

// per head slot of a pair: the last layer's bias, the three hidden biases and the ReLU masks of the three hidden layers (128 bits per point).
// The last layer's weights live next to the tables as tcgen05 B operands (fp16 hi / lo): see load_tables.
struct TbHeadTab {
  float b4[16];
  float bias[3][TQ_H];
  uint32_t mask[3][4][TQ_M];
};
constexpr int TB_W4F_PLANE = 2 * 16 * TQ_KC * 2;          // W4 as B of the forward product: [16 outputs x 128 units], two 64-unit K chunks, 128-byte swizzle
constexpr int TB_W4B_PLANE = TQ_H * 16 * 2;               // W4 as B of the backward product: [128 units x 16 outputs], 32-byte rows, 32-byte swizzle
static_assert(sizeof(TbHeadTab) % 16 == 0 && 2 * sizeof(TbHeadTab) + 2 * TQ_M * 2 * 16 <= TQ_SLOT, "two head tables and the mailboxes share one 32 KB slot");
static_assert(4 * TB_W4F_PLANE + 4 * TB_W4B_PLANE == TQ_SLOT, "the last-layer operands of two head slots fill one 32 KB slot");

// tcgen05.mma with the A operand in tensor memory (fp16 pairs packed K-contiguous: element k of row m at lane m, column k / 2, half k & 1)
__device__ __forceinline__ void tb_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate, uint32_t idesc = TQ_IDESC) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
constexpr uint32_t TB_IDESC_N16 = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(TQ_M >> 4) << 24);     // M = 128, N = 16 (the head outputs)
// K-major operand with 32-byte rows (K = 16 fp16), SWIZZLE_32B: groups of 8 rows are 256 bytes apart
__device__ __forceinline__ uint64_t tb_desc_sw32(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
}
__device__ __forceinline__ void tb_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tb_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tb_ld16f(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  tb_ld16(taddr, r);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tq_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tb_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
// 2^e for |e| <= 126 (the renormalisation exponents are clamped to [-100, 128]; 127 / 128 only for non-finite input, where any factor does)
__device__ __forceinline__ float tb_pow2(int e) { return __int_as_float((min(max(e, -126), 127) + 127) << 23); }

// power-of-two normalisation of a non-negative maximum: returns e with 2^-e * m in [0.5, 1) (0 for m == 0 / non-finite)
__device__ __forceinline__ int tb_norm_exp(float m) {
  if (!(m > 0.f) || !(m < 3.0e38f)) return 0;
  int e;
  frexpf(m, &e);
  return max(e, -100);
}

__device__ __forceinline__ float4 tb_ld4_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// the four bilinear taps (4 channels each) of one point in one chunk.  ALLV: the caller knows that every tap is inside the map (warp-uniform
// fast path): plain loads at 32-bit element offsets; otherwise the predicated, zero-filled form.  Same addresses, same values.
template <bool ALLV>
__device__ __forceinline__ void tb_load_taps(const TqTapTable& T, const TqChunkSrc& s, int pp, float4& t00, float4& t01, float4& t10, float4& t11,
                                             float& tx, float& ty) {
  const uint32_t offv = T.offv[s.combo][pp];
  tx = T.tx[s.combo][pp]; ty = T.ty[s.combo][pp];
  const int rs = s.W * s.C;
  if (ALLV) {
    const int e = ((int)(offv & 0x0FFFFFFFu) - (s.W + 1)) * s.C + s.ch;       // inside one frame's map: < 2^31 elements
    t00 = ld4(s.base + e); t01 = ld4(s.base + (e + s.C)); t10 = ld4(s.base + (e + rs)); t11 = ld4(s.base + (e + rs + s.C));
  } else {
    const unsigned valid = s.sampled ? (offv >> 28) : 0u;
    const float* p = s.base + ((long long)((int)(offv & 0x0FFFFFFFu) - (s.W + 1)) * s.C + s.ch);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    t00 = (valid & 1u) ? ld4(p) : z;
    t01 = (valid & 2u) ? ld4(p + s.C) : z;
    t10 = (valid & 4u) ? ld4(p + rs) : z;
    t11 = (valid & 8u) ? ld4(p + rs + s.C) : z;
  }
}

// 14 warps are allocated registers as 16 (groups of four): 65536 / 512 = 128 per thread
__global__ void __launch_bounds__(TB_THREADS, 1)
query_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_w1_hi, const __grid_constant__ CUtensorMap tm_w1_lo,
                    const __grid_constant__ CUtensorMap tm_w23_hi, const __grid_constant__ CUtensorMap tm_w23_lo,
                    const __grid_constant__ CUtensorMap tm_w23t_hi, const __grid_constant__ CUtensorMap tm_w23t_lo,
                    const __grid_constant__ CUtensorMap tm_w1t_hi, const __grid_constant__ CUtensorMap tm_w1t_lo,
                    const float* __restrict__ points, const float* __restrict__ crop_center, const float* __restrict__ body_center,
                    int B, int N, TqMaps m, TqCam cam, const float* __restrict__ wpack, int wpack_head_stride, TbParams prm,
                    int* __restrict__ overflow) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t feat_full[2], feat_empty[2], w_full[4], w_empty[4], f1_done, acc_full[2], act_full[2], gf_full[TB_NGF],
      gf_empty[TB_NGF], stg_full[2], stg_empty[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ TqTapTable s_tap;
  __shared__ float s_xyz[TQ_M][4];                          // x, y, z - z0, z
  __shared__ unsigned char s_in_img[TQ_M];
  __shared__ int s_scale_e[TQ_M];                           // exponent of the per-point renormalisation of the current head
  __shared__ float s_wloss[2];                              // merged heads: the two loss weights (device scalar x host factor)
  __shared__ unsigned char s_label[TQ_M];                   // mode 2: part label of every point (14 classes)
  __shared__ float s_dfc[TQ_M];                             // clamp(df, max=threshold) (projection step)
  __shared__ float s_gp[TQ_M][2];                           // current head, before the per-point scale: d/d(u, v) of the perspective samples ...
  __shared__ float s_g3[TQ_M][3];                           // ... and d/d(x, y, z) through the three orthographic views and the direct inputs

  const uint32_t smem_base = (tq_smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - tq_smem_u32(smem_raw));
  const uint32_t feat_base = smem_base, w_base = smem_base + 2 * TQ_SLOT;
  uint8_t* feat_ptr = smem_al;                              // also the gf staging ring: slot = fp32 [128 points][64 features]
  // the former activation buffer (the A operands live in tensor memory now), two 32 KB slots: [0] the last layer's weights of the two head slots as
  // tcgen05 B operands -- per slot pj: forward hi | lo planes at pj * 8 KB, backward hi | lo planes at 16 KB + pj * 8 KB --, [1] the two head
  // tables, then the 16-byte mailboxes of the column-half pairs
  uint8_t* w4_ptr = smem_al + 2 * TQ_SLOT + TQ_NW * TQ_SLOT;
  const uint32_t w4_base = smem_base + 2 * TQ_SLOT + TQ_NW * TQ_SLOT;
  TbHeadTab* tab[2] = {reinterpret_cast<TbHeadTab*>(w4_ptr + TQ_SLOT), reinterpret_cast<TbHeadTab*>(w4_ptr + TQ_SLOT) + 1};
  uint8_t* mail_ptr = w4_ptr + TQ_SLOT + 2 * sizeof(TbHeadTab);
  // the warp index through a shuffle: the compiler then knows the role branches are warp-uniform (needed for straight-line UTCHMMA issue below)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int heads = prm.mode == 1 ? 1 : prm.mode == 2 ? ((prm.labels ? 5 : 1) | prm.fwd_mask) : prm.head_mask;
  auto fwd_only = [&](int h) { return prm.mode == 2 && ((prm.fwd_mask >> h) & 1) != 0; };
  const bool merge = prm.mode == 2 && prm.labels != nullptr && prm.w_df_ptr != nullptr && prm.w_ce_ptr != nullptr;   // heads 0 and 2 as ONE pair
  // heads are processed in pairs that share ONE forward gather: both first layers accumulate from the same feature chunks (TMEM
  // columns 0-127 and 128-255), then each head runs its own forward / backward chain and backward gather; gf slots start at column 256
  const int n_heads = __popc((unsigned)heads), n_pairs = (n_heads + 1) >> 1;
  // a head's chain stages may borrow the idle feature ring for their weight tiles when no backward gather (which stages through that
  // ring) runs between the pair's forward gather and this chain: always for the first head of a pair, for the second when the heads are
  // merged or the first was forward-only
  // tensor-memory regions (128 columns each): head slot pj accumulates in region pj, its activations live in region 2 + pj.  The feature
  // gradients gf of a head go through a ring of the two regions that are dead by then: merged pair: both accumulators; otherwise the
  // head's own accumulator and the OTHER slot's activation region
  auto gf_region = [&](int pj, int gs) { return merge ? gs : (gs == 0 ? pj : 2 + (pj ^ 1)); };
  auto chain_wide = [&](int pi, int pj) {
    if (pj == 0 || merge) return true;
    const int hA = 2 * pi < n_heads ? (int)__fns((unsigned)heads, 0, 2 * pi + 1) : -1;
    return hA >= 0 && prm.mode == 2 && ((prm.fwd_mask >> hA) & 1) != 0;
  };
  // merged pair: the chains of the two heads run interleaved, stage by stage (each head has its own accumulator and activation regions)
  const bool interleave = merge;
  auto pair_head = [&](int pi, int j) { return 2 * pi + j < n_heads ? (int)__fns((unsigned)heads, 0, 2 * pi + j + 1) : -1; };

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w1_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w1_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w23_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w23_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w23t_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w23t_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w1t_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w1t_lo) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tq_smem_u32(&s_tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // one tile of 128 points per CTA (a persistent loop over tiles with a static round-robin assignment was measured 8 % SLOWER: tiles differ in
  // cost -- points outside the image skip their taps -- and the hardware CTA scheduler balances them dynamically)
  const int b = blockIdx.y, n0 = blockIdx.x * TQ_M;
  if (warp == 5 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      tq_mbar_init(tq_smem_u32(&feat_full[s]), TB_GATHER_WARPS); tq_mbar_init(tq_smem_u32(&feat_empty[s]), 1);
      tq_mbar_init(tq_smem_u32(&stg_full[s]), 4); tq_mbar_init(tq_smem_u32(&stg_empty[s]), TB_GATHER_WARPS);
    }
    for (int s = 0; s < 4; ++s) { tq_mbar_init(tq_smem_u32(&w_full[s]), 1); tq_mbar_init(tq_smem_u32(&w_empty[s]), 1); }
    tq_mbar_init(tq_smem_u32(&f1_done), 1);
    for (int s = 0; s < TB_NGF; ++s) { tq_mbar_init(tq_smem_u32(&gf_full[s]), 1); tq_mbar_init(tq_smem_u32(&gf_empty[s]), 4); }
    for (int s = 0; s < 2; ++s) { tq_mbar_init(tq_smem_u32(&acc_full[s]), 1); tq_mbar_init(tq_smem_u32(&act_full[s]), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // projections of the tile's points (gather warps, one thread per point) -- same arithmetic as query_fwd_tc_kernel
  if (warp >= 6 && (threadIdx.x - 6 * 32) < TQ_M) {
    const int pp = threadIdx.x - 6 * 32, n = n0 + pp;
    TqProj q; float x = 0.f, y = 0.f, z = 1.f; int in_img = 1;
    if (n < N) {
      const float* pt = points + ((size_t)b * N + n) * 3;
      x = pt[0]; y = pt[1]; z = pt[2];
      float px = __fadd_rn(__fdiv_rn(__fmul_rn(cam.fx, x), z), cam.cx);
      float py = __fadd_rn(__fdiv_rn(__fmul_rn(cam.fy, y), z), cam.cy);
      px = __fadd_rn(__fadd_rn(cam.crop * 0.5f, px), -crop_center[b * 2 + 0]);
      py = __fadd_rn(__fadd_rn(cam.crop * 0.5f, py), -crop_center[b * 2 + 1]);
      q.nx = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, px), cam.crop), -1.f);
      q.ny = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, py), cam.crop), -1.f);
      in_img = (q.nx >= -1.f && q.nx <= 1.f && q.ny >= -1.f && q.ny <= 1.f) ? 1 : 0;
      const float cx = __fadd_rn(x, -body_center[b * 3 + 0]), cy = __fadd_rn(y, -body_center[b * 3 + 1]), cz = __fadd_rn(z, -body_center[b * 3 + 2]);
      q.tu0 = cz; q.tv0 = cy; q.tu1 = -cx; q.tv1 = cy; q.tu2 = cx; q.tv2 = -cz;
    } else {
      q.nx = q.ny = q.tu0 = q.tv0 = q.tu1 = q.tv1 = q.tu2 = q.tv2 = 1e30f;       // every tap out of range -> zero features
    }
    if (prm.mode == 2 && prm.labels != nullptr) s_label[pp] = n < N ? (unsigned char)prm.labels[(size_t)b * N + n] : 0;
    if (merge && pp < 2) s_wloss[pp] = pp == 0 ? __ldg(prm.w_df_ptr) * prm.w_df_mul : __ldg(prm.w_ce_ptr) * prm.w_ce_mul;
    tq_tap_fill(s_tap, pp, q, m); s_xyz[pp][0] = x; s_xyz[pp][1] = y; s_xyz[pp][2] = __fadd_rn(z, -cam.z0); s_xyz[pp][3] = z; s_in_img[pp] = in_img;
  }
  tq_fence_before();
  __syncthreads();
  tq_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  // ================================================================== the decoder chain of a head, ONE STAGE per call (epilogue side):
  // thread = point row = TMEM lane, split by COLUMN HALVES between two warps per lane quadrant -- the epilogue warp q (hf = 0: hidden units
  // 0-63) and gather warp 6 + ((q + 2) & 3), idle while the chain runs and in the same TMEM lane quadrant (hf = 1: units 64-127).
  //
  // Tensor memory (four regions of 128 columns): head slot pj of a pair accumulates in region pj and keeps its ACTIVATIONS in region 2 + pj as
  // the A operand of the next tcgen05.mma (packed fp16 pairs, per 32-element K chunk 16 columns of hi | 16 of lo).  An MMA with A in tensor
  // memory reads only the weight tile through the shared-memory port -- with A in shared memory the 128x128x16 MMAs of these stages ran at
  // 75 cycles each against 32 of tensor-pipe time (4 KB of A + 4 KB of B per MMA at 128 B/clk) -- and, with one A region per head, the chains
  // of the two heads of a merged pair are INTERLEAVED stage by stage: while the tensor pipe works on head A's stage the epilogue warps run
  // head B's, so the MMA + barrier round trip (~40 % of a stage) is hidden instead of waited for twelve times per tile.
  //   stage 0 / 1: layers 1 / 2: bias + ReLU (mask kept), activation -> X          stage 2: layer 3 + head outputs + cotangent, g3 -> X
  //   stage 3 / 4: backward through layers 3 / 2: row renormalisation, mask, -> X (stage 4 leaves g1, the A operand of the B1 product)
  // What the two threads of a row share (partial head outputs, cotangent, row maxima) goes through 64-byte mailboxes, double-buffered by
  // exchange parity (a writer may be one exchange ahead of its reader), ordered by a 64-thread named barrier per quadrant.
  int tr = 0;
  const bool tracing = prm.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
#define TB_STAMP() do { if (tracing && tr < 64) prm.trace[tr++] = clock64(); } while (0)
  struct HeadState { int iacc; int e_total; };
  // stage tables of the two head slots of a pair, filled by the epilogue warps before stage 0 (load_tables)
  auto load_tables = [&](int pi) {
    asm volatile("bar.sync 2, 512;" ::: "memory");        // the previous pair has finished reading the tables
    if (warp < 4) {
      for (int pj = 0; pj < 2; ++pj) {
        const int h = pair_head(pi, pj);
        if (h < 0) break;
        const float* hw = wpack + (size_t)h * wpack_head_stride;
        const float* b1 = hw + 616 * 128;
        const float* b2 = b1 + 128 + 128 * 128;
        const float* b3 = b2 + 128 + 128 * 128;
        const float* W4 = b3 + 128;                       // [128 units][16 outputs (zero padded)], then the 16 output biases
        if (threadIdx.x < 96) {
          const int l = threadIdx.x >> 5, qq = threadIdx.x & 31;
          reinterpret_cast<float4*>(tab[pj]->bias[l])[qq] = __ldg(reinterpret_cast<const float4*>(l == 0 ? b1 : l == 1 ? b2 : b3) + qq);
        } else if (threadIdx.x < 100) {
          reinterpret_cast<float4*>(tab[pj]->b4)[threadIdx.x - 96] = __ldg(reinterpret_cast<const float4*>(W4 + TQ_H * 16) + (threadIdx.x - 96));
        }
        // the last layer as tcgen05 operands: thread = hidden unit u, its 16 weights split into fp16 hi / lo
        //   forward  B[n = output][k = unit]: two [16 x 64] K chunks, rows of 128 bytes, 128-byte swizzle (the layout of every other weight tile)
        //   backward B[n = unit][k = output]: [128 x 16], rows of 32 bytes, 32-byte swizzle (16-byte chunk j of row n at j ^ ((n >> 2) & 1))
        const int u = threadIdx.x;
        float w[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(W4 + u * 16) + i);
          w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
        }
        uint32_t hi[8], lo[8];
        float wmax = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tq_split2(w[2 * i], w[2 * i + 1], hi[i], lo[i], wmax);
        uint8_t* fb = w4_ptr + pj * 2 * TB_W4F_PLANE;                        // hi plane (both K chunks), then the lo plane
        uint8_t* bb = w4_ptr + 4 * TB_W4F_PLANE + pj * 2 * TB_W4B_PLANE;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t o = (uint32_t)((u >> 3) * 256 + (u & 7) * 32 + ((j ^ ((u >> 2) & 1)) << 4));
          *reinterpret_cast<uint4*>(bb + o) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
          *reinterpret_cast<uint4*>(bb + TB_W4B_PLANE + o) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
        }
        const int kc = u >> 6, kk = u & 63;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const uint32_t o = (uint32_t)(kc * (16 * TQ_KC * 2) + (c >> 3) * 1024 + (c & 7) * 128 + ((((kk >> 3) ^ (c & 7)) & 7) << 4) + (kk & 7) * 2);
          const uint32_t hv = (c & 1) ? (hi[c >> 1] >> 16) : (hi[c >> 1] & 0xFFFFu), lv = (c & 1) ? (lo[c >> 1] >> 16) : (lo[c >> 1] & 0xFFFFu);
          *reinterpret_cast<unsigned short*>(fb + o) = (unsigned short)hv;
          *reinterpret_cast<unsigned short*>(fb + TB_W4F_PLANE + o) = (unsigned short)lv;
        }
      }
      tq_fence_async();                                   // generic-proxy writes above -> the tensor core's (async proxy) reads
    }
    asm volatile("bar.sync 2, 512;" ::: "memory");
  };
  int xchg = 0;                                           // exchanges done by this thread (mailbox parity); same sequence in both halves
  auto stage = [&](const int h, const int pj, const int hq, const int st, HeadState& hs, float& amax) -> bool {
    const int q = warp & 3, r = q * 32 + lane, n = n0 + r;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t acc_base = lane_base + pj * TQ_H, x_base = lane_base + (2 + pj) * TQ_H;
    TbHeadTab& T = *tab[pj];
    const int col0 = 32 * hq;                              // this thread's quarter of the 128 hidden units
    auto mail = [&]() { return reinterpret_cast<float*>(mail_ptr + ((xchg & 1) * TQ_M + r) * 32); };       // 4 row maxima + the first head's exponent
    auto quad_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(4 + q) : "memory"); ++xchg; };
    // 8 values (K elements c0 + 8 g .. + 7 of this point) into the A-operand region: 4 columns of hi pairs, 4 of lo pairs
    auto store_x8 = [&](const float (&v)[8], int g, int c0) {
      uint32_t hi4[4], lo4[4];
      tq_split2(v[0], v[1], hi4[0], lo4[0], amax); tq_split2(v[2], v[3], hi4[1], lo4[1], amax);
      tq_split2(v[4], v[5], hi4[2], lo4[2], amax); tq_split2(v[6], v[7], hi4[3], lo4[3], amax);
      const uint32_t at = x_base + c0 + (g >> 2) * 32 + (g & 3) * 4;
      tb_st4(at, hi4);
      tb_st4(at + 16, lo4);
    };
    auto publish = [&]() {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tq_fence_before();                                   // this warp's TMEM reads / writes are done before the MMA it releases
      __syncwarp();
      if (lane == 0) tq_mbar_arrive(tq_smem_u32(&act_full[pj]));
    };
    auto wait_acc = [&]() {
      tq_mbar_wait(tq_smem_u32(&acc_full[pj]), (uint32_t)hs.iacc & 1u); ++hs.iacc;
      tq_fence_after();
    };
    if (st < 3) {
      // ---- forward epilogues E1, E2, E3: bias, ReLU (mask kept), activation -> X.  (E3's activation is the A operand of the last layer:
      //      [128 x 128] x W4^T [128 x 16] on the tensor core; as FFMA loops over W4 in shared memory -- 32 broadcast LDS.128 per 8 hidden units and
      //      point for the part head -- the last layer and its transpose took 39 k of a merged tile's 173 k cycles.)
      wait_acc();
      const float* bias = T.bias[st] + col0;
      uint32_t mword = 0;
#pragma unroll 1
      for (int g2 = 0; g2 < 2; ++g2) {                     // 16 accumulator columns per tensor-memory round trip
        float w16[16];
        tb_ld16f(acc_base + col0 + g2 * 16, w16);
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          const int g = 2 * g2 + gg;
          float v[8];
          const float4 ba = *reinterpret_cast<const float4*>(bias + g * 8), bb = *reinterpret_cast<const float4*>(bias + g * 8 + 4);
          v[0] = w16[8 * gg] + ba.x; v[1] = w16[8 * gg + 1] + ba.y; v[2] = w16[8 * gg + 2] + ba.z; v[3] = w16[8 * gg + 3] + ba.w;
          v[4] = w16[8 * gg + 4] + bb.x; v[5] = w16[8 * gg + 5] + bb.y; v[6] = w16[8 * gg + 6] + bb.z; v[7] = w16[8 * gg + 7] + bb.w;
          uint32_t mk = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) { mk |= (v[i] > 0.f ? 1u : 0u) << i; v[i] = fmaxf(v[i], 0.f); }
          mword |= mk << ((g & 3) * 8);
          store_x8(v, g, col0);
        }
      }
      T.mask[st][hq][r] = mword;
      publish();
      TB_STAMP();
      return true;
    }
    if (st == 3) {
      // ---- head outputs (16 accumulator columns; BOTH column halves read them and derive the same cotangent: nothing to exchange), cotangent at
      //      the outputs, normalised per point -> the K = 16 A operand of g3 = g4 W4 (written by the lower half)
      wait_acc();
      float o[16];
      {
        float t0[8], t1[8];
        tq_ld8(acc_base, t0); tq_ld8(acc_base + 8, t1);
#pragma unroll
        for (int c = 0; c < 8; ++c) { o[c] = t0[c]; o[8 + c] = t1[c]; }
#pragma unroll 1
        for (int ks = 1; ks < TQ_H / 16; ++ks) {             // the partial sums of the other K steps (see stage_w4f)
          tq_ld8(acc_base + 16 * ks, t0); tq_ld8(acc_base + 16 * ks + 8, t1);
#pragma unroll
          for (int c = 0; c < 8; ++c) { o[c] += t0[c]; o[8 + c] += t1[c]; }
        }
      }
      const int nout = h == 0 ? 2 : h == 1 ? 9 : h == 2 ? 14 : h == 3 ? 3 : 1;
      const int hoff = h == 0 ? 0 : h == 1 ? 2 : h == 2 ? 11 : h == 3 ? 25 : 28;
      if (fwd_only(h)) {                       // predictions only: write them, hand the accumulator back
        if (!hq && n < N) {
#pragma unroll
          for (int c = 0; c < 14; ++c) {
            if (c >= nout) break;
            float val = o[c] + T.b4[c];
            if (h == 4) val = 1.f / (1.f + expf(-val));
            if (h == 0 && !s_in_img[r]) val = cam.out_dist;
            prm.out_fwd[((size_t)b * 29 + hoff + c) * N + n] = val;
          }
        }
        publish();
        return false;
      }
      // (one compact path per use: this block runs once per head and tile, and the generic form -- 14 unrolled outputs with the head tests inside --
      //  was 1.4 k instructions per inlined copy: the stage spent most of its 8 k cycles fetching them)
      float g4[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) g4[c] = 0.f;
      float gmax = 0.f;
      if (prm.mode == 0) {                                 // generic cotangent (autograd backward of query())
#pragma unroll
        for (int c = 0; c < 14; ++c) {
          float g = 0.f;
          if (c < nout && n < N) {
            g = prm.g_out[((size_t)b * 29 + hoff + c) * N + n];
            if (h == 4) { const float val = 1.f / (1.f + expf(-(o[c] + T.b4[c]))); g *= val * (1.f - val); }
            if (h == 0 && !s_in_img[r]) g = 0.f;
          }
          g4[c] = g;
          gmax = fmaxf(gmax, fabsf(g));
        }
      } else if (h == 0) {                                 // d clamp(df[df_idx], max = threshold): one output
        if (n < N) {
          float val = (prm.df_idx ? o[1] : o[0]) + T.b4[prm.df_idx];
          const bool inside = s_in_img[r] != 0;
          if (!inside) val = cam.out_dist;
          float g = (val <= prm.threshold && inside) ? 1.f : 0.f;
          if (!hq) {
            s_dfc[r] = fminf(val, prm.threshold);
            if (prm.mode == 2) prm.vals_df[(size_t)b * N + n] = fminf(val, prm.threshold);
          }
          if (merge) g *= s_wloss[0];
          if (prm.df_idx) g4[1] = g; else g4[0] = g;
          gmax = fabsf(g);
        }
      } else if (prm.mode == 2 && h == 2) {                // F.cross_entropy(parts, labels, reduction='none') and its logit gradient
        if (n < N) {
          const int lab = s_label[r];
          float mx = -3.0e38f;
#pragma unroll
          for (int c = 0; c < 14; ++c) { g4[c] = o[c] + T.b4[c]; mx = fmaxf(mx, g4[c]); }
          float sum = 0.f, l_lab = 0.f;
#pragma unroll
          for (int c = 0; c < 14; ++c) { if (c == lab) l_lab = g4[c]; g4[c] = __expf(g4[c] - mx); sum += g4[c]; }
          if (!hq) prm.vals_ce[(size_t)b * N + n] = logf(sum) - (l_lab - mx);
          const float inv_sum = 1.f / sum;
          const float wce = merge ? s_wloss[1] : 1.f;
#pragma unroll
          for (int c = 0; c < 14; ++c) { g4[c] = (g4[c] * inv_sum - (c == lab ? 1.f : 0.f)) * wce; gmax = fmaxf(gmax, fabsf(g4[c])); }
        }
      }
      hs.e_total = tb_norm_exp(gmax);
      if (!hq) {                                           // (warp-uniform: tcgen05.st is warp-collective)
        const float inv = tb_pow2(-hs.e_total);
        float v0[8], v1[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { v0[c] = g4[c] * inv; v1[c] = g4[8 + c] * inv; }
        store_x8(v0, 0, 0);
        store_x8(v1, 1, 0);
      }
      publish();
      TB_STAMP();
      return true;
    }
    if (st == 4) {
      // ---- g3 = relu'(h3) . (g4 W4): mask the tensor core's product, -> X (the A operand of B3)
      wait_acc();
#pragma unroll 1
      for (int g2 = 0; g2 < 2; ++g2) {
        float w16[16];
        tb_ld16f(acc_base + col0 + g2 * 16, w16);
        const uint32_t mk16 = T.mask[2][hq][r] >> (g2 * 16);
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = ((mk16 >> (8 * gg + i)) & 1u) ? w16[8 * gg + i] : 0.f;
          store_x8(v, 2 * g2 + gg, col0);
        }
      }
      publish();
      TB_STAMP();
      return true;
    }
    // ---- backward epilogues EB3 (st 5: mask of layer 2), EB2 (st 6: mask of layer 1): renormalise by the ROW maximum (both halves), mask, split
    const int bl = 6 - st;
    wait_acc();
    float vmax = 0.f;
#pragma unroll 1
    for (int g2 = 0; g2 < 2; ++g2) {
      float w16[16];
      tb_ld16f(acc_base + col0 + g2 * 16, w16);
      const uint32_t mk16 = T.mask[bl][hq][r] >> (g2 * 16);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if ((mk16 >> i) & 1u) vmax = fmaxf(vmax, fabsf(w16[i]));
    }
    // merged pair, second head: the first head's exponent.  The lower half wrote it (same thread, earlier stage) and hands it to the upper
    // half with the row maximum, so that no read of s_scale_e crosses threads without a barrier
    int e_first = hq ? 0 : s_scale_e[r];
    {
      float* mb = mail();
      mb[hq] = vmax;
      if (!hq) mb[4] = __int_as_float(e_first);
      quad_sync();
      const float4 t = *reinterpret_cast<const float4*>(mb);
      vmax = fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w));
      if (hq) e_first = __float_as_int(mb[4]);
    }
    const int e = tb_norm_exp(vmax);
    hs.e_total += e;
    // merged pair: both heads' g1 enter ONE product and need a common per-point exponent E = max(e_A, e_B): the second head scales its own
    // values on the way in and, if it raised E, rescales the first head's g1 (powers of two on fp16 pairs: exact up to underflow)
    int E = hs.e_total;
    const bool last_merged = merge && bl == 0;
    if (last_merged && pj == 1) E = max(E, e_first);
    // the lower half publishes the exponent right after the row-maximum exchange: both halves have read the first head's value (e_first)
    // before that barrier, and with interleaved chains the upper half may reach the second head's read with no further barrier in between
    if (bl == 0 && !hq) s_scale_e[r] = E;                  // also read by the gather warps after the first staging chunk is published
    const float inv = tb_pow2(-e + (hs.e_total - E));
#pragma unroll 1
    for (int g2 = 0; g2 < 2; ++g2) {
      float w16[16];
      tb_ld16f(acc_base + col0 + g2 * 16, w16);
      const uint32_t mk16 = T.mask[bl][hq][r] >> (g2 * 16);
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = ((mk16 >> (8 * gg + i)) & 1u) ? w16[8 * gg + i] * inv : 0.f;
        store_x8(v, 2 * g2 + gg, col0);
      }
    }
    if (last_merged) {
      // (tcgen05.ld / st are warp-collective: the branch must be warp-uniform, rows that need no rescale multiply by one)
      const bool rescale = pj == 1 && e_first < E;
      if (__any_sync(0xffffffffu, rescale)) {
        const __half2 sc2 = __float2half2_rn(rescale ? tb_pow2(max(e_first - E, -30)) : 1.f);
        const uint32_t other = lane_base + 2 * TQ_H + col0;       // the first head's g1, this thread's quarter
#pragma unroll 1
        for (int qq = 0; qq < 2; ++qq) {
          uint32_t w[16];
          tb_ld16(other + qq * 16, w);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const __half2 t = __hmul2(*reinterpret_cast<const __half2*>(&w[i]), sc2);
            w[i] = *reinterpret_cast<const uint32_t*>(&t);
          }
          tb_st16(other + qq * 16, w);
        }
      }
    }
    publish();
    TB_STAMP();
    return true;
  };
  constexpr int TB_NSTAGE = 7;

  if (warp >= 6) {
    // ================================================================== gather warps
    const int gw = warp - 6;
    const int sub = lane >> 4, k = (lane & 15) * 4;
    constexpr int PB = 2;                                   // point PAIRS in flight per warp
    constexpr int PW = TQ_M / TB_GATHER_WARPS;              // points per warp
    constexpr int PBB = 2, NGRP = 1;                        // ... and in the backward contraction: NGRP groups of PBB point pairs;
    static_assert(PBB == 2, "the butterfly reduction below handles exactly two points per half-warp");
    for (int i = lane; i < PW * 3; i += 32) (&s_g3[gw * PW][0])[i] = 0.f;       // each warp owns the rows of its 16 points
    if (lane < PW * 2) (&s_gp[gw * PW][0])[lane] = 0.f;
    float acc_x = 0.f, acc_y = 0.f, acc_z = 0.f;           // lanes 0-15: point gw * PW + lane, summed over heads (modes 0 and 1)
    __syncwarp();
    int it = 0, sc = 0;
    // bit c: all four taps of all 16 points of this warp are inside the maps chunk c samples (warp-uniform; tap table of the prologue)
    uint32_t fastc = 0;
    {
      uint32_t fm = 0;
#pragma unroll
      for (int cb = 0; cb < TQ_NCOMBO; ++cb)
        if (__all_sync(0xffffffffu, (s_tap.offv[cb][gw * PW + (lane & (PW - 1))] >> 28) == 0xFu)) fm |= 1u << cb;
      const uint32_t b = fm;
      fastc = ((b & 1u) ? 0xFu : 0u) | (((b >> 1) & 1u) << 4) | (((b >> 2) & 7u) << 5) | ((((b >> 5) & 3u) == 3u ? 1u : 0u) << 8) | (((b >> 7) & 1u) << 9);
    }
    HeadState hs_h[2] = {{0, 0}, {0, 0}};                  // gather warps 0-3: upper column half of the chain stages
    float amax = 0.f;
    int trg = 64;
    const bool tracing_g = prm.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && gw == 0 && lane == 0;
#define TB_STAMP_G() do { if (tracing_g && trg < 128) prm.trace[trg++] = clock64(); } while (0)
    TB_STAMP_G();
    for (int pi = 0; pi < n_pairs; ++pi) {
      asm volatile("bar.sync 3, %0;" ::"n"(TB_GATHER_WARPS * 32) : "memory");       // every gather warp is done reading the previous head's staging slots
      // ---- forward: 10 feature chunks into the A-operand ring (once per pair of heads)
      for (int c = 0; c < TQ_NCHUNK; ++c, ++it) {
        const int slot = it & 1;
        tq_mbar_wait(tq_smem_u32(&feat_empty[slot]), ((uint32_t)(it >> 1) & 1u) ^ 1u);
        uint8_t* dst = feat_ptr + slot * TQ_SLOT;
        const TqChunkSrc src = tq_chunk_src(c, k, m, b, B);
        // ALLV: every tap of the warp's 16 points lies inside the maps of this chunk (fastc, set up once per tile): unconditional loads --
        // no zero-filled destination registers, no per-tap predicates (the predicated form spends ~10 instructions of set-up per load, and
        // the gather loops are instruction-issue-bound at two gather warps per scheduler)
        auto rounds = [&](auto allv) {
#pragma unroll
          for (int i0 = 0; i0 < PW; i0 += 2 * PB) {
            float4 t00[PB], t01[PB], t10[PB], t11[PB];
            float tx[PB], ty[PB];
#pragma unroll
            for (int j = 0; j < PB; ++j)
              tb_load_taps<decltype(allv)::value>(s_tap, src, gw * PW + i0 + 2 * j + sub, t00[j], t01[j], t10[j], t11[j], tx[j], ty[j]);
#pragma unroll
            for (int j = 0; j < PB; ++j) {
              const int pp = gw * PW + i0 + 2 * j + sub;
              const float w00 = (1.f - tx[j]) * (1.f - ty[j]), w01 = tx[j] * (1.f - ty[j]), w10 = (1.f - tx[j]) * ty[j], w11 = tx[j] * ty[j];
              float v[4];
              v[0] = t00[j].x * w00; v[1] = t00[j].y * w00; v[2] = t00[j].z * w00; v[3] = t00[j].w * w00;
              v[0] += t01[j].x * w01; v[1] += t01[j].y * w01; v[2] += t01[j].z * w01; v[3] += t01[j].w * w01;
              v[0] += t10[j].x * w10; v[1] += t10[j].y * w10; v[2] += t10[j].z * w10; v[3] += t10[j].w * w10;
              v[0] += t11[j].x * w11; v[1] += t11[j].y * w11; v[2] += t11[j].z * w11; v[3] += t11[j].w * w11;
              if (!src.sampled) {
                v[0] = v[1] = v[2] = v[3] = 0.f;
                if (src.direct) { v[0] = s_xyz[pp][0]; v[1] = s_xyz[pp][1]; v[2] = s_xyz[pp][2]; }
              }
              uint2 hh, ll;
              tq_split2(v[0], v[1], hh.x, ll.x, amax);
              tq_split2(v[2], v[3], hh.y, ll.y, amax);
              const uint32_t off = tq_sw_off(pp, k);
              *reinterpret_cast<uint2*>(dst + off) = hh;
              *reinterpret_cast<uint2*>(dst + TQ_PLANE + off) = ll;
            }
          }
        };
        if ((fastc >> c) & 1u) rounds(std::true_type{}); else rounds(std::false_type{});
        tq_fence_async();
        __syncwarp();
        if (lane == 0) tq_mbar_arrive(tq_smem_u32(&feat_full[slot]));
      }
      TB_STAMP_G();
      if (gw < 12) load_tables(pi);
#pragma unroll 1
      for (int pj = 0; pj < 2; ++pj) {
      const int h = pair_head(pi, pj);
      if (h < 0) break;
      // the chain: a merged pair runs both head slots stage by stage (once, at slot 0's turn), otherwise slot pj alone.  ONE call site of the
      // stage code per warp role, head slot and stage as run-time values: inlined at six sites the stages were a third of the kernel's 25 k
      // instructions, each copy executed once per tile -- the chain was instruction-fetch-bound
      if (gw < 12 && (!interleave || pj == 0)) {
        const int nsj = interleave ? 2 * TB_NSTAGE : TB_NSTAGE;
#pragma unroll 1
        for (int sj = 0; sj < nsj; ++sj) {
          const int st = interleave ? (sj >> 1) : sj, pq = interleave ? (sj & 1) : pj;
          HeadState cur = pq ? hs_h[1] : hs_h[0];
          const bool live = stage(pair_head(pi, pq), pq, 1 + (gw >> 2), st, cur, amax);
          if (pq) hs_h[1] = cur; else hs_h[0] = cur;
          if (!live) break;
        }
      }
      if (fwd_only(h)) continue;
      if (merge && pj == 0) continue;                       // merged heads: one staged feature-gradient tile, after the second head's chain
      // ---- backward: contract the staged feature gradients with d(feature)/d(u, v) (second gather of the same taps)
      for (int c = 0; c < TQ_NCHUNK; ++c, ++sc) {
        const int slot = sc & 1;
        tq_mbar_wait(tq_smem_u32(&stg_full[slot]), (uint32_t)(sc >> 1) & 1u);
        TB_STAMP_G();
        const uint8_t* stg = feat_ptr + slot * TQ_SLOT;
        const TqChunkSrc src = tq_chunk_src(c, k, m, b, B);
        const bool full_res = (c < 4) || (c >= 5 && c < 8);                // im_feat / tri_feat maps (Hf x Wf); else tmpx-sized maps
        const float su = 0.5f * (float)((full_res ? m.Wf : m.Wt) - 1), sv = 0.5f * (float)((full_res ? m.Hf : m.Ht) - 1);
        auto rounds = [&](auto allv) {
#pragma unroll 1
        for (int i0 = 0; i0 < PW; i0 += 2 * PBB * NGRP) {
          // d(feature)/d(u, v) of a bilinear sample is linear in the four taps, so the contraction with the staged feature gradient reduces to
          // four dot products per point -- D_ab = sum_k gf_k * tap_ab_k, 16 FMAs per lane -- and the (1 - t, t) blend is applied to the four
          // scalars afterwards (the previous form blended every feature: ~44 operations per lane).  The taps of NGRP groups of two point
          // pairs are requested before the first group is reduced: twice the loads in flight per warp.
          float4 t00[NGRP][PBB], t01[NGRP][PBB], t10[NGRP][PBB], t11[NGRP][PBB];
          float txs[NGRP][PBB], tys[NGRP][PBB];
#pragma unroll
          for (int gq = 0; gq < NGRP; ++gq)
#pragma unroll
            for (int j = 0; j < PBB; ++j)
              tb_load_taps<decltype(allv)::value>(s_tap, src, gw * PW + i0 + gq * 2 * PBB + 2 * j + sub, t00[gq][j], t01[gq][j], t10[gq][j], t11[gq][j],
                                                  txs[gq][j], tys[gq][j]);
#pragma unroll
          for (int gq = 0; gq < NGRP; ++gq) {
          const int ib = i0 + gq * 2 * PBB;
          // per point only (d/du, d/dv) in map pixels are reduced over the lanes -- 4 values per half-warp through a halving butterfly (5
          // shuffles; 4 in the chunks whose 8-lane halves sample different views) -- and summed per projection in shared memory; the
          // projection Jacobians and the per-point scale are applied once per point and head in finish_head below
          float red[4];
#pragma unroll
          for (int j = 0; j < PBB; ++j) {
            const int pp = gw * PW + ib + 2 * j + sub;
            const float4 g = *reinterpret_cast<const float4*>(stg + pp * 256 + (((lane & 15) ^ (pp & 15)) << 4));
            const float tx = txs[gq][j], ty = tys[gq][j];
            const float d00 = fmaf(g.w, t00[gq][j].w, fmaf(g.z, t00[gq][j].z, fmaf(g.y, t00[gq][j].y, g.x * t00[gq][j].x)));
            const float d01 = fmaf(g.w, t01[gq][j].w, fmaf(g.z, t01[gq][j].z, fmaf(g.y, t01[gq][j].y, g.x * t01[gq][j].x)));
            const float d10 = fmaf(g.w, t10[gq][j].w, fmaf(g.z, t10[gq][j].z, fmaf(g.y, t10[gq][j].y, g.x * t10[gq][j].x)));
            const float d11 = fmaf(g.w, t11[gq][j].w, fmaf(g.z, t11[gq][j].z, fmaf(g.y, t11[gq][j].y, g.x * t11[gq][j].x)));
            red[2 * j] = ((d01 - d00) * (1.f - ty) + (d11 - d10) * ty) * su;
            red[2 * j + 1] = ((d10 - d00) * (1.f - tx) + (d11 - d01) * tx) * sv;
            if (src.direct) { s_g3[pp][0] += g.x; s_g3[pp][1] += g.y; s_g3[pp][2] += g.z; }     // one lane of chunk 9: the (x, y, z - z0) inputs
          }
          if (c == 9) __syncwarp();                                        // ... before the top-view sums of the same points below
          {
            const bool u8 = (lane & 8) != 0, u4 = (lane & 4) != 0, u2 = (lane & 2) != 0;
            int idx, cls;
            if (c < 8) {                                                   // one projection for the 16 lanes
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const float send = u8 ? red[i] : red[i + 2], keep = u8 ? red[i + 2] : red[i];
                red[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
              }
              {
                const float send = u4 ? red[0] : red[1], keep = u4 ? red[1] : red[0];
                red[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
              }
              red[0] += __shfl_xor_sync(0xffffffffu, red[0], 2);
              red[0] += __shfl_xor_sync(0xffffffffu, red[0], 1);
              idx = (u8 ? 2 : 0) + (u4 ? 1 : 0);
              cls = (lane & 3) == 0 ? src.view + 1 : -1;
            } else {                                                       // chunks 8 / 9: lanes 0-7 and 8-15 sample different views
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const float send = u4 ? red[i] : red[i + 2], keep = u4 ? red[i + 2] : red[i];
                red[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
              }
              {
                const float send = u2 ? red[0] : red[1], keep = u2 ? red[1] : red[0];
                red[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
              }
              red[0] += __shfl_xor_sync(0xffffffffu, red[0], 1);
              idx = (u4 ? 2 : 0) + (u2 ? 1 : 0);
              cls = ((lane & 1) == 0 && src.sampled) ? src.view + 1 : -1;
            }
            // right view samples (z, y), back (-x, y), top (x, -z); in chunk 8 the two 8-lane halves (right | back) both add to y: in turn
            const int pp = gw * PW + ib + 2 * (idx >> 1) + sub, uv = idx & 1;
#pragma unroll
            for (int turn = 0; turn < 2; ++turn) {
              if (cls >= 0 && (c != 8 || (int)u8 == turn)) {
                if (cls == 0) s_gp[pp][uv] += red[0];
                else if (cls == 1) s_g3[pp][uv ? 1 : 2] += red[0];
                else if (cls == 2) { if (uv) s_g3[pp][1] += red[0]; else s_g3[pp][0] -= red[0]; }
                else { if (uv) s_g3[pp][2] -= red[0]; else s_g3[pp][0] += red[0]; }
              }
              if (c != 8) break;
              __syncwarp();
            }
          }
          }   // groups
        }
        };
        if ((fastc >> c) & 1u) rounds(std::true_type{}); else rounds(std::false_type{});
        __syncwarp();
        TB_STAMP_G();
        if (lane == 0) tq_mbar_arrive(tq_smem_u32(&stg_empty[slot]));
      }
      TB_STAMP_G();
      // ---- finish_head: projection Jacobians and the per-point scale, once per point (one lane per point)
      __syncwarp();
      if (lane < PW) {
        const int pp = gw * PW + lane, n = n0 + pp;
        const float scale = ldexpf(1.f, s_scale_e[pp]);
        const float kk = 2.f / cam.crop, x = s_xyz[pp][0], y = s_xyz[pp][1], iz = 1.f / s_xyz[pp][3];
        const float pu = s_gp[pp][0] * (kk * cam.fx * iz), pv = s_gp[pp][1] * (kk * cam.fy * iz);   // nx = 2 (crop/2 + fx x / z + cx - ccx) / crop - 1
        const float gx = (pu + s_g3[pp][0]) * scale, gy = (pv + s_g3[pp][1]) * scale, gz = (s_g3[pp][2] - (pu * x + pv * y) * iz) * scale;
        if (prm.mode == 2) {      // one gradient tensor per loss term
          if (n < N) {
            float* gp = ((h == 0 || merge) ? prm.g_points : prm.g_points2) + ((size_t)b * N + n) * 3;
            gp[0] = gx; gp[1] = gy; gp[2] = gz;
          }
        } else {
          acc_x += gx; acc_y += gy; acc_z += gz;
        }
        s_gp[pp][0] = 0.f; s_gp[pp][1] = 0.f; s_g3[pp][0] = 0.f; s_g3[pp][1] = 0.f; s_g3[pp][2] = 0.f;
      }
      __syncwarp();
      }   // heads of the pair
    }
    // ---- write the point gradients / the projected points (one lane per point)
    __syncwarp();
    if (prm.mode != 2 && lane < PW) {
      const int pp = gw * PW + lane, n = n0 + pp;
      if (n < N) {
        const float gx = acc_x, gy = acc_y, gz = acc_z;
        if (prm.g_points) { float* gp = prm.g_points + ((size_t)b * N + n) * 3; gp[0] = gx; gp[1] = gy; gp[2] = gz; }
        if (prm.mode == 1) {     // samples - F.normalize(gradient, dim=2) * df_target   (eps 1e-12, generator.py:96)
          const float inv = 1.f / fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-12f), dd = s_dfc[pp];
          float* po = prm.points_out + ((size_t)b * N + n) * 3;
          po[0] = s_xyz[pp][0] - gx * inv * dd; po[1] = s_xyz[pp][1] - gy * inv * dd; po[2] = s_xyz[pp][3] - gz * inv * dd;
        }
      }
    }
    if (amax > 65504.f) atomicAdd(overflow, 1);
  } else if (warp == 4) {
    // ================================================================== TMA producer (weights), same order as the MMA issuer
    if (lane == 0) {
      // Weight tiles go through a ring of two 32 KB slots -- or of FOUR while a head's chain stages run: the feature ring is idle between
      // the last first-layer MMA (f1_done) and the first staging write of the backward gather, and with only two slots the tiles of stage
      // s + 1 could not be requested before the MMAs of stage s had drained, which put an L2 round trip of 64 KB on every one of the
      // (latency-chained) stages.  Producer and MMA issuer walk the same tile sequence and pick slots with the same rule.
      uint32_t par = 0;                                     // bit s: uses of slot s so far, mod 2
      int nsel = 0, wsel = 0;
      bool f1_ok = false;
      int pair_idx = 0;
      auto load = [&](const CUtensorMap* hi, const CUtensorMap* lo, int col, int row, bool wide) {
        const int s = wide ? (wsel++ & 3) : (nsel++ & 1);
        if (s >= 2 && !f1_ok) { tq_mbar_wait(tq_smem_u32(&f1_done), (uint32_t)pair_idx & 1u); f1_ok = true; }
        tq_mbar_wait(tq_smem_u32(&w_empty[s]), ((par >> s) & 1u) ^ 1u);
        par ^= 1u << s;
        const uint32_t full = tq_smem_u32(&w_full[s]), dst = s < 2 ? w_base + s * TQ_SLOT : feat_base + (s - 2) * TQ_SLOT;
        tq_mbar_expect_tx(full, TQ_SLOT);
        tq_tma_2d(dst, hi, full, col, row);
        tq_tma_2d(dst + TQ_PLANE, lo, full, col, row);
      };
      for (int pi = 0; pi < n_pairs; ++pi) {
        const int hA = pair_head(pi, 0), hB = pair_head(pi, 1);
        pair_idx = pi; f1_ok = false;
        for (int c = 0; c < TQ_NCHUNK; ++c) {
          load(&tm_w1_hi, &tm_w1_lo, c * TQ_KC, hA * TQ_H, false);
          if (hB >= 0) load(&tm_w1_hi, &tm_w1_lo, c * TQ_KC, hB * TQ_H, false);
        }
        if (interleave) {                                   // merged pair: stage by stage, head A then head B (the issuer's order)
          for (int st = 0; st < 4; ++st)
            for (int pj = 0; pj < 2; ++pj) {
              const int h = pj == 0 ? hA : hB, layer = st < 2 ? st : 3 - st;
              for (int kc = 0; kc < 2; ++kc) {
                if (st < 2) load(&tm_w23_hi, &tm_w23_lo, kc * TQ_KC, (layer * 5 + h) * TQ_H, true);
                else load(&tm_w23t_hi, &tm_w23t_lo, kc * TQ_KC, (layer * 5 + h) * TQ_H, true);
              }
            }
          for (int u = 0; u < TQ_NCHUNK / 2; ++u)                        // merged B1: per column group the W1^T tiles of both heads
            for (int q = 0; q < 4; ++q)
              load(&tm_w1t_hi, &tm_w1t_lo, (q & 1) * TQ_KC, (q < 2 ? hA : hB) * TQ_NCHUNK * TQ_KC + u * TQ_H, false);
          continue;
        }
        for (int pj = 0; pj < 2; ++pj) {
          const int h = pj == 0 ? hA : hB;
          if (h < 0) break;
          const bool wide = chain_wide(pi, pj);
          for (int layer = 0; layer < 2; ++layer)
            for (int kc = 0; kc < 2; ++kc) load(&tm_w23_hi, &tm_w23_lo, kc * TQ_KC, (layer * 5 + h) * TQ_H, wide);
          if (fwd_only(h)) continue;
          for (int layer = 1; layer >= 0; --layer)
            for (int kc = 0; kc < 2; ++kc) load(&tm_w23t_hi, &tm_w23t_lo, kc * TQ_KC, (layer * 5 + h) * TQ_H, wide);
          for (int u = 0; u < TQ_NCHUNK / 2; ++u)
            for (int kc = 0; kc < 2; ++kc) load(&tm_w1t_hi, &tm_w1t_lo, kc * TQ_KC, h * TQ_NCHUNK * TQ_KC + u * TQ_H, false);
        }
      }
    }
  } else if (warp == 5) {
    // ================================================================== MMA issuer: the whole warp walks the (uniform) control flow and polls the
    // barriers, ONE elected lane issues -- under `if (lane == 0)` the compiler wraps every tcgen05.mma in an ELECT / BRA.U.ANY loop over
    // the active lanes (~10 instructions and ~80 cycles per 32-cycle MMA: the issue thread, not the tensor pipe, set the stage times)
    {
      int it = 0, iact[2] = {0, 0}, gfi = 0, nsel = 0, wsel = 0;
      uint32_t par = 0;
      int trm = 128;
      const bool tracing_m = prm.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
#define TB_STAMP_M() do { if (tracing_m && trm < 256) prm.trace[trm++] = clock64(); } while (0)
      auto next_w = [&](bool wide, uint32_t& addr) {         // wait for the next weight tile of the sequence; returns its slot
        const int s = wide ? (wsel++ & 3) : (nsel++ & 1);
        tq_mbar_wait(tq_smem_u32(&w_full[s]), (par >> s) & 1u);
        par ^= 1u << s;
        tq_fence_after();
        addr = s < 2 ? w_base + s * TQ_SLOT : feat_base + (s - 2) * TQ_SLOT;
        return s;
      };
      auto mma_tile = [&](uint32_t a_addr, uint32_t acc, bool first, bool wide) {
        uint32_t waddr;
        const int s = next_w(wide, waddr);
        const uint64_t a_hi = tq_desc(a_addr), a_lo = tq_desc(a_addr + TQ_PLANE);
        const uint64_t b_hi = tq_desc(waddr), b_lo = tq_desc(waddr + TQ_PLANE);
        if (tq_elect_one()) {
#pragma unroll
          for (int kk = 0; kk < TQ_KC / 16; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);
            tq_mma(acc, a_hi + adv, b_hi + adv, (first && kk == 0) ? 0u : 1u);
            tq_mma(acc, a_hi + adv, b_lo + adv, 1u);
            tq_mma(acc, a_lo + adv, b_hi + adv, 1u);
          }
          tq_commit(tq_smem_u32(&w_empty[s]));
        }
        __syncwarp();
      };
      // K chunk kc (64 elements) of an A operand held in tensor memory (per 32-element chunk: 16 columns of hi pairs | 16 of lo pairs)
      auto mma_tile_ts = [&](uint32_t x_region, int kc, uint32_t acc, bool first, bool wide) {
        uint32_t waddr;
        const int s = next_w(wide, waddr);
        const uint64_t b_hi = tq_desc(waddr), b_lo = tq_desc(waddr + TQ_PLANE);
        if (tq_elect_one()) {
#pragma unroll
          for (int kk = 0; kk < TQ_KC / 16; ++kk) {
            const int kg = kc * TQ_KC + kk * 16;                       // K step of 16 elements = 8 columns
            const uint32_t a_hi = x_region + (kg >> 5) * 32 + ((kg >> 4) & 1) * 8, a_lo = a_hi + 16;
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);
            tb_mma_ts(acc, a_hi, b_hi + adv, (first && kk == 0) ? 0u : 1u);
            tb_mma_ts(acc, a_hi, b_lo + adv, 1u);
            tb_mma_ts(acc, a_lo, b_hi + adv, 1u);
          }
          tq_commit(tq_smem_u32(&w_empty[s]));
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) { if (tq_elect_one()) tq_commit(tq_smem_u32(bar)); __syncwarp(); };
      for (int pi = 0; pi < n_pairs; ++pi) {
        const bool two = pair_head(pi, 1) >= 0;
        for (int c = 0; c < TQ_NCHUNK; ++c, ++it) {                     // F1 of both heads of the pair from the same feature chunk
          const int slot = it & 1;
          tq_mbar_wait(tq_smem_u32(&feat_full[slot]), (uint32_t)(it >> 1) & 1u);
          tq_fence_after();
          mma_tile(feat_base + slot * TQ_SLOT, tmem_base, c == 0, false);
          if (two) mma_tile(feat_base + slot * TQ_SLOT, tmem_base + TQ_H, c == 0, false);
          commit(&feat_empty[slot]);
        }
        commit(&acc_full[0]);
        if (two) commit(&acc_full[1]);
        commit(&f1_done);                               // the feature ring is free: the weight producer may borrow it
        // one chain stage of head slot pj (F2, F3, B3, B2): A = the head's activation region in tensor memory, D = its accumulator
        auto stage_mma = [&](int pj, bool wide) {
          tq_mbar_wait(tq_smem_u32(&act_full[pj]), (uint32_t)iact[pj] & 1u); ++iact[pj];
          tq_fence_after();
          TB_STAMP_M();
          for (int kc = 0; kc < 2; ++kc) { mma_tile_ts(tmem_base + (2 + pj) * TQ_H, kc, tmem_base + pj * TQ_H, kc == 0, wide); TB_STAMP_M(); }
          commit(&acc_full[pj]);
        };
        // the last layer and its transpose: operands resident in shared memory (load_tables), A in the head's activation region
        //   forward   out[128 x 16]  = h3[128 x 128] W4^T   (8 K steps x 3 MMAs, N = 16) -> eight partial sums in accumulator columns 16 ks .. 16 ks + 15
        //   backward  g3[128 x 128]  = g4[128 x 16]  W4     (1 K step  x 3 MMAs)        -> the whole accumulator region
        auto stage_w4f = [&](int pj) {
          tq_mbar_wait(tq_smem_u32(&act_full[pj]), (uint32_t)iact[pj] & 1u); ++iact[pj];
          tq_fence_after();
          const uint32_t fb = w4_base + pj * 2 * TB_W4F_PLANE, x_region = tmem_base + (2 + pj) * TQ_H, acc = tmem_base + pj * TQ_H;
          // every K step accumulates into its OWN 16 columns (8 x 16 = the accumulator region; the epilogue adds the eight partial sums): 24
          // N = 16 MMAs chained on one accumulator ran at ~250 cycles each -- a 16-column update is shorter than the tensor pipe's depth, so
          // each waited for the previous one -- while here consecutive MMAs are independent (dependency distance 8)
          if (tq_elect_one()) {
#pragma unroll
            for (int term = 0; term < 3; ++term)
#pragma unroll
              for (int ks = 0; ks < TQ_H / 16; ++ks) {
                const int kg = ks * 16;
                const uint32_t a_hi = x_region + (kg >> 5) * 32 + ((kg >> 4) & 1) * 8, a_lo = a_hi + 16;
                const uint32_t boff = (uint32_t)((kg >> 6) * (16 * TQ_KC * 2));
                const uint64_t adv = (uint64_t)((kg & 63) >> 3);
                const uint64_t b_hi = tq_desc(fb + boff) + adv, b_lo = tq_desc(fb + TB_W4F_PLANE + boff) + adv;
                tb_mma_ts(acc + 16 * ks, term == 2 ? a_lo : a_hi, term == 1 ? b_lo : b_hi, term == 0 ? 0u : 1u, TB_IDESC_N16);
              }
          }
          __syncwarp();
          commit(&acc_full[pj]);
        };
        auto stage_w4b = [&](int pj) {
          tq_mbar_wait(tq_smem_u32(&act_full[pj]), (uint32_t)iact[pj] & 1u); ++iact[pj];
          tq_fence_after();
          const uint32_t bb = w4_base + 4 * TB_W4F_PLANE + pj * 2 * TB_W4B_PLANE, x_region = tmem_base + (2 + pj) * TQ_H, acc = tmem_base + pj * TQ_H;
          if (tq_elect_one()) {
            const uint64_t b_hi = tb_desc_sw32(bb), b_lo = tb_desc_sw32(bb + TB_W4B_PLANE);
            tb_mma_ts(acc, x_region, b_hi, 0u);
            tb_mma_ts(acc, x_region, b_lo, 1u);
            tb_mma_ts(acc, x_region + 16, b_hi, 1u);
          }
          __syncwarp();
          commit(&acc_full[pj]);
        };
        if (interleave) {
          for (int st = 0; st < 2; ++st)
            for (int pj = 0; pj < 2; ++pj) stage_mma(pj, true);
          for (int pj = 0; pj < 2; ++pj) stage_w4f(pj);
          for (int pj = 0; pj < 2; ++pj) stage_w4b(pj);
          for (int st = 2; st < 4; ++st)
            for (int pj = 0; pj < 2; ++pj) stage_mma(pj, true);
          for (int pj = 0; pj < 2; ++pj) {                                // both heads' g1 are in their activation regions
            tq_mbar_wait(tq_smem_u32(&act_full[pj]), (uint32_t)iact[pj] & 1u); ++iact[pj];
          }
          tq_fence_after();
          for (int u = 0; u < TQ_NCHUNK / 2; ++u, ++gfi) {              // merged B1: gf = g1_A W1_A + g1_B W1_B into the (dead) accumulator regions
            const int gs = gfi % TB_NGF;
            tq_mbar_wait(tq_smem_u32(&gf_empty[gs]), ((uint32_t)(gfi / TB_NGF) & 1u) ^ 1u);
            tq_fence_after();
            for (int q = 0; q < 4; ++q) mma_tile_ts(tmem_base + (2 + (q >> 1)) * TQ_H, q & 1, tmem_base + gf_region(0, gs) * TQ_H, q == 0, false);
            commit(&gf_full[gs]);
          }
          continue;
        }
        for (int pj = 0; pj < (two ? 2 : 1); ++pj) {
          const bool wide = chain_wide(pi, pj);
          const bool fo = fwd_only(pair_head(pi, pj));
          for (int st = 0; st < 2; ++st) stage_mma(pj, wide);
          stage_w4f(pj);
          if (!fo) {
            stage_w4b(pj);
            for (int st = 2; st < 4; ++st) stage_mma(pj, wide);
          }
          tq_mbar_wait(tq_smem_u32(&act_full[pj]), (uint32_t)iact[pj] & 1u); ++iact[pj];   // g1 is in the activation region (forward-only head:
          tq_fence_after();                                                              // its accumulator has been read out and may be reused)
          if (fo) continue;
          for (int u = 0; u < TQ_NCHUNK / 2; ++u, ++gfi) {              // B1: five groups of 128 feature-gradient columns
            const int gs = gfi % TB_NGF;
            tq_mbar_wait(tq_smem_u32(&gf_empty[gs]), ((uint32_t)(gfi / TB_NGF) & 1u) ^ 1u);
            tq_fence_after();
            for (int kc = 0; kc < 2; ++kc) mma_tile_ts(tmem_base + (2 + pj) * TQ_H, kc, tmem_base + gf_region(pj, gs) * TQ_H, kc == 0, false);
            commit(&gf_full[gs]);
          }
        }
      }
    }
  } else {
    // ================================================================== epilogue warps: lower column half of every chain stage, then the drain
    const int r = warp * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    int gfi = 0, sc = 0;
    float amax = 0.f;
    HeadState hs[2] = {{0, 0}, {0, 0}};
    TB_STAMP();
    for (int pi = 0; pi < n_pairs; ++pi) {
    load_tables(pi);
#pragma unroll 1
    for (int pj = 0; pj < 2; ++pj) {
      const int h = pair_head(pi, pj);
      if (h < 0) break;
      bool backward = true;
      if (!interleave || pj == 0) {                        // (one call site: see the gather warps' copy of this loop)
        const int nsj = interleave ? 2 * TB_NSTAGE : TB_NSTAGE;
#pragma unroll 1
        for (int sj = 0; sj < nsj && backward; ++sj) {
          const int st = interleave ? (sj >> 1) : sj, pq = interleave ? (sj & 1) : pj;
          HeadState cur = pq ? hs[1] : hs[0];
          backward = stage(pair_head(pi, pq), pq, 0, st, cur, amax);
          if (pq) hs[1] = cur; else hs[0] = cur;
        }
      }
      if (interleave && pj == 0) continue;                 // the feature gradients of both heads are drained together
      if (!backward) continue;
      // ---- drain gf (five 128-column groups) into the fp32 staging ring for the gather warps
      for (int u = 0; u < TQ_NCHUNK / 2; ++u, ++gfi) {
        const int gs = gfi % TB_NGF;
        tq_mbar_wait(tq_smem_u32(&gf_full[gs]), (uint32_t)(gfi / TB_NGF) & 1u);
        tq_fence_after();
        TB_STAMP();
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          if ((ch & 1) == 0) {
            tq_mbar_wait(tq_smem_u32(&stg_empty[sc & 1]), ((uint32_t)(sc >> 1) & 1u) ^ 1u);
          }
          float v[32];
          tq_ld32(lane_base + gf_region(pj, gs) * TQ_H + ch * 32, v);
          uint8_t* stg = feat_ptr + (sc & 1) * TQ_SLOT + r * 256;
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4)
            *reinterpret_cast<float4*>(stg + ((((ch & 1) * 8 + q4) ^ (r & 15)) << 4)) = make_float4(v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]);
          if ((ch & 1) == 1) {
            __syncwarp();
            if (lane == 0) tq_mbar_arrive(tq_smem_u32(&stg_full[sc & 1]));
            ++sc;
          }
        }
        tq_fence_before();
        __syncwarp();
        if (lane == 0) tq_mbar_arrive(tq_smem_u32(&gf_empty[gs]));
        TB_STAMP();
      }
    }   // head slots
    }   // pairs
    if (amax > 65504.f) atomicAdd(overflow, 1);
  }
  tq_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem_base), "r"(512u) : "memory");
  }
}

static int launch_bwd_tc(const float* points, const float* crop_center, const float* body_center, int B, int N, const float* im_feat,
                         const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht, int Wt, const float* cam7,
                         const float* wpack, const void* const* planes /*w1 hi lo, w23 hi lo, w23t hi lo, w1t hi lo*/, const TbParams& prm,
                         int* overflow, cudaStream_t stream, const char* who) {
  CUtensorMap mp[8];
  int rc;
  const int kcols[4] = {TQ_NCHUNK * TQ_KC, TQ_H, TQ_H, TQ_H};
  const int rows[4] = {5 * TQ_H, 2 * 5 * TQ_H, 2 * 5 * TQ_H, 5 * TQ_NCHUNK * TQ_KC};
  for (int i = 0; i < 8; ++i)
    if ((rc = tq_make_map(&mp[i], planes[i], kcols[i / 2], rows[i / 2]))) return rc;
  TqMaps m{im_feat, tmpx, tri_tmpx, tri_feat, Hf, Wf, Ht, Wt};
  TqCam cam{cam7[0], cam7[1], cam7[2], cam7[3], cam7[4], cam7[5], cam7[6]};
  cudaError_t e = cudaFuncSetAttribute(query_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM);
  if (e != cudaSuccess) return cuda_fail(e, who);
  dim3 grid(ceil_div(N, TQ_M), B);
  const int head_stride = 616 * 128 + 128 + 2 * (128 * 128 + 128) + 128 * 16 + 16;
  TbParams prm2 = prm;
  const bool trace = getenv("VT_QUERY_TRACE") != nullptr;       // debug only: synchronises and prints the phase stamps of CTA (0,0)
  if (trace) { cudaMalloc(&prm2.trace, 256 * sizeof(long long)); cudaMemset(prm2.trace, 0, 256 * sizeof(long long)); }
  query_bwd_tc_kernel<<<grid, TB_THREADS, TB_SMEM, stream>>>(mp[0], mp[1], mp[2], mp[3], mp[4], mp[5], mp[6], mp[7], points, crop_center,
                                                             body_center, B, N, m, cam, wpack, head_stride, prm2, overflow);
  VT_CHECK_LAUNCH(who);
  if (trace) {
    long long h[256];
    cudaDeviceSynchronize();
    cudaMemcpy(h, prm2.trace, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(prm2.trace);
    fprintf(stderr, "[%s trace, cycles since CTA start] epilogue:", who);
    for (int i = 1; i < 64 && h[i]; ++i) fprintf(stderr, " %lld", h[i] - h[0]);
    fprintf(stderr, " | gather:");
    for (int i = 65; i < 128 && h[i]; ++i) fprintf(stderr, " %lld", h[i] - h[0]);
    fprintf(stderr, " | mma:");
    for (int i = 128; i < 256 && h[i]; ++i) fprintf(stderr, " %lld", h[i] - h[0]);
    fprintf(stderr, "\n");
  }
  return 0;
}

}  // namespace vt

using namespace vt;

extern "C" {

int vt_query_bwd_tc(const float* points, const float* crop_center, const float* body_center, int B, int N, const float* im_feat,
                    const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht, int Wt, const float* cam7,
                    const float* wpack, const void* w1_hi, const void* w1_lo, const void* w23_hi, const void* w23_lo, const void* w23t_hi,
                    const void* w23t_lo, const void* w1t_hi, const void* w1t_lo, const float* g_out, int head_mask, float* g_points,
                    int* overflow, void* stream) {
  VT_CHECK_ARG(head_mask >= 0 && head_mask < 32, "vt_query_bwd_tc: head mask %d", head_mask);
  VT_CHECK_ARG(g_out != nullptr && g_points != nullptr && overflow != nullptr, "vt_query_bwd_tc: g_out, g_points and overflow are required");
  if (B <= 0 || N <= 0) return 0;
  if (head_mask == 0) return cudaMemsetAsync(g_points, 0, (size_t)B * N * 3 * sizeof(float), (cudaStream_t)stream) == cudaSuccess ? 0 : -3;
  const void* planes[8] = {w1_hi, w1_lo, w23_hi, w23_lo, w23t_hi, w23t_lo, w1t_hi, w1t_lo};
  TbParams prm{g_out, g_points, nullptr, 0, 0, head_mask, 0.f, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr};
  return launch_bwd_tc(points, crop_center, body_center, B, N, im_feat, tmpx, tri_tmpx, tri_feat, Hf, Wf, Ht, Wt, cam7, wpack, planes, prm,
                       overflow, (cudaStream_t)stream, "vt_query_bwd_tc");
}

int vt_query_project_step_tc(const float* points, const float* crop_center, const float* body_center, int B, int N, const float* im_feat,
                             const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht, int Wt,
                             const float* cam7, const float* wpack, const void* w1_hi, const void* w1_lo, const void* w23_hi,
                             const void* w23_lo, const void* w23t_hi, const void* w23t_lo, const void* w1t_hi, const void* w1t_lo,
                             int df_idx, float threshold, float* points_out, float* g_points, int* overflow, void* stream) {
  VT_CHECK_ARG(df_idx == 0 || df_idx == 1, "vt_query_project_step_tc: df_idx %d (0 human, 1 object)", df_idx);
  VT_CHECK_ARG(points_out != nullptr && overflow != nullptr, "vt_query_project_step_tc: points_out and overflow are required");
  if (B <= 0 || N <= 0) return 0;
  const void* planes[8] = {w1_hi, w1_lo, w23_hi, w23_lo, w23t_hi, w23t_lo, w1t_hi, w1t_lo};
  TbParams prm{nullptr, g_points, points_out, 1, df_idx, 1, threshold, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr};
  return launch_bwd_tc(points, crop_center, body_center, B, N, im_feat, tmpx, tri_tmpx, tri_feat, Hf, Wf, Ht, Wt, cam7, wpack, planes, prm,
                       overflow, (cudaStream_t)stream, "vt_query_project_step_tc");
}

int vt_query_losses_tc(const float* points, const float* crop_center, const float* body_center, int B, int N, const float* im_feat,
                       const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht, int Wt, const float* cam7,
                       const float* wpack, const void* w1_hi, const void* w1_lo, const void* w23_hi, const void* w23_lo, const void* w23t_hi,
                       const void* w23t_lo, const void* w1t_hi, const void* w1t_lo, int df_idx, float clamp_max, const long long* part_labels,
                       float* vals_df, float* g_df, float* vals_ce, float* g_ce, int fwd_mask, float* out_fwd, int* overflow, void* stream) {
  VT_CHECK_ARG(df_idx == 0 || df_idx == 1, "vt_query_losses_tc: df_idx %d (0 human, 1 object)", df_idx);
  VT_CHECK_ARG(fwd_mask >= 0 && fwd_mask < 32 && !(fwd_mask & 1) && !(part_labels && (fwd_mask & 4)) && (fwd_mask == 0 || out_fwd != nullptr),
               "vt_query_losses_tc: forward-only head mask %d (not the loss heads; needs out_fwd)", fwd_mask);
  VT_CHECK_ARG(vals_df != nullptr && g_df != nullptr && overflow != nullptr, "vt_query_losses_tc: vals_df, g_df and overflow are required");
  VT_CHECK_ARG(part_labels == nullptr || (vals_ce != nullptr && g_ce != nullptr), "vt_query_losses_tc: part_labels need vals_ce and g_ce");
  if (B <= 0 || N <= 0) return 0;
  const void* planes[8] = {w1_hi, w1_lo, w23_hi, w23_lo, w23t_hi, w23t_lo, w1t_hi, w1t_lo};
  TbParams prm{nullptr, g_df, nullptr, 2, df_idx, 0, clamp_max, part_labels, vals_df, vals_ce, g_ce, fwd_mask, out_fwd, nullptr};
  return launch_bwd_tc(points, crop_center, body_center, B, N, im_feat, tmpx, tri_tmpx, tri_feat, Hf, Wf, Ht, Wt, cam7, wpack, planes, prm,
                       overflow, (cudaStream_t)stream, "vt_query_losses_tc");
}

int vt_query_losses_merged_tc(const float* points, const float* crop_center, const float* body_center, int B, int N, const float* im_feat,
                              const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht, int Wt, const float* cam7,
                              const float* wpack, const void* w1_hi, const void* w1_lo, const void* w23_hi, const void* w23_lo, const void* w23t_hi,
                              const void* w23t_lo, const void* w1t_hi, const void* w1t_lo, int df_idx, float clamp_max, const long long* part_labels,
                              const float* w_df, float w_df_mul, const float* w_ce, float w_ce_mul, float* vals_df, float* vals_ce, float* g_points,
                              int* overflow, void* stream) {
  VT_CHECK_ARG(df_idx == 0 || df_idx == 1, "vt_query_losses_merged_tc: df_idx %d (0 human, 1 object)", df_idx);
  VT_CHECK_ARG(part_labels != nullptr && w_df != nullptr && w_ce != nullptr, "vt_query_losses_merged_tc: part_labels and both weight pointers are required");
  VT_CHECK_ARG(vals_df != nullptr && vals_ce != nullptr && g_points != nullptr && overflow != nullptr,
               "vt_query_losses_merged_tc: vals_df, vals_ce, g_points and overflow are required");
  if (B <= 0 || N <= 0) return 0;
  const void* planes[8] = {w1_hi, w1_lo, w23_hi, w23_lo, w23t_hi, w23t_lo, w1t_hi, w1t_lo};
  TbParams prm{nullptr, g_points, nullptr, 2, df_idx, 0, clamp_max, part_labels, vals_df, vals_ce, nullptr, 0, nullptr, nullptr, w_df, w_ce, w_df_mul, w_ce_mul};
  return launch_bwd_tc(points, crop_center, body_center, B, N, im_feat, tmpx, tri_tmpx, tri_feat, Hf, Wf, Ht, Wt, cam7, wpack, planes, prm,
                       overflow, (cudaStream_t)stream, "vt_query_losses_merged_tc");
}

}  // extern "C"
