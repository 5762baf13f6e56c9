// SIF-Net point query, forward, with the five decoder MLPs on the tcgen05 tensor cores
// (model/chore_triplane.py:97-164, model/chore_tri_vis.py:31-50 -- same contract as query_fwd_kernel in query.cu, which stays
// as the CUDA-core cross-check and as the forward recompute inside the backward kernel).
//
// One CTA = 128 query points = the M dimension of every MMA.  fp32 parity comes from the same fp16 split as the convolution
// (x = hi + lo, three MMAs per K-step: hi*hi + hi*lo + lo*hi); here lo is NOT rescaled (features and weights are O(1e-2..1e1),
// so lo stays at or just below the fp16 normal range and its absolute error is <= 2^-25), which lets all three MMAs accumulate
// into ONE fp32 TMEM accumulator per head: 128 columns per head, four heads fill the 512 TMEM columns.  The kernel therefore
// makes two passes over the feature chunks: heads {df, pca, parts, centers}, then {visibility}.
//
// Warp roles (448 threads):
//   warps 0-3   epilogue: TMEM -> registers, bias + ReLU, fp16 split, write the next layer's A operand into shared memory in the
//               UMMA K-major 128-byte-swizzled layout; the last layer (128 -> <=14 outputs) runs on the CUDA cores from registers
//   warp  4     TMA producer: weight tiles [128 out x 64 in] (hi and lo planes) through a 2-slot ring
//   warp  5     MMA issuer (one thread)
//   warps 6-13  gather: projections once, then per 64-feature chunk 4 bilinear taps per point from the NHWC maps (half a warp per
//               point, 16-byte loads, four point pairs in flight), split, and store straight into the swizzled A-operand slot
// Feature order for this kernel (10 chunks of 64, every chunk maps to whole channel runs of one or two maps):
//   [ im_feat 4x64 | tmpx 64 | tri_feat right 64 | back 64 | top 64 | tri_tmpx right 32, back 32 | tri_tmpx top 32, x, y, z-2.2, 0... ]
#include "query_tc_common.cuh"
#include "vt_internal.h"

namespace vt {

__global__ void __launch_bounds__(TQ_THREADS, 1)
query_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_w1_hi, const __grid_constant__ CUtensorMap tm_w1_lo,
                    const __grid_constant__ CUtensorMap tm_w23_hi, const __grid_constant__ CUtensorMap tm_w23_lo,
                    const float* __restrict__ points, const float* __restrict__ crop_center, const float* __restrict__ body_center,
                    int B, int N, TqMaps m, TqCam cam, const float* __restrict__ wpack /*fp32 pack of query.cu: biases + W4*/,
                    int wpack_head_stride, float* __restrict__ out, float* __restrict__ xy_out, int* __restrict__ overflow, int head_mask) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t feat_full[TQ_NF], feat_empty[TQ_NF], w_full[TQ_NW], w_empty[TQ_NW], acc_full, act_full;
  __shared__ uint32_t s_tmem_base;
  __shared__ TqTapTable s_tap;
  __shared__ float s_xyz[TQ_M][3];
  __shared__ int s_in_img[TQ_M];
  __shared__ __align__(16) float s_w4[TQ_H * 16 + 16];      // last layer of the current head: W4[128][16] + b4[16]
  __shared__ __align__(16) float s_bias[3][TQ_H];           // b1, b2, b3 of the current head

  const uint32_t smem_base = (tq_smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - tq_smem_u32(smem_raw));
  const uint32_t feat_base = smem_base, w_base = smem_base + TQ_NF * TQ_SLOT, act_base = w_base + TQ_NW * TQ_SLOT;
  uint8_t* feat_ptr = smem_al;
  uint8_t* act_ptr = smem_al + TQ_NF * TQ_SLOT + TQ_NW * TQ_SLOT;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int b = blockIdx.y, n0 = blockIdx.x * TQ_M;
  // active heads (bit h of head_mask), four per pass: the g-th head of pass p is the (4p+g+1)-th set bit
  const int n_heads = __popc((unsigned)head_mask), n_pass = (n_heads + 3) >> 2;
  auto pass_heads = [&](int pass) { return min(4, n_heads - 4 * pass); };
  auto head_of = [&](int pass, int g) { return (int)__fns((unsigned)head_mask, 0, 4 * pass + g + 1); };

  if (warp == 5 && lane == 0) {
    for (int s = 0; s < TQ_NF; ++s) { tq_mbar_init(tq_smem_u32(&feat_full[s]), TQ_GATHER_WARPS); tq_mbar_init(tq_smem_u32(&feat_empty[s]), 1); }
    for (int s = 0; s < TQ_NW; ++s) { tq_mbar_init(tq_smem_u32(&w_full[s]), 1); tq_mbar_init(tq_smem_u32(&w_empty[s]), 1); }
    tq_mbar_init(tq_smem_u32(&acc_full), 1);
    tq_mbar_init(tq_smem_u32(&act_full), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w1_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w1_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w23_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w23_lo) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tq_smem_u32(&s_tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // projections of the tile's points (gather warps, one thread per point)
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
      if (xy_out) { xy_out[((size_t)b * 2 + 0) * N + n] = q.nx; xy_out[((size_t)b * 2 + 1) * N + n] = q.ny; }
    } else {
      q.nx = q.ny = q.tu0 = q.tv0 = q.tu1 = q.tv1 = q.tu2 = q.tv2 = 1e30f;       // every tap out of range -> zero features
    }
    tq_tap_fill(s_tap, pp, q, m); s_xyz[pp][0] = x; s_xyz[pp][1] = y; s_xyz[pp][2] = __fadd_rn(z, -cam.z0); s_in_img[pp] = in_img;
  }
  tq_fence_before();
  __syncthreads();
  tq_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  // head groups of the two passes: {0,1,2,3} then {4}; accumulator of head g in the group sits at column 128 * g
  if (warp >= 6) {
    // ================================================================== gather warps
    const int gw = warp - 6;
    int it = 0;
    float amax = 0.f;
    for (int pass = 0; pass < n_pass; ++pass) {
      for (int c = 0; c < TQ_NCHUNK; ++c, ++it) {
        const int slot = it % TQ_NF;
        tq_mbar_wait(tq_smem_u32(&feat_empty[slot]), ((uint32_t)(it / TQ_NF) & 1u) ^ 1u);
        uint8_t* dst = feat_ptr + slot * TQ_SLOT;
        // half a warp per point: lane -> (point of the pair, 4 consecutive features k..k+3 of the chunk), 16-byte tap loads
        const int sub = lane >> 4, k = (lane & 15) * 4;
        const TqChunkSrc src = tq_chunk_src(c, k, m, b, B);
        const bool sampled = src.sampled;
        constexpr int PB = 4;                             // point PAIRS in flight per warp
        for (int i0 = 0; i0 < TQ_M / TQ_GATHER_WARPS; i0 += 2 * PB) {
          TqTap tap[PB];
          float4 direct[PB];
#pragma unroll
          for (int j = 0; j < PB; ++j) {
            const int pp = gw * (TQ_M / TQ_GATHER_WARPS) + i0 + 2 * j + sub;
            tap[j] = tq_tap_get(s_tap, src, pp);
            direct[j] = src.direct ? make_float4(s_xyz[pp][0], s_xyz[pp][1], s_xyz[pp][2], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          float4 t00[PB], t01[PB], t10[PB], t11[PB];
#pragma unroll
          for (int j = 0; j < PB; ++j) {                  // all loads first
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            t00[j] = (tap[j].valid & 1u) ? ld4(tap[j].p) : z;
            t01[j] = (tap[j].valid & 2u) ? ld4(tap[j].p + tap[j].C) : z;
            t10[j] = (tap[j].valid & 4u) ? ld4(tap[j].p + tap[j].rowstride) : z;
            t11[j] = (tap[j].valid & 8u) ? ld4(tap[j].p + tap[j].rowstride + tap[j].C) : z;
          }
#pragma unroll
          for (int j = 0; j < PB; ++j) {
            const int pp = gw * (TQ_M / TQ_GATHER_WARPS) + i0 + 2 * j + sub;
            float v[4];
            // same accumulation order as the CUDA-core kernel
            v[0] = t00[j].x * tap[j].w00; v[1] = t00[j].y * tap[j].w00; v[2] = t00[j].z * tap[j].w00; v[3] = t00[j].w * tap[j].w00;
            v[0] += t01[j].x * tap[j].w01; v[1] += t01[j].y * tap[j].w01; v[2] += t01[j].z * tap[j].w01; v[3] += t01[j].w * tap[j].w01;
            v[0] += t10[j].x * tap[j].w10; v[1] += t10[j].y * tap[j].w10; v[2] += t10[j].z * tap[j].w10; v[3] += t10[j].w * tap[j].w10;
            v[0] += t11[j].x * tap[j].w11; v[1] += t11[j].y * tap[j].w11; v[2] += t11[j].z * tap[j].w11; v[3] += t11[j].w * tap[j].w11;
            if (!sampled) { v[0] = direct[j].x; v[1] = direct[j].y; v[2] = direct[j].z; v[3] = direct[j].w; }
            uint2 hh, ll;
            tq_split2(v[0], v[1], hh.x, ll.x, amax);
            tq_split2(v[2], v[3], hh.y, ll.y, amax);
            const uint32_t off = tq_sw_off(pp, k);
            *reinterpret_cast<uint2*>(dst + off) = hh;
            *reinterpret_cast<uint2*>(dst + TQ_PLANE + off) = ll;
          }
        }
        tq_fence_async();                                 // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) tq_mbar_arrive(tq_smem_u32(&feat_full[slot]));
      }
    }
    if (amax > 65504.f) atomicAdd(overflow, 1);
  } else if (warp == 4) {
    // ================================================================== TMA producer (weights)
    if (lane == 0) {
      int iw = 0;
      auto load = [&](const CUtensorMap* hi, const CUtensorMap* lo, int col, int row) {
        const int s = iw % TQ_NW;
        tq_mbar_wait(tq_smem_u32(&w_empty[s]), ((uint32_t)(iw / TQ_NW) & 1u) ^ 1u);
        const uint32_t full = tq_smem_u32(&w_full[s]);
        tq_mbar_expect_tx(full, TQ_SLOT);
        tq_tma_2d(w_base + s * TQ_SLOT, hi, full, col, row);
        tq_tma_2d(w_base + s * TQ_SLOT + TQ_PLANE, lo, full, col, row);
        ++iw;
      };
      for (int pass = 0; pass < n_pass; ++pass) {
        const int nh = pass_heads(pass);
        for (int c = 0; c < TQ_NCHUNK; ++c)
          for (int g = 0; g < nh; ++g) load(&tm_w1_hi, &tm_w1_lo, c * TQ_KC, head_of(pass, g) * TQ_H);
        for (int g = 0; g < nh; ++g)
          for (int layer = 0; layer < 2; ++layer)
            for (int kc = 0; kc < 2; ++kc) load(&tm_w23_hi, &tm_w23_lo, kc * TQ_KC, (layer * 5 + head_of(pass, g)) * TQ_H);
      }
    }
  } else if (warp == 5) {
    // ================================================================== MMA issuer (whole warp walks the control flow; one elected lane issues:
    // straight-line UTCHMMA instead of an ELECT / BRA.U.ANY loop per instruction, see tq_elect_one)
    {
      int it = 0, iw = 0, iact = 0;
      auto commit = [&](uint64_t* bar) { if (tq_elect_one()) tq_commit(tq_smem_u32(bar)); __syncwarp(); };
      auto mma_tile = [&](uint32_t a_addr, uint32_t acc, bool first) {
        const int s = iw % TQ_NW;
        tq_mbar_wait(tq_smem_u32(&w_full[s]), (uint32_t)(iw / TQ_NW) & 1u);
        tq_fence_after();
        const uint64_t a_hi = tq_desc(a_addr), a_lo = tq_desc(a_addr + TQ_PLANE);
        const uint64_t b_hi = tq_desc(w_base + s * TQ_SLOT), b_lo = tq_desc(w_base + s * TQ_SLOT + TQ_PLANE);
        if (tq_elect_one()) {
#pragma unroll
          for (int k = 0; k < TQ_KC / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            tq_mma(acc, a_hi + adv, b_hi + adv, (first && k == 0) ? 0u : 1u);
            tq_mma(acc, a_hi + adv, b_lo + adv, 1u);
            tq_mma(acc, a_lo + adv, b_hi + adv, 1u);
          }
          tq_commit(tq_smem_u32(&w_empty[s]));
        }
        __syncwarp();
        ++iw;
      };
      for (int pass = 0; pass < n_pass; ++pass) {
        const int nh = pass_heads(pass);
        for (int c = 0; c < TQ_NCHUNK; ++c, ++it) {
          const int slot = it % TQ_NF;
          tq_mbar_wait(tq_smem_u32(&feat_full[slot]), (uint32_t)(it / TQ_NF) & 1u);
          tq_fence_after();
          for (int g = 0; g < nh; ++g) mma_tile(feat_base + slot * TQ_SLOT, tmem_base + g * TQ_H, c == 0);
          commit(&feat_empty[slot]);
        }
        commit(&acc_full);                                 // layer 1 of the whole group is complete
        for (int g = 0; g < nh; ++g)
          for (int layer = 0; layer < 2; ++layer) {
            tq_mbar_wait(tq_smem_u32(&act_full), (uint32_t)iact & 1u); ++iact;     // epilogue wrote this layer's input
            tq_fence_after();
            for (int kc = 0; kc < 2; ++kc) mma_tile(act_base + kc * TQ_SLOT, tmem_base + g * TQ_H, kc == 0);
            commit(&acc_full);
          }
      }
    }
  } else {
    // ================================================================== epilogue warps: thread = point row = TMEM lane
    const int r = warp * 32 + lane, n = n0 + r;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    int iacc = 0;
    float amax = 0.f;
    for (int pass = 0; pass < n_pass; ++pass) {
      const int nh = pass_heads(pass);
      tq_mbar_wait(tq_smem_u32(&acc_full), (uint32_t)iacc & 1u); ++iacc;            // layer 1 done
      tq_fence_after();
      for (int g = 0; g < nh; ++g) {
        const int h = head_of(pass, g);
        const float* hw = wpack + (size_t)h * wpack_head_stride;
        const float* b1 = hw + 616 * 128;
        const float* b2 = b1 + 128 + 128 * 128;
        const float* b3 = b2 + 128 + 128 * 128;
        const float* W4 = b3 + 128;
        asm volatile("bar.sync 2, 128;" ::: "memory");      // previous head's last layer has finished reading s_w4
        for (int i = threadIdx.x; i < (TQ_H * 16 + 16) / 4; i += 128)
          reinterpret_cast<float4*>(s_w4)[i] = __ldg(reinterpret_cast<const float4*>(W4) + i);     // W4 and b4 are contiguous in the pack
        if (threadIdx.x < 96) {
          const int l = threadIdx.x >> 5, q = threadIdx.x & 31;
          reinterpret_cast<float4*>(s_bias[l])[q] = __ldg(reinterpret_cast<const float4*>(l == 0 ? b1 : l == 1 ? b2 : b3) + q);
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        for (int layer = 0; layer < 3; ++layer) {
          if (layer > 0) { tq_mbar_wait(tq_smem_u32(&acc_full), (uint32_t)iacc & 1u); ++iacc; tq_fence_after(); }
          const float* bias = s_bias[layer];
          float o[14];
#pragma unroll
          for (int c = 0; c < 14; ++c) o[c] = 0.f;
#pragma unroll 1
          for (int ch = 0; ch < 4; ++ch) {
            float v[32];
            tq_ld32(lane_base + g * TQ_H + ch * 32, v);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(bias + ch * 32 + i);
              v[i] = fmaxf(v[i] + bb.x, 0.f); v[i + 1] = fmaxf(v[i + 1] + bb.y, 0.f);
              v[i + 2] = fmaxf(v[i + 2] + bb.z, 0.f); v[i + 3] = fmaxf(v[i + 3] + bb.w, 0.f);
            }
            if (layer < 2) {
              // next layer's A operand: K index = ch*32 + i -> chunk kc = ch / 2, k = (ch & 1) * 32 + i
              uint8_t* dst = act_ptr + (ch >> 1) * TQ_SLOT;
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 hh, ll;
                tq_split2(v[i], v[i + 1], hh.x, ll.x, amax); tq_split2(v[i + 2], v[i + 3], hh.y, ll.y, amax);
                tq_split2(v[i + 4], v[i + 5], hh.z, ll.z, amax); tq_split2(v[i + 6], v[i + 7], hh.w, ll.w, amax);
                const uint32_t off = tq_sw_off(r, (ch & 1) * 32 + i);
                *reinterpret_cast<uint4*>(dst + off) = hh;
                *reinterpret_cast<uint4*>(dst + TQ_PLANE + off) = ll;
              }
            } else if (h == 0 || h == 3 || h == 4) {
              // last layer on the CUDA cores, heads with <= 4 outputs (W4 columns 4..15 are zero padding): one 16-byte load per unit
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float4 w0 = *reinterpret_cast<const float4*>(s_w4 + (ch * 32 + i) * 16);
                o[0] = fmaf(v[i], w0.x, o[0]); o[1] = fmaf(v[i], w0.y, o[1]); o[2] = fmaf(v[i], w0.z, o[2]); o[3] = fmaf(v[i], w0.w, o[3]);
              }
            } else {
              // last layer on the CUDA cores: out[c] += h3[k] * W4[k][c]
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float4* wr = reinterpret_cast<const float4*>(s_w4 + (ch * 32 + i) * 16);
                const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
                o[0] = fmaf(v[i], w0.x, o[0]); o[1] = fmaf(v[i], w0.y, o[1]); o[2] = fmaf(v[i], w0.z, o[2]); o[3] = fmaf(v[i], w0.w, o[3]);
                o[4] = fmaf(v[i], w1.x, o[4]); o[5] = fmaf(v[i], w1.y, o[5]); o[6] = fmaf(v[i], w1.z, o[6]); o[7] = fmaf(v[i], w1.w, o[7]);
                o[8] = fmaf(v[i], w2.x, o[8]); o[9] = fmaf(v[i], w2.y, o[9]); o[10] = fmaf(v[i], w2.z, o[10]); o[11] = fmaf(v[i], w2.w, o[11]);
                o[12] = fmaf(v[i], w3.x, o[12]); o[13] = fmaf(v[i], w3.y, o[13]);
              }
            }
          }
          if (layer < 2) {
            tq_fence_before();                             // TMEM reads of this accumulator are done before the MMA overwrites it
            tq_fence_async();
            __syncwarp();
            if (lane == 0) tq_mbar_arrive(tq_smem_u32(&act_full));
          } else if (n < N) {
            const int nout = h == 0 ? 2 : h == 1 ? 9 : h == 2 ? 14 : h == 3 ? 3 : 1;
            const int hoff = h == 0 ? 0 : h == 1 ? 2 : h == 2 ? 11 : h == 3 ? 25 : 28;
#pragma unroll
            for (int c = 0; c < 14; ++c) {
              if (c >= nout) break;
              float a = o[c] + s_w4[TQ_H * 16 + c];
              if (h == 4) a = 1.f / (1.f + expf(-a));
              if (h == 0 && !s_in_img[r]) a = cam.out_dist;
              out[((size_t)b * 29 + hoff + c) * N + n] = a;
            }
          }
        }
      }
      // the group's accumulators are drained: order these TMEM reads before the next pass's first MMAs
      tq_fence_before();
    }
    if (amax > 65504.f) atomicAdd(overflow, 1);
  }
  tq_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace vt

using namespace vt;

extern "C" {

int vt_query_fwd_tc(const float* points, const float* crop_center, const float* body_center, int B, int N, const float* im_feat,
                    const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht, int Wt, const float* cam7,
                    const float* wpack, const void* w1_hi, const void* w1_lo, const void* w23_hi, const void* w23_lo, float* out,
                    float* xy_out, int* overflow, void* stream) {
  return vt_query_fwd_tc_heads(points, crop_center, body_center, B, N, im_feat, tmpx, tri_tmpx, tri_feat, Hf, Wf, Ht, Wt, cam7, wpack, w1_hi,
                               w1_lo, w23_hi, w23_lo, 31, out, xy_out, overflow, stream);
}

int vt_query_fwd_tc_heads(const float* points, const float* crop_center, const float* body_center, int B, int N, const float* im_feat,
                          const float* tmpx, const float* tri_tmpx, const float* tri_feat, int Hf, int Wf, int Ht, int Wt, const float* cam7,
                          const float* wpack, const void* w1_hi, const void* w1_lo, const void* w23_hi, const void* w23_lo, int head_mask,
                          float* out, float* xy_out, int* overflow, void* stream) {
  VT_CHECK_ARG(head_mask > 0 && head_mask < 32, "vt_query_fwd_tc_heads: head mask %d", head_mask);
  if (B <= 0 || N <= 0) return 0;
  CUtensorMap m1h, m1l, m2h, m2l;
  int rc;
  if ((rc = tq_make_map(&m1h, w1_hi, TQ_NCHUNK * TQ_KC, 5 * TQ_H))) return rc;
  if ((rc = tq_make_map(&m1l, w1_lo, TQ_NCHUNK * TQ_KC, 5 * TQ_H))) return rc;
  if ((rc = tq_make_map(&m2h, w23_hi, TQ_H, 2 * 5 * TQ_H))) return rc;
  if ((rc = tq_make_map(&m2l, w23_lo, TQ_H, 2 * 5 * TQ_H))) return rc;
  TqMaps m{im_feat, tmpx, tri_tmpx, tri_feat, Hf, Wf, Ht, Wt};
  TqCam cam{cam7[0], cam7[1], cam7[2], cam7[3], cam7[4], cam7[5], cam7[6]};
  cudaError_t e = cudaFuncSetAttribute(query_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TQ_SMEM);
  if (e != cudaSuccess) return cuda_fail(e, "vt_query_fwd_tc smem attr");
  dim3 grid(ceil_div(N, TQ_M), B);
  const int head_stride = 616 * 128 + 128 + 2 * (128 * 128 + 128) + 128 * 16 + 16;
  query_fwd_tc_kernel<<<grid, TQ_THREADS, TQ_SMEM, (cudaStream_t)stream>>>(m1h, m1l, m2h, m2l, points, crop_center, body_center, B, N, m, cam,
                                                                          wpack, head_stride, out, xy_out, overflow, head_mask);
  VT_CHECK_LAUNCH("vt_query_fwd_tc");
  return 0;
}

}  // extern "C"
