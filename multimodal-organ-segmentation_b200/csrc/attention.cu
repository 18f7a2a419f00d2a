// Fused multi-head cross attention over voxel tokens for CrossAttentionFusion (reference
// src/models/fusion/attention_fusion.py:144-155): for every (batch, head)
//     out[n, :] = softmax_m( Q[n, :] . K[m, :] * scale ) @ V[m, :]
// on tcgen05 tensor cores, flash style: the [N x N] attention matrix (0.5 - 34 GB in the reference) never exists.
//
// Q, K, V, out are blocked token tensors [n_img * cbt][N tokens][8] bf16 (the conv kernels' activation layout with the
// voxels flattened), head h owning the channel blocks [h*hd/8, (h+1)*hd/8).  One CTA = 128 queries of one (batch, head):
//   warp 0      TMA: Q tile once, K / V tiles of 128 keys through a 2-stage ring
//   warp 1      MMA: S = Q K^T  (A, B K-major: rows of 16 B, K halves one 2 KB plane apart), then O_t = P V_t
//               (A = P K-major from shared memory, B = V MN-major straight from the blocked layout)
//   warps 2-5   one query row per thread: tcgen05.ld the score row, online softmax (exp2, running max / sum), write
//               P as bf16 into the K-major operand layout, rescale the fp32 output row kept in registers and add O_t.
// S and O_t live in TMEM (128 + hd columns).
#include <cuda.h>

#include "common.h"
#include "ptx.cuh"

namespace mmseg {

struct AttnKParams {
  int n_tok, hd, heads, q_cbt, kv_cbt, o_cbt;   // hd = padded head dim (multiple of 16)
  int q_cb0, k_cb0, v_cb0, o_cb0;               // first channel block of head 0 in each tensor
  float scale_log2e;                            // hd_real^-0.5 * log2(e)
  __nv_bfloat16* out;
  float* lse;                                   // optional [n_img][heads][n_tok]: log2-domain log-sum-exp of every query row
                                                // (m + log2 l), what the backward needs to recompute P without a second pass
};

constexpr int kAThreads = 192;
constexpr int kTile = 128;

struct __align__(16) ASmemHeader {
  uint64_t q_full, kv_full[2], kv_empty[2], s_full, p_full, pv_full;
  uint32_t tmem_ptr;
};

__host__ __device__ constexpr uint32_t idesc_kk(uint32_t n) {  // A, B K-major
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_k_mn(uint32_t n) {  // A K-major, B MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void umma_acc(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}

// Two CTAs per SM for head_dim <= 64 (256 TMEM columns and <= 113 KB of shared memory each): inside a CTA the score MMA, the
// softmax and the P V MMA of a key tile are serialised, so a second resident CTA is what keeps the tensor pipe busy during
// the other one's softmax (and the MUFU busy during its MMAs).
template <int HD>
__global__ void __launch_bounds__(kAThreads, HD <= 64 ? 2 : 1)
cross_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnKParams p) {
  constexpr uint32_t kPlane = kTile * 16;            // one channel block of a 128-token tile
  constexpr uint32_t kTileBytes = (HD / 8) * kPlane;  // Q / K / V tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  ASmemHeader* hdr = reinterpret_cast<ASmemHeader*>(smem);
  const uint32_t base = smem_u32(smem) + 128;
  const uint32_t q_smem = base;
  const uint32_t k_smem = q_smem + kTileBytes;             // 2 stages
  const uint32_t v_smem = k_smem + 2 * kTileBytes;         // 2 stages
  const uint32_t p_smem = v_smem + 2 * kTileBytes;         // [16 key blocks][128 rows][16 B]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTile, head = blockIdx.y, img = blockIdx.z;
  const int n_kt = (p.n_tok + kTile - 1) / kTile;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&hdr->q_full), 1);
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&hdr->kv_full[s]), 1); mbar_init(smem_u32(&hdr->kv_empty[s]), 1); }
    mbar_init(smem_u32(&hdr->s_full), 1);
    mbar_init(smem_u32(&hdr->p_full), 128);
    mbar_init(smem_u32(&hdr->pv_full), 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
  if (warp == 2) { tmem_alloc(smem_u32(&hdr->tmem_ptr), 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = hdr->tmem_ptr;          // 128 columns of scores
  const uint32_t tmem_o = tmem_s + 128;           // HD columns of P V_t

  if (warp == 0) {
    if (elect_one()) {
      const int qcb = img * p.q_cbt + p.q_cb0 + head * (HD / 8);
      const int kcb = img * p.kv_cbt + p.k_cb0 + head * (HD / 8);
      const int vcb = img * p.kv_cbt + p.v_cb0 + head * (HD / 8);
      mbar_arrive_expect_tx(smem_u32(&hdr->q_full), kTileBytes);
      tma_load_2d(q_smem, &tmQ, smem_u32(&hdr->q_full), 2 * q0, qcb);
      for (int t = 0; t < n_kt; ++t) {
        const uint32_t s = t & 1, ph = (t >> 1) & 1;
        mbar_wait(smem_u32(&hdr->kv_empty[s]), ph ^ 1);
        const uint32_t full = smem_u32(&hdr->kv_full[s]);
        mbar_arrive_expect_tx(full, 2 * kTileBytes);
        tma_load_2d(k_smem + s * kTileBytes, &tmK, full, 2 * t * kTile, kcb);
        tma_load_2d(v_smem + s * kTileBytes, &tmV, full, 2 * t * kTile, vcb);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t hi_k = (128u >> 4) | (1u << 14);               // K-major: SBO = 128 B (8-row groups)
      const uint32_t lbo_k = (kPlane >> 4) << 16;                   //          LBO = one plane (K halves)
      const uint32_t hi_mn = ((kPlane >> 4) & 0x3FFFu) | (1u << 14);  // MN-major V: SBO = plane (next 8 channels)
      const uint32_t lbo_mn = (128u >> 4) << 16;                    //             LBO = 8 keys
      mbar_wait(smem_u32(&hdr->q_full), 0);
      for (int t = 0; t < n_kt; ++t) {
        const uint32_t s = t & 1, ph = (t >> 1) & 1;
        mbar_wait(smem_u32(&hdr->kv_full[s]), ph);
        // S_t may overwrite S_{t-1}: the softmax warps finished reading it before arriving on p_full(t-1), which this
        // thread waited for below before issuing P V_{t-1}
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < HD / 16; ++j) {
          const uint32_t a = (((q_smem + j * 2 * kPlane) >> 4) & 0x3FFFu) | lbo_k;
          const uint32_t b = (((k_smem + s * kTileBytes + j * 2 * kPlane) >> 4) & 0x3FFFu) | lbo_k;
          umma_acc(tmem_s, a, hi_k, b, hi_k, idesc_kk(128), j > 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&hdr->s_full));
        mbar_wait(smem_u32(&hdr->p_full), t & 1);   // P_t in shared memory; O_{t-1} consumed
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kTile / 16; ++c) {
          const uint32_t a = (((p_smem + c * 2 * kPlane) >> 4) & 0x3FFFu) | lbo_k;
          const uint32_t b = (((v_smem + s * kTileBytes + c * 256) >> 4) & 0x3FFFu) | lbo_mn;
          umma_acc(tmem_o, a, hi_k, b, hi_mn, idesc_k_mn(HD), c > 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&hdr->kv_empty[s]));
        umma_commit(smem_u32(&hdr->pv_full));
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;                 // query row inside the tile = TMEM lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    float o[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* p_gen = smem + 128 + (size_t)5 * kTileBytes;   // generic pointer to p_smem
    for (int t = 0; t < n_kt; ++t) {
      mbar_wait(smem_u32(&hdr->s_full), t & 1);
      tc_fence_after();
      const int key0 = t * kTile;
      // pass 1 over the score row (TMEM reads are cheap; keeps only 16 scores in registers at a time): row maximum
      float mx = -INFINITY;
      const bool full = key0 + kTile <= p.n_tok;     // every key of this tile exists: no masking (all tiles but the last)
#pragma unroll
      for (int c = 0; c < kTile / 16; ++c) {
        float sv[16];
        tmem_ld16(tmem_s + lane_off + c * 16, sv);
        if (full) {
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, sv[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (key0 + c * 16 + i < p.n_tok) mx = fmaxf(mx, sv[i]);
        }
      }
      mx *= p.scale_log2e;                           // (scale > 0: the maximum commutes with the scaling)
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f(m_run - m_new);      // 0 on the first tile (m_run = -inf)
      // pass 2: p = exp2(s - m), row sum, P as bf16 into the K-major operand layout [key block][row][16 B]
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < kTile / 16; ++c) {
        float sv[16];
        tmem_ld16(tmem_s + lane_off + c * 16, sv);
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int key = key0 + c * 16 + hb * 8 + i;
            const float e = exp2f(fmaf(sv[hb * 8 + i], p.scale_log2e, -m_new));
            pv[i] = (full || key < p.n_tok) ? e : 0.f;
            rs += pv[i];
          }
          __nv_bfloat162 a = __floats2bfloat162_rn(pv[0], pv[1]), b = __floats2bfloat162_rn(pv[2], pv[3]);
          __nv_bfloat162 c2 = __floats2bfloat162_rn(pv[4], pv[5]), d = __floats2bfloat162_rn(pv[6], pv[7]);
          uint4 u;
          u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
          u.z = *reinterpret_cast<uint32_t*>(&c2); u.w = *reinterpret_cast<uint32_t*>(&d);
          *reinterpret_cast<uint4*>(p_gen + ((size_t)(c * 2 + hb) * kTile + row) * 16) = u;
        }
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      // make P visible to the tensor core (async proxy), then hand it over; this also releases S and O_{t-1}
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      mbar_arrive(smem_u32(&hdr->p_full));
      // rescale the running output while the P V_t MMAs execute, then add O_t
#pragma unroll
      for (int i = 0; i < HD; ++i) o[i] *= alpha;
      mbar_wait(smem_u32(&hdr->pv_full), t & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < HD / 16; ++c) {
        float v[16];
        tmem_ld16(tmem_o + lane_off + c * 16, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[c * 16 + i] += v[i];
      }
      tc_fence_before();
    }
    const int tok = q0 + row;
    if (tok < p.n_tok) {
      if (p.lse) p.lse[((size_t)img * p.heads + head) * p.n_tok + tok] = m_run + log2f(l_run);
      const float inv = 1.f / l_run;
      const int cb0 = img * p.o_cbt + p.o_cb0 + head * (HD / 8);
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) {
        __nv_bfloat162 a = __floats2bfloat162_rn(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv);
        __nv_bfloat162 b = __floats2bfloat162_rn(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
        __nv_bfloat162 c2 = __floats2bfloat162_rn(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv);
        __nv_bfloat162 d = __floats2bfloat162_rn(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
        u.z = *reinterpret_cast<uint32_t*>(&c2); u.w = *reinterpret_cast<uint32_t*>(&d);
        *reinterpret_cast<uint4*>(p.out + ((size_t)(cb0 + c) * p.n_tok + tok) * 8) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(hdr->tmem_ptr, 256);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// Flash-style backward by recomputation (no N x N matrix), deterministic (no atomics): two kernels.
//   P  = exp2(S * scale*log2e - LSE)            S = Q K^T (raw scores), LSE from the forward (log2 domain)
//   dP = dO V^T,   D = rowsum(dO o O),   dS = P o (dP - D) * scale
//   dV = P^T dO,   dK = dS^T Q                  (attention_bwd_dkv_kernel: one CTA per 128-key tile, loops over query tiles)
//   dQ = dS K                                   (attention_bwd_dq_kernel:  one CTA per 128-query tile, loops over key tiles)
// All five contractions run on tcgen05 with fp32 accumulators in TMEM; the softmax warps own one query row per thread
// (TMEM lane), recompute P / dS from the score and dP rows and write them to shared memory in the SAME [key block][row]
// [16 B] layout the forward uses for P — read K-major it is the A operand of dS K (rows = queries), read MN-major it is
// the A operand of P^T dO / dS^T Q (rows = keys): no transposition anywhere.
struct AttnBwdParams {
  int n_tok, hd, heads;
  int q_cbt, kv_cbt, do_cbt, dq_cbt, dkv_cbt;
  int q_cb0, k_cb0, v_cb0, do_cb0, dq_cb0, dk_cb0, dv_cb0;
  float scale_log2e, scale;
  const float* lse;     // [n_img][heads][n_tok]
  const float* dsum;    // [n_img][heads][n_tok]  D = rowsum(dO o O)
  __nv_bfloat16* dq;
  __nv_bfloat16* dkv;
};

struct __align__(16) ABSmemHeader {
  uint64_t res_full, ring_full[2], ring_empty[2], sp_full, pds_full, pds_free, acc_full;
  uint32_t tmem_ptr;
};

__host__ __device__ constexpr uint32_t idesc_mn_mn(uint32_t n) {  // A, B MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ uint4 pack8_bf16_attn(const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  return u;
}

// D[img][head][tok] = sum_d dO[tok, d] * O[tok, d]   (one thread per (img, head, tok))
__global__ void __launch_bounds__(256)
attention_rowdot_kernel(const __nv_bfloat16* __restrict__ o, int o_cbt, int o_cb0, const __nv_bfloat16* __restrict__ d_o,
                        int do_cbt, int do_cb0, int heads, int hd, int n_tok, float* __restrict__ dsum) {
  const int tok = blockIdx.x * blockDim.x + threadIdx.x;
  const int head = blockIdx.y, img = blockIdx.z;
  if (tok >= n_tok) return;
  float acc = 0.f;
  for (int c = 0; c < hd / 8; ++c) {
    const uint4 a = *reinterpret_cast<const uint4*>(o + ((size_t)(img * o_cbt + o_cb0 + head * (hd / 8) + c) * n_tok + tok) * 8);
    const uint4 b = *reinterpret_cast<const uint4*>(d_o + ((size_t)(img * do_cbt + do_cb0 + head * (hd / 8) + c) * n_tok + tok) * 8);
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 fa = __bfloat1622float2(ha[i]), fb = __bfloat1622float2(hb[i]);
      acc = fmaf(fa.x, fb.x, acc);
      acc = fmaf(fa.y, fb.y, acc);
    }
  }
  dsum[((size_t)img * heads + head) * n_tok + tok] = acc;
}

// MODE 0: dK / dV (resident = the K, V tiles of this CTA's 128 keys; ring = Q, dO tiles of the query loop)
// MODE 1: dQ      (resident = the Q, dO tiles of this CTA's 128 queries; ring = K, V tiles of the key loop)
template <int HD, int MODE>
__global__ void __launch_bounds__(kAThreads, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                     const __grid_constant__ AttnBwdParams p) {
  constexpr uint32_t kPlane = kTile * 16;
  constexpr uint32_t kTileBytes = (HD / 8) * kPlane;
  constexpr int STAGES = HD <= 64 ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  ABSmemHeader* hdr = reinterpret_cast<ABSmemHeader*>(smem);
  const uint32_t base = smem_u32(smem) + 128;
  const uint32_t res_a = base;                                  // MODE 0: K tile   | MODE 1: Q tile
  const uint32_t res_b = res_a + kTileBytes;                    // MODE 0: V tile   | MODE 1: dO tile
  const uint32_t ring_a = res_b + kTileBytes;                   // MODE 0: Q tiles  | MODE 1: K tiles   (STAGES)
  const uint32_t ring_b = ring_a + STAGES * kTileBytes;         // MODE 0: dO tiles | MODE 1: V tiles
  const uint32_t p_smem = ring_b + STAGES * kTileBytes;         // P  [16 key blocks][128 query rows][16 B]
  const uint32_t ds_smem = p_smem + 16 * kPlane;                // dS, same layout
  uint8_t* p_gen = smem + 128 + (size_t)(2 + 2 * STAGES) * kTileBytes;
  uint8_t* ds_gen = p_gen + 16 * kPlane;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * kTile, head = blockIdx.y, img = blockIdx.z;
  const int n_it = (p.n_tok + kTile - 1) / kTile;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&hdr->res_full), 1);
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&hdr->ring_full[s]), 1); mbar_init(smem_u32(&hdr->ring_empty[s]), 1); }
    mbar_init(smem_u32(&hdr->sp_full), 1);
    mbar_init(smem_u32(&hdr->pds_full), 128);
    mbar_init(smem_u32(&hdr->pds_free), 1);
    mbar_init(smem_u32(&hdr->acc_full), 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO); }
  if (warp == 2) { tmem_alloc(smem_u32(&hdr->tmem_ptr), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = hdr->tmem_ptr;            // 128 columns: S
  const uint32_t tmem_dp = tmem_s + 128;            // 128 columns: dP
  const uint32_t tmem_acc0 = tmem_s + 256;          // HD columns: dV (MODE 0) / dQ (MODE 1)
  const uint32_t tmem_acc1 = tmem_acc0 + HD;        // HD columns: dK (MODE 0)

  const int qcb = img * p.q_cbt + p.q_cb0 + head * (HD / 8);
  const int kcb = img * p.kv_cbt + p.k_cb0 + head * (HD / 8);
  const int vcb = img * p.kv_cbt + p.v_cb0 + head * (HD / 8);
  const int docb = img * p.do_cbt + p.do_cb0 + head * (HD / 8);

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(smem_u32(&hdr->res_full), 2 * kTileBytes);
      if (MODE == 0) {
        tma_load_2d(res_a, &tmK, smem_u32(&hdr->res_full), 2 * t0, kcb);
        tma_load_2d(res_b, &tmV, smem_u32(&hdr->res_full), 2 * t0, vcb);
      } else {
        tma_load_2d(res_a, &tmQ, smem_u32(&hdr->res_full), 2 * t0, qcb);
        tma_load_2d(res_b, &tmDO, smem_u32(&hdr->res_full), 2 * t0, docb);
      }
      for (int t = 0; t < n_it; ++t) {
        const uint32_t s = t % STAGES, ph = (t / STAGES) & 1;
        mbar_wait(smem_u32(&hdr->ring_empty[s]), ph ^ 1);
        const uint32_t full = smem_u32(&hdr->ring_full[s]);
        mbar_arrive_expect_tx(full, 2 * kTileBytes);
        if (MODE == 0) {
          tma_load_2d(ring_a + s * kTileBytes, &tmQ, full, 2 * t * kTile, qcb);
          tma_load_2d(ring_b + s * kTileBytes, &tmDO, full, 2 * t * kTile, docb);
        } else {
          tma_load_2d(ring_a + s * kTileBytes, &tmK, full, 2 * t * kTile, kcb);
          tma_load_2d(ring_b + s * kTileBytes, &tmV, full, 2 * t * kTile, vcb);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t hi_k = (128u >> 4) | (1u << 14);                 // K-major: SBO = 128 B (8-row groups)
      const uint32_t lbo_k = (kPlane >> 4) << 16;                     //          LBO = one plane (K halves)
      const uint32_t hi_mn = ((kPlane >> 4) & 0x3FFFu) | (1u << 14);  // MN-major: SBO = plane (next 8 MN elements)
      const uint32_t lbo_mn = (128u >> 4) << 16;                      //           LBO = 8 K rows
      mbar_wait(smem_u32(&hdr->res_full), 0);
      for (int t = 0; t < n_it; ++t) {
        const uint32_t s = t % STAGES, ph = (t / STAGES) & 1;
        mbar_wait(smem_u32(&hdr->ring_full[s]), ph);
        tc_fence_after();
        const uint32_t q_t = MODE == 0 ? ring_a + s * kTileBytes : res_a;
        const uint32_t do_t = MODE == 0 ? ring_b + s * kTileBytes : res_b;
        const uint32_t k_t = MODE == 0 ? res_a : ring_a + s * kTileBytes;
        const uint32_t v_t = MODE == 0 ? res_b : ring_b + s * kTileBytes;
        // S = Q K^T and dP = dO V^T (rows = queries): the softmax warps finished reading the previous S / dP before they
        // arrived on pds_full(t-1), which this thread waited for below
#pragma unroll
        for (int j = 0; j < HD / 16; ++j) {
          const uint32_t a = (((q_t + j * 2 * kPlane) >> 4) & 0x3FFFu) | lbo_k;
          const uint32_t b = (((k_t + j * 2 * kPlane) >> 4) & 0x3FFFu) | lbo_k;
          umma_acc(tmem_s, a, hi_k, b, hi_k, idesc_kk(128), j > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int j = 0; j < HD / 16; ++j) {
          const uint32_t a = (((do_t + j * 2 * kPlane) >> 4) & 0x3FFFu) | lbo_k;
          const uint32_t b = (((v_t + j * 2 * kPlane) >> 4) & 0x3FFFu) | lbo_k;
          umma_acc(tmem_dp, a, hi_k, b, hi_k, idesc_kk(128), j > 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&hdr->sp_full));
        mbar_wait(smem_u32(&hdr->pds_full), t & 1);     // P / dS of this pair in shared memory
        tc_fence_after();
        if (MODE == 0) {
          // dV += P^T dO, dK += dS^T Q: contraction over the 128 queries; A = P / dS read MN-major (rows = keys)
#pragma unroll
          for (int c = 0; c < kTile / 16; ++c) {
            const uint32_t a = (((p_smem + c * 256) >> 4) & 0x3FFFu) | lbo_mn;
            const uint32_t b = (((do_t + c * 256) >> 4) & 0x3FFFu) | lbo_mn;
            umma_acc(tmem_acc0, a, hi_mn, b, hi_mn, idesc_mn_mn(HD), (t > 0 || c > 0) ? 1u : 0u);
          }
#pragma unroll
          for (int c = 0; c < kTile / 16; ++c) {
            const uint32_t a = (((ds_smem + c * 256) >> 4) & 0x3FFFu) | lbo_mn;
            const uint32_t b = (((q_t + c * 256) >> 4) & 0x3FFFu) | lbo_mn;
            umma_acc(tmem_acc1, a, hi_mn, b, hi_mn, idesc_mn_mn(HD), (t > 0 || c > 0) ? 1u : 0u);
          }
        } else {
          // dQ += dS K: contraction over the 128 keys; A = dS read K-major (rows = queries), B = K MN-major
#pragma unroll
          for (int c = 0; c < kTile / 16; ++c) {
            const uint32_t a = (((ds_smem + c * 2 * kPlane) >> 4) & 0x3FFFu) | lbo_k;
            const uint32_t b = (((k_t + c * 256) >> 4) & 0x3FFFu) | lbo_mn;
            umma_acc(tmem_acc0, a, hi_k, b, hi_mn, idesc_k_mn(HD), (t > 0 || c > 0) ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&hdr->ring_empty[s]));
        umma_commit(smem_u32(&hdr->pds_free));
      }
      umma_commit(smem_u32(&hdr->acc_full));
    }
  } else {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;                   // TMEM lane: query row of the current query tile
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const size_t stat_base = ((size_t)img * p.heads + head) * p.n_tok;
    for (int t = 0; t < n_it; ++t) {
      const int qtok = (MODE == 0 ? t * kTile : t0) + row;
      const int key0 = MODE == 0 ? t0 : t * kTile;
      const bool qok = qtok < p.n_tok;
      const float lse = qok ? p.lse[stat_base + qtok] : 0.f;
      const float dsum = qok ? p.dsum[stat_base + qtok] : 0.f;
      mbar_wait(smem_u32(&hdr->sp_full), t & 1);
      tc_fence_after();
      mbar_wait(smem_u32(&hdr->pds_free), (t & 1) ^ 1);   // the MMAs that read the previous P / dS have completed
#pragma unroll
      for (int c = 0; c < kTile / 16; ++c) {
        float sv[16], dv[16];
        tmem_ld16(tmem_s + lane_off + c * 16, sv);
        tmem_ld16(tmem_dp + lane_off + c * 16, dv);
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          float pv[8], gv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int key = key0 + c * 16 + hb * 8 + i;
            const float pr = (qok && key < p.n_tok) ? exp2f(sv[hb * 8 + i] * p.scale_log2e - lse) : 0.f;
            pv[i] = pr;
            gv[i] = pr * (dv[hb * 8 + i] - dsum) * p.scale;
          }
          const size_t off = ((size_t)(c * 2 + hb) * kTile + row) * 16;
          if (MODE == 0) *reinterpret_cast<uint4*>(p_gen + off) = pack8_bf16_attn(pv);
          *reinterpret_cast<uint4*>(ds_gen + off) = pack8_bf16_attn(gv);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      mbar_arrive(smem_u32(&hdr->pds_full));
    }
    // epilogue: accumulator rows = this CTA's keys (MODE 0) / queries (MODE 1)
    mbar_wait(smem_u32(&hdr->acc_full), 0);
    tc_fence_after();
    const int tok = t0 + row;
    const bool tok_ok = tok < p.n_tok;
    // (tcgen05.ld is warp-collective: every lane executes it, only the global stores are guarded — a ragged last tile
    // has lanes beyond n_tok)
    auto store = [&](uint32_t tm, __nv_bfloat16* dst, int cb0) {
#pragma unroll
      for (int c = 0; c < HD / 16; ++c) {
        float v[16];
        tmem_ld16(tm + lane_off + c * 16, v);
        if (tok_ok) {
          *reinterpret_cast<uint4*>(dst + ((size_t)(cb0 + 2 * c) * p.n_tok + tok) * 8) = pack8_bf16_attn(v);
          *reinterpret_cast<uint4*>(dst + ((size_t)(cb0 + 2 * c + 1) * p.n_tok + tok) * 8) = pack8_bf16_attn(v + 8);
        }
      }
    };
    if (MODE == 0) {
      store(tmem_acc0, p.dkv, img * p.dkv_cbt + p.dv_cb0 + head * (HD / 8));
      store(tmem_acc1, p.dkv, img * p.dkv_cbt + p.dk_cb0 + head * (HD / 8));
    } else {
      store(tmem_acc0, p.dq, img * p.dq_cbt + p.dq_cb0 + head * (HD / 8));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(hdr->tmem_ptr, 512);
  }
}

// y = a + b on blocked bf16 tensors -> fp32 blocked, plus per-(image, chunk, channel) partial (sum, sum of squares) in
// the layout mmseg_instnorm_finalize reads: the residual + InstanceNorm3d of CrossAttentionFusion
// (attention_fusion.py:162).  grid (n_chunks, n_img*cb)
__global__ void __launch_bounds__(256)
add_stats_kernel(const __nv_bfloat16* __restrict__ a, int a_cbt, int a_cb0, const __nv_bfloat16* __restrict__ b,
                 int b_cbt, int b_cb0, int cb, size_t nvox, float* __restrict__ y, float* __restrict__ partial) {
  const int blk = blockIdx.y;
  const int img = blk / cb, c = blk - img * cb;
  const size_t abase = (size_t)(img * a_cbt + a_cb0 + c) * nvox * 8, bbase = (size_t)(img * b_cbt + b_cb0 + c) * nvox * 8;
  const size_t ybase = (size_t)blk * nvox * 8;
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (size_t)gridDim.x * blockDim.x) {
    const uint4 ua = *reinterpret_cast<const uint4*>(a + abase + v * 8), ub = *reinterpret_cast<const uint4*>(b + bbase + v * 8);
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&ua);
    const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&ub);
    float r[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 fa = __bfloat1622float2(ha[i]), fb = __bfloat1622float2(hb[i]);
      r[2 * i] = fa.x + fb.x;
      r[2 * i + 1] = fa.y + fb.y;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1[i] += r[i]; s2[i] = fmaf(r[i], r[i], s2[i]); }
    float4* d = reinterpret_cast<float4*>(y + ybase + v * 8);
    d[0] = make_float4(r[0], r[1], r[2], r[3]);
    d[1] = make_float4(r[4], r[5], r[6], r[7]);
  }
  __shared__ float red[8][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], o);
      s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], o);
    }
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < 8; ++i) { red[warp][2 * i] = s1[i]; red[warp][2 * i + 1] = s2[i]; }
  __syncthreads();
  if (threadIdx.x < 16) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    // partial[img][chunk][C][2], C = cb*8
    partial[(((size_t)img * gridDim.x + blockIdx.x) * cb * 8 + c * 8) * 2 + threadIdx.x] = t;
  }
}

typedef CUresult (*PFN_encodeTiledA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledA get_encode_fn_a() {
  static PFN_encodeTiledA fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiledA>(f);
    cudaGetLastError();
  }
  return fn;
}

static int make_tok_map(PFN_encodeTiledA enc, CUtensorMap* tm, const void* ptr, int n_tok, int blocks, int hd) {
  cuuint64_t dims[2] = {(cuuint64_t)2 * n_tok, (cuuint64_t)blocks};
  cuuint64_t strides[1] = {(cuuint64_t)n_tok * 16};
  cuuint32_t box[2] = {(cuuint32_t)(2 * kTile), (cuuint32_t)(hd / 8)};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return cr == CUDA_SUCCESS ? 0 : (int)cr;
}

}  // namespace mmseg

using namespace mmseg;

extern "C" int mmseg_cross_attention_fwd(const void* q, int32_t q_cbt, int32_t q_cb0, const void* kv, int32_t kv_cbt,
                                         int32_t k_cb0, int32_t v_cb0, void* out, int32_t o_cbt, int32_t o_cb0,
                                         int32_t n_img, int32_t heads, int32_t head_dim, int64_t n_tok, float scale,
                                         float* lse, void* stream) {
  if (!q || !kv || !out || n_img < 1 || heads < 1 || n_tok < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "cross_attention: bad arguments");
  if (head_dim != 16 && head_dim != 32 && head_dim != 64 && head_dim != 128)
    return fail(MMSEG_ERR_UNSUPPORTED, "cross_attention: head_dim=%d (supported: 16, 32, 64, 128; pad 8 to 16)", head_dim);
  if (n_tok > (1 << 30)) return fail(MMSEG_ERR_INVALID_ARG, "cross_attention: too many tokens");
  PFN_encodeTiledA enc = get_encode_fn_a();
  if (!enc) return fail(MMSEG_ERR_NO_DRIVER, "cross_attention: cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  CUtensorMap tq, tk, tv;
  int rc = make_tok_map(enc, &tq, q, (int)n_tok, n_img * q_cbt, head_dim);
  if (!rc) rc = make_tok_map(enc, &tk, kv, (int)n_tok, n_img * kv_cbt, head_dim);
  if (!rc) rc = make_tok_map(enc, &tv, kv, (int)n_tok, n_img * kv_cbt, head_dim);
  if (rc) return fail(MMSEG_ERR_CUDA, "cross_attention: cuTensorMapEncodeTiled failed (%d)", rc);
  AttnKParams p;
  p.n_tok = (int)n_tok; p.hd = head_dim; p.heads = heads; p.q_cbt = q_cbt; p.kv_cbt = kv_cbt; p.o_cbt = o_cbt;
  p.q_cb0 = q_cb0; p.k_cb0 = k_cb0; p.v_cb0 = v_cb0; p.o_cb0 = o_cb0;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  const uint32_t tile_bytes = (uint32_t)(head_dim / 8) * kTile * 16;
  const uint32_t smem = 128 + 128 + 5 * tile_bytes + 16 * kTile * 16 + 128;
  dim3 grid((unsigned)((n_tok + kTile - 1) / kTile), (unsigned)heads, (unsigned)n_img);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define MMSEG_ATTN(HD)                                                                                               \
  {                                                                                                                  \
    static bool set = false;                                                                                         \
    if (!set) {                                                                                                      \
      cudaError_t e = cudaFuncSetAttribute(cross_attention_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
      if (e != cudaSuccess) return fail(MMSEG_ERR_CUDA, "cross_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));     \
      set = true;                                                                                                    \
    }                                                                                                                \
    cross_attention_kernel<HD><<<grid, kAThreads, smem, st>>>(tq, tk, tv, p);                                        \
  }
  switch (head_dim) {
    case 16: MMSEG_ATTN(16); break;
    case 32: MMSEG_ATTN(32); break;
    case 64: MMSEG_ATTN(64); break;
    default: MMSEG_ATTN(128); break;
  }
#undef MMSEG_ATTN
  return check_launch("cross_attention_kernel");
}

extern "C" int mmseg_add_stats(const void* a, int32_t a_cbt, int32_t a_cb0, const void* b, int32_t b_cbt,
                               int32_t b_cb0, int32_t n_img, int32_t cb, int64_t voxels, float* y, float* partial,
                               int32_t n_chunks, void* stream) {
  if (!a || !b || !y || !partial || n_img < 1 || cb < 1 || voxels < 1 || n_chunks < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "add_stats: bad arguments");
  dim3 grid((unsigned)n_chunks, (unsigned)(n_img * cb));
  add_stats_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(a), a_cbt, a_cb0, reinterpret_cast<const __nv_bfloat16*>(b), b_cbt, b_cb0,
      cb, (size_t)voxels, y, partial);
  return check_launch("add_stats_kernel");
}

extern "C" int mmseg_cross_attention_bwd(const void* q, int32_t q_cbt, int32_t q_cb0, const void* kv, int32_t kv_cbt,
                                         int32_t k_cb0, int32_t v_cb0, const void* out, int32_t o_cbt, int32_t o_cb0,
                                         const void* d_out, int32_t do_cbt, int32_t do_cb0, const float* lse, float* dsum,
                                         void* dq, int32_t dq_cbt, int32_t dq_cb0, void* dkv, int32_t dkv_cbt,
                                         int32_t dk_cb0, int32_t dv_cb0, int32_t n_img, int32_t heads, int32_t head_dim,
                                         int64_t n_tok, float scale, void* stream) {
  if (!q || !kv || !out || !d_out || !lse || !dsum || !dq || !dkv || n_img < 1 || heads < 1 || n_tok < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "cross_attention_bwd: bad arguments");
  if (head_dim != 16 && head_dim != 32 && head_dim != 64 && head_dim != 128)
    return fail(MMSEG_ERR_UNSUPPORTED, "cross_attention_bwd: head_dim=%d (supported: 16, 32, 64, 128)", head_dim);
  if (n_tok > (1 << 30)) return fail(MMSEG_ERR_INVALID_ARG, "cross_attention_bwd: too many tokens");
  PFN_encodeTiledA enc = get_encode_fn_a();
  if (!enc) return fail(MMSEG_ERR_NO_DRIVER, "cross_attention_bwd: cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  CUtensorMap tq, tk, tv, tdo;
  int rc = make_tok_map(enc, &tq, q, (int)n_tok, n_img * q_cbt, head_dim);
  if (!rc) rc = make_tok_map(enc, &tk, kv, (int)n_tok, n_img * kv_cbt, head_dim);
  if (!rc) rc = make_tok_map(enc, &tv, kv, (int)n_tok, n_img * kv_cbt, head_dim);
  if (!rc) rc = make_tok_map(enc, &tdo, d_out, (int)n_tok, n_img * do_cbt, head_dim);
  if (rc) return fail(MMSEG_ERR_CUDA, "cross_attention_bwd: cuTensorMapEncodeTiled failed (%d)", rc);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  {
    dim3 g((unsigned)((n_tok + 255) / 256), (unsigned)heads, (unsigned)n_img);
    attention_rowdot_kernel<<<g, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(out), o_cbt, o_cb0,
                                               reinterpret_cast<const __nv_bfloat16*>(d_out), do_cbt, do_cb0, heads, head_dim,
                                               (int)n_tok, dsum);
    rc = check_launch("attention_rowdot_kernel");
    if (rc) return rc;
  }
  AttnBwdParams p;
  p.n_tok = (int)n_tok; p.hd = head_dim; p.heads = heads;
  p.q_cbt = q_cbt; p.kv_cbt = kv_cbt; p.do_cbt = do_cbt; p.dq_cbt = dq_cbt; p.dkv_cbt = dkv_cbt;
  p.q_cb0 = q_cb0; p.k_cb0 = k_cb0; p.v_cb0 = v_cb0; p.do_cb0 = do_cb0; p.dq_cb0 = dq_cb0; p.dk_cb0 = dk_cb0; p.dv_cb0 = dv_cb0;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.lse = lse; p.dsum = dsum;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.dkv = reinterpret_cast<__nv_bfloat16*>(dkv);
  const uint32_t tile_bytes = (uint32_t)(head_dim / 8) * kTile * 16;
  const int stages = head_dim <= 64 ? 2 : 1;
  uint32_t smem = 128 + 128 + (2 + 2 * stages) * tile_bytes + 2 * 16 * kTile * 16 + 128;
  // every CTA allocates all 512 TMEM columns: ask for more than half of the shared memory so that a second CTA can never
  // be co-resident and block inside tcgen05.alloc (same rule as conv_tc.cu / wgrad_tc.cu)
  if (smem < 118u * 1024u) smem = 118u * 1024u;
  dim3 grid((unsigned)((n_tok + kTile - 1) / kTile), (unsigned)heads, (unsigned)n_img);
#define MMSEG_ATTN_BWD(HD)                                                                                           \
  {                                                                                                                  \
    static bool set = false;                                                                                         \
    if (!set) {                                                                                                      \
      cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel<HD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
      if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_bwd_kernel<HD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
      if (e != cudaSuccess) return fail(MMSEG_ERR_CUDA, "cross_attention_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); \
      set = true;                                                                                                    \
    }                                                                                                                \
    attention_bwd_kernel<HD, 0><<<grid, kAThreads, smem, st>>>(tq, tk, tv, tdo, p);               \
    rc = check_launch("attention_bwd_kernel<dkv>");                                                                  \
    if (rc) return rc;                                                                                               \
    attention_bwd_kernel<HD, 1><<<grid, kAThreads, smem, st>>>(tq, tk, tv, tdo, p);               \
  }
  switch (head_dim) {
    case 16: MMSEG_ATTN_BWD(16); break;
    case 32: MMSEG_ATTN_BWD(32); break;
    case 64: MMSEG_ATTN_BWD(64); break;
    default: MMSEG_ATTN_BWD(128); break;
  }
#undef MMSEG_ATTN_BWD
  return check_launch("attention_bwd_kernel<dq>");
}
