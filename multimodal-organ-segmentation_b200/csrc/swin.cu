// SwinUNETR pieces that are not convolutions (BASELINE.json configs[3]; reference src/models/backbones/swin_unetr.py:80-117
// -> monai.networks.nets.SwinUNETR): patch embedding, token LayerNorm (+ residual add), patch merging, shifted-window
// attention with relative-position bias, and the residual InstanceNorm + LeakyReLU of UnetResBlock.  The linear layers
// (qkv / proj / mlp / reduction) are 1x1x1 GEMMs on the tcgen05 conv kernel (conv_tc.cu); everything stays in the blocked
// layout [n_img * cb][Z][Y][X][8]: the residual stream in fp32, GEMM operands in the 16-bit element format.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <type_traits>

#include "common.h"
#include "ptx.cuh"

namespace mmseg {

// ------------------------------------------------------------------------------------------------ patch embedding
// Conv3d(Cin, F, k=2, s=2) + bias: NCDHW fp32 image -> blocked fp32 tokens.  K = Cin*8 <= 64 products per output.
__global__ void __launch_bounds__(256)
swin_patch_embed_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                        float* __restrict__ xs, int n_img, int Cin, int F, int Z, int Y, int X) {
  extern __shared__ float sw[];   // [F][Cin*8] + [F]
  const int K = Cin * 8;
  for (int i = threadIdx.x; i < F * K; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < F; i += blockDim.x) sw[F * K + i] = b ? b[i] : 0.f;
  __syncthreads();
  const size_t nvox = (size_t)Z * Y * X;
  const int cb = F / 8;
  const int blk = blockIdx.y;   // img*cb + c
  const int img = blk / cb, c = blk - img * cb;
  const size_t ivox = nvox * 8;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (size_t)gridDim.x * blockDim.x) {
    const int xo = (int)(v % X);
    const size_t r = v / X;
    const int yo = (int)(r % Y), zo = (int)(r / Y);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = sw[F * K + c * 8 + j];
    for (int ci = 0; ci < Cin; ++ci) {
      const float* p = x + ((size_t)img * Cin + ci) * ivox;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int dz = t >> 2, dy = (t >> 1) & 1, dx = t & 1;
        const float in = p[((size_t)(2 * zo + dz) * (2 * Y) + (2 * yo + dy)) * (2 * X) + (2 * xo + dx)];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(in, sw[(c * 8 + j) * K + ci * 8 + t], acc[j]);
      }
    }
    float4* d = reinterpret_cast<float4*>(xs + ((size_t)blk * nvox + v) * 8);
    d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// ------------------------------------------------------------------------------------------------ token LayerNorm
__device__ __forceinline__ void ld8f(const float* p, float* v) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8f(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// xs (+= add) ; dst = LayerNorm_C(xs) * gamma + beta.  One token per thread, consecutive threads = consecutive voxels, so
// every load instruction of a warp covers 1 KB of one channel block.  Three passes (sum / centred squares / normalise):
// the token's C*4 bytes stay in L1/L2 between them.
__global__ void __launch_bounds__(128)
swin_layernorm_kernel(float* __restrict__ xs, const float* __restrict__ add, const float* __restrict__ gamma,
                      const float* __restrict__ beta, void* __restrict__ dst, int n_img, int cb, size_t nvox, int dst_cbt,
                      int dst_cb_off, float eps, int fp16) {
  const size_t tok = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= (size_t)n_img * nvox) return;
  const int img = (int)(tok / nvox);
  const size_t v = tok - (size_t)img * nvox;
  const size_t bs = nvox * 8;
  float* px = xs + (size_t)img * cb * bs + v * 8;
  float s = 0.f;
  if (add) {
    const float* pa = add + (size_t)img * cb * bs + v * 8;
    for (int b = 0; b < cb; ++b) {
      float a[8], y[8];
      ld8f(px + b * bs, a);
      ld8f(pa + b * bs, y);
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] += y[i]; s += a[i]; }
      st8f(px + b * bs, a);
    }
  } else {
    for (int b = 0; b < cb; ++b) {
      float a[8];
      ld8f(px + b * bs, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += a[i];
    }
  }
  if (!dst) return;
  const float inv_c = 1.f / (float)(cb * 8);
  const float mean = s * inv_c;
  float q = 0.f;
  for (int b = 0; b < cb; ++b) {
    float a[8];
    ld8f(px + b * bs, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = a[i] - mean; q = fmaf(d, d, q); }
  }
  const float rstd = rsqrtf(q * inv_c + eps);
  uint16_t* pd = reinterpret_cast<uint16_t*>(dst);
  const size_t dbase = ((size_t)img * dst_cbt + dst_cb_off) * bs + v * 8;
  for (int b = 0; b < cb; ++b) {
    float a[8];
    ld8f(px + b * bs, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = (a[i] - mean) * rstd;
      if (gamma) y = fmaf(y, gamma[b * 8 + i], beta ? beta[b * 8 + i] : 0.f);
      a[i] = y;
    }
    store8_act(pd, dbase + b * bs, 0, a, fp16 != 0);
  }
}

// Warp-per-token variant for the deep stages (few tokens, C >= 192): lane l owns channel blocks l, l+32, ...; the
// thread-per-token kernel would leave most SMs idle there (216 tokens x 384 channels on one warp's worth of threads).
__global__ void __launch_bounds__(256)
swin_layernorm_warp_kernel(float* __restrict__ xs, const float* __restrict__ add, const float* __restrict__ gamma,
                           const float* __restrict__ beta, void* __restrict__ dst, int n_img, int cb, size_t nvox,
                           int dst_cbt, int dst_cb_off, float eps, int fp16) {
  const size_t tok = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tok >= (size_t)n_img * nvox) return;
  const int img = (int)(tok / nvox);
  const size_t v = tok - (size_t)img * nvox;
  const size_t bs = nvox * 8;
  float* px = xs + (size_t)img * cb * bs + v * 8;
  const float* pa = add ? add + (size_t)img * cb * bs + v * 8 : nullptr;
  float s = 0.f;
  for (int b = lane; b < cb; b += 32) {
    float a[8];
    ld8f(px + b * bs, a);
    if (pa) {
      float y[8];
      ld8f(pa + b * bs, y);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] += y[i];
      st8f(px + b * bs, a);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
  }
  if (!dst) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv_c = 1.f / (float)(cb * 8);
  const float mean = s * inv_c;
  float q = 0.f;
  for (int b = lane; b < cb; b += 32) {
    float a[8];
    ld8f(px + b * bs, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = a[i] - mean; q = fmaf(d, d, q); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * inv_c + eps);
  uint16_t* pd = reinterpret_cast<uint16_t*>(dst);
  const size_t dbase = ((size_t)img * dst_cbt + dst_cb_off) * bs + v * 8;
  for (int b = lane; b < cb; b += 32) {
    float a[8];
    ld8f(px + b * bs, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = (a[i] - mean) * rstd;
      if (gamma) y = fmaf(y, gamma[b * 8 + i], beta ? beta[b * 8 + i] : 0.f);
      a[i] = y;
    }
    store8_act(pd, dbase + b * bs, 0, a, fp16 != 0);
  }
}

// ------------------------------------------------------------------------------------------------ patch merging
// MONAI PatchMerging (v1, 3-D): the eight gathered sub-grids in its legacy order, concatenated on channels, LayerNorm(8C)
// with affine; the bias-free Linear(8C -> 2C) that follows is a 1x1x1 GEMM on the conv kernel.
__constant__ int kMergeOff[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {0, 1, 0}, {0, 0, 1}, {1, 1, 1}};

// One warp per output token: lane l owns the (octant, channel block) pairs l, l+32, ... of the 8C-channel row.
__global__ void __launch_bounds__(256)
swin_merge_ln_kernel(const float* __restrict__ xs, const float* __restrict__ gamma, const float* __restrict__ beta,
                     void* __restrict__ dst, int n_img, int cb, int Z, int Y, int X, float eps, int fp16) {
  const int Zo = Z / 2, Yo = Y / 2, Xo = X / 2;
  const size_t nout = (size_t)Zo * Yo * Xo, nvox = (size_t)Z * Y * X;
  const size_t tok = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tok >= (size_t)n_img * nout) return;
  const int img = (int)(tok / nout);
  const size_t v = tok - (size_t)img * nout;
  const int xo = (int)(v % Xo);
  const size_t r = v / Xo;
  const int yo = (int)(r % Yo), zo = (int)(r / Yo);
  const float* base = xs + (size_t)img * cb * nvox * 8;
  const int npair = 8 * cb;
  auto src_of = [&](int p) -> const float* {
    const int o = p / cb, b = p - o * cb;
    const size_t off = (((size_t)(2 * zo + kMergeOff[o][0]) * Y + (2 * yo + kMergeOff[o][1])) * X + (2 * xo + kMergeOff[o][2])) * 8;
    return base + (size_t)b * nvox * 8 + off;
  };
  float s = 0.f;
  for (int p = lane; p < npair; p += 32) {
    float a[8];
    ld8f(src_of(p), a);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv_c = 1.f / (float)(cb * 64);
  const float mean = s * inv_c;
  float q = 0.f;
  for (int p = lane; p < npair; p += 32) {
    float a[8];
    ld8f(src_of(p), a);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = a[i] - mean; q = fmaf(d, d, q); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * inv_c + eps);
  uint16_t* pd = reinterpret_cast<uint16_t*>(dst);
  const size_t dbase = (size_t)img * npair * nout * 8 + v * 8;
  for (int p = lane; p < npair; p += 32) {
    float a[8];
    ld8f(src_of(p), a);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaf((a[i] - mean) * rstd, gamma[p * 8 + i], beta[p * 8 + i]);
    store8_act(pd, dbase + (size_t)p * nout * 8, 0, a, fp16 != 0);
  }
}

// ------------------------------------------------------------------------------------------------ window attention
// One CTA = one (window, head, image); head_dim = 16 (every SwinUNETR stage: C / heads = 16), so Q K^T is ONE
// m16n8k16 tensor-core MMA per 16 x 8 score tile and P V one per 16 keys x 8 dims.  Flash-style: a warp owns 16 queries,
// walks the keys in blocks of 64 with an online softmax, P goes from the score fragments straight into the A fragments of
// the P V MMA (no shared-memory round trip).  The window's tokens are gathered with the cyclic shift and the zero padding
// folded into the addressing: a padded token carries q/k/v = the qkv bias (MONAI pads AFTER norm1, so its qkv input is 0).
// relative-position bias: table[h][base(i) - base(j) + centre] from a 2197-entry shared-memory table (base(i) = the
// (2w-1)-radix code of i's coordinates in the CONFIGURED window — which is also what MONAI's index[:n, :n] slice gives
// when the window shrank to the feature-map size); shift mask: -100 between tokens of different shift regions.
struct SwinAttnK {
  const void* qkv;
  void* out;
  const float* table;      // [table_len][heads]
  const float* qkv_bias;   // [3C] or NULL
  float* lse;              // optional [img][head][window][352]: log2-domain log-sum-exp rows (saved for the backward)
  int n_img, D, H, W;
  int ws0, ws1, ws2;       // actual window (min(configured, extent))
  int cw1, cw2;            // configured window extents along h, w (relative-position radix)
  int s0, s1, s2;          // cyclic shift (0 = none)
  int Dp, Hp, Wp;          // padded extents
  int heads, qkv_cbt, out_cbt, out_cb_off;
  int table_len, centre;
  float scale_log2e;
};

constexpr int kAttnMaxTok = 352;       // 343 rounded up to 16
constexpr int kAttnQS = 24;            // halfs per Q / K row (16 + pad: conflict-free fragment loads)
constexpr int kAttnVS = 360;           // halfs per V^T row
constexpr int kAttnTab = 2200;         // (2*7-1)^3 = 2197 table rows
constexpr size_t kAttnSmem = (size_t)kAttnMaxTok * kAttnQS * 2 + 16 * kAttnVS * 2 + kAttnTab * 4 + kAttnMaxTok * 8 + 32 * 4 + 16;

template <bool FP16>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (FP16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
template <bool FP16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if (FP16) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <bool FP16>
__device__ __forceinline__ uint16_t cvt1(float v) {
  if (FP16) { const __half h = __float2half_rn(v); return *reinterpret_cast<const uint16_t*>(&h); }
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<const uint16_t*>(&h);
}

template <bool FP16>
__global__ void __launch_bounds__(256) swin_window_attention_kernel(const SwinAttnK k) {
  extern __shared__ __align__(16) uint8_t attn_smem[];
  // (the Q fragments of a warp's 16 queries are read straight from global memory, once: no shared-memory copy — 40 KB per
  // CTA instead of 57 KB lets five CTAs share an SM instead of three)
  uint16_t* sK = reinterpret_cast<uint16_t*>(attn_smem);                 // [kAttnMaxTok][kAttnQS]
  uint16_t* sVt = sK + kAttnMaxTok * kAttnQS;                            // [16][kAttnVS]
  float* sTab = reinterpret_cast<float*>(sVt + 16 * kAttnVS);            // [kAttnTab] relative-position bias of this head
  int* sPos = reinterpret_cast<int*>(sTab + kAttnTab);                   // voxel index of the token, -1 = padded token
  int* sKey = sPos + kAttnMaxTok;                                        // relative-position code | shift-mask region << 16

  const int n_tok = k.ws0 * k.ws1 * k.ws2;
  const int np = (n_tok + 15) & ~15;
  const int nW1 = k.Hp / k.ws1, nW2 = k.Wp / k.ws2;
  const int win = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
  const int w2 = win % nW2, w1 = (win / nW2) % nW1, w0 = win / (nW2 * nW1);
  const size_t nvox = (size_t)k.D * k.H * k.W;
  const int C = k.heads * 16;
  const bool shifted = (k.s0 | k.s1 | k.s2) != 0;

  for (int i = threadIdx.x; i < k.table_len; i += blockDim.x) sTab[i] = k.table[(size_t)i * k.heads + head] * 1.4426950408889634f;
  for (int i = threadIdx.x; i < kAttnMaxTok + 32; i += blockDim.x) {   // (+32: the last 64-key block reads past np)
    int pos = -1, base = 0, reg = 0;
    if (i < n_tok) {
      const int t2 = i % k.ws2, t1 = (i / k.ws2) % k.ws1, t0 = i / (k.ws2 * k.ws1);
      const int g0 = w0 * k.ws0 + t0, g1 = w1 * k.ws1 + t1, g2 = w2 * k.ws2 + t2;
      int p0 = g0 + k.s0, p1 = g1 + k.s1, p2 = g2 + k.s2;
      if (p0 >= k.Dp) p0 -= k.Dp;
      if (p1 >= k.Hp) p1 -= k.Hp;
      if (p2 >= k.Wp) p2 -= k.Wp;
      if (p0 < k.D && p1 < k.H && p2 < k.W) pos = (p0 * k.H + p1) * k.W + p2;
      // MONAI: relative_position_index[:n, :n] of the configured window, addressed by the FLAT token index
      const int c2 = i % k.cw2, c1 = (i / k.cw2) % k.cw1, c0 = i / (k.cw2 * k.cw1);
      base = (c0 * (2 * k.cw1 - 1) + c1) * (2 * k.cw2 - 1) + c2;
      if (shifted) {
        // MONAI compute_mask: slices [0, P-ws), [P-ws, P-s), [P-s, P) per axis; with s = 0 the last slice covers the
        // whole axis, i.e. one region
        const int r0 = k.s0 == 0 ? 0 : (g0 < k.Dp - k.ws0 ? 0 : (g0 < k.Dp - k.s0 ? 1 : 2));
        const int r1 = k.s1 == 0 ? 0 : (g1 < k.Hp - k.ws1 ? 0 : (g1 < k.Hp - k.s1 ? 1 : 2));
        const int r2 = k.s2 == 0 ? 0 : (g2 < k.Wp - k.ws2 ? 0 : (g2 < k.Wp - k.s2 ? 1 : 2));
        reg = (r0 * 3 + r1) * 3 + r2;
      }
    }
    if (i < kAttnMaxTok) sPos[i] = pos;
    sKey[i] = base | (reg << 16);
  }
  __syncthreads();
  // gather q / k / v of this head: 2 channel blocks each (16 dims)
  const uint16_t* qkv = reinterpret_cast<const uint16_t*>(k.qkv);
  for (int e = threadIdx.x; e < np * 4; e += blockDim.x) {
    const int i = e >> 2, part = e & 3;          // part: 0,1 = k blocks; 2,3 = v
    const int which = 1 + (part >> 1), half = part & 1;
    const int ch0 = which * C + head * 16 + half * 8;
    const int pos = i < n_tok ? sPos[i] : -2;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (pos >= 0) {
      u = *reinterpret_cast<const uint4*>(qkv + (((size_t)img * k.qkv_cbt + (ch0 >> 3)) * nvox + pos) * 8);
    } else if (pos == -1 && k.qkv_bias) {
      float b[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = k.qkv_bias[ch0 + j];
      u = cvt8_from_f32(b, FP16);
    }
    if (which == 1) {
      *reinterpret_cast<uint4*>(sK + i * kAttnQS + half * 8) = u;
    } else {
      const uint16_t* h = reinterpret_cast<const uint16_t*>(&u);
#pragma unroll
      for (int j = 0; j < 8; ++j) sVt[(half * 8 + j) * kAttnVS + i] = h[j];
    }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n_qt = np >> 4;
  uint16_t* out = reinterpret_cast<uint16_t*>(k.out);
  for (int qt = warp; qt < n_qt; qt += (blockDim.x >> 5)) {
    const int q0 = qt * 16;
    uint32_t qa[4];
    {
      const int pq0 = (q0 + g) < n_tok ? sPos[q0 + g] : -1, pq1 = (q0 + g + 8) < n_tok ? sPos[q0 + g + 8] : -1;
      const size_t qb = ((size_t)img * k.qkv_cbt + head * 2) * nvox;   // first of the head's two q channel blocks
      // a padded query's row is cropped from the output: its q does not matter
      qa[0] = pq0 >= 0 ? *reinterpret_cast<const uint32_t*>(qkv + (qb + pq0) * 8 + 2 * t) : 0u;
      qa[1] = pq1 >= 0 ? *reinterpret_cast<const uint32_t*>(qkv + (qb + pq1) * 8 + 2 * t) : 0u;
      qa[2] = pq0 >= 0 ? *reinterpret_cast<const uint32_t*>(qkv + (qb + nvox + pq0) * 8 + 2 * t) : 0u;
      qa[3] = pq1 >= 0 ? *reinterpret_cast<const uint32_t*>(qkv + (qb + nvox + pq1) * 8 + 2 * t) : 0u;
    }
    const int bq0 = (sKey[q0 + g] & 0xffff) + k.centre, bq1 = (sKey[q0 + g + 8] & 0xffff) + k.centre;
    const int rq0 = sKey[q0 + g] >> 16, rq1 = sKey[q0 + g + 8] >> 16;
    float m0 = -1e30f, m1 = -1e30f, l0 = 0.f, l1 = 0.f;
    float o[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    for (int kb = 0; kb < np; kb += 64) {
      float s[8][4];
      // scores of 64 keys: scale (log2 domain), relative-position bias, shift mask, key padding.  FULL = every key of the
      // block exists (all blocks but the last): no bounds checks.  The per-key code (base | region << 16) of the two keys a
      // lane owns in an n-tile comes with one 8-byte load.
      auto score_block = [&](auto full_c) {
        constexpr bool FULL = decltype(full_c)::value;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int key0 = kb + nt * 8;
#pragma unroll
          for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
          if (FULL || key0 < np) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(sK + (key0 + g) * kAttnQS + 2 * t);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(sK + (key0 + g) * kAttnQS + 2 * t + 8);
            mma16816<FP16>(s[nt], qa, b0, b1);
          }
          const int kc = key0 + 2 * t;
          const int2 ki = *reinterpret_cast<const int2*>(sKey + kc);
          const int kinfo[2] = {ki.x, ki.y};
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (FULL || kc + j < n_tok) {
              const int bk = kinfo[j] & 0xffff;
              float v0 = fmaf(s[nt][j], k.scale_log2e, sTab[bq0 - bk]);
              float v1 = fmaf(s[nt][2 + j], k.scale_log2e, sTab[bq1 - bk]);
              if (shifted) {
                const int rk = kinfo[j] >> 16;
                v0 = rk == rq0 ? v0 : v0 - 144.26950408889634f;   // -100 in the log2 domain
                v1 = rk == rq1 ? v1 : v1 - 144.26950408889634f;
              }
              s[nt][j] = v0;
              s[nt][2 + j] = v1;
            } else {
              s[nt][j] = -1e30f;
              s[nt][2 + j] = -1e30f;
            }
          }
        }
      };
      if (kb + 64 <= n_tok) score_block(std::true_type{});
      else score_block(std::false_type{});
      float mx0 = m0, mx1 = m1;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float c0 = exp2f(m0 - mx0), c1 = exp2f(m1 - mx1);
      m0 = mx0; m1 = mx1;
      l0 *= c0; l1 *= c1;
#pragma unroll
      for (int i = 0; i < 2; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = exp2f(s[nt][0] - m0); s[nt][1] = exp2f(s[nt][1] - m0);
        s[nt][2] = exp2f(s[nt][2] - m1); s[nt][3] = exp2f(s[nt][3] - m1);
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {   // 16 keys per P V step
        const int key0 = kb + kk * 16;
        if (key0 < np) {
          uint32_t pa[4];
          pa[0] = pack2<FP16>(s[2 * kk][0], s[2 * kk][1]);
          pa[1] = pack2<FP16>(s[2 * kk][2], s[2 * kk][3]);
          pa[2] = pack2<FP16>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          pa[3] = pack2<FP16>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
          for (int nd = 0; nd < 2; ++nd) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(sVt + (nd * 8 + g) * kAttnVS + key0 + 2 * t);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(sVt + (nd * 8 + g) * kAttnVS + key0 + 2 * t + 8);
            mma16816<FP16>(o[nd], pa, b0, b1);
          }
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    if (k.lse && t == 0) {
      float* lr = k.lse + (((size_t)img * k.heads + head) * gridDim.x + win) * kAttnMaxTok;
      lr[q0 + g] = m0 + log2f(l0);
      lr[q0 + g + 8] = m1 + log2f(l1);
    }
    const int p0 = (q0 + g) < n_tok ? sPos[q0 + g] : -1, p1 = (q0 + g + 8) < n_tok ? sPos[q0 + g + 8] : -1;
#pragma unroll
    for (int nd = 0; nd < 2; ++nd) {
      const size_t cbase = ((size_t)img * k.out_cbt + k.out_cb_off + head * 2 + nd) * nvox;
      if (p0 >= 0) *reinterpret_cast<uint32_t*>(out + (cbase + p0) * 8 + 2 * t) = pack2<FP16>(o[nd][0] * i0, o[nd][1] * i0);
      if (p1 >= 0) *reinterpret_cast<uint32_t*>(out + (cbase + p1) * 8 + 2 * t) = pack2<FP16>(o[nd][2] * i1, o[nd][3] * i1);
    }
  }
}

// ------------------------------------------------------------------------------------------------ residual IN + act
// UnetResBlock tail (MONAI dynunet_block.py): y = LeakyReLU( IN(a) + r' ), r' = IN(r) when the block has a 1x1x1
// residual conv (channel change), else the block input itself.  a / r raw conv outputs (fp32 or 16-bit), r may also be a
// 16-bit activation inside a wider buffer (r_cbt / r_cb_off).
struct ResNormK {
  const void* a;
  const void* r;
  const float* a_mr;   // [n_img][C][2]
  const float* r_mr;   // [n_img][C][2] or NULL (identity)
  void* dst;
  int n_img, cb;
  size_t nvox;
  int a_f32, r_f32, r_cbt, r_cb_off, dst_cbt, dst_cb_off;
  float slope;
};

template <bool FP16>
__global__ void __launch_bounds__(256) instnorm_residual_act_kernel(const ResNormK k) {
  const int blk = blockIdx.y;
  const int img = blk / k.cb, c = blk - img * k.cb;
  float am[8], ar[8], rm[8], rr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 m = *reinterpret_cast<const float2*>(k.a_mr + ((size_t)img * k.cb * 8 + c * 8 + i) * 2);
    am[i] = m.x; ar[i] = m.y;
    rm[i] = 0.f; rr[i] = 1.f;
    if (k.r_mr) {
      const float2 q = *reinterpret_cast<const float2*>(k.r_mr + ((size_t)img * k.cb * 8 + c * 8 + i) * 2);
      rm[i] = q.x; rr[i] = q.y;
    }
  }
  const size_t a_base = (size_t)blk * k.nvox * 8;
  const size_t r_base = ((size_t)img * k.r_cbt + k.r_cb_off + c) * k.nvox * 8;
  const size_t d_base = ((size_t)img * k.dst_cbt + k.dst_cb_off + c) * k.nvox * 8;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < k.nvox; v += (size_t)gridDim.x * blockDim.x) {
    float a[8], r[8];
    if (k.a_f32) ld8f(reinterpret_cast<const float*>(k.a) + a_base + v * 8, a); else load8_act(k.a, a_base + v * 8, a, FP16);
    if (k.r_f32) ld8f(reinterpret_cast<const float*>(k.r) + r_base + v * 8, r); else load8_act(k.r, r_base + v * 8, r, FP16);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float y = (a[i] - am[i]) * ar[i] + (r[i] - rm[i]) * rr[i];
      a[i] = y > 0.f ? y : y * k.slope;
    }
    store8_act(k.dst, d_base + v * 8, 0, a, FP16);
  }
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace mmseg

using namespace mmseg;

extern "C" int mmseg_swin_patch_embed(const float* x, const float* w, const float* b, float* xs, int32_t n_img, int32_t Cin,
                                      int32_t F, int32_t Z, int32_t Y, int32_t X, void* stream) {
  if (!x || !w || !xs || n_img < 1 || Cin < 1 || Cin > 8 || F < 8 || F % 8 || F > 256 || Z < 1 || Y < 1 || X < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "swin_patch_embed: bad arguments (Cin 1..8, F multiple of 8 <= 256)");
  const size_t nvox = (size_t)Z * Y * X;
  const size_t smem = ((size_t)F * Cin * 8 + F) * sizeof(float);
  size_t gx = (nvox + 255) / 256;
  const size_t cap = (size_t)sm_count() * 8;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)(n_img * (F / 8)));
  swin_patch_embed_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(x, w, b, xs, n_img, Cin, F, Z, Y, X);
  return check_launch("swin_patch_embed_kernel");
}

extern "C" int mmseg_swin_layernorm(float* xs, const float* add, const float* gamma, const float* beta, void* dst,
                                    int32_t n_img, int32_t cb, int64_t voxels, int32_t dst_cbt, int32_t dst_cb_off, float eps,
                                    int32_t elem_fmt, void* stream) {
  if (!xs || (!add && !dst) || n_img < 1 || cb < 1 || voxels < 1 || (beta && !gamma))
    return fail(MMSEG_ERR_INVALID_ARG, "swin_layernorm: bad arguments");
  if (elem_fmt != MMSEG_FMT_BF16 && elem_fmt != MMSEG_FMT_FP16) return fail(MMSEG_ERR_INVALID_ARG, "swin_layernorm: elem_fmt");
  const size_t tok = (size_t)n_img * voxels;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cb >= 24 || tok < 4096) {   // deep stages: a warp per token
    swin_layernorm_warp_kernel<<<(unsigned)((tok + 7) / 8), 256, 0, st>>>(
        xs, add, gamma, beta, dst, n_img, cb, (size_t)voxels, dst_cbt, dst_cb_off, eps, elem_fmt == MMSEG_FMT_FP16);
  } else {
    swin_layernorm_kernel<<<(unsigned)((tok + 127) / 128), 128, 0, st>>>(
        xs, add, gamma, beta, dst, n_img, cb, (size_t)voxels, dst_cbt, dst_cb_off, eps, elem_fmt == MMSEG_FMT_FP16);
  }
  return check_launch("swin_layernorm_kernel");
}

extern "C" int mmseg_swin_merge_ln(const float* xs, const float* gamma, const float* beta, void* dst, int32_t n_img, int32_t cb,
                                   int32_t Z, int32_t Y, int32_t X, float eps, int32_t elem_fmt, void* stream) {
  if (!xs || !gamma || !beta || !dst || n_img < 1 || cb < 1 || Z < 2 || Y < 2 || X < 2)
    return fail(MMSEG_ERR_INVALID_ARG, "swin_merge_ln: bad arguments");
  if ((Z | Y | X) & 1) return fail(MMSEG_ERR_UNSUPPORTED, "swin_merge_ln: odd extents (the zero-padded merge) are not built");
  if (elem_fmt != MMSEG_FMT_BF16 && elem_fmt != MMSEG_FMT_FP16) return fail(MMSEG_ERR_INVALID_ARG, "swin_merge_ln: elem_fmt");
  const size_t tok = (size_t)n_img * (Z / 2) * (Y / 2) * (X / 2);
  swin_merge_ln_kernel<<<(unsigned)((tok + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      xs, gamma, beta, dst, n_img, cb, Z, Y, X, eps, elem_fmt == MMSEG_FMT_FP16);
  return check_launch("swin_merge_ln_kernel");
}

extern "C" int mmseg_swin_window_attention(const mmseg_swin_attn_args* a, void* stream) {
  if (!a || !a->qkv || !a->out || !a->table) return fail(MMSEG_ERR_INVALID_ARG, "swin_window_attention: null pointer");
  if (a->n_img < 1 || a->D < 1 || a->H < 1 || a->W < 1 || a->heads < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "swin_window_attention: bad extents");
  if (a->head_dim != 16) return fail(MMSEG_ERR_UNSUPPORTED, "swin_window_attention: head_dim %d (built for 16)", a->head_dim);
  SwinAttnK k;
  k.qkv = a->qkv; k.out = a->out; k.table = a->table; k.qkv_bias = a->qkv_bias; k.lse = a->lse;
  k.n_img = a->n_img; k.D = a->D; k.H = a->H; k.W = a->W;
  const int ext[3] = {a->D, a->H, a->W};
  int ws[3], ss[3], pp[3];
  for (int i = 0; i < 3; ++i) {
    if (a->window[i] < 1 || a->shift[i] < 0 || a->shift[i] >= a->window[i])
      return fail(MMSEG_ERR_INVALID_ARG, "swin_window_attention: window / shift");
    // MONAI get_window_size: an extent that does not exceed the window shrinks the window and cancels the shift
    ws[i] = ext[i] <= a->window[i] ? ext[i] : a->window[i];
    ss[i] = ext[i] <= a->window[i] ? 0 : a->shift[i];
    pp[i] = (ext[i] + ws[i] - 1) / ws[i] * ws[i];
  }
  if (ws[0] * ws[1] * ws[2] > 343) return fail(MMSEG_ERR_UNSUPPORTED, "swin_window_attention: more than 343 tokens per window");
  k.ws0 = ws[0]; k.ws1 = ws[1]; k.ws2 = ws[2];
  k.s0 = ss[0]; k.s1 = ss[1]; k.s2 = ss[2];
  k.Dp = pp[0]; k.Hp = pp[1]; k.Wp = pp[2];
  k.cw1 = a->window[1]; k.cw2 = a->window[2];
  k.heads = a->heads; k.qkv_cbt = a->qkv_cbt; k.out_cbt = a->out_cbt; k.out_cb_off = a->out_cb_off;
  k.table_len = (2 * a->window[0] - 1) * (2 * a->window[1] - 1) * (2 * a->window[2] - 1);
  if (k.table_len > kAttnTab) return fail(MMSEG_ERR_UNSUPPORTED, "swin_window_attention: bias table of %d rows", k.table_len);
  k.centre = ((a->window[0] - 1) * (2 * a->window[1] - 1) + (a->window[1] - 1)) * (2 * a->window[2] - 1) + (a->window[2] - 1);
  k.scale_log2e = a->scale * 1.4426950408889634f;
  if (a->qkv_cbt < 6 * a->heads || a->out_cbt < a->out_cb_off + 2 * a->heads)
    return fail(MMSEG_ERR_INVALID_ARG, "swin_window_attention: channel blocks");
  if (a->elem_fmt != MMSEG_FMT_BF16 && a->elem_fmt != MMSEG_FMT_FP16) return fail(MMSEG_ERR_INVALID_ARG, "swin_window_attention: elem_fmt");
  dim3 grid((unsigned)((pp[0] / ws[0]) * (pp[1] / ws[1]) * (pp[2] / ws[2])), (unsigned)a->heads, (unsigned)a->n_img);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(swin_window_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnSmem);
    cudaFuncSetAttribute(swin_window_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnSmem);
    attr_set = true;
  }
  if (a->elem_fmt == MMSEG_FMT_FP16) swin_window_attention_kernel<true><<<grid, 256, kAttnSmem, st>>>(k);
  else swin_window_attention_kernel<false><<<grid, 256, kAttnSmem, st>>>(k);
  return check_launch("swin_window_attention_kernel");
}

extern "C" int mmseg_instnorm_residual_act(const void* a, int32_t a_is_f32, const float* a_mean_rstd, const void* r,
                                           int32_t r_is_f32, const float* r_mean_rstd, int32_t r_cbt, int32_t r_cb_off,
                                           void* dst, int32_t dst_cbt, int32_t dst_cb_off, int32_t n_img, int32_t cb,
                                           int64_t voxels, float slope, int32_t elem_fmt, void* stream) {
  if (!a || !a_mean_rstd || !r || !dst || n_img < 1 || cb < 1 || voxels < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "instnorm_residual_act: bad arguments");
  if (elem_fmt != MMSEG_FMT_BF16 && elem_fmt != MMSEG_FMT_FP16) return fail(MMSEG_ERR_INVALID_ARG, "instnorm_residual_act: elem_fmt");
  ResNormK k;
  k.a = a; k.r = r; k.a_mr = a_mean_rstd; k.r_mr = r_mean_rstd; k.dst = dst;
  k.n_img = n_img; k.cb = cb; k.nvox = (size_t)voxels;
  k.a_f32 = a_is_f32; k.r_f32 = r_is_f32; k.r_cbt = r_cbt; k.r_cb_off = r_cb_off; k.dst_cbt = dst_cbt; k.dst_cb_off = dst_cb_off;
  k.slope = slope;
  const int rows = n_img * cb;
  size_t want = ((size_t)sm_count() * 8 + rows - 1) / rows, need = ((size_t)voxels + 255) / 256;
  size_t gx = want < need ? want : need;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)rows);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (elem_fmt == MMSEG_FMT_FP16) instnorm_residual_act_kernel<true><<<grid, 256, 0, st>>>(k);
  else instnorm_residual_act_kernel<false><<<grid, 256, 0, st>>>(k);
  return check_launch("instnorm_residual_act_kernel");
}
