// Backward (training) kernels of the SwinUNETR path — the ops autograd would issue for loss.backward()
// (reference src/trainer/trainer.py:243) through monai.networks.nets.SwinUNETR (src/models/backbones/swin_unetr.py:80-117):
// token LayerNorm forward-with-statistics / backward / parameter gradients, GELU backward, the LeakyReLU mask of the
// UnetResBlock tail, patch-merging gather / scatter, patch-embedding weight gradient and the window-attention backward.
// bf16 operands, fp32 residual-stream gradients.  Every buffer is written by exactly one thread (no global float atomics);
// the relative-position-bias gradient is accumulated per CTA in shared memory in fixed point (exact integer adds) and reduced
// across CTAs in a fixed order.
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.h"
#include "ptx.cuh"

namespace mmseg {

__device__ __forceinline__ void ld8(const float* p, float* v) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// ------------------------------------------------------------------------------------------------ LayerNorm (training)
// xs_out = xs_in (+ add16);  ln = LN(xs_out) * gamma + beta (bf16);  stats[token] = (mean, rstd).  One token per thread.
__global__ void __launch_bounds__(128)
swin_ln_fwd_train_kernel(const float* __restrict__ xs_in, const void* __restrict__ add16, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float* __restrict__ xs_out, void* __restrict__ ln,
                         float* __restrict__ stats, int n_img, int cb, size_t nvox, float eps) {
  const size_t tok = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= (size_t)n_img * nvox) return;
  const int img = (int)(tok / nvox);
  const size_t v = tok - (size_t)img * nvox;
  const size_t bs = nvox * 8;
  const size_t base = (size_t)img * cb * bs + v * 8;
  float s = 0.f;
  for (int b = 0; b < cb; ++b) {
    float a[8];
    ld8(xs_in + base + b * bs, a);
    if (add16) {
      float y[8];
      load8_act(add16, base + b * bs, y, false);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] += y[i];
    }
    if (xs_out != xs_in || add16) st8(xs_out + base + b * bs, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
  }
  const float inv_c = 1.f / (float)(cb * 8);
  const float mean = s * inv_c;
  float q = 0.f;
  for (int b = 0; b < cb; ++b) {
    float a[8];
    ld8(xs_out + base + b * bs, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = a[i] - mean; q = fmaf(d, d, q); }
  }
  const float rstd = rsqrtf(q * inv_c + eps);
  if (stats) { stats[tok * 2] = mean; stats[tok * 2 + 1] = rstd; }
  if (!ln) return;
  for (int b = 0; b < cb; ++b) {
    float a[8];
    ld8(xs_out + base + b * bs, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = (a[i] - mean) * rstd;
      if (gamma) y = fmaf(y, gamma[b * 8 + i], beta ? beta[b * 8 + i] : 0.f);
      a[i] = y;
    }
    store8_act(ln, base + b * bs, 0, a, false);
  }
}

// dxs_out = (dxs_in | 0) + rstd * (g' - mean_C(g') - xhat * mean_C(g' xhat)),  g' = dy * gamma;  dxs16 = bf16 copy (the
// gradient operand of the Linear that produced the residual branch).  dy == NULL: pure pass-through of dxs_in (+ copy).
__global__ void __launch_bounds__(128)
swin_ln_bwd_kernel(const float* __restrict__ xs, const float* __restrict__ stats, const void* __restrict__ dy16,
                   const float* __restrict__ gamma, const float* __restrict__ dxs_in, float* __restrict__ dxs_out,
                   void* __restrict__ dxs16, int n_img, int cb, size_t nvox) {
  const size_t tok = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= (size_t)n_img * nvox) return;
  const int img = (int)(tok / nvox);
  const size_t v = tok - (size_t)img * nvox;
  const size_t bs = nvox * 8;
  const size_t base = (size_t)img * cb * bs + v * 8;
  float mean = 0.f, rstd = 0.f, m1 = 0.f, m2 = 0.f;
  if (dy16) {
    mean = stats[tok * 2];
    rstd = stats[tok * 2 + 1];
    for (int b = 0; b < cb; ++b) {
      float a[8], g[8];
      ld8(xs + base + b * bs, a);
      load8_act(dy16, base + b * bs, g, false);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float gp = gamma ? g[i] * gamma[b * 8 + i] : g[i];
        m1 += gp;
        m2 = fmaf(gp, (a[i] - mean) * rstd, m2);
      }
    }
    const float inv_c = 1.f / (float)(cb * 8);
    m1 *= inv_c;
    m2 *= inv_c;
  }
  for (int b = 0; b < cb; ++b) {
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = 0.f;
    if (dxs_in) ld8(dxs_in + base + b * bs, d);
    if (dy16) {
      float a[8], g[8];
      ld8(xs + base + b * bs, a);
      load8_act(dy16, base + b * bs, g, false);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float gp = gamma ? g[i] * gamma[b * 8 + i] : g[i];
        d[i] += rstd * (gp - m1 - (a[i] - mean) * rstd * m2);
      }
    }
    if (dxs_out) st8(dxs_out + base + b * bs, d);
    if (dxs16) store8_act(dxs16, base + b * bs, 0, d, false);
  }
}

// Warp-per-token variants for the deep stages (C >= 192, few tokens): lane l owns channel blocks l, l + 32, ...
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
swin_ln_fwd_train_warp_kernel(const float* __restrict__ xs_in, const void* __restrict__ add16, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float* __restrict__ xs_out, void* __restrict__ ln,
                              float* __restrict__ stats, int n_img, int cb, size_t nvox, float eps) {
  const size_t tok = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tok >= (size_t)n_img * nvox) return;
  const int img = (int)(tok / nvox);
  const size_t v = tok - (size_t)img * nvox;
  const size_t bs = nvox * 8;
  const size_t base = (size_t)img * cb * bs + v * 8;
  float s = 0.f;
  for (int b = lane; b < cb; b += 32) {
    float a[8];
    ld8(xs_in + base + b * bs, a);
    if (add16) {
      float y[8];
      load8_act(add16, base + b * bs, y, false);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] += y[i];
    }
    if (xs_out != xs_in || add16) st8(xs_out + base + b * bs, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
  }
  s = warp_sum(s);
  const float inv_c = 1.f / (float)(cb * 8);
  const float mean = s * inv_c;
  float q = 0.f;
  for (int b = lane; b < cb; b += 32) {
    float a[8];
    ld8(xs_out + base + b * bs, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = a[i] - mean; q = fmaf(d, d, q); }
  }
  q = warp_sum(q);
  const float rstd = rsqrtf(q * inv_c + eps);
  if (stats && lane == 0) { stats[tok * 2] = mean; stats[tok * 2 + 1] = rstd; }
  if (!ln) return;
  for (int b = lane; b < cb; b += 32) {
    float a[8];
    ld8(xs_out + base + b * bs, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = (a[i] - mean) * rstd;
      if (gamma) y = fmaf(y, gamma[b * 8 + i], beta ? beta[b * 8 + i] : 0.f);
      a[i] = y;
    }
    store8_act(ln, base + b * bs, 0, a, false);
  }
}

__global__ void __launch_bounds__(256)
swin_ln_bwd_warp_kernel(const float* __restrict__ xs, const float* __restrict__ stats, const void* __restrict__ dy16,
                        const float* __restrict__ gamma, const float* __restrict__ dxs_in, float* __restrict__ dxs_out,
                        void* __restrict__ dxs16, int n_img, int cb, size_t nvox) {
  const size_t tok = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tok >= (size_t)n_img * nvox) return;
  const int img = (int)(tok / nvox);
  const size_t v = tok - (size_t)img * nvox;
  const size_t bs = nvox * 8;
  const size_t base = (size_t)img * cb * bs + v * 8;
  float mean = 0.f, rstd = 0.f, m1 = 0.f, m2 = 0.f;
  if (dy16) {
    mean = stats[tok * 2];
    rstd = stats[tok * 2 + 1];
    for (int b = lane; b < cb; b += 32) {
      float a[8], g[8];
      ld8(xs + base + b * bs, a);
      load8_act(dy16, base + b * bs, g, false);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float gp = gamma ? g[i] * gamma[b * 8 + i] : g[i];
        m1 += gp;
        m2 = fmaf(gp, (a[i] - mean) * rstd, m2);
      }
    }
    const float inv_c = 1.f / (float)(cb * 8);
    m1 = warp_sum(m1) * inv_c;
    m2 = warp_sum(m2) * inv_c;
  }
  for (int b = lane; b < cb; b += 32) {
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = 0.f;
    if (dxs_in) ld8(dxs_in + base + b * bs, d);
    if (dy16) {
      float a[8], g[8];
      ld8(xs + base + b * bs, a);
      load8_act(dy16, base + b * bs, g, false);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float gp = gamma ? g[i] * gamma[b * 8 + i] : g[i];
        d[i] += rstd * (gp - m1 - (a[i] - mean) * rstd * m2);
      }
    }
    if (dxs_out) st8(dxs_out + base + b * bs, d);
    if (dxs16) store8_act(dxs16, base + b * bs, 0, d, false);
  }
}

// dgamma[c] = sum_tokens dy * xhat, dbeta[c] = sum_tokens dy: grid (chunks, cb); partial[chunk][C][2], summed by the caller
// in chunk order (deterministic).
__global__ void __launch_bounds__(256)
swin_ln_param_grad_kernel(const float* __restrict__ xs, const float* __restrict__ stats, const void* __restrict__ dy16,
                          float* __restrict__ partial, int n_img, int cb, size_t nvox) {
  __shared__ float sm[8][16];
  const int b = blockIdx.y;
  const size_t ntok = (size_t)n_img * nvox;
  const size_t bs = nvox * 8;
  float dg[8], db[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { dg[i] = 0.f; db[i] = 0.f; }
  for (size_t tok = (size_t)blockIdx.x * blockDim.x + threadIdx.x; tok < ntok; tok += (size_t)gridDim.x * blockDim.x) {
    const int img = (int)(tok / nvox);
    const size_t v = tok - (size_t)img * nvox;
    const size_t off = ((size_t)img * cb + b) * bs + v * 8;
    float a[8], g[8];
    ld8(xs + off, a);
    load8_act(dy16, off, g, false);
    const float mean = stats[tok * 2], rstd = stats[tok * 2 + 1];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dg[i] = fmaf(g[i], (a[i] - mean) * rstd, dg[i]); db[i] += g[i]; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dg[i] += __shfl_xor_sync(0xffffffffu, dg[i], o);
      db[i] += __shfl_xor_sync(0xffffffffu, db[i], o);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { sm[warp][i] = dg[i]; sm[warp][8 + i] = db[i]; }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += sm[w][threadIdx.x];
    const int i = threadIdx.x & 7, which = threadIdx.x >> 3;
    partial[((size_t)blockIdx.x * cb * 8 + b * 8 + i) * 2 + which] = s;
  }
}

// ------------------------------------------------------------------------------------------------ element-wise backward
// GELU (exact erf form): dx = dy * (Phi(x) + x * phi(x)), x = the saved pre-activation (bf16).
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const void* __restrict__ x16, const void* __restrict__ dy16,
                                                       void* __restrict__ dx16, size_t n8) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float x[8], g[8];
    load8_act(x16, i * 8, x, false);
    load8_act(dy16, i * 8, g, false);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float cdf = 0.5f * (1.f + erff(x[k] * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * __expf(-0.5f * x[k] * x[k]);
      g[k] *= cdf + x[k] * pdf;
    }
    store8_act(dx16, i * 8, 0, g, false);
  }
}

// UnetResBlock tail y = LeakyReLU(pre): g_pre = dy * (y > 0 ? 1 : slope)  (sign(y) == sign(pre) for slope > 0)
__global__ void __launch_bounds__(256) lrelu_mask_mul_kernel(const void* __restrict__ y16, const void* __restrict__ dy16,
                                                             void* __restrict__ out16, size_t n8, float slope) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float y[8], g[8];
    load8_act(y16, i * 8, y, false);
    load8_act(dy16, i * 8, g, false);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = y[k] > 0.f ? g[k] : g[k] * slope;
    store8_act(out16, i * 8, 0, g, false);
  }
}

// ------------------------------------------------------------------------------------------------ patch merging
__constant__ int kMergeOffB[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {0, 1, 0}, {0, 0, 1}, {1, 1, 1}};

// cat[img][slot*cb + b][zo][yo][xo] = xs[img][b][2zo+dz][2yo+dy][2xo+dx]   (fp32 blocked both sides)
__global__ void __launch_bounds__(256)
swin_merge_gather_kernel(const float* __restrict__ xs, float* __restrict__ cat, int n_img, int cb, int Z, int Y, int X) {
  const int Zo = Z / 2, Yo = Y / 2, Xo = X / 2;
  const size_t nout = (size_t)Zo * Yo * Xo, nvox = (size_t)Z * Y * X;
  const int blk = blockIdx.y;                       // img * 8cb + slot*cb + b
  const int img = blk / (8 * cb), sb = blk - img * 8 * cb;
  const int slot = sb / cb, b = sb - slot * cb;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nout; v += (size_t)gridDim.x * blockDim.x) {
    const int xo = (int)(v % Xo);
    const size_t r = v / Xo;
    const int yo = (int)(r % Yo), zo = (int)(r / Yo);
    const size_t src = (((size_t)(2 * zo + kMergeOffB[slot][0]) * Y + (2 * yo + kMergeOffB[slot][1])) * X + (2 * xo + kMergeOffB[slot][2]));
    float a[8];
    ld8(xs + (((size_t)img * cb + b) * nvox + src) * 8, a);
    st8(cat + ((size_t)blk * nout + v) * 8, a);
  }
}

// transpose of the gather: every source voxel is written once — octants (0,1,0) and (0,0,1) receive the sum of their two
// slots, octants (1,1,0) and (0,1,1) (never gathered by MONAI's legacy order) receive zero.
__global__ void __launch_bounds__(256)
swin_merge_scatter_kernel(const float* __restrict__ dcat, float* __restrict__ dxs, int n_img, int cb, int Z, int Y, int X) {
  const int Zo = Z / 2, Yo = Y / 2, Xo = X / 2;
  const size_t nout = (size_t)Zo * Yo * Xo, nvox = (size_t)Z * Y * X;
  const int blk = blockIdx.y;                       // img * cb + b
  const int img = blk / cb, b = blk - img * cb;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(v % X);
    const size_t r = v / X;
    const int y = (int)(r % Y), z = (int)(r / Y);
    const int oz = z & 1, oy = y & 1, ox = x & 1;
    const size_t vo = ((size_t)(z >> 1) * Yo + (y >> 1)) * Xo + (x >> 1);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int slot = 0; slot < 8; ++slot) {
      if (kMergeOffB[slot][0] == oz && kMergeOffB[slot][1] == oy && kMergeOffB[slot][2] == ox) {
        float a[8];
        ld8(dcat + ((((size_t)img * 8 + slot) * cb + b) * nout + vo) * 8, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += a[i];
      }
    }
    st8(dxs + ((size_t)blk * nvox + v) * 8, acc);
  }
}

// ------------------------------------------------------------------------------------------------ patch embedding wgrad
// dW[f][ci*8 + t] = sum_tokens dxs[f][token] * x[ci][2*token + t],  db[f] = sum_tokens dxs[f][token].
// grid = chunks of 64 tokens per iteration; partial[chunk][F][K + 1] (last column = bias), summed by the caller.
__global__ void __launch_bounds__(256)
swin_patch_embed_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dxs, float* __restrict__ partial,
                              int n_img, int Cin, int F, int Z, int Y, int X) {
  extern __shared__ float sm[];                 // dtile[64][F] | xtile[64][K + 1]
  const int K = Cin * 8, K1 = K + 1;
  float* dt = sm;
  float* xt = sm + 64 * F;
  const size_t nvox = (size_t)Z * Y * X, ntok = (size_t)n_img * nvox;
  const int n_el = F * K1;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};          // up to 4 (f, k) elements per thread (F*K1 <= 1024)
  for (size_t t0 = (size_t)blockIdx.x * 64; t0 < ntok; t0 += (size_t)gridDim.x * 64) {
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * F; e += blockDim.x) {
      const int tl = e % 64, f = e / 64;
      const size_t tok = t0 + tl;
      float v = 0.f;
      if (tok < ntok) {
        const int img = (int)(tok / nvox);
        const size_t vv = tok - (size_t)img * nvox;
        v = dxs[(((size_t)img * (F / 8) + (f >> 3)) * nvox + vv) * 8 + (f & 7)];
      }
      dt[tl * F + f] = v;
    }
    for (int e = threadIdx.x; e < 64 * K1; e += blockDim.x) {
      const int tl = e % 64, k = e / 64;
      const size_t tok = t0 + tl;
      float v = 0.f;
      if (tok < ntok) {
        if (k == K) {
          v = 1.f;
        } else {
          const int img = (int)(tok / nvox);
          const size_t vv = tok - (size_t)img * nvox;
          const int xo = (int)(vv % X);
          const size_t r = vv / X;
          const int yo = (int)(r % Y), zo = (int)(r / Y);
          const int ci = k >> 3, tp = k & 7;
          const int dz = tp >> 2, dy = (tp >> 1) & 1, dx = tp & 1;
          v = x[((size_t)img * Cin + ci) * nvox * 8 + ((size_t)(2 * zo + dz) * (2 * Y) + (2 * yo + dy)) * (2 * X) + (2 * xo + dx)];
        }
      }
      xt[tl * K1 + k] = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = threadIdx.x + j * 256;
      if (e < n_el) {
        const int f = e / K1, k = e - f * K1;
        float s = 0.f;
        for (int tl = 0; tl < 64; ++tl) s = fmaf(dt[tl * F + f], xt[tl * K1 + k], s);
        acc[j] += s;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int e = threadIdx.x + j * 256;
    if (e < n_el) partial[(size_t)blockIdx.x * n_el + e] = acc[j];
  }
}

// ------------------------------------------------------------------------------------------------ window attention backward
// One CTA = one (window, head, image), like the forward; head_dim = 16, so every score-shaped product is ONE
// mma.sync.m16n8k16 per 16 x 8 tile.  Probabilities are recomputed from q, k, the bias table and the saved log-sum-exp rows;
// delta_i = dO_i . O_i.
//   pass A (a warp owns 16 queries, walks the keys 16 at a time):  S = Q K^T and dP = dO V^T on the tensor cores, dS = P (dP -
//           delta) in the accumulator fragments -> bias-table gradient (shared-memory accumulation, one partial table per
//           CTA) and, re-used as the A fragments of a second MMA, dQ += dS K;
//   pass B (a warp owns 16 keys, walks the queries): S^T = K Q^T and dP^T = V dO^T, then dK += dS^T Q and dV += P^T dO.
// Padded tokens (MONAI pads AFTER norm1: their k / v ARE the qkv bias) contribute their dk / dv to the qkv-bias gradient;
// padded queries are cropped from the output, so their rows carry no gradient.
struct SwinAttnBwdK {
  const void* qkv;
  const void* out;        // forward output O (bf16)
  const void* dout;       // dO (bf16)
  const float* lse;       // [img][head][win][352] log-sum-exp rows in the log2 domain
  const float* table;
  const float* qkv_bias;
  void* dqkv;             // bf16 blocked, same shape as qkv
  float* dtable;          // [img][win][heads][table_len] partial
  float* dbias;           // [img][win][heads][32] partial: dk (16) | dv (16) of the padded tokens
  int n_img, D, H, W;
  int ws0, ws1, ws2, cw1, cw2, s0, s1, s2, Dp, Hp, Wp;
  int heads, qkv_cbt, out_cbt, out_cb_off, dout_cbt, dout_cb_off;
  int table_len, centre;
  float scale, scale_log2e;
};

constexpr int kBwdTok = 352;
constexpr int kBwdRS = 16;     // halfs per row of the row-major operand copies (2-way conflicts on fragment loads, but two CTAs fit per SM)
constexpr int kBwdTS = 360;    // halfs per row of the transposed copies
constexpr int kBwdThreads = 384;
constexpr size_t kBwdSmem = (size_t)4 * kBwdTok * kBwdRS * 2 + (size_t)3 * 16 * kBwdTS * 2 + 2 * 2200 * 4 + 32 + 2 * kBwdTok * 4 +
                            kBwdTok * 4 + kBwdTok * 2 + kBwdTok + 32 * 4 + 64;

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void load_afrag(const uint16_t* base, int row0, int g, int t, uint32_t (&a)[4]) {
  a[0] = *reinterpret_cast<const uint32_t*>(base + (row0 + g) * kBwdRS + 2 * t);
  a[1] = *reinterpret_cast<const uint32_t*>(base + (row0 + g + 8) * kBwdRS + 2 * t);
  a[2] = *reinterpret_cast<const uint32_t*>(base + (row0 + g) * kBwdRS + 2 * t + 8);
  a[3] = *reinterpret_cast<const uint32_t*>(base + (row0 + g + 8) * kBwdRS + 2 * t + 8);
}

__global__ void __launch_bounds__(kBwdThreads) swin_window_attention_bwd_kernel(const SwinAttnBwdK k) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint16_t* sQ = reinterpret_cast<uint16_t*>(smem);          // [352][24] bf16, row = token
  uint16_t* sK = sQ + kBwdTok * kBwdRS;
  uint16_t* sV = sK + kBwdTok * kBwdRS;
  uint16_t* sdO = sV + kBwdTok * kBwdRS;
  uint16_t* sKt = sdO + kBwdTok * kBwdRS;                    // [16][360] bf16, row = dim
  uint16_t* sQt = sKt + 16 * kBwdTS;
  uint16_t* sdOt = sQt + 16 * kBwdTS;
  float* sTab = reinterpret_cast<float*>(sdOt + 16 * kBwdTS);  // bias (log2 domain), 2200
  // gradient accumulation in int32 fixed point: 32-bit integer shared-memory atomics are native (fp32 / 64-bit ones are CAS
  // loops) and exact, so the bias-table gradient is bitwise reproducible whatever the order the warps arrive in.  The scale
  // comes from a per-CTA bound: |dS_ij| <= |dO_i| |v_j| + |delta_i|, at most 343 addends per table entry.
  int* sdTab = reinterpret_cast<int*>(sTab + 2200);            // 2200 x int32
  unsigned* sMax = reinterpret_cast<unsigned*>(sdTab + 2200);  // max |dO_i|^2, max |v_j|^2, max |delta_i| (float bits), scale
  float* sLse = reinterpret_cast<float*>(sdTab + 2200 + 8);    // [352]
  float* sDelta = sLse + kBwdTok;                              // [352]
  int* sPos = reinterpret_cast<int*>(sDelta + kBwdTok);        // [352]
  int16_t* sBase = reinterpret_cast<int16_t*>(sPos + kBwdTok);
  uint8_t* sReg = reinterpret_cast<uint8_t*>(sBase + kBwdTok);
  float* sPad = reinterpret_cast<float*>(smem + kBwdSmem - 32 * 4 - 16);   // dk | dv of the padded tokens

  const int n_tok = k.ws0 * k.ws1 * k.ws2;
  const int np = (n_tok + 15) & ~15;
  const int nW1 = k.Hp / k.ws1, nW2 = k.Wp / k.ws2;
  const int win = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
  const int w2 = win % nW2, w1 = (win / nW2) % nW1, w0 = win / (nW2 * nW1);
  const size_t nvox = (size_t)k.D * k.H * k.W;
  const int C = k.heads * 16;
  const bool shifted = (k.s0 | k.s1 | k.s2) != 0;
  const int tid = threadIdx.x;

  for (int i = tid; i < k.table_len; i += blockDim.x) {
    sTab[i] = k.table[(size_t)i * k.heads + head] * 1.4426950408889634f;
    sdTab[i] = 0;
  }
  if (tid < 32) sPad[tid] = 0.f;
  if (tid < 8) sMax[tid] = 0u;
  for (int i = tid; i < kBwdTok; i += blockDim.x) {
    int pos = -2, base = 0, reg = 0;
    if (i < n_tok) {
      pos = -1;
      const int t2 = i % k.ws2, t1 = (i / k.ws2) % k.ws1, t0 = i / (k.ws2 * k.ws1);
      const int g0 = w0 * k.ws0 + t0, g1 = w1 * k.ws1 + t1, g2 = w2 * k.ws2 + t2;
      int p0 = g0 + k.s0, p1 = g1 + k.s1, p2 = g2 + k.s2;
      if (p0 >= k.Dp) p0 -= k.Dp;
      if (p1 >= k.Hp) p1 -= k.Hp;
      if (p2 >= k.Wp) p2 -= k.Wp;
      if (p0 < k.D && p1 < k.H && p2 < k.W) pos = (p0 * k.H + p1) * k.W + p2;
      const int c2 = i % k.cw2, c1 = (i / k.cw2) % k.cw1, c0 = i / (k.cw2 * k.cw1);
      base = (c0 * (2 * k.cw1 - 1) + c1) * (2 * k.cw2 - 1) + c2;
      if (shifted) {
        const int r0 = k.s0 == 0 ? 0 : (g0 < k.Dp - k.ws0 ? 0 : (g0 < k.Dp - k.s0 ? 1 : 2));
        const int r1 = k.s1 == 0 ? 0 : (g1 < k.Hp - k.ws1 ? 0 : (g1 < k.Hp - k.s1 ? 1 : 2));
        const int r2 = k.s2 == 0 ? 0 : (g2 < k.Wp - k.ws2 ? 0 : (g2 < k.Wp - k.s2 ? 1 : 2));
        reg = (r0 * 3 + r1) * 3 + r2;
      }
    }
    sPos[i] = pos;
    sBase[i] = (int16_t)base;
    sReg[i] = (uint8_t)reg;
  }
  __syncthreads();
  const uint16_t* qkv = reinterpret_cast<const uint16_t*>(k.qkv);
  // q, k, v, dO rows (bf16) and the transposed copies the second-stage MMAs read as B operands
  for (int e = tid; e < np * 8; e += blockDim.x) {
    const int i = e >> 3, part = e & 7;            // part: 0,1 q | 2,3 k | 4,5 v | 6,7 dO
    const int which = part >> 1, half = part & 1;
    const int pos = sPos[i];
    uint4 u = make_uint4(0, 0, 0, 0);
    if (which < 3) {
      const int ch0 = which * C + head * 16 + half * 8;
      if (pos >= 0) {
        u = *reinterpret_cast<const uint4*>(qkv + (((size_t)img * k.qkv_cbt + (ch0 >> 3)) * nvox + pos) * 8);
      } else if (pos == -1 && k.qkv_bias) {
        float b[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = k.qkv_bias[ch0 + j];
        u = cvt8_from_f32(b, false);
      }
    } else if (pos >= 0) {
      u = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(k.dout) +
                                          (((size_t)img * k.dout_cbt + k.dout_cb_off + head * 2 + half) * nvox + pos) * 8);
    }
    uint16_t* rowp = (which == 0 ? sQ : which == 1 ? sK : which == 2 ? sV : sdO) + i * kBwdRS + half * 8;
    *reinterpret_cast<uint4*>(rowp) = u;
    if (which != 2) {
      uint16_t* tp = (which == 0 ? sQt : which == 1 ? sKt : sdOt);
      const uint16_t* h = reinterpret_cast<const uint16_t*>(&u);
#pragma unroll
      for (int j = 0; j < 8; ++j) tp[(half * 8 + j) * kBwdTS + i] = h[j];
    }
  }
  const size_t row0 = (((size_t)img * k.heads + head) * gridDim.x + win) * kBwdTok;
  for (int i = tid; i < np; i += blockDim.x) sLse[i] = i < n_tok ? k.lse[row0 + i] : 0.f;
  __syncthreads();
  for (int i = tid; i < np; i += blockDim.x) {
    const int pos = sPos[i];
    float d = 0.f;
    if (pos >= 0) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float o[8], g8[8];
        load8_act(k.out, (((size_t)img * k.out_cbt + k.out_cb_off + head * 2 + half) * nvox + pos) * 8, o, false);
        cvt8_to_f32(*reinterpret_cast<const uint4*>(sdO + i * kBwdRS + half * 8), g8, false);
#pragma unroll
        for (int j = 0; j < 8; ++j) d = fmaf(o[j], g8[j], d);
      }
    }
    sDelta[i] = d;
    float g16[16], v16[16];
    cvt8_to_f32(*reinterpret_cast<const uint4*>(sdO + i * kBwdRS), g16, false);
    cvt8_to_f32(*reinterpret_cast<const uint4*>(sdO + i * kBwdRS + 8), g16 + 8, false);
    cvt8_to_f32(*reinterpret_cast<const uint4*>(sV + i * kBwdRS), v16, false);
    cvt8_to_f32(*reinterpret_cast<const uint4*>(sV + i * kBwdRS + 8), v16 + 8, false);
    float ng = 0.f, nv = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { ng = fmaf(g16[j], g16[j], ng); nv = fmaf(v16[j], v16[j], nv); }
    atomicMax(&sMax[0], __float_as_uint(ng));      // non-negative floats order like their bit patterns
    atomicMax(&sMax[1], __float_as_uint(nv));
    atomicMax(&sMax[2], __float_as_uint(fabsf(d)));
  }
  __syncthreads();
  if (tid == 0) {
    const float bound = sqrtf(__uint_as_float(sMax[0])) * sqrtf(__uint_as_float(sMax[1])) + __uint_as_float(sMax[2]);
    reinterpret_cast<float*>(sMax)[3] = bound > 0.f ? 1073741824.f / (343.f * bound) : 0.f;   // 2^30 / (addends * bound)
  }
  __syncthreads();
  const float fix = reinterpret_cast<const float*>(sMax)[3];

  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n_warps = blockDim.x >> 5;
  const int n_tiles = np >> 4;
  uint16_t* dqkv = reinterpret_cast<uint16_t*>(k.dqkv);

  // ---- pass A: 16 queries per warp -> dq, bias-table gradient
  for (int qt = warp; qt < n_tiles; qt += n_warps) {
    const int q0 = qt * 16;
    uint32_t qa[4], ga[4];
    load_afrag(sQ, q0, g, t, qa);
    load_afrag(sdO, q0, g, t, ga);
    const int i0 = q0 + g, i1 = q0 + g + 8;
    const bool v0 = sPos[i0] >= 0, v1 = sPos[i1] >= 0;
    const float lse0 = sLse[i0], lse1 = sLse[i1], de0 = sDelta[i0], de1 = sDelta[i1];
    const int bq0 = sBase[i0] + k.centre, bq1 = sBase[i1] + k.centre;
    const int rq0 = sReg[i0], rq1 = sReg[i1];
    float dq[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) dq[a][b] = 0.f;
    for (int key0 = 0; key0 < np; key0 += 16) {
      float ds[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int kr = key0 + nt * 8 + g;           // B-fragment row of this lane
        float sc[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        mma_bf16(sc, qa, *reinterpret_cast<const uint32_t*>(sK + kr * kBwdRS + 2 * t),
                 *reinterpret_cast<const uint32_t*>(sK + kr * kBwdRS + 2 * t + 8));
        mma_bf16(dp, ga, *reinterpret_cast<const uint32_t*>(sV + kr * kBwdRS + 2 * t),
                 *reinterpret_cast<const uint32_t*>(sV + kr * kBwdRS + 2 * t + 8));
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int key = key0 + nt * 8 + 2 * t + j;
          float d0 = 0.f, d1 = 0.f;
          if (key < n_tok) {
            const int bk = sBase[key], rk = sReg[key];
            if (v0) {
              float sl = fmaf(sc[j], k.scale_log2e, sTab[bq0 - bk]);
              if (shifted && rk != rq0) sl -= 144.26950408889634f;
              d0 = exp2f(sl - lse0) * (dp[j] - de0);
              atomicAdd(&sdTab[bq0 - bk], __float2int_rn(d0 * fix));
            }
            if (v1) {
              float sl = fmaf(sc[2 + j], k.scale_log2e, sTab[bq1 - bk]);
              if (shifted && rk != rq1) sl -= 144.26950408889634f;
              d1 = exp2f(sl - lse1) * (dp[2 + j] - de1);
              atomicAdd(&sdTab[bq1 - bk], __float2int_rn(d1 * fix));
            }
          }
          ds[nt][j] = d0;
          ds[nt][2 + j] = d1;
        }
      }
      uint32_t pa[4] = {pack_bf16(ds[0][0], ds[0][1]), pack_bf16(ds[0][2], ds[0][3]), pack_bf16(ds[1][0], ds[1][1]),
                        pack_bf16(ds[1][2], ds[1][3])};
#pragma unroll
      for (int nd = 0; nd < 2; ++nd)
        mma_bf16(dq[nd], pa, *reinterpret_cast<const uint32_t*>(sKt + (nd * 8 + g) * kBwdTS + key0 + 2 * t),
                 *reinterpret_cast<const uint32_t*>(sKt + (nd * 8 + g) * kBwdTS + key0 + 2 * t + 8));
    }
    const int p0 = sPos[i0], p1 = sPos[i1];
#pragma unroll
    for (int nd = 0; nd < 2; ++nd) {
      const size_t cbase = ((size_t)img * k.qkv_cbt + head * 2 + nd) * nvox;
      if (p0 >= 0) *reinterpret_cast<uint32_t*>(dqkv + (cbase + p0) * 8 + 2 * t) = pack_bf16(dq[nd][0] * k.scale, dq[nd][1] * k.scale);
      if (p1 >= 0) *reinterpret_cast<uint32_t*>(dqkv + (cbase + p1) * 8 + 2 * t) = pack_bf16(dq[nd][2] * k.scale, dq[nd][3] * k.scale);
    }
  }

  // ---- pass B: 16 keys per warp -> dk, dv
  for (int kt = warp; kt < n_tiles; kt += n_warps) {
    const int k0 = kt * 16;
    uint32_t ka[4], va[4];
    load_afrag(sK, k0, g, t, ka);
    load_afrag(sV, k0, g, t, va);
    const int j0 = k0 + g, j1 = k0 + g + 8;
    const bool kv0 = j0 < n_tok, kv1 = j1 < n_tok;
    const int bk0 = sBase[j0], bk1 = sBase[j1], rk0 = sReg[j0], rk1 = sReg[j1];
    float dk[2][4], dv[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) { dk[a][b] = 0.f; dv[a][b] = 0.f; }
    for (int q0 = 0; q0 < np; q0 += 16) {
      float pT[2][4], dsT[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int qr = q0 + nt * 8 + g;
        float sc[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        mma_bf16(sc, ka, *reinterpret_cast<const uint32_t*>(sQ + qr * kBwdRS + 2 * t),
                 *reinterpret_cast<const uint32_t*>(sQ + qr * kBwdRS + 2 * t + 8));
        mma_bf16(dp, va, *reinterpret_cast<const uint32_t*>(sdO + qr * kBwdRS + 2 * t),
                 *reinterpret_cast<const uint32_t*>(sdO + qr * kBwdRS + 2 * t + 8));
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int qi = q0 + nt * 8 + 2 * t + j;    // column = query
          float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
          if (sPos[qi] >= 0) {
            const int bq = sBase[qi] + k.centre, rq = sReg[qi];
            const float lse = sLse[qi], de = sDelta[qi];
            if (kv0) {
              float sl = fmaf(sc[j], k.scale_log2e, sTab[bq - bk0]);
              if (shifted && rq != rk0) sl -= 144.26950408889634f;
              p0 = exp2f(sl - lse);
              d0 = p0 * (dp[j] - de);
            }
            if (kv1) {
              float sl = fmaf(sc[2 + j], k.scale_log2e, sTab[bq - bk1]);
              if (shifted && rq != rk1) sl -= 144.26950408889634f;
              p1 = exp2f(sl - lse);
              d1 = p1 * (dp[2 + j] - de);
            }
          }
          pT[nt][j] = p0; pT[nt][2 + j] = p1;
          dsT[nt][j] = d0; dsT[nt][2 + j] = d1;
        }
      }
      uint32_t pa[4] = {pack_bf16(pT[0][0], pT[0][1]), pack_bf16(pT[0][2], pT[0][3]), pack_bf16(pT[1][0], pT[1][1]),
                        pack_bf16(pT[1][2], pT[1][3])};
      uint32_t da[4] = {pack_bf16(dsT[0][0], dsT[0][1]), pack_bf16(dsT[0][2], dsT[0][3]), pack_bf16(dsT[1][0], dsT[1][1]),
                        pack_bf16(dsT[1][2], dsT[1][3])};
#pragma unroll
      for (int nd = 0; nd < 2; ++nd) {
        mma_bf16(dk[nd], da, *reinterpret_cast<const uint32_t*>(sQt + (nd * 8 + g) * kBwdTS + q0 + 2 * t),
                 *reinterpret_cast<const uint32_t*>(sQt + (nd * 8 + g) * kBwdTS + q0 + 2 * t + 8));
        mma_bf16(dv[nd], pa, *reinterpret_cast<const uint32_t*>(sdOt + (nd * 8 + g) * kBwdTS + q0 + 2 * t),
                 *reinterpret_cast<const uint32_t*>(sdOt + (nd * 8 + g) * kBwdTS + q0 + 2 * t + 8));
      }
    }
    const int p0 = kv0 ? sPos[j0] : -2, p1 = kv1 ? sPos[j1] : -2;
#pragma unroll
    for (int nd = 0; nd < 2; ++nd) {
      const size_t kb = ((size_t)img * k.qkv_cbt + ((C + head * 16) >> 3) + nd) * nvox;
      const size_t vb = ((size_t)img * k.qkv_cbt + ((2 * C + head * 16) >> 3) + nd) * nvox;
      if (p0 >= 0) {
        *reinterpret_cast<uint32_t*>(dqkv + (kb + p0) * 8 + 2 * t) = pack_bf16(dk[nd][0] * k.scale, dk[nd][1] * k.scale);
        *reinterpret_cast<uint32_t*>(dqkv + (vb + p0) * 8 + 2 * t) = pack_bf16(dv[nd][0], dv[nd][1]);
      } else if (p0 == -1) {
        atomicAdd(&sPad[nd * 8 + 2 * t], dk[nd][0] * k.scale); atomicAdd(&sPad[nd * 8 + 2 * t + 1], dk[nd][1] * k.scale);
        atomicAdd(&sPad[16 + nd * 8 + 2 * t], dv[nd][0]);      atomicAdd(&sPad[16 + nd * 8 + 2 * t + 1], dv[nd][1]);
      }
      if (p1 >= 0) {
        *reinterpret_cast<uint32_t*>(dqkv + (kb + p1) * 8 + 2 * t) = pack_bf16(dk[nd][2] * k.scale, dk[nd][3] * k.scale);
        *reinterpret_cast<uint32_t*>(dqkv + (vb + p1) * 8 + 2 * t) = pack_bf16(dv[nd][2], dv[nd][3]);
      } else if (p1 == -1) {
        atomicAdd(&sPad[nd * 8 + 2 * t], dk[nd][2] * k.scale); atomicAdd(&sPad[nd * 8 + 2 * t + 1], dk[nd][3] * k.scale);
        atomicAdd(&sPad[16 + nd * 8 + 2 * t], dv[nd][2]);      atomicAdd(&sPad[16 + nd * 8 + 2 * t + 1], dv[nd][3]);
      }
    }
  }
  __syncthreads();
  const size_t cta = ((size_t)img * gridDim.x + win) * k.heads + head;
  if (tid < 32) k.dbias[cta * 32 + tid] = sPad[tid];
  const float unfix = fix > 0.f ? 1.f / fix : 0.f;
  for (int i = tid; i < k.table_len; i += blockDim.x) k.dtable[cta * k.table_len + i] = (float)sdTab[i] * unfix;
}

static int sm_count_b() {
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

extern "C" int mmseg_swin_ln_fwd_train(const float* xs_in, const void* add16, const float* gamma, const float* beta,
                                       float* xs_out, void* ln16, float* stats, int32_t n_img, int32_t cb, int64_t voxels,
                                       float eps, void* stream) {
  if (!xs_in || !xs_out || n_img < 1 || cb < 1 || voxels < 1 || (beta && !gamma))
    return fail(MMSEG_ERR_INVALID_ARG, "swin_ln_fwd_train: bad arguments");
  const size_t tok = (size_t)n_img * voxels;
  if (cb >= 24 || tok < 4096)
    swin_ln_fwd_train_warp_kernel<<<(unsigned)((tok + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        xs_in, add16, gamma, beta, xs_out, ln16, stats, n_img, cb, (size_t)voxels, eps);
  else
    swin_ln_fwd_train_kernel<<<(unsigned)((tok + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        xs_in, add16, gamma, beta, xs_out, ln16, stats, n_img, cb, (size_t)voxels, eps);
  return check_launch("swin_ln_fwd_train_kernel");
}

extern "C" int mmseg_swin_ln_bwd(const float* xs, const float* stats, const void* dy16, const float* gamma,
                                 const float* dxs_in, float* dxs_out, void* dxs16, int32_t n_img, int32_t cb, int64_t voxels,
                                 void* stream) {
  if ((!dxs_out && !dxs16) || n_img < 1 || cb < 1 || voxels < 1 || (dy16 && (!xs || !stats)) || (!dy16 && !dxs_in))
    return fail(MMSEG_ERR_INVALID_ARG, "swin_ln_bwd: bad arguments");
  const size_t tok = (size_t)n_img * voxels;
  if (cb >= 24 || tok < 4096)
    swin_ln_bwd_warp_kernel<<<(unsigned)((tok + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        xs, stats, dy16, gamma, dxs_in, dxs_out, dxs16, n_img, cb, (size_t)voxels);
  else
    swin_ln_bwd_kernel<<<(unsigned)((tok + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        xs, stats, dy16, gamma, dxs_in, dxs_out, dxs16, n_img, cb, (size_t)voxels);
  return check_launch("swin_ln_bwd_kernel");
}

extern "C" int mmseg_swin_ln_param_grad(const float* xs, const float* stats, const void* dy16, float* partial,
                                        int32_t n_chunks, int32_t n_img, int32_t cb, int64_t voxels, void* stream) {
  if (!xs || !stats || !dy16 || !partial || n_chunks < 1 || n_img < 1 || cb < 1 || voxels < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "swin_ln_param_grad: bad arguments");
  dim3 grid((unsigned)n_chunks, (unsigned)cb);
  swin_ln_param_grad_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(xs, stats, dy16, partial, n_img, cb,
                                                                                      (size_t)voxels);
  return check_launch("swin_ln_param_grad_kernel");
}

extern "C" int mmseg_gelu_bwd(const void* x16, const void* dy16, void* dx16, int64_t n_elems, void* stream) {
  if (!x16 || !dy16 || !dx16 || n_elems < 8 || (n_elems & 7)) return fail(MMSEG_ERR_INVALID_ARG, "gelu_bwd: bad arguments");
  const size_t n8 = (size_t)n_elems / 8;
  size_t g = (n8 + 255) / 256, cap = (size_t)sm_count_b() * 16;
  if (g > cap) g = cap;
  gelu_bwd_kernel<<<(unsigned)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x16, dy16, dx16, n8);
  return check_launch("gelu_bwd_kernel");
}

extern "C" int mmseg_lrelu_mask_mul(const void* y16, const void* dy16, void* out16, int64_t n_elems, float slope,
                                    void* stream) {
  if (!y16 || !dy16 || !out16 || n_elems < 8 || (n_elems & 7)) return fail(MMSEG_ERR_INVALID_ARG, "lrelu_mask_mul: bad arguments");
  const size_t n8 = (size_t)n_elems / 8;
  size_t g = (n8 + 255) / 256, cap = (size_t)sm_count_b() * 16;
  if (g > cap) g = cap;
  lrelu_mask_mul_kernel<<<(unsigned)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(y16, dy16, out16, n8, slope);
  return check_launch("lrelu_mask_mul_kernel");
}

extern "C" int mmseg_swin_merge_gather(const float* xs, float* cat, int32_t n_img, int32_t cb, int32_t Z, int32_t Y, int32_t X,
                                       void* stream) {
  if (!xs || !cat || n_img < 1 || cb < 1 || Z < 2 || Y < 2 || X < 2 || ((Z | Y | X) & 1))
    return fail(MMSEG_ERR_INVALID_ARG, "swin_merge_gather: bad arguments (even extents)");
  const size_t nout = (size_t)(Z / 2) * (Y / 2) * (X / 2);
  size_t gx = (nout + 255) / 256;
  if (gx > 1024) gx = 1024;
  dim3 grid((unsigned)gx, (unsigned)(n_img * 8 * cb));
  swin_merge_gather_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(xs, cat, n_img, cb, Z, Y, X);
  return check_launch("swin_merge_gather_kernel");
}

extern "C" int mmseg_swin_merge_scatter(const float* dcat, float* dxs, int32_t n_img, int32_t cb, int32_t Z, int32_t Y,
                                        int32_t X, void* stream) {
  if (!dcat || !dxs || n_img < 1 || cb < 1 || Z < 2 || Y < 2 || X < 2 || ((Z | Y | X) & 1))
    return fail(MMSEG_ERR_INVALID_ARG, "swin_merge_scatter: bad arguments (even extents)");
  const size_t nvox = (size_t)Z * Y * X;
  size_t gx = (nvox + 255) / 256;
  if (gx > 2048) gx = 2048;
  dim3 grid((unsigned)gx, (unsigned)(n_img * cb));
  swin_merge_scatter_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dcat, dxs, n_img, cb, Z, Y, X);
  return check_launch("swin_merge_scatter_kernel");
}

extern "C" int mmseg_swin_patch_embed_wgrad(const float* x, const float* dxs, float* partial, int32_t n_chunks, int32_t n_img,
                                            int32_t Cin, int32_t F, int32_t Z, int32_t Y, int32_t X, void* stream) {
  if (!x || !dxs || !partial || n_chunks < 1 || n_img < 1 || Cin < 1 || Cin > 8 || F < 8 || (F % 8) ||
      F * (Cin * 8 + 1) > 1024)
    return fail(MMSEG_ERR_INVALID_ARG, "swin_patch_embed_wgrad: bad arguments (F * (8 Cin + 1) <= 1024)");
  const size_t smem = (size_t)64 * (F + Cin * 8 + 1) * sizeof(float);
  swin_patch_embed_wgrad_kernel<<<(unsigned)n_chunks, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, dxs, partial, n_img, Cin, F, Z, Y, X);
  return check_launch("swin_patch_embed_wgrad_kernel");
}

extern "C" int mmseg_swin_window_attention_bwd(const mmseg_swin_attn_args* a, const void* dout, int32_t dout_cbt,
                                               int32_t dout_cb_off, const float* lse, void* dqkv, float* dtable,
                                               float* dbias, void* stream) {
  if (!a || !a->qkv || !a->out || !a->table || !dout || !lse || !dqkv || !dtable || !dbias)
    return fail(MMSEG_ERR_INVALID_ARG, "swin_window_attention_bwd: null pointer");
  if (a->head_dim != 16) return fail(MMSEG_ERR_UNSUPPORTED, "swin_window_attention_bwd: head_dim %d (built for 16)", a->head_dim);
  if (a->elem_fmt != MMSEG_FMT_BF16) return fail(MMSEG_ERR_UNSUPPORTED, "swin_window_attention_bwd: bf16 only");
  SwinAttnBwdK k;
  k.qkv = a->qkv; k.out = a->out; k.dout = dout; k.lse = lse; k.table = a->table; k.qkv_bias = a->qkv_bias;
  k.dqkv = dqkv; k.dtable = dtable; k.dbias = dbias;
  k.n_img = a->n_img; k.D = a->D; k.H = a->H; k.W = a->W;
  const int ext[3] = {a->D, a->H, a->W};
  int ws[3], ss[3], pp[3];
  for (int i = 0; i < 3; ++i) {
    if (a->window[i] < 1 || a->shift[i] < 0 || a->shift[i] >= a->window[i])
      return fail(MMSEG_ERR_INVALID_ARG, "swin_window_attention_bwd: window / shift");
    ws[i] = ext[i] <= a->window[i] ? ext[i] : a->window[i];
    ss[i] = ext[i] <= a->window[i] ? 0 : a->shift[i];
    pp[i] = (ext[i] + ws[i] - 1) / ws[i] * ws[i];
  }
  if (ws[0] * ws[1] * ws[2] > 343) return fail(MMSEG_ERR_UNSUPPORTED, "swin_window_attention_bwd: more than 343 tokens per window");
  k.ws0 = ws[0]; k.ws1 = ws[1]; k.ws2 = ws[2];
  k.s0 = ss[0]; k.s1 = ss[1]; k.s2 = ss[2];
  k.Dp = pp[0]; k.Hp = pp[1]; k.Wp = pp[2];
  k.cw1 = a->window[1]; k.cw2 = a->window[2];
  k.heads = a->heads; k.qkv_cbt = a->qkv_cbt; k.out_cbt = a->out_cbt; k.out_cb_off = a->out_cb_off;
  k.dout_cbt = dout_cbt; k.dout_cb_off = dout_cb_off;
  k.table_len = (2 * a->window[0] - 1) * (2 * a->window[1] - 1) * (2 * a->window[2] - 1);
  if (k.table_len > 2200) return fail(MMSEG_ERR_UNSUPPORTED, "swin_window_attention_bwd: bias table of %d rows", k.table_len);
  k.centre = ((a->window[0] - 1) * (2 * a->window[1] - 1) + (a->window[1] - 1)) * (2 * a->window[2] - 1) + (a->window[2] - 1);
  k.scale = a->scale;
  k.scale_log2e = a->scale * 1.4426950408889634f;
  const size_t smem = kBwdSmem;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(swin_window_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  dim3 grid((unsigned)((pp[0] / ws[0]) * (pp[1] / ws[1]) * (pp[2] / ws[2])), (unsigned)a->heads, (unsigned)a->n_img);
  swin_window_attention_bwd_kernel<<<grid, kBwdThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(k);
  return check_launch("swin_window_attention_bwd_kernel");
}
