// Backward (training) kernels of the SwinUNETR path — the ops autograd would issue for loss.backward()
// (reference src/trainer/trainer.py:243) through monai.networks.nets.SwinUNETR (src/models/backbones/swin_unetr.py:80-117):
// token LayerNorm forward-with-statistics / backward / parameter gradients, GELU backward, the LeakyReLU mask of the
// UnetResBlock tail, patch-merging gather / scatter, patch-embedding weight gradient and the window-attention backward.
// bf16 operands, fp32 residual-stream gradients.  Every buffer is written by exactly one thread (no global float atomics);
// the relative-position-bias gradient is accumulated per CTA in shared memory and reduced across CTAs in a fixed order.
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
// One CTA = one (window, head, image), like the forward.  Probabilities are recomputed from q, k, the bias table and the
// saved log-sum-exp rows; delta_i = dO_i . O_i.  Pass A (one query per thread): dq_i and the bias-table gradient
// (shared-memory accumulation, one partial table per CTA); pass B (one key per thread): dk_j, dv_j.  Padded tokens
// contribute their dk / dv to the qkv-bias gradient (their k / v ARE the bias).
struct SwinAttnBwdK {
  const void* qkv;
  const void* out;        // forward output O (bf16)
  const void* dout;       // dO (bf16)
  const float* lse;       // [img][head][win][352] natural-log-sum-exp in the log2 domain
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

__global__ void __launch_bounds__(384) swin_window_attention_bwd_kernel(const SwinAttnBwdK k) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* sQ = reinterpret_cast<float*>(smem);            // [352][16]
  float* sK = sQ + kBwdTok * 16;
  float* sV = sK + kBwdTok * 16;
  float* sdO = sV + kBwdTok * 16;
  float* sTab = sdO + kBwdTok * 16;                      // bias (log2 domain), 2200
  float* sdTab = sTab + 2200;                            // gradient accumulation, 2200
  float* sLse = sdTab + 2200;                            // [352]
  float* sDelta = sLse + kBwdTok;                        // [352]
  int* sPos = reinterpret_cast<int*>(sDelta + kBwdTok);  // [352]
  int16_t* sBase = reinterpret_cast<int16_t*>(sPos + kBwdTok);
  uint8_t* sReg = reinterpret_cast<uint8_t*>(sBase + kBwdTok);

  const int n_tok = k.ws0 * k.ws1 * k.ws2;
  const int nW1 = k.Hp / k.ws1, nW2 = k.Wp / k.ws2;
  const int win = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
  const int w2 = win % nW2, w1 = (win / nW2) % nW1, w0 = win / (nW2 * nW1);
  const size_t nvox = (size_t)k.D * k.H * k.W;
  const int C = k.heads * 16;
  const bool shifted = (k.s0 | k.s1 | k.s2) != 0;
  const int tid = threadIdx.x;

  for (int i = tid; i < k.table_len; i += blockDim.x) {
    sTab[i] = k.table[(size_t)i * k.heads + head] * 1.4426950408889634f;
    sdTab[i] = 0.f;
  }
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
  // q, k, v, dO rows as fp32 in shared memory; delta = dO . O
  for (int e = tid; e < n_tok * 8; e += blockDim.x) {
    const int i = e >> 3, part = e & 7;            // part: 0,1 q | 2,3 k | 4,5 v | 6,7 dO
    const int which = part >> 1, half = part & 1;
    const int pos = sPos[i];
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (which < 3) {
      const int ch0 = which * C + head * 16 + half * 8;
      if (pos >= 0) {
        load8_act(qkv, (((size_t)img * k.qkv_cbt + (ch0 >> 3)) * nvox + pos) * 8, v, false);
      } else if (k.qkv_bias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(k.qkv_bias[ch0 + j]));
      }
    } else if (pos >= 0) {
      load8_act(k.dout, (((size_t)img * k.dout_cbt + k.dout_cb_off + head * 2 + half) * nvox + pos) * 8, v, false);
    }
    float* dstp = (which == 0 ? sQ : which == 1 ? sK : which == 2 ? sV : sdO) + i * 16 + half * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) dstp[j] = v[j];
  }
  const size_t row0 = (((size_t)img * k.heads + head) * gridDim.x + win) * kBwdTok;
  for (int i = tid; i < n_tok; i += blockDim.x) sLse[i] = k.lse[row0 + i];
  __syncthreads();
  for (int i = tid; i < n_tok; i += blockDim.x) {
    const int pos = sPos[i];
    float d = 0.f;
    if (pos >= 0) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float o[8];
        load8_act(k.out, (((size_t)img * k.out_cbt + k.out_cb_off + head * 2 + half) * nvox + pos) * 8, o, false);
#pragma unroll
        for (int j = 0; j < 8; ++j) d = fmaf(o[j], sdO[i * 16 + half * 8 + j], d);
      }
    }
    sDelta[i] = d;
  }
  __syncthreads();

  uint16_t* dqkv = reinterpret_cast<uint16_t*>(k.dqkv);
  // ---- pass A: one query row per thread -> dq, bias-table gradient
  for (int i = tid; i < n_tok; i += blockDim.x) {
    const int pos = sPos[i];
    if (pos < 0) continue;                         // a padded query's output is cropped: no gradient flows through its row
    float q[16], dO[16], dq[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { q[j] = sQ[i * 16 + j]; dO[j] = sdO[i * 16 + j]; dq[j] = 0.f; }
    const float lse = sLse[i], delta = sDelta[i];
    const int bq = sBase[i] + k.centre, rq = sReg[i];
    for (int j = 0; j < n_tok; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) { s = fmaf(q[c], sK[j * 16 + c], s); dp = fmaf(dO[c], sV[j * 16 + c], dp); }
      const int ti = bq - sBase[j];
      float sl = fmaf(s, k.scale_log2e, sTab[ti]);
      if (shifted && sReg[j] != rq) sl -= 144.26950408889634f;
      const float p = exp2f(sl - lse);
      const float ds = p * (dp - delta);
      atomicAdd(&sdTab[ti], ds);
#pragma unroll
      for (int c = 0; c < 16; ++c) dq[c] = fmaf(ds, sK[j * 16 + c], dq[c]);
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) dq[c] *= k.scale;
#pragma unroll
    for (int half = 0; half < 2; ++half)
      store8_act(dqkv, (((size_t)img * k.qkv_cbt + ((head * 16 + half * 8) >> 3)) * nvox + pos) * 8, 0, dq + half * 8, false);
  }
  // ---- pass B: one key per thread -> dk, dv
  float padk[16], padv[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) { padk[c] = 0.f; padv[c] = 0.f; }
  for (int j = tid; j < n_tok; j += blockDim.x) {
    float kk[16], vv[16], dk[16], dv[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { kk[c] = sK[j * 16 + c]; vv[c] = sV[j * 16 + c]; dk[c] = 0.f; dv[c] = 0.f; }
    const int bk = sBase[j], rk = sReg[j];
    for (int i = 0; i < n_tok; ++i) {
      if (sPos[i] < 0) continue;                   // padded queries carry no gradient
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) { s = fmaf(sQ[i * 16 + c], kk[c], s); dp = fmaf(sdO[i * 16 + c], vv[c], dp); }
      float sl = fmaf(s, k.scale_log2e, sTab[sBase[i] + k.centre - bk]);
      if (shifted && sReg[i] != rk) sl -= 144.26950408889634f;
      const float p = exp2f(sl - sLse[i]);
      const float ds = p * (dp - sDelta[i]);
#pragma unroll
      for (int c = 0; c < 16; ++c) { dk[c] = fmaf(ds, sQ[i * 16 + c], dk[c]); dv[c] = fmaf(p, sdO[i * 16 + c], dv[c]); }
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) dk[c] *= k.scale;
    const int pos = sPos[j];
    if (pos >= 0) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        store8_act(dqkv, (((size_t)img * k.qkv_cbt + ((C + head * 16 + half * 8) >> 3)) * nvox + pos) * 8, 0, dk + half * 8, false);
        store8_act(dqkv, (((size_t)img * k.qkv_cbt + ((2 * C + head * 16 + half * 8) >> 3)) * nvox + pos) * 8, 0, dv + half * 8, false);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) { padk[c] += dk[c]; padv[c] += dv[c]; }
    }
  }
  // padded tokens -> qkv-bias gradient: fixed-order reduction over the CTA's threads (shared memory reused after a sync)
  __syncthreads();
  float* red = sQ;                                 // [384][32] floats = 48 KB >= needs 352*16*... reuse sQ..sV region
  for (int c = 0; c < 16; ++c) { red[tid * 32 + c] = padk[c]; red[tid * 32 + 16 + c] = padv[c]; }
  __syncthreads();
  const size_t cta = ((size_t)img * gridDim.x + win) * k.heads + head;
  if (tid < 32) {
    float s = 0.f;
    for (int t = 0; t < (int)blockDim.x; ++t) s += red[t * 32 + tid];
    k.dbias[cta * 32 + tid] = s;
  }
  for (int i = tid; i < k.table_len; i += blockDim.x) k.dtable[cta * k.table_len + i] = sdTab[i];
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
  const size_t smem = (size_t)kBwdTok * 16 * 4 * 4 + 2200 * 4 * 2 + kBwdTok * 4 * 2 + kBwdTok * 4 + kBwdTok * 2 + kBwdTok + 64;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(swin_window_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  dim3 grid((unsigned)((pp[0] / ws[0]) * (pp[1] / ws[1]) * (pp[2] / ws[2])), (unsigned)a->heads, (unsigned)a->n_img);
  swin_window_attention_bwd_kernel<<<grid, 384, smem, reinterpret_cast<cudaStream_t>(stream)>>>(k);
  return check_launch("swin_window_attention_bwd_kernel");
}
