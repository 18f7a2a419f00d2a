// Backward of InstanceNorm3d(affine=False) + ReLU/LeakyReLU (+ the MaxPool3d(2) that may follow) on blocked tensors,
// and the pixel-unshuffle that turns the gradient of a ConvTranspose3d(k2,s2) output into its k=1 GEMM view.
// HBM-bound: 16-byte vectors, two-stage deterministic reductions (no atomics).
//
//   y^ = (x - mean) * rstd,  a = act(y^),  g' = (s * gA [+ route(gP)]) * act'(y^)
//   dx = rstd * (g' - mean_v(g') - y^ * mean_v(g' * y^))            
// route(gP): MaxPool3d(2) backward — the pooled gradient goes to the first maximum of each 2x2x2 cell in scan order.
#include "common.h"
#include "ptx.cuh"

namespace mmseg {

int num_sms();

struct NormBwdK {
  const __nv_bfloat16* x;
  const float* mr;
  const __nv_bfloat16* gA;
  const __nv_bfloat16* gP;
  float* partial;
  __nv_bfloat16* dx;
  const float* chan_scale;
  const float* chan_bias;
  const float* m12;   // optional [n_img][C][2]: the two subtraction terms, given instead of being re-reduced from `partial`
  int n_img, cb, Z, Y, X;
  int gA_cbt, gA_cb_off, gP_cbt, gP_cb_off, dx_cbt, dx_cb_off;
  int n_chunks;
  float gA_scale, slope;
};

__device__ __forceinline__ void ldb8(const __nv_bfloat16* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void unpk8(const uint4& u, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pkb8(const float* v) {
  uint4 r;
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
  r.z = *reinterpret_cast<uint32_t*>(&c); r.w = *reinterpret_cast<uint32_t*>(&d);
  return r;
}

// Per-thread work item: one voxel (POOL = false) or one 2x2x2 cell (POOL = true).  Calls f(voxel_index, yhat[8], gprime[8]).
template <bool POOL, int U, typename F>
__device__ __forceinline__ void for_each_item(const NormBwdK& k, int img, int c, const float* mean, const float* rstd, F f) {
  const size_t nvox = (size_t)k.Z * k.Y * k.X;
  float cs[8];  // gA_scale x optional per-(image, channel) scale (Dropout3d: 0 or 1/(1-p))
#pragma unroll
  for (int i = 0; i < 8; ++i) cs[i] = k.gA_scale * (k.chan_scale ? k.chan_scale[(size_t)img * k.cb * 8 + c * 8 + i] : 1.f);
  float cbias[8];  // optional per-(image, channel) constant added to the activation gradient (gate: d pooled / N)
#pragma unroll
  for (int i = 0; i < 8; ++i) cbias[i] = k.chan_bias ? k.chan_bias[(size_t)img * k.cb * 8 + c * 8 + i] : 0.f;
  const size_t xbase = (size_t)(img * k.cb + c) * nvox * 8;
  const size_t abase = (size_t)(img * k.gA_cbt + k.gA_cb_off + c) * nvox * 8;
  if (!POOL) {
    // U voxels in flight per thread (all 2U 16-byte loads issued before any arithmetic): with one voxel per iteration
    // the kernels ran at 40-58 % of HBM bandwidth.  U = 4 for the read-only reduce, 2 for the apply (registers).
    const size_t step = (size_t)gridDim.x * blockDim.x;
    size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; v + (U - 1) * step < nvox; v += U * step) {
      uint4 rx[U], rg[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        rx[u] = *reinterpret_cast<const uint4*>(k.x + xbase + (v + u * step) * 8);
        rg[u] = *reinterpret_cast<const uint4*>(k.gA + abase + (v + u * step) * 8);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float x[8], g[8];
        unpk8(rx[u], x);
        unpk8(rg[u], g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          x[i] = (x[i] - mean[i]) * rstd[i];
          g[i] = (g[i] * cs[i] + cbias[i]) * (x[i] > 0.f ? 1.f : k.slope);
        }
        f(v + u * step, x, g);
      }
    }
    for (; v < nvox; v += step) {
      float x[8], g[8];
      ldb8(k.x + xbase + v * 8, x);
      ldb8(k.gA + abase + v * 8, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        x[i] = (x[i] - mean[i]) * rstd[i];
        g[i] = (g[i] * cs[i] + cbias[i]) * (x[i] > 0.f ? 1.f : k.slope);
      }
      f(v, x, g);
    }
  } else {
    const int Zh = k.Z / 2, Yh = k.Y / 2, Xh = k.X / 2;
    const size_t ncell = (size_t)Zh * Yh * Xh;
    const size_t pbase = (size_t)(img * k.gP_cbt + k.gP_cb_off + c) * ncell * 8;
    for (size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x; cell < ncell;
         cell += (size_t)gridDim.x * blockDim.x) {
      const int xh = (int)(cell % Xh);
      const size_t r = cell / Xh;
      const int yh = (int)(r % Yh), zh = (int)(r / Yh);
      float gp[8];
      ldb8(k.gP + pbase + cell * 8, gp);
      // pass 1: which of the 8 voxels holds each channel's maximum (3 bits per channel packed in one register).
      // The normalised values are NOT kept (8 x 8 floats pushed the kernel to 198 / 251 registers = one block per SM
      // and ~1.2 TB/s); pass 2 re-reads the 8 vectors, which are still in L1.
      float mx[8];
      uint32_t arg = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
      const size_t v000 = ((size_t)(2 * zh) * k.Y + (2 * yh)) * k.X + (2 * xh);
      // the cell's 8 gA vectors are only consumed in pass 2 (two in flight per thread there): request them now, next to
      // the 8 loads of pass 1, without holding 32 more registers (the lane pair d & 1 shares a 32-byte sector)
      if (k.gA) {
#pragma unroll
        for (int d = 0; d < 8; d += 2) {
          const size_t v = v000 + ((size_t)(d >> 2) * k.Y + ((d >> 1) & 1)) * k.X;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(k.gA + abase + v * 8));
        }
      }
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        const size_t v = v000 + ((size_t)(d >> 2) * k.Y + ((d >> 1) & 1)) * k.X + (d & 1);
        float x[8];
        ldb8(k.x + xbase + v * 8, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float y = (x[i] - mean[i]) * rstd[i];
          const float a = y > 0.f ? y : y * k.slope;
          if (a > mx[i]) { mx[i] = a; arg = (arg & ~(7u << (3 * i))) | ((uint32_t)d << (3 * i)); }   // strict >: first maximum in scan order
        }
      }
#pragma unroll 2   // (fully unrolled the compiler hoists all 16 loads: 254 registers or spills)
      for (int d = 0; d < 8; ++d) {
        const size_t v = v000 + ((size_t)(d >> 2) * k.Y + ((d >> 1) & 1)) * k.X + (d & 1);
        float y[8], g[8];
        ldb8(k.x + xbase + v * 8, y);
        if (k.gA) {
          ldb8(k.gA + abase + v * 8, g);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] = g[i] * cs[i] + cbias[i];
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] = cbias[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          y[i] = (y[i] - mean[i]) * rstd[i];
          if (((arg >> (3 * i)) & 7u) == (uint32_t)d) g[i] += gp[i];
          g[i] *= (y[i] > 0.f ? 1.f : k.slope);
        }
        f(v, y, g);
      }
    }
  }
}

__device__ __forceinline__ void load_mr(const NormBwdK& k, int img, int c, float* mean, float* rstd) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 m = *reinterpret_cast<const float2*>(k.mr + ((size_t)img * k.cb * 8 + c * 8 + i) * 2);
    mean[i] = m.x;
    rstd[i] = m.y;
  }
}

// grid (n_chunks, n_img*cb): partial[(blk*n_chunks + chunk)*16 + {i, 8+i}] = sum g', sum g'*y^
template <bool POOL>
__global__ void __launch_bounds__(256, 2) norm_bwd_reduce_kernel(const NormBwdK k) {
  const int blk = blockIdx.y;
  const int img = blk / k.cb, c = blk - img * k.cb;
  float mean[8], rstd[8];
  load_mr(k, img, c, mean, rstd);
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  for_each_item<POOL, 4>(k, img, c, mean, rstd, [&](size_t, const float* y, const float* g) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1[i] += g[i]; s2[i] = fmaf(g[i], y[i], s2[i]); }
  });
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
    for (int i = 0; i < 8; ++i) { red[warp][i] = s1[i]; red[warp][8 + i] = s2[i]; }
  __syncthreads();
  if (threadIdx.x < 16) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    k.partial[((size_t)blk * gridDim.x + blockIdx.x) * 16 + threadIdx.x] = t;
  }
}

// grid (any, n_img*cb): every block first re-reduces the n_chunks partials of its (img, cb) in a fixed order
template <bool POOL>
__global__ void __launch_bounds__(256, 2) norm_bwd_apply_kernel(const NormBwdK k) {
  const int blk = blockIdx.y;
  const int img = blk / k.cb, c = blk - img * k.cb;
  __shared__ float m12[16];
  if (threadIdx.x < 16) {
    if (k.m12) {   // affine / group / batch norms: the caller combined the sums over the norm's reduction set (train_engine.py)
      const int i = threadIdx.x & 7, which = threadIdx.x >> 3;
      m12[threadIdx.x] = k.m12[((size_t)img * k.cb * 8 + c * 8 + i) * 2 + which];
    } else {
      double s = 0.0;
      for (int j = 0; j < k.n_chunks; ++j) s += (double)k.partial[((size_t)blk * k.n_chunks + j) * 16 + threadIdx.x];
      m12[threadIdx.x] = (float)(s / ((double)k.Z * k.Y * k.X));
    }
  }
  __syncthreads();
  float mean[8], rstd[8], m1[8], m2[8];
  load_mr(k, img, c, mean, rstd);
#pragma unroll
  for (int i = 0; i < 8; ++i) { m1[i] = m12[i]; m2[i] = m12[8 + i]; }
  const size_t nvox = (size_t)k.Z * k.Y * k.X;
  const size_t dbase = (size_t)(img * k.dx_cbt + k.dx_cb_off + c) * nvox * 8;
  for_each_item<POOL, 2>(k, img, c, mean, rstd, [&](size_t v, const float* y, const float* g) {
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = rstd[i] * (g[i] - m1[i] - y[i] * m2[i]);
    *reinterpret_cast<uint4*>(k.dx + dbase + v * 8) = pkb8(d);
  });
}

// Weight gradient of a 3x3x3 conv with ONE input channel (DualEncoder's per-modality first layers): as a wgrad GEMM its
// accumulator has 3 live rows of 128 and still costs the MMAs of a 32-channel layer (0.33 ms at 2 x 128^3).  The 27
// shifted copies of the channel are written as a 32-channel blocked tensor instead (channel t = tap (dz*3+dy)*3+dx, zero
// outside the volume = the conv's padding, channels 27..31 zero); dW[co][0][tap] is then the k = 1 weight gradient of that
// tensor, an HBM-bound GEMM.
__global__ void __launch_bounds__(256)
im2col_k3_c1_kernel(const uint16_t* __restrict__ src, int src_cbt, int src_cb, int src_lane, int Z, int Y, int X,
                    uint16_t* __restrict__ dst) {
  const int img = blockIdx.y;
  const size_t nvox = (size_t)Z * Y * X;
  const uint16_t* s = src + (size_t)(img * src_cbt + src_cb) * nvox * 8 + src_lane;
  uint16_t* d = dst + (size_t)img * 4 * nvox * 8;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(v % X);
    const size_t r = v / X;
    const int y = (int)(r % Y), z = (int)(r / Y);
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = 0u;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const int zz = z + t / 9 - 1, yy = y + (t / 3) % 3 - 1, xx = x + t % 3 - 1;
      uint32_t val = 0u;
      if (zz >= 0 && zz < Z && yy >= 0 && yy < Y && xx >= 0 && xx < X) val = s[(((size_t)zz * Y + yy) * X + xx) * 8];
      w[t >> 1] |= val << ((t & 1) * 16);
    }
#pragma unroll
    for (int b = 0; b < 4; ++b)
      *reinterpret_cast<uint4*>(d + ((size_t)b * nvox + v) * 8) = make_uint4(w[4 * b], w[4 * b + 1], w[4 * b + 2], w[4 * b + 3]);
  }
}

// gradient of a ConvTranspose3d(k2,s2) output [n_img*src_cbt][2Z][2Y][2X][8] (channels cb_off*8 .. +C) ->
// its GEMM view [n_img*8*cb][Z][Y][X][8] with channel = tap*C + co, tap = (dz*2 + dy)*2 + dx
__global__ void __launch_bounds__(256)
unshuffle_k2s2_kernel(const __nv_bfloat16* __restrict__ src, int src_cbt, int src_cb_off, int cb, int Z, int Y, int X,
                      __nv_bfloat16* __restrict__ dst) {
  const int blk = blockIdx.y;  // img*cb + c
  const int img = blk / cb, c = blk - img * cb;
  const size_t nlo = (size_t)Z * Y * X;
  const size_t sbase = (size_t)(img * src_cbt + src_cb_off + c) * nlo * 8 * 8;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nlo; v += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(v % X);
    const size_t r = v / X;
    const int y = (int)(r % Y), z = (int)(r / Y);
#pragma unroll
    for (int tap = 0; tap < 8; ++tap) {
      const size_t sv = ((size_t)(2 * z + (tap >> 2)) * (2 * Y) + (2 * y + ((tap >> 1) & 1))) * (2 * X) + (2 * x + (tap & 1));
      const uint4 u = *reinterpret_cast<const uint4*>(src + sbase + sv * 8);
      *reinterpret_cast<uint4*>(dst + ((size_t)(img * 8 * cb + tap * cb + c) * nlo + v) * 8) = u;
    }
  }
}

// dw[b, m] = sum_{c, v} g[b, c, v] * x[b, m*C + c, v]: gradient of DualEncoder's modality gate weights
// (dual_encoder.py:251-252).  grid (n_chunks, n_img*M*cb) -> partial[(img*M + m)*cb + c][chunk]
__global__ void __launch_bounds__(256)
modality_dot_partial_kernel(const __nv_bfloat16* __restrict__ x, int x_cbt, const __nv_bfloat16* __restrict__ g, int g_cbt,
                            int g_cb_off, int M, int cb, size_t nvox, float* __restrict__ partial) {
  const int blk = blockIdx.y;             // (img*M + m)*cb + c
  const int c = blk % cb;
  const int im = blk / cb;
  const int m = im % M, img = im / M;
  const size_t xbase = (size_t)(img * x_cbt + m * cb + c) * nvox * 8;
  const size_t gbase = (size_t)(img * g_cbt + g_cb_off + c) * nvox * 8;
  float s = 0.f;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (size_t)gridDim.x * blockDim.x) {
    float a[8], b[8];
    ldb8(x + xbase + v * 8, a);
    ldb8(g + gbase + v * 8, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(a[i], b[i], s);
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[(size_t)blk * gridDim.x + blockIdx.x] = t;
  }
}

__global__ void modality_dot_final_kernel(const float* __restrict__ partial, int n, int per, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < per; ++k) s += (double)partial[(size_t)i * per + k];
  out[i] = (float)s;
}

static unsigned gxb(size_t items, int rows, int cap_per_row) {
  size_t want = ((size_t)num_sms() * 8 + rows - 1) / rows;
  size_t need = (items + 255) / 256;
  size_t g = want < need ? want : need;
  if (g > (size_t)cap_per_row) g = cap_per_row;
  return (unsigned)(g < 1 ? 1 : g);
}

static int fill_norm_bwd(const mmseg_norm_bwd_args* a, NormBwdK* k) {
  if (!a || !a->x || !a->mean_rstd || !a->partial) return fail(MMSEG_ERR_INVALID_ARG, "instnorm_act_bwd: null pointer");
  if (!a->gA && !a->gP) return fail(MMSEG_ERR_INVALID_ARG, "instnorm_act_bwd: no incoming gradient");
  if (a->n_img < 1 || a->cb < 1 || a->Z < 1 || a->Y < 1 || a->X < 1 || a->n_chunks < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "instnorm_act_bwd: bad extents");
  if (a->gP && ((a->Z | a->Y | a->X) & 1)) return fail(MMSEG_ERR_UNSUPPORTED, "instnorm_act_bwd: MaxPool3d(2) routing needs even extents");
  k->x = reinterpret_cast<const __nv_bfloat16*>(a->x); k->mr = a->mean_rstd;
  k->gA = reinterpret_cast<const __nv_bfloat16*>(a->gA); k->gP = reinterpret_cast<const __nv_bfloat16*>(a->gP);
  k->partial = a->partial; k->dx = reinterpret_cast<__nv_bfloat16*>(a->dx);
  k->n_img = a->n_img; k->cb = a->cb; k->Z = a->Z; k->Y = a->Y; k->X = a->X;
  k->gA_cbt = a->gA_cbt; k->gA_cb_off = a->gA_cb_off; k->gP_cbt = a->gP_cbt; k->gP_cb_off = a->gP_cb_off;
  k->dx_cbt = a->dx_cbt; k->dx_cb_off = a->dx_cb_off; k->n_chunks = a->n_chunks;
  k->gA_scale = a->gA_scale; k->slope = a->slope; k->chan_scale = a->chan_scale; k->chan_bias = a->chan_bias;
  k->m12 = a->m12;
  return MMSEG_OK;
}

}  // namespace mmseg

using namespace mmseg;

extern "C" int mmseg_instnorm_act_bwd_reduce(const mmseg_norm_bwd_args* a, void* stream) {
  NormBwdK k;
  int rc = fill_norm_bwd(a, &k);
  if (rc) return rc;
  dim3 grid((unsigned)a->n_chunks, (unsigned)(a->n_img * a->cb));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->gP) norm_bwd_reduce_kernel<true><<<grid, 256, 0, st>>>(k);
  else norm_bwd_reduce_kernel<false><<<grid, 256, 0, st>>>(k);
  return check_launch("norm_bwd_reduce_kernel");
}

extern "C" int mmseg_instnorm_act_bwd_apply(const mmseg_norm_bwd_args* a, void* stream) {
  NormBwdK k;
  int rc = fill_norm_bwd(a, &k);
  if (rc) return rc;
  if (!a->dx) return fail(MMSEG_ERR_INVALID_ARG, "instnorm_act_bwd_apply: null dx");
  const int rows = a->n_img * a->cb;
  const size_t items = a->gP ? (size_t)(a->Z / 2) * (a->Y / 2) * (a->X / 2) : (size_t)a->Z * a->Y * a->X;
  dim3 grid(gxb(items, rows, 65535), (unsigned)rows);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->gP) norm_bwd_apply_kernel<true><<<grid, 256, 0, st>>>(k);
  else norm_bwd_apply_kernel<false><<<grid, 256, 0, st>>>(k);
  return check_launch("norm_bwd_apply_kernel");
}

extern "C" int mmseg_modality_dot(const void* x, int32_t x_cbt, const void* g, int32_t g_cbt, int32_t g_cb_off,
                                  int32_t n_img, int32_t M, int32_t cb, int64_t voxels, float* partial, int32_t n_chunks,
                                  float* out, void* stream) {
  if (!x || !g || !partial || !out || n_img < 1 || M < 1 || cb < 1 || voxels < 1 || n_chunks < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "modality_dot: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)n_chunks, (unsigned)(n_img * M * cb));
  modality_dot_partial_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), x_cbt,
                                                    reinterpret_cast<const __nv_bfloat16*>(g), g_cbt, g_cb_off, M, cb,
                                                    (size_t)voxels, partial);
  int rc = check_launch("modality_dot_partial_kernel");
  if (rc) return rc;
  const int n = n_img * M;
  modality_dot_final_kernel<<<(n + 63) / 64, 64, 0, st>>>(partial, n, cb * n_chunks, out);
  return check_launch("modality_dot_final_kernel");
}

extern "C" int mmseg_im2col_k3_c1(const void* src, int32_t n_img, int32_t src_cbt, int32_t src_cb, int32_t src_lane,
                                  int32_t Z, int32_t Y, int32_t X, void* dst, void* stream) {
  if (!src || !dst || n_img < 1 || src_cbt < 1 || src_cb < 0 || src_cb >= src_cbt || src_lane < 0 || src_lane > 7 || Z < 1 ||
      Y < 1 || X < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "im2col_k3_c1: bad arguments");
  dim3 grid(gxb((size_t)Z * Y * X, n_img, 65535), (unsigned)n_img);
  im2col_k3_c1_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint16_t*>(src), src_cbt, src_cb, src_lane, Z, Y, X, reinterpret_cast<uint16_t*>(dst));
  return check_launch("im2col_k3_c1_kernel");
}

extern "C" int mmseg_unshuffle_k2s2(const void* src, int32_t n_img, int32_t src_cbt, int32_t src_cb_off, int32_t cb,
                                    int32_t Z, int32_t Y, int32_t X, void* dst, void* stream) {
  if (!src || !dst || n_img < 1 || cb < 1 || Z < 1 || Y < 1 || X < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "unshuffle_k2s2: bad arguments");
  const int rows = n_img * cb;
  dim3 grid(gxb((size_t)Z * Y * X, rows, 65535), (unsigned)rows);
  unshuffle_k2s2_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), src_cbt, src_cb_off, cb, Z, Y, X,
      reinterpret_cast<__nv_bfloat16*>(dst));
  return check_launch("unshuffle_k2s2_kernel");
}
