// InstanceNorm3d statistics finalize + fused normalise / activation / (MaxPool3d(2)) apply, and the NCDHW <-> blocked
// layout converters at the module boundary.  All HBM-bound: 16-byte vector accesses, grid sized from the SM count.
#include "common.h"
#include "ptx.cuh"

namespace mmseg {

static int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// ---------------------------------------------------------------------------------------------- finalize
// one warp per (img, channel): lanes stride over the conv CTAs' partials (fp64), fixed-order shuffle tree -> mean, rstd.
// Deterministic: the lane a partial lands in and the tree order depend only on the tile count.
__global__ void __launch_bounds__(256)
instnorm_finalize_kernel(const float* __restrict__ part, int n_img, int tiles, int C, double inv_n, float eps,
                         float* __restrict__ mean_rstd) {
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (idx >= n_img * C) return;
  const int img = idx / C, c = idx - img * C;
  const float* p = part + ((size_t)img * tiles * C + c) * 2;
  double s1 = 0.0, s2 = 0.0;
  // the loads of 8 iterations are in flight together (a lane walks up to 60 partial rows, each a different sector: the
  // rolled loop paid one L2 / DRAM latency per row, ~14 us per launch); the additions keep their order, bit for bit
#pragma unroll 8
  for (int t = lane; t < tiles; t += 32) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p + (size_t)t * C * 2));
    s1 += (double)v.x;
    s2 += (double)v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) {
    const double mean = s1 * inv_n;
    double var = s2 * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_rstd[2 * idx] = (float)mean;
    mean_rstd[2 * idx + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// GroupNorm(G, C) statistics (ConvBlock3D norm="group", unet.py:36-38): one warp per (image, group) sums the conv
// epilogue's per-tile partials of the group's C/G channels in a fixed order (fp64) and writes, per channel, the table
// the apply kernel consumes: (mean_g, rstd_g * gamma_c) and shift = beta_c, i.e. y = (x - mean_g) * rstd_g * gamma + beta.
__global__ void __launch_bounds__(256)
groupnorm_finalize_kernel(const float* __restrict__ part, int n_img, int tiles, int C, int G, double inv_n, float eps,
                          const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ mean_rstd,
                          float* __restrict__ shift) {
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (idx >= n_img * G) return;
  const int img = idx / G, g = idx - img * G;
  const int cpg = C / G;
  const float* p = part + ((size_t)img * tiles * C + (size_t)g * cpg) * 2;
  double s1 = 0.0, s2 = 0.0;
  for (int t = lane; t < tiles; t += 32) {
    const float* q = p + (size_t)t * C * 2;
    for (int c = 0; c < cpg; ++c) {
      s1 += (double)q[2 * c];
      s2 += (double)q[2 * c + 1];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const double m = s1 * inv_n;
  double var = s2 * inv_n - m * m;
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  for (int c = lane; c < cpg; c += 32) {
    const int ch = g * cpg + c;
    const float ga = gamma ? gamma[ch] : 1.f;
    mean_rstd[((size_t)img * C + ch) * 2] = (float)m;
    mean_rstd[((size_t)img * C + ch) * 2 + 1] = (float)(rstd * (double)ga);
    shift[(size_t)img * C + ch] = beta ? beta[ch] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------- apply
struct NormK {   // kept at exactly 128 bytes: one more field and the default apply kernel goes from 80 to 88 registers
  const void* src;
  const float* mr;
  const float* stats;   // optional: per-tile (sum, sum of squares) partials [n_img][tiles][C][2] -> finalized in the prologue
  float* mr_out;        // optional: where block x == 0 of each row publishes the (mean, rstd) it derived
  void* dst;
  void* pooled;
  const float* shift;   // optional [n_img][C]: y = (x - mean) * rstd + shift (affine norms: GroupNorm / BatchNorm beta)
  double inv_n;
  int tiles;
  int n_img, cb, Z, Y, X;
  int dst_cbt, dst_cb_off, dst_lo_off;
  int pool_cbt, pool_cb_off, pool_lo_off;
  float eps;
  float slope;
  int act;              // 0: ReLU / LeakyReLU(slope); 1: exact GELU (nn.GELU(), ConvBlock3D activation="gelu", unet.py:47-48)
};
static_assert(sizeof(NormK) <= 128, "NormK must stay within 128 bytes (register allocation of the apply kernels)");

template <bool F32, bool FP16>
__device__ __forceinline__ void load8(const void* base, size_t elem_off, float* v) {
  if (F32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
    const float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    load8_act(base, elem_off, v, FP16);
  }
}

// mean / rstd of this block's 8 channels: read from the finalized table, or (stats != NULL) finalized HERE from the conv
// epilogue's per-tile partials in a fixed order with fp64 accumulation — every block of a row derives bit-identical
// values, and the 18 instnorm_finalize launches of a UNet3D forward (latency-bound, ~9 us each) disappear.  The partials
// (<= 2 MB) are L2-resident; a block reads tiles x 64 B of them.
__device__ __forceinline__ void block_mean_rstd(const NormK& k, int img, int c, float* mean, float* rstd) {
  if (k.stats == nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 m = *reinterpret_cast<const float2*>(k.mr + ((size_t)img * k.cb * 8 + c * 8 + i) * 2);
      mean[i] = m.x;
      rstd[i] = m.y;
    }
    return;
  }
  __shared__ double sred[16][17];
  __shared__ float smr[16];
  const int C = k.cb * 8;
  const int j = threadIdx.x & 15, r = threadIdx.x >> 4;   // value j = (channel i, {sum, sumsq}) x 16 row groups
  const float* p = k.stats + ((size_t)img * k.tiles * C + (size_t)c * 8) * 2 + j;
  double acc = 0.0;
  int t = r;
  for (; t + 7 * 16 < k.tiles; t += 8 * 16) {   // eight independent L2 loads in flight, summed in a fixed order
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(p + (size_t)(t + u * 16) * C * 2);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += (double)v[u];
  }
  for (; t < k.tiles; t += 16) acc += (double)__ldg(p + (size_t)t * C * 2);
  sred[r][j] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int g = 0; g < 16; ++g) { s1 += sred[g][2 * threadIdx.x]; s2 += sred[g][2 * threadIdx.x + 1]; }
    const double m = s1 * k.inv_n;
    double var = s2 * k.inv_n - m * m;
    if (var < 0.0) var = 0.0;
    const float mf = (float)m, rf = (float)(1.0 / sqrt(var + (double)k.eps));
    smr[2 * threadIdx.x] = mf;
    smr[2 * threadIdx.x + 1] = rf;
    if (k.mr_out && blockIdx.x == 0) {
      k.mr_out[((size_t)img * C + c * 8 + threadIdx.x) * 2] = mf;
      k.mr_out[((size_t)img * C + c * 8 + threadIdx.x) * 2 + 1] = rf;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) { mean[i] = smr[2 * i]; rstd[i] = smr[2 * i + 1]; }
}

__device__ __forceinline__ float act_fn(float y, float slope, int act) {
  if (act == 1) return 0.5f * y * (1.f + erff(y * 0.70710678118654752440f));
  return y > 0.f ? y : y * slope;
}

__device__ __forceinline__ void block_shift(const NormK& k, int img, int c, float* sh) {
#pragma unroll
  for (int i = 0; i < 8; ++i) sh[i] = k.shift ? k.shift[(size_t)img * k.cb * 8 + c * 8 + i] : 0.f;
}

// grid: (chunks over voxels, n_img*cb).  One 8-channel vector per thread-iteration.
// EXT = additive shift and / or GELU (the affine norms / gelu option): kept out of the default instantiation, whose
// register count decides how many loads are in flight (80 -> 93 registers cost a resident block per SM)
template <bool F32, bool EXT, bool FP16>
__global__ void __launch_bounds__(256) instnorm_apply_kernel(const NormK k) {
  const int blk = blockIdx.y;  // img*cb + c
  const int img = blk / k.cb, c = blk - img * k.cb;
  float mean[8], rstd[8], sh[8];
  block_mean_rstd(k, img, c, mean, rstd);
  if constexpr (EXT) block_shift(k, img, c, sh);
  const size_t nvox = (size_t)k.Z * k.Y * k.X;
  const size_t src_base = (size_t)blk * nvox * 8;
  const size_t dst_base = (size_t)(img * k.dst_cbt + k.dst_cb_off + c) * nvox * 8;
  const size_t lo_delta = (size_t)k.dst_lo_off * nvox * 8;
  // four voxels per thread and iteration: four independent 16-byte loads in flight (HBM latency hiding)
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvox; v0 += 4 * stride) {
    float x[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t v = v0 + u * stride;
      if (v < nvox) load8<F32, FP16>(k.src, src_base + v * 8, x[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t v = v0 + u * stride;
      if (v < nvox) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if constexpr (EXT) {
            x[u][i] = act_fn((x[u][i] - mean[i]) * rstd[i] + sh[i], k.slope, k.act);
          } else {
            const float y = (x[u][i] - mean[i]) * rstd[i];
            x[u][i] = y > 0.f ? y : y * k.slope;
          }
        }
        store8_act(k.dst, dst_base + v * 8, lo_delta, x[u], FP16);
      }
    }
  }
}

// Variant that also emits MaxPool3d(2): one 2x2x2 cell per thread-iteration (Z, Y, X even).
template <bool F32, bool EXT, bool FP16>
__global__ void __launch_bounds__(256) instnorm_apply_pool_kernel(const NormK k) {
  const int blk = blockIdx.y;
  const int img = blk / k.cb, c = blk - img * k.cb;
  float mean[8], rstd[8], sh[8];
  block_mean_rstd(k, img, c, mean, rstd);
  if constexpr (EXT) block_shift(k, img, c, sh);
  const int Zh = k.Z / 2, Yh = k.Y / 2, Xh = k.X / 2;
  const size_t nvox = (size_t)k.Z * k.Y * k.X;
  const size_t ncell = (size_t)Zh * Yh * Xh;
  const size_t src_base = (size_t)blk * nvox * 8;
  const size_t dst_base = (size_t)(img * k.dst_cbt + k.dst_cb_off + c) * nvox * 8;
  const size_t lo_delta = (size_t)k.dst_lo_off * nvox * 8;
  const size_t pool_base = (size_t)(img * k.pool_cbt + k.pool_cb_off + c) * ncell * 8;
  const size_t pool_lo = (size_t)k.pool_lo_off * ncell * 8;
  for (size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x; cell < ncell;
       cell += (size_t)gridDim.x * blockDim.x) {
    const int xh = (int)(cell % Xh);
    const size_t r = cell / Xh;
    const int yh = (int)(r % Yh);
    const int zh = (int)(r / Yh);
    float mx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const size_t v = ((size_t)(2 * zh + dz) * k.Y + (2 * yh + dy)) * k.X + (2 * xh + dx);
          float x[8];
          load8<F32, FP16>(k.src, src_base + v * 8, x);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if constexpr (EXT) {
              x[i] = act_fn((x[i] - mean[i]) * rstd[i] + sh[i], k.slope, k.act);
            } else {
              const float y = (x[i] - mean[i]) * rstd[i];
              x[i] = y > 0.f ? y : y * k.slope;
            }
            mx[i] = fmaxf(mx[i], x[i]);
          }
          store8_act(k.dst, dst_base + v * 8, lo_delta, x, FP16);
        }
    store8_act(k.pooled, pool_base + cell * 8, pool_lo, mx, FP16);
  }
}

// ---------------------------------------------------------------------------------------------- layout converters
__global__ void __launch_bounds__(256)
pack_ncdhw_kernel(const float* __restrict__ src, void* __restrict__ dst, int n_img, int C, size_t nvox,
                  int dst_cbt, int dst_cb_off, int dst_lo_off, int cb, int fp16) {
  const int blk = blockIdx.y;  // img*cb + c
  const int img = blk / cb, c = blk - img * cb;
  const size_t dst_base = (size_t)(img * dst_cbt + dst_cb_off + c) * nvox * 8;
  const size_t lo_delta = dst_lo_off > 0 ? (size_t)dst_lo_off * nvox * 8 : 0;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (size_t)gridDim.x * blockDim.x) {
    float x[8];
    if (dst_lo_off < 0) {   // packed split: virtual channels [hi(C) | lo(C) | hi(C)] in one 16-channel K chunk (see swi_gather_kernel)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int vc = c * 8 + i;
        const int part = vc / C, ch = vc - part * C;
        const float t = part < 3 ? src[((size_t)img * C + ch) * nvox + v] : 0.f;
        const float hi = fp16 ? __half2float(__float2half_rn(t)) : __bfloat162float(__float2bfloat16_rn(t));
        x[i] = part == 1 ? t - hi : hi;
      }
      store8_act(dst, dst_base + v * 8, 0, x, fp16 != 0);
      continue;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = c * 8 + i;
      x[i] = ch < C ? src[((size_t)img * C + ch) * nvox + v] : 0.f;
    }
    store8_act(dst, dst_base + v * 8, lo_delta, x, fp16 != 0);
  }
}

// pack with the two element-wise steps of SUVGuidedAttention (fusion/attention_fusion.py:283-292) folded in:
//   pre:  x <- sigmoid((x - pre_sub) * pre_mul)          (the soft SUV mask of the 1-channel PET image), and / or
//   gate: x <- x * (1 + sigmoid(gate[img][voxel]))        (CT features modulated by the spatial attention LOGITS)
__global__ void __launch_bounds__(256)
pack_ncdhw_ex_kernel(const float* __restrict__ src, void* __restrict__ dst, int n_img, int C, size_t nvox,
                     int dst_cbt, int dst_cb_off, int dst_lo_off, int cb, int pre, float pre_sub, float pre_mul,
                     const float* __restrict__ gate, int fp16) {
  const int blk = blockIdx.y;  // img*cb + c
  const int img = blk / cb, c = blk - img * cb;
  const size_t dst_base = (size_t)(img * dst_cbt + dst_cb_off + c) * nvox * 8;
  const size_t lo_delta = (size_t)dst_lo_off * nvox * 8;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (size_t)gridDim.x * blockDim.x) {
    float g = 1.f;
    if (gate) g = 1.f + 1.f / (1.f + expf(-gate[(size_t)img * nvox + v]));
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = c * 8 + i;
      float t = ch < C ? src[((size_t)img * C + ch) * nvox + v] : 0.f;
      if (pre && ch < C) t = 1.f / (1.f + expf(-(t - pre_sub) * pre_mul));
      x[i] = t * g;
    }
    store8_act(dst, dst_base + v * 8, lo_delta, x, fp16 != 0);
  }
}

__global__ void __launch_bounds__(256)
unpack_ncdhw_kernel(const void* __restrict__ src, float* __restrict__ dst, int n_img, int C, size_t nvox,
                    int src_cbt, int src_cb_off, int src_lo_off, int fp16) {
  const int cbn = (C + 7) / 8;
  const int blk = blockIdx.y;
  const int img = blk / cbn, c = blk - img * cbn;
  const size_t src_base = (size_t)(img * src_cbt + src_cb_off + c) * nvox * 8;
  const size_t lo_delta = (size_t)src_lo_off * nvox * 8;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (size_t)gridDim.x * blockDim.x) {
    float x[8];
    load8_act(src, src_base + v * 8, x, fp16 != 0);
    if (lo_delta) {
      float l[8];
      load8_act(src, src_base + lo_delta + v * 8, l, fp16 != 0);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] += l[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = c * 8 + i;
      if (ch < C) dst[((size_t)img * C + ch) * nvox + v] = x[i];
    }
  }
}

static unsigned grid_x_for(size_t items, int rows) {
  // enough CTAs for ~8 resident per SM across all rows, never more than the work
  size_t want = ((size_t)num_sms() * 8 + rows - 1) / rows;
  size_t need = (items + 255) / 256;
  size_t g = want < need ? want : need;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace mmseg

using namespace mmseg;

extern "C" int mmseg_instnorm_finalize(const float* stats_partial, int32_t n_img, int32_t tiles_per_img,
                                       int32_t channels, int64_t voxels, float eps, float* mean_rstd, void* stream) {
  if (!stats_partial || !mean_rstd || n_img < 1 || tiles_per_img < 1 || channels < 1 || voxels < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "instnorm_finalize: bad arguments");
  const int n = n_img * channels;
  instnorm_finalize_kernel<<<(n + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      stats_partial, n_img, tiles_per_img, channels, 1.0 / (double)voxels, eps, mean_rstd);
  return check_launch("instnorm_finalize_kernel");
}

extern "C" int mmseg_groupnorm_finalize(const float* stats_partial, int32_t n_img, int32_t tiles_per_img, int32_t channels,
                                        int32_t groups, int64_t voxels, float eps, const float* gamma, const float* beta,
                                        float* mean_rstd, float* shift, void* stream) {
  if (!stats_partial || !mean_rstd || !shift || n_img < 1 || tiles_per_img < 1 || channels < 1 || groups < 1 || voxels < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "groupnorm_finalize: bad arguments");
  if (channels % groups) return fail(MMSEG_ERR_INVALID_ARG, "groupnorm_finalize: channels %d not divisible by groups %d", channels, groups);
  const int n = n_img * groups;
  const double inv_n = 1.0 / ((double)voxels * (double)(channels / groups));
  groupnorm_finalize_kernel<<<(n + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      stats_partial, n_img, tiles_per_img, channels, groups, inv_n, eps, gamma, beta, mean_rstd, shift);
  return check_launch("groupnorm_finalize_kernel");
}

extern "C" int mmseg_instnorm_act_apply(const mmseg_norm_args* a, void* stream) {
  if (!a || !a->src || !a->dst || (!a->mean_rstd && !a->stats_partial))
    return fail(MMSEG_ERR_INVALID_ARG, "instnorm_apply: null pointer");
  if (a->stats_partial && a->tiles_per_img < 1) return fail(MMSEG_ERR_INVALID_ARG, "instnorm_apply: tiles_per_img");
  if (a->n_img < 1 || a->cb < 1 || a->Z < 1 || a->Y < 1 || a->X < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "instnorm_apply: bad extents");
  NormK k;
  k.src = a->src; k.mr = a->mean_rstd;
  k.shift = a->shift;
  k.stats = a->stats_partial; k.mr_out = a->stats_partial ? a->mean_rstd_out : nullptr;
  k.tiles = a->tiles_per_img; k.eps = a->eps;
  k.inv_n = 1.0 / ((double)a->Z * a->Y * a->X);
  k.dst = a->dst;
  k.pooled = a->pooled;
  k.n_img = a->n_img; k.cb = a->cb; k.Z = a->Z; k.Y = a->Y; k.X = a->X;
  k.dst_cbt = a->dst_cbt; k.dst_cb_off = a->dst_cb_off; k.dst_lo_off = a->dst_lo_off;
  k.pool_cbt = a->pool_cbt; k.pool_cb_off = a->pool_cb_off; k.pool_lo_off = a->pool_lo_off;
  k.slope = a->slope;
  k.act = a->act;
  if (k.act != 0 && k.act != 1) return fail(MMSEG_ERR_INVALID_ARG, "instnorm_apply: act=%d (0 relu/leaky_relu, 1 gelu)", k.act);
  if (a->elem_fmt != MMSEG_FMT_BF16 && a->elem_fmt != MMSEG_FMT_FP16)
    return fail(MMSEG_ERR_INVALID_ARG, "instnorm_apply: elem_fmt=%d", a->elem_fmt);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rows = a->n_img * a->cb;
  const bool ext = k.shift != nullptr || k.act != 0;
  const bool f32 = a->src_is_f32 != 0, fp16 = a->elem_fmt == MMSEG_FMT_FP16;
  // instantiation index: (raw fp32, affine shift / gelu, fp16 elements)
  const int inst = (f32 ? 4 : 0) | (ext ? 2 : 0) | (fp16 ? 1 : 0);
  typedef void (*ApplyFn)(const NormK);
  if (a->pooled) {
    if ((a->Z | a->Y | a->X) & 1) return fail(MMSEG_ERR_UNSUPPORTED, "instnorm_apply: fused MaxPool3d(2) needs even extents");
    static const ApplyFn pool_fns[8] = {
        instnorm_apply_pool_kernel<false, false, false>, instnorm_apply_pool_kernel<false, false, true>,
        instnorm_apply_pool_kernel<false, true, false>,  instnorm_apply_pool_kernel<false, true, true>,
        instnorm_apply_pool_kernel<true, false, false>,  instnorm_apply_pool_kernel<true, false, true>,
        instnorm_apply_pool_kernel<true, true, false>,   instnorm_apply_pool_kernel<true, true, true>};
    const size_t ncell = (size_t)(a->Z / 2) * (a->Y / 2) * (a->X / 2);
    dim3 grid(grid_x_for(ncell, rows), rows);
    pool_fns[inst]<<<grid, 256, 0, st>>>(k);
  } else {
    static const ApplyFn fns[8] = {
        instnorm_apply_kernel<false, false, false>, instnorm_apply_kernel<false, false, true>,
        instnorm_apply_kernel<false, true, false>,  instnorm_apply_kernel<false, true, true>,
        instnorm_apply_kernel<true, false, false>,  instnorm_apply_kernel<true, false, true>,
        instnorm_apply_kernel<true, true, false>,   instnorm_apply_kernel<true, true, true>};
    const size_t nvox = (size_t)a->Z * a->Y * a->X;
    dim3 grid(grid_x_for(nvox, rows), rows);
    fns[inst]<<<grid, 256, 0, st>>>(k);
  }
  return check_launch("instnorm_apply_kernel");
}

extern "C" int mmseg_pack_ncdhw(const float* src, void* dst, int32_t n_img, int32_t C, int32_t Z, int32_t Y,
                                int32_t X, int32_t dst_cbt, int32_t dst_cb_off, int32_t dst_lo_off, int32_t cb,
                                int32_t fmt, void* stream) {
  if (!src || !dst || n_img < 1 || C < 1 || cb * 8 < (dst_lo_off < 0 ? 3 * C : C) ||
      (fmt != MMSEG_FMT_BF16 && fmt != MMSEG_FMT_FP16))
    return fail(MMSEG_ERR_INVALID_ARG, "pack_ncdhw: bad arguments");
  const size_t nvox = (size_t)Z * Y * X;
  const int rows = n_img * cb;
  dim3 grid(grid_x_for(nvox, rows), rows);
  pack_ncdhw_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, dst, n_img, C, nvox, dst_cbt, dst_cb_off, dst_lo_off, cb, fmt);
  return check_launch("pack_ncdhw_kernel");
}

extern "C" int mmseg_pack_ncdhw_ex(const float* src, void* dst, int32_t n_img, int32_t C, int32_t Z, int32_t Y, int32_t X,
                                   int32_t dst_cbt, int32_t dst_cb_off, int32_t dst_lo_off, int32_t cb, int32_t pre_sigmoid,
                                   float pre_sub, float pre_mul, const float* gate_logits, int32_t fmt, void* stream) {
  if (!src || !dst || n_img < 1 || C < 1 || cb * 8 < C || (fmt != MMSEG_FMT_BF16 && fmt != MMSEG_FMT_FP16))
    return fail(MMSEG_ERR_INVALID_ARG, "pack_ncdhw_ex: bad arguments");
  const size_t nvox = (size_t)Z * Y * X;
  const int rows = n_img * cb;
  dim3 grid(grid_x_for(nvox, rows), rows);
  pack_ncdhw_ex_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, dst, n_img, C, nvox, dst_cbt, dst_cb_off, dst_lo_off, cb, pre_sigmoid ? 1 : 0, pre_sub, pre_mul, gate_logits,
      fmt);
  return check_launch("pack_ncdhw_ex_kernel");
}

extern "C" int mmseg_unpack_ncdhw(const void* src, float* dst, int32_t n_img, int32_t C, int32_t Z, int32_t Y,
                                  int32_t X, int32_t src_cbt, int32_t src_cb_off, int32_t src_lo_off, int32_t fmt,
                                  void* stream) {
  if (!src || !dst || n_img < 1 || C < 1 || (fmt != MMSEG_FMT_BF16 && fmt != MMSEG_FMT_FP16))
    return fail(MMSEG_ERR_INVALID_ARG, "unpack_ncdhw: bad arguments");
  const size_t nvox = (size_t)Z * Y * X;
  const int rows = n_img * ((C + 7) / 8);
  dim3 grid(grid_x_for(nvox, rows), rows);
  unpack_ncdhw_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, dst, n_img, C, nvox, src_cbt, src_cb_off, src_lo_off, fmt);
  return check_launch("unpack_ncdhw_kernel");
}
