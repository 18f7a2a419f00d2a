// Modality fusion kernels for DualEncoder: per-channel global average pool (deterministic two-stage), the SE-style
// gate MLP + softmax over modalities, the weighted sum over modalities (mean / add / gate), and a stand-alone
// MaxPool3d(2) on blocked tensors.  All HBM-bound, 16-byte vectors.
#include "common.h"
#include "ptx.cuh"

namespace mmseg {

int num_sms();

// 16-byte (8-channel) vector accessors in the blocked buffer's element format (bf16 / fp16), 2-byte element offsets
__device__ __forceinline__ void ld8_e(const uint16_t* p, float* v, bool fp16) {
  cvt8_to_f32(*reinterpret_cast<const uint4*>(p), v, fp16);
}

// grid (n_chunks, n_img*cb): per-chunk partial sums of 8 channels -> partial[(blk*n_chunks + chunk)*8 + i]
__global__ void __launch_bounds__(256)
channel_sum_partial_kernel(const uint16_t* __restrict__ src, int src_cbt, int cb_off, int lo_off, int cb,
                           size_t nvox, float* __restrict__ partial, bool fp16) {
  const int blk = blockIdx.y;
  const int img = blk / cb, c = blk - img * cb;
  const size_t base = (size_t)(img * src_cbt + cb_off + c) * nvox * 8;
  const size_t lo_delta = (size_t)lo_off * nvox * 8;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (size_t)gridDim.x * blockDim.x) {
    float x[8];
    ld8_e(src + base + v * 8, x, fp16);
    if (lo_delta) {
      float l[8];
      ld8_e(src + base + lo_delta + v * 8, l, fp16);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] += l[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] += x[i];
  }
  __shared__ float red[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][i] = s[i];
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    partial[((size_t)blk * gridDim.x + blockIdx.x) * 8 + threadIdx.x] = t;
  }
}

__global__ void channel_mean_final_kernel(const float* __restrict__ partial, int n_rows /*n_img*cb*/, int n_chunks,
                                          double inv_n, float* __restrict__ mean) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over n_rows*8
  if (idx >= n_rows * 8) return;
  const int blk = idx >> 3, i = idx & 7;
  double s = 0.0;
  for (int k = 0; k < n_chunks; ++k) s += (double)partial[((size_t)blk * n_chunks + k) * 8 + i];
  mean[idx] = (float)(s * inv_n);
}

// one block per image: h = relu(W1 p + b1); logits = W2 h + b2; softmax over M  (reference dual_encoder.py:226-254)
__global__ void gate_mlp_kernel(const float* __restrict__ pooled, const float* __restrict__ w1,
                                const float* __restrict__ b1, const float* __restrict__ w2,
                                const float* __restrict__ b2, int MC, int H, int M, float* __restrict__ weights) {
  extern __shared__ float sh[];  // H hidden + M logits
  float* hid = sh;
  float* lg = sh + H;
  const int img = blockIdx.x;
  const float* p = pooled + (size_t)img * MC;
  // one warp per hidden unit, lanes along the M*C inputs (coalesced reads of the W1 row, fixed-order shuffle reduction)
  for (int j = threadIdx.x >> 5; j < H; j += blockDim.x >> 5) {
    const float* w = w1 + (size_t)j * MC;
    float acc = 0.f;
    for (int k = threadIdx.x & 31; k < MC; k += 32) acc = fmaf(w[k], p[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    acc += b1[j];
    if ((threadIdx.x & 31) == 0) hid[j] = acc > 0.f ? acc : 0.f;
  }
  __syncthreads();
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    float acc = b2[m];
    const float* w = w2 + (size_t)m * H;
    for (int k = 0; k < H; ++k) acc = fmaf(w[k], hid[k], acc);
    lg[m] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = -INFINITY;
    for (int m = 0; m < M; ++m) mx = fmaxf(mx, lg[m]);
    float se = 0.f;
    for (int m = 0; m < M; ++m) se += expf(lg[m] - mx);
    for (int m = 0; m < M; ++m) weights[(size_t)img * M + m] = expf(lg[m] - mx) / se;
  }
}

// Backward of the gate MLP (autograd through nn.Linear - ReLU - nn.Linear - Softmax of CrossModalAttention.attention,
// dual_encoder.py:226-233), given dweights [n_img, M] = d loss / d softmax output.  Two launches:
//   _hidden (one block per image): recompute h and the softmax, ds = w * (dw - sum_m w dw), dh = (h > 0) * W2^T ds,
//            dpooled = W1^T dh; keeps h, dh, ds for
//   _params (grid over the weight elements): dW1 = sum_img dh p^T, db1 = sum_img dh, dW2 = sum_img ds h^T, db2 = sum_img ds.
__global__ void gate_mlp_bwd_hidden_kernel(const float* __restrict__ pooled, const float* __restrict__ w1,
                                           const float* __restrict__ b1, const float* __restrict__ w2,
                                           const float* __restrict__ b2, const float* __restrict__ dweights, int MC, int H,
                                           int M, float* __restrict__ hid_out, float* __restrict__ dh_out,
                                           float* __restrict__ ds_out, float* __restrict__ dpooled) {
  extern __shared__ float sh[];  // H hidden | H dh | M logits | M ds
  float* hid = sh;
  float* dh = sh + H;
  float* lg = sh + 2 * H;
  float* ds = lg + M;
  const int img = blockIdx.x;
  const float* p = pooled + (size_t)img * MC;
  // one warp per hidden unit, lanes along the M*C inputs (coalesced reads of the W1 row, fixed-order shuffle reduction)
  for (int j = threadIdx.x >> 5; j < H; j += blockDim.x >> 5) {
    const float* w = w1 + (size_t)j * MC;
    float acc = 0.f;
    for (int k = threadIdx.x & 31; k < MC; k += 32) acc = fmaf(w[k], p[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    acc += b1[j];
    if ((threadIdx.x & 31) == 0) hid[j] = acc > 0.f ? acc : 0.f;
  }
  __syncthreads();
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    float acc = b2[m];
    const float* w = w2 + (size_t)m * H;
    for (int k = 0; k < H; ++k) acc = fmaf(w[k], hid[k], acc);
    lg[m] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = -INFINITY;
    for (int m = 0; m < M; ++m) mx = fmaxf(mx, lg[m]);
    float se = 0.f;
    for (int m = 0; m < M; ++m) se += expf(lg[m] - mx);
    float dot = 0.f;
    for (int m = 0; m < M; ++m) {
      lg[m] = expf(lg[m] - mx) / se;                  // softmax weight
      dot = fmaf(lg[m], dweights[(size_t)img * M + m], dot);
    }
    for (int m = 0; m < M; ++m) {
      ds[m] = lg[m] * (dweights[(size_t)img * M + m] - dot);
      ds_out[(size_t)img * M + m] = ds[m];
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    float acc = 0.f;
    for (int m = 0; m < M; ++m) acc = fmaf(ds[m], w2[(size_t)m * H + j], acc);
    const float g = hid[j] > 0.f ? acc : 0.f;
    dh[j] = g;
    dh_out[(size_t)img * H + j] = g;
    hid_out[(size_t)img * H + j] = hid[j];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < MC; k += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < H; ++j) acc = fmaf(dh[j], w1[(size_t)j * MC + k], acc);
    dpooled[(size_t)img * MC + k] = acc;
  }
}

__global__ void __launch_bounds__(256)
gate_mlp_bwd_params_kernel(const float* __restrict__ pooled, const float* __restrict__ hid, const float* __restrict__ dh,
                           const float* __restrict__ ds, int n_img, int MC, int H, int M, float* __restrict__ dw1,
                           float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2) {
  const size_t n1 = (size_t)H * MC, n2 = (size_t)M * H;
  const size_t total = n1 + H + n2 + M;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    if (i < n1) {
      const int j = (int)(i / MC), k = (int)(i - (size_t)j * MC);
      for (int b = 0; b < n_img; ++b) acc = fmaf(dh[(size_t)b * H + j], pooled[(size_t)b * MC + k], acc);
      dw1[i] = acc;
    } else if (i < n1 + H) {
      const int j = (int)(i - n1);
      for (int b = 0; b < n_img; ++b) acc += dh[(size_t)b * H + j];
      db1[j] = acc;
    } else if (i < n1 + H + n2) {
      const size_t r = i - n1 - H;
      const int m = (int)(r / H), j = (int)(r - (size_t)m * H);
      for (int b = 0; b < n_img; ++b) acc = fmaf(ds[(size_t)b * M + m], hid[(size_t)b * H + j], acc);
      dw2[r] = acc;
    } else {
      const int m = (int)(i - n1 - H - n2);
      for (int b = 0; b < n_img; ++b) acc += ds[(size_t)b * M + m];
      db2[m] = acc;
    }
  }
}

// dst[b, c] = sum_m w[b, m] * src[b, m*cb + c]; grid (chunks, n_img*cb)
__global__ void __launch_bounds__(256)
modality_combine_kernel(const uint16_t* __restrict__ src, int src_cbt, int src_lo_off, int M, int cb, size_t nvox,
                        const float* __restrict__ weights, float uniform_w, uint16_t* __restrict__ dst,
                        int dst_cbt, int dst_cb_off, int dst_lo_off, bool fp16) {
  const int blk = blockIdx.y;
  const int img = blk / cb, c = blk - img * cb;
  const size_t src_lo = (size_t)src_lo_off * nvox * 8;
  const size_t dst_base = (size_t)(img * dst_cbt + dst_cb_off + c) * nvox * 8;
  const size_t dst_lo = (size_t)dst_lo_off * nvox * 8;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (size_t)gridDim.x * blockDim.x) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int m = 0; m < M; ++m) {
      const float w = weights ? weights[(size_t)img * M + m] : uniform_w;
      const size_t base = (size_t)(img * src_cbt + m * cb + c) * nvox * 8 + v * 8;
      float x[8];
      ld8_e(src + base, x, fp16);
      if (src_lo) {
        float l[8];
        ld8_e(src + base + src_lo, l, fp16);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] += l[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, x[i], acc[i]);
    }
    store8_act(dst, dst_base + v * 8, dst_lo, acc, fp16);
  }
}

// dst[b, c] = max_m src[b, m*cb + c]  (LateFusion 'max', src/models/fusion/late_fusion.py:62-64); grid (chunks, n_img*cb)
__global__ void __launch_bounds__(256)
modality_max_kernel(const uint16_t* __restrict__ src, int src_cbt, int M, int cb, size_t nvox,
                    uint16_t* __restrict__ dst, int dst_cbt, int dst_cb_off) {
  const int blk = blockIdx.y;
  const int img = blk / cb, c = blk - img * cb;
  const size_t dst_base = (size_t)(img * dst_cbt + dst_cb_off + c) * nvox * 8;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (size_t)gridDim.x * blockDim.x) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = -INFINITY;
    for (int m = 0; m < M; ++m) {
      float x[8];
      ld8_e(src + (size_t)(img * src_cbt + m * cb + c) * nvox * 8 + v * 8, x, false);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaxf(acc[i], x[i]);
    }
    store8_act(dst, dst_base + v * 8, 0, acc, false);
  }
}

// MaxPool3d(2) on a blocked tensor (stand-alone DownBlock3D use); grid (chunks, n_img*cb)
__global__ void __launch_bounds__(256)
maxpool2_kernel(const uint16_t* __restrict__ src, int src_cbt, int src_cb_off, int src_lo_off, int cb, int Z, int Y,
                int X, uint16_t* __restrict__ dst, int dst_cbt, int dst_cb_off, int dst_lo_off, bool fp16) {
  const int blk = blockIdx.y;
  const int img = blk / cb, c = blk - img * cb;
  const int Zh = Z / 2, Yh = Y / 2, Xh = X / 2;
  const size_t nvox = (size_t)Z * Y * X, ncell = (size_t)Zh * Yh * Xh;
  const size_t sbase = (size_t)(img * src_cbt + src_cb_off + c) * nvox * 8;
  const size_t slo = (size_t)src_lo_off * nvox * 8;
  const size_t dbase = (size_t)(img * dst_cbt + dst_cb_off + c) * ncell * 8;
  const size_t dlo = (size_t)dst_lo_off * ncell * 8;
  for (size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x; cell < ncell;
       cell += (size_t)gridDim.x * blockDim.x) {
    const int xh = (int)(cell % Xh);
    const size_t r = cell / Xh;
    const int yh = (int)(r % Yh), zh = (int)(r / Yh);
    float mx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
    for (int d = 0; d < 8; ++d) {
      const size_t v = ((size_t)(2 * zh + (d >> 2)) * Y + (2 * yh + ((d >> 1) & 1))) * X + (2 * xh + (d & 1));
      float x[8];
      ld8_e(src + sbase + v * 8, x, fp16);
      if (slo) {
        float l[8];
        ld8_e(src + sbase + slo + v * 8, l, fp16);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] += l[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = fmaxf(mx[i], x[i]);
    }
    store8_act(dst, dbase + cell * 8, dlo, mx, fp16);
  }
}

static unsigned gx_for(size_t items, int rows) {
  size_t want = ((size_t)num_sms() * 8 + rows - 1) / rows;
  size_t need = (items + 255) / 256;
  size_t g = want < need ? want : need;
  return (unsigned)(g < 1 ? 1 : g);
}

// ------------------------------------------------------------------------------------------------ 1x1x1 logits conv
// out_conv = nn.Conv3d(features[0], num_classes, 1) (unet.py:163, dual_encoder.py:118): K = 32, N = 8 is far too thin
// for the tensor-core tile (N padded to 16, fp32 NCDHW scatter from 4 epilogue warps: 178 us per 8 windows) but it is
// only 512 FLOP per voxel, so it runs on the CUDA cores at HBM speed: each thread owns 4 consecutive voxels, reads their
// 8-channel blocks as 64 contiguous bytes, keeps COUT x 4 fp32 accumulators, and writes one float4 per class plane.
// The weights sit in shared memory as [cin][COUT] so one LDS.128 feeds 16 FMAs.  fp32 accumulation of exact bf16 (or
// hi + lo in parity mode) inputs with fp32 weights.
template <int COUT>
__global__ void __launch_bounds__(256) conv1x1_logits_kernel(const uint16_t* __restrict__ src, int src_cbt, int cb_off,
                                                            int lo_off, int cin_blocks, size_t nvox,
                                                            const float* __restrict__ weight, const float* __restrict__ bias,
                                                            int cout, float* __restrict__ dst, bool fp16) {
  extern __shared__ float wsm[];   // [cin][COUT], zero-padded classes
  const int cin = cin_blocks * 8;
  for (int i = threadIdx.x; i < cin * COUT; i += blockDim.x) {
    const int ci = i / COUT, co = i - ci * COUT;
    wsm[i] = co < cout ? weight[(size_t)co * cin + ci] : 0.f;
  }
  __syncthreads();
  const int img = blockIdx.y;
  const size_t v0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (v0 >= nvox) return;
  const bool full = v0 + 4 <= nvox;
  float acc[COUT][4];
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    const float b = (bias && co < cout) ? bias[co] : 0.f;
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[co][v] = b;
  }
  for (int cb = 0; cb < cin_blocks; ++cb) {
    const uint16_t* base = src + (((size_t)img * src_cbt + cb_off + cb) * nvox + v0) * 8;
    float x[4][8];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (full || v0 + v < nvox) {
        ld8_e(base + v * 8, x[v], fp16);
        if (lo_off > 0) {
          float l[8];
          ld8_e(base + (size_t)lo_off * nvox * 8 + v * 8, l, fp16);
#pragma unroll
          for (int j = 0; j < 8; ++j) x[v][j] += l[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[v][j] = 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4* wr = reinterpret_cast<const float4*>(wsm + (size_t)(cb * 8 + j) * COUT);
#pragma unroll
      for (int c4 = 0; c4 < COUT / 4; ++c4) {
        const float4 w = wr[c4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          acc[4 * c4 + 0][v] = fmaf(w.x, x[v][j], acc[4 * c4 + 0][v]);
          acc[4 * c4 + 1][v] = fmaf(w.y, x[v][j], acc[4 * c4 + 1][v]);
          acc[4 * c4 + 2][v] = fmaf(w.z, x[v][j], acc[4 * c4 + 2][v]);
          acc[4 * c4 + 3][v] = fmaf(w.w, x[v][j], acc[4 * c4 + 3][v]);
        }
      }
    }
  }
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    if (co < cout) {
      float* o = dst + ((size_t)img * cout + co) * nvox + v0;
      if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
        *reinterpret_cast<float4*>(o) = make_float4(acc[co][0], acc[co][1], acc[co][2], acc[co][3]);
      } else {
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (v0 + v < nvox) o[v] = acc[co][v];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ trilinear resize
// F.interpolate(x, size, mode="trilinear", align_corners=True) on NCDHW fp32 (DeepSupervisionHead, segmentation.py:
// 108-113; the logits of a coarse scale brought to the target size).  Same index arithmetic as ATen's
// upsample_trilinear3d: src = dst * (in-1)/(out-1) in fp32, i0 = (int)src, lambda1 = src - i0, i1 = i0 + (i0 < in-1).
__global__ void __launch_bounds__(256) trilinear_resize_kernel(const float* __restrict__ src, int Zi, int Yi, int Xi,
                                                              float* __restrict__ dst, int Zo, int Yo, int Xo,
                                                              float sz, float sy, float sx) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y % Yo, z = blockIdx.y / Yo;
  const size_t nc = blockIdx.z;
  if (x >= Xo) return;
  const float fz = sz * z, fy = sy * y, fx = sx * x;
  const int z0 = (int)fz, y0 = (int)fy, x0 = (int)fx;
  const int zp = z0 < Zi - 1 ? 1 : 0, yp = y0 < Yi - 1 ? 1 : 0, xp = x0 < Xi - 1 ? 1 : 0;
  const float z1l = fz - z0, y1l = fy - y0, x1l = fx - x0;
  const float z0l = 1.f - z1l, y0l = 1.f - y1l, x0l = 1.f - x1l;
  const float* p = src + ((nc * Zi + z0) * Yi + y0) * (size_t)Xi + x0;
  const size_t sY = (size_t)Xi * yp, sZ = (size_t)Yi * Xi * zp;
  const float v =
      z0l * (y0l * (x0l * p[0] + x1l * p[xp]) + y1l * (x0l * p[sY] + x1l * p[sY + xp])) +
      z1l * (y0l * (x0l * p[sZ] + x1l * p[sZ + xp]) + y1l * (x0l * p[sZ + sY] + x1l * p[sZ + sY + xp]));
  dst[((nc * Zo + z) * Yo + y) * (size_t)Xo + x] = v;
}

}  // namespace mmseg

using namespace mmseg;

extern "C" int mmseg_trilinear_resize(const float* src, int32_t n_planes, int32_t Zi, int32_t Yi, int32_t Xi, float* dst,
                                      int32_t Zo, int32_t Yo, int32_t Xo, void* stream) {
  if (!src || !dst || n_planes < 1 || Zi < 1 || Yi < 1 || Xi < 1 || Zo < 1 || Yo < 1 || Xo < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "trilinear_resize: bad arguments");
  if (n_planes > 65535 || (int64_t)Zo * Yo > 2147483647LL) return fail(MMSEG_ERR_INVALID_ARG, "trilinear_resize: grid too large");
  const float sz = Zo > 1 ? (float)(Zi - 1) / (float)(Zo - 1) : 0.f;
  const float sy = Yo > 1 ? (float)(Yi - 1) / (float)(Yo - 1) : 0.f;
  const float sx = Xo > 1 ? (float)(Xi - 1) / (float)(Xo - 1) : 0.f;
  const int bx = Xo >= 256 ? 256 : (Xo >= 128 ? 128 : (Xo >= 64 ? 64 : 32));
  dim3 grid((unsigned)((Xo + bx - 1) / bx), (unsigned)(Zo * Yo), (unsigned)n_planes);
  trilinear_resize_kernel<<<grid, bx, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, Zi, Yi, Xi, dst, Zo, Yo, Xo, sz, sy, sx);
  return check_launch("trilinear_resize_kernel");
}

extern "C" int mmseg_conv1x1_logits(const void* src, int32_t n_img, int32_t src_cbt, int32_t cb_off, int32_t lo_off,
                                    int32_t cin, int64_t voxels, const float* weight, const float* bias, int32_t cout,
                                    float* dst, int32_t fmt, void* stream) {
  if (!src || !weight || !dst || n_img < 1 || voxels < 1 || (fmt != MMSEG_FMT_BF16 && fmt != MMSEG_FMT_FP16))
    return fail(MMSEG_ERR_INVALID_ARG, "conv1x1_logits: bad arguments");
  if (cin < 8 || (cin % 8) || cin > 256) return fail(MMSEG_ERR_UNSUPPORTED, "conv1x1_logits: cin=%d (multiple of 8, <= 256)", cin);
  if (cout < 1 || cout > 16) return fail(MMSEG_ERR_UNSUPPORTED, "conv1x1_logits: cout=%d (1..16)", cout);
  if (n_img > 65535) return fail(MMSEG_ERR_INVALID_ARG, "conv1x1_logits: n_img");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t nv = (size_t)voxels;
  dim3 grid((unsigned)((nv + 1023) / 1024), (unsigned)n_img);
  const uint16_t* s = reinterpret_cast<const uint16_t*>(src);
  const bool fp16 = fmt == MMSEG_FMT_FP16;
  if (cout <= 4)
    conv1x1_logits_kernel<4><<<grid, 256, (size_t)cin * 4 * sizeof(float), st>>>(s, src_cbt, cb_off, lo_off, cin / 8, nv, weight, bias, cout, dst, fp16);
  else if (cout <= 8)
    conv1x1_logits_kernel<8><<<grid, 256, (size_t)cin * 8 * sizeof(float), st>>>(s, src_cbt, cb_off, lo_off, cin / 8, nv, weight, bias, cout, dst, fp16);
  else
    conv1x1_logits_kernel<16><<<grid, 256, (size_t)cin * 16 * sizeof(float), st>>>(s, src_cbt, cb_off, lo_off, cin / 8, nv, weight, bias, cout, dst, fp16);
  return check_launch("conv1x1_logits_kernel");
}

extern "C" int mmseg_channel_mean(const void* src, int32_t n_img, int32_t src_cbt, int32_t cb_off, int32_t lo_off,
                                  int32_t cb, int64_t voxels, float* partial, int32_t n_chunks, float* mean,
                                  int32_t fmt, void* stream) {
  if (!src || !partial || !mean || n_img < 1 || cb < 1 || voxels < 1 || n_chunks < 1 ||
      (fmt != MMSEG_FMT_BF16 && fmt != MMSEG_FMT_FP16))
    return fail(MMSEG_ERR_INVALID_ARG, "channel_mean: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(n_chunks, n_img * cb);
  channel_sum_partial_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint16_t*>(src), src_cbt, cb_off, lo_off, cb,
                                                   (size_t)voxels, partial, fmt == MMSEG_FMT_FP16);
  int rc = check_launch("channel_sum_partial_kernel");
  if (rc) return rc;
  const int n = n_img * cb * 8;
  channel_mean_final_kernel<<<(n + 127) / 128, 128, 0, st>>>(partial, n_img * cb, n_chunks, 1.0 / (double)voxels, mean);
  return check_launch("channel_mean_final_kernel");
}

extern "C" int mmseg_gate_mlp(const float* pooled, const float* w1, const float* b1, const float* w2, const float* b2,
                              int32_t n_img, int32_t MC, int32_t H, int32_t M, float* weights, void* stream) {
  if (!pooled || !w1 || !b1 || !w2 || !b2 || !weights || n_img < 1 || MC < 1 || H < 1 || M < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "gate_mlp: bad arguments");
  const size_t sh = (size_t)(H + M) * sizeof(float);
  if (sh > 48 * 1024) return fail(MMSEG_ERR_UNSUPPORTED, "gate_mlp: hidden size too large");
  gate_mlp_kernel<<<n_img, 1024, sh, reinterpret_cast<cudaStream_t>(stream)>>>(pooled, w1, b1, w2, b2, MC, H, M, weights);
  return check_launch("gate_mlp_kernel");
}

extern "C" int mmseg_gate_mlp_bwd(const float* pooled, const float* w1, const float* b1, const float* w2, const float* b2,
                                  const float* dweights, int32_t n_img, int32_t MC, int32_t H, int32_t M, float* workspace,
                                  float* dpooled, float* dw1, float* db1, float* dw2, float* db2, void* stream) {
  if (!pooled || !w1 || !b1 || !w2 || !b2 || !dweights || !workspace || !dpooled || !dw1 || !db1 || !dw2 || !db2 || n_img < 1 ||
      MC < 1 || H < 1 || M < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "gate_mlp_bwd: bad arguments");
  const size_t sh = (size_t)(2 * H + 2 * M) * sizeof(float);
  if (sh > 48 * 1024) return fail(MMSEG_ERR_UNSUPPORTED, "gate_mlp_bwd: hidden size too large");
  float* hid = workspace;                        // [n_img][H]
  float* dh = hid + (size_t)n_img * H;           // [n_img][H]
  float* ds = dh + (size_t)n_img * H;            // [n_img][M]
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  gate_mlp_bwd_hidden_kernel<<<n_img, 1024, sh, st>>>(pooled, w1, b1, w2, b2, dweights, MC, H, M, hid, dh, ds, dpooled);
  int rc = check_launch("gate_mlp_bwd_hidden_kernel");
  if (rc) return rc;
  const size_t total = (size_t)H * MC + H + (size_t)M * H + M;
  size_t blocks = (total + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  gate_mlp_bwd_params_kernel<<<(unsigned)blocks, 256, 0, st>>>(pooled, hid, dh, ds, n_img, MC, H, M, dw1, db1, dw2, db2);
  return check_launch("gate_mlp_bwd_params_kernel");
}

extern "C" int mmseg_modality_combine(const void* src, int32_t n_img, int32_t src_cbt, int32_t src_lo_off, int32_t M,
                                      int32_t cb, int64_t voxels, const float* weights, float uniform_weight, void* dst,
                                      int32_t dst_cbt, int32_t dst_cb_off, int32_t dst_lo_off, int32_t fmt, void* stream) {
  if (!src || !dst || n_img < 1 || M < 1 || cb < 1 || voxels < 1 || (fmt != MMSEG_FMT_BF16 && fmt != MMSEG_FMT_FP16))
    return fail(MMSEG_ERR_INVALID_ARG, "modality_combine: bad arguments");
  const int rows = n_img * cb;
  dim3 grid(gx_for((size_t)voxels, rows), rows);
  modality_combine_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint16_t*>(src), src_cbt, src_lo_off, M, cb, (size_t)voxels, weights, uniform_weight,
      reinterpret_cast<uint16_t*>(dst), dst_cbt, dst_cb_off, dst_lo_off, fmt == MMSEG_FMT_FP16);
  return check_launch("modality_combine_kernel");
}

extern "C" int mmseg_modality_max(const void* src, int32_t n_img, int32_t src_cbt, int32_t M, int32_t cb, int64_t voxels,
                                  void* dst, int32_t dst_cbt, int32_t dst_cb_off, void* stream) {
  if (!src || !dst || n_img < 1 || M < 1 || cb < 1 || voxels < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "modality_max: bad arguments");
  const int rows = n_img * cb;
  dim3 grid(gx_for((size_t)voxels, rows), rows);
  modality_max_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint16_t*>(src), src_cbt, M, cb, (size_t)voxels,
      reinterpret_cast<uint16_t*>(dst), dst_cbt, dst_cb_off);
  return check_launch("modality_max_kernel");
}

extern "C" int mmseg_maxpool3d_2(const void* src, int32_t n_img, int32_t src_cbt, int32_t src_cb_off,
                                 int32_t src_lo_off, int32_t cb, int32_t Z, int32_t Y, int32_t X, void* dst,
                                 int32_t dst_cbt, int32_t dst_cb_off, int32_t dst_lo_off, int32_t fmt, void* stream) {
  if (!src || !dst || n_img < 1 || cb < 1 || Z < 2 || Y < 2 || X < 2 || (fmt != MMSEG_FMT_BF16 && fmt != MMSEG_FMT_FP16))
    return fail(MMSEG_ERR_INVALID_ARG, "maxpool3d_2: bad arguments");
  const size_t ncell = (size_t)(Z / 2) * (Y / 2) * (X / 2);
  const int rows = n_img * cb;
  dim3 grid(gx_for(ncell, rows), rows);
  maxpool2_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint16_t*>(src), src_cbt, src_cb_off, src_lo_off, cb, Z, Y, X,
      reinterpret_cast<uint16_t*>(dst), dst_cbt, dst_cb_off, dst_lo_off, fmt == MMSEG_FMT_FP16);
  return check_launch("maxpool2_kernel");
}
