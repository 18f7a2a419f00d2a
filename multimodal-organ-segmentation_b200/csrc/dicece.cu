// One-pass DiceCE forward: read logits + labels once, per-voxel softmax in registers, per-(batch, class) sums of
// p, p*onehot, onehot and the (weighted) negative log-likelihood; deterministic two-stage reduction (no atomics).
#include "common.h"
#include "ptx.cuh"

namespace mmseg {

constexpr int kMaxC = 16;

// grid: (n_blocks, B).  partial layout per (b, block): [3*C + 2] = I[c], P[c], T[c], nll_sum, weight_sum
template <int C>
__global__ void __launch_bounds__(256)
dicece_partial_kernel(const float* __restrict__ logits, const long long* __restrict__ target, size_t N,
                      const float* __restrict__ cw, float* __restrict__ partial) {
  const int b = blockIdx.y;
  const float* lg = logits + (size_t)b * C * N;
  const long long* tg = target + (size_t)b * N;
  float aI[C], aP[C], aT[C];
#pragma unroll
  for (int c = 0; c < C; ++c) { aI[c] = 0.f; aP[c] = 0.f; aT[c] = 0.f; }
  float nll = 0.f, wsum = 0.f;
  for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x) {
    float z[C];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) { z[c] = lg[(size_t)c * N + n]; mx = fmaxf(mx, z[c]); }
    const int t = (int)tg[n];
    float se = 0.f, zt = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float e = expf(z[c] - mx);
      if (c == t) zt = z[c] - mx;
      z[c] = e;
      se += e;
    }
    const float inv = 1.f / se;
    // a label outside [0, C) (ignore values such as -100 / 255) selects no class: no one-hot term and zero CE weight
    // (the reference raises in F.one_hot; here it must at least never index class_weights out of bounds)
    const bool tv = t >= 0 && t < C;
    const float w = tv ? (cw ? cw[t] : 1.f) : 0.f;
    nll += w * (logf(se) - zt);
    wsum += w;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float p = z[c] * inv;
      aP[c] += p;
      if (c == t) { aI[c] += p; aT[c] += 1.f; }
    }
  }
  __shared__ float red[8][3 * C + 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      aI[c] += __shfl_xor_sync(0xffffffffu, aI[c], o);
      aP[c] += __shfl_xor_sync(0xffffffffu, aP[c], o);
      aT[c] += __shfl_xor_sync(0xffffffffu, aT[c], o);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nll += __shfl_xor_sync(0xffffffffu, nll, o);
    wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < C; ++c) { red[warp][c] = aI[c]; red[warp][C + c] = aP[c]; red[warp][2 * C + c] = aT[c]; }
    red[warp][3 * C] = nll;
    red[warp][3 * C + 1] = wsum;
  }
  __syncthreads();
  if (threadIdx.x < 3 * C + 2) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    partial[((size_t)b * gridDim.x + blockIdx.x) * (3 * C + 2) + threadIdx.x] = s;
  }
}

// single block: fixed-order fp64 reduction over blocks, then the loss.
__global__ void dicece_final_kernel(const float* __restrict__ partial, int B, int C, int n_blocks, float dice_w,
                                    float ce_w, float smooth, int include_bg, float* __restrict__ result,
                                    float* __restrict__ sums) {
  __shared__ double acc[4 * (3 * kMaxC + 2)];  // B <= 4 per pass handled by loop below
  __shared__ double dice_sum, nll_sum, w_sum;
  const int stride = 3 * C + 2;
  if (threadIdx.x == 0) { dice_sum = 0.0; nll_sum = 0.0; w_sum = 0.0; }
  __syncthreads();
  for (int b = 0; b < B; ++b) {
    if (threadIdx.x < stride) {
      double s = 0.0;
      for (int k = 0; k < n_blocks; ++k) s += (double)partial[((size_t)b * n_blocks + k) * stride + threadIdx.x];
      acc[threadIdx.x] = s;
      if (sums) sums[(size_t)b * stride + threadIdx.x] = (float)s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int c = include_bg ? 0 : 1; c < C; ++c) {
        const double I = acc[c], P = acc[C + c], T = acc[2 * C + c];
        dice_sum += 1.0 - (2.0 * I + smooth) / (P + T + smooth);
      }
      nll_sum += acc[3 * C];
      w_sum += acc[3 * C + 1];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int nc = include_bg ? C : C - 1;
    const double dice = dice_sum / (double)(B * nc);
    const double ce = nll_sum / w_sum;
    result[0] = (float)(dice_w * dice + ce_w * ce);
    result[1] = (float)dice;
    result[2] = (float)ce;
  }
}

// Backward: dz_c = go * [ dice_w * p_c (g_c - sum_k g_k p_k) + ce_w * w_t (p_c - t_c) / W ],
// g_c = -(2 t_c (U_c + s) - (2 I_c + s)) / (B nc (U_c + s)^2)   (losses.py:39-80,216-228)
template <int C>
__global__ void __launch_bounds__(256)
dicece_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ target, size_t N, int B,
                  const float* __restrict__ cw, const float* __restrict__ sums, float dice_w, float ce_w, float smooth,
                  int include_bg, const float* __restrict__ grad_out, float* __restrict__ dlogits) {
  const int b = blockIdx.y;
  const float go = grad_out ? grad_out[0] : 1.f;
  const int stride = 3 * C + 2;
  float gA[C], gB[C];
  const int nc = include_bg ? C : C - 1;
  float wtot = 0.f;
  for (int bb = 0; bb < B; ++bb) wtot += sums[(size_t)bb * stride + 3 * C + 1];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float I = sums[(size_t)b * stride + c];
    const float U = sums[(size_t)b * stride + C + c] + sums[(size_t)b * stride + 2 * C + c];
    const float den = U + smooth;
    const float k = 1.f / ((float)(B * nc) * den * den);
    const bool on = include_bg || c > 0;
    gA[c] = on ? -2.f * den * k : 0.f;         // multiplies t_c
    gB[c] = on ? (2.f * I + smooth) * k : 0.f;  // constant part
  }
  const float* lg = logits + (size_t)b * C * N;
  float* dl = dlogits + (size_t)b * C * N;
  const long long* tg = target + (size_t)b * N;
  for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x) {
    float z[C];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) { z[c] = lg[(size_t)c * N + n]; mx = fmaxf(mx, z[c]); }
    const int t = (int)tg[n];
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { z[c] = expf(z[c] - mx); se += z[c]; }
    const float inv = 1.f / se;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      z[c] *= inv;
      const float g = gB[c] + (c == t ? gA[c] : 0.f);
      dot = fmaf(g, z[c], dot);
    }
    const float wce = (t >= 0 && t < C) ? ce_w * (cw ? cw[t] : 1.f) / wtot : 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float g = gB[c] + (c == t ? gA[c] : 0.f);
      const float d = dice_w * z[c] * (g - dot) + wce * (z[c] - (c == t ? 1.f : 0.f));
      dl[(size_t)c * N + n] = go * d;
    }
  }
}

// ---------------------------------------------------------------------------------------------- Tversky / Focal
// TverskyLoss (src/trainer/losses.py:156-185): tv = (TP + s) / (TP + a FP + b FN + s) per (batch, class) with TP = I,
// FP = P - I, FN = T - I from the same one-pass sums as Dice; loss = mean(1 - tv).
__global__ void tversky_final_kernel(const float* __restrict__ partial, int B, int C, int n_blocks, float alpha, float beta,
                                     float smooth, float* __restrict__ result, float* __restrict__ sums) {
  __shared__ double acc[3 * kMaxC + 2];
  __shared__ double tot;
  const int stride = 3 * C + 2;
  if (threadIdx.x == 0) tot = 0.0;
  __syncthreads();
  for (int b = 0; b < B; ++b) {
    if (threadIdx.x < stride) {
      double s = 0.0;
      for (int k = 0; k < n_blocks; ++k) s += (double)partial[((size_t)b * n_blocks + k) * stride + threadIdx.x];
      acc[threadIdx.x] = s;
      if (sums) sums[(size_t)b * stride + threadIdx.x] = (float)s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int c = 0; c < C; ++c) {
        const double I = acc[c], P = acc[C + c], T = acc[2 * C + c];
        tot += 1.0 - (I + smooth) / (I + alpha * (P - I) + beta * (T - I) + smooth);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) result[0] = (float)(tot / (double)(B * C));
}

// dz_c = go * p_c (g_c - sum_k g_k p_k),  g_c = gB_c + t_c gA_c  with the Tversky coefficients
template <int C>
__global__ void __launch_bounds__(256)
tversky_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ target, size_t N, int B,
                   const float* __restrict__ sums, float alpha, float beta, float smooth,
                   const float* __restrict__ grad_out, float* __restrict__ dlogits) {
  const int b = blockIdx.y;
  const float go = grad_out ? grad_out[0] : 1.f;
  const int stride = 3 * C + 2;
  float gA[C], gB[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float I = sums[(size_t)b * stride + c], P = sums[(size_t)b * stride + C + c], T = sums[(size_t)b * stride + 2 * C + c];
    const float D = I + alpha * (P - I) + beta * (T - I) + smooth;
    const float k = 1.f / ((float)(B * C) * D * D);
    gA[c] = -(D - (I + smooth) * (1.f - alpha - beta)) * k;   // multiplies t_c
    gB[c] = (I + smooth) * alpha * k;
  }
  const float* lg = logits + (size_t)b * C * N;
  float* dl = dlogits + (size_t)b * C * N;
  const long long* tg = target + (size_t)b * N;
  for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x) {
    float z[C];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) { z[c] = lg[(size_t)c * N + n]; mx = fmaxf(mx, z[c]); }
    const int t = (int)tg[n];
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { z[c] = expf(z[c] - mx); se += z[c]; }
    const float inv = 1.f / se;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      z[c] *= inv;
      dot = fmaf(gB[c] + (c == t ? gA[c] : 0.f), z[c], dot);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) dl[(size_t)c * N + n] = go * z[c] * (gB[c] + (c == t ? gA[c] : 0.f) - dot);
  }
}

// FocalLoss (src/trainer/losses.py:106-125): ce = -w_t log p_t, pt = exp(-ce), loss = mean((1 - pt)^gamma ce).
// grid (n_blocks, 1): per-block partial sums over ALL B*N voxels (logits [B][C][N]).
template <int C, bool BWD>
__global__ void __launch_bounds__(256)
focal_kernel(const float* __restrict__ logits, const long long* __restrict__ target, int B, size_t N,
             const float* __restrict__ cw, float gamma, const float* __restrict__ grad_out, float* __restrict__ partial,
             float* __restrict__ dlogits) {
  const size_t total = (size_t)B * N;
  const float go = BWD ? (grad_out ? grad_out[0] : 1.f) / (float)total : 0.f;
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / N, n = i - b * N;
    const float* lg = logits + b * C * N + n;
    float z[C];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) { z[c] = lg[(size_t)c * N]; mx = fmaxf(mx, z[c]); }
    const int t = (int)target[i];
    float se = 0.f, zt = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { if (c == t) zt = z[c] - mx; z[c] = expf(z[c] - mx); se += z[c]; }
    const float w = (t >= 0 && t < C) ? (cw ? cw[t] : 1.f) : 0.f;   // out-of-range label: ignored (never read out of bounds)
    const float ce = w * (logf(se) - zt);     // = -w log p_t
    const float pt = expf(-ce);
    const float om = 1.f - pt;
    if (!BWD) {
      acc += powf(om, gamma) * ce;
    } else {
      // d/dce [(1-pt)^g ce] = (1-pt)^g + g (1-pt)^(g-1) pt ce ;   dce/dz_c = w (p_c - [c == t])
      const float dfdce = powf(om, gamma) + (om > 0.f ? gamma * powf(om, gamma - 1.f) * pt * ce : 0.f);
      const float inv = 1.f / se;
      float* dl = dlogits + b * C * N + n;
#pragma unroll
      for (int c = 0; c < C; ++c) dl[(size_t)c * N] = go * dfdce * w * (z[c] * inv - (c == t ? 1.f : 0.f));
    }
  }
  if (!BWD) {
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w2 = 0; w2 < 8; ++w2) t += red[w2];
      partial[blockIdx.x] = t;
    }
  }
}

__global__ void focal_final_kernel(const float* __restrict__ partial, int n_blocks, double inv_total, float* __restrict__ result) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < n_blocks; ++k) s += (double)partial[k];
    result[0] = (float)(s * inv_total);
  }
}

// One pass over (prediction, target) label maps -> K x K confusion counts (rows = target, columns = prediction).
// Integer atomics only (associative -> deterministic).  DiceMetric / ConfusionMatrix (src/trainer/metrics.py:42-65,
// 184-196 — the latter a per-voxel Python loop in the reference) are read off this matrix.
template <typename PT>
__global__ void __launch_bounds__(256)
confusion_kernel(const PT* __restrict__ pred, const long long* __restrict__ target, size_t N, int K,
                 unsigned long long* __restrict__ counts) {
  extern __shared__ unsigned int hist[];  // K*K
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)pred[n], t = (int)target[n];
    if (p >= 0 && p < K && t >= 0 && t < K) atomicAdd(&hist[t * K + p], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K; i += blockDim.x)
    if (hist[i]) atomicAdd(&counts[i], (unsigned long long)hist[i]);
}

}  // namespace mmseg

using namespace mmseg;

extern "C" int mmseg_confusion_hist(const void* pred, int32_t pred_is_u8, const int64_t* target, int64_t N, int32_t K,
                                    uint64_t* counts, void* stream) {
  if (!pred || !target || !counts || N < 1 || K < 1 || K > 64)
    return fail(MMSEG_ERR_INVALID_ARG, "confusion_hist: bad arguments");
  int64_t nb = (N + 255) / 256;
  if (nb > 148 * 8) nb = 148 * 8;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t sh = (size_t)K * K * sizeof(unsigned int);
  const long long* tg = reinterpret_cast<const long long*>(target);
  unsigned long long* c = reinterpret_cast<unsigned long long*>(counts);
  // blocks cover at most 2^32 voxels each in a 32-bit smem counter: N / nb < 2^32 always holds here
  if (pred_is_u8) confusion_kernel<uint8_t><<<(unsigned)nb, 256, sh, st>>>(reinterpret_cast<const uint8_t*>(pred), tg, (size_t)N, K, c);
  else confusion_kernel<long long><<<(unsigned)nb, 256, sh, st>>>(reinterpret_cast<const long long*>(pred), tg, (size_t)N, K, c);
  return check_launch("confusion_kernel");
}

extern "C" int mmseg_dicece_fwd(const float* logits, const int64_t* target, int32_t B, int32_t C, int64_t N,
                                float dice_weight, float ce_weight, float smooth, int32_t include_background,
                                const float* class_weights, float* partial, int32_t n_blocks, float* result,
                                float* sums, void* stream) {
  if (!logits || !target || !partial || !result || B < 1 || N < 1 || n_blocks < 1 || n_blocks > 65535)
    return fail(MMSEG_ERR_INVALID_ARG, "dicece_fwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(n_blocks, B);
  const long long* tg = reinterpret_cast<const long long*>(target);
  switch (C) {
    case 2: dicece_partial_kernel<2><<<grid, 256, 0, st>>>(logits, tg, (size_t)N, class_weights, partial); break;
    case 3: dicece_partial_kernel<3><<<grid, 256, 0, st>>>(logits, tg, (size_t)N, class_weights, partial); break;
    case 4: dicece_partial_kernel<4><<<grid, 256, 0, st>>>(logits, tg, (size_t)N, class_weights, partial); break;
    case 8: dicece_partial_kernel<8><<<grid, 256, 0, st>>>(logits, tg, (size_t)N, class_weights, partial); break;
    case 16: dicece_partial_kernel<16><<<grid, 256, 0, st>>>(logits, tg, (size_t)N, class_weights, partial); break;
    default: return fail(MMSEG_ERR_UNSUPPORTED, "dicece_fwd: C=%d (supported: 2,3,4,8,16)", C);
  }
  int rc = check_launch("dicece_partial_kernel");
  if (rc != MMSEG_OK) return rc;
  dicece_final_kernel<<<1, 64, 0, st>>>(partial, B, C, n_blocks, dice_weight, ce_weight, smooth, include_background,
                                        result, sums);
  return check_launch("dicece_final_kernel");
}

extern "C" int mmseg_dicece_bwd(const float* logits, const int64_t* target, int32_t B, int32_t C, int64_t N,
                                float dice_weight, float ce_weight, float smooth, int32_t include_background,
                                const float* class_weights, const float* sums, const float* grad_out, float* dlogits,
                                void* stream) {
  if (!logits || !target || !sums || !dlogits || B < 1 || N < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "dicece_bwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int64_t nb = (N + 255) / 256;
  if (nb > 148 * 8) nb = 148 * 8;
  dim3 grid((unsigned)nb, B);
  const long long* tg = reinterpret_cast<const long long*>(target);
#define MMSEG_BWD(CC)                                                                                              \
  dicece_bwd_kernel<CC><<<grid, 256, 0, st>>>(logits, tg, (size_t)N, B, class_weights, sums, dice_weight, ce_weight, \
                                              smooth, include_background, grad_out, dlogits)
  switch (C) {
    case 2: MMSEG_BWD(2); break;
    case 3: MMSEG_BWD(3); break;
    case 4: MMSEG_BWD(4); break;
    case 8: MMSEG_BWD(8); break;
    case 16: MMSEG_BWD(16); break;
    default: return fail(MMSEG_ERR_UNSUPPORTED, "dicece_bwd: C=%d (supported: 2,3,4,8,16)", C);
  }
#undef MMSEG_BWD
  return check_launch("dicece_bwd_kernel");
}

extern "C" int mmseg_tversky_fwd(const float* logits, const int64_t* target, int32_t B, int32_t C, int64_t N, float alpha,
                                 float beta, float smooth, float* partial, int32_t n_blocks, float* result, float* sums,
                                 void* stream) {
  if (!logits || !target || !partial || !result || !sums || B < 1 || N < 1 || n_blocks < 1 || n_blocks > 65535)
    return fail(MMSEG_ERR_INVALID_ARG, "tversky_fwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(n_blocks, B);
  const long long* tg = reinterpret_cast<const long long*>(target);
  switch (C) {
    case 2: dicece_partial_kernel<2><<<grid, 256, 0, st>>>(logits, tg, (size_t)N, nullptr, partial); break;
    case 3: dicece_partial_kernel<3><<<grid, 256, 0, st>>>(logits, tg, (size_t)N, nullptr, partial); break;
    case 4: dicece_partial_kernel<4><<<grid, 256, 0, st>>>(logits, tg, (size_t)N, nullptr, partial); break;
    case 8: dicece_partial_kernel<8><<<grid, 256, 0, st>>>(logits, tg, (size_t)N, nullptr, partial); break;
    case 16: dicece_partial_kernel<16><<<grid, 256, 0, st>>>(logits, tg, (size_t)N, nullptr, partial); break;
    default: return fail(MMSEG_ERR_UNSUPPORTED, "tversky_fwd: C=%d (supported: 2,3,4,8,16)", C);
  }
  int rc = check_launch("dicece_partial_kernel");
  if (rc != MMSEG_OK) return rc;
  tversky_final_kernel<<<1, 64, 0, st>>>(partial, B, C, n_blocks, alpha, beta, smooth, result, sums);
  return check_launch("tversky_final_kernel");
}

extern "C" int mmseg_tversky_bwd(const float* logits, const int64_t* target, int32_t B, int32_t C, int64_t N, float alpha,
                                 float beta, float smooth, const float* sums, const float* grad_out, float* dlogits,
                                 void* stream) {
  if (!logits || !target || !sums || !dlogits || B < 1 || N < 1) return fail(MMSEG_ERR_INVALID_ARG, "tversky_bwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int64_t nb = (N + 255) / 256;
  if (nb > 148 * 8) nb = 148 * 8;
  dim3 grid((unsigned)nb, B);
  const long long* tg = reinterpret_cast<const long long*>(target);
#define MMSEG_TV(CC) tversky_bwd_kernel<CC><<<grid, 256, 0, st>>>(logits, tg, (size_t)N, B, sums, alpha, beta, smooth, grad_out, dlogits)
  switch (C) {
    case 2: MMSEG_TV(2); break;
    case 3: MMSEG_TV(3); break;
    case 4: MMSEG_TV(4); break;
    case 8: MMSEG_TV(8); break;
    case 16: MMSEG_TV(16); break;
    default: return fail(MMSEG_ERR_UNSUPPORTED, "tversky_bwd: C=%d (supported: 2,3,4,8,16)", C);
  }
#undef MMSEG_TV
  return check_launch("tversky_bwd_kernel");
}

extern "C" int mmseg_focal(const float* logits, const int64_t* target, int32_t B, int32_t C, int64_t N,
                           const float* class_weights, float gamma, float* partial, int32_t n_blocks, float* result,
                           const float* grad_out, float* dlogits, void* stream) {
  if (!logits || !target || B < 1 || N < 1 || n_blocks < 1) return fail(MMSEG_ERR_INVALID_ARG, "focal: bad arguments");
  if (!dlogits && (!partial || !result)) return fail(MMSEG_ERR_INVALID_ARG, "focal: forward needs partial and result");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long* tg = reinterpret_cast<const long long*>(target);
#define MMSEG_FOCAL(CC)                                                                                                  \
  if (dlogits) focal_kernel<CC, true><<<n_blocks, 256, 0, st>>>(logits, tg, B, (size_t)N, class_weights, gamma, grad_out, partial, dlogits); \
  else focal_kernel<CC, false><<<n_blocks, 256, 0, st>>>(logits, tg, B, (size_t)N, class_weights, gamma, grad_out, partial, dlogits)
  switch (C) {
    case 2: MMSEG_FOCAL(2); break;
    case 3: MMSEG_FOCAL(3); break;
    case 4: MMSEG_FOCAL(4); break;
    case 8: MMSEG_FOCAL(8); break;
    case 16: MMSEG_FOCAL(16); break;
    default: return fail(MMSEG_ERR_UNSUPPORTED, "focal: C=%d (supported: 2,3,4,8,16)", C);
  }
#undef MMSEG_FOCAL
  int rc = check_launch("focal_kernel");
  if (rc != MMSEG_OK || dlogits) return rc;
  focal_final_kernel<<<1, 32, 0, st>>>(partial, n_blocks, 1.0 / ((double)B * (double)N), result);
  return check_launch("focal_final_kernel");
}
