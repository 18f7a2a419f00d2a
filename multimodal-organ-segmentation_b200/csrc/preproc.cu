// Inference-edge preprocessing on the device: the reference normalises intensities per modality on
// the host with numpy (src/data/transforms.py:362-404, ModalitySpecificNormalize) before a volume reaches the model; at
// ~0.5 s of GPU time per 512x512x300 volume that host pass (three numpy sweeps over 629 MB) would dominate, so it runs
// here as two HBM-bound kernels: a deterministic per-channel statistics pass (max, sum, sum of squares; fixed-order
// two-stage reduction, fp64 combine) and an apply pass.  Resize(order=1) (transforms.py:215-250, scipy.ndimage.zoom ==
// trilinear interpolation with aligned corners) is mmseg_trilinear_resize (fusion.cu).
#include "common.h"

namespace mmseg {

int num_sms();

// partial[c][block][3] = (max, sum, sum of squares) of this block's slice of channel c
__global__ void __launch_bounds__(256)
channel_stats_partial_kernel(const float* __restrict__ vol, size_t nvox, float* __restrict__ partial) {
  const int c = blockIdx.y;
  const float* p = vol + (size_t)c * nvox;
  float mx = -INFINITY;
  double s1 = 0.0, s2 = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += stride) {
    const float v = p[i];
    mx = fmaxf(mx, v);
    s1 += (double)v;
    s2 += (double)v * (double)v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  __shared__ float smx[8];
  __shared__ double ss1[8], ss2[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { smx[warp] = mx; ss1[warp] = s1; ss2[warp] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { mx = fmaxf(mx, smx[w]); s1 += ss1[w]; s2 += ss2[w]; }
    double* out = reinterpret_cast<double*>(partial) + ((size_t)c * gridDim.x + blockIdx.x) * 3;
    out[0] = (double)mx; out[1] = s1; out[2] = s2;
  }
}

// stats[c] = (max, mean, population std) — fixed order over the blocks
__global__ void channel_stats_final_kernel(const double* __restrict__ partial, int n_blocks, double inv_n,
                                           float* __restrict__ stats) {
  const int c = blockIdx.x;
  if (threadIdx.x != 0) return;
  double mx = -INFINITY, s1 = 0.0, s2 = 0.0;
  for (int b = 0; b < n_blocks; ++b) {
    const double* q = partial + ((size_t)c * n_blocks + b) * 3;
    mx = fmax(mx, q[0]); s1 += q[1]; s2 += q[2];
  }
  const double mean = s1 * inv_n;
  double var = s2 * inv_n - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[c * 3 + 0] = (float)mx;
  stats[c * 3 + 1] = (float)mean;
  stats[c * 3 + 2] = (float)sqrt(var);
}

// kind[c]: 0 copy, 1 CT window: clip(v, a, b) then (v - a) / (b - a); 2 PET: v / max when max > 0;
//          3 z-score: (v - mean) / (std + 1e-8)      (reference transforms.py:380-401, fp32 arithmetic as numpy does)
__global__ void __launch_bounds__(256)
modality_normalize_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t nvox, const int* __restrict__ kind,
                          const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ stats) {
  const int c = blockIdx.y;
  const int k = kind[c];
  const float lo = a[c], hi = b[c];
  const float mx = stats[c * 3], mean = stats[c * 3 + 1], sd = stats[c * 3 + 2] + 1e-8f;
  const float* p = src + (size_t)c * nvox;
  float* q = dst + (size_t)c * nvox;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += stride) {
    float v = p[i];
    if (k == 1) {
      v = fminf(fmaxf(v, lo), hi);
      v = __fdiv_rn(__fsub_rn(v, lo), __fsub_rn(hi, lo));
    } else if (k == 2) {
      if (mx > 0.f) v = __fdiv_rn(v, mx);
    } else if (k == 3) {
      v = __fdiv_rn(__fsub_rn(v, mean), sd);
    }
    q[i] = v;
  }
}

}  // namespace mmseg

using namespace mmseg;

extern "C" int mmseg_channel_stats(const float* vol, int32_t C, int64_t voxels, void* partial, int32_t n_blocks,
                                   float* stats, void* stream) {
  if (!vol || !partial || !stats || C < 1 || C > 65535 || voxels < 1 || n_blocks < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "channel_stats: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)n_blocks, (unsigned)C);
  channel_stats_partial_kernel<<<grid, 256, 0, st>>>(vol, (size_t)voxels, reinterpret_cast<float*>(partial));
  int rc = check_launch("channel_stats_partial_kernel");
  if (rc) return rc;
  channel_stats_final_kernel<<<C, 32, 0, st>>>(reinterpret_cast<const double*>(partial), n_blocks, 1.0 / (double)voxels, stats);
  return check_launch("channel_stats_final_kernel");
}

extern "C" int mmseg_modality_normalize(const float* src, float* dst, int32_t C, int64_t voxels, const int32_t* kind,
                                        const float* a, const float* b, const float* stats, void* stream) {
  if (!src || !dst || !kind || !a || !b || !stats || C < 1 || C > 65535 || voxels < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "modality_normalize: bad arguments");
  int64_t blocks = (voxels + 255) / 256;
  const int64_t cap = ((int64_t)num_sms() * 16 + C - 1) / C;
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks, (unsigned)C);
  modality_normalize_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, dst, (size_t)voxels, kind, a, b, stats);
  return check_launch("modality_normalize_kernel");
}
