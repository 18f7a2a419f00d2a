// Trainer step glue: the kernels either side of forward / backward.
//   * mmseg_weights_repack  — fp32 PyTorch-layout parameter -> kernel-layout 16-bit GEMM operand of conv_tc.cu (forward
//     form, the spatially flipped / channel-transposed dgrad form, the ConvTranspose GEMM forms, hi / lo splits), in one
//     launch per weight: replaces the flip / permute / cat / copy chains ATen ran on every training step.
//   * mmseg_adamw_multi     — torch.optim.AdamW (reference src/trainer/trainer.py:115-117, stepped at :245-248) for a
//     whole list of parameter tensors in one launch (decoupled weight decay, bias correction, fp32 state).
// Both are HBM-bound streaming kernels; neither allocates memory (tables and state are caller-owned).
#include "common.h"
#include "ptx.cuh"

namespace mmseg {

// One thread = one (GEMM column n, 8-channel K group kg, filter tap) = one 16-byte vector of the packed operand.
// Threads run tap-fastest so that the reads of a warp walk the source's contiguous filter taps (the PyTorch layout keeps
// the k^3 taps of one (out, in) pair adjacent).
//   n_off[n]  : element offset of GEMM column n in the source (-1: zero padding column)
//   k_off[kk] : element offset of GEMM K index kk (16 per chunk, chunk-major) in the source (-1: zero padding)
//   source element = w[n_off[n] + k_off[kk] + tap_src],  tap_src = flip ? taps-1-tap : tap
// Destination layout (conv_tc.cu): [n_ntiles][n_kc_total][taps2d][2 k halves][rows][8], rows = KZ*NT with the z taps
// stored dz-DESCENDING (row = (KZ-1-dz)*NT + n_local); taps2d = 9 for k=3, 1 for k=1.
// Split modes: the hi part is written `hi_copies` times at chunk offsets 0, n_kc, ...; the lo part (w - hi) once after.
__device__ __forceinline__ void repack_one(const float* __restrict__ w, const int* __restrict__ n_off,
                                           const int* __restrict__ k_off, uint4* __restrict__ dst, int n_out, int NT, int n_kc,
                                           int n_kc_total, int ksize, int flip, int hi_copies, int has_lo, int fp16, float scale,
                                           long long idx) {
  const int taps = ksize * ksize * ksize;
  const long long total = (long long)n_out * n_kc * 2 * taps;
  if (idx >= total) return;
  const int tap = (int)(idx % taps);
  long long r = idx / taps;
  const int kg = (int)(r % (n_kc * 2));
  const int n = (int)(r / (n_kc * 2));
  const int kc = kg >> 1, half = kg & 1;
  const int tap_src = flip ? taps - 1 - tap : tap;
  const int no = n_off[n];
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ko = k_off[kg * 8 + j];
    v[j] = (no >= 0 && ko >= 0) ? w[(size_t)no + (size_t)ko + tap_src] * scale : 0.f;
  }
  const int nt = n / NT, nl = n - nt * NT;
  int tap2d, row, rows;
  if (ksize == 3) {
    const int dz = tap / 9;
    tap2d = tap - dz * 9;
    rows = 3 * NT;
    row = (2 - dz) * NT + nl;
  } else {
    tap2d = 0;
    rows = NT;
    row = nl;
  }
  const int taps2d = ksize == 3 ? 9 : 1;
  auto at = [&](int kc_dst) -> size_t {
    return ((((size_t)nt * n_kc_total + kc_dst) * taps2d + tap2d) * 2 + half) * rows + row;
  };
  if (!has_lo) {
    const uint4 hi = cvt8_from_f32(v, fp16 != 0);
    for (int c = 0; c < hi_copies; ++c) dst[at(kc + c * n_kc)] = hi;
  } else {
    uint4 hi, lo;
    split8_from_f32(v, hi, lo, fp16 != 0);
    for (int c = 0; c < hi_copies; ++c) dst[at(kc + c * n_kc)] = hi;
    dst[at(kc + hi_copies * n_kc)] = lo;
  }
}

__global__ void __launch_bounds__(256)
weights_repack_kernel(const float* __restrict__ w, const int* __restrict__ n_off, const int* __restrict__ k_off,
                      uint4* __restrict__ dst, int n_out, int NT, int n_kc, int n_kc_total, int ksize, int flip,
                      int hi_copies, int has_lo, int fp16, float scale) {
  repack_one(w, n_off, k_off, dst, n_out, NT, n_kc, n_kc_total, ksize, flip, hi_copies, has_lo, fp16, scale,
             (long long)blockIdx.x * blockDim.x + threadIdx.x);
}

// Every weight of a model in ONE launch (a training step repacks ~90 operands; as separate launches they cost 0.6 - 1.5 ms of
// launch latency): block b works on descriptor block_desc[b], at block offset b - first_block of that descriptor.
struct RepackDesc {
  const float* w;
  const int* n_off;
  const int* k_off;
  uint4* dst;
  int n_out, NT, n_kc, n_kc_total, ksize, flip, hi_copies, has_lo;
  long long first_block;
};
__global__ void __launch_bounds__(256)
weights_repack_multi_kernel(const RepackDesc* __restrict__ descs, const int* __restrict__ block_desc, int fp16, float scale) {
  const RepackDesc d = descs[block_desc[blockIdx.x]];
  repack_one(d.w, d.n_off, d.k_off, d.dst, d.n_out, d.NT, d.n_kc, d.n_kc_total, d.ksize, d.flip, d.hi_copies, d.has_lo, fp16,
             scale, ((long long)blockIdx.x - d.first_block) * blockDim.x + threadIdx.x);
}

// dst[i] = idx[i] >= 0 ? src[idx[i]] : 0  (bias expanded / padded to the GEMM columns)
__global__ void gather_f32_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = idx[i] >= 0 ? src[idx[i]] : 0.f;
}

// ---------------------------------------------------------------------------------------------- fused AdamW
// torch.optim.AdamW semantics (decoupled decay, no amsgrad, maximize = false):
//   p <- p * (1 - lr * wd);  m <- b1 m + (1 - b1) g;  v <- b2 v + (1 - b2) g^2;
//   p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Every tensor has its own DEVICE step counter (fp32; a small kernel increments them stream-ordered behind the update)
// and the hyper-parameters are a device array {lr, beta1, beta2, eps, weight_decay, grad_scale}, so the whole optimizer step
// can sit in a captured CUDA graph and still follow a learning-rate schedule.  One launch covers every tensor of the list: chunk c of the
// launch = (tensor, offset) from the caller-built chunk table.
struct AdamwTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* step;     // this tensor's own device-side step counter (fp32 scalar), like torch.optim.AdamW(capturable=True)
  long long n;
};
constexpr int kAdamChunk = 8192;   // elements per block

__global__ void __launch_bounds__(256)
adamw_multi_kernel(const AdamwTensor* __restrict__ tensors, const int2* __restrict__ chunks, int n_chunks,
                   const float* __restrict__ hyper, int zero_grad) {
  // hyper-parameters live on the device (lr changes under a scheduler; a captured graph must see the new value)
  const float lr = hyper[0], beta1 = hyper[1], beta2 = hyper[2], eps = hyper[3], wd = hyper[4], grad_scale = hyper[5];
  const int2 ch = chunks[blockIdx.x];
  const AdamwTensor t = tensors[ch.x];
  const float step = t.step[0] + 1.f;
  const float bc1 = 1.f - powf(beta1, step);
  const float bc2 = 1.f - powf(beta2, step);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const long long base = (long long)ch.y * kAdamChunk;
  const long long end = base + kAdamChunk < t.n ? base + kAdamChunk : t.n;
  const float decay = 1.f - lr * wd;
  const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) |
                     reinterpret_cast<uintptr_t>(t.v)) & 15) == 0;
  auto upd = [&](float& p, float g, float& m, float& v) {
    g *= grad_scale;
    p *= decay;
    m = beta1 * m + (1.f - beta1) * g;
    v = beta2 * v + (1.f - beta2) * g * g;
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    p -= step_size * (m / denom);
  };
  if (vec) {
    const long long b4 = base >> 2, e4 = end >> 2;   // base is a multiple of kAdamChunk (so of 4)
    for (long long i = b4 + threadIdx.x; i < e4; i += blockDim.x) {
      float4 p = reinterpret_cast<float4*>(t.p)[i];
      const float4 g = reinterpret_cast<const float4*>(t.g)[i];
      float4 m = reinterpret_cast<float4*>(t.m)[i];
      float4 v = reinterpret_cast<float4*>(t.v)[i];
      upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
      reinterpret_cast<float4*>(t.p)[i] = p;
      reinterpret_cast<float4*>(t.m)[i] = m;
      reinterpret_cast<float4*>(t.v)[i] = v;
      if (zero_grad) reinterpret_cast<float4*>(const_cast<float*>(t.g))[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (long long i = (e4 << 2) + threadIdx.x; i < end; i += blockDim.x) {
      upd(t.p[i], t.g[i], t.m[i], t.v[i]);
      if (zero_grad) const_cast<float*>(t.g)[i] = 0.f;
    }
  } else {
    for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
      upd(t.p[i], t.g[i], t.m[i], t.v[i]);
      if (zero_grad) const_cast<float*>(t.g)[i] = 0.f;
    }
  }
}

__global__ void adamw_step_inc_kernel(const AdamwTensor* __restrict__ tensors, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tensors[i].step[0] += 1.f;
}

}  // namespace mmseg

using namespace mmseg;

extern "C" int mmseg_weights_repack(const float* w, const int32_t* n_off, const int32_t* k_off, void* dst, int32_t n_out,
                                    int32_t NT, int32_t n_kc, int32_t n_kc_total, int32_t ksize, int32_t flip,
                                    int32_t hi_copies, int32_t has_lo, int32_t fmt, float scale, void* stream) {
  if (!w || !n_off || !k_off || !dst) return fail(MMSEG_ERR_INVALID_ARG, "weights_repack: null pointer");
  if (ksize != 1 && ksize != 3) return fail(MMSEG_ERR_UNSUPPORTED, "weights_repack: ksize %d (only 1, 3)", ksize);
  if (n_out < 16 || NT < 16 || (NT % 16) || (n_out % NT) || n_kc < 1 || hi_copies < 1 || hi_copies > 2 ||
      n_kc_total != n_kc * (hi_copies + (has_lo ? 1 : 0)))
    return fail(MMSEG_ERR_INVALID_ARG, "weights_repack: bad shape (n_out=%d NT=%d n_kc=%d total=%d hi=%d lo=%d)", n_out, NT,
                n_kc, n_kc_total, hi_copies, has_lo);
  if (fmt != MMSEG_FMT_BF16 && fmt != MMSEG_FMT_FP16) return fail(MMSEG_ERR_INVALID_ARG, "weights_repack: fmt");
  if (reinterpret_cast<uintptr_t>(dst) & 15) return fail(MMSEG_ERR_INVALID_ARG, "weights_repack: dst must be 16-byte aligned");
  const long long total = (long long)n_out * n_kc * 2 * ksize * ksize * ksize;
  const long long blocks = (total + 255) / 256;
  if (blocks > 2147483647LL) return fail(MMSEG_ERR_INVALID_ARG, "weights_repack: too large");
  weights_repack_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w, n_off, k_off, reinterpret_cast<uint4*>(dst), n_out, NT, n_kc, n_kc_total, ksize, flip ? 1 : 0, hi_copies,
      has_lo ? 1 : 0, fmt == MMSEG_FMT_FP16 ? 1 : 0, scale);
  return check_launch("weights_repack_kernel");
}

extern "C" int mmseg_weights_repack_multi(const void* descs, int32_t n_descs, const int32_t* block_desc, int64_t n_blocks,
                                          int32_t fmt, float scale, void* stream) {
  static_assert(sizeof(RepackDesc) == sizeof(mmseg_repack_desc), "mmseg_repack_desc layout");
  if (!descs || !block_desc || n_descs < 1 || n_blocks < 1 || n_blocks > 2147483647LL)
    return fail(MMSEG_ERR_INVALID_ARG, "weights_repack_multi: bad arguments");
  if (fmt != MMSEG_FMT_BF16 && fmt != MMSEG_FMT_FP16) return fail(MMSEG_ERR_INVALID_ARG, "weights_repack_multi: fmt");
  weights_repack_multi_kernel<<<(unsigned)n_blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const RepackDesc*>(descs), block_desc, fmt == MMSEG_FMT_FP16 ? 1 : 0, scale);
  return check_launch("weights_repack_multi_kernel");
}

extern "C" int mmseg_gather_f32(const float* src, const int32_t* idx, float* dst, int32_t n, void* stream) {
  if (!src || !idx || !dst || n < 1) return fail(MMSEG_ERR_INVALID_ARG, "gather_f32: bad arguments");
  gather_f32_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, idx, dst, n);
  return check_launch("gather_f32_kernel");
}

extern "C" int mmseg_adamw_multi(const void* tensors, int32_t n_tensors, const int32_t* chunks, int32_t n_chunks,
                                 const float* hyper, int32_t zero_grad, void* stream) {
  if (!tensors || !chunks || !hyper || n_chunks < 1 || n_tensors < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "adamw_multi: bad arguments");
  static_assert(sizeof(AdamwTensor) == sizeof(mmseg_adamw_tensor), "mmseg_adamw_tensor layout");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  adamw_multi_kernel<<<(unsigned)n_chunks, 256, 0, st>>>(reinterpret_cast<const AdamwTensor*>(tensors),
                                                        reinterpret_cast<const int2*>(chunks), n_chunks, hyper,
                                                        zero_grad ? 1 : 0);
  int rc = check_launch("adamw_multi_kernel");
  if (rc) return rc;
  // the counters advance stream-ordered behind every block's read of them
  adamw_step_inc_kernel<<<(n_tensors + 255) / 256, 256, 0, st>>>(reinterpret_cast<const AdamwTensor*>(tensors), n_tensors);
  return check_launch("adamw_step_inc_kernel");
}
