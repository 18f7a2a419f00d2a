// Sliding-window inference kernels: window gather (volume -> blocked bf16 batch), importance-weighted accumulation of
// window logits into the output volume (owner-thread per voxel, windows applied in index order: deterministic and in
// the same order as the reference loop), and the final normalise + argmax.  HBM-bound, coalesced along X.
#include "common.h"
#include "ptx.cuh"

namespace mmseg {

int num_sms();

// grid: (x-chunks, RZ*RY rows, n_win*cb)
__global__ void __launch_bounds__(128)
swi_gather_kernel(const float* __restrict__ vol, int C, int VZ, int VY, int VX, const int* __restrict__ starts,
                  int RZ, int RY, int RX, void* __restrict__ dst, int dst_cbt, int dst_lo_off, int cb, int fp16) {
  const int wc = blockIdx.z;
  const int win = wc / cb, c = wc - win * cb;
  const int row = blockIdx.y;
  const int lz = row / RY, ly = row - lz * RY;
  const int sz = starts[win * 3 + 0], sy = starts[win * 3 + 1], sx = starts[win * 3 + 2];
  const int gz = sz + lz, gy = sy + ly;
  const bool row_in = (gz >= 0) && (gz < VZ) && (gy >= 0) && (gy < VY);
  const size_t nvox_v = (size_t)VZ * VY * VX;
  const size_t nvox_r = (size_t)RZ * RY * RX;
  const size_t dst_base = ((size_t)(win * dst_cbt + c) * nvox_r + ((size_t)lz * RY + ly) * RX) * 8;
  const size_t lo_delta = dst_lo_off > 0 ? (size_t)dst_lo_off * nvox_r * 8 : 0;
  for (int lx = blockIdx.x * blockDim.x + threadIdx.x; lx < RX; lx += gridDim.x * blockDim.x) {
    const int gx = sx + lx;
    float v[8];
    if (dst_lo_off < 0) {
      // packed split (numeric modes that keep the INPUT hi + lo AND split the first conv's weights): the three operand
      // passes A_hi*W_hi + A_lo*W_hi + A_hi*W_lo of a C <= 5 channel input fit ONE 16-channel K chunk as the virtual
      // channels [hi(C) | lo(C) | hi(C)] against the weights [W_hi | W_hi | W_lo] (engine.py)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int vc = c * 8 + i;
        const int part = vc / C, ch = vc - part * C;
        float x = 0.f;
        if (row_in && gx >= 0 && gx < VX && part < 3) x = vol[(size_t)ch * nvox_v + ((size_t)gz * VY + gy) * VX + gx];
        const float hi = fp16 ? __half2float(__float2half_rn(x)) : __bfloat162float(__float2bfloat16_rn(x));
        v[i] = part == 1 ? x - hi : hi;
      }
      store8_act(dst, dst_base + (size_t)lx * 8, 0, v, fp16 != 0);
      continue;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = c * 8 + i;
      v[i] = (row_in && gx >= 0 && gx < VX && ch < C)
                 ? vol[(size_t)ch * nvox_v + ((size_t)gz * VY + gy) * VX + gx]
                 : 0.f;
    }
    store8_act(dst, dst_base + (size_t)lx * 8, lo_delta, v, fp16 != 0);
  }
}

// Windows as a plain NCDHW fp32 batch (zero outside the volume): the input of SwinUNETR's strided patch embedding, which
// reads the image itself rather than the 16-bit blocked copy.
__global__ void __launch_bounds__(128)
swi_gather_ncdhw_kernel(const float* __restrict__ vol, int C, int VZ, int VY, int VX, const int* __restrict__ starts,
                        int RZ, int RY, int RX, float* __restrict__ dst) {
  const int wc = blockIdx.z;
  const int win = wc / C, ch = wc - win * C;
  const int row = blockIdx.y;
  const int lz = row / RY, ly = row - lz * RY;
  const int gz = starts[win * 3 + 0] + lz, gy = starts[win * 3 + 1] + ly, sx = starts[win * 3 + 2];
  const bool row_in = (gz >= 0) && (gz < VZ) && (gy >= 0) && (gy < VY);
  const float* src = vol + (size_t)ch * VZ * VY * VX + ((size_t)gz * VY + gy) * VX;
  float* d = dst + ((size_t)wc * RZ * RY + row) * RX;
  for (int lx = blockIdx.x * blockDim.x + threadIdx.x; lx < RX; lx += gridDim.x * blockDim.x) {
    const int gx = sx + lx;
    d[lx] = (row_in && gx >= 0 && gx < VX) ? src[gx] : 0.f;
  }
}

// One owner thread per output voxel of the box; loops over the batch's windows in index order.
// seg*w and the accumulate are separate roundings (__fmul_rn / __fadd_rn) exactly like `seg *= w; out += seg`.
__global__ void __launch_bounds__(128)
swi_blend_kernel(const float* __restrict__ logits, const int* __restrict__ starts, int n_win, int K, int RZ, int RY,
                 int RX, const float* __restrict__ wz, const float* __restrict__ wy, const float* __restrict__ wx,
                 float w_floor, float* __restrict__ out, float* __restrict__ count, int VZ, int VY, int VX, int bz0,
                 int by0, int bx0, int bx1, int by_n) {
  const int row = blockIdx.y;
  const int z = bz0 + row / by_n, y = by0 + row % by_n;
  const size_t nvox_v = (size_t)VZ * VY * VX;
  const size_t nvox_r = (size_t)RZ * RY * RX;
  for (int x = bx0 + blockIdx.x * blockDim.x + threadIdx.x; x < bx1; x += gridDim.x * blockDim.x) {
    const size_t vox = ((size_t)z * VY + y) * VX + x;
    for (int j = 0; j < n_win; ++j) {
      const int lz = z - starts[j * 3 + 0], ly = y - starts[j * 3 + 1], lx = x - starts[j * 3 + 2];
      if (lz < 0 || lz >= RZ || ly < 0 || ly >= RY || lx < 0 || lx >= RX) continue;
      const float w = fmaxf(__fmul_rn(__fmul_rn(wz[lz], wy[ly]), wx[lx]), w_floor);
      const float* seg = logits + (size_t)j * K * nvox_r + ((size_t)lz * RY + ly) * RX + lx;
      for (int c = 0; c < K; ++c) {
        float* o = out + (size_t)c * nvox_v + vox;
        *o = __fadd_rn(*o, __fmul_rn(seg[(size_t)c * nvox_r], w));
      }
      count[vox] = __fadd_rn(count[vox], w);
    }
  }
}

// Window mode: the launch covers exactly one window (extent read from starts on the device, so the launch arguments
// are batch-independent and the call can sit in a CUDA graph).  grid: (x-chunks, RZ*RY)
__global__ void __launch_bounds__(256)
swi_blend_window_kernel(const float* __restrict__ logits, const int* __restrict__ starts, int K, int RZ, int RY, int RX,
                        const float* __restrict__ wz, const float* __restrict__ wy, const float* __restrict__ wx,
                        float w_floor, float* __restrict__ out, float* __restrict__ count, int VZ, int VY, int VX) {
  // block = (x lanes, rows): blockDim.y window rows per block so that every lane of every warp has work
  const int row = blockIdx.y * blockDim.y + threadIdx.y;
  if (row >= RZ * RY) return;
  const int lz = row / RY, ly = row - lz * RY;
  const int z = starts[0] + lz, y = starts[1] + ly;
  if (z < 0 || z >= VZ || y < 0 || y >= VY) return;
  const int sx = starts[2];
  const size_t nvox_v = (size_t)VZ * VY * VX;
  const size_t nvox_r = (size_t)RZ * RY * RX;
  const float wzy = __fmul_rn(wz[lz], wy[ly]);
  const size_t row_v = ((size_t)z * VY + y) * VX;
  const size_t row_r = ((size_t)lz * RY + ly) * RX;
  // 4 consecutive x per thread as float4 when the window row is 16-byte aligned in both tensors and inside the volume
  const bool vec = ((sx | RX | VX) & 3) == 0 && sx >= 0 && sx + RX <= VX;
  if (vec) {
    for (int lx = 4 * (blockIdx.x * blockDim.x + threadIdx.x); lx < RX; lx += 4 * gridDim.x * blockDim.x) {
      const float4 w4 = *reinterpret_cast<const float4*>(wx + lx);
      const float w[4] = {fmaxf(__fmul_rn(wzy, w4.x), w_floor), fmaxf(__fmul_rn(wzy, w4.y), w_floor),
                          fmaxf(__fmul_rn(wzy, w4.z), w_floor), fmaxf(__fmul_rn(wzy, w4.w), w_floor)};
      const size_t vox = row_v + sx + lx;
      float4* cp = reinterpret_cast<float4*>(count + vox);
      // channel planes in groups of 8: all 16 loads (window logits + accumulator) of a group are issued before the first
      // store — the per-channel load / add / store chain ran at 55 % of HBM bandwidth
      for (int c0 = 0; c0 < K; c0 += 8) {
        float4 s4[8], a4[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c0 + c < K) {
            s4[c] = *reinterpret_cast<const float4*>(logits + (size_t)(c0 + c) * nvox_r + row_r + lx);
            a4[c] = *reinterpret_cast<const float4*>(out + (size_t)(c0 + c) * nvox_v + vox);
          }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c0 + c < K) {
            float4 a = a4[c];
            a.x = __fadd_rn(a.x, __fmul_rn(s4[c].x, w[0]));
            a.y = __fadd_rn(a.y, __fmul_rn(s4[c].y, w[1]));
            a.z = __fadd_rn(a.z, __fmul_rn(s4[c].z, w[2]));
            a.w = __fadd_rn(a.w, __fmul_rn(s4[c].w, w[3]));
            *reinterpret_cast<float4*>(out + (size_t)(c0 + c) * nvox_v + vox) = a;
          }
        }
      }
      float4 cv = *cp;
      cv.x = __fadd_rn(cv.x, w[0]); cv.y = __fadd_rn(cv.y, w[1]); cv.z = __fadd_rn(cv.z, w[2]); cv.w = __fadd_rn(cv.w, w[3]);
      *cp = cv;
    }
    return;
  }
  for (int lx = blockIdx.x * blockDim.x + threadIdx.x; lx < RX; lx += gridDim.x * blockDim.x) {
    const int x = sx + lx;
    if (x < 0 || x >= VX) continue;
    const size_t vox = row_v + x;
    const float w = fmaxf(__fmul_rn(wzy, wx[lx]), w_floor);
    const float* seg = logits + row_r + lx;
    for (int c = 0; c < K; ++c) {
      float* o = out + (size_t)c * nvox_v + vox;
      *o = __fadd_rn(*o, __fmul_rn(seg[(size_t)c * nvox_r], w));
    }
    count[vox] = __fadd_rn(count[vox], w);
  }
}

// out_conv fused into the blend: logits = W x features + b are computed in registers from the window's last feature map
// (blocked bf16, optional lo plane) and blended straight into the volume accumulator — the NCDHW logits tensor is never
// written or read (28 + 28 MB per 96^3 window).  Same arithmetic, in the same order, as conv1x1_logits_kernel followed
// by swi_blend_window_kernel (fp32 FMAs over the channels in index order, then seg * w and out + seg as separate
// roundings), so the accumulator is bit-identical to the two-kernel path.  One thread = 4 consecutive x voxels of a
// window row (float4 accumulator accesses); needs RX, VX and every window's x origin to be multiples of 4.
template <int KP>   // classes padded to a multiple of 4 (<= 8)
__global__ void __launch_bounds__(128)
swi_logits_blend_kernel(const uint16_t* __restrict__ feat, int src_cbt, int cb_off, int lo_off, int cin_blocks, int win,
                        const float* __restrict__ weight, const float* __restrict__ bias, int K,
                        const int* __restrict__ starts, int RZ, int RY, int RX, const float* __restrict__ wz,
                        const float* __restrict__ wy, const float* __restrict__ wx, float w_floor, float* __restrict__ out,
                        float* __restrict__ count, int VZ, int VY, int VX, int fp16) {
  extern __shared__ float wsm[];   // [cin][KP], zero-padded classes
  const int cin = cin_blocks * 8;
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < cin * KP; i += blockDim.x * blockDim.y) {
    const int ci = i / KP, co = i - ci * KP;
    wsm[i] = co < K ? weight[(size_t)co * cin + ci] : 0.f;
  }
  __syncthreads();
  const int row = blockIdx.y * blockDim.y + threadIdx.y;
  if (row >= RZ * RY) return;
  const int lz = row / RY, ly = row - lz * RY;
  const int z = starts[0] + lz, y = starts[1] + ly;
  const int sx = starts[2];
  if (z < 0 || z >= VZ || y < 0 || y >= VY) return;
  const size_t nvox_v = (size_t)VZ * VY * VX;
  const size_t nvox_r = (size_t)RZ * RY * RX;
  const float wzy = __fmul_rn(wz[lz], wy[ly]);
  const size_t row_v = ((size_t)z * VY + y) * VX;
  const size_t row_r = ((size_t)lz * RY + ly) * RX;
  for (int lx = 4 * (blockIdx.x * blockDim.x + threadIdx.x); lx < RX; lx += 4 * gridDim.x * blockDim.x) {
    float acc[KP][4];
#pragma unroll
    for (int co = 0; co < KP; ++co) {
      const float b = (bias && co < K) ? bias[co] : 0.f;
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[co][v] = b;
    }
    for (int cb = 0; cb < cin_blocks; ++cb) {
      const uint16_t* base = feat + (((size_t)win * src_cbt + cb_off + cb) * nvox_r + row_r + lx) * 8;
      float x[4][8];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        cvt8_to_f32(*reinterpret_cast<const uint4*>(base + v * 8), x[v], fp16 != 0);
        if (lo_off > 0) {
          float l[8];
          cvt8_to_f32(*reinterpret_cast<const uint4*>(base + (size_t)lo_off * nvox_r * 8 + v * 8), l, fp16 != 0);
#pragma unroll
          for (int j = 0; j < 8; ++j) x[v][j] += l[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* wr = reinterpret_cast<const float4*>(wsm + (size_t)(cb * 8 + j) * KP);
#pragma unroll
        for (int c4 = 0; c4 < KP / 4; ++c4) {
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
    const float4 w4 = *reinterpret_cast<const float4*>(wx + lx);
    const float w[4] = {fmaxf(__fmul_rn(wzy, w4.x), w_floor), fmaxf(__fmul_rn(wzy, w4.y), w_floor),
                        fmaxf(__fmul_rn(wzy, w4.z), w_floor), fmaxf(__fmul_rn(wzy, w4.w), w_floor)};
    const size_t vox = row_v + sx + lx;
    float4 a4[KP];
#pragma unroll
    for (int c = 0; c < KP; ++c)
      if (c < K) a4[c] = *reinterpret_cast<const float4*>(out + (size_t)c * nvox_v + vox);
#pragma unroll
    for (int c = 0; c < KP; ++c) {
      if (c < K) {
        float4 a = a4[c];
        a.x = __fadd_rn(a.x, __fmul_rn(acc[c][0], w[0]));
        a.y = __fadd_rn(a.y, __fmul_rn(acc[c][1], w[1]));
        a.z = __fadd_rn(a.z, __fmul_rn(acc[c][2], w[2]));
        a.w = __fadd_rn(a.w, __fmul_rn(acc[c][3], w[3]));
        *reinterpret_cast<float4*>(out + (size_t)c * nvox_v + vox) = a;
      }
    }
    float4* cp = reinterpret_cast<float4*>(count + vox);
    float4 cv = *cp;
    cv.x = __fadd_rn(cv.x, w[0]); cv.y = __fadd_rn(cv.y, w[1]); cv.z = __fadd_rn(cv.z, w[2]); cv.w = __fadd_rn(cv.w, w[3]);
    *cp = cv;
  }
}

// `out` addresses the first voxel of the range in class plane 0; plane c starts `plane` elements further (a z-slab of the
// [K][VZ][VY][VX] accumulator is finalized in place: nvox = slab voxels, plane = VZ*VY*VX).
__global__ void __launch_bounds__(256)
swi_finalize_kernel(float* __restrict__ out, const float* __restrict__ count, int K, size_t nvox, size_t plane,
                    int normalize, uint8_t* __restrict__ labels) {
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (size_t)gridDim.x * blockDim.x) {
    const float cnt = count[v];
    float best = -INFINITY;
    int arg = 0;
    for (int c = 0; c < K; ++c) {
      const float q = __fdiv_rn(out[(size_t)c * plane + v], cnt);
      if (normalize) out[(size_t)c * plane + v] = q;
      if (q > best) { best = q; arg = c; }
    }
    if (labels) labels[v] = (uint8_t)arg;
  }
}

// acc[p][i] += part[p][i] for p < planes, i < n (plane strides in elements): the owner's rank-ordered add of another
// rank's partial sums in the sharded exchange.  float4 when both ranges are 16-byte aligned.
__global__ void __launch_bounds__(256)
swi_add_partial_kernel(float* __restrict__ acc, size_t acc_plane, const float* __restrict__ part, size_t part_plane,
                       size_t n, int vec) {
  float* a = acc + (size_t)blockIdx.y * acc_plane;
  const float* b = part + (size_t)blockIdx.y * part_plane;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (vec) {
    const size_t n4 = n >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 x = reinterpret_cast<float4*>(a)[i];
      const float4 y = reinterpret_cast<const float4*>(b)[i];
      x.x = __fadd_rn(x.x, y.x); x.y = __fadd_rn(x.y, y.y); x.z = __fadd_rn(x.z, y.z); x.w = __fadd_rn(x.w, y.w);
      reinterpret_cast<float4*>(a)[i] = x;
    }
    for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] = __fadd_rn(a[i], b[i]);
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] = __fadd_rn(a[i], b[i]);
  }
}

}  // namespace mmseg

using namespace mmseg;

extern "C" int mmseg_swi_gather(const float* volume, int32_t C, int32_t VZ, int32_t VY, int32_t VX,
                                const int32_t* starts_dev, int32_t n_win, int32_t RZ, int32_t RY, int32_t RX,
                                void* dst, int32_t dst_cbt, int32_t dst_lo_off, int32_t cb, int32_t fmt, void* stream) {
  if (!volume || !starts_dev || !dst || n_win < 1 || C < 1 || cb * 8 < (dst_lo_off < 0 ? 3 * C : C) || RZ * RY > 65535 || n_win * cb > 65535 ||
      (fmt != MMSEG_FMT_BF16 && fmt != MMSEG_FMT_FP16))
    return fail(MMSEG_ERR_INVALID_ARG, "swi_gather: bad arguments");
  dim3 grid((RX + 127) / 128, RZ * RY, n_win * cb);
  swi_gather_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      volume, C, VZ, VY, VX, starts_dev, RZ, RY, RX, dst, dst_cbt, dst_lo_off, cb, fmt);
  return check_launch("swi_gather_kernel");
}

extern "C" int mmseg_swi_gather_ncdhw(const float* volume, int32_t C, int32_t VZ, int32_t VY, int32_t VX,
                                      const int32_t* starts_dev, int32_t n_win, int32_t RZ, int32_t RY, int32_t RX,
                                      float* dst, void* stream) {
  if (!volume || !starts_dev || !dst || n_win < 1 || C < 1 || RZ * RY > 65535 || n_win * C > 65535)
    return fail(MMSEG_ERR_INVALID_ARG, "swi_gather_ncdhw: bad arguments");
  dim3 grid((RX + 127) / 128, RZ * RY, n_win * C);
  swi_gather_ncdhw_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(volume, C, VZ, VY, VX, starts_dev, RZ,
                                                                                    RY, RX, dst);
  return check_launch("swi_gather_ncdhw_kernel");
}

extern "C" int mmseg_swi_blend(const float* win_logits, const int32_t* starts_dev, int32_t n_win, int32_t K,
                               int32_t RZ, int32_t RY, int32_t RX, const float* wz, const float* wy, const float* wx,
                               float w_floor, float* out, float* count, int32_t VZ, int32_t VY, int32_t VX,
                               int32_t bz0, int32_t bz1, int32_t by0, int32_t by1, int32_t bx0, int32_t bx1,
                               void* stream) {
  if (!win_logits || !starts_dev || !wz || !wy || !wx || !out || !count || n_win < 1 || K < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "swi_blend: bad arguments");
  if (bz0 < 0) {  // window mode: one window, extent taken from starts_dev on the device
    if (n_win != 1) return fail(MMSEG_ERR_INVALID_ARG, "swi_blend: window mode takes exactly one window");
    if ((int64_t)RZ * RY > 65535) return fail(MMSEG_ERR_INVALID_ARG, "swi_blend: roi has too many rows");
    // every thread handles 4 voxels of a row (float4) when the row is aligned: RX/4 lanes per row, several rows per block
    const bool vec_ok = (RX & 3) == 0 && (VX & 3) == 0 && RX / 4 <= 128;
    const int bx = vec_ok ? RX / 4 : (RX < 128 ? RX : 128);
    const int by = bx >= 128 ? 1 : 128 / bx;
    dim3 wblock(bx, by);
    dim3 wgrid(vec_ok ? 1 : (RX + bx - 1) / bx, (RZ * RY + by - 1) / by);
    swi_blend_window_kernel<<<wgrid, wblock, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        win_logits, starts_dev, K, RZ, RY, RX, wz, wy, wx, w_floor, out, count, VZ, VY, VX);
    return check_launch("swi_blend_window_kernel");
  }
  if (bz0 < 0 || by0 < 0 || bx0 < 0 || bz1 > VZ || by1 > VY || bx1 > VX || bz1 <= bz0 || by1 <= by0 || bx1 <= bx0)
    return fail(MMSEG_ERR_INVALID_ARG, "swi_blend: box outside the volume");
  const int64_t rows = (int64_t)(bz1 - bz0) * (by1 - by0);
  if (rows > 65535) return fail(MMSEG_ERR_INVALID_ARG, "swi_blend: box has too many rows for one launch");
  dim3 grid((bx1 - bx0 + 127) / 128, (unsigned)rows);
  swi_blend_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      win_logits, starts_dev, n_win, K, RZ, RY, RX, wz, wy, wx, w_floor, out, count, VZ, VY, VX, bz0, by0, bx0, bx1,
      by1 - by0);
  return check_launch("swi_blend_kernel");
}

extern "C" int mmseg_swi_logits_blend(const void* feat, int32_t src_cbt, int32_t cb_off, int32_t lo_off, int32_t cin,
                                      int32_t window, const float* weight, const float* bias, int32_t K,
                                      const int32_t* starts_dev, int32_t RZ, int32_t RY, int32_t RX, const float* wz,
                                      const float* wy, const float* wx, float w_floor, float* out, float* count,
                                      int32_t VZ, int32_t VY, int32_t VX, int32_t fmt, void* stream) {
  if (fmt != MMSEG_FMT_BF16 && fmt != MMSEG_FMT_FP16) return fail(MMSEG_ERR_INVALID_ARG, "swi_logits_blend: fmt");
  if (!feat || !weight || !starts_dev || !wz || !wy || !wx || !out || !count || window < 0)
    return fail(MMSEG_ERR_INVALID_ARG, "swi_logits_blend: bad arguments");
  if (K < 1 || K > 8 || cin < 8 || (cin % 8) || cin > 128)
    return fail(MMSEG_ERR_UNSUPPORTED, "swi_logits_blend: K=%d (1..8), cin=%d (multiple of 8, <= 128)", K, cin);
  if ((RX & 3) || (VX & 3) || RX / 4 > 128 || (int64_t)RZ * RY > 65535 * 8)
    return fail(MMSEG_ERR_UNSUPPORTED, "swi_logits_blend: RX and VX must be multiples of 4 (RX <= 512)");
  const int bx = RX / 4;
  const int by = bx >= 128 ? 1 : 128 / bx;
  dim3 block(bx, by), grid(1, (RZ * RY + by - 1) / by);
  const uint16_t* f = reinterpret_cast<const uint16_t*>(feat);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (K <= 4)
    swi_logits_blend_kernel<4><<<grid, block, (size_t)cin * 4 * sizeof(float), st>>>(
        f, src_cbt, cb_off, lo_off, cin / 8, window, weight, bias, K, starts_dev, RZ, RY, RX, wz, wy, wx, w_floor, out, count,
        VZ, VY, VX, fmt);
  else
    swi_logits_blend_kernel<8><<<grid, block, (size_t)cin * 8 * sizeof(float), st>>>(
        f, src_cbt, cb_off, lo_off, cin / 8, window, weight, bias, K, starts_dev, RZ, RY, RX, wz, wy, wx, w_floor, out, count,
        VZ, VY, VX, fmt);
  return check_launch("swi_logits_blend_kernel");
}

extern "C" int mmseg_swi_finalize(float* out, const float* count, int32_t K, int64_t voxels, int64_t plane_stride,
                                  int32_t normalize_in_place, uint8_t* labels, void* stream) {
  if (!out || !count || K < 1 || voxels < 1 || plane_stride < voxels)
    return fail(MMSEG_ERR_INVALID_ARG, "swi_finalize: bad arguments");
  int64_t blocks = (voxels + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  swi_finalize_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      out, count, K, (size_t)voxels, (size_t)plane_stride, normalize_in_place, labels);
  return check_launch("swi_finalize_kernel");
}

extern "C" int mmseg_swi_add_partial(float* acc, int64_t acc_plane_stride, const float* part, int64_t part_plane_stride,
                                     int32_t planes, int64_t n, void* stream) {
  if (!acc || !part || planes < 1 || planes > 65535 || n < 1 || acc_plane_stride < n || part_plane_stride < n)
    return fail(MMSEG_ERR_INVALID_ARG, "swi_add_partial: bad arguments");
  const int vec = ((reinterpret_cast<uintptr_t>(acc) | reinterpret_cast<uintptr_t>(part)) & 15) == 0 &&
                  (acc_plane_stride & 3) == 0 && (part_plane_stride & 3) == 0;
  int64_t blocks = ((vec ? n / 4 : n) + 255) / 256;
  const int64_t cap = ((int64_t)num_sms() * 16 + planes - 1) / planes;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dim3 grid((unsigned)blocks, (unsigned)planes);
  swi_add_partial_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      acc, (size_t)acc_plane_stride, part, (size_t)part_plane_stride, (size_t)n, vec);
  return check_launch("swi_add_partial_kernel");
}
