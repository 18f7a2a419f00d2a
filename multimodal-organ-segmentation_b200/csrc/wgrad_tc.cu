// Conv3d weight gradient (k=3 pad 1, k=1; ConvTranspose3d k2 s2 through its k=1 GEMM view) as a tcgen05 / TMEM GEMM
// whose contraction (K) dimension runs over VOXELS:   dW[co, ci, tap] = sum_vox dY[vox, co] * X[vox + tap, ci].
//
// Both operands are read MN-major straight from the blocked activation layout ([cb][Z][Y][X][8] bf16): a voxel row is
// 16 bytes = 8 channels, 8 consecutive voxel rows form one SWIZZLE_NONE core matrix (K direction, LBO = 128 B) and the
// next channel block (MN direction) is one smem plane further (SBO = plane bytes) — no transposition anywhere.
//   A (M = 128 rows) = X halo tile, loaded three times by TMA with x shifted by dx = -1, 0, +1 (rows of TX voxels, no x
//       halo, so A and B share the row pitch): M index = (dx, ci) -> 3*CIG valid rows (the rest multiply garbage and
//       are never stored).  The dy tap is a start-address offset of dy*TX rows.
//   B (N columns) = dY tile; the dz taps are folded into N: the planes z-1, z, z+1 of dY sit in consecutive ring slots,
//       so one MMA of N = 3*NTc columns pairs X plane p with dY planes p+1, p, p-1  (dz = 0, 1, 2).
//   D[dy] (TMEM, fp32) = [(dx, ci)] x [(dz-descending, co)], kept resident while a PERSISTENT CTA sweeps many voxel
//       tiles; each CTA then writes one partial and mmseg_wgrad_reduce sums the partials in a fixed order
//       (deterministic split-K, no atomics) straight into the PyTorch-layout fp32 gradient.
#include <cstdlib>
#include <cuda.h>

#include "common.h"
#include "ptx.cuh"

namespace mmseg {

struct WgradKParams {
  int X, Y, Z, n_img;
  int TX, TY, TZ;
  int tiles_x, tiles_y, tiles_z, n_tiles;
  int cig_blocks, cot_blocks, n_cig, n_cot;
  int x_cbt, y_cbt, y_cb0;
  uint32_t xplane_bytes, yplane_bytes, xslot_bytes, yslot_bytes, x_off, y_off;
  uint32_t tmem_cols;
  int ncols;
  float* partial;
  int16_t x_cb[MMSEG_MAX_WGRAD_GROUPS];
};

constexpr int kWThreads = 192;
constexpr int kRX = 2;  // X-plane ring slots
constexpr int kRY = 4;  // dY-plane ring slots (three live + one in flight)

struct __align__(16) WSmemHeader {
  uint64_t x_full[kRX], x_empty[kRX];
  uint64_t y_full[kRY], y_empty[kRY];
  uint64_t acc_full, acc_zero;
  uint32_t tmem_ptr;
};
constexpr uint32_t kWHeaderBytes = 256;
static_assert(sizeof(WSmemHeader) <= kWHeaderBytes, "header too large");

// idesc: bf16 x bf16 -> f32, A and B both MN-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_mn(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void umma_mn(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
      : "memory");
}

template <int KT>
__global__ void __launch_bounds__(kWThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                const __grid_constant__ WgradKParams p) {
  constexpr int H = KT / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  WSmemHeader* hdr = reinterpret_cast<WSmemHeader*>(smem);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t x_smem = smem_base + p.x_off;
  const uint32_t y_smem = smem_base + p.y_off;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.y;
  const int cig = pair / p.n_cot, cot = pair - cig * p.n_cot;
  const int NTc = p.cot_blocks * 8;
  const int PY = p.TY + 2 * H;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRX; ++s) { mbar_init(smem_u32(&hdr->x_full[s]), 1); mbar_init(smem_u32(&hdr->x_empty[s]), 1); }
    for (int s = 0; s < kRY; ++s) { mbar_init(smem_u32(&hdr->y_full[s]), 1); mbar_init(smem_u32(&hdr->y_empty[s]), 1); }
    mbar_init(smem_u32(&hdr->acc_full), 1);
    mbar_init(smem_u32(&hdr->acc_zero), 128);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmY); }
  if (warp == 2) { tmem_alloc(smem_u32(&hdr->tmem_ptr), p.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_ptr;
  const uint32_t x_bytes = (uint32_t)KT * p.cig_blocks * p.xplane_bytes;   // bytes landing per X plane (KT dx copies)
  const uint32_t x_dx_bytes = (uint32_t)p.cig_blocks * p.xplane_bytes;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t xc = 0, yc = 0;  // planes loaded so far (ring position + phase)
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        int r = t;
        const int tx = r % p.tiles_x; r /= p.tiles_x;
        const int ty = r % p.tiles_y; r /= p.tiles_y;
        const int tz = r % p.tiles_z;
        const int img = r / p.tiles_z;
        const int x0 = tx * p.TX, y0 = ty * p.TY, z0 = tz * p.TZ;
        const int tzv = min(p.TZ, p.Z - z0);
        const int xcb = img * p.x_cbt + p.x_cb[cig];
        const int ycb = img * p.y_cbt + p.y_cb0 + cot * p.cot_blocks;
        for (int i = 0; i < tzv + 2 * H; ++i) {
          if (i < tzv) {  // dY plane z0 + i
            const uint32_t s = yc % kRY, ph = (yc / kRY) & 1;
            mbar_wait(smem_u32(&hdr->y_empty[s]), ph ^ 1);
            const uint32_t full = smem_u32(&hdr->y_full[s]);
            mbar_arrive_expect_tx(full, p.yslot_bytes);
            tma_load_4d(y_smem + s * p.yslot_bytes, &tmY, full, 2 * x0, y0, z0 + i, ycb);
            ++yc;
          }
          const int pz = z0 - H + i;  // X plane
          if (pz >= 0 && pz < p.Z) {
            const uint32_t s = xc % kRX, ph = (xc / kRX) & 1;
            mbar_wait(smem_u32(&hdr->x_empty[s]), ph ^ 1);
            const uint32_t full = smem_u32(&hdr->x_full[s]);
            mbar_arrive_expect_tx(full, x_bytes);
#pragma unroll
            for (int dx = 0; dx < KT; ++dx)
              tma_load_4d(x_smem + s * p.xslot_bytes + dx * x_dx_bytes, &tmX, full, 2 * (x0 - H + dx), y0 - H, pz, xcb);
            ++xc;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t a_hi = ((p.xplane_bytes >> 4) & 0x3FFFu) | (1u << 14);  // SBO = MN-group stride = one X plane
      const uint32_t b_hi = ((p.yplane_bytes >> 4) & 0x3FFFu) | (1u << 14);
      const uint32_t lbo = (128u >> 4) << 16;                                  // K-group stride = 8 voxel rows
      const int nchunks = (p.TX * p.TY) >> 4;
      uint32_t xc = 0, yc = 0, yw = 0;  // yc: first dY plane counter of this tile; yw: dY planes waited so far
      mbar_wait(smem_u32(&hdr->acc_zero), 0);
      tc_fence_after();
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        int r = t / (p.tiles_x * p.tiles_y);
        const int tz = r % p.tiles_z;
        const int z0 = tz * p.TZ;
        const int tzv = min(p.TZ, p.Z - z0);
        for (int i = 0; i < tzv + 2 * H; ++i) {
          const int pz = z0 - H + i;
          if (pz >= 0 && pz < p.Z) {
            const int za = max(pz - H, z0), zb = min(pz + H, z0 + tzv - 1);
            // dY planes up to zb must have landed
            while ((int)(yw - yc) <= zb - z0) {
              mbar_wait(smem_u32(&hdr->y_full[yw % kRY]), (yw / kRY) & 1);
              ++yw;
            }
            const uint32_t sx = xc % kRX;
            mbar_wait(smem_u32(&hdr->x_full[sx]), (xc / kRX) & 1);
            tc_fence_after();
            if (za <= zb) {
              const uint32_t ya = (yc + (uint32_t)(za - z0)) % kRY;      // ring slot of plane za
              const int n_pl = zb - za + 1;
              const int n_first = min(n_pl, kRY - (int)ya);              // planes before the ring wraps
              const uint32_t col0 = (uint32_t)(za - (pz - H)) * NTc;      // dz-descending column block of plane za
              const uint32_t a_base = (((x_smem + sx * p.xslot_bytes) >> 4) & 0x3FFFu) | lbo;
              const uint32_t b1 = (((y_smem + ya * p.yslot_bytes) >> 4) & 0x3FFFu) | lbo;
              const uint32_t b2 = ((y_smem >> 4) & 0x3FFFu) | lbo;        // slot 0 after the wrap
              const uint32_t id1 = make_idesc_mn((uint32_t)(n_first * NTc));
              const uint32_t id2 = make_idesc_mn((uint32_t)((n_pl - n_first) * NTc));
              for (int kc = 0; kc < nchunks; ++kc) {
#pragma unroll
                for (int dy = 0; dy < KT; ++dy) {
                  const uint32_t a = a_base + (uint32_t)(dy * p.TX + kc * 16);
                  const uint32_t d = tmem_base + (uint32_t)(dy * KT * NTc) + col0;
                  umma_mn(d, a, a_hi, b1 + (uint32_t)(kc * 16), b_hi, id1);
                  if (n_first < n_pl) umma_mn(d + (uint32_t)(n_first * NTc), a, a_hi, b2 + (uint32_t)(kc * 16), b_hi, id2);
                }
              }
            }
            umma_commit(smem_u32(&hdr->x_empty[sx]));
            ++xc;
          }
          // dY plane pz - H is not needed by any later X plane
          const int zr = pz - H;
          if (zr >= z0 && zr < z0 + tzv) umma_commit(smem_u32(&hdr->y_empty[(yc + (uint32_t)(zr - z0)) % kRY]));
        }
        yc += (uint32_t)tzv;
      }
      umma_commit(smem_u32(&hdr->acc_full));
    }
  } else {
    // ===================== epilogue warps: zero TMEM first, store the partial at the end =====================
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    for (uint32_t c = 0; c < p.tmem_cols; c += 16) tmem_st16_zero(lane_base + c);
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive(smem_u32(&hdr->acc_zero));

    mbar_wait(smem_u32(&hdr->acc_full), 0);
    tc_fence_after();
    float* dst = p.partial + (((size_t)pair * gridDim.x + blockIdx.x) * 128 + (size_t)(q * 32 + lane)) * p.ncols;
    for (int c = 0; c < p.ncols; c += 16) {
      float v[16];
      tmem_ld16(lane_base + (uint32_t)c, v);
      float4* d4 = reinterpret_cast<float4*>(dst + c);
      d4[0] = make_float4(v[0], v[1], v[2], v[3]);
      d4[1] = make_float4(v[4], v[5], v[6], v[7]);
      d4[2] = make_float4(v[8], v[9], v[10], v[11]);
      d4[3] = make_float4(v[12], v[13], v[14], v[15]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// dW[...] = sum over the persistent CTAs' partials in a fixed order (deterministic split-K), written straight into the
// PyTorch-layout fp32 gradient.
//   mode 0 (Conv3d):          dst[(co*Cin + ci)*taps + tap],            tap = (dz*KT + dy)*KT + dx
//   mode 1 (ConvTranspose3d): GEMM column n = tap8*Cout + co,            dst[(ci*Cout + co)*8 + tap8]
// The partials are read in THEIR memory order: a warp owns 32 consecutive fp32 columns of one partial row (128-byte
// coalesced loads), the block's 8 warps each sum a slice k = w, w + 8, ... of the n_part partials (independent loads in
// flight), and the slices are combined in warp order through shared memory — every element sees the same association
// whatever the grid.  (The first version had one thread per weight element walking all partials with a 128*ncols
// stride: 175-500 GB/s, 1.05 ms of a 20 ms training step.)
// S slices x R rows per block (S * R = 8 warps): S = 8 when there are many partials, 1 when a single CTA produced the
// partial (the deep layers: pure layout change).  Only the 3*CIG valid accumulator rows get blocks.
template <int S>
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int n_part, int ncols, int n_cot, int NTc, int CIG, int KT,
                    int Cin, int Cout_gemm, int Cout, int mode, const int* __restrict__ ci_of_pos,
                    float* __restrict__ dst) {
  constexpr int R = 8 / S;
  __shared__ double sm[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slice = warp % S, rl = warp / S;
  const int valid_rows = KT * CIG;
  const int row_blocks = (valid_rows + R - 1) / R;
  const int cols_blocks = (ncols + 31) / 32;
  int b = blockIdx.x;
  const int cblk = b % cols_blocks; b /= cols_blocks;
  const int rblk = b % row_blocks;
  const int pair = b / row_blocks;                // (input-channel group, output-channel group)
  const int row = rblk * R + rl;                  // accumulator row = (dx, ci inside the group)
  const int col = cblk * 32 + lane;
  const int dx = row / CIG, cil = row - dx * CIG;
  const int cig = pair / n_cot, cot = pair - cig * n_cot;
  const int ci = (row < valid_rows) ? ci_of_pos[cig * CIG + cil] : -1;
  double s = 0.0;
  if (col < ncols && ci >= 0) {
    const float* p = partial + (((size_t)pair * n_part) * 128 + row) * ncols + col;
    const size_t kstride = (size_t)128 * ncols;
    int k = slice;
    for (; k + 3 * S < n_part; k += 4 * S) {
      const float v0 = p[(size_t)k * kstride], v1 = p[(size_t)(k + S) * kstride];
      const float v2 = p[(size_t)(k + 2 * S) * kstride], v3 = p[(size_t)(k + 3 * S) * kstride];
      s += (double)v0; s += (double)v1; s += (double)v2; s += (double)v3;
    }
    for (; k < n_part; k += S) s += (double)p[(size_t)k * kstride];
  }
  if (S > 1) {
    sm[warp][lane] = s;
    __syncthreads();
    if (slice != 0) return;
    s = 0.0;
#pragma unroll
    for (int w = 0; w < S; ++w) s += sm[rl * S + w][lane];
  }
  if (col >= ncols || ci < 0) return;
  const int dy = col / (KT * NTc), r2 = col - dy * (KT * NTc);
  const int dzr = r2 / NTc, col_l = r2 - dzr * NTc;
  const int dz = KT - 1 - dzr;
  const int n = cot * NTc + col_l;                // GEMM column (co, or tap8*Cout + co)
  if (n >= Cout_gemm) return;
  const int taps = KT * KT * KT;
  const int tap = (dz * KT + dy) * KT + dx;
  size_t o;
  if (mode == 0) {
    o = ((size_t)n * Cin + ci) * taps + tap;
  } else {
    const int tap8 = n / Cout, co = n - tap8 * Cout;
    o = ((size_t)ci * Cout + co) * 8 + tap8;
  }
  dst[o] = (float)s;
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiledW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledW get_encode_fn_w() {
  static PFN_encodeTiledW fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiledW>(f);
    cudaGetLastError();
  }
  return fn;
}

static inline uint32_t rup(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

struct WgradPlan {
  WgradKParams k;
  uint32_t smem_bytes;
  int KT;
};

static int plan_wgrad(const mmseg_wgrad_args* a, WgradPlan* out) {
  if (!a) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: null args");
  if (a->ksize != 1 && a->ksize != 3) return fail(MMSEG_ERR_UNSUPPORTED, "wgrad: ksize %d (only 1, 3)", a->ksize);
  if (a->n_img < 1 || a->X < 1 || a->Y < 1 || a->Z < 1) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: bad extents");
  if (a->TX < 1 || a->TY < 1 || a->TZ < 1 || ((a->TX * a->TY) & 15))
    return fail(MMSEG_ERR_INVALID_ARG, "wgrad: TX*TY=%d must be a positive multiple of 16", a->TX * a->TY);
  if (a->TX > 128) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: TX > 128 (TMA box limit)");
  const int KT = a->ksize, H = KT / 2;
  if (a->cig_blocks < 1 || KT * a->cig_blocks > 16) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: %d ci blocks x %d dx copies > 16 M groups", a->cig_blocks, KT);
  if (a->cot_blocks & 1) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: cot_blocks must be even (MMA N is a multiple of 16)");
  if (a->cot_blocks < 1 || KT * a->cot_blocks * 8 > 256) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: folded N = %d > 256", KT * a->cot_blocks * 8);
  if (a->n_cig < 1 || a->n_cig > MMSEG_MAX_WGRAD_GROUPS || a->n_cot < 1) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: group counts");
  if (a->n_part < 1) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: n_part");
  WgradKParams& k = out->k;
  out->KT = KT;
  k.X = a->X; k.Y = a->Y; k.Z = a->Z; k.n_img = a->n_img;
  k.TX = a->TX; k.TY = a->TY; k.TZ = a->TZ;
  k.tiles_x = (a->X + a->TX - 1) / a->TX; k.tiles_y = (a->Y + a->TY - 1) / a->TY; k.tiles_z = (a->Z + a->TZ - 1) / a->TZ;
  k.n_tiles = k.tiles_x * k.tiles_y * k.tiles_z * a->n_img;
  k.cig_blocks = a->cig_blocks; k.cot_blocks = a->cot_blocks; k.n_cig = a->n_cig; k.n_cot = a->n_cot;
  k.x_cbt = a->x_cbt; k.y_cbt = a->y_cbt; k.y_cb0 = a->y_cb0;
  const int PY = a->TY + 2 * H;
  k.xplane_bytes = (uint32_t)PY * a->TX * 16u;
  k.yplane_bytes = (uint32_t)a->TY * a->TX * 16u;
  if ((k.xplane_bytes >> 4) > 0x3FFF) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: X plane too large for SBO");
  k.xslot_bytes = (uint32_t)KT * a->cig_blocks * k.xplane_bytes;
  // every TMA destination (the KT x-shifted copies of an X plane, the ring slots) must be 128-byte aligned
  if (((uint32_t)a->cig_blocks * k.xplane_bytes) & 127u)
    return fail(MMSEG_ERR_INVALID_ARG, "wgrad: %d ci blocks x %u-byte X planes: the x-shifted copies would not be 128-byte aligned",
                a->cig_blocks, k.xplane_bytes);
  k.yslot_bytes = (uint32_t)a->cot_blocks * k.yplane_bytes;
  k.ncols = KT * KT * a->cot_blocks * 8;
  if (k.ncols > 512) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: %d TMEM columns > 512", k.ncols);
  uint32_t tc = 32;
  while ((int)tc < k.ncols) tc <<= 1;
  k.tmem_cols = tc;
  k.x_off = kWHeaderBytes;
  k.y_off = k.x_off + kRX * k.xslot_bytes;
  // A reads 16 MN groups from the start of a slot (+ the dy/kc row offset): keep that inside the allocation
  const uint32_t a_reach = (kRX - 1) * k.xslot_bytes + 15u * k.xplane_bytes + (uint32_t)(2 * H * a->TX + a->TX * a->TY) * 16u;
  uint32_t total = k.y_off + kRY * k.yslot_bytes;
  if (k.x_off + a_reach > total) total = k.x_off + a_reach;
  total = rup(total, 128) + 128;
  if (total < 116u * 1024u) total = 116u * 1024u;  // one CTA per SM (each allocates up to all 512 TMEM columns)
  if (total > 227u * 1024u) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: %u bytes of shared memory > 227 KB", total);
  if (k.xslot_bytes >= (1u << 20) || k.yslot_bytes >= (1u << 20)) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: tx bytes");
  out->smem_bytes = total;
  k.partial = a->partial;
  for (int i = 0; i < MMSEG_MAX_WGRAD_GROUPS; ++i) k.x_cb[i] = i < a->n_cig ? a->x_cb[i] : 0;
  for (int i = 0; i < a->n_cig; ++i)
    if (a->x_cb[i] < 0 || a->x_cb[i] + a->cig_blocks > a->x_cbt) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: x_cb[%d] outside x_cbt", i);
  if (a->y_cb0 < 0 || a->y_cb0 + a->n_cot * a->cot_blocks > a->y_cbt) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: dY blocks outside y_cbt");
  return MMSEG_OK;
}

}  // namespace mmseg

using namespace mmseg;

extern "C" int64_t mmseg_conv3d_wgrad_smem_bytes(const mmseg_wgrad_args* a) {
  WgradPlan pl;
  int rc = plan_wgrad(a, &pl);
  if (rc != MMSEG_OK) return rc;
  return pl.smem_bytes;
}

extern "C" int mmseg_conv3d_wgrad(const mmseg_wgrad_args* a, void* stream) {
  WgradPlan pl;
  int rc = plan_wgrad(a, &pl);
  if (rc != MMSEG_OK) return rc;
  if (!a->x || !a->dy || !a->partial) return fail(MMSEG_ERR_INVALID_ARG, "wgrad: null pointer");
  PFN_encodeTiledW enc = get_encode_fn_w();
  if (!enc) return fail(MMSEG_ERR_NO_DRIVER, "wgrad: cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  const WgradKParams& k = pl.k;
  const int H = pl.KT / 2;
  CUtensorMap tmx, tmy;
  cuuint64_t strides[3] = {(cuuint64_t)k.X * 16, (cuuint64_t)k.X * k.Y * 16, (cuuint64_t)k.X * k.Y * k.Z * 16};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  cuuint64_t dimsx[4] = {(cuuint64_t)2 * k.X, (cuuint64_t)k.Y, (cuuint64_t)k.Z, (cuuint64_t)k.n_img * k.x_cbt};
  cuuint32_t boxx[4] = {(cuuint32_t)(2 * k.TX), (cuuint32_t)(k.TY + 2 * H), 1, (cuuint32_t)k.cig_blocks};
  CUresult cr = enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(a->x), dimsx, strides, boxx, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(MMSEG_ERR_CUDA, "wgrad: cuTensorMapEncodeTiled(X) failed (%d)", (int)cr);
  cuuint64_t dimsy[4] = {(cuuint64_t)2 * k.X, (cuuint64_t)k.Y, (cuuint64_t)k.Z, (cuuint64_t)k.n_img * k.y_cbt};
  cuuint32_t boxy[4] = {(cuuint32_t)(2 * k.TX), (cuuint32_t)k.TY, 1, (cuuint32_t)k.cot_blocks};
  cr = enc(&tmy, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(a->dy), dimsy, strides, boxy, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(MMSEG_ERR_CUDA, "wgrad: cuTensorMapEncodeTiled(dY) failed (%d)", (int)cr);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(MMSEG_ERR_CUDA, "wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)a->n_part, (unsigned)(k.n_cig * k.n_cot));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pl.KT == 3) wgrad_tc_kernel<3><<<grid, kWThreads, pl.smem_bytes, st>>>(tmx, tmy, k);
  else wgrad_tc_kernel<1><<<grid, kWThreads, pl.smem_bytes, st>>>(tmx, tmy, k);
  return check_launch("wgrad_tc_kernel");
}

// Conv3d 3x3x3 layers whose gradient is large and whose split-K is shallow (n_part < 16: 128 channels and up — where
// the bytes of a step's weight gradients are).  The kernel above writes 4-byte elements 108 bytes apart (a warp's lanes
// are output channels, 27 * Cin floats apart), so every 32-byte sector of dW is completed by 8 different blocks.  Here a
// block owns, for 8 input channels and 8 output channels, ALL 27 taps — accumulator rows (dx, ci) x columns (dy, dz, co),
// read as 32-byte segments — sums the partials in index order (fp64), transposes through shared memory and writes runs of
// 8 * 27 consecutive floats of the PyTorch layout.
constexpr int kRedRC = 8, kRedNC = 8, kRedElems = 3 * kRedRC * 9 * kRedNC;      // 1,728 elements per block
constexpr int kRedPerThread = (kRedElems + 255) / 256;
// element e of the block's tile, in LOAD order (row, 32-byte column segment, column) -> its slot in STORE order
// (output channel, input channel, tap) and its offset inside one partial
__device__ __forceinline__ void red_tile_elem(int e, int CIG, int NTc, int ncols, int cil0, int n0, int& off, int& slot) {
  const int r = e / (9 * kRedNC), c = e - r * (9 * kRedNC);
  const int dx = r / kRedRC, cil = r - dx * kRedRC;
  const int seg = c / kRedNC, nl = c - seg * kRedNC;              // seg = dy * 3 + dzr (the partial's column order)
  const int dy = seg / 3, dz = 2 - (seg - dy * 3);
  off = (dx * CIG + cil0 + cil) * ncols + seg * NTc + n0 + nl;
  slot = (nl * kRedRC + cil) * 27 + (dz * 3 + dy) * 3 + dx;
}
__global__ void __launch_bounds__(256, 4)
wgrad_reduce_tile_kernel(const float* __restrict__ partial, int n_part, int ncols, int n_cot, int NTc, int CIG, int Cin,
                         int Cout_gemm, const int* __restrict__ ci_of_pos, float* __restrict__ dst) {
  __shared__ float tile[kRedElems];
  int b = blockIdx.x;
  const int n_nblk = NTc / kRedNC, n_rblk = CIG / kRedRC;
  const int nblk = b % n_nblk; b /= n_nblk;
  const int rblk = b % n_rblk;
  const int pair = b / n_rblk;
  const int cig = pair / n_cot, cot = pair - cig * n_cot;
  const int cil0 = rblk * kRedRC, n0 = nblk * kRedNC;
  const float* base = partial + (size_t)pair * n_part * 128 * ncols;
  const size_t kstride = (size_t)128 * ncols;
  int off[kRedPerThread];
  float acc[kRedPerThread];           // n_part < 16 addends in index order: fp32 is exact enough (the deep split-K
                                      // layers go through the fp64 kernel above)
#pragma unroll
  for (int i = 0; i < kRedPerThread; ++i) {
    const int e = threadIdx.x + i * 256;
    int slot;
    red_tile_elem(e < kRedElems ? e : 0, CIG, NTc, ncols, cil0, n0, off[i], slot);
    acc[i] = (e < kRedElems) ? base[off[i]] : 0.f;
  }
  for (int k = 1; k < n_part; ++k) {
    float v[kRedPerThread];
#pragma unroll
    for (int i = 0; i < kRedPerThread; ++i) v[i] = base[(size_t)k * kstride + off[i]];
#pragma unroll
    for (int i = 0; i < kRedPerThread; ++i) acc[i] += v[i];
  }
#pragma unroll
  for (int i = 0; i < kRedPerThread; ++i) {
    const int e = threadIdx.x + i * 256;
    if (e < kRedElems) {
      int o, slot;
      red_tile_elem(e, CIG, NTc, ncols, cil0, n0, o, slot);
      tile[slot] = acc[i];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kRedElems; e += 256) {
    const int nl = e / (kRedRC * 27), r = e - nl * (kRedRC * 27);
    const int cil = r / 27, tap = r - cil * 27;
    const int n = cot * NTc + n0 + nl;
    const int ci = ci_of_pos[cig * CIG + cil0 + cil];
    if (n < Cout_gemm && ci >= 0) dst[((size_t)n * Cin + ci) * 27 + tap] = tile[e];
  }
}

// The same for 1x1x1 convs / Linear layers (accumulator rows = input channels, columns = output channels, dW[co][ci]):
// a 32 x 32 tile read along the columns, transposed through shared memory, written along ci.
__global__ void __launch_bounds__(256)
wgrad_reduce_tile_k1_kernel(const float* __restrict__ partial, int n_part, int ncols, int n_cot, int NTc, int CIG, int Cin,
                            int Cout_gemm, const int* __restrict__ ci_of_pos, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  int b = blockIdx.x;
  const int n_nblk = (NTc + 31) / 32, n_rblk = (CIG + 31) / 32;
  const int nblk = b % n_nblk; b /= n_nblk;
  const int rblk = b % n_rblk;
  const int pair = b / n_rblk;
  const int cig = pair / n_cot, cot = pair - cig * n_cot;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = nblk * 32 + tx;
  const float* base = partial + (size_t)pair * n_part * 128 * ncols + col;
  const size_t kstride = (size_t)128 * ncols;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  bool ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) ok[j] = col < NTc && rblk * 32 + ty + 8 * j < CIG;
  for (int k = 0; k < n_part; ++k) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ok[j] ? base[(size_t)k * kstride + (size_t)(rblk * 32 + ty + 8 * j) * ncols] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] += v[j];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) tile[ty + 8 * j][tx] = acc[j];
  __syncthreads();
  const int cil = rblk * 32 + tx;
  const int ci = cil < CIG ? ci_of_pos[cig * CIG + cil] : -1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int nl = nblk * 32 + ty + 8 * j;
    const int n = cot * NTc + nl;
    if (nl < NTc && n < Cout_gemm && ci >= 0) dst[(size_t)n * Cin + ci] = tile[tx][ty + 8 * j];
  }
}

extern "C" int mmseg_wgrad_reduce(const float* partial, int32_t n_part, int32_t ksize, int32_t cig_blocks,
                                  int32_t cot_blocks, int32_t n_cig, int32_t n_cot, int32_t Cin, int32_t Cout_gemm,
                                  int32_t Cout, int32_t transposed, const int32_t* ci_of_pos, float* dst, void* stream) {
  if (!partial || !ci_of_pos || !dst || n_part < 1 || Cin < 1 || Cout_gemm < 1 || n_cig < 1 || n_cot < 1)
    return fail(MMSEG_ERR_INVALID_ARG, "wgrad_reduce: bad arguments");
  const int KT = ksize;
  const int ncols = KT * KT * cot_blocks * 8;
  const int S = n_part >= 16 ? 8 : (n_part >= 4 ? 4 : (n_part >= 2 ? 2 : 1));
  const int R = 8 / S;
  const int valid_rows = KT * cig_blocks * 8;
  const long long blocks = (long long)n_cig * n_cot * ((valid_rows + R - 1) / R) * ((ncols + 31) / 32);
  if (blocks > 2147483647LL) return fail(MMSEG_ERR_INVALID_ARG, "wgrad_reduce: grid too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const bool tile_ok = [] { const char* e = getenv("MMSEG_WGRAD_REDUCE_TILE"); return !(e && e[0] == '0'); }();
  if (tile_ok && KT == 3 && !transposed && n_part < 16) {     // cig_blocks * 8 and cot_blocks * 8 are multiples of 8
    const long long tb = (long long)n_cig * n_cot * cig_blocks * cot_blocks;
    wgrad_reduce_tile_kernel<<<(unsigned)tb, 256, 0, st>>>(partial, n_part, ncols, n_cot, cot_blocks * 8, cig_blocks * 8, Cin,
                                                           Cout_gemm, ci_of_pos, dst);
    return check_launch("wgrad_reduce_tile_kernel");
  }
  if (tile_ok && KT == 1 && !transposed && n_part < 16) {
    const long long tb = (long long)n_cig * n_cot * ((cig_blocks * 8 + 31) / 32) * ((cot_blocks * 8 + 31) / 32);
    wgrad_reduce_tile_k1_kernel<<<(unsigned)tb, 256, 0, st>>>(partial, n_part, ncols, n_cot, cot_blocks * 8, cig_blocks * 8,
                                                              Cin, Cout_gemm, ci_of_pos, dst);
    return check_launch("wgrad_reduce_tile_k1_kernel");
  }
#define MMSEG_RED(SV)                                                                                                  \
  wgrad_reduce_kernel<SV><<<(unsigned)blocks, 256, 0, st>>>(partial, n_part, ncols, n_cot, cot_blocks * 8, cig_blocks * 8, \
                                                            KT, Cin, Cout_gemm, Cout, transposed ? 1 : 0, ci_of_pos, dst)
  if (S == 8) MMSEG_RED(8);
  else if (S == 4) MMSEG_RED(4);
  else if (S == 2) MMSEG_RED(2);
  else MMSEG_RED(1);
#undef MMSEG_RED
  return check_launch("wgrad_reduce_kernel");
}
