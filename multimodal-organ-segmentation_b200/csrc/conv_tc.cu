// Conv3d forward (k=3 pad 1, k=1, and ConvTranspose3d k2 s2 as a 1x1 GEMM + pixel shuffle) as a tcgen05 / TMEM
// implicit GEMM for sm_100a.  See include/mmseg_b200.h for the contract and DESIGN.md for the derivation.
//
// conv3d_tc_kernel<MT, KS> (tile kernel): one persistent CTA sweeps output tiles of TZ x TY x TX voxels x NT channels.
//   * activations live in HBM as [cb][Z][Y][X][8] bf16 ("blocked"); a TMA box (2 channel blocks x PY x PX voxels,
//     zero-filled outside the volume = the conv's zero padding) lands in shared memory as two contiguous planes of
//     16-byte voxel rows, which IS the SWIZZLE_NONE K-major UMMA operand layout (row pitch 16 B, LBO = plane size).
//   * every filter tap is the same smem plane read at a row offset (dy*PX + dx): the A descriptor's start address
//     moves, nothing is re-loaded.  128-row M tiles run over the flattened (y, x) plane; rows that fall in the halo
//     columns are computed and discarded (2/PX of the rows).
//   * K loop = 16-channel chunks (outer) x input z-planes (ring of `stages` smem slots) x (dy, dx) taps; the TZ*mt
//     accumulators (NT fp32 columns each, <= 512 columns, two sets when they fit) stay resident in TMEM for the whole
//     loop.  The dz taps are folded into the MMA N dimension: one MMA adds an input plane to up to three adjacent
//     output-plane accumulators.
//   * k = 1 (ConvTranspose GEMM, 1x1 projections): the whole TX x TY x TZ tile is ONE stage (4-D box) whose voxels are the
//     flattened GEMM rows, the [K x NT] weight panel is resident, stages are issued in groups of up to 16.
//   * warp 0: TMA producer, warp 1: MMA issuer (one elected lane issues), warps 2-5: epilogue (TMEM -> registers ->
//     HBM, plus per-channel sum / sum-of-squares partials for InstanceNorm), instantiated per output mode.
// conv3d_roll_kernel (rolling-z, further down): the full-resolution C_out = 32 layers — z segments of a column, a TMEM
//     ring of 16 output planes, all weights resident, per-plane hand-over to 8 epilogue warps.
#include <cuda.h>

#include "common.h"
#include <type_traits>
#include "ptx.cuh"

namespace mmseg {

struct ConvKParams {
  int X, Y, Z, n_img;
  int TX, TY, TZ, PX, PY;
  int tiles_x, tiles_y, tiles_z;
  int halo;
  int mt, NT, n_kchunks, src_cbt, stages, n_ntiles;
  uint32_t stage_bytes, w_bytes, plane_bytes, a_tx_bytes, tmem_cols;
  uint32_t w_off, a_off;  // smem offsets (from the 128-aligned base)
  int out_mode, out_channels, dst_cbt, dst_cb_off, dst_lo_off;
  int desc_swap;
  int fp16;           // 16-bit element format of activations / weights / 16-bit outputs: 0 bf16, 1 fp16 (MMSEG_CONV_FP16)
  int w_stages;       // weight ring slots (2 for k=3: 27 taps per chunk; >= 8 for k=1: tiny chunks, latency-bound)
  int w_resident;     // k=1 only: every K chunk's weights stay in shared memory for the CTA's lifetime (one load)
  int n_tiles;        // voxel tiles (all images) swept by the persistent CTAs of one N tile
  int roll;           // rolling-z kernel (conv3d_roll_kernel): TZ = z-segment length, tiles_z = segments per column
  int kpb;            // rolling-z: K chunks per TMA stage (2: one 4-block box per plane step, half the stage operations)
  int R;              // rolling-z: TMEM ring slots (output planes in flight) = tmem_cols / NT
  int acc_bufs;       // 1 or 2 accumulator sets in TMEM (2: the epilogue of tile i overlaps the MMAs of tile i+1)
  uint32_t buf_cols;  // TMEM columns between the two sets
  long long* dbg;     // optional per-CTA cycle counters (flags bit1): MMA warp waits / epilogue waits
  int dbg_flags;      // timing experiments only (results invalid): bit2 skip epilogue work, bit3 skip TMA after 1st ring pass
  const uint8_t* W;
  const float* bias;
  void* dst;
  float* stats;
  int16_t a_cb[MMSEG_MAX_KCHUNKS];
};

constexpr int kModeConvtHiLo = 100;   // epilogue-internal: MMSEG_OUT_CONVT_K2S2 with a lo plane (parity mode)
constexpr int kMaxStages = 24;
constexpr int kThreads = 192;
constexpr int kMaxMT = 8;
// smem header: barriers + tmem pointer + stats scratch
struct __align__(16) SmemHeader {
  uint64_t a_full[kMaxStages];
  uint64_t a_empty[kMaxStages];
  uint64_t w_full[8];
  uint64_t w_empty[8];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_ptr;
  uint32_t pad;
  float red[4][32];
};
constexpr uint32_t kHeaderBytes = 2048;
static_assert(sizeof(SmemHeader) <= kHeaderBytes, "header too large");

__device__ __forceinline__ size_t blocked_off(int blk, int Z, int Y, int X, int z, int y, int x) {
  return ((((size_t)blk * Z + z) * Y + y) * X + x) * 8;
}

// UMMA with the two descriptor words passed separately: only the low word (start address) changes per MMA.
__device__ __forceinline__ void umma_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
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

// MT = 128-row M tiles per output plane, KS = filter size (3: padding 1, taps along z folded into the MMA N dimension),
// FP16 = 16-bit element format of operands and 16-bit outputs (false: bf16, true: fp16 — same MMA rate, 11-bit significand).
//
// Accumulator (m, zo) lives at TMEM column (m*TZ + zo)*NT, so the accumulators of consecutive output planes are
// adjacent: ONE MMA of N = nz*NT columns adds input plane `pl`'s contribution for the filter taps dz = dz_hi..dz_lo to the
// nz output planes zo = pl-dz_hi .. pl-dz_lo (the weight rows are stored dz-descending).  This triples the MMA N for
// the C_out = 32 / 64 layers (an M=128, K=16 MMA costs >= ~51 clk whatever N is, measured).  All accumulators are
// zeroed by the epilogue warps while the first TMA loads are in flight, so every MMA accumulates.
template <int MT, int KS, bool FP16>
__global__ void __launch_bounds__(kThreads, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ ConvKParams p) {
  constexpr int KT = KS;  // taps per axis
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t w_smem = smem_base + p.w_off;
  const uint32_t a_smem = smem_base + p.a_off;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int nt = blockIdx.y;
  // k=3: one accumulator / one stage per z plane of the tile (+ halo planes).  k=1: the whole TX x TY x TZ tile is ONE
  // stage (a 4-D TMA box) whose voxels are the flattened GEMM rows, so there is a single "plane" of MT M tiles
  constexpr bool FLAT = KS == 1;
  const int acc_z = FLAT ? 1 : p.TZ;
  const int n_planes = acc_z + 2 * p.halo;
  const int tiles_per_img = p.tiles_x * p.tiles_y * p.tiles_z;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&hdr->a_full[s]), 1);
      mbar_init(smem_u32(&hdr->a_empty[s]), 1);
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(smem_u32(&hdr->w_full[s]), 1);
      mbar_init(smem_u32(&hdr->w_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&hdr->acc_full[s]), 1);
      mbar_init(smem_u32(&hdr->acc_empty[s]), 128);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmA);
  if (warp == 2) {
    tmem_alloc(smem_u32(&hdr->tmem_ptr), p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_ptr;

  // PERSISTENT: every role sweeps the same tile sequence blockIdx.x, blockIdx.x + gridDim.x, ...
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t ph = 0, wc = 0, n_loaded = 0;
      if (p.w_resident) {
        const uint32_t wfull = smem_u32(&hdr->w_full[0]);
        mbar_arrive_expect_tx(wfull, p.w_bytes * (uint32_t)p.n_kchunks);
        for (int kc = 0; kc < p.n_kchunks; ++kc)
          bulk_load_1d(w_smem + kc * p.w_bytes, p.W + ((size_t)nt * p.n_kchunks + kc) * p.w_bytes, p.w_bytes, wfull);
      }
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        int t = tile;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y; t /= p.tiles_y;
        const int tz = t % p.tiles_z;
        const int img = t / p.tiles_z;
        const int x0 = tx * p.TX, y0 = ty * p.TY, z0 = tz * p.TZ;
        for (int kc = 0; kc < p.n_kchunks; ++kc, ++wc) {
          if (!p.w_resident) {
            const uint32_t ws = wc % (uint32_t)p.w_stages, wph = (wc / (uint32_t)p.w_stages) & 1;
            const uint32_t wfull = smem_u32(&hdr->w_full[ws]);
            mbar_wait(smem_u32(&hdr->w_empty[ws]), wph ^ 1);
            mbar_arrive_expect_tx(wfull, p.w_bytes);
            bulk_load_1d(w_smem + ws * p.w_bytes, p.W + ((size_t)nt * p.n_kchunks + kc) * p.w_bytes, p.w_bytes, wfull);
          }
          const int cb = img * p.src_cbt + p.a_cb[kc];
          for (int pl = 0; pl < n_planes; ++pl) {
            const int z = z0 - p.halo + pl;
            if (z < 0 || z >= p.Z) continue;
            const uint32_t afull = smem_u32(&hdr->a_full[stage]);
            mbar_wait(smem_u32(&hdr->a_empty[stage]), ph ^ 1);
            if ((p.dbg_flags & 8) && n_loaded >= (uint32_t)p.stages) {
              mbar_arrive(afull);
            } else {
              mbar_arrive_expect_tx(afull, p.a_tx_bytes);
              tma_load_4d(a_smem + stage * p.stage_bytes, &tmA, afull, 2 * (x0 - p.halo), y0 - p.halo, z, cb);
            }
            ++n_loaded;
            if (++stage == p.stages) { stage = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the warp-uniform loops (so the descriptor arithmetic stays in the uniform datapath and the
    // barrier waits are converged); ONE elected lane issues the MMAs and the commits.  The issue stream of that lane is
    // the critical resource of this kernel (measured: every extra dependent instruction per plane shows up 1:1 in the
    // kernel time), hence the per-tap offsets are precomputed and only two adds per MMA remain.
    const bool leader = elect_one();
    constexpr bool fp16 = FP16;
    const uint32_t NT = (uint32_t)p.NT;
    const uint32_t brows = (uint32_t)KT * NT;                   // weight rows per (tap9, k half): dz-descending x NT
    const uint32_t a_hi = (128u >> 4) | (1u << 14);             // SBO = 128 B, descriptor version 1
    const uint32_t b_hi = a_hi;
    const uint32_t a_lbo = (p.plane_bytes >> 4) << 16;          // K halves one plane apart
    const uint32_t b_lbo = ((brows * 16u) >> 4) << 16;
    const uint32_t m_cols = (uint32_t)acc_z * NT;                // TMEM columns between consecutive m tiles
    uint32_t a_tap[KT * KT], b_tap[KT * KT];
#pragma unroll
    for (int i = 0; i < KT * KT; ++i) {
      a_tap[i] = (uint32_t)((i / KT) * p.PX + (i % KT));         // row offset of tap (dy, dx), in 16-byte units
      b_tap[i] = (uint32_t)i * 2u * brows;
    }
    int stage = 0;
    uint32_t ph = 0, wc = 0, it = 0;
    const long long t_begin = clock64();
    long long dbg_wait = 0, dbg_pre = 0, dbg_issue = 0, dbg_acc = 0, dbg_w = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int tz = (tile / (p.tiles_x * p.tiles_y)) % p.tiles_z;
      const int z0 = tz * p.TZ;
      const int tz_valid = FLAT ? 1 : min(p.TZ, p.Z - z0);
      const uint32_t as = p.acc_bufs == 2 ? (it & 1u) : 0u;
      const uint32_t aph = p.acc_bufs == 2 ? ((it >> 1) & 1u) : (it & 1u);
      const long long ca = clock64();
      mbar_wait(smem_u32(&hdr->acc_empty[as]), aph);   // accumulator set drained and re-zeroed by the epilogue
      tc_fence_after();
      dbg_acc += clock64() - ca;
      const uint32_t acc_base = tmem_base + as * p.buf_cols;
      if (KS == 1 && p.w_resident) {
        // k=1 with resident weights: the (K chunk, plane) items of the tile are consecutive ring stages and nothing
        // else gates them, so they are issued in groups of up to 16 stages (one barrier round + one issue pause per
        // group; with one group per K chunk of TZ <= 4 MMAs the ~400-cycle pause cost 3/4 of the tile time, measured)
        if (it == 0) mbar_wait(smem_u32(&hdr->w_full[0]), 0);
        constexpr int G1 = 16;
        const int g1 = min(G1, p.stages);
        const int total = p.n_kchunks * tz_valid;
        const uint32_t idesc1 = make_idesc_16(128, NT, fp16);
        int kc = 0, pl = 0;
        for (int j0 = 0; j0 < total; j0 += g1) {
          const int cnt = min(g1, total - j0);
          if (lane < cnt) {
            int si = stage + lane;
            uint32_t pp = ph;
            if (si >= p.stages) { si -= p.stages; pp ^= 1u; }
            mbar_wait(smem_u32(&hdr->a_full[si]), pp);
          }
          __syncwarp();
          tc_fence_after();
          uint32_t d0[G1], at0[G1], bt0[G1], abar[G1];
#pragma unroll
          for (int h = 0; h < G1; ++h) {
            int si = stage + h;
            if (si >= p.stages) si -= p.stages;
            d0[h] = acc_base + (uint32_t)pl * NT;
            at0[h] = (((a_smem + si * p.stage_bytes) >> 4) & 0x3FFFu) | a_lbo;
            bt0[h] = (((w_smem + (uint32_t)kc * p.w_bytes) >> 4) & 0x3FFFu) | b_lbo;
            abar[h] = smem_u32(&hdr->a_empty[si]);
            if (h < cnt && ++pl == tz_valid) { pl = 0; ++kc; }
          }
          if (leader) {
#pragma unroll
            for (int h = 0; h < G1; ++h) {
              if (h < cnt) {
#pragma unroll
                for (int m = 0; m < MT; ++m)
                  umma_lohi(d0[h] + (uint32_t)m * m_cols, at0[h] + (uint32_t)m * 128u, a_hi, bt0[h], b_hi, idesc1);
                umma_commit(abar[h]);
              }
            }
          }
          __syncwarp();
          stage += cnt;
          if (stage >= p.stages) { stage -= p.stages; ph ^= 1; }
        }
      } else
      for (int kc = 0; kc < p.n_kchunks; ++kc, ++wc) {
        const uint32_t ws = wc % (uint32_t)p.w_stages, wph = (wc / (uint32_t)p.w_stages) & 1;
        const long long cw = clock64();
        mbar_wait(smem_u32(&hdr->w_full[ws]), wph);
        dbg_w += clock64() - cw;
        const uint32_t b_base = (((w_smem + ws * p.w_bytes) >> 4) & 0x3FFFu) | b_lbo;
        // MEASURED (tools/micro/umma_bench.cu): a pause of the issuing lane between MMAs costs ~390 cycles + the pause
        // itself (8 back-to-back N=96 MMAs = 448 cycles; with a 100-cycle pause after them = 934), so barrier waits and
        // descriptor set-up between planes used to double the plane time.  Hence: ALL planes of a K chunk are waited
        // for at once (lane i polls the barrier of plane i, so the ~100-cycle mbarrier latencies overlap), every
        // per-plane descriptor is precomputed, and the elected lane then issues the whole group — up to 10 planes x
        // 9 taps x MT MMAs, with the per-plane commits inline (free) — without a single pause.
        constexpr int GP = 10;
        const int gp = min(GP, p.stages);
        const int pl_lo = max(0, p.halo - z0);                       // first / one-past-last plane inside the volume
        const int pl_hi = min(n_planes, p.Z - z0 + p.halo);
        for (int g0 = pl_lo; g0 < pl_hi; g0 += gp) {
          const int cnt = min(gp, pl_hi - g0);
          const long long c0 = clock64();
          if (lane < cnt) {
            int si = stage + lane;
            uint32_t pp = ph;
            if (si >= p.stages) { si -= p.stages; pp ^= 1u; }
            mbar_wait(smem_u32(&hdr->a_full[si]), pp);
          }
          __syncwarp();
          tc_fence_after();
          const long long c1 = clock64();
          uint32_t idesc[GP], d0[GP], at0[GP], bt0[GP], nzv[GP], abar[GP];
#pragma unroll
          for (int h = 0; h < GP; ++h) {
            int si = stage + h;
            if (si >= p.stages) si -= p.stages;
            const int q = g0 + h;
            const int dz_hi = min(KT - 1, q);
            const int dz_lo = max(0, q - tz_valid + 1);
            nzv[h] = (uint32_t)max(dz_hi - dz_lo + 1, 0);
            idesc[h] = make_idesc_16(128, nzv[h] * NT, fp16);
            d0[h] = acc_base + (uint32_t)(q - dz_hi) * NT;
            at0[h] = (((a_smem + si * p.stage_bytes) >> 4) & 0x3FFFu) | a_lbo;
            bt0[h] = b_base + (uint32_t)(KT - 1 - dz_hi) * NT;   // first weight row of the dz range
            abar[h] = smem_u32(&hdr->a_empty[si]);
          }
          const long long c2 = clock64();
          if (leader) {
#pragma unroll
            for (int h = 0; h < GP; ++h) {
              if (h < cnt) {
                if (nzv[h] > 0) {
#pragma unroll
                  for (int i = 0; i < KT * KT; ++i) {
#pragma unroll
                    for (int m = 0; m < MT; ++m)
                      umma_lohi(d0[h] + (uint32_t)m * m_cols, at0[h] + a_tap[i] + (uint32_t)m * 128u, a_hi, bt0[h] + b_tap[i], b_hi, idesc[h]);
                  }
                }
                umma_commit(abar[h]);
              }
            }
          }
          __syncwarp();
          const long long c3 = clock64();
          dbg_wait += c1 - c0; dbg_pre += c2 - c1; dbg_issue += c3 - c2;
          stage += cnt;
          if (stage >= p.stages) { stage -= p.stages; ph ^= 1; }
        }
        if (leader) umma_commit(smem_u32(&hdr->w_empty[ws]));
        __syncwarp();
      }
      if (leader) umma_commit(smem_u32(&hdr->acc_full[as]));
      __syncwarp();
    }
    if (p.dbg && leader) {
      long long* d = p.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8;
      d[0] = clock64() - t_begin; d[1] = dbg_wait; d[2] = dbg_pre; d[3] = dbg_issue; d[4] = dbg_acc; d[7] = dbg_w;
    }
  } else {
    // ===================== epilogue (4 warps, one TMEM lane quarter each) =====================
    const int q = warp & 3;
    const int ew = warp - 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // zero this warp's quarter of every accumulator while the first loads are in flight
    for (uint32_t c = 0; c < p.tmem_cols; c += 16) tmem_st16_zero(lane_base + c);
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive(smem_u32(&hdr->acc_empty[0]));
    if (p.acc_bufs == 2) mbar_arrive(smem_u32(&hdr->acc_empty[1]));

    const int n_base = nt * p.NT;
    const bool want_stats = p.stats != nullptr;
    // accumulator row L = m*128 + q*32 + lane is tile voxel (yy, xx) = (L / PX, L % PX): position for m = 0 and the step
    // per M tile (no division inside the tile loop)
    // (k=1: rows are the flattened (zz, yy, xx) voxels of the TX x TY x TZ tile, PX = TX, PY = TY)
    const int pxy = p.PX * p.PY;
    const int L0 = q * 32 + lane;
    const int zz0 = FLAT ? L0 / pxy : 0;
    const int yy0 = (L0 - zz0 * pxy) / p.PX, xx0 = (L0 - zz0 * pxy) - yy0 * p.PX;
    const int dz128 = FLAT ? 128 / pxy : 0;
    const int dy128 = (128 - dz128 * pxy) / p.PX, dx128 = (128 - dz128 * pxy) - dy128 * p.PX;
    const int n_cg = p.NT / 16;
    uint32_t it = 0;
    long long e_wait = 0;
    const long long e_begin = clock64();
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      int t = tile;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; t /= p.tiles_y;
      const int tz = t % p.tiles_z;
      const int img = t / p.tiles_z;
      const int tile_in_img = tile - img * tiles_per_img;
      const int x0 = tx * p.TX, y0 = ty * p.TY, z0 = tz * p.TZ;
      const int tz_valid = FLAT ? 1 : min(p.TZ, p.Z - z0);
      const uint32_t as = p.acc_bufs == 2 ? (it & 1u) : 0u;
      const uint32_t aph = p.acc_bufs == 2 ? ((it >> 1) & 1u) : (it & 1u);
      const uint32_t acc_lane = lane_base + as * p.buf_cols;
      const long long eq = clock64();
      mbar_wait(smem_u32(&hdr->acc_full[as]), aph);
      tc_fence_after();
      e_wait += clock64() - eq;
      const size_t plane = (size_t)p.Y * p.X;
      const size_t nvox = plane * p.Z;
      // The column-group loop is instantiated per (output mode, statistics) and selected ONCE per tile: with the mode
      // tested inside the chunk loop the k1 / ConvTranspose / first-layer launches were bound by the ~220 instructions
      // per 16-column chunk of these 4 warps (ncu), not by TMEM or HBM.
      auto epi = [&](auto mode_c, auto stats_c) {
        constexpr int MODE = decltype(mode_c)::value;   // MMSEG_OUT_*; CONVT with a lo plane = kModeConvtHiLo
        constexpr bool STATS = decltype(stats_c)::value;
        constexpr bool F32 = MODE == MMSEG_OUT_BLOCKED_F32 || MODE == MMSEG_OUT_NCDHW_F32;
        constexpr bool CONVT = MODE == MMSEG_OUT_CONVT_K2S2 || MODE == kModeConvtHiLo;
        constexpr size_t ES = F32 ? 4 : 2;
        const bool has_bias = p.bias != nullptr;
        const size_t lo_bytes = (size_t)p.dst_lo_off * nvox * 8 * 2 * (CONVT ? 8 : 1);
        for (int cg = 0; cg < ((p.dbg_flags & 4) ? 0 : n_cg); ++cg) {
          const int n0 = n_base + cg * 16;
          float bias_v[16];
          if (has_bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) bias_v[i] = p.bias[n0 + i];
          }
          float s1[16], s2[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
          // destination of this column group at (z0, y = 0, x = 0) and the byte strides of a voxel step in z / y / x:
          // everything mode-specific about addressing is hoisted out of the chunk loop
          char* cg_base;
          size_t zstride;
          uint32_t ystride, xstride;
          if constexpr (CONVT) {
            const int CB = p.out_channels >> 3;
            const int g = n0 >> 4;
            const int tzy = g / CB, cb = g - tzy * CB;
            const int ct_dz = tzy >> 1, ct_dy = tzy & 1;
            const size_t cbase = (size_t)(img * p.dst_cbt + p.dst_cb_off + cb) * nvox * 8 * 8;   // output volume is 8x larger
            const size_t oplane = (size_t)(2 * p.Y) * (2 * p.X);
            zstride = 2 * oplane * 8 * ES;
            ystride = (uint32_t)(2 * (2 * p.X) * 8 * ES);
            xstride = (uint32_t)(2 * 8 * ES);
            cg_base = reinterpret_cast<char*>(p.dst) +
                      (cbase + ((size_t)(2 * z0 + ct_dz) * oplane + (size_t)ct_dy * (2 * p.X)) * 8) * ES;
          } else if constexpr (MODE == MMSEG_OUT_NCDHW_F32) {
            const size_t cbase = ((size_t)img * p.out_channels + n0) * nvox;
            zstride = plane * ES;
            ystride = (uint32_t)(p.X * ES);
            xstride = (uint32_t)ES;
            cg_base = reinterpret_cast<char*>(p.dst) + (cbase + (size_t)z0 * plane) * ES;
          } else {
            const size_t cbase = (size_t)(img * p.dst_cbt + p.dst_cb_off + (n0 >> 3)) * nvox * 8;
            zstride = plane * 8 * ES;
            ystride = (uint32_t)(p.X * 8 * ES);
            xstride = (uint32_t)(8 * ES);
            cg_base = reinterpret_cast<char*>(p.dst) + (cbase + (size_t)z0 * plane * 8) * ES;
          }
          const size_t half = nvox * 8 * ES;   // blocked layouts: the second 8-channel block of the 16 columns
          // one (zo, m) accumulator chunk: bias, statistics, convert, store
          auto consume = [&](const bool okm, const size_t off, float* v) {
            if (has_bias) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += bias_v[i];
            }
            if (okm) {
              if constexpr (STATS) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { s1[i] += v[i]; s2[i] = fmaf(v[i], v[i], s2[i]); }
              }
              char* o = cg_base + off;
              if constexpr (MODE == MMSEG_OUT_BLOCKED_BF16) {
                *reinterpret_cast<uint4*>(o) = cvt8_from_f32(v, FP16);
                *reinterpret_cast<uint4*>(o + half) = cvt8_from_f32(v + 8, FP16);
              } else if constexpr (MODE == MMSEG_OUT_BLOCKED_F32) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  float4* d = reinterpret_cast<float4*>(o + h * half);
                  d[0] = make_float4(v[8 * h + 0], v[8 * h + 1], v[8 * h + 2], v[8 * h + 3]);
                  d[1] = make_float4(v[8 * h + 4], v[8 * h + 5], v[8 * h + 6], v[8 * h + 7]);
                }
              } else if constexpr (MODE == MMSEG_OUT_BLOCKED_BF16_HILO) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  uint4 hi, lo;
                  split8_from_f32(v + 8 * h, hi, lo, FP16);
                  *reinterpret_cast<uint4*>(o + h * half) = hi;
                  *reinterpret_cast<uint4*>(o + h * half + lo_bytes) = lo;
                }
              } else if constexpr (MODE == MMSEG_OUT_CONVT_K2S2) {
                // column n = (((dz*2 + dy)*CB + cb)*2 + dx)*8 + j: a thread's 16 columns are the SAME 8 output channels
                // at the two x-adjacent output voxels -> one contiguous 32-byte store per thread, 1 KB per warp
                *reinterpret_cast<uint4*>(o) = cvt8_from_f32(v, FP16);
                *reinterpret_cast<uint4*>(o + 16) = cvt8_from_f32(v + 8, FP16);
              } else if constexpr (MODE == kModeConvtHiLo) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  uint4 hi, l;
                  split8_from_f32(v + 8 * h, hi, l, FP16);
                  *reinterpret_cast<uint4*>(o + 16 * h) = hi;
                  *reinterpret_cast<uint4*>(o + 16 * h + lo_bytes) = l;
                }
              } else {  // MMSEG_OUT_NCDHW_F32
                float* of = reinterpret_cast<float*>(o);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  if (n0 + i < p.out_channels) of[(size_t)i * nvox] = v[i];
                }
              }
            }
          };
          // TMEM -> registers: two chunks (planes zo, zo+1 of the same M tile) per step, and the loads of the NEXT step
          // are issued before this step's arithmetic and stores
          auto chunk_addr = [&](const int zo, const int m) {
            return acc_lane + (uint32_t)((m * acc_z + zo) * p.NT + cg * 16);
          };
          if constexpr (FLAT) {
            // one chunk per M tile; the load of tile m+1 is in flight while tile m is converted and stored
            uint32_t ra[16], rb[16];
            int zz = zz0, yy = yy0, xx = xx0;
            auto next_pos = [&](bool& okm, size_t& off) {
              okm = (zz < p.TZ) && (z0 + zz < p.Z) && (y0 + yy < p.Y) && (x0 + xx < p.X);
              off = (size_t)zz * zstride + (uint32_t)(y0 + yy) * ystride + (uint32_t)(x0 + xx) * xstride;
              xx += dx128;
              if (xx >= p.PX) { xx -= p.PX; ++yy; }
              yy += dy128;
              if (yy >= p.PY) { yy -= p.PY; ++zz; }
              zz += dz128;
            };
            tmem_ld16_issue(chunk_addr(0, 0), ra);
#pragma unroll 1
            for (int m = 0; m < MT; m += 2) {
              bool okm;
              size_t off;
              float v[16];
              next_pos(okm, off);
              tmem_ld_wait16(ra);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(ra[i]);
              if (m + 1 < MT) tmem_ld16_issue(chunk_addr(0, m + 1), rb);
              tmem_st16_zero(chunk_addr(0, m));
              consume(okm, off, v);
              if (m + 1 < MT) {
                next_pos(okm, off);
                tmem_ld_wait16(rb);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rb[i]);
                if (m + 2 < MT) tmem_ld16_issue(chunk_addr(0, m + 2), ra);
                tmem_st16_zero(chunk_addr(0, m + 1));
                consume(okm, off, v);
              }
            }
          } else {
            uint32_t ra[16], rb[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) rb[i] = 0u;
            tmem_ld16_issue(chunk_addr(0, 0), ra);
            if (tz_valid > 1) tmem_ld16_issue(chunk_addr(1, 0), rb);
            for (int zo = 0; zo < tz_valid; zo += 2) {
              const bool has_b = zo + 1 < tz_valid;
              // the M-tile loop stays ROLLED (unrolled x MT x modes the epilogue no longer fits the instruction cache:
              // stall_no_inst dominated the MT = 5 logits launch); the row position advances by 128 rows per M tile
              int yy = yy0, xx = xx0;
#pragma unroll 1
              for (int m = 0; m < MT; ++m) {
                const bool okm = (xx < p.TX) && (yy < p.TY) && (x0 + xx < p.X) && (y0 + yy < p.Y);
                const uint32_t row_off = (uint32_t)(y0 + yy) * ystride + (uint32_t)(x0 + xx) * xstride;
                xx += dx128; yy += dy128;
                if (xx >= p.PX) { xx -= p.PX; ++yy; }
                tmem_ld_wait16(ra);
                tmem_ld_tie16(rb);
                float va[16], vb[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { va[i] = __uint_as_float(ra[i]); vb[i] = __uint_as_float(rb[i]); }
                if (m + 1 < MT) {
                  tmem_ld16_issue(chunk_addr(zo, m + 1), ra);
                  if (has_b) tmem_ld16_issue(chunk_addr(zo + 1, m + 1), rb);
                } else if (zo + 2 < tz_valid) {
                  tmem_ld16_issue(chunk_addr(zo + 2, 0), ra);
                  if (zo + 3 < tz_valid) tmem_ld16_issue(chunk_addr(zo + 3, 0), rb);
                }
                // re-zero for the tile after next (every MMA accumulates); the loads of these columns have completed
                tmem_st16_zero(chunk_addr(zo, m));
                if (has_b) tmem_st16_zero(chunk_addr(zo + 1, m));
                consume(okm, (size_t)zo * zstride + row_off, va);
                if (has_b) consume(okm, (size_t)(zo + 1) * zstride + row_off, vb);
              }
            }
          }
          if constexpr (STATS) {
            // 32 per-thread sums (16 columns x {sum, sum of squares}) -> lane l holds the warp total of value l: a
            // halving butterfly (31 shuffles) instead of 32 full reductions (160)
            float r[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) { r[i] = s1[i]; r[16 + i] = s2[i]; }
#pragma unroll
            for (int off = 16, n = 16; off > 0; off >>= 1, n >>= 1) {
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < n; ++i) {
                const float send = up ? r[i] : r[i + n];
                const float keep = up ? r[i + n] : r[i];
                r[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            named_bar_sync(1, 128);  // previous column group's readers are done with hdr->red
            hdr->red[ew][lane] = r[0];
            named_bar_sync(1, 128);
            if (ew == 0) {
              const float tot = hdr->red[0][lane] + hdr->red[1][lane] + hdr->red[2][lane] + hdr->red[3][lane];
              const int ch = n0 + (lane & 15);
              const int C = p.n_ntiles * p.NT;
              float* dstp = p.stats + (((size_t)img * tiles_per_img + tile_in_img) * C + ch) * 2 + (lane >> 4);
              *dstp = tot;
            }
          }
        }
      };
      using std::integral_constant;
      using std::true_type;
      using std::false_type;
      switch (p.out_mode) {
        case MMSEG_OUT_BLOCKED_BF16:
          if (want_stats) epi(integral_constant<int, MMSEG_OUT_BLOCKED_BF16>{}, true_type{});
          else epi(integral_constant<int, MMSEG_OUT_BLOCKED_BF16>{}, false_type{});
          break;
        case MMSEG_OUT_BLOCKED_F32:
          if (want_stats) epi(integral_constant<int, MMSEG_OUT_BLOCKED_F32>{}, true_type{});
          else epi(integral_constant<int, MMSEG_OUT_BLOCKED_F32>{}, false_type{});
          break;
        case MMSEG_OUT_BLOCKED_BF16_HILO:
          if (want_stats) epi(integral_constant<int, MMSEG_OUT_BLOCKED_BF16_HILO>{}, true_type{});
          else epi(integral_constant<int, MMSEG_OUT_BLOCKED_BF16_HILO>{}, false_type{});
          break;
        case MMSEG_OUT_CONVT_K2S2:
          if (p.dst_lo_off > 0) epi(integral_constant<int, kModeConvtHiLo>{}, false_type{});
          else epi(integral_constant<int, MMSEG_OUT_CONVT_K2S2>{}, false_type{});
          break;
        default:
          epi(integral_constant<int, MMSEG_OUT_NCDHW_F32>{}, false_type{});
          break;
      }
      // (planes zo >= tz_valid never receive MMAs and stay zero) hand the re-zeroed set back to the MMA warp
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(smem_u32(&hdr->acc_empty[as]));
    }
    if (p.dbg && warp == 2 && lane == 0) {
      long long* d = p.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8;
      d[5] = clock64() - e_begin; d[6] = e_wait;
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ rolling-z variant
// For the full-resolution C_out = 32 layers (k=3, one 128-row M tile per plane, every K chunk's weights resident in
// shared memory) the tile is a whole z SEGMENT of a (TX x TY) column, and the TMEM accumulators are a RING of
// R = 512 / NT output planes: input plane q (outer loop; K chunks inner) adds its three z taps to ring slots
// q-2 .. q with ONE N = 3*NT MMA per (dy, dx) tap, output plane q-2 is then complete and is handed to the epilogue
// (z_full[slot]) while the MMAs run on, and the slot returns through z_empty[slot].  Versus the TZ = 8 tiles of
// conv3d_tc_kernel this removes the four partial-N boundary planes of every tile (20 % of the MMA cycles: an N = 32
// MMA costs the same ~51 cycles as an N = 96 one) and the per-tile accumulator hand-over.  A ring wrap splits one
// plane's MMAs in two (1 plane in 16).  InstanceNorm partial sums are kept in registers over the whole segment.
struct __align__(16) RollHeader {
  uint64_t a_full[kMaxStages];
  uint64_t a_empty[kMaxStages];
  uint64_t z_full[16];
  uint64_t z_empty[16];
  uint64_t w_full;
  uint32_t tmem_ptr;
  uint32_t pad;
  float red[4][64];
};
constexpr uint32_t kRollHeaderBytes = 2048;
static_assert(sizeof(RollHeader) <= kRollHeaderBytes, "roll header too large");

// warps: 0 TMA, 1 MMA, 2-9 epilogue (two warps per TMEM lane quarter, one 16-column group each: with four warps the
// per-plane epilogue (~700 cycles) was slower than a plane's MMAs for C_in <= 32)
constexpr int kRollThreads = 320;

// F32OUT: raw output blocked fp32 (MMSEG_OUT_BLOCKED_F32: the modes that keep the InstanceNorm input in fp32) instead of
// the 16-bit element format.
template <bool FP16, bool F32OUT>
__global__ void __launch_bounds__(kRollThreads, 1)
conv3d_roll_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ ConvKParams p) {
  constexpr int KT = 3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  RollHeader* hdr = reinterpret_cast<RollHeader*>(smem);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t w_smem = smem_base + p.w_off;
  const uint32_t a_smem = smem_base + p.a_off;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool fp16 = FP16;
  constexpr uint32_t R = 16;   // ring slots: 512 TMEM columns / NT (NT = 32, checked on the host); a power of two so
                               // that the ring arithmetic of the issue loop is shifts and masks, not divisions
  const int ZS = p.TZ;
  const int items_per_img = p.tiles_x * p.tiles_y * p.tiles_z;
  const int n_items = items_per_img * p.n_img;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&hdr->a_full[s]), 1);
      mbar_init(smem_u32(&hdr->a_empty[s]), 1);
    }
    for (int s = 0; s < 16; ++s) {
      mbar_init(smem_u32(&hdr->z_full[s]), 1);
      mbar_init(smem_u32(&hdr->z_empty[s]), (kRollThreads - 64) / 32);   // one arrive per epilogue WARP (per plane: 256
                                                                           // per-thread arrives would swamp the barrier unit)
    }
    mbar_init(smem_u32(&hdr->w_full), 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmA);
  if (warp == 2) {
    tmem_alloc(smem_u32(&hdr->tmem_ptr), p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer: weights once, then (item, plane, K chunk) stages =====================
    if (elect_one()) {
      const uint32_t wfull = smem_u32(&hdr->w_full);
      mbar_arrive_expect_tx(wfull, p.w_bytes * (uint32_t)p.n_kchunks);
      for (int kc = 0; kc < p.n_kchunks; ++kc)
        bulk_load_1d(w_smem + kc * p.w_bytes, p.W + (size_t)kc * p.w_bytes, p.w_bytes, wfull);
      int stage = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int t = item;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y; t /= p.tiles_y;
        const int zs = t % p.tiles_z;
        const int img = t / p.tiles_z;
        const int x0 = tx * p.TX, y0 = ty * p.TY, zs0 = zs * ZS;
        const int zsv = min(ZS, p.Z - zs0);
        for (int q = 0; q < zsv + 2; ++q) {
          const int z = zs0 - 1 + q;
          if (z < 0 || z >= p.Z) continue;
          for (int kc = 0; kc < p.n_kchunks; kc += p.kpb) {   // one box = kpb K chunks = 2*kpb consecutive channel blocks
            const uint32_t afull = smem_u32(&hdr->a_full[stage]);
            mbar_wait(smem_u32(&hdr->a_empty[stage]), ph ^ 1);
            mbar_arrive_expect_tx(afull, p.a_tx_bytes);
            tma_load_4d(a_smem + stage * p.stage_bytes, &tmA, afull, 2 * (x0 - 1), y0 - 1, z, img * p.src_cbt + p.a_cb[kc]);
            if (++stage == p.stages) { stage = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // A software-pipelined per-stage loop: the tensor pipe queues only ~2 MMAs behind the issuing thread (measured,
    // tools/micro/umma_bench.cu: a pause of P cycles costs P - 110), so nothing longer than ~100 cycles may sit between
    // two MMAs.  Hence the descriptors of stage s+1 are computed in the middle of stage s's nine MMAs, stage s+1's
    // "full" barrier is probed there too, and there are no per-group barrier rounds at all.  ONE elected lane runs the
    // whole loop (measured: the same loop run warp-uniformly with the MMAs under `if (leader)` is 35 % slower).
    const bool leader = elect_one();
    if (leader) {
      const uint32_t NT = (uint32_t)p.NT;
      const uint32_t brows = (uint32_t)KT * NT;
      const uint32_t a_hi = (128u >> 4) | (1u << 14);
      const uint32_t b_hi = a_hi;
      const uint32_t a_lbo = (p.plane_bytes >> 4) << 16;
      const uint32_t b_lbo = ((brows * 16u) >> 4) << 16;
      uint32_t a_tap[KT * KT], b_tap[KT * KT];
#pragma unroll
      for (int i = 0; i < KT * KT; ++i) {
        a_tap[i] = (uint32_t)((i / KT) * p.PX + (i % KT));
        b_tap[i] = (uint32_t)i * 2u * brows;
      }
      const uint32_t a_stage0 = ((a_smem >> 4) & 0x3FFFu) | a_lbo;
      const uint32_t a_stage_step = p.stage_bytes >> 4;
      const uint32_t w_base0 = ((w_smem >> 4) & 0x3FFFu) | b_lbo;
      const uint32_t w_step = p.w_bytes >> 4;                  // one K chunk of weights
      const uint32_t a_chunk = 2u * (p.plane_bytes >> 4);      // one K chunk (two channel blocks) inside a stage
      const int kpb = p.kpb;
      const int n_st = p.n_kchunks / kpb;                      // stages per input plane
      const uint32_t a_full0 = smem_u32(&hdr->a_full[0]), a_empty0 = smem_u32(&hdr->a_empty[0]);
      const uint32_t z_full0 = smem_u32(&hdr->z_full[0]), z_empty0 = smem_u32(&hdr->z_empty[0]);
      const int txy = p.tiles_x * p.tiles_y;

      // ---- stage generator state: (item, plane q, K chunk kc), ring stage / phase, ring position gz
      int item = blockIdx.x;
      int zsv = 0, q = 0, q_last = -1, kc = 0;
      uint32_t gz = 0, stage = 0, ph = 0;
      bool more = item < n_items;
      auto enter_item = [&]() {
        const int zs = (item / txy) % p.tiles_z;
        const int zs0 = zs * ZS;
        zsv = min(ZS, p.Z - zs0);
        q = zs0 == 0 ? 1 : 0;                                 // input planes inside the volume: q .. q_last
        q_last = (zs0 + zsv >= p.Z) ? zsv : zsv + 1;
        kc = 0;
      };
      if (more) enter_item();
      // descriptor set of one stage
      struct Stage { uint32_t at0, abar, afull, apar, bta, d0a, ida, btb, idb, zc0, zc1, zneed; };
      auto gen = [&](Stage& d) {   // fills d for the current cursor and advances it; call only while `more`
        const int dz_hi = min(KT - 1, q);
        const int dz_lo = max(0, q - (zsv - 1));
        const uint32_t n = (uint32_t)(dz_hi - dz_lo + 1);
        const uint32_t g_lo = gz + (uint32_t)(q - dz_hi);
        const uint32_t col = g_lo & (R - 1);
        const uint32_t n1 = min(n, R - col);
        d.at0 = a_stage0 + stage * a_stage_step;
        d.abar = a_empty0 + stage * 8u;
        d.afull = a_full0 + stage * 8u;
        d.apar = ph;
        const uint32_t bb = w_base0 + (uint32_t)(kc * kpb) * w_step + (uint32_t)(KT - 1 - dz_hi) * NT;
        d.bta = bb;
        d.d0a = tmem_base + col * NT;
        d.ida = make_idesc_16(128, n1 * NT, fp16);
        d.btb = bb + n1 * NT;
        d.idb = n > n1 ? make_idesc_16(128, (n - n1) * NT, fp16) : 0u;
        d.zneed = gz + (uint32_t)min(q, zsv - 1) + 1u;        // ring slots [.., zneed) must be acquired before this stage
        d.zc0 = 0u; d.zc1 = 0u;
        if (kc == n_st - 1) {                                 // last stage of plane q: output plane q-2 is complete
          if (q >= 2) d.zc0 = z_full0 + ((gz + (uint32_t)(q - 2)) & (R - 1)) * 8u;
          if (q == q_last && q_last == zsv)                   // top halo plane outside the volume: zsv-1 completes too
            d.zc1 = z_full0 + ((gz + (uint32_t)(zsv - 1)) & (R - 1)) * 8u;
        }
        // advance
        if (++stage == (uint32_t)p.stages) { stage = 0; ph ^= 1u; }
        if (++kc == n_st) {
          kc = 0;
          if (++q > q_last) {
            gz += (uint32_t)zsv;
            item += gridDim.x;
            more = item < n_items;
            if (more) enter_item();
          }
        }
      };

      mbar_wait(smem_u32(&hdr->w_full), 0);
      uint32_t g_acq = 0;   // ring slots acquired so far (z_empty waited), a running count over all items
      Stage cur, nxt;
      bool have = more;
      if (have) gen(cur);
      bool cur_ready = false;
      while (have) {
        if (!cur_ready) mbar_wait(cur.afull, cur.apar);
        while (g_acq < cur.zneed) {
          mbar_wait(z_empty0 + (g_acq & (R - 1)) * 8u, (g_acq >> 4) & 1u);
          ++g_acq;
        }
        tc_fence_after();
        const bool have_next = more;
        bool next_ready = false;
#pragma unroll
        for (int i = 0; i < KT * KT; ++i) {
          umma_lohi(cur.d0a, cur.at0 + a_tap[i], a_hi, cur.bta + b_tap[i], b_hi, cur.ida);
          // the descriptor arithmetic for stage s+1 and the probe of its "full" barrier sit between this stage's MMAs
          // (each gap stays below the ~110 cycles the tensor pipe has queued)
          if (i == 2 && have_next) gen(nxt);
          if (i == 6 && have_next) next_ready = mbar_test_wait(nxt.afull, nxt.apar);
        }
        if (cur.idb) {
#pragma unroll
          for (int i = 0; i < KT * KT; ++i) umma_lohi(tmem_base, cur.at0 + a_tap[i], a_hi, cur.btb + b_tap[i], b_hi, cur.idb);
        }
        if (kpb == 2) {   // second K chunk of the stage: channel blocks 2, 3 of the box, next chunk of weights
          const uint32_t a1 = cur.at0 + a_chunk, b1 = cur.bta + w_step, b1b = cur.btb + w_step;
#pragma unroll
          for (int i = 0; i < KT * KT; ++i) umma_lohi(cur.d0a, a1 + a_tap[i], a_hi, b1 + b_tap[i], b_hi, cur.ida);
          if (cur.idb) {
#pragma unroll
            for (int i = 0; i < KT * KT; ++i) umma_lohi(tmem_base, a1 + a_tap[i], a_hi, b1b + b_tap[i], b_hi, cur.idb);
          }
        }
        umma_commit(cur.abar);
        if (cur.zc0) umma_commit(cur.zc0);
        if (cur.zc1) umma_commit(cur.zc1);
        cur = nxt;
        cur_ready = next_ready;
        have = have_next;
      }
    }
  } else {
    // ===================== epilogue (8 warps: TMEM lane quarter = warp & 3, column group = (warp - 2) / 4) ==========
    const int qd = warp & 3;
    const int ew = warp - 2;
    const int cg = ew >> 2;          // 16-column group of this warp (NT = 32: two groups)
    const int w4 = ew & 3;           // rank among the four warps of the group
    const uint32_t lane_base = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)cg * 16u;
    for (uint32_t c = 0; c < p.tmem_cols; c += 32) tmem_st16_zero(lane_base + c);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int s = 0; s < (int)R; ++s) mbar_arrive(smem_u32(&hdr->z_empty[s]));

    const bool want_stats = p.stats != nullptr;
    const int L0 = qd * 32 + lane;
    const int yy = L0 / p.PX, xx = L0 - yy * p.PX;
    const size_t plane = (size_t)p.Y * p.X;
    const size_t nvox = plane * p.Z;
    uint32_t gz = 0;
    long long e_wait = 0;
    const long long e_begin = clock64();
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int t = item;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; t /= p.tiles_y;
      const int zs = t % p.tiles_z;
      const int img = t / p.tiles_z;
      const int item_in_img = item - img * items_per_img;
      const int x0 = tx * p.TX, y0 = ty * p.TY, zs0 = zs * ZS;
      const int zsv = min(ZS, p.Z - zs0);
      const bool ok = (xx < p.TX) && (yy < p.TY) && (x0 + xx < p.X) && (y0 + yy < p.Y);
      constexpr size_t ES = F32OUT ? 4 : 2;
      char* row = reinterpret_cast<char*>(p.dst) +
                  ((size_t)(img * p.dst_cbt + p.dst_cb_off + 2 * cg) * nvox + (size_t)zs0 * plane + (size_t)(y0 + yy) * p.X + (x0 + xx)) * 8 * ES;
      float s1[16], s2[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
      int prev_slot = -1;
      for (int zo = 0; zo < zsv; ++zo) {
        const uint32_t g = gz + (uint32_t)zo;
        const uint32_t slot = (g & (R - 1));
        const long long eq = clock64();
        mbar_wait(smem_u32(&hdr->z_full[slot]), (g >> 4) & 1u);
        tc_fence_after();
        e_wait += clock64() - eq;
        const uint32_t taddr = lane_base + slot * (uint32_t)p.NT;
        uint32_t ra[16];
        tmem_ld16_issue(taddr, ra);
        // the previous plane's re-zeroing store has long completed: hand its slot back now (keeps wait::st off the
        // critical path of that plane)
        if (prev_slot >= 0) {
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&hdr->z_empty[prev_slot]));
        }
        tmem_ld_wait16(ra);
        float va[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) va[i] = __uint_as_float(ra[i]);
        tmem_st16_zero(taddr);
        prev_slot = (int)slot;
        if (ok) {
          if (want_stats) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { s1[i] += va[i]; s2[i] = fmaf(va[i], va[i], s2[i]); }
          }
          char* o = row + (size_t)zo * plane * 8 * ES;
          if constexpr (F32OUT) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float4* d = reinterpret_cast<float4*>(o + (size_t)h * nvox * 8 * ES);
              d[0] = make_float4(va[8 * h + 0], va[8 * h + 1], va[8 * h + 2], va[8 * h + 3]);
              d[1] = make_float4(va[8 * h + 4], va[8 * h + 5], va[8 * h + 6], va[8 * h + 7]);
            }
          } else {
            *reinterpret_cast<uint4*>(o) = cvt8_from_f32(va, fp16);
            *reinterpret_cast<uint4*>(o + nvox * 8 * ES) = cvt8_from_f32(va + 8, fp16);
          }
        }
      }
      if (prev_slot >= 0) {
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&hdr->z_empty[prev_slot]));
      }
      gz += (uint32_t)zsv;
      if (want_stats) {
        float r[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) { r[i] = s1[i]; r[16 + i] = s2[i]; }
#pragma unroll
        for (int off = 16, n = 16; off > 0; off >>= 1, n >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            const float send = up ? r[i] : r[i + n];
            const float keep = up ? r[i + n] : r[i];
            r[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        hdr->red[w4][cg * 32 + lane] = r[0];
        named_bar_sync(1, kRollThreads - 64);
        if (w4 == 0) {   // one warp per column group: lane l<16 = sum of channel l, l>=16 = sum of squares
          const float tot = hdr->red[0][cg * 32 + lane] + hdr->red[1][cg * 32 + lane] + hdr->red[2][cg * 32 + lane] +
                            hdr->red[3][cg * 32 + lane];
          const int ch = cg * 16 + (lane & 15);
          float* dstp = p.stats + (((size_t)img * items_per_img + item_in_img) * p.NT + ch) * 2 + (lane >> 4);
          *dstp = tot;
        }
        named_bar_sync(1, kRollThreads - 64);   // red is reused by the next item
      }
    }
    if (p.dbg && warp == 2 && lane == 0) {
      long long* d = p.dbg + (size_t)blockIdx.x * 8;
      d[5] = clock64() - e_begin; d[6] = e_wait;
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(f);
    }
    cudaGetLastError();
  }
  return fn;
}

static inline uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

struct ConvPlan {
  ConvKParams k;
  uint32_t smem_bytes;
};

static int plan_conv(const mmseg_conv_args* a, ConvPlan* out) {
  if (!a) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: null args");
  if (a->ksize != 1 && a->ksize != 3) return fail(MMSEG_ERR_UNSUPPORTED, "conv3d: ksize %d (only 1, 3)", a->ksize);
  if (a->n_img < 1 || a->X < 1 || a->Y < 1 || a->Z < 1) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: bad extents");
  if (a->NT < 16 || a->NT > 256 || (a->NT % 16)) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: NT=%d must be a multiple of 16 in [16,256]", a->NT);
  if (a->n_ntiles < 1) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: n_ntiles");
  if (a->n_kchunks < 1 || a->n_kchunks > MMSEG_MAX_KCHUNKS) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: n_kchunks=%d", a->n_kchunks);
  if (a->TX < 1 || a->TY < 1 || a->TZ < 1) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: bad tile");
  if (a->stages < 2 || a->stages > kMaxStages) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: stages=%d", a->stages);
  if (a->out_mode < 0 || a->out_mode > MMSEG_OUT_NCDHW_F32) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: out_mode");
  if (a->out_mode == MMSEG_OUT_CONVT_K2S2 && (a->out_channels % 16)) return fail(MMSEG_ERR_UNSUPPORTED, "convT: out_channels must be a multiple of 16");
  if (a->out_mode == MMSEG_OUT_CONVT_K2S2 && a->ksize != 1) return fail(MMSEG_ERR_INVALID_ARG, "convT runs as ksize 1");
  for (int i = 0; i < a->n_kchunks; ++i)
    if (a->a_cb[i] < 0 || a->a_cb[i] + 2 > a->src_cbt) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: a_cb[%d]=%d outside src_cbt=%d", i, a->a_cb[i], a->src_cbt);
  ConvKParams& k = out->k;
  k.X = a->X; k.Y = a->Y; k.Z = a->Z; k.n_img = a->n_img;
  k.halo = a->ksize / 2;
  k.TX = a->TX; k.TY = a->TY; k.TZ = a->TZ;
  k.PX = a->TX + 2 * k.halo; k.PY = a->TY + 2 * k.halo;
  if (k.PX > 128) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: TX+halo=%d > 128 (TMA box limit)", k.PX);
  if (k.PY > 256) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: TY too large");
  k.tiles_x = (a->X + a->TX - 1) / a->TX;
  k.tiles_y = (a->Y + a->TY - 1) / a->TY;
  k.tiles_z = (a->Z + a->TZ - 1) / a->TZ;
  // k=3: rows of one padded plane up to the last needed voxel; k=1: every voxel of the TX x TY x TZ tile (one stage)
  const int flat = a->ksize == 1 ? a->TX * a->TY * a->TZ : (a->TY - 1) * k.PX + a->TX;
  k.mt = (flat + 127) / 128;
  k.roll = (a->flags & MMSEG_CONV_ROLL_Z) ? 1 : 0;
  if (k.roll) {
    if (a->ksize != 3 || (a->out_mode != MMSEG_OUT_BLOCKED_BF16 && a->out_mode != MMSEG_OUT_BLOCKED_F32) || a->n_ntiles != 1 ||
        a->NT != 32 || a->dst_lo_off != 0 || a->bias)
      return fail(MMSEG_ERR_UNSUPPORTED, "conv3d: rolling-z needs ksize 3, NT = C_out = 32, blocked 16-bit or fp32 raw output, no bias");
    if (k.mt != 1) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: rolling-z needs one M tile per plane (got %d)", k.mt);
  }
  const int n_acc = k.roll ? 512 / a->NT : (a->ksize == 1 ? k.mt : k.mt * a->TZ);
  if (a->ksize == 1 && a->TZ > 256) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: TZ=%d > 256 (TMA box)", a->TZ);
  if (k.mt > kMaxMT) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: %d M tiles per plane > %d", k.mt, kMaxMT);
  if (a->ksize == 3 && 3 * a->NT > 256) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: NT=%d > 80 with ksize 3 (folded MMA N = 3*NT <= 256)", a->NT);
  const int cols = n_acc * a->NT;
  if (cols > 512) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: %d TMEM columns > 512", cols);
  // two accumulator sets when they fit: the epilogue of one tile overlaps the MMAs of the next
  k.acc_bufs = (k.roll || 2 * cols > 512) ? 1 : 2;
  k.R = k.roll ? 16 : 0;
  k.buf_cols = (uint32_t)cols;
  uint32_t tc = 32;
  while ((int)tc < cols * k.acc_bufs) tc <<= 1;
  k.tmem_cols = tc;
  k.NT = a->NT; k.n_ntiles = a->n_ntiles; k.n_kchunks = a->n_kchunks; k.src_cbt = a->src_cbt; k.stages = a->stages;
  const int taps = a->ksize * a->ksize * a->ksize;
  k.w_bytes = (uint32_t)taps * a->NT * 32u;
  k.plane_bytes = (uint32_t)k.PX * k.PY * 16u * (a->ksize == 1 ? (uint32_t)a->TZ : 1u);   // one K half of a stage
  if ((k.plane_bytes >> 4) > 0x3FFF) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: plane too large for LBO");
  k.kpb = 1;
  if (k.roll && (a->flags & MMSEG_CONV_ROLL_KPAIR)) {
    if (a->n_kchunks % 2) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: paired K chunks need an even chunk count");
    for (int i = 0; i + 1 < a->n_kchunks; i += 2)
      if (a->a_cb[i + 1] != a->a_cb[i] + 2) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: paired K chunks must be adjacent channel blocks");
    k.kpb = 2;
  }
  k.a_tx_bytes = 2u * k.plane_bytes * (uint32_t)k.kpb;
  k.stage_bytes = round_up(k.a_tx_bytes, 128);
  const uint32_t rows_needed = (uint32_t)(k.mt * 128 + 2 * k.halo * k.PX + 2 * k.halo);
  const uint32_t overflow = rows_needed * 16u > k.plane_bytes ? rows_needed * 16u - k.plane_bytes : 0u;
  k.w_off = kHeaderBytes;
  // k=1: the whole [K x NT] weight panel of this CTA's column tile is a few KB -> keep it resident (one load per CTA,
  // no per-chunk weight barrier, and the MMA warp can issue all (chunk, plane) items of a tile as one group)
  k.w_resident = a->ksize == 1 && (uint32_t)a->n_kchunks * round_up(k.w_bytes, 128) <= 64u * 1024u;
  k.w_stages = a->ksize == 1 ? (k.w_resident && a->n_kchunks > 8 ? a->n_kchunks : 8) : 2;
  if (k.roll) {   // every K chunk's 27-tap weights resident
    k.w_resident = 1;
    k.w_stages = a->n_kchunks;
    static_assert(kRollHeaderBytes == kHeaderBytes, "roll header shares the offset arithmetic");
  }
  k.a_off = k.w_off + k.w_stages * round_up(k.w_bytes, 128);
  uint32_t total = k.a_off + a->stages * k.stage_bytes + round_up(overflow, 128) + 128 /*align slack*/;
  // two CTAs share an SM only when both fit in TMEM: a CTA that needs more than 256 columns asks for more than half of
  // the shared memory so that a second CTA can never be co-resident and block in tcgen05.alloc
  if (tc > 256 && total < 116u * 1024u) total = 116u * 1024u;
  if (total > 227u * 1024u) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: %u bytes of shared memory > 227 KB", total);
  if (k.w_bytes >= (1u << 20) || k.a_tx_bytes >= (1u << 20)) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: tx bytes");
  out->smem_bytes = total;
  k.out_mode = a->out_mode; k.out_channels = a->out_channels;
  k.dst_cbt = a->dst_cbt; k.dst_cb_off = a->dst_cb_off; k.dst_lo_off = a->dst_lo_off;
  k.desc_swap = a->flags & 1;
  k.fp16 = (a->flags & MMSEG_CONV_FP16) ? 1 : 0;
  k.dbg_flags = a->flags;
  k.dbg = (a->flags & 2) ? reinterpret_cast<long long*>(a->stats_partial) : nullptr;  // debug: counters replace stats
  k.n_tiles = k.tiles_x * k.tiles_y * k.tiles_z * a->n_img;
  k.W = reinterpret_cast<const uint8_t*>(a->weights); k.bias = a->bias; k.dst = a->dst;
  k.stats = (a->flags & 2) ? nullptr : a->stats_partial;
  for (int i = 0; i < MMSEG_MAX_KCHUNKS; ++i) k.a_cb[i] = i < a->n_kchunks ? a->a_cb[i] : 0;
  return MMSEG_OK;
}

}  // namespace mmseg

using namespace mmseg;

extern "C" int64_t mmseg_conv3d_smem_bytes(const mmseg_conv_args* a) {
  ConvPlan pl;
  int rc = plan_conv(a, &pl);
  if (rc != MMSEG_OK) return rc;
  return pl.smem_bytes;
}

extern "C" int32_t mmseg_conv3d_tiles_per_img(const mmseg_conv_args* a) {
  ConvPlan pl;
  int rc = plan_conv(a, &pl);
  if (rc != MMSEG_OK) return rc;
  return pl.k.tiles_x * pl.k.tiles_y * pl.k.tiles_z;
}

extern "C" int mmseg_conv3d_fwd(const mmseg_conv_args* a, void* stream) {
  ConvPlan pl;
  int rc = plan_conv(a, &pl);
  if (rc != MMSEG_OK) return rc;
  if (!a->src || !a->weights || !a->dst) return fail(MMSEG_ERR_INVALID_ARG, "conv3d: null pointer");
  if ((reinterpret_cast<uintptr_t>(a->src) & 15) || (reinterpret_cast<uintptr_t>(a->weights) & 15) ||
      (reinterpret_cast<uintptr_t>(a->dst) & 15))
    return fail(MMSEG_ERR_INVALID_ARG, "conv3d: pointers must be 16-byte aligned");
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(MMSEG_ERR_NO_DRIVER, "conv3d: cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  const ConvKParams& k = pl.k;
  CUtensorMap tm;
  // activations viewed as 8-byte elements: dim0 = 2*X (one voxel's 8 bf16 channels = 2 elements)
  cuuint64_t dims[4] = {(cuuint64_t)2 * k.X, (cuuint64_t)k.Y, (cuuint64_t)k.Z, (cuuint64_t)k.n_img * k.src_cbt};
  cuuint64_t strides[3] = {(cuuint64_t)k.X * 16, (cuuint64_t)k.X * k.Y * 16, (cuuint64_t)k.X * k.Y * k.Z * 16};
  cuuint32_t box[4] = {(cuuint32_t)(2 * k.PX), (cuuint32_t)k.PY, (cuuint32_t)(a->ksize == 1 ? k.TZ : 1), (cuuint32_t)(2 * k.kpb)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(a->src), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(MMSEG_ERR_CUDA, "conv3d: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  typedef void (*KernelFn)(const CUtensorMap, const ConvKParams);
#define MMSEG_ROW(KS, F) {conv3d_tc_kernel<1, KS, F>, conv3d_tc_kernel<2, KS, F>, conv3d_tc_kernel<3, KS, F>, conv3d_tc_kernel<4, KS, F>, \
                          conv3d_tc_kernel<5, KS, F>, conv3d_tc_kernel<6, KS, F>, conv3d_tc_kernel<7, KS, F>, conv3d_tc_kernel<8, KS, F>}
  static const KernelFn table[2][2][kMaxMT] = {{MMSEG_ROW(1, false), MMSEG_ROW(3, false)}, {MMSEG_ROW(1, true), MMSEG_ROW(3, true)}};
#undef MMSEG_ROW
  static bool attr_set = false;
  if (!attr_set) {
    for (int f = 0; f < 2; ++f)
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < kMaxMT; ++j) {
          cudaError_t e = cudaFuncSetAttribute(table[f][i][j], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
          if (e != cudaSuccess) return fail(MMSEG_ERR_CUDA, "conv3d: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        }
    attr_set = true;
  }
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sms <= 0) n_sms = 148;
  }
  if (k.roll) {
    typedef void (*RollFn)(const CUtensorMap, const ConvKParams);
    static const RollFn roll_fns[2][2] = {{conv3d_roll_kernel<false, false>, conv3d_roll_kernel<false, true>},
                                          {conv3d_roll_kernel<true, false>, conv3d_roll_kernel<true, true>}};
    static bool roll_attr = false;
    if (!roll_attr) {
      for (int i = 0; i < 4; ++i) {
        cudaError_t e = cudaFuncSetAttribute(roll_fns[i >> 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return fail(MMSEG_ERR_CUDA, "conv3d: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      }
      roll_attr = true;
    }
    const int n_ctas = k.n_tiles < n_sms ? k.n_tiles : n_sms;
    roll_fns[k.fp16][a->out_mode == MMSEG_OUT_BLOCKED_F32 ? 1 : 0]<<<n_ctas, kRollThreads, pl.smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(tm, k);
    return check_launch("conv3d_roll_kernel");
  }
  // persistent CTAs: about one per SM in total, each sweeping its share of the voxel tiles of one N tile
  int ctas = n_sms / k.n_ntiles;
  if (ctas < 1) ctas = 1;
  if (ctas > k.n_tiles) ctas = k.n_tiles;
  dim3 grid((unsigned)ctas, (unsigned)k.n_ntiles);
  table[k.fp16][a->ksize == 3 ? 1 : 0][k.mt - 1]<<<grid, kThreads, pl.smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(tm, k);
  return check_launch("conv3d_tc_kernel");
}
