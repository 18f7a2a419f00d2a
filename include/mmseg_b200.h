/*
 * mmseg_b200 — C ABI of the B200 (sm_100a) kernels behind the multimodal-organ-segmentation hot path.
 *
 * The reference (wittyseok/multimodal-organ-segmentation) is pure PyTorch and has no FFI of its own; the
 * boundary it binds for this path is "torch.nn module -> ATen op".  Each entry point below therefore cites
 * the reference call site whose ATen ops it replaces (paths relative to the reference root).  INTEGRATION.md
 * shows the ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *  - plain C types only: device pointers as void*, sizes as int32/int64, the CUDA stream as void* (cudaStream_t).
 *  - the caller (PyTorch) owns ALL memory, including workspaces; the library never allocates device memory.
 *  - every call is asynchronous on `stream`; return 0 on success, negative mmseg_status on rejected arguments or
 *    launch failure; mmseg_last_error() returns a thread-local message.  No exceptions cross the ABI.
 *  - spatial axes are named (Z, Y, X) = tensor dims (2, 3, 4) of the reference's [B, C, H, W, D]; X is stride-1.
 *
 * Device data layouts
 *  - "blocked" activations: [n_img * cbt][Z][Y][X][8] bf16 (or fp16: MMSEG_FMT_FP16 / MMSEG_CONV_FP16, the "fp16"
 *    numeric modes; every comment below that says bf16 means "the 16-bit element format") — channels in blocks of 8
 *    (16 bytes per voxel per block);
 *    cbt = channel blocks per image held by the buffer (a buffer may hold a concat of several producers, and, in the
 *    3-pass "parity" numeric mode, a bf16 hi plane followed by a bf16 lo plane of the same channels).
 *  - raw conv output before InstanceNorm: same blocked shape, bf16 (fast mode) or fp32 (parity mode).
 *  - packed conv weights: [n_ntiles][n_kchunks][taps][2][NT][8] bf16 (K-major, SWIZZLE_NONE UMMA core matrices).
 */
#ifndef MMSEG_B200_H_
#define MMSEG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMSEG_ABI_VERSION 2
/* 16-bit element format of blocked activations / packed weights (`fmt` arguments): the tensor cores run kind::f16 on
 * either at the same rate; fp16 carries an 11-bit significand (bf16: 8) at a narrower exponent range. */
#define MMSEG_FMT_BF16 0
#define MMSEG_FMT_FP16 1
#define MMSEG_MAX_KCHUNKS 256
#define MMSEG_MAX_WGRAD_GROUPS 64

typedef enum {
  MMSEG_OK = 0,
  MMSEG_ERR_INVALID_ARG = -1,
  MMSEG_ERR_UNSUPPORTED = -2,
  MMSEG_ERR_CUDA = -3,
  MMSEG_ERR_NO_DRIVER = -4
} mmseg_status;

/* epilogue / output modes of mmseg_conv3d_fwd */
enum {
  MMSEG_OUT_BLOCKED_BF16 = 0,   /* blocked bf16 (raw conv output in fast mode, or activation)            */
  MMSEG_OUT_BLOCKED_F32 = 1,    /* blocked fp32 (raw conv output in parity mode)                          */
  MMSEG_OUT_BLOCKED_BF16_HILO = 2, /* blocked bf16 hi plane + lo plane at dst_lo_off (activation, parity) */
  MMSEG_OUT_CONVT_K2S2 = 3,     /* ConvTranspose3d(k2,s2) pixel-shuffle scatter into a blocked bf16 buffer;
                                   GEMM column n = (((dz*2+dy)*CB + cb)*2 + dx)*8 + j, CB = out_channels/8        */
  MMSEG_OUT_NCDHW_F32 = 4       /* [n_img][out_channels][Z][Y][X] fp32 (logits)                            */
};

int mmseg_version(void);
const char* mmseg_last_error(void);
/* 1 when the process can see a CUDA device of compute capability 10.x, else 0 (never raises). */
int mmseg_device_ok(void);
/* sizeof() of the ABI's argument structs (0 mmseg_conv_args, 1 mmseg_wgrad_args, 2 mmseg_norm_args, 3 mmseg_norm_bwd_args,
 * 4 mmseg_adamw_tensor, 5 mmseg_repack_desc, 6 mmseg_swin_attn_args; -1 otherwise): lets a binding verify its layout. */
int mmseg_sizeof(int which);

/*
 * Conv3d forward as a tcgen05/TMEM implicit GEMM fed by TMA halo tiles.
 * Replaces: nn.Conv3d(k=3,p=1) in ConvBlock3D (src/models/backbones/unet.py:26-27,54,57), nn.Conv3d(k=1)
 * (unet.py:163; src/models/backbones/dual_encoder.py:75,84; src/models/fusion/attention_fusion.py:108-113) and, with
 * MMSEG_OUT_CONVT_K2S2, nn.ConvTranspose3d(k=2,s=2) in UpBlock3D (unet.py:95,105).  torch.cat([up, skip], 1)
 * (unet.py:111) costs nothing: a_cb[] lets one conv read its K chunks from any channel blocks of a shared buffer.
 *
 * GEMM view: M = output voxels (128-row tiles over a flattened halo tile), N = NT output channels per CTA,
 * K = taps x 16-channel chunks.  Accumulators stay in TMEM for the whole K loop.
 * stats_partial (optional): per-CTA per-channel (sum, sum of squares) of the fp32 accumulators over valid voxels —
 * the InstanceNorm3d statistics (unet.py:34-35) are produced by the conv epilogue, deterministically (no atomics).
 */
#define MMSEG_CONV_ROLL_Z 16
#define MMSEG_CONV_ROLL_KPAIR 32   /* rolling-z: one TMA stage = two adjacent K chunks (4 channel blocks) */
#define MMSEG_CONV_FP16 64         /* activations, weights and 16-bit outputs are fp16 instead of bf16 (MMSEG_FMT_FP16) */
typedef struct {
  const void* src;       /* blocked bf16 [n_img*src_cbt][Z][Y][X][8]                                  */
  const void* weights;   /* packed bf16, see layout above                                             */
  const float* bias;     /* [n_ntiles*NT] fp32 or NULL                                                */
  void* dst;             /* per out_mode                                                              */
  float* stats_partial;  /* [n_img][tiles_per_img][n_ntiles*NT][2] fp32 or NULL                       */
  int32_t n_img, Z, Y, X;
  int32_t src_cbt;       /* channel blocks per image in src                                           */
  int32_t ksize;         /* 3 (padding 1) or 1                                                        */
  int32_t n_kchunks;     /* K chunks of 16 channels (<= MMSEG_MAX_KCHUNKS)                            */
  int32_t NT, n_ntiles;  /* N tile (multiple of 16, <= 256) and number of N tiles (grid.y)            */
  int32_t TX, TY, TZ;    /* output tile per CTA; TX + 2*(ksize/2) <= 128                              */
  int32_t stages;        /* depth of the activation-plane ring in shared memory (>= 2)                */
  int32_t out_mode;
  int32_t out_channels;  /* real channels (masks N padding); for CONVT: channels per tap              */
  int32_t dst_cbt, dst_cb_off, dst_lo_off; /* blocked destinations: blocks per image, first block, lo-plane offset */
  int32_t flags;         /* bit0: debug — swap LBO/SBO in the smem descriptors; MMSEG_CONV_ROLL_Z (16): run the
                            rolling-accumulator kernel — TZ is then the z-SEGMENT length of a (TX x TY) column
                            (needs ksize 3, NT = C_out = 32, one 128-row M tile per plane, blocked 16-bit or fp32
                            raw output, no bias, all K-chunk weights resident in shared memory; see conv_tc.cu);
                            MMSEG_CONV_FP16 (64): fp16 instead of bf16 elements                               */
  int16_t a_cb[MMSEG_MAX_KCHUNKS]; /* first channel block (of 2) in src for each K chunk              */
} mmseg_conv_args;

int mmseg_conv3d_fwd(const mmseg_conv_args* args, void* stream);
/* bytes of dynamic shared memory / TMEM columns / CTAs the call would use; <0 when the tiling is rejected */
int64_t mmseg_conv3d_smem_bytes(const mmseg_conv_args* args);
int32_t mmseg_conv3d_tiles_per_img(const mmseg_conv_args* args);

/*
 * Conv3d weight gradient (the wgrad half of autograd's convolution_backward for nn.Conv3d k=3 p=1 / k=1 at
 * src/models/backbones/unet.py:26-27,54,57,163 and, through the k=1 GEMM view, nn.ConvTranspose3d(k2,s2) at unet.py:95;
 * the reference reaches it via loss.backward(), src/trainer/trainer.py:243).
 *   dW[co, ci, tap] = sum over images and voxels of dY[vox, co] * X[vox + tap, ci]
 * tcgen05 GEMM with the voxels as the contraction dimension, both operands MN-major straight from the blocked layout.
 * grid = n_part persistent CTAs x (n_cig * n_cot) channel-group pairs; every CTA sweeps its share of the voxel tiles
 * with the accumulators resident in TMEM and writes ONE fp32 partial; mmseg_wgrad_reduce sums the partials in a fixed
 * order (deterministic split-K) into the PyTorch-layout gradient.  dgrad needs no entry point of its own: it is
 * mmseg_conv3d_fwd with the spatially flipped, channel-transposed weights.
 */
typedef struct {
  const void* x;        /* conv input, blocked bf16 [n_img*x_cbt][Z][Y][X][8]                                 */
  const void* dy;       /* gradient of the raw conv output, blocked bf16 [n_img*y_cbt][Z][Y][X][8]            */
  float* partial;       /* workspace [n_cig*n_cot][n_part][128][ksize^2 * cot_blocks*8] fp32                  */
  int32_t n_img, Z, Y, X;
  int32_t ksize;        /* 3 (padding 1) or 1                                                                 */
  int32_t TX, TY, TZ;   /* voxel tile per sweep step; TX*TY % 16 == 0, TX <= 128                              */
  int32_t cig_blocks;   /* input-channel blocks (of 8) per group; ksize*cig_blocks <= 16                      */
  int32_t cot_blocks;   /* output-channel blocks per group (even); ksize*cot_blocks*8 <= 256                  */
  int32_t n_cig, n_cot; /* number of input / output channel groups                                            */
  int32_t x_cbt, y_cbt; /* channel blocks per image in x / dy                                                 */
  int32_t y_cb0;        /* first channel block of dy used                                                     */
  int32_t n_part;       /* persistent CTAs per group pair                                                     */
  int16_t x_cb[MMSEG_MAX_WGRAD_GROUPS]; /* first channel block in x of each input-channel group (concat-aware) */
} mmseg_wgrad_args;
int mmseg_conv3d_wgrad(const mmseg_wgrad_args* args, void* stream);
int64_t mmseg_conv3d_wgrad_smem_bytes(const mmseg_wgrad_args* args);
/* ci_of_pos[group*CIG + index inside the group] = weight input channel held by that accumulator row, or -1 for a padded
 * row (CIG = cig_blocks*8; n_cig groups).  dst must be zero-initialised only where no accumulator row maps (never: every
 * real (co, ci, tap) is written exactly once).  transposed: ConvTranspose3d(k2,s2) layout [Cin][Cout][2][2][2] from the
 * GEMM columns n = tap8*Cout + co (Cout_gemm = 8*Cout). */
int mmseg_wgrad_reduce(const float* partial, int32_t n_part, int32_t ksize, int32_t cig_blocks, int32_t cot_blocks,
                       int32_t n_cig, int32_t n_cot, int32_t Cin, int32_t Cout_gemm, int32_t Cout, int32_t transposed,
                       const int32_t* ci_of_pos, float* dst, void* stream);

/*
 * InstanceNorm3d(affine=False, eps) statistics: reduce the conv epilogue's per-CTA partials in a fixed order (fp64)
 * into mean / rstd per (image, channel).  Replaces the statistics half of nn.InstanceNorm3d (unet.py:34-35,55,58).
 */
int mmseg_instnorm_finalize(const float* stats_partial, int32_t n_img, int32_t tiles_per_img, int32_t channels,
                            int64_t voxels, float eps, float* mean_rstd /* [n_img][channels][2] */, void* stream);

/*
 * GroupNorm(groups, channels) statistics (ConvBlock3D norm="group": nn.GroupNorm(8, C), unet.py:36-38) from the conv
 * epilogue's partials (the conv must have added its bias): writes the apply kernel's table (mean_g, rstd_g * gamma_c)
 * and shift = beta_c.  gamma / beta may be NULL (no affine).
 */
int mmseg_groupnorm_finalize(const float* stats_partial, int32_t n_img, int32_t tiles_per_img, int32_t channels,
                             int32_t groups, int64_t voxels, float eps, const float* gamma, const float* beta,
                             float* mean_rstd /* [n_img][channels][2] */, float* shift /* [n_img][channels] */,
                             void* stream);

/*
 * y = act((x - mean) * rstd) over a blocked tensor, act = ReLU (slope 0) / LeakyReLU(slope); writes bf16 (hi[, lo]).
 * Replaces nn.InstanceNorm3d apply + nn.ReLU (unet.py:45,55-59).  src_is_f32 selects the raw dtype.
 * pooled (optional): also writes MaxPool3d(2) (unet.py:73,77) of y into a second blocked buffer.
 * The struct is zero-initialised by callers; the trailing fused-finalize fields default to "off".
 */
typedef struct {
  const void* src;         /* raw conv output, blocked [n_img*cb][Z][Y][X][8] bf16 or fp32                 */
  const float* mean_rstd;  /* [n_img][cb*8][2]                                                             */
  void* dst;               /* blocked bf16 [n_img*dst_cbt]...                                              */
  void* pooled;            /* blocked bf16 [n_img*pool_cbt][Z/2][Y/2][X/2][8] or NULL                      */
  int32_t n_img, cb, Z, Y, X;
  int32_t src_is_f32;
  int32_t dst_cbt, dst_cb_off, dst_lo_off;
  int32_t pool_cbt, pool_cb_off, pool_lo_off;
  float slope;
  /* fused statistics finalize (optional, replaces the mmseg_instnorm_finalize launch): when stats_partial is set,
   * every block derives mean / rstd of its 8 channels from the conv epilogue's partials (fixed order, fp64) and
   * mean_rstd is not read; mean_rstd_out (optional) receives the table for later consumers (backward). */
  const float* stats_partial; /* [n_img][tiles_per_img][cb*8][2] or NULL                                   */
  float* mean_rstd_out;       /* [n_img][cb*8][2] or NULL                                                  */
  int32_t tiles_per_img;
  float eps;
  const float* shift;         /* [n_img][cb*8] or NULL: y = (x - mean) * rstd + shift before the activation — the
                                 affine norms of ConvBlock3D (GroupNorm / BatchNorm, unet.py:30-38) fold gamma into
                                 rstd and pass beta here; not combined with stats_partial                         */
  int32_t act;                /* 0: ReLU / LeakyReLU(slope); 1: exact GELU (ConvBlock3D activation="gelu", unet.py:47-48) */
  int32_t elem_fmt;           /* MMSEG_FMT_*: element type of the 16-bit tensors (src when !src_is_f32, dst, pooled) */
} mmseg_norm_args;
int mmseg_instnorm_act_apply(const mmseg_norm_args* args, void* stream);

/*
 * Backward of InstanceNorm3d(affine=False) + ReLU/LeakyReLU (+ the MaxPool3d(2) that may follow) — the
 * native_batch_norm_backward / threshold_backward / max_pool3d_with_indices_backward ops autograd issues for
 * ConvBlock3D / DownBlock3D (unet.py:53-60,76-79 via trainer.py:243).  Two launches: _reduce writes per-block partial
 * sums of g' and g'*y^, _apply re-reduces them in a fixed order and writes dx (bf16, blocked) = gradient of the raw
 * conv output.  gA: gradient w.r.t. the activation (scaled by gA_scale, e.g. 1/M for DualEncoder's mean fusion);
 * gP: gradient w.r.t. the pooled activation (routed to the first maximum of each 2x2x2 cell).  Either may be NULL.
 */
typedef struct {
  const void* x;           /* raw conv output saved by the forward, blocked bf16 [n_img*cb][Z][Y][X][8]       */
  const float* mean_rstd;  /* [n_img][cb*8][2] saved by the forward                                           */
  const void* gA;          /* blocked bf16 [n_img*gA_cbt][Z][Y][X][8] or NULL                                 */
  const void* gP;          /* blocked bf16 [n_img*gP_cbt][Z/2][Y/2][X/2][8] or NULL                           */
  float* partial;          /* [n_img*cb][n_chunks][16] fp32                                                   */
  void* dx;                /* blocked bf16 [n_img*dx_cbt][Z][Y][X][8] (apply only)                            */
  const float* chan_scale; /* optional [n_img][cb*8] multiplier of gA (Dropout3d mask * 1/(1-p); gate weight) or NULL */
  const float* chan_bias;  /* optional [n_img][cb*8] constant added to the activation gradient (gate: dpooled/N) or NULL */
  int32_t n_img, cb, Z, Y, X;
  int32_t gA_cbt, gA_cb_off, gP_cbt, gP_cb_off, dx_cbt, dx_cb_off;
  int32_t n_chunks;
  float gA_scale, slope;
  const float* m12;        /* optional [n_img][cb*8][2] (apply only): dx = rstd * (g' - m12[0] - y^ * m12[1]) with these terms instead
                              of the per-(image, channel) means of `partial` — how GroupNorm / BatchNorm / affine norms reuse the
                              kernel: the caller folds gamma / beta into mean_rstd and combines the sums over the norm's
                              reduction set (train_engine.py) */
} mmseg_norm_bwd_args;
int mmseg_instnorm_act_bwd_reduce(const mmseg_norm_bwd_args* args, void* stream);
int mmseg_instnorm_act_bwd_apply(const mmseg_norm_bwd_args* args, void* stream);
/* out[b][m] = sum over channels and voxels of g[b, c, v] * x[b, m*C + c, v]: gradient w.r.t. the modality-gate weights
 * of CrossModalAttention (src/models/backbones/dual_encoder.py:251-252).  partial: [n_img*M*cb][n_chunks] fp32. */
int mmseg_modality_dot(const void* x, int32_t x_cbt, const void* g, int32_t g_cbt, int32_t g_cb_off, int32_t n_img,
                       int32_t M, int32_t cb, int64_t voxels, float* partial, int32_t n_chunks, float* out, void* stream);
/* The 27 shifted copies of ONE channel (channel block src_cb, lane src_lane of a blocked 16-bit tensor) as a 32-channel
 * blocked tensor dst [n_img*4][Z][Y][X][8]: channel t = tap (dz*3+dy)*3+dx, zero outside the volume, channels 27..31 zero.
 * The weight gradient of a 1-input-channel Conv3d(k3,p1) (the per-modality first layers,
 * src/models/backbones/dual_encoder.py:62-71 -> unet.py:33-39) is then mmseg_conv3d_wgrad with ksize 1 on dst. */
int mmseg_im2col_k3_c1(const void* src, int32_t n_img, int32_t src_cbt, int32_t src_cb, int32_t src_lane, int32_t Z,
                       int32_t Y, int32_t X, void* dst, void* stream);
/* Gradient of a ConvTranspose3d(k2,s2) output (blocked, high resolution, channel blocks [src_cb_off, +cb)) -> its k=1
 * GEMM view at low resolution with channel = tap*C + co: the dY operand of the transposed conv's dgrad / wgrad. */
int mmseg_unshuffle_k2s2(const void* src, int32_t n_img, int32_t src_cbt, int32_t src_cb_off, int32_t cb, int32_t Z,
                         int32_t Y, int32_t X, void* dst, void* stream);

/* NCDHW fp32 [n_img][C][Z][Y][X] -> blocked bf16 (hi[, lo]) with channels zero-padded to cb*8.  Module boundary.  * dst_lo_off < 0 ("packed split"): instead of separate hi / lo planes the destination receives the VIRTUAL channels
 * [hi(C) | lo(C) | hi(C)] (3C <= 8*cb), so that A_hi*W_hi + A_lo*W_hi + A_hi*W_lo of a thin first layer is ONE K chunk
 * against the weights [W_hi | W_hi | W_lo]; mmseg_swi_gather takes the same convention.
 */
int mmseg_pack_ncdhw(const float* src, void* dst, int32_t n_img, int32_t C, int32_t Z, int32_t Y, int32_t X,
                     int32_t dst_cbt, int32_t dst_cb_off, int32_t dst_lo_off, int32_t cb, int32_t fmt, void* stream);
/* blocked bf16 (hi[, lo]) -> NCDHW fp32 (feature taps for return_features / hooks). */
/* mmseg_pack_ncdhw with SUVGuidedAttention's element-wise steps folded in (fusion/attention_fusion.py:283-292):
 * pre_sigmoid: x <- sigmoid((x - pre_sub) * pre_mul) (soft SUV mask); gate_logits [n_img][voxels] fp32 or NULL:
 * x <- x * (1 + sigmoid(gate)) (CT features modulated by the spatial attention). */
int mmseg_pack_ncdhw_ex(const float* src, void* dst, int32_t n_img, int32_t C, int32_t Z, int32_t Y, int32_t X,
                        int32_t dst_cbt, int32_t dst_cb_off, int32_t dst_lo_off, int32_t cb, int32_t pre_sigmoid,
                        float pre_sub, float pre_mul, const float* gate_logits, int32_t fmt, void* stream);
int mmseg_unpack_ncdhw(const void* src, float* dst, int32_t n_img, int32_t C, int32_t Z, int32_t Y, int32_t X,
                       int32_t src_cbt, int32_t src_cb_off, int32_t src_lo_off, int32_t fmt, void* stream);

/*
 * Sliding-window inference pieces (monai.inferers.sliding_window_inference as called at
 * src/trainer/trainer.py:381-392; algorithm in SURVEY.md Appendix C).
 *  gather:  windows of the NCDHW fp32 volume -> blocked bf16 batch (one image per window).
 *  blend:   out[:, window] += w * logits[window]; count[window] += w — windows applied in index order per voxel,
 *           one owner thread per voxel (deterministic, same order as the reference loop; no atomics).
 *           Box mode (bz0 >= 0): the launch covers the box [bz0,bz1)x[by0,by1)x[bx0,bx1) and loops over n_win windows.
 *           Window mode (bz0 < 0, n_win == 1): the launch covers the window whose origin is read from starts_dev —
 *           all launch arguments are batch-independent, so the call can be captured in a CUDA graph.
 *  finalize: out / count (in place, optional) and argmax over channels -> uint8 labels (trainer.py:364-367).
 */
int mmseg_swi_gather(const float* volume, int32_t C, int32_t VZ, int32_t VY, int32_t VX, const int32_t* starts_dev,
                     int32_t n_win, int32_t RZ, int32_t RY, int32_t RX, void* dst, int32_t dst_cbt, int32_t dst_lo_off,
                     int32_t cb, int32_t fmt, void* stream);
/* The same windows as a plain fp32 batch dst [n_win][C][RZ][RY][RX] (zero outside the volume): the image input of
 * SwinUNETR's patch embedding under monai.inferers.sliding_window_inference (trainer.py:381-392). */
int mmseg_swi_gather_ncdhw(const float* volume, int32_t C, int32_t VZ, int32_t VY, int32_t VX, const int32_t* starts_dev,
                           int32_t n_win, int32_t RZ, int32_t RY, int32_t RX, float* dst, void* stream);
/* out_conv (1x1x1, C -> K <= 8 classes, unet.py:163,199) fused into the blend of ONE window: the logits of window
 * `window` of the blocked feature batch are computed in registers and blended into out / count (same arithmetic and order
 * as mmseg_conv1x1_logits + mmseg_swi_blend in window mode, bit-identical accumulators) — the logits tensor is never
 * materialised.  RX, VX and the window's x origin must be multiples of 4. */
int mmseg_swi_logits_blend(const void* feat, int32_t src_cbt, int32_t cb_off, int32_t lo_off, int32_t cin, int32_t window,
                           const float* weight /* [K][cin] fp32 */, const float* bias, int32_t K,
                           const int32_t* starts_dev /* this window's origin */, int32_t RZ, int32_t RY, int32_t RX,
                           const float* wz, const float* wy, const float* wx, float w_floor, float* out, float* count,
                           int32_t VZ, int32_t VY, int32_t VX, int32_t fmt, void* stream);
int mmseg_swi_blend(const float* win_logits /* [n_win][K][RZ][RY][RX] */, const int32_t* starts_dev, int32_t n_win,
                    int32_t K, int32_t RZ, int32_t RY, int32_t RX, const float* wz, const float* wy, const float* wx,
                    float w_floor, float* out /* [K][VZ][VY][VX] */, float* count /* [VZ][VY][VX] */, int32_t VZ,
                    int32_t VY, int32_t VX, int32_t bz0, int32_t bz1, int32_t by0, int32_t by1, int32_t bx0,
                    int32_t bx1, void* stream);
/* finalize: `out` / `count` / `labels` address the first voxel of the range; class plane c of `out` starts
 * c * plane_stride elements further, so an axis-0 slab of the [K][VZ][VY][VX] accumulator is finalized in place
 * (whole volume: plane_stride == voxels). */
int mmseg_swi_finalize(float* out, const float* count, int32_t K, int64_t voxels, int64_t plane_stride,
                       int32_t normalize_in_place, uint8_t* labels /* or NULL */, void* stream);
/* acc[p][i] += part[p][i], p < planes, i < n (plane strides in elements): the one add of the sharded (multi-GPU) path —
 * the owner of an axis-0 slab adds, in rank order, the partial sums (K weighted-logit planes + the count map) another
 * rank accumulated over that slab.  No reference counterpart: the reference has no distributed path. */
int mmseg_swi_add_partial(float* acc, int64_t acc_plane_stride, const float* part, int64_t part_plane_stride,
                          int32_t planes, int64_t n, void* stream);

/*
 * DiceCE forward in one pass over the logits (src/trainer/losses.py:216-228, DiceLoss :39-80, CrossEntropyLoss).
 * result[0..2] = total, dice part, ce part.  C in {2,3,4,8,16}.
 */
int mmseg_dicece_fwd(const float* logits /* [B][C][N] */, const int64_t* target /* [B][N] */, int32_t B, int32_t C,
                     int64_t N, float dice_weight, float ce_weight, float smooth, int32_t include_background,
                     const float* class_weights /* [C] or NULL */, float* partial /* [B][n_blocks][3*C+2] */,
                     int32_t n_blocks, float* result /* [3] */, float* sums /* [B][3*C+2] or NULL */, void* stream);
/* d(DiceCE)/d(logits) * grad_out[0]; `sums` is the forward's per-batch (I, P, T, nll, weight) reduction. */
int mmseg_dicece_bwd(const float* logits, const int64_t* target, int32_t B, int32_t C, int64_t N, float dice_weight,
                     float ce_weight, float smooth, int32_t include_background, const float* class_weights,
                     const float* sums, const float* grad_out /* [1] or NULL */, float* dlogits, void* stream);

/*
 * TverskyLoss (src/trainer/losses.py:156-185, reduction mean) from the same one-pass per-class sums, and FocalLoss
 * (losses.py:106-125, reduction mean, optional class weights alpha) as a one-pass kernel.  mmseg_focal runs the
 * forward when dlogits is NULL (result[0] = loss) and the backward otherwise.
 */
int mmseg_tversky_fwd(const float* logits, const int64_t* target, int32_t B, int32_t C, int64_t N, float alpha, float beta,
                      float smooth, float* partial /* [B][n_blocks][3*C+2] */, int32_t n_blocks, float* result /* [1] */,
                      float* sums /* [B][3*C+2] */, void* stream);
int mmseg_tversky_bwd(const float* logits, const int64_t* target, int32_t B, int32_t C, int64_t N, float alpha, float beta,
                      float smooth, const float* sums, const float* grad_out, float* dlogits, void* stream);
int mmseg_focal(const float* logits, const int64_t* target, int32_t B, int32_t C, int64_t N, const float* class_weights,
                float gamma, float* partial /* [n_blocks] */, int32_t n_blocks, float* result, const float* grad_out,
                float* dlogits, void* stream);

/*
 * Fused multi-head cross attention over voxel tokens — the einsum / softmax / einsum of CrossAttentionFusion.forward
 * (src/models/fusion/attention_fusion.py:144-155) as one flash-style tcgen05 kernel; the q/k/v/out 1x1 projections
 * (:138-140,159) run through mmseg_conv3d_fwd and the residual + InstanceNorm3d (:162) through mmseg_add_stats +
 * mmseg_instnorm_finalize + mmseg_instnorm_act_apply (slope 1 = no activation).
 * q / kv / out: blocked token tensors [n_img * cbt][n_tok][8] bf16; head h owns channel blocks [cb0 + h*head_dim/8, ...).
 * head_dim in {16, 32, 64, 128} (8 is run zero-padded to 16); scale = real_head_dim^-0.5.
 */
int mmseg_cross_attention_fwd(const void* q, int32_t q_cbt, int32_t q_cb0, const void* kv, int32_t kv_cbt,
                              int32_t k_cb0, int32_t v_cb0, void* out, int32_t o_cbt, int32_t o_cb0, int32_t n_img,
                              int32_t heads, int32_t head_dim, int64_t n_tok, float scale,
                              float* lse /* optional [n_img][heads][n_tok]: saved for the backward */, void* stream);
/*
 * Backward of the fused attention (autograd through attention_fusion.py:144-155 in the reference), flash style by
 * recomputation, deterministic: dsum = rowsum(dO o O) (workspace [n_img][heads][n_tok]), then one kernel per 128-key
 * tile writes dK / dV and one per 128-query tile writes dQ (all blocked bf16 token tensors, same head layout as the
 * forward; dK and dV live in one tensor `dkv` like K and V do).  lse: the forward's log-sum-exp output.
 */
int mmseg_cross_attention_bwd(const void* q, int32_t q_cbt, int32_t q_cb0, const void* kv, int32_t kv_cbt, int32_t k_cb0,
                              int32_t v_cb0, const void* out, int32_t o_cbt, int32_t o_cb0, const void* d_out,
                              int32_t do_cbt, int32_t do_cb0, const float* lse, float* dsum, void* dq, int32_t dq_cbt,
                              int32_t dq_cb0, void* dkv, int32_t dkv_cbt, int32_t dk_cb0, int32_t dv_cb0, int32_t n_img,
                              int32_t heads, int32_t head_dim, int64_t n_tok, float scale, void* stream);
/* y (fp32 blocked [n_img*cb][voxels][8]) = a + b (blocked bf16) and per-chunk (sum, sum of squares) partials
 * [n_img][n_chunks][cb*8][2] for mmseg_instnorm_finalize. */
int mmseg_add_stats(const void* a, int32_t a_cbt, int32_t a_cb0, const void* b, int32_t b_cbt, int32_t b_cb0,
                    int32_t n_img, int32_t cb, int64_t voxels, float* y, float* partial, int32_t n_chunks, void* stream);

/*
 * K x K confusion counts (rows = target, columns = prediction) of two label maps in one pass; `counts` accumulates
 * (zero it first).  DiceMetric.update / ConfusionMatrix.update (src/trainer/metrics.py:42-65,184-196) read their
 * per-class intersections and unions off this matrix.  pred is int64 or uint8.
 */
int mmseg_confusion_hist(const void* pred, int32_t pred_is_u8, const int64_t* target, int64_t N, int32_t K,
                         uint64_t* counts /* [K*K] */, void* stream);

/*
 * DualEncoder modality fusion (src/models/backbones/dual_encoder.py:167-199, CrossModalAttention :207-254; same maths
 * as AttentionFusion, src/models/fusion/attention_fusion.py:48-74).  The M encoders write their level outputs into one
 * blocked buffer, modality-major (channel m*C + c), so torch.stack / torch.cat cost nothing.
 *  channel_mean:     AdaptiveAvgPool3d(1) per (image, channel), deterministic two-stage reduction.
 *  gate_mlp:         Linear(MC, MC/4) -> ReLU -> Linear(MC/4, M) -> Softmax over modalities.
 *  modality_combine: dst[b,c] = sum_m w[b,m] * src[b, m*C + c]  (w = gate, or uniform 1/M for mean, 1 for add).
 *  maxpool3d_2:      nn.MaxPool3d(2) (unet.py:73,77) on a blocked tensor (stand-alone DownBlock3D).
 */
int mmseg_channel_mean(const void* src, int32_t n_img, int32_t src_cbt, int32_t cb_off, int32_t lo_off, int32_t cb,
                       int64_t voxels, float* partial /* [n_img*cb][n_chunks][8] */, int32_t n_chunks,
                       float* mean /* [n_img][cb*8] */, int32_t fmt, void* stream);
int mmseg_gate_mlp(const float* pooled, const float* w1, const float* b1, const float* w2, const float* b2,
                   int32_t n_img, int32_t MC, int32_t H, int32_t M, float* weights /* [n_img][M] */, void* stream);
/* Backward of the gate MLP (autograd through the Linear - ReLU - Linear - Softmax of CrossModalAttention.attention,
 * dual_encoder.py:226-233): dweights [n_img][M] -> dpooled [n_img][MC], dw1 [H][MC], db1 [H], dw2 [M][H], db2 [M] (summed over the
 * images).  workspace: (2 H + M) * n_img floats. */
int mmseg_gate_mlp_bwd(const float* pooled, const float* w1, const float* b1, const float* w2, const float* b2,
                       const float* dweights, int32_t n_img, int32_t MC, int32_t H, int32_t M, float* workspace, float* dpooled,
                       float* dw1, float* db1, float* dw2, float* db2, void* stream);
int mmseg_modality_combine(const void* src, int32_t n_img, int32_t src_cbt, int32_t src_lo_off, int32_t M, int32_t cb,
                           int64_t voxels, const float* weights /* [n_img][M] or NULL */, float uniform_weight,
                           void* dst, int32_t dst_cbt, int32_t dst_cb_off, int32_t dst_lo_off, int32_t fmt, void* stream);
/* F.interpolate(x, size=(Zo,Yo,Xo), mode="trilinear", align_corners=True) on fp32 [n_planes = N*C][Zi][Yi][Xi]:
 * DeepSupervisionHead's resize of coarse-scale logits (src/models/heads/segmentation.py:108-113). */
int mmseg_trilinear_resize(const float* src, int32_t n_planes, int32_t Zi, int32_t Yi, int32_t Xi, float* dst, int32_t Zo,
                           int32_t Yo, int32_t Xo, void* stream);
/* out_conv / SegmentationHead k=1: nn.Conv3d(C, num_classes, 1) (unet.py:163,199; dual_encoder.py:118,164) on the CUDA
 * cores at HBM speed: src blocked bf16 (channels cb_off*8 .. +cin, optional lo plane lo_off blocks away, parity mode),
 * weight fp32 [cout][cin], bias fp32 [cout] or NULL, dst fp32 NCDHW [n_img][cout][voxels].  cin % 8 == 0, cout <= 16. */
int mmseg_conv1x1_logits(const void* src, int32_t n_img, int32_t src_cbt, int32_t cb_off, int32_t lo_off, int32_t cin,
                         int64_t voxels, const float* weight, const float* bias, int32_t cout, float* dst, int32_t fmt,
                         void* stream);
/* dst[b, c] = max over modalities (LateFusion fusion_method="max", src/models/fusion/late_fusion.py:62-64); bf16 mode. */
int mmseg_modality_max(const void* src, int32_t n_img, int32_t src_cbt, int32_t M, int32_t cb, int64_t voxels, void* dst,
                       int32_t dst_cbt, int32_t dst_cb_off, void* stream);
int mmseg_maxpool3d_2(const void* src, int32_t n_img, int32_t src_cbt, int32_t src_cb_off, int32_t src_lo_off,
                      int32_t cb, int32_t Z, int32_t Y, int32_t X, void* dst, int32_t dst_cbt, int32_t dst_cb_off,
                      int32_t dst_lo_off, int32_t fmt, void* stream);

/*
 * Inference-edge preprocessing on the device: ModalitySpecificNormalize of the reference's data pipeline
 * (src/data/transforms.py:362-404) on a [C][voxels] fp32 volume that is already in HBM.
 *  channel_stats: per channel (max, mean, population std) -> stats[C][3]; deterministic two-stage reduction;
 *                 partial = workspace of C * n_blocks * 3 doubles.
 *  modality_normalize: kind[c] 0 copy | 1 CT window clip(v, a, b) -> (v - a) / (b - a) (transforms.py:380-387) |
 *                 2 PET v / max when max > 0 (:389-394) | 3 z-score (v - mean) / (std + 1e-8) (MRI / US, :396-401).
 * Resize(order=1) (transforms.py:215-250, scipy.ndimage.zoom) is mmseg_trilinear_resize (aligned corners).
 */
int mmseg_channel_stats(const float* vol, int32_t C, int64_t voxels, void* partial, int32_t n_blocks, float* stats,
                        void* stream);
int mmseg_modality_normalize(const float* src, float* dst, int32_t C, int64_t voxels, const int32_t* kind, const float* a,
                             const float* b, const float* stats, void* stream);

/*
 * Trainer step glue — the kernels either side of forward / backward.
 *
 * mmseg_weights_repack: fp32 PyTorch-layout parameter (nn.Conv3d [Cout][Cin][k][k][k], nn.ConvTranspose3d
 * [Cin][Cout][2][2][2]) -> the packed 16-bit GEMM operand of mmseg_conv3d_fwd, in one launch: the forward form, the
 * spatially flipped + channel-transposed form the dgrad launch needs (autograd's convolution_backward input gradient,
 * reached from src/trainer/trainer.py:243), the ConvTranspose GEMM forms and the hi / lo splits of the split numeric
 * modes.  It replaces the per-step flip / permute / cat / copy chains.  The caller supplies two int32 device tables:
 *   n_off[n_out]   element offset of GEMM column n in the source, -1 = zero-padding column
 *   k_off[n_kc*16] element offset of GEMM K index (16 per chunk) in the source, -1 = zero padding
 * element = w[n_off[n] + k_off[k] + tap] * scale (tap mirrored when flip).  dst layout as consumed by conv_tc.cu:
 * [n_out/NT][n_kc_total][9 | 1][2][3*NT | NT][8], n_kc_total = n_kc * (hi_copies + has_lo).
 */
int mmseg_weights_repack(const float* w, const int32_t* n_off, const int32_t* k_off, void* dst, int32_t n_out, int32_t NT,
                         int32_t n_kc, int32_t n_kc_total, int32_t ksize, int32_t flip, int32_t hi_copies, int32_t has_lo,
                         int32_t fmt, float scale, void* stream);
/* The same for a whole list of weights in ONE launch: descs = device array of mmseg_repack_desc (the per-weight arguments of
 * mmseg_weights_repack; first_block = index of the weight's first 256-thread block), block_desc[b] = descriptor of block b. */
typedef struct {
  const float* w;
  const int32_t* n_off;
  const int32_t* k_off;
  void* dst;
  int32_t n_out, NT, n_kc, n_kc_total, ksize, flip, hi_copies, has_lo;
  int64_t first_block;
} mmseg_repack_desc;
int mmseg_weights_repack_multi(const void* descs, int32_t n_descs, const int32_t* block_desc, int64_t n_blocks, int32_t fmt,
                               float scale, void* stream);
/* dst[i] = idx[i] >= 0 ? src[idx[i]] : 0 — conv bias padded / expanded to the GEMM columns (ConvTranspose: x8 taps). */
int mmseg_gather_f32(const float* src, const int32_t* idx, float* dst, int32_t n, void* stream);

/*
 * torch.optim.AdamW.step() (reference src/trainer/trainer.py:115-117, stepped at :245-248) over a whole list of fp32
 * tensors in one launch: decoupled weight decay, bias correction, optional zeroing of the gradients (optimizer.zero_grad).
 * tensors: device array of n_tensors mmseg_adamw_tensor (each with its own device fp32 step counter, incremented by the
 * call); chunks: device array of (tensor index, chunk index) pairs, one per CTA, 8192 elements per chunk; hyper: device fp32[6] =
 * {lr, beta1, beta2, eps, weight_decay, grad_scale} (grad_scale multiplies every gradient, e.g. 1/world after a summing
 * all-reduce).  Graph-capturable: no host-side state.
 */
typedef struct {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* step;
  int64_t n;
} mmseg_adamw_tensor;
#define MMSEG_ADAMW_CHUNK 8192
int mmseg_adamw_multi(const void* tensors, int32_t n_tensors, const int32_t* chunks, int32_t n_chunks, const float* hyper,
                      int32_t zero_grad, void* stream);

/*
 * SwinUNETR (BASELINE.json configs[3]).  The reference's SwinUNETR (src/models/backbones/swin_unetr.py:20-200) builds
 * monai.networks.nets.SwinUNETR in its ctor (:80-96) and forward is self.model(x) (:117); these entry points replace
 * the non-convolution ATen ops MONAI issues under that call.  Linear layers (qkv / proj / mlp / patch-merging
 * reduction) are 1x1x1 GEMMs through mmseg_conv3d_fwd; tokens stay in the blocked layout, residual stream fp32.
 */
/* swinViT.patch_embed: nn.Conv3d(Cin, F, kernel 2, stride 2) + bias.  x NCDHW fp32 [n][Cin][2Z][2Y][2X] ->
 * xs blocked fp32 [n * F/8][Z][Y][X][8]. */
int mmseg_swin_patch_embed(const float* x, const float* w /* [F][Cin][2][2][2] */, const float* b /* [F] or NULL */, float* xs,
                           int32_t n_img, int32_t Cin, int32_t F, int32_t Z, int32_t Y, int32_t X, void* stream);
/* Residual add + token LayerNorm over channels (SwinTransformerBlock: x = shortcut + attn(...), norm2(x); x = x + mlp,
 * next block's norm1(x); SwinTransformer.proj_out: F.layer_norm without affine):
 *   xs += add (when add != NULL);  dst = LN(xs) * gamma + beta (when dst != NULL; gamma / beta NULL = no affine).
 * xs / add blocked fp32 [n * cb][voxels][8]; dst blocked 16-bit (elem_fmt) at channel block dst_cb_off of dst_cbt. */
int mmseg_swin_layernorm(float* xs, const float* add, const float* gamma, const float* beta, void* dst, int32_t n_img,
                         int32_t cb, int64_t voxels, int32_t dst_cbt, int32_t dst_cb_off, float eps, int32_t elem_fmt,
                         void* stream);
/* PatchMerging (MONAI "merging", 3-D legacy gather order) + LayerNorm(8C) with affine: xs blocked fp32 [n*cb][Z][Y][X][8] ->
 * dst blocked 16-bit [n * 8cb][Z/2][Y/2][X/2][8] (channel = octant * C + c); the Linear(8C -> 2C) follows as a GEMM. */
int mmseg_swin_merge_ln(const float* xs, const float* gamma, const float* beta, void* dst, int32_t n_img, int32_t cb, int32_t Z,
                        int32_t Y, int32_t X, float eps, int32_t elem_fmt, void* stream);
/* WindowAttention inside SwinTransformerBlock.forward_part1: zero padding to a multiple of the window, cyclic shift,
 * window partition, softmax(q k^T * scale + relative_position_bias [+ shift mask]) v, window reverse, un-shift, crop —
 * all as addressing inside one kernel (mma.sync m16n8k16 tensor-core tiles, head_dim 16, online softmax). */
typedef struct {
  const void* qkv;        /* blocked 16-bit [n_img * qkv_cbt][D][H][W][8]: channels [q | k | v], each heads x 16          */
  void* out;              /* blocked 16-bit [n_img * out_cbt][D][H][W][8]: attention output before proj                   */
  const float* table;     /* relative_position_bias_table [(2w0-1)(2w1-1)(2w2-1)][heads] fp32                             */
  const float* qkv_bias;  /* [3 * heads * 16] fp32 or NULL: q / k / v of a zero-padded token                              */
  int32_t n_img, D, H, W;
  int32_t window[3];      /* configured window (7, 7, 7); an axis not longer than it shrinks the window, cancels the shift */
  int32_t shift[3];       /* cyclic shift of this block (0 or window / 2)                                                 */
  int32_t heads, head_dim;
  int32_t qkv_cbt, out_cbt, out_cb_off;
  float scale;            /* head_dim^-0.5                                                                                 */
  int32_t elem_fmt;
  float* lse;             /* optional [n_img][heads][windows][352] fp32: log2-domain log-sum-exp rows for the backward      */
} mmseg_swin_attn_args;
int mmseg_swin_window_attention(const mmseg_swin_attn_args* args, void* stream);
/* UnetResBlock tail (MONAI dynunet_block.UnetResBlock, the block of every SwinUNETR encoder / decoder stage):
 *   y = LeakyReLU(slope)( IN(a) + r' ),  r' = IN(r) with r_mean_rstd (1x1x1 residual conv) or r itself (NULL).
 * a: raw conv output blocked [n*cb] (fp32 or 16-bit); r: blocked at channel block r_cb_off of r_cbt. */
int mmseg_instnorm_residual_act(const void* a, int32_t a_is_f32, const float* a_mean_rstd, const void* r, int32_t r_is_f32,
                                const float* r_mean_rstd, int32_t r_cbt, int32_t r_cb_off, void* dst, int32_t dst_cbt,
                                int32_t dst_cb_off, int32_t n_img, int32_t cb, int64_t voxels, float slope, int32_t elem_fmt,
                                void* stream);

/*
 * SwinUNETR training: the backward ops autograd issues for loss.backward() (reference src/trainer/trainer.py:243) through
 * monai.networks.nets.SwinUNETR (src/models/backbones/swin_unetr.py:80-117).  bf16 operands, fp32 stream gradients.
 */
/* LayerNorm forward that also saves (mean, rstd) per token: xs_out = xs_in (+ add16, blocked bf16); ln16 = LN(xs_out). */
int mmseg_swin_ln_fwd_train(const float* xs_in, const void* add16, const float* gamma, const float* beta, float* xs_out,
                            void* ln16, float* stats /* [tokens][2] */, int32_t n_img, int32_t cb, int64_t voxels, float eps,
                            void* stream);
/* native_layer_norm_backward (input gradient): dxs_out = (dxs_in | 0) + LN'(dy16); dxs16 = bf16 copy of dxs_out. */
int mmseg_swin_ln_bwd(const float* xs, const float* stats, const void* dy16, const float* gamma, const float* dxs_in,
                      float* dxs_out, void* dxs16, int32_t n_img, int32_t cb, int64_t voxels, void* stream);
/* native_layer_norm_backward (weight / bias gradients): partial [n_chunks][C][2] = (sum dy*xhat, sum dy) per chunk. */
int mmseg_swin_ln_param_grad(const float* xs, const float* stats, const void* dy16, float* partial, int32_t n_chunks,
                             int32_t n_img, int32_t cb, int64_t voxels, void* stream);
/* gelu_backward (MLPBlock, exact erf form) and the LeakyReLU mask of the UnetResBlock tail (out = dy * (y > 0 ? 1 : slope)). */
int mmseg_gelu_bwd(const void* x16, const void* dy16, void* dx16, int64_t n_elems, void* stream);
int mmseg_lrelu_mask_mul(const void* y16, const void* dy16, void* out16, int64_t n_elems, float slope, void* stream);
/* PatchMerging gather (xs -> [8 slots x C] channels at half resolution, fp32 blocked) and its transpose (backward). */
int mmseg_swin_merge_gather(const float* xs, float* cat, int32_t n_img, int32_t cb, int32_t Z, int32_t Y, int32_t X, void* stream);
int mmseg_swin_merge_scatter(const float* dcat, float* dxs, int32_t n_img, int32_t cb, int32_t Z, int32_t Y, int32_t X,
                             void* stream);
/* convolution_backward (weight, bias) of the patch embedding: partial [n_chunks][F][8 Cin + 1]. */
int mmseg_swin_patch_embed_wgrad(const float* x, const float* dxs, float* partial, int32_t n_chunks, int32_t n_img, int32_t Cin,
                                 int32_t F, int32_t Z, int32_t Y, int32_t X, void* stream);
/* Backward of mmseg_swin_window_attention (args as in the forward, `out` = the forward output, `lse` = its saved rows):
 * dqkv (blocked bf16 like qkv), dtable partial [n_img][windows][heads][table rows], dbias partial [n_img][windows][heads][32]
 * (dk | dv of the zero-padded tokens, i.e. their share of the qkv-bias gradient). */
int mmseg_swin_window_attention_bwd(const mmseg_swin_attn_args* args, const void* dout, int32_t dout_cbt, int32_t dout_cb_off,
                                    const float* lse, void* dqkv, float* dtable, float* dbias, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMSEG_B200_H_ */
