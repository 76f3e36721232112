/*
 * hifigan_b200.h — C ABI of the B200-native HiFi-GAN generator (libhifigan_b200.so).
 *
 * The reference (diff7/tts-king) has no FFI: its boundary for this path is the Python
 * nn.Module protocol of hifi/models.py::Generator and the hifiapi.py::HIFIapi wrapper
 * (SURVEY.md §8b).  This header is the C surface that sits directly under that protocol;
 * tts_king_b200/hifi/models.py binds it with ctypes (see INTEGRATION.md for the stub a
 * reference maintainer would add).  Every entry point names the reference interface it
 * stands in for.
 *
 * Conventions: extern "C"; plain pointers and sizes only (no torch / C++ types); every
 * function returns 0 on success or a negative HG_E* code, with a thread-local message from
 * hg_last_error(); nothing throws; hg_forward performs no device allocation — the caller owns
 * the mel, the output and the workspace (PyTorch's caching allocator in the Python binding),
 * and all work is enqueued on the caller's stream.  A plan is immutable after
 * hg_plan_finalize; concurrent hg_forward calls on distinct streams + workspaces are safe.
 *
 * There is no CPU fallback: every compute entry point runs hand-written sm_100a kernels and
 * fails with HG_ENODEVICE / HG_ECUDA otherwise.
 */
#ifndef HIFIGAN_B200_H_
#define HIFIGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HG_ABI_VERSION 1

#if defined(__GNUC__)
#define HG_API __attribute__((visibility("default")))
#else
#define HG_API
#endif

#define HG_MAX_UPS 8
#define HG_MAX_KERNELS 8
#define HG_MAX_DIL 4

/* error codes */
#define HG_OK 0
#define HG_EINVAL (-1)    /* bad argument / shape mismatch (RuntimeError in the reference) */
#define HG_ENODEVICE (-2) /* no sm_100 device */
#define HG_ECUDA (-3)     /* CUDA runtime / driver error */
#define HG_ESTATE (-4)    /* call order: missing weights, plan not finalized, ... */
#define HG_ENOMEM (-5)    /* workspace too small / host allocation failed */

/* arithmetic mode of the contraction (north_star: an fp32 path within 1e-4 of the reference and
 * a bf16 path with SNR >= 40 dB) */
#define HG_PREC_BF16 0      /* bf16 operands, fp32 accumulate (TMEM), fp32 residual stream */
#define HG_PREC_FP32 1      /* fp32-accurate on tensor cores: bf16x3 split operands (hi/lo)    */
#define HG_PREC_FP32_FFMA 2 /* exact fp32 FFMA on CUDA cores (slow; on-device cross-check)      */

#define HG_OUT_F32 0 /* Generator.forward: float wav in [-1,1]           hifi/models.py:199 */
#define HG_OUT_I16 1 /* HIFIapi.generate: wav*scale, truncating int16     hifiapi.py:50-51  */

/*
 * The fields Generator.__init__ reads from `h` — hifi/models.py:150-181 (values for V1 in
 * config.yaml:16,25-29).  resblock_type: 1 = ResBlock1 (hifi/models.py:12-101, uses 3
 * dilations), 2 = ResBlock2 (:104-143, uses 2).
 */
typedef struct HgConfig {
  int32_t num_mels;                 /* 80, hard-coded at hifi/models.py:153 */
  int32_t upsample_initial_channel;
  int32_t num_upsamples;
  int32_t upsample_rates[HG_MAX_UPS];
  int32_t upsample_kernel_sizes[HG_MAX_UPS];
  int32_t num_kernels;
  int32_t resblock_kernel_sizes[HG_MAX_KERNELS];
  int32_t resblock_dilation_sizes[HG_MAX_KERNELS][HG_MAX_DIL];
  int32_t resblock_type;
} HgConfig;

typedef struct HgPlan HgPlan;

HG_API int hg_abi_version(void);

/* Message for the last failing call on this thread ("" if none). */
HG_API const char* hg_last_error(void);

/* Number of visible sm_100 devices (0 if none / no driver).  Never fails. */
HG_API int hg_device_count(void);

/* Generator.__init__(h) — hifi/models.py:147-183: fixes the layer list for `cfg` on `device`. */
HG_API int hg_plan_create(const HgConfig* cfg, int device, HgPlan** plan);

/*
 * load_state_dict + remove_weight_norm — hifiapi.py:20-28, hifi/models.py:203-210.
 * `name` is the state_dict prefix ("conv_pre", "ups.1", "resblocks.4.convs2.0", "conv_post");
 * `weight` is the FOLDED fp32 tensor in the reference's own layout on the HOST
 * (Conv1d [C_out,C_in,k], ConvTranspose1d [C_in,C_out,k]); `bias` [C_out].  The library repacks
 * (tap-major, K-major, bf16 hi/lo, 128B-swizzled tiles) and uploads.
 */
HG_API int hg_plan_upload_weight(HgPlan* plan, const char* name, const float* weight,
                          const int64_t* shape, int ndim, const float* bias, int64_t bias_len);

/* Checks every layer has weights; after this the plan is immutable. */
HG_API int hg_plan_finalize(HgPlan* plan);

/* Bytes of device scratch hg_forward needs for a [B,80,T] input in `precision`. */
HG_API int hg_workspace_bytes(const HgPlan* plan, int B, int T, int precision, size_t* bytes);

/* Kernel launches one hg_forward(B,T,precision) enqueues (bench.py's gpu_launches). */
HG_API int hg_forward_launches(const HgPlan* plan, int B, int T, int precision, int* launches);

/*
 * Generator.forward(x) — hifi/models.py:185-201 (and, with HG_OUT_I16, the tail of
 * HIFIapi.generate, hifiapi.py:50-51).
 *   mel      device fp32, logical shape [B, num_mels, T] with ELEMENT strides (sB, sC, sT) — the
 *            caller's tensor may be the non-contiguous transpose of a time-major [B,T,80]
 *            (tts_king.py:48); it is read in place, never modified.
 *   out      device, [B, 1, T*prod(rates)] contiguous; float (HG_OUT_F32) or int16 (HG_OUT_I16,
 *            value = (int16)trunc(wav * out_scale) with numpy's wrap-around cast).
 *   workspace  device scratch of at least hg_workspace_bytes(); 1024-byte aligned.
 *   stream   cudaStream_t (as void*); all kernels are enqueued there, nothing synchronizes.
 */
HG_API int hg_forward(HgPlan* plan, const float* mel, int64_t sB, int64_t sC, int64_t sT, int B, int T,
               void* out, int out_dtype, float out_scale, int precision, void* workspace,
               size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Conv stacks — the step before the vocoder: FastSpeech2's PostNet
 * (fs_two/transformer/Layers.py:71-143: 5 x [Conv1d k5 + BatchNorm1d (+ tanh)]) and mel_linear
 * (fs_two/model/fastspeech2.py:101-104, a k = 1 conv) on the same conv kernels.  A stack is a plain
 * chain of same-length Conv1d layers; BatchNorm in eval mode is an affine map per channel and is
 * folded into the conv's weight and bias by the caller before upload.
 * Build: hg_stack_create -> hg_plan_upload_weight(plan, "<layer index>", ...) per layer ->
 * hg_plan_finalize; free with hg_plan_destroy.
 */
enum { HG_ACT_NONE = 0, HG_ACT_LRELU = 1, HG_ACT_TANH = 2 };

typedef struct HgStackLayer {
  int32_t c_in, c_out;
  int32_t k;         /* odd, padding = dilation * (k - 1) / 2 (ConvNorm's default, Layers.py:49-51) */
  int32_t dilation;
  int32_t act;       /* HG_ACT_*: applied to this layer's output on its way into the next layer */
  float slope;       /* HG_ACT_LRELU only */
} HgStackLayer;

HG_API int hg_stack_create(const HgStackLayer* layers, int n_layers, int device, HgPlan** plan);

HG_API int hg_stack_workspace_bytes(const HgPlan* plan, int B, int T, int precision, size_t* bytes);

/*
 * y = stack(x) (+ residual) — PostNet.forward (Layers.py:133-143) without its two transposes, and,
 * with residual = x, the `postnet(output) + output` of fastspeech2.py:104 in the last epilogue.
 *   x         device fp32, logical shape [B, c_in, T] with ELEMENT strides (sB, sC, sT); the
 *             time-major [B,T,c_in] tensor FastSpeech2 produces is sB = T*c_in, sC = 1, sT = c_in.
 *   residual  device fp32 [B][T][c_out] contiguous, or NULL.
 *   out       device fp32 [B][T][c_out] contiguous (time-major: what tts_king.py:48 transposes and
 *             hg_forward reads in place through its strides).
 */
HG_API int hg_stack_forward(HgPlan* plan, const float* x, int64_t sB, int64_t sC, int64_t sT, int B, int T,
                     const float* residual, float* out, int precision, void* workspace,
                     size_t workspace_bytes, void* stream);

/*
 * hg_forward whose last layer writes only the samples [skip_samples, skip_samples + keep_samples) of
 * every item, at out[b * out_item_stride + (t - skip_samples)] — the "compute, then gather" of
 * time-chunked long-form synthesis as one step: a chunk computed with halo
 * frames drops its halo samples in conv_post's epilogue and stores its owned samples directly at
 * their place in the final waveform.  `out` may be a peer-mapped address of a buffer on ANOTHER GPU
 * of the node (CUDA IPC + hg_enable_peer_access): the stores then travel over NVLink / NVSwitch and
 * no separate gather collective runs.
 */
HG_API int hg_forward_window(HgPlan* plan, const float* mel, int64_t sB, int64_t sC, int64_t sT, int B, int T,
                      void* out, int64_t out_item_stride, int64_t skip_samples, int64_t keep_samples,
                      int out_dtype, float out_scale, int precision, void* workspace,
                      size_t workspace_bytes, void* stream);

/*
 * CUDA IPC for the direct-store gather: the owner of the final waveform buffer exports it
 * (hg_ipc_export: 64-byte handle of the underlying allocation + the byte offset of `device_ptr` in it),
 * the handle travels to the other ranks' processes by any means (the Python layer uses the process
 * group), and each of them maps it FOR ITS OWN GPU (hg_ipc_import(device, ...) -> `ptr`, usable as
 * `out` of hg_forward_window on `device`; `base` is what hg_ipc_close takes).  Single node.
 */
HG_API int hg_ipc_export(const void* device_ptr, unsigned char* handle64, int64_t* offset);
HG_API int hg_ipc_import(int device, const unsigned char* handle64, int64_t offset, void** base, void** ptr);
HG_API int hg_ipc_close(int device, void* base);

/* cudaDeviceEnablePeerAccess(device -> peer_device), idempotent; HG_ENODEVICE when there is no peer path. */
HG_API int hg_enable_peer_access(int device, int peer_device);

/*
 * Mel frames of right context a sample needs (13 for V1): the generator's
 * receptive reach, rounded up to frames.  What hg_forward_ragged adds to every item's length.
 */
HG_API int hg_halo_frames(const HgPlan* plan, int* frames);

/*
 * hg_forward for a padded batch whose items have their own lengths — the vocoder call of
 * synth_samples, fs_two/utils/tools.py:257-268 + fs_two/utils/model.py:85-100, where the reference
 * runs the generator over the padding and trims the waveforms to `lengths` afterwards.
 *   frames   HOST array [B]: mel frames item b keeps, 1 <= frames[b] <= T.  B <= 64.
 * Every layer computes item b's rows only up to (frames[b] + hg_halo_frames) frames (the last
 * one up to frames[b]); `out` is written for b on [0, frames[b] * hop), rounded up to the last
 * kernel's 256-sample tile, and left untouched beyond.  Inside [0, frames[b] * hop) the result is
 * bit-identical to hg_forward on the same padded input, whatever the padding and the scratch hold;
 * samples of that last tile past frames[b] * hop are unspecified (none when hop is a multiple of 256).
 */
HG_API int hg_forward_ragged(HgPlan* plan, const float* mel, int64_t sB, int64_t sC, int64_t sT, int B, int T,
                      const int32_t* frames, void* out, int out_dtype, float out_scale, int precision,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Frees device weights and host state. */
HG_API int hg_plan_destroy(HgPlan* plan);

/*
 * Single-layer entry points for op-level parity tests (SURVEY.md §4 tier T1).  They run the same
 * kernels hg_forward uses on one layer.  Layout is the library's native channels-last:
 *   x  device fp32 [B][L][C_in]      y  device fp32 [B][L_out][C_out]
 * `weight`/`bias` are HOST fp32 in the reference layout.  in_slope: leaky_relu slope applied to
 * x before the convolution (1.0 = none).  residual: optional device fp32 [B][L_out][C_out] added
 * to the result.  Allocates temporaries and synchronizes — not a hot-path call.
 *   hg_op_conv1d            torch Conv1d(C_in,C_out,k,1,dilation=d,padding=get_padding(k,d))
 *                           as built at hifi/models.py:19-81,152-154
 *   hg_op_conv_transpose1d  torch ConvTranspose1d(C_in,C_out,k,s,padding=(k-s)//2),
 *                           hifi/models.py:161-171
 *   hg_op_conv_post         leaky_relu(0.01) -> Conv1d(C,1,7,padding=3) -> tanh,
 *                           hifi/models.py:197-199; y device fp32 [B][L]
 */
HG_API int hg_op_conv1d(int device, int precision, const float* x, int B, int L, int C_in,
                 const float* weight, const float* bias, int C_out, int k, int dilation,
                 float in_slope, const float* residual, float* y, void* stream);
HG_API int hg_op_conv_transpose1d(int device, int precision, const float* x, int B, int L, int C_in,
                           const float* weight, const float* bias, int C_out, int k, int stride,
                           float in_slope, float* y, void* stream);
HG_API int hg_op_conv_post(int device, const float* x, int B, int L, int C, const float* weight,
                    const float* bias, float* y, void* stream);
/*
 * One ResBlock1 pair, hifi/models.py:90-94, through the fused kernel (bf16 operands):
 *   y = c2(leaky_relu(c1(leaky_relu(x, in_slope)), in_slope)) + residual
 * c1: Conv1d(C,C,k,dilation=d1), c2: Conv1d(C,C,k,dilation=1); C in {32, 64}.  Fails with HG_EINVAL
 * when the shape is not covered by the fused kernel.
 */
HG_API int hg_op_conv_pair(int device, const float* x, int B, int L, int C, int k, int d1,
                           const float* w1, const float* b1, const float* w2, const float* b2,
                           float in_slope, const float* residual, float* y, void* stream);
/* A whole ResBlock1 (reference hifi/models.py:88-95): np pairs, weights w1[m] / w2[m] of shape [C, C, k], first-conv
 * dilations d1[m]; y = block(x).  Runs the fused-ResBlock kernel (csrc/conv_chain_tc.cu) when the shape is covered
 * (C in {32, 64}, k = 3, bf16 arithmetic) and otherwise the pair kernels; `fused` (nullable) reports which. */
HG_API int hg_op_resblock1(int device, const float* x, int B, int L, int C, int k, int np, const int32_t* d1,
                           const float* const* w1, const float* const* b1, const float* const* w2,
                           const float* const* b2, float slope, float* y, void* stream, int32_t* fused);

/*
 * Per-layer timing (bench.py's roofline pass; the reference has no profiler).
 * hg_layer_count / hg_layer_info describe the plan's layer table; hg_profile_forward runs one
 * hg_forward with CUDA events recorded on `stream` around every launch and, after synchronizing,
 * returns for launch i the plan layer it ran (layer_index[i], -1 = the mel repack) and its device
 * time in milliseconds.
 */
typedef struct HgLayerInfo {
  char name[64];     /* state_dict prefix */
  int32_t kind;      /* 0 Conv1d, 1 ConvTranspose1d, 2 conv_post(+tanh) */
  int32_t c_in, c_out, k, dilation, stride;
  int32_t tensor_core; /* 1: tcgen05 path available for this layer */
  int32_t n_tile, k_chunk, m_subtiles, stages, smem_bytes; /* tcgen05 tiling at `precision` */
  int32_t weights_resident, slab_buffers;
  int32_t kernel_path; /* HG_PATH_*: filled by hg_profile_launch_info (what the launch actually ran) */
} HgLayerInfo;
enum {
  HG_PATH_CUDA_CORE = 0,   /* conv_ffma.cu */
  HG_PATH_TC = 1,          /* conv_tc.cu: tcgen05, one CTA per tile */
  HG_PATH_TC_CTA_PAIR = 2, /* conv_tc2.cu: tcgen05 cta_group::2 */
  HG_PATH_FUSED_PAIR = 3,  /* conv_pair_tc.cu: c1 + c2 of a ResBlock pair in one launch (reported under c2's name) */
  HG_PATH_NARROW = 4,      /* conv_narrow.cu: 8 / 16 channels */
  HG_PATH_POST = 5,        /* tail.cu: conv_post + tanh (+ int16) */
  HG_PATH_REPACK = 6,      /* tail.cu: mel -> operand layout */
  HG_PATH_FUSED_BLOCK = 7  /* conv_chain_tc.cu: a whole ResBlock1 (all pairs) in one launch (reported under its last c2) */
};
HG_API int hg_layer_count(const HgPlan* plan, int* count);
/* Static description of plan layer `index`; the tiling fields are the single-CTA tcgen05 kernel's
 * default choice — the launch-time choice (CTA-pair kernel, fused pair, epilogue slots) is what
 * hg_profile_launch_info reports. */
HG_API int hg_layer_info(const HgPlan* plan, int index, int precision, HgLayerInfo* info);
/* Launch `launch` (0-based) of this thread's last hg_profile_forward: which kernel family ran and with
 * which tiling. */
HG_API int hg_profile_launch_info(const HgPlan* plan, int launch, HgLayerInfo* info);
HG_API int hg_profile_forward(HgPlan* plan, const float* mel, int64_t sB, int64_t sC, int64_t sT, int B,
                              int T, void* out, int out_dtype, float out_scale, int precision,
                              void* workspace, size_t workspace_bytes, void* stream,
                              int* layer_index, float* layer_ms, int max_launches, int* n_launches);

/*
 * Host-only diagnostic (no device needed): the tiling and MMA schedule the time-folded ResBlock-pair
 * kernel (csrc/conv_pair_fold.cu) uses for a pair of C -> C convs with k taps, first-conv dilation d1, on
 * sequences of L rows — the two convs of one iteration of ResBlock1.forward (reference hifi/models.py:88-95).
 * fusable = 0 when the shape is not covered (the N = C pair kernel or two launches run instead).
 * ops[i] = {A phase slab, first A row, first tap of B, taps stacked along N, first accumulator column,
 * weight block released}.  tests/test_host.py replays the whole dataflow from this description in numpy
 * against a direct convolution.
 */
typedef struct HgFoldInfo {
  int32_t fusable;
  int32_t f;                 /* time rows folded into N: 128 / C */
  int32_t r_out;             /* output rows per tile */
  int32_t delta;             /* tile t covers xt rows [t*r_out - delta, ...) */
  int32_t fdiv;              /* f * d1: rows per block group of the de-interleaved input */
  int32_t blk_off;           /* first block group of a tile's slab, relative to its origin (<= 0) */
  int32_t nb_slab;           /* block groups per slab */
  int32_t slab_phase_bytes, xt_phase_bytes;
  int32_t t_bufs, stages, weights_resident, smem_bytes;
  int32_t n_ops1, n_ops2;
  int32_t ops1[24][6];
  int32_t ops2[24][6];
} HgFoldInfo;
HG_API int hg_fold_info(int C, int k, int d1, int L, HgFoldInfo* info);

/*
 * Host-side view of the streamed-weight ring of the time-folded pair kernel (conv_pair_fold.cu::FoldRingPlan<k, period>),
 * evaluated by the same constexpr functions the kernel is compiled from: for tap `tap` of a conv whose position in the
 * ring order G1(0) G1(1) G2(0) G1(2) ... has parity `conv_parity`, the shared-memory slot, the mbarrier parity the
 * consumer waits for, and whether the tap is also copied into the mirror slot behind the ring.  *slots receives the
 * number of weight blocks held in shared memory.  Lets the CPU tests check that every slot's barrier is used with
 * strictly alternating parity and that every run of two consecutive taps is contiguous.  Returns HG_EINVAL for a
 * (k, period) the kernel is not instantiated for.
 */
HG_API int hg_fold_ring_query(int k, int period, int tap, int conv_parity, int32_t* slot, int32_t* parity, int32_t* mirror,
                              int32_t* slots);

/*
 * Debug / bring-up: runs the tcgen05 descriptor self-test (shifted-row UMMA descriptors against a
 * CUDA-core reference) and writes a report into `buf`.  Returns the number of failing cases.
 */
HG_API int hg_selftest_tcgen05(int device, char* buf, size_t buf_len);

#ifdef __cplusplus
}
#endif
#endif /* HIFIGAN_B200_H_ */
