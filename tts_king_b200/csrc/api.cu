// api.cu — C ABI of libhifigan_b200 (include/hifigan_b200.h): plan construction, weight
// repacking, the layer schedule of Generator.forward (reference hifi/models.py:185-201) and the
// op-level / self-test entry points.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "plan.h"

namespace hg {
// kernels.cu-side launchers
size_t conv_tc_smem_bytes(int n_t, int kc, bool split, int slab_rows, int nbuf, int stages, int epi_slot_bytes);
cudaError_t launch_conv_tc(int n_t, int kc, int ms, bool split, const CUtensorMap* maps, const TcConvParams& p,
                           int n_blocks, size_t smem, int grid_ctas, cudaStream_t st);
cudaError_t launch_conv_ffma(FfmaConvParams p, cudaStream_t st, const RaggedItems* items = nullptr);
cudaError_t launch_conv_post(const float* x, int B, int L, int C, const float* w_tapmajor, float bias, float* out_f32,
                             int16_t* out_i16, float out_scale, cudaStream_t st, const float* w_host_tapmajor,
                             const RaggedItems* items = nullptr, long long item_stride = 0, int skip = 0, int keep = 0);
cudaError_t launch_mel_to_operand(const float* mel, long long sB, long long sC, long long sT, int B, int C, int T,
                                  int c_pad, int a_fmt, void* a0, void* a1, cudaStream_t st, int act = 0, float slope = 1.f);
cudaError_t launch_f32_to_operand(const float* x, long long n, float slope, int a_fmt, void* a0, void* a1,
                                  cudaStream_t st);
int run_tcgen05_selftest(char* buf, size_t len);
cudaError_t launch_conv_narrow(int c, NarrowConvParams p, cudaStream_t st, const RaggedItems* items = nullptr);
cudaError_t launch_convt_narrow16(const void* a, const float* w_tap_cin_n, int B, int L_in, int stride, int pad, EpiParams epi,
                                  cudaStream_t st, const RaggedItems* items = nullptr);
size_t conv_pair_smem_bytes(int c, int slab_rows, int t_rows, int t_bufs, int stages);
size_t conv_tc2_smem_bytes(int n_t, int slab_rows, int nbuf, int stages, int epi_slot_bytes);
cudaError_t launch_conv_tc2(int n_t, int ms, const CUtensorMap* maps, const TcConvParams& p, size_t smem, int grid,
                            cudaStream_t st);
cudaError_t launch_conv_pair_tc(int c, const CUtensorMap& m, const CUtensorMap& mr, const TcPairParams& p, size_t smem,
                                int grid, cudaStream_t st);
size_t conv_chain_smem_bytes(int c, int buf_rows, int stages);
int conv_chain_guard_rows();
cudaError_t launch_conv_chain_tc(int c, const CUtensorMap& m, const CUtensorMap& mr, const TcChainParams& p, size_t smem,
                                 int grid, cudaStream_t st);
size_t conv_fold_smem_bytes(int c, int slab_phase_bytes, int xt_phase_bytes, int t_bufs, int stages);
bool conv_fold_has_kernel(int c, int k, int ring_period);
int conv_fold_weight_slots(int k, int ring_period);
bool conv_fold_ring_query(int k, int s, int tap, int cpar, int* slot, int* parity, int* mirror, int* slots);
cudaError_t launch_conv_pair_fold(int c, int k, int ring_period, const CUtensorMap& m, const CUtensorMap& mr, const TcFoldParams& p,
                                  size_t smem, int grid, cudaStream_t st);
}  // namespace hg

using namespace hg;

// ------------------------------------------------------------------------------------------------
// errors
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) return fail(HG_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e));   \
  } while (0)

// Every entry point that touches a device runs with that device current and puts the caller's device
// back on return: a process that drives several GPUs from one thread (or whose garbage collector destroys
// a plan at an arbitrary point) must not find its current device changed behind its back.
// see common.cuh::ragged_finish — consecutive launches of a forward walk their tiles in opposite directions
namespace hg {
thread_local int hg_reverse_tiles = -1;
}
struct TileOrderScope {
  explicit TileOrderScope(bool alternate) { hg_reverse_tiles = alternate ? 0 : -1; }
  ~TileOrderScope() { hg_reverse_tiles = -1; }
};

struct DeviceScope {
  int prev = -1;
  bool changed = false;
  cudaError_t err;
  explicit DeviceScope(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) {
      err = cudaSetDevice(dev);
      changed = err == cudaSuccess;
    }
  }
  ~DeviceScope() {
    if (changed) cudaSetDevice(prev);
  }
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope& operator=(const DeviceScope&) = delete;
};
#define DEVICE_SCOPE(dev)                                                                                  \
  DeviceScope _device_scope(dev);                                                                          \
  if (_device_scope.err != cudaSuccess)                                                                    \
  return fail(HG_ECUDA, "cudaSetDevice(%d): %s", static_cast<int>(dev), cudaGetErrorString(_device_scope.err))

extern "C" int hg_abi_version(void) { return HG_ABI_VERSION; }
extern "C" const char* hg_last_error(void) { return g_err.c_str(); }

extern "C" int hg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, d) == cudaSuccess && pr.major == 10) ++ok;
  }
  return ok;
}

static int check_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(HG_ENODEVICE, "no CUDA device visible; libhifigan_b200 has no CPU fallback");
  }
  if (device < 0 || device >= n) return fail(HG_EINVAL, "device %d out of range (0..%d)", device, n - 1);
  cudaDeviceProp pr;
  CUDA_TRY(cudaGetDeviceProperties(&pr, device));
  if (pr.major != 10)
    return fail(HG_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, pr.major, pr.minor);
  return HG_OK;
}

// ------------------------------------------------------------------------------------------------
// TMA descriptor encoding through the driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 operand plane [B][L][cpitch] -> 3-D map with box {kc, box_rows, 1}; rows outside [0,L) read
// as zero (that is the convolution's zero padding, and it keeps batch items from bleeding).
static int make_operand_map(HgPlan* plan, const void* ptr, int L, int B, int cpitch, int kc, int box_rows,
                            CUtensorMap* out) {
  MapKey key(ptr, L, B, cpitch, kc, box_rows);
  if (plan->maps.get(key, out)) return HG_OK;
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(HG_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cpitch), static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(cpitch) * 2, static_cast<cuuint64_t>(L) * cpitch * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(kc), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(HG_ECUDA, "cuTensorMapEncodeTiled failed (%d) for L=%d B=%d C=%d kc=%d box=%d", static_cast<int>(r), L,
                B, cpitch, kc, box_rows);
  plan->maps.put(key, m);
  *out = m;
  return HG_OK;
}

// Epilogue tile maps over a channels-last tensor [B][L][C]: box {cols, 32 rows, 1}.
//   kind 0: fp32, 32 columns, 128B swizzle  (fused-pair residual tiles)
//   kind 1: fp32, 16 columns,  64B swizzle  (conv_tc residual in / x out)
//   kind 2: bf16, 16 columns,  32B swizzle  (conv_tc operand copies out)
// rstride > 1: the phase-major 4-D view [B][L / rstride][rstride][c] of a polyphase ConvTranspose1d's output (L is a
// multiple of rstride); a box is 32 rows of ONE phase
static int make_tile_map(HgPlan* plan, const void* ptr, int L, int B, int c, int kind, CUtensorMap* out, int rstride = 1) {
  MapKey key(ptr, L, B, c, -(kind + 1), 32 * rstride);
  if (plan->maps.get(key, out)) return HG_OK;
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(HG_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const int esz = kind == 2 ? 2 : 4;
  const int cols = kind == 0 ? 32 : 16;
  const CUtensorMapDataType dt = kind == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = kind == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : kind == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap m;
  CUresult r;
  if (rstride > 1) {
    if (L % rstride) return fail(HG_ESTATE, "internal: phase-major map over %d rows, stride %d", L, rstride);
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(rstride), static_cast<cuuint64_t>(L / rstride),
                          static_cast<cuuint64_t>(B)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(c) * esz, static_cast<cuuint64_t>(rstride) * c * esz,
                             static_cast<cuuint64_t>(L) * c * esz};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(cols), 1, 32, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    r = enc(&m, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(c) * esz, static_cast<cuuint64_t>(L) * c * esz};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(cols), 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    r = enc(&m, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) return fail(HG_ECUDA, "cuTensorMapEncodeTiled (tile kind %d) failed (%d)", kind, static_cast<int>(r));
  plan->maps.put(key, m);
  *out = m;
  return HG_OK;
}
// packed (pre-swizzled) bf16 weight image seen as a plain 2-D tensor [rows][64]: a box of `box_rows`
// rows is one CTA's half of a weight tile in the CTA-pair kernel (no TMA swizzle: bytes land as packed)
static int make_weight_map(HgPlan* plan, const void* ptr, long long rows, int box_rows, CUtensorMap* out) {
  MapKey key(ptr, static_cast<int>(rows), 0, 64, -10, box_rows);
  if (plan->maps.get(key, out)) return HG_OK;
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(HG_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {64, static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HG_ECUDA, "cuTensorMapEncodeTiled (weights) failed (%d)", static_cast<int>(r));
  plan->maps.put(key, m);
  *out = m;
  return HG_OK;
}

// De-interleaved input slab of the time-folded pair kernel (conv_pair_fold.cu): the bf16 operand plane
// [B][L][C] seen as [B][blk][phase][r][C] with time row = (blk*F + phase)*d + r.  One box = nb block
// groups of one phase = nb*d consecutive shared-memory rows.  Block groups outside [0, ceil(L/(F*d)))
// read as zero; rows >= L inside the last group are inside this map's extent and are zeroed by the kernel.
static int make_fold_slab_map(HgPlan* plan, const void* ptr, int L, int B, int c, int d, int nb, CUtensorMap* out) {
  MapKey key(ptr, L, B, c, -(100 + d), nb);
  if (plan->maps.get(key, out)) return HG_OK;
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(HG_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const int f = 128 / c;
  const cuuint64_t rowb = static_cast<cuuint64_t>(c) * 2;
  const cuuint64_t nblk = (static_cast<cuuint64_t>(L) + static_cast<cuuint64_t>(f) * d - 1) / (static_cast<cuuint64_t>(f) * d);
  cuuint64_t dims[5] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(f), nblk,
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[4] = {rowb, rowb * d, rowb * d * f, rowb * static_cast<cuuint64_t>(L)};
  cuuint32_t box[5] = {static_cast<cuuint32_t>(c), static_cast<cuuint32_t>(d), 1, static_cast<cuuint32_t>(nb), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(HG_ECUDA, "cuTensorMapEncodeTiled (folded slab) failed (%d) for L=%d B=%d C=%d d=%d nb=%d", static_cast<int>(r), L, B,
                c, d, nb);
  plan->maps.put(key, m);
  *out = m;
  return HG_OK;
}

static int make_f32_tile_map(HgPlan* plan, const void* ptr, int L, int B, int c, CUtensorMap* out) {
  return make_tile_map(plan, ptr, L, B, c, 0, out);
}

// ------------------------------------------------------------------------------------------------
// layer table
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

static void setup_gemm_view(Layer& l) {
  if (l.kind == L_CONV) {
    l.ntaps = l.k;
    for (int j = 0; j < l.k; ++j) l.tap_off[j] = j * l.dil - l.pad;
    l.n_total = l.cout;
  } else if (l.kind == L_CONVT) {
    l.ntaps = (l.k + l.stride - 1) / l.stride;
    for (int m = 0; m < l.ntaps; ++m) l.tap_off[m] = -m;
    l.n_total = l.stride * l.cout;
  }
  l.tc = false;
  if (l.kind == L_POST) return;
  int cin_pad = 0;
  if (l.cin % 32 == 0) cin_pad = l.cin;
  else if (l.cin >= 64) cin_pad = (l.cin + 63) / 64 * 64;  // conv_pre: 80 mel bins -> 128
  if (cin_pad && l.n_total % 32 == 0 && l.cout % 4 == 0 && l.ntaps <= kMaxTaps) {
    l.tc = true;
    l.cin_pad = cin_pad;
    l.kc = cin_pad % 64 == 0 ? 64 : 32;
    l.nc = cin_pad / l.kc;
    l.n_tile = l.n_total % 256 == 0 ? 256 : l.n_total % 128 == 0 ? 128 : l.n_total % 64 == 0 ? 64 : 32;
    l.n_blocks = l.n_total / l.n_tile;
  } else {
    l.cin_pad = l.cin;
  }
}

static Layer make_conv(const std::string& name, int cin, int cout, int k, int dil) {
  Layer l;
  l.name = name; l.kind = L_CONV; l.cin = cin; l.cout = cout; l.k = k; l.dil = dil;
  l.pad = (k * dil - dil) / 2;  // get_padding, hifi/vocoder/utils.py:36-37
  setup_gemm_view(l);
  return l;
}
static Layer make_convT(const std::string& name, int cin, int cout, int k, int stride) {
  Layer l;
  l.name = name; l.kind = L_CONVT; l.cin = cin; l.cout = cout; l.k = k; l.stride = stride;
  l.pad = (k - stride) / 2;  // hifi/models.py:169
  setup_gemm_view(l);
  return l;
}

// device properties + the bring-up switches (HG_* environment variables)
static void init_plan_env(HgPlan* p, int device) {
  p->device = device;
  cudaDeviceProp pr;
  if (cudaGetDeviceProperties(&pr, device) == cudaSuccess) p->sm_count = pr.multiProcessorCount;
  p->desc_mode = env_int("HG_DESC_MODE", 0);
  p->force_ms = env_int("HG_TC_MS", 0);
  p->force_stages = env_int("HG_TC_STAGES", 0);
  p->force_ffma = env_int("HG_FORCE_FFMA", 0) != 0;
  p->ctas_per_sm = env_int("HG_TC_CTAS_PER_SM", 1);
  p->fuse_pairs = env_int("HG_FUSE_PAIRS", 1) != 0;
  p->fold_pairs = env_int("HG_FOLD", 1) != 0;
  p->tile_alternate = env_int("HG_TILE_ORDER", 1) != 0;
  p->tc2_in_bufs = env_int("HG_TC2_INBUFS", 2) == 1 ? 1 : 2;
  p->tc2_convt = env_int("HG_TC2_CONVT", 1) != 0;
  p->epi_tma_convt = env_int("HG_EPI_TMA_CONVT", 1) != 0;
  p->concurrent_elems = static_cast<long long>(env_int("HG_CONCURRENT_KELEMS", 2560)) * 1024;
  p->fold_force = env_int("HG_FOLD", 1) == 2;
  // Off by default: measured on B200 (16 x 800 frames) the fused ResBlock is SLOWER than its three fused pairs
  // (0.96 vs 0.73 ms at C = 64, 0.88 vs 0.71 ms at C = 32; profiles/r2_resblock_fusion_experiment.md) — the twelve
  // GEMM / epilogue steps of a tile are strictly sequential, so tensor pipe and epilogue warps never overlap, and
  // the HBM bytes it saves (36 -> 12 per element) do not buy that back.  Kept, tested bit-for-bit, behind HG_CHAIN=1.
  p->fuse_blocks = env_int("HG_CHAIN", 0) != 0;
  p->epi_tma = env_int("HG_EPI_TMA", 1) != 0;
  p->use_tc2 = env_int("HG_TC2", 1) != 0;
}

static int rb_dilations(const HgConfig& c) { return c.resblock_type == 1 ? 3 : 2; }

extern "C" int hg_plan_create(const HgConfig* cfg, int device, HgPlan** out) {
  if (!cfg || !out) return fail(HG_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->num_upsamples < 1 || cfg->num_upsamples > HG_MAX_UPS || cfg->num_kernels < 1 ||
      cfg->num_kernels > HG_MAX_KERNELS || cfg->num_mels < 1 || cfg->upsample_initial_channel < 1 ||
      (cfg->resblock_type != 1 && cfg->resblock_type != 2))
    return fail(HG_EINVAL, "bad HgConfig");
  if ((cfg->upsample_initial_channel >> cfg->num_upsamples) < 1)
    return fail(HG_EINVAL, "upsample_initial_channel too small for %d upsamplers", cfg->num_upsamples);
  int rc = check_device(device);
  if (rc) return rc;
  HgPlan* p = new HgPlan();
  p->cfg = *cfg;
  init_plan_env(p, device);
  const int uic = cfg->upsample_initial_channel;
  p->layers.push_back(make_conv("conv_pre", cfg->num_mels, uic, 7, 1));
  for (int i = 0; i < cfg->num_upsamples; ++i) {
    // The schedule assumes every stage is exactly rate x longer than its input (hifi/models.py:169 pads
    // (k - u) // 2, which gives L*u only for even k - u) and that the polyphase GEMM needs L + 1 rows
    // (true for k <= 3u).  Anything else is rejected here rather than silently mis-sized downstream.
    const int u = cfg->upsample_rates[i], k = cfg->upsample_kernel_sizes[i];
    if (u < 1 || k < u || ((k - u) & 1) || k > 3 * u || (k + u - 1) / u > kMaxTaps) {
      delete p;
      return fail(HG_EINVAL, "upsampler %d: (kernel %d, rate %d) unsupported — need rate >= 1, rate <= kernel <= 3*rate, "
                  "kernel - rate even (output length = rate x input length)", i, k, u);
    }
    p->layers.push_back(make_convT("ups." + std::to_string(i), uic >> i, uic >> (i + 1),
                                   cfg->upsample_kernel_sizes[i], cfg->upsample_rates[i]));
  }
  const int D = rb_dilations(*cfg);
  int ch = uic;
  for (int i = 0; i < cfg->num_upsamples; ++i) {
    ch = uic >> (i + 1);
    for (int j = 0; j < cfg->num_kernels; ++j) {
      const std::string base = "resblocks." + std::to_string(i * cfg->num_kernels + j);
      const int k = cfg->resblock_kernel_sizes[j];
      if (k < 1 || k > kMaxTaps || (k & 1) == 0) {
        delete p;
        return fail(HG_EINVAL, "resblock kernel size %d unsupported (odd, <= %d)", k, kMaxTaps);
      }
      if (cfg->resblock_type == 1) {
        for (int m = 0; m < D; ++m)
          p->layers.push_back(make_conv(base + ".convs1." + std::to_string(m), ch, ch, k, cfg->resblock_dilation_sizes[j][m]));
        for (int m = 0; m < D; ++m) p->layers.push_back(make_conv(base + ".convs2." + std::to_string(m), ch, ch, k, 1));
      } else {
        for (int m = 0; m < D; ++m)
          p->layers.push_back(make_conv(base + ".convs." + std::to_string(m), ch, ch, k, cfg->resblock_dilation_sizes[j][m]));
      }
    }
  }
  Layer post;
  post.name = "conv_post"; post.kind = L_POST; post.cin = ch; post.cout = 1; post.k = 7; post.pad = 3;
  p->layers.push_back(post);
  for (size_t i = 0; i < p->layers.size(); ++i) p->by_name[p->layers[i].name] = static_cast<int>(i);
  // Stages narrower than 16 channels (V2-style configs end at 8) have no tensor-core tiling: K = 16 is the
  // smallest bf16 MMA.  For the bf16 schedule they are carried zero-padded to 16 channels — weights, bias and
  // activations; a padded channel is 0 everywhere, so nothing changes for the real ones — and run on the
  // 16-channel kernels.  The fp32 schedules keep the exact shapes (CUDA-core kernels).
  if (env_int("HG_PAD_NARROW", 1)) {
    std::vector<Layer> padded = p->layers;
    bool any = false;
    int li = 1 + cfg->num_upsamples;
    const int per_stage = cfg->num_kernels * (cfg->resblock_type == 1 ? 2 * D : D);
    for (int i = 0; i < cfg->num_upsamples; ++i, li += per_stage) {
      const int c = uic >> (i + 1);
      if (c >= 16) continue;
      any = true;
      auto repad = [&](Layer& l, int cin, int cout) {
        Layer n = l.kind == L_CONVT ? make_convT(l.name, cin, cout, l.k, l.stride) : make_conv(l.name, cin, cout, l.k, l.dil);
        n.cin_w = l.cin_w ? l.cin_w : l.cin;
        n.cout_w = l.cout_w ? l.cout_w : l.cout;
        l = n;
      };
      Layer& up = padded[1 + i];
      repad(up, up.cin, 16);
      for (int q = 0; q < per_stage; ++q) repad(padded[li + q], 16, 16);
      if (i + 1 < cfg->num_upsamples) {
        Layer& nx = padded[2 + i];
        repad(nx, 16, nx.cout);
      } else {
        Layer& po = padded.back();
        po.cin_w = po.cin;
        po.cin = 16;
      }
    }
    if (any) p->layers_pad = padded;
  }
  // Latency schedule (BASELINE cfg-1: one 256-frame utterance): a 256-channel conv over 2 048 rows is 8 CTA-pair
  // tiles of 256 x 256 — 16 of 148 SMs busy, each issuing the whole K loop.  With 64-column N tiles the same layer is
  // 64 work items (16 M tiles x 4 N blocks) with a quarter of the MMAs and of the weight stream each.
  if (env_int("HG_SMALL_TILES", 1)) {
    bool any = false;
    std::vector<Layer> small(p->layers.size());
    for (size_t i = 0; i < p->layers.size(); ++i) {
      const Layer& l = p->layers[i];
      if (l.kind != L_CONV || !l.tc || l.n_tile != 256) continue;
      small[i] = l;
      small[i].n_tile = 64;
      small[i].n_blocks = l.n_total / 64;
      any = true;
    }
    if (any) p->layers_small = small;
  }
  *out = p;
  return HG_OK;
}

// the layer table a forward in `precision` runs on (see hg_plan_create)
static const std::vector<Layer>& active_layers(const HgPlan* plan, int precision) {
  return (precision == HG_PREC_BF16 && !plan->layers_pad.empty() && !plan->force_ffma) ? plan->layers_pad : plan->layers;
}
static int layer_index(const HgPlan* plan, const Layer* l) {
  if (!plan->layers_small.empty() && l >= plan->layers_small.data() && l < plan->layers_small.data() + plan->layers_small.size())
    return static_cast<int>(l - plan->layers_small.data());
  if (!plan->layers_pad.empty() && l >= plan->layers_pad.data() && l < plan->layers_pad.data() + plan->layers_pad.size())
    return static_cast<int>(l - plan->layers_pad.data());
  return static_cast<int>(l - plan->layers.data());
}

// ------------------------------------------------------------------------------------------------
// weight repacking
static inline uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// GEMM-view weight Wg(tap, n, c) from the reference layouts.
static inline float gemm_weight(const Layer& l, const float* W, int t, int n, int c) {
  const int cin_w = l.cin_w ? l.cin_w : l.cin, cout_w = l.cout_w ? l.cout_w : l.cout;  // dims of W itself
  if (c >= cin_w) return 0.f;
  if (l.kind == L_CONV) return n < cout_w ? W[(static_cast<size_t>(n) * cin_w + c) * l.k + t] : 0.f;
  const int r = n / l.cout, o = n % l.cout;
  const int j = r + l.stride * t;
  return (j < l.k && o < cout_w) ? W[(static_cast<size_t>(c) * cout_w + o) * l.k + j] : 0.f;
}

static int upload(const void* host, size_t bytes, void** dev) {
  CUDA_TRY(cudaMalloc(dev, bytes));
  CUDA_TRY(cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice));
  return HG_OK;
}

static void free_layer(Layer& l) {
  cudaFree(l.w_hi); cudaFree(l.w_lo); cudaFree(l.w_ffma); cudaFree(l.bias); cudaFree(l.w_post); cudaFree(l.bias_fold);
  l.w_hi = l.w_lo = nullptr; l.w_ffma = nullptr; l.bias = nullptr; l.w_post = nullptr; l.bias_fold = nullptr;
  l.loaded = false;
}

static int pack_layer(Layer& l, const float* W, const float* bias) {
  int rc;
  const int cin_w = l.cin_w ? l.cin_w : l.cin, cout_w = l.cout_w ? l.cout_w : l.cout;
  if (l.kind == L_POST) {
    std::vector<float> wp(static_cast<size_t>(l.k) * l.cin, 0.f);
    for (int j = 0; j < l.k; ++j)
      for (int c = 0; c < cin_w; ++c) wp[static_cast<size_t>(j) * l.cin + c] = W[static_cast<size_t>(c) * l.k + j];
    if ((rc = upload(wp.data(), wp.size() * 4, reinterpret_cast<void**>(&l.w_post)))) return rc;
    l.w_post_host = wp;
    l.bias_post = bias[0];
    l.loaded = true;
    return HG_OK;
  }
  std::vector<float> bfull(l.n_total);
  for (int n = 0; n < l.n_total; ++n) bfull[n] = (n % l.cout) < cout_w ? bias[n % l.cout] : 0.f;
  if ((rc = upload(bfull.data(), bfull.size() * 4, reinterpret_cast<void**>(&l.bias)))) return rc;
  if (l.kind == L_CONV && (l.cout == 16 || l.cout == 32 || l.cout == 64)) {
    // the time-folded pair kernel's epilogue sees F = 128 / C output rows as one 128-column row
    std::vector<float> bf(128);
    for (int n = 0; n < 128; ++n) bf[n] = (n % l.cout) < cout_w ? bias[n % l.cout] : 0.f;
    if ((rc = upload(bf.data(), bf.size() * 4, reinterpret_cast<void**>(&l.bias_fold)))) return rc;
  }
  {  // CUDA-core layout [tap][cin][n_total]
    std::vector<float> wf(static_cast<size_t>(l.ntaps) * l.cin * l.n_total);
    for (int t = 0; t < l.ntaps; ++t)
      for (int c = 0; c < l.cin; ++c)
        for (int n = 0; n < l.n_total; ++n)
          wf[(static_cast<size_t>(t) * l.cin + c) * l.n_total + n] = gemm_weight(l, W, t, n, c);
    if ((rc = upload(wf.data(), wf.size() * 4, reinterpret_cast<void**>(&l.w_ffma)))) return rc;
  }
  if (!l.tc && l.kind == L_CONV && l.cin == 16 && l.cout == 16) {
    // 16-channel convs have no tiling in conv_tc.cu (K chunks of 32 / 64); the time-folded pair kernel takes
    // them as [tap][16 rows][16] tiles of 32-byte rows in the 32B-swizzle layout (chunk ^= (row >> 2) & 1)
    const size_t tile = 16 * 32;
    std::vector<uint8_t> hi(tile * l.k);
    for (int t = 0; t < l.k; ++t)
      for (int row = 0; row < 16; ++row)
        for (int kk = 0; kk < 16; ++kk) {
          const uint16_t h = f32_to_bf16_rn(gemm_weight(l, W, t, row, kk));
          const size_t off = t * tile + static_cast<size_t>(row) * 32 + ((((kk >> 3) ^ ((row >> 2) & 1))) << 4) + ((kk & 7) << 1);
          memcpy(&hi[off], &h, 2);
        }
    if ((rc = upload(hi.data(), hi.size(), reinterpret_cast<void**>(&l.w_hi)))) return rc;
  }
  if (l.tc) {  // swizzled bf16 tiles [n_blk][chunk][tap][n_tile rows][kc], hi and lo planes
    const int rowb = l.kc * 2;
    const size_t tile = static_cast<size_t>(l.n_tile) * rowb;
    const size_t total = tile * l.n_blocks * l.nc * l.ntaps;
    std::vector<uint8_t> hi(total), lo(total);
    for (int nb = 0; nb < l.n_blocks; ++nb)
      for (int c = 0; c < l.nc; ++c)
        for (int t = 0; t < l.ntaps; ++t) {
          const size_t base = ((static_cast<size_t>(nb) * l.nc + c) * l.ntaps + t) * tile;
          for (int row = 0; row < l.n_tile; ++row) {
            const int swz = l.kc == 64 ? (row & 7) : ((row >> 1) & 3);
            for (int kk = 0; kk < l.kc; ++kk) {
              const float v = gemm_weight(l, W, t, nb * l.n_tile + row, c * l.kc + kk);
              const uint16_t h = f32_to_bf16_rn(v);
              const uint16_t lw = f32_to_bf16_rn(v - bf16_to_f32(h));
              const size_t off = base + static_cast<size_t>(row) * rowb + (((kk >> 3) ^ swz) << 4) + ((kk & 7) << 1);
              memcpy(&hi[off], &h, 2);
              memcpy(&lo[off], &lw, 2);
            }
          }
        }
    if ((rc = upload(hi.data(), total, reinterpret_cast<void**>(&l.w_hi)))) return rc;
    if ((rc = upload(lo.data(), total, reinterpret_cast<void**>(&l.w_lo)))) return rc;
  }
  l.loaded = true;
  return HG_OK;
}

extern "C" int hg_plan_upload_weight(HgPlan* plan, const char* name, const float* weight, const int64_t* shape,
                                     int ndim, const float* bias, int64_t bias_len) {
  if (!plan || !name || !weight || !shape || !bias) return fail(HG_EINVAL, "null argument");
  if (plan->finalized) return fail(HG_ESTATE, "plan is finalized (immutable)");
  auto it = plan->by_name.find(name);
  if (it == plan->by_name.end()) return fail(HG_EINVAL, "unexpected key in state_dict: %s", name);
  Layer& l = plan->layers[it->second];
  const int cin_sd = l.cin_w ? l.cin_w : l.cin, cout_sd = l.cout_w ? l.cout_w : l.cout;  // state_dict dims
  int64_t want[3];
  if (l.kind == L_CONVT) { want[0] = cin_sd; want[1] = cout_sd; want[2] = l.k; }
  else { want[0] = cout_sd; want[1] = cin_sd; want[2] = l.k; }
  if (ndim != 3 || shape[0] != want[0] || shape[1] != want[1] || shape[2] != want[2])
    return fail(HG_EINVAL, "size mismatch for %s.weight: expected [%lld,%lld,%lld]", name,
                static_cast<long long>(want[0]), static_cast<long long>(want[1]), static_cast<long long>(want[2]));
  if (bias_len != cout_sd) return fail(HG_EINVAL, "size mismatch for %s.bias: expected [%d]", name, cout_sd);
  DEVICE_SCOPE(plan->device);
  if (l.loaded) free_layer(l);
  int rc = pack_layer(l, weight, bias);
  if (!rc && !plan->layers_small.empty() && plan->layers_small[it->second].tc) {
    Layer& ls = plan->layers_small[it->second];  // the same tensors in 64-column N tiles (latency schedule)
    if (ls.loaded) free_layer(ls);
    ls.w_hi = ls.w_lo = nullptr; ls.w_ffma = nullptr; ls.bias = nullptr; ls.bias_fold = nullptr; ls.w_post = nullptr;
    rc = pack_layer(ls, weight, bias);
  }
  if (rc || plan->layers_pad.empty()) return rc;
  Layer& lp = plan->layers_pad[it->second];  // the same tensors, zero-padded, for the bf16 schedule
  if (lp.loaded) free_layer(lp);
  return pack_layer(lp, weight, bias);
}

extern "C" int hg_plan_finalize(HgPlan* plan) {
  if (!plan) return fail(HG_EINVAL, "null plan");
  for (auto& l : plan->layers)
    if (!l.loaded) return fail(HG_ESTATE, "missing key in state_dict: %s", l.name.c_str());
  plan->finalized = true;
  return HG_OK;
}

extern "C" int hg_plan_destroy(HgPlan* plan) {
  if (!plan) return HG_OK;
  DeviceScope scope(plan->device);
  for (auto& l : plan->layers) free_layer(l);
  for (auto& l : plan->layers_pad) free_layer(l);
  for (auto& l : plan->layers_small) if (l.loaded) free_layer(l);
  for (cudaStream_t sst : plan->side_stream) if (sst) cudaStreamDestroy(sst);
  for (cudaEvent_t ev : plan->ev_done) if (ev) cudaEventDestroy(ev);
  if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
  delete plan;
  return HG_OK;
}

// ------------------------------------------------------------------------------------------------
// optional per-launch event recording (hg_profile_forward)
// what a launch actually ran (kernel family + tiling), reported through hg_profile_launch_info
struct LaunchRec {
  int path = HG_PATH_CUDA_CORE, n_tile = 0, kc = 0, ms = 0, stages = 0, nbuf = 0, resident = 0;
  size_t smem = 0;
};
struct ProfSink {
  cudaStream_t st;
  std::vector<cudaEvent_t> ev;   // ev[0] = before the first launch, ev[i+1] = after launch i
  std::vector<int> layer;        // plan layer index per launch (-1 = mel repack)
  std::vector<LaunchRec> rec;
};
static thread_local ProfSink* g_prof = nullptr;
static thread_local std::vector<std::pair<int, LaunchRec>> g_last_profile;  // (layer, record) of the last hg_profile_forward
static void prof_mark(int layer_index, const LaunchRec& rec = LaunchRec()) {
  if (!g_prof) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, g_prof->st);
  g_prof->ev.push_back(e);
  if (layer_index > -2) { g_prof->layer.push_back(layer_index); g_prof->rec.push_back(rec); }
}
static LaunchRec make_rec(int path, int n_tile = 0, int kc = 0, int ms = 0, int stages = 0, int nbuf = 0, bool resident = false,
                          size_t smem = 0) {
  LaunchRec r;
  r.path = path; r.n_tile = n_tile; r.kc = kc; r.ms = ms; r.stages = stages; r.nbuf = nbuf; r.resident = resident ? 1 : 0; r.smem = smem;
  return r;
}

// ------------------------------------------------------------------------------------------------
// ragged batches (hg_forward_ragged): item b needs (frames[b] + halo) mel frames' worth of rows at
// every layer — the generator's receptive reach, SURVEY.md App. E — and nothing beyond
struct RaggedCtx {
  const int32_t* frames;  // host, [B]
  int halo;               // mel frames
  int T;
};
static thread_local const RaggedCtx* g_rag = nullptr;

// output window of the last layer (hg_forward_window)
struct WindowCtx {
  long long item_stride;
  int skip, keep;
};
static thread_local const WindowCtx* g_win = nullptr;

// valid GEMM rows per item for a launch whose input has L_in rows per item (rows = L_in + extra)
static const RaggedItems* ragged_items(RaggedItems* store, int B, int L_in, int extra, bool with_halo = true) {
  if (!g_rag) return nullptr;
  const long long per_frame = L_in / g_rag->T;  // exact: every stage length is T * prod(rates so far)
  store->n = B;
  for (int b = 0; b < B; ++b) {
    const long long v = (static_cast<long long>(g_rag->frames[b]) + (with_halo ? g_rag->halo : 0)) * per_frame + extra;
    store->valid_rows[b] = static_cast<int>(std::min<long long>(v, static_cast<long long>(L_in) + extra));
  }
  return store;
}

// receptive reach of one output sample beyond the frames that own it, in samples per side
// (3 258 for V1; same walk as tts_king_b200/parallel.py::receptive_reach_samples)
static long long reach_samples(const HgConfig& c) {
  long long reach = 3;  // conv_pre k7
  const int nd = c.resblock_type == 1 ? 3 : 2;
  for (int i = 0; i < c.num_upsamples; ++i) {
    const int u = c.upsample_rates[i], k = c.upsample_kernel_sizes[i];
    reach = reach * u + (k - u + 1) / 2;
    long long blk = 0;
    for (int j = 0; j < c.num_kernels; ++j) {
      const int half = (c.resblock_kernel_sizes[j] - 1) / 2;
      long long r = 0;
      for (int d = 0; d < nd; ++d) r += static_cast<long long>(half) * c.resblock_dilation_sizes[j][d];
      if (c.resblock_type == 1) r += static_cast<long long>(nd) * half;
      blk = std::max(blk, r);
    }
    reach += blk;
  }
  return reach + 3;  // conv_post k7
}

static int halo_frames_of(const HgConfig& c) {
  long long hop = 1;
  for (int i = 0; i < c.num_upsamples; ++i) hop *= c.upsample_rates[i];
  return static_cast<int>((reach_samples(c) + hop - 1) / hop);
}

// ------------------------------------------------------------------------------------------------
// launching one GEMM-shaped layer
struct OperandBuf {
  void* a0 = nullptr;
  void* a1 = nullptr;
};

static int a_fmt_of(int precision) {
  return precision == HG_PREC_BF16 ? A_BF16 : precision == HG_PREC_FP32 ? A_BF16_SPLIT : A_F32;
}

static bool use_tc(const HgPlan* plan, const Layer& l, int precision) {
  return l.tc && precision != HG_PREC_FP32_FFMA && !plan->force_ffma;
}

static TcTiling choose_tiling(const HgPlan* plan, const Layer& l, bool split, int slot = 6144, int max_ms = 0) {
  TcTiling t;
  int min_off = l.tap_off[0], max_off = l.tap_off[0];
  for (int j = 1; j < l.ntaps; ++j) {
    min_off = std::min(min_off, l.tap_off[j]);
    max_off = std::max(max_off, l.tap_off[j]);
  }
  t.min_off = min_off;
  const int span = max_off - min_off;
  // two accumulator buffers of MS x N_T fp32 columns must fit the 512 TMEM columns
  int ms = l.n_tile == 256 ? 1 : l.n_tile == 128 ? 2 : l.n_tile == 64 ? 2 : 4;
  if (plan->force_ms) ms = plan->force_ms;
  if (max_ms && ms > max_ms) ms = max_ms;
  while (2 * ms * l.n_tile > 512) ms >>= 1;
  const size_t kMaxSmem = 227 * 1024;
  const int total_stages = l.nc * l.ntaps * (split ? 2 : 1);
  for (;; ms >>= 1) {
    t.ms = ms;
    const int need = ms * 128 + span;
    t.nboxes = (need + 255) / 256;
    t.box_rows = (((need + t.nboxes - 1) / t.nboxes) + 7) / 8 * 8;
    t.slab_rows = t.nboxes * t.box_rows;
    t.stages = 0;
    // 1) weights resident for the CTA's lifetime (single N block only) with a double-buffered slab
    if (l.n_blocks == 1 && !plan->force_stages &&
        conv_tc_smem_bytes(l.n_tile, l.kc, split, t.slab_rows, 2, total_stages, slot) <= kMaxSmem) {
      t.resident = true;
      t.stages = total_stages;
      t.nbuf = 2;
      if (conv_tc_smem_bytes(l.n_tile, l.kc, split, t.slab_rows, 3, total_stages, slot) <= kMaxSmem) t.nbuf = 3;
      break;
    }
    // 2) weights streamed through a ring: at least 3 stages next to a double-buffered slab
    t.resident = false;
    t.nbuf = 2;
    int s = 8;
    while (s >= 2 && conv_tc_smem_bytes(l.n_tile, l.kc, split, t.slab_rows, t.nbuf, s, slot) > kMaxSmem) --s;
    if (plan->force_stages) s = plan->force_stages;
    if (s >= 2 && conv_tc_smem_bytes(l.n_tile, l.kc, split, t.slab_rows, t.nbuf, s, slot) <= kMaxSmem) {
      t.stages = s;
      if (s >= 4 && conv_tc_smem_bytes(l.n_tile, l.kc, split, t.slab_rows, 3, s, slot) <= kMaxSmem) t.nbuf = 3;
      break;
    }
    if (ms == 1) break;  // does not fit at all -> CUDA-core path
  }
  t.smem = conv_tc_smem_bytes(l.n_tile, l.kc, split, t.slab_rows, t.nbuf, t.stages, slot);
  return t;
}

// in: operand planes [B][L_in][cin_pad].  epi: fully populated except bias/a_fmt.
static int run_layer(HgPlan* plan, const Layer& l, int precision, int B, int L_in, const OperandBuf& in, EpiParams epi,
                     cudaStream_t st) {
  const int rows = l.kind == L_CONVT ? L_in + 1 : L_in;
  const long long L_out = l.kind == L_CONVT ? static_cast<long long>(L_in - 1) * l.stride - 2 * l.pad + l.k : L_in;
  epi.bias = l.bias;
  epi.a_fmt = a_fmt_of(precision);
  const int c_store = l.n_store ? l.n_store : l.cout;  // width of the output tensor in memory
  epi.out_batch_stride = L_out * c_store;
  epi.out_extent = L_out * c_store;
  epi.out_row_stride = l.n_store ? l.n_store : l.n_total;
  epi.n_valid = l.n_store;
  epi.out_offset = l.kind == L_CONVT ? -static_cast<long long>(l.pad) * l.cout : 0;
  RaggedItems rag_store;
  const RaggedItems* rag = ragged_items(&rag_store, B, L_in, rows - L_in);
  // how many 256-row tiles the launch has: below an eighth of the SM count it is latency-, not throughput-bound
  const bool few_tiles = static_cast<long long>(B) * ((rows + 255) / 256) * 8 <= plan->sm_count;
  if (few_tiles && precision == HG_PREC_BF16 && !plan->layers_small.empty() && l.kind == L_CONV && l.n_tile == 256 && !plan->force_ffma) {
    const int idx = layer_index(plan, &l);
    if (idx >= 0 && idx < static_cast<int>(plan->layers_small.size()) && plan->layers_small[idx].loaded)
      return run_layer(plan, plan->layers_small[idx], precision, B, L_in, in, epi, st);
  }
  if (use_tc(plan, l, precision)) {
    const bool split = precision == HG_PREC_FP32;
    // TMA epilogue for same-length convs without MRF accumulate (68 of the 78 layers of V1); its
    // per-warp slot holds the residual-in, x-out and operand-out tiles the layer actually uses
    // ... and for the polyphase upsamplers, whose phases are row-strided boxes of the output (HG_EPI_TMA_CONVT=0: generic)
    const bool convt_tma = l.kind == L_CONVT && plan->epi_tma_convt && !epi.res && l.stride > 1 && L_out % l.stride == 0;
    bool tma_epi = plan->epi_tma && (l.kind == L_CONV || convt_tma) && !epi.acc_in && epi.post_div <= 0.f && l.cout % 16 == 0 && !l.n_store;
    int slot = 2048;  // generic epilogue: one 32x16 fp32 transpose tile per warp
    if (tma_epi) {
      slot = (epi.res ? 2048 : 0) + (epi.out_x ? 2048 : 0) + (epi.out_a0 ? (split ? 2048 : 1024) : 0);
      slot = std::max(slot, 1024);
    }
    // CTA-pair kernel (conv_tc2.cu) for the wide same-length convs in bf16 mode
    // — whenever the single-CTA kernel could not keep the layer's weights resident in shared memory
    // (its TMA epilogue also takes the MRF running sum as a second input tile, so the last conv of a
    // ResBlock — xs += x, / num_kernels — stays on this path)
    // ... and the polyphase upsamplers with several N blocks (N_total = stride * C_out): their weight stream per
    // work item (N_T x K: 256 KB at ups.1) starves a single CTA's ring; the pair halves it per SM (HG_TC2_CONVT=0: off)
    const bool convt2 = convt_tma && plan->tc2_convt && l.n_blocks > 1 && !epi.acc_in;
    const bool tma_epi2 = plan->epi_tma && (l.kind == L_CONV || convt2) && l.cout % 16 == 0 && !l.n_store;
    // conv_tc2.cu keeps TWO residual tiles in flight per epilogue warp where the layer only has a residual input; with
    // the MRF running sum as a second input the doubled slot would cost the weight ring its depth (HG_TC2_INBUFS=1: one)
    const int in_tile = (epi.res ? 2048 : 0) + (epi.acc_in ? 2048 : 0);
    const int in_bufs = (plan->tc2_in_bufs == 2 && epi.res && !epi.acc_in) ? 2 : 1;
    const int slot2 = std::max(1024, in_bufs * in_tile + (epi.out_x ? 2048 : 0) + (epi.out_a0 ? 1024 : 0));
    if (plan->use_tc2 && tma_epi2 && !split && (l.n_blocks == 1 || convt2) && l.kc == 64 && (l.n_tile == 128 || l.n_tile == 256) &&
        !choose_tiling(plan, l, split, tma_epi ? slot : 2048).resident) {
      const int slot = slot2;  // shadows the single-CTA kernel's slot size inside this branch
      int min_off = l.tap_off[0], max_off = l.tap_off[0];
      for (int j = 1; j < l.ntaps; ++j) { min_off = std::min(min_off, l.tap_off[j]); max_off = std::max(max_off, l.tap_off[j]); }
      // two 128-row sub-tiles per CTA unless that leaves most of the GPU without a tile (short inputs)
      const int ms = (l.n_tile == 256 || static_cast<long long>(B) * ((rows + 511) / 512) * 2 <= plan->sm_count / 2) ? 1 : 2;
      const int need = ms * 128 + (max_off - min_off);
      const int nboxes = (need + 255) / 256;
      const int box_rows = (((need + nboxes - 1) / nboxes) + 7) / 8 * 8;
      const int slab_rows = nboxes * box_rows;
      const size_t kMaxSmem = 227 * 1024;
      int nbuf = 2, stages = 12;
      while (stages >= 3 && conv_tc2_smem_bytes(l.n_tile, slab_rows, nbuf, stages, slot) > kMaxSmem) --stages;
      if (stages >= 3) {
        if (conv_tc2_smem_bytes(l.n_tile, slab_rows, 3, stages, slot) <= kMaxSmem && stages >= 6) nbuf = 3;
        TcConvParams p;
        memset(&p, 0, sizeof(p));
        p.B = B; p.rows = rows;
        p.tiles_per_item = (rows + 2 * ms * 128 - 1) / (2 * ms * 128);
        p.nc = l.nc; p.ntaps = l.ntaps;
        for (int j = 0; j < l.ntaps; ++j) p.tap_row[j] = l.tap_off[j] - min_off;
        p.min_off = min_off; p.slab_rows = slab_rows; p.box_rows = box_rows; p.nboxes = nboxes;
        p.nbuf = nbuf; p.stages = stages; p.n_blocks = l.n_blocks;
        p.total_work = ragged_fill(&p.rag, rag, B, rows, 2 * ms * 128) * l.n_blocks;
        p.epi = epi; p.epi_tma = 1; p.epi_slot_bytes = slot; p.in_bufs = in_bufs;
        CUtensorMap maps[6];
        int rc = make_operand_map(plan, in.a0, L_in, B, l.cin_pad, 64, box_rows, &maps[0]);
        if (rc) return rc;
        if ((rc = make_weight_map(plan, l.w_hi, static_cast<long long>(l.n_blocks) * l.nc * l.ntaps * l.n_tile, l.n_tile / 2, &maps[1]))) return rc;
        maps[2] = maps[3] = maps[4] = maps[5] = maps[0];
        const int rs = l.kind == L_CONVT ? l.stride : 1;
        const int Lo = static_cast<int>(L_out);
        if (l.kind == L_CONVT) { p.out_cmod = l.cout; p.out_rstride = l.stride; p.out_roff = -l.pad; }
        if (epi.acc_in) { p.has_acc = 1; if ((rc = make_tile_map(plan, epi.acc_in, L_in, B, l.cout, 1, &maps[5]))) return rc; }
        if (epi.res) { p.has_res = 1; if ((rc = make_tile_map(plan, epi.res, L_in, B, l.cout, 1, &maps[2]))) return rc; }
        if (epi.out_x) { p.has_x = 1; if ((rc = make_tile_map(plan, epi.out_x, Lo, B, l.cout, 1, &maps[3], rs))) return rc; }
        if (epi.out_a0) { p.has_a = 1; if ((rc = make_tile_map(plan, epi.out_a0, Lo, B, l.cout, 2, &maps[4], rs))) return rc; }
        const int pairs = std::min(p.total_work, plan->sm_count / 2);
        const size_t smem = conv_tc2_smem_bytes(l.n_tile, slab_rows, nbuf, stages, slot);
        cudaError_t e = launch_conv_tc2(l.n_tile, ms, maps, p, smem, 2 * pairs, st);
        if (e != cudaSuccess) return fail(HG_ECUDA, "conv_tc2 launch (%s): %s", l.name.c_str(), cudaGetErrorString(e));
        if (g_prof) prof_mark(layer_index(plan, &l), make_rec(HG_PATH_TC_CTA_PAIR, l.n_tile, 64, ms, stages, nbuf, false, smem));
        return HG_OK;
      }
    }
    TcTiling t = choose_tiling(plan, l, split, slot, few_tiles ? 1 : 0);
    if (tma_epi && (t.stages < 3 && !t.resident)) {
      // the TMA epilogue's tiles squeeze the weight ring too far (fp32 split mode at 256 channels):
      // use the generic epilogue's 2 KB transpose slots instead
      TcTiling t2 = choose_tiling(plan, l, split, 2048);
      if (t2.stages > t.stages) { t = t2; tma_epi = false; slot = 2048; }
    }
    if (t.stages >= 2) {
      TcConvParams p;
      memset(&p, 0, sizeof(p));
      p.B = B; p.rows = rows;
      p.tiles_per_item = (rows + t.ms * 128 - 1) / (t.ms * 128);
      p.nc = l.nc; p.ntaps = l.ntaps;
      for (int j = 0; j < l.ntaps; ++j) p.tap_row[j] = l.tap_off[j] - t.min_off;
      p.min_off = t.min_off; p.slab_rows = t.slab_rows; p.box_rows = t.box_rows; p.nboxes = t.nboxes;
      p.nbuf = t.nbuf; p.stages = t.stages; p.desc_mode = plan->desc_mode;
      p.w_resident = t.resident ? 1 : 0;
      p.n_blocks = l.n_blocks;
      p.total_work = ragged_fill(&p.rag, rag, B, rows, t.ms * 128) * l.n_blocks;
      p.w_hi = l.w_hi; p.w_lo = l.w_lo; p.epi = epi;
      CUtensorMap maps[6];
      int rc = make_operand_map(plan, in.a0, L_in, B, l.cin_pad, l.kc, t.box_rows, &maps[0]);
      if (rc) return rc;
      maps[1] = maps[0];
      if (split && (rc = make_operand_map(plan, in.a1, L_in, B, l.cin_pad, l.kc, t.box_rows, &maps[1]))) return rc;
      maps[2] = maps[3] = maps[4] = maps[5] = maps[0];
      // TMA epilogue for same-length convs without MRF accumulate (68 of the 78 layers of V1)
      p.epi_slot_bytes = slot;
      if (tma_epi) {
        p.epi_tma = 1;
        const int rs = l.kind == L_CONVT ? l.stride : 1;
        const int Lo = static_cast<int>(L_out);
        if (l.kind == L_CONVT) { p.out_cmod = l.cout; p.out_rstride = l.stride; p.out_roff = -l.pad; }
        if (epi.res) { p.has_res = 1; if ((rc = make_tile_map(plan, epi.res, L_in, B, l.cout, 1, &maps[2]))) return rc; }
        if (epi.out_x) { p.has_x = 1; if ((rc = make_tile_map(plan, epi.out_x, Lo, B, l.cout, 1, &maps[3], rs))) return rc; }
        if (epi.out_a0) {
          p.has_a = 1;
          if ((rc = make_tile_map(plan, epi.out_a0, Lo, B, l.cout, 2, &maps[4], rs))) return rc;
          if (split && (rc = make_tile_map(plan, epi.out_a1, Lo, B, l.cout, 2, &maps[5], rs))) return rc;
        }
      }
      const int grid = std::min(p.total_work, plan->sm_count * std::max(1, plan->ctas_per_sm));
      static long long* dbg_buf = nullptr;
      const char* dbg_layer = getenv("HG_TC_DEBUG_TIMING");  // layer name to instrument (bring-up only)
      if (dbg_layer && l.name == dbg_layer) {
        if (!dbg_buf) cudaMalloc(&dbg_buf, 256 * 8 * sizeof(long long));
        cudaMemsetAsync(dbg_buf, 0, 256 * 8 * sizeof(long long), st);
        p.dbg = dbg_buf;
      }
      cudaError_t e = launch_conv_tc(l.n_tile, l.kc, t.ms, split, maps, p, l.n_blocks, t.smem, grid, st);
      if (p.dbg && e == cudaSuccess) {
        std::vector<long long> h(256 * 8);
        cudaStreamSynchronize(st);
        cudaMemcpy(h.data(), dbg_buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        double tot = 0, a = 0, sl = 0, w = 0, n = 0;
        for (int i = 0; i < grid; ++i) { tot += h[i * 8]; a += h[i * 8 + 1]; sl += h[i * 8 + 2]; w += h[i * 8 + 3]; n += h[i * 8 + 4]; }
        fprintf(stderr, "[hg dbg] %s grid=%d tiles/cta=%.1f MMA-warp cycles: total=%.0f wait acc_empty=%.0f (%.1f%%) slab=%.0f (%.1f%%) weights=%.0f (%.1f%%)\n",
                l.name.c_str(), grid, n / grid, tot / grid, a / grid, 100 * a / tot, sl / grid, 100 * sl / tot, w / grid, 100 * w / tot);
      }
      if (e != cudaSuccess) return fail(HG_ECUDA, "conv_tc launch (%s): %s", l.name.c_str(), cudaGetErrorString(e));
      if (g_prof) prof_mark(layer_index(plan, &l), make_rec(HG_PATH_TC, l.n_tile, l.kc, t.ms, t.stages, t.nbuf, t.resident, t.smem));
      return HG_OK;
    }
  }
  // very narrow same-length convs (V2-style 16- / 8-channel stages): dedicated CUDA-core kernel
  if (l.kind == L_CONV && l.cin == l.cout && (l.cin == 8 || l.cin == 16) && (l.k & 1) && !plan->force_ffma) {
    NarrowConvParams np;
    memset(&np, 0, sizeof(np));
    np.B = B; np.L = L_in; np.k = l.k; np.dil = l.dil;
    np.a0 = in.a0; np.a1 = in.a1; np.a_fmt = a_fmt_of(precision);
    np.w = l.w_ffma; np.epi = epi;
    cudaError_t e = launch_conv_narrow(l.cin, np, st, rag);
    if (e != cudaSuccess) return fail(HG_ECUDA, "conv_narrow launch (%s): %s", l.name.c_str(), cudaGetErrorString(e));
    if (g_prof) prof_mark(layer_index(plan, &l), make_rec(HG_PATH_NARROW));
    return HG_OK;
  }
  // 16 -> 16 channel upsampler with two taps per phase (the step into a stage that runs padded to 16 channels)
  if (l.kind == L_CONVT && l.cin == 16 && l.cout == 16 && l.ntaps == 2 && precision == HG_PREC_BF16 && !plan->force_ffma &&
      !epi.res && !epi.acc_in && epi.post_div <= 0.f) {
    cudaError_t e = launch_convt_narrow16(in.a0, l.w_ffma, B, L_in, l.stride, l.pad, epi, st, rag);
    if (e != cudaSuccess) return fail(HG_ECUDA, "convt_narrow16 launch (%s): %s", l.name.c_str(), cudaGetErrorString(e));
    if (g_prof) prof_mark(layer_index(plan, &l), make_rec(HG_PATH_NARROW));
    return HG_OK;
  }
  FfmaConvParams f;
  memset(&f, 0, sizeof(f));
  f.B = B; f.L_in = L_in; f.rows = rows; f.cin = l.cin; f.n_total = l.n_total; f.ntaps = l.ntaps;
  f.a_pitch = use_tc(plan, l, precision) ? l.cin_pad : l.cin;  // matches mel_pitch() for conv_pre
  for (int j = 0; j < l.ntaps; ++j) f.tap_off[j] = l.tap_off[j];
  f.a0 = in.a0; f.a1 = in.a1; f.a_fmt = a_fmt_of(precision);
  f.w = l.w_ffma; f.epi = epi;
  cudaError_t e = launch_conv_ffma(f, st, rag);
  if (e != cudaSuccess) return fail(HG_ECUDA, "conv_ffma launch (%s): %s", l.name.c_str(), cudaGetErrorString(e));
  if (g_prof) prof_mark(layer_index(plan, &l), make_rec(HG_PATH_CUDA_CORE));
  return HG_OK;
}

// ------------------------------------------------------------------------------------------------
// fused ResBlock1 pair (conv_pair_tc.cu): bf16 mode, C in {32, 64}
struct PairTiling {
  int ms = 0, slab_rows = 0, box_rows = 0, nboxes = 0, t_rows = 0, t_bufs = 1, stages = 0, r_out = 0;
  bool resident = false;
  size_t smem = 0;
};

static bool pair_fusable(const HgPlan* plan, const Layer& l1, const Layer& l2, int precision, PairTiling* t) {
  if (!plan->fuse_pairs || precision != HG_PREC_BF16 || plan->force_ffma) return false;
  if (l1.kind != L_CONV || l2.kind != L_CONV || !l1.tc || !l2.tc) return false;
  const int c = l1.cin;
  if ((c != 32 && c != 64) || l1.cout != c || l2.cin != c || l2.cout != c) return false;
  if (l1.k != l2.k || l2.dil != 1 || (l1.k & 1) == 0 || l1.k > kMaxTaps) return false;
  const int rowb = c * 2;
  t->ms = 128 / c;  // MS * N_T = 128 accumulator columns per buffer
  const int mt = t->ms * 128;
  t->r_out = mt - (l1.k - 1);
  if (t->r_out < 64) return false;
  const int need = mt + l1.dil * (l1.k - 1);
  t->nboxes = (need + 255) / 256;
  const int align_rows = 1024 / rowb * 2;  // keeps every buffer a multiple of 1024 bytes
  t->box_rows = (((need + t->nboxes - 1) / t->nboxes) + align_rows - 1) / align_rows * align_rows;
  if (t->box_rows > 256) return false;
  t->slab_rows = t->nboxes * t->box_rows;
  t->t_rows = (mt + l1.k - 1 + 15) / 16 * 16;
  const size_t kMaxSmem = 227 * 1024;
  const int all = 2 * l1.k;
  // preference order: resident weights + double-buffered xt, resident + single xt, streamed ring
  t->stages = 0;
  for (int tb : {2, 1}) {
    if (conv_pair_smem_bytes(c, t->slab_rows, t->t_rows, tb, all) <= kMaxSmem) {
      t->resident = true; t->stages = all; t->t_bufs = tb;
      break;
    }
  }
  if (!t->stages) {
    t->resident = false;
    t->t_bufs = 1;
    int s = 8;
    while (s >= 2 && conv_pair_smem_bytes(c, t->slab_rows, t->t_rows, 1, s) > kMaxSmem) --s;
    if (s < 2) return false;
    t->stages = s;
    if (conv_pair_smem_bytes(c, t->slab_rows, t->t_rows, 2, std::min(s, 4)) <= kMaxSmem && s >= 4) {
      t->t_bufs = 2;
      while (conv_pair_smem_bytes(c, t->slab_rows, t->t_rows, 2, t->stages) > kMaxSmem) --t->stages;
    }
  }
  t->smem = conv_pair_smem_bytes(c, t->slab_rows, t->t_rows, t->t_bufs, t->stages);
  return true;
}

static int run_pair(HgPlan* plan, const Layer& l1, const Layer& l2, const PairTiling& t, int B, int L,
                    const OperandBuf& in, EpiParams epi, float slope, cudaStream_t st) {
  const int c = l1.cin;
  epi.bias = l2.bias;
  epi.a_fmt = A_BF16;
  epi.out_batch_stride = static_cast<long long>(L) * c;
  epi.out_extent = static_cast<long long>(L) * c;
  epi.out_row_stride = c;
  epi.out_offset = 0;
  TcPairParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.L = L; p.r_out = t.r_out;
  p.tiles_per_item = (L + t.r_out - 1) / t.r_out;
  RaggedItems rag_store;
  p.total_work = ragged_fill(&p.rag, ragged_items(&rag_store, B, L, 0), B, L, t.r_out);
  p.k = l1.k; p.d1 = l1.dil;
  p.slab_rows = t.slab_rows; p.box_rows = t.box_rows; p.nboxes = t.nboxes; p.t_rows = t.t_rows; p.t_bufs = t.t_bufs;
  p.stages = t.stages; p.w_resident = t.resident ? 1 : 0;
  p.w1 = l1.w_hi; p.w2 = l2.w_hi; p.bias1 = l1.bias; p.slope = slope;
  p.epi = epi;
  if (!epi.res) return fail(HG_ESTATE, "internal: fused pair without a residual");
  CUtensorMap m, mr;
  int rc = make_operand_map(plan, in.a0, L, B, c, c, t.box_rows, &m);
  if (rc) return rc;
  if ((rc = make_f32_tile_map(plan, epi.res, L, B, c, &mr))) return rc;
  epi.res = nullptr;  // the kernel adds the residual from its TMA-loaded tile
  p.epi = epi;
  const int grid = std::min(p.total_work, plan->sm_count);
  static long long* dbg_buf = nullptr;
  const char* dbg_layer = getenv("HG_TC_DEBUG_TIMING");  // layer name (the pair's c2) to instrument (bring-up only)
  if (dbg_layer && l2.name == dbg_layer) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 256 * 16 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 256 * 16 * sizeof(long long), st);
    p.dbg = dbg_buf;
  }
  cudaError_t e = launch_conv_pair_tc(c, m, mr, p, t.smem, grid, st);
  if (p.dbg && e == cudaSuccess) {
    std::vector<long long> h(256 * 16);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), dbg_buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double s[16] = {0};
    for (int i = 0; i < grid; ++i)
      for (int j = 0; j < 16; ++j) s[j] += static_cast<double>(h[i * 16 + j]) / grid;
    fprintf(stderr,
            "[hg dbg] %s pair k=%d d=%d grid=%d tiles/cta=%.1f resident=%d stages=%d t_bufs=%d | MMA warp: total=%.0f wait weights=%.0f d1_empty=%.0f "
            "slab=%.0f t_full=%.0f d2_empty=%.0f | epi warp 0: total=%.0f wait d1_full=%.0f t_empty=%.0f d2_full=%.0f res=%.0f; busy E1=%.0f E2=%.0f\n",
            l2.name.c_str(), p.k, p.d1, grid, s[6], p.w_resident, p.stages, p.t_bufs, s[0], s[1], s[2], s[3], s[4], s[5], s[8], s[9], s[10],
            s[11], s[12], s[13], s[14]);
  }
  if (e != cudaSuccess) return fail(HG_ECUDA, "conv_pair_tc launch (%s): %s", l2.name.c_str(), cudaGetErrorString(e));
  if (g_prof) prof_mark(layer_index(plan, &l2), make_rec(HG_PATH_FUSED_PAIR, c, c, t.ms, t.stages, 2, t.resident, t.smem));
  return HG_OK;
}

// ------------------------------------------------------------------------------------------------
// the same pair with F = 128 / C time rows folded into N (conv_pair_fold.cu)
struct FoldTiling {
  int f = 0, c_half = 0, smin = 0, smax = 0, nb_slab = 0, slab_phase_bytes = 0, xt_phase_bytes = 0, delta = 0, r_out = 0, fdiv = 0;
  int t_bufs = 1, stages = 0;  // stages: weight blocks held in shared memory
  int ring_period = 0;         // streamed weights: period of the ring (kernel template parameter); 0 = resident
  bool resident = false;
  size_t smem = 0;
};

// pure geometry (no device state): also what hg_fold_info reports, so the CPU tests can replay the dataflow
static bool fold_geometry(int c, int k, int d1, int L, FoldTiling* t) {
  if (c != 16 && c != 32 && c != 64) return false;
  const int f = 128 / c;
  if ((k & 1) == 0 || k > kMaxTaps || k + f - 1 > kFoldMaxOps) return false;
  if (L < 1 || L % f != 0 || d1 < 1 || d1 > 16) return false;
  const int rowb = c * 2, align = 1024 / rowb;
  const int ch = (k - 1) / 2;             // taps each side
  const int s0 = (ch + f - 1) / f;        // block-group shifts reach -s0 .. smax
  t->f = f; t->c_half = ch; t->smin = -s0; t->smax = (f - 1 + ch) / f;
  const int nblk_t = (128 + d1 - 1) / d1;  // block groups one 128-row M tile spans
  t->nb_slab = nblk_t + t->smax - t->smin;
  if (t->nb_slab > 256) return false;
  t->slab_phase_bytes = (t->nb_slab * d1 + align - 1) / align * align * rowb;
  t->fdiv = f * d1;
  const int xv = (128 / d1) * d1 * f;                 // xt rows of a tile that are complete block groups
  t->delta = t->fdiv * ((s0 + d1 - 1) / d1);           // first kept output row, a multiple of fdiv >= f*s0
  t->r_out = (xv - ch - t->delta) / t->fdiv * t->fdiv;
  if (t->r_out < 64) return false;
  const int tau_max = t->fdiv * (127 / d1) + (f - 1) * d1 + 127 % d1;  // last xt row E1 writes
  const int idx_r = t->delta / f + 127 + t->smax;                      // last xt row (per phase) G2 reads
  const int xt_rows = std::max(tau_max / f, idx_r) + 1;
  t->xt_phase_bytes = (xt_rows + align - 1) / align * align * rowb;
  // preference order: resident weights (2k stages) with two xt buffers, resident with one, then a ring as deep
  // as fits (at least F + 2 stages: a group holds F blocks while the next ones arrive)
  const size_t kMaxSmem = 227 * 1024;
  t->stages = 0;
  t->ring_period = 0;
  if (conv_fold_has_kernel(c, k, 0)) {
    for (int tb : {2, 1}) {
      if (conv_fold_smem_bytes(c, t->slab_phase_bytes, t->xt_phase_bytes, tb, 2 * k) <= kMaxSmem) {
        t->resident = true; t->stages = 2 * k; t->t_bufs = tb;
        break;
      }
    }
  }
  if (!t->stages) {  // streamed: the longest ring period (a kernel template parameter) whose blocks fit
    for (int s : {k, 6, 5, 4}) {
      if (!conv_fold_has_kernel(c, k, s)) continue;
      const int slots = conv_fold_weight_slots(k, s);
      if (conv_fold_smem_bytes(c, t->slab_phase_bytes, t->xt_phase_bytes, 1, slots) > kMaxSmem) continue;
      t->resident = false; t->t_bufs = 1; t->stages = slots; t->ring_period = s;
      break;
    }
  }
  if (!t->stages) return false;
  t->smem = conv_fold_smem_bytes(c, t->slab_phase_bytes, t->xt_phase_bytes, t->t_bufs, t->stages);
  return true;
}

// Where the time-folded kernel is the faster of the two pair kernels (interleaved A/B on B200, 16 x 800 frames,
// tools/pair_modes.py, profiles/r2_v11_pair_kernel_ab.txt): it wins where the N = C kernel is bound by MMA operand
// fetch — 32 channels from k = 5 up (0.43 -> 0.27 ms at k = 11) and 64 channels from k = 7 up (0.34 -> 0.31-0.34 ms
// at k = 7, 0.51 -> 0.39-0.51 ms at k = 11) — and loses a few percent where the pair is HBM-bound (k = 3).
// HG_FOLD=2 forces it wherever it applies (tests, A/B).
static bool fold_pays(const HgPlan* plan, const Layer& l2, bool mrf_accumulate) {
  (void)mrf_accumulate;
  if (plan->fold_force || l2.cin == 16) return true;  // 16 channels: the alternative is the CUDA-core kernel
  if (l2.cin == 32) return l2.k >= 5;
  return l2.k >= 7;
}

static bool fold_fusable(const HgPlan* plan, const Layer& l1, const Layer& l2, int precision, int L, FoldTiling* t) {
  if (!plan->fuse_pairs || !plan->fold_pairs || precision != HG_PREC_BF16 || plan->force_ffma) return false;
  if (l1.kind != L_CONV || l2.kind != L_CONV || !l1.w_hi || !l2.w_hi || !l2.bias_fold) return false;
  const int c = l1.cin;
  if (!(l1.tc && l2.tc) && c != 16) return false;
  if (l1.cout != c || l2.cin != c || l2.cout != c || l1.k != l2.k || l2.dil != 1) return false;
  return fold_geometry(c, l1.k, l1.dil, L, t);
}

// MMA groups of one conv of the folded pair.  Phase h of output M row i is time row (relative to the row
// the M tile starts at) F*d*blk + h*d + r; tap j reads input row that + (j - ch)*d, i.e. phase
// u = h + j - ch of the de-interleaved operand.  For every u the phases that use it form a run, and their
// taps j = u + ch - h a run of consecutive taps: one MMA group.  Accumulator column block q = phase F-1-q,
// so the run ascends in tap index.  Groups run in ascending u: every output element then accumulates its
// taps in the order 0..k-1 whatever its phase (position-independent, bit-identical to conv_pair_tc.cu), and
// the streamed weight blocks are consumed in ascending order.  The first groups cover only some phases,
// so the kernel clears the accumulator with a zero-operand MMA instead of an overwrite flag.
static int fold_schedule(int c, int f, int k, int row_bytes_shift16, int phase_bytes, int first_row, int rows_per_shift,
                         FoldOp* ops) {
  const int ch = (k - 1) / 2;
  int n = 0;
  for (int u = -ch; u <= f - 1 + ch; ++u) {
    const int s = u >= 0 ? u / f : -((-u + f - 1) / f);  // floor(u / f)
    const int hp = u - s * f;
    const int h_lo = std::max(0, u + ch - (k - 1)), h_hi = std::min(f - 1, u + ch);
    FoldOp& op = ops[n++];
    op.a_off16 = (hp * phase_bytes >> 4) + (first_row + s * rows_per_shift) * row_bytes_shift16;
    op.b_blk = u + ch - h_hi;
    op.nblk = h_hi - h_lo + 1;
    op.d_col = (f - 1 - h_hi) * c;
    // tap j has its last use (phase f-1) in the group u = j - ch + f - 1
    op.rel = (u + ch - f + 1 >= 0 && u + ch - f + 1 <= k - 1) ? 1 : 0;
  }
  return n;
}

static int run_pair_fold(HgPlan* plan, const Layer& l1, const Layer& l2, const FoldTiling& t, int B, int L,
                         const OperandBuf& in, EpiParams epi, float slope, cudaStream_t st) {
  const int c = l1.cin, f = t.f, rowb = c * 2;
  epi.bias = l2.bias_fold;
  epi.a_fmt = A_BF16;
  epi.out_batch_stride = static_cast<long long>(L) * c;
  epi.out_extent = static_cast<long long>(L) * c;
  epi.out_row_stride = 128;  // the folded view [L/F][128]
  epi.out_offset = 0;
  TcFoldParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.L = L; p.r_out = t.r_out;
  p.tiles_per_item = (L + t.r_out - 1) / t.r_out;
  RaggedItems rag_store;
  p.total_work = ragged_fill(&p.rag, ragged_items(&rag_store, B, L, 0), B, L, t.r_out);
  p.d1 = l1.dil;
  p.delta = t.delta; p.fdiv = t.fdiv; p.blk_off = t.smin;
  p.nblk_item = (L + t.fdiv - 1) / t.fdiv;
  p.nb_slab = t.nb_slab;
  p.slab_phase_bytes = t.slab_phase_bytes; p.xt_phase_bytes = t.xt_phase_bytes;
  p.a1_row0 = -t.smin * l1.dil;  // conv 1: slab row 0 = block group (origin / fdiv + smin)
  p.a2_row0 = t.delta / f;       // conv 2: M row i is output row origin + delta + F*i (+ phase)
  p.t_bufs = t.t_bufs;
  p.w1 = l1.w_hi; p.w2 = l2.w_hi; p.bias1 = l1.bias; p.slope = slope;
  if (!epi.res) return fail(HG_ESTATE, "internal: fused pair without a residual");
  (void)rowb;
  p.stages = t.stages;
  p.dbg = env_int("HG_FOLD_DBG", 0);
  CUtensorMap m, mr;
  int rc = make_fold_slab_map(plan, in.a0, L, B, c, l1.dil, t.nb_slab, &m);
  if (rc) return rc;
  if ((rc = make_f32_tile_map(plan, epi.res, L / f, B, 128, &mr))) return rc;
  epi.res = nullptr;  // the kernel adds the residual from its TMA-loaded tile
  p.epi = epi;
  const int grid = std::min(p.total_work, plan->sm_count);
  cudaError_t e = launch_conv_pair_fold(c, l1.k, t.ring_period, m, mr, p, t.smem, grid, st);
  if (e != cudaSuccess) return fail(HG_ECUDA, "conv_pair_fold launch (%s): %s", l2.name.c_str(), cudaGetErrorString(e));
  if (g_prof) prof_mark(layer_index(plan, &l2), make_rec(HG_PATH_FUSED_PAIR, 128, c, 1, t.stages, 2, t.resident, t.smem));
  return HG_OK;
}

// ------------------------------------------------------------------------------------------------
// a whole ResBlock1 in one launch (conv_chain_tc.cu): C in {32, 64}, every conv k = 3 (any small odd k whose reach
// fits the guard rows), bf16 mode
struct ChainTiling {
  int ms = 0, buf_rows = 0, box_rows = 0, nboxes = 0, r_out = 0, halo = 0, stages = 0, np = 0;
  bool resident = false;
  size_t smem = 0;
};

// c1 = first conv-1 layer of the block, c2 = first conv-2 layer; np pairs
static bool chain_fusable(const HgPlan* plan, const Layer* c1, const Layer* c2, int np, int precision, ChainTiling* t) {
  if (!plan->fuse_pairs || !plan->fuse_blocks || precision != HG_PREC_BF16 || plan->force_ffma) return false;
  if (np < 2 || np > kChainMaxPairs) return false;
  const int c = c1[0].cin, k = c1[0].k;
  if ((c != 32 && c != 64) || (k & 1) == 0 || k > kMaxTaps) return false;
  const int h2 = (k - 1) / 2, guard = conv_chain_guard_rows();
  int halo = 0;
  for (int m = 0; m < np; ++m) {
    const Layer& a = c1[m];
    const Layer& b = c2[m];
    if (a.kind != L_CONV || b.kind != L_CONV || !a.tc || !b.tc) return false;
    if (a.cin != c || a.cout != c || b.cin != c || b.cout != c || a.k != k || b.k != k || b.dil != 1) return false;
    if (a.dil * h2 > guard) return false;
    halo += (a.dil + 1) * h2;
  }
  t->np = np;
  t->ms = 128 / c;
  const int mt = t->ms * 128;
  t->halo = halo;
  t->r_out = mt - 2 * halo;
  if (t->r_out * 4 < mt * 3) return false;  // keep at least 3/4 of every tile (k = 3: 232 of 256, 488 of 512)
  const int rowb = c * 2, need = mt + 2 * guard;
  t->nboxes = (need + 255) / 256;
  const int align_rows = 1024 / rowb;
  t->box_rows = (((need + t->nboxes - 1) / t->nboxes) + align_rows - 1) / align_rows * align_rows;
  if (t->box_rows > 256) return false;
  t->buf_rows = t->nboxes * t->box_rows;
  const size_t kMaxSmem = 227 * 1024;
  const int all = 2 * np * k;
  if (conv_chain_smem_bytes(c, t->buf_rows, all) <= kMaxSmem) {
    t->resident = true; t->stages = all;
  } else {
    t->resident = false;
    int s = std::min(all - 1, 12);
    while (s >= 3 && conv_chain_smem_bytes(c, t->buf_rows, s) > kMaxSmem) --s;
    if (s < 3) return false;
    t->stages = s;
  }
  t->smem = conv_chain_smem_bytes(c, t->buf_rows, t->stages);
  return true;
}

static int run_chain(HgPlan* plan, const Layer* c1, const Layer* c2, const ChainTiling& t, int B, int L, const OperandBuf& in,
                     EpiParams epi, float slope, cudaStream_t st) {
  const int c = c1[0].cin;
  const Layer& last = c2[t.np - 1];
  epi.bias = last.bias;
  epi.a_fmt = A_BF16;
  epi.out_batch_stride = static_cast<long long>(L) * c;
  epi.out_extent = static_cast<long long>(L) * c;
  epi.out_row_stride = c;
  epi.out_offset = 0;
  TcChainParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.L = L; p.r_out = t.r_out;
  p.tiles_per_item = (L + t.r_out - 1) / t.r_out;
  RaggedItems rag_store;
  p.total_work = ragged_fill(&p.rag, ragged_items(&rag_store, B, L, 0), B, L, t.r_out);
  p.k = c1[0].k; p.np = t.np; p.halo = t.halo;
  p.buf_rows = t.buf_rows; p.box_rows = t.box_rows; p.nboxes = t.nboxes;
  p.stages = t.stages; p.w_resident = t.resident ? 1 : 0;
  for (int m = 0; m < t.np; ++m) {
    p.d1[m] = c1[m].dil;
    p.w1[m] = c1[m].w_hi; p.w2[m] = c2[m].w_hi;
    p.bias1[m] = c1[m].bias; p.bias2[m] = c2[m].bias;
  }
  p.slope = slope;
  if (!epi.res) return fail(HG_ESTATE, "internal: fused ResBlock without a residual");
  CUtensorMap m, mr;
  int rc = make_operand_map(plan, in.a0, L, B, c, c, t.box_rows, &m);
  if (rc) return rc;
  if ((rc = make_f32_tile_map(plan, epi.res, L, B, c, &mr))) return rc;
  epi.res = nullptr;  // the kernel keeps the residual in its TMA-loaded tile
  p.epi = epi;
  const int grid = std::min(p.total_work, plan->sm_count);
  cudaError_t e = launch_conv_chain_tc(c, m, mr, p, t.smem, grid, st);
  if (e != cudaSuccess) return fail(HG_ECUDA, "conv_chain_tc launch (%s): %s", last.name.c_str(), cudaGetErrorString(e));
  if (g_prof) prof_mark(layer_index(plan, &last), make_rec(HG_PATH_FUSED_BLOCK, c, c, t.ms, t.stages, 2, t.resident, t.smem));
  return HG_OK;
}

// ------------------------------------------------------------------------------------------------
// workspace
struct Workspace {
  float* F[3];
  OperandBuf A[3];
  OperandBuf mel;
  // short inputs only (concurrent_eligible): private x / operand buffers of ResBlocks 1..K-1 of a stage
  float* FB[kMaxSideStreams];
  OperandBuf AB[kMaxSideStreams][2];
  size_t bytes;
};

// Short inputs leave most of the GPU idle in every launch (one 256-frame utterance: 8 .. 132 tiles for 148 SMs) and
// the forward is a chain of ~11 us launches.  The K ResBlocks of a stage only meet in the MRF sum, so they run on
// separate streams there: the chain is one block long instead of K.  Decided per stage on its activation size
// (items x rows x channels; a stage of one 256-frame V1 utterance has 0.5 / 2.1 / 2.1 / 2.1 M elements): launches
// that fill the GPU by themselves only get in each other's way (tools/concurrent_blocks_ab.py).
static bool concurrent_stage(const HgPlan* plan, long long elems) {
  return plan->concurrent_elems > 0 && plan->cfg.num_kernels > 1 && plan->cfg.num_kernels - 1 <= kMaxSideStreams &&
         elems <= plan->concurrent_elems;
}
// whether any stage of this forward may run that way (then the workspace holds the blocks' private buffers)
static bool concurrent_eligible(const HgPlan* plan, int B, int T, int precision) {
  const std::vector<Layer>& LY = active_layers(plan, precision);
  long long L = T;
  for (int i = 0; i < plan->cfg.num_upsamples; ++i) {
    const Layer& up = LY[1 + i];
    L = (L - 1) * up.stride - 2 * up.pad + up.k;
    if (L >= 1 && concurrent_stage(plan, static_cast<long long>(B) * L * up.cout)) return true;
  }
  return false;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// channel pitch of the mel operand: padded only when conv_pre runs on the tensor-core path
static int mel_pitch(const HgPlan* plan, int precision) {
  const Layer& pre = active_layers(plan, precision)[0];
  return use_tc(plan, pre, precision) ? pre.cin_pad : pre.cin;
}

static int layout_workspace(const HgPlan* plan, int B, int T, int precision, void* base, Workspace* ws) {
  const HgConfig& c = plan->cfg;
  long long L = T;
  long long max_e = static_cast<long long>(B) * T * c.upsample_initial_channel;
  for (int i = 0; i < c.num_upsamples; ++i) {
    const Layer& up = active_layers(plan, precision)[1 + i];
    L = (L - 1) * up.stride - 2 * up.pad + up.k;
    if (L < 1) return fail(HG_EINVAL, "input too short for upsampler %d", i);
    max_e = std::max(max_e, static_cast<long long>(B) * L * up.cout);
  }
  const size_t fb = align_up(static_cast<size_t>(max_e) * 4, 1024);
  const size_t plane = align_up(static_cast<size_t>(max_e) * 2, 1024);
  const size_t ab = precision == HG_PREC_BF16 ? plane : precision == HG_PREC_FP32 ? 2 * plane : fb;
  const size_t mel_e = static_cast<size_t>(B) * T * mel_pitch(plan, precision);
  const size_t mplane = align_up(mel_e * 2, 1024);
  const size_t mb = precision == HG_PREC_BF16 ? mplane : precision == HG_PREC_FP32 ? 2 * mplane : align_up(mel_e * 4, 1024);
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  for (int i = 0; i < 3; ++i) { ws->F[i] = reinterpret_cast<float*>(p + off); off += fb; }
  for (int i = 0; i < 3; ++i) {
    ws->A[i].a0 = p + off;
    ws->A[i].a1 = precision == HG_PREC_FP32 ? p + off + plane : nullptr;
    off += ab;
  }
  ws->mel.a0 = p + off;
  ws->mel.a1 = precision == HG_PREC_FP32 ? p + off + mplane : nullptr;
  off += mb;
  if (concurrent_eligible(plan, B, T, precision)) {
    for (int j = 0; j + 1 < c.num_kernels; ++j) {
      ws->FB[j] = reinterpret_cast<float*>(p + off); off += fb;
      for (int i = 0; i < 2; ++i) {
        ws->AB[j][i].a0 = p + off;
        ws->AB[j][i].a1 = precision == HG_PREC_FP32 ? p + off + plane : nullptr;
        off += ab;
      }
    }
  }
  ws->bytes = off + 4096;  // the folded pair kernel's input map may read (and discard) a few rows past an operand buffer
  return HG_OK;
}

static int check_fwd_args(const HgPlan* plan, int B, int T, int precision) {
  if (!plan) return fail(HG_EINVAL, "null plan");
  if (plan->is_stack) return fail(HG_ESTATE, "this plan is a conv stack (hg_stack_create): use hg_stack_forward");
  if (!plan->finalized) return fail(HG_ESTATE, "plan not finalized");
  if (B < 1 || T < 1) return fail(HG_EINVAL, "expected B >= 1 and T >= 1, got B=%d T=%d", B, T);
  if (precision < HG_PREC_BF16 || precision > HG_PREC_FP32_FFMA) return fail(HG_EINVAL, "unknown precision %d", precision);
  long long hop = 1;
  for (int i = 0; i < plan->cfg.num_upsamples; ++i) hop *= plan->cfg.upsample_rates[i];
  if (static_cast<long long>(T) * hop > 0x7fffffffLL) return fail(HG_EINVAL, "T*hop exceeds 2^31-1 samples");
  return HG_OK;
}

extern "C" int hg_workspace_bytes(const HgPlan* plan, int B, int T, int precision, size_t* bytes) {
  int rc = check_fwd_args(plan, B, T, precision);
  if (rc) return rc;
  if (!bytes) return fail(HG_EINVAL, "null bytes");
  Workspace ws;
  if ((rc = layout_workspace(plan, B, T, precision, nullptr, &ws))) return rc;
  *bytes = ws.bytes;
  return HG_OK;
}

extern "C" int hg_forward_launches(const HgPlan* plan, int B, int T, int precision, int* launches) {
  int rc = check_fwd_args(plan, B, T, precision);
  if (rc) return rc;
  if (!launches) return fail(HG_EINVAL, "null launches");
  const std::vector<Layer>& LY = active_layers(plan, precision);
  int n = static_cast<int>(LY.size()) + 1;  // every layer + the mel repack
  if (plan->cfg.resblock_type == 1) {
    const int D = 3, U = plan->cfg.num_upsamples, K = plan->cfg.num_kernels;
    int li = 1 + U;
    for (int i = 0; i < U * K; ++i, li += 2 * D) {
      ChainTiling ct;
      if (chain_fusable(plan, &LY[li], &LY[li + D], D, precision, &ct)) { n -= 2 * D - 1; continue; }
      for (int m = 0; m < D; ++m) {
        PairTiling pt;
        FoldTiling ft;
        int Ls = T;  // sequence length of stage i / K
        for (int s = 0; s <= i / K; ++s) Ls *= plan->cfg.upsample_rates[s];
        if (pair_fusable(plan, LY[li + m], LY[li + D + m], precision, &pt) ||
            fold_fusable(plan, LY[li + m], LY[li + D + m], precision, Ls, &ft))
          --n;
      }
    }
  }
  *launches = n;
  return HG_OK;
}

// ------------------------------------------------------------------------------------------------
// Generator.forward — hifi/models.py:185-201
static int ensure_side_streams(HgPlan* plan, int n) {
  if (!plan->ev_fork && cudaEventCreateWithFlags(&plan->ev_fork, cudaEventDisableTiming) != cudaSuccess)
    return fail(HG_ECUDA, "cudaEventCreate: %s", cudaGetErrorString(cudaGetLastError()));
  for (int j = 0; j <= n; ++j)
    if (!plan->ev_done[j] && cudaEventCreateWithFlags(&plan->ev_done[j], cudaEventDisableTiming) != cudaSuccess)
      return fail(HG_ECUDA, "cudaEventCreate: %s", cudaGetErrorString(cudaGetLastError()));
  for (int j = 0; j < n; ++j)
    if (!plan->side_stream[j] && cudaStreamCreateWithFlags(&plan->side_stream[j], cudaStreamNonBlocking) != cudaSuccess)
      return fail(HG_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(cudaGetLastError()));
  return HG_OK;
}

extern "C" int hg_forward(HgPlan* plan, const float* mel, int64_t sB, int64_t sC, int64_t sT, int B, int T, void* out,
                          int out_dtype, float out_scale, int precision, void* workspace, size_t workspace_bytes,
                          void* stream) {
  int rc = check_fwd_args(plan, B, T, precision);
  if (rc) return rc;
  if (!mel || !out || !workspace) return fail(HG_EINVAL, "null buffer");
  if (out_dtype != HG_OUT_F32 && out_dtype != HG_OUT_I16) return fail(HG_EINVAL, "unknown out_dtype %d", out_dtype);
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(HG_EINVAL, "workspace must be 1024-byte aligned");
  Workspace ws;
  if ((rc = layout_workspace(plan, B, T, precision, workspace, &ws))) return rc;
  if (workspace_bytes < ws.bytes)
    return fail(HG_ENOMEM, "workspace too small: %zu < %zu bytes", workspace_bytes, ws.bytes);
  DEVICE_SCOPE(plan->device);
  TileOrderScope tile_order(plan->tile_alternate);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const HgConfig& c = plan->cfg;
  const std::vector<Layer>& LY = active_layers(plan, precision);
  const int fmt = a_fmt_of(precision);
  const int U = c.num_upsamples, K = c.num_kernels, D = rb_dilations(c);
  const float slope = 0.1f;  // LRELU_SLOPE, hifi/models.py:9
  // short inputs: the ResBlocks of a stage on separate streams (concurrent_eligible above); per-launch profiling and
  // the whole-block kernel keep the single-stream schedule
  const bool conc_any = concurrent_eligible(plan, B, T, precision) && !g_prof && !plan->fuse_blocks;
  std::unique_lock<std::mutex> side_lock(plan->side_mutex, std::defer_lock);
  if (conc_any) {
    side_lock.lock();
    if ((rc = ensure_side_streams(plan, K - 1))) return rc;
  }
  auto cuda_ok = [&](cudaError_t ce, const char* what) { return ce == cudaSuccess ? HG_OK : fail(HG_ECUDA, "%s: %s", what, cudaGetErrorString(ce)); };

  prof_mark(-2);
  // mel [B,80,T] (any strides) -> channels-last operand
  cudaError_t e = launch_mel_to_operand(mel, sB, sC, sT, B, c.num_mels, T, mel_pitch(plan, precision), fmt, ws.mel.a0,
                                        ws.mel.a1, st);
  if (e != cudaSuccess) return fail(HG_ECUDA, "mel_to_operand: %s", cudaGetErrorString(e));
  prof_mark(-1, make_rec(HG_PATH_REPACK));

  // x = conv_pre(x)  :186 ; only leaky_relu(x) is consumed (by ups[0], :188-189)
  int a_cur = 1;  // operand buffer holding leaky_relu(current x)
  {
    EpiParams ep; memset(&ep, 0, sizeof(ep));
    ep.out_a0 = ws.A[a_cur].a0; ep.out_a1 = ws.A[a_cur].a1; ep.slope = slope;
    if ((rc = run_layer(plan, LY[0], precision, B, T, ws.mel, ep, st))) return rc;
  }
  int L = T;
  int li = 1 + U;  // first resblock layer
  float* x_final = nullptr;
  for (int i = 0; i < U; ++i) {
    const Layer& up = LY[1 + i];
    const bool last_stage = i == U - 1;
    // x = ups[i](leaky_relu(x))  :188-189  -> residual stream F0 and operand A0
    {
      EpiParams ep; memset(&ep, 0, sizeof(ep));
      ep.out_x = ws.F[0]; ep.out_a0 = ws.A[0].a0; ep.out_a1 = ws.A[0].a1; ep.slope = slope;
      if ((rc = run_layer(plan, up, precision, B, L, ws.A[a_cur], ep, st))) return rc;
    }
    L = (L - 1) * up.stride - 2 * up.pad + up.k;
    const bool conc = conc_any && concurrent_stage(plan, static_cast<long long>(B) * L * up.cout);
    if (conc) {  // fork: every block starts from the upsampler's output
      if ((rc = cuda_ok(cudaEventRecord(plan->ev_fork, st), "cudaEventRecord"))) return rc;
      for (int j = 1; j < K; ++j)
        if ((rc = cuda_ok(cudaStreamWaitEvent(plan->side_stream[j - 1], plan->ev_fork, 0), "cudaStreamWaitEvent"))) return rc;
    }
    // xs = sum_j resblocks[i*K+j](x) ; x = xs / K   :190-196
    for (int j = 0; j < K; ++j) {
      // block j's stream and private buffers (index 0 = the shared stage input); single-stream schedule: everything shared
      cudaStream_t sj = (conc && j > 0) ? plan->side_stream[j - 1] : st;
      float* const Fj = (conc && j > 0) ? ws.FB[j - 1] : ws.F[1];
      const OperandBuf Aj[3] = {ws.A[0], (conc && j > 0) ? ws.AB[j - 1][0] : ws.A[1], (conc && j > 0) ? ws.AB[j - 1][1] : ws.A[2]};
      ChainTiling ct;
      if (c.resblock_type == 1 && chain_fusable(plan, &LY[li], &LY[li + D], D, precision, &ct)) {
        // the whole ResBlock in one launch (conv_chain_tc.cu): x and the operand stay on chip between its pairs;
        // its last epilogue is the last pair's, MRF combine included  :88-95, :193-196
        EpiParams ep; memset(&ep, 0, sizeof(ep));
        ep.res = ws.F[0]; ep.slope = slope;
        if (j > 0) ep.acc_in = ws.F[2];
        if (j < K - 1) {
          ep.out_x = ws.F[2];
        } else {
          if (K > 1) ep.post_div = static_cast<float>(K);
          if (last_stage) {
            ep.out_x = ws.F[1];
            x_final = ws.F[1];
          } else {
            ep.out_a0 = ws.A[1].a0; ep.out_a1 = ws.A[1].a1;
            a_cur = 1;
          }
        }
        if ((rc = run_chain(plan, &LY[li], &LY[li + D], ct, B, L, ws.A[0], ep, slope, st))) return rc;
        li += 2 * D;
        continue;
      }
      int a_in = 0;          // operand of the block's running x (A0 = the shared stage input)
      const float* res = ws.F[0];
      for (int m = 0; m < D; ++m) {
        const bool last_pair = m == D - 1;
        int a_conv_in = a_in;
        const Layer& l2 = LY[c.resblock_type == 1 ? li + D + m : li + m];
        PairTiling pt;
        FoldTiling ft;
        const bool pair_ok = c.resblock_type == 1 && pair_fusable(plan, LY[li + m], l2, precision, &pt);
        const bool fold_ok = c.resblock_type == 1 && fold_fusable(plan, LY[li + m], l2, precision, L, &ft);
        const bool fused = pair_ok || fold_ok;
        if (c.resblock_type == 1 && !fused) {
          // xt = c1(leaky_relu(x)); only leaky_relu(xt) is consumed  :90-92
          const int a_t = 2;
          EpiParams ep; memset(&ep, 0, sizeof(ep));
          ep.out_a0 = Aj[a_t].a0; ep.out_a1 = Aj[a_t].a1; ep.slope = slope;
          if ((rc = run_layer(plan, LY[li + m], precision, B, L, Aj[a_in], ep, sj))) return rc;
          a_conv_in = a_t;
        }
        // x = c2(xt) + x  :93-94 (ResBlock2: x = c(leaky_relu(x)) + x  :136-138)
        EpiParams ep; memset(&ep, 0, sizeof(ep));
        ep.res = res; ep.slope = slope;
        // next operand buffer must differ from the one this conv reads (neighbouring CTAs read halos)
        const int a_next = (a_conv_in == 1) ? 2 : 1;
        if (!last_pair) {
          ep.out_x = Fj; ep.out_a0 = Aj[a_next].a0; ep.out_a1 = Aj[a_next].a1;
          a_in = a_next;
          res = Fj;
        } else {
          // MRF combine fused into the block's last epilogue  :193-196
          if (j > 0) ep.acc_in = ws.F[2];
          if (j < K - 1) {
            ep.out_x = ws.F[2];
          } else {
            if (K > 1) ep.post_div = static_cast<float>(K);
            if (last_stage) {
              ep.out_x = ws.F[1];  // conv_post applies its own leaky_relu(0.01)  :197
              x_final = ws.F[1];
            } else {
              // concurrent schedule: the last block reads its private buffers, so the stage output can always go to A[1]
              const int a_out = conc ? 1 : a_next;
              ep.out_a0 = ws.A[a_out].a0; ep.out_a1 = ws.A[a_out].a1;
              a_cur = a_out;
            }
          }
          // join the MRF chain: this epilogue reads block j-1's sum (and, for the last block, writes buffers block 0 used)
          if (conc && j > 0 && (rc = cuda_ok(cudaStreamWaitEvent(sj, plan->ev_done[j - 1], 0), "cudaStreamWaitEvent"))) return rc;
        }
        if (fold_ok && (!pair_ok || fold_pays(plan, l2, ep.acc_in != nullptr))) {
          // ... with F = 128 / C time rows folded into the MMA's N dimension  (conv_pair_fold.cu)
          if ((rc = run_pair_fold(plan, LY[li + m], l2, ft, B, L, Aj[a_conv_in], ep, slope, sj))) return rc;
        } else if (fused) {
          // c1 and c2 in one kernel; xt never leaves shared memory  (conv_pair_tc.cu)
          if ((rc = run_pair(plan, LY[li + m], l2, pt, B, L, Aj[a_conv_in], ep, slope, sj))) return rc;
        } else if ((rc = run_layer(plan, l2, precision, B, L, Aj[a_conv_in], ep, sj))) {
          return rc;
        }
      }
      li += c.resblock_type == 1 ? 2 * D : D;
      if (conc && (rc = cuda_ok(cudaEventRecord(plan->ev_done[j], sj), "cudaEventRecord"))) return rc;
    }
    if (conc && (rc = cuda_ok(cudaStreamWaitEvent(st, plan->ev_done[K - 1], 0), "cudaStreamWaitEvent"))) return rc;  // join
  }
  // x = tanh(conv_post(leaky_relu(x)))  :197-199  (+ optional int16 tail, hifiapi.py:50-51)
  const Layer& post = LY.back();
  RaggedItems rag_store;
  e = launch_conv_post(x_final, B, L, post.cin, post.w_post, post.bias_post,
                       out_dtype == HG_OUT_F32 ? static_cast<float*>(out) : nullptr,
                       out_dtype == HG_OUT_I16 ? static_cast<int16_t*>(out) : nullptr, out_scale, st,
                       post.w_post_host.empty() ? nullptr : post.w_post_host.data(),
                       ragged_items(&rag_store, B, L, 0, /*with_halo=*/false),  // the last layer feeds nobody
                       g_win ? g_win->item_stride : 0, g_win ? g_win->skip : 0, g_win ? g_win->keep : 0);
  if (e != cudaSuccess) return fail(HG_ECUDA, "conv_post: %s", cudaGetErrorString(e));
  prof_mark(static_cast<int>(LY.size()) - 1, make_rec(HG_PATH_POST));
  return HG_OK;
}

extern "C" int hg_forward_window(HgPlan* plan, const float* mel, int64_t sB, int64_t sC, int64_t sT, int B, int T, void* out,
                                 int64_t out_item_stride, int64_t skip_samples, int64_t keep_samples, int out_dtype,
                                 float out_scale, int precision, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_fwd_args(plan, B, T, precision);
  if (rc) return rc;
  long long hop = 1;
  for (int i = 0; i < plan->cfg.num_upsamples; ++i) hop *= plan->cfg.upsample_rates[i];
  const long long n = static_cast<long long>(T) * hop;
  if (skip_samples < 0 || keep_samples < 1 || skip_samples + keep_samples > n)
    return fail(HG_EINVAL, "window [%lld, %lld) outside the %lld samples of the forward", static_cast<long long>(skip_samples),
                static_cast<long long>(skip_samples + keep_samples), n);
  if (out_item_stride < keep_samples) return fail(HG_EINVAL, "out_item_stride %lld < keep_samples %lld", static_cast<long long>(out_item_stride), static_cast<long long>(keep_samples));
  WindowCtx ctx{out_item_stride, static_cast<int>(skip_samples), static_cast<int>(keep_samples)};
  g_win = &ctx;
  rc = hg_forward(plan, mel, sB, sC, sT, B, T, out, out_dtype, out_scale, precision, workspace, workspace_bytes, stream);
  g_win = nullptr;
  return rc;
}

extern "C" int hg_enable_peer_access(int device, int peer_device) {
  int rc = check_device(device);
  if (rc) return rc;
  if (device == peer_device) return HG_OK;
  int can = 0;
  CUDA_TRY(cudaDeviceCanAccessPeer(&can, device, peer_device));
  if (!can) return fail(HG_ENODEVICE, "device %d cannot access device %d's memory (no NVLink / PCIe peer path)", device, peer_device);
  DEVICE_SCOPE(device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
  if (e != cudaSuccess) return fail(HG_ECUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", device, peer_device, cudaGetErrorString(e));
  return HG_OK;
}

// ------------------------------------------------------------------------------------------------
// one buffer, every GPU of the node: CUDA IPC export / import for the direct-store gather
extern "C" int hg_ipc_export(const void* device_ptr, unsigned char* handle64, int64_t* offset) {
  if (!device_ptr || !handle64 || !offset) return fail(HG_EINVAL, "null argument");
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static RangeFn range = nullptr;
  if (!range) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
      return fail(HG_ECUDA, "cuMemGetAddressRange entry point unavailable");
    range = reinterpret_cast<RangeFn>(p);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  if (range(&base, &size, reinterpret_cast<CUdeviceptr>(device_ptr)) != CUDA_SUCCESS)
    return fail(HG_EINVAL, "not a device allocation");
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CUDA_TRY(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  memcpy(handle64, &h, 64);
  *offset = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(device_ptr) - base);
  return HG_OK;
}

extern "C" int hg_ipc_import(int device, const unsigned char* handle64, int64_t offset, void** base, void** ptr) {
  if (!handle64 || !base || !ptr) return fail(HG_EINVAL, "null argument");
  int rc = check_device(device);
  if (rc) return rc;
  // opened with `device` current: the mapping is made for THIS device's kernels, with peer access to
  // the owner enabled lazily (a mapping opened under the owner's device is not reachable from here)
  DEVICE_SCOPE(device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* b = nullptr;
  CUDA_TRY(cudaIpcOpenMemHandle(&b, h, cudaIpcMemLazyEnablePeerAccess));
  *base = b;
  *ptr = static_cast<unsigned char*>(b) + offset;
  return HG_OK;
}

extern "C" int hg_ipc_close(int device, void* base) {
  if (!base) return HG_OK;
  int rc = check_device(device);
  if (rc) return rc;
  DEVICE_SCOPE(device);
  CUDA_TRY(cudaIpcCloseMemHandle(base));
  return HG_OK;
}

extern "C" int hg_halo_frames(const HgPlan* plan, int* frames) {
  if (!plan || !frames) return fail(HG_EINVAL, "null argument");
  *frames = halo_frames_of(plan->cfg);
  return HG_OK;
}

extern "C" int hg_forward_ragged(HgPlan* plan, const float* mel, int64_t sB, int64_t sC, int64_t sT, int B, int T,
                                 const int32_t* frames, void* out, int out_dtype, float out_scale, int precision,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!plan) return fail(HG_EINVAL, "null plan");
  if (!frames) return fail(HG_EINVAL, "null frames");
  if (B > kMaxRaggedItems) return fail(HG_EINVAL, "ragged batches hold at most %d items, got %d", kMaxRaggedItems, B);
  for (int b = 0; b < B; ++b)
    if (frames[b] < 1 || frames[b] > T) return fail(HG_EINVAL, "frames[%d] = %d outside [1, T = %d]", b, frames[b], T);
  RaggedCtx ctx{frames, halo_frames_of(plan->cfg), T};
  g_rag = &ctx;
  const int rc = hg_forward(plan, mel, sB, sC, sT, B, T, out, out_dtype, out_scale, precision, workspace, workspace_bytes, stream);
  g_rag = nullptr;
  return rc;
}

// ------------------------------------------------------------------------------------------------
// conv stacks: FastSpeech2's PostNet (fs_two/transformer/Layers.py:71-143, BatchNorm folded by the
// caller) and mel_linear (fs_two/model/fastspeech2.py:101-104) on the generator's conv kernels

extern "C" int hg_stack_create(const HgStackLayer* layers, int n_layers, int device, HgPlan** out) {
  if (!layers || !out) return fail(HG_EINVAL, "null argument");
  *out = nullptr;
  if (n_layers < 1 || n_layers > 64) return fail(HG_EINVAL, "a stack holds 1..64 layers, got %d", n_layers);
  for (int i = 0; i < n_layers; ++i) {
    const HgStackLayer& d = layers[i];
    if (d.c_in < 1 || d.c_out < 1 || d.k < 1 || d.k > kMaxTaps || (d.k & 1) == 0 || d.dilation < 1)
      return fail(HG_EINVAL, "stack layer %d: bad shape (odd k <= %d, dilation >= 1)", i, kMaxTaps);
    if (d.act < HG_ACT_NONE || d.act > HG_ACT_TANH) return fail(HG_EINVAL, "stack layer %d: unknown activation %d", i, d.act);
    if (i + 1 < n_layers && layers[i + 1].c_in != d.c_out)
      return fail(HG_EINVAL, "stack layer %d takes %d channels but layer %d produces %d", i + 1, layers[i + 1].c_in, i, d.c_out);
  }
  if (layers[n_layers - 1].act != HG_ACT_NONE) return fail(HG_EINVAL, "the last layer's output is returned as is: act must be HG_ACT_NONE");
  int rc = check_device(device);
  if (rc) return rc;
  HgPlan* p = new HgPlan();
  memset(&p->cfg, 0, sizeof(p->cfg));
  p->is_stack = true;
  init_plan_env(p, device);
  for (int i = 0; i < n_layers; ++i) {
    // an output width that is not a multiple of 32 (80 mel bins) has no tensor-core tiling: the LAST layer may run
    // with its N padded up on zero weights, the padded columns never stored (intermediate layers would change the
    // operand pitch of their consumer, so only the final 80-wide projection is treated this way)
    const int co = layers[i].c_out;
    const bool pad_n = i == n_layers - 1 && co % 32 != 0 && co >= 48 && co % 4 == 0 && env_int("HG_STACK_PAD_N", 1);
    Layer sl = make_conv(std::to_string(i), layers[i].c_in, pad_n ? (co + 31) / 32 * 32 : co, layers[i].k, layers[i].dilation);
    if (pad_n) { sl.cout_w = co; sl.cin_w = layers[i].c_in; sl.n_store = co; }
    p->layers.push_back(sl);
    p->stack_act.push_back(layers[i].act);
    p->stack_slope.push_back(layers[i].slope);
    p->by_name[p->layers.back().name] = i;
  }
  *out = p;
  return HG_OK;
}

struct StackWorkspace {
  OperandBuf A;
  float* F;
  size_t bytes;
};

static int in_pitch(const HgPlan* plan, const Layer& l, int precision) { return use_tc(plan, l, precision) ? l.cin_pad : l.cin; }

static void layout_stack(const HgPlan* plan, int B, int T, int precision, void* base, StackWorkspace* ws) {
  size_t max_a = 0, max_f = 0;
  for (const Layer& l : plan->layers) {
    max_a = std::max(max_a, static_cast<size_t>(B) * T * in_pitch(plan, l, precision));
    max_f = std::max(max_f, static_cast<size_t>(B) * T * (l.n_store ? l.n_store : l.cout));
  }
  const size_t plane = align_up(max_a * 2, 1024);
  const size_t ab = precision == HG_PREC_BF16 ? plane : precision == HG_PREC_FP32 ? 2 * plane : align_up(max_a * 4, 1024);
  uint8_t* p = static_cast<uint8_t*>(base);
  ws->A.a0 = p;
  ws->A.a1 = precision == HG_PREC_FP32 ? p + plane : nullptr;
  ws->F = reinterpret_cast<float*>(p + ab);
  ws->bytes = ab + align_up(max_f * 4, 1024);
}

static int check_stack_args(const HgPlan* plan, int B, int T, int precision) {
  if (!plan) return fail(HG_EINVAL, "null plan");
  if (!plan->is_stack) return fail(HG_ESTATE, "this plan is a generator (hg_plan_create): use hg_forward");
  if (!plan->finalized) return fail(HG_ESTATE, "plan not finalized");
  if (B < 1 || T < 1) return fail(HG_EINVAL, "expected B >= 1 and T >= 1, got B=%d T=%d", B, T);
  if (precision < HG_PREC_BF16 || precision > HG_PREC_FP32_FFMA) return fail(HG_EINVAL, "unknown precision %d", precision);
  return HG_OK;
}

extern "C" int hg_stack_workspace_bytes(const HgPlan* plan, int B, int T, int precision, size_t* bytes) {
  int rc = check_stack_args(plan, B, T, precision);
  if (rc) return rc;
  if (!bytes) return fail(HG_EINVAL, "null bytes");
  StackWorkspace ws;
  layout_stack(plan, B, T, precision, nullptr, &ws);
  *bytes = ws.bytes;
  return HG_OK;
}

extern "C" int hg_stack_forward(HgPlan* plan, const float* x, int64_t sB, int64_t sC, int64_t sT, int B, int T,
                                const float* residual, float* out, int precision, void* workspace, size_t workspace_bytes,
                                void* stream) {
  int rc = check_stack_args(plan, B, T, precision);
  if (rc) return rc;
  if (!x || !out || !workspace) return fail(HG_EINVAL, "null buffer");
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(HG_EINVAL, "workspace must be 1024-byte aligned");
  StackWorkspace ws;
  layout_stack(plan, B, T, precision, workspace, &ws);
  if (workspace_bytes < ws.bytes) return fail(HG_ENOMEM, "workspace too small: %zu < %zu bytes", workspace_bytes, ws.bytes);
  DEVICE_SCOPE(plan->device);
  TileOrderScope tile_order(plan->tile_alternate);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int fmt = a_fmt_of(precision);
  const int n = static_cast<int>(plan->layers.size());
  // x [B, C, T] (any strides; a time-major [B,T,C] tensor is sC = 1, sT = C) -> operand of layer 0
  cudaError_t e = launch_mel_to_operand(x, sB, sC, sT, B, plan->layers[0].cin, T, in_pitch(plan, plan->layers[0], precision), fmt,
                                        ws.A.a0, ws.A.a1, st);
  if (e != cudaSuccess) return fail(HG_ECUDA, "stack input repack: %s", cudaGetErrorString(e));
  for (int i = 0; i < n; ++i) {
    const Layer& l = plan->layers[i];
    const bool last = i == n - 1;
    EpiParams ep; memset(&ep, 0, sizeof(ep));
    ep.slope = 1.f;
    ep.out_x = last ? out : ws.F;
    ep.res = last ? residual : nullptr;
    if ((rc = run_layer(plan, l, precision, B, T, ws.A, ep, st))) return rc;
    if (!last) {
      // activation between the layers + operand format of the next one (a pass over B*T*C floats:
      // the stack is ~1 % of the vocoder's work, so this is not fused into the conv epilogues)
      const Layer& nx = plan->layers[i + 1];
      e = launch_mel_to_operand(ws.F, static_cast<long long>(T) * l.cout, 1, l.cout, B, l.cout, T, in_pitch(plan, nx, precision), fmt,
                                ws.A.a0, ws.A.a1, st, plan->stack_act[i], plan->stack_slope[i]);
      if (e != cudaSuccess) return fail(HG_ECUDA, "stack activation (layer %d): %s", i, cudaGetErrorString(e));
    }
  }
  return HG_OK;
}

extern "C" int hg_layer_count(const HgPlan* plan, int* count) {
  if (!plan || !count) return fail(HG_EINVAL, "null argument");
  *count = static_cast<int>(plan->layers.size());
  return HG_OK;
}

extern "C" int hg_layer_info(const HgPlan* plan, int index, int precision, HgLayerInfo* info) {
  if (!plan || !info) return fail(HG_EINVAL, "null argument");
  if (index < 0 || index >= static_cast<int>(plan->layers.size())) return fail(HG_EINVAL, "layer index out of range");
  if (precision < HG_PREC_BF16 || precision > HG_PREC_FP32_FFMA) return fail(HG_EINVAL, "unknown precision");
  const Layer& l = plan->layers[index];
  memset(info, 0, sizeof(*info));
  snprintf(info->name, sizeof(info->name), "%s", l.name.c_str());
  info->kind = l.kind; info->c_in = l.cin; info->c_out = l.cout; info->k = l.k;
  info->dilation = l.dil; info->stride = l.stride;
  info->tensor_core = use_tc(plan, l, precision) ? 1 : 0;
  if (info->tensor_core) {
    TcTiling t = choose_tiling(plan, l, precision == HG_PREC_FP32);
    info->tensor_core = t.stages >= 2 ? 1 : 0;
    info->n_tile = l.n_tile; info->k_chunk = l.kc; info->m_subtiles = t.ms; info->stages = t.stages;
    info->smem_bytes = static_cast<int32_t>(t.smem);
    info->weights_resident = t.resident ? 1 : 0; info->slab_buffers = t.nbuf;
  }
  return HG_OK;
}

extern "C" int hg_profile_forward(HgPlan* plan, const float* mel, int64_t sB, int64_t sC, int64_t sT, int B, int T,
                                  void* out, int out_dtype, float out_scale, int precision, void* workspace,
                                  size_t workspace_bytes, void* stream, int* layer_index, float* layer_ms,
                                  int max_launches, int* n_launches) {
  if (!layer_index || !layer_ms || !n_launches) return fail(HG_EINVAL, "null argument");
  ProfSink sink;
  sink.st = static_cast<cudaStream_t>(stream);
  g_prof = &sink;
  int rc = hg_forward(plan, mel, sB, sC, sT, B, T, out, out_dtype, out_scale, precision, workspace, workspace_bytes, stream);
  g_prof = nullptr;
  cudaError_t es = cudaStreamSynchronize(sink.st);
  int n = 0;
  g_last_profile.clear();
  for (size_t i = 0; i < sink.layer.size() && i < sink.rec.size(); ++i) g_last_profile.emplace_back(sink.layer[i], sink.rec[i]);
  if (!rc && es == cudaSuccess) {
    n = static_cast<int>(sink.layer.size());
    if (n > max_launches) n = max_launches;
    for (int i = 0; i < n; ++i) {
      layer_index[i] = sink.layer[i];
      float ms = 0.f;
      cudaEventElapsedTime(&ms, sink.ev[i], sink.ev[i + 1]);
      layer_ms[i] = ms;
    }
  }
  for (cudaEvent_t e : sink.ev) cudaEventDestroy(e);
  if (rc) return rc;
  if (es != cudaSuccess) return fail(HG_ECUDA, "profile_forward: %s", cudaGetErrorString(es));
  *n_launches = n;
  return HG_OK;
}

extern "C" int hg_profile_launch_info(const HgPlan* plan, int launch, HgLayerInfo* info) {
  if (!plan || !info) return fail(HG_EINVAL, "null argument");
  if (launch < 0 || launch >= static_cast<int>(g_last_profile.size()))
    return fail(HG_EINVAL, "launch %d outside the last hg_profile_forward of this thread (%zu launches)", launch, g_last_profile.size());
  const int li = g_last_profile[launch].first;
  const LaunchRec& r = g_last_profile[launch].second;
  memset(info, 0, sizeof(*info));
  if (li >= 0 && li < static_cast<int>(plan->layers.size())) {
    const Layer& l = plan->layers[li];
    snprintf(info->name, sizeof(info->name), "%s", l.name.c_str());
    info->kind = l.kind; info->c_in = l.cin; info->c_out = l.cout; info->k = l.k; info->dilation = l.dil; info->stride = l.stride;
  } else {
    snprintf(info->name, sizeof(info->name), "mel_to_operand");
    info->kind = -1;
  }
  info->kernel_path = r.path;
  info->tensor_core = (r.path == HG_PATH_TC || r.path == HG_PATH_TC_CTA_PAIR || r.path == HG_PATH_FUSED_PAIR ||
                       r.path == HG_PATH_FUSED_BLOCK) ? 1 : 0;
  info->n_tile = r.n_tile; info->k_chunk = r.kc; info->m_subtiles = r.ms; info->stages = r.stages;
  info->smem_bytes = static_cast<int32_t>(r.smem); info->weights_resident = r.resident; info->slab_buffers = r.nbuf;
  return HG_OK;
}

// ------------------------------------------------------------------------------------------------
// op-level entry points (parity tier T1)
static int op_layer(int device, int precision, Layer& l, const float* x, int B, int L, const float* weight,
                    const float* bias, float in_slope, const float* residual, float* y, void* stream) {
  int rc = check_device(device);
  if (rc) return rc;
  if (precision < HG_PREC_BF16 || precision > HG_PREC_FP32_FFMA) return fail(HG_EINVAL, "unknown precision");
  DEVICE_SCOPE(device);
  HgPlan plan;
  plan.device = device;
  plan.desc_mode = env_int("HG_DESC_MODE", 0);
  plan.force_ms = env_int("HG_TC_MS", 0);
  plan.force_stages = env_int("HG_TC_STAGES", 0);
  plan.force_ffma = env_int("HG_FORCE_FFMA", 0) != 0;
  plan.ctas_per_sm = env_int("HG_TC_CTAS_PER_SM", 1);
  plan.epi_tma = env_int("HG_EPI_TMA", 1) != 0;
  plan.use_tc2 = env_int("HG_TC2", 1) != 0;
  {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, device) == cudaSuccess) plan.sm_count = pr.multiProcessorCount;
  }
  if ((rc = pack_layer(l, weight, bias))) { free_layer(l); return rc; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int fmt = a_fmt_of(precision);
  const bool tc = use_tc(&plan, l, precision);
  const int pitch = tc ? l.cin_pad : l.cin;
  // pitch > C_in: the tensor-core path zero-pads K to a multiple of 64 (conv_pre: 80 mel bins -> 128)
  const long long n = static_cast<long long>(B) * L * pitch;
  void* a = nullptr;
  const size_t plane = align_up(static_cast<size_t>(n) * 2, 1024);
  cudaError_t e = cudaMalloc(&a, fmt == A_F32 ? static_cast<size_t>(n) * 4 : 2 * plane);
  if (e != cudaSuccess) { free_layer(l); return fail(HG_ECUDA, "cudaMalloc: %s", cudaGetErrorString(e)); }
  OperandBuf in;
  in.a0 = a;
  in.a1 = fmt == A_BF16_SPLIT ? static_cast<uint8_t*>(a) + plane : nullptr;
  if (pitch == l.cin)
    e = launch_f32_to_operand(x, n, in_slope, fmt, in.a0, in.a1, st);
  else  // the strided repack the generator uses for the mel: x is [B][L][C_in] channels-last
    e = launch_mel_to_operand(x, static_cast<long long>(L) * l.cin, 1, l.cin, B, l.cin, L, pitch, fmt, in.a0, in.a1, st,
                              in_slope == 1.f ? 0 : 1, in_slope);
  EpiParams ep; memset(&ep, 0, sizeof(ep));
  ep.res = residual; ep.out_x = y; ep.slope = 1.f;
  rc = e == cudaSuccess ? run_layer(&plan, l, precision, B, L, in, ep, st) : fail(HG_ECUDA, "f32_to_operand: %s", cudaGetErrorString(e));
  cudaError_t es = cudaStreamSynchronize(st);
  cudaFree(a);
  free_layer(l);
  if (rc) return rc;
  if (es != cudaSuccess) return fail(HG_ECUDA, "op execution failed: %s", cudaGetErrorString(es));
  return HG_OK;
}

extern "C" int hg_op_conv1d(int device, int precision, const float* x, int B, int L, int C_in, const float* weight,
                            const float* bias, int C_out, int k, int dilation, float in_slope, const float* residual,
                            float* y, void* stream) {
  if (!x || !weight || !bias || !y) return fail(HG_EINVAL, "null argument");
  if (B < 1 || L < 1 || C_in < 1 || C_out < 1 || k < 1 || k > kMaxTaps || dilation < 1) return fail(HG_EINVAL, "bad shape");
  Layer l = make_conv("op.conv1d", C_in, C_out, k, dilation);
  return op_layer(device, precision, l, x, B, L, weight, bias, in_slope, residual, y, stream);
}

extern "C" int hg_op_conv_transpose1d(int device, int precision, const float* x, int B, int L, int C_in,
                                      const float* weight, const float* bias, int C_out, int k, int stride,
                                      float in_slope, float* y, void* stream) {
  if (!x || !weight || !bias || !y) return fail(HG_EINVAL, "null argument");
  if (B < 1 || L < 1 || C_in < 1 || C_out < 1 || k < stride || stride < 1 || (k + stride - 1) / stride > kMaxTaps)
    return fail(HG_EINVAL, "bad shape");
  Layer l = make_convT("op.conv_transpose1d", C_in, C_out, k, stride);
  return op_layer(device, precision, l, x, B, L, weight, bias, in_slope, nullptr, y, stream);
}

extern "C" int hg_op_conv_post(int device, const float* x, int B, int L, int C, const float* weight, const float* bias,
                               float* y, void* stream) {
  if (!x || !weight || !bias || !y) return fail(HG_EINVAL, "null argument");
  int rc = check_device(device);
  if (rc) return rc;
  DEVICE_SCOPE(device);
  Layer l;
  l.name = "op.conv_post"; l.kind = L_POST; l.cin = C; l.cout = 1; l.k = 7; l.pad = 3;
  if ((rc = pack_layer(l, weight, bias))) { free_layer(l); return rc; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = launch_conv_post(x, B, L, C, l.w_post, l.bias_post, y, nullptr, 1.f, st,
                                   l.w_post_host.empty() ? nullptr : l.w_post_host.data());
  cudaError_t es = cudaStreamSynchronize(st);
  free_layer(l);
  if (e != cudaSuccess) return fail(HG_ECUDA, "conv_post: %s", cudaGetErrorString(e));
  if (es != cudaSuccess) return fail(HG_ECUDA, "conv_post execution: %s", cudaGetErrorString(es));
  return HG_OK;
}

extern "C" int hg_op_conv_pair(int device, const float* x, int B, int L, int C, int k, int d1, const float* w1,
                               const float* b1, const float* w2, const float* b2, float in_slope, const float* residual,
                               float* y, void* stream) {
  if (!x || !w1 || !b1 || !w2 || !b2 || !y) return fail(HG_EINVAL, "null argument");
  if (B < 1 || L < 1 || k < 1 || k > kMaxTaps || d1 < 1) return fail(HG_EINVAL, "bad shape");
  int rc = check_device(device);
  if (rc) return rc;
  DEVICE_SCOPE(device);
  HgPlan plan;
  plan.device = device;
  plan.desc_mode = env_int("HG_DESC_MODE", 0);
  {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, device) == cudaSuccess) plan.sm_count = pr.multiProcessorCount;
  }
  plan.fold_pairs = env_int("HG_FOLD", 1) != 0;  // the op-level entry runs the folded kernel wherever it applies
  plan.layers.push_back(make_conv("op.pair.c1", C, C, k, d1));
  plan.layers.push_back(make_conv("op.pair.c2", C, C, k, 1));
  Layer& l1 = plan.layers[0];
  Layer& l2 = plan.layers[1];
  PairTiling pt;
  const bool pair_ok = pair_fusable(&plan, l1, l2, HG_PREC_BF16, &pt);
  if (!pair_ok && C != 16) return fail(HG_EINVAL, "shape not covered by the fused pair kernels");
  if ((rc = pack_layer(l1, w1, b1)) || (rc = pack_layer(l2, w2, b2))) { free_layer(l1); free_layer(l2); return rc; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = static_cast<long long>(B) * L * C;
  void* a = nullptr;
  // + 4 KB: the folded kernel's input map covers whole block groups, up to F*d1 - 1 rows past the last item
  cudaError_t e = cudaMalloc(&a, static_cast<size_t>(n) * 2 + 4096);
  if (e != cudaSuccess) { free_layer(l1); free_layer(l2); return fail(HG_ECUDA, "cudaMalloc: %s", cudaGetErrorString(e)); }
  OperandBuf in;
  in.a0 = a;
  e = launch_f32_to_operand(x, n, in_slope, A_BF16, in.a0, nullptr, st);
  EpiParams ep; memset(&ep, 0, sizeof(ep));
  ep.res = residual; ep.out_x = y; ep.slope = 1.f;
  FoldTiling ft;
  if (e != cudaSuccess) rc = fail(HG_ECUDA, "f32_to_operand: %s", cudaGetErrorString(e));
  else if (fold_fusable(&plan, l1, l2, HG_PREC_BF16, L, &ft)) rc = run_pair_fold(&plan, l1, l2, ft, B, L, in, ep, in_slope, st);
  else if (pair_ok) rc = run_pair(&plan, l1, l2, pt, B, L, in, ep, in_slope, st);
  else rc = fail(HG_EINVAL, "shape not covered by the fused pair kernels (16 channels need L %% 8 == 0)");
  cudaError_t es = cudaStreamSynchronize(st);
  cudaFree(a);
  free_layer(l1); free_layer(l2);
  if (rc) return rc;
  if (es != cudaSuccess) return fail(HG_ECUDA, "op execution failed: %s", cudaGetErrorString(es));
  return HG_OK;
}

extern "C" int hg_op_resblock1(int device, const float* x, int B, int L, int C, int k, int np, const int32_t* d1,
                               const float* const* w1, const float* const* b1, const float* const* w2,
                               const float* const* b2, float slope, float* y, void* stream, int32_t* fused) {
  if (!x || !d1 || !w1 || !b1 || !w2 || !b2 || !y) return fail(HG_EINVAL, "null argument");
  if (B < 1 || L < 1 || k < 1 || k > kMaxTaps || np < 1 || np > kChainMaxPairs) return fail(HG_EINVAL, "bad shape");
  int rc = check_device(device);
  if (rc) return rc;
  DEVICE_SCOPE(device);
  HgPlan plan;
  init_plan_env(&plan, device);
  for (int m = 0; m < np; ++m) plan.layers.push_back(make_conv("op.block.c1." + std::to_string(m), C, C, k, d1[m]));
  for (int m = 0; m < np; ++m) plan.layers.push_back(make_conv("op.block.c2." + std::to_string(m), C, C, k, 1));
  auto cleanup = [&]() { for (auto& l : plan.layers) free_layer(l); };
  for (int m = 0; m < np; ++m)
    if ((rc = pack_layer(plan.layers[m], w1[m], b1[m])) || (rc = pack_layer(plan.layers[np + m], w2[m], b2[m]))) { cleanup(); return rc; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = static_cast<long long>(B) * L * C;
  // scratch: operand planes a (bf16, + slack for the folded kernel's map) x2, fp32 residual stream x2
  void* a[2] = {nullptr, nullptr};
  float* f[2] = {nullptr, nullptr};
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaMalloc(&a[i], static_cast<size_t>(n) * 2 + 4096);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&f[i]), static_cast<size_t>(n) * 4);
  }
  auto release = [&]() { for (int i = 0; i < 2; ++i) { cudaFree(a[i]); cudaFree(f[i]); } cleanup(); };
  if (e != cudaSuccess) { release(); return fail(HG_ECUDA, "cudaMalloc: %s", cudaGetErrorString(e)); }
  e = launch_f32_to_operand(x, n, slope, A_BF16, a[0], nullptr, st);
  if (e != cudaSuccess) { release(); return fail(HG_ECUDA, "f32_to_operand: %s", cudaGetErrorString(e)); }
  ChainTiling ct;
  const bool chain = chain_fusable(&plan, &plan.layers[0], &plan.layers[np], np, HG_PREC_BF16, &ct);
  if (fused) *fused = chain ? 1 : 0;
  if (chain) {
    EpiParams ep; memset(&ep, 0, sizeof(ep));
    ep.res = x; ep.out_x = y; ep.slope = slope;
    OperandBuf in; in.a0 = a[0];
    rc = run_chain(&plan, &plan.layers[0], &plan.layers[np], ct, B, L, in, ep, slope, st);
  } else {
    // pair by pair, as hg_forward schedules an unfused block: x_m in f[], operand in a[]
    const float* res = x;
    int cur = 0;
    for (int m = 0; m < np && !rc; ++m) {
      const Layer& l1 = plan.layers[m];
      const Layer& l2 = plan.layers[np + m];
      const bool last = m == np - 1;
      EpiParams ep; memset(&ep, 0, sizeof(ep));
      ep.res = res; ep.slope = slope;
      ep.out_x = last ? y : f[m & 1];
      if (!last) ep.out_a0 = a[cur ^ 1];
      OperandBuf in; in.a0 = a[cur];
      PairTiling pt;
      FoldTiling ft;
      if (fold_fusable(&plan, l1, l2, HG_PREC_BF16, L, &ft) && (plan.fold_force || fold_pays(&plan, l2, false)))
        rc = run_pair_fold(&plan, l1, l2, ft, B, L, in, ep, slope, st);
      else if (pair_fusable(&plan, l1, l2, HG_PREC_BF16, &pt))
        rc = run_pair(&plan, l1, l2, pt, B, L, in, ep, slope, st);
      else
        rc = fail(HG_EINVAL, "shape not covered by the fused pair kernels");
      res = f[m & 1];
      cur ^= 1;
    }
  }
  cudaError_t es = cudaStreamSynchronize(st);
  release();
  if (rc) return rc;
  if (es != cudaSuccess) return fail(HG_ECUDA, "op execution failed: %s", cudaGetErrorString(es));
  return HG_OK;
}

extern "C" int hg_fold_info(int C, int k, int d1, int L, HgFoldInfo* info) {
  if (!info) return fail(HG_EINVAL, "null info");
  memset(info, 0, sizeof(*info));
  FoldTiling t;
  if (!fold_geometry(C, k, d1, L, &t)) return HG_OK;  // fusable = 0: the N = C pair kernel (or two launches) runs instead
  info->fusable = 1;
  info->f = t.f; info->r_out = t.r_out; info->delta = t.delta; info->fdiv = t.fdiv; info->blk_off = t.smin;
  info->nb_slab = t.nb_slab; info->slab_phase_bytes = t.slab_phase_bytes; info->xt_phase_bytes = t.xt_phase_bytes;
  info->t_bufs = t.t_bufs; info->stages = t.stages; info->weights_resident = t.resident ? 1 : 0;
  info->smem_bytes = static_cast<int32_t>(t.smem);
  FoldOp ops[kFoldMaxOps];
  const int rowb = C * 2;
  for (int conv = 0; conv < 2; ++conv) {
    const int phase_bytes = conv ? t.xt_phase_bytes : t.slab_phase_bytes;
    const int n = conv ? fold_schedule(C, t.f, k, rowb >> 4, t.xt_phase_bytes, t.delta / t.f, 1, ops)
                       : fold_schedule(C, t.f, k, rowb >> 4, t.slab_phase_bytes, -t.smin * d1, d1, ops);
    (conv ? info->n_ops2 : info->n_ops1) = n;
    for (int i = 0; i < n; ++i) {
      int32_t* o = conv ? info->ops2[i] : info->ops1[i];
      const int bytes = ops[i].a_off16 * 16;
      o[0] = bytes / phase_bytes;            // phase slab of the A operand
      o[1] = (bytes % phase_bytes) / rowb;   // first row inside it
      o[2] = ops[i].b_blk; o[3] = ops[i].nblk; o[4] = ops[i].d_col; o[5] = ops[i].rel;
    }
  }
  return HG_OK;
}

extern "C" int hg_fold_ring_query(int k, int period, int tap, int conv_parity, int32_t* slot, int32_t* parity, int32_t* mirror,
                                  int32_t* slots) {
  if (!slot || !parity || !mirror || !slots) return fail(HG_EINVAL, "null output");
  int a, b, c, d;
  if (!conv_fold_ring_query(k, period, tap, conv_parity, &a, &b, &c, &d))
    return fail(HG_EINVAL, "no streamed-weight kernel for k = %d, ring period %d (tap %d)", k, period, tap);
  *slot = a; *parity = b; *mirror = c; *slots = d;
  return HG_OK;
}

extern "C" int hg_selftest_tcgen05(int device, char* buf, size_t buf_len) {
  if (buf && buf_len) buf[0] = 0;
  int rc = check_device(device);
  if (rc) return rc;
  DEVICE_SCOPE(device);
  return run_tcgen05_selftest(buf, buf_len);
}
