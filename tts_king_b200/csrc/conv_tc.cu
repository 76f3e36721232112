// conv_tc.cu — tap-shifted implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// One kernel covers every GEMM-shaped layer of the generator:
//   * same-length dilated Conv1d  (hifi/models.py:19-81 ResBlock convs, :152-154 conv_pre)
//   * ConvTranspose1d in polyphase form (hifi/models.py:161-171): a 2-tap
//     convolution whose N dimension enumerates (phase, C_out) and whose output rows are the
//     phases interleaved in time.
//
// GEMM mapping:  D[M = 128 time rows][N = N_T out channels] += A[M][K] * B[N][K]^T with
// K = (tap, input channel).  Activations are channels-last bf16, so a [rows x KC] slab is
// K-major in shared memory exactly as TMA lands it (128B- or 64B-swizzled rows).  The slab for a
// tile (tile rows + dilation halo) is loaded ONCE per K chunk; each tap is the same slab read
// through a UMMA descriptor whose start address is advanced by tap_row*ROW_BYTES — the dilation
// shift costs no memory traffic, and rows outside [0,L) are TMA zero fill = the conv's padding.
// Weights are pre-packed on the host into swizzled [N_T x KC] tiles and streamed through a ring of
// stages with 1-D bulk copies.  Accumulators (MS sub-tiles of 128 rows) live in TMEM.
//
// Warp roles (224 threads): 0 weight producer | 1 MMA issuer + TMEM owner | 2 slab producer |
// 3..6 epilogue (TMEM lane quarter = warp % 4).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace hg {

constexpr int kTcThreads = 224;

template <int KC>
struct SwzOf {
  static constexpr uint32_t layout = (KC == 64) ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
};

__device__ __forceinline__ uint32_t desc_base_offset(uint32_t saddr, int desc_mode) {
  // desc_mode 0 (default, verified on B200 by hg_selftest_tcgen05 for every row shift, SW128 and
  //   SW64): the XOR swizzle phase comes from absolute smem address bits, base_offset stays 0.
  // desc_mode 1: base_offset = (start address >> 7) & 7 — measured WRONG for shifts that are not a
  //   multiple of 8 rows; kept only so the self-test can document it.
  return desc_mode == 1 ? ((saddr >> 7) & 7u) : 0u;
}

template <int N_T, int KC, int MS, bool SPLIT>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
               const TcConvParams p) {
  constexpr int ROWB = KC * 2;                 // bytes per slab / weight row (= swizzle span)
  constexpr int PLANES = SPLIT ? 2 : 1;
  constexpr int STAGE_BYTES = N_T * ROWB;
  constexpr int KSTEPS = KC / 16;              // UMMA K = 16 bf16
  constexpr uint32_t TMEM_COLS = MS * N_T;     // fp32 accumulator columns
  constexpr uint32_t SBO = 8 * ROWB;           // bytes between 8-row groups
  static_assert(TMEM_COLS >= 32 && TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int slab_bytes = p.slab_rows * ROWB;
  uint8_t* slab = smem;
  uint8_t* wst = smem + ((p.nbuf * PLANES * slab_bytes + 1023) & ~1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(wst + p.stages * STAGE_BYTES);
  uint64_t* slab_full = bars;          // [2]
  uint64_t* slab_empty = bars + 2;     // [2]
  uint64_t* acc_full = bars + 4;       // [1]
  uint64_t* w_full = bars + 5;         // [stages]
  uint64_t* w_empty = w_full + p.stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_empty + p.stages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.tiles_per_item;
  const int tile = blockIdx.x - b * p.tiles_per_item;
  const int m0 = tile * (MS * 128);
  const int nblk = blockIdx.y;
  int ms_count = (p.rows - m0 + 127) / 128;
  ms_count = ms_count > MS ? MS : ms_count;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_hi);
    if (SPLIT) prefetch_tensormap(&map_lo);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(&slab_full[0], 1); mbar_init(&slab_full[1], 1);
      mbar_init(&slab_empty[0], 1); mbar_init(&slab_empty[1], 1);
      mbar_init(acc_full, 1);
      for (int s = 0; s < p.stages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------------------------ weight producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const size_t blk_off = static_cast<size_t>(nblk) * p.nc * p.ntaps * STAGE_BYTES;
      for (int c = 0; c < p.nc; ++c) {
        for (int t = 0; t < p.ntaps; ++t) {
#pragma unroll
          for (int wp = 0; wp < PLANES; ++wp) {
            mbar_wait(&w_empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&w_full[stage], STAGE_BYTES);
            const uint8_t* src = (wp ? p.w_lo : p.w_hi) + blk_off + static_cast<size_t>(c * p.ntaps + t) * STAGE_BYTES;
            bulk_load_1d(wst + stage * STAGE_BYTES, src, STAGE_BYTES, &w_full[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------ activation slab producer (TMA)
    if (lane == 0) {
      for (int c = 0; c < p.nc; ++c) {
        const int buf = c % p.nbuf;
        const uint32_t use = static_cast<uint32_t>(c / p.nbuf);
        mbar_wait(&slab_empty[buf], (use & 1) ^ 1);
        mbar_arrive_expect_tx(&slab_full[buf], PLANES * slab_bytes);
#pragma unroll
        for (int pl = 0; pl < PLANES; ++pl) {
          uint8_t* dst = slab + (buf * PLANES + pl) * slab_bytes;
          for (int bx = 0; bx < p.nboxes; ++bx)
            tma_load_3d(dst + bx * p.box_rows * ROWB, pl ? &map_lo : &map_hi, &slab_full[buf], c * KC,
                        m0 + p.min_off + bx * p.box_rows, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, N_T);
      int stage = 0; uint32_t phase = 0;
      for (int c = 0; c < p.nc; ++c) {
        const int buf = c % p.nbuf;
        const uint32_t use = static_cast<uint32_t>(c / p.nbuf);
        mbar_wait(&slab_full[buf], use & 1);
        tc_fence_after();
        for (int t = 0; t < p.ntaps; ++t) {
#pragma unroll
          for (int wp = 0; wp < PLANES; ++wp) {
            mbar_wait(&w_full[stage], phase);
            tc_fence_after();
            const uint32_t w_base = smem_u32(wst + stage * STAGE_BYTES);
            const int n_a = (SPLIT && wp == 0) ? 2 : 1;  // W_hi meets A_hi and A_lo; W_lo meets A_hi
            for (int ap = 0; ap < n_a; ++ap) {
              const uint32_t a_plane = smem_u32(slab + (buf * PLANES + ap) * slab_bytes);
              for (int ms = 0; ms < ms_count; ++ms) {
                const uint32_t a_base = a_plane + static_cast<uint32_t>(ms * 128 + p.tap_row[t]) * ROWB;
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks) {
                  const uint32_t a_addr = a_base + ks * 32;
                  const uint64_t adesc =
                      umma_smem_desc(a_addr, 0, SBO, SwzOf<KC>::layout, desc_base_offset(a_addr, p.desc_mode));
                  const uint64_t bdesc = umma_smem_desc(w_base + ks * 32, 0, SBO, SwzOf<KC>::layout, 0);
                  const uint32_t accumulate = (c | t | wp | ap | ks) != 0 ? 1u : 0u;
                  umma_bf16(tmem_base + ms * N_T, adesc, bdesc, idesc, accumulate);
                }
              }
            }
            umma_commit(&w_empty[stage]);  // frees the weight stage once these MMAs retire
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(&slab_empty[buf]);
      }
      umma_commit(acc_full);
    }
  } else {
    // ------------------------------------------------ epilogue: TMEM -> registers -> HBM
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int ms = 0; ms < ms_count; ++ms) {
      const long long q = static_cast<long long>(m0) + ms * 128 + row_in_tile;
#pragma unroll 1
      for (int c0 = 0; c0 < N_T; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + ms * N_T + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          epilogue_vec4(p.epi, b, q, nblk * N_T + c0 + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                        __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// host side

size_t conv_tc_smem_bytes(int n_t, int kc, bool split, int slab_rows, int nbuf, int stages) {
  const int rowb = kc * 2, planes = split ? 2 : 1;
  size_t slab = (static_cast<size_t>(nbuf) * planes * slab_rows * rowb + 1023) & ~size_t(1023);
  return 1024 + slab + static_cast<size_t>(stages) * n_t * rowb + (5 + 2 * stages) * 8 + 16;
}

template <int N_T, int KC, int MS, bool SPLIT>
static cudaError_t launch_one(const CUtensorMap& mh, const CUtensorMap& ml, const TcConvParams& p, int n_blocks,
                              size_t smem, cudaStream_t st) {
  auto kern = conv_tc_kernel<N_T, KC, MS, SPLIT>;
  static size_t configured = 0;  // per-instantiation high-water mark
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  dim3 grid(static_cast<unsigned>(p.B * p.tiles_per_item), static_cast<unsigned>(n_blocks));
  kern<<<grid, kTcThreads, smem, st>>>(mh, ml, p);
  return cudaGetLastError();
}

template <int N_T, int KC, int MS>
static cudaError_t launch_split(bool split, const CUtensorMap& mh, const CUtensorMap& ml, const TcConvParams& p,
                                int n_blocks, size_t smem, cudaStream_t st) {
  return split ? launch_one<N_T, KC, MS, true>(mh, ml, p, n_blocks, smem, st)
               : launch_one<N_T, KC, MS, false>(mh, ml, p, n_blocks, smem, st);
}

// Valid (N_T, KC, MS): N_T in {32,64,128,256}, KC in {32,64}, MS*N_T <= 512, MS in {1,2,4}.
cudaError_t launch_conv_tc(int n_t, int kc, int ms, bool split, const CUtensorMap& mh, const CUtensorMap& ml,
                           const TcConvParams& p, int n_blocks, size_t smem, cudaStream_t st) {
#define HG_CASE(NT, KCV, MSV) \
  if (n_t == NT && kc == KCV && ms == MSV) return launch_split<NT, KCV, MSV>(split, mh, ml, p, n_blocks, smem, st);
  HG_CASE(256, 64, 1) HG_CASE(256, 64, 2)
  HG_CASE(128, 64, 1) HG_CASE(128, 64, 2) HG_CASE(128, 64, 4)
  HG_CASE(64, 64, 1) HG_CASE(64, 64, 2) HG_CASE(64, 64, 4)
  HG_CASE(64, 32, 1) HG_CASE(64, 32, 2) HG_CASE(64, 32, 4)
  HG_CASE(32, 64, 1) HG_CASE(32, 64, 2) HG_CASE(32, 64, 4)
  HG_CASE(32, 32, 1) HG_CASE(32, 32, 2) HG_CASE(32, 32, 4)
  HG_CASE(128, 32, 1) HG_CASE(128, 32, 2)
#undef HG_CASE
  return cudaErrorInvalidValue;
}

}  // namespace hg
