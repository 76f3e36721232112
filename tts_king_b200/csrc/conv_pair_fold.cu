// conv_pair_fold.cu — one ResBlock1 pair as a single tcgen05 kernel with TIME FOLDED INTO N
// (bf16 mode, C = 16, 32 or 64):
//
//     xt = leaky_relu(c1(a) + b1)          dilated Conv1d, hifi/models.py:90-92
//     y  = c2(xt) + b2 + x                 Conv1d d=1 + residual, hifi/models.py:93-94
//
// Why a second pair kernel: conv_pair_tc.cu issues M=128 x N=C MMAs, and an SS-mode MMA pulls
// 32 + N/4 operand wavefronts through the shared-memory pipe for N/2 cycles of math
// (profiles/r1_umma_issue_floor.txt) — at N = 64 / 32 the A-operand fetch, not the tensor pipe, sets
// the pace (48 / 40 cycles per MMA against 32 / 16 of math).  The A fetch is the same 32 wavefronts
// whatever N is, so the cure is more columns per fetch.
//
// How: a tap-shifted conv lets F = 128 / C output rows share one A operand.  With dilation d, output
// rows t and t + d use the same input rows one tap apart:
//     y[t + h*d] = sum_j W_j x[t + (h + j - c) d]        (c = (k-1)/2)
// so stacking the phases h = 0..F-1 along N, the A operand "x[t + u*d]" (u = h + j - c) meets the
// weight block [W_{u+c-h}]_h — a run of CONSECUTIVE taps, i.e. a contiguous slice of the packed weight
// image.  A conv becomes k + F - 1 MMA groups of N <= 128 instead of F * k groups of N = C: the same
// FLOPs, (k + F - 1) / (k F) of the A fetches.  For that, M rows must step by F*d in time: the input
// slab is loaded DE-INTERLEAVED — time row g = F*d*blk + h*d + r lives in phase slab h at row
// blk*d + r — which one 5-D TMA box per phase does (dims: channel, r, h, blk, item), and a tap shift
// is again a row shift of the same slab (s*d rows for u = s*F + h').  E1 writes xt the same way for
// c2 (d = 1: phase = row mod F), and the accumulator tile [128 M rows][F phases x C] is exactly a
// contiguous 64 KB block of the channels-last output, so E2 works on the folded view [L/F][128].
// Accumulator columns hold the phases in REVERSE order (column block q = phase F-1-q) so that the
// stacked weight run ascends in tap index and the Layer's packed image [tap][C][C] is used as is.
//
// Arithmetic order: the MMA groups run in ascending u, so every output element — whatever its phase, tile or
// position in the sequence — accumulates its taps in the order 0..k-1 (K steps inside), exactly like
// conv_pair_tc.cu: results are bit-identical between the two kernels, between a time chunk and the
// monolithic forward, and between batch items.  Ascending order means the first groups cover only some of
// the phases, so they cannot carry the "overwrite" flag; instead one extra MMA with an all-zero A operand
// (a 1 KB block whose 8-row groups alias through SBO = 0) clears the accumulator first.
//
// The MMA issuer is ONE thread running straight-line code.  With N = 128 the tensor pipe wants a new
// MMA every 64 cycles; the first version of this kernel walked a host-built schedule table (constant-bank
// loads, ring-slot arithmetic with a modulo, a warp-wide elect / reconverge per group) and spent ~240 cycles
// per MMA in its own instruction stream — slower than the kernel it was meant to replace
// (profiles/r2_fold_issue_bound.md).  So the schedule is compile-time here: the kernel is instantiated per
// tap count K, every group's operand offsets, N and accumulator column are constants of the unrolled code.
// Weight stages: resident weights sit at slot conv*K + tap; streamed weights (C = 64, k >= 5: 2k blocks of
// 8 KB do not fit) go through a ring whose slot cursor is carried from group to group (one compare per
// step, no division) — a run of taps that wraps around the ring end is issued as two MMA groups.
//
// E2 is conv_pair_tc.cu's: the fp32 residual tile of a warp's 32 x 32 item is TMA-prefetched a whole tile
// ahead into the warp's 4 KB slot (over the folded view), the accumulator row is added in place and the slot
// is read back transposed so that eight lanes cover one 128-byte row segment.  Two shared-memory-free
// variants (register transpose by butterfly shuffles, residual by direct loads with an L2 bulk prefetch, or
// by TMA) were measured and lost 0.03-0.06 ms per launch on the HBM-bound k = 3 pairs
// (profiles/r2_fold_issue_bound.md).
//
// Geometry and the input map are laid out by the host (api.cu::fold_geometry); pipeline, barriers and
// warp roles are conv_pair_tc.cu's.  Rows beyond the sequence end inside the last (partial) block group
// cannot be expressed as TMA out-of-bounds, so the slab producer zeroes them in shared memory for the one
// tile per item that sees them.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace hg {

// Timing experiments (tools/fold_dbg.py; build with HG_NVCC_EXTRA=-DHG_FOLD_DBG): TcFoldParams::dbg bits drop parts
// of the kernel's work — 1 MMAs, 2 global stores, 4 residual TMA, 8 slab TMA, 16 staging read-modify-write, 32 xt
// stores, 64 weight streaming (the producer only signals), 256 weight waits and producer, 512 weight-stage releases,
// 1024 epilogue bodies (hand-shakes only).  Results are wrong by construction; the shipped build compiles none of it.
#ifdef HG_FOLD_DBG
#define FOLD_DBG(bit) ((p.dbg & (bit)) != 0)
#else
#define FOLD_DBG(bit) false
#endif

constexpr int kFoldEpiWarps = 16;
constexpr int kFoldThreads = (3 + kFoldEpiWarps) * 32;
constexpr int kFoldZeroBytes = 1024;       // all-zero A operand of the accumulator-clearing MMA
constexpr int kFoldStageFloats = 32 * 32;  // per-warp residual / transpose tile: 32 rows x 32 fp32 columns

__host__ __device__ constexpr int fold_floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
__host__ __device__ constexpr int fold_min(int a, int b) { return a < b ? a : b; }
__host__ __device__ constexpr int fold_max(int a, int b) { return a > b ? a : b; }

// Streamed weights (S > 0): a ring with a COMPILE-TIME period.  Tap j of every conv sits in slot j mod S, so each
// slot, each barrier and each barrier parity of the unrolled schedule is a constant of the code (the first version
// carried a run-time cursor: ~45 dependent instructions per group of four MMAs, and the issuing thread, not the
// tensor pipe, set the pace — profiles/r2_pair_epilogue_issue_bound.md).  Convs go through the ring in the order
// G1(0) G1(1) G2(0) G1(2) G2(1) ...; slot s is used uses(s) times per conv, so use u of the c-th conv is phase
// c * uses(s) + u of its barriers: parity = (uses(s) odd ? c & 1 : 0) ^ (u & 1) — one run-time bit.
// A run of F = 2 consecutive taps (j, j + 1) must be contiguous for the stacked B operand: when j mod S = S - 1 the
// tap j + 1 is ALSO loaded into a mirror slot S behind the ring (same barrier as slot 0, twice the bytes).
template <int K, int S>
struct FoldRingPlan {
  static constexpr bool kMirror = S > 0 && K > S;
  static constexpr int kSlots = S == 0 ? 2 * K : S + (kMirror ? 1 : 0);  // physical weight blocks in shared memory
  static constexpr int kBars = S == 0 ? 2 * K : S;
  __host__ __device__ static constexpr int uses(int s) { return (K - 1 - s) / (S > 0 ? S : 1) + 1; }
  __host__ __device__ static constexpr uint32_t parity(int j, uint32_t cpar) {
    return ((uses(j % (S > 0 ? S : 1)) & 1) ? cpar : 0u) ^ static_cast<uint32_t>((j / (S > 0 ? S : 1)) & 1);
  }
};

// One conv of the pair, issued by a single thread.  Group oi (u = oi - CH) contracts the A operand
// "phase u mod F, shifted by floor(u / F) block groups" against the taps u + CH - h of the phases h that
// use it (consecutive taps, stacked along N, ascending with the accumulator column).  Group oi is the
// first to touch weight block oi (oi < K); its run starts at block max(0, oi - F + 1), which is also the
// block whose last use it is (oi >= F - 1).
//   a_base16   operand buffer base (smem address >> 4)
//   a_phase16  bytes >> 4 between phase slabs
//   a_row0_16  (first row) * ROWB >> 4   — the slab row of shift 0
//   a_shift16  (rows per block-group shift) * ROWB >> 4
//   w_lo       resident: weight stage 0 of this conv; ring: stage 0 of the ring (smem address >> 4)
template <int C, int K, int S>
__device__ __forceinline__ void fold_issue_conv(uint32_t acc, uint32_t a_base16, uint32_t a_phase16, int a_row0_16,
                                                int a_shift16, uint32_t w_lo, uint32_t zero_lo, uint64_t* w_full,
                                                uint64_t* w_empty, uint32_t cpar, bool wait_weights, int nosync = 0) {
  constexpr int F = 128 / C, CH = (K - 1) / 2, ROWB = C * 2, KSTEPS = C / 16, WB16 = (C * ROWB) >> 4;
  constexpr bool RING = S > 0;
  static_assert(!RING || F == 2, "the mirror slot covers runs of two taps");
  using Plan = FoldRingPlan<K, S>;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr uint32_t LAYOUT = (C == 64) ? UMMA_LAYOUT_SW128 : (C == 32) ? UMMA_LAYOUT_SW64 : UMMA_LAYOUT_SW32;
  constexpr uint32_t desc_hi = ((SBO >> 4) & 0x3FFFu) | (1u << 14) | (LAYOUT << 29);
  constexpr uint32_t desc_hi_alias = (1u << 14) | (LAYOUT << 29);  // SBO = 0: every 8-row group is the same 8 rows
  constexpr uint32_t idesc0 = umma_idesc_bf16(128, 0);
#pragma unroll
  for (int oi = 0; oi < K + F - 1; ++oi) {
    const int u = oi - CH;
    const int s = fold_floor_div(u, F), hp = u - s * F;
    const int h_lo = fold_max(0, u + CH - (K - 1)), h_hi = fold_min(F - 1, u + CH);
    const int b_blk = u + CH - h_hi, nblk = h_hi - h_lo + 1, d_col = (F - 1 - h_hi) * C;
    if (oi < K) {  // block oi is first used here
      if (RING) {
        if (!(nosync & 1)) mbar_wait(&w_full[oi % (RING ? S : 1)], Plan::parity(oi, cpar));
        tc_fence_after();
      } else if (wait_weights) {
        mbar_wait(&w_full[oi], 0u);
        tc_fence_after();
      }
    }
    // the run's blocks are contiguous: resident at slot b_blk, streamed at slot b_blk mod S (mirror slot S behind S - 1)
    const uint32_t b0 = w_lo + (RING ? b_blk % (RING ? S : 1) : b_blk) * WB16;
    if (oi == 0)  // D[128 x 128] = 0 * (first rows of weight block 0): one K = 16 MMA with the overwrite flag
      umma_bf16_lohi(acc, zero_lo, b0, desc_hi_alias, idesc0 | (static_cast<uint32_t>(128 >> 3) << 17), 0u);
    const uint32_t a_lo = a_base16 + hp * a_phase16 + static_cast<uint32_t>(a_row0_16 + s * a_shift16);
    const uint32_t idesc = idesc0 | (static_cast<uint32_t>((nblk * C) >> 3) << 17);
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) umma_bf16_lohi(acc + d_col, a_lo + ks * 2, b0 + ks * 2, desc_hi, idesc, 1u);
    if (RING && oi >= F - 1 && oi - F + 1 <= K - 1) {  // block oi - F + 1 (= the run's first) has had its last use
      if (!(nosync & 2)) umma_commit(&w_empty[(oi - F + 1) % (RING ? S : 1)]);
    }
  }
}

// S = 0: all 2K weight blocks stay in shared memory (slot conv*K + tap); S > 0: streamed through a ring of period S.
template <int C, int K, int S>
__global__ void __launch_bounds__(kFoldThreads, 1)
conv_pair_fold_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_res,
                      const __grid_constant__ TcFoldParams p) {
  constexpr int F = 128 / C;
  constexpr int LOG2F = (F == 2) ? 1 : (F == 4) ? 2 : 3;
  constexpr int ROWB = C * 2;
  constexpr int WBLK = C * ROWB;  // one tap's [C rows][C] tile
  constexpr bool RING = S > 0;
  using Plan = FoldRingPlan<K, S>;
  constexpr int STAGES = Plan::kBars;  // weight barriers; Plan::kSlots blocks of shared memory
  constexpr uint32_t ACC_COLS = 128;
  constexpr uint32_t TMEM_COLS = 4 * ACC_COLS;
  static_assert(C == 64 || C == 32 || C == 16, "N = 128 = F x C");
  static_assert(K & 1, "odd taps");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int slab_bytes = F * p.slab_phase_bytes;  // one slab buffer: F phase slabs
  const int t_bytes = F * p.xt_phase_bytes;       // one xt buffer:   F phase slabs
  uint8_t* zero_a = smem;                         // [1 KB] zeros
  uint8_t* slab = smem + kFoldZeroBytes;          // [2][slab_bytes]
  uint8_t* tbuf = slab + 2 * slab_bytes;          // [t_bufs][t_bytes]
  float* staging = reinterpret_cast<float*>(tbuf + p.t_bufs * t_bytes);  // [16][4 KB], 1024-aligned (TMA dst)
  uint8_t* wst = reinterpret_cast<uint8_t*>(staging + kFoldEpiWarps * kFoldStageFloats);  // [Plan::kSlots][WBLK]
  uint64_t* bars = reinterpret_cast<uint64_t*>(wst + Plan::kSlots * WBLK);
  uint64_t* slab_full = bars;        // [2]
  uint64_t* slab_empty = bars + 2;   // [2]
  uint64_t* slab_land = bars + 4;    // [2]  TMA landing barrier of a slab that needs its tail rows zeroed
  uint64_t* d1_full = bars + 6;      // [2]  G1 -> E1
  uint64_t* d1_empty = bars + 8;     // [2]  E1 -> G1
  uint64_t* t_full = bars + 10;      // [2]  E1 -> G2
  uint64_t* t_empty = bars + 12;     // [2]  G2 -> E1
  uint64_t* d2_full = bars + 14;     // [2]  G2 -> E2
  uint64_t* d2_empty = bars + 16;    // [2]  E2 -> G2
  uint64_t* res_bar = bars + 18;     // [16] residual tile landed in a warp's staging slot (TMA)
  uint64_t* w_full = bars + 34;      // [STAGES]
  uint64_t* w_empty = w_full + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_empty + STAGES);
  float* bias1_s = reinterpret_cast<float*>(tmem_ptr_smem + 4);  // [C] conv 1 bias: E1 reads it with broadcast LDS

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = p.total_work > static_cast<int>(blockIdx.x)
                       ? (p.total_work - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                       : 0;

  if (warp == 0 && lane == 0) { prefetch_tensormap(&map_in); prefetch_tensormap(&map_res); }
  if (warp == 2) {  // the zero operand (generic-proxy stores, made visible to the tensor core below)
    *reinterpret_cast<uint4*>(zero_a + lane * 32) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(zero_a + lane * 32 + 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    // a weight, not a product of the previous launch: may be read before griddepcontrol.wait.  (From global memory
    // the same-address float4 loads of E1 cost four L1 wavefronts each — 512 per tile on a saturated data pipe.)
    if (lane < C / 4) reinterpret_cast<float4*>(bias1_s)[lane] = __ldg(reinterpret_cast<const float4*>(p.bias1) + lane);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&slab_full[i], 1); mbar_init(&slab_empty[i], 1); mbar_init(&slab_land[i], 1);
        mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], kFoldEpiWarps);
        mbar_init(&t_full[i], kFoldEpiWarps); mbar_init(&t_empty[i], 1);
        mbar_init(&d2_full[i], 1); mbar_init(&d2_empty[i], kFoldEpiWarps);
      }
      for (int s = 0; s < STAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
      for (int w = 0; w < kFoldEpiWarps; ++w) mbar_init(&res_bar[w], 1);
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
  // programmatic dependent launch: setup and the first weight stages overlap the previous layer's tail
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------ weight producer: blocks in the order the MMA issuer needs them
    if (lane == 0 && n_my > 0 && !(RING && (FOLD_DBG(1) || FOLD_DBG(256)))) {
      if (!RING) {
        for (int cv = 0; cv < 2; ++cv) {
          const uint8_t* w = cv ? p.w2 : p.w1;
          for (int j = 0; j < K; ++j) {
            const int st = cv * K + j;
            mbar_arrive_expect_tx(&w_full[st], WBLK);
            bulk_load_1d(wst + st * WBLK, w + static_cast<size_t>(j) * WBLK, WBLK, &w_full[st]);
          }
        }
      } else {
        uint32_t cpar = 0;  // parity of the conv's position in the ring order
        auto load_conv = [&](const uint8_t* w) {
#pragma unroll
          for (int j = 0; j < K; ++j) {
            constexpr int SS = RING ? S : 1;
            const int slot = j % SS;
            const bool mirror = Plan::kMirror && j >= SS && slot == 0;  // tap j also closes the run (j - 1, j) behind slot S - 1
            mbar_wait(&w_empty[slot], Plan::parity(j, cpar) ^ 1);
            if (FOLD_DBG(64)) {
              mbar_arrive(&w_full[slot]);
            } else {
              mbar_arrive_expect_tx(&w_full[slot], mirror ? 2 * WBLK : WBLK);
              bulk_load_1d(wst + slot * WBLK, w + static_cast<size_t>(j) * WBLK, WBLK, &w_full[slot]);
              if (mirror) bulk_load_1d(wst + SS * WBLK, w + static_cast<size_t>(j) * WBLK, WBLK, &w_full[slot]);
            }
          }
          cpar ^= 1;
        };
        load_conv(p.w1);                       // G1(0)
        for (int i = 0; i < n_my; ++i) {
          if (i + 1 < n_my) load_conv(p.w1);   // G1(i+1)
          load_conv(p.w2);                     // G2(i)
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------ input slab producer (TMA), all lanes walk the loop
    const uint32_t box_bytes = static_cast<uint32_t>(p.nb_slab) * p.d1 * ROWB;
    const int tail = p.L - (p.nblk_item - 1) * p.fdiv;  // valid rows of the item's last block group (fdiv when full)
    uint32_t land_uses0 = 0, land_uses1 = 0;
    for (int i = 0; i < n_my; ++i) {
      const int work = blockIdx.x + i * gridDim.x;
      int b, tile;
      decode_tile(p.rag, p.tiles_per_item, work, b, tile);
      const int g0 = tile * p.r_out - p.delta;       // a multiple of fdiv (possibly negative)
      const int blk_first = g0 / p.fdiv + p.blk_off;
      const int buf = i & 1;
      const int last_rel = p.nblk_item - 1 - blk_first;  // the item's last block group, slab-relative
      const bool fix = tail < p.fdiv && last_rel >= 0 && last_rel < p.nb_slab;
      uint8_t* dst = slab + buf * slab_bytes;
      if (lane == 0) {
        mbar_wait(&slab_empty[buf], ((i >> 1) & 1) ^ 1);
        uint64_t* bar = fix ? &slab_land[buf] : &slab_full[buf];
        if (FOLD_DBG(8)) {
          mbar_arrive(bar);
        } else {
          mbar_arrive_expect_tx(bar, F * box_bytes);
          for (int h = 0; h < F; ++h) tma_load_5d(dst + h * p.slab_phase_bytes, &map_in, bar, 0, 0, h, blk_first, b);
        }
      }
      if (fix) {
        // rows L .. nblk_item*fdiv - 1 sit inside the tensor map's extent (they are the next item's first
        // rows, or whatever follows the buffer): the convolution must see zeros there
        const uint32_t uses = buf ? land_uses1 : land_uses0;
        mbar_wait(&slab_land[buf], uses & 1);
        if (buf) ++land_uses1; else ++land_uses0;
        constexpr int CH16 = ROWB / 16;
        const int nz = (p.fdiv - tail) * CH16;
        for (int e = lane; e < nz; e += 32) {
          const int o = tail + e / CH16;             // row offset inside the block group
          const int h = o / p.d1, r = o - h * p.d1;
          uint8_t* rp = dst + h * p.slab_phase_bytes + (last_rel * p.d1 + r) * ROWB + (e % CH16) * 16;
          *reinterpret_cast<uint4*>(rp) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slab_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer: one thread, straight-line groups.  The guard is
    // elect.sync, not `lane == 0`: only then does the compiler know a single lane is active and feed tcgen05.mma's
    // uniform-register operands with plain R2UR moves instead of a per-instruction "waterfall" loop.
    if (elect_one()) {
      const uint32_t slab_lo = (smem_u32(slab) & 0x3FFFFu) >> 4;
      const uint32_t t_lo = (smem_u32(tbuf) & 0x3FFFFu) >> 4;
      const uint32_t wst_lo = (smem_u32(wst) & 0x3FFFFu) >> 4;
      const uint32_t zero_lo = (smem_u32(zero_a) & 0x3FFFFu) >> 4;
      const uint32_t slab_buf16 = static_cast<uint32_t>(slab_bytes) >> 4, t_buf16 = static_cast<uint32_t>(t_bytes) >> 4;
      const uint32_t slab_phase16 = static_cast<uint32_t>(p.slab_phase_bytes) >> 4, xt_phase16 = static_cast<uint32_t>(p.xt_phase_bytes) >> 4;
      const int a1_row0_16 = p.a1_row0 * (ROWB >> 4), a2_row0_16 = p.a2_row0 * (ROWB >> 4);
      const int a1_shift16 = p.d1 * (ROWB >> 4), a2_shift16 = ROWB >> 4;
      uint32_t cpar = 0;  // streamed weights: parity of the conv's position in the ring order G1(0) G1(1) G2(0) G1(2) ...
      bool w_seen1 = false, w_seen2 = false;  // resident weights: waited for during the first conv 1 / conv 2 only
      auto g1 = [&](int i) {
        const int buf = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        mbar_wait(&d1_empty[buf], ph ^ 1);
        mbar_wait(&slab_full[buf], ph);
        tc_fence_after();
        if (!FOLD_DBG(1))
          fold_issue_conv<C, K, S>(tmem_base + buf * ACC_COLS, slab_lo + buf * slab_buf16, slab_phase16, a1_row0_16, a1_shift16,
                                   wst_lo, zero_lo, w_full, w_empty, cpar, !w_seen1, (FOLD_DBG(256) ? 1 : 0) | (FOLD_DBG(512) ? 2 : 0));
        cpar ^= 1;
        umma_commit(&slab_empty[buf]);
        umma_commit(&d1_full[buf]);
        w_seen1 = true;
      };
      auto g2 = [&](int i) {
        const int buf = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        const int tbi = p.t_bufs == 2 ? buf : 0;
        const uint32_t tph = p.t_bufs == 2 ? ph : static_cast<uint32_t>(i & 1);
        mbar_wait(&t_full[tbi], tph);
        mbar_wait(&d2_empty[buf], ph ^ 1);
        tc_fence_after();
        if (!FOLD_DBG(1))
          fold_issue_conv<C, K, S>(tmem_base + (2 + buf) * ACC_COLS, t_lo + tbi * t_buf16, xt_phase16, a2_row0_16, a2_shift16,
                                   RING ? wst_lo : wst_lo + K * (WBLK >> 4), zero_lo, RING ? w_full : w_full + K,
                                   RING ? w_empty : w_empty + K, cpar, !w_seen2, (FOLD_DBG(256) ? 1 : 0) | (FOLD_DBG(512) ? 2 : 0));
        cpar ^= 1;
        umma_commit(&t_empty[tbi]);
        umma_commit(&d2_full[buf]);
        w_seen2 = true;
      };
      if (n_my > 0) g1(0);
      for (int i = 0; i < n_my; ++i) {
        if (i + 1 < n_my) g1(i + 1);
        g2(i);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue warps (all 16 do E1 then E2)
    const int e = warp - 3;
    const int quarter = warp & 3;
    const int sub = e >> 2;  // 0..3: which of the four warps sharing this TMEM lane quarter
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    // E1: this thread's M row and where its F output rows of xt go (relative to the tile origin)
    const int mrow = quarter * 32 + lane;
    const int blk1 = mrow / p.d1, r1 = mrow - blk1 * p.d1;
    // E2: this warp's 32-column item; accumulator column block q holds phase F-1-q
    const int c02 = sub * 32;                                   // accumulator columns
    // columns of the folded output row this item lands on.  At C = 16 the item spans two phases, stored in
    // the accumulator in descending phase order: its two 16-column halves are swapped relative to memory.
    const int c02m = (C >= 32) ? (F - 1 - c02 / C) * C + (c02 % C) : (F - (c02 + 32) / C) * C;

    // The epilogue warps are INSTRUCTION-ISSUE bound (profiles/r2_pair_epilogue_issue_bound.md: with every MMA,
    // global access and shared-memory transpose switched off the kernel still took 57 % of its time, 16 warps x
    // ~1 260 SASS instructions per tile): everything below is written for instruction count — tile coordinates
    // are decoded once per tile, shared memory is addressed through 32-bit window addresses with the swizzle
    // folded into per-thread constants, global rows step by compile-time strides from one 64-bit base, row
    // validity is one 32-bit count, and the biases live in registers.
    struct TileAt { int b, g0, q0; };  // item, first xt row (global time row), first folded output row
    auto locate = [&](int i) {
      int b, tile;
      decode_tile(p.rag, p.tiles_per_item, blockIdx.x + i * gridDim.x, b, tile);
      return TileAt{b, tile * p.r_out - p.delta, (tile * p.r_out) >> LOG2F};
    };
    const uint32_t tb_s = smem_u32(tbuf);
    float* stg = staging + e * kFoldStageFloats;
    const uint32_t stg_s = smem_u32(stg);
    const int c4 = lane & 7, rsub = lane >> 3;
    const int n2 = c02m + c4 * 4;
    const int tau0 = p.fdiv * blk1 + r1;  // xt row of (this M row, phase 0), tile-relative; phase h adds h * d1
    const float slope1 = p.slope;
    const uint32_t bias1_a = smem_u32(bias1_s);

    // E1: D1 -> (+b1, leaky_relu, bf16) -> xt phase slabs in UMMA layout; two 16-column items per warp
    auto e1 = [&](int i, const TileAt& at) {
      const int buf = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      const int tbi = p.t_bufs == 2 ? buf : 0;
      const uint32_t tph = p.t_bufs == 2 ? ph : static_cast<uint32_t>(i & 1);
      mbar_wait(&d1_full[buf], ph);
      mbar_wait(&t_empty[tbi], tph ^ 1);  // the G2 that last read this xt buffer has retired
      tc_fence_after();
      const uint32_t tbs = tb_s + tbi * t_bytes;
      const uint32_t tmem_acc = tmem_base + buf * ACC_COLS + lane_base;
      if (FOLD_DBG(1024)) {  // hand-shakes only
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&d1_empty[buf]); mbar_arrive(&t_full[tbi]); }
        return;
      }
#pragma unroll 1
      for (int j = sub; j < 8; j += 4) {
        uint32_t r[16];
        tmem_ld_32x16(tmem_acc + 16 * j, r);
        const int q = (16 * j) / C, c0 = (16 * j) % C;
        const int tau = tau0 + (F - 1 - q) * p.d1;
        const bool inside = static_cast<unsigned>(at.g0 + tau) < static_cast<unsigned>(p.L);
        const uint32_t bp = bias1_a + c0 * 4;
        const float4 b0 = lds128(bp), b1 = lds128(bp + 16), b2 = lds128(bp + 32), b3 = lds128(bp + 48);
        tmem_ld_wait();
        if (j + 4 >= 8) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&d1_empty[buf]);
        }
        auto item4 = [&](int q, const float4 bb) {  // bf16x4 of leaky_relu(acc + bias)
          return lrelu_pack4(add4(make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                              __uint_as_float(r[4 * q + 3])), bb), slope1);
        };
        uint2 u0 = item4(0, b0), u1 = item4(1, b1), u2 = item4(2, b2), u3 = item4(3, b3);
        if (!inside) u0 = u1 = u2 = u3 = make_uint2(0u, 0u);  // rows outside the sequence are the conv's zero padding
        // c2 has dilation 1: xt row tau lives in phase slab tau mod F at row tau / F
        const int row = tau >> LOG2F;
        const uint32_t swz = (C == 64) ? (row & 7) : (C == 32) ? ((row >> 1) & 3) : ((row >> 2) & 1);
        const uint32_t ch = c0 >> 3;  // first 16-byte chunk of this item within the row
        const uint32_t rp = tbs + (tau & (F - 1)) * p.xt_phase_bytes + row * ROWB;
        if (!FOLD_DBG(32)) {
          sts128u(rp + ((ch ^ swz) << 4), u0.x, u0.y, u1.x, u1.y);
          sts128u(rp + (((ch + 1) ^ swz) << 4), u2.x, u2.y, u3.x, u3.y);
        }
      }
      fence_proxy_async();  // generic-proxy writes of xt -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_full[tbi]);
    };

    // E2 on the folded view: M row i of tile t is folded output row t*r_out/F + i (F time rows x C channels =
    // 128 fp32 = 512 contiguous bytes).  The fp32 residual tile of this warp's item is TMA-loaded into the
    // warp's 4 KB staging slot (128B-swizzled box = the staging swizzle) a whole tile ahead; the accumulator
    // row is added to it in place, and after the transpose 8 lanes cover one 128-byte row segment.
    auto prefetch_res = [&](const TileAt& at) {  // lane 0 only
      if (FOLD_DBG(4)) { mbar_arrive(&res_bar[e]); return; }
      mbar_arrive_expect_tx(&res_bar[e], kFoldStageFloats * 4);
      tma_load_3d(stg, &map_res, &res_bar[e], c02m, at.q0 + quarter * 32, at.b);
      // the MRF running sum of the same tile (last pair of ResBlocks 1, 2 of a stage) is read by plain loads
      // below: pull its contiguous 64 KB block into L2 now so that they do not wait on HBM
      if (e == 0 && p.epi.acc_in) {
        const long long first = static_cast<long long>(at.q0) * 128;
        const long long left = (p.epi.out_extent - first) * 4;
        if (left > 0)
          bulk_prefetch_l2(p.epi.acc_in + static_cast<long long>(at.b) * p.epi.out_batch_stride + first,
                           static_cast<uint32_t>(left < 65536 ? left : 65536) & ~15u);
      }
    };
    const float4 bias2 = __ldg(reinterpret_cast<const float4*>(p.epi.bias + n2));
    const int rows_item = static_cast<int>(p.epi.out_extent >> 7);  // folded rows per item
    const int rows_tile = p.r_out >> LOG2F;                          // folded rows a tile owns (the rest is halo)
    // lane-per-row view of the slot (the TMEM register layout) and its transpose (8 lanes per 128-byte row segment)
    const uint32_t row_s = stg_s + lane * 128, x7 = static_cast<uint32_t>(lane & 7) << 4;
    const uint32_t tr0_s = stg_s + rsub * 128 + ((c4 ^ rsub) << 4);               // rows rsub, rsub + 8, ...
    const uint32_t tr1_s = stg_s + (rsub + 4) * 128 + ((c4 ^ (rsub + 4)) << 4);   // rows rsub + 4, rsub + 12, ...
    auto e2 = [&](int i, const TileAt& at, const TileAt* next) {
      const int buf = i & 1;
      // rows 4*ii + rsub (ii = 0..7) of this warp's 32: valid while inside the tile's own rows and the item
      int nv = (rows_tile < rows_item - at.q0 ? rows_tile : rows_item - at.q0) - (quarter * 32 + rsub);
      if (FOLD_DBG(2)) nv = 0;
      const long long off = static_cast<long long>(at.b) * p.epi.out_batch_stride +
                            (static_cast<long long>(at.q0 + quarter * 32 + rsub) << 7) + n2;
      mbar_wait(&d2_full[buf], (i >> 1) & 1);
      tc_fence_after();
      if (FOLD_DBG(1024)) {  // hand-shakes only
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d2_empty[buf]);
        mbar_wait(&res_bar[e], i & 1);
        if (lane == 0 && next) prefetch_res(*next);
        return;
      }
      {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (2 + buf) * ACC_COLS + lane_base + c02, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d2_empty[buf]);
        mbar_wait(&res_bar[e], i & 1);
        if (!FOLD_DBG(16))
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const uint32_t a = row_s + ((k4 << 4) ^ x7);
          float4 t = lds128(a);
          const int ka = (C >= 32) ? k4 : (k4 ^ 4);  // accumulator float4 that belongs to memory float4 k4
          t = add4(t, make_float4(__uint_as_float(r[4 * ka]), __uint_as_float(r[4 * ka + 1]), __uint_as_float(r[4 * ka + 2]),
                                  __uint_as_float(r[4 * ka + 3])));
          sts128(a, t);
        }
      }
      __syncwarp();
      float4 v[8];
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) v[ii] = lds128(((ii & 1) ? tr1_s : tr0_s) + (ii >> 1) * 1024);
      fence_proxy_async();  // our generic reads of the slot happen-before the next TMA write into it
      __syncwarp();
      if (lane == 0 && next) prefetch_res(*next);
      epilogue_tail8<512>(p.epi, off, nv, bias2, v);
    };
    if (n_my > 0) {
      TileAt cur = locate(0);
      if (lane == 0) prefetch_res(cur);
      e1(0, cur);
      for (int i = 0; i < n_my; ++i) {
        TileAt nxt = cur;
        const bool more = i + 1 < n_my;
        if (more) {
          nxt = locate(i + 1);
          e1(i + 1, nxt);
        }
        e2(i, cur, more ? &nxt : nullptr);
        cur = nxt;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// stages: 2k when resident, the ring depth otherwise
size_t conv_fold_smem_bytes(int c, int slab_phase_bytes, int xt_phase_bytes, int t_bufs, int stages) {
  const int f = 128 / c;
  return 1024 + kFoldZeroBytes + 2 * static_cast<size_t>(f) * slab_phase_bytes +
         static_cast<size_t>(t_bufs) * f * xt_phase_bytes + kFoldEpiWarps * kFoldStageFloats * 4 +
         static_cast<size_t>(stages) * c * c * 2 + (34 + 2 * stages) * 8 + 16 + 256;
}

// (C, k, ring period) combinations the kernel is instantiated for; period 0 = resident weights.  C = 16 / 32 keep
// their weights resident for every k; C = 64 does for k = 3 and streams from k = 5 on, through a ring of period k (no
// mirror slot), 6, 5 or 4 (+ mirror) — whichever fits next to the slabs (api.cu::fold_geometry).  (C = 32, k = 3 is left
// to conv_pair_tc.cu.)
bool conv_fold_has_kernel(int c, int k, int s) {
  if (c == 16) return s == 0 && (k == 3 || k == 5 || k == 7 || k == 9 || k == 11);
  if (c == 32) return s == 0 && (k == 5 || k == 7 || k == 9 || k == 11);
  if (c == 64) {
    if (k == 3) return s == 0;
    if (k == 5) return s == 5;
    if (k == 7) return s == 7 || s == 5 || s == 4;
    if (k == 9 || k == 11) return s == 6 || s == 5 || s == 4;
  }
  return false;
}
// physical weight blocks of shared memory the (k, period) kernel uses
int conv_fold_weight_slots(int k, int s) { return s == 0 ? 2 * k : s + (k > s ? 1 : 0); }

// host-side view of FoldRingPlan<k, s> (hg_fold_ring_query): the constexpr functions the kernels are compiled from
template <int K, int S>
static void ring_query(int tap, int cpar, int* slot, int* parity, int* mirror, int* slots) {
  using Plan = FoldRingPlan<K, S>;
  *slot = tap % S;
  *parity = static_cast<int>(Plan::parity(tap, static_cast<uint32_t>(cpar & 1)));
  *mirror = (Plan::kMirror && tap >= S && tap % S == 0) ? 1 : 0;
  *slots = Plan::kSlots;
}
bool conv_fold_ring_query(int k, int s, int tap, int cpar, int* slot, int* parity, int* mirror, int* slots) {
  if (s <= 0 || tap < 0 || tap >= k || !conv_fold_has_kernel(64, k, s)) return false;
#define HG_RING_CASE(KV, SV) if (k == KV && s == SV) { ring_query<KV, SV>(tap, cpar, slot, parity, mirror, slots); return true; }
  HG_RING_CASE(5, 5)
  HG_RING_CASE(7, 7) HG_RING_CASE(7, 5) HG_RING_CASE(7, 4)
  HG_RING_CASE(9, 6) HG_RING_CASE(9, 5) HG_RING_CASE(9, 4)
  HG_RING_CASE(11, 6) HG_RING_CASE(11, 5) HG_RING_CASE(11, 4)
#undef HG_RING_CASE
  return false;
}

template <int C, int K, int S>
static cudaError_t launch_fold(const CUtensorMap& m, const CUtensorMap& mr, const TcFoldParams& p, size_t smem, int grid,
                               cudaStream_t st) {
  auto kern = conv_pair_fold_kernel<C, K, S>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kFoldThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, m, mr, p);
}

// mr: fp32 tile map over the folded residual view [B][L/F][128]; s: ring period (0 = resident weights)
cudaError_t launch_conv_pair_fold(int c, int k, int s, const CUtensorMap& m, const CUtensorMap& mr, const TcFoldParams& p,
                                  size_t smem, int grid, cudaStream_t st) {
  if (!conv_fold_has_kernel(c, k, s)) return cudaErrorInvalidValue;
#define HG_FOLD_CASE(CV, KV, SV) \
  if (c == CV && k == KV && s == SV) return launch_fold<CV, KV, SV>(m, mr, p, smem, grid, st);
  HG_FOLD_CASE(16, 3, 0) HG_FOLD_CASE(16, 5, 0) HG_FOLD_CASE(16, 7, 0) HG_FOLD_CASE(16, 9, 0) HG_FOLD_CASE(16, 11, 0)
  HG_FOLD_CASE(32, 5, 0) HG_FOLD_CASE(32, 7, 0) HG_FOLD_CASE(32, 9, 0) HG_FOLD_CASE(32, 11, 0)
  HG_FOLD_CASE(64, 3, 0)
  HG_FOLD_CASE(64, 5, 5)
  HG_FOLD_CASE(64, 7, 7) HG_FOLD_CASE(64, 7, 5) HG_FOLD_CASE(64, 7, 4)
  HG_FOLD_CASE(64, 9, 6) HG_FOLD_CASE(64, 9, 5) HG_FOLD_CASE(64, 9, 4)
  HG_FOLD_CASE(64, 11, 6) HG_FOLD_CASE(64, 11, 5) HG_FOLD_CASE(64, 11, 4)
#undef HG_FOLD_CASE
  return cudaErrorInvalidValue;
}

}  // namespace hg
