// conv_pair_fold.cu — one ResBlock1 pair as a single tcgen05 kernel with TIME FOLDED INTO N
// (bf16 mode, C = 32 or 64):
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
// E2 needs no shared memory: the 32 x 32 accumulator item of a warp is transposed in registers (three
// butterfly exchanges among the eight lanes that share lane & 3) into the layout in which eight lanes cover
// one 128-byte row segment; residual loads and the x / operand stores are then plain coalesced 16-byte
// accesses.  That frees the 64 KB the staging slots took in conv_pair_tc.cu for weight stages.
//
// Geometry, schedule and the input map are laid out by the host (api.cu::fold_geometry /
// fold_schedule); pipeline, barriers and warp roles are conv_pair_tc.cu's.  Rows beyond the sequence end inside the last (partial) block group cannot be
// expressed as TMA out-of-bounds, so the slab producer zeroes them in shared memory for the one tile
// per item that sees them.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace hg {

constexpr int kFoldEpiWarps = 16;
constexpr int kFoldThreads = (3 + kFoldEpiWarps) * 32;
constexpr int kFoldZeroBytes = 1024;       // all-zero A operand of the accumulator-clearing MMA
constexpr int kFoldStageFloats = 32 * 32;  // per-warp residual tile (E2 modes 0 and 2): 32 rows x 32 fp32 columns

// E2M selects how E2 gets the residual in and the rows out (all three give identical bits):
//   0  residual tile TMA-prefetched a tile ahead into a 4 KB per-warp slot, accumulator added in place, transposed
//      read of the slot, coalesced stores (conv_pair_tc.cu's E2)
//   1  no shared memory at all: register transpose, then plain coalesced residual loads and stores
//   2  TMA-prefetched residual slot read row-wise into registers, register transpose, coalesced stores
// DBG = true is the HG_TC_DEBUG_TIMING build (cycle counters around every wait).
template <int C, int E2M, bool DBG>
__global__ void __launch_bounds__(kFoldThreads, 1)
conv_pair_fold_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_res,
                      const __grid_constant__ TcFoldParams p) {
  constexpr bool STAGED = E2M != 1;
  constexpr int F = 128 / C;
  constexpr int LOG2F = (F == 2) ? 1 : (F == 4) ? 2 : 3;
  constexpr int ROWB = C * 2;
  constexpr int KSTEPS = C / 16;
  constexpr int WBLK = C * ROWB;  // one tap's [C rows][C] tile
  constexpr uint32_t ACC_COLS = 128;
  constexpr uint32_t TMEM_COLS = 4 * ACC_COLS;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr uint32_t LAYOUT = (C == 64) ? UMMA_LAYOUT_SW128 : (C == 32) ? UMMA_LAYOUT_SW64 : UMMA_LAYOUT_SW32;
  static_assert(C == 64 || C == 32, "C = 16 needs the half-swapped E2 item (see DESIGN.md)");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int slab_bytes = F * p.slab_phase_bytes;  // one slab buffer: F phase slabs
  const int t_bytes = F * p.xt_phase_bytes;       // one xt buffer:   F phase slabs
  uint8_t* zero_a = smem;                         // [1 KB] zeros
  uint8_t* slab = smem + kFoldZeroBytes;          // [2][slab_bytes]
  uint8_t* tbuf = slab + 2 * slab_bytes;          // [t_bufs][t_bytes]
  float* staging = reinterpret_cast<float*>(tbuf + p.t_bufs * t_bytes);  // [16][4 KB] (STAGED), 1024-aligned (TMA dst)
  uint8_t* wst = reinterpret_cast<uint8_t*>(staging + (STAGED ? kFoldEpiWarps * kFoldStageFloats : 0));  // [stages][WBLK]
  uint64_t* bars = reinterpret_cast<uint64_t*>(wst + p.stages * WBLK);
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
  uint64_t* w_full = bars + 34;      // [stages]
  uint64_t* w_empty = w_full + p.stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_empty + p.stages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = p.total_work > static_cast<int>(blockIdx.x)
                       ? (p.total_work - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                       : 0;

  if (warp == 0 && lane == 0) { prefetch_tensormap(&map_in); if (STAGED) prefetch_tensormap(&map_res); }
  if (warp == 2) {  // the zero operand (generic-proxy stores, made visible to the tensor core below)
    *reinterpret_cast<uint4*>(zero_a + lane * 32) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(zero_a + lane * 32 + 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&slab_full[i], 1); mbar_init(&slab_empty[i], 1); mbar_init(&slab_land[i], 1);
        mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], kFoldEpiWarps);
        mbar_init(&t_full[i], kFoldEpiWarps); mbar_init(&t_empty[i], 1);
        mbar_init(&d2_full[i], 1); mbar_init(&d2_empty[i], kFoldEpiWarps);
      }
      for (int s = 0; s < p.stages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
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
    if (lane == 0 && n_my > 0) {
      if (p.w_resident) {
        for (int cv = 0; cv < 2; ++cv) {
          const uint8_t* w = cv ? p.w2 : p.w1;
          for (int j = 0; j < p.k; ++j) {
            const int st = cv * p.k + j;
            mbar_arrive_expect_tx(&w_full[st], WBLK);
            bulk_load_1d(wst + st * WBLK, w + static_cast<size_t>(j) * WBLK, WBLK, &w_full[st]);
          }
        }
      } else {
        int stage = 0; uint32_t phase = 0;
        auto load_conv = [&](const uint8_t* w) {
          for (int j = 0; j < p.k; ++j) {
            mbar_wait(&w_empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&w_full[stage], WBLK);
            bulk_load_1d(wst + stage * WBLK, w + static_cast<size_t>(j) * WBLK, WBLK, &w_full[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
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
    uint32_t land_uses[2] = {0, 0};
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
        mbar_arrive_expect_tx(bar, F * box_bytes);
        for (int h = 0; h < F; ++h) tma_load_5d(dst + h * p.slab_phase_bytes, &map_in, bar, 0, 0, h, blk_first, b);
      }
      if (fix) {
        // rows L .. nblk_item*fdiv - 1 sit inside the tensor map's extent (they are the next item's first
        // rows, or whatever follows the buffer): the convolution must see zeros there
        mbar_wait(&slab_land[buf], land_uses[buf] & 1);
        ++land_uses[buf];
        constexpr int CH = ROWB / 16;
        const int nz = (p.fdiv - tail) * CH;
        for (int e = lane; e < nz; e += 32) {
          const int o = tail + e / CH;               // row offset inside the block group
          const int h = o / p.d1, r = o - h * p.d1;
          uint8_t* rp = dst + h * p.slab_phase_bytes + (last_rel * p.d1 + r) * ROWB + (e % CH) * 16;
          *reinterpret_cast<uint4*>(rp) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slab_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc0 = umma_idesc_bf16(128, 0);
    constexpr uint32_t desc_hi = ((SBO >> 4) & 0x3FFFu) | (1u << 14) | (LAYOUT << 29);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t slab_lo = (smem_u32(slab) & 0x3FFFFu) >> 4;
    const uint32_t t_lo = (smem_u32(tbuf) & 0x3FFFFu) >> 4;
    const uint32_t wst_lo = (smem_u32(wst) & 0x3FFFFu) >> 4;
    // streamed weights: blocks are numbered along the conv sequence G1(0) G1(1) G2(0) G1(2) ...; block n sits
    // in ring slot n % stages.  base = slot of the current conv's block 0, avail = next block to wait for,
    // head = oldest block not yet released.
    int base_slot = 0, avail_slot = 0, head_slot = 0;
    uint32_t avail_phase = 0;
    int avail_ahead = 0;  // blocks of the current conv already waited for
    bool w_seen = false;  // resident mode: every stage has been waited for once
    // bring-up instrumentation (HG_TC_DEBUG_TIMING): cycles spent in each wait, kept in global memory
    long long* dbg = (DBG && p.dbg) ? p.dbg + static_cast<size_t>(blockIdx.x) * 16 : nullptr;
    const long long c_t0 = (DBG && dbg) ? clock64() : 0;
    auto timed_wait = [&](uint64_t* bar, uint32_t ph, int slot) {
      if (DBG && dbg) { const long long t0 = clock64(); mbar_wait(bar, ph); if (lane == 0) dbg[slot] += clock64() - t0; }
      else mbar_wait(bar, ph);
    };

    constexpr uint32_t desc_hi_alias = (1u << 14) | (LAYOUT << 29);  // SBO = 0: every 8-row group is the same 8 rows
    const uint32_t zero_lo = (smem_u32(zero_a) & 0x3FFFFu) >> 4;
    auto issue = [&](uint32_t acc, uint32_t a_lo, uint32_t b_lo, int n) {
      const uint32_t idesc = idesc0 | (static_cast<uint32_t>(n >> 3) << 17);
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) umma_bf16_lohi(acc, a_lo + ks * 2, b_lo + ks * 2, desc_hi, idesc, 1u);
    };
    // D[128 x 128] = 0 * (the first rows of weight block `b_lo`): one K = 16 MMA, overwrite
    auto clear_acc = [&](uint32_t acc, uint32_t b_lo) {
      umma_bf16_lohi(acc, zero_lo, b_lo, desc_hi_alias, idesc0 | (static_cast<uint32_t>(128 >> 3) << 17), 0u);
    };
    auto run_ops = [&](const FoldOp* ops, int n_ops, uint32_t a_base_lo, uint32_t acc, int conv) {
      if (!p.w_resident) avail_ahead = 0;
      for (int o = 0; o < n_ops; ++o) {
        const int a_off = ops[o].a_off16, b_blk = ops[o].b_blk, nblk = ops[o].nblk, d_col = ops[o].d_col, rel = ops[o].rel;
        if (p.w_resident) {
          const int st = conv * p.k + b_blk;
          if (!w_seen) {
            for (int j = 0; j < nblk; ++j) timed_wait(&w_full[st + j], 0u, 1);
            tc_fence_after();
          }
          if (elect_one()) {
            if (o == 0) clear_acc(acc, wst_lo + static_cast<uint32_t>(st) * (WBLK >> 4));
            issue(acc + d_col, a_base_lo + a_off, wst_lo + static_cast<uint32_t>(st) * (WBLK >> 4), nblk * C);
          }
          __syncwarp();
        } else {
          while (avail_ahead < b_blk + nblk) {
            timed_wait(&w_full[avail_slot], avail_phase, 1);
            if (++avail_slot == p.stages) { avail_slot = 0; avail_phase ^= 1; }
            ++avail_ahead;
          }
          tc_fence_after();
          const int s0 = (base_slot + b_blk) % p.stages;
          const int n1 = (s0 + nblk <= p.stages) ? nblk : p.stages - s0;  // blocks before the ring wraps
          if (elect_one()) {
            if (o == 0) clear_acc(acc, wst_lo + static_cast<uint32_t>(s0) * (WBLK >> 4));
            issue(acc + d_col, a_base_lo + a_off, wst_lo + static_cast<uint32_t>(s0) * (WBLK >> 4), n1 * C);
            // a run that wraps around the ring is two MMA groups; the second half starts at slot 0
            if (n1 < nblk) issue(acc + d_col + n1 * C, a_base_lo + a_off, wst_lo, (nblk - n1) * C);
            if (rel) umma_commit(&w_empty[head_slot]);
          }
          __syncwarp();
          if (rel && ++head_slot == p.stages) head_slot = 0;
        }
      }
      if (!p.w_resident) {
        base_slot += p.k % p.stages;
        if (base_slot >= p.stages) base_slot -= p.stages;
      }
    };
    auto g1 = [&](int i) {
      const int buf = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      timed_wait(&d1_empty[buf], ph ^ 1, 2);
      timed_wait(&slab_full[buf], ph, 3);
      tc_fence_after();
      run_ops(p.ops1, p.n_ops1, slab_lo + static_cast<uint32_t>(buf) * (static_cast<uint32_t>(slab_bytes) >> 4),
              tmem_u + buf * ACC_COLS, 0);
      if (elect_one()) { umma_commit(&slab_empty[buf]); umma_commit(&d1_full[buf]); }
      __syncwarp();
    };
    auto g2 = [&](int i) {
      const int buf = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      const int tbi = p.t_bufs == 2 ? buf : 0;
      const uint32_t tph = p.t_bufs == 2 ? ph : static_cast<uint32_t>(i & 1);
      timed_wait(&t_full[tbi], tph, 4);
      timed_wait(&d2_empty[buf], ph ^ 1, 5);
      tc_fence_after();
      run_ops(p.ops2, p.n_ops2, t_lo + static_cast<uint32_t>(tbi) * (static_cast<uint32_t>(t_bytes) >> 4),
              tmem_u + (2 + buf) * ACC_COLS, 1);
      if (elect_one()) { umma_commit(&t_empty[tbi]); umma_commit(&d2_full[buf]); }
      __syncwarp();
    };
    if (n_my > 0) g1(0);
    for (int i = 0; i < n_my; ++i) {
      if (i + 1 < n_my) g1(i + 1);
      g2(i);
      w_seen = true;
    }
    if (DBG && dbg && lane == 0) { dbg[0] = clock64() - c_t0; dbg[6] = n_my; }
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
    const int c02m = (F - 1 - c02 / C) * C + (c02 % C);         // columns of the folded output row
    float* stg = staging + e * kFoldStageFloats;
    long long* dbg = (DBG && p.dbg && e == 0) ? p.dbg + static_cast<size_t>(blockIdx.x) * 16 : nullptr;
    const long long c_t0 = (DBG && dbg) ? clock64() : 0;
    auto timed_wait = [&](uint64_t* bar, uint32_t ph, int slot) {
      if (DBG && dbg) { const long long t0 = clock64(); mbar_wait(bar, ph); if (lane == 0) dbg[slot] += clock64() - t0; }
      else mbar_wait(bar, ph);
    };

    // E1: D1 -> (+b1, leaky_relu, bf16) -> xt phase slabs in UMMA layout; two 16-column items per warp
    auto e1 = [&](int i) {
      const int work = blockIdx.x + i * gridDim.x;
      int b, tile;
      decode_tile(p.rag, p.tiles_per_item, work, b, tile);
      const int g0 = tile * p.r_out - p.delta;
      const int buf = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      const int tbi = p.t_bufs == 2 ? buf : 0;
      const uint32_t tph = p.t_bufs == 2 ? ph : static_cast<uint32_t>(i & 1);
      timed_wait(&d1_full[buf], ph, 9);
      timed_wait(&t_empty[tbi], tph ^ 1, 10);  // the G2 that last read this xt buffer has retired
      tc_fence_after();
      const long long c_s = (DBG && dbg) ? clock64() : 0;
      uint8_t* tb = tbuf + tbi * t_bytes;
      const uint32_t tmem_acc = tmem_base + buf * ACC_COLS + lane_base;
#pragma unroll 1
      for (int j = sub; j < 8; j += 4) {
        uint32_t r[16];
        tmem_ld_32x16(tmem_acc + 16 * j, r);
        tmem_ld_wait();
        if (j + 4 >= 8) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&d1_empty[buf]);
        }
        const int q = (16 * j) / C, c0 = (16 * j) % C;
        const int h = F - 1 - q;
        const int tau = p.fdiv * blk1 + h * p.d1 + r1;  // xt row of (M row, phase h), tile-relative
        const int grow = g0 + tau;
        const bool inside = grow >= 0 && grow < p.L;
        uint32_t pk[8];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const float4 bb = *reinterpret_cast<const float4*>(p.bias1 + c0 + 4 * qq);
          const float v0 = inside ? lrelu_fast(__uint_as_float(r[4 * qq]) + bb.x, p.slope) : 0.f;
          const float v1 = inside ? lrelu_fast(__uint_as_float(r[4 * qq + 1]) + bb.y, p.slope) : 0.f;
          const float v2 = inside ? lrelu_fast(__uint_as_float(r[4 * qq + 2]) + bb.z, p.slope) : 0.f;
          const float v3 = inside ? lrelu_fast(__uint_as_float(r[4 * qq + 3]) + bb.w, p.slope) : 0.f;
          const uint2 u = pack_bf16x4(v0, v1, v2, v3);
          pk[2 * qq] = u.x; pk[2 * qq + 1] = u.y;
        }
        // c2 has dilation 1: xt row tau lives in phase slab tau mod F at row tau / F
        const int row = tau >> LOG2F;
        const uint32_t swz = (C == 64) ? (row & 7) : (C == 32) ? ((row >> 1) & 3) : ((row >> 2) & 1);
        const int ch = c0 >> 3;  // first 16-byte chunk of this item within the row
        uint8_t* rp = tb + (tau & (F - 1)) * p.xt_phase_bytes + row * ROWB;
        *reinterpret_cast<uint4*>(rp + (((ch) ^ swz) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(rp + (((ch + 1) ^ swz) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      fence_proxy_async();  // generic-proxy writes of xt -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_full[tbi]);
      if (DBG && dbg && lane == 0) dbg[13] += clock64() - c_s;
    };

    // E2 on the folded view: M row i of tile t is folded output row t*r_out/F + i (F time rows x C channels =
    // 128 fp32 = 512 contiguous bytes); tcgen05.ld hands lane l row l of the warp's 32 x 32 item.
    auto coords = [&](int i, int& b, int& q0) {
      const int work = blockIdx.x + i * gridDim.x;
      int tile;
      decode_tile(p.rag, p.tiles_per_item, work, b, tile);
      q0 = (tile * p.r_out) >> LOG2F;
    };
    auto prefetch_res = [&](int i) {  // lane 0 only; STAGED
      int b, q0;
      coords(i, b, q0);
      mbar_arrive_expect_tx(&res_bar[e], kFoldStageFloats * 4);
      tma_load_3d(stg, &map_res, &res_bar[e], c02m, q0 + quarter * 32, b);
    };
    // register transpose in 4-column units among the eight lanes that share lane & 3: afterwards lane l holds
    // columns 4*(l>>2)..+3 of rows 4*ii + (l&3), ii = 0..7 (eight lanes per 128-byte row segment)
    auto transpose_units = [&](uint32_t (&r)[32]) {
      const int unit = lane >> 2;
#pragma unroll
      for (int bit = 1; bit <= 4; bit <<= 1) {
        const bool up = (unit & bit) != 0;
#pragma unroll
        for (int c_lo = 0; c_lo < 8; ++c_lo) {
          if (c_lo & bit) continue;
          const int c_hi = c_lo | bit;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const uint32_t send = up ? r[4 * c_lo + w] : r[4 * c_hi + w];
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 4 * bit);
            if (up) r[4 * c_lo + w] = recv; else r[4 * c_hi + w] = recv;
          }
        }
      }
    };
    auto e2 = [&](int i) {
      int b, q0;
      coords(i, b, q0);
      const int buf = i & 1;
      timed_wait(&d2_full[buf], (i >> 1) & 1, 11);
      tc_fence_after();
      const long long c_s = (DBG && dbg) ? clock64() : 0;
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (2 + buf) * ACC_COLS + lane_base + c02, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d2_empty[buf]);
      const long long q_lim = static_cast<long long>(q0) + (p.r_out >> LOG2F);
      float v[8][4];
      if (E2M == 0) {
        timed_wait(&res_bar[e], i & 1, 12);
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          float4* sp = reinterpret_cast<float4*>(stg + lane * 32 + ((k4 ^ (lane & 7)) << 2));
          float4 t = *sp;
          t.x += __uint_as_float(r[4 * k4]); t.y += __uint_as_float(r[4 * k4 + 1]);
          t.z += __uint_as_float(r[4 * k4 + 2]); t.w += __uint_as_float(r[4 * k4 + 3]);
          *sp = t;
        }
        __syncwarp();
        const int c4 = lane & 7, rsub = lane >> 3;
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
          const int row = ii * 4 + rsub;
          const float4 t4 = *reinterpret_cast<const float4*>(stg + row * 32 + ((c4 ^ (row & 7)) << 2));
          v[ii][0] = t4.x; v[ii][1] = t4.y; v[ii][2] = t4.z; v[ii][3] = t4.w;
        }
        fence_proxy_async();  // our generic reads of the slot happen-before the next TMA write into it
        __syncwarp();
        if (lane == 0 && i + 1 < n_my) prefetch_res(i + 1);
        // (accumulator + residual) + bias here; epilogue_rows<.., true> does (accumulator + bias) + residual.
        // Both pair kernels use this order, so they stay bit-identical to each other.
        epilogue_rows<8, false>(p.epi, b, static_cast<long long>(q0) + quarter * 32 + rsub, 4, c02m + c4 * 4, v, q_lim);
      } else {
        if (E2M == 2) {
          timed_wait(&res_bar[e], i & 1, 12);
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 t = *reinterpret_cast<const float4*>(stg + lane * 32 + ((k4 ^ (lane & 7)) << 2));
            r[4 * k4] = __float_as_uint(t.x + __uint_as_float(r[4 * k4]));
            r[4 * k4 + 1] = __float_as_uint(t.y + __uint_as_float(r[4 * k4 + 1]));
            r[4 * k4 + 2] = __float_as_uint(t.z + __uint_as_float(r[4 * k4 + 2]));
            r[4 * k4 + 3] = __float_as_uint(t.w + __uint_as_float(r[4 * k4 + 3]));
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && i + 1 < n_my) prefetch_res(i + 1);
        }
        transpose_units(r);
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
          v[ii][0] = __uint_as_float(r[4 * ii]); v[ii][1] = __uint_as_float(r[4 * ii + 1]);
          v[ii][2] = __uint_as_float(r[4 * ii + 2]); v[ii][3] = __uint_as_float(r[4 * ii + 3]);
        }
        const long long qrow = static_cast<long long>(q0) + quarter * 32 + (lane & 3);
        const int ncol = c02m + (lane >> 2) * 4;
        if (E2M == 2) epilogue_rows<8, false>(p.epi, b, qrow, 4, ncol, v, q_lim);
        else epilogue_rows_res_first<8>(p.epi, b, qrow, 4, ncol, v, q_lim);
      }
      if (DBG && dbg && lane == 0) dbg[14] += clock64() - c_s;
    };
    if (n_my > 0 && STAGED && lane == 0) prefetch_res(0);
    if (n_my > 0) e1(0);
    for (int i = 0; i < n_my; ++i) {
      if (i + 1 < n_my) e1(i + 1);
      e2(i);
    }
    if (DBG && dbg && lane == 0) dbg[8] = clock64() - c_t0;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
size_t conv_fold_smem_bytes(int c, int slab_phase_bytes, int xt_phase_bytes, int t_bufs, int stages, int e2_mode) {
  const int f = 128 / c;
  return 1024 + kFoldZeroBytes + 2 * static_cast<size_t>(f) * slab_phase_bytes +
         static_cast<size_t>(t_bufs) * f * xt_phase_bytes + (e2_mode != 1 ? kFoldEpiWarps * kFoldStageFloats * 4 : 0) +
         static_cast<size_t>(stages) * c * c * 2 + (34 + 2 * stages) * 8 + 16;
}

template <int C, int E2M, bool DBG>
static cudaError_t launch_fold(const CUtensorMap& m, const CUtensorMap& mr, const TcFoldParams& p, size_t smem, int grid,
                               cudaStream_t st) {
  auto kern = conv_pair_fold_kernel<C, E2M, DBG>;
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

template <int C>
static cudaError_t launch_fold_c(const CUtensorMap& m, const CUtensorMap& mr, const TcFoldParams& p, size_t smem, int grid,
                                 cudaStream_t st) {
  if (p.dbg) {
    if (p.e2_mode == 0) return launch_fold<C, 0, true>(m, mr, p, smem, grid, st);
    if (p.e2_mode == 1) return launch_fold<C, 1, true>(m, mr, p, smem, grid, st);
    return launch_fold<C, 2, true>(m, mr, p, smem, grid, st);
  }
  if (p.e2_mode == 0) return launch_fold<C, 0, false>(m, mr, p, smem, grid, st);
  if (p.e2_mode == 1) return launch_fold<C, 1, false>(m, mr, p, smem, grid, st);
  return launch_fold<C, 2, false>(m, mr, p, smem, grid, st);
}

// mr: fp32 tile map over the folded residual view [B][L/F][128] (unused when p.e2_mode == 1)
cudaError_t launch_conv_pair_fold(int c, const CUtensorMap& m, const CUtensorMap& mr, const TcFoldParams& p, size_t smem,
                                  int grid, cudaStream_t st) {
  if (c == 64) return launch_fold_c<64>(m, mr, p, smem, grid, st);
  if (c == 32) return launch_fold_c<32>(m, mr, p, smem, grid, st);
  return cudaErrorInvalidValue;
}

}  // namespace hg
