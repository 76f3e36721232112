// conv_pair_tc.cu — one ResBlock1 pair as a single tcgen05 kernel (bf16 mode, C = 32 or 64):
//
//     xt = leaky_relu(c1(a) + b1)          dilated Conv1d, hifi/models.py:90-92
//     y  = c2(xt) + b2 + x                 Conv1d d=1 + residual, hifi/models.py:93-94
//     (+ MRF accumulate / divide / next operand copy, fused as in conv_tc.cu)
//
// Why: at 64/32 channels both convs are HBM-bound on their own; run separately
// the pair moves 16 B per element (a, xt written, xt read, x read, x written, a written), fused it
// moves 12 B and the first conv's whole kernel time disappears under the second's memory time.
//
// How: the first GEMM accumulates in TMEM; its epilogue applies bias + leaky_relu, rounds to bf16 and
// writes the tile straight into shared memory in the swizzled K-major layout a UMMA descriptor
// reads, so the second GEMM uses it as its A operand with the same tap-shifted descriptors.  A tile
// computes MT = MS*128 rows of xt and keeps the R = MT - (k-1) output rows that have their full c2
// footprint; tiles advance by R rows (96-98 % useful work).  Rows of xt outside [0,L) are forced to
// zero — c2 zero-pads xt, it does not see c1 evaluated beyond the sequence.
//
// Pipeline per CTA (persistent, grid-strided tiles; D1, D2 and the slab double-buffered, xt single- or
// double-buffered as shared memory allows):
//     tensor pipe : G1(0) G1(1) G2(0) G1(2) G2(1) G1(3) G2(2) ...
//     epilogue    : E1(0) E1(1) E2(0) E1(2) E2(1) ...
//       E1 = TMEM -> bias, leaky_relu, bf16 -> xt tile in smem;  E2 = TMEM -> transpose -> residual -> HBM
// Warp roles (608 threads): 0 weight producer | 1 MMA issuer + TMEM owner | 2 slab producer (TMA) |
// 3..18 epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace hg {

constexpr int kPairEpiWarps = 16;
constexpr int kPairThreads = (3 + kPairEpiWarps) * 32;
constexpr int kPairStageFloats = 32 * 32;        // per-warp transpose tile: 32 rows x 32 fp32 columns

// DBG = true is the HG_TC_DEBUG_TIMING build of the same kernel (cycle counters around every wait);
// the production instantiation carries none of it.
template <int C, int MS, bool DBG>
__global__ void __launch_bounds__(kPairThreads, 1)
conv_pair_tc_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_res,
                    const TcPairParams p) {
  constexpr int KC = C, N_T = C;
  constexpr int ROWB = KC * 2;
  constexpr int STAGE_BYTES = N_T * ROWB;
  constexpr int KSTEPS = KC / 16;
  constexpr uint32_t ACC_COLS = MS * N_T;  // 128
  constexpr uint32_t TMEM_COLS = 4 * ACC_COLS;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr uint32_t LAYOUT = (KC == 64) ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
  static_assert(TMEM_COLS == 512, "four accumulator buffers of 128 columns");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int slab_bytes = p.slab_rows * ROWB;  // multiple of 1024 (host rounds the rows)
  const int t_bytes = p.t_rows * ROWB;        // multiple of 1024
  uint8_t* slab = smem;                       // [2][slab_bytes]
  uint8_t* tbuf = slab + 2 * slab_bytes;      // [t_bufs][t_bytes]
  float* staging = reinterpret_cast<float*>(tbuf + p.t_bufs * t_bytes);  // [16][4 KB], 1024-aligned (TMA dst)
  uint8_t* wst = reinterpret_cast<uint8_t*>(staging + kPairEpiWarps * kPairStageFloats);  // [stages][STAGE_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(wst + p.stages * STAGE_BYTES);
  uint64_t* slab_full = bars;        // [2]
  uint64_t* slab_empty = bars + 2;   // [2]
  uint64_t* d1_full = bars + 4;      // [2]  G1 -> E1
  uint64_t* d1_empty = bars + 6;     // [2]  E1 -> G1
  uint64_t* t_full = bars + 8;       // [2]  E1 -> G2
  uint64_t* t_empty = bars + 10;     // [2]  G2 -> E1
  uint64_t* d2_full = bars + 12;     // [2]  G2 -> E2
  uint64_t* d2_empty = bars + 14;    // [2]  E2 -> G2
  uint64_t* res_bar = bars + 16;     // [16] residual tile landed in a warp's staging slot (TMA)
  uint64_t* w_full = bars + 32;      // [stages]
  uint64_t* w_empty = w_full + p.stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_empty + p.stages);
  float* bias1_s = reinterpret_cast<float*>(tmem_ptr_smem + 4);  // [C] conv 1 bias: E1 reads it with broadcast LDS

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h2 = (p.k - 1) >> 1;
  const int p1 = p.d1 * h2;
  // number of tiles this CTA owns
  const int n_my = p.total_work > static_cast<int>(blockIdx.x)
                       ? (p.total_work - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                       : 0;

  if (warp == 0 && lane == 0) { prefetch_tensormap(&map_in); prefetch_tensormap(&map_res); }
  // conv 1's bias is a weight, not a product of the previous launch: it may be read before griddepcontrol.wait.  (From
  // global memory the same-address float4 loads of E1 cost four L1 wavefronts each on a saturated data pipe.)
  if (warp == 2 && lane < C / 4) reinterpret_cast<float4*>(bias1_s)[lane] = __ldg(reinterpret_cast<const float4*>(p.bias1) + lane);
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&slab_full[i], 1); mbar_init(&slab_empty[i], 1);
        mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], kPairEpiWarps);
        mbar_init(&t_full[i], kPairEpiWarps); mbar_init(&t_empty[i], 1);
        mbar_init(&d2_full[i], 1); mbar_init(&d2_empty[i], kPairEpiWarps);
      }
      for (int s = 0; s < p.stages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
      for (int w = 0; w < kPairEpiWarps; ++w) mbar_init(&res_bar[w], 1);
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
    // ------------------------------------------------ weight producer: same order as the MMA issuer
    if (lane == 0 && n_my > 0) {
      int stage = 0; uint32_t phase = 0;
      auto load_conv = [&](const uint8_t* w) {
        for (int t = 0; t < p.k; ++t) {
          mbar_wait(&w_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&w_full[stage], STAGE_BYTES);
          bulk_load_1d(wst + stage * STAGE_BYTES, w + static_cast<size_t>(t) * STAGE_BYTES, STAGE_BYTES, &w_full[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      };
      if (p.w_resident) {
        load_conv(p.w1);
        load_conv(p.w2);
      } else {
        load_conv(p.w1);                       // G1(0)
        for (int i = 0; i < n_my; ++i) {
          if (i + 1 < n_my) load_conv(p.w1);   // G1(i+1)
          load_conv(p.w2);                     // G2(i)
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------ input slab producer (TMA)
    if (lane == 0) {
      for (int i = 0; i < n_my; ++i) {
        const int work = blockIdx.x + i * gridDim.x;
        int b, tile;
        decode_tile(p.rag, p.tiles_per_item, work, b, tile);
        const int m0 = tile * p.r_out;
        const int buf = i & 1;
        mbar_wait(&slab_empty[buf], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&slab_full[buf], slab_bytes);
        uint8_t* dst = slab + buf * slab_bytes;
        for (int bx = 0; bx < p.nboxes; ++bx)
          tma_load_3d(dst + bx * p.box_rows * ROWB, &map_in, &slab_full[buf], 0, m0 - h2 - p1 + bx * p.box_rows, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, N_T);
    constexpr uint32_t desc_hi = ((SBO >> 4) & 0x3FFFu) | (1u << 14) | (LAYOUT << 29);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t slab_lo = (smem_u32(slab) & 0x3FFFFu) >> 4;
    const uint32_t t_lo = (smem_u32(tbuf) & 0x3FFFFu) >> 4;
    const uint32_t wst_lo = (smem_u32(wst) & 0x3FFFFu) >> 4;
    int stage = 0; uint32_t wphase = 0;
    bool w_seen = false;  // resident mode: every stage has been waited for once
    // bring-up instrumentation (HG_TC_DEBUG_TIMING): cycles spent in each wait, kept in global memory
    long long* dbg = (DBG && p.dbg) ? p.dbg + static_cast<size_t>(blockIdx.x) * 16 : nullptr;
    const long long c_t0 = (DBG && dbg) ? clock64() : 0;
    auto timed_wait = [&](uint64_t* bar, uint32_t ph, int slot) {
      if (DBG && dbg) { const long long t0 = clock64(); mbar_wait(bar, ph); if (lane == 0) dbg[slot] += clock64() - t0; }
      else mbar_wait(bar, ph);
    };

    // one GEMM: k taps, A = base + tap*row_step rows, accumulate into `acc`
    auto gemm = [&](uint32_t a_base_lo, uint32_t tap_step_lo, uint32_t acc, int conv) {
      for (int t = 0; t < p.k; ++t) {
        const int st = p.w_resident ? conv * p.k + t : stage;
        if (!p.w_resident || !w_seen) {
          timed_wait(&w_full[st], p.w_resident ? 0u : wphase, 1);
          tc_fence_after();
        }
        const uint32_t b_lo = wst_lo + static_cast<uint32_t>(st) * (STAGE_BYTES >> 4);
        const uint32_t a_lo0 = a_base_lo + static_cast<uint32_t>(t) * tap_step_lo;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
            for (int ms = 0; ms < MS; ++ms)
              umma_bf16_lohi(acc + ms * N_T, a_lo0 + static_cast<uint32_t>(ms * ((128 * ROWB) >> 4) + ks * 2), b_lo + ks * 2,
                             desc_hi, idesc, (ks == 0 && t == 0) ? 0u : 1u);
          }
          if (!p.w_resident) umma_commit(&w_empty[stage]);
        }
        __syncwarp();
        if (!p.w_resident && ++stage == p.stages) { stage = 0; wphase ^= 1; }
      }
    };
    auto g1 = [&](int i) {
      const int buf = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      timed_wait(&d1_empty[buf], ph ^ 1, 2);
      timed_wait(&slab_full[buf], ph, 3);
      tc_fence_after();
      gemm(slab_lo + static_cast<uint32_t>(buf) * (static_cast<uint32_t>(slab_bytes) >> 4),
           (static_cast<uint32_t>(p.d1) * ROWB) >> 4, tmem_u + buf * ACC_COLS, 0);
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
      gemm(t_lo + static_cast<uint32_t>(tbi) * (static_cast<uint32_t>(t_bytes) >> 4), ROWB >> 4,
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
    float* stg = staging + e * kPairStageFloats;
    const int c4 = lane & 7, rsub = lane >> 3;
    // the one (sub-tile, 32-column) item of every tile this warp finishes in E2
    constexpr int E2_CH = N_T / 32;
    const int ms2 = sub / E2_CH, c02 = (sub - ms2 * E2_CH) * 32;
    const int n2 = c02 + c4 * 4;
    // bring-up instrumentation (HG_TC_DEBUG_TIMING): cycles spent in each wait, kept in global memory
    long long* dbg = (DBG && p.dbg && e == 0) ? p.dbg + static_cast<size_t>(blockIdx.x) * 16 : nullptr;
    const long long c_t0 = (DBG && dbg) ? clock64() : 0;
    auto timed_wait = [&](uint64_t* bar, uint32_t ph, int slot) {
      if (DBG && dbg) { const long long t0 = clock64(); mbar_wait(bar, ph); if (lane == 0) dbg[slot] += clock64() - t0; }
      else mbar_wait(bar, ph);
    };

    // The epilogue warps are INSTRUCTION-ISSUE bound (profiles/r2_pair_epilogue_issue_bound.md), so this code is
    // written for instruction count: tile coordinates are decoded once per tile, shared memory is addressed through
    // 32-bit window addresses with the swizzle folded into per-thread constants, global rows step by compile-time
    // strides from one 64-bit base, row validity is one 32-bit count, and the c2 bias lives in registers.
    struct TileAt { int b, m0; };  // item, first output row
    auto locate = [&](int i) {
      int b, tile;
      decode_tile(p.rag, p.tiles_per_item, blockIdx.x + i * gridDim.x, b, tile);
      return TileAt{b, tile * p.r_out};
    };
    const uint32_t tb_s = smem_u32(tbuf);
    const uint32_t stg_s = smem_u32(stg);
    const float slope1 = p.slope;
    const uint32_t bias1_a = smem_u32(bias1_s);

    // E1: D1 -> (+b1, leaky_relu, bf16) -> xt tile in UMMA layout; two (sub-tile, 16-column) items
    auto e1 = [&](int i, const TileAt& at) {
      const int buf = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      const int tbi = p.t_bufs == 2 ? buf : 0;
      const uint32_t tph = p.t_bufs == 2 ? ph : static_cast<uint32_t>(i & 1);
      timed_wait(&d1_full[buf], ph, 9);
      timed_wait(&t_empty[tbi], tph ^ 1, 10);  // the G2 that last read this xt buffer has retired
      tc_fence_after();
      const long long c_s = (DBG && dbg) ? clock64() : 0;
      const uint32_t tbs = tb_s + tbi * t_bytes;
      const uint32_t tmem_acc = tmem_base + buf * ACC_COLS + lane_base;
      constexpr int ITEMS = MS * (N_T / 16);
#pragma unroll 1
      for (int j = sub; j < ITEMS; j += 4) {
        const int ms = j / (N_T / 16), c0 = (j - ms * (N_T / 16)) * 16;
        uint32_t r[16];
        tmem_ld_32x16(tmem_acc + ms * N_T + c0, r);
        const int row = ms * 128 + quarter * 32 + lane;  // row of the xt tile; global time index m0 - h2 + row
        const bool inside = static_cast<unsigned>(at.m0 - h2 + row) < static_cast<unsigned>(p.L);
        const uint32_t bp = bias1_a + c0 * 4;
        const float4 b0 = lds128(bp), b1 = lds128(bp + 16), b2 = lds128(bp + 32), b3 = lds128(bp + 48);
        tmem_ld_wait();
        if (j + 4 >= ITEMS) {
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
        const uint32_t swz = (KC == 64) ? (row & 7) : ((row >> 1) & 3);
        const uint32_t ch = c0 >> 3;  // first 16-byte chunk of this item within the row
        const uint32_t rp = tbs + row * ROWB;
        sts128u(rp + ((ch ^ swz) << 4), u0.x, u0.y, u1.x, u1.y);
        sts128u(rp + (((ch + 1) ^ swz) << 4), u2.x, u2.y, u3.x, u3.y);
      }
      fence_proxy_async();  // generic-proxy writes of xt -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_full[tbi]);
      if (DBG && dbg && lane == 0) dbg[13] += clock64() - c_s;
    };

    // E2: D2 + residual -> smem transpose -> fused epilogue of c2.  The fp32 residual tile of this
    // warp's item is TMA-loaded straight into the warp's 4 KB staging slot (128B-swizzled box = the
    // staging swizzle) a whole tile ahead, so no thread ever waits on a global load: the accumulator
    // row is added to it in place, and after the transpose 8 lanes cover one 128-byte row segment.
    auto prefetch_res = [&](const TileAt& at) {  // lane 0 only
      mbar_arrive_expect_tx(&res_bar[e], kPairStageFloats * 4);
      tma_load_3d(stg, &map_res, &res_bar[e], c02, at.m0 + ms2 * 128 + quarter * 32, at.b);
      // the MRF running sum of the same tile is read by plain loads in the tail: pull its contiguous block
      // (r_out rows x C fp32) into L2 a tile ahead so that they do not wait on HBM
      if (e == 0 && p.epi.acc_in) {
        const long long first = static_cast<long long>(at.m0) * C;
        long long left = (p.epi.out_extent - first) * 4;
        const long long want = static_cast<long long>(p.r_out) * C * 4;
        if (left > want) left = want;
        if (left > 0)
          bulk_prefetch_l2(p.epi.acc_in + static_cast<long long>(at.b) * p.epi.out_batch_stride + first,
                           static_cast<uint32_t>(left) & ~15u);
      }
    };
    const float4 bias2 = __ldg(reinterpret_cast<const float4*>(p.epi.bias + n2));
    // lane-per-row view of the slot (the TMEM register layout) and its transpose (8 lanes per 128-byte row segment)
    const uint32_t row_s = stg_s + lane * 128, x7 = static_cast<uint32_t>(lane & 7) << 4;
    const uint32_t tr0_s = stg_s + rsub * 128 + ((c4 ^ rsub) << 4);               // rows rsub, rsub + 8, ...
    const uint32_t tr1_s = stg_s + (rsub + 4) * 128 + ((c4 ^ (rsub + 4)) << 4);   // rows rsub + 4, rsub + 12, ...
    const int row2 = ms2 * 128 + quarter * 32 + rsub;  // this thread's first row within the tile
    auto e2 = [&](int i, const TileAt& at, const TileAt* next) {
      const int buf = i & 1;
      // rows row2 + 4*ii (ii = 0..7): stored while inside the tile's own r_out rows and the sequence
      const int nv = (p.r_out < p.L - at.m0 ? p.r_out : p.L - at.m0) - row2;
      const long long off = static_cast<long long>(at.b) * p.epi.out_batch_stride +
                            static_cast<long long>(at.m0 + row2) * C + n2;
      timed_wait(&d2_full[buf], (i >> 1) & 1, 11);
      tc_fence_after();
      const long long c_s = (DBG && dbg) ? clock64() : 0;
      {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (2 + buf) * ACC_COLS + lane_base + ms2 * N_T + c02, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d2_empty[buf]);
        timed_wait(&res_bar[e], i & 1, 12);
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const uint32_t a = row_s + ((k4 << 4) ^ x7);
          float4 t = lds128(a);
          t = add4(t, make_float4(__uint_as_float(r[4 * k4]), __uint_as_float(r[4 * k4 + 1]), __uint_as_float(r[4 * k4 + 2]),
                                  __uint_as_float(r[4 * k4 + 3])));
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
      epilogue_tail8<4 * C>(p.epi, off, nv, bias2, v);
      if (DBG && dbg && lane == 0) dbg[14] += clock64() - c_s;
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
    if (DBG && dbg && lane == 0) dbg[8] = clock64() - c_t0;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
size_t conv_pair_smem_bytes(int c, int slab_rows, int t_rows, int t_bufs, int stages) {
  const int rowb = c * 2;
  return 1024 + 2 * static_cast<size_t>(slab_rows) * rowb + static_cast<size_t>(t_bufs) * t_rows * rowb +
         static_cast<size_t>(stages) * c * rowb + kPairEpiWarps * kPairStageFloats * 4 + (32 + 2 * stages) * 8 + 16 + 256;
}

template <int C, int MS, bool DBG>
static cudaError_t launch_pair(const CUtensorMap& m, const CUtensorMap& mr, const TcPairParams& p, size_t smem, int grid,
                               cudaStream_t st) {
  auto kern = conv_pair_tc_kernel<C, MS, DBG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, m, mr, p);
}

cudaError_t launch_conv_pair_tc(int c, const CUtensorMap& m, const CUtensorMap& mr, const TcPairParams& p, size_t smem,
                                int grid, cudaStream_t st) {
  if (c == 64) return p.dbg ? launch_pair<64, 2, true>(m, mr, p, smem, grid, st) : launch_pair<64, 2, false>(m, mr, p, smem, grid, st);
  if (c == 32) return p.dbg ? launch_pair<32, 4, true>(m, mr, p, smem, grid, st) : launch_pair<32, 4, false>(m, mr, p, smem, grid, st);
  return cudaErrorInvalidValue;
}

}  // namespace hg
