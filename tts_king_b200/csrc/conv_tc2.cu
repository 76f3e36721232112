// conv_tc2.cu — the tap-shifted implicit-GEMM convolution of conv_tc.cu on CTA PAIRS
// (tcgen05.mma.cta_group::2), for the wide layers (C = 256 / 128) whose weights cannot stay
// resident in one SM's shared memory.
//
// Why: with one CTA per SM every SM streams the layer's whole weight tensor through its own shared
// memory for every 128/256-row tile (1.44 MB per tile at C = 256, k = 11).  Those TMA writes and the
// B-operand reads share the L1/smem data pipe with the A-operand reads and the epilogue, and measured
// MMA issue sits at ~89 cycles instead of the 64-cycle floor at N = 128.  A CTA pair
// computes M = 256 rows per MMA (128 from each SM's slab) against ONE copy of the weight tile that is
// split between the two SMs: per SM, weight smem writes, B-operand reads and L2 weight traffic halve.
//
// Protocol (both CTAs run the same code; rank 0 is the leader):
//   * full barriers live in the leader: its producer arms them with the byte count of BOTH CTAs'
//     loads, and both CTAs' TMA loads (.cta_group::2 form) complete_tx on the leader's barrier;
//   * the leader's MMA warp issues tcgen05.mma.cta_group::2 and releases slots / publishes
//     accumulators with multicast commits that arrive on both CTAs' barriers;
//   * each CTA's epilogue drains its own TMEM half (its 128 rows per sub-tile) and arrives on the
//     leader's acc_empty barrier (remote mbarrier arrive);
//   * cluster barriers bracket the kernel so no CTA touches a peer that is not initialised / gone.
// Epilogue = conv_tc.cu's TMA epilogue (same-length convs without MRF accumulate, bf16 operands).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace hg {

constexpr int kTc2EpiWarps = 16;
constexpr int kTc2Threads = (3 + kTc2EpiWarps) * 32;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* local, uint32_t rank) {
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(local)), "r"(rank));
  return a;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA loads whose completion is signalled on the barrier at `leader_bar` (a shared::cluster address)
__device__ __forceinline__ void tma2_load_3d(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1,
                                             int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
      :
      : "r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same offset in both CTAs once all prior tcgen05 ops have completed
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

template <int N_T, int MS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTc2Threads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ CUtensorMap map_res, const __grid_constant__ CUtensorMap map_x,
                const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_acc,
                const TcConvParams p) {
  constexpr int KC = 64, ROWB = 128;
  constexpr int HALF_N = N_T / 2;
  constexpr int STAGE_BYTES = HALF_N * ROWB;     // this CTA's half of a weight tile
  constexpr int KSTEPS = KC / 16;
  constexpr uint32_t ACC_COLS = MS * N_T;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int CHUNKS = N_T / 16;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int slab_bytes = p.slab_rows * ROWB;
  uint8_t* slab = smem;
  uint8_t* epi_smem = smem + ((p.nbuf * slab_bytes + 1023) & ~1023);
  uint8_t* wst = epi_smem + kTc2EpiWarps * p.epi_slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wst + p.stages * STAGE_BYTES);
  uint64_t* slab_full = bars;           // [4]  (leader's are the live ones)
  uint64_t* slab_empty = bars + 4;      // [4]
  uint64_t* acc_full = bars + 8;        // [2]
  uint64_t* acc_empty = bars + 10;      // [2]  (leader's)
  uint64_t* res_bar = bars + 12;        // [16][2]: two input tiles in flight per epilogue warp
  uint64_t* w_full = bars + 44;         // [stages] (leader's)
  uint64_t* w_empty = w_full + p.stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_empty + p.stages);
  // [N_T] bias: the epilogue's same-address float4 loads cost four L1 wavefronts each from global memory — as much as
  // the operand tile a c1 layer stores — and one as a shared-memory broadcast
  float* bias_s = reinterpret_cast<float*>(tmem_ptr_smem + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_w);
    if (p.has_res) prefetch_tensormap(&map_res);
    if (p.has_acc) prefetch_tensormap(&map_acc);
    if (p.has_x) prefetch_tensormap(&map_x);
    if (p.has_a) prefetch_tensormap(&map_ahi);
  }
  if (warp == 2) {  // a weight, not a product of the previous launch: may be read before griddepcontrol.wait
    for (int i = lane; i < N_T / 4; i += 32) reinterpret_cast<float4*>(bias_s)[i] = __ldg(reinterpret_cast<const float4*>(p.epi.bias) + i);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 4; ++i) { mbar_init(&slab_full[i], 1); mbar_init(&slab_empty[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * kTc2EpiWarps); }
      for (int s = 0; s < p.stages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
      for (int w = 0; w < 2 * kTc2EpiWarps; ++w) mbar_init(&res_bar[w], 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc2(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();

  // work item -> (N block, batch item, first row of THIS CTA's 128*MS rows); N blocks of one pair tile are consecutive
  // work items (the polyphase upsamplers: N_total = stride * C_out), so their slab reloads hit L2
  auto tile_coords = [&](int work, int& nblk, int& b, int& m0) {
    nblk = work % p.n_blocks;
    int tile;
    decode_tile(p.rag, p.tiles_per_item, work / p.n_blocks, b, tile);
    m0 = tile * (2 * MS * 128) + static_cast<int>(rank) * (MS * 128);
  };

  if (warp == 0) {
    // ------------------------------------------------ weight producer: this CTA's half of every tile
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int work = pair; work < p.total_work; work += npairs) {
        const int nblk = work % p.n_blocks;
        for (int c = 0; c < p.nc; ++c) {
          for (int t = 0; t < p.ntaps; ++t) {
            mbar_wait(&w_empty[stage], phase ^ 1);
            if (leader) mbar_arrive_expect_tx(&w_full[stage], 2 * STAGE_BYTES);
            // packed image rows: (((n_blk*nc + chunk)*ntaps + tap) * N_T + row); this CTA takes rows [rank*HALF_N, +HALF_N)
            tma2_load_2d(wst + stage * STAGE_BYTES, &map_w, map_to_cta(&w_full[stage], 0), 0,
                         ((nblk * p.nc + c) * p.ntaps + t) * N_T + static_cast<int>(rank) * HALF_N);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------ activation slab producer: this CTA's rows
    if (lane == 0) {
      int buf = 0; uint32_t phase = 0;
      for (int work = pair; work < p.total_work; work += npairs) {
        int nblk, b, m0;
        tile_coords(work, nblk, b, m0);
        for (int c = 0; c < p.nc; ++c) {
          mbar_wait(&slab_empty[buf], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&slab_full[buf], 2 * slab_bytes);
          const uint32_t lbar = map_to_cta(&slab_full[buf], 0);
          uint8_t* dst = slab + buf * slab_bytes;
          for (int bx = 0; bx < p.nboxes; ++bx)
            tma2_load_3d(dst + bx * p.box_rows * ROWB, &map_a, lbar, c * KC, m0 + p.min_off + bx * p.box_rows, b);
          if (++buf == p.nbuf) { buf = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, N_T);
      constexpr uint32_t desc_hi = ((SBO >> 4) & 0x3FFFu) | (1u << 14) | (UMMA_LAYOUT_SW128 << 29);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t slab_lo = (smem_u32(slab) & 0x3FFFFu) >> 4;
      const uint32_t wst_lo = (smem_u32(wst) & 0x3FFFFu) >> 4;
      const uint32_t slab_step = static_cast<uint32_t>(slab_bytes) >> 4;
      int stage = 0; uint32_t wphase = 0;
      int buf = 0; uint32_t sphase = 0;
      int it = 0;
      for (int work = pair; work < p.total_work; work += npairs, ++it) {
        const int ab = it & 1;
        mbar_wait(&acc_empty[ab], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_u + ab * ACC_COLS;
        for (int c = 0; c < p.nc; ++c) {
          mbar_wait(&slab_full[buf], sphase);
          tc_fence_after();
          for (int t = 0; t < p.ntaps; ++t) {
            const uint32_t tap_lo = (static_cast<uint32_t>(p.tap_row[t]) * ROWB) >> 4;
            mbar_wait(&w_full[stage], wphase);
            tc_fence_after();
            const uint32_t b_lo = wst_lo + static_cast<uint32_t>(stage) * (STAGE_BYTES >> 4);
            if (elect_one()) {
              const uint32_t a_lo0 = slab_lo + static_cast<uint32_t>(buf) * slab_step + tap_lo;
              const uint32_t first = (c | t) != 0 ? 1u : 0u;
#pragma unroll
              for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
                for (int ms = 0; ms < MS; ++ms)
                  umma2_bf16_lohi(tmem_acc + ms * N_T, a_lo0 + static_cast<uint32_t>(ms * ((128 * ROWB) >> 4) + ks * 2),
                                  b_lo + ks * 2, desc_hi, idesc, ks == 0 ? first : 1u);
              }
              umma2_commit_both(&w_empty[stage]);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; wphase ^= 1; }
          }
          if (elect_one()) umma2_commit_both(&slab_empty[buf]);
          __syncwarp();
          if (++buf == p.nbuf) { buf = 0; sphase ^= 1; }
        }
        if (elect_one()) umma2_commit_both(&acc_full[ab]);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------ TMA epilogue over this CTA's TMEM half
    const int e = warp - 3;
    const int quarter = warp & 3;
    const int sub = e >> 2;
    const uint32_t bias_a = smem_u32(bias_s);
    uint8_t* slot = epi_smem + e * p.epi_slot_bytes;
    // slot: in_bufs x ([residual in 2 KB] [MRF running sum in 2 KB]) [x out 2 KB] [operand out 1 KB], each only if used.
    // TWO input buffers where they fit: with one, every warp had a single 2 KB tile in flight and an item could not start before a
    // whole HBM round trip — 16 warps x 512 elements per ~2.5 us is the 0.24 ms the residual-carrying C = 128 layers
    // took for their 0.19 ms of HBM traffic.
    const int in_bytes = (p.has_res ? 2048 : 0) + (p.has_acc ? 2048 : 0);
    const uint32_t ib_mask = p.in_bufs == 2 ? 1u : 0u;
    float* xb = reinterpret_cast<float*>(slot + p.in_bufs * in_bytes);
    const bool has_in = p.has_res || p.has_acc;
    uint8_t* ab_hi = reinterpret_cast<uint8_t*>(xb) + (p.has_x ? 2048 : 0);
    constexpr int ITEMS = MS * CHUNKS;
    const uint32_t acc_empty_leader0 = map_to_cta(&acc_empty[0], 0);
    const uint32_t acc_empty_leader1 = map_to_cta(&acc_empty[1], 0);
    // lane 0's prefetch cursor: the next item (work, j) of this warp to request, and how many have been requested
    int pf_work = pair, pf_j = sub, pf_n = 0;
    auto prefetch_next = [&]() {  // lane 0 only
      if (pf_work >= p.total_work) return;
      int nblk, b, m0;
      tile_coords(pf_work, nblk, b, m0);
      const int ms = pf_j / CHUNKS, c0 = nblk * N_T + (pf_j - ms * CHUNKS) * 16;
      uint8_t* dst = slot + (pf_n & ib_mask) * in_bytes;
      uint64_t* bar = &res_bar[2 * e + (pf_n & ib_mask)];
      mbar_arrive_expect_tx(bar, in_bytes);
      if (p.has_res) tma_load_3d(dst, &map_res, bar, c0, m0 + ms * 128 + quarter * 32, b);
      if (p.has_acc) tma_load_3d(dst + (p.has_res ? 2048 : 0), &map_acc, bar, c0, m0 + ms * 128 + quarter * 32, b);
      ++pf_n;
      if (pf_j + 4 < ITEMS) pf_j += 4; else { pf_work += npairs; pf_j = sub; }
    };
    uint32_t n_item = 0;  // items consumed by this warp
    if (has_in && lane == 0 && sub < ITEMS) { prefetch_next(); if (p.in_bufs == 2) prefetch_next(); }
    const uint32_t swz64 = (lane >> 1) & 3, swz32 = (lane >> 2) & 1;
    int it = 0;
    for (int work = pair; work < p.total_work; work += npairs, ++it) {
      int nblk, b, m0;
      tile_coords(work, nblk, b, m0);
      const int ab = it & 1;
      const uint32_t tmem_acc = tmem_base + ab * ACC_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
      mbar_wait(&acc_full[ab], (it >> 1) & 1);
      tc_fence_after();
      bool released = false;
#pragma unroll 1
      for (int j = sub; j < ITEMS; j += 4) {
        const int ms = j / CHUNKS, c0 = (j - ms * CHUNKS) * 16;
        float v[16];
        {
          uint32_t r[16];
          tmem_ld_32x16(tmem_acc + ms * N_T + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        }
        if (j + 4 >= ITEMS) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(ab ? acc_empty_leader1 : acc_empty_leader0);
          released = true;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 bb = p.n_blocks == 1 ? lds128(bias_a + (c0 + 4 * q) * 4)
                                            : *reinterpret_cast<const float4*>(p.epi.bias + nblk * N_T + c0 + 4 * q);
          v[4 * q] += bb.x; v[4 * q + 1] += bb.y; v[4 * q + 2] += bb.z; v[4 * q + 3] += bb.w;
        }
        if (has_in) {
          const uint32_t ib = n_item & ib_mask;
          mbar_wait(&res_bar[2 * e + ib], (p.in_bufs == 2 ? n_item >> 1 : n_item) & 1);
          ++n_item;
          const float* rb = reinterpret_cast<const float*>(slot + ib * in_bytes);
          const float* accb = reinterpret_cast<const float*>(slot + ib * in_bytes + (p.has_res ? 2048 : 0));
          if (p.has_res) {  // x = xt + x   (hifi/models.py:94)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 t = *reinterpret_cast<const float4*>(rb + lane * 16 + ((q ^ swz64) << 2));
              v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
            }
          }
          if (p.has_acc) {  // xs += resblock(x)   (:193-195), same operand order as epilogue_rows
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 t = *reinterpret_cast<const float4*>(accb + lane * 16 + ((q ^ swz64) << 2));
              v[4 * q] = t.x + v[4 * q]; v[4 * q + 1] = t.y + v[4 * q + 1];
              v[4 * q + 2] = t.z + v[4 * q + 2]; v[4 * q + 3] = t.w + v[4 * q + 3];
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) prefetch_next();  // into the buffer just read: the item after next
        }
        if (p.epi.post_div == 3.f) {  // x = xs / num_kernels   (:196), V1's three kernels: common.cuh::div3_rn
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t = div3_rn(make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
          }
        } else if (p.epi.post_div > 0.f) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __fdiv_rn(v[i], p.epi.post_div);
        }
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        if (p.has_x) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(xb + lane * 16 + ((q ^ swz64) << 2)) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        if (p.has_a) {
          const float sl = p.epi.slope;
          uint32_t hi[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(lrelu_fast(v[2 * q], sl), lrelu_fast(v[2 * q + 1], sl));
            hi[q] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(ab_hi + lane * 32 + ((0 ^ swz32) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(ab_hi + lane * 32 + ((1 ^ swz32) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        }
        fence_proxy_async();
        __syncwarp();
        const int row0 = m0 + ms * 128 + quarter * 32;
        if (p.out_cmod) {
          // polyphase ConvTranspose1d (conv_tc.cu has the same epilogue): this item is one phase ph of 32 GEMM rows q;
          // output row q*s + ph + roff = (q + dq)*s + ph2 — a box of the phase-major 4-D view.  A negative start
          // coordinate faults in a TMA store: the one item per phase that begins before row 0 uses plain stores.
          const int n0 = nblk * N_T + c0;
          const int ph = n0 / p.out_cmod, col0 = n0 - ph * p.out_cmod;
          const int t = ph + p.out_roff;
          const int dq = t >= 0 ? t / p.out_rstride : -((-t + p.out_rstride - 1) / p.out_rstride);
          const int ph2 = t - dq * p.out_rstride;
          if (row0 + dq >= 0) {
            if (lane == 0) {
              if (p.has_x) tma_store_4d(&map_x, xb, col0, ph2, row0 + dq, b);
              if (p.has_a) tma_store_4d(&map_ahi, ab_hi, col0, ph2, row0 + dq, b);
              tma_store_commit();
            }
          } else {
            const long long orow = static_cast<long long>(row0 + lane) * p.out_rstride + t;  // this lane's output row
            const long long off = static_cast<long long>(b) * p.epi.out_batch_stride + orow * p.out_cmod + col0;
            if (orow >= 0 && (orow + 1) * p.out_cmod <= p.epi.out_extent) {
              if (p.has_x) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  *reinterpret_cast<float4*>(p.epi.out_x + off + 4 * q) = *reinterpret_cast<const float4*>(xb + lane * 16 + ((q ^ swz64) << 2));
              }
              if (p.has_a) {
#pragma unroll
                for (int q = 0; q < 2; ++q)
                  *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.epi.out_a0) + off + 8 * q) =
                      *reinterpret_cast<const uint4*>(ab_hi + lane * 32 + ((q ^ swz32) << 4));
              }
            }
            __syncwarp();  // the tiles are free again once every lane has read its row
          }
        } else if (lane == 0) {
          if (p.has_x) tma_store_3d(&map_x, xb, nblk * N_T + c0, row0, b);
          if (p.has_a) tma_store_3d(&map_ahi, ab_hi, nblk * N_T + c0, row0, b);
          tma_store_commit();
        }
      }
      if (!released) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(ab ? acc_empty_leader1 : acc_empty_leader0);
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still signal / be read
  if (warp == 1) tmem_dealloc2(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
size_t conv_tc2_smem_bytes(int n_t, int slab_rows, int nbuf, int stages, int epi_slot_bytes) {
  size_t slab = (static_cast<size_t>(nbuf) * slab_rows * 128 + 1023) & ~size_t(1023);
  return 1024 + slab + static_cast<size_t>(stages) * (n_t / 2) * 128 + static_cast<size_t>(kTc2EpiWarps) * epi_slot_bytes +
         (44 + 2 * stages) * 8 + 16 + 1024;
}

template <int N_T, int MS>
static cudaError_t launch_tc2(const CUtensorMap* maps, const TcConvParams& p, size_t smem, int grid, cudaStream_t st) {
  auto kern = conv_tc2_kernel<N_T, MS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kTc2Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
}

// maps: [0] operand (bf16), [1] packed weights (2-D), [2] residual, [3] x out, [4] a out.  grid must be even.
cudaError_t launch_conv_tc2(int n_t, int ms, const CUtensorMap* maps, const TcConvParams& p, size_t smem, int grid,
                            cudaStream_t st) {
  if (n_t == 256 && ms == 1) return launch_tc2<256, 1>(maps, p, smem, grid, st);
  if (n_t == 128 && ms == 2) return launch_tc2<128, 2>(maps, p, smem, grid, st);
  if (n_t == 128 && ms == 1) return launch_tc2<128, 1>(maps, p, smem, grid, st);
  return cudaErrorInvalidValue;
}

}  // namespace hg
