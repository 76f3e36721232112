// plan.h — host-side plan structures of libhifigan_b200 (internal).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <list>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/hifigan_b200.h"
#include "common.cuh"

namespace hg {

enum LayerKind { L_CONV = 0, L_CONVT = 1, L_POST = 2 };

// One GEMM-shaped layer (or conv_post).  See conv_tc.cu for the mapping.
struct Layer {
  std::string name;  // state_dict prefix
  int kind = L_CONV;
  int cin = 0, cout = 0, k = 0, dil = 1, stride = 1, pad = 0;
  // channel counts of the state_dict tensors when this layer runs zero-padded to cin / cout channels in
  // memory (the bf16 schedule of stages narrower than 16 channels); 0 = not padded
  int cin_w = 0, cout_w = 0;
  // > 0: the output tensor is only n_store channels wide although the GEMM computes cout (zero weights beyond):
  // conv-stack layers whose width is not a multiple of 32 (the 80-mel outputs of the FastSpeech2 tail)
  int n_store = 0;
  // GEMM view
  int ntaps = 0;
  int tap_off[kMaxTaps] = {0};  // input row = GEMM row + tap_off
  int n_total = 0;              // cout (conv) or stride*cout (convT)
  // tensor-core tiling (tc == false: CUDA-core path only)
  bool tc = false;
  int cin_pad = 0, kc = 0, nc = 0, n_tile = 0, n_blocks = 0;
  // device weights
  bool loaded = false;
  uint8_t* w_hi = nullptr;   // packed bf16 tiles (tc)
  uint8_t* w_lo = nullptr;
  float* w_ffma = nullptr;   // [tap][cin][n_total]
  float* bias = nullptr;     // [n_total] (bias[n % cout]); conv_post: host scalar below
  float* bias_fold = nullptr;  // [128] bias[n % cout], C = 32 / 64 convs: epilogue of the time-folded pair kernel
  float* w_post = nullptr;   // conv_post: [7][C]
  std::vector<float> w_post_host;  // same, host copy (passed by value in the kernel-parameter bank when C = 32)
  float bias_post = 0.f;
};

struct TcTiling {
  int ms = 1, stages = 2, nbuf = 1, slab_rows = 0, box_rows = 0, nboxes = 1, min_off = 0;
  bool resident = false;  // all weight tiles stay in shared memory
  size_t smem = 0;
};

using MapKey = std::tuple<const void*, int, int, int, int, int>;  // ptr, L, B, cpitch, kc, box_rows

// Bounded least-recently-used cache of encoded TMA descriptors.  Keys hold raw pointers (workspace
// and caller tensors), so a server whose requests vary in T or whose output tensors are freshly
// allocated keeps producing new keys: the oldest entries fall out one at a time, the hot ones (the
// current shapes, the static weight maps) stay.
class TensorMapCache {
 public:
  explicit TensorMapCache(size_t capacity = 2048) : cap_(capacity) {}
  bool get(const MapKey& k, CUtensorMap* out) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = map_.find(k);
    if (it == map_.end()) return false;
    order_.splice(order_.begin(), order_, it->second.second);  // most recently used first
    *out = it->second.first;
    return true;
  }
  void put(const MapKey& k, const CUtensorMap& m) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = map_.find(k);
    if (it != map_.end()) {
      it->second.first = m;
      order_.splice(order_.begin(), order_, it->second.second);
      return;
    }
    order_.push_front(k);
    map_.emplace(k, std::make_pair(m, order_.begin()));
    while (map_.size() > cap_) {
      map_.erase(order_.back());
      order_.pop_back();
    }
  }
  size_t size() {
    std::lock_guard<std::mutex> g(mu_);
    return map_.size();
  }

 private:
  size_t cap_;
  std::mutex mu_;
  std::list<MapKey> order_;
  std::map<MapKey, std::pair<CUtensorMap, std::list<MapKey>::iterator>> map_;
};

}  // namespace hg

namespace hg { constexpr int kMaxSideStreams = 7; }  // resblock kernels per stage - 1 (HG_MAX_KERNELS = 8)

struct HgPlan {
  HgConfig cfg;
  int device = 0;
  int sm_count = 148;
  bool finalized = false;
  std::vector<hg::Layer> layers;
  // bf16 schedule when a stage has fewer than 16 channels: the same layers with that stage's tensors padded
  // to 16 channels (zero weights / bias), so that it runs on the tensor-core kernels; empty otherwise
  std::vector<hg::Layer> layers_pad;
  // latency schedule: the 256-channel convs again with 64-column N tiles (4 N blocks), used when a launch has too
  // few M tiles to occupy the GPU (one short utterance); entries for other layers stay unloaded
  std::vector<hg::Layer> layers_small;
  std::map<std::string, int> by_name;
  int desc_mode = 0;  // measured on B200 (selftest.cu): UMMA swizzle phase comes from absolute smem address bits
  int force_ms = 0, force_stages = 0;
  int ctas_per_sm = 1;  // persistent grid = min(work, SMs * ctas_per_sm)
  bool fuse_pairs = true;  // HG_FUSE_PAIRS=0: never use the fused ResBlock-pair kernel
  bool fold_pairs = true;  // HG_FOLD=0: fused pairs run on conv_pair_tc.cu (N = C) instead of conv_pair_fold.cu (N = 128)
  bool fuse_blocks = false;  // HG_CHAIN=1: fuse a whole k = 3 ResBlock1 into one launch (conv_chain_tc.cu; slower than its pairs, see api.cu)
  bool tile_alternate = true;  // consecutive launches walk their tiles in opposite directions (HG_TILE_ORDER=0: all forward)
  bool fold_force = false;  // HG_FOLD=2: ... and on conv_pair_fold.cu wherever it applies, not only where it is faster
  bool epi_tma = true;     // HG_EPI_TMA=0: always use the generic (LSU) epilogue in conv_tc
  bool epi_tma_convt = true;  // HG_EPI_TMA_CONVT=0: the polyphase upsamplers keep the generic epilogue
  int tc2_in_bufs = 2;     // HG_TC2_INBUFS=1: one residual tile in flight per epilogue warp of the CTA-pair kernel (round 1)
  bool tc2_convt = true;   // HG_TC2_CONVT=0: upsamplers with several N blocks stay on the single-CTA kernel
  bool use_tc2 = true;     // HG_TC2=0: never use the CTA-pair (cta_group::2) kernel for the 256/128-channel convs
  bool force_ffma = false;  // HG_FORCE_FFMA=1: route every layer to the CUDA-core kernel
  // hg_stack_create: a plain chain of same-length Conv1d layers (no generator schedule); act/slope[i] is
  // the activation between layer i and layer i+1
  bool is_stack = false;
  std::vector<int> stack_act;
  std::vector<float> stack_slope;
  hg::TensorMapCache maps;
  // short inputs: the ResBlocks of a stage whose activations hold at most concurrent_elems elements run on separate
  // streams (they only meet in the MRF sum), api.cu::hg_forward.  Streams / events are created on first use and owned by the plan; `side_mutex`
  // serialises the enqueue of two such forwards on one plan (events are re-recorded per stage).
  long long concurrent_elems = 2560 * 1024;  // HG_CONCURRENT_KELEMS (in units of 1024 elements; 0 switches it off)
  cudaStream_t side_stream[hg::kMaxSideStreams] = {};
  cudaEvent_t ev_fork = nullptr, ev_done[hg::kMaxSideStreams + 1] = {};
  std::mutex side_mutex;
};
