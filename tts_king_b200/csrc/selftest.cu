// selftest.cu — bring-up probe for the one non-standard thing conv_tc.cu relies on: reading the
// SAME swizzled K-major slab through UMMA descriptors whose start address is shifted by an
// arbitrary number of rows (the dilated-tap shift).  For each swizzle mode (128B / 64B), row shift
// and base_offset convention it runs D[128 x 64] = A[shift .. shift+128) * B^T on tcgen05 and
// compares with a host reference.  Operands are written with ordinary st.shared using the same
// address-bit XOR swizzle TMA applies, so this isolates the descriptor semantics from TMA.
#include <cuda_bf16.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "sm100_ptx.cuh"

namespace hg {

constexpr int ST_N = 64;
constexpr int ST_ROWS = 160;  // slab rows available (128 + up to 32 of shift)

// mode: 0 -> base_offset 0; 1 -> (addr>>7)&7; 2 -> (addr>>7)&3 (only differs for SW64)
template <int KC>
__global__ void __launch_bounds__(128) selftest_kernel(const __nv_bfloat16* __restrict__ A,  // [ST_ROWS][KC]
                                                       const __nv_bfloat16* __restrict__ Bm,  // [ST_N][KC]
                                                       int shift, int mode, float* __restrict__ D) {  // [128][ST_N]
  constexpr int ROWB = KC * 2;
  constexpr uint32_t LAYOUT = KC == 64 ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
  __shared__ __align__(1024) uint8_t sa[ST_ROWS * ROWB];
  __shared__ __align__(1024) uint8_t sb[ST_N * ROWB];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // swizzled fill: 16-byte chunk index XOR address bits [7,10) (SW128) or [7,9) (SW64)
  for (int e = threadIdx.x; e < ST_ROWS * KC; e += 128) {
    const int r = e / KC, kk = e % KC;
    const uint32_t lin = r * ROWB + (kk >> 3) * 16;
    const uint32_t x = KC == 64 ? ((lin >> 7) & 7u) : ((lin >> 7) & 3u);
    const uint32_t off = (lin ^ (x << 4)) + (kk & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sa + off) = A[e];
  }
  for (int e = threadIdx.x; e < ST_N * KC; e += 128) {
    const int r = e / KC, kk = e % KC;
    const uint32_t lin = r * ROWB + (kk >> 3) * 16;
    const uint32_t x = KC == 64 ? ((lin >> 7) & 7u) : ((lin >> 7) & 3u);
    const uint32_t off = (lin ^ (x << 4)) + (kk & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sb + off) = Bm[e];
  }
  fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc(&tmem_ptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, ST_N);
    const uint32_t a0 = smem_u32(sa) + shift * ROWB, b0 = smem_u32(sb);
#pragma unroll
    for (int ks = 0; ks < KC / 16; ++ks) {
      const uint32_t aa = a0 + ks * 32;
      const uint32_t bo = mode == 0 ? 0u : mode == 1 ? ((aa >> 7) & 7u) : ((aa >> 7) & 3u);
      umma_bf16(tmem, umma_smem_desc(aa, 0, 8 * ROWB, LAYOUT, bo), umma_smem_desc(b0 + ks * 32, 0, 8 * ROWB, LAYOUT, 0),
                idesc, ks ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < ST_N; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * ST_N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// ------------------------------------------------------------------------------------------------
// UMMA issue/throughput microbenchmark: one CTA, one elected lane issues `n_mma` back-to-back
// tcgen05.mma (M=128, N, K=16, bf16, SS operands from a 128B-swizzled slab) and the warp times the
// span from first issue to commit-arrival with clock64.  n_acc = how many TMEM accumulators the
// chain rotates over (1 = fully dependent accumulation).
template <int N>
__global__ void __launch_bounds__(128) umma_bench_kernel(int n_mma, int n_acc, int shift_rows, long long* out_cycles) {
  extern __shared__ uint8_t bsm_raw[];
  uint8_t* bsm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bsm_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = bsm;                   // 192 rows x 128 B
  uint8_t* sb = bsm + 192 * 128;       // N rows x 128 B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < (192 + N) * 32; e += 128) reinterpret_cast<uint32_t*>(bsm)[e] = 0x3c003c00u + e;
  fence_proxy_async();
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc(&tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    constexpr uint32_t desc_hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (UMMA_LAYOUT_SW128 << 29);
    const uint32_t a_lo = (smem_u32(sa) & 0x3FFFFu) >> 4, b_lo = (smem_u32(sb) & 0x3FFFFu) >> 4;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    long long t0 = clock64();
    if (elect_one()) {
      // tight issue loop: 8 MMAs per iteration, descriptors are loop-invariant adds of constants
      const uint32_t acc1 = tm + (n_acc > 1 ? N : 0), sh = static_cast<uint32_t>(shift_rows) * 8;
      for (int i = 0; i < n_mma; i += 8) {
        umma_bf16_lohi(tm, a_lo, b_lo, desc_hi, idesc, 1u);
        umma_bf16_lohi(acc1, a_lo + 2, b_lo + 2, desc_hi, idesc, 1u);
        umma_bf16_lohi(tm, a_lo + 4, b_lo + 4, desc_hi, idesc, 1u);
        umma_bf16_lohi(acc1, a_lo + 6, b_lo + 6, desc_hi, idesc, 1u);
        umma_bf16_lohi(tm, a_lo + sh, b_lo, desc_hi, idesc, 1u);
        umma_bf16_lohi(acc1, a_lo + sh + 2, b_lo + 2, desc_hi, idesc, 1u);
        umma_bf16_lohi(tm, a_lo + sh + 4, b_lo + 4, desc_hi, idesc, 1u);
        umma_bf16_lohi(acc1, a_lo + sh + 6, b_lo + 6, desc_hi, idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (lane == 0) *out_cycles = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N>
static void bench_one(std::string& rep, long long* d_out) {
  const int smem = 1024 + (192 + N) * 128;
  cudaFuncSetAttribute(umma_bench_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int n_acc : {1, 2}) {
    if (n_acc * N > 512) continue;
    for (int shift : {0, 1}) {
      long long c1 = 0, c2 = 0;
      for (int rep_i = 0; rep_i < 2; ++rep_i) {
        umma_bench_kernel<N><<<1, 128, smem>>>(256, n_acc, shift, d_out);
        cudaDeviceSynchronize();
        cudaMemcpy(&c1, d_out, 8, cudaMemcpyDeviceToHost);
        umma_bench_kernel<N><<<1, 128, smem>>>(1280, n_acc, shift, d_out);
        cudaDeviceSynchronize();
        cudaMemcpy(&c2, d_out, 8, cudaMemcpyDeviceToHost);
      }
      char b[160];
      snprintf(b, sizeof(b), "umma_bench M=128 N=%3d K=16 n_acc=%d row_shift=%d: %.1f cycles/MMA (256: %lld, 1280: %lld)\n", N,
               n_acc, shift, (c2 - c1) / 1024.0, c1, c2);
      rep += b;
    }
  }
}

static void appendf(std::string& s, const char* fmt, ...) {
  char b[256];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(b, sizeof(b), fmt, ap);
  va_end(ap);
  s += b;
}

template <int KC>
static int run_kc(std::string& rep) {
  std::vector<__nv_bfloat16> hA(ST_ROWS * KC), hB(ST_N * KC);
  std::vector<float> fA(ST_ROWS * KC), fB(ST_N * KC);
  uint32_t s = 12345u + KC;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 32768.0f - 1.0f; };
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16(rnd()); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16(rnd()); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA = nullptr, *dB = nullptr;
  float* dD = nullptr;
  if (cudaMalloc(&dA, hA.size() * 2) || cudaMalloc(&dB, hB.size() * 2) || cudaMalloc(&dD, 128 * ST_N * 4)) return -1;
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  std::vector<float> hD(128 * ST_N);
  int failures = 0;
  const int shifts[] = {0, 8, 16, 1, 2, 3, 4, 5, 7, 9, 15, 25, 30};
  for (int mode = 0; mode < 3; ++mode) {
    for (int sh : shifts) {
      cudaMemset(dD, 0xff, 128 * ST_N * 4);
      selftest_kernel<KC><<<1, 128>>>(dA, dB, sh, mode, dD);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        appendf(rep, "KC=%d mode=%d shift=%d CUDA error %s\n", KC, mode, sh, cudaGetErrorString(e));
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
        return failures + 1000;
      }
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < ST_N; ++n) {
          double ref = 0;
          for (int k = 0; k < KC; ++k) ref += static_cast<double>(fA[(m + sh) * KC + k]) * fB[n * KC + k];
          double d = fabs(ref - hD[m * ST_N + n]);
          if (!(d <= maxerr)) maxerr = d;  // NaN-propagating max
        }
      const bool ok = maxerr < 1e-3;
      if (!ok) ++failures;
      appendf(rep, "KC=%d mode=%d shift=%2d maxerr=%.3e %s\n", KC, mode, sh, maxerr, ok ? "OK" : "MISMATCH");
    }
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return failures;
}

int run_tcgen05_selftest(char* buf, size_t len) {
  std::string rep;
  int f = run_kc<64>(rep);
  if (f < 1000) f += run_kc<32>(rep);
  if (f < 1000) {
    long long* d_out = nullptr;
    if (cudaMalloc(&d_out, 8) == cudaSuccess) {
      bench_one<32>(rep, d_out); bench_one<64>(rep, d_out); bench_one<128>(rep, d_out); bench_one<256>(rep, d_out);
      cudaFree(d_out);
    }
  }
  if (buf && len) {
    strncpy(buf, rep.c_str(), len - 1);
    buf[len - 1] = 0;
  }
  return f;
}

}  // namespace hg
