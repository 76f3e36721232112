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
  if (buf && len) {
    strncpy(buf, rep.c_str(), len - 1);
    buf[len - 1] = 0;
  }
  return f;
}

}  // namespace hg
