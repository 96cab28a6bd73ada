// umma_test.cu — single-CTA tcgen05 unit test behind vadc_debug_umma (used by
// tests/test_gpu_umma.py): validates the shared-memory descriptor encoding
// (K-major and MN-major SWIZZLE_128B), the instruction descriptor, TMEM
// allocation and the tcgen05.ld accumulator layout against numpy, for
// kind::tf32 and kind::f16 (bf16), before the fused kernels rely on them.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace vadc {
using namespace tc;

__device__ __forceinline__ void st16(uint8_t* p, float v, bool half) {
  if (half) *reinterpret_cast<__half*>(p) = __float2half_rn(v);
  else *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16(v);
}

// mode bit0: B is MN-major ([Kd, N] row-major) instead of K-major ([N, Kd] row-major)
// mode bit1: bf16 (kind::f16) instead of tf32
// mode bit2: A is MN-major ([Kd, 128] row-major) instead of K-major ([128, Kd])
// mode bit3: A (K-major) uses the un-swizzled core-matrix layout
//            [k-chunk of 16 B][16 row groups][8 rows x 16 B]  (LBO 2048, SBO 128)
// mode bit4: (with bit1) A is fp16 while B stays bf16  -- mixed a_format / b_format in one kind::f16 MMA
// mode bit5: (with bit1) B is fp16 while A stays bf16
// mode bit6: (with bit0|bit1) B is given as [Kd, 2N]; the MMA uses only columns [N, 2N) through a
//            +N*2-byte start-address offset inside the 128-byte swizzle row (N == 32)
__global__ void __launch_bounds__(128)
umma_test_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out,
                 int N, int Kd, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const bool b_mn = mode & 1, bf = mode & 2, a_mn = mode & 4, a_ns = mode & 8;
  const bool a_h = mode & 16, b_h = mode & 32, b_sub = mode & 64;
  const int NB = b_sub ? 2 * N : N;          // columns of B present in shared memory
  const int es = bf ? 2 : 4;                 // element size
  const int epr = 128 / es;                  // elements per 128-byte row
  const int tid = threadIdx.x, warp = tid >> 5;
  // ---- operand A ----
  uint8_t* sA = smem;
  uint32_t a_bytes;
  if (a_ns) {    // K-major, no swizzle: core matrices of 8 rows x 16 bytes
    const int epc = 16 / es;                 // elements per 16-byte chunk
    a_bytes = 128 * Kd * es;
    for (int i = tid; i < 128 * Kd; i += 128) {
      int r = i / Kd, k = i % Kd;
      uint32_t off = (k / epc) * 2048 + (r / 8) * 128 + (r % 8) * 16 + (k % epc) * es;
      if (bf) st16(sA + off, A[i], a_h);
      else *reinterpret_cast<float*>(sA + off) = A[i];
    }
  } else if (!a_mn) {   // K-major: Kd/epr blocks of [128 rows x 128 B]
    a_bytes = (Kd / epr) * 128 * 128;
    for (int i = tid; i < 128 * Kd; i += 128) {
      int r = i / Kd, k = i % Kd;
      uint32_t off = (k / epr) * (128 * 128) + sw128(r, (k % epr) * es);
      if (bf) st16(sA + off, A[i], a_h);
      else *reinterpret_cast<float*>(sA + off) = A[i];
    }
  } else {       // MN-major: 128/epr blocks of [Kd rows(k) x 128 B]; A given as [Kd, 128]
    a_bytes = (128 / epr) * Kd * 128;
    for (int i = tid; i < 128 * Kd; i += 128) {
      int k = i / 128, m = i % 128;
      uint32_t off = (m / epr) * (Kd * 128) + sw128(k, (m % epr) * es);
      if (bf) st16(sA + off, A[i], a_h);
      else *reinterpret_cast<float*>(sA + off) = A[i];
    }
  }
  uint8_t* sB = smem + ((a_bytes + 1023) & ~1023u);
  if (!b_mn) {   // K-major: Kd/epr blocks of [N rows x 128 B]; B given as [N, Kd]
    for (int i = tid; i < N * Kd; i += 128) {
      int r = i / Kd, k = i % Kd;
      uint32_t off = (k / epr) * (N * 128) + sw128(r, (k % epr) * es);
      if (bf) st16(sB + off, B[i], b_h);
      else *reinterpret_cast<float*>(sB + off) = B[i];
    }
  } else {       // MN-major: N/epr blocks of [Kd rows(k) x 128 B]; B given as [Kd, NB]
    for (int i = tid; i < NB * Kd; i += 128) {
      int k = i / NB, n = i % NB;
      uint32_t off = (n / epr) * (Kd * 128) + sw128(k, (n % epr) * es);
      if (bf) st16(sB + off, B[i], b_h);
      else *reinterpret_cast<float*>(sB + off) = B[i];
    }
  }
  uint32_t ncols = 32;
  while (ncols < (uint32_t)N) ncols <<= 1;
  if (warp == 0) tmem_alloc(&tmem_base, ncols);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    uint32_t idesc = instr_desc(bf ? kFmtBF16 : kFmtTF32, 128, N, a_mn ? 1 : 0, b_mn ? 1 : 0);
    if (a_h) idesc &= ~(7u << 7);            // a_format = F16
    if (b_h) idesc &= ~(7u << 10);           // b_format = F16
    const int kstep = bf ? 16 : 8;           // elements per MMA along the contraction
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int k = 0; k < Kd; k += kstep) {
      uint64_t ad, bd;
      if (a_ns) ad = smem_desc_noswz(a0 + (k / (16 / es)) * 2048, 2048, 128);
      else if (!a_mn) ad = smem_desc_sw128(a0 + (k / epr) * (128 * 128) + (k % epr) * es, 0, 1024);
      else ad = smem_desc_sw128(a0 + (k / 8) * 1024, Kd * 128, 1024);
      if (!b_mn) bd = smem_desc_sw128(b0 + (k / epr) * (N * 128) + (k % epr) * es, 0, 1024);
      else bd = smem_desc_sw128(b0 + (k / 8) * 1024 + (b_sub ? N * es : 0), Kd * 128, 1024);
      if (bf) mma_f16(tmem, ad, bd, idesc, k > 0); else mma_tf32(tmem, ad, bd, idesc, k > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 32 && c0 + j < N; ++j) out[(size_t)tid * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}
}  // namespace vadc

extern "C" int vadc_debug_umma(const float* A, const float* B, float* out, int N, int Kd, int mode,
                               void* stream) {
  using namespace vadc;
  VADC_REQUIRE(N >= 8 && N <= 256 && (N % 8) == 0 && Kd >= 32 && (Kd % 64) == 0, VADC_ERR_BAD_SHAPE);
  VADC_REQUIRE(!(mode & 1) || (N % 64) == 0 || N == 32, VADC_ERR_BAD_SHAPE);
  VADC_REQUIRE(!(mode & 64) || ((mode & 3) == 3 && N == 32), VADC_ERR_BAD_SHAPE);
  size_t smem = (size_t)(128 + ((mode & 64) ? 2 * N : N)) * Kd * 4 + 4096;
  VADC_REQUIRE(smem <= 200 * 1024, VADC_ERR_UNSUPPORTED);
  VADC_CUDA(cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_test_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(A, B, out, N, Kd, mode);
  VADC_CHECK_LAUNCH("umma_test_kernel");
  return VADC_OK;
}

// ---------------------------------------------------------------------------
// issue-rate micro-benchmark: one thread issues `reps` x 8 tcgen05.mma (kind::f16, bf16) of one
// shape back to back on uninitialised shared memory and waits for the commit; out[0] = cycles.
// (The fused kernels' tile shapes were chosen from these numbers: scripts/umma_bench.py.)
// ---------------------------------------------------------------------------
namespace vadc {
__global__ void __launch_bounds__(128)
umma_bench_kernel(int M, int N, int a_mn, int b_mn, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;   // bf16 1.0
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = instr_desc(kFmtBF16, M, N, a_mn, b_mn);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        // K-major: 32-byte K slices of [rows x 128 B] blocks; MN-major: 16-row (2 KB) K slices, 8 KB between 64-wide blocks
        const uint64_t ad = a_mn ? smem_desc_sw128(a0 + (k & 3) * 2048, 8192, 1024) : smem_desc_sw128(a0 + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024);
        const uint64_t bd = b_mn ? smem_desc_sw128(b0 + (k & 3) * 2048, 8192, 1024) : smem_desc_sw128(b0 + (k >> 2) * 32768 + (k & 3) * 32, 0, 1024);
        mma_f16(tmem, ad, bd, idesc, (r | k) ? 1u : 0u);
      }
    }
    const long long t1 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t2 - t0;
    out[1] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace vadc

extern "C" int vadc_debug_umma_bench(int M, int N, int a_mn, int b_mn, int reps, long long* out, void* stream) {
  using namespace vadc;
  VADC_REQUIRE((M == 64 || M == 128) && N >= 8 && N <= 256 && (N % 8) == 0 && reps > 0, VADC_ERR_BAD_SHAPE);
  const size_t smem = 161 * 1024 + 1024;
  VADC_CUDA(cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_bench_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(M, N, a_mn, b_mn, reps, out);
  VADC_CHECK_LAUNCH("umma_bench_kernel");
  return VADC_OK;
}
