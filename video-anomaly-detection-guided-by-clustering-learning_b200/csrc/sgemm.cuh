// sgemm.cuh — generic fp32 CUDA-core tile GEMM with a functor epilogue.
// The SIMT (VADC_IMPL_SIMT) kernels for every shape the tcgen05 path does not
// cover are built from this: C[m,n] = sum_k opA[m,k] * opB[k,n], operands
// addressed through (row stride, col stride) so any transposition / batching
// of the caller's tensors is expressed without a copy.
#pragma once
#include "common.cuh"

namespace vadc {

struct Operand {
  const float* p;
  long long rs, cs;   // element (r,c) at p[r*rs + c*cs]
};

template <int BM, int BN, int BK, int TM, int TN>
struct SgemmCfg {
  static constexpr int kThreads = (BM / TM) * (BN / TN);
  static constexpr int kPadA = 4, kPadB = 4;
  static constexpr int kSmemFloats = BK * (BM + kPadA) + BK * (BN + kPadB);
};

// One CTA computes the BM x BN tile at (m0, n0) over k in [k0, k1).
// acc[i][j] is the value for row m0 + ty*TM + i, col n0 + tx*TN + j.
template <int BM, int BN, int BK, int TM, int TN>
__device__ __forceinline__ void sgemm_tile(int M, int N, Operand A, Operand B, int m0, int n0,
                                           int k0, int k1, float (&acc)[TM][TN], float* smem) {
  using Cfg = SgemmCfg<BM, BN, BK, TM, TN>;
  constexpr int NT = Cfg::kThreads;
  constexpr int LDA = BM + Cfg::kPadA, LDB = BN + Cfg::kPadB;
  float* As = smem;                 // [BK][LDA]
  float* Bs = smem + BK * LDA;      // [BK][LDB]
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const bool a_kfast = (A.cs == 1);  // contraction index contiguous in memory
  const bool b_kfast = (B.rs == 1);
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int kb = k0; kb < k1; kb += BK) {
    // ---- stage A tile: As[k][m] ----
#pragma unroll
    for (int it = 0; it < (BM * BK + NT - 1) / NT; ++it) {
      int idx = tid + it * NT;
      if (idx < BM * BK) {
        int m, k;
        if (a_kfast) { m = idx / BK; k = idx % BK; } else { k = idx / BM; m = idx % BM; }
        int gm = m0 + m, gk = kb + k;
        float v = 0.f;
        if (gm < M && gk < k1) v = __ldg(A.p + (long long)gm * A.rs + (long long)gk * A.cs);
        As[k * LDA + m] = v;
      }
    }
    // ---- stage B tile: Bs[k][n] ----
#pragma unroll
    for (int it = 0; it < (BN * BK + NT - 1) / NT; ++it) {
      int idx = tid + it * NT;
      if (idx < BN * BK) {
        int n, k;
        if (b_kfast) { n = idx / BK; k = idx % BK; } else { k = idx / BN; n = idx % BN; }
        int gn = n0 + n, gk = kb + k;
        float v = 0.f;
        if (gn < N && gk < k1) v = __ldg(B.p + (long long)gk * B.rs + (long long)gn * B.cs);
        Bs[k * LDB + n] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(&As[k * LDA + ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[k * LDB + tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
}

// Generic kernel: grid = (ceil(N/BN), ceil(M/BM), batches*splits).  The epilogue
// functor receives (batch, split, m, n, value) for every in-range element and
// carries its own output pointers / per-batch strides.
template <int BM, int BN, int BK, int TM, int TN, class Epi>
__global__ void __launch_bounds__(SgemmCfg<BM, BN, BK, TM, TN>::kThreads)
sgemm_kernel(int M, int N, int Kd, Operand A, Operand B, long long batchA, long long batchB,
             int splits, Epi epi) {
  using Cfg = SgemmCfg<BM, BN, BK, TM, TN>;
  __shared__ __align__(16) float smem[Cfg::kSmemFloats];
  const int batch = blockIdx.z / splits, split = blockIdx.z % splits;
  A.p += (long long)batch * batchA;
  B.p += (long long)batch * batchB;
  // split the contraction range in BK-aligned chunks
  int kper = ((Kd + splits - 1) / splits + BK - 1) / BK * BK;
  int k0 = split * kper, k1 = min(Kd, k0 + kper);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[TM][TN];
  sgemm_tile<BM, BN, BK, TM, TN>(M, N, A, B, m0, n0, k0, k1, acc, smem);
  const int tx = threadIdx.x % (BN / TN), ty = threadIdx.x / (BN / TN);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n < N) epi(batch, split, m, n, acc[i][j]);
    }
  }
}

template <int BM, int BN, class Epi>
inline cudaError_t launch_sgemm(int M, int N, int Kd, Operand A, Operand B, long long batchA,
                                long long batchB, int batches, int splits, Epi epi,
                                cudaStream_t st) {
  constexpr int BK = 16, TM = 4, TN = 4;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batches * splits);
  sgemm_kernel<BM, BN, BK, TM, TN, Epi><<<grid, SgemmCfg<BM, BN, BK, TM, TN>::kThreads, 0, st>>>(
      M, N, Kd, A, B, batchA, batchB, splits, epi);
  count_launch();
  return cudaGetLastError();
}

// pick the tile shape by the width of the output
template <class Epi>
inline cudaError_t sgemm_auto(int M, int N, int Kd, Operand A, Operand B, long long batchA,
                              long long batchB, int batches, int splits, Epi epi, cudaStream_t st) {
  if (N <= 32) return launch_sgemm<128, 32, Epi>(M, N, Kd, A, B, batchA, batchB, batches, splits, epi, st);
  if (M <= 32) return launch_sgemm<32, 128, Epi>(M, N, Kd, A, B, batchA, batchB, batches, splits, epi, st);
  return launch_sgemm<64, 64, Epi>(M, N, Kd, A, B, batchA, batchB, batches, splits, epi, st);
}

}  // namespace vadc
