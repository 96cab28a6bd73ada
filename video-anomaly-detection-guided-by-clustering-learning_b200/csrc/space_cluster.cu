// space_cluster.cu — Space_EuclidDistance_Assign_Module forward / backward (C3)
// model/cluster.py:127-149.  Per-channel spatial-map clustering: for every
// channel c the M = B*D frames are P-vectors (P = H*W) compared with K
// centroids of that channel: a batched [M,P] x [P,K] contraction over C
// batches.
#include <cuda_bf16.h>
#include "common.cuh"
#include "sgemm.cuh"
#include "tc_gemm.cuh"
#include "rows.cuh"
#include "cluster.h"

namespace vadc {

// LayerNorm + transpose: x [T, C] tokens -> zt [C, T].  A block normalises 32
// tokens (one warp per 4 tokens), parks them in shared memory and writes 32
// consecutive tokens per channel (128-byte segments) so both sides coalesce.
__global__ void __launch_bounds__(256)
ln_transpose_kernel(const float* __restrict__ x, const float* __restrict__ w,
                    const float* __restrict__ b, long long T, int C, float eps,
                    float* __restrict__ zt, float* __restrict__ mu, float* __restrict__ rstd,
                    __nv_bfloat16* __restrict__ terms /* optional: the three bf16 terms of zt, [3][C*T] */) {
  extern __shared__ float tile[];            // [32][C+1]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long t0 = (long long)blockIdx.x * 32;
  const int ld = C + 1;
  for (int rr = wid; rr < 32; rr += 8) {
    long long row = t0 + rr;
    if (row >= T) break;
    const float* xr = x + row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) { float v = __ldg(xr + c); tile[rr * ld + c] = v; s += v; }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { float d = tile[rr * ld + c] - mean; q += d * d; }
    const float rs = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
    for (int c = lane; c < C; c += 32)
      tile[rr * ld + c] = (tile[rr * ld + c] - mean) * rs * __ldg(w + c) + __ldg(b + c);
    if (lane == 0) { mu[row] = mean; rstd[row] = rs; }
  }
  __syncthreads();
  const long long row = t0 + lane;
  if (row < T) {
    const long long n = (long long)C * T;
    for (int c = wid; c < C; c += 8) {
      const float v = tile[lane * ld + c];
      const long long i = (long long)c * T + row;
      zt[i] = v;
      if (terms) {                                // the operand split of the batched tcgen05 GEMM, written in the same pass
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v);
        const float r1 = v - __bfloat162float(h0);
        const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
        terms[i] = h0; terms[n + i] = h1; terms[2 * n + i] = __float2bfloat16_rn(r1 - __bfloat162float(h1));
      }
    }
  }
}

// LayerNorm backward from a transposed upstream gradient gzt [C, T]:
// stage a 32-token tile of gzt into shared memory, then one warp per token.
__global__ void __launch_bounds__(256)
ln_bwd_transposed_kernel(const float* __restrict__ gzt, const float* __restrict__ x,
                         const float* __restrict__ mu, const float* __restrict__ rstd,
                         const float* __restrict__ w, long long T, int C,
                         float* __restrict__ gx, float* __restrict__ partial /*[grid,2C]*/) {
  extern __shared__ float sm[];              // tile [32][C+1] + acc [8][2C]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ld = C + 1;
  float* tile = sm;
  float* acc = sm + 32 * ld;                 // per-warp gamma/beta partials
  for (int i = threadIdx.x; i < 8 * 2 * C; i += 256) acc[i] = 0.f;
  for (long long t0 = (long long)blockIdx.x * 32; t0 < T; t0 += (long long)gridDim.x * 32) {
    __syncthreads();
    {
      const long long row = t0 + lane;
      if (row < T) {
#pragma unroll 8
        for (int c = wid; c < C; c += 8) tile[lane * ld + c] = __ldg(gzt + (long long)c * T + row);
      }
    }
    __syncthreads();
    for (int rr = wid; rr < 32; rr += 8) {
      long long row = t0 + rr;
      if (row >= T) break;
      const float m = mu[row], rs = rstd[row];
      const float* xr = x + row * C;
      float s1 = 0.f, s2 = 0.f;
      for (int c = lane; c < C; c += 32) {
        float gv = tile[rr * ld + c];
        float xh = (__ldg(xr + c) - m) * rs;
        float g = gv * __ldg(w + c);
        s1 += g; s2 += g * xh;
        acc[wid * 2 * C + c] += gv * xh;
        acc[wid * 2 * C + C + c] += gv;
      }
      s1 = warp_sum(s1) / (float)C;
      s2 = warp_sum(s2) / (float)C;
      for (int c = lane; c < C; c += 32) {
        float xh = (__ldg(xr + c) - m) * rs;
        float g = tile[rr * ld + c] * __ldg(w + c);
        gx[row * C + c] = (g - s1 - xh * s2) * rs;
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    float s = 0.f;
    for (int ww = 0; ww < 8; ++ww) s += acc[ww * 2 * C + c];
    partial[(size_t)blockIdx.x * 2 * C + c] = s;
  }
}

__global__ void __launch_bounds__(256)
ln_bwd_finalize2_kernel(const float* __restrict__ partial, int nblocks, int C,
                        float* __restrict__ gw, float* __restrict__ gb) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * C) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * 2 * C + c];
  if (c < C) gw[c] = s; else gb[c - C] = s;
}

// Ds[m,c,k] = sqrt(max(0, |zt[c,m]|^2 + |centers[c,k]|^2 - 2 acc))   ('C (B D) CN -> B D C CN')
struct SpaceDistEpilogue {
  float* out; const float* aa; const float* bb; long long CK; int K, M;
  __device__ __forceinline__ void operator()(int batch, int, int m, int n, float v) const {
    float sq = aa[(long long)batch * M + m] + bb[(long long)batch * K + n] - 2.0f * v;
    out[(long long)m * CK + (long long)batch * K + n] = sqrtf(fmaxf(sq, 0.f));
  }
};

// gzt[c,m,p] = zt[c,m,p] * rsum[m,c] - acc
struct SpaceGzEpilogue {
  float* out; const float* zt; const float* rsum; long long MP; int P, C;
  __device__ __forceinline__ void operator()(int batch, int, int m, int n, float v) const {
    long long i = (long long)batch * MP + (long long)m * P + n;
    out[i] = zt[i] * rsum[(long long)m * C + batch] - v;
  }
};

// gcenters[c,k,p] = centers[c,k,p] * rcol[c,k] - acc
struct SpaceGcEpilogue {
  float* out; const float* centers; const float* rcol; long long KP; int P, K;
  __device__ __forceinline__ void operator()(int batch, int, int m, int n, float v) const {
    long long i = (long long)batch * KP + (long long)m * P + n;
    out[i] = centers[i] * rcol[(long long)batch * K + m] - v;
  }
};

// The three contractions of the head as batched tcgen05 GEMMs (three-term bf16 split, fp32-faithful): worth it
// from a few GFLOP up; P, K multiples of 8 for the TMA row pitches, K a multiple of 64 so that a batch's window
// of the contraction over centroids ends on a k-block boundary (the next batch's columns must not leak in).
static bool space_tc_ok(long long M, int P, int C, int K) {
  return M >= 1 && (P % 8) == 0 && (K % 64) == 0 && C <= 65535 && (long long)C * M < (1ll << 31) &&
         (long long)C * K < (1ll << 31) && (double)M * P * K * C >= (double)(1ll << 28) && vadc_device_ok() != 0 &&
         !env_on("VADC_NO_TC_GEMM");
}

static int space_ln_bwd_blocks(long long T) {
  long long b = (T + 31) / 32;
  long long cap = (long long)sm_count() * 6;      // ~37 KB of shared memory per block at C = 192: six resident blocks per SM
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace vadc

using namespace vadc;

extern "C" size_t vadc_space_cluster_fwd_workspace_bytes(int64_t M, int P, int C, int K) {
  size_t b = 0;
  b += align_up((size_t)C * (M > 0 ? M : 1) * sizeof(float), 256);    // |zt row|^2
  b += align_up((size_t)C * K * sizeof(float), 256);                  // |center row|^2
  b += align_up((size_t)(softmin_blocks(M * C, K) + 1) * sizeof(double), 256);
  if (M > 0 && (P % 8) == 0)
    b += tc_gemm_split_bytes((long long)C * M, P) + tc_gemm_split_bytes((long long)C * K, P);
  return b + 256;
}

static int space_cluster_fwd_impl(const float* x, const float* ln_w, const float* ln_b,
                                  const float* centers, int64_t M, int P, int C, int K, int k_valid,
                                  float alpha, float eps, float* Ds, float* As, float* zt,
                                  float* mu, float* rstd, float* loss_sq, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  VADC_REQUIRE(M >= 0 && P > 0 && C > 0 && K > 0 && (K % 4) == 0, VADC_ERR_BAD_SHAPE);
  VADC_REQUIRE(M * (int64_t)P < (1ll << 31) && M * (int64_t)C < (1ll << 31), VADC_ERR_UNSUPPORTED);
  VADC_REQUIRE(ln_w && ln_b && centers && loss_sq && workspace, VADC_ERR_NULL_POINTER);
  VADC_REQUIRE(M == 0 || (x && Ds && As && zt && mu && rstd), VADC_ERR_NULL_POINTER);
  VADC_REQUIRE(aligned16(Ds) && aligned16(As), VADC_ERR_MISALIGNED);
  VADC_REQUIRE(workspace_bytes >= vadc_space_cluster_fwd_workspace_bytes(M, P, C, K), VADC_ERR_WORKSPACE);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver ws(workspace, workspace_bytes);
  float* zz = ws.take<float>((size_t)C * (M > 0 ? M : 1));
  float* cc = ws.take<float>((size_t)C * K);
  double* partial = ws.take<double>(softmin_blocks(M * C, K) + 1);
  const long long T = (long long)M * P;
  int rc;
  if (M > 0) {
    size_t smem = (size_t)32 * (C + 1) * sizeof(float);
    if (smem > 48 * 1024)
      VADC_CUDA(cudaFuncSetAttribute(ln_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool tc = space_tc_ok(M, P, C, K);
    __nv_bfloat16* zs = tc ? reinterpret_cast<__nv_bfloat16*>(ws.take<uint8_t>(tc_gemm_split_bytes((long long)C * M, P))) : nullptr;
    ln_transpose_kernel<<<(unsigned)((T + 31) / 32), 256, smem, st>>>(x, ln_w, ln_b, T, C, eps, zt, mu, rstd, zs);
    VADC_CHECK_LAUNCH("ln_transpose_kernel");
    if ((rc = launch_row_sqnorm(zt, (long long)C * M, P, zz, st))) return rc;
    if ((rc = launch_row_sqnorm(centers, (long long)C * K, P, cc, st))) return rc;
    // batch c: A = zt[c] [M,P], B = centers[c]^T; out Ds[m, c, k]
    if (tc) {
      void* cs = ws.take<uint8_t>(tc_gemm_split_bytes((long long)C * K, P));
      if ((rc = tc_split3(centers, (long long)C * K, P, cs, st))) return rc;
      TcBatchDistEpi epi{Ds, zz, cc, (long long)C * K, K, M, K};
      if ((rc = launch_tc_gemm_batched<false, false>(zs, (long long)C * M, P, cs, (long long)C * K, P, M, K, P, C,
                                                     TcBatchOffsets{(int)M, 0, K, 0}, epi, st))) return rc;
    } else {
      Operand Aop{zt, P, 1}, Bop{centers, 1, P};
      SpaceDistEpilogue epi{Ds, zz, cc, (long long)C * K, K, (int)M};
      cudaError_t e = sgemm_auto((int)M, K, P, Aop, Bop, (long long)M * P, (long long)K * P, C, 1, epi, st);
      if (e != cudaSuccess) return record_cuda_error(e, "space dist sgemm");
    }
  }
  return launch_softmin_rows(Ds, M * (long long)C, K, alpha, As, nullptr, partial, loss_sq, st, k_valid);
}

extern "C" int vadc_space_cluster_fwd(const float* x, const float* ln_w, const float* ln_b,
                                      const float* centers, int64_t M, int P, int C, int K,
                                      float alpha, float eps, float* Ds, float* As, float* zt,
                                      float* mu, float* rstd, float* loss_sq, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  return space_cluster_fwd_impl(x, ln_w, ln_b, centers, M, P, C, K, K, alpha, eps, Ds, As, zt, mu, rstd, loss_sq, workspace,
                                workspace_bytes, stream);
}

// any cluster_num: centers [C, K, P] padded by the host to K % 4 == 0; the trailing K - K_valid rows of every channel are excluded
extern "C" int vadc_space_cluster_fwd_padded(const float* x, const float* ln_w, const float* ln_b,
                                             const float* centers, int64_t M, int P, int C, int K, int K_valid,
                                             float alpha, float eps, float* Ds, float* As, float* zt,
                                             float* mu, float* rstd, float* loss_sq, void* workspace,
                                             size_t workspace_bytes, void* stream) {
  VADC_REQUIRE(K_valid >= 1 && K_valid <= K, VADC_ERR_BAD_SHAPE);
  return space_cluster_fwd_impl(x, ln_w, ln_b, centers, M, P, C, K, K_valid, alpha, eps, Ds, As, zt, mu, rstd, loss_sq,
                                workspace, workspace_bytes, stream);
}

extern "C" size_t vadc_space_cluster_bwd_workspace_bytes(int64_t M, int P, int C, int K) {
  size_t m = (size_t)(M > 0 ? M : 1);
  size_t b = 0;
  b += align_up(m * C * K * sizeof(float), 256);                       // r
  b += align_up(m * C * sizeof(float), 256);                           // rsum
  b += align_up((size_t)C * m * P * sizeof(float), 256);               // gzt
  b += align_up((size_t)colsum_chunks(M) * C * K * sizeof(float), 256);
  b += align_up((size_t)C * K * sizeof(float), 256);                   // rcol
  b += align_up((size_t)space_ln_bwd_blocks(M * (int64_t)P) * 2 * C * sizeof(float), 256);
  if ((P % 8) == 0 && (K % 64) == 0)
    b += tc_gemm_split_bytes((long long)m, (long long)C * K) + tc_gemm_split_bytes((long long)C * K, P) +
         tc_gemm_split_bytes((long long)C * m, P);
  return b + 256;
}

extern "C" int vadc_space_cluster_bwd(const float* x, const float* mu, const float* rstd,
                                      const float* zt, const float* ln_w, const float* centers,
                                      const float* Ds, const float* As, const float* gD,
                                      const float* gA, const float* g_loss_sq,
                                      int64_t M, int P, int C, int K, float alpha, float* gx,
                                      float* gcenters, float* g_ln_w, float* g_ln_b,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  VADC_REQUIRE(M > 0 && P > 0 && C > 0 && K > 0 && (K % 4) == 0, VADC_ERR_BAD_SHAPE);
  VADC_REQUIRE(M * (int64_t)P < (1ll << 31) && M * (int64_t)C < (1ll << 31), VADC_ERR_UNSUPPORTED);
  VADC_REQUIRE(x && mu && rstd && zt && ln_w && centers && Ds && As && gx && gcenters && g_ln_w &&
               g_ln_b && workspace, VADC_ERR_NULL_POINTER);
  VADC_REQUIRE(workspace_bytes >= vadc_space_cluster_bwd_workspace_bytes(M, P, C, K), VADC_ERR_WORKSPACE);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver ws(workspace, workspace_bytes);
  const long long T = (long long)M * P;
  float* r = ws.take<float>((size_t)M * C * K);
  float* rsum = ws.take<float>((size_t)M * C);
  float* gzt = ws.take<float>((size_t)C * T);
  float* cpart = ws.take<float>((size_t)colsum_chunks(M) * C * K);
  float* rcol = ws.take<float>((size_t)C * K);
  int nb = space_ln_bwd_blocks(T);
  float* lnpart = ws.take<float>((size_t)nb * 2 * C);
  int rc;
  cudaError_t e;
  if ((rc = launch_bwd_rows(Ds, As, nullptr, gD, gA, g_loss_sq, M * (long long)C, K, alpha, r, rsum, st))) return rc;
  const bool tc = space_tc_ok(M, P, C, K);
  if (tc) {
    // r as [M, C K] (batch c = its K-column window), centers as [C K, P], zt as [C M, P]: three split passes, then
    //   gzt[c] = zt[c] rsum[:,c] - r[:,c,:] centers[c]         A K-major (window offset along k), B MN-major
    //   gcenters[c] = centers[c] rcol[c] - r[:,c,:]^T zt[c]    A MN-major (window offset along m), B MN-major
    void* rs = ws.take<uint8_t>(tc_gemm_split_bytes(M, (long long)C * K));
    void* cs = ws.take<uint8_t>(tc_gemm_split_bytes((long long)C * K, P));
    void* zs = ws.take<uint8_t>(tc_gemm_split_bytes((long long)C * M, P));
    if ((rc = tc_split3(r, M, (long long)C * K, rs, st))) return rc;
    if ((rc = tc_split3(centers, (long long)C * K, P, cs, st))) return rc;
    if ((rc = tc_split3(zt, (long long)C * M, P, zs, st))) return rc;
    TcSpaceGzEpi egz{gzt, zt, rsum, T, P, C};
    if ((rc = launch_tc_gemm_batched<false, true>(rs, M, (long long)C * K, cs, (long long)C * K, P, M, P, K, C,
                                                  TcBatchOffsets{0, K, 0, K}, egz, st))) return rc;
    e = launch_colsum(r, M, C * K, cpart, rcol, st);
    if (e != cudaSuccess) return record_cuda_error(e, "space colsum r");
    TcSpaceGcEpi egc{gcenters, centers, rcol, (long long)K * P, P, K};
    if ((rc = launch_tc_gemm_batched<true, true>(rs, M, (long long)C * K, zs, (long long)C * M, P, K, P, M, C,
                                                 TcBatchOffsets{K, 0, 0, (int)M}, egc, st))) return rc;
  }
  // gzt[c] = zt[c] * rsum[:,c] - r[:,c,:] @ centers[c]
  if (!tc) {
    Operand Aop{r, (long long)C * K, 1}, Bop{centers, P, 1};
    SpaceGzEpilogue epi{gzt, zt, rsum, T, P, C};
    e = sgemm_auto((int)M, P, K, Aop, Bop, K, (long long)K * P, C, 1, epi, st);
    if (e != cudaSuccess) return record_cuda_error(e, "space bwd sgemm r.c");
  }
  // gcenters[c] = centers[c] * rcol[c] - r[:,c,:]^T @ zt[c]
  if (!tc) {
    e = launch_colsum(r, M, C * K, cpart, rcol, st);
    if (e != cudaSuccess) return record_cuda_error(e, "space colsum r");
    Operand Aop{r, 1, (long long)C * K}, Bop{zt, P, 1};
    SpaceGcEpilogue epi{gcenters, centers, rcol, (long long)K * P, P, K};
    e = sgemm_auto(K, P, (int)M, Aop, Bop, K, T, C, 1, epi, st);
    if (e != cudaSuccess) return record_cuda_error(e, "space bwd sgemm rT.zt");
  }
  size_t smem = ((size_t)32 * (C + 1) + (size_t)8 * 2 * C) * sizeof(float);
  if (smem > 48 * 1024)
    VADC_CUDA(cudaFuncSetAttribute(ln_bwd_transposed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ln_bwd_transposed_kernel<<<nb, 256, smem, st>>>(gzt, x, mu, rstd, ln_w, T, C, gx, lnpart);
  VADC_CHECK_LAUNCH("ln_bwd_transposed_kernel");
  ln_bwd_finalize2_kernel<<<(2 * C + 255) / 256, 256, 0, st>>>(lnpart, nb, C, g_ln_w, g_ln_b);
  VADC_CHECK_LAUNCH("ln_bwd_finalize2_kernel");
  return VADC_OK;
}
