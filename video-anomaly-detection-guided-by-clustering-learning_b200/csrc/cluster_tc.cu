// cluster_tc.cu — fused cluster forward (C1 + L1) on tcgen05 / TMEM / TMA.
//
//   model/cluster.py:81-99 + model/backbone.py:98 in ONE persistent kernel:
//   LayerNorm -> distance GEMM (tcgen05, TMEM accumulator) -> epilogue
//   {sqrt/clamp, row min + argmin, softmin, sum (D*A)^2} -> x_rec GEMM
//   (tcgen05) -> TMA tensor stores of D, A, x_rec.
//
// fp32-faithful tensor-core arithmetic: every fp32 operand v is split exactly
// into three bf16 terms v = v1 + v2 + v3 (8+8+8 mantissa bits) and the product
// is accumulated in fp32 TMEM from the six terms of order <= 2^-16
// (v1w1, v1w2, v2w1, v1w3, v2w2, v3w1); the dropped terms are <= 2^-24
// relative, i.e. below fp32 rounding.  kind::f16 is used rather than kind::tf32
// because bf16 K-major and MN-major SWIZZLE_128B tiles are byte-identical, so
// one shared-memory image of the centroids feeds both GEMMs (tf32 MN-major
// needs a different swizzle; tests/test_gpu_umma.py).
//
// One CTA (256 threads) per SM, persistent over 128-token tiles:
//   P1  8 warps: x rows (L2-prefetched one tile ahead by cp.async.bulk.prefetch)
//       -> LayerNorm in registers -> feature rows to HBM (coalesced float4) ->
//       3 bf16 terms into SWIZZLE_128B K-major operand tiles in shared memory
//   P2  1 thread: 6 x C/16 tcgen05.mma (M=128, N=K, K=16) -> TMEM cols [0,K)
//   P3  warps 0-3 (thread = token row): tcgen05.ld, distance / argmin / softmin /
//       loss, D and A rows into swizzled staging tiles (TMA tensor store), A's
//       3 bf16 terms as the (un-swizzled K-major) A operand of GEMM2
//   P4  1 thread: 6 x K/16 tcgen05.mma (M=128, N=C) against the SAME centroid
//       tiles read MN-major -> TMEM cols [K, K+C)
//   P5  8 warps: tcgen05.ld -> swizzled staging -> TMA tensor store of x_rec
// Shared memory (C=192, K=32): 144 KB token operand (re-used as staging after
// GEMM1), 36 KB centroid operand, loaded once per CTA by a bulk TMA copy.
#include <cuda.h>
#include <algorithm>
#include <stdio.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "cluster.h"
#include "rows.cuh"

namespace vadc {
using namespace tc;

constexpr int kTileM = 128;
constexpr int kTcThreads = 256;

struct TcSmemPlan {
  uint32_t z_bytes, c_bytes, a_off, dst_off, ast_off, xst_off, alias_end;
  uint32_t cc_off, zz_off, misc_off, total;
};

__host__ __device__ inline TcSmemPlan tc_plan(int C, int K) {
  TcSmemPlan p;
  p.z_bytes = 3u * kTileM * C * 2u;                 // three bf16 terms of the 128 x C token tile
  p.c_bytes = 3u * K * C * 2u;                      // three bf16 terms of the K x C centroids
  p.a_off = 0;                                      // aliases the token operand after GEMM1
  p.dst_off = p.a_off + 3u * kTileM * K * 2u;       // D staging  [K/32 blocks][128 x 128 B]
  p.ast_off = p.dst_off + kTileM * K * 4u;          // A staging
  p.xst_off = p.ast_off + kTileM * K * 4u;          // x_rec staging, 2 slots of [128 x 128 B]
  p.alias_end = p.xst_off + 2u * kTileM * 128u;
  uint32_t off = p.z_bytes + p.c_bytes;
  p.cc_off = off; off += K * 4u;
  p.zz_off = off; off += kTileM * 4u;
  p.misc_off = off; off += 128u;
  p.total = off;
  return p;
}

// ---------------------------------------------------------------------------
// centroid prologue: 3-term bf16 split written as the shared-memory IMAGE
// (per term: C/64 blocks of [K rows x 128 B], SWIZZLE_128B), plus |c_k|^2.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
centroid_prep_kernel(const float* __restrict__ centers, int K, int C, uint8_t* __restrict__ image,
                     float* __restrict__ cc) {
  const int k = blockIdx.x;
  __shared__ float red[32];
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = centers[(size_t)k * C + c];
    s += v * v;
    __nv_bfloat16 t1 = __float2bfloat16_rn(v);
    float r1 = v - __bfloat162float(t1);
    __nv_bfloat16 t2 = __float2bfloat16_rn(r1);
    float r2 = r1 - __bfloat162float(t2);
    __nv_bfloat16 t3 = __float2bfloat16_rn(r2);
    uint32_t off = (uint32_t)(c / 64) * (K * 128u) + sw128(k, (c % 64) * 2);
    const uint32_t term = (uint32_t)K * C * 2u;
    *reinterpret_cast<__nv_bfloat16*>(image + off) = t1;
    *reinterpret_cast<__nv_bfloat16*>(image + term + off) = t2;
    *reinterpret_cast<__nv_bfloat16*>(image + 2 * term + off) = t3;
  }
  s = block_sum<float>(s, red);
  if (threadIdx.x == 0) cc[k] = s;
}

__device__ __forceinline__ void split3(float v, __nv_bfloat16& t1, __nv_bfloat16& t2, __nv_bfloat16& t3) {
  t1 = __float2bfloat16_rn(v);
  float r1 = v - __bfloat162float(t1);
  t2 = __float2bfloat16_rn(r1);
  float r2 = r1 - __bfloat162float(t2);
  t3 = __float2bfloat16_rn(r2);
}

__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 1D bulk copy global -> shared with mbarrier completion
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct TcParams {
  const float* x; const float* ln_w; const float* ln_b;
  const uint8_t* cimage; const float* cc;
  float* feature; long long* label; float* mu; float* rstd; double* partial;
  long long N; int C, K; float alpha, eps;
};

__global__ void __launch_bounds__(kTcThreads, 1)
cluster_fwd_tc_kernel(const __grid_constant__ CUtensorMap mapD, const __grid_constant__ CUtensorMap mapA,
                      const __grid_constant__ CUtensorMap mapR, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int C = p.C, K = p.K;
  const TcSmemPlan pl = tc_plan(C, K);
  uint8_t* sZ = smem;                                   // [3][C/64][128 x 128 B]
  uint8_t* sC = smem + pl.z_bytes;                      // [3][C/64][K x 128 B]
  float* sCC = reinterpret_cast<float*>(smem + pl.cc_off);
  float* sZZ = reinterpret_cast<float*>(smem + pl.zz_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.misc_off);     // [0] centroids, [1] gemm1, [2] gemm2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + pl.misc_off + 32);
  double* loss_acc = reinterpret_cast<double*>(smem + pl.misc_off + 40);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ntiles = (p.N + kTileM - 1) / kTileM;
  const uint32_t zterm = (uint32_t)kTileM * C * 2u;     // bytes per bf16 term of the token tile
  const uint32_t cterm = (uint32_t)K * C * 2u;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)(K + C)) ncols <<= 1;

  if (tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
    *loss_acc = 0.0;
    prefetch_tmap(&mapD); prefetch_tmap(&mapA); prefetch_tmap(&mapR);
  }
  if (warp == 1) tmem_alloc(tmem_slot, ncols);
  for (int k = tid; k < K; k += kTcThreads) sCC[k] = p.cc[k];
  __syncthreads();
  if (tid == 0) {                                       // centroid operand image: bulk TMA copy, once per CTA
    mbar_expect_tx(&bars[0], pl.c_bytes);
    for (uint32_t off = 0; off < pl.c_bytes; off += 32768u) {
      uint32_t n = min(32768u, pl.c_bytes - off);
      bulk_g2s(sC + off, p.cimage + off, n, &bars[0]);
    }
    long long t0 = blockIdx.x;
    if (t0 < ntiles) {
      long long rows = min((long long)kTileM, p.N - t0 * kTileM);
      prefetch_l2_bulk(p.x + t0 * kTileM * C, (uint32_t)(rows * C * 4));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmemD = tmem, tmemR = tmem + (uint32_t)K;

  // LayerNorm affine parameters of this lane's channels (2 float4 per lane: C <= 256)
  const int nv = C >> 2;
  float4 gam[2], bet[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int c4 = lane + 32 * i;
    gam[i] = (c4 < nv) ? __ldg(reinterpret_cast<const float4*>(p.ln_w) + c4) : make_float4(0, 0, 0, 0);
    bet[i] = (c4 < nv) ? __ldg(reinterpret_cast<const float4*>(p.ln_b) + c4) : make_float4(0, 0, 0, 0);
  }
  const uint32_t idesc1 = instr_desc(kFmtBF16, 128, K, 0, 0);
  const uint32_t idesc2 = instr_desc(kFmtBF16, 128, C, 0, 1);
  const float invC = 1.0f / (float)C;
  float loss_thread = 0.f;
  uint32_t phase = 0;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, phase ^= 1) {
    const long long row0 = tile * kTileM;
    if (tid == 0) {                                     // L2 prefetch of the next tile's tokens
      long long tn = tile + gridDim.x;
      if (tn < ntiles) {
        long long rows = min((long long)kTileM, p.N - tn * kTileM);
        prefetch_l2_bulk(p.x + tn * kTileM * C, (uint32_t)(rows * C * 4));
      }
    }
    // ------------------------------------------------------------------ P1
    for (int rr = warp; rr < kTileM; rr += 4 * (kTcThreads / 32)) {
      float4 v[4][2];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long row = row0 + rr + u * (kTcThreads / 32);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          int c4 = lane + 32 * i;
          v[u][i] = (row < p.N && c4 < nv) ? ld_stream(reinterpret_cast<const float4*>(p.x + row * C) + c4)
                                           : make_float4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = rr + u * (kTcThreads / 32);
        const long long row = row0 + r;
        float s = (v[u][0].x + v[u][0].y) + (v[u][0].z + v[u][0].w) + (v[u][1].x + v[u][1].y) + (v[u][1].z + v[u][1].w);
        const float mean = warp_sum(s) * invC;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (lane + 32 * i < nv) {
            float a = v[u][i].x - mean, b = v[u][i].y - mean, c = v[u][i].z - mean, d = v[u][i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
          }
        }
        const float rs = 1.0f / sqrtf(warp_sum(q) * invC + p.eps);
        float nz = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c4 = lane + 32 * i;
          if (c4 < nv) {
            float4 o;
            o.x = (v[u][i].x - mean) * rs * gam[i].x + bet[i].x;
            o.y = (v[u][i].y - mean) * rs * gam[i].y + bet[i].y;
            o.z = (v[u][i].z - mean) * rs * gam[i].z + bet[i].z;
            o.w = (v[u][i].w - mean) * rs * gam[i].w + bet[i].w;
            if (row >= p.N) o = make_float4(0, 0, 0, 0);
            else reinterpret_cast<float4*>(p.feature + row * C)[c4] = o;
            nz += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
            __nv_bfloat16 a1, a2, a3, b1, b2, b3, c1, c2, c3, d1, d2, d3;
            split3(o.x, a1, a2, a3); split3(o.y, b1, b2, b3); split3(o.z, c1, c2, c3); split3(o.w, d1, d2, d3);
            const uint32_t off = (uint32_t)(c4 >> 4) * (kTileM * 128u) + sw128(r, (c4 & 15) * 8);
            *reinterpret_cast<uint2*>(sZ + off) = make_uint2(pack2(a1, b1), pack2(c1, d1));
            *reinterpret_cast<uint2*>(sZ + zterm + off) = make_uint2(pack2(a2, b2), pack2(c2, d2));
            *reinterpret_cast<uint2*>(sZ + 2 * zterm + off) = make_uint2(pack2(a3, b3), pack2(c3, d3));
          }
        }
        nz = warp_sum(nz);
        if (lane == 0) {
          sZZ[r] = nz;
          if (row < p.N) { p.mu[row] = mean; p.rstd[row] = rs; }
        }
      }
    }
    fence_async_smem();
    __syncthreads();
    // ------------------------------------------------------------------ P2
    if (tid == 0) {
      if (tile == (long long)blockIdx.x) mbar_wait(&bars[0], 0);        // centroid image landed
      tc_fence_after();
      const uint32_t zb = smem_u32(sZ), cb = smem_u32(sC);
      uint32_t acc = 0;
      const int pi[6] = {0, 0, 1, 0, 1, 2}, pj[6] = {0, 1, 0, 2, 1, 0};
#pragma unroll 1
      for (int t = 0; t < 6; ++t) {
        const uint32_t za = zb + pi[t] * zterm, ca = cb + pj[t] * cterm;
        for (int kk = 0; kk < C / 16; ++kk) {
          uint64_t ad = smem_desc_sw128(za + (kk >> 2) * (kTileM * 128u) + (kk & 3) * 32u, 0, 1024);
          uint64_t bd = smem_desc_sw128(ca + (kk >> 2) * (K * 128u) + (kk & 3) * 32u, 0, 1024);
          mma_f16(tmemD, ad, bd, idesc1, acc);
          acc = 1;
        }
      }
      mma_commit(&bars[1]);
    }
    // ------------------------------------------------------------------ P3
    if (warp < 4) {
      mbar_wait(&bars[1], phase);
      tc_fence_after();
      const int r = tid;                                 // token row of this thread (TMEM lane)
      const long long row = row0 + r;
      const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
      const float zz = sZZ[r];
      float best = INFINITY; int bidx = 0;
      float dv[32];
      for (int c0 = 0; c0 < K; c0 += 32) {
        tmem_ld32(tmemD + lane_addr + c0, dv);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float d = sqrtf(fmaxf(zz + sCC[c0 + j] - 2.0f * dv[j], 0.f));
          if (d < best) { best = d; bidx = c0 + j; }
        }
      }
      float sum = 0.f;
      for (int c0 = 0; c0 < K; c0 += 32) {
        if (K > 32) tmem_ld32(tmemD + lane_addr + c0, dv);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float d = sqrtf(fmaxf(zz + sCC[c0 + j] - 2.0f * dv[j], 0.f));
          sum += expf(-p.alpha * (d - best));
        }
      }
      for (int c0 = 0; c0 < K; c0 += 32) {
        if (K > 32) tmem_ld32(tmemD + lane_addr + c0, dv);
        uint8_t* dblk = smem + pl.dst_off + (c0 >> 5) * (kTileM * 128u);
        uint8_t* ablk = smem + pl.ast_off + (c0 >> 5) * (kTileM * 128u);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float d[4], a[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            d[e] = sqrtf(fmaxf(zz + sCC[c0 + 4 * q + e] - 2.0f * dv[4 * q + e], 0.f));
            a[e] = expf(-p.alpha * (d[e] - best)) / sum;
            float pr = d[e] * a[e];
            if (row < p.N) loss_thread += pr * pr;
          }
          *reinterpret_cast<float4*>(dblk + sw128(r, q * 16)) = make_float4(d[0], d[1], d[2], d[3]);
          *reinterpret_cast<float4*>(ablk + sw128(r, q * 16)) = make_float4(a[0], a[1], a[2], a[3]);
          dv[4 * q + 0] = a[0]; dv[4 * q + 1] = a[1]; dv[4 * q + 2] = a[2]; dv[4 * q + 3] = a[3];
        }
        // A operand of GEMM2: 3 bf16 terms, un-swizzled K-major core matrices (8 rows x 16 B)
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          uint32_t w1[4], w2[4], w3[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat16 x1, x2, x3, y1, y2, y3;
            split3(dv[8 * kc + 2 * e], x1, x2, x3);
            split3(dv[8 * kc + 2 * e + 1], y1, y2, y3);
            w1[e] = pack2(x1, y1); w2[e] = pack2(x2, y2); w3[e] = pack2(x3, y3);
          }
          const uint32_t off = (uint32_t)((c0 >> 3) + kc) * 2048u + (r >> 3) * 128u + (r & 7) * 16u;
          const uint32_t aterm = (uint32_t)kTileM * K * 2u;
          *reinterpret_cast<uint4*>(smem + pl.a_off + off) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
          *reinterpret_cast<uint4*>(smem + pl.a_off + aterm + off) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
          *reinterpret_cast<uint4*>(smem + pl.a_off + 2 * aterm + off) = make_uint4(w3[0], w3[1], w3[2], w3[3]);
        }
      }
      if (row < p.N) p.label[row] = bidx;
      fence_async_smem();
      tc_fence_before();
      named_bar(1, 128);
      // ---------------------------------------------------------------- P4
      if (tid == 0) {
        tc_fence_after();
        for (int b = 0; b < K / 32; ++b) {
          tma_store_2d(&mapD, smem + pl.dst_off + b * (kTileM * 128u), b * 32, (int)row0);
          tma_store_2d(&mapA, smem + pl.ast_off + b * (kTileM * 128u), b * 32, (int)row0);
        }
        bulk_commit();
        const uint32_t ab = smem_u32(smem + pl.a_off), cb = smem_u32(sC);
        const uint32_t aterm = (uint32_t)kTileM * K * 2u;
        uint32_t acc = 0;
        const int pi[6] = {0, 0, 1, 0, 1, 2}, pj[6] = {0, 1, 0, 2, 1, 0};
#pragma unroll 1
        for (int t = 0; t < 6; ++t) {
          for (int ks = 0; ks < K / 16; ++ks) {
            uint64_t ad = smem_desc_noswz(ab + pi[t] * aterm + (uint32_t)(2 * ks) * 2048u, 2048, 128);
            uint64_t bd = smem_desc_sw128(cb + pj[t] * cterm + (uint32_t)(2 * ks) * 1024u, (uint32_t)K * 128u, 1024);
            mma_f16(tmemR, ad, bd, idesc2, acc);
            acc = 1;
          }
        }
        mma_commit(&bars[2]);
      }
    }
    // ------------------------------------------------------------------ P5
    {
      mbar_wait(&bars[2], phase);
      tc_fence_after();
      const int half = warp >> 2, q4 = warp & 3;
      const int r = q4 * 32 + lane;
      uint8_t* slot = smem + pl.xst_off + half * (kTileM * 128u);
      const bool issuer = (q4 == 0 && lane == 0);
      float xv[32];
      for (int ch = half; ch < C / 32; ch += 2) {
        tmem_ld32(tmemR + ((uint32_t)(q4 * 32) << 16) + ch * 32, xv);
        if (issuer) bulk_wait_read0();                   // previous store out of this slot has been read
        named_bar(2 + half, 128);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(slot + sw128(r, q * 16)) = make_float4(xv[4 * q], xv[4 * q + 1], xv[4 * q + 2], xv[4 * q + 3]);
        fence_async_smem();
        named_bar(2 + half, 128);
        if (issuer) { tma_store_2d(&mapR, slot, ch * 32, (int)row0); bulk_commit(); }
      }
      if (issuer || tid == 0) bulk_wait_read0();         // staging aliases the next tile's operand tiles
    }
    tc_fence_before();
    __syncthreads();
  }
  // ---- teardown
  loss_thread = warp_sum(loss_thread);
  if (lane == 0 && warp < 4) atomicAdd(loss_acc, (double)loss_thread);
  if (tid == 0 || tid == 128) bulk_wait0();
  __syncthreads();
  if (tid == 0) {
    if (ntiles <= (long long)blockIdx.x) mbar_wait(&bars[0], 0);     // never consumed: drain the bulk copy
    p.partial[blockIdx.x] = *loss_acc;
  }
  if (warp == 1) tmem_dealloc(tmem, ncols);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// [rows, cols] fp32 row-major tensor, box = 32 columns x 128 rows, SWIZZLE_128B
static int make_map(CUtensorMap* m, float* base, long long rows, int cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return VADC_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)kTileM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "cuTensorMapEncodeTiled failed: %d", (int)r);
    return VADC_ERR_CUDA;
  }
  return VADC_OK;
}

static bool tc_shape_ok(long long N, int C, int K) {
  if (N < 1 || N >= (1ll << 31)) return false;
  if (C % 64 || C < 64 || C > 256) return false;
  if (K % 32 || K < 32 || K > 256 || K + C > 512) return false;
  TcSmemPlan pl = tc_plan(C, K);
  if (pl.alias_end > pl.z_bytes) return false;          // staging must fit in the dead operand tiles
  if (pl.total + 1024 > 227 * 1024) return false;
  return true;
}

}  // namespace vadc

using namespace vadc;

extern "C" size_t vadc_cluster_tc_extra_workspace_bytes(int64_t N, int C, int K) {
  if (!tc_shape_ok(N, C, K)) return 0;
  return align_up((size_t)3 * K * C * 2, 256) + align_up((size_t)(sm_count() + 1) * sizeof(double), 256) +
         align_up((size_t)K * sizeof(float), 256) + 256;
}

int vadc_cluster_fwd_tc(const float* x, const float* ln_w, const float* ln_b, const float* centers,
                        int64_t N, int C, int K, float alpha, float eps, float* D, float* A,
                        float* x_rec, float* feature, int64_t* label, float* mu, float* rstd,
                        float* loss_sq, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (!tc_shape_ok(N, C, K)) return VADC_ERR_UNSUPPORTED;
  if (!vadc_device_ok()) return VADC_ERR_NO_DEVICE;
  if (workspace_bytes < vadc_cluster_tc_extra_workspace_bytes(N, C, K)) return VADC_ERR_WORKSPACE;
  Carver ws(workspace, workspace_bytes);       // the SIMT carve is not live on this path
  uint8_t* image = ws.take<uint8_t>((size_t)3 * K * C * 2);
  const int grid = (int)std::min<long long>((N + kTileM - 1) / kTileM, sm_count());
  double* partial = ws.take<double>(sm_count() + 1);
  float* cc = ws.take<float>(K);

  CUtensorMap mD, mA, mR;
  int rc;
  if ((rc = make_map(&mD, D, N, K))) return rc;
  if ((rc = make_map(&mA, A, N, K))) return rc;
  if ((rc = make_map(&mR, x_rec, N, C))) return rc;

  centroid_prep_kernel<<<K, 256, 0, st>>>(centers, K, C, image, cc);
  VADC_CHECK_LAUNCH("centroid_prep_kernel");

  const TcSmemPlan pl = tc_plan(C, K);
  const size_t smem = pl.total + 1024;
  static size_t attr_set = 0;
  if (smem > attr_set) {
    VADC_CUDA(cudaFuncSetAttribute(cluster_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = smem;
  }
  TcParams p{x, ln_w, ln_b, image, cc, feature, reinterpret_cast<long long*>(label), mu, rstd, partial,
             (long long)N, C, K, alpha, eps};
  cluster_fwd_tc_kernel<<<grid, kTcThreads, smem, st>>>(mD, mA, mR, p);
  VADC_CHECK_LAUNCH("cluster_fwd_tc_kernel");
  finalize_sum_kernel<<<1, 256, 0, st>>>(partial, grid, loss_sq);
  VADC_CHECK_LAUNCH("finalize_sum_kernel");
  return VADC_OK;
}
