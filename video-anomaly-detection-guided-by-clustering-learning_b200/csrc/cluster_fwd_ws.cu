// cluster_fwd_ws.cu — fused cluster forward (C1 + L1), warp-specialised tcgen05 pipeline.
//
//   model/cluster.py:81-99 + model/backbone.py:98 in ONE persistent kernel, K == 32:
//   LayerNorm -> distance GEMM (tcgen05, TMEM accumulator) -> {sqrt/clamp, argmin, softmin,
//   sum (D*A)^2} -> x_rec GEMM (tcgen05) -> TMA tensor stores of D, A, x_rec.
//
// Unlike cluster_tc.cu (every warp walks every phase, phases serialised inside the CTA) the three
// kinds of work run concurrently on different 128-token tiles:
//   warps 0-7   PRODUCERS  x rows (registers, software-pipelined one 8-row batch ahead) -> LayerNorm
//               -> feature rows to HBM -> two fp16 terms into the SWIZZLE_128B K-major operand tile
//   warps 8-11  EPILOGUE   thread = token row = TMEM lane: D from TMEM -> distance / argmin /
//               softmin / loss -> swizzled staging (TMA stores of D and A) + fp16 A operand;
//               then x_rec from TMEM -> staging -> TMA stores
//   warp 12     MMA        one thread: GEMM1 (M=128, N=K) into a double-buffered TMEM D, GEMM2
//               (M=128, N=C) against the SAME centroid image read MN-major
// mbarriers: zfull / zempty (operand tile), dfull[2] / dempty[2] (TMEM D), afull (A operand +
// staging ready), rfull (x_rec accumulator ready).
//
// fp32-faithful tensor-core arithmetic: every operand v is scaled by a power of two into [-8, 8]
// (LayerNorm output bound sqrt(C) max|gamma| + max|beta|; max |centroid|; softmin in [0,1]) and split
// exactly into two fp16 terms v = h1 + h2 (11 + 11 mantissa bits, absolute floor 2^-25); the product
// keeps h1*g1 + h1*g2 + h2*g1 in fp32 TMEM (dropped term <= 2^-22 relative per element, far below
// the fp32 rounding of the |z|^2 + |c|^2 - 2 z.c cancellation).  fp16 K-major and MN-major
// SWIZZLE_128B tiles are byte-identical, so one centroid image feeds both GEMMs.
#include <cuda.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <type_traits>
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "cluster.h"
#include "rows.cuh"

namespace vadc {
using namespace tc;

namespace ws {

constexpr int kTileM = 128;
constexpr int kK = 32;                       // centroids (this kernel is specialised for K == 32)
constexpr int kProdWarps = 11;               // warps 0-10; 8-row blocks are dealt round-robin
constexpr int kMmaWarp = 11;
constexpr int kEpiWarp0 = 12;                // 4 epilogue warps: TMEM lane quadrant = warp % 4
constexpr int kThreads = 16 * 32;            // 4 warps per scheduler -> 128 registers per thread
constexpr int kBlocksPerTile = kTileM / 8;   // 16 producer blocks of 8 rows
constexpr uint32_t kFmtF16 = 0;
constexpr float kAScale = 8.0f;              // softmin assignment scaled into [0, 8] before the fp16 split

struct SmemPlan {
  uint32_t z_off, c_off, aop_off, dst_off, ast_off, xst_off, cc_off, zz_off, gam_off, bet_off, misc_off, total;
  uint32_t z_bytes, c_bytes;
};

__host__ __device__ inline SmemPlan plan(int C) {
  SmemPlan p;
  uint32_t off = 0;
  p.z_bytes = 2u * kTileM * C * 2u;                      // two fp16 terms of the 128 x C token tile
  p.c_bytes = 2u * kK * C * 2u;                          // two fp16 terms of the K x C centroids
  p.z_off = off; off += p.z_bytes;
  p.c_off = off; off += p.c_bytes;
  p.aop_off = off; off += 2u * kTileM * kK * 2u;         // two fp16 terms of A, un-swizzled K-major
  p.dst_off = off; off += kTileM * kK * 4u;              // D staging [128 x 128 B], SWIZZLE_128B
  p.ast_off = off; off += kTileM * kK * 4u;              // A staging
  p.xst_off = off; off += 2u * kTileM * 128u;            // x_rec staging, two slots
  p.cc_off = off; off += kK * 4u;
  p.zz_off = off; off += 4u * kTileM * 4u;               // |z|^2 per row, ring of 4 tiles
  p.gam_off = off; off += (uint32_t)C * 4u;
  p.bet_off = off; off += (uint32_t)C * 4u;
  p.misc_off = off; off += 256u;
  p.total = off;
  return p;
}

// barrier slots inside misc
enum { B_CEN = 0, B_ZFULL, B_ZEMPTY, B_DFULL0, B_DFULL1, B_DEMPTY0, B_DEMPTY1, B_AFULL, B_RFULL, B_COUNT };

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_hint(float4* p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const void* tmap, const void* smem_src, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);                   // a -> low half, b -> high half
  return *reinterpret_cast<uint32_t*>(&h);
}
// exact two-term fp16 split of a pair: (a, b) = (lo(p1), hi(p1)) + (lo(p2), hi(p2)) + O(2^-22)
__device__ __forceinline__ void split2_h(float a, float b, uint32_t& p1, uint32_t& p2) {
  p1 = pack_h2(a, b);
  const float2 l = sub2(make_float2(a, b), __half22float2(*reinterpret_cast<__half2*>(&p1)));
  p2 = pack_h2(l.x, l.y);
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// power of two s with s * bound in [4, 8); 1 for a zero / non-finite bound
__device__ __forceinline__ float pow2_scale(float bound) {
  if (!(bound > 0.f) || !isfinite(bound)) return 1.0f;
  int e;
  (void)frexpf(bound, &e);                               // bound = m 2^e, m in [0.5, 1)
  e = max(-96, min(96, e));
  return ldexpf(1.0f, 3 - e);
}

// ---------------------------------------------------------------------------
// centroid prologue: scales, two-term fp16 split written as the shared-memory IMAGE
// (per term: C/64 blocks of [K rows x 128 B], SWIZZLE_128B), |c_k|^2 (fp32, unsplit values).
// scales[0] = s_z, [1] = s_c, [2] = 1/(s_z s_c), [3] = 1/(kAScale s_c)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
centroid_prep_ws_kernel(const float* __restrict__ centers, const float* __restrict__ ln_w,
                        const float* __restrict__ ln_b, int K, int C, uint8_t* __restrict__ image,
                        float* __restrict__ cc, float* __restrict__ scales) {
  const int k = blockIdx.x;
  __shared__ float red[32];
  __shared__ float bc[3];
  float mc = 0.f, mg = 0.f, mb = 0.f;
  // every block needs max |centers| (24 KB, L2-resident): 128-bit loads, all of a thread's loads independent (this prologue
  // sits on the step's critical path: 12 us with scalar loads in a dependent chain, measured under ncu)
  if (((K * C) & 3) == 0 && (reinterpret_cast<uintptr_t>(centers) & 15u) == 0) {
    float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* c4 = reinterpret_cast<const float4*>(centers);
#pragma unroll 8
    for (int i = threadIdx.x; i < (K * C) / 4; i += blockDim.x) {
      const float4 v = __ldg(c4 + i);
      m4.x = fmaxf(m4.x, fabsf(v.x)); m4.y = fmaxf(m4.y, fabsf(v.y)); m4.z = fmaxf(m4.z, fabsf(v.z)); m4.w = fmaxf(m4.w, fabsf(v.w));
    }
    mc = fmaxf(fmaxf(m4.x, m4.y), fmaxf(m4.z, m4.w));
  } else {
    for (int i = threadIdx.x; i < K * C; i += blockDim.x) mc = fmaxf(mc, fabsf(centers[i]));
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) { mg = fmaxf(mg, fabsf(ln_w[i])); mb = fmaxf(mb, fabsf(ln_b[i])); }
  mc = warp_max(mc); mg = warp_max(mg); mb = warp_max(mb);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[w] = mc; red[8 + w] = mg; red[16 + w] = mb; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float m = 0.f;
    for (int i = 0; i < 8; ++i) m = fmaxf(m, red[8 * threadIdx.x + i]);
    bc[threadIdx.x] = m;
  }
  __syncthreads();
  const float s_c = pow2_scale(bc[0]);
  const float zb = sqrtf((float)C) * bc[1] + bc[2];
  const float s_z = (zb >= 0.25f && zb <= 4096.f) ? 1.0f : pow2_scale(zb);
  if (k == 0 && threadIdx.x == 0) {
    scales[0] = s_z; scales[1] = s_c; scales[2] = 1.0f / (s_z * s_c); scales[3] = 1.0f / (kAScale * s_c);
  }
  float s = 0.f;
  const uint32_t term = (uint32_t)K * C * 2u;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = centers[(size_t)k * C + c];
    s += v * v;
    const float vs = v * s_c;
    const __half h1 = __float2half_rn(vs);
    const __half h2 = __float2half_rn(vs - __half2float(h1));
    const uint32_t off = (uint32_t)(c / 64) * (K * 128u) + sw128(k, (c % 64) * 2);
    *reinterpret_cast<__half*>(image + off) = h1;
    *reinterpret_cast<__half*>(image + term + off) = h2;
  }
  s = block_sum<float>(s, red);
  if (threadIdx.x == 0) cc[k] = s;
}

struct Params {
  const float* x; const float* ln_w; const float* ln_b;
  const uint8_t* cimage; const float* cc; const float* scales;
  float* feature; long long* label; float* mu; float* rstd; float* rowstats; double* partial;
  long long N; float alpha, eps; int pf, hint;
};

template <int F4>
__global__ void __launch_bounds__(kThreads, 1)
cluster_fwd_ws_kernel(const __grid_constant__ CUtensorMap mapD, const __grid_constant__ CUtensorMap mapA,
                      const __grid_constant__ CUtensorMap mapR, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int C = F4 * 32, K = kK;
  const SmemPlan pl = plan(C);
  uint8_t* sZ = smem + pl.z_off;                        // [2][C/64][128 x 128 B]
  uint8_t* sC = smem + pl.c_off;                        // [2][C/64][K x 128 B]
  float* sCC = reinterpret_cast<float*>(smem + pl.cc_off);
  float* sZZ = reinterpret_cast<float*>(smem + pl.zz_off);
  float* sGam = reinterpret_cast<float*>(smem + pl.gam_off);
  float* sBet = reinterpret_cast<float*>(smem + pl.bet_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.misc_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + pl.misc_off + 128);
  double* loss_acc = reinterpret_cast<double*>(smem + pl.misc_off + 136);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ntiles = (p.N + kTileM - 1) / kTileM;
  const int nmine = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles of this CTA (>= 1)
  constexpr uint32_t zterm = (uint32_t)kTileM * C * 2u;  // bytes per fp16 term of the token tile
  constexpr uint32_t cterm = (uint32_t)K * C * 2u;
  constexpr uint32_t ncols = (64 + C <= 128) ? 128u : 256u;       // D0 [0,32) D1 [32,64) R [64, 64+C)

  if (tid == 0) {
    mbar_init(&bars[B_CEN], 1);
    mbar_init(&bars[B_ZFULL], kBlocksPerTile);
    mbar_init(&bars[B_ZEMPTY], 1);
    mbar_init(&bars[B_DFULL0], 1); mbar_init(&bars[B_DFULL1], 1);
    mbar_init(&bars[B_DEMPTY0], 4); mbar_init(&bars[B_DEMPTY1], 4);
    mbar_init(&bars[B_AFULL], 1);
    mbar_init(&bars[B_RFULL], 1);
    fence_mbar_init();
    *loss_acc = 0.0;
    prefetch_tmap(&mapD); prefetch_tmap(&mapA); prefetch_tmap(&mapR);
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, ncols);
  for (int k = tid; k < K; k += kThreads) sCC[k] = p.cc[k];
  for (int c = tid; c < C; c += kThreads) { sGam[c] = p.ln_w[c]; sBet[c] = p.ln_b[c]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float s_z = __ldg(p.scales + 0);

  if (warp < kProdWarps) {
    // ======================================================================= PRODUCERS
    // 8 lanes per token row (lane j owns float4 chunks j, j+8, ...), 4 rows per warp instruction;
    // the two rows sharing a 16-lane store phase differ by 4 so that their swizzled 64-byte
    // pieces fall in disjoint banks.  A block = 8 rows (2 per lane group); the 16 blocks of every
    // tile are dealt round-robin to the producer warps, latency is hidden by the other warps.
    const int lj = lane & 7, lg = lane >> 3;
    const uint64_t pol = policy_evict_first();
    const int rsel = (lg & 1) * 4 + (lg >> 1);           // rows +0, +4, +1, +5 within an 8-row block
    const float invC = 1.0f / (float)C;
    const int nblocks = kBlocksPerTile * nmine;
    const uint32_t sZ32 = smem_u32(sZ), sGam32 = smem_u32(sGam) + lj * 16, sBet32 = smem_u32(sBet) + lj * 16;
    auto produce = [&](auto scaled_tag) {
    constexpr bool SCALED = decltype(scaled_tag)::value;
    for (int g = warp; g < nblocks; g += kProdWarps) {
      const int it = g / kBlocksPerTile, blk = g % kBlocksPerTile;
      const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
      const long long row0 = tile * kTileM;
      if (p.pf > 0) {
        if (blk == 0 && lane == 0 && it + p.pf < nmine) {  // L2 prefetch pf tiles ahead
          const long long tn = tile + (long long)p.pf * gridDim.x;
          const long long rows = min((long long)kTileM, p.N - tn * kTileM);
          prefetch_l2_bulk(p.x + tn * kTileM * C, (uint32_t)(rows * C * 4));
        }
      } else if (p.pf < 0 && lane == 0) {                  // this warp's own block, -pf rounds ahead
        const int gn = g - p.pf * kProdWarps;
        if (gn < nblocks) {
          const long long rn = ((long long)blockIdx.x + (long long)(gn / kBlocksPerTile) * gridDim.x) * kTileM +
                               (gn % kBlocksPerTile) * 8;
          const long long rows = min(8ll, p.N - rn);
          if (rows > 0) prefetch_l2_bulk(p.x + rn * C, (uint32_t)(rows * C * 4));
        }
      }
      float4 v[2][F4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const long long row = row0 + blk * 8 + rsel + 2 * u;
        const float4* xr = reinterpret_cast<const float4*>(p.x + row * C) + lj;
#pragma unroll
        for (int i = 0; i < F4; ++i)
          v[u][i] = (row < p.N) ? ld_stream(xr + 8 * i) : make_float4(0, 0, 0, 0);
      }
      float mean[2], rs[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < F4; ++i) s2 = add2(s2, add2(make_float2(v[u][i].x, v[u][i].y), make_float2(v[u][i].z, v[u][i].w)));
        float s = s2.x + s2.y;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        mean[u] = s * invC;
        const float2 nm2 = bcast2(-mean[u]);
        float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < F4; ++i) {
          const float2 a = add2(make_float2(v[u][i].x, v[u][i].y), nm2), b = add2(make_float2(v[u][i].z, v[u][i].w), nm2);
          v[u][i] = make_float4(a.x, a.y, b.x, b.y);
          q2 = fma2(a, a, fma2(b, b, q2));
        }
        float q = q2.x + q2.y;
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        q += __shfl_xor_sync(0xffffffffu, q, 4);
        rs[u] = 1.0f / sqrtf(q * invC + p.eps);
      }
      // operand tile free: GEMM1 of the previous tile has completed
      mbar_wait_spin(&bars[B_ZEMPTY], (uint32_t)((it & 1) ^ 1));
      float* zzbuf = sZZ + (it & 3) * kTileM;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = blk * 8 + rsel + 2 * u;
        const long long row = row0 + r;
        const bool live = row < p.N;
        float2 nz2 = make_float2(0.f, 0.f), p12 = nz2, p22 = nz2;   // |z|^2, sum z gamma, sum z gamma xhat (for the backward)
        const uint32_t rx = (uint32_t)(r & 7);
        const uint32_t zrow = sZ32 + (uint32_t)r * 128u;
        float4* frow = reinterpret_cast<float4*>(p.feature + row * C) + lj;
#pragma unroll
        for (int i = 0; i < F4; ++i) {
          const float4 gm = lds128f(sGam32 + i * 128);
          const float4 be = lds128f(sBet32 + i * 128);
          const float2 rs2 = bcast2(rs[u]);
          const float2 g01 = make_float2(gm.x, gm.y), g23 = make_float2(gm.z, gm.w);
          const float2 t01 = mul2(make_float2(v[u][i].x, v[u][i].y), rs2), t23 = mul2(make_float2(v[u][i].z, v[u][i].w), rs2);
          float2 o01 = fma2(t01, g01, make_float2(be.x, be.y)), o23 = fma2(t23, g23, make_float2(be.z, be.w));
          if (live) {
            const float4 o = make_float4(o01.x, o01.y, o23.x, o23.y);
            if (p.hint & 1) st_hint(frow + 8 * i, o, pol); else frow[8 * i] = o;
          } else {
            o01 = make_float2(0.f, 0.f); o23 = o01;
          }
          nz2 = fma2(o01, o01, fma2(o23, o23, nz2));
          if (p.rowstats) {
            const float2 og01 = mul2(o01, g01), og23 = mul2(o23, g23);
            p12 = add2(p12, add2(og01, og23));
            p22 = fma2(og01, t01, fma2(og23, t23, p22));
          }
          uint32_t a1, a2, b1, b2;
          if constexpr (SCALED) {
            const float2 sz2 = bcast2(s_z);
            const float2 w01 = mul2(o01, sz2), w23 = mul2(o23, sz2);
            split2_h(w01.x, w01.y, a1, a2);
            split2_h(w23.x, w23.y, b1, b2);
          } else {
            split2_h(o01.x, o01.y, a1, a2);
            split2_h(o23.x, o23.y, b1, b2);
          }
          const int f = lj + 8 * i;                          // float4 index in the row: channels 4f..4f+3
          const uint32_t byte = (uint32_t)(f & 15) * 8u;     // 8 bytes of fp16 in the 128-byte row of block f/16
          const uint32_t off = (uint32_t)(f >> 4) * (kTileM * 128u) + ((((byte >> 4) ^ rx) << 4) | (byte & 15u));
          sts64(zrow + off, a1, b1);
          sts64(zrow + zterm + off, a2, b2);
        }
        float nz = nz2.x + nz2.y, p1 = p12.x + p12.y, p2 = p22.x + p22.y;
        nz += __shfl_xor_sync(0xffffffffu, nz, 1);
        nz += __shfl_xor_sync(0xffffffffu, nz, 2);
        nz += __shfl_xor_sync(0xffffffffu, nz, 4);
        if (p.rowstats) {
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            p1 += __shfl_xor_sync(0xffffffffu, p1, o);
            p2 += __shfl_xor_sync(0xffffffffu, p2, o);
          }
        }
        if (lj == 0) {
          zzbuf[r] = nz;
          if (live) {
            p.mu[row] = mean[u]; p.rstd[row] = rs[u];
            if (p.rowstats) reinterpret_cast<float4*>(p.rowstats)[row] = make_float4(nz, p1, p2, 0.f);
          }
        }
      }
      fence_async_smem();                                // this block's 8 rows of the tile are in place
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_ZFULL]);
    }
    };
    if (s_z == 1.0f) produce(std::false_type{}); else produce(std::true_type{});
  } else if (warp >= kEpiWarp0) {
    // ======================================================================= EPILOGUE
    const int et = tid - kEpiWarp0 * 32;                 // 0..127 = token row of this thread = TMEM lane
    const int q4 = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const bool issuer = (et == 0);
    const uint64_t pol = policy_evict_first();
    const float inv_zc = __ldg(p.scales + 2), inv_r = __ldg(p.scales + 3);
    const float m2 = -2.0f * inv_zc;
    const float nalpha = -p.alpha * 1.4426950408889634f; // exp(-a t) = 2^(-a log2(e) t)
    float loss_thread = 0.f;
    uint8_t* dblk = smem + pl.dst_off;
    uint8_t* ablk = smem + pl.ast_off;
    const uint32_t dblk32 = smem_u32(dblk), ablk32 = smem_u32(ablk), aop32 = smem_u32(smem + pl.aop_off);
    const uint32_t xst32 = smem_u32(smem + pl.xst_off);
    constexpr uint32_t aterm = (uint32_t)kTileM * K * 2u;
    for (int it = 0; it < nmine; ++it) {
      const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
      const long long row0 = tile * kTileM;
      const long long row = row0 + et;
      const bool live = row < p.N;
      const int buf = it & 1;
      mbar_wait(&bars[B_DFULL0 + buf], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float dv[32];
      tmem_ld32(tmem + lane_addr + (uint32_t)(buf * 32), dv);
      const float zz = sZZ[(it & 3) * kTileM + et];
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_DEMPTY0 + buf]);
      float best = INFINITY; int bidx = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        dv[j] = fast_sqrt(fmaxf(fmaf(dv[j], m2, zz + sCC[j]), 0.f));
        if (dv[j] < best) { best = dv[j]; bidx = j; }
      }
      float ev[32];
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) { ev[j] = exp2f(nalpha * (dv[j] - best)); sum += ev[j]; }
      const float inv = fast_rcp(sum);
      float lrow = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        ev[j] *= inv;
        const float pr = dv[j] * ev[j];
        lrow = fmaf(pr, pr, lrow);
      }
      if (live) { loss_thread += lrow; p.label[row] = bidx; }
      // staging + A operand are free once the previous tile's TMA stores have read them
      if (issuer) bulk_wait_read0();
      named_bar(1, 128);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        sts128f(dblk32 + sw128(et, q * 16), dv[4 * q], dv[4 * q + 1], dv[4 * q + 2], dv[4 * q + 3]);
        sts128f(ablk32 + sw128(et, q * 16), ev[4 * q], ev[4 * q + 1], ev[4 * q + 2], ev[4 * q + 3]);
      }
      // A operand of GEMM2: 2 fp16 terms, un-swizzled K-major core matrices (8 rows x 16 B)
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        uint32_t w1[4], w2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          split2_h(ev[8 * kc + 2 * e] * kAScale, ev[8 * kc + 2 * e + 1] * kAScale, w1[e], w2[e]);
        const uint32_t off = (uint32_t)kc * 2048u + (uint32_t)et * 16u;
        sts128(aop32 + off, w1[0], w1[1], w1[2], w1[3]);
        sts128(aop32 + aterm + off, w2[0], w2[1], w2[2], w2[3]);
      }
      fence_async_smem();
      named_bar(1, 128);
      if (issuer) {
        mbar_arrive(&bars[B_AFULL]);
        if (p.hint & 2) {
          tma_store_2d_hint(&mapD, dblk, 0, (int)row0, pol);
          tma_store_2d_hint(&mapA, ablk, 0, (int)row0, pol);
        } else {
          tma_store_2d(&mapD, dblk, 0, (int)row0);
          tma_store_2d(&mapA, ablk, 0, (int)row0);
        }
        bulk_commit();
      }
      // ---- x_rec
      mbar_wait(&bars[B_RFULL], (uint32_t)(it & 1));
      tc_fence_after();
      float xv[32];
#pragma unroll 1
      for (int ch = 0; ch < C / 32; ++ch) {
        uint8_t* slot = smem + pl.xst_off + (ch & 1) * (kTileM * 128u);
        const uint32_t slot32 = xst32 + (ch & 1) * (kTileM * 128u);
        tmem_ld32(tmem + lane_addr + 64u + (uint32_t)(ch * 32), xv);
        if (issuer) bulk_wait_read1();                   // the store that last used this slot has been read
        named_bar(2, 128);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          sts128f(slot32 + sw128(et, q * 16), xv[4 * q] * inv_r, xv[4 * q + 1] * inv_r, xv[4 * q + 2] * inv_r, xv[4 * q + 3] * inv_r);
        fence_async_smem();
        named_bar(2, 128);
        if (issuer) {
          if (p.hint & 2) tma_store_2d_hint(&mapR, slot, ch * 32, (int)row0, pol); else tma_store_2d(&mapR, slot, ch * 32, (int)row0);
          bulk_commit();
        }
      }
      tc_fence_before();
    }
    loss_thread = warp_sum(loss_thread);
    if (lane == 0) atomicAdd(loss_acc, (double)loss_thread);
    if (issuer) bulk_wait0();
  } else if (warp == kMmaWarp && lane == 0) {
    // ======================================================================= MMA
    // centroid operand image: one bulk TMA copy per CTA
    mbar_expect_tx(&bars[B_CEN], pl.c_bytes);
    for (uint32_t off = 0; off < pl.c_bytes; off += 32768u)
      bulk_g2s(sC + off, p.cimage + off, min(32768u, pl.c_bytes - off), &bars[B_CEN]);
    const uint32_t idesc1 = instr_desc(kFmtF16, 128, K, 0, 0);
    const uint32_t idesc2 = instr_desc(kFmtF16, 128, C, 0, 1);
    const uint32_t zb = smem_u32(sZ), cb = smem_u32(sC), ab = smem_u32(smem + pl.aop_off);
    constexpr uint32_t aterm = (uint32_t)kTileM * K * 2u;
    constexpr int pi[3] = {0, 1, 0}, pj[3] = {1, 0, 0};  // small terms first
    mbar_wait(&bars[B_CEN], 0);
    int g1 = 0, g2 = 0;
    while (g2 < nmine) {
      if (g1 < nmine && mbar_try_wait(&bars[B_ZFULL], (uint32_t)(g1 & 1)) &&
          mbar_try_wait(&bars[B_DEMPTY0 + (g1 & 1)], (uint32_t)(((g1 >> 1) & 1) ^ 1))) {
        tc_fence_after();
        const uint32_t tmemD = tmem + (uint32_t)((g1 & 1) * 32);
        uint32_t acc = 0;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const uint32_t za = zb + pi[t] * zterm, ca = cb + pj[t] * cterm;
#pragma unroll
          for (int kk = 0; kk < C / 16; ++kk) {
            const uint64_t ad = smem_desc_sw128(za + (kk >> 2) * (kTileM * 128u) + (kk & 3) * 32u, 0, 1024);
            const uint64_t bd = smem_desc_sw128(ca + (kk >> 2) * (K * 128u) + (kk & 3) * 32u, 0, 1024);
            mma_f16(tmemD, ad, bd, idesc1, acc);
            acc = 1;
          }
        }
        mma_commit(&bars[B_ZEMPTY]);
        mma_commit(&bars[B_DFULL0 + (g1 & 1)]);
        ++g1;
      }
      if (g2 < g1 && mbar_try_wait(&bars[B_AFULL], (uint32_t)(g2 & 1))) {
        tc_fence_after();
        const uint32_t tmemR = tmem + 64u;
        uint32_t acc = 0;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
#pragma unroll
          for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t ad = smem_desc_noswz(ab + pi[t] * aterm + (uint32_t)(2 * ks) * 2048u, 2048, 128);
            const uint64_t bd = smem_desc_sw128(cb + pj[t] * cterm + (uint32_t)(2 * ks) * 1024u, (uint32_t)K * 128u, 1024);
            mma_f16(tmemR, ad, bd, idesc2, acc);
            acc = 1;
          }
        }
        mma_commit(&bars[B_RFULL]);
        ++g2;
      }
    }
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (tid == 0) p.partial[blockIdx.x] = *loss_acc;
  if (warp == kMmaWarp) { tc_fence_after(); tmem_dealloc(tmem, ncols); }
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

static bool shape_ok(long long N, int C, int K) {
  if (N < 1 || N >= (1ll << 31)) return false;
  if (K != kK) return false;
  if (C != 64 && C != 128 && C != 192) return false;
  return plan(C).total + 1024 <= 227u * 1024u;
}

}  // namespace ws
}  // namespace vadc

using namespace vadc;

size_t vadc_cluster_ws_extra_workspace_bytes(int64_t N, int C, int K) {
  if (!ws::shape_ok(N, C, K)) return 0;
  return align_up((size_t)2 * K * C * 2, 256) + align_up((size_t)(sm_count() + 1) * sizeof(double), 256) +
         align_up((size_t)K * sizeof(float), 256) + 256 + 256;
}

int vadc_cluster_fwd_ws(const float* x, const float* ln_w, const float* ln_b, const float* centers,
                        int64_t N, int C, int K, float alpha, float eps, float* D, float* A,
                        float* x_rec, float* feature, int64_t* label, float* mu, float* rstd,
                        float* rowstats, float* loss_sq, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (!ws::shape_ok(N, C, K)) return VADC_ERR_UNSUPPORTED;
  if (!vadc_device_ok()) return VADC_ERR_NO_DEVICE;
  if (workspace_bytes < vadc_cluster_ws_extra_workspace_bytes(N, C, K)) return VADC_ERR_WORKSPACE;
  Carver wsp(workspace, workspace_bytes);
  uint8_t* image = wsp.take<uint8_t>((size_t)2 * K * C * 2);
  double* partial = wsp.take<double>(sm_count() + 1);
  float* cc = wsp.take<float>(K);
  float* scales = wsp.take<float>(4);
  const int grid = (int)std::min<long long>((N + ws::kTileM - 1) / ws::kTileM, sm_count());

  CUtensorMap mD, mA, mR;
  int rc;
  if ((rc = ws::make_map(&mD, D, N, K))) return rc;
  if ((rc = ws::make_map(&mA, A, N, K))) return rc;
  if ((rc = ws::make_map(&mR, x_rec, N, C))) return rc;

  ws::centroid_prep_ws_kernel<<<K, 256, 0, st>>>(centers, ln_w, ln_b, K, C, image, cc, scales);
  VADC_CHECK_LAUNCH("centroid_prep_ws_kernel");

  const size_t smem = ws::plan(C).total + 1024;
  ws::Params p{x, ln_w, ln_b, image, cc, scales, feature, reinterpret_cast<long long*>(label), mu, rstd, rowstats,
               partial, (long long)N, alpha, eps, env_int("VADC_WS_PF", -1),
               env_int("VADC_WS_HINT", 3)};
#define WS_CASE(F4_)                                                                                   \
  if (C == 32 * F4_) {                                                                                 \
    VADC_CUDA(cudaFuncSetAttribute(ws::cluster_fwd_ws_kernel<F4_>,                                      \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    ws::cluster_fwd_ws_kernel<F4_><<<grid, ws::kThreads, smem, st>>>(mD, mA, mR, p);                    \
    launched = true;                                                                                   \
  }
  bool launched = false;
  timing_begin(VADC_TIMING_CLUSTER_FWD, st);
  WS_CASE(2) WS_CASE(4) WS_CASE(6)
#undef WS_CASE
  timing_end(VADC_TIMING_CLUSTER_FWD, st);
  if (!launched) return VADC_ERR_UNSUPPORTED;
  VADC_CHECK_LAUNCH("cluster_fwd_ws_kernel");
  finalize_sum_kernel<<<1, 1024, 0, st>>>(partial, grid, loss_sq);
  VADC_CHECK_LAUNCH("finalize_sum_kernel");
  return VADC_OK;
}
