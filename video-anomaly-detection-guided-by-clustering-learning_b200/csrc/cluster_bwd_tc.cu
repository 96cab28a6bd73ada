// cluster_bwd_tc.cu — C2: fused backward of the cluster head on tcgen05 / TMEM, K == 32,
// training-graph case (gradients arrive through x_rec and the fused cluster loss only).
//
//   autograd of model/cluster.py:81-99 + model/backbone.py:98 (loss.backward(), main_predict.py:296)
//   in ONE persistent warp-specialised kernel; per token it reads x, gR, D, A (+ mu, rstd) once and
//   writes gx once: 12 C + 8 K bytes (fused-loss variant).  64-token tiles:
//
//   warps 2,3,6,7,10,11,14  PRODUCERS  x, gR rows -> xhat = (x - mu) rstd -> two-term bf16 splits of
//                           xhat and gR into SWIZZLE_128B operand tiles + three row sums of xhat
//   warp 15 (one thread)    MMA        S1  G1 = gR cen^T            [tokens x K]   (TMEM, double buffer)
//                                      S5a P1 += gR^T A             [channels x K] (TMEM, whole kernel)
//                                      S3  acc = r cen              [tokens x C]
//                                      S5b P2 += xhat^T r           [channels x K] (TMEM, whole kernel)
//   warps 0,1               E1         thread = token: softmin backward + cdist ratio -> r; the
//                                      LayerNorm-backward row statistics in closed form; bf16 splits
//                                      of A and r into the shared [A | r] operand tile
//   warps 4,5,8,9,12,13     E3         thread = token x 64 channels: gz = z rsum - acc, LayerNorm
//                                      backward, gx through swizzled staging + TMA tensor stores,
//                                      Q[c] = sum_n xhat^2 rsum (register butterfly) for g_gamma
//
// The token tile is the MMA's M (S1, S3: M = 128 with rows 64..127 reading past the tile — their
// TMEM lanes are never read) and its K (S5: the same shared-memory bytes viewed MN-major, M =
// channels).  All operands are exact two-term bf16 splits v = h + l (16 significant bits, full fp32
// exponent range — gradients have no a-priori scale, which rules out fp16); a product keeps
// h*h + h*l + l*h in the fp32 accumulator (~2^-16 relative per element; the fp32 reference's own
// error against fp64 is larger, scripts/bwd_algebra_check.py).
//
// LayerNorm backward needs the row means of gg = gz*gamma and gg*xhat BEFORE gz exists per thread.
// They follow in closed form from quantities E1 already has (z = gamma xhat + beta):
//   sum_c gg        = rsum * sum_c z gamma      - sum_k r_k (cen_k . gamma)
//   sum_c gg xhat   = rsum * sum_c z gamma xhat - sum_k r_k T_k,  T_k = xhat . (gamma * cen_k)
//   T_k = z.cen_k - beta.cen_k = (|z|^2 + |cen_k|^2 - D_k^2) / 2 - beta.cen_k      (D is an input)
// so E3 is a single pass over the accumulator.  gcenters = P1^T - gamma * P2^T + (cen - beta) colsum(r),
//   g_beta  = gamma * rowsum_k(P2) + beta * sum(rcol) - rcol . cen
//   g_gamma = gamma * Q + beta * rowsum_k(P2) - sum_k cen[k,:] * P2[:,k]       (no column sums of gz).
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "cluster.h"

namespace vadc {
using namespace tc;

namespace bt {

constexpr int kTok = 64;                     // tokens per tile
constexpr int kK = 32;                       // centroids
constexpr int kThreads = 512;
constexpr int kProd = 7;                     // producer warps
constexpr int kUnits = kTok / 4;             // 4-row producer units per tile
constexpr int kMmaWarp = 15;
constexpr uint32_t kBlk = kTok * 128u;       // one [64 rows x 128 B] operand block

struct Plan {
  uint32_t x_off, g_off, ar_off, cen_off, stg_off, prow_off, scal_off, gam_off, bet_off, g2_off, bg_off,
      cvec_off, misc_off, total;
  uint32_t xbuf, xterm, gterm, cterm;
};

__host__ __device__ inline Plan plan(int C) {
  Plan p;
  const uint32_t ncb = (uint32_t)C / 64u;
  uint32_t off = 0;
  p.xterm = ncb * kBlk; p.xbuf = 2u * p.xterm; p.gterm = p.xterm; p.cterm = ncb * (kK * 128u);
  p.x_off = off; off += 2u * p.xbuf;                       // [2 bufs][2 terms][C/64][64 x 128 B]
  p.g_off = off; off += 2u * p.gterm;                      // [2 terms][C/64][64 x 128 B]
  p.ar_off = off; off += 2u * kBlk;                        // [2 terms][64 x 128 B]: bytes 0..63 A, 64..127 r
  p.cen_off = off; off += 2u * p.cterm;                    // [2 terms][C/64][32 x 128 B]
  p.stg_off = off; off += 6u * 4096u;                      // per E3 warp: [32 rows x 128 B]
  p.prow_off = off; off += 2u * 3u * kTok * 4u;            // producer row sums [buf][zz, p1, p2][64]
  p.scal_off = off; off += 2u * 4u * kTok * 4u;            // E1 row scalars [parity][rsum, s1r, s2r, rs][64]
  p.gam_off = off; off += (uint32_t)C * 4u;
  p.bet_off = off; off += (uint32_t)C * 4u;
  p.g2_off = off; off += (uint32_t)C * 4u;                 // gamma^2
  p.bg_off = off; off += (uint32_t)C * 4u;                 // beta gamma
  p.cvec_off = off; off += 3u * kK * 4u;                   // hc = |c|^2/2 - beta.c ; cg = gamma.c
  p.misc_off = off; off += 256u;
  p.total = off;
  return p;
}

enum { B_CEN = 0, B_PFULL, B_GEMPTY, B_XEMPTY0, B_XEMPTY1, B_G1FULL0, B_G1FULL1, B_G1EMPTY0, B_G1EMPTY1,
       B_AFULL, B_RFULL, B_AREMPTY, B_ACCFULL, B_ACCEMPTY, B_DONE, B_COUNT };

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
// read-only tables (gamma, beta, ...): not volatile, so the compiler may hoist / schedule the loads
__device__ __forceinline__ float4 lds128f_const(uint32_t addr) {
  float4 r;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ float lds32f_const(uint32_t addr) {
  float r;
  asm("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr));
  return r;
}
// descriptor = {hi, lo}; lo carries the 14-bit (address >> 4) field, so a byte offset is one add
__device__ __forceinline__ uint64_t desc_at(uint32_t lo, uint32_t hi, uint32_t byte_off) {
  return ((uint64_t)hi << 32) | (uint64_t)(lo + (byte_off >> 4));
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  if (accumulate)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_store_2d_hint(const void* tmap, uint32_t smem_src, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(tmap), "r"(smem_src), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 ldg_nc(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// exact two-term bf16 split of a pair: (a, b) = (lo16(p1), hi16(p1)) + (lo16(p2), hi16(p2)) + O(2^-17)
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);        // a -> low half, b -> high half
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf_hi(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }
__device__ __forceinline__ void split2_bf(float a, float b, uint32_t& p1, uint32_t& p2) {
  p1 = pack_bf2(a, b);
  p2 = pack_bf2(a - bf_lo(p1), b - bf_hi(p1));
}

// butterfly transpose-reduce over the 32 lanes (rows) of a warp: stages with xor distance >= STOP.
// After stage s, slot j of a lane holds the partial column sum for column  (lane & ~(s-1) & 31 bits
// already consumed) + j; after all five stages lane l holds the sum of column l in v[0].
template <int STOP>
__device__ __forceinline__ void butterfly(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= STOP; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float keep = up ? v[j + s] : v[j];
      const float send = up ? v[j] : v[j + s];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}
template <int START, int N>
__device__ __forceinline__ void butterfly_tail(float (&v)[N], int lane) {
#pragma unroll
  for (int s = START; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float keep = up ? v[j + s] : v[j];
      const float send = up ? v[j] : v[j + s];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

// ---------------------------------------------------------------------------
// prologue: two-term bf16 centroid image ([term][C/64][32 rows x 128 B], SWIZZLE_128B) and the
// per-centroid constants  hc_k = |c_k|^2 / 2 - beta.c_k,  cg_k = gamma.c_k
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
centroid_prep_bwd_kernel(const float* __restrict__ centers, const float* __restrict__ ln_w,
                         const float* __restrict__ ln_b, int K, int C, uint8_t* __restrict__ image,
                         float* __restrict__ cvec) {
  const int k = blockIdx.x;
  __shared__ float red[32];
  float s = 0.f, sb = 0.f, sg = 0.f;
  const uint32_t term = (uint32_t)K * C * 2u;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = centers[(size_t)k * C + c];
    s += v * v; sb += v * ln_b[c]; sg += v * ln_w[c];
    const __nv_bfloat16 h1 = __float2bfloat16_rn(v);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(v - __bfloat162float(h1));
    const uint32_t off = (uint32_t)(c / 64) * (K * 128u) + sw128(k, (c % 64) * 2);
    *reinterpret_cast<__nv_bfloat16*>(image + off) = h1;
    *reinterpret_cast<__nv_bfloat16*>(image + term + off) = h2;
  }
  s = block_sum<float>(s, red);
  sb = block_sum<float>(sb, red);
  sg = block_sum<float>(sg, red);
  if (threadIdx.x == 0) { cvec[k] = 0.5f * s - sb; cvec[K + k] = sg; }
}

struct Params {
  const float* x; const float* gR; const float* D; const float* A; const float* mu; const float* rstd;
  const float* ln_w; const float* ln_b; const uint8_t* cimage; const float* cvec; const float* g_loss_sq;
  float* part_p; float* part_rcol; float* part_ln;       // [grid][2][C][32], [grid][2][32], Q: [grid][2][C]
  long long N; float alpha; int pf;
  unsigned long long* trace;                             // debugging: per-warp event log of CTA 0 (VADC_BWD_TRACE)
};

template <int F4, bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
cluster_bwd_tc_kernel(const __grid_constant__ CUtensorMap mapGx, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int C = F4 * 32, K = kK, NCB = C / 64, MB = (C + 127) / 128;
  const Plan pl = plan(C);
  float* sProw = reinterpret_cast<float*>(smem + pl.prow_off);
  float* sScal = reinterpret_cast<float*>(smem + pl.scal_off);
  float* sGam = reinterpret_cast<float*>(smem + pl.gam_off);
  float* sBet = reinterpret_cast<float*>(smem + pl.bet_off);
  float* sG2 = reinterpret_cast<float*>(smem + pl.g2_off);
  float* sBG = reinterpret_cast<float*>(smem + pl.bg_off);
  float* sHc = reinterpret_cast<float*>(smem + pl.cvec_off);
  float* sCg = sHc + K;
  float* sConst = sCg + K;                               // [0] = sum beta gamma, [1] = sum beta^2
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.misc_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + pl.misc_off + 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // event trace (debug): entry = {code << 32 | tile, clock64}
  constexpr int kTraceCap = 1024;
  int trace_n = 0;
  auto TR = [&](int code, int it) {
    if constexpr (!TRACE) return;
    if (p.trace && blockIdx.x == 0 && lane == 0 && trace_n < kTraceCap) {
      unsigned long long* e = p.trace + ((size_t)warp * kTraceCap + trace_n) * 2;
      e[0] = ((unsigned long long)code << 32) | (unsigned)it; e[1] = (unsigned long long)clock64();
      ++trace_n;
    }
  };
  const long long ntiles = (p.N + kTok - 1) / kTok;
  const int nmine = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles of this CTA (>= 1)
  constexpr uint32_t kColG1 = 0, kColAcc = 64, kColP1 = 64 + C, kColP2 = kColP1 + 32 * MB;
  constexpr uint32_t ncols = (kColP2 + 32 * MB <= 256) ? 256u : 512u;
  constexpr int kE3Warps = 2 * NCB;

  if (tid == 0) {
    mbar_init(&bars[B_CEN], 1);
    mbar_init(&bars[B_PFULL], kUnits);
    mbar_init(&bars[B_GEMPTY], 1);
    mbar_init(&bars[B_XEMPTY0], kE3Warps + 1); mbar_init(&bars[B_XEMPTY1], kE3Warps + 1);
    mbar_init(&bars[B_G1FULL0], 1); mbar_init(&bars[B_G1FULL1], 1);
    mbar_init(&bars[B_G1EMPTY0], 2); mbar_init(&bars[B_G1EMPTY1], 2);
    mbar_init(&bars[B_AFULL], 1);
    mbar_init(&bars[B_RFULL], 1);
    mbar_init(&bars[B_AREMPTY], 1);
    mbar_init(&bars[B_ACCFULL], 1);
    mbar_init(&bars[B_ACCEMPTY], kE3Warps);
    mbar_init(&bars[B_DONE], 1);
    fence_mbar_init();
    prefetch_tmap(&mapGx);
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, ncols);
  for (int c = tid; c < C; c += kThreads) {
    const float g = p.ln_w[c], b = p.ln_b[c];
    sGam[c] = g; sBet[c] = b; sG2[c] = g * g; sBG[c] = b * g;
  }
  for (int k = tid; k < 2 * K; k += kThreads) sHc[k] = p.cvec[k];
  if (warp == 0) {
    float a = 0.f, b = 0.f;
    for (int c = lane; c < C; c += 32) { const float g = p.ln_w[c], be = p.ln_b[c]; a += be * g; b += be * be; }
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { sConst[0] = a; sConst[1] = b; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sX32 = smem_u32(smem + pl.x_off), sG32 = smem_u32(smem + pl.g_off);
  const uint32_t sAR32 = smem_u32(smem + pl.ar_off), sCen32 = smem_u32(smem + pl.cen_off);

  const bool is_e = (warp & 3) < 2;
  if (!is_e && warp != kMmaWarp) {
    // ======================================================================= PRODUCERS
    // 8 lanes per token row (lane j owns float4 chunks j, j+8, ...), 4 rows per warp instruction = one
    // unit; the two rows sharing a 16-lane store phase differ by 4 (disjoint banks after the 128B
    // swizzle): unit u of a tile covers rows 8 (u/2) + 2 (u%2) + {0, 4, 1, 5}.
    const int pw = (warp >> 2) * 2 + (warp & 1);         // 0..6
    const int lj = lane & 7, lg = lane >> 3;
    const int rsel = (lg & 1) * 4 + (lg >> 1);
    const int nunits = kUnits * nmine;
    const float cB1 = sConst[0], cB2 = sConst[1];
    const uint32_t sG2a = smem_u32(sG2) + lj * 16, sBGa = smem_u32(sBG) + lj * 16;
    for (int g = pw; g < nunits; g += kProd) {
      const int it = g / kUnits, u = g % kUnits;
      const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
      const long long row0 = tile * kTok;
      if (p.pf > 0 && lane == 0 && (u & 1) == 0 && it + p.pf < nmine) {   // L2 prefetch: same 8-row group, pf tiles ahead
        const long long rn = (tile + (long long)p.pf * gridDim.x) * kTok + (u >> 1) * 8;
        const long long rows = min(8ll, p.N - rn);
        if (rows > 0) {
          prefetch_l2_bulk(p.x + rn * C, (uint32_t)(rows * C * 4));
          prefetch_l2_bulk(p.gR + rn * C, (uint32_t)(rows * C * 4));
        }
      }
      const int r = (u >> 1) * 8 + (u & 1) * 2 + rsel;
      const long long row = row0 + r;
      const bool live = row < p.N;
      float4 xv[F4], gv[F4];
      {
        const float4* xr = reinterpret_cast<const float4*>(p.x + row * C) + lj;
        const float4* gr = reinterpret_cast<const float4*>(p.gR + row * C) + lj;
#pragma unroll
        for (int i = 0; i < F4; ++i) xv[i] = live ? ld_stream(xr + 8 * i) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < F4; ++i) gv[i] = live ? ld_stream(gr + 8 * i) : make_float4(0, 0, 0, 0);
      }
      const float rs = live ? __ldg(p.rstd + row) : 0.f;
      const float nmr = live ? -__ldg(p.mu + row) * rs : 0.f;
      TR(0, it);
      const int buf = it & 1;
      const uint32_t rx = (uint32_t)(r & 7);
      const uint32_t rowoff = (uint32_t)r * 128u;
      // ---- xhat: X[buf] is free once S5b and E3 of tile it-2 have finished
      mbar_wait_spin(&bars[B_XEMPTY0 + buf], (uint32_t)(((it >> 1) & 1) ^ 1));
      {
        const uint32_t xb = sX32 + buf * pl.xbuf;
        float q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
        for (int i = 0; i < F4; ++i) {
          const int f = lj + 8 * i;                          // float4 index in the row: channels 4f..4f+3
          const uint32_t byte = (uint32_t)(f & 15) * 8u;     // 8 bytes of bf16 in the 128-byte row of block f/16
          const uint32_t off = (uint32_t)(f >> 4) * kBlk + rowoff + ((((byte >> 4) ^ rx) << 4) | (byte & 15u));
          float4 h;
          h.x = fmaf(xv[i].x, rs, nmr); h.y = fmaf(xv[i].y, rs, nmr);
          h.z = fmaf(xv[i].z, rs, nmr); h.w = fmaf(xv[i].w, rs, nmr);
          const float4 g2 = lds128f_const(sG2a + i * 128);
          const float4 bg = lds128f_const(sBGa + i * 128);
          float w;
          w = g2.x * h.x; q1 += w; q2 = fmaf(w, h.x, q2); q3 = fmaf(bg.x, h.x, q3);
          w = g2.y * h.y; q1 += w; q2 = fmaf(w, h.y, q2); q3 = fmaf(bg.y, h.y, q3);
          w = g2.z * h.z; q1 += w; q2 = fmaf(w, h.z, q2); q3 = fmaf(bg.z, h.z, q3);
          w = g2.w * h.w; q1 += w; q2 = fmaf(w, h.w, q2); q3 = fmaf(bg.w, h.w, q3);
          uint32_t a1, a2, b1, b2;
          split2_bf(h.x, h.y, a1, a2);
          split2_bf(h.z, h.w, b1, b2);
          sts64(xb + off, a1, b1);
          sts64(xb + pl.xterm + off, a2, b2);
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          q1 += __shfl_xor_sync(0xffffffffu, q1, o);
          q2 += __shfl_xor_sync(0xffffffffu, q2, o);
          q3 += __shfl_xor_sync(0xffffffffu, q3, o);
        }
        if (lj == 0) {
          float* prow = sProw + buf * 3 * kTok;
          prow[r] = live ? q2 + 2.f * q3 + cB2 : 0.f;        // |z|^2
          prow[kTok + r] = live ? q1 + cB1 : 0.f;            // sum z gamma
          prow[2 * kTok + r] = live ? q2 + q3 : 0.f;         // sum z gamma xhat
        }
      }
      // ---- gR: G is free once S1 / S5a of the previous tile have completed
      mbar_wait_spin(&bars[B_GEMPTY], (uint32_t)((it & 1) ^ 1));
      TR(1, it);
#pragma unroll
      for (int i = 0; i < F4; ++i) {
        const int f = lj + 8 * i;
        const uint32_t byte = (uint32_t)(f & 15) * 8u;
        const uint32_t off = (uint32_t)(f >> 4) * kBlk + rowoff + ((((byte >> 4) ^ rx) << 4) | (byte & 15u));
        uint32_t a1, a2, b1, b2;
        split2_bf(gv[i].x, gv[i].y, a1, a2);
        split2_bf(gv[i].z, gv[i].w, b1, b2);
        sts64(sG32 + off, a1, b1);
        sts64(sG32 + pl.gterm + off, a2, b2);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_PFULL]);
      TR(2, it);
    }
  } else if (warp < 2) {
    // ======================================================================= E1
    const int et = tid;                                  // 0..63 = token row = TMEM lane
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const float sc = p.g_loss_sq ? 2.0f * __ldg(p.g_loss_sq) : 0.f;
    const float invC = 1.0f / (float)C;
    const uint32_t sHc32 = smem_u32(sHc), sCg32 = smem_u32(sCg);
    float rcol_acc = 0.f;
    float dv[32], av[32];
    float rs_next = 0.f;
    auto load_da = [&](int it) {                         // D / A rows of tile `it` (software-pipelined one tile ahead)
      const long long row = ((long long)blockIdx.x + (long long)it * gridDim.x) * kTok + et;
      const bool live = it < nmine && row < p.N;
      const float4* dp = reinterpret_cast<const float4*>(p.D + row * K);
      const float4* ap = reinterpret_cast<const float4*>(p.A + row * K);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 d4 = live ? ldg_nc(dp + q) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 a4 = live ? ldg_nc(ap + q) : make_float4(0, 0, 0, 0);
        dv[4 * q] = d4.x; dv[4 * q + 1] = d4.y; dv[4 * q + 2] = d4.z; dv[4 * q + 3] = d4.w;
        av[4 * q] = a4.x; av[4 * q + 1] = a4.y; av[4 * q + 2] = a4.z; av[4 * q + 3] = a4.w;
      }
      rs_next = live ? __ldg(p.rstd + row) : 0.f;
    };
    auto prefetch_da = [&](int it) {                     // L2 prefetch of a later tile's D / A rows (8 KB each)
      if (et == 0 && it < nmine) {
        const long long rn = ((long long)blockIdx.x + (long long)it * gridDim.x) * kTok;
        const long long rows = min((long long)kTok, p.N - rn);
        if (rows > 0) {
          prefetch_l2_bulk(p.D + rn * K, (uint32_t)(rows * K * 4));
          prefetch_l2_bulk(p.A + rn * K, (uint32_t)(rows * K * 4));
        }
      }
    };
    prefetch_da(1);
    load_da(0);
    for (int it = 0; it < nmine; ++it) {
      const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
      const long long row = tile * kTok + et;
      const bool live = row < p.N;
      const float rs = rs_next;
      TR(10, it);
      prefetch_da(it + 2);
      // [A | r] tile free: S3 / S5 of the previous tile have completed
      mbar_wait(&bars[B_AREMPTY], (uint32_t)((it & 1) ^ 1));
      TR(11, it);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t w1[4], w2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2_bf(av[8 * j + 2 * e], av[8 * j + 2 * e + 1], w1[e], w2[e]);
        const uint32_t off = sw128((uint32_t)et, (uint32_t)j * 16u);
        sts128(sAR32 + off, w1[0], w1[1], w1[2], w1[3]);
        sts128(sAR32 + kBlk + off, w2[0], w2[1], w2[2], w2[3]);
      }
      fence_async_smem();
      named_bar(1, 64);
      if (et == 0) mbar_arrive(&bars[B_AFULL]);
      TR(12, it);
      // producer row sums of this tile
      mbar_wait(&bars[B_PFULL], (uint32_t)(it & 1));
      TR(13, it);
      const int buf = it & 1;
      const float zz = sProw[buf * 3 * kTok + et];
      const float p1 = sProw[buf * 3 * kTok + kTok + et];
      const float p2 = sProw[buf * 3 * kTok + 2 * kTok + et];
      mbar_wait(&bars[B_G1FULL0 + buf], (uint32_t)((it >> 1) & 1));
      TR(14, it);
      tc_fence_after();
      float gv[32];
      tmem_ld32(tmem + lane_addr + kColG1 + (uint32_t)(buf * 32), gv);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_G1EMPTY0 + buf]);
      // softmin backward + cdist ratio
      float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
      for (int k = 0; k < 32; k += 2) {
        gv[k] = fmaf(sc * dv[k], dv[k] * av[k], gv[k]);              // gA_tot = G1 + sc D^2 A
        gv[k + 1] = fmaf(sc * dv[k + 1], dv[k + 1] * av[k + 1], gv[k + 1]);
        dot0 = fmaf(gv[k], av[k], dot0);
        dot1 = fmaf(gv[k + 1], av[k + 1], dot1);
      }
      const float dot = dot0 + dot1;
      float rsum0 = 0.f, rsum1 = 0.f, sT0 = 0.f, sT1 = 0.f, sG0 = 0.f, sG1 = 0.f;
      const float hz = 0.5f * zz;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float gd = av[k] * fmaf(sc * dv[k], av[k], -p.alpha * (gv[k] - dot));
        const float r = (dv[k] == 0.f || !live) ? 0.f : gd * fast_rcp(dv[k]);
        gv[k] = r;
        const float T = fmaf(-0.5f * dv[k], dv[k], hz + lds32f_const(sHc32 + 4 * k));
        if (k & 1) { rsum1 += r; sT1 = fmaf(r, T, sT1); sG1 = fmaf(r, lds32f_const(sCg32 + 4 * k), sG1); }
        else { rsum0 += r; sT0 = fmaf(r, T, sT0); sG0 = fmaf(r, lds32f_const(sCg32 + 4 * k), sG0); }
      }
      const float rsum = rsum0 + rsum1, sT = sT0 + sT1, sG = sG0 + sG1;
      const float s1 = (rsum * p1 - sG) * invC, s2 = (rsum * p2 - sT) * invC;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t w1[4], w2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2_bf(gv[8 * j + 2 * e], gv[8 * j + 2 * e + 1], w1[e], w2[e]);
        const uint32_t off = sw128((uint32_t)et, 64u + (uint32_t)j * 16u);
        sts128(sAR32 + off, w1[0], w1[1], w1[2], w1[3]);
        sts128(sAR32 + kBlk + off, w2[0], w2[1], w2[2], w2[3]);
      }
      float* scal = sScal + (it & 1) * 4 * kTok;
      scal[et] = rsum; scal[kTok + et] = s1 * rs; scal[2 * kTok + et] = s2 * rs; scal[3 * kTok + et] = rs;
      fence_async_smem();
      named_bar(1, 64);
      if (et == 0) mbar_arrive(&bars[B_RFULL]);
      TR(15, it);
      // next tile's D / A rows: issued after the fence / barrier above (which would wait for them),
      // their latency is covered by the butterfly and the wait for the [A | r] tile
      load_da(it + 1);
      butterfly<1>(gv, lane);                            // lane l: sum over this warp's 32 rows of r[:, l]
      rcol_acc += gv[0];
      TR(16, it);
    }
    p.part_rcol[((size_t)blockIdx.x * 2 + warp) * K + lane] = rcol_acc;
  } else if (is_e) {
    // ======================================================================= E3
    const int e3 = (warp >> 2) - 1;                      // 64-channel group
    const int q = warp & 1;                              // token half
    float accq[2][8];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) accq[c][j] = 0.f;
    if (e3 < NCB) {
      const int rl = q * 32 + lane;                      // token row in the tile = TMEM lane
      const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
      const uint64_t pol = policy_evict_first();
      const uint32_t stg32 = smem_u32(smem + pl.stg_off) + (uint32_t)(e3 * 2 + q) * 4096u;
      const uint32_t rx = (uint32_t)(rl & 7);
      const uint32_t sGam32 = smem_u32(sGam), sBet32 = smem_u32(sBet);
      for (int it = 0; it < nmine; ++it) {
        const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
        const long long row0 = tile * kTok;
        const int buf = it & 1;
        TR(19, it);
        mbar_wait(&bars[B_RFULL], (uint32_t)(it & 1));
        TR(20, it);
        const float* scal = sScal + (it & 1) * 4 * kTok;
        const float rsum = scal[rl], s1r = scal[kTok + rl], s2r = scal[2 * kTok + rl], rs = scal[3 * kTok + rl];
        mbar_wait(&bars[B_ACCFULL], (uint32_t)(it & 1));
        TR(21, it);
        tc_fence_after();
        const uint32_t xrow = sX32 + buf * pl.xbuf + (uint32_t)e3 * kBlk + (uint32_t)rl * 128u;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const int c0 = e3 * 64 + ch * 32;
          float acc[32], xh[32];
          tmem_ld32(tmem + lane_addr + kColAcc + (uint32_t)c0, acc);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t off = (((uint32_t)(ch * 4 + j) ^ rx) << 4);
            const uint4 h = lds128u(xrow + off);
            const uint4 l = lds128u(xrow + pl.xterm + off);
            xh[8 * j + 0] = bf_lo(h.x) + bf_lo(l.x); xh[8 * j + 1] = bf_hi(h.x) + bf_hi(l.x);
            xh[8 * j + 2] = bf_lo(h.y) + bf_lo(l.y); xh[8 * j + 3] = bf_hi(h.y) + bf_hi(l.y);
            xh[8 * j + 4] = bf_lo(h.z) + bf_lo(l.z); xh[8 * j + 5] = bf_hi(h.z) + bf_hi(l.z);
            xh[8 * j + 6] = bf_lo(h.w) + bf_lo(l.w); xh[8 * j + 7] = bf_hi(h.w) + bf_hi(l.w);
          }
          // the TMA store that last read this warp's staging block has finished reading
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 gm = lds128f_const(sGam32 + (uint32_t)(c0 + 4 * j) * 4u);
            const float4 be = lds128f_const(sBet32 + (uint32_t)(c0 + 4 * j) * 4u);
            float o[4];
            const float gmv[4] = {gm.x, gm.y, gm.z, gm.w}, bev[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float xhv = xh[4 * j + e];
              const float z = fmaf(xhv, gmv[e], bev[e]);
              const float gz = fmaf(z, rsum, -acc[4 * j + e]);
              o[e] = fmaf(-xhv, s2r, fmaf(gz * gmv[e], rs, -s1r));
              xh[4 * j + e] = xhv * xhv * rsum;          // Q contribution
            }
            sts128f(stg32 + (uint32_t)lane * 128u + ((((uint32_t)j) ^ (uint32_t)(lane & 7)) << 4), o[0], o[1], o[2], o[3]);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d_hint(&mapGx, stg32, c0, (int)(row0 + q * 32), pol);
            bulk_commit();
          }
          if (ch == 1) {                                 // accumulator and xhat fully consumed: release them early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&bars[B_ACCEMPTY]); mbar_arrive(&bars[B_XEMPTY0 + buf]); }
          }
          butterfly<8>(xh, lane);
#pragma unroll
          for (int j = 0; j < 8; ++j) accq[ch][j] += xh[j];
        }
        TR(22, it);
      }
      if (lane == 0) bulk_wait0();
      // finish the butterflies: lane l ends with Q of channel c0 + l over this warp's rows
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        butterfly_tail<4, 8>(accq[ch], lane);
        p.part_ln[((size_t)blockIdx.x * 2 + q) * C + e3 * 64 + ch * 32 + lane] = accq[ch][0];
      }
    }
  } else if (lane == 0) {
    // ======================================================================= MMA (warp 15, one thread)
    mbar_expect_tx(&bars[B_CEN], 2u * pl.cterm);
    for (uint32_t off = 0; off < 2u * pl.cterm; off += 16384u)
      bulk_g2s(smem + pl.cen_off + off, p.cimage + off, min(16384u, 2u * pl.cterm - off), &bars[B_CEN]);
    const uint32_t idesc1 = instr_desc(kFmtBF16, 128, K, 0, 0);      // S1: tokens x K, both K-major
    const uint32_t idesc3 = instr_desc(kFmtBF16, 128, C, 0, 1);      // S3: tokens x C, cen MN-major
    const uint32_t idesc5 = instr_desc(kFmtBF16, 128, K, 1, 1);      // S5: channels x K, both MN-major
    constexpr int pi[3] = {0, 1, 0}, pj[3] = {1, 0, 0};              // small terms first
    // descriptor bases: hi word is constant per layout, lo word = (address >> 4) | LBO field
    constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024, version 1, SWIZZLE_128B
    const uint32_t loG_k = (sG32 >> 4) & 0x3FFFu, loG_mn = loG_k | ((kBlk >> 4) << 16);
    const uint32_t loAR_k = (sAR32 >> 4) & 0x3FFFu, loAR_mn = loAR_k | ((kBlk >> 4) << 16);
    const uint32_t loC_k = (sCen32 >> 4) & 0x3FFFu, loC_mn = loC_k | (((uint32_t)K * 128u >> 4) << 16);
    const uint32_t loX_mn0 = ((sX32 >> 4) & 0x3FFFu) | ((kBlk >> 4) << 16);
    mbar_wait(&bars[B_CEN], 0);
    int n1 = 0, n5 = 0, n3 = 0;
    while (n3 < nmine) {
      if (n1 < nmine && mbar_try_wait(&bars[B_PFULL], (uint32_t)(n1 & 1)) &&
          mbar_try_wait(&bars[B_G1EMPTY0 + (n1 & 1)], (uint32_t)(((n1 >> 1) & 1) ^ 1))) {
        tc_fence_after();
        const uint32_t d = tmem + kColG1 + (uint32_t)((n1 & 1) * 32);
#pragma unroll
        for (int t = 0; t < 3; ++t) {
#pragma unroll
          for (int kk = 0; kk < C / 16; ++kk) {
            const uint64_t ad = desc_at(loG_k, kHi, pi[t] * pl.gterm + (kk >> 2) * kBlk + (kk & 3) * 32u);
            const uint64_t bd = desc_at(loC_k, kHi, pj[t] * pl.cterm + (kk >> 2) * (K * 128u) + (kk & 3) * 32u);
            mma_bf16(d, ad, bd, idesc1, t > 0 || kk > 0);
          }
        }
        mma_commit(&bars[B_G1FULL0 + (n1 & 1)]);
        TR(30, n1);
        ++n1;
      }
      if (n5 < n1 && mbar_try_wait(&bars[B_AFULL], (uint32_t)(n5 & 1))) {
        tc_fence_after();
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          const uint32_t d = tmem + kColP1 + (uint32_t)(mb * 32);
#pragma unroll
          for (int t = 0; t < 3; ++t) {
#pragma unroll
            for (int ks = 0; ks < kTok / 16; ++ks) {
              const uint64_t ad = desc_at(loG_mn, kHi, pi[t] * pl.gterm + (uint32_t)mb * 2u * kBlk + (uint32_t)ks * 2048u);
              const uint64_t bd = desc_at(loAR_mn, kHi, pj[t] * kBlk + (uint32_t)ks * 2048u);
              mma_f16(d, ad, bd, idesc5, (t > 0 || ks > 0 || n5 > 0) ? 1u : 0u);
            }
          }
        }
        mma_commit(&bars[B_GEMPTY]);
        TR(31, n5);
        ++n5;
      }
      if (n3 < n5 && mbar_try_wait(&bars[B_RFULL], (uint32_t)(n3 & 1)) &&
          mbar_try_wait(&bars[B_ACCEMPTY], (uint32_t)((n3 & 1) ^ 1))) {
        tc_fence_after();
        {
          const uint32_t d = tmem + kColAcc;
#pragma unroll
          for (int t = 0; t < 3; ++t) {
#pragma unroll
            for (int ks = 0; ks < K / 16; ++ks) {
              const uint64_t ad = desc_at(loAR_k, kHi, pi[t] * kBlk + 64u + (uint32_t)ks * 32u);
              const uint64_t bd = desc_at(loC_mn, kHi, pj[t] * pl.cterm + (uint32_t)(2 * ks) * 1024u);
              mma_bf16(d, ad, bd, idesc3, t > 0 || ks > 0);
            }
          }
          mma_commit(&bars[B_ACCFULL]);
        }
        const uint32_t loX_mn = loX_mn0 + (uint32_t)(n3 & 1) * (pl.xbuf >> 4);
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          const uint32_t d = tmem + kColP2 + (uint32_t)(mb * 32);
#pragma unroll
          for (int t = 0; t < 3; ++t) {
#pragma unroll
            for (int ks = 0; ks < kTok / 16; ++ks) {
              const uint64_t ad = desc_at(loX_mn, kHi, pi[t] * pl.xterm + (uint32_t)mb * 2u * kBlk + (uint32_t)ks * 2048u);
              const uint64_t bd = desc_at(loAR_mn, kHi, pj[t] * kBlk + 64u + (uint32_t)ks * 2048u);
              mma_f16(d, ad, bd, idesc5, (t > 0 || ks > 0 || n3 > 0) ? 1u : 0u);
            }
          }
        }
        mma_commit(&bars[B_XEMPTY0 + (n3 & 1)]);
        mma_commit(&bars[B_AREMPTY]);
        TR(32, n3);
        ++n3;
      }
    }
    mma_commit(&bars[B_DONE]);
    mbar_wait(&bars[B_DONE], 0);
  }
  // ---- drain the whole-kernel accumulators: lane = channel, 32 columns = centroids
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 4) {
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
      const int c = mb * 128 + warp * 32 + lane;
      float v[32];
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        tmem_ld32(tmem + lane_addr + (which ? kColP2 : kColP1) + (uint32_t)(mb * 32), v);
        if (c < C) {
          float4* o = reinterpret_cast<float4*>(p.part_p + (((size_t)blockIdx.x * 2 + which) * C + c) * K);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) { tc_fence_after(); tmem_dealloc(tmem, ncols); }
}

// One warp per channel c, lane = centroid k (K == 32), partials added in fixed order (deterministic):
//   gcenters[k,c] = sum_b P1[b][c][k] - gamma_c sum_b P2[b][c][k] + (cen[k,c] - beta_c) rcol_k
//   g_beta[c]  = gamma_c sum_k P2[c,k] + beta_c sum_k rcol_k - sum_k rcol_k cen[k,c]
//   g_gamma[c] = gamma_c Q[c] + beta_c sum_k P2[c,k] - sum_k cen[k,c] P2[c,k]
__global__ void __launch_bounds__(256)
cluster_bwd_tc_finalize_kernel(const float* __restrict__ part_p, const float* __restrict__ part_rcol,
                               const float* __restrict__ part_q, const float* __restrict__ centers,
                               const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                               int nb, int K, int C, float* __restrict__ gcenters,
                               float* __restrict__ gw, float* __restrict__ gb) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), k = threadIdx.x & 31;
  if (c >= C) return;
  const int KC = K * C;
  float s1 = 0.f, s2 = 0.f, rc = 0.f, qv = 0.f;
  for (int b = 0; b < nb; ++b) {
    s1 += part_p[((size_t)b * 2) * KC + c * K + k];
    s2 += part_p[((size_t)b * 2 + 1) * KC + c * K + k];
    rc += part_rcol[((size_t)b * 2) * K + k] + part_rcol[((size_t)b * 2 + 1) * K + k];
  }
  for (int b = k; b < 2 * nb; b += 32) qv += part_q[(size_t)b * C + c];
  const float g = ln_w[c], be = ln_b[c], cen = centers[(size_t)k * C + c];
  gcenters[(size_t)k * C + c] = s1 - g * s2 + (cen - be) * rc;
  const float sp2 = warp_sum(s2), scp = warp_sum(cen * s2), src = warp_sum(rc), scr = warp_sum(rc * cen);
  qv = warp_sum(qv);
  if (k == 0) {
    gb[c] = g * sp2 + be * src - scr;
    gw[c] = g * qv + be * sp2 - scp;
  }
}

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

// [rows, cols] fp32 row-major tensor, box = 32 columns x 32 rows, SWIZZLE_128B
static int make_map(CUtensorMap* m, float* base, long long rows, int cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return VADC_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, 32};
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

}  // namespace bt

bool bwd_tc_shape_ok(long long N, int C, int K) { return bt::shape_ok(N, C, K); }

size_t bwd_tc_workspace_bytes(long long N, int C, int K) {
  if (!bt::shape_ok(N, C, K)) return 0;
  const size_t g = (size_t)sm_count();
  return align_up((size_t)2 * K * C * 2, 256) + align_up((size_t)3 * K * sizeof(float), 256) +
         align_up(g * 2 * C * K * sizeof(float), 256) + align_up(g * 2 * K * sizeof(float), 256) +
         align_up(g * 4 * C * sizeof(float), 256) + 256;
}

int launch_cluster_bwd_tc(const float* x, const float* mu, const float* rstd, const float* ln_w,
                          const float* ln_b, const float* centers, const float* D, const float* A,
                          const float* gR, const float* g_loss_sq, long long N, int C, int K, float alpha,
                          float* gx, float* gcenters, float* g_ln_w, float* g_ln_b, void* workspace,
                          size_t workspace_bytes, cudaStream_t st) {
  if (!bt::shape_ok(N, C, K)) return VADC_ERR_UNSUPPORTED;
  if (!vadc_device_ok()) return VADC_ERR_NO_DEVICE;
  if (workspace_bytes < bwd_tc_workspace_bytes(N, C, K)) return VADC_ERR_WORKSPACE;
  Carver ws(workspace, workspace_bytes);
  const size_t g = (size_t)sm_count();
  uint8_t* image = ws.take<uint8_t>((size_t)2 * K * C * 2);
  float* cvec = ws.take<float>(3 * K);
  float* part_p = ws.take<float>(g * 2 * C * K);
  float* part_rcol = ws.take<float>(g * 2 * K);
  float* part_ln = ws.take<float>(g * 4 * C);
  const int grid = (int)std::min<long long>((N + bt::kTok - 1) / bt::kTok, (long long)g);

  CUtensorMap mGx;
  int rc;
  if ((rc = bt::make_map(&mGx, gx, N, C))) return rc;
  bt::centroid_prep_bwd_kernel<<<K, 256, 0, st>>>(centers, ln_w, ln_b, K, C, image, cvec);
  VADC_CHECK_LAUNCH("centroid_prep_bwd_kernel");

  const size_t smem = bt::plan(C).total + 1024;
  unsigned long long* trace = nullptr;
  const char* trace_path = getenv("VADC_BWD_TRACE");       // debugging only: synchronises and writes a text file
  const size_t trace_bytes = (size_t)16 * 1024 * 2 * sizeof(unsigned long long);
  if (trace_path) { VADC_CUDA(cudaMalloc(&trace, trace_bytes)); VADC_CUDA(cudaMemsetAsync(trace, 0, trace_bytes, st)); }
  bt::Params p{x, gR, D, A, mu, rstd, ln_w, ln_b, image, cvec, g_loss_sq, part_p, part_rcol, part_ln,
               N, alpha, getenv("VADC_BWD_PF") ? atoi(getenv("VADC_BWD_PF")) : 1, trace};
  bool launched = false;
#define BT_CASE(F4_)                                                                                   \
  if (C == 32 * F4_) {                                                                                 \
    auto kern = trace ? bt::cluster_bwd_tc_kernel<F4_, true> : bt::cluster_bwd_tc_kernel<F4_, false>;    \
    VADC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    kern<<<grid, bt::kThreads, smem, st>>>(mGx, p);                                                    \
    launched = true;                                                                                   \
  }
  BT_CASE(2) BT_CASE(4) BT_CASE(6)
#undef BT_CASE
  if (!launched) return VADC_ERR_UNSUPPORTED;
  VADC_CHECK_LAUNCH("cluster_bwd_tc_kernel");
  if (trace) {
    VADC_CUDA(cudaStreamSynchronize(st));
    unsigned long long* h = (unsigned long long*)malloc(trace_bytes);
    VADC_CUDA(cudaMemcpy(h, trace, trace_bytes, cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(trace_path, "w")) {
      for (int w = 0; w < 16; ++w)
        for (int i = 0; i < 1024; ++i) {
          const unsigned long long a = h[((size_t)w * 1024 + i) * 2], c = h[((size_t)w * 1024 + i) * 2 + 1];
          if (c) fprintf(f, "%d %d %d %llu\n", w, (int)(a >> 32), (int)(a & 0xffffffffu), c);
        }
      fclose(f);
    }
    free(h);
    cudaFree(trace);
  }
  bt::cluster_bwd_tc_finalize_kernel<<<(C + 7) / 8, 256, 0, st>>>(part_p, part_rcol, part_ln, centers, ln_w,
                                                                        ln_b, grid, K, C, gcenters, g_ln_w, g_ln_b);
  VADC_CHECK_LAUNCH("cluster_bwd_tc_finalize_kernel");
  return VADC_OK;
}

}  // namespace vadc
