// cluster_bwd_tc2.cu — C2: fused backward of the cluster head on tcgen05 / TMEM / TMA, K == 32,
// training-graph case (gradients arrive through x_rec and the fused cluster loss only).  Second generation.
//
//   autograd of model/cluster.py:81-99 + model/backbone.py:98 (loss.backward(), main_predict.py:296)
//   in ONE persistent warp-specialised kernel; per token it reads x, gR, D, A (+ mu, rstd, rowstats) once and
//   writes gx once: 12 C + 8 K bytes (fused-loss variant).  64-token tiles.
//
// What changed against the first generation (cluster_bwd_tc.cu, 0.57 of the HBM roofline: its seven producer warps
// converted BOTH x and gR through a register pipeline and set the pace of the kernel):
//   * x never passes through the producers.  Each E3 warp TMA-loads the raw fp32 [32 tokens x 32 channels] boxes it
//     owns (SWIZZLE_128B, one tile ahead, mbarrier complete_tx), forms xhat = (x - mu) rstd exactly in fp32 for the
//     LayerNorm backward, writes the two-term bf16 split of xhat into the operand tile of the r^T xhat contraction
//     (which therefore runs AFTER E3 of its tile: it is background work nothing waits for), and stages gx in the SAME
//     box it read x from (same swizzled bytes) for the TMA store — the separate 24 KB store staging is gone and the
//     store's shared-memory read is followed by the next tile's load into the box.
//   * the producers convert gR (and the A rows) only: half the work.  Their global loads are bulk copies
//     (cp.async.bulk, mbarrier complete_tx) into a private 4-row slot per warp, issued one unit ahead right after
//     the slot was read — no load sits in registers, nothing waits on a scoreboard, and a load is never tied to
//     the availability of the operand tile it will be converted into.
//   * the r operand tile is double buffered (E1 of tile t+1 must not wait for the r^T xhat contraction of tile t).
//
//   warps 0,1,6,7,10,11,14  PRODUCERS  gR rows -> two-term bf16 split into the SWIZZLE_128B operand tile (4-row
//                           units out of the warp's slot); the bf16 split of the A rows into [A_hi | A_lo]
//   warp 15 (one thread)    MMA        S1  G1 = gR [cen_hi; cen_lo]^T        [tokens x 2K]  (TMEM, double buffer)
//                                      S3  acc = r cen                       [tokens x C]
//                                      S5a PT[0:64]   += [A_hi | A_lo]^T gR  [2K x C]  (TMEM, whole kernel)
//                                      S5b PT[64:128] += [r_hi | r_lo]^T xhat
//   warps 2,3               E1         thread = token: softmin backward + cdist ratio -> r; the LayerNorm-backward
//                                      row statistics in closed form; bf16 split of r into [r_hi | r_lo].  S1's A
//                                      descriptor starts 64 rows BEFORE the gR tile, so G1 lands in TMEM lanes
//                                      64..127: E1 runs on the schedulers of warps 2,3 (mod 4)
//   warps 4,5,8,9,12,13     E3         thread = token x 64 channels: raw x box -> xhat; gz = z rsum - acc, LayerNorm
//                                      backward, gx -> box -> TMA store; xhat hi / lo -> operand tile;
//                                      Q[c] = sum_n xhat^2 rsum (register butterfly) for g_gamma
//
// Contraction shapes, operand precision (exact two-term bf16 splits, fp32 accumulate) and the closed forms of the
// LayerNorm-backward row means / g_gamma / g_beta are those of the first generation:
//   sum_c gg        = rsum * sum_c z gamma      - sum_k r_k (cen_k . gamma)
//   sum_c gg xhat   = rsum * sum_c z gamma xhat - sum_k r_k T_k,  T_k = (|z|^2 + |cen_k|^2 - D_k^2) / 2 - beta.cen_k
//   gcenters = P1 - gamma * P2 + (cen - beta) rcol,  P1 = A^T gR, P2 = r^T xhat, rcol = colsum(r)
//   g_beta  = gamma * sum_k P2[k,:] + beta * sum(rcol) - rcol . cen
//   g_gamma = gamma * Q + beta * sum_k P2[k,:] - sum_k cen[k,:] * P2[k,:]
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

namespace bt2 {

// the S5 contractions are background work of the MMA thread (nothing waits for their issue): their loops may stay rolled
#ifdef VADC_BWD_ROLL_S5
#define VADC_S5_UNROLL _Pragma("unroll 1")
#else
#define VADC_S5_UNROLL _Pragma("unroll")
#endif

constexpr int kTok = 64;                     // tokens per tile
constexpr int kK = 32;                       // centroids
constexpr int kThreads = 512;
constexpr int kProd = 7;                     // producer warps: 6 convert gR, the last one the A rows
constexpr int kMmaWarp = 15;
constexpr uint32_t kBlk = kTok * 128u;       // one [64 rows x 128 B] operand block
constexpr uint32_t kBox = 32u * 128u;        // one [32 rows x 128 B] fp32 box (TMA load of x / TMA store of gx)

struct Plan {
  uint32_t xr_off, xh_off, g_off, ta_off, zero_off, tr_off, cen_off, da_off, scal_off, gam_off, bet_off, cvec_off, misc_off, total;
  uint32_t xterm;
};

__host__ __device__ inline Plan plan(int C) {
  Plan p;
  const uint32_t ncb = (uint32_t)C / 64u;
  uint32_t off = 0;
  p.xterm = ncb * kBlk;
  p.xr_off = off; off += 2u * ncb * 2u * kBox;             // [E3 warp = (channel group, token half)][2 boxes]: raw x, then gx
  p.xh_off = off; off += 2u * p.xterm;                     // xhat operand [2 terms][C/64][64 x 128 B]
  p.g_off = off; off += 2u * p.xterm;                      // gR operand   [2 terms][C/64][64 x 128 B]
  p.ta_off = off; off += kBlk;                             // [64 tokens x 128 B]: A_hi (32 slots) | A_lo
  p.zero_off = off; off += kBlk;                           // constant zeros (the other half of S5's M)
  p.tr_off = off; off += 2u * kBlk;                        // [parity][64 tokens x 128 B]: r_hi | r_lo
  p.cen_off = off; off += ncb * kBlk;                      // [C/64][64 rows x 128 B]: rows 0..31 cen_hi, 32..63 cen_lo
  p.da_off = off; off += 2u * kBlk;                        // E1: the tile's D rows and A rows, [64 x 128 B] each (TMA, SWIZZLE_128B)
  p.scal_off = off; off += 2u * 5u * kTok * 4u;            // E1 row scalars [parity][rsum, s1r, s2r, rs, -mu rs][64]
  p.gam_off = off; off += (uint32_t)C * 4u;
  p.bet_off = off; off += (uint32_t)C * 4u;
  p.cvec_off = off; off += 2u * kK * 4u;                   // hc = |c|^2/2 - beta.c ; cg = gamma.c
  p.misc_off = off; off += 512u;                           // mbarriers [0, 8 B_COUNT), TMEM slot at +448
  p.total = off;
  return p;
}

enum { B_CEN = 0, B_GFULL, B_GEMPTY, B_G1FULL0, B_G1FULL1, B_G1EMPTY0, B_G1EMPTY1, B_AFULL, B_AEMPTY,
       B_RFULL0, B_RFULL1, B_REMPTY0, B_REMPTY1, B_ACCFULL, B_ACCEMPTY, B_XHFULL, B_XHEMPTY, B_DONE,
       B_XR0, B_DAFULL = B_XR0 + 12, B_COUNT };
static_assert(B_COUNT * 8 <= 448, "mbarriers overlap the TMEM slot");

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
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
// non-blocking phase test (mbarrier.try_wait may suspend the thread for a system-dependent time before it returns false)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// six non-blocking probes issued back to back: ONE ~150-cycle round trip (separate probes serialise on their predicates)
__device__ __forceinline__ uint32_t mbar_test6(const uint32_t (&bar)[6], const uint32_t (&parity)[6]) {
  uint32_t mask;
  asm volatile(
      "{\n\t.reg .pred q0, q1, q2, q3, q4, q5;\n\t.reg .u32 t0, t1, t2, t3, t4, t5;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q0, [%1], %7;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q1, [%2], %8;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q2, [%3], %9;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q3, [%4], %10;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q4, [%5], %11;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q5, [%6], %12;\n\t"
      "selp.u32 t0, 1, 0, q0;\n\tselp.u32 t1, 2, 0, q1;\n\tselp.u32 t2, 4, 0, q2;\n\t"
      "selp.u32 t3, 8, 0, q3;\n\tselp.u32 t4, 16, 0, q4;\n\tselp.u32 t5, 32, 0, q5;\n\t"
      "or.b32 t0, t0, t1;\n\tor.b32 t2, t2, t3;\n\tor.b32 t4, t4, t5;\n\tor.b32 t0, t0, t2;\n\tor.b32 %0, t0, t4;\n\t}"
      : "=r"(mask)
      : "r"(bar[0]), "r"(bar[1]), "r"(bar[2]), "r"(bar[3]), "r"(bar[4]), "r"(bar[5]),
        "r"(parity[0]), "r"(parity[1]), "r"(parity[2]), "r"(parity[3]), "r"(parity[4]), "r"(parity[5]) : "memory");
  return mask;
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
__device__ __forceinline__ void tma_load_2d_u32(const void* tmap, uint32_t smem_dst, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s_u32(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));   // one MUFU (the non-ftz form expands into a subnormal-safe branch)
  return r;
}

// exact two-term bf16 split of a pair: (a, b) = (lo16(p1), hi16(p1)) + (lo16(p2), hi16(p2)) + O(2^-17)
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);        // a -> low half, b -> high half
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf_hi(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }
__device__ __forceinline__ float2 bf_pair(uint32_t p) { return make_float2(bf_lo(p), bf_hi(p)); }
__device__ __forceinline__ void split2_bf(float2 v, uint32_t& p1, uint32_t& p2) {
  p1 = pack_bf2(v.x, v.y);
  const float2 l = sub2(v, bf_pair(p1));
  p2 = pack_bf2(l.x, l.y);
}
__device__ __forceinline__ void split2_bf(float a, float b, uint32_t& p1, uint32_t& p2) {
  split2_bf(make_float2(a, b), p1, p2);
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
// prologue: two-term bf16 centroid image ([C/64][64 rows x 128 B], SWIZZLE_128B: rows 0..31 the
// high terms, rows 32..63 the low terms) and the per-centroid constants
// hc_k = |c_k|^2 / 2 - beta.c_k,  cg_k = gamma.c_k
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
centroid_prep_bwd_kernel(const float* __restrict__ centers, const float* __restrict__ ln_w,
                         const float* __restrict__ ln_b, int K, int C, uint8_t* __restrict__ image,
                         float* __restrict__ cvec) {
  const int k = blockIdx.x;
  __shared__ float red[32];
  float s = 0.f, sb = 0.f, sg = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = centers[(size_t)k * C + c];
    s += v * v; sb += v * ln_b[c]; sg += v * ln_w[c];
    const __nv_bfloat16 h1 = __float2bfloat16_rn(v);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(v - __bfloat162float(h1));
    const uint32_t blk = (uint32_t)(c / 64) * kBlk;
    *reinterpret_cast<__nv_bfloat16*>(image + blk + sw128(k, (c % 64) * 2)) = h1;
    *reinterpret_cast<__nv_bfloat16*>(image + blk + sw128(K + k, (c % 64) * 2)) = h2;
  }
  s = block_sum<float>(s, red);
  sb = block_sum<float>(sb, red);
  sg = block_sum<float>(sg, red);
  if (threadIdx.x == 0) { cvec[k] = 0.5f * s - sb; cvec[K + k] = sg; }
}

struct Params {
  const float* x; const float* gR; const float* D; const float* A; const float* mu; const float* rstd;
  const float* rowstats;                                 // [N,4]: |z|^2, sum z gamma, sum z gamma xhat, 0 (from the forward)
  const float* ln_w; const float* ln_b; const uint8_t* cimage; const float* cvec; const float* g_loss_sq;
  float* part_p; float* part_rcol; float* part_q;        // [grid][128 slots][C], [grid][2][32], [grid][2][C]
  long long N; float alpha; int pf;
  int mma_sleep;                                         // ns the MMA thread sleeps when nothing is ready (0 = spin)
  unsigned long long* trace;                             // debugging: per-warp event log of one CTA (VADC_BWD_TRACE builds)
  int trace_cta;
};

template <int F4, bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
cluster_bwd_tc2_kernel(const __grid_constant__ CUtensorMap mapGx, const __grid_constant__ CUtensorMap mapX,
                       const __grid_constant__ CUtensorMap mapD, const __grid_constant__ CUtensorMap mapA, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int C = F4 * 32, K = kK, NCB = C / 64;
  const Plan pl = plan(C);
  float* sScal = reinterpret_cast<float*>(smem + pl.scal_off);
  float* sGam = reinterpret_cast<float*>(smem + pl.gam_off);
  float* sBet = reinterpret_cast<float*>(smem + pl.bet_off);
  float* sHc = reinterpret_cast<float*>(smem + pl.cvec_off);
  float* sCg = sHc + K;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.misc_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + pl.misc_off + 448);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // event trace (debug): entry = {code << 32 | tile, clock64}
  constexpr int kTraceCap = 1024;
  int trace_n = 0;
  auto TR = [&](int code, int it) {
    if constexpr (!TRACE) return;
    if (p.trace && blockIdx.x == p.trace_cta && lane == 0 && trace_n < kTraceCap) {
      unsigned long long* e = p.trace + ((size_t)warp * kTraceCap + trace_n) * 2;
      e[0] = ((unsigned long long)code << 32) | (unsigned)it; e[1] = (unsigned long long)clock64();
      ++trace_n;
    }
  };
  const long long ntiles = (p.N + kTok - 1) / kTok;
  const int nmine = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles of this CTA (>= 1)
  constexpr uint32_t kColG1 = 0, kColAcc = 128, kColPT = 128 + C;   // G1: two buffers of 64 columns
  constexpr uint32_t ncols = (kColPT + C <= 256) ? 256u : 512u;
  constexpr int kE3Warps = 2 * NCB;

  if (tid == 0) {
    mbar_init(&bars[B_CEN], 1);
    mbar_init(&bars[B_GFULL], kTok / 4);
    mbar_init(&bars[B_GEMPTY], 1);
    mbar_init(&bars[B_G1FULL0], 1); mbar_init(&bars[B_G1FULL1], 1);
    mbar_init(&bars[B_G1EMPTY0], 2); mbar_init(&bars[B_G1EMPTY1], 2);
    mbar_init(&bars[B_AFULL], 1); mbar_init(&bars[B_AEMPTY], 1);
    mbar_init(&bars[B_RFULL0], 1); mbar_init(&bars[B_RFULL1], 1);
    mbar_init(&bars[B_REMPTY0], 2); mbar_init(&bars[B_REMPTY1], 2);   // S5b's commit + the column sums of producer warp 6
    mbar_init(&bars[B_ACCFULL], 1);
    mbar_init(&bars[B_ACCEMPTY], kE3Warps);
    mbar_init(&bars[B_XHFULL], kE3Warps);
    mbar_init(&bars[B_XHEMPTY], 1);
    mbar_init(&bars[B_DONE], 1);
    for (int i = 0; i < 12; ++i) mbar_init(&bars[B_XR0 + i], 1);
    mbar_init(&bars[B_DAFULL], 1);
    fence_mbar_init();
    prefetch_tmap(&mapGx);
    prefetch_tmap(&mapX);
    prefetch_tmap(&mapD);
    prefetch_tmap(&mapA);
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, ncols);
  for (int c = tid; c < C; c += kThreads) {
    const float g = p.ln_w[c], b = p.ln_b[c];
    sGam[c] = g; sBet[c] = b;
  }
  for (int k = tid; k < 2 * K; k += kThreads) sHc[k] = p.cvec[k];
  for (uint32_t i = tid; i < kBlk / 16u; i += kThreads)
    *reinterpret_cast<uint4*>(smem + pl.zero_off + i * 16u) = make_uint4(0, 0, 0, 0);
  fence_async_smem();                                    // the zero block is an MMA operand
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sXH32 = smem_u32(smem + pl.xh_off), sG32 = smem_u32(smem + pl.g_off);
  const uint32_t sTA32 = smem_u32(smem + pl.ta_off), sTR32 = smem_u32(smem + pl.tr_off);
  const uint32_t sZero32 = smem_u32(smem + pl.zero_off);
  const uint32_t sCen32 = smem_u32(smem + pl.cen_off);

  // roles by warp (a warp reads TMEM lanes 32 (warp % 4)..; warp % 4 is also its scheduler / sub-partition):
  //   E1 2,3 (G1 in lanes 64..127)   E3 4,5,8,9,12,13 (accumulator lanes 0..63)   P 0,1,6,7,10,11 gR, 14 A rows   MMA 15
  // (tried: S3 split in two so that two E3 warps could read lanes 64..127 on the sub-partitions of warps 2,3 — 2+2+1+1
  //  E3 warps per sub-partition instead of 3+3+0+0: 13 % SLOWER, E1 then shares its sub-partition with an E3 warp and
  //  becomes the pace setter: the kernel is bound by issue / pipe throughput, moving work between sub-partitions is zero-sum)
  const bool is_e1 = warp == 2 || warp == 3;
  const bool is_e3 = (warp & 3) < 2 && warp >= 4;
  if (!is_e1 && !is_e3 && warp != kMmaWarp) {
    // ======================================================================= PRODUCERS
    // gR: producer warps 0..5, 16 four-row units per tile (warp w owns units w, w+6, w+12 of every tile).  A unit is the
    // rows {b, b+1, b+4, b+5}, b = 8 (v/2) + 2 (v%2): 8 lanes per row (lane j owns float4 chunks j, j+8, ...), and the two
    // rows that share a 16-lane store phase differ by 4 (disjoint banks after the 128B swizzle).  The raw rows arrive in
    // the warp's private slot by two bulk copies (rows b, b+1 and rows b+4, b+5 are contiguous in global memory).
    // The gR operand tile is single buffered, so between the release of tile t and the completion of tile t+1 nothing
    // but conversion should happen: a warp keeps its next TWO units in registers and a third one in (or on its way to) the
    // slot — all of a tile's units are on chip before the tile is released, whatever the load latency is at that moment
    // (first version: one unit in registers, one in the slot; every third unit of a warp paid the full latency, 6.7 k
    // cycles from release to completion in the event trace).
    // A rows: producer warp 6, two 32-row units per tile, loaded into registers while it waits for the tile.
    const int pw = warp < 2 ? warp : (warp >> 2) * 2 + (warp & 1);   // warps 0,1,6,7,10,11,14 -> 0..6
    const int lj = lane & 7, lg = lane >> 3;
    const int rsel = (lg & 1) * 4 + (lg >> 1);                       // row of this lane group inside the 8-row group: 0,4,1,5
    if (pw == kProd - 1) {
      float4 a[16];
      auto load_a = [&](int it) {
        const long long rb = ((long long)blockIdx.x + (long long)it * gridDim.x) * kTok;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const long long row = rb + (j >> 3) * 32 + ((j & 7) >> 1) * 8 + (j & 1) * 2 + rsel;
          a[j] = (it < nmine && row < p.N) ? ldg_nc(reinterpret_cast<const float4*>(p.A + row * K) + lj) : make_float4(0, 0, 0, 0);
        }
      };
      // ... and rcol = colsum(r) (for gcenters / g_beta): lane l adds up the 32-bit word l of every row of the r operand
      // tile, i.e. the slots 2l, 2l+1 of [r_hi | r_lo] (conflict free: a row's 32 words are a permutation of the banks),
      // instead of a 32 x 32 register butterfly in E1 (2.2 k cycles of E1's 10 k per tile in the event trace).  Order of
      // work: A rows of tile it+1 (wanted ~2 k cycles after S1 of tile it), then the sums of tile it-1 (its r tile
      // completes ~5 k cycles after S1 of tile it).
      float2 rc = make_float2(0.f, 0.f);
      auto colsum_r = [&](int t) {                       // tile t: wait for E1, add, hand the tile back
        mbar_wait(&bars[B_RFULL0 + (t & 1)], (uint32_t)((t >> 1) & 1));
        const uint32_t tr32 = sTR32 + (uint32_t)(t & 1) * kBlk;
        float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
#pragma unroll 8
        for (int r = 0; r < kTok; r += 2) {
          uint32_t w0, w1;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(tr32 + sw128((uint32_t)r, (uint32_t)lane * 4u)));
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w1) : "r"(tr32 + sw128((uint32_t)r + 1u, (uint32_t)lane * 4u)));
          acc0 = add2(acc0, bf_pair(w0));
          acc1 = add2(acc1, bf_pair(w1));
        }
        rc = add2(rc, add2(acc0, acc1));
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_REMPTY0 + (t & 1)]);
      };
      load_a(0);
#pragma unroll 1
      for (int it = 0; it < nmine; ++it) {
        mbar_wait(&bars[B_AEMPTY], (uint32_t)((it & 1) ^ 1));        // tile A is free once S5a of the previous tile has completed
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t r = (uint32_t)((j >> 3) * 32 + ((j & 7) >> 1) * 8 + (j & 1) * 2 + rsel);
          uint32_t a1, a2, b1, b2;
          split2_bf(a[j].x, a[j].y, a1, a2);
          split2_bf(a[j].z, a[j].w, b1, b2);
          sts64(sTA32 + sw128(r, (uint32_t)lj * 8u), a1, b1);            // A_hi: slots 4 lj .. 4 lj + 3
          sts64(sTA32 + sw128(r, 64u + (uint32_t)lj * 8u), a2, b2);      // A_lo
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_AFULL]);
        load_a(it + 1);
        if (it >= 2) colsum_r(it - 2);
      }
      if (nmine >= 2) colsum_r(nmine - 2);
      colsum_r(nmine - 1);
      // lanes 0..15 hold the r_hi slots 2l, 2l+1, lanes 16..31 the r_lo slots: rcol_k = hi + lo
      rc.x += __shfl_down_sync(0xffffffffu, rc.x, 16);
      rc.y += __shfl_down_sync(0xffffffffu, rc.y, 16);
      float* pr = p.part_rcol + (size_t)blockIdx.x * 2 * K;
      if (lane < 16) { pr[2 * lane] = rc.x; pr[2 * lane + 1] = rc.y; }
      pr[K + lane] = 0.f;
    } else {
      // gR: plain 128-bit global loads into register sets (one per unit of a tile), no shared-memory staging.  (Second version of this kernel: bulk copies into a private slot per warp.  The event
      // trace showed the issue of a bulk copy taking 1-2 k cycles, and a TMA store's shared-memory read 2.4 k: every
      // TMA operation of the CTA — the x boxes, the gx stores, the slots — queues in one unit, and a store that waits
      // for the memory system blocks the copies behind it.  The x boxes and gx stores stay on TMA, one tile ahead of
      // their use; the latency-critical gR path does not go through it any more.)
      struct Pos { int it, v; };
      auto load = [&](float4 (&R)[F4], const Pos& q) {               // global -> registers, asynchronous (scoreboard)
        if (q.it >= nmine) return;
        const int r = (q.v >> 1) * 8 + (q.v & 1) * 2 + rsel;
        const long long row = ((long long)blockIdx.x + (long long)q.it * gridDim.x) * kTok + r;
        const bool live = row < p.N;
        const float4* sr = reinterpret_cast<const float4*>(p.gR + row * C) + lj;
        TR(0, q.it);
#pragma unroll
        for (int i = 0; i < F4; ++i) R[i] = live ? ld_stream(sr + 8 * i) : make_float4(0, 0, 0, 0);
        if (p.pf > 0 && lg == 0 && (q.v & 1) == 0 && q.it + p.pf < nmine) {   // L2 prefetch: same 8-row group, pf tiles ahead
          const long long rn = ((long long)blockIdx.x + (long long)(q.it + p.pf) * gridDim.x) * kTok + (q.v >> 1) * 8;
          const char* base = reinterpret_cast<const char*>(p.gR + rn * C);
          const long long bytes = min(8ll, p.N - rn) * C * 4;
          for (long long o = (long long)lj * 128; o < bytes; o += 8 * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
        }
        TR(3, q.it);
      };
      auto convert = [&](const float4 (&R)[F4], const Pos& q) {      // registers -> two-term bf16 split in the operand tile
        const int it = q.it;
        const int r = (q.v >> 1) * 8 + (q.v & 1) * 2 + rsel;
        // byte offset of this lane's float4 number lj inside its row's first 64-channel block; float4 lj + 8 i sits in
        // block i/2 at chunk (lj/2 + 4 (i%2)) ^ (r%8): bit 6 of the offset flips with i%2 (all operand bases are 1 KB aligned)
        const uint32_t e0 = (uint32_t)r * 128u + (((uint32_t)(lj >> 1) ^ (uint32_t)(r & 7)) << 4) + (uint32_t)(lj & 1) * 8u;
        // G is free once S1 / S5a of the previous tile have completed
        mbar_wait(&bars[B_GEMPTY], (uint32_t)((it & 1) ^ 1));         // (hardware sleep, not a spin: the wait is long)
        TR(4, it);
        const uint32_t gb = sG32 + e0;
#pragma unroll
        for (int i = 0; i < F4; ++i) {
          uint32_t a1, a2, b1, b2;
          split2_bf(R[i].x, R[i].y, a1, a2);
          split2_bf(R[i].z, R[i].w, b1, b2);
          sts64((gb ^ ((i & 1) ? 64u : 0u)) + (uint32_t)(i >> 1) * kBlk, a1, b1);
          sts64((gb ^ ((i & 1) ? 64u : 0u)) + (uint32_t)(i >> 1) * kBlk + pl.xterm, a2, b2);
        }
        TR(1, it);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_GFULL]);
        TR(2, it);
      };
      // Tile by tile: ALL of this warp's units of the next tile (warps 0..3: units w, w+6, w+12; warps 4, 5: w, w+6) are
      // requested right after the current tile's units were converted, so they sit in registers long before the gR tile
      // is released; after the release only conversions happen (measured with two sets and a rolling order: a load
      // issued between two conversions stalls for 1-2 k cycles whenever the E3 warps' end-of-tile burst of TMA traffic
      // has filled the path to L2, and the conversion of a unit that WAS already loaded waited behind it — 6 k cycles
      // from release to completion of the tile).
      float4 R0[F4], R1[F4], R2[F4];
      const bool three = pw + 2 * (kProd - 1) < kTok / 4;
      load(R0, Pos{0, pw});
      load(R1, Pos{0, pw + (kProd - 1)});
      if (three) load(R2, Pos{0, pw + 2 * (kProd - 1)});
#pragma unroll 1
      for (int it = 0; it < nmine; ++it) {
        convert(R0, Pos{it, pw});
        convert(R1, Pos{it, pw + (kProd - 1)});
        if (three) convert(R2, Pos{it, pw + 2 * (kProd - 1)});
        load(R0, Pos{it + 1, pw});
        load(R1, Pos{it + 1, pw + (kProd - 1)});
        if (three) load(R2, Pos{it + 1, pw + 2 * (kProd - 1)});
      }
    }
  } else if (is_e1) {
    // ======================================================================= E1
    const int et = tid - 64;                             // 0..63 = token row; its G1 row sits in TMEM lane 64 + et
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const float sc = p.g_loss_sq ? 2.0f * __ldg(p.g_loss_sq) : 0.f;
    const float invC = 1.0f / (float)C;
    const uint32_t sHc32 = smem_u32(sHc), sCg32 = smem_u32(sCg);
    float dv[32], av[32];
    float rs_next = 0.f, mu_next = 0.f;
    float4 st_next = make_float4(0, 0, 0, 0);
    // The tile's D and A rows arrive by TMA (two [64 x 32] fp32 boxes, SWIZZLE_128B) one tile ahead and are read from
    // shared memory, thread = row, conflict free.  (Before: 16 global 128-bit loads per thread with a 128-byte lane stride,
    // software-pipelined through 64 registers; the event trace showed E1 — 6 k cycles of work per tile — taking 10.6 k per
    // tile with those loads and setting the pace of the whole kernel once the gR path had been fixed.)
    const uint32_t sD32 = smem_u32(smem + pl.da_off), sA32 = sD32 + kBlk;
    auto request_da = [&](int it) {                      // one thread: the boxes of tile `it`
      const long long r0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * kTok;
      mbar_expect_tx(&bars[B_DAFULL], 2u * kBlk);
      tma_load_2d_u32(&mapD, sD32, &bars[B_DAFULL], 0, (int)r0);
      tma_load_2d_u32(&mapA, sA32, &bars[B_DAFULL], 0, (int)r0);
    };
    auto load_row_stats = [&](int it) {                  // rstd, mu, rowstats of tile `it` (three small loads per thread)
      const long long row = ((long long)blockIdx.x + (long long)it * gridDim.x) * kTok + et;
      const bool live = it < nmine && row < p.N;
      rs_next = live ? __ldg(p.rstd + row) : 0.f;
      mu_next = live ? __ldg(p.mu + row) : 0.f;
      st_next = live ? ldg_nc(reinterpret_cast<const float4*>(p.rowstats) + row) : make_float4(0, 0, 0, 0);
    };
    auto prefetch_da = [&](int it) {                     // L2 prefetch of a later tile's row statistics (12 lines, 12 lanes)
      if (et < 12 && it < nmine) {
        const long long rn = ((long long)blockIdx.x + (long long)it * gridDim.x) * kTok;
        const char* a = et < 8 ? reinterpret_cast<const char*>(p.rowstats + rn * 4) + et * 128
                               : reinterpret_cast<const char*>((et < 10 ? p.mu : p.rstd) + rn) + (et & 1) * 128;
        const long long lim = et < 8 ? (p.N - rn) * 16 - et * 128 : (p.N - rn) * 4 - (et & 1) * 128;
        if (lim > 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
      }
    };
    if (et == 0) request_da(0);
    prefetch_da(1);
    load_row_stats(0);
    for (int it = 0; it < nmine; ++it) {
      const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
      const long long row = tile * kTok + et;
      const bool live = row < p.N;
      const float rs = rs_next, nmr = -mu_next * rs_next;
      const float zz = st_next.x, p1 = st_next.y, p2 = st_next.z;
      TR(10, it);
      prefetch_da(it + 2);
      const int buf = it & 1;
      // D / A rows of this tile: shared memory -> registers; then the boxes go back to the copy engine for the next tile
      mbar_wait(&bars[B_DAFULL], (uint32_t)(it & 1));
      {
        const uint32_t ro = (uint32_t)et * 128u, rx = (uint32_t)(et & 7);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 d4 = lds128f(sD32 + ro + ((((uint32_t)q) ^ rx) << 4));
          const float4 a4 = lds128f(sA32 + ro + ((((uint32_t)q) ^ rx) << 4));
          dv[4 * q] = d4.x; dv[4 * q + 1] = d4.y; dv[4 * q + 2] = d4.z; dv[4 * q + 3] = d4.w;
          av[4 * q] = a4.x; av[4 * q + 1] = a4.y; av[4 * q + 2] = a4.z; av[4 * q + 3] = a4.w;
        }
        fence_async_smem();                              // generic reads before the async-proxy refill (see the note in the producers)
        named_bar(2, 64);
        if (et == 0 && it + 1 < nmine) request_da(it + 1);
      }
      mbar_wait(&bars[B_G1FULL0 + buf], (uint32_t)((it >> 1) & 1));
      TR(14, it);
      tc_fence_after();
      float gv[32];
      tmem_ld32(tmem + lane_addr + kColG1 + (uint32_t)(buf * 64), gv);           // gR . cen_hi (+ gR_lo . cen_hi)
#pragma unroll
      for (int h = 0; h < 2; ++h) {                                               // + gR_hi . cen_lo
        float g2[16];
        tmem_ld16(tmem + lane_addr + kColG1 + (uint32_t)(buf * 64 + 32 + 16 * h), g2);
#pragma unroll
        for (int k = 0; k < 16; ++k) gv[16 * h + k] += g2[k];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_G1EMPTY0 + buf]);
      TR(40, it);
      // softmin backward + cdist ratio (packed fp32: element pairs k, k+1)
      const float2 sc2 = bcast2(sc);
      float2 dot2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 32; k += 2) {
        const float2 d2 = make_float2(dv[k], dv[k + 1]), a2 = make_float2(av[k], av[k + 1]);
        const float2 g2 = fma2(mul2(sc2, d2), mul2(d2, a2), make_float2(gv[k], gv[k + 1]));   // gA_tot = G1 + sc D^2 A
        gv[k] = g2.x; gv[k + 1] = g2.y;
        dot2 = fma2(g2, a2, dot2);
      }
      const float2 dotb = bcast2(dot2.x + dot2.y), nalpha2 = bcast2(-p.alpha), hz2 = bcast2(0.5f * zz);
      float2 rsum2 = make_float2(0.f, 0.f), sT2 = rsum2, sG2 = rsum2;
#pragma unroll
      for (int k = 0; k < 32; k += 2) {
        const float2 d2 = make_float2(dv[k], dv[k + 1]), a2 = make_float2(av[k], av[k + 1]);
        const float2 t = sub2(make_float2(gv[k], gv[k + 1]), dotb);
        const float2 gd = mul2(a2, fma2(mul2(sc2, d2), a2, mul2(nalpha2, t)));
        float2 r = mul2(gd, make_float2(fast_rcp(d2.x), fast_rcp(d2.y)));   // ATen cdist backward: grad / dist, 0 where dist == 0
        r.x = (d2.x == 0.f || !live) ? 0.f : r.x;
        r.y = (d2.y == 0.f || !live) ? 0.f : r.y;
        gv[k] = r.x; gv[k + 1] = r.y;
        const float2 hc = make_float2(lds32f_const(sHc32 + 4 * k), lds32f_const(sHc32 + 4 * k + 4));
        const float2 cg = make_float2(lds32f_const(sCg32 + 4 * k), lds32f_const(sCg32 + 4 * k + 4));
        const float2 T = fma2(mul2(d2, bcast2(-0.5f)), d2, add2(hz2, hc));
        rsum2 = add2(rsum2, r); sT2 = fma2(r, T, sT2); sG2 = fma2(r, cg, sG2);
      }
      const float rsum = rsum2.x + rsum2.y, sT = sT2.x + sT2.y, sG = sG2.x + sG2.y;
      const float s1 = (rsum * p1 - sG) * invC, s2 = (rsum * p2 - sT) * invC;
      TR(41, it);
      // tile R[buf] free: S3 / S5b of tile it-2 have completed
      mbar_wait(&bars[B_REMPTY0 + buf], (uint32_t)(((it >> 1) & 1) ^ 1));
      TR(42, it);
      const uint32_t tr32 = sTR32 + (uint32_t)buf * kBlk;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t w1[4], w2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2_bf(gv[8 * j + 2 * e], gv[8 * j + 2 * e + 1], w1[e], w2[e]);
        sts128(tr32 + sw128((uint32_t)et, (uint32_t)j * 16u), w1[0], w1[1], w1[2], w1[3]);
        sts128(tr32 + sw128((uint32_t)et, 64u + (uint32_t)j * 16u), w2[0], w2[1], w2[2], w2[3]);
      }
      float* scal = sScal + buf * 5 * kTok;
      scal[et] = rsum; scal[kTok + et] = s1 * rs; scal[2 * kTok + et] = s2 * rs; scal[3 * kTok + et] = rs;
      scal[4 * kTok + et] = nmr;
      TR(43, it);
      fence_async_smem();
      TR(44, it);
      named_bar(1, 64);
      if (et == 0) mbar_arrive(&bars[B_RFULL0 + buf]);
      TR(15, it);
      // next tile's row statistics: issued after the fence / barrier above (which would wait for them)
      load_row_stats(it + 1);
      TR(16, it);                                        // (colsum(r) is taken from the r operand tile by producer warp 6)
    }
  } else if (is_e3) {
    // ======================================================================= E3
    const int e3 = (warp >> 2) - 1;                      // 64-channel group
    const int q = warp & 1;                              // token half
    float accq[2][8];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) accq[c][j] = 0.f;
    if (e3 < NCB) {
      const int wi = e3 * 2 + q;
      const int rl = q * 32 + lane;                      // token row in the tile = TMEM lane
      const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
      const uint64_t pol = policy_evict_first();
      const uint32_t box32 = smem_u32(smem + pl.xr_off) + (uint32_t)wi * 2u * kBox;   // this warp's two boxes (channel halves)
      uint64_t* xrbar = &bars[B_XR0 + wi * 2];
      const uint32_t rx = (uint32_t)(lane & 7);          // == rl & 7
      const uint32_t xhrow = sXH32 + (uint32_t)e3 * kBlk + (uint32_t)rl * 128u;
      const uint32_t sGam32 = smem_u32(sGam), sBet32 = smem_u32(sBet);
      auto load_box = [&](int it, int ch) {              // lane 0: raw x of tile `it`, rows q 32.., channels e3 64 + ch 32..
        const long long r0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * kTok + q * 32;
        mbar_expect_tx(&xrbar[ch], kBox);
        tma_load_2d_u32(&mapX, box32 + (uint32_t)ch * kBox, &xrbar[ch], e3 * 64 + ch * 32, (int)r0);
      };
      if (lane == 0) { load_box(0, 0); load_box(0, 1); }
      for (int it = 0; it < nmine; ++it) {
        const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
        const long long row0 = tile * kTok;
        const int par = it & 1;
        // second box of THIS tile: its bytes held gx of the previous tile until that store had read them (the wait
        // overlaps the wait for this tile's accumulator); the first box was requested before the previous tile ended
        if (it > 0 && lane == 0) { bulk_wait_read0(); load_box(it, 1); }
        TR(19, it);
        mbar_wait(&bars[B_RFULL0 + par], (uint32_t)((it >> 1) & 1));
        TR(20, it);
        const float* scal = sScal + par * 5 * kTok;
        const float rsum = scal[rl], s1r = scal[kTok + rl], s2r = scal[2 * kTok + rl], rs = scal[3 * kTok + rl];
        const float nmr = scal[4 * kTok + rl];
        mbar_wait(&bars[B_ACCFULL], (uint32_t)par);
        TR(21, it);
        tc_fence_after();
        // (tried: both 32-column halves of the accumulator pulled into registers at once and the accumulator handed back
        //  immediately, so that S3 of the next tile runs under this tile's epilogue: the S3 gap disappears, but the 64 extra
        //  live registers cost 204 B of spills and this role's per-tile time went from ~8 k to ~10 k cycles: 341 -> 355 us)
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const int c0 = e3 * 64 + ch * 32;
          float acc[32], xh[32];
          tmem_ld32(tmem + lane_addr + kColAcc + (uint32_t)c0, acc);
          mbar_wait(&xrbar[ch], (uint32_t)par);          // the raw x box of this tile has landed
          const uint32_t bx = box32 + (uint32_t)ch * kBox + (uint32_t)lane * 128u;
          {
            const float2 rs2 = bcast2(rs), nmr2 = bcast2(nmr);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = lds128f(bx + ((((uint32_t)j) ^ rx) << 4));
              const float2 a = fma2(make_float2(t.x, t.y), rs2, nmr2), b = fma2(make_float2(t.z, t.w), rs2, nmr2);
              xh[4 * j] = a.x; xh[4 * j + 1] = a.y; xh[4 * j + 2] = b.x; xh[4 * j + 3] = b.y;   // xhat, exact in fp32
            }
          }
          // packed fp32 (element pairs): gz = z rsum - acc;  o = (gz gamma) rs - s1 rs - xhat s2 rs
          const float2 nrsum2 = bcast2(-rsum), rsum2 = bcast2(rsum), nrs2 = bcast2(-rs), ns1r2 = bcast2(-s1r), ns2r2 = bcast2(-s2r);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 gm = lds128f_const(sGam32 + (uint32_t)(c0 + 4 * j) * 4u);
            const float4 be = lds128f_const(sBet32 + (uint32_t)(c0 + 4 * j) * 4u);
            float2 o[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float2 g2 = e ? make_float2(gm.z, gm.w) : make_float2(gm.x, gm.y);
              const float2 b2 = e ? make_float2(be.z, be.w) : make_float2(be.x, be.y);
              const float2 x2 = make_float2(xh[4 * j + 2 * e], xh[4 * j + 2 * e + 1]);
              const float2 a2 = make_float2(acc[4 * j + 2 * e], acc[4 * j + 2 * e + 1]);
              const float2 z = fma2(x2, g2, b2);
              const float2 ngz = fma2(z, nrsum2, a2);                      // -(gz) = acc - z rsum
              o[e] = fma2(x2, ns2r2, fma2(mul2(ngz, g2), nrs2, ns1r2));
            }
            // gx goes back into the bytes x was read from (this thread's own row of the box)
            sts128f(bx + ((((uint32_t)j) ^ rx) << 4), o[0].x, o[0].y, o[1].x, o[1].y);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d_hint(&mapGx, box32 + (uint32_t)ch * kBox, c0, (int)(row0 + q * 32), pol);
            bulk_commit();
            // the first box's store was issued half a tile ago: it has read its bytes, the next tile's x may land there
            if (ch == 1 && it + 1 < nmine) { bulk_wait_read1(); load_box(it + 1, 0); }
          }
          if (ch == 1) {                                 // the accumulator is consumed: S3 of the next tile may start
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_ACCEMPTY]);
            TR(23, it);
          }
          // the xhat operand tile is free once S5b of the previous tile has completed — asked for as late as possible: S5b
          // of tile t-1 is issued after E3 of tile t-1 has ended and shares its issuing thread with S3 of this tile
          if (ch == 0) mbar_wait(&bars[B_XHEMPTY], (uint32_t)(par ^ 1));
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {               // 8 channels per step
            // two-term bf16 split of the 8 xhat values -> one 16-byte chunk of the hi block and of the lo block
            uint32_t w1[4], w2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split2_bf(xh[8 * jj + 2 * e], xh[8 * jj + 2 * e + 1], w1[e], w2[e]);
            const uint32_t xo = xhrow + ((((uint32_t)(ch * 4 + jj)) ^ rx) << 4);
            sts128(xo, w1[0], w1[1], w1[2], w1[3]);
            sts128(xo + pl.xterm, w2[0], w2[1], w2[2], w2[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {                // Q contributions replace xhat
              const float2 x2 = make_float2(xh[8 * jj + 2 * e], xh[8 * jj + 2 * e + 1]);
              const float2 qq = mul2(mul2(x2, rsum2), x2);
              xh[8 * jj + 2 * e] = qq.x; xh[8 * jj + 2 * e + 1] = qq.y;
            }
          }
          if (ch == 1) {                                 // xhat operand rows written
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_XHFULL]);
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
        p.part_q[((size_t)blockIdx.x * 2 + q) * C + e3 * 64 + ch * 32 + lane] = accq[ch][0];
      }
    }
  } else if (lane == 0) {
    // ======================================================================= MMA (warp 15, one thread)
    mbar_expect_tx(&bars[B_CEN], NCB * kBlk);
    for (uint32_t off = 0; off < NCB * kBlk; off += 8192u)
      bulk_g2s(smem + pl.cen_off + off, p.cimage + off, 8192u, &bars[B_CEN]);
    const uint32_t idesc1a = instr_desc(kFmtBF16, 128, 2 * K, 0, 0);  // S1: tokens x [cen_hi; cen_lo], both K-major
    const uint32_t idesc1b = instr_desc(kFmtBF16, 128, K, 0, 0);      //     gR_lo x cen_hi
    const uint32_t idesc3 = instr_desc(kFmtBF16, 128, C, 0, 1);       // S3: tokens x C, cen MN-major
    const uint32_t idesc5 = instr_desc(kFmtBF16, 128, C, 1, 1);       // S5: slots x C, both MN-major, K = tokens
    // descriptor bases: hi word is constant (SBO 1024, version 1, SWIZZLE_128B); lo = (address >> 4) | LBO field
    constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t kLbo = (kBlk >> 4) << 16;                      // 8 KB between the 64-wide blocks of an MN-major operand
    const uint32_t loG_k = (sG32 >> 4) & 0x3FFFu, loG_mn = loG_k | kLbo;
    const uint32_t loG_k64 = loG_k - (kBlk >> 4);                    // S1: MMA rows 64..127 = the tile's rows 0..63
    const uint32_t loX_mn = ((sXH32 >> 4) & 0x3FFFu) | kLbo;
    const uint32_t loTA_mn = ((sTA32 >> 4) & 0x3FFFu) | kLbo;         // blocks: tile A, ZERO
    const uint32_t loZR_mn0 = ((sZero32 >> 4) & 0x3FFFu) | kLbo;      // blocks: ZERO, tile R[0]
    const uint32_t loZR_mn1 = ((sZero32 >> 4) & 0x3FFFu) | (2u * kLbo);   //      ZERO, tile R[1] (16 KB further)
    const uint32_t loTR_k = (sTR32 >> 4) & 0x3FFFu;
    const uint32_t loC_k = (sCen32 >> 4) & 0x3FFFu, loC_mn = loC_k | kLbo;
    mbar_wait(&bars[B_CEN], 0);
    int n1 = 0, n5a = 0, n3 = 0, n5b = 0;
    // One thread issues every MMA of the CTA.  What the event trace (VADC_BWD_TRACE builds) taught about its loop:
    //  * mbarrier.try_wait is a hardware sleep with a time limit: a failed probe of one barrier delays the look at the next
    //    by ~2 k cycles — every probe here is the non-blocking test_wait;
    //  * test_wait probes of one thread do not overlap (~150 cycles EACH, also when issued back to back from one asm
    //    block): six probes per look made every reaction ~0.9 k cycles late, and a look between chunks of a contraction
    //    more than doubled its issue time.  So a condition that has been seen true is remembered until its counter moves,
    //    and a look probes only what is still unknown and currently decides something (typically one or two barriers);
    //  * E3 sets the pace and waits for S3, and, a few hundred cycles into its next tile, for the xhat tile that S5b frees:
    //    priority S3 > S5b > S1 > S5a.
    enum { kRF = 1, kAE = 2, kGF = 4, kG1E = 8, kAF = 16, kXF = 32 };
    uint32_t seen = 0;
    auto probe = [&](uint32_t bit, int b, uint32_t parity) {
      if (!(seen & bit) && mbar_test(&bars[b], parity)) {
        seen |= bit;
        if constexpr (TRACE) TR(64 + (int)seen, n3);
      }
    };
    while (n5b < nmine) {
      // ---- S3 (feeds E3, urgent): acc = r_lo cen_hi + r_hi cen_lo + r_hi cen_hi
      if (n3 < n1) {
        probe(kRF, B_RFULL0 + (n3 & 1), (uint32_t)((n3 >> 1) & 1));
        if (seen & kRF) probe(kAE, B_ACCEMPTY, (uint32_t)((n3 & 1) ^ 1));
        if ((seen & (kRF | kAE)) == (kRF | kAE)) {
          tc_fence_after();
          const uint32_t d = tmem + kColAcc;
          const uint32_t loTR = loTR_k + (uint32_t)(n3 & 1) * (kBlk >> 4);
          constexpr uint32_t ro[3] = {64u, 0u, 0u}, co[3] = {0u, 4096u, 0u};
#pragma unroll
          for (int t = 0; t < 3; ++t) {
#pragma unroll
            for (int ks = 0; ks < K / 16; ++ks) {
              const uint64_t ad = desc_at(loTR, kHi, ro[t] + (uint32_t)ks * 32u);
              const uint64_t bd = desc_at(loC_mn, kHi, co[t] + (uint32_t)(2 * ks) * 1024u);
              mma_f16(d, ad, bd, idesc3, (t > 0 || ks > 0) ? 1u : 0u);
            }
          }
          mma_commit(&bars[B_ACCFULL]);
          TR(32, n3);
          ++n3;
          seen &= ~(uint32_t)(kRF | kAE);
          continue;
        }
      }
      // ---- S5b (after E3 of its tile; frees the xhat tile and r[parity]): PT[64:128] += [r_hi | r_lo]^T (xhat_hi + xhat_lo)
      if (n5b < n3 && n5b < n5a) {
        probe(kXF, B_XHFULL, (uint32_t)(n5b & 1));
        if (seen & kXF) {
          tc_fence_after();
          const uint32_t loZR_mn = (n5b & 1) ? loZR_mn1 : loZR_mn0;
VADC_S5_UNROLL
          for (int t = 0; t < 2; ++t) {
VADC_S5_UNROLL
            for (int ks = 0; ks < kTok / 16; ++ks) {
              const uint64_t ad = desc_at(loZR_mn, kHi, (uint32_t)ks * 2048u);
              const uint64_t bd = desc_at(loX_mn, kHi, (uint32_t)t * pl.xterm + (uint32_t)ks * 2048u);
              mma_f16(tmem + kColPT, ad, bd, idesc5, 1u);
            }
          }
          mma_commit(&bars[B_XHEMPTY]);
          mma_commit(&bars[B_REMPTY0 + (n5b & 1)]);
          TR(33, n5b);
          ++n5b;
          seen &= ~(uint32_t)kXF;
          continue;
        }
      }
      // ---- S1 (feeds E1): G1[:, 0:32] = gR_hi cen_hi + gR_lo cen_hi, G1[:, 32:64] = gR_hi cen_lo
      if (n1 < nmine) {
        probe(kGF, B_GFULL, (uint32_t)(n1 & 1));
        if (seen & kGF) probe(kG1E, B_G1EMPTY0 + (n1 & 1), (uint32_t)(((n1 >> 1) & 1) ^ 1));
        if ((seen & (kGF | kG1E)) == (kGF | kG1E)) {
          tc_fence_after();
          const uint32_t d = tmem + kColG1 + (uint32_t)((n1 & 1) * 64);
#pragma unroll
          for (int kk = 0; kk < C / 16; ++kk) {
            const uint64_t ad = desc_at(loG_k64, kHi, (kk >> 2) * kBlk + (kk & 3) * 32u);
            const uint64_t bd = desc_at(loC_k, kHi, (kk >> 2) * kBlk + (kk & 3) * 32u);
            mma_f16(d, ad, bd, idesc1a, kk > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int kk = 0; kk < C / 16; ++kk) {
            const uint64_t ad = desc_at(loG_k64, kHi, pl.xterm + (kk >> 2) * kBlk + (kk & 3) * 32u);
            const uint64_t bd = desc_at(loC_k, kHi, (kk >> 2) * kBlk + (kk & 3) * 32u);
            mma_f16(d, ad, bd, idesc1b, 1u);
          }
          mma_commit(&bars[B_G1FULL0 + (n1 & 1)]);
          TR(30, n1);
          ++n1;
          seen &= ~(uint32_t)(kGF | kG1E);
          continue;
        }
      }
      // ---- S5a (background; its completion frees the gR tile for the producers): PT[0:64] += [A_hi | A_lo]^T (gR_hi + gR_lo)
      if (n5a < n1) {
        probe(kAF, B_AFULL, (uint32_t)(n5a & 1));
        if (seen & kAF) {
          tc_fence_after();
VADC_S5_UNROLL
          for (int t = 0; t < 2; ++t) {
VADC_S5_UNROLL
            for (int ks = 0; ks < kTok / 16; ++ks) {
              const uint64_t ad = desc_at(loTA_mn, kHi, (uint32_t)ks * 2048u);
              const uint64_t bd = desc_at(loG_mn, kHi, (uint32_t)t * pl.xterm + (uint32_t)ks * 2048u);
              mma_f16(tmem + kColPT, ad, bd, idesc5, (n5a > 0 || t > 0 || ks > 0) ? 1u : 0u);
            }
          }
          mma_commit(&bars[B_GEMPTY]);
          mma_commit(&bars[B_AEMPTY]);
          TR(31, n5a);
          ++n5a;
          seen &= ~(uint32_t)kAF;
          continue;
        }
      }
      if (p.mma_sleep > 0) __nanosleep((unsigned)p.mma_sleep);   // (a sleep costs far more than its nominal length: default spin)
    }
    mma_commit(&bars[B_DONE]);
    mbar_wait(&bars[B_DONE], 0);
  }
  // ---- drain the whole-kernel accumulator: lane = slot (A_hi, A_lo, r_hi, r_lo x 32 centroids), columns = channels
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 4) {
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float4* o = reinterpret_cast<float4*>(p.part_p + ((size_t)blockIdx.x * 128 + warp * 32 + lane) * C);
#pragma unroll 1
    for (int ch = 0; ch < C / 32; ++ch) {
      float v[32];
      tmem_ld32(tmem + lane_addr + kColPT + (uint32_t)(ch * 32), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[ch * 8 + j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) { tc_fence_after(); tmem_dealloc(tmem, ncols); }
}

// finalize 1: block = (centroid k, 32-channel chunk), 256 threads = 32 channels x 8 groups of CTAs; every group adds
// its CTAs' partials in fixed order (four independent accumulation chains: the 14.5 MB of partials are L2-resident and
// the kernel is latency-bound) and the groups are added in fixed order (deterministic):
//   P1 = sum_b (PT[b][k] + PT[b][32+k]),  P2 = sum_b (PT[b][64+k] + PT[b][96+k]),  rcol_k = sum_b rcol[b]
//   gcenters[k,c] = P1 - gamma_c P2 + (cen[k,c] - beta_c) rcol_k;   P2 and rcol_k are kept for finalize 2
__global__ void __launch_bounds__(256)
cluster_bwd_tc_finalize1_kernel(const float* __restrict__ part_p, const float* __restrict__ part_rcol,
                                const float* __restrict__ centers, const float* __restrict__ ln_w,
                                const float* __restrict__ ln_b, int nb, int K, int C,
                                float* __restrict__ gcenters, float* __restrict__ p2buf, float* __restrict__ rcol) {
  const int k = blockIdx.x, cl = threadIdx.x & 31, c = blockIdx.y * 32 + cl, grp = threadIdx.x >> 5;
  __shared__ float s1s[8][32], s2s[8][32], rcs[256];
  float rc = 0.f;
  for (int b = threadIdx.x; b < 2 * nb; b += 256) rc += part_rcol[(size_t)b * K + k];
  float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    const size_t cta = (size_t)128 * C;
    const float* base = part_p + c;
    int b = grp;
    for (; b + 24 < nb; b += 32) {                       // four CTAs of this group per pass: 16 independent loads
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* pp = base + (size_t)(b + 8 * u) * cta;
        a1[u] += pp[(size_t)k * C] + pp[(size_t)(K + k) * C];
        a2[u] += pp[(size_t)(2 * K + k) * C] + pp[(size_t)(3 * K + k) * C];
      }
    }
    for (; b < nb; b += 8) {
      const float* pp = base + (size_t)b * cta;
      a1[0] += pp[(size_t)k * C] + pp[(size_t)(K + k) * C];
      a2[0] += pp[(size_t)(2 * K + k) * C] + pp[(size_t)(3 * K + k) * C];
    }
  }
  s1s[grp][cl] = (a1[0] + a1[1]) + (a1[2] + a1[3]);
  s2s[grp][cl] = (a2[0] + a2[1]) + (a2[2] + a2[3]);
  rcs[threadIdx.x] = rc;
  __syncthreads();
  if (threadIdx.x < 32) {
    rc = 0.f;
    for (int i = 0; i < 256; ++i) rc += rcs[i];          // broadcast reads, fixed order
    float s1 = 0.f, s2 = 0.f;
    for (int g = 0; g < 8; ++g) { s1 += s1s[g][threadIdx.x]; s2 += s2s[g][threadIdx.x]; }
    if (blockIdx.y == 0 && threadIdx.x == 0) rcol[k] = rc;
    if (c < C) {
      gcenters[(size_t)k * C + c] = s1 - ln_w[c] * s2 + (centers[(size_t)k * C + c] - ln_b[c]) * rc;
      p2buf[(size_t)k * C + c] = s2;
    }
  }
}

// finalize 2: one warp per channel, lane = centroid (K == 32):
//   g_beta[c]  = gamma_c sum_k P2[k,c] + beta_c sum_k rcol_k - sum_k rcol_k cen[k,c]
//   g_gamma[c] = gamma_c Q[c] + beta_c sum_k P2[k,c] - sum_k cen[k,c] P2[k,c]
__global__ void __launch_bounds__(256)
cluster_bwd_tc_finalize2_kernel(const float* __restrict__ p2buf, const float* __restrict__ rcol,
                                const float* __restrict__ part_q, const float* __restrict__ centers,
                                const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                int nb, int K, int C, float* __restrict__ gw, float* __restrict__ gb) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), k = threadIdx.x & 31;
  if (c >= C) return;
  const float s2 = p2buf[(size_t)k * C + c], cen = centers[(size_t)k * C + c], rc = rcol[k];
  float qv = 0.f;
  for (int b = k; b < 2 * nb; b += 32) qv += part_q[(size_t)b * C + c];
  const float sp2 = warp_sum(s2), scp = warp_sum(cen * s2), src = warp_sum(rc), scr = warp_sum(rc * cen);
  qv = warp_sum(qv);
  if (k == 0) {
    const float g = ln_w[c], be = ln_b[c];
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

// [rows, cols] fp32 row-major tensor, box = 32 columns x box_rows rows, SWIZZLE_128B
static int make_map(CUtensorMap* m, float* base, long long rows, int cols, int box_rows = 32) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return VADC_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
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

}  // namespace bt2

bool bwd_tc2_shape_ok(long long N, int C, int K) { return bt2::shape_ok(N, C, K); }

size_t bwd_tc2_workspace_bytes(long long N, int C, int K) {
  if (!bt2::shape_ok(N, C, K)) return 0;
  const size_t g = (size_t)sm_count();
  return align_up((size_t)2 * K * C * 2, 256) + align_up((size_t)3 * K * sizeof(float), 256) +
         align_up(g * 128 * C * sizeof(float), 256) + align_up(g * 2 * K * sizeof(float), 256) +
         align_up(g * 2 * C * sizeof(float), 256) + align_up((size_t)K * C * sizeof(float), 256) +
         align_up((size_t)K * sizeof(float), 256) + 256;
}

int launch_cluster_bwd_tc2(const float* x, const float* mu, const float* rstd, const float* rowstats, const float* ln_w,
                           const float* ln_b, const float* centers, const float* D, const float* A,
                           const float* gR, const float* g_loss_sq, long long N, int C, int K, float alpha,
                           float* gx, float* gcenters, float* g_ln_w, float* g_ln_b, void* workspace,
                           size_t workspace_bytes, cudaStream_t st) {
  if (!bt2::shape_ok(N, C, K)) return VADC_ERR_UNSUPPORTED;
  if (!vadc_device_ok()) return VADC_ERR_NO_DEVICE;
  if (workspace_bytes < bwd_tc2_workspace_bytes(N, C, K)) return VADC_ERR_WORKSPACE;
  Carver ws(workspace, workspace_bytes);
  const size_t g = (size_t)sm_count();
  uint8_t* image = ws.take<uint8_t>((size_t)2 * K * C * 2);
  float* cvec = ws.take<float>(3 * K);
  float* part_p = ws.take<float>(g * 128 * C);
  float* part_rcol = ws.take<float>(g * 2 * K);
  float* part_q = ws.take<float>(g * 2 * C);
  float* p2buf = ws.take<float>((size_t)K * C);
  float* rcol = ws.take<float>(K);
  const int grid = (int)std::min<long long>((N + bt2::kTok - 1) / bt2::kTok, (long long)g);

  CUtensorMap mGx, mX, mD, mA;
  int rc;
  if ((rc = bt2::make_map(&mGx, gx, N, C))) return rc;
  if ((rc = bt2::make_map(&mX, const_cast<float*>(x), N, C))) return rc;
  if ((rc = bt2::make_map(&mD, const_cast<float*>(D), N, K, bt2::kTok))) return rc;
  if ((rc = bt2::make_map(&mA, const_cast<float*>(A), N, K, bt2::kTok))) return rc;
  bt2::centroid_prep_bwd_kernel<<<K, 256, 0, st>>>(centers, ln_w, ln_b, K, C, image, cvec);
  VADC_CHECK_LAUNCH("centroid_prep_bwd_kernel");

  const size_t smem = bt2::plan(C).total + 1024;
  unsigned long long* trace = nullptr;
#ifdef VADC_BWD_TRACE_BUILD
  // debugging builds only (-DVADC_BWD_TRACE_BUILD): allocates, synchronises and writes a text file — never in the
  // product library, whose entry points must stay capturable in a CUDA graph
  const char* trace_path = getenv("VADC_BWD_TRACE");
  const size_t trace_bytes = (size_t)16 * 1024 * 2 * sizeof(unsigned long long);
  if (trace_path) { VADC_CUDA(cudaMalloc(&trace, trace_bytes)); VADC_CUDA(cudaMemsetAsync(trace, 0, trace_bytes, st)); }
  const int trace_cta = getenv("VADC_BWD_TRACE_CTA") ? atoi(getenv("VADC_BWD_TRACE_CTA")) : 0;
#else
  const int trace_cta = 0;
#endif
  bt2::Params p{x, gR, D, A, mu, rstd, rowstats, ln_w, ln_b, image, cvec, g_loss_sq, part_p, part_rcol, part_q,
                N, alpha, env_int("VADC_BWD_PF", 1), env_int("VADC_BWD_SLEEP", 0), trace, trace_cta};
  bool launched = false;
#ifdef VADC_BWD_TRACE_BUILD
#define BT_KERN(F4_) (trace ? bt2::cluster_bwd_tc2_kernel<F4_, true> : bt2::cluster_bwd_tc2_kernel<F4_, false>)
#else
#define BT_KERN(F4_) bt2::cluster_bwd_tc2_kernel<F4_, false>
#endif
#define BT_CASE(F4_)                                                                                   \
  if (C == 32 * F4_) {                                                                                 \
    auto kern = BT_KERN(F4_);                                                                          \
    VADC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    kern<<<grid, bt2::kThreads, smem, st>>>(mGx, mX, mD, mA, p);                                               \
    launched = true;                                                                                   \
  }
  timing_begin(VADC_TIMING_CLUSTER_BWD, st);
  BT_CASE(2) BT_CASE(4) BT_CASE(6)
#undef BT_CASE
#undef BT_KERN
  timing_end(VADC_TIMING_CLUSTER_BWD, st);
  if (!launched) return VADC_ERR_UNSUPPORTED;
  VADC_CHECK_LAUNCH("cluster_bwd_tc2_kernel");
#ifdef VADC_BWD_TRACE_BUILD
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
#endif
  bt2::cluster_bwd_tc_finalize1_kernel<<<dim3(K, (C + 31) / 32), 256, 0, st>>>(part_p, part_rcol, centers, ln_w, ln_b, grid, K, C,
                                                         gcenters, p2buf, rcol);
  VADC_CHECK_LAUNCH("cluster_bwd_tc_finalize1_kernel");
  bt2::cluster_bwd_tc_finalize2_kernel<<<(C + 7) / 8, 256, 0, st>>>(p2buf, rcol, part_q, centers, ln_w, ln_b, grid, K, C,
                                                                     g_ln_w, g_ln_b);
  VADC_CHECK_LAUNCH("cluster_bwd_tc_finalize2_kernel");
  return VADC_OK;
}

}  // namespace vadc
