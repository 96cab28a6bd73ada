// cluster_bwd_fused.cu — C2: fused backward of the cluster head for small K
// (K == 32, C in {64,128,192,256}).
//
// One persistent CTA (512 threads, 16 warps) walks 32-token panels; every input
// is read from HBM exactly once (x, feature, gR, D, A [+ gD, gA, gF]) and only
// gx is written per token:
//   S1  G1[t,k]  = sum_c gR[t,c] cen[k,c]                      (x_rec = A @ centers)
//   S2  softmin backward + cdist ratio -> r[t,k], rsum[t]      (rows.cuh::bwd_rows_kernel math)
//   S3  acc[t,c] = sum_k r[t,k] cen[k,c];  gz = z rsum - acc + gF
//   S4  LayerNorm backward -> gx[t,:]; gamma/beta partials in registers
//   S5  gcen[k,c] += A[t,k] gR[t,c] - r[t,k] z[t,c]            (register fragments, whole kernel)
// The three contractions run on the tensor cores as warp-level
// mma.sync.m16n8k8 tf32 with a 2-term split (x = hi + lo, hi = tf32-truncated;
// hi*hi + hi*lo + lo*hi, fp32 accumulate: ~2^-19 relative), because the fp32
// CUDA-core version of this kernel was shared-memory-bandwidth bound (ncu:
// l1tex 90 %, FMA pipe 22 %): MMA fragments are re-used from registers.
// (Pre-splitting the operand tiles into {hi,lo} pairs in shared memory was tried: it doubles the
// shared wavefronts per fragment and was 1.6x slower, so the split stays in registers.)
// Per-CTA partials (gcenters, colsum r, gamma, beta) go to the workspace and a
// small second kernel adds them in fixed order (deterministic).
#include "common.cuh"
#include "cluster.h"
#include <algorithm>

namespace vadc {

constexpr int kBT = 32;          // tokens per panel
constexpr int kBK = 32;          // centroids (this kernel is specialised for K == 32)
constexpr int kBThreads = 512;   // 16 warps; S2/S4 use 16 lanes per token

__device__ __forceinline__ float red16(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v;
}

// D(16x8) += A(16x8, row) * B(8x8, col), tf32 inputs (fp32 bit patterns), fp32 accumulate
__device__ __forceinline__ void mma_tf32_16n8k8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// split fp32 into tf32-exact hi (mantissa truncated to 10 bits) and the exact remainder lo
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// 3-pass product: d += (ah + al) * (bh + bl) without the al*bl term
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                     const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  mma_tf32_16n8k8(d, al, bh);
  mma_tf32_16n8k8(d, ah, bl);
  mma_tf32_16n8k8(d, ah, bh);
}

template <int F4>
struct BwdSmem {
  static constexpr int C = F4 * 32;
  static constexpr int LDT = C + 8;                      // row stride of the [32][C] tiles (bank spread for B frags)
  static constexpr int LDK = kBK + 4;                    // row stride of the [32][32] tiles
  static constexpr int kTile = kBT * LDT;
  static constexpr int kTK = kBT * LDK;
  // cen, gR, z, acc tiles + A, r, G1a, G1b
  static constexpr size_t bytes = sizeof(float) * (4 * kTile + 4 * kTK + 64);
};

struct BwdParams {
  const float* x; const float* mu; const float* rstd; const float* feature; const float* ln_w;
  const float* centers; const float* D; const float* A;
  const float* gD; const float* gA; const float* gR; const float* gF; const float* g_loss_sq;
  long long N; float alpha;
  float* gx; float* part_gc; float* part_rcol; float* part_ln;   // [grid][K*C], [grid][K], [grid][2C]
};

// Element-wise stages use tid = 16*t + q (t = token in panel, q = 0..15): thread (t,q) owns float4
// chunks q + 16j (j < H = C/64) of token t and centroids q, q+16.  MMA stages use warp = tid/32,
// g = lane/4, t4 = lane%4 (PTX m16n8k8 fragment layout).
template <int F4>
__global__ void __launch_bounds__(kBThreads, 1)
cluster_bwd_fused_kernel(const BwdParams p) {
  using S = BwdSmem<F4>;
  constexpr int C = S::C, LDT = S::LDT, LDK = S::LDK, H = F4 / 2;
  constexpr int NT = C / 8;                 // 8-wide column tiles of a [.,C] matrix
  constexpr int NTW = NT / 8;               // column tiles per warp (8 warps per m-tile)
  extern __shared__ __align__(16) float sm[];
  float* sCen = sm;                         // [32][LDT]
  float* sG = sCen + S::kTile;              // gR tile
  float* sZ = sG + S::kTile;                // feature tile
  float* sAcc = sZ + S::kTile;              // r @ cen
  float* sA = sAcc + S::kTile;              // A tile [32][36]
  float* sR = sA + S::kTK;                  // r tile
  float* sG1a = sR + S::kTK;                // G1 partial (first half of the channels)
  float* sG1b = sG1a + S::kTK;              // G1 partial (second half)
  const int tid = threadIdx.x, t = tid >> 4, q = tid & 15;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int wm = warp & 1;                  // m-tile (rows 16*wm .. +15) for every MMA stage
  const int wn = warp >> 1;                 // 0..7
  const float sc = p.g_loss_sq ? 2.0f * __ldg(p.g_loss_sq) : 0.f;

  for (int i = tid; i < kBK * (C / 4); i += kBThreads) {
    int k = i / (C / 4), c4 = i % (C / 4);
    *reinterpret_cast<float4*>(sCen + k * LDT + 4 * c4) = __ldg(reinterpret_cast<const float4*>(p.centers + (size_t)k * C) + c4);
  }
  float4 gam[H];
#pragma unroll
  for (int j = 0; j < H; ++j) gam[j] = __ldg(reinterpret_cast<const float4*>(p.ln_w) + q + 16 * j);

  float acc_gc[NTW][4];                     // S5 fragments: centroids 16*wm + {g, g+8}, channels 8*(wn*NTW+j) + {2t4, 2t4+1}
  float4 acc_gw[H], acc_gb[H];
  float acc_rcol[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < NTW; ++j) { acc_gc[j][0] = acc_gc[j][1] = acc_gc[j][2] = acc_gc[j][3] = 0.f; }
#pragma unroll
  for (int j = 0; j < H; ++j) { acc_gw[j] = make_float4(0, 0, 0, 0); acc_gb[j] = acc_gw[j]; }
  const long long npanels = (p.N + kBT - 1) / kBT;

  for (long long pn = blockIdx.x; pn < npanels; pn += gridDim.x) {
    const long long row = pn * kBT + t;
    const bool live = row < p.N;
    __syncthreads();                                     // previous panel's tiles are no longer read
    // ---- stage gR / z tiles; x stays in registers for S4
    float4 xv[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const int c4 = q + 16 * j;
      float4 gv = make_float4(0, 0, 0, 0), z = gv;
      xv[j] = gv;
      if (live) {
        if (p.gR) gv = ld_stream(reinterpret_cast<const float4*>(p.gR + row * C) + c4);
        z = ld_stream(reinterpret_cast<const float4*>(p.feature + row * C) + c4);
        xv[j] = ld_stream(reinterpret_cast<const float4*>(p.x + row * C) + c4);
      }
      *reinterpret_cast<float4*>(sG + t * LDT + 4 * c4) = gv;
      *reinterpret_cast<float4*>(sZ + t * LDT + 4 * c4) = z;
    }
    float dk[2], ak[2], gdk[2], gak[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = q + 16 * i;
      dk[i] = live ? __ldg(p.D + row * kBK + k) : 1.f;
      ak[i] = live ? __ldg(p.A + row * kBK + k) : 0.f;
      gdk[i] = (live && p.gD) ? __ldg(p.gD + row * kBK + k) : 0.f;
      gak[i] = (live && p.gA) ? __ldg(p.gA + row * kBK + k) : 0.f;
      sA[t * LDK + k] = ak[i];
    }
    __syncthreads();
    // ---- S1: G1[32 x 32] = gR[32 x C] . cen^T : warp -> (m-tile wm, n-tile wn&3, channel half wn>>2)
    if (p.gR) {
      const int n0 = (wn & 3) * 8, half = wn >> 2;
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      const float* ga = sG + (16 * wm + g) * LDT + t4;          // A frag: (token 16wm+g [+8], channel 8s+t4 [+4])
      const float* cb = sCen + (n0 + g) * LDT + t4;             // B frag: (channel 8s+t4 [+4], centroid n0+g)
#pragma unroll 4
      for (int s8 = half * (NT / 2); s8 < (half + 1) * (NT / 2); ++s8) {
        uint32_t ah[4], al[4], bh[2], bl[2];
        split_tf32(ga[8 * s8], ah[0], al[0]);
        split_tf32(ga[8 * LDT + 8 * s8], ah[1], al[1]);
        split_tf32(ga[8 * s8 + 4], ah[2], al[2]);
        split_tf32(ga[8 * LDT + 8 * s8 + 4], ah[3], al[3]);
        split_tf32(cb[8 * s8], bh[0], bl[0]);
        split_tf32(cb[8 * s8 + 4], bh[1], bl[1]);
        mma3(d, ah, al, bh, bl);
      }
      float* o = (half ? sG1b : sG1a) + (16 * wm + g) * LDK + n0 + 2 * t4;   // C frag: (token g [+8], centroid 2t4 [+1])
      o[0] = d[0]; o[1] = d[1]; o[8 * LDK] = d[2]; o[8 * LDK + 1] = d[3];
    }
    __syncthreads();
    // ---- S2: softmin backward + cdist ratio over the 16 lanes of token t
    float g1[2] = {0.f, 0.f};
    if (p.gR) {
      g1[0] = sG1a[t * LDK + q] + sG1b[t * LDK + q];
      g1[1] = sG1a[t * LDK + q + 16] + sG1b[t * LDK + q + 16];
    }
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      g1[i] += gak[i] + sc * dk[i] * dk[i] * ak[i];      // gA_tot
      dot = fmaf(g1[i], ak[i], dot);
    }
    dot = red16(dot);
    float rsum = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float gd = gdk[i] + sc * dk[i] * ak[i] * ak[i] - p.alpha * ak[i] * (g1[i] - dot);
      float r = (dk[i] == 0.f || !live) ? 0.f : gd / dk[i];
      sR[t * LDK + q + 16 * i] = r;
      acc_rcol[i] += r;
      rsum += r;
    }
    rsum = red16(rsum);
    __syncthreads();
    // ---- S3: acc[32 x C] = r[32 x 32] . cen[32 x C] ; S5: gcen += A^T gR - r^T z  (same warp tiling)
    {
      float d3[NTW][4];
#pragma unroll
      for (int j = 0; j < NTW; ++j) { d3[j][0] = d3[j][1] = d3[j][2] = d3[j][3] = 0.f; }
      const float* ra = sR + (16 * wm + g) * LDK + t4;          // S3 A frag: (token 16wm+g [+8], centroid 8s+t4 [+4])
#pragma unroll
      for (int s8 = 0; s8 < 4; ++s8) {
        uint32_t ah[4], al[4];
        split_tf32(ra[8 * s8], ah[0], al[0]);
        split_tf32(ra[8 * LDK + 8 * s8], ah[1], al[1]);
        split_tf32(ra[8 * s8 + 4], ah[2], al[2]);
        split_tf32(ra[8 * LDK + 8 * s8 + 4], ah[3], al[3]);
        // S5 A frags: element (m = centroid 16wm+g [+8], kk = token 8s+t4 [+4]) = A[token][centroid]
        uint32_t aah[4], aal[4], rrh[4], rrl[4];
        const float* aa = sA + (8 * s8 + t4) * LDK + 16 * wm + g;
        const float* rr = sR + (8 * s8 + t4) * LDK + 16 * wm + g;
        split_tf32(aa[0], aah[0], aal[0]);
        split_tf32(aa[8], aah[1], aal[1]);
        split_tf32(aa[4 * LDK], aah[2], aal[2]);
        split_tf32(aa[4 * LDK + 8], aah[3], aal[3]);
        split_tf32(-rr[0], rrh[0], rrl[0]);
        split_tf32(-rr[8], rrh[1], rrl[1]);
        split_tf32(-rr[4 * LDK], rrh[2], rrl[2]);
        split_tf32(-rr[4 * LDK + 8], rrh[3], rrl[3]);
#pragma unroll
        for (int j = 0; j < NTW; ++j) {
          const int c0 = 8 * (wn * NTW + j);
          uint32_t bh[2], bl[2];
          // S3 B frag: (centroid 8s+t4 [+4], channel c0+g)
          split_tf32(sCen[(8 * s8 + t4) * LDT + c0 + g], bh[0], bl[0]);
          split_tf32(sCen[(8 * s8 + t4 + 4) * LDT + c0 + g], bh[1], bl[1]);
          mma3(d3[j], ah, al, bh, bl);
          // S5 B frags: (token 8s+t4 [+4], channel c0+g) of gR and z
          split_tf32(sG[(8 * s8 + t4) * LDT + c0 + g], bh[0], bl[0]);
          split_tf32(sG[(8 * s8 + t4 + 4) * LDT + c0 + g], bh[1], bl[1]);
          mma3(acc_gc[j], aah, aal, bh, bl);
          split_tf32(sZ[(8 * s8 + t4) * LDT + c0 + g], bh[0], bl[0]);
          split_tf32(sZ[(8 * s8 + t4 + 4) * LDT + c0 + g], bh[1], bl[1]);
          mma3(acc_gc[j], rrh, rrl, bh, bl);
        }
      }
#pragma unroll
      for (int j = 0; j < NTW; ++j) {
        float* o = sAcc + (16 * wm + g) * LDT + 8 * (wn * NTW + j) + 2 * t4;
        *reinterpret_cast<float2*>(o) = make_float2(d3[j][0], d3[j][1]);
        *reinterpret_cast<float2*>(o + 8 * LDT) = make_float2(d3[j][2], d3[j][3]);
      }
    }
    __syncthreads();
    // ---- S4: LayerNorm backward of token t (16 lanes), gamma / beta partials
    {
      const float m = live ? __ldg(p.mu + row) : 0.f, rs = live ? __ldg(p.rstd + row) : 0.f;
      float s1 = 0.f, s2 = 0.f;
      float4 gg[H], xh[H];
#pragma unroll
      for (int j = 0; j < H; ++j) {
        const int c4 = q + 16 * j;
        const float4 z = *reinterpret_cast<const float4*>(sZ + t * LDT + 4 * c4);
        const float4 ac = *reinterpret_cast<const float4*>(sAcc + t * LDT + 4 * c4);
        float4 gv;
        gv.x = z.x * rsum - ac.x; gv.y = z.y * rsum - ac.y;
        gv.z = z.z * rsum - ac.z; gv.w = z.w * rsum - ac.w;
        if (p.gF && live) {
          const float4 f = ld_stream(reinterpret_cast<const float4*>(p.gF + row * C) + c4);
          gv.x += f.x; gv.y += f.y; gv.z += f.z; gv.w += f.w;
        }
        xh[j].x = (xv[j].x - m) * rs; xh[j].y = (xv[j].y - m) * rs;
        xh[j].z = (xv[j].z - m) * rs; xh[j].w = (xv[j].w - m) * rs;
        acc_gw[j].x = fmaf(gv.x, xh[j].x, acc_gw[j].x); acc_gw[j].y = fmaf(gv.y, xh[j].y, acc_gw[j].y);
        acc_gw[j].z = fmaf(gv.z, xh[j].z, acc_gw[j].z); acc_gw[j].w = fmaf(gv.w, xh[j].w, acc_gw[j].w);
        acc_gb[j].x += gv.x; acc_gb[j].y += gv.y; acc_gb[j].z += gv.z; acc_gb[j].w += gv.w;
        gg[j].x = gv.x * gam[j].x; gg[j].y = gv.y * gam[j].y; gg[j].z = gv.z * gam[j].z; gg[j].w = gv.w * gam[j].w;
        s1 += (gg[j].x + gg[j].y) + (gg[j].z + gg[j].w);
        s2 += (gg[j].x * xh[j].x + gg[j].y * xh[j].y) + (gg[j].z * xh[j].z + gg[j].w * xh[j].w);
      }
      s1 = red16(s1) * (1.0f / C);
      s2 = red16(s2) * (1.0f / C);
      if (live) {
#pragma unroll
        for (int j = 0; j < H; ++j) {
          float4 o;
          o.x = (gg[j].x - s1 - xh[j].x * s2) * rs; o.y = (gg[j].y - s1 - xh[j].y * s2) * rs;
          o.z = (gg[j].z - s1 - xh[j].z * s2) * rs; o.w = (gg[j].w - s1 - xh[j].w * s2) * rs;
          reinterpret_cast<float4*>(p.gx + row * C)[q + 16 * j] = o;
        }
      }
    }
  }
  // ---- per-CTA partials: gcenters fragments -> [K][C]
  {
    float* gc = p.part_gc + (size_t)blockIdx.x * kBK * C;
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      const int c = 8 * (wn * NTW + j) + 2 * t4, k = 16 * wm + g;
      *reinterpret_cast<float2*>(gc + (size_t)k * C + c) = make_float2(acc_gc[j][0], acc_gc[j][1]);
      *reinterpret_cast<float2*>(gc + (size_t)(k + 8) * C + c) = make_float2(acc_gc[j][2], acc_gc[j][3]);
    }
  }
  __syncthreads();
  // gamma / beta: reduce over the 32 token slots through shared memory (re-using the gR / z tiles)
  float* red_w = sG;     // [32][LDT]
  float* red_b = sZ;
#pragma unroll
  for (int j = 0; j < H; ++j) {
    *reinterpret_cast<float4*>(red_w + t * LDT + 4 * (q + 16 * j)) = acc_gw[j];
    *reinterpret_cast<float4*>(red_b + t * LDT + 4 * (q + 16 * j)) = acc_gb[j];
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) sA[t * LDK + q + 16 * i] = acc_rcol[i];
  __syncthreads();
  for (int c = tid; c < C; c += kBThreads) {
    float sw = 0.f, sb = 0.f;
    for (int tt = 0; tt < kBT; ++tt) { sw += red_w[tt * LDT + c]; sb += red_b[tt * LDT + c]; }
    p.part_ln[(size_t)blockIdx.x * 2 * C + c] = sw;
    p.part_ln[(size_t)blockIdx.x * 2 * C + C + c] = sb;
  }
  if (tid < kBK) {
    float s = 0.f;
    for (int tt = 0; tt < kBT; ++tt) s += sA[tt * LDK + tid];
    p.part_rcol[(size_t)blockIdx.x * kBK + tid] = s;
  }
}

// gcenters = sum_b part_gc[b] + centers * (sum_b part_rcol[b]);  g_ln_w / g_ln_b = sum_b part_ln[b]
__global__ void __launch_bounds__(256)
cluster_bwd_finalize_kernel(const float* __restrict__ part_gc, const float* __restrict__ part_rcol,
                            const float* __restrict__ part_ln, const float* __restrict__ centers,
                            int nb, int K, int C, float* __restrict__ gcenters,
                            float* __restrict__ gw, float* __restrict__ gb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int KC = K * C;
  if (i < KC) {
    float s = 0.f, rc = 0.f;
    const int k = i / C;
    for (int b = 0; b < nb; ++b) { s += part_gc[(size_t)b * KC + i]; rc += part_rcol[(size_t)b * K + k]; }
    gcenters[i] = s + centers[i] * rc;
  } else if (i < KC + 2 * C) {
    const int c = i - KC;
    float s = 0.f;
    for (int b = 0; b < nb; ++b) s += part_ln[(size_t)b * 2 * C + c];
    if (c < C) gw[c] = s; else gb[c - C] = s;
  }
}

bool bwd_fused_shape_ok(long long N, int C, int K) {
  return N >= 1 && K == kBK && (C == 64 || C == 128 || C == 192 || C == 256);
}

int bwd_fused_grid(long long N) {
  long long panels = (N + kBT - 1) / kBT;
  long long g = std::min<long long>(panels, (long long)sm_count());
  return (int)std::max<long long>(g, 1);
}

size_t bwd_fused_workspace_bytes(long long N, int C, int K) {
  if (!bwd_fused_shape_ok(N, C, K)) return 0;
  size_t g = (size_t)sm_count() * 2;
  return align_up(g * K * C * sizeof(float), 256) + align_up(g * K * sizeof(float), 256) +
         align_up(g * 2 * C * sizeof(float), 256) + 256;
}

template <int F4>
static int launch_bwd(const BwdParams& p, int grid, cudaStream_t st) {
  const size_t smem = BwdSmem<F4>::bytes;
  VADC_CUDA(cudaFuncSetAttribute(cluster_bwd_fused_kernel<F4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cluster_bwd_fused_kernel<F4><<<grid, kBThreads, smem, st>>>(p);
  VADC_CHECK_LAUNCH("cluster_bwd_fused_kernel");
  return VADC_OK;
}

int launch_cluster_bwd_fused(const float* x, const float* mu, const float* rstd, const float* feature,
                             const float* ln_w, const float* centers, const float* D, const float* A,
                             const float* gD, const float* gA, const float* gR, const float* gF,
                             const float* g_loss_sq, long long N, int C, int K, float alpha, float* gx,
                             float* gcenters, float* g_ln_w, float* g_ln_b, void* workspace,
                             size_t workspace_bytes, cudaStream_t st) {
  if (!bwd_fused_shape_ok(N, C, K)) return VADC_ERR_UNSUPPORTED;
  if (workspace_bytes < bwd_fused_workspace_bytes(N, C, K)) return VADC_ERR_WORKSPACE;
  Carver ws(workspace, workspace_bytes);
  const size_t g = (size_t)sm_count() * 2;
  float* part_gc = ws.take<float>(g * K * C);
  float* part_rcol = ws.take<float>(g * K);
  float* part_ln = ws.take<float>(g * 2 * C);
  const int grid = bwd_fused_grid(N);
  BwdParams p{x, mu, rstd, feature, ln_w, centers, D, A, gD, gA, gR, gF, g_loss_sq, N, alpha,
              gx, part_gc, part_rcol, part_ln};
  int rc;
  switch (C) {
    case 64: rc = launch_bwd<2>(p, grid, st); break;
    case 128: rc = launch_bwd<4>(p, grid, st); break;
    case 192: rc = launch_bwd<6>(p, grid, st); break;
    default: rc = launch_bwd<8>(p, grid, st); break;
  }
  if (rc) return rc;
  const int tot = K * C + 2 * C;
  cluster_bwd_finalize_kernel<<<(tot + 255) / 256, 256, 0, st>>>(part_gc, part_rcol, part_ln, centers, grid, K, C,
                                                                  gcenters, g_ln_w, g_ln_b);
  VADC_CHECK_LAUNCH("cluster_bwd_finalize_kernel");
  return VADC_OK;
}

}  // namespace vadc
