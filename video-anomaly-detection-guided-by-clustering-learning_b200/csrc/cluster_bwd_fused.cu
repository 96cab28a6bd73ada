// cluster_bwd_fused.cu — C2: fused backward of the cluster head for small K
// (K == 32, C in {64,128,192,256}), fp32 CUDA cores.
//
// One persistent CTA (256 threads) walks 32-token panels; every input is read
// from HBM exactly once (x, feature, gR, D, A [+ gD, gA, gF]) and only gx is
// written per token:
//   S1  G1[t,k]  = sum_c gR[t,c] cen[k,c]                      (x_rec = A @ centers)
//   S2  softmin backward + cdist ratio -> r[t,k], rsum[t]      (rows.cuh::bwd_rows_kernel math)
//   S3  gz[t,c]  = z[t,c] rsum[t] - sum_k r[t,k] cen[k,c] + gF
//   S4  LayerNorm backward -> gx[t,:]; gamma/beta partials in registers
//   S5  gcen[k,c] += A[t,k] gR[t,c] - r[t,k] z[t,c]            (register tile, whole kernel)
// Per-CTA partials (gcenters, colsum r, gamma, beta) go to the workspace and a
// small second kernel adds them in fixed order (deterministic).
//
// Thread maps (tid = 8*t + q, t = token in panel, q = 0..7):
//   S1/S2: thread (t, q) owns centroids k = q, q+8, q+16, q+24
//   S3/S4: thread (t, q) owns float4 chunks q, q+8, ... (F4 per thread) of token t
//   S5   : thread (k = tid/8, q) owns chunks q+8j of centroid k
#include "common.cuh"
#include "cluster.h"
#include <algorithm>

namespace vadc {

constexpr int kBT = 32;          // tokens per panel
constexpr int kBK = 32;          // centroids (this kernel is specialised for K == 32)
constexpr int kBThreads = 256;

template <int F4>
struct BwdSmem {
  static constexpr int C = F4 * 32;
  static constexpr int LDC = C + 4;                      // centroid row stride (conflict-free float4 rows)
  static constexpr int kCen = kBK * LDC;
  static constexpr int kTile = kBT * C;                  // gR / z tiles
  static constexpr int kTK = kBT * (kBK + 4);            // A / r tiles, padded rows
  static constexpr size_t bytes = sizeof(float) * (kCen + 2 * kTile + 2 * kTK + 64);
};

struct BwdParams {
  const float* x; const float* mu; const float* rstd; const float* feature; const float* ln_w;
  const float* centers; const float* D; const float* A;
  const float* gD; const float* gA; const float* gR; const float* gF; const float* g_loss_sq;
  long long N; float alpha;
  float* gx; float* part_gc; float* part_rcol; float* part_ln;   // [grid][K*C], [grid][K], [grid][2C]
};

template <int F4>
__global__ void __launch_bounds__(kBThreads, 1)
cluster_bwd_fused_kernel(const BwdParams p) {
  using S = BwdSmem<F4>;
  constexpr int C = S::C, LDC = S::LDC, LDK = kBK + 4;
  extern __shared__ __align__(16) float sm[];
  float* sCen = sm;                        // [32][C+4]
  float* sG = sCen + S::kCen;              // gR tile [32][C]
  float* sZ = sG + S::kTile;               // feature tile [32][C]
  float* sA = sZ + S::kTile;               // A tile [32][36]
  float* sR = sA + S::kTK;                 // r tile [32][36]
  const int tid = threadIdx.x, t = tid >> 3, q = tid & 7;
  const float sc = p.g_loss_sq ? 2.0f * __ldg(p.g_loss_sq) : 0.f;

  for (int i = tid; i < kBK * (C / 4); i += kBThreads) {
    int k = i / (C / 4), c4 = i % (C / 4);
    *reinterpret_cast<float4*>(sCen + k * LDC + 4 * c4) = __ldg(reinterpret_cast<const float4*>(p.centers + (size_t)k * C) + c4);
  }
  float4 gam[F4];
#pragma unroll
  for (int j = 0; j < F4; ++j) gam[j] = __ldg(reinterpret_cast<const float4*>(p.ln_w) + q + 8 * j);

  float4 acc_gc[F4];          // S5: centroid k5 = tid/8, chunks q + 8j
  float4 acc_gw[F4], acc_gb[F4];
  float acc_rcol[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < F4; ++j) {
    acc_gc[j] = make_float4(0, 0, 0, 0); acc_gw[j] = acc_gc[j]; acc_gb[j] = acc_gc[j];
  }
  const long long npanels = (p.N + kBT - 1) / kBT;

  for (long long pn = blockIdx.x; pn < npanels; pn += gridDim.x) {
    const long long row = pn * kBT + t;
    const bool live = row < p.N;
    __syncthreads();                                     // previous panel's tiles are no longer read
    // ---- stage gR / z tiles (thread (t,q): chunks q+8j of token t), x kept in registers for S4
    float4 xv[F4];
#pragma unroll
    for (int j = 0; j < F4; ++j) {
      const int c4 = q + 8 * j;
      float4 g = make_float4(0, 0, 0, 0), z = g;
      xv[j] = g;
      if (live) {
        if (p.gR) g = ld_stream(reinterpret_cast<const float4*>(p.gR + row * C) + c4);
        z = ld_stream(reinterpret_cast<const float4*>(p.feature + row * C) + c4);
        xv[j] = ld_stream(reinterpret_cast<const float4*>(p.x + row * C) + c4);
      }
      *reinterpret_cast<float4*>(sG + t * C + 4 * c4) = g;
      *reinterpret_cast<float4*>(sZ + t * C + 4 * c4) = z;
    }
    // D / A of this thread's four centroids k = q + 8i
    float dk[4], ak[4], gdk[4], gak[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = q + 8 * i;
      dk[i] = live ? __ldg(p.D + row * kBK + k) : 1.f;
      ak[i] = live ? __ldg(p.A + row * kBK + k) : 0.f;
      gdk[i] = (live && p.gD) ? __ldg(p.gD + row * kBK + k) : 0.f;
      gak[i] = (live && p.gA) ? __ldg(p.gA + row * kBK + k) : 0.f;
      sA[t * LDK + k] = ak[i];
    }
    __syncthreads();
    // ---- S1: G1[t, q+8i] = gR[t,:] . cen[q+8i,:]
    float g1[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.gR) {
      const float* gr = sG + t * C;
#pragma unroll 4
      for (int c = 0; c < C; c += 4) {
        const float4 g = *reinterpret_cast<const float4*>(gr + c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 w = *reinterpret_cast<const float4*>(sCen + (q + 8 * i) * LDC + c);
          g1[i] = fmaf(g.x, w.x, g1[i]); g1[i] = fmaf(g.y, w.y, g1[i]);
          g1[i] = fmaf(g.z, w.z, g1[i]); g1[i] = fmaf(g.w, w.w, g1[i]);
        }
      }
    }
    // ---- S2: softmin backward + cdist ratio over the 8 lanes of token t
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      g1[i] += gak[i] + sc * dk[i] * dk[i] * ak[i];      // gA_tot
      dot = fmaf(g1[i], ak[i], dot);
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    float rsum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float gd = gdk[i] + sc * dk[i] * ak[i] * ak[i] - p.alpha * ak[i] * (g1[i] - dot);
      float r = (dk[i] == 0.f || !live) ? 0.f : gd / dk[i];
      sR[t * LDK + q + 8 * i] = r;
      acc_rcol[i] += r;
      rsum += r;
    }
    rsum += __shfl_xor_sync(0xffffffffu, rsum, 1);
    rsum += __shfl_xor_sync(0xffffffffu, rsum, 2);
    rsum += __shfl_xor_sync(0xffffffffu, rsum, 4);
    __syncthreads();
    // ---- S3: gz[t, chunks] = z * rsum - r[t,:] @ cen + gF
    float4 gz[F4];
#pragma unroll
    for (int j = 0; j < F4; ++j) gz[j] = make_float4(0, 0, 0, 0);
    {
      const float* rr = sR + t * LDK;
#pragma unroll 2
      for (int k = 0; k < kBK; k += 4) {
        const float4 r4 = *reinterpret_cast<const float4*>(rr + k);
        const float rk[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float* cr = sCen + (k + kk) * LDC;
#pragma unroll
          for (int j = 0; j < F4; ++j) {
            const float4 w = *reinterpret_cast<const float4*>(cr + 4 * (q + 8 * j));
            gz[j].x = fmaf(rk[kk], w.x, gz[j].x); gz[j].y = fmaf(rk[kk], w.y, gz[j].y);
            gz[j].z = fmaf(rk[kk], w.z, gz[j].z); gz[j].w = fmaf(rk[kk], w.w, gz[j].w);
          }
        }
      }
    }
    // ---- S4: LayerNorm backward of token t (8 lanes), gamma / beta partials
    {
      const float m = live ? __ldg(p.mu + row) : 0.f, rs = live ? __ldg(p.rstd + row) : 0.f;
      float s1 = 0.f, s2 = 0.f;
      float4 gg[F4], xh[F4];
#pragma unroll
      for (int j = 0; j < F4; ++j) {
        const int c4 = q + 8 * j;
        const float4 z = *reinterpret_cast<const float4*>(sZ + t * C + 4 * c4);
        float4 g;
        g.x = z.x * rsum - gz[j].x; g.y = z.y * rsum - gz[j].y;
        g.z = z.z * rsum - gz[j].z; g.w = z.w * rsum - gz[j].w;
        if (p.gF && live) {
          const float4 f = ld_stream(reinterpret_cast<const float4*>(p.gF + row * C) + c4);
          g.x += f.x; g.y += f.y; g.z += f.z; g.w += f.w;
        }
        xh[j].x = (xv[j].x - m) * rs; xh[j].y = (xv[j].y - m) * rs;
        xh[j].z = (xv[j].z - m) * rs; xh[j].w = (xv[j].w - m) * rs;
        acc_gw[j].x = fmaf(g.x, xh[j].x, acc_gw[j].x); acc_gw[j].y = fmaf(g.y, xh[j].y, acc_gw[j].y);
        acc_gw[j].z = fmaf(g.z, xh[j].z, acc_gw[j].z); acc_gw[j].w = fmaf(g.w, xh[j].w, acc_gw[j].w);
        acc_gb[j].x += g.x; acc_gb[j].y += g.y; acc_gb[j].z += g.z; acc_gb[j].w += g.w;
        gg[j].x = g.x * gam[j].x; gg[j].y = g.y * gam[j].y; gg[j].z = g.z * gam[j].z; gg[j].w = g.w * gam[j].w;
        s1 += (gg[j].x + gg[j].y) + (gg[j].z + gg[j].w);
        s2 += (gg[j].x * xh[j].x + gg[j].y * xh[j].y) + (gg[j].z * xh[j].z + gg[j].w * xh[j].w);
      }
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 2); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 4); s2 += __shfl_xor_sync(0xffffffffu, s2, 4);
      s1 *= (1.0f / C); s2 *= (1.0f / C);
      if (live) {
#pragma unroll
        for (int j = 0; j < F4; ++j) {
          float4 o;
          o.x = (gg[j].x - s1 - xh[j].x * s2) * rs; o.y = (gg[j].y - s1 - xh[j].y * s2) * rs;
          o.z = (gg[j].z - s1 - xh[j].z * s2) * rs; o.w = (gg[j].w - s1 - xh[j].w * s2) * rs;
          reinterpret_cast<float4*>(p.gx + row * C)[q + 8 * j] = o;
        }
      }
    }
    // ---- S5: gcen[k5, chunks] += A[tt,k5] gR[tt,chunks] - r[tt,k5] z[tt,chunks] over the panel
    {
      const int k5 = t;                                  // tid/8 doubles as the centroid index here
#pragma unroll 2
      for (int tt = 0; tt < kBT; ++tt) {
        const float a = sA[tt * LDK + k5], r = sR[tt * LDK + k5];
#pragma unroll
        for (int j = 0; j < F4; ++j) {
          const float4 g = *reinterpret_cast<const float4*>(sG + tt * C + 4 * (q + 8 * j));
          const float4 z = *reinterpret_cast<const float4*>(sZ + tt * C + 4 * (q + 8 * j));
          acc_gc[j].x = fmaf(a, g.x, acc_gc[j].x); acc_gc[j].x = fmaf(-r, z.x, acc_gc[j].x);
          acc_gc[j].y = fmaf(a, g.y, acc_gc[j].y); acc_gc[j].y = fmaf(-r, z.y, acc_gc[j].y);
          acc_gc[j].z = fmaf(a, g.z, acc_gc[j].z); acc_gc[j].z = fmaf(-r, z.z, acc_gc[j].z);
          acc_gc[j].w = fmaf(a, g.w, acc_gc[j].w); acc_gc[j].w = fmaf(-r, z.w, acc_gc[j].w);
        }
      }
    }
  }
  // ---- per-CTA partials
  {
    float* gc = p.part_gc + (size_t)blockIdx.x * kBK * C + (size_t)t * C;
#pragma unroll
    for (int j = 0; j < F4; ++j) reinterpret_cast<float4*>(gc)[q + 8 * j] = acc_gc[j];
  }
  __syncthreads();
  // gamma / beta: reduce over the 32 token slots through shared memory (re-using the gR / z tiles)
  float* red_w = sG;     // [32][C]
  float* red_b = sZ;
#pragma unroll
  for (int j = 0; j < F4; ++j) {
    *reinterpret_cast<float4*>(red_w + t * C + 4 * (q + 8 * j)) = acc_gw[j];
    *reinterpret_cast<float4*>(red_b + t * C + 4 * (q + 8 * j)) = acc_gb[j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) sA[t * LDK + q + 8 * i] = acc_rcol[i];
  __syncthreads();
  for (int c = tid; c < C; c += kBThreads) {
    float sw = 0.f, sb = 0.f;
    for (int tt = 0; tt < kBT; ++tt) { sw += red_w[tt * C + c]; sb += red_b[tt * C + c]; }
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
  long long g = std::min<long long>(panels, (long long)sm_count() * 2);
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
