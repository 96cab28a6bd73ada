// reduce.cu — bandwidth-bound streams: pixel losses (L2, L3) and per-frame
// reconstruction error / PSNR / regularity score (E1-E3).
//
// Each is one pass over its inputs: 128-bit coalesced, L1-bypassing loads,
// four independent loads in flight per thread, warp-shuffle + shared-memory
// block reduction, fp64 accumulation across blocks, fixed-order final sum.
#include "common.cuh"

namespace vadc {

template <int MODE>
__device__ __forceinline__ float term(float x, float t) {
  float e = x - t;
  if (MODE == VADC_LOSS_L1_MEAN) return fabsf(e);
  float e2 = e * e;
  if (MODE == VADC_LOSS_MSE_MEAN) return e2;
  return e2 * e2;
}

constexpr int kRedThreads = 512;
constexpr int kRedUnroll = 4;

template <int MODE>
__global__ void __launch_bounds__(kRedThreads)
pixel_loss_kernel(const float* __restrict__ x, const float* __restrict__ t, long long n,
                  const float* __restrict__ xp, long long npad, double* __restrict__ partial) {
  __shared__ double red[32];
  const long long nv = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* t4 = reinterpret_cast<const float4*>(t);
  const long long stride = (long long)gridDim.x * kRedThreads;
  long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x;
  float acc[kRedUnroll] = {0.f, 0.f, 0.f, 0.f};
  for (; i + (kRedUnroll - 1) * stride < nv; i += kRedUnroll * stride) {
    float4 a[kRedUnroll], b[kRedUnroll];
#pragma unroll
    for (int u = 0; u < kRedUnroll; ++u) { a[u] = ld_stream(x4 + i + u * stride); b[u] = ld_stream(t4 + i + u * stride); }
#pragma unroll
    for (int u = 0; u < kRedUnroll; ++u)
      acc[u] += (term<MODE>(a[u].x, b[u].x) + term<MODE>(a[u].y, b[u].y)) +
                (term<MODE>(a[u].z, b[u].z) + term<MODE>(a[u].w, b[u].w));
  }
  for (; i < nv; i += stride) {
    float4 a = ld_stream(x4 + i), b = ld_stream(t4 + i);
    acc[0] += (term<MODE>(a.x, b.x) + term<MODE>(a.y, b.y)) + (term<MODE>(a.z, b.z) + term<MODE>(a.w, b.w));
  }
  // scalar tail and the zero-padded part of the target (Recon_Loss.py:25-26)
  long long g = (long long)blockIdx.x * kRedThreads + threadIdx.x;
  for (long long j = (nv << 2) + g; j < n; j += stride) acc[1] += term<MODE>(x[j], t[j]);
  for (long long j = g; j < npad; j += stride) acc[2] += term<MODE>(xp[j], 0.f);
  double s = (double)((acc[0] + acc[1]) + (acc[2] + acc[3]));
  s = block_sum<double>(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256)
pixel_loss_finalize_kernel(const double* __restrict__ partial, int nb, long long n_total, int mode,
                           float* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partial[i];
  s = block_sum<double>(s, red);
  if (threadIdx.x == 0) {
    out[1] = (float)s;
    out[0] = (mode == VADC_LOSS_E4_NORM) ? (float)sqrt(s) : (float)(s / (double)n_total);
  }
}

// d loss / d x, elementwise
template <int MODE>
__global__ void __launch_bounds__(256)
pixel_loss_bwd_kernel(const float* __restrict__ x, const float* __restrict__ t, long long n,
                      const float* __restrict__ gout, const float* __restrict__ fwd,
                      long long n_total, float* __restrict__ gx) {
  float g = __ldg(gout), sc;
  if (MODE == VADC_LOSS_L1_MEAN) sc = g / (float)n_total;
  else if (MODE == VADC_LOSS_MSE_MEAN) sc = 2.0f * g / (float)n_total;
  else { float L = __ldg(fwd); sc = (L > 0.f) ? 2.0f * g / L : 0.f; }
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float tv = t ? t[i] : 0.f;
    float e = x[i] - tv, o;
    if (MODE == VADC_LOSS_L1_MEAN) o = (e > 0.f) ? sc : ((e < 0.f) ? -sc : 0.f);
    else if (MODE == VADC_LOSS_MSE_MEAN) o = sc * e;
    else o = sc * e * e * e;
    gx[i] = o;
  }
}

// ---------------------------------------------------------------------------
// E1: per-frame MSE.  recon/clip [B,Cc,T,HW]; frame (b,t) is Cc segments of HW
// contiguous floats.  gridDim = (slices, B*T): each block reduces a slice of
// one frame; partials go to a [B*T, slices] buffer and are summed in fixed
// order by the finalize kernel (deterministic, no atomics).
// ---------------------------------------------------------------------------
struct FrameStrides { long long rb, rc, rt, cb, cc, ct; };   // element strides of recon / clip along batch, channel, frame

__global__ void __launch_bounds__(256)
frame_mse_kernel(const float* __restrict__ recon, const float* __restrict__ clip, int Cc, int T,
                 long long HW, int slices, const FrameStrides fs, double* __restrict__ partial) {
  __shared__ double red[32];
  const int frame = blockIdx.y;           // b*T + t
  const int b = frame / T, t = frame % T;
  const int slice = blockIdx.x;
  const long long nv = HW >> 2;           // float4 per channel plane (HW % 4 == 0 on this path)
  const long long per = (nv + slices - 1) / slices;
  const long long v0 = slice * per, v1 = min(nv, v0 + per);
  float acc0 = 0.f, acc1 = 0.f;
  for (int c = 0; c < Cc; ++c) {
    const float4* r4 = reinterpret_cast<const float4*>(recon + b * fs.rb + c * fs.rc + t * fs.rt);
    const float4* c4 = reinterpret_cast<const float4*>(clip + b * fs.cb + c * fs.cc + t * fs.ct);
    long long i = v0 + threadIdx.x;
    for (; i + 256 < v1; i += 512) {
      float4 a0 = ld_stream(r4 + i), b0 = ld_stream(c4 + i);
      float4 a1 = ld_stream(r4 + i + 256), b1 = ld_stream(c4 + i + 256);
      acc0 += (term<1>(a0.x, b0.x) + term<1>(a0.y, b0.y)) + (term<1>(a0.z, b0.z) + term<1>(a0.w, b0.w));
      acc1 += (term<1>(a1.x, b1.x) + term<1>(a1.y, b1.y)) + (term<1>(a1.z, b1.z) + term<1>(a1.w, b1.w));
    }
    for (; i < v1; i += 256) {
      float4 a0 = ld_stream(r4 + i), b0 = ld_stream(c4 + i);
      acc0 += (term<1>(a0.x, b0.x) + term<1>(a0.y, b0.y)) + (term<1>(a0.z, b0.z) + term<1>(a0.w, b0.w));
    }
  }
  double s = block_sum<double>((double)(acc0 + acc1), red);
  if (threadIdx.x == 0) partial[(long long)frame * slices + slice] = s;
}

// generic (HW % 4 != 0) fallback: scalar loads
__global__ void __launch_bounds__(256)
frame_mse_scalar_kernel(const float* __restrict__ recon, const float* __restrict__ clip, int Cc,
                        int T, long long HW, const FrameStrides fs, double* __restrict__ partial) {
  __shared__ double red[32];
  const int frame = blockIdx.x;
  const int b = frame / T, t = frame % T;
  float acc = 0.f;
  for (int c = 0; c < Cc; ++c) {
    const float* rp = recon + b * fs.rb + c * fs.rc + t * fs.rt;
    const float* cp = clip + b * fs.cb + c * fs.cc + t * fs.ct;
    for (long long i = threadIdx.x; i < HW; i += 256) acc += term<1>(rp[i], cp[i]);
  }
  double s = block_sum<double>((double)acc, red);
  if (threadIdx.x == 0) partial[frame] = s;
}

__global__ void __launch_bounds__(256)
frame_mse_finalize_kernel(const double* __restrict__ partial, int frames, int slices, double denom,
                          float* __restrict__ mse, double* __restrict__ psnr) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= frames) return;
  double s = 0.0;
  for (int i = 0; i < slices; ++i) s += partial[(long long)f * slices + i];
  float m = (float)(s / denom);
  mse[f] = m;
  if (psnr) psnr[f] = 10.0 * log10(1.0 / (double)m);   // misc/utils.py:128 on the fp32 value
}

// E3: per-video 1 - (p - min) / (max - min), one block per video
__global__ void __launch_bounds__(256)
minmax_score_kernel(const double* __restrict__ psnr, const long long* __restrict__ off,
                    double* __restrict__ score) {
  __shared__ double smin[256], smax[256];
  const long long a = off[blockIdx.x], b = off[blockIdx.x + 1];
  double lo = INFINITY, hi = -INFINITY;
  for (long long i = a + threadIdx.x; i < b; i += 256) { double v = psnr[i]; lo = fmin(lo, v); hi = fmax(hi, v); }
  smin[threadIdx.x] = lo; smax[threadIdx.x] = hi;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + s]);
      smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + s]);
    }
    __syncthreads();
  }
  lo = smin[0]; hi = smax[0];
  for (long long i = a + threadIdx.x; i < b; i += 256) score[i] = 1.0 - (psnr[i] - lo) / (hi - lo);
}

static int reduce_blocks(long long n) {
  long long b = (n / 4 + (long long)kRedThreads * kRedUnroll - 1) / ((long long)kRedThreads * kRedUnroll);
  long long cap = (long long)sm_count() * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace vadc

using namespace vadc;

extern "C" size_t vadc_pixel_loss_workspace_bytes(int64_t n) {
  (void)n;
  return align_up((size_t)(sm_count() * 4 + 1) * sizeof(double), 256) + 256;
}

extern "C" int vadc_pixel_loss(const float* x, const float* t, int64_t n, const float* x_pad,
                               int64_t n_pad, int mode, float* out, void* workspace,
                               size_t workspace_bytes, void* stream) {
  VADC_REQUIRE(n >= 0 && n_pad >= 0 && n + n_pad > 0, VADC_ERR_BAD_SHAPE);
  VADC_REQUIRE(mode >= 0 && mode <= 2, VADC_ERR_BAD_SHAPE);
  VADC_REQUIRE(out && workspace && (n == 0 || (x && t)) && (n_pad == 0 || x_pad), VADC_ERR_NULL_POINTER);
  VADC_REQUIRE(aligned16(x) && aligned16(t), VADC_ERR_MISALIGNED);
  VADC_REQUIRE(workspace_bytes >= vadc_pixel_loss_workspace_bytes(n), VADC_ERR_WORKSPACE);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(workspace);
  int nb = reduce_blocks(n > n_pad ? n : n_pad);
  switch (mode) {
    case VADC_LOSS_L1_MEAN: pixel_loss_kernel<0><<<nb, kRedThreads, 0, st>>>(x, t, n, x_pad, n_pad, partial); break;
    case VADC_LOSS_MSE_MEAN: pixel_loss_kernel<1><<<nb, kRedThreads, 0, st>>>(x, t, n, x_pad, n_pad, partial); break;
    default: pixel_loss_kernel<2><<<nb, kRedThreads, 0, st>>>(x, t, n, x_pad, n_pad, partial); break;
  }
  VADC_CHECK_LAUNCH("pixel_loss_kernel");
  pixel_loss_finalize_kernel<<<1, 256, 0, st>>>(partial, nb, n + n_pad, mode, out);
  VADC_CHECK_LAUNCH("pixel_loss_finalize_kernel");
  return VADC_OK;
}

extern "C" int vadc_pixel_loss_bwd(const float* x, const float* t, int64_t n, int mode,
                                   const float* gout, const float* out_fwd, int64_t n_total,
                                   float* gx, void* stream) {
  VADC_REQUIRE(n >= 0 && n_total > 0 && mode >= 0 && mode <= 2, VADC_ERR_BAD_SHAPE);
  if (n == 0) return VADC_OK;
  VADC_REQUIRE(x && gout && out_fwd && gx, VADC_ERR_NULL_POINTER);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long nb = (n + 1023) / 1024;
  long long cap = (long long)sm_count() * 16;
  if (nb > cap) nb = cap;
  switch (mode) {
    case VADC_LOSS_L1_MEAN: pixel_loss_bwd_kernel<0><<<(unsigned)nb, 256, 0, st>>>(x, t, n, gout, out_fwd, n_total, gx); break;
    case VADC_LOSS_MSE_MEAN: pixel_loss_bwd_kernel<1><<<(unsigned)nb, 256, 0, st>>>(x, t, n, gout, out_fwd, n_total, gx); break;
    default: pixel_loss_bwd_kernel<2><<<(unsigned)nb, 256, 0, st>>>(x, t, n, gout, out_fwd, n_total, gx); break;
  }
  VADC_CHECK_LAUNCH("pixel_loss_bwd_kernel");
  return VADC_OK;
}

// slices per frame so that small batches still fill the GPU
static int frame_slices(int frames, long long HW, int Cc) {
  long long per_frame_v4 = (HW / 4) * Cc;
  int s = 1;
  long long want = (long long)sm_count() * 8;
  while ((long long)frames * s < want && s < 64 && per_frame_v4 / (s * 2) >= 2048) s *= 2;
  return s;
}

extern "C" size_t vadc_frame_mse_workspace_bytes(int B, int T, int64_t HW, int Cc) {
  return align_up((size_t)B * T * frame_slices(B * T, HW, Cc) * sizeof(double), 256) + 256;
}

static int frame_mse_impl(const float* recon, const float* clip, int B, int Cc, int T, int64_t HW, const FrameStrides fs,
                          float* mse, double* psnr, void* workspace, size_t workspace_bytes, void* stream) {
  VADC_REQUIRE(B >= 0 && Cc > 0 && T >= 0 && HW > 0, VADC_ERR_BAD_SHAPE);
  if (B == 0 || T == 0) return VADC_OK;
  VADC_REQUIRE(recon && clip && mse && workspace, VADC_ERR_NULL_POINTER);
  VADC_REQUIRE((long long)B * T < 65536ll * 1024, VADC_ERR_UNSUPPORTED);
  VADC_REQUIRE(workspace_bytes >= vadc_frame_mse_workspace_bytes(B, T, HW, Cc), VADC_ERR_WORKSPACE);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int frames = B * T;
  double* partial = static_cast<double*>(workspace);
  int slices = 1;
  const bool vec = (HW & 3) == 0 && aligned16(recon) && aligned16(clip) &&
                   ((fs.rb | fs.rc | fs.rt | fs.cb | fs.cc | fs.ct) & 3) == 0;    // every plane starts on a 16-byte boundary
  if (vec && frames <= 65535) {
    slices = frame_slices(frames, HW, Cc);
    dim3 grid(slices, frames);
    frame_mse_kernel<<<grid, 256, 0, st>>>(recon, clip, Cc, T, HW, slices, fs, partial);
    VADC_CHECK_LAUNCH("frame_mse_kernel");
  } else {
    frame_mse_scalar_kernel<<<frames, 256, 0, st>>>(recon, clip, Cc, T, HW, fs, partial);
    VADC_CHECK_LAUNCH("frame_mse_scalar_kernel");
  }
  frame_mse_finalize_kernel<<<(frames + 255) / 256, 256, 0, st>>>(partial, frames, slices, (double)Cc * (double)HW, mse, psnr);
  VADC_CHECK_LAUNCH("frame_mse_finalize_kernel");
  return VADC_OK;
}

extern "C" int vadc_frame_mse(const float* recon, const float* clip, int B, int Cc, int T,
                                 int64_t HW, float* mse, double* psnr, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  VADC_REQUIRE(aligned16(recon) && aligned16(clip), VADC_ERR_MISALIGNED);
  const FrameStrides fs{(long long)Cc * T * HW, (long long)T * HW, HW, (long long)Cc * T * HW, (long long)T * HW, HW};
  return frame_mse_impl(recon, clip, B, Cc, T, HW, fs, mse, psnr, workspace, workspace_bytes, stream);
}

// the same reduction over STRIDED views: clip batches taken out of a device-resident video without a copy (consecutive
// or overlapping clips), single frames of a reconstruction (recon[:, :, 0]); strides in elements
extern "C" int vadc_frame_mse_strided(const float* recon, int64_t recon_stride_b, int64_t recon_stride_c, int64_t recon_stride_t,
                                      const float* clip, int64_t clip_stride_b, int64_t clip_stride_c, int64_t clip_stride_t,
                                      int B, int Cc, int T, int64_t HW, float* mse, double* psnr,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  VADC_REQUIRE(recon_stride_b >= 0 && recon_stride_c >= 0 && recon_stride_t >= 0 && clip_stride_b >= 0 &&
               clip_stride_c >= 0 && clip_stride_t >= 0, VADC_ERR_BAD_SHAPE);
  const FrameStrides fs{recon_stride_b, recon_stride_c, recon_stride_t, clip_stride_b, clip_stride_c, clip_stride_t};
  return frame_mse_impl(recon, clip, B, Cc, T, HW, fs, mse, psnr, workspace, workspace_bytes, stream);
}

extern "C" int vadc_minmax_score(const double* psnr, const int64_t* seg_offsets, int n_videos,
                                 double* score, void* stream) {
  VADC_REQUIRE(n_videos >= 0, VADC_ERR_BAD_SHAPE);
  if (n_videos == 0) return VADC_OK;
  VADC_REQUIRE(psnr && seg_offsets && score, VADC_ERR_NULL_POINTER);
  minmax_score_kernel<<<n_videos, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      psnr, reinterpret_cast<const long long*>(seg_offsets), score);
  VADC_CHECK_LAUNCH("minmax_score_kernel");
  return VADC_OK;
}
