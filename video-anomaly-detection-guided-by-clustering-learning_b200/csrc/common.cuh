// common.cuh — shared helpers for libvadc (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/vadc.h"

namespace vadc {

extern thread_local char g_last_cuda_error[256];
extern unsigned long long g_launch_count;

int record_cuda_error(cudaError_t e, const char* what);
inline void count_launch(int n = 1) { __atomic_fetch_add(&g_launch_count, (unsigned long long)n, __ATOMIC_RELAXED); }

#define VADC_CHECK_LAUNCH(what)                                        \
  do {                                                                 \
    ::vadc::count_launch();                                            \
    cudaError_t e__ = cudaGetLastError();                              \
    if (e__ != cudaSuccess) return ::vadc::record_cuda_error(e__, what); \
  } while (0)

#define VADC_CUDA(call)                                                \
  do {                                                                 \
    cudaError_t e__ = (call);                                          \
    if (e__ != cudaSuccess) return ::vadc::record_cuda_error(e__, #call); \
  } while (0)

#define VADC_REQUIRE(cond, code) do { if (!(cond)) return (code); } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// carve a caller-provided workspace
struct Carver {
  char* base; size_t off, cap;
  Carver(void* p, size_t c) : base(static_cast<char*>(p)), off(0), cap(c) {}
  template <class T> T* take(size_t n) {
    size_t b = align_up(n * sizeof(T), 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += b;
    return r;
  }
  bool ok() const { return off <= cap; }
};

int sm_count();
// Environment switches (kernel-variant selection for A/B runs and debugging) are read ONCE per process, at their first
// use: an entry point costs a table lookup, never a getenv().  vadc_refresh_env() (tests) forgets the cached values.
const char* env_str(const char* name);               // value or nullptr
inline bool env_on(const char* name) { return env_str(name) != nullptr; }
int env_int(const char* name, int dflt);
// measurement aid (abi.cu): events around the dominant kernels when vadc_timing_enable(1) is in effect
#define VADC_TIMING_SLOTS 2
#define VADC_TIMING_CLUSTER_FWD 0
#define VADC_TIMING_CLUSTER_BWD 1
void timing_begin(int slot, cudaStream_t st);
void timing_end(int slot, cudaStream_t st);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; every thread gets the result. `red` >= 32 elements of smem.
template <class T>
__device__ __forceinline__ T block_sum(T v, T* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  T r = (lane < nw) ? red[lane] : T(0);
  r = warp_sum(r);
  return r;
}

// streaming 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}


// packed fp32 arithmetic on register pairs (sm_100: FFMA2 / FMUL2 / FADD2 — one issue slot for two lanes of work)
__device__ __forceinline__ unsigned long long f2_bits(float2 a) { return *reinterpret_cast<unsigned long long*>(&a); }
__device__ __forceinline__ float2 f2_from(unsigned long long a) { return *reinterpret_cast<float2*>(&a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
  return f2_from(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return f2_from(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return f2_from(d);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return f2_from(d);
}
__device__ __forceinline__ float2 bcast2(float a) { return make_float2(a, a); }

}  // namespace vadc
