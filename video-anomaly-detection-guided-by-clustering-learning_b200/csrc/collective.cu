// collective.cu — one-shot all-reduce(sum) of small fp32 messages over NVLink peer memory.
//
// The data-parallel training step of the cluster head exchanges two latency-bound messages per step: the scalar
// sum (D*A)^2 of the cluster loss between forward and backward (full-batch Frobenius norm: model/backbone.py:98 under
// main_predict.py:171's DDP), and [g cluster_center | g gamma | g beta] after the backward (25 KB at K = 32, C = 192).
// Through NCCL each is a separate collective launch plus pack / unpack kernels (round 1: +40 us per step at N = 2,
// +75 us at N = 8 on a 0.62 ms step).  Here ONE kernel per message does everything:
//   1. every rank copies its contribution into its half of a SYMMETRIC block (allocated, zeroed and exchanged once by
//      torch.distributed._symmetric_memory — plumbing; every rank holds the device pointers of all peers; the first 1 KB
//      of a block are its flag words, 2 x capacity floats follow),
//   2. raises a flag in every peer's block (st.release.sys over NVLink) and waits for all peers' flags in its own,
//   3. reads the contribution of EVERY rank with peer loads and adds them in rank order — the same order on every rank,
//      so the result is bit-identical everywhere and run-to-run deterministic — straight into the destination tensors
//      (up to four segments: no concatenated staging tensor, no unpack).
// Flags are generation numbers kept in device memory (the kernel is captured in the step's CUDA graph and replayed), the
// buffer has two halves used alternately: a rank that starts generation g+2 has seen every peer's g+1 flag, which a peer
// raises only after it finished reading generation g.
#include "common.cuh"

namespace vadc {

struct ArSeg { float* ptr; long long n; };
struct ArParams {
  void* const* bufs;          // device array [world]: every rank's symmetric block: 256 uint32 flag words, then 2 * cap floats
  int rank, world;
  long long cap;
  ArSeg seg[4];
  int nseg;
};

constexpr int kGenSlot = 96;  // word of the rank's OWN flag area holding its generation counter (words 0..world-1 are the flags)
constexpr int kPadWords = 256;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

constexpr int kMaxWorld = 16;

__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// dynamic shared memory: world x (padded message) floats — every peer's contribution is fetched with asynchronous
// 16-byte copies, ALL in flight at once (a peer load is a ~2 us round trip over NVLink: loads interleaved with the
// additions that consume them, or limited by the registers that hold them, run one after the other)
__global__ void __launch_bounds__(1024)
oneshot_allreduce_kernel(const ArParams p, const long long padded) {
  extern __shared__ __align__(16) float stage[];
  __shared__ uint32_t gen_s;
  __shared__ void* buf_s[kMaxWorld];
  const int tid = threadIdx.x;
  if (tid < p.world) buf_s[tid] = p.bufs[tid];
  __syncthreads();
  uint32_t* mypad = static_cast<uint32_t*>(buf_s[p.rank]);
  if (tid == 0) gen_s = mypad[kGenSlot] + 1u;
  __syncthreads();
  const uint32_t gen = gen_s;
  const long long half = (long long)(gen & 1u) * p.cap;
  float* mine = static_cast<float*>(buf_s[p.rank]) + kPadWords + half;
  // 1. my contribution -> my block (every segment starts on a 16-byte boundary of the block; the tail of a segment's last
  //    16-byte group is zero-filled so that whole groups can be copied and added)
  long long off = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const float* src = p.seg[s].ptr;
    const long long n = p.seg[s].n, np4 = (n + 3) & ~3ll;
    for (long long i = tid; i < np4; i += blockDim.x) mine[off + i] = i < n ? src[i] : 0.f;
    off += np4;
  }
  __syncthreads();                               // (bar.sync orders the block's writes before the release stores below)
  // 2. flags: mine to every peer (release, system scope: cumulative over the writes above), every peer's to me
  if (tid < p.world) {
    st_release_sys(static_cast<uint32_t*>(buf_s[tid]) + p.rank, gen);
    while ((int32_t)(ld_acquire_sys(mypad + tid) - gen) < 0) {}
  }
  __syncthreads();
  // 3. every rank's message -> shared memory, all copies in flight together
  const long long groups = padded >> 2;
  const uint32_t st32 = static_cast<uint32_t>(__cvta_generic_to_shared(stage));
  for (int q = 0; q < p.world; ++q) {
    const float* peer = static_cast<const float*>(buf_s[q]) + kPadWords + half;
    for (long long g = tid; g < groups; g += blockDim.x)
      cp_async16_cg(st32 + (uint32_t)(((long long)q * padded + g * 4) * 4), peer + g * 4);
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // 4. sum in rank order (the same order on every rank: bit-identical results), straight into the destination tensors
  off = 0;
  for (int s = 0; s < p.nseg; ++s) {
    float* dst = p.seg[s].ptr;
    const long long n = p.seg[s].n;
    for (long long i = tid; i < n; i += blockDim.x) {
      float acc = stage[off + i];
      for (int q = 1; q < p.world; ++q) acc += stage[(long long)q * padded + off + i];
      dst[i] = acc;
    }
    off += (n + 3) & ~3ll;
  }
  if (tid == 0) mypad[kGenSlot] = gen;
}

}  // namespace vadc

using namespace vadc;

extern "C" int vadc_oneshot_allreduce(const void* peer_blocks_dev, int rank, int world,
                                      int64_t capacity, float* t0, int64_t n0, float* t1, int64_t n1, float* t2, int64_t n2,
                                      float* t3, int64_t n3, void* stream) {
  VADC_REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world && capacity > 0, VADC_ERR_BAD_SHAPE);
  VADC_REQUIRE(n0 >= 0 && n1 >= 0 && n2 >= 0 && n3 >= 0 && n0 + n1 + n2 + n3 + 12 <= capacity, VADC_ERR_BAD_SHAPE);
  VADC_REQUIRE(peer_blocks_dev, VADC_ERR_NULL_POINTER);
  VADC_REQUIRE((n0 == 0 || t0) && (n1 == 0 || t1) && (n2 == 0 || t2) && (n3 == 0 || t3), VADC_ERR_NULL_POINTER);
  ArParams p{};
  p.bufs = static_cast<void* const*>(peer_blocks_dev);
  p.rank = rank; p.world = world; p.cap = capacity;
  float* ts[4] = {t0, t1, t2, t3};
  const int64_t ns[4] = {n0, n1, n2, n3};
  for (int i = 0; i < 4; ++i)
    if (ns[i] > 0) { p.seg[p.nseg].ptr = ts[i]; p.seg[p.nseg].n = ns[i]; ++p.nseg; }
  if (p.nseg == 0) return VADC_OK;
  long long padded = 0;
  for (int i = 0; i < p.nseg; ++i) padded += (p.seg[i].n + 3) & ~3ll;
  const size_t smem = (size_t)world * padded * sizeof(float);
  VADC_REQUIRE(smem <= 200u * 1024u, VADC_ERR_UNSUPPORTED);          // 25 KB x 8 ranks at the cfg2 head
  static bool attr_set = false;
  if (!attr_set) {
    VADC_CUDA(cudaFuncSetAttribute(oneshot_allreduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const long long lanes = padded / 4;                // one 16-byte group per thread and pass
  const int threads = lanes >= 1024 ? 1024 : (lanes > 32 ? (int)((lanes + 31) / 32 * 32) : 32);
  oneshot_allreduce_kernel<<<1, threads, smem, static_cast<cudaStream_t>(stream)>>>(p, padded);
  VADC_CHECK_LAUNCH("oneshot_allreduce_kernel");
  return VADC_OK;
}
