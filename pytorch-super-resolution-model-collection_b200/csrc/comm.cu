// Gradient all-reduce over NVLink / NVSwitch peer memory, in place, as ONE kernel per rank (no NCCL call on this path).
//
// The flat fp32 gradient buffer of every rank lives in symmetric memory (each rank holds device pointers to all peers' buffers
// and to all peers' signal pads; torch.distributed._symmetric_memory does the allocation + handle exchange, nothing else).
// Two-phase, in place:
//   A  every rank tells every peer "my gradients are complete" (system-scope release store into the peer's signal pad) and
//      waits for the same word from all peers;
//   R  rank r owns slice r of the buffer: it loads that slice from all W buffers over NVLink (W independent 16-byte loads in
//      flight per thread), adds them in rank order 0..W-1 -- the same order on every rank, so all replicas end bit-identical --
//      and stores the sum into slice r of all W buffers (reduce-scatter + all-gather fused, each element crosses every link once
//      in each direction);
//   B  when all blocks of a rank have issued their stores, the rank signals the peers again and waits for theirs: at kernel end
//      every slice of the local buffer holds the global sum.
// Nobody reads a slice while its owner rewrites it (only the owner ever reads slice r), so no staging copy exists.
// Latency-bound for the small nets (ESPCN 149 KB: two NVLink round trips); bandwidth-bound for EDSR-256 (172 MB).
// Replaces the tail ncclAllReduce whose ~130 us at 8 ranks capped the round-1 scaling curve at 0.83.
#include "srb_common.cuh"

namespace srb {

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float4 *p) {  // bypasses the (non-coherent) L1
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer1(const float *p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

constexpr int kMaxWorld = 16;

// wait until pad[slot] reaches `want` (values only grow); a peer that never arrives traps after ~10 s instead of hanging the GPU
__device__ __forceinline__ void wait_flag(const uint32_t *slot, uint32_t want) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(slot) - want) < 0) {
    if (clock64() - t0 > 20000000000LL) __trap();
  }
}

struct ArArgs {
  float *const *bufs;      // [W] device pointers to every rank's flat buffer (symmetric memory)
  uint32_t *const *pads;   // [W] device pointers to every rank's signal pad (>= 2 * kMaxWorld words each)
  int rank, world;
  long long n;             // floats
  uint32_t *epoch;         // local: number of completed all-reduces
  uint32_t *arrived;       // local: blocks of this launch that have issued their stores
};

template <int WM>  // WM = world size rounded up to 2 / 4 / 8 / 16 (register arrays)
__global__ void __launch_bounds__(512) k_allreduce_inplace(ArArgs a) {
  __shared__ int is_last;
  const int W = a.world;
  const uint32_t e = *(volatile uint32_t *)a.epoch + 1u;
  // ---- A: my gradients are complete (all earlier kernels of the stream have finished); tell every peer, wait for every peer
  if (blockIdx.x == 0 && (int)threadIdx.x < W) {
    __threadfence_system();
    st_release_sys(a.pads[threadIdx.x] + a.rank, e);
  }
  if ((int)threadIdx.x < W) wait_flag(a.pads[a.rank] + threadIdx.x, e);
  __syncthreads();
  // ---- R: reduce my slice from all buffers, store the sum into all buffers
  const long long n4 = a.n >> 2;
  const long long per = (n4 + W - 1) / W;
  const long long lo = (long long)a.rank * per, hi = lo + per < n4 ? lo + per : n4;
  float4 *ptr[WM];
#pragma unroll
  for (int r = 0; r < WM; ++r) ptr[r] = r < W ? (float4 *)a.bufs[r] : nullptr;
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
    float4 v[WM];
#pragma unroll
    for (int r = 0; r < WM; ++r)
      if (r < W) v[r] = ld_peer4(ptr[r] + i);
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < WM; ++r)
      if (r < W) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
#pragma unroll
    for (int r = 0; r < WM; ++r)
      if (r < W) ptr[r][i] = s;
  }
  if (a.rank == W - 1 && blockIdx.x == 0) {  // the n % 4 tail
    for (long long i = (n4 << 2) + threadIdx.x; i < a.n; i += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < W; ++r) s += ld_peer1(a.bufs[r] + i);
      for (int r = 0; r < W; ++r) a.bufs[r][i] = s;
    }
  }
  // ---- B: all my stores are issued -> signal; the last block of this rank waits until every peer has stored into my buffer
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(a.arrived, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (is_last) {
    __threadfence_system();
    if ((int)threadIdx.x < W) {
      st_release_sys(a.pads[threadIdx.x] + kMaxWorld + a.rank, e);
      wait_flag(a.pads[a.rank] + kMaxWorld + threadIdx.x, e);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      *a.arrived = 0u;
      *(volatile uint32_t *)a.epoch = e;
      __threadfence();
    }
  }
}

}  // namespace

}  // namespace srb

using namespace srb;

extern "C" int srb_allreduce_inplace(float *const *bufs_dev, uint32_t *const *pads_dev, int32_t rank, int32_t world, int64_t n,
                                     uint32_t *state2, void *stream) {
  SRB_REQUIRE(bufs_dev && pads_dev && state2 && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world && n >= 0, SRB_EINVAL,
              "bad all-reduce arguments (world <= 16)");
  if (n == 0 || world == 1) return SRB_OK;
  ArArgs a;
  a.bufs = bufs_dev;
  a.pads = pads_dev;
  a.rank = rank;
  a.world = world;
  a.n = n;
  a.epoch = state2;
  a.arrived = state2 + 1;
  // every block must be resident at once (they spin on flags): at most one block per SM, far fewer for the small nets
  const long long per4 = ((n >> 2) + world - 1) / world;
  long long blocks = (per4 + 511) / 512;
  if (blocks > 96) blocks = 96;
  if (blocks < 1) blocks = 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (world <= 2) k_allreduce_inplace<2><<<(unsigned)blocks, 512, 0, st>>>(a);
  else if (world <= 4) k_allreduce_inplace<4><<<(unsigned)blocks, 512, 0, st>>>(a);
  else if (world <= 8) k_allreduce_inplace<8><<<(unsigned)blocks, 512, 0, st>>>(a);
  else k_allreduce_inplace<16><<<(unsigned)blocks, 512, 0, st>>>(a);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}
