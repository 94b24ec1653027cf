// Gradient exchange of the implicit-differentiation training step, fused with the optimizer:
// ONE kernel per rank = cross-GPU barrier + one-shot all-reduce of the flat gradient over NVLink peer
// memory + 1/world scale + Adam update of the rank's own parameter copy.
//
// The reference has no distributed backend; config 5 adds a data-parallel gradient average
// between loss.backward() and optimizer.step() (reference training/sci_equilibrium_training.py:66-75,
// Adam built at video_sci_proxgrad.py:201).  The gradient is 486,080 floats (1.9 MB): latency-bound, so the
// "one-shot" scheme is the right one on NVSwitch -- every rank reads every peer's gradient buffer directly
// (P2P loads through NVLink, each GPU has full bandwidth to each peer), sums in rank order (so all ranks
// compute bit-identical sums and their parameter copies never diverge) and applies Adam while the loads
// stream in.  No staging copy, no separate scale kernel, no optimizer launch set.
//
// Memory: the gradient buffer of each rank is a cudaMalloc allocation exported with CUDA IPC
// (deqsci_comm_alloc / deqsci_comm_open); the same allocation carries the barrier flags.  With world = 1 (or
// when the caller already reduced the gradient, e.g. with NCCL) the kernel is the plain fused Adam step.
#include "common.cuh"

namespace deqsci {

constexpr int kCommMaxWorld = 16;
constexpr int kCommMaxBlocks = 160;
constexpr int kCommFlagWords = 2 * kCommMaxBlocks * kCommMaxWorld;   // [phase][block][peer]
constexpr size_t kCommFlagBytes = kCommFlagWords * sizeof(unsigned) + 256;   // + error word, padded

struct AdamArgs {
  float* p;
  float* m;
  float* v;
  const float* g[kCommMaxWorld];        // gradient buffer of every rank (peer pointers), g[rank] = own
  unsigned* flags[kCommMaxWorld];       // flag block of every rank
  int rank, world;
  long long n4;                         // float4 elements
  float lr, beta1, beta2, eps, bc1, bc2_sqrt, grad_scale;
  unsigned epoch;
  unsigned long long timeout_ns;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// All ranks launch the same grid.  Block b of rank r tells block b of every peer "phase reached" and waits
// for the same word from every peer.  Flags only ever grow (epoch = step counter), so no reset is needed.
__device__ __forceinline__ void cross_gpu_barrier(const AdamArgs& a, int phase) {
  __syncthreads();
  if (threadIdx.x < a.world) {
    const int slot = (phase * kCommMaxBlocks + blockIdx.x) * kCommMaxWorld;
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + slot + a.rank, a.epoch);
    const unsigned* mine = a.flags[a.rank] + slot + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(mine) < a.epoch) {
      if (globaltimer_ns() - t0 > a.timeout_ns) {          // a peer never arrived: flag it, do not hang the GPU
        atomicExch(a.flags[a.rank] + kCommFlagWords, 1u);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512) adam_allreduce_kernel(const AdamArgs a) {
  if (a.world > 1) cross_gpu_barrier(a, 0);      // every rank's backward has finished writing its gradient
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += stride) {
    float4 g = reinterpret_cast<const float4*>(a.g[0])[i];
    for (int r = 1; r < a.world; ++r) {                    // fixed order: identical sums on every rank
      const float4 q = reinterpret_cast<const float4*>(a.g[r])[i];
      g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
    }
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    // torch.optim.Adam (no weight decay, no amsgrad): m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g^2;
    // p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
#define DEQSCI_ADAM1(c)                                                   \
  {                                                                       \
    const float gg = g.c * a.grad_scale;                                  \
    m.c = m.c + (gg - m.c) * (1.0f - a.beta1);                            \
    v.c = v.c * a.beta2 + (gg * gg) * (1.0f - a.beta2);                   \
    const float denom = sqrtf(v.c) / a.bc2_sqrt + a.eps;                  \
    p.c = p.c - (a.lr / a.bc1) * (m.c / denom);                           \
  }
    DEQSCI_ADAM1(x) DEQSCI_ADAM1(y) DEQSCI_ADAM1(z) DEQSCI_ADAM1(w)
#undef DEQSCI_ADAM1
    reinterpret_cast<float4*>(a.p)[i] = p;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
  }
  if (a.world > 1) cross_gpu_barrier(a, 1);      // nobody overwrites a gradient buffer a peer is still reading
}

}  // namespace deqsci

using namespace deqsci;

extern "C" size_t deqsci_comm_bytes(long long n_floats) {
  if (n_floats <= 0) return 0;
  const size_t grad = ((size_t)n_floats * 4 + 255) / 256 * 256;
  return grad + kCommFlagBytes;
}

extern "C" int deqsci_comm_alloc(long long n_floats, void** dev_ptr, void* ipc_handle_64) {
  DEQSCI_CHECK_ARG(n_floats > 0 && dev_ptr && ipc_handle_64, "deqsci_comm_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  const size_t bytes = deqsci_comm_bytes(n_floats);
  DEQSCI_CUDA(cudaMalloc(&p, bytes));
  DEQSCI_CUDA(cudaMemset(p, 0, bytes));
  DEQSCI_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {                         // IPC unavailable (e.g. restricted container): buffer still usable locally
    cudaGetLastError();
    memset(&h, 0, sizeof(h));
  }
  memcpy(ipc_handle_64, &h, 64);
  *dev_ptr = p;
  return DEQSCI_OK;
}

extern "C" int deqsci_comm_open(const void* ipc_handle_64, void** dev_ptr) {
  DEQSCI_CHECK_ARG(ipc_handle_64 && dev_ptr, "deqsci_comm_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64, 64);
  DEQSCI_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return DEQSCI_OK;
}

extern "C" int deqsci_comm_close(void* peer_ptr) {
  DEQSCI_CHECK_ARG(peer_ptr, "deqsci_comm_close: null pointer");
  DEQSCI_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return DEQSCI_OK;
}

extern "C" int deqsci_comm_free(void* dev_ptr) {
  DEQSCI_CHECK_ARG(dev_ptr, "deqsci_comm_free: null pointer");
  DEQSCI_CUDA(cudaFree(dev_ptr));
  return DEQSCI_OK;
}

extern "C" int deqsci_comm_error(const void* comm_base, long long n_floats, int* error_host) {
  DEQSCI_CHECK_ARG(comm_base && error_host && n_floats > 0, "deqsci_comm_error: bad arguments");
  const size_t grad = ((size_t)n_floats * 4 + 255) / 256 * 256;
  unsigned e = 0;
  DEQSCI_CUDA(cudaMemcpy(&e, (const char*)comm_base + grad + kCommFlagWords * sizeof(unsigned), 4, cudaMemcpyDeviceToHost));
  *error_host = (int)e;
  return DEQSCI_OK;
}

extern "C" int deqsci_adam_allreduce_step(float* params, float* exp_avg, float* exp_avg_sq,
                                          const void* const* comm_bases_host, int rank, int world,
                                          long long n_floats, float lr, float beta1, float beta2, float eps,
                                          int step, float grad_scale, unsigned epoch, void* stream) {
  DEQSCI_CHECK_ARG(params && exp_avg && exp_avg_sq && comm_bases_host, "deqsci_adam_allreduce_step: null pointer");
  DEQSCI_CHECK_ARG(world >= 1 && world <= kCommMaxWorld && rank >= 0 && rank < world,
                   "deqsci_adam_allreduce_step: rank %d / world %d unsupported (max %d)", rank, world, kCommMaxWorld);
  DEQSCI_CHECK_ARG(n_floats > 0 && n_floats % 4 == 0, "deqsci_adam_allreduce_step: n=%lld must be a positive multiple of 4",
                   n_floats);
  DEQSCI_CHECK_ARG(step >= 1, "deqsci_adam_allreduce_step: step counts from 1");
  DEQSCI_CHECK_ARG((((uintptr_t)params | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                   "deqsci_adam_allreduce_step: buffers must be 16-byte aligned");
  AdamArgs a{};
  a.p = params;
  a.m = exp_avg;
  a.v = exp_avg_sq;
  const size_t grad = ((size_t)n_floats * 4 + 255) / 256 * 256;
  for (int r = 0; r < world; ++r) {
    DEQSCI_CHECK_ARG(comm_bases_host[r], "deqsci_adam_allreduce_step: comm buffer of rank %d is null", r);
    a.g[r] = reinterpret_cast<const float*>(comm_bases_host[r]);
    a.flags[r] = reinterpret_cast<unsigned*>((char*)comm_bases_host[r] + grad);
  }
  a.rank = rank;
  a.world = world;
  a.n4 = n_floats / 4;
  a.lr = lr;
  a.beta1 = beta1;
  a.beta2 = beta2;
  a.eps = eps;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  a.grad_scale = grad_scale;
  a.epoch = epoch;
  a.timeout_ns = 10ull * 1000 * 1000 * 1000;
  const int threads = 512;
  long long blocks = (a.n4 + threads - 1) / threads;
  const int cap = num_sms() < kCommMaxBlocks ? num_sms() : kCommMaxBlocks;
  if (blocks > cap) blocks = cap;                       // co-resident: every block takes part in the barrier
  adam_allreduce_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(a);
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}
