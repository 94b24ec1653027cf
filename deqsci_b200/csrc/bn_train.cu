// Train-mode BatchNorm2d for the hidden conv layers (nn.BatchNorm2d(64), eps 1e-5, momentum 0.1,
// networks/ffdnet/models.py:58): batch statistics over (frames, H, W) per channel.
//   bn_finalize : adds up the per-CTA partial sums / sums of squares the conv kernel wrote (fp64, fixed order):
//                 mean, biased variance -> scale = gamma * rsqrt(var + eps), shift = beta - mean*scale;
//                 running_mean / running_var momentum update (unbiased variance)
//   bn_apply    : y = relu(x * scale[c] + shift[c]) on the raw conv output planes, in place
//                 (fp16 hi/lo pair -> fp32 -> affine -> re-split); 256 B per pixel read + written.
#include "common.cuh"

namespace deqsci {

__global__ void bn_finalize_kernel(const double* __restrict__ stats, int n_partials, float* __restrict__ scale_shift,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                   float eps, double count, float* __restrict__ record) {
  pdl_launch_dependents();       // launched with programmatic stream serialization: see launch_pdl (common.cuh)
  pdl_wait_predecessor();
  // 1024 threads: 8 row groups x 128 columns of the per-CTA partials; every column is added up in a fixed
  // order (group-strided rows, then the 8 groups), so the statistics are deterministic
  __shared__ double part[8][2 * kHidden];
  __shared__ double tot[2 * kHidden];
  const int t = threadIdx.x & (2 * kHidden - 1), grp = threadIdx.x >> 7;
  double acc = 0.0;
#pragma unroll 4
  for (int k = grp; k < n_partials; k += 8) acc += stats[(size_t)k * 2 * kHidden + t];
  part[grp][t] = acc;
  __syncthreads();
  if (grp != 0) return;
  acc = 0.0;
#pragma unroll
  for (int g = 0; g < 8; ++g) acc += part[g][t];
  tot[t] = acc;
  __syncthreads();
  const int c = t;
  if (c >= kHidden) return;
  const double mean = tot[c] / count;
  double var = tot[kHidden + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float sc = g * rsqrtf((float)var + eps);
  scale_shift[c] = sc;
  scale_shift[kHidden + c] = b - (float)mean * sc;
  if (record) {                      // kept for the backward pass: scale, shift, batch mean, 1/sqrt(var + eps)
    record[c] = sc;
    record[kHidden + c] = b - (float)mean * sc;
    record[2 * kHidden + c] = (float)mean;
    record[3 * kHidden + c] = rsqrtf((float)var + eps);
  }
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// src == act: in place; otherwise the raw conv output planes `src` are kept (backward pass) and `act` receives the result
__global__ void __launch_bounds__(256) bn_apply_kernel(__half* act, const __half* src, long long plane_elems,
                                                       const float* __restrict__ scale_shift, int relu) {
  __shared__ float ss[2 * kHidden];
  pdl_launch_dependents();
  pdl_wait_predecessor();
  if (threadIdx.x < 2 * kHidden) ss[threadIdx.x] = scale_shift[threadIdx.x];
  __syncthreads();
  const long long n_vec = plane_elems / 8;     // 8 channels (16 bytes per plane) per step
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec;
       i += (long long)gridDim.x * blockDim.x) {
    uint4 h4 = *reinterpret_cast<const uint4*>(src + i * 8);
    uint4 l4 = *reinterpret_cast<const uint4*>(src + plane_elems + i * 8);
    __half* hh = reinterpret_cast<__half*>(&h4);
    __half* ll = reinterpret_cast<__half*>(&l4);
    const int c0 = (int)((i * 8) & (kHidden - 1));
    uint32_t* hp = reinterpret_cast<uint32_t*>(&h4);
    uint32_t* lp = reinterpret_cast<uint32_t*>(&l4);
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float v0 = fmaf(join_f16(hh[e], ll[e]), ss[c0 + e], ss[kHidden + c0 + e]);
      float v1 = fmaf(join_f16(hh[e + 1], ll[e + 1]), ss[c0 + e + 1], ss[kHidden + c0 + e + 1]);
      if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
      split_f16x2(v0, v1, hp[e >> 1], lp[e >> 1]);
    }
    *reinterpret_cast<uint4*>(act + i * 8) = h4;
    *reinterpret_cast<uint4*>(act + plane_elems + i * 8) = l4;
  }
}

// stats: device double[n_partials][128], the conv kernel's per-CTA partial sums (rows of CTAs that did not
// run are zero); scale_shift: device float[128] scratch
int bn_train_launch(__half* act, long long plane_elems, const double* stats, int n_partials, float* scale_shift, const float* gamma,
                    const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                    long long count, int relu, cudaStream_t st, const __half* src, float* record) {
  if (!src) src = act;
  DEQSCI_CUDA(launch_pdl(bn_finalize_kernel, 1, 8 * 2 * kHidden, 0, st, stats, n_partials, scale_shift, gamma, beta,
                         running_mean, running_var, momentum, eps, (double)count, record));
  long long blocks = (plane_elems / 8 + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  ProfScope prof(PK_GAP, st);
  DEQSCI_CUDA(launch_pdl(bn_apply_kernel, (unsigned)blocks, 256, 0, st, act, src, plane_elems, (const float*)scale_shift, relu));
  return DEQSCI_OK;
}

// Running-statistics snapshot / restore for the device-resident train-mode driver: an iteration that
// was queued speculatively (the previous one had already converged) must not leave its momentum update
// behind.  One block per conv layer; layers without BatchNorm have null pointers.
struct BnRunningTable {
  float* mean[kMaxBnLayers];
  float* var[kMaxBnLayers];
};

__global__ void bn_running_copy_kernel(BnRunningTable t, float* __restrict__ backup, int restore) {
  const int l = blockIdx.x, c = threadIdx.x;
  float* b = backup + (size_t)l * 2 * kHidden;
  if (t.mean[l]) { if (restore) t.mean[l][c] = b[c]; else b[c] = t.mean[l][c]; }
  if (t.var[l]) { if (restore) t.var[l][c] = b[kHidden + c]; else b[kHidden + c] = t.var[l][c]; }
}

int bn_running_snapshot(const deqsci_bn_params* bn, int n_layers, float* backup, int restore, cudaStream_t st) {
  if (n_layers > kMaxBnLayers) { set_error("train-mode driver: %d layers (max %d)", n_layers, kMaxBnLayers); return DEQSCI_ERR_INVALID; }
  BnRunningTable t;
  for (int i = 0; i < kMaxBnLayers; ++i) {
    t.mean[i] = i < n_layers ? bn[i].running_mean : nullptr;
    t.var[i] = i < n_layers ? bn[i].running_var : nullptr;
  }
  bn_running_copy_kernel<<<n_layers, kHidden, 0, st>>>(t, backup, restore);
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

}  // namespace deqsci
