// Kernel group 2a: the thin ends of the denoiser conv stacks on CUDA cores (fp32 FMA), plus an
// fp32 CUDA-core version of the hidden 64->64 layer used as the validation / DEQSCI_PREC_FP32 path.
//
//   conv_first : cube z' [B,H,W,T] fp32 (+ sigma)  -> hidden activation planes [B*T,Hc,Wc,64] (fp16 hi/lo)
//                FFDNet: pixel-unshuffle + constant noise-map channel folded into the gather
//                (networks/ffdnet/functions.py:16-53), conv 5->64 + ReLU (networks/ffdnet/models.py:46-51)
//                DnCNN : conv 1->64 + ReLU (networks/provable/model/SimpleCNN_models.py:44-45)
//                Optionally fuses the GAP step in front (z' computed on the fly, written once).
//   conv_mid   : 64->64 + per-channel affine (folded eval BatchNorm) + ReLU, fp32 FMA
//   conv_last  : hidden planes -> cout (4 | 1) -> pixel-shuffle (networks/ffdnet/functions.py:63-81)
//                -> out = z' - noise in the cube layout (solvers/equilibrium_solvers_yaping.py:417,420)
//
// The K = 45 / 9 first layer and N = 4 / 1 last layer are < 1.2 % of the stack's FLOPs and are
// hostile to tensor-core tiles, so they stay on the FMA pipe; the 64->64 layers (98.8 %) run on
// tcgen05 (conv_tc.cu).
#include "common.cuh"

namespace deqsci {

constexpr int TH = 4;    // tile rows   (conv-resolution pixels)
constexpr int TW = 32;   // tile cols   = one warp per row -> conflict-free smem rows
constexpr int kTileThreads = TH * TW;
constexpr int kPixStride = 68;   // floats per pixel in smem for 64-channel tiles (64 + 4 pad: conflict-free LDS.128)

__device__ __forceinline__ void store_split64(__half* hi_p, __half* lo_p, const float* v) {
  // 64 channels of one pixel -> 8 x 16-byte stores per plane
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_f16(v[q * 8 + e], h[e], l[e]);
    reinterpret_cast<uint4*>(hi_p)[q] = *reinterpret_cast<const uint4*>(h);
    reinterpret_cast<uint4*>(lo_p)[q] = *reinterpret_cast<const uint4*>(l);
  }
}

// ------------------------------------------------------------------------------------------------
// first layer
// ------------------------------------------------------------------------------------------------
template <int KIND, bool FUSE_GAP>
__global__ void __launch_bounds__(kTileThreads)
conv_first_kernel(const float* __restrict__ zin, const float* __restrict__ y, const float* __restrict__ phi,
                  const float* __restrict__ phi_sum, float* __restrict__ zprime_out, float sigma,
                  const float* __restrict__ wpack,   // [9][CIN][64]
                  const float* __restrict__ scale, const float* __restrict__ bias, int relu,
                  __half* __restrict__ act_out, long long plane_elems, int B, int H, int W, int T) {
  constexpr bool FFD = (KIND == DEQSCI_NET_FFDNET);
  constexpr int CIN = FFD ? 5 : 1;
  constexpr int SC = FFD ? 2 : 1;                 // fine pixels per conv pixel per axis
  constexpr int FR = SC * TH + 2 * SC;            // fine tile rows incl. halo
  constexpr int FC = SC * TW + 2 * SC;            // fine tile cols incl. halo
  constexpr int ROWLEN = TW + 2;                  // per-parity row length
  extern __shared__ __align__(16) float smem[];
  float* w_s = smem;                              // 9*CIN*64
  float* aff_s = w_s + 9 * CIN * 64;              // scale[64], bias[64]
  float* tile = aff_s + 128;                      // [T][FR][SC][ROWLEN]
  const int Hc = H / SC, Wc = W / SC;
  const int b = blockIdx.z;
  const int ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;
  const int tid = threadIdx.x;

  for (int i = tid; i < 9 * CIN * 64; i += kTileThreads) w_s[i] = wpack[i];
  if (tid < 64) {
    aff_s[tid] = scale ? scale[tid] : 1.f;
    aff_s[64 + tid] = bias ? bias[tid] : 0.f;
  }
  // phase 1: stage z' for the fine tile (all T frames), zero outside the image
  const int fr0 = SC * (ty0 - 1), fc0 = SC * (tx0 - 1);
  for (int p = tid; p < FR * FC; p += kTileThreads) {
    const int fr = p / FC, fc = p - fr * FC;
    const int gr = fr0 + fr, gc = fc0 + fc;
    const bool in = (gr >= 0 && gr < H && gc >= 0 && gc < W);
    const int par = fc % SC, half = fc / SC;
    float* dst = tile + (fr * SC + par) * ROWLEN + half;
    const long long pix = ((long long)b * H + (in ? gr : 0)) * W + (in ? gc : 0);
    float r = 0.f;
    if (FUSE_GAP && in) {
      float s = 0.f;
      for (int t = 0; t < T; ++t) s = __fadd_rn(s, __fmul_rn(zin[pix * T + t], phi[pix * T + t]));
      r = __fdiv_rn(__fsub_rn(y[pix], s), phi_sum[pix]);
    }
    // owned (non-halo) pixels of this tile write z' back once
    const bool owned = in && fr >= SC && fr < FR - SC && fc >= SC && fc < FC - SC;
    for (int t = 0; t < T; ++t) {
      float v = 0.f;
      if (in) {
        v = zin[pix * T + t];
        if (FUSE_GAP) {
          v = __fadd_rn(v, __fmul_rn(r, phi[pix * T + t]));
          if (owned) zprime_out[pix * T + t] = v;
        }
      }
      dst[(long long)t * FR * SC * ROWLEN] = v;
    }
  }
  __syncthreads();

  const int li = tid / TW, lj = tid % TW;
  const int ci = ty0 + li, cj = tx0 + lj;          // conv-resolution pixel of this thread
  const bool active = (ci < Hc && cj < Wc);
  for (int t = 0; t < T; ++t) {
    float acc[64];
#pragma unroll
    for (int o = 0; o < 64; ++o) acc[o] = 0.f;
    const float* tt = tile + (long long)t * FR * SC * ROWLEN;
#pragma unroll 1
    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int tap = dy * 3 + dx;
#pragma unroll
        for (int ch = 0; ch < CIN; ++ch) {
          float v;
          if (FFD && ch == 0) {
            const int ni = ci + dy - 1, nj = cj + dx - 1;   // zero padding also pads the noise map
            v = (ni >= 0 && ni < Hc && nj >= 0 && nj < Wc) ? sigma : 0.f;
          } else {
            const int sub = FFD ? ch - 1 : 0;
            const int rr = FFD ? (sub >> 1) : 0, cc = FFD ? (sub & 1) : 0;
            const int fr = SC * (li + dy) + rr;              // = SC*(li+dy-1) + rr + SC (halo offset)
            v = tt[(fr * SC + cc) * ROWLEN + (lj + dx)];
          }
          const float4* wv = reinterpret_cast<const float4*>(w_s + (tap * CIN + ch) * 64);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float4 w4 = wv[q];
            acc[4 * q + 0] = fmaf(v, w4.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
          }
        }
      }
    }
    if (active) {
#pragma unroll
      for (int o = 0; o < 64; ++o) {
        float v = fmaf(acc[o], aff_s[o], aff_s[64 + o]);
        acc[o] = relu ? fmaxf(v, 0.f) : v;
      }
      const long long nf = (long long)b * T + t;
      const long long off = ((nf * Hc + ci) * Wc + cj) * 64;
      store_split64(act_out + off, act_out + plane_elems + off, acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// hidden layer, fp32 CUDA cores (validation path).  Persistent CTAs: weights are staged once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads)
conv_mid_fp32_kernel(const __half* __restrict__ act_in, __half* __restrict__ act_out, long long plane_elems,
                     const float* __restrict__ wpack,   // [9][64 cin][64 cout]
                     const float* __restrict__ scale, const float* __restrict__ bias, int relu, int NF, int Hc,
                     int Wc) {
  extern __shared__ __align__(16) float smem[];
  float* w_s = smem;                         // 9*64*64
  float* aff_s = w_s + 9 * 64 * 64;          // 128
  float* tile = aff_s + 128;                 // (TH+2)*(TW+2)*kPixStride
  const int tid = threadIdx.x;
  for (int i = tid; i < 9 * 64 * 64 / 4; i += kTileThreads)
    reinterpret_cast<float4*>(w_s)[i] = reinterpret_cast<const float4*>(wpack)[i];
  if (tid < 64) {
    aff_s[tid] = scale ? scale[tid] : 1.f;
    aff_s[64 + tid] = bias ? bias[tid] : 0.f;
  }
  const int tiles_x = (Wc + TW - 1) / TW, tiles_y = (Hc + TH - 1) / TH;
  const long long n_tiles = (long long)NF * tiles_y * tiles_x;
  const int li = tid / TW, lj = tid % TW;
  for (long long tile_id = blockIdx.x; tile_id < n_tiles; tile_id += gridDim.x) {
    const int nf = (int)(tile_id / (tiles_y * tiles_x));
    const int rem = (int)(tile_id - (long long)nf * tiles_y * tiles_x);
    const int ty0 = (rem / tiles_x) * TH, tx0 = (rem % tiles_x) * TW;
    __syncthreads();   // previous tile fully consumed (also covers the weight staging)
    // stage the halo tile as fp32: 8 channels (one 16-byte hi + one 16-byte lo load) per step
    for (int p = tid; p < (TH + 2) * (TW + 2) * 8; p += kTileThreads) {
      const int pix = p >> 3, q = p & 7;
      const int r = pix / (TW + 2), c = pix - r * (TW + 2);
      const int gi = ty0 + r - 1, gj = tx0 + c - 1;
      float v[8];
      if (gi >= 0 && gi < Hc && gj >= 0 && gj < Wc) {
        const long long off = (((long long)nf * Hc + gi) * Wc + gj) * 64 + q * 8;
        const uint4 h4 = __ldg(reinterpret_cast<const uint4*>(act_in + off));
        const uint4 l4 = __ldg(reinterpret_cast<const uint4*>(act_in + plane_elems + off));
        const __half* hh = reinterpret_cast<const __half*>(&h4);
        const __half* ll = reinterpret_cast<const __half*>(&l4);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = join_f16(hh[e], ll[e]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      }
      float4* d = reinterpret_cast<float4*>(tile + pix * kPixStride + q * 8);
      d[0] = make_float4(v[0], v[1], v[2], v[3]);
      d[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    float acc[64];
#pragma unroll
    for (int o = 0; o < 64; ++o) acc[o] = 0.f;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dx = tap - dy * 3;
      const float4* in4 = reinterpret_cast<const float4*>(tile + ((li + dy) * (TW + 2) + lj + dx) * kPixStride);
      const float4* wt = reinterpret_cast<const float4*>(w_s + tap * 64 * 64);
#pragma unroll 2
      for (int c4 = 0; c4 < 16; ++c4) {
        const float4 x4 = in4[c4];
        const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const float4* wv = wt + (c4 * 4 + cc) * 16;
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float4 w4 = wv[q];
            acc[4 * q + 0] = fmaf(xs[cc], w4.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(xs[cc], w4.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(xs[cc], w4.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(xs[cc], w4.w, acc[4 * q + 3]);
          }
        }
      }
    }
    const int ci = ty0 + li, cj = tx0 + lj;
    if (ci < Hc && cj < Wc) {
#pragma unroll
      for (int o = 0; o < 64; ++o) {
        float v = fmaf(acc[o], aff_s[o], aff_s[64 + o]);
        acc[o] = relu ? fmaxf(v, 0.f) : v;
      }
      const long long off = (((long long)nf * Hc + ci) * Wc + cj) * 64;
      store_split64(act_out + off, act_out + plane_elems + off, acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// last layer + pixel shuffle + residual subtract, written back in the cube layout
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(kTileThreads)
conv_last_kernel(const __half* __restrict__ act_in, long long plane_elems,
                 const float* __restrict__ wpack,   // [9][64 cin][COUT]
                 const float* __restrict__ scale, const float* __restrict__ bias, int relu,
                 const float* __restrict__ zprime, float* __restrict__ out, int B, int H, int W, int T) {
  constexpr bool FFD = (KIND == DEQSCI_NET_FFDNET);
  constexpr int COUT = FFD ? 4 : 1;
  constexpr int SC = FFD ? 2 : 1;
  extern __shared__ __align__(16) float smem[];
  float* w_s = smem;                                   // 9*64*COUT
  float* aff_s = w_s + 9 * 64 * COUT;                  // scale[COUT], bias[COUT] (padded to 8)
  float* tile = aff_s + 8;                             // (TH+2)*(TW+2)*kPixStride
  float* outs = tile + (TH + 2) * (TW + 2) * kPixStride;   // [kTileThreads][COUT][T]
  const int Hc = H / SC, Wc = W / SC;
  const int b = blockIdx.z;
  const int ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;
  const int tid = threadIdx.x;
  for (int i = tid; i < 9 * 64 * COUT; i += kTileThreads) w_s[i] = wpack[i];
  if (tid < COUT) {
    aff_s[tid] = scale ? scale[tid] : 1.f;
    aff_s[4 + tid] = bias ? bias[tid] : 0.f;
  }
  const int li = tid / TW, lj = tid % TW;
  for (int t = 0; t < T; ++t) {
    const long long nf = (long long)b * T + t;
    __syncthreads();
    for (int p = tid; p < (TH + 2) * (TW + 2) * 8; p += kTileThreads) {
      const int pix = p >> 3, q = p & 7;
      const int r = pix / (TW + 2), c = pix - r * (TW + 2);
      const int gi = ty0 + r - 1, gj = tx0 + c - 1;
      float v[8];
      if (gi >= 0 && gi < Hc && gj >= 0 && gj < Wc) {
        const long long off = ((nf * Hc + gi) * Wc + gj) * 64 + q * 8;
        const uint4 h4 = __ldg(reinterpret_cast<const uint4*>(act_in + off));
        const uint4 l4 = __ldg(reinterpret_cast<const uint4*>(act_in + plane_elems + off));
        const __half* hh = reinterpret_cast<const __half*>(&h4);
        const __half* ll = reinterpret_cast<const __half*>(&l4);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = join_f16(hh[e], ll[e]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      }
      float4* d = reinterpret_cast<float4*>(tile + pix * kPixStride + q * 8);
      d[0] = make_float4(v[0], v[1], v[2], v[3]);
      d[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    float acc[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dx = tap - dy * 3;
      const float4* in4 = reinterpret_cast<const float4*>(tile + ((li + dy) * (TW + 2) + lj + dx) * kPixStride);
      const float* wt = w_s + tap * 64 * COUT;
#pragma unroll 4
      for (int c4 = 0; c4 < 16; ++c4) {
        const float4 x4 = in4[c4];
        const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          if (COUT == 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wt + (c4 * 4 + cc) * 4);
            acc[0] = fmaf(xs[cc], w4.x, acc[0]);
            acc[1 % COUT] = fmaf(xs[cc], w4.y, acc[1 % COUT]);
            acc[2 % COUT] = fmaf(xs[cc], w4.z, acc[2 % COUT]);
            acc[3 % COUT] = fmaf(xs[cc], w4.w, acc[3 % COUT]);
          } else {
            acc[0] = fmaf(xs[cc], wt[c4 * 4 + cc], acc[0]);
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      float v = fmaf(acc[o], aff_s[o], aff_s[4 + o]);
      if (relu) v = fmaxf(v, 0.f);
      outs[(tid * COUT + o) * T + t] = v;
    }
  }
  __syncthreads();
  // out[b, fine pixel, t] = z'[...] - noise ; the tile's fine pixels: (SC*TH) x (SC*TW) x T floats
  const int n_fine = SC * TH * SC * TW;
  for (int p = tid; p < n_fine * T; p += kTileThreads) {
    const int t = p % T, fp = p / T;
    const int fr = fp / (SC * TW), fc = fp - fr * (SC * TW);
    const int gr = SC * ty0 + fr, gc = SC * tx0 + fc;
    if (gr < H && gc < W) {
      const int pl = (fr / SC) * TW + (fc / SC);                  // owning conv pixel (thread id)
      const int sub = FFD ? ((fr & 1) * 2 + (fc & 1)) : 0;        // channel idx = 2*r + c
      const long long g = (((long long)b * H + gr) * W + gc) * T + t;
      out[g] = __fsub_rn(zprime[g], outs[(pl * COUT + sub) * T + t]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static size_t first_smem_bytes(int kind, int T) {
  const int CIN = kind == DEQSCI_NET_FFDNET ? 5 : 1, SC = kind == DEQSCI_NET_FFDNET ? 2 : 1;
  const int FR = SC * TH + 2 * SC;
  return sizeof(float) * ((size_t)9 * CIN * 64 + 128 + (size_t)T * FR * SC * (TW + 2));
}

int conv_first_launch(int kind, bool fuse_gap, const float* zin, const float* y, const float* phi,
                      const float* phi_sum, float* zprime_out, float sigma, const float* wpack,
                      const float* scale, const float* bias, int relu, __half* act_out, long long plane_elems,
                      int B, int H, int W, int T, cudaStream_t st) {
  const int SC = kind == DEQSCI_NET_FFDNET ? 2 : 1;
  const int Hc = H / SC, Wc = W / SC;
  dim3 grid((Wc + TW - 1) / TW, (Hc + TH - 1) / TH, B);
  const size_t smem = first_smem_bytes(kind, T);
  DEQSCI_CHECK_ARG(smem <= 200 * 1024, "conv_first: T=%d needs %zu bytes of shared memory", T, smem);
  ProfScope prof(PK_CONV_FIRST, st);
#define LAUNCH_FIRST(K, F)                                                                                       \
  do {                                                                                                           \
    DEQSCI_CUDA(cudaFuncSetAttribute(conv_first_kernel<K, F>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                     (int)smem));                                                                \
    conv_first_kernel<K, F><<<grid, kTileThreads, smem, st>>>(zin, y, phi, phi_sum, zprime_out, sigma, wpack,   \
                                                              scale, bias, relu, act_out, plane_elems, B, H, W, \
                                                              T);                                                \
  } while (0)
  if (kind == DEQSCI_NET_FFDNET) {
    if (fuse_gap) LAUNCH_FIRST(DEQSCI_NET_FFDNET, true); else LAUNCH_FIRST(DEQSCI_NET_FFDNET, false);
  } else {
    if (fuse_gap) LAUNCH_FIRST(DEQSCI_NET_DNCNN, true); else LAUNCH_FIRST(DEQSCI_NET_DNCNN, false);
  }
#undef LAUNCH_FIRST
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

int conv_mid_fp32_launch(const __half* act_in, __half* act_out, long long plane_elems, const float* wpack,
                         const float* scale, const float* bias, int relu, int NF, int Hc, int Wc,
                         cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)9 * 64 * 64 + 128 + (size_t)(TH + 2) * (TW + 2) * kPixStride);
  DEQSCI_CUDA(cudaFuncSetAttribute(conv_mid_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long n_tiles = (long long)NF * ((Hc + TH - 1) / TH) * ((Wc + TW - 1) / TW);
  const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  ProfScope prof(PK_CONV_HIDDEN, st);
  conv_mid_fp32_kernel<<<grid, kTileThreads, smem, st>>>(act_in, act_out, plane_elems, wpack, scale, bias, relu,
                                                         NF, Hc, Wc);
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

int conv_last_launch(int kind, const __half* act_in, long long plane_elems, const float* wpack,
                     const float* scale, const float* bias, int relu, const float* zprime, float* out, int B,
                     int H, int W, int T, cudaStream_t st) {
  const int SC = kind == DEQSCI_NET_FFDNET ? 2 : 1, COUT = kind == DEQSCI_NET_FFDNET ? 4 : 1;
  const int Hc = H / SC, Wc = W / SC;
  dim3 grid((Wc + TW - 1) / TW, (Hc + TH - 1) / TH, B);
  const size_t smem = sizeof(float) * ((size_t)9 * 64 * COUT + 8 + (size_t)(TH + 2) * (TW + 2) * kPixStride +
                                       (size_t)kTileThreads * COUT * T);
  DEQSCI_CHECK_ARG(smem <= 200 * 1024, "conv_last: T=%d needs %zu bytes of shared memory", T, smem);
  ProfScope prof(PK_CONV_LAST, st);
  if (kind == DEQSCI_NET_FFDNET) {
    DEQSCI_CUDA(cudaFuncSetAttribute(conv_last_kernel<DEQSCI_NET_FFDNET>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_last_kernel<DEQSCI_NET_FFDNET><<<grid, kTileThreads, smem, st>>>(act_in, plane_elems, wpack, scale, bias,
                                                                          relu, zprime, out, B, H, W, T);
  } else {
    DEQSCI_CUDA(cudaFuncSetAttribute(conv_last_kernel<DEQSCI_NET_DNCNN>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_last_kernel<DEQSCI_NET_DNCNN><<<grid, kTileThreads, smem, st>>>(act_in, plane_elems, wpack, scale, bias,
                                                                         relu, zprime, out, B, H, W, T);
  }
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

}  // namespace deqsci
