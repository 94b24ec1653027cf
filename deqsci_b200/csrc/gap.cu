// Kernel group 1: SCI operator and fused GAP data-consistency step (HBM-bound, fp32).
//
// Layout (reference): cube [B,H,W,T] with T innermost, snapshot [B,H,W].  For T == 8 a pixel's
// 8 frames are 32 contiguous bytes: two adjacent lanes take one float4 each, so every warp-wide
// load/store is a fully coalesced 512-byte access, and the 8-term sum over T finishes with a
// single shuffle.  Any other T uses one thread per pixel.
//
// Algorithmic bytes per pixel (T=8, fp32): read z 32 + Phi 32 + y 4 + phi_sum 4, write 32 = 104.
// Multiplications and additions are kept as separate roundings (no FMA contraction) to follow the
// reference's `x*Phi` -> sum, `y - fb`, `/ Phi_sum`, `y[...,None]*Phi`, `z + ...` sequence
// (utils/cg_utils.py:90,129; solvers/equilibrium_solvers_yaping.py:399-400).
#include "common.cuh"

namespace deqsci {

enum GapOp { OP_FORWARD = 0, OP_ADJOINT = 1, OP_STEP = 2, OP_VJP = 3, OP_PHISUM = 4 };

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float dot4_nofma(float4 a, float4 b) {
  float s = __fmul_rn(a.x, b.x);
  s = __fadd_rn(s, __fmul_rn(a.y, b.y));
  s = __fadd_rn(s, __fmul_rn(a.z, b.z));
  s = __fadd_rn(s, __fmul_rn(a.w, b.w));
  return s;
}

// T == 8: item = half pixel (4 frames); lanes 2k and 2k+1 share a pixel.
template <int OP>
__global__ void __launch_bounds__(256) gap_t8_kernel(const float* __restrict__ a,      // z / x / v (cube) or y for ADJOINT
                                                     const float* __restrict__ y,      // snapshot (STEP)
                                                     const float* __restrict__ phi,
                                                     const float* __restrict__ phi_sum,
                                                     const float* __restrict__ add,    // VJP optional addend
                                                     float* __restrict__ out, long long n_items) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  // Warp-uniform trip count: every lane of a warp runs the same number of iterations so the
  // full-mask shuffles are always convergent; lanes past the end recompute the last pair (n_items
  // is even, so a lane and its partner i^1 are clamped together) and store nothing.
  const long long first = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  for (long long base = first; base < n_items; base += stride) {
    long long i = base + (threadIdx.x & 31);
    const bool valid = i < n_items;          // clamped lanes compute but never store
    if (!valid) i = n_items - 2 + (i & 1);
    const long long pix = i >> 1;
    const float4 p = ldg4(phi + i * 4);
    if (OP == OP_ADJOINT) {
      const float yy = __ldg(a + pix);
      float4 o;
      o.x = __fmul_rn(yy, p.x); o.y = __fmul_rn(yy, p.y); o.z = __fmul_rn(yy, p.z); o.w = __fmul_rn(yy, p.w);
      if (valid) *reinterpret_cast<float4*>(out + i * 4) = o;
      continue;
    }
    if (OP == OP_PHISUM) {
      float s = __fadd_rn(__fadd_rn(p.x, p.y), __fadd_rn(p.z, p.w));
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
      if (valid && (i & 1) == 0) out[pix] = (s == 0.0f) ? 1.0f : s;
      continue;
    }
    const float4 zv = ldg4(a + i * 4);
    float s = dot4_nofma(zv, p);
    const float other = __shfl_xor_sync(0xffffffffu, s, 1);
    // same association on both lanes: (frames 0-3) + (frames 4-7)
    s = (i & 1) ? __fadd_rn(other, s) : __fadd_rn(s, other);
    if (OP == OP_FORWARD) {
      if (valid && (i & 1) == 0) out[pix] = s;
      continue;
    }
    float r;
    if (OP == OP_STEP) r = __fdiv_rn(__fsub_rn(__ldg(y + pix), s), __ldg(phi_sum + pix));
    else               r = -__fdiv_rn(s, __ldg(phi_sum + pix));
    float4 o;
    o.x = __fadd_rn(zv.x, __fmul_rn(r, p.x));
    o.y = __fadd_rn(zv.y, __fmul_rn(r, p.y));
    o.z = __fadd_rn(zv.z, __fmul_rn(r, p.z));
    o.w = __fadd_rn(zv.w, __fmul_rn(r, p.w));
    if (OP == OP_VJP && add != nullptr) {
      const float4 ad = ldg4(add + i * 4);
      o.x = __fadd_rn(o.x, ad.x); o.y = __fadd_rn(o.y, ad.y); o.z = __fadd_rn(o.z, ad.z); o.w = __fadd_rn(o.w, ad.w);
    }
    if (valid) *reinterpret_cast<float4*>(out + i * 4) = o;
  }
}

// any T: one thread per pixel
template <int OP>
__global__ void __launch_bounds__(256) gap_generic_kernel(const float* __restrict__ a, const float* __restrict__ y,
                                                          const float* __restrict__ phi,
                                                          const float* __restrict__ phi_sum,
                                                          const float* __restrict__ add, float* __restrict__ out,
                                                          long long n_pix, int T) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < n_pix; pix += stride) {
    const float* pp = phi + pix * T;
    if (OP == OP_ADJOINT) {
      const float yy = a[pix];
      for (int t = 0; t < T; ++t) out[pix * T + t] = __fmul_rn(yy, pp[t]);
      continue;
    }
    if (OP == OP_PHISUM) {
      float s = 0.f;
      for (int t = 0; t < T; ++t) s = __fadd_rn(s, pp[t]);
      out[pix] = (s == 0.0f) ? 1.0f : s;
      continue;
    }
    const float* zp = a + pix * T;
    float s = 0.f;
    for (int t = 0; t < T; ++t) s = __fadd_rn(s, __fmul_rn(zp[t], pp[t]));
    if (OP == OP_FORWARD) { out[pix] = s; continue; }
    float r;
    if (OP == OP_STEP) r = __fdiv_rn(__fsub_rn(y[pix], s), phi_sum[pix]);
    else               r = -__fdiv_rn(s, phi_sum[pix]);
    for (int t = 0; t < T; ++t) {
      float o = __fadd_rn(zp[t], __fmul_rn(r, pp[t]));
      if (OP == OP_VJP && add != nullptr) o = __fadd_rn(o, add[pix * T + t]);
      out[pix * T + t] = o;
    }
  }
}

template <int OP>
static int launch_gap(const float* a, const float* y, const float* phi, const float* phi_sum, const float* add,
                      float* out, int B, int H, int W, int T, void* stream) {
  DEQSCI_CHECK_ARG(B > 0 && H > 0 && W > 0 && T > 0, "gap: non-positive dimension B=%d H=%d W=%d T=%d", B, H, W, T);
  DEQSCI_CHECK_ARG(phi != nullptr && out != nullptr, "gap: null pointer");
  if (OP != OP_PHISUM) DEQSCI_CHECK_ARG(a != nullptr, "gap: null input");
  if (OP == OP_STEP) DEQSCI_CHECK_ARG(y != nullptr && phi_sum != nullptr, "gap_step: null y / phi_sum");
  if (OP == OP_VJP) DEQSCI_CHECK_ARG(phi_sum != nullptr, "gap_vjp: null phi_sum");
  const long long n_pix = (long long)B * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = 256;
  // grid: a multiple of the SM count, 8 resident CTAs of 256 threads per SM, 2-4 items per thread
  const long long max_blocks = (long long)num_sms() * 8 * 4;
  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool fast = (T == 8) && aligned16(phi) && aligned16(out) && (OP == OP_PHISUM || aligned16(a)) &&
                    (add == nullptr || aligned16(add));
  ProfScope prof(PK_GAP, st);
  if (fast) {
    const long long n_items = n_pix * 2;
    long long blocks = (n_items + threads - 1) / threads;
    if (blocks > max_blocks) blocks = max_blocks;
    gap_t8_kernel<OP><<<(unsigned)blocks, threads, 0, st>>>(a, y, phi, phi_sum, add, out, n_items);
  } else {
    long long blocks = (n_pix + threads - 1) / threads;
    if (blocks > max_blocks) blocks = max_blocks;
    gap_generic_kernel<OP><<<(unsigned)blocks, threads, 0, st>>>(a, y, phi, phi_sum, add, out, n_pix, T);
  }
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

int gap_step_launch(const float* z, const float* y, const float* phi, const float* phi_sum, float* out, int B,
                    int H, int W, int T, void* stream) {
  return launch_gap<OP_STEP>(z, y, phi, phi_sum, nullptr, out, B, H, W, T, stream);
}

}  // namespace deqsci

using namespace deqsci;

extern "C" int deqsci_gap_forward(const float* x, const float* phi, float* out, int B, int H, int W, int T,
                                  void* stream) {
  return launch_gap<OP_FORWARD>(x, nullptr, phi, nullptr, nullptr, out, B, H, W, T, stream);
}
extern "C" int deqsci_gap_adjoint(const float* y, const float* phi, float* out, int B, int H, int W, int T,
                                  void* stream) {
  return launch_gap<OP_ADJOINT>(y, nullptr, phi, nullptr, nullptr, out, B, H, W, T, stream);
}
extern "C" int deqsci_phi_sum(const float* phi, float* out, int B, int H, int W, int T, void* stream) {
  return launch_gap<OP_PHISUM>(nullptr, nullptr, phi, nullptr, nullptr, out, B, H, W, T, stream);
}
extern "C" int deqsci_gap_step(const float* z, const float* y, const float* phi, const float* phi_sum, float* out,
                               int B, int H, int W, int T, void* stream) {
  return launch_gap<OP_STEP>(z, y, phi, phi_sum, nullptr, out, B, H, W, T, stream);
}
extern "C" int deqsci_gap_vjp(const float* v, const float* phi, const float* phi_sum, const float* add, float* out,
                              int B, int H, int W, int T, void* stream) {
  return launch_gap<OP_VJP>(v, nullptr, phi, phi_sum, add, out, B, H, W, T, stream);
}

// ------------------------------------------------------------------------------------------------
// Prologue of the tensor-core first conv layer: (optional) GAP step, then the frame-major,
// channels-last K-packed input plane that conv_tc_first.cu loads with TMA.
//   FFDNet: one thread per half-resolution pixel (b,i,j): z' of its 2x2 fine pixels for all T frames,
//           plane row [b*T+t, i, j, :] = {sigma, z'(0,0), z'(0,1), z'(1,0), z'(1,1), 0 x 11}
//           (pixel-unshuffle + noise map of networks/ffdnet/functions.py:16-53; because sigma is a real
//           channel, the conv's zero padding of the noise map is just TMA out-of-bounds fill)
//   DnCNN : one thread per pixel, plane row = {z', 0 x 15}
// Values are stored as fp16 pairs (hi plane, lo plane), like every activation of the stack.
// Bytes per measurement (256x256x8): read z, Phi 4 MiB (+y, phi_sum), write z' 2 MiB + planes 8 MiB.
// ------------------------------------------------------------------------------------------------
namespace deqsci {

template <int KIND>
__global__ void __launch_bounds__(128) gap_prep_kernel(const float* __restrict__ z, const float* __restrict__ y,
                                                       const float* __restrict__ phi,
                                                       const float* __restrict__ phi_sum,
                                                       float* __restrict__ zprime_out, __half* __restrict__ planes,
                                                       long long plane_elems, float sigma, int B, int H, int W,
                                                       int T, int do_gap) {
  pdl_launch_dependents();       // launched with programmatic stream serialization: see launch_pdl (common.cuh)
  pdl_wait_predecessor();
  constexpr int SC = (KIND == DEQSCI_NET_FFDNET) ? 2 : 1;
  constexpr int NSUB = SC * SC;
  const int Hc = H / SC, Wc = W / SC;
  const long long n = (long long)B * Hc * Wc;
  __half s_hi, s_lo;
  split_f16(sigma, s_hi, s_lo);
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < n;
       id += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(id % Wc);
    const int i = (int)((id / Wc) % Hc);
    const int b = (int)(id / ((long long)Wc * Hc));
    float r[NSUB];
    long long pix[NSUB];
#pragma unroll
    for (int s = 0; s < NSUB; ++s) {
      pix[s] = ((long long)b * H + SC * i + s / SC) * W + SC * j + s % SC;
      r[s] = 0.f;
      if (do_gap) {
        float acc = 0.f;
        for (int t = 0; t < T; ++t) acc = __fadd_rn(acc, __fmul_rn(z[pix[s] * T + t], phi[pix[s] * T + t]));
        r[s] = __fdiv_rn(__fsub_rn(y[pix[s]], acc), phi_sum[pix[s]]);
      }
    }
    for (int t = 0; t < T; ++t) {
      __align__(16) __half hi[8];
      __align__(16) __half lo[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) { hi[c] = __float2half_rn(0.f); lo[c] = hi[c]; }
      if (KIND == DEQSCI_NET_FFDNET) { hi[0] = s_hi; lo[0] = s_lo; }
#pragma unroll
      for (int s = 0; s < NSUB; ++s) {
        float v = z[pix[s] * T + t];
        if (do_gap) {
          v = __fadd_rn(v, __fmul_rn(r[s], phi[pix[s] * T + t]));
          if (do_gap == 2)               // planar z' [B,T,H,W] for the tensor-core last layer (coalesced there)
            zprime_out[(((long long)b * T + t) * H + SC * i + s / SC) * W + SC * j + s % SC] = v;
          else
            zprime_out[pix[s] * T + t] = v;
        }
        const int c = (KIND == DEQSCI_NET_FFDNET) ? 1 + s : 0;
        split_f16(v, hi[c], lo[c]);
      }
      // one K = 16 row per pixel: [hi x C | hi x C | lo' x C | 0], C = 5 (FFDNet) or 1 (DnCNN) -- see conv_tc_first.cu
      constexpr int C = (KIND == DEQSCI_NET_FFDNET) ? 5 : 1;
      __align__(16) __half krow[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) krow[c] = __float2half_rn(0.f);
#pragma unroll
      for (int c = 0; c < C; ++c) { krow[c] = hi[c]; krow[C + c] = hi[c]; krow[2 * C + c] = lo[c]; }
      const long long row = ((((long long)b * T + t) * Hc + i) * Wc + j) * kPrepChannels;
      reinterpret_cast<uint4*>(planes + row)[0] = reinterpret_cast<const uint4*>(krow)[0];
      reinterpret_cast<uint4*>(planes + row)[1] = reinterpret_cast<const uint4*>(krow)[1];
    }
  }
}

// T == 8 fast path: every load is a float4 issued up front (a thread's fine pixels are 2 x 64
// contiguous bytes per array and row for FFDNet), every store a 16-byte vector.
template <int KIND>
__global__ void __launch_bounds__(128) gap_prep_t8_kernel(const float* __restrict__ z, const float* __restrict__ y,
                                                          const float* __restrict__ phi,
                                                          const float* __restrict__ phi_sum,
                                                          float* __restrict__ zprime_out, __half* __restrict__ planes,
                                                          long long plane_elems, float sigma, int B, int H, int W,
                                                          int do_gap) {
  pdl_launch_dependents();
  pdl_wait_predecessor();
  constexpr int SC = (KIND == DEQSCI_NET_FFDNET) ? 2 : 1;
  constexpr int NSUB = SC * SC;
  constexpr int T = 8;
  const int Hc = H / SC, Wc = W / SC;
  const long long n = (long long)B * Hc * Wc;
  __half s_hi, s_lo;
  split_f16(sigma, s_hi, s_lo);
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < n;
       id += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(id % Wc);
    const int i = (int)((id / Wc) % Hc);
    const int b = (int)(id / ((long long)Wc * Hc));
    float zv[NSUB][T];
    long long pix[NSUB];
#pragma unroll
    for (int s = 0; s < NSUB; ++s) {
      pix[s] = ((long long)b * H + SC * i + s / SC) * W + SC * j + s % SC;
      const float4 a0 = ldg4(z + pix[s] * T), a1 = ldg4(z + pix[s] * T + 4);
      zv[s][0] = a0.x; zv[s][1] = a0.y; zv[s][2] = a0.z; zv[s][3] = a0.w;
      zv[s][4] = a1.x; zv[s][5] = a1.y; zv[s][6] = a1.z; zv[s][7] = a1.w;
    }
    if (do_gap) {
#pragma unroll
      for (int s = 0; s < NSUB; ++s) {
        const float4 p0 = ldg4(phi + pix[s] * T), p1 = ldg4(phi + pix[s] * T + 4);
        const float pv[T] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        // same association as gap_t8_kernel: (frames 0-3) + (frames 4-7)
        float lo4 = __fmul_rn(zv[s][0], pv[0]), hi4 = __fmul_rn(zv[s][4], pv[4]);
#pragma unroll
        for (int t = 1; t < 4; ++t) {
          lo4 = __fadd_rn(lo4, __fmul_rn(zv[s][t], pv[t]));
          hi4 = __fadd_rn(hi4, __fmul_rn(zv[s][4 + t], pv[4 + t]));
        }
        const float r = __fdiv_rn(__fsub_rn(__ldg(y + pix[s]), __fadd_rn(lo4, hi4)), __ldg(phi_sum + pix[s]));
#pragma unroll
        for (int t = 0; t < T; ++t) zv[s][t] = __fadd_rn(zv[s][t], __fmul_rn(r, pv[t]));
        if (do_gap == 1) {
          float4* zo = reinterpret_cast<float4*>(zprime_out + pix[s] * T);
          zo[0] = make_float4(zv[s][0], zv[s][1], zv[s][2], zv[s][3]);
          zo[1] = make_float4(zv[s][4], zv[s][5], zv[s][6], zv[s][7]);
        }
      }
      if (do_gap == 2) {
        // planar z' [B,T,H,W]: what the tensor-core last layer reads back per frame -- there one lane owns one
        // pixel of ONE frame, so the frame-innermost cube layout costs it a 32-byte sector per float
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
          for (int dy = 0; dy < SC; ++dy) {
            float* dst = zprime_out + (((long long)b * T + t) * H + SC * i + dy) * W + SC * j;
            if (SC == 2) *reinterpret_cast<float2*>(dst) = make_float2(zv[dy * 2][t], zv[dy * 2 + 1][t]);
            else         *dst = zv[0][t];
          }
      }
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
      __align__(16) __half hi[8];
      __align__(16) __half lo[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) { hi[c] = __float2half_rn(0.f); lo[c] = hi[c]; }
      if (KIND == DEQSCI_NET_FFDNET) { hi[0] = s_hi; lo[0] = s_lo; }
#pragma unroll
      for (int s = 0; s < NSUB; ++s) {
        const int c = (KIND == DEQSCI_NET_FFDNET) ? 1 + s : 0;
        split_f16(zv[s][t], hi[c], lo[c]);
      }
      // one K = 16 row per pixel: [hi x C | hi x C | lo' x C | 0], C = 5 (FFDNet) or 1 (DnCNN) -- see conv_tc_first.cu
      constexpr int C = (KIND == DEQSCI_NET_FFDNET) ? 5 : 1;
      __align__(16) __half krow[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) krow[c] = __float2half_rn(0.f);
#pragma unroll
      for (int c = 0; c < C; ++c) { krow[c] = hi[c]; krow[C + c] = hi[c]; krow[2 * C + c] = lo[c]; }
      const long long row = ((((long long)b * T + t) * Hc + i) * Wc + j) * kPrepChannels;
      reinterpret_cast<uint4*>(planes + row)[0] = reinterpret_cast<const uint4*>(krow)[0];
      reinterpret_cast<uint4*>(planes + row)[1] = reinterpret_cast<const uint4*>(krow)[1];
    }
  }
}

// planes: [B*T, Hc, Wc, 16] fp16 (kPrepChannels): one K-packed row per pixel; plane_elems = B*T*Hc*Wc*16.  y/phi/phi_sum/zprime_out may be null
// when do_gap == 0 (the planes are then built from z itself).
int gap_prep_launch(int kind, const float* z, const float* y, const float* phi, const float* phi_sum,
                    float* zprime_out, __half* planes, long long plane_elems, float sigma, int B, int H, int W,
                    int T, bool do_gap_flag, bool zprime_planar, cudaStream_t st) {
  const int do_gap = do_gap_flag ? (zprime_planar ? 2 : 1) : 0;
  const int SC = kind == DEQSCI_NET_FFDNET ? 2 : 1;
  const long long n = (long long)B * (H / SC) * (W / SC);
  long long blocks = (n + 127) / 128;
  const long long cap = (long long)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  ProfScope prof(PK_GAP, st);
  auto a16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (T == 8 && a16(z) && a16(phi) && a16(zprime_out)) {
    if (kind == DEQSCI_NET_FFDNET)
      DEQSCI_CUDA(launch_pdl(gap_prep_t8_kernel<DEQSCI_NET_FFDNET>, (unsigned)blocks, 128, 0, st, z, y, phi, phi_sum,
                             zprime_out, planes, plane_elems, sigma, B, H, W, do_gap));
    else
      DEQSCI_CUDA(launch_pdl(gap_prep_t8_kernel<DEQSCI_NET_DNCNN>, (unsigned)blocks, 128, 0, st, z, y, phi, phi_sum,
                             zprime_out, planes, plane_elems, sigma, B, H, W, do_gap));
    DEQSCI_LAUNCH_CHECK();
    return DEQSCI_OK;
  }
  if (kind == DEQSCI_NET_FFDNET)
    DEQSCI_CUDA(launch_pdl(gap_prep_kernel<DEQSCI_NET_FFDNET>, (unsigned)blocks, 128, 0, st, z, y, phi, phi_sum, zprime_out,
                           planes, plane_elems, sigma, B, H, W, T, do_gap));
  else
    DEQSCI_CUDA(launch_pdl(gap_prep_kernel<DEQSCI_NET_DNCNN>, (unsigned)blocks, 128, 0, st, z, y, phi, phi_sum, zprime_out,
                           planes, plane_elems, sigma, B, H, W, T, do_gap));
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

}  // namespace deqsci
