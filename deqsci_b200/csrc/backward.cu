// Weight gradients of ONE iterate-map call: the backward pass of the graph-attached call of DEQFixedPoint.forward
// (reference solvers/new_equilibrium_utils_yaping.py:268: z = f(z*) with the autograd tape; loss.backward() then runs
// cuDNN dgrad / wgrad / BatchNorm-backward kernels through it).  Here the same pass on this library's kernels:
//
//   layer L-1 (last conv):   dW = corr(a_{L-2}, gn),  gn = -(pixel-unshuffled upstream gradient)
//                            da_{L-2} = conv(gn, W^T flipped)       first-layer tensor-core kernel on the adjoint plan
//   layers L-2 .. 1:         dy  = da_i * (a_i > 0)                              ReLU
//                            dc  = gamma*invstd * (dy - mean(dy) - xhat*mean(dy*xhat))   train-mode BatchNorm (or dc = dy)
//                            dgamma = sum(dy*xhat), dbeta = sum(dy)
//                            dW_i = corr(a_{i-1}, dc)                            wgrad_hidden_kernel (fp32 CUDA cores)
//                            da_{i-1} = conv(dc, W_i^T flipped)                  CTA-pair tcgen05 kernel, adjoint plan
//   layer 0:                 dy = da_0 * (a_0 > 0);  dW_0 = corr(input, dy)
//
// Gradient planes use the activations' storage (fp16 hi + lo*2^11 pairs), so every layer's dc is multiplied by a
// power of two chosen on the device from max|dy| (the running product is carried in `cum` and divided out of every
// result): loss gradients are ~1e-7 per element, fp16's normal range starts at 6e-5.
// Reductions are two-stage with a fixed order (per-CTA partials, then one finalize kernel): deterministic.
#include "common.cuh"
#include <algorithm>

namespace deqsci {

constexpr int kRecFloats = 4 * kHidden;        // per BatchNorm layer: scale, shift, mean, invstd
constexpr int kRedCtas = 296;                  // partial rows of the per-channel reductions
constexpr int kWgTile = 32;                    // wgrad tile: kWgRows x 32 pixels
constexpr int kWgRows = 4;
constexpr int kWgMaxCtas = 148;
constexpr int kThinCtas = 592;                 // thin-layer wgrads: latency-bound streaming loops, 4 CTAs per SM

// scal[0] = cumulative scale of the gradient planes currently in flight, scal[1] = this layer's factor
struct ActBwdParams {
  const __half* g;            // da_i planes (hi, lo)
  const __half* act;          // a_i planes: ReLU mask source (hi plane)
  const __half* pre;          // c_i planes (conv output before BatchNorm) or nullptr
  const float* rec;           // BatchNorm record (scale, shift, mean, invstd) or nullptr
  const float* gamma;         // or nullptr (= 1)
  __half* out;                // dc planes
  long long plane_elems;
  double count;
};

__device__ __forceinline__ float load_pair(const __half* planes, long long plane_elems, long long i) {
  return join_f16(planes[i], planes[plane_elems + i]);
}

// per channel: sum dy, sum dy*xhat, max |dy|   -> partial[blockIdx][3][64]
__global__ void __launch_bounds__(256) act_bwd_reduce_kernel(const ActBwdParams p, float* __restrict__ partial) {
  __shared__ float s_sum[4][kHidden], s_sx[4][kHidden], s_mx[4][kHidden];
  const int c = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const long long n_px = p.plane_elems / kHidden;
  float mean = 0.f, invstd = 1.f;
  if (p.rec) { mean = p.rec[2 * kHidden + c]; invstd = p.rec[3 * kHidden + c]; }
  float sum = 0.f, sx = 0.f, mx = 0.f;
  for (long long px = (long long)blockIdx.x * 4 + sub; px < n_px; px += (long long)gridDim.x * 4) {
    const long long i = px * kHidden + c;
    const float a = __half2float(p.act[i]);
    float dy = load_pair(p.g, p.plane_elems, i);
    dy = a > 0.f ? dy : 0.f;
    sum += dy;
    mx = fmaxf(mx, fabsf(dy));
    if (p.pre) sx = fmaf(dy, (load_pair(p.pre, p.plane_elems, i) - mean) * invstd, sx);
  }
  s_sum[sub][c] = sum; s_sx[sub][c] = sx; s_mx[sub][c] = mx;
  __syncthreads();
  if (sub == 0) {
    float* o = partial + (size_t)blockIdx.x * 3 * kHidden;
    o[c] = (s_sum[0][c] + s_sum[1][c]) + (s_sum[2][c] + s_sum[3][c]);
    o[kHidden + c] = (s_sx[0][c] + s_sx[1][c]) + (s_sx[2][c] + s_sx[3][c]);
    o[2 * kHidden + c] = fmaxf(fmaxf(s_mx[0][c], s_mx[1][c]), fmaxf(s_mx[2][c], s_mx[3][c]));
  }
}

// one block of 1024 threads (16 row groups x 64 channels): totals in fp64 in a fixed order, BatchNorm parameter
// gradients, the layer's scale factor
// coef[0..63] = per-channel multiplier, coef[64..127] = m1, coef[128..191] = m2 (means of dy and dy*xhat)
__global__ void __launch_bounds__(1024) act_bwd_finalize_kernel(const float* __restrict__ partial, int n_partials,
                                                                 const float* __restrict__ rec,
                                                                 const float* __restrict__ gamma, double count,
                                                                 float* __restrict__ coef, float* __restrict__ scal,
                                                                 float* __restrict__ d_gamma, float* __restrict__ d_beta) {
  __shared__ double s_sum[16][kHidden], s_sx[16][kHidden];
  __shared__ float s_mx[16][kHidden];
  __shared__ float s_bound[kHidden];
  const int c = threadIdx.x & 63, grp = threadIdx.x >> 6;
  {
    double a = 0.0, b = 0.0;
    float m = 0.f;
    for (int k = grp; k < n_partials; k += 16) {
      const float* o = partial + (size_t)k * 3 * kHidden;
      a += (double)o[c];
      b += (double)o[kHidden + c];
      m = fmaxf(m, o[2 * kHidden + c]);
    }
    s_sum[grp][c] = a; s_sx[grp][c] = b; s_mx[grp][c] = m;
  }
  __syncthreads();
  if (grp != 0) return;
  double sum = 0.0, sx = 0.0;
  float mx = 0.f;
#pragma unroll
  for (int g = 0; g < 16; ++g) { sum += s_sum[g][c]; sx += s_sx[g][c]; mx = fmaxf(mx, s_mx[g][c]); }
  const float cum = scal[0];
  float mult = 1.f, m1 = 0.f, m2 = 0.f, bound = mx;
  if (rec) {
    const float g = gamma ? gamma[c] : 1.f;
    mult = g * rec[3 * kHidden + c];
    m1 = (float)(sum / count);
    m2 = (float)(sx / count);
    bound = fabsf(mult) * (mx + fabsf(m1) + 8.f * fabsf(m2));
    if (d_gamma) d_gamma[c] = (float)(sx / (double)cum);
    if (d_beta) d_beta[c] = (float)(sum / (double)cum);
  }
  coef[c] = mult; coef[kHidden + c] = m1; coef[2 * kHidden + c] = m2;
  s_bound[c] = bound;
  // only warps 0 and 1 (threads 0..63) are left: a named barrier over those 64 threads
  asm volatile("bar.sync 1, 64;" ::: "memory");
  if (c == 0) {
    float b = 0.f;
    for (int k = 0; k < kHidden; ++k) b = fmaxf(b, s_bound[k]);
    // bring the largest possible |dc| to ~2^8: headroom for the adjoint conv that follows (fp16 max 65504)
    float s = 1.f;
    if (b > 0.f && b < 3.0e38f) {
      int e;
      frexpf(b, &e);                       // b = m * 2^e, m in [0.5, 1)
      s = ldexpf(1.f, 8 - e);
    }
    scal[1] = s;
    scal[0] = cum * s;
  }
}

__global__ void __launch_bounds__(256) act_bwd_apply_kernel(const ActBwdParams p, const float* __restrict__ coef,
                                                            const float* __restrict__ scal) {
  __shared__ float s_c[3 * kHidden], s_mi[2 * kHidden];
  if (threadIdx.x < 3 * kHidden) s_c[threadIdx.x] = coef[threadIdx.x];
  if (threadIdx.x < 2 * kHidden) s_mi[threadIdx.x] = p.rec ? p.rec[2 * kHidden + threadIdx.x] : (threadIdx.x < kHidden ? 0.f : 1.f);
  __syncthreads();
  const float s = scal[1];
  const long long n_vec = p.plane_elems / 8;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (long long)gridDim.x * blockDim.x) {
    const long long i0 = v * 8;
    const int c0 = (int)(i0 & (kHidden - 1));
    const uint4 gh = *reinterpret_cast<const uint4*>(p.g + i0);
    const uint4 gl = *reinterpret_cast<const uint4*>(p.g + p.plane_elems + i0);
    const uint4 ah = *reinterpret_cast<const uint4*>(p.act + i0);
    uint4 ph = make_uint4(0, 0, 0, 0), pl = ph;
    if (p.pre) {
      ph = *reinterpret_cast<const uint4*>(p.pre + i0);
      pl = *reinterpret_cast<const uint4*>(p.pre + p.plane_elems + i0);
    }
    const __half* ghh = reinterpret_cast<const __half*>(&gh);
    const __half* gll = reinterpret_cast<const __half*>(&gl);
    const __half* ahh = reinterpret_cast<const __half*>(&ah);
    const __half* phh = reinterpret_cast<const __half*>(&ph);
    const __half* pll = reinterpret_cast<const __half*>(&pl);
    uint4 oh, ol;
    uint32_t* ohp = reinterpret_cast<uint32_t*>(&oh);
    uint32_t* olp = reinterpret_cast<uint32_t*>(&ol);
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float r[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = c0 + e + u;
        float dy = join_f16(ghh[e + u], gll[e + u]);
        dy = __half2float(ahh[e + u]) > 0.f ? dy : 0.f;
        if (p.pre) {
          const float xhat = (join_f16(phh[e + u], pll[e + u]) - s_mi[c]) * s_mi[kHidden + c];
          dy = s_c[c] * (dy - s_c[kHidden + c] - xhat * s_c[2 * kHidden + c]);
        }
        r[u] = dy * s;
      }
      split_f16x2(r[0], r[1], ohp[e >> 1], olp[e >> 1]);
    }
    *reinterpret_cast<uint4*>(p.out + i0) = oh;
    *reinterpret_cast<uint4*>(p.out + p.plane_elems + i0) = ol;
  }
}

// ---- wgrad of a hidden 64 -> 64 layer on the CUDA cores ------------------------------------------------------
// dW[o][c][ky][kx] = sum over (frame, y, x) of d[y][x][o] * a[y+ky-1][x+kx-1][c]  (zero outside the frame).
// 256 threads = 16 groups of 4 output channels x 16 groups of 4 input channels; a thread keeps 9 x 4 x 4 sums.
// Tiles of kWgRows x 32 pixels are staged in shared memory as fp32 (hi/lo pairs joined on the way in).
struct WgradParams {
  const __half* a;            // input activations of the layer (planes)
  const __half* d;            // gradient w.r.t. the layer's output (planes)
  long long plane_elems;
  int NF, Hc, Wc;
  int tiles_x, tiles_y;
  long long n_tiles;
};

__global__ void __launch_bounds__(256, 1) wgrad_hidden_kernel(const WgradParams p, float* __restrict__ partial) {
  extern __shared__ float wg_smem[];
  float* s_a = wg_smem;                                              // [(kWgRows+2)][kWgTile+2][64]
  float* s_d = wg_smem + (kWgRows + 2) * (kWgTile + 2) * kHidden;      // [kWgRows][kWgTile][64]
  const int tc = threadIdx.x & 15, to = threadIdx.x >> 4;
  float acc[9][4][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.f;
  for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const int per_frame = p.tiles_x * p.tiles_y;
    const int nf = (int)(tile / per_frame);
    const int rem = (int)(tile - (long long)nf * per_frame);
    const int y0 = (rem / p.tiles_x) * kWgRows, x0 = (rem % p.tiles_x) * kWgTile;
    __syncthreads();
    // stage a (with halo) and d; 8 channels (16 B per plane) per thread and step
    const int n_a = (kWgRows + 2) * (kWgTile + 2) * 8;
    for (int v = threadIdx.x; v < n_a; v += 256) {
      const int c8 = v & 7, px = v >> 3;
      const int rx = px % (kWgTile + 2), ry = px / (kWgTile + 2);
      const int y = y0 + ry - 1, x = x0 + rx - 1;
      float vals[8];
      if (y >= 0 && y < p.Hc && x >= 0 && x < p.Wc) {
        const long long i0 = (((long long)nf * p.Hc + y) * p.Wc + x) * kHidden + c8 * 8;
        const uint4 h4 = *reinterpret_cast<const uint4*>(p.a + i0);
        const uint4 l4 = *reinterpret_cast<const uint4*>(p.a + p.plane_elems + i0);
        const __half* hh = reinterpret_cast<const __half*>(&h4);
        const __half* ll = reinterpret_cast<const __half*>(&l4);
#pragma unroll
        for (int e = 0; e < 8; ++e) vals[e] = join_f16(hh[e], ll[e]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) vals[e] = 0.f;
      }
      float4* dst = reinterpret_cast<float4*>(s_a + (size_t)px * kHidden + c8 * 8);
      dst[0] = make_float4(vals[0], vals[1], vals[2], vals[3]);
      dst[1] = make_float4(vals[4], vals[5], vals[6], vals[7]);
    }
    const int n_d = kWgRows * kWgTile * 8;
    for (int v = threadIdx.x; v < n_d; v += 256) {
      const int c8 = v & 7, px = v >> 3;
      const int rx = px % kWgTile, ry = px / kWgTile;
      const int y = y0 + ry, x = x0 + rx;
      float vals[8];
      if (y < p.Hc && x < p.Wc) {
        const long long i0 = (((long long)nf * p.Hc + y) * p.Wc + x) * kHidden + c8 * 8;
        const uint4 h4 = *reinterpret_cast<const uint4*>(p.d + i0);
        const uint4 l4 = *reinterpret_cast<const uint4*>(p.d + p.plane_elems + i0);
        const __half* hh = reinterpret_cast<const __half*>(&h4);
        const __half* ll = reinterpret_cast<const __half*>(&l4);
#pragma unroll
        for (int e = 0; e < 8; ++e) vals[e] = join_f16(hh[e], ll[e]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) vals[e] = 0.f;
      }
      float4* dst = reinterpret_cast<float4*>(s_d + (size_t)px * kHidden + c8 * 8);
      dst[0] = make_float4(vals[0], vals[1], vals[2], vals[3]);
      dst[1] = make_float4(vals[4], vals[5], vals[6], vals[7]);
    }
    __syncthreads();
    for (int ry = 0; ry < kWgRows; ++ry) {
#pragma unroll 2
      for (int rx = 0; rx < kWgTile; ++rx) {
        const float4 d4 = *reinterpret_cast<const float4*>(s_d + (size_t)(ry * kWgTile + rx) * kHidden + to * 4);
        const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4 a4 = *reinterpret_cast<const float4*>(
                s_a + (size_t)((ry + ky) * (kWgTile + 2) + rx + kx) * kHidden + tc * 4);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[ky * 3 + kx][i][j] = fmaf(dv[i], av[j], acc[ky * 3 + kx][i][j]);
          }
      }
    }
  }
  // partial[cta][o][c][tap]  (the weight tensor's own layout)
  float* o = partial + (size_t)blockIdx.x * kHidden * kHidden * 9;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int t = 0; t < 9; ++t) o[((size_t)(to * 4 + i) * kHidden + (tc * 4 + j)) * 9 + t] = acc[t][i][j];
}

// out[e] = sum over CTAs of partial[cta][e] (fp64, fixed order) / cum
__global__ void wgrad_finalize_kernel(const float* __restrict__ partial, int n_partials, int n_elems,
                                      const float* __restrict__ scal, float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elems) return;
  double s = 0.0;
  for (int k = 0; k < n_partials; ++k) s += (double)partial[(size_t)k * n_elems + e];
  out[e] = (float)(s / (double)scal[0]);
}

// ---- thin layers ------------------------------------------------------------------------------------------------
// last conv (64 -> CO, CO = 4 FFDNet / 1 DnCNN): dW[o][c][tap] = sum_p gn[p][o] * a[p+tap][c], gn = -(upstream
// gradient), pixel-unshuffled for FFDNet (conv pixel (y,x), channel o=2r+s <-> cube pixel (2y+r, 2x+s)).
// gsc [B,H,W,T] = the upstream gradient already multiplied by -scale.  One thread per input channel c.
template <int CO>
__global__ void __launch_bounds__(256) wgrad_last_kernel(const __half* __restrict__ a, long long plane_elems,
                                                         const float* __restrict__ gsc, int B, int H, int W, int T,
                                                         float* __restrict__ partial) {
  constexpr int SC = CO == 4 ? 2 : 1;
  const int Hc = H / SC, Wc = W / SC;
  const int c = threadIdx.x & 63, sub = threadIdx.x >> 6;
  float acc[CO][9];
#pragma unroll
  for (int o = 0; o < CO; ++o)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[o][t] = 0.f;
  // every INPUT pixel q once (its 64 channels: one coalesced load per plane); the output pixels it feeds are
  // p = q - tap + 1, whose gradients are a handful of scalars shared by the 64 threads (broadcast loads)
  const long long n_px = (long long)B * T * Hc * Wc;
  for (long long px = (long long)blockIdx.x * 4 + sub; px < n_px; px += (long long)gridDim.x * 4) {
    const int x = (int)(px % Wc);
    const int y = (int)((px / Wc) % Hc);
    const int nf = (int)(px / ((long long)Wc * Hc));
    const int b = nf / T, t = nf % T;
    const long long i = px * kHidden + c;
    const float av = join_f16(a[i], a[plane_elems + i]);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = y - ky + 1, xx = x - kx + 1;          // output pixel that sees q through tap (ky, kx)
        if (yy < 0 || yy >= Hc || xx < 0 || xx >= Wc) continue;
#pragma unroll
        for (int o = 0; o < CO; ++o) {
          const int r = SC == 2 ? (o >> 1) : 0, s = SC == 2 ? (o & 1) : 0;
          const float gv = __ldg(gsc + (((long long)b * H + (SC * yy + r)) * W + (SC * xx + s)) * T + t);
          acc[o][ky * 3 + kx] = fmaf(gv, av, acc[o][ky * 3 + kx]);
        }
      }
  }
  __shared__ float s_red[4][CO * 9][kHidden + 1];
#pragma unroll
  for (int o = 0; o < CO; ++o)
#pragma unroll
    for (int t = 0; t < 9; ++t) s_red[sub][o * 9 + t][c] = acc[o][t];
  __syncthreads();
  // partial[cta][o][c][tap]
  for (int e = threadIdx.x; e < CO * kHidden * 9; e += 256) {
    const int t = e % 9, cc = (e / 9) % kHidden, o = e / (9 * kHidden);
    partial[(size_t)blockIdx.x * CO * kHidden * 9 + e] =
        (s_red[0][o * 9 + t][cc] + s_red[1][o * 9 + t][cc]) + (s_red[2][o * 9 + t][cc] + s_red[3][o * 9 + t][cc]);
  }
}

// first conv (CI -> 64, CI = 5 FFDNet: [sigma map, 4 pixel-unshuffled sub-images] / 1 DnCNN):
// dW[o][ci][tap] = sum_p d[p][o] * in[p+tap][ci].  zp [B,T,H,W] = the layer's input frames (frame-planar z').
// One thread per output channel o.
template <int CI>
__global__ void __launch_bounds__(256) wgrad_first_kernel(const __half* __restrict__ d, long long plane_elems,
                                                          const float* __restrict__ zp, float sigma, int NF, int H, int W,
                                                          float* __restrict__ partial) {
  constexpr int SC = CI == 5 ? 2 : 1;
  const int Hc = H / SC, Wc = W / SC;
  const int o = threadIdx.x & 63, sub = threadIdx.x >> 6;
  float acc[CI][9];
#pragma unroll
  for (int ci = 0; ci < CI; ++ci)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[ci][t] = 0.f;
  const long long n_px = (long long)NF * Hc * Wc;
  for (long long px = (long long)blockIdx.x * 4 + sub; px < n_px; px += (long long)gridDim.x * 4) {
    const int x = (int)(px % Wc);
    const int y = (int)((px / Wc) % Hc);
    const int nf = (int)(px / ((long long)Wc * Hc));
    const long long i = px * kHidden + o;
    const float dv = join_f16(d[i], d[plane_elems + i]);
    const float* frame = zp + (long long)nf * H * W;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = y + ky - 1, xx = x + kx - 1;
        if (yy < 0 || yy >= Hc || xx < 0 || xx >= Wc) continue;      // zero padding (also of the sigma map)
        if (CI == 5) {
          acc[0][ky * 3 + kx] = fmaf(dv, sigma, acc[0][ky * 3 + kx]);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            acc[1 + q][ky * 3 + kx] = fmaf(dv, frame[(long long)(2 * yy + (q >> 1)) * W + 2 * xx + (q & 1)], acc[1 + q][ky * 3 + kx]);
        } else {
          acc[0][ky * 3 + kx] = fmaf(dv, frame[(long long)yy * W + xx], acc[0][ky * 3 + kx]);
        }
      }
  }
  __shared__ float s_red[4][CI * 9][kHidden + 1];
#pragma unroll
  for (int ci = 0; ci < CI; ++ci)
#pragma unroll
    for (int t = 0; t < 9; ++t) s_red[sub][ci * 9 + t][o] = acc[ci][t];
  __syncthreads();
  // partial[cta][o][ci][tap]
  for (int e = threadIdx.x; e < kHidden * CI * 9; e += 256) {
    const int t = e % 9, ci = (e / 9) % CI, oo = e / (9 * CI);
    partial[(size_t)blockIdx.x * kHidden * CI * 9 + e] =
        (s_red[0][ci * 9 + t][oo] + s_red[1][ci * 9 + t][oo]) + (s_red[2][ci * 9 + t][oo] + s_red[3][ci * 9 + t][oo]);
  }
}

__global__ void scale_cube_kernel(float* __restrict__ dst, const float* __restrict__ src, float s, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i] * s;
}
__global__ void set_scal_kernel(float* scal, float cum) { scal[0] = cum; scal[1] = 1.f; }

// ---- launchers (used by api.cu) -----------------------------------------------------------------------------
size_t backward_scratch_floats() {
  // per-channel reduction partials, coefficients, scalars, wgrad partials (hidden layer: the largest)
  return (size_t)kRedCtas * 3 * kHidden + 3 * kHidden + 64 + std::max((size_t)kWgMaxCtas * kHidden * kHidden * 9, (size_t)kThinCtas * kHidden * 5 * 9);
}

int act_bwd_launch(const __half* g, const __half* act, const __half* pre, const float* rec, const float* gamma,
                   __half* out, long long plane_elems, long long count, float* scratch, float* d_gamma, float* d_beta,
                   cudaStream_t st) {
  ActBwdParams p{g, act, pre, rec, gamma, out, plane_elems, (double)count};
  float* partial = scratch;
  float* coef = partial + (size_t)kRedCtas * 3 * kHidden;
  float* scal = coef + 3 * kHidden;
  const long long n_px = plane_elems / kHidden;
  const int blocks = (int)std::min<long long>(kRedCtas, (n_px + 3) / 4);
  act_bwd_reduce_kernel<<<blocks, 256, 0, st>>>(p, partial);
  act_bwd_finalize_kernel<<<1, 1024, 0, st>>>(partial, blocks, rec, gamma, (double)count, coef, scal, d_gamma, d_beta);
  const long long n_vec = plane_elems / 8;
  const int ablocks = (int)std::min<long long>((n_vec + 255) / 256, (long long)num_sms() * 16);
  act_bwd_apply_kernel<<<ablocks, 256, 0, st>>>(p, coef, scal);
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

int wgrad_hidden_launch(const __half* a, const __half* d, long long plane_elems, int NF, int Hc, int Wc, float* scratch,
                        float* d_weight, cudaStream_t st) {
  WgradParams p;
  p.a = a; p.d = d; p.plane_elems = plane_elems; p.NF = NF; p.Hc = Hc; p.Wc = Wc;
  p.tiles_x = (Wc + kWgTile - 1) / kWgTile;
  p.tiles_y = (Hc + kWgRows - 1) / kWgRows;
  p.n_tiles = (long long)NF * p.tiles_x * p.tiles_y;
  float* scal = scratch + (size_t)kRedCtas * 3 * kHidden + 3 * kHidden;
  float* partial = scal + 64;
  const int ctas = (int)std::min<long long>(std::min(kWgMaxCtas, num_sms()), p.n_tiles);
  const size_t smem = ((size_t)(kWgRows + 2) * (kWgTile + 2) + (size_t)kWgRows * kWgTile) * kHidden * sizeof(float);
  DEQSCI_CUDA(cudaFuncSetAttribute(wgrad_hidden_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  wgrad_hidden_kernel<<<ctas, 256, smem, st>>>(p, partial);
  const int n = kHidden * kHidden * 9;
  wgrad_finalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(partial, ctas, n, scal, d_weight);
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

int wgrad_last_launch(int cout, const __half* a, long long plane_elems, const float* gsc, int B, int H, int W, int T,
                      float* scratch, float* d_weight, cudaStream_t st) {
  float* scal = scratch + (size_t)kRedCtas * 3 * kHidden + 3 * kHidden;
  float* partial = scal + 64;
  const int ctas = kThinCtas;
  if (cout == 4) wgrad_last_kernel<4><<<ctas, 256, 0, st>>>(a, plane_elems, gsc, B, H, W, T, partial);
  else if (cout == 1) wgrad_last_kernel<1><<<ctas, 256, 0, st>>>(a, plane_elems, gsc, B, H, W, T, partial);
  else { set_error("backward: last layer with %d outputs unsupported", cout); return DEQSCI_ERR_INVALID; }
  const int n = cout * kHidden * 9;
  wgrad_finalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(partial, ctas, n, scal, d_weight);
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

int wgrad_first_launch(int cin, const __half* d, long long plane_elems, const float* zp, float sigma, int NF, int H, int W,
                       float* scratch, float* d_weight, cudaStream_t st) {
  float* scal = scratch + (size_t)kRedCtas * 3 * kHidden + 3 * kHidden;
  float* partial = scal + 64;
  const int ctas = kThinCtas;
  if (cin == 5) wgrad_first_kernel<5><<<ctas, 256, 0, st>>>(d, plane_elems, zp, sigma, NF, H, W, partial);
  else if (cin == 1) wgrad_first_kernel<1><<<ctas, 256, 0, st>>>(d, plane_elems, zp, sigma, NF, H, W, partial);
  else { set_error("backward: first layer with %d inputs unsupported", cin); return DEQSCI_ERR_INVALID; }
  const int n = kHidden * cin * 9;
  wgrad_finalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(partial, ctas, n, scal, d_weight);
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

int backward_begin(float* scratch, const float* g, float* gsc, float scale, long long n, cudaStream_t st) {
  float* scal = scratch + (size_t)kRedCtas * 3 * kHidden + 3 * kHidden;
  set_scal_kernel<<<1, 1, 0, st>>>(scal, scale);
  scale_cube_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, (long long)num_sms() * 16), 256, 0, st>>>(gsc, g, -scale, n);
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

}  // namespace deqsci
