// Kernel group 3: Anderson acceleration bookkeeping (HBM-bound, fp32 state).
//
// Restates the per-iteration tensor algebra of andersonexp
// (solvers/new_equilibrium_utils_yaping.py:174-184) as three launches:
//   gram  : g = F[slot]-X[slot] -> G[slot]; partial <g, G[j]> (j<n), |F[slot]|^2 per 4096-float chunk
//   solve : fixed-order fp64 reduction of the partials, incremental Gram row/column update,
//           bordered (n+1)x(n+1) LU solve with partial pivoting (fp32, like LAPACK gesv behind
//           torch.solve), whole-batch residual norms -- everything stays on the device
//   mix   : X[slot] = beta * sum_j alpha_j F[j] (+ (1-beta) * sum_j alpha_j X[j])
// Only the row of the Gram matrix that changed is recomputed (the reference recomputes all n^2
// entries with a bmm of K = N each iteration); reductions use a fixed chunking, so results are
// run-to-run deterministic.
//
// Algorithmic bytes per iteration and sample (n = m = 5, beta = 1): gram reads F,X (2N) + G[j!=slot]
// ((n-1)N) and writes G[slot] (N); mix reads F[0..n) (nN) and writes X[slot] (N): (2n+3) N floats.
#include "common.cuh"

namespace deqsci {

constexpr int kMaxM = 8;            // history slots supported (reference default m = 5)
constexpr int kChunk = 4096;        // floats of N reduced by one CTA
constexpr int kGramThreads = 256;

__host__ __device__ inline int anderson_chunks(long long N) { return (int)((N + kChunk - 1) / kChunk); }

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

// partials layout: [B][chunks][kMaxM + 1]; entry j < n: <g, G_j> (j == slot: |g|^2); entry kMaxM: |F_slot|^2
template <int VEC, bool WRITE_G>
__global__ void __launch_bounds__(kGramThreads) anderson_gram_kernel(const float* __restrict__ X,
                                                                     const float* __restrict__ F,
                                                                     float* __restrict__ G,
                                                                     float* __restrict__ partials, int B,
                                                                     long long N, int slot, int n) {
  pdl_launch_dependents();       // launched with programmatic stream serialization: see launch_pdl (common.cuh)
  pdl_wait_predecessor();
  const int b = blockIdx.y;
  const int chunk = blockIdx.x;
  const long long sstride = (long long)B * N;          // history is slot-major: [m][B][N]
  const long long base = (long long)b * N;
  const float* xs = X + base + slot * sstride;
  const float* fs = F + base + slot * sstride;
  float* gs = G + base + slot * sstride;
  float acc[kMaxM + 1];
#pragma unroll
  for (int j = 0; j <= kMaxM; ++j) acc[j] = 0.f;
  const long long lo = (long long)chunk * kChunk;
  const long long hi = (lo + kChunk < N) ? lo + kChunk : N;
  if (VEC == 4) {
    for (long long e = lo + threadIdx.x * 4; e < hi; e += kGramThreads * 4) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(fs + e));
      const float4 x = __ldg(reinterpret_cast<const float4*>(xs + e));
      float4 g;
      g.x = f.x - x.x; g.y = f.y - x.y; g.z = f.z - x.z; g.w = f.w - x.w;
      if (WRITE_G) *reinterpret_cast<float4*>(gs + e) = g;
      acc[kMaxM] += f.x * f.x + f.y * f.y + f.z * f.z + f.w * f.w;
#pragma unroll
      for (int j = 0; j < kMaxM; ++j) {
        if (j < n) {
          float4 o = g;
          if (j != slot) o = __ldg(reinterpret_cast<const float4*>(G + base + j * sstride + e));
          acc[j] += g.x * o.x + g.y * o.y + g.z * o.z + g.w * o.w;
        }
      }
    }
  } else {
    for (long long e = lo + threadIdx.x; e < hi; e += kGramThreads) {
      const float f = fs[e], x = xs[e];
      const float g = f - x;
      if (WRITE_G) gs[e] = g;
      acc[kMaxM] += f * f;
#pragma unroll
      for (int j = 0; j < kMaxM; ++j) {
        if (j < n) {
          const float o = (j == slot) ? g : G[base + j * sstride + e];
          acc[j] += g * o;
        }
      }
    }
  }
  __shared__ float red[kGramThreads / 32][kMaxM + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j <= kMaxM; ++j) {
    const float s = warp_sum(acc[j]);
    if (lane == 0) red[warp][j] = s;
  }
  __syncthreads();
  if (threadIdx.x <= kMaxM) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kGramThreads / 32; ++w) s += red[w][threadIdx.x];
    partials[((long long)b * gridDim.x + chunk) * (kMaxM + 1) + threadIdx.x] = s;
  }
}

// Solves the bordered system H a = e0, H = [[0, 1^T],[1, GG + lam I]] of size S = n+1, returns
// a[1..n] in alpha.  Right-looking LU with partial pivoting (first maximal |entry| wins, pivot
// applied through its reciprocal as LAPACK's getf2 does), then forward/back substitution.
// S is a compile-time constant so the matrix stays in registers (fully unrolled loops).
template <int S>
__device__ __forceinline__ void bordered_solve_fixed(const float* gram_b, int m, float lam, float* alpha_out) {
  constexpr int n = S - 1;
  float Hm[S][S];
  float rhs[S];
#pragma unroll
  for (int i = 0; i < S; ++i) {
    rhs[i] = (i == 0) ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      float v;
      if (i == 0 && j == 0) v = 0.f;
      else if (i == 0 || j == 0) v = 1.f;
      else v = gram_b[(i - 1) * m + (j - 1)] + ((i == j) ? lam : 0.f);
      Hm[i][j] = v;
    }
  }
#pragma unroll
  for (int k = 0; k < S; ++k) {
    int p = k;
    float best = fabsf(Hm[k][k]);
#pragma unroll
    for (int i = k + 1; i < S; ++i) {
      const float v = fabsf(Hm[i][k]);
      if (v > best) { best = v; p = i; }
    }
    // row swap k <-> p without dynamic register indexing
#pragma unroll
    for (int i = k + 1; i < S; ++i) {
      if (i == p) {
#pragma unroll
        for (int j = 0; j < S; ++j) { const float t = Hm[k][j]; Hm[k][j] = Hm[i][j]; Hm[i][j] = t; }
        const float t = rhs[k]; rhs[k] = rhs[i]; rhs[i] = t;
      }
    }
    const float rp = 1.0f / Hm[k][k];
#pragma unroll
    for (int i = k + 1; i < S; ++i) {
      const float l = Hm[i][k] * rp;
      Hm[i][k] = l;
#pragma unroll
      for (int j = k + 1; j < S; ++j) Hm[i][j] = Hm[i][j] - l * Hm[k][j];
    }
  }
#pragma unroll
  for (int i = 1; i < S; ++i) {
    float v = rhs[i];
#pragma unroll
    for (int j = 0; j < i; ++j) v -= Hm[i][j] * rhs[j];
    rhs[i] = v;
  }
#pragma unroll
  for (int i = S - 1; i >= 0; --i) {
    float v = rhs[i];
#pragma unroll
    for (int j = i + 1; j < S; ++j) v -= Hm[i][j] * rhs[j];
    rhs[i] = v / Hm[i][i];
  }
#pragma unroll
  for (int j = 0; j < kMaxM; ++j)
    if (j < m) alpha_out[j] = (j < n) ? rhs[(j + 1 < S) ? j + 1 : 0] : 0.f;
}

__device__ void bordered_solve(const float* gram_b, int m, int n, float lam, float* alpha_out) {
  switch (n) {
    case 1: bordered_solve_fixed<2>(gram_b, m, lam, alpha_out); break;
    case 2: bordered_solve_fixed<3>(gram_b, m, lam, alpha_out); break;
    case 3: bordered_solve_fixed<4>(gram_b, m, lam, alpha_out); break;
    case 4: bordered_solve_fixed<5>(gram_b, m, lam, alpha_out); break;
    case 5: bordered_solve_fixed<6>(gram_b, m, lam, alpha_out); break;
    case 6: bordered_solve_fixed<7>(gram_b, m, lam, alpha_out); break;
    case 7: bordered_solve_fixed<8>(gram_b, m, lam, alpha_out); break;
    default: bordered_solve_fixed<9>(gram_b, m, lam, alpha_out); break;
  }
}

// One CTA; one warp per sample (looping), lane-parallel fp64 reduction over chunks.
__global__ void __launch_bounds__(1024) anderson_solve_kernel(const float* __restrict__ partials,
                                                              float* __restrict__ gram, float* __restrict__ alpha,
                                                              float* __restrict__ res, int B, int m, int chunks,
                                                              int slot, int n, float lam, float res_eps,
                                                              int do_solve) {
  pdl_launch_dependents();
  pdl_wait_predecessor();
  __shared__ double s_g2[32], s_f2[32];
  __shared__ float s_min[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  double g2 = 0.0, f2 = 0.0;   // per-warp running sums over its samples (fixed order)
  float rmin = 3.0e38f;        // smallest PER-SAMPLE residual among this warp's samples
  for (int b = warp; b < B; b += nwarps) {
    double acc[kMaxM + 1];
#pragma unroll
    for (int j = 0; j <= kMaxM; ++j) acc[j] = 0.0;
    for (int c = lane; c < chunks; c += 32) {
      const float* p = partials + ((long long)b * chunks + c) * (kMaxM + 1);
#pragma unroll
      for (int j = 0; j <= kMaxM; ++j) acc[j] += (double)p[j];
    }
#pragma unroll
    for (int j = 0; j <= kMaxM; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
      float* gb = gram + (long long)b * m * m;
#pragma unroll
      for (int j = 0; j < kMaxM; ++j) {
        if (j < n) {
          const float v = (float)acc[j];
          gb[slot * m + j] = v;
          gb[j * m + slot] = v;
        }
      }
      g2 += acc[slot < kMaxM ? slot : 0];
      f2 += acc[kMaxM];
      // what the reference's whole-batch test (:184) computes when this sample is solved on its own (batch 1)
      rmin = fminf(rmin, (float)(sqrt(acc[slot < kMaxM ? slot : 0]) / ((double)res_eps + sqrt(acc[kMaxM]))));
      if (do_solve) bordered_solve(gb, m, n, lam, alpha + (long long)b * m);
    }
  }
  if (lane == 0) { s_g2[warp] = g2; s_f2[warp] = f2; s_min[warp] = rmin; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    float mn = 3.0e38f;
    for (int w = 0; w < nwarps; ++w) { a += s_g2[w]; c += s_f2[w]; mn = fminf(mn, s_min[w]); }
    const float ng = (float)sqrt(a), nf = (float)sqrt(c);
    res[0] = (float)((double)ng / ((double)res_eps + (double)nf));
    res[1] = ng;
    res[2] = nf;
    res[3] = mn;      // min over samples of the per-sample residual
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) anderson_mix_kernel(float* __restrict__ X, const float* __restrict__ F,
                                                           const float* __restrict__ alpha, int B, int m,
                                                           long long N, int slot, int n, float beta) {
  pdl_launch_dependents();
  pdl_wait_predecessor();
  const int b = blockIdx.y;
  const long long sstride = (long long)B * N;          // history is slot-major: [m][B][N]
  float a[kMaxM];
#pragma unroll
  for (int j = 0; j < kMaxM; ++j) a[j] = (j < n) ? __ldg(alpha + (long long)b * m + j) : 0.f;
  const long long base = (long long)b * N;
  const bool use_x = (beta != 1.0f);
  const float omb = 1.0f - beta;
  const long long stride = (long long)gridDim.x * blockDim.x * VEC;
  for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC; e < N; e += stride) {
    if (VEC == 4) {
      float4 sf = make_float4(0.f, 0.f, 0.f, 0.f), sx = sf;
#pragma unroll
      for (int j = 0; j < kMaxM; ++j) {
        if (j < n) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(F + base + j * sstride + e));
          sf.x = fmaf(a[j], f.x, sf.x); sf.y = fmaf(a[j], f.y, sf.y);
          sf.z = fmaf(a[j], f.z, sf.z); sf.w = fmaf(a[j], f.w, sf.w);
          if (use_x) {
            const float4 x = *reinterpret_cast<const float4*>(X + base + j * sstride + e);
            sx.x = fmaf(a[j], x.x, sx.x); sx.y = fmaf(a[j], x.y, sx.y);
            sx.z = fmaf(a[j], x.z, sx.z); sx.w = fmaf(a[j], x.w, sx.w);
          }
        }
      }
      float4 o;
      if (use_x) {
        o.x = beta * sf.x + omb * sx.x; o.y = beta * sf.y + omb * sx.y;
        o.z = beta * sf.z + omb * sx.z; o.w = beta * sf.w + omb * sx.w;
      } else {
        o = sf;
      }
      *reinterpret_cast<float4*>(X + base + slot * sstride + e) = o;
    } else {
      float sf = 0.f, sx = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxM; ++j) {
        if (j < n) {
          sf = fmaf(a[j], F[base + j * sstride + e], sf);
          if (use_x) sx = fmaf(a[j], X[base + j * sstride + e], sx);
        }
      }
      X[base + slot * sstride + e] = use_x ? (beta * sf + omb * sx) : sf;
    }
  }
}

static bool vec4_ok(const void* p, long long N) {
  return (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
}

}  // namespace deqsci

using namespace deqsci;

extern "C" size_t deqsci_anderson_scratch_floats(int B, int m, long long N) {
  if (B <= 0 || N <= 0) return 0;
  (void)m;
  return (size_t)B * (size_t)anderson_chunks(N) * (kMaxM + 1) + 8;
}

extern "C" int deqsci_anderson_update(const float* X, const float* F, float* G, float* gram, float* alpha,
                                      float* res, float* scratch, int B, int m, long long N, int slot, int n,
                                      float lam, float res_eps, void* stream) {
  DEQSCI_CHECK_ARG(X && F && G && gram && alpha && res && scratch, "anderson_update: null pointer");
  DEQSCI_CHECK_ARG(B > 0 && N > 0, "anderson_update: B=%d N=%lld", B, N);
  DEQSCI_CHECK_ARG(m >= 1 && m <= kMaxM, "anderson_update: m=%d unsupported (1..%d)", m, kMaxM);
  DEQSCI_CHECK_ARG(n >= 1 && n <= m && slot >= 0 && slot < n, "anderson_update: slot=%d n=%d m=%d", slot, n, m);
  DEQSCI_CHECK_ARG(B <= 65535, "anderson_update: B=%d too large for one launch", B);
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = anderson_chunks(N);
  dim3 grid(chunks, B);
  {
  ProfScope prof(PK_AND_GRAM, st);
  if (vec4_ok(X, N) && vec4_ok(F, N) && vec4_ok(G, N))
    DEQSCI_CUDA(launch_pdl(anderson_gram_kernel<4, true>, grid, kGramThreads, 0, st, X, F, G, scratch, B, N, slot, n));
  else
    DEQSCI_CUDA(launch_pdl(anderson_gram_kernel<1, true>, grid, kGramThreads, 0, st, X, F, G, scratch, B, N, slot, n));
  }
  DEQSCI_LAUNCH_CHECK();
  ProfScope prof2(PK_AND_SOLVE, st);
  int threads = 32 * (B < 32 ? B : 32);
  DEQSCI_CUDA(launch_pdl(anderson_solve_kernel, 1, threads, 0, st, (const float*)scratch, gram, alpha, res, B, m, chunks, slot, n,
                         lam, res_eps, 1));
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

extern "C" int deqsci_anderson_mix(float* X, const float* F, const float* alpha, int B, int m, long long N,
                                   int slot, int n, float beta, void* stream) {
  DEQSCI_CHECK_ARG(X && F && alpha, "anderson_mix: null pointer");
  DEQSCI_CHECK_ARG(B > 0 && N > 0 && B <= 65535, "anderson_mix: B=%d N=%lld", B, N);
  DEQSCI_CHECK_ARG(m >= 1 && m <= kMaxM && n >= 1 && n <= m && slot >= 0 && slot < m,
                   "anderson_mix: slot=%d n=%d m=%d", slot, n, m);
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = 256;
  const bool v4 = vec4_ok(X, N) && vec4_ok(F, N);
  const long long per_block = (long long)threads * (v4 ? 4 : 1) * 2;   // 2 vectors per thread
  long long bx = (N + per_block - 1) / per_block;
  if (bx < 1) bx = 1;
  if (bx > 65535) bx = 65535;
  dim3 grid((unsigned)bx, B);
  ProfScope prof(PK_AND_MIX, st);
  if (v4) DEQSCI_CUDA(launch_pdl(anderson_mix_kernel<4>, grid, threads, 0, st, X, F, alpha, B, m, N, slot, n, beta));
  else    DEQSCI_CUDA(launch_pdl(anderson_mix_kernel<1>, grid, threads, 0, st, X, F, alpha, B, m, N, slot, n, beta));
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

// Residual of forward_iteration: a plays F (slot 0), b plays X, with m = n = 1; the difference is
// reduced on the fly and never written.  scratch: deqsci_anderson_scratch_floats(1, 1, count) + 8.
extern "C" int deqsci_residual(const float* a, const float* b, float* res, float* scratch, long long count,
                               float res_eps, void* stream) {
  DEQSCI_CHECK_ARG(a && b && res && scratch && count > 0, "residual: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = anderson_chunks(count);
  float* partials = scratch;
  float* gram = scratch + (size_t)chunks * (kMaxM + 1);   // 1 float
  float* alpha = gram + 4;                                 // unused (do_solve = 0)
  dim3 grid(chunks, 1);
  if (vec4_ok(a, count) && vec4_ok(b, count))
    anderson_gram_kernel<4, false><<<grid, kGramThreads, 0, st>>>(b, a, nullptr, partials, 1, count, 0, 1);
  else
    anderson_gram_kernel<1, false><<<grid, kGramThreads, 0, st>>>(b, a, nullptr, partials, 1, count, 0, 1);
  DEQSCI_LAUNCH_CHECK();
  anderson_solve_kernel<<<1, 32, 0, st>>>(partials, gram, alpha, res, 1, 1, chunks, 0, 1, 0.f, res_eps, 0);
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}
