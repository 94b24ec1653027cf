// Device-resident driver: one C-ABI call runs a whole DE-GAP reconstruction of a batch --
// x0 = At(y), andersonexp(f, x0) (solvers/new_equilibrium_utils_yaping.py:153-189), then the final
// f call that DEQFixedPoint.forward returns (:268) -- by queueing the library's own kernels from a
// C++ loop.  No per-iteration host round trip stalls the device: every iteration's residual lands in
// pinned host memory through a 16-byte async copy + event, and the host tests iteration k-1 while
// iteration k is already queued (rolling back one speculative sigma step on convergence).
// The same loop serves the train-mode forward solve (deqsci_reconstruct_train) and the backward solve of
// the implicit-differentiation hook (deqsci_adjoint_solve).
#include <mutex>
#include <vector>

#include "common.cuh"
#include <algorithm>

using namespace deqsci;

namespace {
size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Residual ring (pinned host floats + one event per iteration), pooled across calls: cudaMallocHost and
// event creation cost about a millisecond per reconstruction otherwise.  A ring belongs to one call at a
// time (concurrent calls on different threads each take their own); events belong to the device that was
// current when they were created.
struct ResidualRing {
  float* host = nullptr;
  int capacity = 0, device = -1;
  std::vector<cudaEvent_t> events;
};
std::mutex g_ring_mutex;
std::vector<ResidualRing*> g_free_rings;

ResidualRing* acquire_ring(int iters) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  ResidualRing* r = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_ring_mutex);
    for (size_t i = 0; i < g_free_rings.size(); ++i)
      if (g_free_rings[i]->device == dev) {
        r = g_free_rings[i];
        g_free_rings.erase(g_free_rings.begin() + i);
        break;
      }
  }
  if (!r) { r = new ResidualRing(); r->device = dev; }
  if (r->capacity < iters) {
    if (r->host) cudaFreeHost(r->host);
    r->host = nullptr;
    r->capacity = 0;
    if (cudaMallocHost(&r->host, (size_t)iters * 4 * sizeof(float)) != cudaSuccess) { delete r; return nullptr; }
    r->capacity = iters;
  }
  while ((int)r->events.size() < iters) {
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) break;
    r->events.push_back(e);
  }
  if ((int)r->events.size() < iters) {
    std::lock_guard<std::mutex> lock(g_ring_mutex);
    g_free_rings.push_back(r);
    return nullptr;
  }
  return r;
}
void release_ring(ResidualRing* r) {
  std::lock_guard<std::mutex> lock(g_ring_mutex);
  g_free_rings.push_back(r);
}

__global__ void scale_copy_kernel(float* __restrict__ dst, const float* __restrict__ src, float s, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i] * s;
}

struct Layout {
  size_t hist_floats, gram_floats, alpha_floats, scratch_floats, floats_total;
  size_t den_bytes, total_bytes;
};

Layout layout(const deqsci_denoiser* h, int B, int H, int W, int T, int m) {
  Layout L;
  const size_t N = (size_t)H * W * T;
  L.hist_floats = (size_t)3 * m * B * N;
  L.gram_floats = align_up_sz((size_t)B * m * m, 64);
  L.alpha_floats = align_up_sz((size_t)B * m, 64);
  L.scratch_floats = align_up_sz(deqsci_anderson_scratch_floats(B, m, (long long)N), 64);
  L.floats_total = L.hist_floats + L.gram_floats + L.alpha_floats + L.scratch_floats + 64 /*res*/ +
                   (size_t)kMaxBnLayers * 2 * kHidden /*running-statistics snapshot*/ +
                   align_up_sz((size_t)B * N, 64) /*one cube of scratch: the masked-adjoint map's intermediate*/;
  L.den_bytes = h ? deqsci_denoiser_workspace_bytes(h, B, H, W, T) : 0;
  L.total_bytes = 1024 + align_up_sz(L.floats_total * sizeof(float), 1024) + L.den_bytes;
  return L;
}
}  // namespace

extern "C" size_t deqsci_adjoint_solve_workspace_bytes(int B, int H, int W, int T, int m) {
  if (B <= 0 || H <= 0 || W <= 0 || T <= 0 || m < 2 || m > 8) return 0;
  return layout(nullptr, B, H, W, T, m).total_bytes;
}

extern "C" size_t deqsci_reconstruct_workspace_bytes(const deqsci_denoiser* h, int B, int H, int W, int T, int m) {
  if (!h || B <= 0 || H <= 0 || W <= 0 || T <= 0 || m < 2 || m > 8) return 0;
  Layout L = layout(h, B, H, W, T, m);
  return L.den_bytes == 0 ? 0 : L.total_bytes;
}

namespace {
// The andersonexp loop on one of three maps:
//   h, bn == nullptr : f = the iterate map, eval-mode plan (BatchNorm folded)
//   h, bn            : f = the iterate map in train mode (deqsci_iterate_train)
//   h == nullptr     : f(v) = gap_vjp(v) + adjoint_grad, the backward fixed-point map of tag 'ffdnet'
//                      (solvers/new_equilibrium_utils_yaping.py:274-277); y is unused
//   h, masks         : f(v) = gap_vjp(v - J_D^T v) + adjoint_grad, the same map for tag 'denoiser': h is the ADJOINT
//                      plan (transposed, flipped weights in reverse order) and masks the saved forward activations
int anderson_loop(const deqsci_denoiser* h, const float* y, const float* phi, const float* phi_sum,
                     const float* x0, float* out, const deqsci_solver_opts* o, const deqsci_bn_params* bn,
                     float momentum, float eps, const float* adjoint_grad, void* workspace,
                     size_t workspace_bytes, deqsci_solver_result* result, int B, int H, int W, int T,
                     void* stream, const void* const* masks = nullptr, float vjp_scale = 1.f) {
  DEQSCI_CHECK_ARG((h || adjoint_grad) && (y || adjoint_grad) && phi && phi_sum && out && o && workspace && result,
                   "reconstruct: null pointer");
  DEQSCI_CHECK_ARG(o->m >= 2 && o->m <= 8, "reconstruct: m=%d unsupported (2..8)", o->m);
  DEQSCI_CHECK_ARG(o->max_iter >= 2, "reconstruct: max_iter=%d (need >= 2)", o->max_iter);
  DEQSCI_CHECK_ARG(B > 0 && H > 0 && W > 0 && T > 0, "reconstruct: non-positive dimension");
  const size_t need = h ? deqsci_reconstruct_workspace_bytes(h, B, H, W, T, o->m)
                        : deqsci_adjoint_solve_workspace_bytes(B, H, W, T, o->m);
  DEQSCI_CHECK_ARG(need != 0, "reconstruct: unsupported shape B=%d H=%d W=%d T=%d", B, H, W, T);
  if (workspace_bytes < need) {
    set_error("reconstruct: workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
    return DEQSCI_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int m = o->m;
  const long long N = (long long)H * W * T;
  const Layout L = layout(h, B, H, W, T, m);
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up_sz(reinterpret_cast<uintptr_t>(workspace), 1024));
  float* fl = reinterpret_cast<float*>(base);
  float* X = fl;
  float* F = X + (size_t)m * B * N;
  float* G = F + (size_t)m * B * N;
  float* gram = G + (size_t)m * B * N;
  float* alpha = gram + L.gram_floats;
  float* scratch = alpha + L.alpha_floats;
  float* res_dev = scratch + L.scratch_floats;
  float* bn_backup = res_dev + 64;
  float* tmp_cube = bn_backup + (size_t)kMaxBnLayers * 2 * kHidden;
  const int n_layers = denoiser_num_layers(h);
  void* den_ws = base + align_up_sz(L.floats_total * sizeof(float), 1024);
  const size_t slot = (size_t)B * N;
  auto Xs = [&](int s) { return X + (size_t)s * slot; };
  auto Fs = [&](int s) { return F + (size_t)s * slot; };

  // sigma schedule of tag 'ffdnet' (solvers/equilibrium_solvers_yaping.py:409-413): fp32 multiply per call
  float sigma = o->sigma0, sigma_prev = o->sigma0;
  for (int i = 0; i < o->sigma_start_call; ++i) sigma = sigma * o->sigma_decay;
  int calls = 0;
  auto f_call = [&](const float* zin, float* zout) -> int {
    sigma_prev = sigma;
    sigma = sigma * o->sigma_decay;
    ++calls;
    if (masks) {
      // J_D^T is linear: evaluate it on vjp_scale * v (a power of two chosen by the caller so the operand planes sit
      // in fp16's normal range -- loss gradients are ~1e-7 per element at full size, below fp16's subnormal step) and
      // scale back; the Anderson state itself keeps the reference's magnitudes (its lam * I term is scale-dependent)
      const float* vin = zin;
      if (vjp_scale != 1.f) {
        scale_copy_kernel<<<(unsigned)std::min<long long>(((long long)slot + 1023) / 1024, 148 * 8), 256, 0, st>>>(
            zout, zin, vjp_scale, (long long)slot);
        vin = zout;                                   // zout is free until the projector writes it below
      }
      const int rc_m = deqsci_denoise_residual_masked(h, vin, tmp_cube, den_ws, L.den_bytes, masks, B, H, W, T, stream);
      if (rc_m) return rc_m;
      if (vjp_scale != 1.f)
        scale_copy_kernel<<<(unsigned)std::min<long long>(((long long)slot + 1023) / 1024, 148 * 8), 256, 0, st>>>(
            tmp_cube, tmp_cube, 1.f / vjp_scale, (long long)slot);
      return deqsci_gap_vjp(tmp_cube, phi, phi_sum, adjoint_grad, zout, B, H, W, T, stream);
    }
    if (!h) return deqsci_gap_vjp(zin, phi, phi_sum, adjoint_grad, zout, B, H, W, T, stream);
    if (bn)
      return deqsci_iterate_train(h, zin, y, phi, phi_sum, sigma_prev, zout, den_ws, L.den_bytes, bn, momentum, eps, B,
                                  H, W, T, stream);
    return deqsci_iterate(h, zin, y, phi, phi_sum, sigma_prev, zout, den_ws, L.den_bytes, B, H, W, T, stream);
  };
  auto undo_call = [&]() { sigma = sigma_prev; --calls; };   // a speculative iteration does not advance the schedule

  int rc;
  DEQSCI_CUDA(cudaMemsetAsync(gram, 0, (L.gram_floats + L.alpha_floats) * sizeof(float), st));
  if (x0) DEQSCI_CUDA(cudaMemcpyAsync(Xs(0), x0, slot * sizeof(float), cudaMemcpyDeviceToDevice, st));
  else if (!h) DEQSCI_CUDA(cudaMemcpyAsync(Xs(0), adjoint_grad, slot * sizeof(float), cudaMemcpyDeviceToDevice, st));
  else if ((rc = deqsci_gap_adjoint(y, phi, Xs(0), B, H, W, T, stream))) return rc;      // initial_point = At(y)
  if ((rc = f_call(Xs(0), Fs(0)))) return rc;
  DEQSCI_CUDA(cudaMemcpyAsync(Xs(1), Fs(0), slot * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if ((rc = f_call(Xs(1), Fs(1)))) return rc;
  if ((rc = deqsci_anderson_update(X, F, G, gram, alpha, res_dev, scratch, B, m, N, 0, 1, o->lam, (float)o->res_eps, stream))) return rc;
  if ((rc = deqsci_anderson_update(X, F, G, gram, alpha, res_dev, scratch, B, m, N, 1, 2, o->lam, (float)o->res_eps, stream))) return rc;

  // residual ring in pinned memory, one event per iteration
  ResidualRing* ring = acquire_ring(o->max_iter);
  if (!ring) { set_error("reconstruct: pinned residual ring / events: %s", cudaGetErrorString(cudaGetLastError())); return DEQSCI_ERR_CUDA; }
  float* res_host = ring->host;
  std::vector<cudaEvent_t>& ev = ring->events;
  auto res_of = [&](int k) -> double {
    cudaEventSynchronize(ev[k]);
    return (double)res_host[4 * k + 1] / (o->res_eps + (double)res_host[4 * k + 2]);
  };
  int current_k = 0, stop_k = -1;
  rc = DEQSCI_OK;
  double min_sample = 1e300;          // smallest per-sample residual seen on the counted iterations
  for (int k = 2; k < o->max_iter && rc == DEQSCI_OK; ++k) {
    current_k = k;
    const int n = k < m ? k : m, s = k % m;
    if ((rc = deqsci_anderson_mix(X, F, alpha, B, m, N, s, n, o->beta, stream))) break;
    // train mode: iteration k may turn out speculative (k > 2) -- keep the running statistics it will update
    if (bn && k > 2 && (rc = bn_running_snapshot(bn, n_layers, bn_backup, 0, st))) break;
    if ((rc = f_call(Xs(s), Fs(s)))) break;
    if ((rc = deqsci_anderson_update(X, F, G, gram, alpha, res_dev, scratch, B, m, N, s, (k + 1 < m ? k + 1 : m),
                                     o->lam, (float)o->res_eps, stream))) break;
    if (cudaMemcpyAsync(res_host + 4 * k, res_dev, 4 * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaEventRecord(ev[k], st) != cudaSuccess) {
      set_error("reconstruct: residual copy / event failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = DEQSCI_ERR_CUDA;
      break;
    }
    if (k > 2) min_sample = std::min(min_sample, (cudaEventSynchronize(ev[k - 1]), (double)res_host[4 * (k - 1) + 3]));
    if (k > 2 && res_of(k - 1) < (double)o->tol) {      // iteration k was speculative
      undo_call();
      if (bn) rc = bn_running_snapshot(bn, n_layers, bn_backup, 1, st);
      stop_k = current_k = k - 1;
      break;
    }
  }
  double res = 0.0;
  if (rc == DEQSCI_OK && current_k >= 2) {
    res = res_of(current_k);
    min_sample = std::min(min_sample, (double)res_host[4 * current_k + 3]);
  }
  if (rc == DEQSCI_OK) {
    // z = f(z*): the reconstruction DEQFixedPoint.forward returns
    if (o->final_call) rc = f_call(Xs(current_k % m), out);
    else if (cudaMemcpyAsync(out, Xs(current_k % m), slot * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      rc = DEQSCI_ERR_CUDA;
  }
  cudaError_t e = cudaStreamSynchronize(st);     // the pinned ring and events are released below
  release_ring(ring);                            // after the synchronize: nothing in flight references it
  if (rc != DEQSCI_OK) return rc;
  if (e != cudaSuccess) { set_error("reconstruct: %s", cudaGetErrorString(e)); return DEQSCI_ERR_CUDA; }
  result->residual = res;
  result->iterations = current_k;
  result->f_calls = calls;
  result->converged = stop_k >= 0 ? 1 : 0;
  result->sigma_next = sigma;
  result->min_sample_residual = min_sample < 1e299 ? min_sample : res;
  (void)stop_k;
  return DEQSCI_OK;
}
}  // namespace

extern "C" int deqsci_reconstruct(const deqsci_denoiser* h, const float* y, const float* phi, const float* phi_sum,
                                  const float* x0, float* out, const deqsci_solver_opts* o, void* workspace,
                                  size_t workspace_bytes, deqsci_solver_result* result, int B, int H, int W, int T,
                                  void* stream) {
  DEQSCI_CHECK_ARG(h != nullptr, "reconstruct: null denoiser handle");
  return anderson_loop(h, y, phi, phi_sum, x0, out, o, nullptr, 0.f, 0.f, nullptr, workspace, workspace_bytes, result,
                          B, H, W, T, stream);
}

extern "C" int deqsci_reconstruct_train(const deqsci_denoiser* h, const float* y, const float* phi,
                                        const float* phi_sum, const float* x0, float* out,
                                        const deqsci_solver_opts* o, const deqsci_bn_params* bn, float momentum,
                                        float eps, void* workspace, size_t workspace_bytes,
                                        deqsci_solver_result* result, int B, int H, int W, int T, void* stream) {
  DEQSCI_CHECK_ARG(bn != nullptr, "reconstruct_train: null BatchNorm table");
  DEQSCI_CHECK_ARG(h != nullptr, "reconstruct_train: null denoiser handle");
  return anderson_loop(h, y, phi, phi_sum, x0, out, o, bn, momentum, eps, nullptr, workspace, workspace_bytes, result,
                          B, H, W, T, stream);
}

extern "C" int deqsci_adjoint_solve(const float* grad, const float* phi, const float* phi_sum, float* out,
                                    const deqsci_solver_opts* o, void* workspace, size_t workspace_bytes,
                                    deqsci_solver_result* result, int B, int H, int W, int T, void* stream) {
  DEQSCI_CHECK_ARG(grad != nullptr && o != nullptr, "adjoint_solve: null pointer");
  deqsci_solver_opts opts = *o;
  opts.final_call = 0;                 // the reference returns the solver's iterate, not one more evaluation
  return anderson_loop(nullptr, nullptr, phi, phi_sum, /*x0=*/grad, out, &opts, nullptr, 0.f, 0.f, grad, workspace,
                          workspace_bytes, result, B, H, W, T, stream);
}

extern "C" int deqsci_adjoint_solve_denoiser(const deqsci_denoiser* h_adjoint, const void* const* masks_host,
                                             const float* grad, const float* phi, const float* phi_sum, float* out,
                                             const deqsci_solver_opts* o, float vjp_scale, void* workspace,
                                             size_t workspace_bytes, deqsci_solver_result* result, int B, int H, int W,
                                             int T, void* stream) {
  DEQSCI_CHECK_ARG(h_adjoint != nullptr && masks_host != nullptr && grad != nullptr && o != nullptr,
                   "adjoint_solve_denoiser: null pointer");
  DEQSCI_CHECK_ARG(vjp_scale > 0.f && vjp_scale < 3.0e38f, "adjoint_solve_denoiser: vjp_scale must be a positive finite number");
  deqsci_solver_opts opts = *o;
  opts.final_call = 0;
  return anderson_loop(h_adjoint, nullptr, phi, phi_sum, /*x0=*/grad, out, &opts, nullptr, 0.f, 0.f, grad, workspace,
                       workspace_bytes, result, B, H, W, T, stream, masks_host, vjp_scale);
}
