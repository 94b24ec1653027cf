/*
 * deqsci.h — C-ABI of libdeqsci (sm_100a): the DE-GAP reconstruction hot path of
 * IndigoPurple/DEQSCI, hand-written CUDA behind plain pointers.
 *
 * The reference has no FFI layer (it is pure Python/PyTorch); each entry point below
 * replaces the PyTorch call sites named in its comment (paths relative to the reference root).
 * The Python host mirror of the reference interface (deqsci_b200/) binds these with ctypes; the
 * stub a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *  - tensors are fp32, contiguous, in the reference's own layouts:
 *        cube  z, Phi : [B, H, W, T]   (T innermost)        snapshot y, Phi_sum : [B, H, W]
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - functions never allocate or free caller memory, never synchronise the stream (except
 *    deqsci_denoiser_create, which uploads weights synchronously) and return 0 on success or a
 *    negative deqsci_status; deqsci_last_error() gives the text for the calling thread;
 *  - handles are immutable after creation, so calls are re-entrant per (handle, stream) as long
 *    as each concurrent call has its own workspace.
 */
#ifndef DEQSCI_H_
#define DEQSCI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEQSCI_VERSION 100

typedef enum {
  DEQSCI_OK = 0,
  DEQSCI_ERR_INVALID = -1,   /* bad argument (null pointer, unsupported shape, ...)            */
  DEQSCI_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed                             */
  DEQSCI_ERR_WORKSPACE = -3, /* workspace too small                                             */
  DEQSCI_ERR_ARCH = -4       /* device is not sm_100 (tcgen05 path unavailable)                 */
} deqsci_status;

int deqsci_version(void);
const char* deqsci_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * (1) SCI operator / GAP data-consistency step — one fused, vectorised, memory-bound kernel each.
 * ---------------------------------------------------------------------------------------------- */

/* out[B,H,W] = sum_t x*Phi.            Replaces A_torch_, utils/cg_utils.py:85-90. */
int deqsci_gap_forward(const float* x, const float* phi, float* out,
                       int B, int H, int W, int T, void* stream);

/* out[B,H,W,T] = y[...,None]*Phi.      Replaces At_torch_ / initial_point, utils/cg_utils.py:124-129,228-229. */
int deqsci_gap_adjoint(const float* y, const float* phi, float* out,
                       int B, int H, int W, int T, void* stream);

/* out[B,H,W] = sum_t Phi, zeros -> 1.  Replaces training/sci_equilibrium_training.py:61-62,162-163. */
int deqsci_phi_sum(const float* phi, float* out, int B, int H, int W, int T, void* stream);

/* out = z + Phi * ((y - sum_t z*Phi) / phi_sum)[...,None].
 * Replaces solvers/equilibrium_solvers_yaping.py:399-400 (A, subtract, divide, At, add: 6 launches). */
int deqsci_gap_step(const float* z, const float* y, const float* phi, const float* phi_sum,
                    float* out, int B, int H, int W, int T, void* stream);

/* out = v - Phi * ((sum_t v*Phi) / phi_sum)[...,None]  (+ add[B,H,W,T] if add != NULL).
 * Vector-Jacobian product of deqsci_gap_step w.r.t. z; for tag 'ffdnet' it is the whole VJP of the
 * iterate map (the denoiser input is detached, networks/ffdnet/models.py:103-104), used by the
 * backward solve at solvers/new_equilibrium_utils_yaping.py:274-277. */
int deqsci_gap_vjp(const float* v, const float* phi, const float* phi_sum, const float* add,
                   float* out, int B, int H, int W, int T, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (2) Learned denoiser: 3x3 conv stack (pad 1, stride 1, no conv bias), 64 hidden features.
 * ---------------------------------------------------------------------------------------------- */

typedef struct deqsci_denoiser deqsci_denoiser; /* opaque */

typedef enum {
  DEQSCI_NET_FFDNET = 0, /* networks/ffdnet/models.py:70-108: pixel-unshuffle + noise map (5 ch) ->
                            convs at H/2 x W/2 -> 4 ch -> pixel-shuffle                              */
  DEQSCI_NET_DNCNN = 1   /* networks/provable/model/SimpleCNN_models.py:6-61: 1 ch -> convs -> 1 ch  */
} deqsci_net_kind;

typedef enum {
  DEQSCI_PREC_TC_SPLIT = 0, /* hidden 64->64 layers on tcgen05 tensor cores, fp16 hi/lo operand
                               split (3 MMA products, fp32 TMEM accumulation) — the parity mode     */
  DEQSCI_PREC_FP32 = 1,     /* hidden layers on CUDA cores in fp32 (slow; validation)               */
  DEQSCI_PREC_TC_SINGLE = 2 /* single fp16 MMA pass (fast, NOT parity: reported as such)            */
} deqsci_precision;

/* One conv layer: weight_host [cout,cin,3,3] fp32 (PyTorch OIHW); optional per-output-channel
 * affine applied after the conv (folded eval-mode BatchNorm, networks/ffdnet/models.py:58):
 * out = conv*scale + bias; then ReLU if relu != 0. scale_host/bias_host may be NULL (= 1 / 0). */
typedef struct {
  int cin, cout, relu;
  const float* weight_host;
  const float* scale_host;
  const float* bias_host;
} deqsci_conv_layer;

/* Builds the device-side plan (packs / splits weights). layers[0].cin must be 5 (FFDNET) or 1
 * (DNCNN), hidden layers 64->64, last layer cout 4 (FFDNET) or 1 (DNCNN).  Synchronous. */
int deqsci_denoiser_create(int net_kind, int precision, int num_layers,
                           const deqsci_conv_layer* layers_host, deqsci_denoiser** out);
int deqsci_denoiser_destroy(deqsci_denoiser* h);

/* Refreshes the conv weights of an existing plan from DEVICE tensors (fp32 [cout][cin][3][3] each, e.g.
 * the parameters an optimizer step just updated in place), stream-ordered on `stream`: no host round
 * trip, no synchronisation, no allocation after the first call (which builds the gather maps).  Folded
 * scale/bias are left as they are, so this is for plans whose BatchNorm is not folded (train plans) or
 * absent.  The caller orders it against launches that use the plan on other streams. */
int deqsci_denoiser_update_weights(deqsci_denoiser* h, int num_layers, const float* const* weight_dev,
                                   void* stream);

/* Bytes of scratch (activation ping-pong planes) for a [B,H,W,T] cube. */
size_t deqsci_denoiser_workspace_bytes(const deqsci_denoiser* h, int B, int H, int W, int T);

/* out[B,H,W,T] = zin - D(zin) where D runs on the B*T frames zin[b,:,:,t] (frame index b*T+t) and,
 * for FFDNET, the constant noise map sigma (same for every frame of the call).
 * Replaces the permute/view + net(...) + `z - noise.view().permute()` of
 * solvers/equilibrium_solvers_yaping.py:415-420 and everything under networks/ffdnet/models.py:98-108.
 * `out` may alias `zin`. */
int deqsci_denoise_residual(const deqsci_denoiser* h, const float* zin, float sigma, float* out,
                            void* workspace, size_t workspace_bytes,
                            int B, int H, int W, int T, void* stream);

/* The whole iterate map f(z) = denoise_residual(gap_step(z)):
 * EquilibriumProxGradSCI.forward, solvers/equilibrium_solvers_yaping.py:396-436.
 * `out` may alias `z`. */
int deqsci_iterate(const deqsci_denoiser* h, const float* z, const float* y, const float* phi,
                   const float* phi_sum, float sigma, float* out,
                   void* workspace, size_t workspace_bytes,
                   int B, int H, int W, int T, void* stream);

/* The iterate map with the denoiser in TRAIN mode (nn.Module.train(): batch-statistics BatchNorm,
 * as the reference runs its forward solve while training).  The plan must have been
 * created WITHOUT folded BatchNorm (scale_host = bias_host = NULL on the BatchNorm layers).
 * bn_host[i] describes the BatchNorm2d that follows conv layer i (all NULL = none): device pointers to
 * gamma, beta (may be NULL = 1 / 0) and to running_mean / running_var, which are UPDATED in place with
 * `momentum` exactly once per call, like PyTorch (unbiased variance; the caller increments
 * num_batches_tracked).  Per BatchNorm layer: conv with per-channel sum / sum-of-squares epilogue,
 * a finalize kernel, an in-place normalise + ReLU pass.  Needs precision TC_SPLIT and conv images wider
 * than 64 pixels. */
typedef struct {
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
} deqsci_bn_params;

int deqsci_iterate_train(const deqsci_denoiser* h, const float* z, const float* y, const float* phi,
                         const float* phi_sum, float sigma, float* out,
                         void* workspace, size_t workspace_bytes,
                         const deqsci_bn_params* bn_host, float momentum, float eps,
                         int B, int H, int W, int T, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (3) Anderson acceleration state update (andersonexp / anderson,
 *     solvers/new_equilibrium_utils_yaping.py:114-189).  History X, F, G : [m, B, N] fp32 — slot-major,
 *     so that slot s is a contiguous [B,H,W,T] cube the iterate map reads / writes in place (the
 *     reference's [B,m,N] buffer is private to andersonexp, :159-160).
 * ---------------------------------------------------------------------------------------------- */

/* Number of floats of `scratch` needed by deqsci_anderson_update. */
size_t deqsci_anderson_scratch_floats(int B, int m, long long N);

/* After F[slot] = f(X[slot]) has been written, with n valid slots (0..n-1, slot < n):
 *   G[slot]     = F[slot] - X[slot]                                     (:177)
 *   gram[b,slot,j] = gram[b,j,slot] = <G[slot,b], G[j,b]>  for j < n     (:178, incremental row)
 *   alpha[b,0:n] = solution[1:n+1] of the bordered system [[0,1^T],[1,GG^T+lam I]] a = e0   (:168-171,180)
 *                  (LU with partial pivoting, fp32, one lane per sample)
 *   res[0] = ||G[slot]|| / (res_eps + ||F[slot]||) over the whole batch  (:184);  res[1], res[2] = the two norms
 * gram [B,m,m], alpha [B,m], res [4] are device buffers owned by the caller.  Two launches. */
int deqsci_anderson_update(const float* X, const float* F, float* G, float* gram, float* alpha,
                           float* res, float* scratch, int B, int m, long long N, int slot, int n,
                           float lam, float res_eps, void* stream);

/* X[slot,b] = beta * sum_{j<n} alpha[b,j] F[j,b] + (1-beta) * sum_{j<n} alpha[b,j] X[j,b]   (:182).
 * When beta == 1 the X term is not read (0*X dropped: differs from the reference only if X holds
 * NaN/Inf). */
int deqsci_anderson_mix(float* X, const float* F, const float* alpha, int B, int m, long long N,
                        int slot, int n, float beta, void* stream);

/* res[0] = ||a-b|| / (res_eps + ||a||), res[1] = ||a-b||, res[2] = ||a|| over all `count` floats
 * (forward_iteration, solvers/new_equilibrium_utils_yaping.py:219).  scratch as above (B=1,N=count). */
int deqsci_residual(const float* a, const float* b, float* res, float* scratch, long long count,
                    float res_eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (4) Whole reconstruction in one call: the device-resident driver.
 *     z* = andersonexp(f, x0, m, lam, max_iter, tol, beta); out = f(z*)        with f = deqsci_iterate
 *     = DEQFixedPoint.forward at inference (solvers/new_equilibrium_utils_yaping.py:248-268; the
 *     reference's second post-solver call only feeds the backward hook and is not run).
 *     Queues the library's kernels from a C++ loop; residuals come back through pinned memory one
 *     iteration behind, so the device never waits for the host.  Synchronises `stream` before returning.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int m;               /* Anderson history (2..8); entry script: 5                                   */
  float lam;           /* regulariser of the bordered system; entry script: 1e-2                       */
  float beta;          /* mixing; entry script: 1.0                                                    */
  int max_iter;        /* and_maxiters                                                                 */
  float tol;           /* stop when ||F-X|| / (res_eps + ||F||) < tol (whole batch); entry script 1e-5 */
  float sigma0;        /* FFDNet noise level of the first call after a reset: 60/255                   */
  float sigma_decay;   /* 0.971 per call (fp32 multiply); both ignored by DnCNN plans                  */
  int sigma_start_call;/* schedule position of the first f call of this solve (0 = fresh measurement)  */
  int final_call;      /* 1: out = f(z*) (the reconstruction), 0: out = z*                             */
  double res_eps;      /* 1e-5 in andersonexp (a python double there, so a double here)                */
} deqsci_solver_opts;

typedef struct {
  double residual;     /* last residual tested (python float `forward_res` of the reference)          */
  int iterations;      /* last solver iteration index k                                                */
  int f_calls;         /* iterate-map evaluations that count towards the sigma schedule                */
  int converged;       /* 1 if the tolerance stopped the loop                                          */
  float sigma_next;    /* sigma the next call would use                                                */
  double min_sample_residual; /* smallest PER-SAMPLE residual ||F_b-X_b||/(eps+||F_b||) over the counted iterations:
                          below tol means a sample solved on its own (the reference's batch 1) would have stopped
                          earlier than the whole-batch test did                                                   */
} deqsci_solver_result;

size_t deqsci_reconstruct_workspace_bytes(const deqsci_denoiser* h, int B, int H, int W, int T, int m);

/* y, phi_sum [B,H,W]; phi, x0, out [B,H,W,T]; x0 may be NULL (= At(y, Phi), utils/cg_utils.py:228-229). */
int deqsci_reconstruct(const deqsci_denoiser* h, const float* y, const float* phi, const float* phi_sum,
                       const float* x0, float* out, const deqsci_solver_opts* opts,
                       void* workspace, size_t workspace_bytes, deqsci_solver_result* result,
                       int B, int H, int W, int T, void* stream);

/* The same loop with the denoiser in TRAIN mode (every f call = deqsci_iterate_train): the no_grad
 * forward solve DEQFixedPoint.forward runs while training (solvers/new_equilibrium_utils_yaping.py:265-266).
 * `h` must be a train plan, `bn` as for deqsci_iterate_train; running statistics receive one momentum
 * update per counted f call (result->f_calls; a speculative iteration's update is rolled back).
 * Normally called with opts->final_call = 0: the caller's graph-attached f(z*) follows. */
int deqsci_reconstruct_train(const deqsci_denoiser* h, const float* y, const float* phi, const float* phi_sum,
                             const float* x0, float* out, const deqsci_solver_opts* opts,
                             const deqsci_bn_params* bn, float momentum, float eps,
                             void* workspace, size_t workspace_bytes, deqsci_solver_result* result,
                             int B, int H, int W, int T, void* stream);

/* The backward fixed-point solve of the implicit-differentiation hook for tag 'ffdnet'
 * (solvers/new_equilibrium_utils_yaping.py:274-277): andersonexp on g -> VJP_f(g) + grad starting at grad,
 * where VJP_f is the GAP projector (deqsci_gap_vjp) because FFDNet detaches its input.  out = the solver's
 * last iterate (what the reference's hook returns as the gradient w.r.t. z); result->residual is
 * `backward_res`.  opts: m, lam, beta, max_iter, tol, res_eps (sigma fields and final_call are ignored). */
size_t deqsci_adjoint_solve_workspace_bytes(int B, int H, int W, int T, int m);
int deqsci_adjoint_solve(const float* grad, const float* phi, const float* phi_sum, float* out,
                         const deqsci_solver_opts* opts, void* workspace, size_t workspace_bytes,
                         deqsci_solver_result* result, int B, int H, int W, int T, void* stream);

/* The implicit-differentiation hook for tag 'denoiser' (DE-GAP-CNN, solvers/new_equilibrium_utils_yaping.py:271-277):
 * the reference evaluates f0 = f(z0) with a graph and runs andersonexp on g -> autograd.grad(f0, z0, g) + grad, i.e.
 * one full backward pass of the conv stack per solver iteration.  Here:
 *  - deqsci_iterate_save is deqsci_iterate that also keeps every hidden activation (deqsci_saved_forward below):
 *    save->acts[i] (i = 0 .. num_layers-2, device buffers of deqsci_denoiser_activation_bytes() each) receives the
 *    output planes of conv layer i ([hi plane | lo plane], channels-last [B*T,Hc,Wc,64] fp16);
 *  - the VJP of the conv / ReLU stack is the SAME kernels run on an ADJOINT plan -- created with
 *    deqsci_denoiser_create from the layers in reverse order, weights transposed and flipped
 *    (W'[c][o][ky][kx] = W[o][c][2-ky][2-kx]) -- whose ReLUs are replaced by the sign of the saved activations:
 *    masks_host[i] gates the output of adjoint layer i (= acts of forward layer num_layers-2-i).
 *    deqsci_denoise_residual_masked computes out = v - J_D^T v;
 *  - deqsci_adjoint_solve_denoiser is the whole backward solve: andersonexp on g -> gap_vjp(g - J_D^T g) + grad
 *    (workspace: deqsci_reconstruct_workspace_bytes(h_adjoint, ...)).  vjp_scale: a power of two (1 = none); J_D^T is
 *    evaluated on vjp_scale * g and scaled back, so that loss gradients of ~1e-7 per element sit in fp16's normal
 *    range inside the conv kernels; the solver state keeps its own magnitudes (lam * I is scale-dependent).
 * Plain conv / ReLU stacks (no folded BatchNorm), precision TC_SPLIT, conv images wider than 64 pixels. */
/* What a forward call keeps for the backward passes (all device buffers, owned by the caller; the tables themselves
 * are host arrays with one entry per conv layer 0 .. num_layers-2):
 *   acts[i]    output planes of conv layer i AFTER BatchNorm / ReLU (deqsci_denoiser_activation_bytes() each);
 *   pre[i]     train mode only: the raw conv output planes of a layer followed by BatchNorm (NULL entries / NULL
 *              table otherwise);
 *   bn_record  train mode only: 256 floats per conv layer: scale, shift, batch mean, 1/sqrt(var+eps) (64 each);
 *   zprime     [B,T,H,W] floats: the network's input frames z' (frame-planar), the first layer's wgrad input
 *              (may be NULL when only the activations are wanted). */
typedef struct {
  void* const* acts;
  void* const* pre;
  float* bn_record;
  float* zprime;
} deqsci_saved_forward;

size_t deqsci_denoiser_activation_bytes(const deqsci_denoiser* h, int B, int H, int W, int T);
int deqsci_iterate_save(const deqsci_denoiser* h, const float* z, const float* y, const float* phi,
                        const float* phi_sum, float sigma, float* out, void* workspace, size_t workspace_bytes,
                        const deqsci_saved_forward* save, int B, int H, int W, int T, void* stream);
/* deqsci_iterate_train (batch-statistics BatchNorm, running statistics updated) that keeps the same things. */
int deqsci_iterate_train_save(const deqsci_denoiser* h, const float* z, const float* y, const float* phi,
                              const float* phi_sum, float sigma, float* out, void* workspace, size_t workspace_bytes,
                              const deqsci_bn_params* bn_host, float momentum, float eps,
                              const deqsci_saved_forward* save, int B, int H, int W, int T, void* stream);

/* Weight gradients of that call (csrc/backward.cu): the backward pass the reference runs through autograd / cuDNN for
 * its ONE graph-attached f evaluation (solvers/new_equilibrium_utils_yaping.py:268; loss.backward()).  `grad` [B,H,W,T]
 * is dL/d(out); the gradient w.r.t. the iterate itself is not produced (the reference's z* is detached).  Per conv
 * layer i: d_weight[i] <- dL/dW_i ([cout,cin,3,3] fp32); for a layer followed by train-mode BatchNorm (save->pre[i]
 * non-NULL) gamma[i] is its weight (NULL = 1) and d_gamma[i] / d_beta[i] receive its parameter gradients.
 * dgrad = the tensor-core conv kernels on the ADJOINT plan `adjoint` (layers reversed, weights transposed and flipped;
 * its LAST layer is never run and may hold zeros); wgrad, ReLU and BatchNorm backward = CUDA-core kernels with
 * two-stage fixed-order reductions.  grad_scale: power of two applied to `grad` on entry and divided out of every
 * result (the gradient planes are fp16 pairs; deeper layers are re-scaled on the device).  sigma as in the forward call.
 * Needs precision TC_SPLIT and conv images wider than 64 pixels. */
size_t deqsci_backward_workspace_bytes(const deqsci_denoiser* h, int B, int H, int W, int T);
int deqsci_backward_weights(const deqsci_denoiser* h, const deqsci_denoiser* adjoint, const deqsci_saved_forward* save,
                            const float* const* gamma, const float* grad, float grad_scale, float sigma,
                            float* const* d_weight, float* const* d_gamma, float* const* d_beta,
                            void* workspace, size_t workspace_bytes, int B, int H, int W, int T, void* stream);
int deqsci_denoise_residual_masked(const deqsci_denoiser* h_adjoint, const float* vin, float* out,
                                   void* workspace, size_t workspace_bytes, const void* const* masks_host,
                                   int B, int H, int W, int T, void* stream);
int deqsci_adjoint_solve_denoiser(const deqsci_denoiser* h_adjoint, const void* const* masks_host,
                                  const float* grad, const float* phi, const float* phi_sum, float* out,
                                  const deqsci_solver_opts* opts, float vjp_scale, void* workspace,
                                  size_t workspace_bytes, deqsci_solver_result* result, int B, int H, int W, int T,
                                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gradient exchange of the training step fused with the optimizer (csrc/optim.cu).  The reference has no
 * distributed backend; this is the one exchange step data-parallel training adds between loss.backward() and
 * optimizer.step() (training/sci_equilibrium_training.py:66-75; Adam(lr) built at video_sci_proxgrad.py:201).
 *
 * Every rank owns a communication buffer: [n_floats gradient | barrier flags], allocated by deqsci_comm_alloc
 * and exported as a 64-byte CUDA IPC handle; peers map it with deqsci_comm_open.  deqsci_adam_allreduce_step
 * launches ONE kernel: cross-GPU barrier, g = grad_scale * sum_r grad_r (P2P loads over NVLink, summed in rank
 * order so every rank gets identical bits), torch.optim.Adam update (no weight decay / amsgrad) of this rank's
 * params / exp_avg / exp_avg_sq, closing barrier.  comm_bases_host[r] = base of rank r's buffer as mapped in
 * THIS process (own allocation for r = rank).  world = 1 (or a gradient already reduced by the caller, with
 * comm_bases_host = {own}) is the plain fused Adam step.  `epoch` must grow by one per call on every rank
 * (1, 2, ...); `step` is Adam's bias-correction step.  All ranks must make the call with the same n_floats.
 * ---------------------------------------------------------------------------------------------- */
size_t deqsci_comm_bytes(long long n_floats);
int deqsci_comm_alloc(long long n_floats, void** dev_ptr, void* ipc_handle_64);
int deqsci_comm_open(const void* ipc_handle_64, void** dev_ptr);
int deqsci_comm_close(void* peer_ptr);
int deqsci_comm_free(void* dev_ptr);
/* error_host = 1 when a barrier of an earlier step timed out (a peer never arrived); synchronous copy. */
int deqsci_comm_error(const void* comm_base, long long n_floats, int* error_host);
int deqsci_adam_allreduce_step(float* params, float* exp_avg, float* exp_avg_sq,
                               const void* const* comm_bases_host, int rank, int world, long long n_floats,
                               float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                               unsigned epoch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Launch accounting and sampled device timing of the library's own kernels.
 * Kernel classes: 0 gap, 1 conv_first, 2 conv_hidden, 3 conv_last, 4 anderson_gram,
 * 5 anderson_solve, 6 anderson_mix  (DEQSCI_PROFILE_KINDS entries in every array below).
 * deqsci_profile_begin resets the counters; with sample_every = k > 0 every k-th launch is bracketed
 * by a pair of CUDA events on its own stream.  deqsci_profile_end synchronises the device and
 * returns, per class, the summed milliseconds of the sampled launches, how many were sampled and
 * how many were launched.
 * ---------------------------------------------------------------------------------------------- */
#define DEQSCI_PROFILE_KINDS 7
int deqsci_profile_begin(int sample_every);
int deqsci_profile_end(double* ms_sum, long long* n_sampled, long long* n_launched);

/* ------------------------------------------------------------------------------------------------
 * Testing hook: runs ONE hidden 64->64 layer (index `layer`, 0 < layer < num_layers-1) of a plan on
 * caller-provided activation planes: channels-last [NF,Hc,Wc,64], fp16 hi plane followed by the
 * fp16 lo plane (value = hi + lo*2^-11).  Used by tests/ to compare the tcgen05 kernel against the
 * fp32 CUDA-core kernel layer by layer.
 * ---------------------------------------------------------------------------------------------- */
int deqsci_debug_hidden_layer(const deqsci_denoiser* h, int layer, const void* act_in, void* act_out,
                              int NF, int Hc, int Wc, void* stream);

/* Testing hook (host only, no device work): the strip height the CTA-pair hidden kernel picks for NF frames of
 * Hc x Wc conv pixels on a device with `num_sms` SMs -- the cost model of csrc/tma_host.cu, so the CPU test
 * suite can pin its choices. */
int deqsci_debug_pair_strip_rows(int NF, int Hc, int Wc, int num_sms);

/* Testing hook (host only, no device work): the gather map of the CTA-pair hidden kernel's weight image for this
 * process's issue mode (DEQSCI_TC_RS; csrc/conv_tc2.cu tc2_layout): map[e] = 4 * index into w[64][64][3][3] + kind
 * (0 = fp16 hi half, 1 = scaled lo half) for every fp16 element e of the image [2 ranks][9 taps][rows][64].  Returns
 * the number of elements; map may be NULL (or count too small) to query it. */
long long deqsci_debug_pair_weight_map(int* map, long long count);

#ifdef __cplusplus
}
#endif
#endif /* DEQSCI_H_ */
