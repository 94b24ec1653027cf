/* Plain-C host program against include/deqsci.h: proves the boundary is a C-ABI with raw pointers.
 * Builds a 4-layer DnCNN plan from deterministic pseudo-random weights, reconstructs a small
 * synthetic measurement with deqsci_reconstruct and prints the residual and a checksum; the GPU test
 * (tests/test_gpu_parity.py::test_c_abi_from_plain_c) runs the same problem through the Python host
 * mirror and compares.   cc smoke.c -I include -L deqsci_b200 -ldeqsci -lcudart */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "deqsci.h"

static unsigned int lcg_state = 12345u;
static float lcg(void) { /* uniform in [-0.5, 0.5) */
  lcg_state = lcg_state * 1664525u + 1013904223u;
  return (float)(lcg_state >> 8) / 16777216.0f - 0.5f;
}
#define CK(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, deqsci_last_error()); return 1; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(void) {
  const int B = 2, H = 16, W = 144, T = 8, NL = 4;
  const size_t n_cube = (size_t)B * H * W * T, n_snap = (size_t)B * H * W;
  const int cin[4] = {1, 64, 64, 64}, cout[4] = {64, 64, 64, 1};
  float* wts[4];
  deqsci_conv_layer layers[4];
  for (int l = 0; l < NL; ++l) {
    const size_t n = (size_t)cout[l] * cin[l] * 9;
    wts[l] = (float*)malloc(n * sizeof(float));
    for (size_t i = 0; i < n; ++i) wts[l][i] = lcg() * (l == 0 ? 0.5f : 0.08f);
    layers[l].cin = cin[l]; layers[l].cout = cout[l]; layers[l].relu = l < NL - 1;
    layers[l].weight_host = wts[l]; layers[l].scale_host = NULL; layers[l].bias_host = NULL;
  }
  deqsci_denoiser* plan = NULL;
  CK(deqsci_denoiser_create(DEQSCI_NET_DNCNN, DEQSCI_PREC_TC_SPLIT, NL, layers, &plan));

  float *x_h = (float*)malloc(n_cube * 4), *phi_h = (float*)malloc(n_cube * 4), *out_h = (float*)malloc(n_cube * 4);
  for (size_t i = 0; i < n_cube; ++i) { x_h[i] = lcg() + 0.5f; phi_h[i] = lcg() < 0.f ? 0.f : 1.f; }
  float *x_d, *phi_d, *y_d, *ps_d, *out_d;
  CU(cudaMalloc((void**)&x_d, n_cube * 4)); CU(cudaMalloc((void**)&phi_d, n_cube * 4)); CU(cudaMalloc((void**)&out_d, n_cube * 4));
  CU(cudaMalloc((void**)&y_d, n_snap * 4)); CU(cudaMalloc((void**)&ps_d, n_snap * 4));
  CU(cudaMemcpy(x_d, x_h, n_cube * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(phi_d, phi_h, n_cube * 4, cudaMemcpyHostToDevice));
  CK(deqsci_gap_forward(x_d, phi_d, y_d, B, H, W, T, NULL));          /* y = A(x) */
  CK(deqsci_phi_sum(phi_d, ps_d, B, H, W, T, NULL));

  deqsci_solver_opts o;
  o.m = 5; o.lam = 1e-2f; o.beta = 1.0f; o.max_iter = 12; o.tol = 1e-5f; o.sigma0 = 60.f / 255.f;
  o.sigma_decay = 0.971f; o.sigma_start_call = 0; o.final_call = 1; o.res_eps = 1e-5;
  deqsci_solver_result r;
  const size_t ws_bytes = deqsci_reconstruct_workspace_bytes(plan, B, H, W, T, o.m);
  void* ws;
  CU(cudaMalloc(&ws, ws_bytes));
  CK(deqsci_reconstruct(plan, y_d, phi_d, ps_d, NULL, out_d, &o, ws, ws_bytes, &r, B, H, W, T, NULL));
  CU(cudaMemcpy(out_h, out_d, n_cube * 4, cudaMemcpyDeviceToHost));
  double sum = 0.0, sq = 0.0;
  for (size_t i = 0; i < n_cube; ++i) { sum += out_h[i]; sq += (double)out_h[i] * out_h[i]; }
  printf("C_ABI_SMOKE residual=%.9e iterations=%d f_calls=%d sum=%.9e sumsq=%.9e\n", r.residual, r.iterations, r.f_calls, sum, sq);
  CK(deqsci_denoiser_destroy(plan));
  return 0;
}
