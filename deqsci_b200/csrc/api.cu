// C-ABI glue: error text, the denoiser plan (weight packing / upload) and the layer sequencing of
// deqsci_denoise_residual / deqsci_iterate.  See include/deqsci.h for the contract.
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace deqsci {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// launchers implemented in conv_cc.cu / conv_tc.cu / gap.cu
int conv_first_launch(int kind, bool fuse_gap, const float* zin, const float* y, const float* phi,
                      const float* phi_sum, float* zprime_out, float sigma, const float* wpack,
                      const float* scale, const float* bias, int relu, __half* act_out, long long plane_elems,
                      int B, int H, int W, int T, cudaStream_t st);
int conv_mid_fp32_launch(const __half* act_in, __half* act_out, long long plane_elems, const float* wpack,
                         const float* scale, const float* bias, int relu, int NF, int Hc, int Wc, cudaStream_t st);
int conv_last_launch(int kind, const __half* act_in, long long plane_elems, const float* wpack,
                     const float* scale, const float* bias, int relu, const float* zprime, float* out, int B,
                     int H, int W, int T, cudaStream_t st);
int conv_mid_tc_launch(bool split, const __half* act_in, __half* act_out, long long plane_elems,
                       const uint8_t* wimg, const float* scale, const float* bias, int relu, int NF, int Hc, int Wc,
                       cudaStream_t st);
int conv_tc_launch(int mode, bool split, const __half* act_in, __half* act_out, long long plane_elems,
                   const uint8_t* wimg, const float* scale, const float* bias, int relu, int NF, int Hc, int Wc,
                   const float* zprime, float* out_cube, int H, int W, int T, cudaStream_t st);
int gap_prep_launch(int kind, const float* z, const float* y, const float* phi, const float* phi_sum,
                    float* zprime_out, __half* planes, long long plane_elems, float sigma, int B, int H, int W,
                    int T, bool do_gap, bool zprime_planar, cudaStream_t st);
size_t tcf_weight_image_bytes();
void tcf_pack_weights(const float* w, int cin, uint8_t* img);
void tcf_pack_map(int cin, int32_t* map);
bool tcf_supported(int Wc);
int conv_first_tc_launch(const __half* planes_in, long long in_plane_elems, __half* act_out, long long plane_elems,
                         const uint8_t* wimg, const float* scale, const float* bias, int relu, int NF, int Hc,
                         int Wc, cudaStream_t st, const __half* mask = nullptr);
size_t tcl_weight_image_bytes();
void tcl_pack_weights(const float* w, int cout, uint8_t* img);
void tcl_pack_map(int cout, int32_t* map);
bool tcl_supported(int Wc);
int conv_last_tc_launch(int cout, const __half* act_in, long long plane_elems, const uint8_t* wimg,
                        const float* scale, const float* bias, int relu, int NF, int Hc, int Wc,
                        const float* zprime, bool zprime_planar, float* out_cube, int H, int W, int T, cudaStream_t st);
size_t tc2_weight_image_bytes();
void tc2_pack_weights(const float* w, uint8_t* img);
void tc2_pack_map(int32_t* map);
bool tc2_supported(int Hc, int Wc);
int conv_hidden_2cta_launch(const __half* act_in, __half* act_out, long long plane_elems, const uint8_t* wimg,
                            const float* scale, const float* bias, int relu, int NF, int Hc, int Wc,
                            cudaStream_t st, double* stats = nullptr, const __half* mask = nullptr);
// backward.cu
size_t backward_scratch_floats();
int act_bwd_launch(const __half* g, const __half* act, const __half* pre, const float* rec, const float* gamma,
                   __half* out, long long plane_elems, long long count, float* scratch, float* d_gamma, float* d_beta,
                   cudaStream_t st);
int wgrad_hidden_launch(const __half* a, const __half* d, long long plane_elems, int NF, int Hc, int Wc, float* scratch,
                        float* d_weight, cudaStream_t st);
int wgrad_last_launch(int cout, const __half* a, long long plane_elems, const float* gsc, int B, int H, int W, int T,
                      float* scratch, float* d_weight, cudaStream_t st);
int wgrad_first_launch(int cin, const __half* d, long long plane_elems, const float* zp, float sigma, int NF, int H, int W,
                       float* scratch, float* d_weight, cudaStream_t st);
int backward_begin(float* scratch, const float* g, float* gsc, float scale, long long n, cudaStream_t st);
bool tc2_chain_supported(int NF, int Hc, int Wc, int n_layers);
size_t tc2_chain_flag_count(int NF, int Hc);
int conv_hidden_chain_launch(__half* buf0, __half* buf1, long long plane_elems, int n_layers, const uint8_t* const* wimg,
                             const float* const* scale, const float* const* bias, const int* relu, int NF, int Hc, int Wc,
                             uint32_t* flags, uint32_t* epoch, cudaStream_t st);
int bn_train_launch(__half* act, long long plane_elems, const double* stats, int n_partials, float* scale_shift,
                    const float* gamma,
                    const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                    long long count, int relu, cudaStream_t st, const __half* src = nullptr, float* record = nullptr);
size_t tc_weight_image_bytes(bool split, int cout);
void tc_pack_weights(const float* w, int cout, bool split, uint8_t* img);
void tc_pack_map(int cout, bool split, int32_t* map);

struct Layer {
  int cin = 0, cout = 0, relu = 0;
  float* w_cc = nullptr;      // CUDA-core packing [9][cin][cout] fp32
  uint8_t* w_tc = nullptr;    // tcgen05 shared-memory image (hidden and last layers, TC modes)
  uint8_t* w_tc2 = nullptr;   // second image: CTA-pair layout (hidden layers, split precision) or the
                              // ky-transposed layout (last layer)
  float* scale = nullptr;     // [cout] or null
  float* bias = nullptr;      // [cout] or null
  // gather maps of the device-side repack (owned by the handle, shared between layers of one shape)
  const int32_t *map_cc = nullptr, *map_tc = nullptr, *map_tc2 = nullptr;
  int n_cc = 0, n_tc = 0, n_tc2 = 0;      // elements (fp32 for w_cc, fp16 for the images)
};

}  // namespace deqsci

// per-stream state of the chained hidden-layer kernel (conv_tc2.cu): strip-ready flags + the launch epoch they count from
struct ChainState {
  uint32_t* flags = nullptr;
  size_t count = 0;
  uint32_t epoch = 0;
};

struct deqsci_denoiser {
  int kind = 0, precision = 0;
  std::vector<deqsci::Layer> layers;
  std::map<std::string, int32_t*> maps;   // device gather maps, built on the first update_weights call
  mutable std::map<cudaStream_t, ChainState> chain;     // launches on one stream are ordered: they may share flags
  mutable std::mutex chain_mutex;
};

using namespace deqsci;

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// train-mode BatchNorm scratch at the end of the workspace: per-CTA statistics partials + scale/shift
constexpr int kMaxStatCtas = 256;
constexpr size_t kTrainScratchBytes = (size_t)kMaxStatCtas * 2 * kHidden * sizeof(double) + 1024;

int deqsci::denoiser_num_layers(const deqsci_denoiser* h) { return h ? (int)h->layers.size() : 0; }

extern "C" int deqsci_version(void) { return DEQSCI_VERSION; }
extern "C" const char* deqsci_last_error(void) { return g_err; }

extern "C" int deqsci_denoiser_destroy(deqsci_denoiser* h) {
  if (!h) return DEQSCI_OK;
  for (auto& L : h->layers) {
    if (L.w_cc) cudaFree(L.w_cc);
    if (L.w_tc) cudaFree(L.w_tc);
    if (L.w_tc2) cudaFree(L.w_tc2);
    if (L.scale) cudaFree(L.scale);
    if (L.bias) cudaFree(L.bias);
  }
  for (auto& kv : h->maps) cudaFree(kv.second);
  for (auto& kv : h->chain) if (kv.second.flags) cudaFree(kv.second.flags);
  delete h;
  return DEQSCI_OK;
}

static int upload(const void* host, size_t bytes, void** dev) {
  DEQSCI_CUDA(cudaMalloc(dev, bytes));
  DEQSCI_CUDA(cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice));
  return DEQSCI_OK;
}

extern "C" int deqsci_denoiser_create(int net_kind, int precision, int num_layers,
                                      const deqsci_conv_layer* layers_host, deqsci_denoiser** out) {
  DEQSCI_CHECK_ARG(out != nullptr && layers_host != nullptr, "denoiser_create: null pointer");
  *out = nullptr;
  DEQSCI_CHECK_ARG(net_kind == DEQSCI_NET_FFDNET || net_kind == DEQSCI_NET_DNCNN, "denoiser_create: net_kind=%d",
                   net_kind);
  DEQSCI_CHECK_ARG(precision >= DEQSCI_PREC_TC_SPLIT && precision <= DEQSCI_PREC_TC_SINGLE,
                   "denoiser_create: precision=%d", precision);
  DEQSCI_CHECK_ARG(num_layers >= 2 && num_layers <= 64, "denoiser_create: num_layers=%d (need 2..64)", num_layers);
  const int cin0 = net_kind == DEQSCI_NET_FFDNET ? 5 : 1, coutL = net_kind == DEQSCI_NET_FFDNET ? 4 : 1;
  for (int i = 0; i < num_layers; ++i) {
    const deqsci_conv_layer& L = layers_host[i];
    const int want_in = i == 0 ? cin0 : kHidden, want_out = i == num_layers - 1 ? coutL : kHidden;
    DEQSCI_CHECK_ARG(L.cin == want_in && L.cout == want_out, "denoiser_create: layer %d is %d->%d, expected %d->%d",
                     i, L.cin, L.cout, want_in, want_out);
    DEQSCI_CHECK_ARG(L.weight_host != nullptr, "denoiser_create: layer %d has no weights", i);
  }
  if (precision != DEQSCI_PREC_FP32) {
    int dev = 0, major = 0;
    DEQSCI_CUDA(cudaGetDevice(&dev));
    DEQSCI_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
      set_error("denoiser_create: tcgen05 precision modes need an sm_100 device (found sm_%d0)", major);
      return DEQSCI_ERR_ARCH;
    }
  }
  deqsci_denoiser* h = new deqsci_denoiser();
  h->kind = net_kind;
  h->precision = precision;
  h->layers.resize(num_layers);
  int rc = DEQSCI_OK;
  for (int i = 0; i < num_layers && rc == DEQSCI_OK; ++i) {
    const deqsci_conv_layer& S = layers_host[i];
    Layer& L = h->layers[i];
    L.cin = S.cin; L.cout = S.cout; L.relu = S.relu;
    // [cout][cin][3][3] -> [tap][cin][cout]
    std::vector<float> pk((size_t)9 * S.cin * S.cout);
    for (int o = 0; o < S.cout; ++o)
      for (int c = 0; c < S.cin; ++c)
        for (int t = 0; t < 9; ++t) pk[((size_t)t * S.cin + c) * S.cout + o] = S.weight_host[((size_t)o * S.cin + c) * 9 + t];
    rc = upload(pk.data(), pk.size() * sizeof(float), (void**)&L.w_cc);
    if (rc == DEQSCI_OK && S.scale_host) rc = upload(S.scale_host, S.cout * sizeof(float), (void**)&L.scale);
    if (rc == DEQSCI_OK && S.bias_host) rc = upload(S.bias_host, S.cout * sizeof(float), (void**)&L.bias);
    if (rc == DEQSCI_OK && i == 0 && precision != DEQSCI_PREC_FP32) {     // first layer: K padded to 16 per tap
      std::vector<uint8_t> img(tcf_weight_image_bytes());
      tcf_pack_weights(S.weight_host, S.cin, img.data());
      rc = upload(img.data(), img.size(), (void**)&L.w_tc);
    }
    // every layer with 64 input channels has a tensor-core image (hidden layers and the last layer)
    if (rc == DEQSCI_OK && i > 0 && precision != DEQSCI_PREC_FP32) {
      const bool split = precision == DEQSCI_PREC_TC_SPLIT;
      std::vector<uint8_t> img(tc_weight_image_bytes(split, S.cout));
      tc_pack_weights(S.weight_host, S.cout, split, img.data());
      rc = upload(img.data(), img.size(), (void**)&L.w_tc);
      if (rc == DEQSCI_OK && i == num_layers - 1) {
        std::vector<uint8_t> img2(tcl_weight_image_bytes());
        tcl_pack_weights(S.weight_host, S.cout, img2.data());
        rc = upload(img2.data(), img2.size(), (void**)&L.w_tc2);
      }
      if (rc == DEQSCI_OK && split && S.cout == kHidden && i < num_layers - 1) {
        std::vector<uint8_t> img2(tc2_weight_image_bytes());
        tc2_pack_weights(S.weight_host, img2.data());
        rc = upload(img2.data(), img2.size(), (void**)&L.w_tc2);
      }
    }
  }
  if (rc != DEQSCI_OK) {
    deqsci_denoiser_destroy(h);
    return rc;
  }
  *out = h;
  return DEQSCI_OK;
}

// ---- device-side weight refresh ------------------------------------------------------------------
namespace {
__global__ void repack_kernel(const float* __restrict__ w, const int32_t* __restrict__ map, void* __restrict__ dst,
                              int n, int as_f32) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t m = map[i];
  if (as_f32) {
    reinterpret_cast<float*>(dst)[i] = m < 0 ? 0.f : w[m >> 2];
    return;
  }
  const __half out = m >= 0 ? pack_half(w[m >> 2], m & 3) : __float2half_rn(0.f);
  reinterpret_cast<__half*>(dst)[i] = out;
}

template <class Fill>
int get_map(deqsci_denoiser* h, const std::string& key, int n, Fill fill, const int32_t** out) {
  auto it = h->maps.find(key);
  if (it == h->maps.end()) {
    std::vector<int32_t> host((size_t)n, -1);
    fill(host.data());
    int32_t* dev = nullptr;
    DEQSCI_CUDA(cudaMalloc(&dev, (size_t)n * sizeof(int32_t)));
    cudaError_t e = cudaMemcpy(dev, host.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(dev); set_error("update_weights: %s", cudaGetErrorString(e)); return DEQSCI_ERR_CUDA; }
    it = h->maps.emplace(key, dev).first;
  }
  *out = it->second;
  return DEQSCI_OK;
}

int build_layer_maps(deqsci_denoiser* h, int i) {
  Layer& L = h->layers[i];
  if (L.map_cc) return DEQSCI_OK;
  const int cin = L.cin, cout = L.cout, n_layers = (int)h->layers.size();
  const bool split = h->precision == DEQSCI_PREC_TC_SPLIT;
  char key[64];
  int rc;
  L.n_cc = 9 * cin * cout;
  snprintf(key, sizeof key, "cc:%d:%d", cin, cout);
  if ((rc = get_map(h, key, L.n_cc, [&](int32_t* m) {
        for (int o = 0; o < cout; ++o)
          for (int c = 0; c < cin; ++c)
            for (int t = 0; t < 9; ++t) m[(t * cin + c) * cout + o] = 4 * ((o * cin + c) * 9 + t);
      }, &L.map_cc))) return rc;
  if (L.w_tc && i == 0) {
    L.n_tc = (int)(tcf_weight_image_bytes() / 2);
    snprintf(key, sizeof key, "tcf:%d", cin);
    if ((rc = get_map(h, key, L.n_tc, [&](int32_t* m) { tcf_pack_map(cin, m); }, &L.map_tc))) return rc;
  } else if (L.w_tc) {
    L.n_tc = (int)(tc_weight_image_bytes(split, cout) / 2);
    snprintf(key, sizeof key, "tc:%d:%d", cout, (int)split);
    if ((rc = get_map(h, key, L.n_tc, [&](int32_t* m) { tc_pack_map(cout, split, m); }, &L.map_tc))) return rc;
  }
  if (L.w_tc2 && i == n_layers - 1) {
    L.n_tc2 = (int)(tcl_weight_image_bytes() / 2);
    snprintf(key, sizeof key, "tcl:%d", cout);
    if ((rc = get_map(h, key, L.n_tc2, [&](int32_t* m) { tcl_pack_map(cout, m); }, &L.map_tc2))) return rc;
  } else if (L.w_tc2) {
    L.n_tc2 = (int)(tc2_weight_image_bytes() / 2);
    if ((rc = get_map(h, "tc2", L.n_tc2, [&](int32_t* m) { tc2_pack_map(m); }, &L.map_tc2))) return rc;
  }
  return DEQSCI_OK;
}
}  // namespace

extern "C" int deqsci_denoiser_update_weights(deqsci_denoiser* h, int num_layers, const float* const* weight_dev,
                                              void* stream) {
  DEQSCI_CHECK_ARG(h && weight_dev, "update_weights: null pointer");
  DEQSCI_CHECK_ARG(num_layers == (int)h->layers.size(), "update_weights: %d layers given, the plan has %d", num_layers,
                   (int)h->layers.size());
  for (int i = 0; i < num_layers; ++i) DEQSCI_CHECK_ARG(weight_dev[i] != nullptr, "update_weights: layer %d is null", i);
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < num_layers; ++i) {
    int rc = build_layer_maps(h, i);
    if (rc != DEQSCI_OK) return rc;
    Layer& L = h->layers[i];
    repack_kernel<<<(L.n_cc + 255) / 256, 256, 0, st>>>(weight_dev[i], L.map_cc, L.w_cc, L.n_cc, 1);
    if (L.w_tc) repack_kernel<<<(L.n_tc + 255) / 256, 256, 0, st>>>(weight_dev[i], L.map_tc, L.w_tc, L.n_tc, 0);
    if (L.w_tc2) repack_kernel<<<(L.n_tc2 + 255) / 256, 256, 0, st>>>(weight_dev[i], L.map_tc2, L.w_tc2, L.n_tc2, 0);
    DEQSCI_LAUNCH_CHECK();
  }
  return DEQSCI_OK;
}

namespace {
struct Geometry {
  int SC, Hc, Wc, NF;
  long long plane_elems;
  size_t zprime_bytes, act_bytes;
};
int geometry(const deqsci_denoiser* h, int B, int H, int W, int T, Geometry* g) {
  DEQSCI_CHECK_ARG(h != nullptr, "null denoiser handle");
  DEQSCI_CHECK_ARG(B > 0 && H > 0 && W > 0 && T > 0, "non-positive dimension B=%d H=%d W=%d T=%d", B, H, W, T);
  DEQSCI_CHECK_ARG(B <= 65535, "B=%d exceeds 65535 per call", B);
  g->SC = h->kind == DEQSCI_NET_FFDNET ? 2 : 1;
  DEQSCI_CHECK_ARG(H % g->SC == 0 && W % g->SC == 0, "FFDNet needs even H and W (got %dx%d)", H, W);
  g->Hc = H / g->SC;
  g->Wc = W / g->SC;
  g->NF = B * T;
  g->plane_elems = (long long)g->NF * g->Hc * g->Wc * kHidden;
  g->zprime_bytes = align_up((size_t)B * H * W * T * sizeof(float), 1024);
  g->act_bytes = align_up((size_t)g->plane_elems * 2 * sizeof(__half), 1024);
  return DEQSCI_OK;
}
}  // namespace

extern "C" size_t deqsci_denoiser_workspace_bytes(const deqsci_denoiser* h, int B, int H, int W, int T) {
  Geometry g;
  if (geometry(h, B, H, W, T, &g) != DEQSCI_OK) return 0;
  return 1024 + g.zprime_bytes + 2 * g.act_bytes + kTrainScratchBytes;
}

// save  (optional, num_layers-1 entries): layer i writes its output planes to save[i] instead of the ping-pong
//       buffers (and layer i+1 reads them): the activations a backward pass needs.
// masks (optional, num_layers-1 entries): layer i's output is gated by the sign of masks[i] (a saved activation's hi
//       plane) instead of ReLU -- the adjoint stack of a conv / ReLU network (tensor-core kernels only).
static int run_stack(const deqsci_denoiser* h, bool fuse_gap, const float* z, const float* y, const float* phi,
                     const float* phi_sum, float sigma, float* out, void* workspace, size_t workspace_bytes, int B,
                     int H, int W, int T, void* stream, const deqsci_bn_params* bn = nullptr, float momentum = 0.f,
                     float eps = 0.f, const deqsci_saved_forward* sv = nullptr, const void* const* masks = nullptr) {
  void* const* save = sv ? sv->acts : nullptr;
  Geometry g;
  int rc = geometry(h, B, H, W, T, &g);
  if (rc) return rc;
  DEQSCI_CHECK_ARG(z != nullptr && out != nullptr && workspace != nullptr, "null pointer");
  if (fuse_gap) DEQSCI_CHECK_ARG(y && phi && phi_sum, "iterate: null y / phi / phi_sum");
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
  const size_t need = (size_t)(ws - reinterpret_cast<uint8_t*>(workspace)) + g.zprime_bytes + 2 * g.act_bytes + kTrainScratchBytes;
  if (workspace_bytes < need) {
    set_error("workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
    return DEQSCI_ERR_WORKSPACE;
  }
  float* zprime_ws = reinterpret_cast<float*>(ws);
  __half* act[2] = {reinterpret_cast<__half*>(ws + g.zprime_bytes),
                    reinterpret_cast<__half*>(ws + g.zprime_bytes + g.act_bytes)};
  cudaStream_t st = (cudaStream_t)stream;
  const int nl = (int)h->layers.size();
  if (save)
    for (int i = 0; i < nl - 1; ++i) DEQSCI_CHECK_ARG(save[i] != nullptr, "activation buffer of layer %d is null", i);
  if (masks) {
    for (int i = 0; i < nl - 1; ++i) DEQSCI_CHECK_ARG(masks[i] != nullptr, "mask plane of layer %d is null", i);
    DEQSCI_CHECK_ARG(!bn && h->precision == DEQSCI_PREC_TC_SPLIT && tcf_supported(g.Wc) && tc2_supported(g.Hc, g.Wc),
                     "masked (adjoint) stacks need precision tc_split and conv images wider than 64 pixels (got %dx%d)",
                     g.Hc, g.Wc);
  }
  // output buffer of conv layer i / input buffer of conv layer i (i >= 1)
  auto out_buf = [&](int i, int cur) { return save ? reinterpret_cast<__half*>(save[i]) : act[cur ^ 1]; };
  auto in_buf = [&](int i, int cur) { return save ? reinterpret_cast<__half*>(save[i - 1]) : act[cur]; };
  auto mask_of = [&](int i) { return masks ? reinterpret_cast<const __half*>(masks[i]) : nullptr; };
  auto relu_of = [&](int i, int relu) { return masks ? 2 : relu; };
  double* bn_stats = reinterpret_cast<double*>(ws + g.zprime_bytes + 2 * g.act_bytes);      // [kMaxStatCtas][128]
  float* bn_scale_shift = reinterpret_cast<float*>(bn_stats + (size_t)kMaxStatCtas * 2 * kHidden);   // [128]
  if (bn) {
    DEQSCI_CHECK_ARG(h->precision == DEQSCI_PREC_TC_SPLIT && tc2_supported(g.Hc, g.Wc) && tcf_supported(g.Wc),
                     "train-mode BatchNorm path needs precision tc_split and conv images wider than 64 pixels "
                     "(got %dx%d)", g.Hc, g.Wc);
    DEQSCI_CHECK_ARG(num_sms() <= kMaxStatCtas, "train-mode BatchNorm path: %d SMs (max %d)", num_sms(), kMaxStatCtas);
    // rows of CTAs a launch does not start stay zero (every hidden layer of this call uses the same grid)
    DEQSCI_CUDA(cudaMemsetAsync(bn_stats, 0, (size_t)kMaxStatCtas * 2 * kHidden * sizeof(double), st));
  }
  const Layer& L0 = h->layers[0];
  // z' goes from gap_prep straight to the tensor-core last layer: frame-planar [B,T,H,W] suits both (the last
  // layer handles one frame per tile); every other producer / consumer keeps the cube layout [B,H,W,T]
  const bool zprime_planar = fuse_gap && h->precision != DEQSCI_PREC_FP32 && tcf_supported(g.Wc) && tcl_supported(g.Wc);
  if (h->precision != DEQSCI_PREC_FP32 && tcf_supported(g.Wc)) {
    // tensor-core first layer: GAP + unshuffle + K-packing into one 16-channel plane (parked in the second
    // ping-pong buffer, which is free until the first hidden layer writes it), then the MMA kernel
    const long long in_plane = (long long)g.NF * g.Hc * g.Wc * kPrepChannels;      // ONE plane: [Ah | Ah | Al']
    rc = gap_prep_launch(h->kind, z, y, phi, phi_sum, zprime_ws, act[1], in_plane, sigma, B, H, W, T, fuse_gap,
                         zprime_planar, st);
    if (rc) return rc;
    if (sv && sv->zprime) {          // the first layer's wgrad input: the frames z' (frame-planar)
      DEQSCI_CHECK_ARG(zprime_planar, "saving z' needs the fused GAP step and the tensor-core first / last layers");
      DEQSCI_CUDA(cudaMemcpyAsync(sv->zprime, zprime_ws, (size_t)B * H * W * T * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    rc = conv_first_tc_launch(act[1], in_plane, save ? reinterpret_cast<__half*>(save[0]) : act[0], g.plane_elems,
                              L0.w_tc, L0.scale, L0.bias, relu_of(0, L0.relu), g.NF, g.Hc, g.Wc, st, mask_of(0));
  } else {
    rc = conv_first_launch(h->kind, fuse_gap, z, y, phi, phi_sum, zprime_ws, sigma, L0.w_cc, L0.scale, L0.bias,
                           L0.relu, save ? reinterpret_cast<__half*>(save[0]) : act[0], g.plane_elems, B, H, W, T, st);
  }
  if (rc) return rc;
  int cur = 0;
  // all hidden layers in ONE launch (CTA pairs stay resident, per-strip ready flags instead of kernel boundaries) when
  // nothing between them needs a kernel boundary: no train-mode statistics, no saved activations, no masks
  const bool chained = !bn && !save && !masks && h->precision == DEQSCI_PREC_TC_SPLIT && nl - 2 >= 2 &&
                       tc2_chain_supported(g.NF, g.Hc, g.Wc, nl - 2);
  if (chained) {
    const uint8_t* wimg[32];
    const float *sc[32], *bi[32];
    int relu[32];
    for (int i = 1; i < nl - 1; ++i) {
      const Layer& L = h->layers[i];
      wimg[i - 1] = L.w_tc2; sc[i - 1] = L.scale; bi[i - 1] = L.bias; relu[i - 1] = L.relu;
    }
    ChainState* cs;
    {
      std::lock_guard<std::mutex> lock(h->chain_mutex);
      cs = &h->chain[st];
      const size_t want = tc2_chain_flag_count(g.NF, g.Hc);
      if (cs->count < want) {
        if (cs->flags) { DEQSCI_CUDA(cudaStreamSynchronize(st)); DEQSCI_CUDA(cudaFree(cs->flags)); cs->flags = nullptr; cs->count = 0; }
        DEQSCI_CUDA(cudaMalloc((void**)&cs->flags, want * sizeof(uint32_t)));
        DEQSCI_CUDA(cudaMemsetAsync(cs->flags, 0, want * sizeof(uint32_t), st));
        cs->count = want;
        cs->epoch = 0;
      }
    }
    rc = conv_hidden_chain_launch(act[0], act[1], g.plane_elems, nl - 2, wimg, sc, bi, relu, g.NF, g.Hc, g.Wc, cs->flags,
                                  &cs->epoch, st);
    if (rc) return rc;
    cur = (nl - 2) & 1;
  }
  for (int i = 1; i < nl - 1 && !chained; ++i) {
    const Layer& L = h->layers[i];
    const __half* a_in = in_buf(i, cur);
    __half* a_out = out_buf(i, cur);
    if (h->precision == DEQSCI_PREC_FP32)
      rc = conv_mid_fp32_launch(a_in, a_out, g.plane_elems, L.w_cc, L.scale, L.bias, L.relu, g.NF, g.Hc,
                                g.Wc, st);
    else if (bn && bn[i].running_mean) {      // a BatchNorm follows this conv (gamma / beta are NULL for affine=False)
      // train mode: raw conv + per-channel statistics, then batch-statistics BatchNorm (+ ReLU) in place
      // backward pass wanted: the raw conv output is kept in pre[i], the normalised activation goes to acts[i]
      __half* c_out = (sv && sv->pre && sv->pre[i]) ? reinterpret_cast<__half*>(sv->pre[i]) : a_out;
      rc = conv_hidden_2cta_launch(a_in, c_out, g.plane_elems, L.w_tc2, nullptr, nullptr, 0, g.NF, g.Hc,
                                   g.Wc, st, bn_stats);
      if (rc == DEQSCI_OK)
        rc = bn_train_launch(a_out, g.plane_elems, bn_stats, num_sms(), bn_scale_shift, bn[i].gamma, bn[i].beta,
                             bn[i].running_mean, bn[i].running_var, momentum, eps, (long long)g.NF * g.Hc * g.Wc,
                             L.relu, st, c_out, (sv && sv->bn_record) ? sv->bn_record + (size_t)i * 4 * kHidden : nullptr);
    } else if (h->precision == DEQSCI_PREC_TC_SPLIT && tc2_supported(g.Hc, g.Wc))
      rc = conv_hidden_2cta_launch(a_in, a_out, g.plane_elems, L.w_tc2, L.scale, L.bias, relu_of(i, L.relu), g.NF,
                                   g.Hc, g.Wc, st, nullptr, mask_of(i));
    else
      rc = conv_mid_tc_launch(h->precision == DEQSCI_PREC_TC_SPLIT, a_in, a_out, g.plane_elems, L.w_tc,
                              L.scale, L.bias, L.relu, g.NF, g.Hc, g.Wc, st);
    if (rc) return rc;
    cur ^= 1;
  }
  const Layer& LL = h->layers[nl - 1];
  const __half* a_last = save ? reinterpret_cast<const __half*>(save[nl - 2]) : act[cur];
  if (h->precision != DEQSCI_PREC_FP32 && tcl_supported(g.Wc))
    return conv_last_tc_launch(LL.cout, a_last, g.plane_elems, LL.w_tc2, LL.scale, LL.bias, LL.relu, g.NF, g.Hc,
                               g.Wc, fuse_gap ? zprime_ws : z, zprime_planar, out, H, W, T, st);
  if (h->precision != DEQSCI_PREC_FP32)
    return conv_tc_launch(h->kind == DEQSCI_NET_FFDNET ? 1 : 2, h->precision == DEQSCI_PREC_TC_SPLIT, a_last, nullptr,
                          g.plane_elems, LL.w_tc, LL.scale, LL.bias, LL.relu, g.NF, g.Hc, g.Wc,
                          fuse_gap ? zprime_ws : z, out, H, W, T, st);
  return conv_last_launch(h->kind, a_last, g.plane_elems, LL.w_cc, LL.scale, LL.bias, LL.relu,
                          fuse_gap ? zprime_ws : z, out, B, H, W, T, st);
}

extern "C" int deqsci_denoise_residual(const deqsci_denoiser* h, const float* zin, float sigma, float* out,
                                       void* workspace, size_t workspace_bytes, int B, int H, int W, int T,
                                       void* stream) {
  return run_stack(h, false, zin, nullptr, nullptr, nullptr, sigma, out, workspace, workspace_bytes, B, H, W, T,
                   stream);
}

extern "C" int deqsci_iterate(const deqsci_denoiser* h, const float* z, const float* y, const float* phi,
                              const float* phi_sum, float sigma, float* out, void* workspace, size_t workspace_bytes,
                              int B, int H, int W, int T, void* stream) {
  return run_stack(h, true, z, y, phi, phi_sum, sigma, out, workspace, workspace_bytes, B, H, W, T, stream);
}

extern "C" size_t deqsci_denoiser_activation_bytes(const deqsci_denoiser* h, int B, int H, int W, int T) {
  Geometry g;
  if (geometry(h, B, H, W, T, &g) != DEQSCI_OK) return 0;
  return g.act_bytes;
}

extern "C" int deqsci_iterate_save(const deqsci_denoiser* h, const float* z, const float* y, const float* phi,
                                   const float* phi_sum, float sigma, float* out, void* workspace,
                                   size_t workspace_bytes, const deqsci_saved_forward* save, int B, int H, int W, int T,
                                   void* stream) {
  DEQSCI_CHECK_ARG(save != nullptr && save->acts != nullptr, "iterate_save: null activation table");
  return run_stack(h, true, z, y, phi, phi_sum, sigma, out, workspace, workspace_bytes, B, H, W, T, stream, nullptr,
                   0.f, 0.f, save, nullptr);
}

extern "C" int deqsci_iterate_train_save(const deqsci_denoiser* h, const float* z, const float* y, const float* phi,
                                         const float* phi_sum, float sigma, float* out, void* workspace,
                                         size_t workspace_bytes, const deqsci_bn_params* bn_host, float momentum,
                                         float eps, const deqsci_saved_forward* save, int B, int H, int W, int T,
                                         void* stream) {
  DEQSCI_CHECK_ARG(bn_host != nullptr, "iterate_train_save: null BatchNorm table");
  DEQSCI_CHECK_ARG(save != nullptr && save->acts != nullptr, "iterate_train_save: null activation table");
  return run_stack(h, true, z, y, phi, phi_sum, sigma, out, workspace, workspace_bytes, B, H, W, T, stream, bn_host,
                   momentum, eps, save, nullptr);
}

// ---- weight gradients of one iterate-map call (kernels in backward.cu) --------------------------------------
extern "C" size_t deqsci_backward_workspace_bytes(const deqsci_denoiser* h, int B, int H, int W, int T) {
  Geometry g;
  if (geometry(h, B, H, W, T, &g) != DEQSCI_OK) return 0;
  // the adjoint plan's own stack workspace (K-packed plane of the last layer's dgrad), three gradient plane pairs,
  // the scaled upstream cube and the reduction scratch
  return 2048 + deqsci_denoiser_workspace_bytes(h, B, H, W, T) + 3 * g.act_bytes + g.zprime_bytes +
         align_up(backward_scratch_floats() * sizeof(float), 1024);
}

extern "C" int deqsci_backward_weights(const deqsci_denoiser* h, const deqsci_denoiser* adj,
                                       const deqsci_saved_forward* sv, const float* const* gamma, const float* grad,
                                       float grad_scale, float sigma, float* const* d_weight, float* const* d_gamma,
                                       float* const* d_beta, void* workspace, size_t workspace_bytes, int B, int H, int W,
                                       int T, void* stream) {
  DEQSCI_CHECK_ARG(h && adj && sv && sv->acts && sv->zprime && grad && d_weight && workspace, "backward_weights: null pointer");
  Geometry g;
  int rc = geometry(h, B, H, W, T, &g);
  if (rc) return rc;
  const int nl = (int)h->layers.size();
  DEQSCI_CHECK_ARG((int)adj->layers.size() == nl && adj->kind == h->kind, "backward_weights: adjoint plan does not match");
  DEQSCI_CHECK_ARG(h->precision == DEQSCI_PREC_TC_SPLIT && adj->precision == DEQSCI_PREC_TC_SPLIT && tcf_supported(g.Wc) &&
                       tc2_supported(g.Hc, g.Wc),
                   "backward_weights needs precision tc_split and conv images wider than 64 pixels (got %dx%d)", g.Hc, g.Wc);
  DEQSCI_CHECK_ARG(grad_scale > 0.f && grad_scale < 3.0e38f, "backward_weights: grad_scale must be positive and finite");
  const size_t need = deqsci_backward_workspace_bytes(h, B, H, W, T);
  if (workspace_bytes < need) {
    set_error("backward_weights: workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
    return DEQSCI_ERR_WORKSPACE;
  }
  for (int i = 0; i < nl; ++i) DEQSCI_CHECK_ARG(d_weight[i] != nullptr, "backward_weights: d_weight[%d] is null", i);
  for (int i = 0; i < nl - 1; ++i) DEQSCI_CHECK_ARG(sv->acts[i] != nullptr, "backward_weights: acts[%d] is null", i);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
  const size_t den_bytes = deqsci_denoiser_workspace_bytes(h, B, H, W, T);
  uint8_t* den_ws = ws;
  __half* G[2] = {reinterpret_cast<__half*>(ws + align_up(den_bytes, 1024)),
                  reinterpret_cast<__half*>(ws + align_up(den_bytes, 1024) + g.act_bytes)};
  __half* D = reinterpret_cast<__half*>(ws + align_up(den_bytes, 1024) + 2 * g.act_bytes);
  float* gsc = reinterpret_cast<float*>(ws + align_up(den_bytes, 1024) + 3 * g.act_bytes);
  float* scratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(gsc) + g.zprime_bytes);
  const long long count = (long long)g.NF * g.Hc * g.Wc;
  const long long ncube = (long long)B * H * W * T;
  auto act_of = [&](int i) { return reinterpret_cast<const __half*>(sv->acts[i]); };
  auto pre_of = [&](int i) { return (sv->pre && sv->pre[i]) ? reinterpret_cast<const __half*>(sv->pre[i]) : nullptr; };

  // upstream: the network's output enters f as  z' - noise  ->  gn = -grad (scaled)
  if ((rc = backward_begin(scratch, grad, gsc, grad_scale, ncube, st))) return rc;
  const Layer& LL = h->layers[nl - 1];
  if ((rc = wgrad_last_launch(LL.cout, act_of(nl - 2), g.plane_elems, gsc, B, H, W, T, scratch, d_weight[nl - 1], st))) return rc;
  // dgrad of the last layer: the adjoint plan's FIRST layer (gap_prep builds its K-packed input from the cube; for
  // FFDNet the sigma channel's adjoint weights are zero)
  {
    uint8_t* dws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(den_ws), 1024));
    float* zp_dummy = reinterpret_cast<float*>(dws);
    __half* kplane = reinterpret_cast<__half*>(dws + g.zprime_bytes + g.act_bytes);      // the stack's act[1]
    const long long in_plane = (long long)g.NF * g.Hc * g.Wc * kPrepChannels;
    if ((rc = gap_prep_launch(h->kind, gsc, nullptr, nullptr, nullptr, zp_dummy, kplane, in_plane, 0.f, B, H, W, T, false,
                              false, st))) return rc;
    const Layer& A0 = adj->layers[0];
    if ((rc = conv_first_tc_launch(kplane, in_plane, G[0], g.plane_elems, A0.w_tc, nullptr, nullptr, 0, g.NF, g.Hc, g.Wc,
                                   st))) return rc;
  }
  int cur = 0;
  for (int i = nl - 2; i >= 0; --i) {
    const __half* pre = pre_of(i);
    const float* rec = pre ? sv->bn_record + (size_t)i * 4 * kHidden : nullptr;
    if (pre) DEQSCI_CHECK_ARG(sv->bn_record != nullptr, "backward_weights: BatchNorm layer %d without a record", i);
    if ((rc = act_bwd_launch(G[cur], act_of(i), pre, rec, (pre && gamma) ? gamma[i] : nullptr, D, g.plane_elems, count,
                             scratch, (pre && d_gamma) ? d_gamma[i] : nullptr, (pre && d_beta) ? d_beta[i] : nullptr, st)))
      return rc;
    if (i == 0) {
      const Layer& L0 = h->layers[0];
      return wgrad_first_launch(L0.cin, D, g.plane_elems, sv->zprime, sigma, g.NF, H, W, scratch, d_weight[0], st);
    }
    if ((rc = wgrad_hidden_launch(act_of(i - 1), D, g.plane_elems, g.NF, g.Hc, g.Wc, scratch, d_weight[i], st))) return rc;
    // da_{i-1} = conv(dc_i, W_i^T flipped): adjoint layer nl-1-i
    const Layer& A = adj->layers[nl - 1 - i];
    if ((rc = conv_hidden_2cta_launch(D, G[cur ^ 1], g.plane_elems, A.w_tc2, nullptr, nullptr, 0, g.NF, g.Hc, g.Wc, st)))
      return rc;
    cur ^= 1;
  }
  return DEQSCI_OK;
}

extern "C" int deqsci_denoise_residual_masked(const deqsci_denoiser* h_adjoint, const float* vin, float* out,
                                              void* workspace, size_t workspace_bytes, const void* const* masks_host,
                                              int B, int H, int W, int T, void* stream) {
  DEQSCI_CHECK_ARG(masks_host != nullptr, "denoise_residual_masked: null mask table");
  return run_stack(h_adjoint, false, vin, nullptr, nullptr, nullptr, 0.f, out, workspace, workspace_bytes, B, H, W, T,
                   stream, nullptr, 0.f, 0.f, nullptr, masks_host);
}

extern "C" int deqsci_iterate_train(const deqsci_denoiser* h, const float* z, const float* y, const float* phi,
                                    const float* phi_sum, float sigma, float* out, void* workspace,
                                    size_t workspace_bytes, const deqsci_bn_params* bn_host, float momentum, float eps,
                                    int B, int H, int W, int T, void* stream) {
  DEQSCI_CHECK_ARG(bn_host != nullptr, "iterate_train: null BatchNorm table");
  return run_stack(h, true, z, y, phi, phi_sum, sigma, out, workspace, workspace_bytes, B, H, W, T, stream, bn_host,
                   momentum, eps);
}

// Testing hook (declared in deqsci.h): the pair kernel's strip-height choice, host arithmetic only.
extern "C" int deqsci_debug_pair_strip_rows(int NF, int Hc, int Wc, int n_sms) {
  if (NF <= 0 || Hc <= 0 || Wc <= 0 || n_sms < 2) return 0;
  return pick_strip_rows_balanced(NF, (Wc + 127) / 128, Hc, false, n_sms / 2, 2, 1, 1);
}

extern "C" long long deqsci_debug_pair_weight_map(int* map, long long count) {
  const long long n = (long long)(tc2_weight_image_bytes() / 2);
  if (map && count >= n) tc2_pack_map(map);
  return n;
}

// Testing hook (declared in deqsci.h): one hidden 64->64 layer on caller-provided planes.
extern "C" int deqsci_debug_hidden_layer(const deqsci_denoiser* h, int layer, const void* act_in, void* act_out,
                                         int NF, int Hc, int Wc, void* stream) {
  DEQSCI_CHECK_ARG(h && act_in && act_out, "debug_hidden_layer: null pointer");
  DEQSCI_CHECK_ARG(layer > 0 && layer < (int)h->layers.size() - 1, "debug_hidden_layer: layer %d is not hidden", layer);
  DEQSCI_CHECK_ARG(NF > 0 && Hc > 0 && Wc > 0, "debug_hidden_layer: bad shape");
  const Layer& L = h->layers[layer];
  const long long plane = (long long)NF * Hc * Wc * kHidden;
  cudaStream_t st = (cudaStream_t)stream;
  if (h->precision == DEQSCI_PREC_FP32)
    return conv_mid_fp32_launch((const __half*)act_in, (__half*)act_out, plane, L.w_cc, L.scale, L.bias, L.relu, NF,
                                Hc, Wc, st);
  if (h->precision == DEQSCI_PREC_TC_SPLIT && tc2_supported(Hc, Wc))
    return conv_hidden_2cta_launch((const __half*)act_in, (__half*)act_out, plane, L.w_tc2, L.scale, L.bias, L.relu,
                                   NF, Hc, Wc, st);
  return conv_mid_tc_launch(h->precision == DEQSCI_PREC_TC_SPLIT, (const __half*)act_in, (__half*)act_out, plane,
                            L.w_tc, L.scale, L.bias, L.relu, NF, Hc, Wc, st);
}
