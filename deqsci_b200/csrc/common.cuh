// Shared host/device helpers for libdeqsci (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <utility>

#include "../../include/deqsci.h"

struct CUtensorMap_st;   // <cuda.h>'s CUtensorMap (driver API type; only tma_host.cu and the TMA kernels include it)

namespace deqsci {

// thread-local error text behind deqsci_last_error()
void set_error(const char* fmt, ...);

#define DEQSCI_CHECK_ARG(cond, ...)                  \
  do {                                               \
    if (!(cond)) {                                   \
      ::deqsci::set_error(__VA_ARGS__);              \
      return DEQSCI_ERR_INVALID;                     \
    }                                                \
  } while (0)

#define DEQSCI_CUDA(call)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      ::deqsci::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),     \
                          __FILE__, __LINE__);                                         \
      return DEQSCI_ERR_CUDA;                                                          \
    }                                                                                  \
  } while (0)

#define DEQSCI_LAUNCH_CHECK() DEQSCI_CUDA(cudaGetLastError())

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// kernel classes for launch accounting / sampled event timing (profile.cu); order = deqsci.h
enum ProfKind { PK_GAP = 0, PK_CONV_FIRST, PK_CONV_HIDDEN, PK_CONV_LAST, PK_AND_GRAM, PK_AND_SOLVE, PK_AND_MIX, PK_COUNT };
struct ProfScope {
  ProfScope(int kind, cudaStream_t st);
  ~ProfScope();
  int kind_;
  cudaStream_t st_;
  long long idx_;
};

int env_int(const char* name, int dflt);

// Launch with programmatic stream serialization (PDL): the kernel may start while its stream predecessor is
// still draining.  Only for kernels that execute `griddepcontrol.wait` before touching anything the
// predecessor wrote.  DEQSCI_TC_PDL=0 falls back to plain stream order.
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  static const int pdl = env_int("DEQSCI_TC_PDL", 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait_predecessor() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// tma_host.cu
int make_plane_map(::CUtensorMap_st* map, const __half* plane, int channels, int NF, int Hc, int Wc, int box_c,
                   int box_w, int box_h, int swizzle_bytes);
int pick_strip_rows(int NF, int tiles_x, int Hc, bool must_divide, long long min_items, int floor_rows);
int pick_strip_rows_balanced(int NF, int tiles_x, int Hc, bool must_divide, int workers, int strips_per_item,
                             int overhead_half_rows, int min_rows);
int env_int(const char* name, int dflt);

// Activation storage between conv layers: channels-last [frames, H, W, 64], each value stored as
// an fp16 pair  v ~= hi + lo * 2^-11  in two planes (hi plane then lo plane).  Keeps ~22 mantissa
// bits (this split reproduces the fp32 trajectory at its noise floor) in exactly
// the operand format the tcgen05 kind::f16 MMA consumes.
constexpr int kHidden = 64;
constexpr int kPrepChannels = 16;             // the first layer's input plane: one K = 16 row per pixel, [Ah | Ah | Al' | 0]
constexpr int kMaxBnLayers = 32;              // conv layers a train-mode driver call can snapshot

// bn_train.cu / api.cu helpers used by the driver
int bn_running_snapshot(const deqsci_bn_params* bn, int n_layers, float* backup, int restore, cudaStream_t st);
int denoiser_num_layers(const deqsci_denoiser* h);
constexpr float kLoScale = 2048.0f;           // 2^11
constexpr float kLoInvScale = 1.0f / 2048.0f;

__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn((v - __half2float(hi)) * kLoScale);
}
// Two values at once with the packed conversion (cvt.rn.f16x2.f32 = F2FP, a full-rate ALU instruction; the scalar
// cvt.rn.f16.f32 is a quarter-rate F2F and was 64 of ~480 instructions per epilogue tile).  Same roundings as
// split_f16; packs v0 into the low half, as the activation planes want consecutive channels.
__device__ __forceinline__ void split_f16x2(float v0, float v1, uint32_t& hi_pk, uint32_t& lo_pk) {
  const __half2 h = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((v0 - hf.x) * kLoScale, (v1 - hf.y) * kLoScale);
  hi_pk = *reinterpret_cast<const uint32_t*>(&h);
  lo_pk = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float join_f16(__half hi, __half lo) {
  return fmaf(__half2float(lo), kLoInvScale, __half2float(hi));
}

// The weight images (shared-memory operand layouts of the tensor-core kernels) are written through an
// emitter -- emit(byte offset in the image, index into the [cout][cin][3][3] weights, kind of half) -- so one
// layout routine serves the host packer (PackWrite) and the gather map of the device-side repack
// (PackMap: element -> 4*index + kind, -1 = zero padding; deqsci_denoiser_update_weights).
// Kinds: 0 = hi(w), 1 = lo'(w) = (w - hi) * 2^11, 2 = the low part at its true scale (w - hi), 3 = hi * 2^-11
// (kinds 2 and 3 live in fp16's subnormal range: the K-packed first layer, conv_tc_first.cu).
enum { kPackHi = 0, kPackLo = 1, kPackLoTrue = 2, kPackHiSmall = 3 };
__host__ __device__ inline __half pack_half(float v, int kind) {
  const __half hi = __float2half_rn(v);
  const float hf = __half2float(hi);
  if (kind == kPackLo) return __float2half_rn((v - hf) * kLoScale);
  if (kind == kPackLoTrue) return __float2half_rn(v - hf);
  if (kind == kPackHiSmall) return __float2half_rn(hf * kLoInvScale);
  return hi;
}
struct PackWrite {
  const float* w;
  uint8_t* img;
  void operator()(size_t byte, int src, int kind) const {
    *reinterpret_cast<__half*>(img + byte) = pack_half(w[src], kind);
  }
};
struct PackMap {
  int32_t* map;
  void operator()(size_t byte, int src, int kind) const { map[byte >> 1] = src * 4 + kind; }
};

}  // namespace deqsci
