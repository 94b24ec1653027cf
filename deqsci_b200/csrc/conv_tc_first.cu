// Kernel group 2d: the FIRST conv layer of the denoiser on tensor cores (split-fp16 precision, K-packed).
//
// The layer has C = 5 (FFDNet: sigma + 4 unshuffled sub-pixels) or 1 (DnCNN) input channels, so one UMMA_K = 16
// row per pixel has room for all three products of the precision split at once:
//     A row (gap_prep, gap.cu)  = [ Ah x C | Ah x C | Al' x C | 0 ]           Al' = (a - Ah) * 2^11
//     B row (tcf_layout below)  = [ Wh x C | Wl x C | Wh * 2^-11 x C | 0 ]    Wl = w - Wh at its TRUE scale
//     A . B = Ah.Wh + Ah.Wl + Al.Wh                                           (fp32 accumulation in TMEM)
// -- ONE MMA per tap into 64 accumulator columns (two MMAs into 128 columns with separate hi / lo planes before).
// Wl and Wh * 2^-11 sit in fp16's subnormal range, which still carries the 6-10 bits the correction terms need
// (CPU emulation: tests/tools/emulate_precision.py first_kpack; the per-iterate error is unchanged).
// A tap is a UMMA descriptor 32 bytes (one pixel) further into a TMA-loaded input row (32-byte swizzle).
//
// Structure = the rolling-row pipeline of conv_tc.cu (LD_ROLL) with the epilogue of conv_tc2.cu; two CTAs per SM
// (84 KB, 96 registers, 256 TMEM columns = 4 accumulator buffers each) hide each other's epilogue latency;
// warp 0 TMA producer (6-slot ring of input rows with 1-pixel halo), warp 1 TMEM allocator + MMA issuer,
// warps 2-9 epilogue (tcgen05.ld -> affine + ReLU -> hi/lo split -> swizzled smem -> TMA store of whole pixels).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace deqsci {
namespace tcf {
using namespace ptx;

constexpr int kTileM = 128;
constexpr int kThreads = 320;
constexpr int kSlots = 6;
constexpr int kRowBytes = 32;                                  // one K = 16 row, fp16
constexpr int kSlotBytes = 5 * 1024;                           // 130 pixels x 32 B = 4160 B, padded
constexpr int kTxBytes = (kTileM + 2) * kRowBytes;
constexpr int kTapBytesB = 64 * kRowBytes;                     // 64 cout rows x 16 k
constexpr int kWBytes = 9 * kTapBytesB;                        // 18 KB
constexpr int kStageBytes = 4096;                              // one staging tile: 32 px x 64 ch fp16 (per warp pair: hi tile, lo tile)
constexpr int kAccCols = 64;
constexpr int kBufs = 4;                                       // accumulator buffers
constexpr int kTmemCols = kBufs * kAccCols;                    // 256
constexpr int kSmemBytes = 1024 + kWBytes + kSlots * kSlotBytes + 8 * kStageBytes + 1024;

struct Params {
  const uint8_t* wimg;
  const float* scale;
  const float* bias;
  int relu;
  int NF, Hc, Wc;
  int tiles_x, strips_y, strip_rows;
  long long n_items;
  const __half* mask;      // relu == 2: keep (pixel, channel) where this hi plane [NF,Hc,Wc,64] is > 0 (see conv_tc2.cu)
};
struct Item { int nf, h0, w0, ntiles; };

__device__ __forceinline__ Item decode(const Params& p, long long item) {
  const int per_frame = p.tiles_x * p.strips_y;
  Item it;
  it.nf = (int)(item / per_frame);
  const int rem = (int)(item - (long long)it.nf * per_frame);
  const int sy = rem / p.tiles_x;
  it.w0 = (rem - sy * p.tiles_x) * kTileM;
  it.h0 = sy * p.strip_rows;
  it.ntiles = min(p.strip_rows, p.Hc - it.h0);
  return it;
}

// AFFINE = false: the layer has no folded scale / bias (FFDNet's and DnCNN's first layers: conv + ReLU), so the
// epilogue skips the per-channel multiply-add and its shared-memory loads
template <bool AFFINE>
__global__ void __launch_bounds__(kThreads, 2)
conv_first_tc_kernel(const __grid_constant__ CUtensorMap in_map,
                     const __grid_constant__ CUtensorMap out_hi, const __grid_constant__ CUtensorMap out_lo,
                     const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* w_s = smem;
  uint8_t* a_s = w_s + kWBytes;
  uint8_t* st_s = a_s + kSlots * kSlotBytes;
  uint8_t* tail = st_s + 8 * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);          // [0] w, full[S], empty[S], tfull[kBufs], tempty[kBufs]
  float* aff_s = reinterpret_cast<float*>(tail + 256);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 256 + 512);
  pdl_launch_dependents();       // the next kernel's prologue may overlap this grid's tail (see conv_tc2.cu)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_w = smem_u32(&bars[0]);
  auto bar_full = [&](int s) { return smem_u32(&bars[1 + s]); };
  auto bar_empty = [&](int s) { return smem_u32(&bars[1 + kSlots + s]); };
  auto bar_tfull = [&](int b) { return smem_u32(&bars[1 + 2 * kSlots + b]); };
  auto bar_tempty = [&](int b) { return smem_u32(&bars[1 + 2 * kSlots + kBufs + b]); };

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < kSlots; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int b = 0; b < kBufs; ++b) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), 8); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) {
    const int c = threadIdx.x - 64;
    aff_s[2 * c] = p.scale ? p.scale[c] : 1.f;          // {scale, bias} pairs
    aff_s[2 * c + 1] = p.bias ? p.bias[c] : 0.f;
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(bar_w, kWBytes);
      bulk_load_1d(smem_u32(w_s), p.wimg, kWBytes, bar_w);
      pdl_wait_predecessor();      // the input planes are the previous kernel's output
      int slot = 0;
      uint32_t phase = 0;
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const Item it = decode(p, item);
        for (int q = 0; q < it.ntiles + 2; ++q) {
          mbar_wait(bar_empty(slot), phase ^ 1);
          mbar_arrive_expect_tx(bar_full(slot), kTxBytes);
          const uint32_t dst = smem_u32(a_s + slot * kSlotBytes);
          tma_load_4d(dst, &in_map, bar_full(slot), 0, it.w0 - 1, it.h0 - 1 + q, it.nf);
          if (++slot == kSlots) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc(kTileM, 64);
      mbar_wait(bar_w, 0);
      const uint32_t a_base = smem_u32(a_s), w_base = smem_u32(w_s);
      int first = 0;
      uint32_t first_phase = 0;
      int buf = 0;
      uint32_t tphase = 0;
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const Item it = decode(p, item);
        int wait_slot = first;
        uint32_t wait_phase = first_phase;
        int rows_ready = 0;
        for (int j = 0; j < it.ntiles; ++j) {
          mbar_wait(bar_tempty(buf), tphase ^ 1);
          while (rows_ready < j + 3) {
            mbar_wait(bar_full(wait_slot), wait_phase);
            if (++wait_slot == kSlots) { wait_slot = 0; wait_phase ^= 1; }
            ++rows_ready;
          }
          tc_fence_after();
          const uint32_t d_main = tmem_base + buf * kAccCols;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int slot = (first + j + ky) % kSlots;
            const uint32_t a_row = a_base + slot * kSlotBytes;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int tap = ky * 3 + kx;
              umma_f16(d_main, sdesc_sw32(a_row + kx * kRowBytes), sdesc_sw32(w_base + tap * kTapBytesB), idesc,
                       tap != 0);
            }
          }
          const int dead = (first + j) % kSlots;
          umma_commit(bar_empty(dead));
          if (j == it.ntiles - 1) {
            umma_commit(bar_empty((dead + 1) % kSlots));
            umma_commit(bar_empty((dead + 2) % kSlots));
          }
          umma_commit(bar_tfull(buf));
          if (++buf == kBufs) { buf = 0; tphase ^= 1; }
        }
        first = wait_slot;
        first_phase = wait_phase;
      }
    }
  } else {
    const int e = warp - 2;
    const int quarter = warp & 3;
    const int half = e >> 2;
    // the two warps that own the same 32 pixels (channel halves 0 / 1) share one staging tile per plane --
    // [32 px][128 B], 128-byte swizzle -- so a TMA store moves whole 128-byte pixels: with 64-byte rows the
    // store engine's request rate (one row per request) capped this kernel at ~12 B/clk/SM
    const int pair = e & 3;
    const uint32_t stage = smem_u32(st_s + pair * 2 * kStageBytes);
    int buf = 0;
    uint32_t tphase = 0;
    for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const Item it = decode(p, item);
      for (int j = 0; j < it.ntiles; ++j) {
        const int h = it.h0 + j;
        uint32_t mkw[16];
        if (p.relu == 2) {
#pragma unroll
          for (int q = 0; q < 16; ++q) mkw[q] = 0u;
          const int wpx = it.w0 + quarter * 32 + lane;
          if (wpx < p.Wc && h < p.Hc) {
            const uint4* mp = reinterpret_cast<const uint4*>(
                p.mask + (((long long)it.nf * p.Hc + h) * p.Wc + wpx) * 64 + half * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 t4 = __ldg(mp + q);
              mkw[4 * q] = t4.x; mkw[4 * q + 1] = t4.y; mkw[4 * q + 2] = t4.z; mkw[4 * q + 3] = t4.w;
            }
          }
        }
        mbar_wait(bar_tfull(buf), tphase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * kAccCols;
        uint32_t hi_pk[16], lo_pk[16];
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          uint32_t acc[16];
          tmem_ld16(t_row + half * 32 + part * 16, acc);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float v[2];
            float sb[4] = {1.f, 0.f, 1.f, 0.f};
            if (AFFINE) {
              const float4 sb4 = *reinterpret_cast<const float4*>(aff_s + 2 * (half * 32 + part * 16 + i));
              sb[0] = sb4.x; sb[1] = sb4.y; sb[2] = sb4.z; sb[3] = sb4.w;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              float a = __uint_as_float(acc[i + u]);
              if (AFFINE) a = fmaf(a, sb[2 * u], sb[2 * u + 1]);
              if (p.relu == 2) {
                const uint32_t mbits = (mkw[part * 8 + (i >> 1)] >> (16 * u)) & 0x7fffu;
                v[u] = mbits ? a : 0.f;
              } else {
                v[u] = p.relu ? fmaxf(a, 0.f) : a;
              }
            }
            split_f16x2(v[0], v[1], hi_pk[part * 8 + (i >> 1)], lo_pk[part * 8 + (i >> 1)]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (elect_one_sync()) mbar_arrive(bar_tempty(buf));
        const uint32_t row_addr = stage + lane * 128;
        const int sw = lane & 7;
        // the previous tile's stores (issued by the half-0 warp) have read the staging tiles; every lane executes
        // the wait (only the issuing lane has a bulk group pending): no lane-0 guard, which ptxas would wrap in a
        // per-lane uniformisation loop (BRA.U.ANY)
        if (half == 0) bulk_wait_read0();
        named_bar_sync(1 + pair, 64);
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
          const uint32_t* pk = plane == 0 ? hi_pk : lo_pk;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + plane * kStageBytes +
                                                                          (((half * 4 + q) ^ sw) << 4)),
                         "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                         : "memory");
          }
        }
        fence_proxy_async();
        named_bar_sync(1 + pair, 64);
        if (half == 0) {
          if (elect_one_sync()) {
            tma_store_4d(&out_hi, stage, 0, it.w0 + quarter * 32, h, it.nf);
            tma_store_4d(&out_lo, stage + kStageBytes, 0, it.w0 + quarter * 32, h, it.nf);
            bulk_commit();
          }
        }
        if (++buf == kBufs) { buf = 0; tphase ^= 1; }
      }
    }
    bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace tcf

size_t tcf_weight_image_bytes() { return tcf::kWBytes; }

// w [64 cout][cin][3][3] fp32 (cin = 5 or 1, 3 * cin <= 16) -> [tap][64 cout rows][16 k] fp16, 32-byte swizzle
// (16-byte chunk index XOR bit 2 of the row index); unused k slots are zero.
template <class Emit>
static void tcf_layout(int cin, Emit emit) {
  for (int tap = 0; tap < 9; ++tap) {
    const int ky = tap / 3, kx = tap % 3;
    for (int n = 0; n < 64; ++n)
      for (int kk = 0; kk < 3 * cin; ++kk) {        // K slots: [Wh x cin | Wl (true scale) x cin | Wh * 2^-11 x cin]
        const int k = kk % cin;
        const int kind = kk < cin ? kPackHi : (kk < 2 * cin ? kPackLoTrue : kPackHiSmall);
        const size_t byte = (size_t)tap * tcf::kTapBytesB + (size_t)n * 32 + (size_t)(((kk >> 3) ^ ((n >> 2) & 1)) << 4) +
                            (size_t)(kk & 7) * 2;
        emit(byte, ((n * cin + k) * 3 + ky) * 3 + kx, kind);
      }
  }
}
void tcf_pack_weights(const float* w, int cin, uint8_t* img) {
  memset(img, 0, tcf::kWBytes);
  tcf_layout(cin, PackWrite{w, img});
}
void tcf_pack_map(int cin, int32_t* map) { tcf_layout(cin, PackMap{map}); }

bool tcf_supported(int Wc) {
  static const int enabled = env_int("DEQSCI_TC_FIRST", 1);
  return enabled && Wc > 64;
}

// planes_in: [NF,Hc,Wc,16] fp16, K-packed rows (gap_prep_kernel); act_out: [2][NF,Hc,Wc,64] fp16
int conv_first_tc_launch(const __half* planes_in, long long in_plane_elems, __half* act_out, long long plane_elems,
                         const uint8_t* wimg, const float* scale, const float* bias, int relu, int NF, int Hc,
                         int Wc, cudaStream_t st, const __half* mask) {
  tcf::Params p;
  p.wimg = wimg; p.scale = scale; p.bias = bias; p.relu = relu;
  p.mask = mask;
  if (relu == 2 && !mask) { set_error("conv_first_tc_launch: masked layer without a mask plane"); return DEQSCI_ERR_INVALID; }
  p.NF = NF; p.Hc = Hc; p.Wc = Wc;
  p.tiles_x = (Wc + tcf::kTileM - 1) / tcf::kTileM;
  const int R = pick_strip_rows_balanced(NF, p.tiles_x, Hc, false, 2 * num_sms(), 1, 1, 2);
  p.strip_rows = R;
  p.strips_y = (Hc + R - 1) / R;
  p.n_items = (long long)NF * p.tiles_x * p.strips_y;
  CUtensorMap in_map, out_hi, out_lo;
  int rc;
  (void)in_plane_elems;
  if ((rc = make_plane_map(&in_map, planes_in, kPrepChannels, NF, Hc, Wc, 16, tcf::kTileM + 2, 1, 32))) return rc;
  if ((rc = make_plane_map(&out_hi, act_out, 64, NF, Hc, Wc, 64, 32, 1, 128))) return rc;
  if ((rc = make_plane_map(&out_lo, act_out + plane_elems, 64, NF, Hc, Wc, 64, 32, 1, 128))) return rc;
  const int grid = (int)(p.n_items < 2 * num_sms() ? p.n_items : 2 * num_sms());      // two CTAs per SM
  ProfScope prof(PK_CONV_FIRST, st);
  if (scale || bias) {
    DEQSCI_CUDA(cudaFuncSetAttribute(tcf::conv_first_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     tcf::kSmemBytes));
    DEQSCI_CUDA(launch_pdl(tcf::conv_first_tc_kernel<true>, (unsigned)grid, tcf::kThreads, tcf::kSmemBytes, st, in_map,
                           out_hi, out_lo, p));
  } else {
    DEQSCI_CUDA(cudaFuncSetAttribute(tcf::conv_first_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     tcf::kSmemBytes));
    DEQSCI_CUDA(launch_pdl(tcf::conv_first_tc_kernel<false>, (unsigned)grid, tcf::kThreads, tcf::kSmemBytes, st, in_map,
                           out_hi, out_lo, p));
  }
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

}  // namespace deqsci
