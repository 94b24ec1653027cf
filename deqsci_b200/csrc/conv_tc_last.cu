// Kernel group 2e: the LAST conv layer (64 -> 4 FFDNet / 64 -> 1 DnCNN) on tensor cores,
// "ky-transposed", fused with the pixel-shuffle and the residual subtract out = z' - noise.
//
// A 3x3 conv with cout = 4 issued tap by tap (conv_tc.cu, N padded to 16) reads the 128x64 A tile
// from shared memory 9 x 2 times per output row and is bound by exactly that (tensor pipe 22 %).
// Here each INPUT row is pushed through the tensor core once:
//     V_q[w][ky*cout + c] = sum_kx sum_k W[c][k][ky][kx] * a[q][w + kx - 1][k]       (N = 3*cout <= 16)
// i.e. the three kx taps are three A descriptors 128 B apart (as everywhere else) but the three ky
// taps are three groups of OUTPUT columns.  The conv output of row r is then
//     noise[r][w][c] = V_{r-1}[w][0*cout + c] + V_r[w][1*cout + c] + V_{r+1}[w][2*cout + c]
// -- same pixel = same TMEM lane, three consecutive accumulator buffers -- which the epilogue adds
// while it applies the affine, the pixel shuffle (networks/ffdnet/functions.py:63-81) and
// z' - noise (solvers/equilibrium_solvers_yaping.py:417,420).  24 MMAs per row instead of 72.
//
// Two CTAs per SM (82 KB of shared memory, 256 TMEM columns each) hide each other's pipeline latencies.
// Warp roles (192 threads): warp 0 TMA producer (2-slot streaming ring of input rows, hi + lo planes,
// every row used once), warp 1 TMEM allocator + MMA issuer (ring of 8 accumulator buffers x 32
// columns: [16 main | 16 corr]), warps 2-5 epilogue.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace deqsci {
namespace tcl {
using namespace ptx;

constexpr int kTileM = 128;
constexpr int kThreads = 192;
constexpr int kSlots = 2;
constexpr int kPlaneBytes = 17 * 1024;              // 130 pixels x 128 B, padded to the 1 KB swizzle atom
constexpr int kSlotBytes = 2 * kPlaneBytes;
constexpr int kTxBytes = 2 * (kTileM + 2) * 128;
constexpr int kKxBytesB = 32 * 128;                 // per kx: [16 hi rows | 16 lo' rows] x 64 k
constexpr int kWBytes = 3 * kKxBytesB;              // 12 KB
constexpr int kBufs = 8;                            // accumulator ring capacity (input rows in flight: Params::bufs <= kBufs)
constexpr int kBufCols = 32;
constexpr int kTmemCols = kBufs * kBufCols;         // 256
constexpr int kSmemBytes = 1024 + kWBytes + kSlots * kSlotBytes + 1024;

struct Params {
  const uint8_t* wimg;
  const float* scale;
  const float* bias;
  int relu;
  int NF, Hc, Wc;
  int tiles_x, strips_y, strip_rows;
  long long n_items;
  const float* zprime;
  float* out_cube;
  int H, W, T;
  int bufs;                   // accumulator buffers in use (4..kBufs)
  int zp_planar;              // z' is frame-planar [B,T,H,W] (written by gap_prep for this kernel) instead of [B,H,W,T]
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

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}

template <int COUT>   // 4: FFDNet (pixel shuffle), 1: DnCNN
__global__ void __launch_bounds__(kThreads, 2)
conv_last_tc_kernel(const __grid_constant__ CUtensorMap in_hi, const __grid_constant__ CUtensorMap in_lo,
                    const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* w_s = smem;
  uint8_t* a_s = w_s + kWBytes;
  uint8_t* tail = a_s + kSlots * kSlotBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);          // [0] w, full[S], empty[S], tfull[4], tempty[4]
  float* aff_s = reinterpret_cast<float*>(tail + 256);         // scale[4], bias[4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 512);
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
    for (int b = 0; b < kBufs; ++b) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), 4); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 4) {
    const int c = threadIdx.x - 64;
    aff_s[c] = (c < COUT && p.scale) ? p.scale[c] : 1.f;
    aff_s[4 + c] = (c < COUT && p.bias) ? p.bias[c] : 0.f;
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: the strip's ntiles + 2 input rows, each loaded and used once
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
          tma_load_4d(dst, &in_hi, bar_full(slot), 0, it.w0 - 1, it.h0 - 1 + q, it.nf);
          tma_load_4d(dst + kPlaneBytes, &in_lo, bar_full(slot), 0, it.w0 - 1, it.h0 - 1 + q, it.nf);
          if (++slot == kSlots) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one accumulator buffer per input row
    if (elect_one_sync()) {
      constexpr uint32_t idesc_main = make_idesc(kTileM, 32);   // [Wh | Wl'] of one kx
      constexpr uint32_t idesc_lo = make_idesc(kTileM, 16);     // Wh
      mbar_wait(bar_w, 0);
      const uint32_t a_base = smem_u32(a_s), w_base = smem_u32(w_s);
      int slot = 0;
      uint32_t phase = 0;
      const uint32_t nb = (uint32_t)p.bufs;
      uint32_t g = 0;                                            // running input-row counter -> buffer ring
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const Item it = decode(p, item);
        for (int q = 0; q < it.ntiles + 2; ++q, ++g) {
          const int buf = g % nb;
          mbar_wait(bar_tempty(buf), ((g / nb) & 1) ^ 1);
          mbar_wait(bar_full(slot), phase);
          tc_fence_after();
          const uint32_t d = tmem_base + buf * kBufCols;
          const uint32_t a_row = a_base + slot * kSlotBytes;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint64_t a_hi = sdesc_sw128(a_row + kx * 128);
            const uint64_t a_lo = sdesc_sw128(a_row + kPlaneBytes + kx * 128);
            const uint64_t b_w = sdesc_sw128(w_base + kx * kKxBytesB);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16(d, a_hi + 2 * k, b_w + 2 * k, idesc_main, (kx | k) != 0);
              umma_f16(d + 16, a_lo + 2 * k, b_w + 2 * k, idesc_lo, 1u);
            }
          }
          umma_commit(bar_empty(slot));
          umma_commit(bar_tfull(buf));
          if (++slot == kSlots) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: out row j = V_j[ky=0] + V_{j+1}[ky=1] + V_{j+2}[ky=2]
    pdl_wait_predecessor();      // z' is read below without passing through the loader's wait
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t nb = (uint32_t)p.bufs;
    uint32_t g0 = 0;                                             // counter of the strip's first input row
    for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const Item it = decode(p, item);
      const int w = it.w0 + m;
      const int b = it.nf / p.T, t = it.nf - b * p.T;
      int rows_ready = 0;
      for (int j = 0; j < it.ntiles; ++j) {
        // z' for this output row does not depend on the MMAs: issue the loads before waiting on them
        const int h = it.h0 + j;
        long long gidx[COUT];
        float zp[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
          if (COUT == 4) gidx[c] = (((long long)b * p.H + 2 * h + (c >> 1)) * p.W + 2 * w + (c & 1)) * p.T + t;
          else           gidx[c] = (((long long)b * p.H + h) * p.W + w) * p.T + t;
        }
        if (p.zp_planar) {
          // frame-planar z': this lane's sub-pixels are two adjacent pairs, the warp's loads are contiguous
          if (COUT == 4) {
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
              const float2 v = (w < p.Wc) ? __ldg(reinterpret_cast<const float2*>(
                                                p.zprime + (((long long)b * p.T + t) * p.H + 2 * h + dy) * p.W + 2 * w))
                                          : make_float2(0.f, 0.f);
              zp[(dy * 2) % COUT] = v.x;
              zp[(dy * 2 + 1) % COUT] = v.y;
            }
          } else {
            zp[0] = (w < p.Wc) ? __ldg(p.zprime + (((long long)b * p.T + t) * p.H + h) * p.W + w) : 0.f;
          }
        } else {
#pragma unroll
          for (int c = 0; c < COUT; ++c) zp[c] = (w < p.Wc) ? __ldg(p.zprime + gidx[c]) : 0.f;
        }
        while (rows_ready < j + 3) {
          const uint32_t g = g0 + rows_ready;
          mbar_wait(bar_tfull(g % nb), (g / nb) & 1);
          ++rows_ready;
        }
        tc_fence_after();
        uint32_t acc[3][4], cor[3][4];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const uint32_t t_addr = lane_base + ((g0 + j + ky) % nb) * kBufCols;
          const int col = (COUT == 4) ? 4 * ky : 0;              // COUT == 1: columns 0,1,2 sit in one x4 load
          tmem_ld4(t_addr + col, acc[ky]);
          tmem_ld4(t_addr + 16 + col, cor[ky]);
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (elect_one_sync()) {
          mbar_arrive(bar_tempty((g0 + j) % nb));              // V_j is dead after output row j
          if (j == it.ntiles - 1) {                               // end of strip: its last two rows too
            mbar_arrive(bar_tempty((g0 + j + 1) % nb));
            mbar_arrive(bar_tempty((g0 + j + 2) % nb));
          }
        }
        if (w < p.Wc) {
#pragma unroll
          for (int c = 0; c < COUT; ++c) {
            float noise = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const int e = (COUT == 4) ? c : ky;
              noise += fmaf(__uint_as_float(cor[ky][e]), kLoInvScale, __uint_as_float(acc[ky][e]));
            }
            float a = fmaf(noise, aff_s[c], aff_s[4 + c]);
            if (p.relu) a = fmaxf(a, 0.f);
            p.out_cube[gidx[c]] = __fsub_rn(zp[c], a);
          }
        }
      }
      g0 += it.ntiles + 2;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace tcl

size_t tcl_weight_image_bytes() { return tcl::kWBytes; }

// w [cout][64][3][3] fp32 (cout = 4 or 1) -> [kx][32 rows][64 k] fp16, 128-byte swizzle;
// row n < 16: hi(W[c][k][ky][kx]) with n = ky*cout + c (zero for n >= 3*cout); row 16+n: lo'.
template <class Emit>
static void tcl_layout(int cout, Emit emit) {
  for (int kx = 0; kx < 3; ++kx)
    for (int n = 0; n < 32; ++n) {
      const int nn = n & 15;
      if (nn >= 3 * cout) continue;
      const int ky = nn / cout, c = nn % cout;
      for (int k = 0; k < 64; ++k) {
        const size_t byte = (size_t)kx * tcl::kKxBytesB + (size_t)n * 128 + (size_t)(((k >> 3) ^ (n & 7)) << 4) +
                            (size_t)(k & 7) * 2;
        emit(byte, ((c * 64 + k) * 3 + ky) * 3 + kx, n >= 16);
      }
    }
}
void tcl_pack_weights(const float* w, int cout, uint8_t* img) {
  memset(img, 0, tcl::kWBytes);
  tcl_layout(cout, PackWrite{w, img});
}
void tcl_pack_map(int cout, int32_t* map) { tcl_layout(cout, PackMap{map}); }

bool tcl_supported(int Wc) {
  static const int enabled = env_int("DEQSCI_TC_LAST", 1);
  return enabled && Wc > 64;
}

int conv_last_tc_launch(int cout, const __half* act_in, long long plane_elems, const uint8_t* wimg,
                        const float* scale, const float* bias, int relu, int NF, int Hc, int Wc,
                        const float* zprime, bool zprime_planar, float* out_cube, int H, int W, int T,
                        cudaStream_t st) {
  tcl::Params p;
  p.wimg = wimg; p.scale = scale; p.bias = bias; p.relu = relu;
  p.NF = NF; p.Hc = Hc; p.Wc = Wc;
  p.tiles_x = (Wc + tcl::kTileM - 1) / tcl::kTileM;
  const int R = pick_strip_rows_balanced(NF, p.tiles_x, Hc, false, 2 * num_sms(), 1, 4, 2);
  p.strip_rows = R;
  p.strips_y = (Hc + R - 1) / R;
  p.n_items = (long long)NF * p.tiles_x * p.strips_y;
  p.zprime = zprime; p.out_cube = out_cube; p.H = H; p.W = W; p.T = T;
  p.zp_planar = zprime_planar ? 1 : 0;
  static const int bufs = env_int("DEQSCI_TCL_BUFS", tcl::kBufs);
  p.bufs = bufs < 4 ? 4 : (bufs > tcl::kBufs ? tcl::kBufs : bufs);
  CUtensorMap in_hi, in_lo;
  int rc;
  if ((rc = make_plane_map(&in_hi, act_in, 64, NF, Hc, Wc, 64, tcl::kTileM + 2, 1, 128))) return rc;
  if ((rc = make_plane_map(&in_lo, act_in + plane_elems, 64, NF, Hc, Wc, 64, tcl::kTileM + 2, 1, 128))) return rc;
  const int grid = (int)(p.n_items < 2 * num_sms() ? p.n_items : 2 * num_sms());      // two CTAs per SM
  ProfScope prof(PK_CONV_LAST, st);
  if (cout == 4) {
    DEQSCI_CUDA(cudaFuncSetAttribute(tcl::conv_last_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     tcl::kSmemBytes));
    DEQSCI_CUDA(launch_pdl(tcl::conv_last_tc_kernel<4>, (unsigned)grid, tcl::kThreads, tcl::kSmemBytes, st, in_hi, in_lo, p));
  } else {
    DEQSCI_CUDA(cudaFuncSetAttribute(tcl::conv_last_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     tcl::kSmemBytes));
    DEQSCI_CUDA(launch_pdl(tcl::conv_last_tc_kernel<1>, (unsigned)grid, tcl::kThreads, tcl::kSmemBytes, st, in_hi, in_lo, p));
  }
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

}  // namespace deqsci
