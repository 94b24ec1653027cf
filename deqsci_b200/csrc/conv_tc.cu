// Kernel group 2b: hidden 3x3 conv layers (64 -> 64 channels) as a tcgen05 / TMEM implicit GEMM
// fed by TMA (sm_100a only).
//
// GEMM view per layer:  D[M = pixels, N = 64 cout] = sum over 9 taps of  A_tap[M, 64 cin] * W_tap[64 cin, N]
//   A_tap is the channels-last activation tile shifted by (ky-1, kx-1): one 4-D TMA box
//   {64 ch, TWm, THm, 1 frame} per tap, zero padding = TMA out-of-bounds fill, frames never mix.
//   One CTA tile = 128 pixels (UMMA_M = 128, cta_group::1), K per tap = 64 = one 128-byte swizzle row.
//
// Precision (the parity mode, DEQSCI_PREC_TC_SPLIT): activations and weights are fp16 pairs
//   v = hi + lo' * 2^-11.  D = Ah*Wh + 2^-11 (Ah*Wl' + Al'*Wh) with fp32 accumulation in TMEM:
//     MMA 1: A = Ah, B = [Wh | Wl'] (N = 128)  -> columns [0,64) main, [64,128) correction
//     MMA 2: A = Al', B = Wh        (N =  64)  -> accumulates into the correction columns
//   i.e. 3 products for 2 instructions (the dropped Al*Wl term is O(2^-22)).  The epilogue combines
//   main + 2^-11 * correction, applies the folded-BatchNorm affine + ReLU and re-splits to hi/lo.
//
// Warp roles (192 threads, persistent over tiles, one CTA per SM):
//   warp 0    : TMA producer (weights once via 1-D bulk copies; per tap the hi and lo A boxes)
//   warp 1    : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2-5 : epilogue (tcgen05.ld -> affine/ReLU/split -> global stores), one TMEM lane quarter each
// Pipelines: smem A ring (full/empty mbarriers, tcgen05.commit frees a slot) and a double-buffered
// TMEM accumulator (tmem_full/tmem_empty) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>   // CUtensorMap types only; the encode entry point is fetched at run time
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace deqsci {

constexpr int kTileM = 128;
constexpr int kTapBytesA = kTileM * 128;        // 16 KB: 128 pixels x 64 fp16
constexpr int kTcThreads = 192;

// How a CTA brings the A operand (activation tile + 3x3 neighbourhood) into shared memory:
//   LD_TAP  : one TMA box per tap (9 per tile).  Any tile shape; used for small images where a
//             128-pixel tile spans several rows.
//   LD_ROW3 : tiles are 128 consecutive pixels of one image row.  A stage is one input ROW with a
//             one-pixel halo each side (130 pixels, hi and lo planes); the three kx taps are three
//             UMMA descriptors into it, 128 bytes (one pixel) apart: 3 row loads per tile.
//   LD_ROLL : as LD_ROW3, but a CTA walks down a strip of R consecutive output rows and keeps the
//             input rows in a ring of >= 4 slots: every input row is loaded once and used by the
//             three output rows around it (ky = 2, 1, 0): (R + 2) / R row loads per tile.
// LD_TAP is bound by L2->SM bandwidth (9 x 32 KB per 3456 MMA cycles), LD_ROW3 cuts that 3x, LD_ROLL
// ~7x; which one runs is decided by how many ring slots fit beside the resident weights.
enum { LD_TAP = 0, LD_ROW3 = 1, LD_ROLL = 2 };

// MODE: 0 = hidden layer (64 -> 64, writes activation planes);
//       1 = FFDNet last layer (64 -> 4, pixel-shuffle, out = z' - noise in the cube layout);
//       2 = DnCNN last layer (64 -> 1, out = z' - noise).  Last layers pad cout to N = 16.
enum { TC_HIDDEN = 0, TC_LAST_FFD = 1, TC_LAST_DN = 2 };

template <bool SPLIT, int LOAD, int MODE = TC_HIDDEN>
struct TcCfg {
  static constexpr bool kRow = (LOAD != LD_TAP);
  static constexpr int kNOut = (MODE == TC_HIDDEN) ? 64 : 16; // UMMA N of one product (M = 128 needs N % 16 == 0)
  static constexpr int kCout = (MODE == TC_HIDDEN) ? 64 : (MODE == TC_LAST_FFD ? 4 : 1);
  static constexpr int kBRows = SPLIT ? 2 * kNOut : kNOut;    // rows of the B tile per tap: [hi | lo']
  static constexpr int kTapBytesB = kBRows * 128;             // 16 KB / 8 KB (hidden)
  static constexpr int kWBytes = 9 * kTapBytesB;              // 144 KB / 72 KB (hidden)
  static constexpr int kRowPix = kRow ? kTileM + 2 : kTileM;  // pixels per stage plane
  static constexpr int kPlaneBytes = kRow ? 17 * 1024 : kTapBytesA;   // 130*128 B rounded up to the 1 KB swizzle atom
  static constexpr int kStageBytes = (SPLIT ? 2 : 1) * kPlaneBytes;
  static constexpr int kTxBytes = (SPLIT ? 2 : 1) * kRowPix * 128;     // bytes TMA delivers per stage
  static constexpr int kSteps = kRow ? 3 : 9;                 // stages consumed per tile (LD_TAP / LD_ROW3)
  static constexpr int kStagesFit = (227 * 1024 - 2048 - kWBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
  static constexpr int kAccCols = SPLIT ? 2 * kNOut : kNOut;  // TMEM columns per accumulator buffer
  static constexpr int kTmemCols = 2 * kAccCols < 32 ? 32 : 2 * kAccCols;   // power of two >= 32
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kWBytes + kStages * kStageBytes + 1024 /*barriers, affine*/;
  static_assert(LOAD != LD_ROLL || kStages >= 4, "rolling rows need a ring of at least 4 slots");
};

using namespace ptx;   // mbarrier / TMA / tcgen05 wrappers (tc_ptx.cuh)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t a) { return sdesc_sw128(a); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}

struct TcParams {
  const uint8_t* wimg;      // pre-swizzled weight image, TcCfg::kWBytes
  const float* scale;       // [cout] or null
  const float* bias;        // [cout] or null
  __half* out_hi;           // output planes (channels-last), hidden mode
  __half* out_lo;
  int relu;
  int NF, Hc, Wc;
  int TWm, THm;             // tile = THm rows x TWm cols, TWm*THm == 128
  int tiles_x, tiles_y;     // LD_ROLL: tiles_y counts strips of `strip_rows` output rows
  int strip_rows;           // LD_ROLL: output rows per work item (1 otherwise)
  long long n_items;
  // last-layer modes: residual epilogue in the cube layout [B,H,W,T]
  const float* zprime;
  float* out_cube;
  int H, W, T;
};

struct TcItem { int nf, h0, w0, ntiles; };

template <int LOAD>
__device__ __forceinline__ TcItem tc_decode(const TcParams& p, long long item) {
  const int per_frame = p.tiles_x * p.tiles_y;
  TcItem it;
  it.nf = (int)(item / per_frame);
  const int rem = (int)(item - (long long)it.nf * per_frame);
  const int ty = rem / p.tiles_x;
  it.w0 = (rem - ty * p.tiles_x) * p.TWm;
  if (LOAD == LD_ROLL) {
    it.h0 = ty * p.strip_rows;
    it.ntiles = min(p.strip_rows, p.Hc - it.h0);
  } else {
    it.h0 = ty * p.THm;
    it.ntiles = 1;
  }
  return it;
}

template <bool SPLIT, int LOAD, int MODE>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_mid_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                   const TcParams p) {
  using Cfg = TcCfg<SPLIT, LOAD, MODE>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* w_s = smem;
  uint8_t* a_s = smem + Cfg::kWBytes;
  uint8_t* tail = a_s + S * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);       // [0] w_full, then full[S], empty[S], tfull[2], tempty[2]
  float* aff_s = reinterpret_cast<float*>(tail + 256);      // scale[64], bias[64]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 256 + 512);

  const uint32_t bar_w = smem_u32(&bars[0]);
  auto bar_full = [&](int s) { return smem_u32(&bars[1 + s]); };
  auto bar_empty = [&](int s) { return smem_u32(&bars[1 + S + s]); };
  auto bar_tfull = [&](int b) { return smem_u32(&bars[1 + 2 * S + b]); };
  auto bar_tempty = [&](int b) { return smem_u32(&bars[3 + 2 * S + b]); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < S; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), 4); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) {
    const int c = threadIdx.x - 64;
    const bool real = c < Cfg::kCout;
    aff_s[c] = (real && p.scale) ? p.scale[c] : 1.f;
    aff_s[64 + c] = (real && p.bias) ? p.bias[c] : 0.f;
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(bar_w, Cfg::kWBytes);
      for (int t = 0; t < 9; ++t)
        bulk_load_1d(smem_u32(w_s + t * Cfg::kTapBytesB), p.wimg + (size_t)t * Cfg::kTapBytesB, Cfg::kTapBytesB, bar_w);
      int stage = 0;
      uint32_t phase = 0;
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const TcItem it = tc_decode<LOAD>(p, item);
        // LD_TAP: 9 tap boxes; LD_ROW3: 3 halo rows; LD_ROLL: the strip's ntiles + 2 halo rows, each once
        const int n_loads = (LOAD == LD_ROLL) ? it.ntiles + 2 : Cfg::kSteps;
        for (int q = 0; q < n_loads; ++q) {
          int dy, dx;
          if (LOAD == LD_TAP) { dy = q / 3 - 1; dx = q - (q / 3) * 3 - 1; }
          else                { dy = q - 1; dx = -1; }          // row boxes start one pixel left of the tile
          mbar_wait(bar_empty(stage), phase ^ 1);
          mbar_arrive_expect_tx(bar_full(stage), Cfg::kTxBytes);
          const uint32_t dst = smem_u32(a_s + stage * Cfg::kStageBytes);
          tma_load_4d(dst, &map_hi, bar_full(stage), 0, it.w0 + dx, it.h0 + dy, it.nf);
          if (SPLIT) tma_load_4d(dst + Cfg::kPlaneBytes, &map_lo, bar_full(stage), 0, it.w0 + dx, it.h0 + dy, it.nf);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one_sync()) {
      constexpr uint32_t idesc_main = make_idesc(kTileM, Cfg::kBRows);
      constexpr uint32_t idesc_lo = make_idesc(kTileM, Cfg::kNOut);
      mbar_wait(bar_w, 0);
      const uint32_t a_base = smem_u32(a_s), w_base = smem_u32(w_s);
      // the MMAs of one tap: A = 128 pixel rows starting at a_addr (hi plane; lo plane kPlaneBytes further)
      auto issue_tap = [&](uint32_t d_main, uint32_t a_addr, int tap) {
        const uint64_t a_hi = make_sdesc(a_addr);
        const uint64_t a_lo = make_sdesc(a_addr + Cfg::kPlaneBytes);
        const uint64_t b_w = make_sdesc(w_base + tap * Cfg::kTapBytesB);
#pragma unroll
        for (int k = 0; k < 4; ++k) {                 // K = 64 per tap = 4 x UMMA_K(16); +32 bytes per step
          umma_f16(d_main, a_hi + 2 * k, b_w + 2 * k, idesc_main, (tap | k) != 0);
          if (SPLIT) umma_f16(d_main + Cfg::kNOut, a_lo + 2 * k, b_w + 2 * k, idesc_lo, 1u);
        }
      };
      int stage = 0;          // LD_TAP / LD_ROW3: ring position; LD_ROLL: slot of the strip's first row
      uint32_t phase = 0;
      int buf = 0;
      uint32_t tphase = 0;
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const TcItem it = tc_decode<LOAD>(p, item);
        if (LOAD == LD_ROLL) {
          // rows q = 0 .. ntiles+1 of the strip live in slots (stage + q) % S, in load order
          int wait_stage = stage;
          uint32_t wait_phase = phase;
          int rows_ready = 0;
          for (int j = 0; j < it.ntiles; ++j) {
            mbar_wait(bar_tempty(buf), tphase ^ 1);
            while (rows_ready < j + 3) {              // output row j needs input rows j, j+1, j+2
              mbar_wait(bar_full(wait_stage), wait_phase);
              if (++wait_stage == S) { wait_stage = 0; wait_phase ^= 1; }
              ++rows_ready;
            }
            tc_fence_after();
            const uint32_t d_main = tmem_base + buf * Cfg::kAccCols;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              int slot = stage + j + ky;
              slot -= (slot / S) * S;
              const uint32_t a_addr = a_base + slot * Cfg::kStageBytes;
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) issue_tap(d_main, a_addr + kx * 128, ky * 3 + kx);
            }
            {
              int slot = stage + j;
              slot -= (slot / S) * S;
              umma_commit(bar_empty(slot));           // input row j is dead after output row j
              if (j == it.ntiles - 1) {               // end of strip: its last two rows too
                int s1 = slot + 1 == S ? 0 : slot + 1;
                umma_commit(bar_empty(s1));
                umma_commit(bar_empty(s1 + 1 == S ? 0 : s1 + 1));
              }
            }
            umma_commit(bar_tfull(buf));
            if (++buf == 2) { buf = 0; tphase ^= 1; }
          }
          stage = wait_stage;                         // next strip starts after this strip's ntiles+2 rows
          phase = wait_phase;
        } else {
          mbar_wait(bar_tempty(buf), tphase ^ 1);     // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_main = tmem_base + buf * Cfg::kAccCols;
          for (int step = 0; step < Cfg::kSteps; ++step) {
            mbar_wait(bar_full(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = a_base + stage * Cfg::kStageBytes;
            if (LOAD == LD_ROW3) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) issue_tap(d_main, a_addr + kx * 128, step * 3 + kx);
            } else {
              issue_tap(d_main, a_addr, step);
            }
            umma_commit(bar_empty(stage));            // frees the smem slot when these MMAs retire
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
          umma_commit(bar_tfull(buf));                // accumulator complete -> epilogue
          if (++buf == 2) { buf = 0; tphase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;                     // TMEM lanes [32*quarter, 32*quarter+32)
    const int m = quarter * 32 + lane;                // pixel row of the tile owned by this thread
    int buf = 0;
    uint32_t tphase = 0;
    for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
     const TcItem it = tc_decode<LOAD>(p, item);
     const int nf = it.nf;
     for (int j = 0; j < it.ntiles; ++j) {
      const int h = (LOAD == LD_TAP) ? it.h0 + m / p.TWm : it.h0 + j;
      const int w = (LOAD == LD_TAP) ? it.w0 + m % p.TWm : it.w0 + m;
      const bool inside = (h < p.Hc && w < p.Wc);
      const long long off = (((long long)nf * p.Hc + h) * p.Wc + w) * 64;
      mbar_wait(bar_tfull(buf), tphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * Cfg::kAccCols;
      if (MODE != TC_HIDDEN) {
        // last layer: cout values per pixel -> out = z' - noise, scattered through the pixel shuffle
        uint32_t acc[4], cor[4];
        tmem_ld4(t_addr, acc);
        if (SPLIT) tmem_ld4(t_addr + Cfg::kNOut, cor);
        tmem_ld_wait();
        if (inside) {
          const int b = nf / p.T, t = nf - b * p.T;
#pragma unroll
          for (int c = 0; c < Cfg::kCout; ++c) {
            float a = __uint_as_float(acc[c]);
            if (SPLIT) a = fmaf(__uint_as_float(cor[c]), kLoInvScale, a);
            a = fmaf(a, aff_s[c], aff_s[64 + c]);
            if (p.relu) a = fmaxf(a, 0.f);
            long long g;
            if (MODE == TC_LAST_FFD) g = (((long long)b * p.H + 2 * h + (c >> 1)) * p.W + 2 * w + (c & 1)) * p.T + t;
            else                     g = (((long long)b * p.H + h) * p.W + w) * p.T + t;
            p.out_cube[g] = __fsub_rn(p.zprime[g], a);
          }
        }
      } else
#pragma unroll
      for (int half = 0; half < 2; ++half) {          // 32 output channels at a time
        uint32_t acc[32], cor[32];
        tmem_ld32(t_addr + half * 32, acc);
        if (SPLIT) tmem_ld32(t_addr + 64 + half * 32, cor);
        tmem_ld_wait();
        uint32_t hi_pk[16], lo_pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float v[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int c = half * 32 + e + u;
            float a = __uint_as_float(acc[e + u]);
            if (SPLIT) a = fmaf(__uint_as_float(cor[e + u]), kLoInvScale, a);
            a = fmaf(a, aff_s[c], aff_s[64 + c]);
            v[u] = p.relu ? fmaxf(a, 0.f) : a;
          }
          split_f16x2(v[0], v[1], hi_pk[e >> 1], lo_pk[e >> 1]);
        }
        if (inside) {
          uint4* dh = reinterpret_cast<uint4*>(p.out_hi + off + half * 32);
          uint4* dl = reinterpret_cast<uint4*>(p.out_lo + off + half * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            dh[q] = make_uint4(hi_pk[4 * q], hi_pk[4 * q + 1], hi_pk[4 * q + 2], hi_pk[4 * q + 3]);
            dl[q] = make_uint4(lo_pk[4 * q], lo_pk[4 * q + 1], lo_pk[4 * q + 2], lo_pk[4 * q + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (elect_one_sync()) mbar_arrive(bar_tempty(buf));    // 4 arrivals (one per epilogue warp) free the buffer
      if (++buf == 2) { buf = 0; tphase ^= 1; }
     }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ---- host side -------------------------------------------------------------------------------
void tc_tile_shape(int Wc, int* TWm, int* THm) {
  int tw = 128;
  while (tw > 8 && tw / 2 >= Wc) tw /= 2;     // smallest power of two >= Wc, clamped to [8,128]
  *TWm = tw;
  *THm = kTileM / tw;
}

size_t tc_weight_image_bytes(bool split, int cout) {
  const int n_out = cout == 64 ? 64 : 16;
  return (size_t)9 * (split ? 2 : 1) * n_out * 128;
}

// Host packing of one layer: w [cout][64 cin][3][3] fp32 -> the exact shared-memory image the kernel
// bulk-copies: per tap a K-major tile [rows n][64 k=cin] fp16 with the 128-byte swizzle (16-byte chunk
// index XOR (row & 7)); rows [0,N) = hi(W) (zero rows for n >= cout), rows [N,2N) = lo'(W) (split
// mode), N = 64 for hidden layers, 16 for the last layer.
template <class Emit>
static void tc_layout(int cout, bool split, Emit emit) {
  const int n_out = cout == 64 ? 64 : 16;
  const int rows = split ? 2 * n_out : n_out;
  for (int tap = 0; tap < 9; ++tap) {
    const int ky = tap / 3, kx = tap % 3;
    for (int n = 0; n < rows; ++n) {
      const int co = n % n_out;
      if (co >= cout) continue;
      for (int k = 0; k < 64; ++k) {
        const size_t byte = (size_t)tap * rows * 128 + (size_t)n * 128 + (size_t)(((k >> 3) ^ (n & 7)) << 4) +
                            (size_t)(k & 7) * 2;
        emit(byte, ((co * 64 + k) * 3 + ky) * 3 + kx, n >= n_out);
      }
    }
  }
}
void tc_pack_weights(const float* w, int cout, bool split, uint8_t* img) {
  memset(img, 0, tc_weight_image_bytes(split, cout));
  tc_layout(cout, split, PackWrite{w, img});
}
void tc_pack_map(int cout, bool split, int32_t* map) { tc_layout(cout, split, PackMap{map}); }

template <bool SPLIT, int LOAD, int MODE>
static int launch_tc(const CUtensorMap& map_hi, const CUtensorMap& map_lo, const TcParams& p, int grid,
                     cudaStream_t st) {
  using Cfg = TcCfg<SPLIT, LOAD, MODE>;
  DEQSCI_CUDA(cudaFuncSetAttribute(conv_mid_tc_kernel<SPLIT, LOAD, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg::kSmemBytes));
  conv_mid_tc_kernel<SPLIT, LOAD, MODE><<<grid, kTcThreads, Cfg::kSmemBytes, st>>>(map_hi, map_lo, p);
  return DEQSCI_OK;
}

// Load mode by what fits: a split-precision hidden layer keeps 144 KB of weights resident, which
// leaves 2 row slots (LD_ROW3); single-pass hidden layers and the N = 16 last layers have room for a
// ring of 6 (LD_ROLL).  DEQSCI_TC_LOAD = 0 | 1 | 2 caps the mode (testing).
static int pick_load_mode(int mode, bool split, int TWm) {
  static const int cap = env_int("DEQSCI_TC_LOAD", LD_ROLL);
  int lm = LD_TAP;
  if (TWm == kTileM) lm = (mode == TC_HIDDEN && split) ? LD_ROW3 : LD_ROLL;
  return lm < cap ? lm : cap;
}

// mode: TC_HIDDEN writes act_out planes; TC_LAST_* writes out_cube = zprime - noise ([B,H,W,T] fp32).
int conv_tc_launch(int mode, bool split, const __half* act_in, __half* act_out, long long plane_elems,
                   const uint8_t* wimg, const float* scale, const float* bias, int relu, int NF, int Hc, int Wc,
                   const float* zprime, float* out_cube, int H, int W, int T, cudaStream_t st) {
  int TWm, THm;
  tc_tile_shape(Wc, &TWm, &THm);
  const int lm = pick_load_mode(mode, split, TWm);
  CUtensorMap map_hi, map_lo;
  int rc = make_plane_map(&map_hi, act_in, 64, NF, Hc, Wc, 64, lm != LD_TAP ? kTileM + 2 : TWm, THm, 128);
  if (rc) return rc;
  rc = make_plane_map(&map_lo, act_in + plane_elems, 64, NF, Hc, Wc, 64, lm != LD_TAP ? kTileM + 2 : TWm, THm, 128);
  if (rc) return rc;
  TcParams p;
  p.wimg = wimg; p.scale = scale; p.bias = bias;
  p.out_hi = act_out; p.out_lo = act_out ? act_out + plane_elems : nullptr;
  p.relu = relu; p.NF = NF; p.Hc = Hc; p.Wc = Wc; p.TWm = TWm; p.THm = THm;
  p.tiles_x = (Wc + TWm - 1) / TWm;
  p.strip_rows = 1;
  if (lm == LD_ROLL) {
    const int R = pick_strip_rows(NF, p.tiles_x, Hc, false, 6LL * num_sms(), 2);
    p.strip_rows = R;
    p.tiles_y = (Hc + R - 1) / R;
  } else {
    p.tiles_y = (Hc + THm - 1) / THm;
  }
  p.n_items = (long long)NF * p.tiles_x * p.tiles_y;
  p.zprime = zprime; p.out_cube = out_cube; p.H = H; p.W = W; p.T = T;
  const int grid = (int)(p.n_items < num_sms() ? p.n_items : num_sms());
  ProfScope prof(mode == TC_HIDDEN ? PK_CONV_HIDDEN : PK_CONV_LAST, st);
#define TC_DISPATCH(SP, LM)                                                                       \
  (mode == TC_HIDDEN     ? launch_tc<SP, LM, TC_HIDDEN>(map_hi, map_lo, p, grid, st)               \
   : mode == TC_LAST_FFD ? launch_tc<SP, LM, TC_LAST_FFD>(map_hi, map_lo, p, grid, st)             \
                         : launch_tc<SP, LM, TC_LAST_DN>(map_hi, map_lo, p, grid, st))
  if (split) {
    if (lm == LD_TAP) rc = TC_DISPATCH(true, LD_TAP);
    else if (lm == LD_ROW3 || mode == TC_HIDDEN) rc = TC_DISPATCH(true, LD_ROW3);
    else rc = (mode == TC_LAST_FFD ? launch_tc<true, LD_ROLL, TC_LAST_FFD>(map_hi, map_lo, p, grid, st)
                                   : launch_tc<true, LD_ROLL, TC_LAST_DN>(map_hi, map_lo, p, grid, st));
  } else {
    if (lm == LD_TAP) rc = TC_DISPATCH(false, LD_TAP);
    else if (lm == LD_ROW3) rc = TC_DISPATCH(false, LD_ROW3);
    else rc = TC_DISPATCH(false, LD_ROLL);
  }
#undef TC_DISPATCH
  if (rc) return rc;
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

int conv_mid_tc_launch(bool split, const __half* act_in, __half* act_out, long long plane_elems,
                       const uint8_t* wimg, const float* scale, const float* bias, int relu, int NF, int Hc, int Wc,
                       cudaStream_t st) {
  return conv_tc_launch(TC_HIDDEN, split, act_in, act_out, plane_elems, wimg, scale, bias, relu, NF, Hc, Wc,
                        nullptr, nullptr, 0, 0, 1, st);
}

}  // namespace deqsci
