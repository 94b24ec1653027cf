// Kernel group 2c: hidden 64->64 3x3 conv layer, split-fp16 precision, as a CTA-PAIR kernel
// (thread-block cluster of 2, tcgen05 cta_group::2, UMMA_M = 256).
//
// Why a pair: the split-precision weights [Wh | Wl'] of one layer are 144 KB.  Resident in ONE CTA
// they leave room for only two activation-row slots and no store staging, so the single-CTA kernel
// (conv_tc.cu, LD_ROW3) is starved by TMA latency and its epilogue pays 32 partial-line stores per
// instruction.  cta_group::2 lets each CTA of the pair keep HALF of the B operand (72 KB; 108 KB in
// the default issue mode): that frees shared memory for a rolling row ring (every input row is loaded
// once per strip and used by the three output rows around it) plus staging buffers for coalesced TMA
// stores, and halves the B-operand shared-memory reads per SM.
//
// Each CTA of the pair owns its own strip of R consecutive output rows (128-pixel row segments) of
// some frame; the two strips only share the weights.  out = main + 2^-11 (corr1 + corr2) with
//     main = Ah x Wh,   corr1 = Ah x Wl',   corr2 = Al' x Wh        (three fp16 products per fp32-accurate one)
//
// Default issue order (MODE 2, "row-stationary wide"; DESIGN.md finding 17): the issuer walks the strip's R + 2 INPUT
// rows.  A slice (row q, tap column kx, 16 channels k) of the hi plane is multiplied with the weights of the three output
// rows it feeds (ky = 0, 1, 2 -> rows q, q-1, q-2) in three consecutive N = 128 instructions (B = [Wh | Wl'], 64 rows from
// each CTA), the same slice of the lo plane in three N = 64 instructions (B = Wh, 32 rows from each CTA): the A tile is
// fetched from shared memory once per three instructions (collector hints).  Accumulator of an output row, 128 columns:
//     [0,32) main 0-31 | [32,96) corr1 + corr2 0-63 | [96,128) main 32-63
// -- the N = 64 instruction writes the middle 64 columns of the N = 128 one's 128, which fixes the weight-tile layout
// (tc2_layout below: 96 rows per tap and CTA, one 32-row block of Wh stored twice).  Four accumulators: three rows in
// flight, one draining.  MODE 1 issues nine N = 64 instructions per slice on the 64-row tiles; MODE 0 is the round-1
// output-stationary order (per output row: 9 taps x {N = 128 on Ah, N = 64 on Al'}, accumulators of 192 columns: main /
// corr1 interleaved in [0,128), corr2 in [128,192), added in the epilogue).
//
// Warp roles per CTA (320 threads): warp 0 TMA producer, warp 1 TMEM alloc (+ MMA issuer in the
// leader CTA only), warps 2-9 epilogue (TMEM lane quarter = warp % 4, channel half = (warp-2) / 4).
// Barriers: full[s] lives in the leader (both CTAs' TMA loads complete_tx on it), empty[s] and
// tmem_full[b] in both CTAs (multicast tcgen05.commit), tmem_empty[b] in the leader (16 arrivals).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace deqsci {

namespace tc2 {

constexpr int kTileM = 128;
constexpr int kThreads = 320;
constexpr int kSlotsMax = 4;                      // rolling ring of input rows (3 in the wide row-stationary mode)
constexpr int kPlaneBytes = 17 * 1024;            // 130 pixels x 128 B, rounded up to the 1 KB swizzle atom
constexpr int kSlotBytes = 2 * kPlaneBytes;       // hi + lo
constexpr int kTxBytes = 2 * (kTileM + 2) * 128;  // bytes one CTA's TMA delivers per row
constexpr int kStageBytes = 2048;                 // per epilogue warp: 32 pixels x 32 channels fp16
constexpr int kAccCols = 192;                     // output-stationary: main/corr1 interleaved (128) + corr2 (64), 2 buffers
constexpr int kAccColsRS = 128;                   // row-stationary: main (64) + both corrections (64), 4 buffers
constexpr int kAccBufsMax = 4;
constexpr int kTmemCols = 512;
// Issue modes (template parameter MODE of the kernel, DEQSCI_TC_RS):
//   0  output-stationary: per output row 9 taps x {N = 128 on Ah, N = 64 on Al'}
//   1  row-stationary, nine N = 64 instructions per (kx, k) slice, A-collector hints
//   2  row-stationary wide: three N = 128 (Ah x [Wh | Wl']) + three N = 64 (Al' x Wh) per slice, hints; needs the
//      96-row weight tiles (one 32-row block duplicated) and gives up one ring slot for them
__host__ __device__ constexpr int tap_rows(int mode) { return mode == 2 ? 96 : 64; }        // this CTA's B rows per tap
__host__ __device__ constexpr int tap_bytes(int mode) { return tap_rows(mode) * 128; }
__host__ __device__ constexpr int w_bytes(int mode) { return 9 * tap_bytes(mode); }         // 72 / 108 KB per CTA
__host__ __device__ constexpr int n_slots(int mode) { return mode == 2 ? 3 : 4; }
__host__ __device__ constexpr int smem_bytes(int mode) {
  return w_bytes(mode) + n_slots(mode) * kSlotBytes + 8 * kStageBytes + 1024;               // mode 2: exactly 227 KB
}

using namespace ptx;   // single-CTA mbarrier / TMA / tcgen05 wrappers (tc_ptx.cuh); below: the cluster / cta_group::2 forms

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `local_addr`'s twin in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on a barrier anywhere in the cluster (address from mapa)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  // default semantics (release at CTA scope): the TMEM reads it orders were already fenced with
  // tcgen05.fence::before_thread_sync; a .release.cluster here costs a GPU-scope MEMBAR + ERRBAR per tile
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA row load into THIS CTA's shared memory, completing on `cluster_bar` (the leader's full barrier)
__device__ __forceinline__ void tma_load_4d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0,
                                                 int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same with an A-operand collector hint: the tensor core keeps the A tile it fetched from shared memory
// (MODE 1 = fill, SASS .A_KEEP) and later instructions take it from there instead of reading shared memory again
// (2 = use, .A_REUSE.A_KEEP; 3 = lastuse, .A_REUSE).  ACC false overwrites D (first product of an accumulator).
template <int MODE, bool ACC>
__device__ __forceinline__ void umma2_f16_c(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
#define DEQSCI_UMMA2(Q)                                                                                   \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                        \
               "tcgen05.mma.cta_group::2.kind::f16" Q " [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),        \
               "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(ACC ? 1u : 0u)                                   \
               : "memory")
  if (MODE == 1) DEQSCI_UMMA2(".collector::a::fill");
  else if (MODE == 2) DEQSCI_UMMA2(".collector::a::use");
  else if (MODE == 3) DEQSCI_UMMA2(".collector::a::lastuse");
  else DEQSCI_UMMA2("");
#undef DEQSCI_UMMA2
}
// arrive (once all prior MMAs of this thread retired) on the barrier at the same offset in both CTAs
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ uint64_t make_sdesc(uint32_t a) { return sdesc_sw128(a); }

struct Params {
  const uint8_t* wimg;        // [2 ranks][9 taps][64 rows][128 B]
  const float* scale;
  const float* bias;
  int relu;
  int NF, Hc, Wc;
  int tiles_x, strips_y, strip_rows;
  long long n_strips;         // real strips; strip ids >= n_strips are padding (computed on a clamped strip, not stored)
  long long n_pair_items;
  double* stats;              // STATS kernels: [gridDim.x][128] per-CTA partials: channel sums, then sums of squares
  uint32_t lo_mask;           // experiment switch (DEQSCI_TC_LO_MASK): AND mask on each packed pair of lo' halves
  int debug_skip_store;       // experiment switch (DEQSCI_TC_DEBUG_SKIP_STORE): 1 = compute but do not store, 2 = direct st.global
  __half* dbg_out_hi;
  __half* dbg_out_lo;
  int lookahead;              // issuer tests the next row's barriers ahead of time (DEQSCI_TC_LOOKAHEAD, default 1)
  const __half* mask;         // relu == 2: output (pixel, channel) is kept where this hi plane [NF,Hc,Wc,64] is > 0, else
                              // zeroed: the ReLU derivative of a saved forward activation (adjoint / VJP stacks)
};

struct Strip { int nf, h0, w0; bool real; };

template <class P>
__device__ __forceinline__ Strip decode(const P& p, long long strip) {
  Strip s;
  s.real = strip < p.n_strips;
  if (!s.real) strip = p.n_strips - 1;
  const int per_frame = p.tiles_x * p.strips_y;
  s.nf = (int)(strip / per_frame);
  const int rem = (int)(strip - (long long)s.nf * per_frame);
  const int sy = rem / p.tiles_x;
  s.w0 = (rem - sy * p.tiles_x) * kTileM;
  s.h0 = sy * p.strip_rows;
  return s;
}


// Collector hint of the idx-th of `total` consecutive instructions that share one A tile.
__host__ __device__ constexpr int collector_mode(int idx, int total) {
  return total == 1 ? 0 : idx == 0 ? 1 : idx == total - 1 ? 3 : 2;
}

// Row-stationary issue of ONE input row (both CTAs' rows, M = 256).  Each (kx, k) slice of the row's hi plane is
// multiplied with Wh and Wl' of every output row the input row feeds (ky = 0, 1, 2 -> output rows q, q-1, q-2; L0..L2
// say which of them exist), the slice of the lo plane with their Wh -- consecutive instructions on the same A tile, so
// the tensor core fetches an activation tile from shared memory once instead of three times (collector hints).
// d0..d2: accumulators of those output rows, 128 columns: main and both correction products (same 2^-11 scale,
// one accumulator).  WIDE = false: nine N = 64 instructions per slice, columns [0,64) main, [64,128) corrections.
// WIDE = true: three N = 128 (Ah x 64 rows of each CTA, tile rows [32,96)) + three N = 64 (Al' x tile rows [0,32)),
// columns [0,32) main 0-31 | [32,96) corrections 0-63 | [96,128) main 32-63.
// KXB, KXE: the kx taps [KXB, KXE) of the row (the issuer splits a row in two to look ahead in between).
template <bool L0, bool L1, bool L2, bool WIDE, int KXB, int KXE>
__device__ __forceinline__ void issue_row_rs(uint32_t a_row, uint32_t w_base, uint32_t d0, uint32_t d1, uint32_t d2) {
  constexpr uint32_t idesc64 = make_idesc(256, 64), idesc128 = make_idesc(256, 128);
  constexpr int n_live = (L0 ? 1 : 0) + (L1 ? 1 : 0) + (L2 ? 1 : 0);
  constexpr int p1 = L0 ? 1 : 0, p2 = p1 + (L1 ? 1 : 0);       // position of ky = 1, 2 among the live rows
  constexpr uint64_t kRows32 = (32 * 128) >> 4;                // descriptor step of 32 tile rows
  constexpr int kTapB = tap_bytes(WIDE ? 2 : 1);
#pragma unroll
  for (int kx = KXB; kx < KXE; ++kx) {
    const uint64_t a_hi = make_sdesc(a_row + kx * 128);
    const uint64_t a_lo = make_sdesc(a_row + kPlaneBytes + kx * 128);
    const uint64_t b0 = make_sdesc(w_base + (0 * 3 + kx) * kTapB);
    const uint64_t b1 = make_sdesc(w_base + (1 * 3 + kx) * kTapB);
    const uint64_t b2 = make_sdesc(w_base + (2 * 3 + kx) * kTapB);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t ah = a_hi + 2 * k, al = a_lo + 2 * k;
      if (WIDE) {
        if (L0) {                     // ky = 0 opens the accumulator of output row q: its first product overwrites
          if (kx == 0 && k == 0) umma2_f16_c<collector_mode(0, n_live), false>(d0, ah, b0 + kRows32 + 2 * k, idesc128);
          else umma2_f16_c<collector_mode(0, n_live), true>(d0, ah, b0 + kRows32 + 2 * k, idesc128);
        }
        if (L1) umma2_f16_c<collector_mode(p1, n_live), true>(d1, ah, b1 + kRows32 + 2 * k, idesc128);
        if (L2) umma2_f16_c<collector_mode(p2, n_live), true>(d2, ah, b2 + kRows32 + 2 * k, idesc128);
        if (L0) umma2_f16_c<collector_mode(0, n_live), true>(d0 + 32, al, b0 + 2 * k, idesc64);
        if (L1) umma2_f16_c<collector_mode(p1, n_live), true>(d1 + 32, al, b1 + 2 * k, idesc64);
        if (L2) umma2_f16_c<collector_mode(p2, n_live), true>(d2 + 32, al, b2 + 2 * k, idesc64);
      } else {
        if (L0) {
          if (kx == 0 && k == 0) {
            umma2_f16_c<collector_mode(0, 2 * n_live), false>(d0, ah, b0 + 2 * k, idesc64);
            umma2_f16_c<collector_mode(1, 2 * n_live), false>(d0 + 64, ah, b0 + kRows32 + 2 * k, idesc64);
          } else {
            umma2_f16_c<collector_mode(0, 2 * n_live), true>(d0, ah, b0 + 2 * k, idesc64);
            umma2_f16_c<collector_mode(1, 2 * n_live), true>(d0 + 64, ah, b0 + kRows32 + 2 * k, idesc64);
          }
        }
        if (L1) {
          umma2_f16_c<collector_mode(2 * p1, 2 * n_live), true>(d1, ah, b1 + 2 * k, idesc64);
          umma2_f16_c<collector_mode(2 * p1 + 1, 2 * n_live), true>(d1 + 64, ah, b1 + kRows32 + 2 * k, idesc64);
        }
        if (L2) {
          umma2_f16_c<collector_mode(2 * p2, 2 * n_live), true>(d2, ah, b2 + 2 * k, idesc64);
          umma2_f16_c<collector_mode(2 * p2 + 1, 2 * n_live), true>(d2 + 64, ah, b2 + kRows32 + 2 * k, idesc64);
        }
        if (L0) umma2_f16_c<collector_mode(0, n_live), true>(d0 + 64, al, b0 + 2 * k, idesc64);
        if (L1) umma2_f16_c<collector_mode(p1, n_live), true>(d1 + 64, al, b1 + 2 * k, idesc64);
        if (L2) umma2_f16_c<collector_mode(p2, n_live), true>(d2 + 64, al, b2 + 2 * k, idesc64);
      }
    }
  }
}

// live: bit ky set = output row q - ky exists
template <bool WIDE, int KXB, int KXE>
__device__ __forceinline__ void issue_row_live(int live, uint32_t a_row, uint32_t w_base, uint32_t d0, uint32_t d1,
                                               uint32_t d2) {
  switch (live) {
    case 1: issue_row_rs<true, false, false, WIDE, KXB, KXE>(a_row, w_base, d0, d1, d2); break;
    case 2: issue_row_rs<false, true, false, WIDE, KXB, KXE>(a_row, w_base, d0, d1, d2); break;
    case 3: issue_row_rs<true, true, false, WIDE, KXB, KXE>(a_row, w_base, d0, d1, d2); break;
    case 4: issue_row_rs<false, false, true, WIDE, KXB, KXE>(a_row, w_base, d0, d1, d2); break;
    case 6: issue_row_rs<false, true, true, WIDE, KXB, KXE>(a_row, w_base, d0, d1, d2); break;
    default: issue_row_rs<true, true, true, WIDE, KXB, KXE>(a_row, w_base, d0, d1, d2); break;
  }
}

// Look-ahead barrier tests of the issuing thread.  A blocking mbarrier wait costs ~250 cycles even when the phase has
// long completed, and the tensor pipe's instruction queue is too shallow to cover two of them per input row (measured:
// 350-1200 idle cycles per row).  So the NEXT row's barriers are tested (mbarrier.test_wait, non-blocking) two thirds
// through the current row's instructions and the predicates are read after the last third: the latency hides behind
// MMA issue, and the blocking wait remains only as the fallback.  The predicates live in two PTX registers declared once
// at kernel scope (inline asm cannot carry a predicate between statements any other way).
#define DEQSCI_TOK_DECL() asm volatile(".reg .pred tokF, tokE;")
#define DEQSCI_TOK_TEST(tok, bar, par) \
  asm volatile("mbarrier.test_wait.parity.shared::cta.b64 " #tok ", [%0], %1;" ::"r"(bar), "r"(par) : "memory")
#define DEQSCI_TOK_GET(tok, out) asm volatile("selp.u32 %0, 1, 0, " #tok ";" : "=r"(out))

// STATS = true (train-mode BatchNorm): additionally accumulates per-output-channel sum and sum of
// squares of the values it writes (over valid pixels); every CTA writes its partial to p.stats[cta][128].
// MODE: issue order, see the constants above (0 output-stationary, 2 accumulator buffers of 192 columns; 1 / 2
// row-stationary, 4 buffers of 128 columns).
template <bool STATS, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_hidden_2cta_kernel(const __grid_constant__ CUtensorMap in_hi, const __grid_constant__ CUtensorMap in_lo,
                        const __grid_constant__ CUtensorMap out_hi, const __grid_constant__ CUtensorMap out_lo,
                        const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];          // swizzle atoms need the 1 KB alignment; no slack to
  if ((smem_u32(smem) & 1023u) != 0) __trap();               // realign by hand in mode 2 (227 KB exactly)
  DEQSCI_TOK_DECL();
  constexpr int kSlots = n_slots(MODE), kWBytes = w_bytes(MODE), kTapBytesB = tap_bytes(MODE);
  uint8_t* w_s = smem;
  uint8_t* a_s = w_s + kWBytes;
  uint8_t* st_s = a_s + kSlots * kSlotBytes;                 // 8 x 2 KB store staging
  uint8_t* tail = st_s + 8 * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);        // [0] w, full[4], empty[4], tfull[4], tempty[4]
  constexpr bool RS = MODE != 0;
  constexpr int kBufs = RS ? 4 : 2;
  constexpr int kCols = RS ? kAccColsRS : kAccCols;
  float* aff_s = reinterpret_cast<float*>(tail + 256);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 256 + 512);

  // programmatic dependent launch: the next layer's CTAs may take over each SM as soon as this grid's CTA
  // leaves it and run their prologue (barriers, TMEM, 72 KB of weights) under this grid's tail
  pdl_launch_dependents();
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_w = smem_u32(&bars[0]);
  auto bar_full = [&](int s) { return smem_u32(&bars[1 + s]); };
  auto bar_empty = [&](int s) { return smem_u32(&bars[1 + kSlotsMax + s]); };
  auto bar_tfull = [&](int b) { return smem_u32(&bars[1 + 2 * kSlotsMax + b]); };
  auto bar_tempty = [&](int b) { return smem_u32(&bars[1 + 2 * kSlotsMax + kAccBufsMax + b]); };

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < kSlots; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int b = 0; b < kBufs; ++b) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), 16); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) {
    const int c = threadIdx.x - 64;
    aff_s[2 * c] = p.scale ? p.scale[c] : 1.f;          // {scale, bias} pairs: one 16-byte load per two channels
    aff_s[2 * c + 1] = p.bias ? p.bias[c] : 0.f;
  }
  if (warp == 1) tmem_alloc2(smem_u32(tmem_slot), kTmemCols);
  if (warp == 0 && elect_one_sync()) {
    // this CTA's half of the weights; visible to the pair after the cluster barrier below
    mbar_arrive_expect_tx(bar_w, kWBytes);
    const uint8_t* src = p.wimg + (size_t)rank * kWBytes;
    for (int t = 0; t < 9; ++t) bulk_load_1d(smem_u32(w_s + t * kTapBytesB), src + (size_t)t * kTapBytesB, kTapBytesB, bar_w);
    mbar_wait(bar_w, 0);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                 // both CTAs: barriers initialised, weights resident, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs, own rows) =====================
    if (elect_one_sync()) {
      // the activations are the previous kernel's output: wait for that grid to complete (its memory is then
      // visible).  Everything this kernel writes depends on these loads, so no other thread needs the wait.
      pdl_wait_predecessor();
      int slot = 0;
      uint32_t phase = 0;
      for (long long item = pair; item < p.n_pair_items; item += n_pairs) {
        const Strip s = decode(p, 2 * item + rank);
        for (int q = 0; q < p.strip_rows + 2; ++q) {
          mbar_wait(bar_empty(slot), phase ^ 1);
          const uint32_t full_leader = mapa(bar_full(slot), 0);
          if (leader) mbar_arrive_expect_tx(bar_full(slot), 2 * kTxBytes);   // both CTAs' rows complete here
          const uint32_t dst = smem_u32(a_s + slot * kSlotBytes);
          tma_load_4d_2cta(dst, &in_hi, full_leader, 0, s.w0 - 1, s.h0 - 1 + q, s.nf);
          tma_load_4d_2cta(dst + kPlaneBytes, &in_lo, full_leader, 0, s.w0 - 1, s.h0 - 1 + q, s.nf);
          if (++slot == kSlots) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread, for the pair) =====================
    if (RS && leader && elect_one_sync()) {
      // row-stationary: walk the strip's R + 2 input rows; input row q feeds output rows q, q-1, q-2 (ky = 0, 1, 2)
      const uint32_t a_base = smem_u32(a_s), w_base = smem_u32(w_s);
      const int R = p.strip_rows;
      int slot = 0;
      uint32_t sphase = 0;
      uint32_t g0 = 0;               // running output-row count: row g lives in accumulator buffer g % 4
      bool tok_full = false, tok_tempty = false;     // a look-ahead test of this row's barrier is in tokF / tokE
      for (long long item = pair; item < p.n_pair_items; item += n_pairs) {
        const bool last_item = item + n_pairs >= p.n_pair_items;
        for (int q = 0; q < R + 2; ++q) {
          uint32_t ok;
          if (q < R) {               // a new output row opens: its buffer must have been drained
            const uint32_t g = g0 + q;
            ok = 0;
            if (tok_tempty) DEQSCI_TOK_GET(tokE, ok);
            if (!ok) mbar_wait(bar_tempty(g & 3), ((g >> 2) & 1) ^ 1);
          }
          ok = 0;
          if (tok_full) DEQSCI_TOK_GET(tokF, ok);
          if (!ok) mbar_wait(bar_full(slot), sphase);
          tc_fence_after();
          const uint32_t a_row = a_base + slot * kSlotBytes;
          const uint32_t d0 = tmem_base + ((g0 + q) & 3) * kAccColsRS;
          const uint32_t d1 = tmem_base + ((g0 + q - 1) & 3) * kAccColsRS;
          const uint32_t d2 = tmem_base + ((g0 + q - 2) & 3) * kAccColsRS;
          const int live = (q < R ? 1 : 0) | (q >= 1 && q - 1 < R ? 2 : 0) | (q >= 2 ? 4 : 0);
          constexpr bool H = MODE == 2;      // wide instructions
          issue_row_live<H, 0, 2>(live, a_row, w_base, d0, d1, d2);
          // look ahead: the next input row's barriers (same strip, or row 0 of this pair's next strip)
          tok_full = tok_tempty = false;
          if (p.lookahead && !(last_item && q == R + 1)) {
            const int nslot = slot + 1 == kSlots ? 0 : slot + 1;
            DEQSCI_TOK_TEST(tokF, bar_full(nslot), nslot == 0 ? sphase ^ 1 : sphase);
            tok_full = true;
            const int nq = q == R + 1 ? 0 : q + 1;
            if (nq < R) {
              const uint32_t ng = (q == R + 1 ? g0 + R : g0) + nq;
              DEQSCI_TOK_TEST(tokE, bar_tempty(ng & 3), ((ng >> 2) & 1) ^ 1);
              tok_tempty = true;
            }
          }
          issue_row_live<H, 2, 3>(live, a_row, w_base, d0, d1, d2);
          umma2_commit_mc(bar_empty(slot));                       // the input row is consumed (both CTAs)
          if (q >= 2) umma2_commit_mc(bar_tfull((g0 + q - 2) & 3));   // output row q-2 is complete
          if (++slot == kSlots) { slot = 0; sphase ^= 1; }
        }
        g0 += R;
      }
    }
    if (!RS && leader && elect_one_sync()) {
      constexpr uint32_t idesc_main = make_idesc(256, 128);
      constexpr uint32_t idesc_lo = make_idesc(256, 64);
      const uint32_t a_base = smem_u32(a_s), w_base = smem_u32(w_s);
      int first = 0;                 // slot of the strip's first row
      uint32_t first_phase = 0;
      int buf = 0;
      uint32_t tphase = 0;
      for (long long item = pair; item < p.n_pair_items; item += n_pairs) {
        int wait_slot = first;
        uint32_t wait_phase = first_phase;
        int rows_ready = 0;
        for (int j = 0; j < p.strip_rows; ++j) {
          mbar_wait(bar_tempty(buf), tphase ^ 1);
          while (rows_ready < j + 3) {               // output row j needs input rows j, j+1, j+2
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
              const uint64_t a_hi = make_sdesc(a_row + kx * 128);
              const uint64_t a_lo = make_sdesc(a_row + kPlaneBytes + kx * 128);
              const uint64_t b_w = make_sdesc(w_base + tap * kTapBytesB);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma2_f16(d_main, a_hi + 2 * k, b_w + 2 * k, idesc_main, (tap | k) != 0);
                umma2_f16(d_main + 128, a_lo + 2 * k, b_w + 2 * k, idesc_lo, (tap | k) != 0);
              }
            }
          }
          const int dead = (first + j) % kSlots;
          umma2_commit_mc(bar_empty(dead));          // input row j is dead after output row j (both CTAs)
          if (j == p.strip_rows - 1) {
            umma2_commit_mc(bar_empty((dead + 1) % kSlots));
            umma2_commit_mc(bar_empty((dead + 2) % kSlots));
          }
          umma2_commit_mc(bar_tfull(buf));
          if (++buf == kBufs) { buf = 0; tphase ^= 1; }
        }
        first = wait_slot;
        first_phase = wait_phase;
      }
    }
  } else {
    // ===================== epilogue (warps 2..9, both CTAs) =====================
    const int e = warp - 2;
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32)
    const int half = e >> 2;                      // output channels [32*half, +32)
    const uint32_t stage = smem_u32(st_s + e * kStageBytes);
    const uint32_t tempty_leader0 = mapa(bar_tempty(0), 0);      // the leader's tempty[b] = this + 8 b
    int buf = 0;
    uint32_t tphase = 0;
    float st_sum[STATS ? 32 : 1], st_sq[STATS ? 32 : 1];     // this thread's 32 channels, over all its pixels
    if (STATS) {
#pragma unroll
      for (int c = 0; c < 32; ++c) { st_sum[c] = 0.f; st_sq[c] = 0.f; }
    }
    for (long long item = pair; item < p.n_pair_items; item += n_pairs) {
      const Strip s = decode(p, 2 * item + rank);
      const bool col_valid = s.real && (s.w0 + quarter * 32 + lane) < p.Wc;
      for (int j = 0; j < p.strip_rows; ++j) {
        const int h = s.h0 + j;
        const bool px_valid = col_valid && h < p.Hc;
        // mask words of this thread's pixel and 32 channels, fetched before the accumulator wait
        uint32_t mkw[16];
        if (p.relu == 2) {
#pragma unroll
          for (int q = 0; q < 16; ++q) mkw[q] = 0u;
          if (px_valid) {
            const uint4* mp = reinterpret_cast<const uint4*>(
                p.mask + (((long long)s.nf * p.Hc + h) * p.Wc + (s.w0 + quarter * 32 + lane)) * 64 + half * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 t4 = __ldg(mp + q);
              mkw[4 * q] = t4.x; mkw[4 * q + 1] = t4.y; mkw[4 * q + 2] = t4.z; mkw[4 * q + 3] = t4.w;
            }
          }
        }
        mbar_wait(bar_tfull(buf), tphase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * kCols;
        uint32_t hi_pk[16], lo_pk[16];
#pragma unroll
        for (int part = 0; part < 2; ++part) {      // 16 channels at a time
          uint32_t acc[16], c1[16], c2[16];
          if (MODE == 2) {
            tmem_ld16(t_row + half * 96 + part * 16, acc);
            tmem_ld16(t_row + 32 + half * 32 + part * 16, c1);
          } else if (MODE == 1) {
            tmem_ld16(t_row + half * 32 + part * 16, acc);
            tmem_ld16(t_row + 64 + half * 32 + part * 16, c1);
          } else {
            tmem_ld16(t_row + half * 64 + part * 16, acc);
            tmem_ld16(t_row + half * 64 + 32 + part * 16, c1);
            tmem_ld16(t_row + 128 + half * 32 + part * 16, c2);
          }
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float v[2];
            const float4 sb4 = *reinterpret_cast<const float4*>(aff_s + 2 * (half * 32 + part * 16 + i));
            const float sb[4] = {sb4.x, sb4.y, sb4.z, sb4.w};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              float a = fmaf(RS ? __uint_as_float(c1[i + u]) : __uint_as_float(c1[i + u]) + __uint_as_float(c2[i + u]),
                             kLoInvScale, __uint_as_float(acc[i + u]));
              a = fmaf(a, sb[2 * u], sb[2 * u + 1]);
              if (p.relu == 2) {         // gate by the saved activation's sign (fp16 hi half != +0)
                const uint32_t mbits = (mkw[part * 8 + (i >> 1)] >> (16 * u)) & 0x7fffu;
                v[u] = mbits ? a : 0.f;
              } else {
                v[u] = p.relu ? fmaxf(a, 0.f) : a;
              }
              if (STATS && px_valid) {
                st_sum[part * 16 + i + u] += v[u];
                st_sq[part * 16 + i + u] = fmaf(v[u], v[u], st_sq[part * 16 + i + u]);
              }
            }
            split_f16x2(v[0], v[1], hi_pk[part * 8 + (i >> 1)], lo_pk[part * 8 + (i >> 1)]);
            lo_pk[part * 8 + (i >> 1)] = lo_pk[part * 8 + (i >> 1)] & p.lo_mask;
          }
        }
        // accumulator drained: hand the TMEM buffer back to the leader's MMA thread
        tc_fence_before();
        __syncwarp();
        if (elect_one_sync()) mbar_arrive_cluster(tempty_leader0 + 8 * buf);
        // stage (64-byte swizzle: chunk ^= (row >> 1) & 3) and store the two planes with TMA
        const uint32_t row_addr = stage + lane * 64;
        const int sw = (lane >> 1) & 3;
        if (p.debug_skip_store == 1) { if (++buf == kBufs) { buf = 0; tphase ^= 1; } continue; }
        if (p.debug_skip_store == 2) {             // experiment: direct global stores, no smem staging
          const int wpx = s.w0 + quarter * 32 + lane;
          if (s.real && wpx < p.Wc && h < p.Hc) {
            const long long off = (((long long)s.nf * p.Hc + h) * p.Wc + wpx) * 64 + half * 32;
            uint4* dh = reinterpret_cast<uint4*>(p.dbg_out_hi + off);
            uint4* dl = reinterpret_cast<uint4*>(p.dbg_out_lo + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              dh[q] = make_uint4(hi_pk[4 * q], hi_pk[4 * q + 1], hi_pk[4 * q + 2], hi_pk[4 * q + 3]);
              dl[q] = make_uint4(lo_pk[4 * q], lo_pk[4 * q + 1], lo_pk[4 * q + 2], lo_pk[4 * q + 3]);
            }
          }
          if (++buf == kBufs) { buf = 0; tphase ^= 1; }
          continue;
        }
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
          const uint32_t* pk = plane == 0 ? hi_pk : lo_pk;
          bulk_wait_read0();      // previous store has finished reading the staging buffer (all lanes execute it;
                                  // only the storing lane has a group pending -- no lane-0 guard, see conv_tc_first.cu)
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((q ^ sw) << 4)), "r"(pk[4 * q]),
                         "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                         : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (s.real && h < p.Hc) {
            if (elect_one_sync()) {
              tma_store_4d(plane == 0 ? &out_hi : &out_lo, stage, half * 32, s.w0 + quarter * 32, h, s.nf);
              bulk_commit();
            }
          }
        }
        if (++buf == kBufs) { buf = 0; tphase ^= 1; }
      }
    }
    bulk_wait0();
    if (STATS) {
      // warp totals of this warp's 32 channels -> its (now idle) staging buffer: [0,32) sums, [32,64) squares
      float* mine = reinterpret_cast<float*>(st_s + e * kStageBytes);
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float a = st_sum[c], b = st_sq[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0) {
          mine[c] = a;
          mine[32 + c] = b;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (STATS && threadIdx.x < 2 * 64) {
    // CTA partial in a fixed summation order (deterministic): p.stats[cta][0..63] sums, [64..127] sums of
    // squares; the finalize kernel adds the CTAs up in fp64.  No atomics: all CTAs finish together, and
    // ~600 same-address fp64 atomics per channel cost 20 us per launch.
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63, half = c >> 5;
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      acc += (double)reinterpret_cast<const float*>(st_s + (half * 4 + q) * kStageBytes)[which * 32 + (c & 31)];
    p.stats[(size_t)blockIdx.x * 128 + threadIdx.x] = acc;
  }
  cluster_sync();                 // the peer may still be reading this CTA's weights / barriers until here
  if (warp == 1) tmem_dealloc2(tmem_base, kTmemCols);
}


// ---- a RUN of consecutive hidden layers in ONE launch ---------------------------------------------------------
// EXPERIMENTAL, OFF BY DEFAULT, NOT PARITY-SAFE (DESIGN.md finding 19; DEQSCI_TC_CHAIN_EXPERIMENTAL=1).  Single calls
// match the per-layer kernel bit for bit, but now and then a TMA load issued right after the acquire of a neighbour's
// ready flag returns that row's PREVIOUS contents: the hand-off below is the documented pattern (store completion ->
// proxy fence -> release; acquire -> proxy fence -> TMA load) and still is not coherent on the B200 without a pause
// after the acquire.  Kept as the record of the experiment and of its measurements.
//
// The per-layer kernel above pays, per layer, the drain of its last tile, the exit, the next grid's prologue (barriers,
// TMEM, 108 KB of weights per CTA) and a cold input ring -- ~8 us, a third of a layer at batch 1 -- because a CTA fills
// its SM (227 KB, 512 TMEM columns): programmatic dependent launch cannot overlap two layers.  Here the CTA pairs stay
// resident for the whole run; barriers, TMEM, the rings and their phases carry over.  The grid-wide dependency between
// layers becomes a per-strip one: a strip of layer l + 1 needs rows h0 - 1 .. h0 + R of layer l's output, i.e. the same
// strip and its two neighbours in the frame.  Every strip has a ready flag in global memory; the epilogue publishes
// "layer l stored" (TMA stores complete -> named barrier of the epilogue warps -> st.release.gpu), the TMA producer
// acquires the three flags before it loads the strip's rows for layer l + 1.  That wait also covers the write-after-read
// hazard of the ping-pong planes (a neighbour has read my boundary rows of layer l - 1's output before it publishes
// layer l, and I overwrite them only after I have seen that flag).  Flags count up from a per-launch epoch, so they are
// never reset.  Deadlock-free: every wait is on a strictly earlier layer and all CTAs are co-resident (grid <= SMs).
// Weights of the next layer replace the current ones as soon as the layer's last MMA has retired (bar_wfree, committed
// by the issuer); the peer CTA reports its half through the leader's bar_wpeer (release / acquire at cluster scope).
// Issue order, accumulator layout and epilogue are those of MODE 2 above.
constexpr int kMaxChain = 16;
struct ChainParams {
  const uint8_t* wimg[kMaxChain];      // per layer: [2 ranks][9 taps][96 rows][128 B]
  const float* scale[kMaxChain];
  const float* bias[kMaxChain];
  int relu[kMaxChain];
  int n_layers;
  int NF, Hc, Wc;
  int tiles_x, strips_y, strip_rows;   // tiles_x == 1 (images of at most 128 pixels per row)
  long long n_strips, n_pair_items;
  uint32_t* flags;                     // [n_strips]; value epoch + l + 1 = layer l of this launch is in memory
  uint32_t epoch;
  int lookahead;
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t target) {   // bounded: a protocol bug traps
  uint32_t spins = 0;
  while ((int32_t)(ld_acquire_gpu(p) - target) < 0) {
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_hidden_chain_kernel(const __grid_constant__ CUtensorMap ld_hi0, const __grid_constant__ CUtensorMap ld_lo0,
                         const __grid_constant__ CUtensorMap ld_hi1, const __grid_constant__ CUtensorMap ld_lo1,
                         const __grid_constant__ CUtensorMap st_hi0, const __grid_constant__ CUtensorMap st_lo0,
                         const __grid_constant__ CUtensorMap st_hi1, const __grid_constant__ CUtensorMap st_lo1,
                         const __grid_constant__ ChainParams p) {
  constexpr int MODE = 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  DEQSCI_TOK_DECL();
  constexpr int kSlots = n_slots(MODE), kWBytes = w_bytes(MODE), kTapBytesB = tap_bytes(MODE);
  uint8_t* w_s = smem;
  uint8_t* a_s = w_s + kWBytes;
  uint8_t* st_s = a_s + kSlots * kSlotBytes;
  uint8_t* tail = st_s + 8 * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);        // [0] w, full[4], empty[4], tfull[4], tempty[4], wfree, wpeer
  float* aff_s = reinterpret_cast<float*>(tail + 256);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 256 + 512);

  pdl_launch_dependents();
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar_full = [&](int s) { return smem_u32(&bars[1 + s]); };
  auto bar_empty = [&](int s) { return smem_u32(&bars[1 + kSlotsMax + s]); };
  auto bar_tfull = [&](int b) { return smem_u32(&bars[1 + 2 * kSlotsMax + b]); };
  auto bar_tempty = [&](int b) { return smem_u32(&bars[1 + 2 * kSlotsMax + kAccBufsMax + b]); };
  // weights travel in three groups (the taps of one ky, 36 KB): group ky is last read by input row R - 1 + ky of a
  // layer's last strip and first read by input row ky of the next layer's first strip, so each group is replaced two
  // row times before it is needed and the reload never stalls the tensor pipe
  constexpr int kBarW = 1 + 2 * kSlotsMax + 2 * kAccBufsMax;
  auto bar_wg = [&](int g) { return smem_u32(&bars[kBarW + g]); };          // group g of the current layer has landed
  auto bar_wfree = [&](int g) { return smem_u32(&bars[kBarW + 3 + g]); };   // every MMA reading group g has retired
  auto bar_wpeer = [&](int g) { return smem_u32(&bars[kBarW + 6 + g]); };   // leader only: the peer's group g has landed
  constexpr int kGroupBytes = 3 * kTapBytesB;
  const int n_layers = p.n_layers;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int b = 0; b < kAccBufsMax; ++b) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), 16); }
    for (int g = 0; g < 3; ++g) { mbar_init(bar_wg(g), 1); mbar_init(bar_wfree(g), 1); mbar_init(bar_wpeer(g), 1); }
    fence_barrier_init();
    fence_proxy_async();
  }
  auto load_affine = [&](int l) {      // threads 64..127: {scale, bias} pairs of layer l
    const int c = threadIdx.x - 64;
    aff_s[2 * c] = p.scale[l] ? p.scale[l][c] : 1.f;
    aff_s[2 * c + 1] = p.bias[l] ? p.bias[l][c] : 0.f;
  };
  if (threadIdx.x >= 64 && threadIdx.x < 128) load_affine(0);
  if (warp == 1) tmem_alloc2(smem_u32(tmem_slot), kTmemCols);
  auto load_weight_group = [&](int l, int g) {     // one thread: this CTA's half of layer l's taps 3g .. 3g + 2
    mbar_arrive_expect_tx(bar_wg(g), kGroupBytes);
    const uint8_t* src = p.wimg[l] + (size_t)rank * kWBytes + (size_t)g * kGroupBytes;
    for (int t = 0; t < 3; ++t)
      bulk_load_1d(smem_u32(w_s + g * kGroupBytes + t * kTapBytesB), src + (size_t)t * kTapBytesB, kTapBytesB, bar_wg(g));
  };
  if (warp == 0 && elect_one_sync()) {
    for (int g = 0; g < 3; ++g) load_weight_group(0, g);
    for (int g = 0; g < 3; ++g) mbar_wait(bar_wg(g), 0);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                 // both CTAs: barriers initialised, first weights resident, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int R = p.strip_rows;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs, own rows) =====================
    if (elect_one_sync()) {
      pdl_wait_predecessor();
      int slot = 0;
      uint32_t phase = 0;
      for (int l = 0; l < n_layers; ++l) {
        const CUtensorMap* m_hi = (l & 1) ? &ld_hi1 : &ld_hi0;
        const CUtensorMap* m_lo = (l & 1) ? &ld_lo1 : &ld_lo0;
        for (long long item = pair; item < p.n_pair_items; item += n_pairs) {
          long long sid = 2 * item + rank;
          if (sid >= p.n_strips) sid = p.n_strips - 1;
          const Strip s = decode(p, sid);
          // Weight group ky of the new layer goes in as soon as layer l - 1 lets go of it (bar_wfree): group 0 before
          // anything that can block, groups 1 and 2 between the first rows -- or, when this CTA has a single strip per
          // layer, all three before the flags (its layer is finished, the flags are what it waits for).
          const bool first_item = item == pair, single = pair + n_pairs >= p.n_pair_items;
          int groups_loaded = 3;
          if (l > 0 && first_item) {
            groups_loaded = single ? 3 : 1;
            for (int g = 0; g < groups_loaded; ++g) {
              mbar_wait(bar_wfree(g), (l - 1) & 1);
              load_weight_group(l, g);
            }
          }
          if (l > 0) {                             // layer l - 1 of this strip and of its neighbours in the frame
            const uint32_t target = p.epoch + (uint32_t)l;
            const int sy = (int)(sid % p.strips_y);
            wait_flag(p.flags + sid, target);
            if (sy > 0) wait_flag(p.flags + sid - 1, target);
            if (sy < p.strips_y - 1) wait_flag(p.flags + sid + 1, target);
            fence_proxy_async_all();               // the acquired data is read through the async proxy (TMA)
          }
          for (int q = 0; q < R + 2; ++q) {
            if (q >= 1 && q < 3 && groups_loaded <= q) {
              mbar_wait(bar_wfree(q), (l - 1) & 1);
              load_weight_group(l, q);
            }
            mbar_wait(bar_empty(slot), phase ^ 1);
            const uint32_t full_leader = mapa(bar_full(slot), 0);
            if (leader) mbar_arrive_expect_tx(bar_full(slot), 2 * kTxBytes);
            const uint32_t dst = smem_u32(a_s + slot * kSlotBytes);
            tma_load_4d_2cta(dst, m_hi, full_leader, 0, s.w0 - 1, s.h0 - 1 + q, s.nf);
            tma_load_4d_2cta(dst + kPlaneBytes, m_lo, full_leader, 0, s.w0 - 1, s.h0 - 1 + q, s.nf);
            if (++slot == kSlots) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (!leader) {
      // peer CTA: tell the leader's issuer when this CTA's half of a layer's weights has landed
      if (elect_one_sync()) {
        const uint32_t wpeer_leader = mapa(bar_wpeer(0), 0);
        for (int l = 1; l < n_layers; ++l)
          for (int g = 0; g < 3; ++g) {
            mbar_wait(bar_wg(g), l & 1);
            mbar_arrive_release_cluster(wpeer_leader + 8 * g);
          }
      }
    } else if (elect_one_sync()) {
      // ===================== MMA issuer (leader CTA, one thread, for the pair) =====================
      const uint32_t a_base = smem_u32(a_s), w_base = smem_u32(w_s);
      int slot = 0;
      uint32_t sphase = 0;
      uint32_t g0 = 0;
      bool tok_full = false, tok_tempty = false;     // a look-ahead test of this row's barrier is in tokF / tokE
      for (int l = 0; l < n_layers; ++l) {
        for (long long item = pair; item < p.n_pair_items; item += n_pairs) {
          const bool last_item = item + n_pairs >= p.n_pair_items;
          for (int q = 0; q < R + 2; ++q) {
            if (l > 0 && item == pair && q < 3) {     // input row q is the first to read weight group q of this layer
              mbar_wait(bar_wg(q), l & 1);
              mbar_wait_acquire_cluster(bar_wpeer(q), (l - 1) & 1);
            }
            uint32_t ok;
            if (q < R) {
              const uint32_t g = g0 + q;
              ok = 0;
              if (tok_tempty) DEQSCI_TOK_GET(tokE, ok);
              if (!ok) mbar_wait(bar_tempty(g & 3), ((g >> 2) & 1) ^ 1);
            }
            ok = 0;
            if (tok_full) DEQSCI_TOK_GET(tokF, ok);
            if (!ok) mbar_wait(bar_full(slot), sphase);
            tc_fence_after();
            const uint32_t a_row = a_base + slot * kSlotBytes;
            const uint32_t d0 = tmem_base + ((g0 + q) & 3) * kAccColsRS;
            const uint32_t d1 = tmem_base + ((g0 + q - 1) & 3) * kAccColsRS;
            const uint32_t d2 = tmem_base + ((g0 + q - 2) & 3) * kAccColsRS;
            const int live = (q < R ? 1 : 0) | (q >= 1 && q - 1 < R ? 2 : 0) | (q >= 2 ? 4 : 0);
            issue_row_live<true, 0, 2>(live, a_row, w_base, d0, d1, d2);
            // look ahead: the next input row's barriers (same strip, this pair's next strip, or the next layer's first)
            tok_full = tok_tempty = false;
            if (p.lookahead && !(last_item && q == R + 1 && l + 1 == n_layers)) {
              const int nslot = slot + 1 == kSlots ? 0 : slot + 1;
              DEQSCI_TOK_TEST(tokF, bar_full(nslot), nslot == 0 ? sphase ^ 1 : sphase);
              tok_full = true;
              const int nq = q == R + 1 ? 0 : q + 1;
              if (nq < R) {
                const uint32_t ng = (q == R + 1 ? g0 + R : g0) + nq;
                DEQSCI_TOK_TEST(tokE, bar_tempty(ng & 3), ((ng >> 2) & 1) ^ 1);
                tok_tempty = true;
              }
            }
            issue_row_live<true, 2, 3>(live, a_row, w_base, d0, d1, d2);
            umma2_commit_mc(bar_empty(slot));
            if (q >= 2) umma2_commit_mc(bar_tfull((g0 + q - 2) & 3));
            // both CTAs: input row R - 1 + ky of the layer's last strip was the last reader of weight group ky
            if (last_item && l + 1 < n_layers && q >= R - 1) umma2_commit_mc(bar_wfree(q - (R - 1)));
            if (++slot == kSlots) { slot = 0; sphase ^= 1; }
          }
          g0 += R;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9, both CTAs) =====================
    const int e = warp - 2;
    const int quarter = warp & 3;
    const int half = e >> 2;
    const uint32_t stage = smem_u32(st_s + e * kStageBytes);
    const uint32_t tempty_leader0 = mapa(bar_tempty(0), 0);
    int buf = 0;
    uint32_t tphase = 0;
    // A strip's rows of layer l are in memory once every epilogue warp's TMA stores have completed (wait_group 0 by the
    // storing lanes), hence the barrier of the 8 warps; then one thread publishes.  sid < 0: nothing to publish, but the
    // barrier is still taken (the next layer's affine is loaded behind the layer's last one).
    auto publish_strip = [&](long long sid, int l) {
      bulk_wait0();
      fence_proxy_async_all();
      __threadfence();
      __syncwarp();
      named_bar_sync(1, 256);
      if (e == 0 && sid >= 0 && elect_one_sync()) {
        __threadfence();
        st_release_gpu(p.flags + sid, p.epoch + (uint32_t)l + 1u);
      }
    };
    for (int l = 0; l < n_layers; ++l) {
      const int relu = p.relu[l];
      const CUtensorMap* o_hi = ((l + 1) & 1) ? &st_hi1 : &st_hi0;
      const CUtensorMap* o_lo = ((l + 1) & 1) ? &st_lo1 : &st_lo0;
      const bool publish = l + 1 < n_layers;
      for (long long item = pair; item < p.n_pair_items; item += n_pairs) {
        const long long sid = 2 * item + rank;
        const Strip s = decode(p, sid);
        for (int j = 0; j < R; ++j) {
          const int h = s.h0 + j;
          mbar_wait(bar_tfull(buf), tphase);
          tc_fence_after();
          const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * kAccColsRS;
          uint32_t hi_pk[16], lo_pk[16];
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            uint32_t acc[16], c1[16];
            tmem_ld16(t_row + half * 96 + part * 16, acc);
            tmem_ld16(t_row + 32 + half * 32 + part * 16, c1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              float v[2];
              const float4 sb4 = *reinterpret_cast<const float4*>(aff_s + 2 * (half * 32 + part * 16 + i));
              const float sb[4] = {sb4.x, sb4.y, sb4.z, sb4.w};
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                float a = fmaf(__uint_as_float(c1[i + u]), kLoInvScale, __uint_as_float(acc[i + u]));
                a = fmaf(a, sb[2 * u], sb[2 * u + 1]);
                v[u] = relu ? fmaxf(a, 0.f) : a;
              }
              split_f16x2(v[0], v[1], hi_pk[part * 8 + (i >> 1)], lo_pk[part * 8 + (i >> 1)]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (elect_one_sync()) mbar_arrive_cluster(tempty_leader0 + 8 * buf);
          const uint32_t row_addr = stage + lane * 64;
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int plane = 0; plane < 2; ++plane) {
            const uint32_t* pk = plane == 0 ? hi_pk : lo_pk;
            bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((q ^ sw) << 4)), "r"(pk[4 * q]),
                           "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                           : "memory");
            }
            fence_proxy_async();
            __syncwarp();
            if (s.real && h < p.Hc) {
              if (elect_one_sync()) {
                tma_store_4d(plane == 0 ? o_hi : o_lo, stage, half * 32, s.w0 + quarter * 32, h, s.nf);
                bulk_commit();
              }
            }
          }
          if (++buf == kAccBufsMax) { buf = 0; tphase ^= 1; }
        }
        // The wait for the stores costs the epilogue warps nothing: the next strip's first accumulator fills only after
        // three more input rows.  (Publishing one strip late instead -- from the next strip's first tile -- starved the
        // producer: it wants the flags as soon as it has queued the layer's last row.)
        if (publish) publish_strip(s.real ? sid : -1, l);
      }
      if (publish) {
        // every epilogue warp is past its last use of layer l's affine (the barrier above): load the next layer's
        if (threadIdx.x >= 64 && threadIdx.x < 128) load_affine(l + 1);
        named_bar_sync(1, 256);
      }
    }
    bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) tmem_dealloc2(tmem_base, kTmemCols);
}

}  // namespace tc2

// ---- host side -------------------------------------------------------------------------------
// issue mode of the pair kernel, fixed for the process (it also decides the weight image layout)
static int tc2_issue_mode() {
  static const int mode = [] { const int m = env_int("DEQSCI_TC_RS", 2); return m < 0 || m > 2 ? 2 : m; }();
  return mode;
}
size_t tc2_weight_image_bytes() { return 2 * (size_t)tc2::w_bytes(tc2_issue_mode()); }

// w [64 cout][64 cin][3][3] fp32 -> [rank][tap][rows][128 B], K-major fp16 rows with the 128-byte swizzle.
// Modes 0, 1 (64 rows): rows of rank r: [0,32) = hi(W[32r + n]), [32,64) = lo'(W[32r + n - 32]).
// Mode 2 (96 rows): rank 0: [0,32) = [32,64) = hi(W[n % 32]), [64,96) = lo'(W[n - 64]);
//                   rank 1: [0,32) = [64,96) = hi(W[32 + n % 32]), [32,64) = lo'(W[n]).
//   The N = 128 instruction reads rows [32,96) of both CTAs, the N = 64 one rows [0,32): its 64 columns land on the
//   64 lo' columns in the middle of the wide instruction's 128.
template <class Emit>
static void tc2_layout(Emit emit) {
  const int mode = tc2_issue_mode(), rows = tc2::tap_rows(mode), tapb = tc2::tap_bytes(mode);
  for (int rank = 0; rank < 2; ++rank)
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
      for (int n = 0; n < rows; ++n) {
        const int co = 32 * rank + (n & 31);
        const bool lo = mode == 2 ? (rank == 0 ? n >= 64 : (n >= 32 && n < 64)) : n >= 32;
        for (int k = 0; k < 64; ++k) {
          const size_t byte = ((size_t)rank * 9 + tap) * tapb + (size_t)n * 128 +
                              (size_t)(((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
          emit(byte, ((co * 64 + k) * 3 + ky) * 3 + kx, lo);
        }
      }
    }
}
void tc2_pack_weights(const float* w, uint8_t* img) {
  tc2_layout(PackWrite{w, img});
  const uint16_t mask = (uint16_t)env_int("DEQSCI_TC_LO_MASK", 0xFFFF);      // experiment switch, see Params::lo_mask
  if (mask != 0xFFFF)
    tc2_layout([&](size_t byte, int, bool lo) { if (lo) *reinterpret_cast<uint16_t*>(img + byte) &= mask; });
}
void tc2_pack_map(int32_t* map) { tc2_layout(PackMap{map}); }

// true when the pair kernel can run this shape: images wider than half a 128-pixel row tile (any height)
bool tc2_supported(int Hc, int Wc) {
  static const int enabled = env_int("DEQSCI_TC_PAIR", 1);
  return enabled && Wc > 64 && Hc >= 1;
}

int conv_hidden_2cta_launch(const __half* act_in, __half* act_out, long long plane_elems, const uint8_t* wimg,
                            const float* scale, const float* bias, int relu, int NF, int Hc, int Wc,
                            cudaStream_t st, double* stats, const __half* mask) {
  tc2::Params p;
  p.wimg = wimg; p.scale = scale; p.bias = bias; p.relu = relu;
  p.mask = mask;
  if (relu == 2 && !mask) { set_error("conv_hidden_2cta_launch: masked layer without a mask plane"); return DEQSCI_ERR_INVALID; }
  p.NF = NF; p.Hc = Hc; p.Wc = Wc;
  p.tiles_x = (Wc + tc2::kTileM - 1) / tc2::kTileM;
  const int pairs_hw = num_sms() / 2;
  // strip height from the cost model in tma_host.cu; DEQSCI_TC_ROUNDS=n > 0 selects the older rule instead
  // (largest power of two <= 16 dividing Hc that leaves >= n strips per SM)
  static const int rounds = env_int("DEQSCI_TC_ROUNDS", 0);
  const int R = rounds > 0 ? pick_strip_rows(NF, p.tiles_x, Hc, true, 2LL * rounds * pairs_hw, 1)
                           : pick_strip_rows_balanced(NF, p.tiles_x, Hc, false, pairs_hw, 2, 1, 1);
  p.strip_rows = R;
  p.strips_y = (Hc + R - 1) / R;       // the last strip of a frame may run past Hc: those rows load zeros (TMA
                                       // out-of-bounds fill), are computed in lockstep with the pair and never stored
  p.n_strips = (long long)NF * p.tiles_x * p.strips_y;
  p.n_pair_items = (p.n_strips + 1) / 2;
  static const int skip_store = env_int("DEQSCI_TC_DEBUG_SKIP_STORE", 0);
  p.debug_skip_store = skip_store;
  static const uint32_t lo_mask16 = (uint32_t)env_int("DEQSCI_TC_LO_MASK", 0xFFFF) & 0xFFFFu;
  p.lo_mask = lo_mask16 | (lo_mask16 << 16);
  p.stats = stats;
  static const int lookahead = env_int("DEQSCI_TC_LOOKAHEAD", 1);
  p.lookahead = lookahead;
  p.dbg_out_hi = act_out;
  p.dbg_out_lo = act_out + plane_elems;
  CUtensorMap in_hi, in_lo, out_hi, out_lo;
  int rc;
  if ((rc = make_plane_map(&in_hi, act_in, 64, NF, Hc, Wc, 64, tc2::kTileM + 2, 1, 128))) return rc;
  if ((rc = make_plane_map(&in_lo, act_in + plane_elems, 64, NF, Hc, Wc, 64, tc2::kTileM + 2, 1, 128))) return rc;
  if ((rc = make_plane_map(&out_hi, act_out, 64, NF, Hc, Wc, 32, 32, 1, 64))) return rc;
  if ((rc = make_plane_map(&out_lo, act_out + plane_elems, 64, NF, Hc, Wc, 32, 32, 1, 64))) return rc;
  const long long pairs = p.n_pair_items < pairs_hw ? p.n_pair_items : pairs_hw;
  const int issue_mode = tc2_issue_mode();
  ProfScope prof(PK_CONV_HIDDEN, st);
  auto launch = [&](auto kernel) {
    const int smem = tc2::smem_bytes(issue_mode);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    return launch_pdl(kernel, (unsigned)(2 * pairs), tc2::kThreads, smem, st, in_hi, in_lo, out_hi, out_lo, p);
  };
  if (stats) {
    if (issue_mode == 0) DEQSCI_CUDA(launch(tc2::conv_hidden_2cta_kernel<true, 0>));
    else if (issue_mode == 1) DEQSCI_CUDA(launch(tc2::conv_hidden_2cta_kernel<true, 1>));
    else DEQSCI_CUDA(launch(tc2::conv_hidden_2cta_kernel<true, 2>));
  } else {
    if (issue_mode == 0) DEQSCI_CUDA(launch(tc2::conv_hidden_2cta_kernel<false, 0>));
    else if (issue_mode == 1) DEQSCI_CUDA(launch(tc2::conv_hidden_2cta_kernel<false, 1>));
    else DEQSCI_CUDA(launch(tc2::conv_hidden_2cta_kernel<false, 2>));
  }
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}


// ---- chained launch: layers [0, n_layers) of a run, ping-pong between two plane pairs -----------------------------
// True when the run kernel can take this shape: wide issue mode (its weight image layout), one 128-pixel tile per row.
bool tc2_chain_supported(int NF, int Hc, int Wc, int n_layers) {
  // OFF by default -- EXPERIMENTAL, NOT PARITY-SAFE (DESIGN.md finding 19): on the B200 a TMA load issued right after
  // the acquire of a strip's ready flag can still return the row's PREVIOUS contents (the data were stored, completed
  // and fenced >= 10 us earlier; a 10 us pause after the acquire hides it, fences at any scope do not close it), so a
  // solve drifts from the per-layer kernels by up to 1e-2.  Kept as the record of that experiment.
  static const int enabled = env_int("DEQSCI_TC_CHAIN_EXPERIMENTAL", 0);
  // only while a CTA pair has at most this many 16-row strip pairs per layer (batches of 1-2 measurements of
  // 256 x 256 x 8): measured on one box, the run kernel is 6 % faster at batch 1, 1.6 % at batch 2, equal at 4 and
  // 2-3 % SLOWER from batch 8 on, where a kernel boundary is under 2 % of a layer
  static const int max_rounds = env_int("DEQSCI_TC_CHAIN_MAX_ROUNDS", 1);
  const long long rounds16 = ((long long)NF * ((Hc + 15) / 16) / 2 + num_sms() / 2 - 1) / (num_sms() / 2);
  return enabled && rounds16 <= max_rounds && tc2_issue_mode() == 2 && tc2_supported(Hc, Wc) && Wc <= tc2::kTileM &&
         n_layers >= 2 && n_layers <= tc2::kMaxChain;
}
size_t tc2_chain_flag_count(int NF, int Hc) { return (size_t)NF * Hc; }       // upper bound: one strip per image row

// Layer l reads plane pair buf[l & 1] and writes buf[(l + 1) & 1]; the result of the run is in buf[n_layers & 1].
// flags: device uint32[>= tc2_chain_flag_count], zero-initialised once; *epoch: host counter owned by the caller with
// the flags (one pair per stream), advanced here; values never need resetting until the counter wraps (handled).
int conv_hidden_chain_launch(__half* buf0, __half* buf1, long long plane_elems, int n_layers, const uint8_t* const* wimg,
                             const float* const* scale, const float* const* bias, const int* relu, int NF, int Hc, int Wc,
                             uint32_t* flags, uint32_t* epoch, cudaStream_t st) {
  tc2::ChainParams p;
  memset(&p, 0, sizeof(p));
  for (int l = 0; l < n_layers; ++l) { p.wimg[l] = wimg[l]; p.scale[l] = scale[l]; p.bias[l] = bias[l]; p.relu[l] = relu[l]; }
  p.n_layers = n_layers;
  p.NF = NF; p.Hc = Hc; p.Wc = Wc;
  p.tiles_x = 1;
  const int pairs_hw = num_sms() / 2;
  // Strip height: rounds x (R + 1/2 row of per-strip cost), as for the per-layer kernel -- but a CTA that has ONE strip
  // per layer sits through the whole store -> flag -> load round trip between two layers (~4 row times), while with two
  // or more its next strip's inputs were published a strip ago: prefer at least two rounds (batch 1: 4-row strips, two
  // per CTA, instead of one of 8).  DEQSCI_TC_CHAIN_ROWS=n forces a height.
  static const int forced_rows = env_int("DEQSCI_TC_CHAIN_ROWS", 0);
  int R = 1;
  long long best_cost = -1;
  for (int r = 16; r >= 1; --r) {
    const long long strips = (long long)NF * ((Hc + r - 1) / r);
    const long long rounds = ((strips + 1) / 2 + pairs_hw - 1) / pairs_hw;
    const long long cost = rounds * (2 * r + 1) + (rounds == 1 ? 8 : 0);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; R = r; }
  }
  if (forced_rows > 0 && forced_rows <= 16) R = forced_rows;
  p.strip_rows = R;
  p.strips_y = (Hc + R - 1) / R;
  p.n_strips = (long long)NF * p.strips_y;
  p.n_pair_items = (p.n_strips + 1) / 2;
  if (*epoch > (1u << 30)) {            // wrap: start over from zeroed flags (stream-ordered behind earlier launches)
    DEQSCI_CUDA(cudaMemsetAsync(flags, 0, tc2_chain_flag_count(NF, Hc) * sizeof(uint32_t), st));
    *epoch = 0;
  }
  p.flags = flags;
  p.epoch = *epoch;
  static const int lookahead = env_int("DEQSCI_TC_LOOKAHEAD", 1);
  p.lookahead = lookahead;
  CUtensorMap m[8];
  __half* bufs[2] = {buf0, buf1};
  int rc;
  for (int b = 0; b < 2; ++b) {
    if ((rc = make_plane_map(&m[2 * b], bufs[b], 64, NF, Hc, Wc, 64, tc2::kTileM + 2, 1, 128))) return rc;
    if ((rc = make_plane_map(&m[2 * b + 1], bufs[b] + plane_elems, 64, NF, Hc, Wc, 64, tc2::kTileM + 2, 1, 128))) return rc;
    if ((rc = make_plane_map(&m[4 + 2 * b], bufs[b], 64, NF, Hc, Wc, 32, 32, 1, 64))) return rc;
    if ((rc = make_plane_map(&m[4 + 2 * b + 1], bufs[b] + plane_elems, 64, NF, Hc, Wc, 32, 32, 1, 64))) return rc;
  }
  const long long pairs = p.n_pair_items < pairs_hw ? p.n_pair_items : pairs_hw;
  const int smem = tc2::smem_bytes(2);
  ProfScope prof(PK_CONV_HIDDEN, st);
  DEQSCI_CUDA(cudaFuncSetAttribute(tc2::conv_hidden_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  DEQSCI_CUDA(launch_pdl(tc2::conv_hidden_chain_kernel, (unsigned)(2 * pairs), tc2::kThreads, smem, st, m[0], m[1], m[2], m[3],
                         m[4], m[5], m[6], m[7], p));
  DEQSCI_LAUNCH_CHECK();
  return DEQSCI_OK;
}

}  // namespace deqsci
