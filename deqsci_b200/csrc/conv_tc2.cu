// Kernel group 2c: hidden 64->64 3x3 conv layer, split-fp16 precision, as a CTA-PAIR kernel
// (thread-block cluster of 2, tcgen05 cta_group::2, UMMA_M = 256).
//
// Why a pair: the split-precision weights [Wh | Wl'] of one layer are 144 KB.  Resident in ONE CTA
// they leave room for only two activation-row slots and no store staging, so the single-CTA kernel
// (conv_tc.cu, LD_ROW3) is starved by TMA latency and its epilogue pays 32 partial-line stores per
// instruction.  cta_group::2 lets each CTA of the pair keep HALF of the B operand (72 KB): that
// frees shared memory for a 4-slot rolling row ring (every input row is loaded once per strip and
// used by the three output rows around it) plus staging buffers for coalesced TMA stores, and halves
// the B-operand shared-memory reads per SM.
//
// Each CTA of the pair owns its own strip of R consecutive output rows (128-pixel row segments) of
// some frame; the two strips only share the weights.  Per output row and CTA:
//     MMA 1: A = Ah (own 128 pixels), B = 128 rows (64 from each CTA), N = 128
//     MMA 2: A = Al',                 B =  64 rows (32 from each CTA), N =  64
// B rows are interleaved so both instructions find their half at the SAME smem address in each CTA:
//     leader tile rows [0,32) = Wh[0:32], [32,64) = Wl'[0:32];  peer: Wh[32:64], Wl'[32:64]
//   => MMA 1 columns: [0,32) main 0-31 | [32,64) corr1 0-31 | [64,96) main 32-63 | [96,128) corr1 32-63
//      MMA 2 columns: [128,192) corr2 0-63           out = main + 2^-11 (corr1 + corr2)
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
  const __half* mask;         // relu == 2: output (pixel, channel) is kept where this hi plane [NF,Hc,Wc,64] is > 0, else
                              // zeroed: the ReLU derivative of a saved forward activation (adjoint / VJP stacks)
};

struct Strip { int nf, h0, w0; bool real; };

__device__ __forceinline__ Strip decode(const Params& p, long long strip) {
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
template <bool L0, bool L1, bool L2, bool WIDE>
__device__ __forceinline__ void issue_row_rs(uint32_t a_row, uint32_t w_base, uint32_t d0, uint32_t d1, uint32_t d2) {
  constexpr uint32_t idesc64 = make_idesc(256, 64), idesc128 = make_idesc(256, 128);
  constexpr int n_live = (L0 ? 1 : 0) + (L1 ? 1 : 0) + (L2 ? 1 : 0);
  constexpr int p1 = L0 ? 1 : 0, p2 = p1 + (L1 ? 1 : 0);       // position of ky = 1, 2 among the live rows
  constexpr uint64_t kRows32 = (32 * 128) >> 4;                // descriptor step of 32 tile rows
  constexpr int kTapB = tap_bytes(WIDE ? 2 : 1);
#pragma unroll
  for (int kx = 0; kx < 3; ++kx) {
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
      for (long long item = pair; item < p.n_pair_items; item += n_pairs) {
        for (int q = 0; q < R + 2; ++q) {
          if (q < R) {               // a new output row opens: its buffer must have been drained
            const uint32_t g = g0 + q;
            mbar_wait(bar_tempty(g & 3), ((g >> 2) & 1) ^ 1);
          }
          mbar_wait(bar_full(slot), sphase);
          tc_fence_after();
          const uint32_t a_row = a_base + slot * kSlotBytes;
          const uint32_t d0 = tmem_base + ((g0 + q) & 3) * kAccColsRS;
          const uint32_t d1 = tmem_base + ((g0 + q - 1) & 3) * kAccColsRS;
          const uint32_t d2 = tmem_base + ((g0 + q - 2) & 3) * kAccColsRS;
          const int live = (q < R ? 1 : 0) | (q >= 1 && q - 1 < R ? 2 : 0) | (q >= 2 ? 4 : 0);
          constexpr bool H = MODE == 2;      // wide instructions
          switch (live) {
            case 1: issue_row_rs<true, false, false, H>(a_row, w_base, d0, d1, d2); break;
            case 2: issue_row_rs<false, true, false, H>(a_row, w_base, d0, d1, d2); break;
            case 3: issue_row_rs<true, true, false, H>(a_row, w_base, d0, d1, d2); break;
            case 4: issue_row_rs<false, false, true, H>(a_row, w_base, d0, d1, d2); break;
            case 6: issue_row_rs<false, true, true, H>(a_row, w_base, d0, d1, d2); break;
            default: issue_row_rs<true, true, true, H>(a_row, w_base, d0, d1, d2); break;
          }
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

}  // namespace deqsci
