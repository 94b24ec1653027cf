// Host-side helpers shared by the tensor-core conv launchers: TMA tensor-map encoding for the
// channels-last fp16 activation planes, and the strip-height heuristic of the rolling-row kernels.
#include <cuda.h>   // CUtensorMap types only; cuTensorMapEncodeTiled is fetched at run time (no -lcuda)

#include "common.cuh"

namespace deqsci {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<PFN_encodeTiled>(ptr);
  }();
  return fn;
}

// 4-D map over one plane [NF][Hc][Wc][channels] fp16 with box {box_c, box_w, box_h, 1}.  Out-of-bounds
// box elements read as zero (= the conv's zero padding) and are clipped on store.  swizzle_bytes must
// equal box_c * 2 (32, 64 or 128): one box row = one swizzle span.
int make_plane_map(CUtensorMap* map, const __half* plane, int channels, int NF, int Hc, int Wc, int box_c, int box_w,
                   int box_h, int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DEQSCI_ERR_CUDA; }
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                                                     : CU_TENSOR_MAP_SWIZZLE_32B;
  const cuuint64_t rb = (cuuint64_t)channels * sizeof(__half);
  cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)Wc, (cuuint64_t)Hc, (cuuint64_t)NF};
  cuuint64_t strides[3] = {rb, (cuuint64_t)Wc * rb, (cuuint64_t)Hc * Wc * rb};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)plane, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: CUresult %d", (int)r); return DEQSCI_ERR_CUDA; }
  return DEQSCI_OK;
}

// Output rows per work item of a rolling-row kernel: the largest power of two <= 16 (dividing Hc when
// must_divide) that still leaves >= min_items work items (load balance over the persistent CTAs);
// longer strips mean fewer halo-row reloads ((R + 2) / R loads per tile).
int pick_strip_rows(int NF, int tiles_x, int Hc, bool must_divide, long long min_items, int floor_rows) {
  int R = 16;
  while (R > floor_rows && ((must_divide && Hc % R != 0) || (long long)NF * tiles_x * ((Hc + R - 1) / R) < min_items))
    R /= 2;
  return R;
}

// Strip height by a cost model instead of a floor on the item count: a persistent worker runs
// ceil(items / workers) strips back to back and each strip costs its R rows plus about half a row of
// pipeline refill (3 start rows cannot be prefetched behind the previous strip), so minimise
// rounds * (R + overhead); ties go to the taller strip (fewer halo reloads).  strips_per_item = strips one
// work item covers (2 for the CTA-pair kernel); overhead_half_rows = per-strip overhead in half rows (1 for
// the refill above; 4 for the ky-transposed last layer, which also pushes the 2 halo rows through the MMA).
// Fitted on batch 1..8 measurements (DESIGN.md finding 11).
int pick_strip_rows_balanced(int NF, int tiles_x, int Hc, bool must_divide, int workers, int strips_per_item,
                             int overhead_half_rows, int min_rows) {
  int best = min_rows;
  long long best_cost = -1;
  for (int R = 16; R >= min_rows; --R) {               // every height, not only powers of two
    if (must_divide && Hc % R != 0) continue;
    const long long strips = (long long)NF * tiles_x * ((Hc + R - 1) / R);
    const long long items = (strips + strips_per_item - 1) / strips_per_item;
    const long long rounds = (items + workers - 1) / workers;
    const long long cost = rounds * (2 * R + overhead_half_rows);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = R; }
  }
  return best;
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

}  // namespace deqsci
