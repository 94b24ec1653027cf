// Launch accounting and sampled CUDA-event timing of libdeqsci's own kernels (bench.py's
// `gpu_launches` and `roofline.achieved`).  Off by default: one relaxed atomic add per launch.
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace deqsci {

static std::atomic<long long> g_launches[PK_COUNT];
static std::atomic<int> g_sample_every{0};
static std::mutex g_mu;
struct Sample { cudaEvent_t a, b; int kind; };
static std::vector<Sample> g_samples;
static std::vector<cudaEvent_t> g_pool;
static std::atomic<long long> g_tick{0};

ProfScope::ProfScope(int kind, cudaStream_t st) : kind_(kind), st_(st), idx_(-1) {
  g_launches[kind].fetch_add(1, std::memory_order_relaxed);
  const int every = g_sample_every.load(std::memory_order_relaxed);
  if (every <= 0) return;
  if (g_tick.fetch_add(1, std::memory_order_relaxed) % every != 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  Sample s;
  s.kind = kind;
  auto get = [&]() {
    cudaEvent_t e;
    if (!g_pool.empty()) { e = g_pool.back(); g_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
  };
  s.a = get();
  s.b = get();
  cudaEventRecord(s.a, st_);
  g_samples.push_back(s);
  idx_ = (long long)g_samples.size() - 1;
}

ProfScope::~ProfScope() {
  if (idx_ < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEventRecord(g_samples[idx_].b, st_);
}

}  // namespace deqsci

using namespace deqsci;

extern "C" int deqsci_profile_begin(int sample_every) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& s : g_samples) { g_pool.push_back(s.a); g_pool.push_back(s.b); }
  g_samples.clear();
  for (int k = 0; k < PK_COUNT; ++k) g_launches[k].store(0);
  g_tick.store(0);
  g_sample_every.store(sample_every);
  return DEQSCI_OK;
}

extern "C" int deqsci_profile_end(double* ms_sum, long long* n_sampled, long long* n_launched) {
  DEQSCI_CHECK_ARG(ms_sum && n_sampled && n_launched, "profile_end: null pointer");
  g_sample_every.store(0);
  DEQSCI_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_mu);
  for (int k = 0; k < PK_COUNT; ++k) { ms_sum[k] = 0.0; n_sampled[k] = 0; n_launched[k] = g_launches[k].load(); }
  for (auto& s : g_samples) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) { ms_sum[s.kind] += ms; n_sampled[s.kind] += 1; }
    g_pool.push_back(s.a);
    g_pool.push_back(s.b);
  }
  g_samples.clear();
  return DEQSCI_OK;
}
