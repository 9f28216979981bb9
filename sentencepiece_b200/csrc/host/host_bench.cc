// host/host_bench.cc -- timing hook for bench.py: the cost of the reference-shaped C++ API itself.
//
// SentencePieceProcessor::Encode(const std::vector<std::string_view>&, std::vector<std::vector<int>>*) is what a C++
// caller of the reference's class uses; on top of the engine call it packs the views into one buffer and
// materialises one std::vector<int> per sentence.  bench.py reports this path next to the packed C ABI so that the
// end-to-end number of the API a user really calls is on record (built into libspm_b200_hostbench.so; not part of
// the product library).
#include <chrono>
#include <cstdint>
#include <string_view>
#include <vector>

#include "host/sentencepiece_processor.h"

extern "C" int spm_hostclass_encode_bench(const char *model_path, int device, const char *bytes, const uint64_t *offsets,
                                          size_t n, int reps, double *best_ms, double *mean_ms, uint64_t *total_ids) {
  sentencepiece::SentencePieceProcessor sp;
  sp.SetDevice(device);
  if (!sp.Load(model_path).ok()) return 1;
  std::vector<std::string_view> views(n);
  for (size_t i = 0; i < n; ++i) views[i] = std::string_view(bytes + offsets[i], offsets[i + 1] - offsets[i]);
  std::vector<std::vector<int>> ids;
  double best = 1e300, sum = 0;
  for (int r = 0; r < reps + 1; ++r) {  // the first call is a warm-up (buffers of the engine grow)
    const auto t0 = std::chrono::steady_clock::now();
    if (!sp.Encode(views, &ids).ok()) return 2;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (r == 0) continue;
    best = ms < best ? ms : best;
    sum += ms;
  }
  uint64_t tot = 0;
  for (const auto &v : ids) tot += v.size();
  *best_ms = best;
  *mean_ms = sum / reps;
  *total_ids = tot;
  return 0;
}
