// oracle/ref_bench.cc -- TEST/BENCH INFRASTRUCTURE ONLY (the "reference arm").
//
// Times the UNMODIFIED reference SentencePieceProcessor::Encode(ids)
// (src/sentencepiece_processor.cc:392-403) on the host cores of this box over a
// text file (one sentence per line), using T std::threads that pull line indices
// from an atomic counter exactly like the reference's own batch entry
// (python/src/sentencepiece/sentencepiece.i:245-267).  T=1 reproduces the
// single-threaded spm_encode loop (src/spm_encode_main.cc:159-165) without the
// text formatting.  Prints one JSON line per run.
//
// usage: ref_bench MODEL INPUT.txt THREADS [REPEAT] [--dump-ids OUT.bin]
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "sentencepiece_processor.h"

int main(int argc, char **argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s MODEL INPUT THREADS [REPEAT] [--dump-ids OUT]\n", argv[0]);
    return 2;
  }
  const std::string model = argv[1], input = argv[2];
  const int threads = std::max(1, atoi(argv[3]));
  int repeat = 1;
  std::string dump;
  for (int i = 4; i < argc; ++i) {
    if (!strcmp(argv[i], "--dump-ids") && i + 1 < argc) dump = argv[++i];
    else repeat = std::max(1, atoi(argv[i]));
  }
  sentencepiece::SentencePieceProcessor sp;
  if (!sp.Load(model).ok()) {
    fprintf(stderr, "cannot load %s\n", model.c_str());
    return 1;
  }
  std::vector<std::string> lines;
  {
    std::ifstream ifs(input);
    std::string l;
    while (std::getline(ifs, l)) lines.push_back(l);
  }
  uint64_t in_bytes = 0;
  for (const auto &l : lines) in_bytes += l.size();
  const size_t n = lines.size();
  std::vector<std::vector<int>> outs(dump.empty() ? 0 : n);
  double best = 1e30;
  uint64_t total_ids = 0;
  for (int r = 0; r < repeat; ++r) {
    std::atomic<size_t> next{0};
    std::atomic<uint64_t> total{0};
    auto work = [&]() {
      size_t i;
      std::vector<int> out;
      uint64_t local = 0;
      while ((i = next.fetch_add(1)) < n) {
        sp.Encode(lines[i], &out).IgnoreError();
        local += out.size();
        if (!dump.empty()) outs[i] = out;
      }
      total.fetch_add(local);
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 1; t < threads; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (s < best) best = s;
    total_ids = total.load();
  }
  if (!dump.empty()) {
    FILE *f = fopen(dump.c_str(), "wb");
    for (size_t i = 0; i < n; ++i) {
      const int32_t c = static_cast<int32_t>(outs[i].size());
      fwrite(&c, 4, 1, f);
      if (c) fwrite(outs[i].data(), 4, c, f);
    }
    fclose(f);
  }
  printf("{\"impl\": \"reference\", \"threads\": %d, \"sentences\": %zu, \"input_bytes\": %llu, "
         "\"ids\": %llu, \"seconds\": %.6f, \"sentences_per_sec\": %.1f, \"input_MBps\": %.3f}\n",
         threads, n, (unsigned long long)in_bytes, (unsigned long long)total_ids, best,
         n / best, in_bytes / best / 1e6);
  return 0;
}
