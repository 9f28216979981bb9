// oracle/ref_shim.cc -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" shim (our own code) over the UNMODIFIED reference library
// compiled by oracle/Makefile into oracle/_ref/libsentencepiece.a.  It lets the
// Python tests and bench.py's reference arm call the reference's own
// SentencePieceProcessor::Encode / Normalize / SetVocabulary
// (src/sentencepiece_processor.h:245-460) through ctypes on packed buffers, so
// that the product and the reference see byte-identical inputs (including NUL
// bytes and malformed UTF-8, which text files/CLIs cannot carry).
//
// Nothing under sentencepiece_b200/ may link or load this.
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "sentencepiece_processor.h"

using sentencepiece::SentencePieceProcessor;

extern "C" {

void *ref_load_serialized(const void *data, size_t len) {
  auto *sp = new SentencePieceProcessor();
  const auto st = sp->LoadFromSerializedProto(
      absl::string_view(static_cast<const char *>(data), len));
  if (!st.ok()) {
    delete sp;
    return nullptr;
  }
  return sp;
}

void *ref_load(const char *path) {
  auto *sp = new SentencePieceProcessor();
  const auto st = sp->Load(path);
  if (!st.ok()) {
    delete sp;
    return nullptr;
  }
  return sp;
}

void ref_free(void *h) { delete static_cast<SentencePieceProcessor *>(h); }
void ref_free_buf(void *p) { free(p); }

int ref_set_encode_extra_options(void *h, const char *opt) {
  return static_cast<SentencePieceProcessor *>(h)->SetEncodeExtraOptions(opt).ok() ? 0 : 1;
}

// pieces: NUL-separated list of n valid pieces.  n == 0 -> ResetVocabulary.
int ref_set_vocabulary(void *h, const char *pieces, size_t n) {
  auto *sp = static_cast<SentencePieceProcessor *>(h);
  if (n == 0) return sp->ResetVocabulary().ok() ? 0 : 1;
  std::vector<absl::string_view> v;
  const char *p = pieces;
  for (size_t i = 0; i < n; ++i) {
    const size_t l = strlen(p);
    v.emplace_back(p, l);
    p += l + 1;
  }
  return sp->SetVocabulary(v).ok() ? 0 : 1;
}

// Batch EncodeAsIds over a packed buffer with `nthreads` std::threads pulling
// sentence indices from an atomic counter -- the same scheme as the reference's
// only batch entry (python/src/sentencepiece/sentencepiece.i:245-267).
// Outputs are malloc'ed: *ids (packed int32), id_offsets[n+1] caller-provided.
// Returns 0 on success, k+1 if sentence k failed.
int ref_encode_ids(void *h, const char *bytes, const uint64_t *offs, size_t n,
                   int nthreads, int32_t **ids_out, uint64_t *id_offsets) {
  auto *sp = static_cast<SentencePieceProcessor *>(h);
  std::vector<std::vector<int>> outs(n);
  std::atomic<size_t> next{0};
  std::atomic<size_t> failed{0};
  if (nthreads < 1) nthreads = 1;
  auto work = [&]() {
    size_t i;
    while ((i = next.fetch_add(1)) < n) {
      const auto st = sp->Encode(
          absl::string_view(bytes + offs[i], offs[i + 1] - offs[i]), &outs[i]);
      if (!st.ok()) failed.store(i + 1);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
  if (failed.load()) return static_cast<int>(failed.load());
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) {
    id_offsets[i] = total;
    total += outs[i].size();
  }
  id_offsets[n] = total;
  int32_t *ids = static_cast<int32_t *>(malloc(sizeof(int32_t) * (total ? total : 1)));
  for (size_t i = 0; i < n; ++i)
    if (!outs[i].empty())
      memcpy(ids + id_offsets[i], outs[i].data(), sizeof(int32_t) * outs[i].size());
  *ids_out = ids;
  return 0;
}

// Timing-only variant: encodes and discards (keeps only the id count) so the
// reference arm is not charged for our packing.  Returns total ids.
uint64_t ref_encode_count(void *h, const char *bytes, const uint64_t *offs,
                          size_t n, int nthreads) {
  auto *sp = static_cast<SentencePieceProcessor *>(h);
  std::atomic<size_t> next{0};
  std::atomic<uint64_t> total{0};
  if (nthreads < 1) nthreads = 1;
  auto work = [&]() {
    size_t i;
    std::vector<int> out;
    uint64_t local = 0;
    while ((i = next.fetch_add(1)) < n) {
      sp->Encode(absl::string_view(bytes + offs[i], offs[i + 1] - offs[i]), &out)
          .IgnoreError();
      local += out.size();
    }
    total.fetch_add(local);
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
  return total.load();
}

// Normalize one string; outputs malloc'ed normalized bytes and norm_to_orig
// (len+1 entries, or 0 entries when the normalized string is empty).
int ref_normalize(void *h, const char *s, size_t len, char **out, size_t *out_len,
                  uint64_t **n2o, size_t *n2o_len) {
  auto *sp = static_cast<SentencePieceProcessor *>(h);
  std::string norm;
  std::vector<size_t> map;
  const auto st = sp->Normalize(absl::string_view(s, len), &norm, &map);
  if (!st.ok()) return 1;
  *out = static_cast<char *>(malloc(norm.size() + 1));
  memcpy(*out, norm.data(), norm.size());
  *out_len = norm.size();
  *n2o = static_cast<uint64_t *>(malloc(sizeof(uint64_t) * (map.size() + 1)));
  for (size_t i = 0; i < map.size(); ++i) (*n2o)[i] = map[i];
  *n2o_len = map.size();
  return 0;
}

// EncodeAsPieces for one sentence; pieces returned NUL-joined (malloc'ed).
int ref_encode_pieces(void *h, const char *s, size_t len, char **out,
                      size_t *out_len, size_t *n_pieces) {
  auto *sp = static_cast<SentencePieceProcessor *>(h);
  std::vector<std::string> pieces;
  const auto st = sp->Encode(absl::string_view(s, len), &pieces);
  if (!st.ok()) return 1;
  std::string joined;
  for (const auto &p : pieces) {
    joined += p;
    joined.push_back('\0');
  }
  *out = static_cast<char *>(malloc(joined.size() + 1));
  memcpy(*out, joined.data(), joined.size());
  *out_len = joined.size();
  *n_pieces = pieces.size();
  return 0;
}

int ref_piece_size(void *h) {
  return static_cast<SentencePieceProcessor *>(h)->GetPieceSize();
}

// SentencePieceProcessor::Decode(const std::vector<int>&, std::string*) over a packed batch of id lists.
// Outputs: *text_out malloc'ed (concatenated), text_offsets[n+1] caller-provided.  Returns 0, or k+1 if list k failed.
int ref_decode_ids(void *h, const int32_t *ids, const uint64_t *id_offs, size_t n, int nthreads, char **text_out,
                   uint64_t *text_offsets) {
  auto *sp = static_cast<SentencePieceProcessor *>(h);
  std::vector<std::string> outs(n);
  std::atomic<size_t> next{0};
  std::atomic<size_t> failed{0};
  if (nthreads < 1) nthreads = 1;
  auto work = [&]() {
    size_t i;
    while ((i = next.fetch_add(1)) < n) {
      std::vector<int> v(ids + id_offs[i], ids + id_offs[i + 1]);
      if (!sp->Decode(v, &outs[i]).ok()) failed.store(i + 1);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
  if (failed.load()) return static_cast<int>(failed.load());
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) { text_offsets[i] = total; total += outs[i].size(); }
  text_offsets[n] = total;
  char *buf = static_cast<char *>(malloc(total ? total : 1));
  for (size_t i = 0; i < n; ++i)
    if (!outs[i].empty()) memcpy(buf + text_offsets[i], outs[i].data(), outs[i].size());
  *text_out = buf;
  return 0;
}

}  // extern "C"

// ---- n-best / sampling (config 5) -------------------------------------------------
#include "sentencepiece.pb.h"

extern "C" {

// NBestEncode for one sentence (src/sentencepiece_processor.cc:653-676).  Outputs (malloc'ed):
// ids of all candidates packed, cand_off[k+1], scores[k]; returns k (candidates) or -1.
int ref_nbest_encode(void *h, const char *s, size_t len, int nbest_size, int32_t **ids_out, uint32_t **cand_off_out,
                     float **scores_out) {
  auto *sp = static_cast<SentencePieceProcessor *>(h);
  sentencepiece::NBestSentencePieceText nb;
  if (!sp->NBestEncode(absl::string_view(s, len), nbest_size, &nb).ok()) return -1;
  const int k = nb.nbests_size();
  size_t total = 0;
  for (int i = 0; i < k; ++i) total += nb.nbests(i).pieces_size();
  int32_t *ids = static_cast<int32_t *>(malloc(sizeof(int32_t) * (total ? total : 1)));
  uint32_t *off = static_cast<uint32_t *>(malloc(sizeof(uint32_t) * (k + 1)));
  float *sc = static_cast<float *>(malloc(sizeof(float) * (k ? k : 1)));
  size_t p = 0;
  for (int i = 0; i < k; ++i) {
    off[i] = static_cast<uint32_t>(p);
    sc[i] = nb.nbests(i).score();
    for (const auto &piece : nb.nbests(i).pieces()) ids[p++] = piece.id();
  }
  off[k] = static_cast<uint32_t>(p);
  *ids_out = ids; *cand_off_out = off; *scores_out = sc;
  return k;
}

// SampleEncode(input, nbest_size, alpha) over a packed batch, sequentially on THIS thread, after
// seeding the thread's generator state deterministically: the reference seeds its thread_local
// mt19937 from the global seed on first use in a thread (src/util.cc:192-205), so the batch is run
// on a fresh std::thread.  Outputs like ref_encode_ids.
int ref_sample_encode_ids(void *h, const char *bytes, const uint64_t *offs, size_t n, int nbest_size, float alpha,
                          unsigned int seed, int32_t **ids_out, uint64_t *id_offsets) {
  auto *sp = static_cast<SentencePieceProcessor *>(h);
  std::vector<std::vector<int>> outs(n);
  int failed = 0;
  sentencepiece::SetRandomGeneratorSeed(seed);
  std::thread t([&]() {
    for (size_t i = 0; i < n; ++i)
      if (!sp->SampleEncode(absl::string_view(bytes + offs[i], offs[i + 1] - offs[i]), nbest_size, alpha, &outs[i]).ok())
        failed = static_cast<int>(i + 1);
  });
  t.join();
  if (failed) return failed;
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) { id_offsets[i] = total; total += outs[i].size(); }
  id_offsets[n] = total;
  int32_t *ids = static_cast<int32_t *>(malloc(sizeof(int32_t) * (total ? total : 1)));
  for (size_t i = 0; i < n; ++i)
    if (!outs[i].empty()) memcpy(ids + id_offsets[i], outs[i].data(), sizeof(int32_t) * outs[i].size());
  *ids_out = ids;
  return 0;
}

// CalculateEntropy(input, alpha) per sentence (src/sentencepiece_processor.cc:747-760).
int ref_entropy(void *h, const char *bytes, const uint64_t *offs, size_t n, float alpha, float *entropy) {
  auto *sp = static_cast<SentencePieceProcessor *>(h);
  for (size_t i = 0; i < n; ++i)
    if (!sp->CalculateEntropy(absl::string_view(bytes + offs[i], offs[i + 1] - offs[i]), alpha, &entropy[i]).ok())
      return static_cast<int>(i + 1);
  return 0;
}

// SampleEncodeAndScore(input, samples, alpha, wor, include_best) over a packed batch, sequentially on a fresh
// thread (see ref_sample_encode_ids).  Outputs: ids of all samples packed (malloc'ed), cand_off[n*samples+1],
// scores[n*samples]; a sentence for which the reference returns fewer results leaves empty slots.
int ref_sample_score(void *h, const char *bytes, const uint64_t *offs, size_t n, int samples, float alpha, int wor,
                     int include_best, unsigned int seed, int32_t **ids_out, uint64_t *cand_off, float *scores) {
  auto *sp = static_cast<SentencePieceProcessor *>(h);
  std::vector<std::vector<std::pair<std::vector<int>, float>>> outs(n);
  sentencepiece::SetRandomGeneratorSeed(seed);
  std::thread t([&]() {
    for (size_t i = 0; i < n; ++i)
      outs[i] = sp->SampleEncodeAndScoreAsIds(absl::string_view(bytes + offs[i], offs[i + 1] - offs[i]), samples, alpha,
                                              wor != 0, include_best != 0);
  });
  t.join();
  uint64_t total = 0;
  size_t c = 0;
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < samples; ++k, ++c) {
      cand_off[c] = total;
      scores[c] = 0.f;
      if (k < static_cast<int>(outs[i].size())) { total += outs[i][k].first.size(); scores[c] = outs[i][k].second; }
    }
  cand_off[c] = total;
  int32_t *ids = static_cast<int32_t *>(malloc(sizeof(int32_t) * (total ? total : 1)));
  c = 0;
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < samples; ++k, ++c)
      if (k < static_cast<int>(outs[i].size()) && !outs[i][k].first.empty())
        memcpy(ids + cand_off[c], outs[i][k].first.data(), sizeof(int32_t) * outs[i][k].first.size());
  *ids_out = ids;
  return 0;
}

}  // extern "C"
