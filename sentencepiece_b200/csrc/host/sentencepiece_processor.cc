// host/sentencepiece_processor.cc -- see the header.  All heavy lifting is behind the C ABI.
#include "host/sentencepiece_processor.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <fstream>
#include <functional>
#include <set>
#include <sstream>
#include <thread>
#include <unordered_map>

#include "model_reader.h"
#include "spm_b200.h"

namespace sentencepiece {

struct SentencePieceProcessor::Impl {
  spm_b200::ModelData model;
  std::vector<std::string> pieces;
  std::unordered_map<std::string, int> piece_to_id;
  std::vector<uint8_t> loaded_types;
};

namespace {
std::atomic<unsigned int> g_seed{0};
std::atomic<bool> g_seed_set{false};
util::Status Internal(const std::string &m) { return util::Status(util::StatusCode::kInternal, m); }
util::Status FromEngine(const spm_engine *e, int rc) {
  if (rc == 0) return util::OkStatus();
  return Internal(std::string("spm_b200(") + std::to_string(rc) + "): " + spm_last_error(e));
}
}  // namespace

void SetRandomGeneratorSeed(unsigned int seed) {
  g_seed = seed;
  g_seed_set = true;
}

SentencePieceProcessor::SentencePieceProcessor() : impl_(new Impl) {}
SentencePieceProcessor::~SentencePieceProcessor() {
  if (engine_) spm_engine_destroy(engine_);
}

util::Status SentencePieceProcessor::status() const {
  if (!engine_) return Internal("Model is not initialized.");  // sentencepiece_processor.cc:293-299
  return util::OkStatus();
}

util::Status SentencePieceProcessor::Load(std::string_view filename) {
  std::ifstream f(std::string(filename), std::ios::binary);
  if (!f) return util::Status(util::StatusCode::kNotFound, "\"" + std::string(filename) + "\": No such file or directory");
  std::stringstream ss;
  ss << f.rdbuf();
  return LoadFromSerializedProto(ss.str());
}

util::Status SentencePieceProcessor::LoadFromSerializedProto(std::string_view serialized) {
  std::string err;
  if (!spm_b200::ParseModelProto(serialized.data(), serialized.size(), &impl_->model, &err)) return Internal(err);
  if (engine_) { spm_engine_destroy(engine_); engine_ = nullptr; }
  const int rc = spm_engine_create_from_serialized(serialized.data(), serialized.size(), device_, &engine_);
  if (rc) { engine_ = nullptr; return FromEngine(nullptr, rc); }
  const auto &m = impl_->model;
  impl_->pieces.clear();
  impl_->piece_to_id.clear();
  unk_id_ = -1;
  for (int i = 0; i < m.vocab_size(); ++i) {
    impl_->pieces.emplace_back(m.piece(i), m.piece_len(i));
    impl_->piece_to_id.emplace(impl_->pieces.back(), i);  // first definition wins, like InsertIfNotPresent
    if (m.types[i] == SPM_UNKNOWN) unk_id_ = i;
  }
  impl_->loaded_types = m.types;
  extra_.clear();
  seeded_ = false;
  return util::OkStatus();
}

bool SentencePieceProcessor::IsControl(int id) const {
  return id >= 0 && id < GetPieceSize() && impl_->model.types[id] == SPM_CONTROL;
}

// the reference's generator is thread_local and takes the global seed when it is first used (util.cc:202-204)
void SentencePieceProcessor::SeedOnce() const {
  if (seeded_) return;
  seeded_ = true;
  if (g_seed_set) spm_set_random_seed(engine_, g_seed);
}

util::Status SentencePieceProcessor::LoadVocabulary(std::string_view filename, int threshold) {
  if (!engine_) return status();
  std::ifstream f(std::string(filename), std::ios::binary);
  if (!f) return util::Status(util::StatusCode::kNotFound, "\"" + std::string(filename) + "\": No such file or directory");
  std::vector<std::string> vocab;
  std::string line;
  while (std::getline(f, line)) {
    const size_t tab = line.find('\t');
    const std::string piece = line.substr(0, tab);
    if (piece.empty()) return Internal("empty piece in the vocabulary file");
    long freq = 1;
    if (tab != std::string::npos) {
      const std::string fs = line.substr(tab + 1, line.find('\t', tab + 1) - tab - 1);
      char *end = nullptr;
      freq = strtol(fs.c_str(), &end, 10);
      if (fs.empty() || *end) return Internal("Could not parse the frequency");
    }
    if (freq >= threshold) vocab.push_back(piece);
  }
  std::vector<std::string_view> views(vocab.begin(), vocab.end());
  return SetVocabulary(views);
}

int SentencePieceProcessor::GetPieceSize() const { return static_cast<int>(impl_->pieces.size()); }
int SentencePieceProcessor::PieceToId(std::string_view piece) const {
  const auto it = impl_->piece_to_id.find(std::string(piece));
  return it == impl_->piece_to_id.end() ? unk_id_ : it->second;  // model_interface.cc:51-61
}
const std::string &SentencePieceProcessor::IdToPiece(int id) const {
  static const std::string kEmpty;
  return id >= 0 && id < GetPieceSize() ? impl_->pieces[id] : kEmpty;
}
int SentencePieceProcessor::bos_id() const {
  const int id = PieceToId(impl_->model.bos_piece);
  return id == unk_id_ ? -1 : id;
}
int SentencePieceProcessor::eos_id() const {
  const int id = PieceToId(impl_->model.eos_piece);
  return id == unk_id_ ? -1 : id;
}

// ParseExtraOptions, sentencepiece_processor.cc:1067-1101
util::Status SentencePieceProcessor::SetEncodeExtraOptions(std::string_view extra_option) {
  extra_.clear();
  if (extra_option.empty()) return util::OkStatus();
  if (!engine_) return status();
  size_t pos = 0;
  while (pos <= extra_option.size()) {
    size_t nxt = extra_option.find(':', pos);
    if (nxt == std::string_view::npos) nxt = extra_option.size();
    const std::string s(extra_option.substr(pos, nxt - pos));
    if (s == "bos") {
      if (IsUnknown(PieceToId(impl_->model.bos_piece))) return Internal("id for `" + impl_->model.bos_piece + "` is not defined.");
      extra_.push_back(BOS);
    } else if (s == "eos") {
      if (IsUnknown(PieceToId(impl_->model.eos_piece))) return Internal("id for `" + impl_->model.eos_piece + "` is not defined.");
      extra_.push_back(EOS);
    } else if (s == "reverse") {
      extra_.push_back(REVERSE);
    } else if (s == "unk" || s == "unk_piece") {
      extra_.push_back(UNK_PIECE);
    } else {
      extra_.clear();
      return Internal("option \"" + s + "\" is not available.");
    }
    pos = nxt + 1;
  }
  return util::OkStatus();
}

// sentencepiece_processor.cc:301-330: pieces of one character are always kept
util::Status SentencePieceProcessor::SetVocabulary(const std::vector<std::string_view> &valid_vocab) {
  if (!engine_) return status();
  auto &m = impl_->model;
  const std::set<std::string_view> vocab(valid_vocab.begin(), valid_vocab.end());
  static const unsigned char kLen[16] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 3, 4};
  for (int i = 0; i < m.vocab_size(); ++i) {
    uint8_t &t = m.types[i];
    if (t == SPM_CONTROL || t == SPM_UNKNOWN || t == SPM_USER_DEFINED) continue;
    const std::string &p = impl_->pieces[i];
    const bool one_char = kLen[static_cast<unsigned char>(p[0]) >> 4] == p.size();
    if (t == SPM_BYTE) continue;  // byte pieces never enter pieces_; their class is fixed
    t = (vocab.count(p) || one_char) ? SPM_NORMAL : SPM_UNUSED;
  }
  return FromEngine(engine_, spm_engine_set_types(engine_, m.types.data()));
}

util::Status SentencePieceProcessor::ResetVocabulary() {
  if (!engine_) return status();
  for (auto &t : impl_->model.types)
    if (t == SPM_UNUSED) t = SPM_NORMAL;
  return FromEngine(engine_, spm_engine_set_types(engine_, impl_->model.types.data()));
}

util::Status SentencePieceProcessor::EncodePacked(const char *bytes, const uint64_t *offsets, size_t n,
                                                  const int32_t **ids, const uint64_t **id_offsets) const {
  if (!engine_) return status();
  return FromEngine(engine_, spm_encode_ids(engine_, bytes, offsets, n, ids, id_offsets));
}

util::Status SentencePieceProcessor::Decode(const std::vector<std::vector<int>> &ids,
                                            std::vector<std::string> *detokenized) const {
  if (!engine_) return status();
  if (!detokenized) return Internal("output container is null");  // CHECK_OR_RETURN_STATUS_STL
  detokenized->clear();
  std::vector<int32_t> packed;
  std::vector<uint64_t> offs(1, 0);
  for (const auto &l : ids) {
    packed.insert(packed.end(), l.begin(), l.end());
    offs.push_back(packed.size());
  }
  const char *text = nullptr;
  const uint64_t *to = nullptr;
  const int rc = spm_decode_ids(engine_, packed.data(), offs.data(), ids.size(), &text, &to);
  if (rc) return FromEngine(engine_, rc);
  detokenized->reserve(ids.size());
  for (size_t i = 0; i < ids.size(); ++i) detokenized->emplace_back(text + to[i], text + to[i + 1]);
  return util::OkStatus();
}

util::Status SentencePieceProcessor::Decode(const std::vector<int> &ids, std::string *detokenized) const {
  if (!engine_) return status();
  if (!detokenized) return Internal("output container is null");
  std::vector<std::string> out;
  const util::Status st = Decode(std::vector<std::vector<int>>{ids}, &out);
  if (!st.ok()) return st;
  *detokenized = std::move(out[0]);
  return util::OkStatus();
}

std::string SentencePieceProcessor::DecodeIds(const std::vector<int> &ids) const {
  std::string out;
  Decode(ids, &out);
  return out;
}

namespace {
// runs fn(lo, hi) over [0, n) on a few host threads (the container conversions around a batch call are the only
// per-sentence host work left; the reference's Python layer uses a thread pool for the same reason,
// python/src/sentencepiece/sentencepiece.i:235-267)
template <typename F>
void ParallelFor(size_t n, F fn) {
  const size_t T = n < 16384 ? 1 : std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), 16));
  if (T == 1) { fn(size_t{0}, n); return; }
  std::vector<std::thread> th;
  const size_t per = (n + T - 1) / T;
  for (size_t t = 1; t < T; ++t) th.emplace_back(fn, std::min(n, t * per), std::min(n, (t + 1) * per));
  fn(size_t{0}, std::min(n, per));
  for (auto &t : th) t.join();
}

void Pack(const std::vector<std::string_view> &in, std::string *bytes, std::vector<uint64_t> *offs) {
  offs->resize(in.size() + 1);
  uint64_t total = 0;
  for (size_t i = 0; i < in.size(); ++i) { (*offs)[i] = total; total += in[i].size(); }
  (*offs)[in.size()] = total;
  bytes->resize(total);
  char *dst = bytes->data();
  const uint64_t *o = offs->data();
  ParallelFor(in.size(), [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i)
      if (!in[i].empty()) memcpy(dst + o[i], in[i].data(), in[i].size());
  });
}
}  // namespace

util::Status SentencePieceProcessor::Encode(const std::vector<std::string_view> &inputs,
                                            std::vector<std::vector<int>> *ids) const {
  if (!engine_) return status();
  if (!ids) return Internal("output container is null");  // CHECK_OR_RETURN_STATUS_STL
  ids->clear();
  std::string bytes;
  std::vector<uint64_t> offs;
  Pack(inputs, &bytes, &offs);
  const int32_t *out;
  const uint64_t *oo;
  const auto st = EncodePacked(bytes.data(), offs.data(), inputs.size(), &out, &oo);
  if (!st.ok()) return st;
  ids->resize(inputs.size());
  ParallelFor(inputs.size(), [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) {
      auto &v = (*ids)[i];
      v.assign(out + oo[i], out + oo[i + 1]);
      ApplyExtraIds(&v);
    }
  });
  return util::OkStatus();
}

// ApplyExtraOptions, sentencepiece_processor.cc:1019-1064
void SentencePieceProcessor::ApplyExtraIds(std::vector<int> *v) const {
  if (extra_.empty()) return;
  const int bos = PieceToId(impl_->model.bos_piece), eos = PieceToId(impl_->model.eos_piece);
  for (const auto o : extra_) {
    if (o == REVERSE) std::reverse(v->begin(), v->end());
    else if (o == EOS) v->push_back(eos);
    else if (o == BOS) v->insert(v->begin(), bos);
  }
}
void SentencePieceProcessor::ApplyExtraPieces(std::vector<std::string> *v, std::vector<int> *pid) const {
  if (extra_.empty()) return;
  const int bos = PieceToId(impl_->model.bos_piece), eos = PieceToId(impl_->model.eos_piece);
  for (const auto o : extra_) {
    if (o == REVERSE) { std::reverse(v->begin(), v->end()); std::reverse(pid->begin(), pid->end()); }
    else if (o == EOS) { v->push_back(impl_->model.eos_piece); pid->push_back(eos); }
    else if (o == BOS) { v->insert(v->begin(), impl_->model.bos_piece); pid->insert(pid->begin(), bos); }
    else if (o == UNK_PIECE)
      for (size_t k = 0; k < v->size(); ++k)
        if ((*pid)[k] == unk_id_) (*v)[k] = impl_->model.unk_piece;
  }
}

util::Status SentencePieceProcessor::Encode(const std::vector<std::string_view> &inputs,
                                            std::vector<std::vector<std::string>> *pieces) const {
  if (!engine_) return status();
  if (!pieces) return Internal("output container is null");
  pieces->clear();
  std::string bytes;
  std::vector<uint64_t> offs;
  Pack(inputs, &bytes, &offs);
  const int32_t *ids;
  const uint32_t *tok_end, *n2o;
  const uint64_t *ido, *no;
  const char *norm;
  const int rc = spm_encode_spans(engine_, bytes.data(), offs.data(), inputs.size(), &ids, &tok_end, &ido, &norm, &no, &n2o);
  if (rc) return FromEngine(engine_, rc);
  pieces->resize(inputs.size());
  std::vector<int> pid;
  for (size_t i = 0; i < inputs.size(); ++i) {
    auto &v = (*pieces)[i];
    pid.clear();
    uint32_t begin = 0;
    for (uint64_t k = ido[i]; k < ido[i + 1]; ++k) {
      // unknown pieces keep their normalized surface (sentencepiece_processor.cc:609-621)
      if (ids[k] == unk_id_) v.emplace_back(norm + no[i] + begin, tok_end[k] - begin);
      else v.push_back(impl_->pieces[ids[k]]);
      pid.push_back(ids[k]);
      begin = tok_end[k];
    }
    ApplyExtraPieces(&v, &pid);
  }
  return util::OkStatus();
}

// ---- n-best and sampling ----
//
// The engine returns id sequences.  The piece forms need, for every UNKNOWN token, its surface in the normalized
// text (sentencepiece_processor.cc:609-621); it is recovered by laying the id sequence over the normalized text: a
// known id must spell its piece (byte pieces their byte) at the current position, an unknown id covers one or more
// whole characters none of which has a one-character piece of its own (only such characters get an UNK node,
// unigram_model.cc:590-594), up to wherever the rest of the sequence fits.
bool SentencePieceProcessor::PiecesFromIds(std::string_view norm, const int32_t *ids, size_t n,
                                           std::vector<std::string> *pieces) const {
  pieces->assign(n, std::string());
  static const unsigned char kLen[16] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 3, 4};
  const auto &m = impl_->model;
  std::function<bool(size_t, size_t)> rec = [&](size_t k, size_t pos) -> bool {
    if (k == n) return pos == norm.size();
    const int id = ids[k];
    if (id < 0 || id >= GetPieceSize()) return false;
    const uint8_t t = m.types[id];
    if (t == SPM_BYTE) {  // <0xXX>: one byte of the text
      const std::string &p = impl_->pieces[id];
      if (pos >= norm.size() || p.size() != 6) return false;
      const int v = static_cast<int>(strtol(p.substr(3, 2).c_str(), nullptr, 16));
      if (static_cast<unsigned char>(norm[pos]) != v) return false;
      (*pieces)[k] = p;
      return rec(k + 1, pos + 1);
    }
    if (id != unk_id_) {
      const std::string &p = impl_->pieces[id];
      if (norm.compare(pos, p.size(), p) != 0) return false;
      (*pieces)[k] = p;
      return rec(k + 1, pos + p.size());
    }
    size_t end = pos;
    while (end < norm.size()) {
      const size_t cl = std::min<size_t>(kLen[static_cast<unsigned char>(norm[end]) >> 4], norm.size() - end);
      const auto it = impl_->piece_to_id.find(std::string(norm.substr(end, cl)));
      if (it != impl_->piece_to_id.end()) {
        const uint8_t ct = m.types[it->second];
        if (ct == SPM_NORMAL || ct == SPM_USER_DEFINED) break;  // this character has its own piece: no UNK node
      }
      end += cl;
      (*pieces)[k].assign(norm.substr(pos, end - pos));
      if (rec(k + 1, end)) return true;
    }
    return false;
  };
  return rec(0, 0);
}

util::Status SentencePieceProcessor::NBestEncode(const std::vector<std::string_view> &inputs, int nbest_size,
                                                 std::vector<std::vector<std::vector<int>>> *ids) const {
  if (!engine_) return status();
  if (!ids) return Internal("output container is null");
  ids->clear();
  std::string bytes;
  std::vector<uint64_t> offs;
  Pack(inputs, &bytes, &offs);
  const int32_t *out;
  const uint64_t *co;
  const float *sc;
  const uint32_t *nc;
  const int rc = spm_nbest_encode(engine_, bytes.data(), offs.data(), inputs.size(), nbest_size, &out, &co, &sc, &nc);
  if (rc) return FromEngine(engine_, rc);
  const size_t K = static_cast<size_t>(std::max(1, std::min(nbest_size, 1024)));
  ids->resize(inputs.size());
  for (size_t i = 0; i < inputs.size(); ++i) {
    auto &cands = (*ids)[i];
    cands.resize(nc[i]);
    for (uint32_t c = 0; c < nc[i]; ++c) {
      cands[c].assign(out + co[i * K + c], out + co[i * K + c + 1]);
      ApplyExtraIds(&cands[c]);
    }
  }
  return util::OkStatus();
}

util::Status SentencePieceProcessor::NBestEncode(const std::vector<std::string_view> &inputs, int nbest_size,
                                                 std::vector<std::vector<std::vector<std::string>>> *pieces) const {
  if (!engine_) return status();
  if (!pieces) return Internal("output container is null");
  pieces->clear();
  std::string bytes;
  std::vector<uint64_t> offs;
  Pack(inputs, &bytes, &offs);
  // normalized text first (the spans call), then the candidates: both results live in engine buffers of their own
  const int32_t *sids;
  const uint32_t *tok_end, *n2o;
  const uint64_t *ido, *no;
  const char *norm;
  int rc = spm_encode_spans(engine_, bytes.data(), offs.data(), inputs.size(), &sids, &tok_end, &ido, &norm, &no, &n2o);
  if (rc) return FromEngine(engine_, rc);
  std::vector<std::string> norms(inputs.size());
  for (size_t i = 0; i < inputs.size(); ++i) norms[i].assign(norm + no[i], norm + no[i + 1]);
  const int32_t *out;
  const uint64_t *co;
  const float *sc;
  const uint32_t *nc;
  rc = spm_nbest_encode(engine_, bytes.data(), offs.data(), inputs.size(), nbest_size, &out, &co, &sc, &nc);
  if (rc) return FromEngine(engine_, rc);
  const size_t K = static_cast<size_t>(std::max(1, std::min(nbest_size, 1024)));
  pieces->resize(inputs.size());
  std::vector<int> pid;
  for (size_t i = 0; i < inputs.size(); ++i) {
    auto &cands = (*pieces)[i];
    cands.resize(nc[i]);
    for (uint32_t c = 0; c < nc[i]; ++c) {
      const uint64_t lo = co[i * K + c], hi = co[i * K + c + 1];
      if (!PiecesFromIds(norms[i], out + lo, hi - lo, &cands[c]))
        return Internal("all normalized characters are not consumed.");  // sentencepiece_processor.cc:628-629
      pid.assign(out + lo, out + hi);
      ApplyExtraPieces(&cands[c], &pid);
    }
  }
  return util::OkStatus();
}

util::Status SentencePieceProcessor::SampleEncode(const std::vector<std::string_view> &inputs, int nbest_size, float alpha,
                                                  std::vector<std::vector<int>> *ids) const {
  if (!engine_) return status();
  if (!ids) return Internal("output container is null");
  ids->clear();
  SeedOnce();
  std::string bytes;
  std::vector<uint64_t> offs;
  Pack(inputs, &bytes, &offs);
  const int32_t *out;
  const uint64_t *oo;
  const int rc = spm_sample_encode_ids(engine_, bytes.data(), offs.data(), inputs.size(), nbest_size, alpha, &out, &oo);
  if (rc) return FromEngine(engine_, rc);
  ids->resize(inputs.size());
  for (size_t i = 0; i < inputs.size(); ++i) {
    (*ids)[i].assign(out + oo[i], out + oo[i + 1]);
    ApplyExtraIds(&(*ids)[i]);
  }
  return util::OkStatus();
}

util::Status SentencePieceProcessor::SampleEncode(const std::vector<std::string_view> &inputs, int nbest_size, float alpha,
                                                  std::vector<std::vector<std::string>> *pieces) const {
  if (!engine_) return status();
  if (!pieces) return Internal("output container is null");
  pieces->clear();
  SeedOnce();
  std::string bytes;
  std::vector<uint64_t> offs;
  Pack(inputs, &bytes, &offs);
  const int32_t *sids;
  const uint32_t *tok_end, *n2o;
  const uint64_t *ido, *no;
  const char *norm;
  int rc = spm_encode_spans(engine_, bytes.data(), offs.data(), inputs.size(), &sids, &tok_end, &ido, &norm, &no, &n2o);
  if (rc) return FromEngine(engine_, rc);
  std::vector<std::string> norms(inputs.size());
  for (size_t i = 0; i < inputs.size(); ++i) norms[i].assign(norm + no[i], norm + no[i + 1]);
  const int32_t *out;
  const uint64_t *oo;
  rc = spm_sample_encode_ids(engine_, bytes.data(), offs.data(), inputs.size(), nbest_size, alpha, &out, &oo);
  if (rc) return FromEngine(engine_, rc);
  pieces->resize(inputs.size());
  std::vector<int> pid;
  for (size_t i = 0; i < inputs.size(); ++i) {
    if (!PiecesFromIds(norms[i], out + oo[i], oo[i + 1] - oo[i], &(*pieces)[i]))
      return Internal("all normalized characters are not consumed.");
    pid.assign(out + oo[i], out + oo[i + 1]);
    ApplyExtraPieces(&(*pieces)[i], &pid);
  }
  return util::OkStatus();
}

util::Status SentencePieceProcessor::CalculateEntropy(const std::vector<std::string_view> &inputs, float alpha,
                                                      std::vector<float> *entropy) const {
  if (!engine_) return status();
  if (!entropy) return Internal("output container is null");
  entropy->clear();
  std::string bytes;
  std::vector<uint64_t> offs;
  Pack(inputs, &bytes, &offs);
  const float *ent = nullptr;
  const int rc = spm_calculate_entropy(engine_, bytes.data(), offs.data(), inputs.size(), alpha, &ent);
  if (rc) return FromEngine(engine_, rc);
  entropy->assign(ent, ent + inputs.size());
  return util::OkStatus();
}
util::Status SentencePieceProcessor::CalculateEntropy(std::string_view input, float alpha, float *entropy) const {
  if (!engine_) return status();
  if (!entropy) return Internal("output container is null");
  std::vector<float> out;
  const auto st = CalculateEntropy(std::vector<std::string_view>{input}, alpha, &out);
  if (!st.ok()) return st;
  *entropy = out[0];
  return util::OkStatus();
}
float SentencePieceProcessor::CalculateEntropy(std::string_view input, float alpha) const {
  float e = 0.f;
  CalculateEntropy(input, alpha, &e).IgnoreError();
  return e;
}

std::vector<std::pair<std::vector<int>, float>> SentencePieceProcessor::SampleEncodeAndScoreAsIds(
    std::string_view input, int num_samples, float alpha, bool wor, bool include_best) const {
  std::vector<std::pair<std::vector<int>, float>> out;
  if (!engine_) return out;
  SeedOnce();
  const uint64_t offs[2] = {0, input.size()};
  const int32_t *ids;
  const uint64_t *co;
  const float *sc;
  if (spm_sample_encode_and_score(engine_, input.data(), offs, 1, num_samples, alpha, wor, include_best, &ids, &co, &sc)) return out;
  for (int c = 0; c < num_samples; ++c) {
    out.emplace_back(std::vector<int>(ids + co[c], ids + co[c + 1]), sc[c]);
    ApplyExtraIds(&out.back().first);
  }
  return out;
}
std::vector<std::pair<std::vector<std::string>, float>> SentencePieceProcessor::SampleEncodeAndScoreAsPieces(
    std::string_view input, int num_samples, float alpha, bool wor, bool include_best) const {
  std::vector<std::pair<std::vector<std::string>, float>> out;
  if (!engine_) return out;
  SeedOnce();
  const uint64_t offs[2] = {0, input.size()};
  const int32_t *sids;
  const uint32_t *tok_end, *n2o;
  const uint64_t *ido, *no;
  const char *norm;
  if (spm_encode_spans(engine_, input.data(), offs, 1, &sids, &tok_end, &ido, &norm, &no, &n2o)) return out;
  const std::string normalized(norm + no[0], norm + no[1]);
  const int32_t *ids;
  const uint64_t *co;
  const float *sc;
  if (spm_sample_encode_and_score(engine_, input.data(), offs, 1, num_samples, alpha, wor, include_best, &ids, &co, &sc)) return out;
  std::vector<int> pid;
  for (int c = 0; c < num_samples; ++c) {
    std::vector<std::string> pcs;
    if (!PiecesFromIds(normalized, ids + co[c], co[c + 1] - co[c], &pcs)) return {};
    pid.assign(ids + co[c], ids + co[c + 1]);
    ApplyExtraPieces(&pcs, &pid);
    out.emplace_back(std::move(pcs), sc[c]);
  }
  return out;
}

util::Status SentencePieceProcessor::NBestEncode(std::string_view input, int nbest_size,
                                                 std::vector<std::vector<int>> *ids) const {
  if (!engine_) return status();
  if (!ids) return Internal("output container is null");
  std::vector<std::vector<std::vector<int>>> out;
  const auto st = NBestEncode(std::vector<std::string_view>{input}, nbest_size, &out);
  if (!st.ok()) return st;
  *ids = std::move(out[0]);
  return util::OkStatus();
}
util::Status SentencePieceProcessor::NBestEncode(std::string_view input, int nbest_size,
                                                 std::vector<std::vector<std::string>> *pieces) const {
  if (!engine_) return status();
  if (!pieces) return Internal("output container is null");
  std::vector<std::vector<std::vector<std::string>>> out;
  const auto st = NBestEncode(std::vector<std::string_view>{input}, nbest_size, &out);
  if (!st.ok()) return st;
  *pieces = std::move(out[0]);
  return util::OkStatus();
}
util::Status SentencePieceProcessor::SampleEncode(std::string_view input, int nbest_size, float alpha,
                                                  std::vector<int> *ids) const {
  if (!engine_) return status();
  if (!ids) return Internal("output container is null");
  std::vector<std::vector<int>> out;
  const auto st = SampleEncode(std::vector<std::string_view>{input}, nbest_size, alpha, &out);
  if (!st.ok()) return st;
  *ids = std::move(out[0]);
  return util::OkStatus();
}
util::Status SentencePieceProcessor::SampleEncode(std::string_view input, int nbest_size, float alpha,
                                                  std::vector<std::string> *pieces) const {
  if (!engine_) return status();
  if (!pieces) return Internal("output container is null");
  std::vector<std::vector<std::string>> out;
  const auto st = SampleEncode(std::vector<std::string_view>{input}, nbest_size, alpha, &out);
  if (!st.ok()) return st;
  *pieces = std::move(out[0]);
  return util::OkStatus();
}
std::vector<std::vector<std::string>> SentencePieceProcessor::NBestEncodeAsPieces(std::string_view input, int nbest_size) const {
  std::vector<std::vector<std::string>> out;
  NBestEncode(input, nbest_size, &out).IgnoreError();
  return out;
}
std::vector<std::vector<int>> SentencePieceProcessor::NBestEncodeAsIds(std::string_view input, int nbest_size) const {
  std::vector<std::vector<int>> out;
  NBestEncode(input, nbest_size, &out).IgnoreError();
  return out;
}
std::vector<std::string> SentencePieceProcessor::SampleEncodeAsPieces(std::string_view input, int nbest_size, float alpha) const {
  std::vector<std::string> out;
  SampleEncode(input, nbest_size, alpha, &out).IgnoreError();
  return out;
}
std::vector<int> SentencePieceProcessor::SampleEncodeAsIds(std::string_view input, int nbest_size, float alpha) const {
  std::vector<int> out;
  SampleEncode(input, nbest_size, alpha, &out).IgnoreError();
  return out;
}

util::Status SentencePieceProcessor::Encode(std::string_view input, std::vector<int> *ids) const {
  if (!engine_) return status();
  if (!ids) return Internal("output container is null");
  std::vector<std::vector<int>> out;
  const auto st = Encode(std::vector<std::string_view>{input}, &out);
  if (!st.ok()) return st;
  *ids = std::move(out[0]);
  return util::OkStatus();
}

util::Status SentencePieceProcessor::Encode(std::string_view input, std::vector<std::string> *pieces) const {
  if (!engine_) return status();
  if (!pieces) return Internal("output container is null");
  std::vector<std::vector<std::string>> out;
  const auto st = Encode(std::vector<std::string_view>{input}, &out);
  if (!st.ok()) return st;
  *pieces = std::move(out[0]);
  return util::OkStatus();
}

// the value-returning forms swallow errors (macro at sentencepiece_processor.h:432-436)
std::vector<std::string> SentencePieceProcessor::EncodeAsPieces(std::string_view input) const {
  std::vector<std::string> out;
  Encode(input, &out).IgnoreError();
  return out;
}
std::vector<int> SentencePieceProcessor::EncodeAsIds(std::string_view input) const {
  std::vector<int> out;
  Encode(input, &out).IgnoreError();
  return out;
}

}  // namespace sentencepiece
