// host/sentencepiece_processor.h -- C++ host layer over the C ABI (include/spm_b200.h).
//
// Keeps the reference's public encode surface for this path -- class name, namespace,
// method names, argument meaning and error behaviour of
// /root/reference/src/sentencepiece_processor.h:238-460 (Load :245,
// LoadFromSerializedProto :261, SetEncodeExtraOptions :267, SetVocabulary :276,
// Encode(pieces) :295, Encode(ids) :299, EncodeAsPieces :453, EncodeAsIds :458) --
// so a caller such as spm_encode compiles against it unchanged, and adds the batch
// overloads that a GPU needs.  Single-sentence calls are batches of one.
// N-best and sampling (NBestEncode :318-324, SampleEncode :345-351, the value-returning forms :465-481,
// SetRandomGeneratorSeed :731) ride on the engine's n-best / sampling kernels; training stays the reference's.
#ifndef SPM_B200_HOST_SENTENCEPIECE_PROCESSOR_H_
#define SPM_B200_HOST_SENTENCEPIECE_PROCESSOR_H_

#include <cstdint>
#include <memory>
#include <string>
#include <string_view>
#include <vector>

struct spm_engine;

namespace sentencepiece {
namespace util {

enum class StatusCode : int { kOk = 0, kNotFound = 5, kInternal = 13 };

// Value-type status, ok() <=> no error (reference: sentencepiece_processor.h:54-76).
class Status {
 public:
  Status() = default;
  Status(StatusCode code, std::string msg) : code_(code), msg_(std::move(msg)) {}
  bool ok() const { return code_ == StatusCode::kOk; }
  StatusCode code() const { return code_; }
  const std::string &message() const { return msg_; }
  std::string ToString() const { return ok() ? "OK" : msg_; }
  void IgnoreError() const {}

 private:
  StatusCode code_ = StatusCode::kOk;
  std::string msg_;
};
inline Status OkStatus() { return Status(); }

}  // namespace util

class SentencePieceProcessor {
 public:
  SentencePieceProcessor();
  ~SentencePieceProcessor();
  SentencePieceProcessor(const SentencePieceProcessor &) = delete;
  SentencePieceProcessor &operator=(const SentencePieceProcessor &) = delete;

  // `device`: CUDA ordinal (one engine per GPU).  Set before Load.
  void SetDevice(int device) { device_ = device; }

  util::Status Load(std::string_view filename);
  util::Status LoadFromSerializedProto(std::string_view serialized);
  util::Status status() const;

  // "bos", "eos", "reverse", "unk"/"unk_piece", colon separated (sentencepiece_processor.cc:1067-1101)
  util::Status SetEncodeExtraOptions(std::string_view extra_option);
  util::Status SetVocabulary(const std::vector<std::string_view> &valid_vocab);
  util::Status ResetVocabulary();
  // "<piece>\t<freq>" lines; pieces with freq < threshold become UNUSED (sentencepiece_processor.cc:342-365)
  util::Status LoadVocabulary(std::string_view filename, int threshold);

  // ---- single sentence (reference signatures) ----
  util::Status Encode(std::string_view input, std::vector<std::string> *pieces) const;
  util::Status Encode(std::string_view input, std::vector<int> *ids) const;
  std::vector<std::string> EncodeAsPieces(std::string_view input) const;
  std::vector<int> EncodeAsIds(std::string_view input) const;

  // ---- n-best and sampling, single sentence (reference signatures, sentencepiece_processor.h:318-351,465-481).
  //      Errors like the reference: n-best is not available for BPE models (sentencepiece_processor.cc:662-663),
  //      nbest_size > 512 fails SampleEncode (:683). ----
  util::Status NBestEncode(std::string_view input, int nbest_size, std::vector<std::vector<std::string>> *pieces) const;
  util::Status NBestEncode(std::string_view input, int nbest_size, std::vector<std::vector<int>> *ids) const;
  util::Status SampleEncode(std::string_view input, int nbest_size, float alpha, std::vector<std::string> *pieces) const;
  util::Status SampleEncode(std::string_view input, int nbest_size, float alpha, std::vector<int> *ids) const;
  std::vector<std::vector<std::string>> NBestEncodeAsPieces(std::string_view input, int nbest_size) const;
  std::vector<std::vector<int>> NBestEncodeAsIds(std::string_view input, int nbest_size) const;
  std::vector<std::string> SampleEncodeAsPieces(std::string_view input, int nbest_size, float alpha) const;
  std::vector<int> SampleEncodeAsIds(std::string_view input, int nbest_size, float alpha) const;
  // ---- full-lattice operations of unigram models (sentencepiece_processor.h:364-379,401-410,483-500,521-526):
  //      num_samples lattice samples with their scores (wor / include_best are not on the accelerated path), and the
  //      entropy of the segmentation lattice ----
  std::vector<std::pair<std::vector<int>, float>> SampleEncodeAndScoreAsIds(std::string_view input, int num_samples,
                                                                            float alpha, bool wor, bool include_best) const;
  std::vector<std::pair<std::vector<std::string>, float>> SampleEncodeAndScoreAsPieces(std::string_view input, int num_samples,
                                                                                       float alpha, bool wor,
                                                                                       bool include_best) const;
  util::Status CalculateEntropy(std::string_view input, float alpha, float *entropy) const;
  float CalculateEntropy(std::string_view input, float alpha) const;
  util::Status CalculateEntropy(const std::vector<std::string_view> &inputs, float alpha, std::vector<float> *entropy) const;

  // ---- batch (what the reference's Python layer does with a thread pool,
  //      python/src/sentencepiece/sentencepiece.i:245-267) ----
  util::Status Encode(const std::vector<std::string_view> &inputs, std::vector<std::vector<int>> *ids) const;
  util::Status Encode(const std::vector<std::string_view> &inputs, std::vector<std::vector<std::string>> *pieces) const;
  // one device call for all sentences; the draws of SampleEncode are taken in input order on one generator, which is
  // what a single-threaded loop over the reference's SampleEncode does
  util::Status NBestEncode(const std::vector<std::string_view> &inputs, int nbest_size,
                           std::vector<std::vector<std::vector<int>>> *ids) const;
  util::Status NBestEncode(const std::vector<std::string_view> &inputs, int nbest_size,
                           std::vector<std::vector<std::vector<std::string>>> *pieces) const;
  util::Status SampleEncode(const std::vector<std::string_view> &inputs, int nbest_size, float alpha,
                            std::vector<std::vector<int>> *ids) const;
  util::Status SampleEncode(const std::vector<std::string_view> &inputs, int nbest_size, float alpha,
                            std::vector<std::vector<std::string>> *pieces) const;
  // zero-copy form: packed input, packed output owned by the engine until the next call
  util::Status EncodePacked(const char *bytes, const uint64_t *offsets, size_t n, const int32_t **ids,
                            const uint64_t **id_offsets) const;

  // ---- Decode(ids) (src/sentencepiece_processor.h Decode(const std::vector<int>&, std::string*)); the batch form is
  //      one device call for all lists.  An id outside [0, GetPieceSize()) fails the call like the reference's
  //      kOutOfRange (sentencepiece_processor.cc:915-918); decode_extra_options are not applied. ----
  util::Status Decode(const std::vector<int> &ids, std::string *detokenized) const;
  util::Status Decode(const std::vector<std::vector<int>> &ids, std::vector<std::string> *detokenized) const;
  std::string DecodeIds(const std::vector<int> &ids) const;

  // ---- vocabulary accessors used by callers of the encode path ----
  int GetPieceSize() const;
  int PieceToId(std::string_view piece) const;
  const std::string &IdToPiece(int id) const;
  bool IsUnknown(int id) const { return id == unk_id_; }
  bool IsControl(int id) const;
  int unk_id() const { return unk_id_; }
  int bos_id() const;
  int eos_id() const;

 private:
  struct Impl;
  enum ExtraOption { REVERSE, BOS, EOS, UNK_PIECE };
  std::unique_ptr<Impl> impl_;
  spm_engine *engine_ = nullptr;
  int device_ = 0;
  int unk_id_ = -1;
  std::vector<ExtraOption> extra_;
  mutable bool seeded_ = false;  // the engine's generator took the global seed (first sampling call)
  void SeedOnce() const;
  void ApplyExtraIds(std::vector<int> *v) const;
  void ApplyExtraPieces(std::vector<std::string> *v, std::vector<int> *pid) const;
  // pieces of an id sequence over the sentence's normalized text (unknown tokens keep their surface)
  bool PiecesFromIds(std::string_view normalized, const int32_t *ids, size_t n, std::vector<std::string> *pieces) const;
};

// sentencepiece_processor.h:731 / util.cc:23-31: the seed every generator created afterwards starts from; this
// layer's generators live in the engines and take it at their first sampling call
void SetRandomGeneratorSeed(unsigned int seed);

}  // namespace sentencepiece
#endif
