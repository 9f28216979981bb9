// model_reader.h -- standalone reader for the reference's .model files.
//
// A .model file is a serialized `ModelProto`
// (reference: src/sentencepiece_model.proto:293-332).  The encode path needs only
// a handful of fields, so instead of linking protobuf we decode the wire format
// directly (varint / length-delimited / fixed32).  Unknown fields are skipped, so
// models written by newer trainers still load.
#ifndef SPM_B200_MODEL_READER_H_
#define SPM_B200_MODEL_READER_H_

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace spm_b200 {

struct ModelData {
  int32_t model_type = 1;  // TrainerSpec.model_type default UNIGRAM (:54)
  std::string piece_bytes;
  std::vector<uint32_t> piece_off;  // [vocab+1]
  std::vector<float> scores;
  std::vector<uint8_t> types;
  bool byte_fallback = false;               // TrainerSpec :194
  bool treat_whitespace_as_suffix = false;  // TrainerSpec :151
  bool add_dummy_prefix = true;             // NormalizerSpec :256
  bool remove_extra_whitespaces = true;     // NormalizerSpec :259
  bool escape_whitespaces = true;           // NormalizerSpec :263
  std::string charsmap;                     // NormalizerSpec.precompiled_charsmap :252
  std::string unk_piece = "<unk>", bos_piece = "<s>", eos_piece = "</s>", pad_piece = "<pad>";  // :220-223
  std::string unk_surface = " \xE2\x81\x87 ";  // TrainerSpec :228 (Decode)
  std::string denormalizer_charsmap;           // ModelProto.denormalizer_spec :327 (Decode; not supported on the device)
  std::vector<std::pair<std::string, std::string>> self_test;  // SelfTestData :277-283

  int vocab_size() const { return static_cast<int>(scores.size()); }
  const char *piece(int i) const { return piece_bytes.data() + piece_off[i]; }
  size_t piece_len(int i) const { return piece_off[i + 1] - piece_off[i]; }
};

// Returns false (and sets *err) when the buffer is not a well-formed ModelProto.
bool ParseModelProto(const void *data, size_t len, ModelData *out, std::string *err);

}  // namespace spm_b200
#endif
