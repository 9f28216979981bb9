// trie_builder.h -- host-side builder of the engine's device trie.
//
// The reference looks pieces up in a Darts-clone double array
// (third_party/darts_clone/darts.h:469-547) built by unigram::Model::BuildTrie
// (src/unigram_model.cc:608-650).  Results depend only on the KEY SET, not on the
// array layout, so the engine uses its own double-array format designed for the
// GPU kernels:
//
//   link[u]  = (base << 11) | (kind << 9) | label9      one 32-bit word per unit
//              child of unit u on byte c is unit  v = base(u) ^ c,
//              valid iff label9(v) == c   (label9 == 0x100 marks an unused unit,
//              so a byte can never match it);
//              kind: 0 = no key ends here, 1 = NORMAL piece, 2 = USER_DEFINED,
//                    3 = UNUSED  -- i.e. the "has_leaf + value + type lookup" of
//              the reference (darts.h:50-80, model_interface.h:217-225) is folded
//              into the transition word, so a non-matching step costs ONE load;
//   val[u]   = float score bits of the piece ending at u (any kind != 0);
//   id[u]    = vocab id of that piece, -1 otherwise.
//
// Units are allocated in decreasing order of expected visit frequency (sum of
// exp(score) below the node), inside a sliding window of open 256-unit blocks, so
// that the FIRST H units are the hot ones: the kernels stage link[0..H) / val[0..H)
// into shared memory with one bulk (TMA) copy each and serve the rest from L2.
#ifndef SPM_B200_TRIE_BUILDER_H_
#define SPM_B200_TRIE_BUILDER_H_

#include <cstdint>
#include <string>
#include <vector>

namespace spm_b200 {

constexpr uint32_t kLinkLabelMask = 0x1FF;
constexpr uint32_t kLinkInvalidLabel = 0x100;
constexpr int kLinkKindShift = 9;
constexpr int kLinkBaseShift = 11;
constexpr uint32_t kMaxTrieUnits = 1u << 21;
enum TrieKind : uint32_t { kKindNone = 0, kKindNormal = 1, kKindUserDefined = 2, kKindUnused = 3 };

struct TrieKey {
  const char *data;
  uint32_t len;
  int32_t id;
  float score;    // stored in val[]
  float weight;   // visit-frequency proxy used for hot-first ordering
  uint32_t kind;  // TrieKind
};

struct DeviceTrie {
  std::vector<uint32_t> link;
  std::vector<uint32_t> val;
  std::vector<int32_t> id;
  std::vector<uint32_t> cmask;  // per unit: OR over children c of 1 << (c & 31); 0 = leaf (early walk termination)
  std::vector<uint32_t> unit_of_id;  // vocab id -> unit (0xFFFFFFFF if the id is not a key)
  uint32_t max_key_len = 0;
  uint32_t max_matches_per_start = 0;  // == trie_results_size_ of unigram_model.cc:635-644
  uint32_t num_nodes = 0;
};

// Returns false with *err set when keys are invalid (duplicate, empty, contain NUL)
// or the array would exceed kMaxTrieUnits.
bool BuildDeviceTrie(const std::vector<TrieKey> &keys, int vocab_size, DeviceTrie *out, std::string *err);

}  // namespace spm_b200
#endif
