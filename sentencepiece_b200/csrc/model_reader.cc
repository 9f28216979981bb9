// model_reader.cc -- see model_reader.h.
#include "model_reader.h"

#include <cstring>

namespace spm_b200 {
namespace {

struct Cursor {
  const uint8_t *p, *end;
  bool ok = true;
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      if (p >= end) { ok = false; return 0; }
      const uint8_t b = *p++;
      v |= static_cast<uint64_t>(b & 0x7F) << shift;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  // Reads one field header + payload; for length-delimited fields sub points at the payload.
  bool field(uint32_t *fno, uint32_t *wt, uint64_t *val, Cursor *sub) {
    const uint64_t key = varint();
    if (!ok) return false;
    *fno = static_cast<uint32_t>(key >> 3);
    *wt = static_cast<uint32_t>(key & 7);
    switch (*wt) {
      case 0: *val = varint(); return ok;
      case 1:
        if (end - p < 8) return ok = false;
        memcpy(val, p, 8); p += 8; return true;
      case 2: {
        const uint64_t l = varint();
        if (!ok || l > static_cast<uint64_t>(end - p)) return ok = false;
        sub->p = p; sub->end = p + l; sub->ok = true;
        p += l;
        return true;
      }
      case 5: {
        if (end - p < 4) return ok = false;
        uint32_t v32; memcpy(&v32, p, 4); p += 4; *val = v32; return true;
      }
      default: return ok = false;  // groups are not used by this schema
    }
  }
  std::string str() const { return std::string(reinterpret_cast<const char *>(p), end - p); }
};

}  // namespace

bool ParseModelProto(const void *data, size_t len, ModelData *m, std::string *err) {
  *m = ModelData();
  Cursor c{static_cast<const uint8_t *>(data), static_cast<const uint8_t *>(data) + len};
  m->piece_off.push_back(0);
  uint32_t fno, wt; uint64_t v; Cursor s{nullptr, nullptr};
  while (!c.done()) {
    if (!c.field(&fno, &wt, &v, &s)) { *err = "malformed ModelProto"; return false; }
    if (wt != 2) continue;
    if (fno == 1) {  // repeated SentencePiece pieces = 1
      std::string piece; float score = 0.f; uint8_t type = 1;
      uint32_t f2, w2; uint64_t v2; Cursor s2{nullptr, nullptr};
      while (!s.done()) {
        if (!s.field(&f2, &w2, &v2, &s2)) { *err = "malformed SentencePiece"; return false; }
        if (f2 == 1 && w2 == 2) piece = s2.str();
        else if (f2 == 2 && w2 == 5) { const uint32_t b = static_cast<uint32_t>(v2); memcpy(&score, &b, 4); }
        else if (f2 == 3 && w2 == 0) type = static_cast<uint8_t>(v2);
      }
      m->piece_bytes += piece;
      m->piece_off.push_back(static_cast<uint32_t>(m->piece_bytes.size()));
      m->scores.push_back(score);
      m->types.push_back(type);
    } else if (fno == 2) {  // TrainerSpec
      uint32_t f2, w2; uint64_t v2; Cursor s2{nullptr, nullptr};
      while (!s.done()) {
        if (!s.field(&f2, &w2, &v2, &s2)) { *err = "malformed TrainerSpec"; return false; }
        if (f2 == 3 && w2 == 0) m->model_type = static_cast<int32_t>(v2);
        else if (f2 == 35 && w2 == 0) m->byte_fallback = v2 != 0;
        else if (f2 == 24 && w2 == 0) m->treat_whitespace_as_suffix = v2 != 0;
        else if (f2 == 44 && w2 == 2) m->unk_surface = s2.str();
        else if (f2 == 45 && w2 == 2) m->unk_piece = s2.str();
        else if (f2 == 46 && w2 == 2) m->bos_piece = s2.str();
        else if (f2 == 47 && w2 == 2) m->eos_piece = s2.str();
        else if (f2 == 48 && w2 == 2) m->pad_piece = s2.str();
      }
    } else if (fno == 3) {  // NormalizerSpec
      uint32_t f2, w2; uint64_t v2; Cursor s2{nullptr, nullptr};
      while (!s.done()) {
        if (!s.field(&f2, &w2, &v2, &s2)) { *err = "malformed NormalizerSpec"; return false; }
        if (f2 == 2 && w2 == 2) m->charsmap = s2.str();
        else if (f2 == 3 && w2 == 0) m->add_dummy_prefix = v2 != 0;
        else if (f2 == 4 && w2 == 0) m->remove_extra_whitespaces = v2 != 0;
        else if (f2 == 5 && w2 == 0) m->escape_whitespaces = v2 != 0;
      }
    } else if (fno == 5) {  // denormalizer_spec (a NormalizerSpec)
      uint32_t f2, w2; uint64_t v2; Cursor s2{nullptr, nullptr};
      while (!s.done()) {
        if (!s.field(&f2, &w2, &v2, &s2)) { *err = "malformed denormalizer_spec"; return false; }
        if (f2 == 2 && w2 == 2) m->denormalizer_charsmap = s2.str();
      }
    } else if (fno == 4) {  // SelfTestData
      uint32_t f2, w2; uint64_t v2; Cursor s2{nullptr, nullptr};
      while (!s.done()) {
        if (!s.field(&f2, &w2, &v2, &s2)) { *err = "malformed SelfTestData"; return false; }
        if (f2 == 1 && w2 == 2) {
          std::string in, expected;
          uint32_t f3, w3; uint64_t v3; Cursor s3{nullptr, nullptr};
          while (!s2.done()) {
            if (!s2.field(&f3, &w3, &v3, &s3)) { *err = "malformed SelfTestData.Sample"; return false; }
            if (f3 == 1 && w3 == 2) in = s3.str();
            else if (f3 == 2 && w3 == 2) expected = s3.str();
          }
          m->self_test.emplace_back(in, expected);
        }
      }
    }
  }
  if (m->scores.empty()) { *err = "model has no pieces"; return false; }
  return true;
}

}  // namespace spm_b200
