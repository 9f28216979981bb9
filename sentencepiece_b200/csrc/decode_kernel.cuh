// decode_kernel.cuh -- K7: batched Decode(ids) -> text (the step after the path).
//
// Reference: SentencePieceProcessor::Decode(const std::vector<int>&, SentencePieceText*)
// (src/sentencepiece_processor.cc:765-925), text only:
//   * every id must be in range (:915-918);
//   * CONTROL pieces are invisible, UNKNOWN pieces become TrainerSpec.unk_surface (:779-790);
//   * while the text is still empty and no leading U+2581 has been consumed yet (`is_bos_ws`,
//     :879-887), a piece that starts with U+2581 loses it when add_dummy_prefix or
//     remove_extra_whitespaces is set; with remove_extra_whitespaces the state is not closed by
//     that, so every leading U+2581 goes until the text is non-empty (:792-807);
//   * U+2581 -> ' ' everywhere else (:809);
//   * a run of BYTE pieces is reassembled one Unicode character at a time, every structurally
//     invalid byte becoming U+FFFD (:817-876).
// The engine precomputes, per id, the decoded bytes (U+2581 already replaced) and an info word, so
// the kernel is a gather: one WARP per id list, one token per lane; the bos rule is evaluated with a
// ballot ("the first token that would end the bos state"), output positions with a warp scan.
// Lists that contain BYTE pieces take a sequential path on lane 0 (the UTF-8 reassembly is a
// character-by-character loop in the reference too).  Text is appended to a temporary buffer in
// completion order; the scan / gather kernels of the encode path put it into list order.
#ifndef SPM_B200_DECODE_KERNEL_CUH_
#define SPM_B200_DECODE_KERNEL_CUH_

#include <cstdint>

namespace spm_b200 {

// dec_info word
constexpr uint32_t kDecKindMask = 3u;     // 0 normal, 1 control, 2 unknown, 3 byte
constexpr uint32_t kDecKindNormal = 0u, kDecKindControl = 1u, kDecKindUnknown = 2u, kDecKindByte = 3u;
constexpr uint32_t kDecLeadWs = 1u << 2;  // normal piece that starts with U+2581
constexpr uint32_t kDecBadByte = 1u << 3; // BYTE piece that is not "<0xXX>" (PieceToByte fails, model_interface.cc:214-230)
constexpr uint32_t kDecByteShift = 8;

struct KDecode {
  const int32_t *ids;
  const unsigned long long *id_offsets;  // [n+1]
  uint32_t n;
  int32_t vocab;
  const uint32_t *dec_off;   // [vocab+1] into dec_bytes
  const uint8_t *dec_bytes;
  const uint32_t *dec_info;  // [vocab]
  uint32_t strip;            // add_dummy_prefix || remove_extra_whitespaces
  uint32_t rm;               // remove_extra_whitespaces
  uint8_t *tmp;              // text in completion order
  unsigned long long tmp_cap;
  unsigned long long *cursor;
  unsigned long long *sent_start;  // [n]
  uint32_t *sent_count;            // [n]
  uint32_t *status;                // [1] error: 2 = id out of range, 1 = malformed byte piece; [2] overflow; [3] the offending id
};

// IsValidDecodeUTF8 (src/util.h:173-176 over DecodeUTF8, src/util.cc:51-84): bytes consumed by one valid character
// at b[0..avail), 0 if the byte is structurally invalid (the caller then consumes one byte).
__device__ __forceinline__ uint32_t dec_utf8_valid(const uint32_t *b, uint32_t avail) {
  const uint32_t b0 = b[0];
  auto trail = [](uint32_t x) { return (x & 0xC0u) == 0x80u; };
  if (b0 < 0x80u) return 1;
  if (avail >= 2 && (b0 & 0xE0u) == 0xC0u) {
    const uint32_t cp = ((b0 & 0x1Fu) << 6) | (b[1] & 0x3Fu);
    return (trail(b[1]) && cp >= 0x80u) ? 2u : 0u;
  }
  if (avail >= 3 && (b0 & 0xF0u) == 0xE0u) {
    const uint32_t cp = ((b0 & 0x0Fu) << 12) | ((b[1] & 0x3Fu) << 6) | (b[2] & 0x3Fu);
    return (trail(b[1]) && trail(b[2]) && cp >= 0x800u && (cp < 0xD800u || cp >= 0xE000u)) ? 3u : 0u;
  }
  if (avail >= 4 && (b0 & 0xF8u) == 0xF0u) {
    const uint32_t cp = ((b0 & 0x07u) << 18) | ((b[1] & 0x3Fu) << 12) | ((b[2] & 0x3Fu) << 6) | (b[3] & 0x3Fu);
    return (trail(b[1]) && trail(b[2]) && trail(b[3]) && cp >= 0x10000u && cp <= 0x10FFFFu) ? 4u : 0u;
  }
  return 0;
}

// Sequential restatement for one list (lists with BYTE pieces); WRITE = false only counts.
template <bool WRITE>
__device__ __forceinline__ uint32_t decode_list_sequential(const KDecode &D, unsigned long long a, uint32_t L, uint8_t *out) {
  uint32_t n = 0;  // bytes of text so far
  bool is_bos_ws = true, bos_ws_seen = false;
  uint32_t i = 0;
  while (i < L) {
    const int32_t id = D.ids[a + i];
    const uint32_t info = __ldg(D.dec_info + id);
    const uint32_t kind = info & kDecKindMask;
    if (kind == kDecKindByte) {
      // one Unicode character from the run of byte pieces that starts here
      uint32_t b[4] = {info >> kDecByteShift, 0, 0, 0};
      uint32_t avail = 1;
      while (avail < 4 && i + avail < L) {
        const uint32_t inf2 = __ldg(D.dec_info + D.ids[a + i + avail]);
        if ((inf2 & kDecKindMask) != kDecKindByte) break;
        b[avail] = inf2 >> kDecByteShift;
        ++avail;
      }
      const uint32_t c = dec_utf8_valid(b, avail);
      if (c == 0) {
        if (WRITE) { out[n] = 0xEF; out[n + 1] = 0xBF; out[n + 2] = 0xBD; }
        n += 3;
        i += 1;
      } else {
        if (WRITE) for (uint32_t k = 0; k < c; ++k) out[n + k] = static_cast<uint8_t>(b[k]);
        n += c;
        i += c;
      }
      continue;
    }
    if (bos_ws_seen || n != 0) is_bos_ws = false;
    const uint32_t off = __ldg(D.dec_off + id);
    uint32_t len = __ldg(D.dec_off + id + 1) - off;
    uint32_t skip = 0;
    bool has_bos_ws = false;
    if (kind == kDecKindNormal && is_bos_ws && D.strip && (info & kDecLeadWs)) {
      skip = 1;  // the ' ' that stands for the leading U+2581
      has_bos_ws = !D.rm;
    }
    len -= skip;
    if (WRITE) for (uint32_t k = 0; k < len; ++k) out[n + k] = __ldg(D.dec_bytes + off + skip + k);
    n += len;
    bos_ws_seen = has_bos_ws;
    ++i;
  }
  return n;
}

__global__ void __launch_bounds__(256) decode_warp_kernel(const KDecode D) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t sent = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; sent < D.n; sent += warps) {
    const unsigned long long a = D.id_offsets[sent];
    const unsigned long long len64 = D.id_offsets[sent + 1] - a;
    const uint32_t L = static_cast<uint32_t>(len64);
    // ---- pass 0: ids in range, any BYTE piece ----
    uint32_t err = len64 > 0x7FFFFFFFull ? 2u : 0u;  // 2 = id out of range (checked first by the reference), 1 = malformed byte piece
    int32_t err_id = 0;
    bool has_byte = false, stop = err != 0;
    for (uint32_t base = 0; base < L && !stop; base += 32) {
      const uint32_t t = base + lane;
      if (t < L) {
        const int32_t id = D.ids[a + t];
        if (id < 0 || id >= D.vocab) { err = 2; err_id = id; }
        else {
          const uint32_t info = __ldg(D.dec_info + id);
          has_byte |= (info & kDecKindMask) == kDecKindByte;
          if ((info & kDecBadByte) && !err) { err = 1; err_id = id; }
        }
      }
      stop = __any_sync(0xFFFFFFFFu, err != 0);
    }
    if (stop) {
      if (err) {
        atomicMax(D.status + 1, err);
        atomicExch(D.status + 3, static_cast<uint32_t>(err_id));
      }
      if (lane == 0) { D.sent_start[sent] = 0; D.sent_count[sent] = 0; }
      continue;
    }
    has_byte = __any_sync(0xFFFFFFFFu, has_byte);
    uint32_t total = 0;
    unsigned long long pos = 0;
    if (has_byte) {
      // ---- sequential path on lane 0 ----
      if (lane == 0) {
        total = decode_list_sequential<false>(D, a, L, nullptr);
        pos = atomicAdd(D.cursor, static_cast<unsigned long long>(total));
        if (pos + total > D.tmp_cap) { atomicOr(D.status + 2, 1u); total = 0; }
        else decode_list_sequential<true>(D, a, L, D.tmp + pos);
        D.sent_start[sent] = pos;
        D.sent_count[sent] = total;
      }
      __syncwarp();
      continue;
    }
    // ---- parallel path: one token per lane, two passes (count, write) ----
    for (int pass = 0; pass < 2; ++pass) {
      bool bos_over = !D.strip;
      uint32_t run = 0;
      for (uint32_t base = 0; base < L; base += 32) {
        const uint32_t t = base + lane;
        const bool valid = t < L;
        uint32_t info = kDecKindControl, off = 0, n0 = 0;
        if (valid) {
          const int32_t id = D.ids[a + t];
          info = __ldg(D.dec_info + id);
          off = __ldg(D.dec_off + id);
          n0 = __ldg(D.dec_off + id + 1) - off;
        }
        const uint32_t kind = info & kDecKindMask;
        const bool lead = kind == kDecKindNormal && (info & kDecLeadWs);
        // would this token end the bos state if it were processed in it?
        const bool ends = valid && (kind == kDecKindUnknown ? n0 > 0
                                    : kind == kDecKindNormal ? (lead ? (!D.rm || n0 > 1) : n0 > 0)
                                                             : false);
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, ends);
        const bool in_bos = !bos_over && (mask & ((1u << lane) - 1u)) == 0u;
        const uint32_t skip = (in_bos && lead) ? 1u : 0u;
        const uint32_t len = n0 - skip;
        bos_over = bos_over || mask != 0u;
        uint32_t incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
          if (lane >= static_cast<uint32_t>(d)) incl += v;
        }
        if (pass == 1) {
          uint8_t *dst = D.tmp + pos + run + (incl - len);
          const uint8_t *src = D.dec_bytes + off + skip;
          const uint32_t mx = __reduce_max_sync(0xFFFFFFFFu, len);
          for (uint32_t k = 0; k < mx; ++k)
            if (k < len) dst[k] = __ldg(src + k);
        }
        run += __shfl_sync(0xFFFFFFFFu, incl, 31);
      }
      if (pass == 0) {
        total = run;
        if (lane == 0) {
          pos = atomicAdd(D.cursor, static_cast<unsigned long long>(total));
          if (pos + total > D.tmp_cap) atomicOr(D.status + 2, 1u);
          D.sent_start[sent] = pos;
          D.sent_count[sent] = pos + total > D.tmp_cap ? 0u : total;
        }
        pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
        if (pos + total > D.tmp_cap) break;
      }
    }
  }
}

}  // namespace spm_b200
#endif
