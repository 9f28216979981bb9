// bpe_lane_kernel.cuh -- K3 fast path: BPE merge, one sentence per LANE, one WORD at a time.
//
// Reference: bpe::Model::SampleEncode with alpha = 0 (src/bpe_model.cc:38-203): repeat
// "merge the live adjacent pair with the greatest score, leftmost on ties" (the agenda's
// order, :51-57) until no adjacent pair is a piece.
//
// Exact decomposition used here (verified there on 20,000 sentences and by
// the parity tests): when no piece contains U+2581 anywhere but at byte 0 (the default
// split_by_whitespace=true vocabulary; checked at load, kFlagBpeWordSplit), a merge can
// never join a symbol with a following "▁..." symbol, so a sentence falls apart into
// independent words (▁ + following characters).  The greedy loop only ever compares raw
// piece scores, so running it per word gives exactly the reference's ids.  Each lane
// walks its sentence word by word with tiny per-word symbol arrays in shared memory
// ([slot][lane], bank == lane), and caches the trie node of every symbol so that
// "is left+right a piece?" walks only the right symbol's bytes.
//
// Engine-side preconditions (else the general warp kernel of bpe_kernel.cuh runs):
// kFlagBpeWordSplit, escape_whitespaces, no user-defined symbols, no UNUSED pieces.
// Words with more than kBpeWordSyms symbols defer the sentence to the general path.
#ifndef SPM_B200_BPE_LANE_KERNEL_CUH_
#define SPM_B200_BPE_LANE_KERNEL_CUH_

#include "lane_kernel.cuh"

namespace spm_b200 {

constexpr uint32_t kBpeWordSyms = 24;                      // symbols of one word held in shared memory
constexpr uint32_t kBpeLaneWarpBytes = kBpeWordSyms * 32 * 4 * 3;  // sym, pn, ps
constexpr uint32_t kBpeDead = 0x3FFFFFu;                   // 22-bit node field: not a trie path / not a piece

__global__ void __launch_bounds__(768, 1) encode_bpe_lane_kernel(const KModel M, const KBatch B, uint8_t *slabs,
                                                                  uint32_t cap) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t *s_tab = reinterpret_cast<uint32_t *>(smem);
  uint8_t *arrays = smem + kLaneTableBytes;
  fill_lane_tables(M, s_tab);
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp_in_cta = threadIdx.x >> 5;
  const uint32_t warp_global = blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
  LaneCtx c;
  c.pol = slab_policy(B.slab_l2);
  uint32_t *sym, *pn;
  float *ps;
  {
    uint8_t *a = arrays + static_cast<size_t>(warp_in_cta) * kBpeLaneWarpBytes;
    sym = reinterpret_cast<uint32_t *>(a) + lane;                             // node(22) | byte_len << 22
    pn = reinterpret_cast<uint32_t *>(a + kBpeWordSyms * 32 * 4) + lane;      // pair node(22) | offset_in_word << 22
    ps = reinterpret_cast<float *>(a + kBpeWordSyms * 32 * 8) + lane;         // pair score
    uint8_t *slab = slabs + static_cast<size_t>(warp_global) * lane_slab_bytes(cap);
    c.text_w = reinterpret_cast<uint32_t *>(slab) + lane;
    c.log = reinterpret_cast<uint32_t *>(slab) + static_cast<size_t>(cap / 4 + kLaneTextSlack) * 32 + lane;
    c.rs = nullptr;
    c.rb = nullptr;
    c.s_lead = s_tab;
    c.s_pair = s_tab + 8;
    c.s_solo = reinterpret_cast<const int32_t *>(s_tab + 8 + 1024);
    c.s_plain = s_tab + 8 + 1024 + 128;
    c.s_plainsp = c.s_plain + 4;
  }
  const uint2 *node2 = M.trie_node2;
  const uint32_t root = __ldg(&node2[0]).x;
  const bool bf = M.flags & kFlagByteFallback;
  // 32-byte register window over the lane's text (words bw .. bw+7): a word's bytes are read many
  // times (character split, every pair evaluation), so they must not go back to L2 each time
  uint32_t bw = 0xFFFFFFF0u, t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, t7 = 0;
  auto load_window = [&](uint32_t k) {
    bw = k >> 2;
    const uint32_t *q = c.text_w + static_cast<size_t>(bw) * 32;
    t0 = q[0]; t1 = q[32]; t2 = q[64]; t3 = q[96];
    t4 = q[128]; t5 = q[160]; t6 = q[192]; t7 = q[224];
  };
  auto text_byte = [&](uint32_t k) -> uint32_t {
    const uint32_t d = (k >> 2) - bw;
    uint32_t w;
    if (d < 8u) {
      const uint32_t lo = (d & 2u) ? ((d & 1u) ? t3 : t2) : ((d & 1u) ? t1 : t0);
      const uint32_t hi = (d & 2u) ? ((d & 1u) ? t7 : t6) : ((d & 1u) ? t5 : t4);
      w = (d & 4u) ? hi : lo;
    } else {
      w = c.text_w[static_cast<size_t>(k >> 2) * 32];
    }
    return (w >> ((k & 3u) * 8u)) & 0xFFu;
  };
  // walks `len` bytes at text offset `off` from link word `l`; returns the node reached (and its
  // link word) or kBpeDead
  auto walk = [&](uint32_t l, uint32_t off, uint32_t len, uint32_t *link_out) -> uint32_t {
    uint32_t v = kBpeDead;
    for (uint32_t i = 0; i < len; ++i) {
      const uint32_t ch = text_byte(off + i);
      v = (l >> kLinkBaseShift) ^ ch;
      l = __ldg(&node2[v]).x;
      if ((l & kLinkLabelMask) != ch) return kBpeDead;
    }
    *link_out = l;
    return v;
  };

  for (;;) {
    uint32_t first = 0;
    if (lane == 0) first = atomicAdd(B.work_counter, 32u);
    first = __shfl_sync(0xFFFFFFFFu, first, 0);
    if (first >= B.n) break;
    lane_wait_input(B, first, lane);
    const bool have = first + lane < B.n;
    const uint32_t sent = have && B.order ? B.order[first + lane] : first + lane;
    // ---------------- K1 ----------------
    uint32_t n = 0;
    bool defer = false;
    if (have) {
      const unsigned long long off = B.offsets[sent];
      const unsigned long long len64 = B.offsets[sent + 1] - off;
      if (len64 > 4ull * cap || off < B.off_lo || off + len64 > B.off_hi) defer = true;
      else {
        n = lane_normalize(M, B.bytes + off, static_cast<uint32_t>(len64), c, cap);
        if (n == 0xFFFFFFFFu) { defer = true; n = 0; }
      }
    }
    __syncwarp();
    bw = 0xFFFFFFF0u;  // the register window still holds the lane's previous sentence
    // ---------------- K3: word by word ----------------
    uint32_t nlog = 0;  // symbols emitted to the log: id(24) | byte_len << 24
    uint32_t p = 0;     // text position of the next word
    while (p < n && !defer) {
      // -- split the word into characters (bpe_model.cc:110-120) and cache their trie nodes --
      uint32_t m = 0, q = p;
      bool first_sym = true;
      if ((p >> 2) - bw >= 4u) load_window(p);  // keep at least 16 bytes of the word in registers
      while (q < n) {
        const uint32_t b0 = text_byte(q);
        uint32_t l = one_char_len(b0);
        if (l > n - q) l = n - q;
        // a new word starts at every U+2581 symbol
        if (!first_sym && l == 3 && b0 == 0xE2u && text_byte(q + 1) == 0x96u && text_byte(q + 2) == 0x81u) break;
        if (m == kBpeWordSyms) { defer = true; break; }
        uint32_t lk = 0;
        const uint32_t node = walk(root, q, l, &lk);
        sym[m * 32] = node | (l << 22);
        pn[m * 32] = kBpeDead | ((q - p) << 22);
        ++m;
        q += l;
        first_sym = false;
      }
      if (defer) break;
      // MaybeAddNewSymbolPair (bpe_model.cc:83-107) for the pair (i, i+1)
      auto eval_pair = [&](uint32_t i) {
        const uint32_t si = sym[i * 32], sj = sym[(i + 1) * 32];
        const uint32_t offj = pn[(i + 1) * 32] >> 22;
        uint32_t res = kBpeDead;
        float score = 0.f;
        if ((si & 0x3FFFFFu) != kBpeDead) {
          uint32_t lk = 0;
          const uint32_t v = walk(__ldg(&node2[si & 0x3FFFFFu]).x, p + offj, sj >> 22, &lk);
          if (v != kBpeDead && ((lk >> kLinkKindShift) & 3u) != kKindNone) {
            res = v;
            score = __uint_as_float(__ldg(M.trie_val + v));
          }
        }
        pn[i * 32] = res | (pn[i * 32] & 0xFFC00000u);
        ps[i * 32] = score;
      };
      for (uint32_t i = 0; i + 1 < m; ++i) eval_pair(i);
      // -- greedy merges: best score, leftmost on ties (bpe_model.cc:51-57,141-173) --
      for (;;) {
        int bi = -1;
        float best = 0.f;
        for (uint32_t i = 0; i + 1 < m; ++i) {
          if ((pn[i * 32] & 0x3FFFFFu) != kBpeDead) {
            const float sc = ps[i * 32];
            if (bi < 0 || sc > best) { best = sc; bi = static_cast<int>(i); }
          }
        }
        if (bi < 0) break;
        const uint32_t i = static_cast<uint32_t>(bi);
        const uint32_t nl = (sym[i * 32] >> 22) + (sym[(i + 1) * 32] >> 22);
        sym[i * 32] = (pn[i * 32] & 0x3FFFFFu) | (nl << 22);
        for (uint32_t j = i + 1; j + 1 < m; ++j) {  // close the gap
          sym[j * 32] = sym[(j + 1) * 32];
          pn[j * 32] = pn[(j + 1) * 32];
          ps[j * 32] = ps[(j + 1) * 32];
        }
        --m;
        if (i > 0) eval_pair(i - 1);
        if (i + 1 < m) eval_pair(i);
        else pn[i * 32] = kBpeDead | (pn[i * 32] & 0xFFC00000u);
      }
      // -- emit the word's symbols: PieceToId (model_interface.cc:51-61) --
      for (uint32_t i = 0; i < m; ++i) {
        const uint32_t s = sym[i * 32];
        int32_t id = M.unk_id;
        if ((s & 0x3FFFFFu) != kBpeDead) {
          const int32_t t = __ldg(M.trie_id + (s & 0x3FFFFFu));
          if (t >= 0) id = t;
        }
        c.log[static_cast<size_t>(nlog) * 32] = static_cast<uint32_t>(id) | ((s >> 22) << 24);
        ++nlog;
      }
      p = q;
    }
    if (have && defer) {
      const uint32_t slot = atomicAdd(B.status, 1u);
      B.deferred[2 * slot] = sent;
      B.deferred[2 * slot + 1] = 0;
      B.sent_count[sent] = 0;  // until a later pass encodes it
      nlog = 0;
    }
    // ---------------- K4: id path of PopulateSentencePieceText over the symbol log ----------------
    const uint32_t unk = static_cast<uint32_t>(M.unk_id);
    const uint32_t max_log = __reduce_max_sync(0xFFFFFFFFu, nlog);
    uint32_t count = 0;
    {
      bool prev_unk = false;
      for (uint32_t t = 0; t < max_log; ++t) {
        if (t < nlog) {
          const uint32_t e = c.log[static_cast<size_t>(t) * 32];
          const bool isunk = (e & 0xFFFFFFu) == unk;
          if (bf) count += isunk ? (e >> 24) : 1u;
          else count += !(isunk && prev_unk);
          prev_unk = isunk;
        }
      }
    }
    uint32_t incl = count;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= static_cast<uint32_t>(d)) incl += t;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    unsigned long long pos = 0;
    if (lane == 0 && total) {
      pos = atomicAdd(B.cursor, static_cast<unsigned long long>(total));
      if (pos + total > B.tmp_cap) atomicOr(B.status + 2, 1u);
    }
    pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
    const bool room = pos + total <= B.tmp_cap;
    pos += incl - count;
    if (have && !defer) {
      B.sent_start[sent] = pos;
      B.sent_count[sent] = room ? count : 0u;
    }
    if (room) {
      bool prev_unk = false;
      uint32_t w = 0, off = 0;
      for (uint32_t t = 0; t < max_log; ++t) {
        if (t < nlog) {
          const uint32_t e = c.log[static_cast<size_t>(t) * 32];
          const uint32_t plen = e >> 24;
          const bool isunk = (e & 0xFFFFFFu) == unk;
          if (isunk) {
            if (bf) {
              for (uint32_t i = 0; i < plen; ++i) B.tmp_ids[pos + (w++)] = __ldg(M.byte_to_id + text_byte(off + i));
            } else if (!prev_unk) {
              B.tmp_ids[pos + (w++)] = M.unk_id;
            }
          } else {
            B.tmp_ids[pos + (w++)] = static_cast<int32_t>(e & 0xFFFFFFu);
          }
          prev_unk = isunk;
          off += plen;
        }
      }
    }
    lane_drain(B, sent, have, lane);  // K6 (fused host path only)
    __syncwarp();
  }
}

}  // namespace spm_b200
#endif
