// unigram_warp.cuh -- the fast unigram path: one sentence per warp, the Viterbi
// window held in REGISTERS and folded with warp shuffles.
//
// Reference semantics: unigram::Model::EncodeOptimized (src/unigram_model.cc:889-1020).
//
// Why this shape (measured on B200, profiles/r01_*): the path is a chain of dependent
// table lookups, so throughput = (independent chains in flight) / (lookup latency).
//   * 32 lanes walk the piece trie from 32 consecutive byte positions at once
//     (32 independent chains per warp, hot trie prefix in shared memory);
//   * best_path_ends_at[] for the 64 positions [w, w+64) lives in two registers per
//     lane (position w+lane and w+32+lane).  Relaxing an edge (s -> e) is
//       base = shfl(best0, s-w);  cand = score + base;  lane (e-w)&31 keeps the better
//     -- no shared-memory round trip on the serial dependency chain, and every
//     instruction is warp-uniform (no divergent tiles);
//   * only the back-pointers (piece length + trie unit, one u32 per byte position)
//     are spilled to shared memory for the back-trace.
// Edges are relaxed in the reference's order (starts ascending), with the reference's
// mixed float/double comparison (Q1/Q2 of SURVEY.md 8a), so ids are bit-identical.
//
// Requires pieces of at most 32 bytes (so e - w < 64); the engine falls back to the
// general tile kernel (kernels.cuh) for models with longer pieces, for the spans API
// and for sentences that do not fit the per-warp shared memory.
#ifndef SPM_B200_UNIGRAM_WARP_CUH_
#define SPM_B200_UNIGRAM_WARP_CUH_

#include "kernels.cuh"

namespace spm_b200 {

constexpr uint32_t kBpUnkIdx = 0xFFFFFFu;  // low 24 bits of a back-pointer: UNK piece

struct WarpMem {
  uint8_t *text;   // [ncap + 16]
  uint32_t *bp;    // [ncap + 4]  (piece_len << 24) | trie_unit ; 0 = unset
  uint32_t *mval;  // [32 * K]
  uint32_t *mli;   // [32 * K]    (piece_len << 24) | trie_unit
  uint8_t *stage;  // input staging, aliases bp/mval/mli
  uint32_t ncap, stage_cap;
};

__host__ __device__ inline uint32_t warp_bytes_for(uint32_t ncap, uint32_t K) {
  return ((ncap + 16) + 4 * (ncap + 4) + 8 * 32 * K + 15u) & ~15u;
}

__device__ __forceinline__ WarpMem carve_warp(uint8_t *base, uint32_t ncap, uint32_t K) {
  WarpMem m;
  uint8_t *p = base;
  m.bp = reinterpret_cast<uint32_t *>(p); p += 4 * (ncap + 4);
  m.mval = reinterpret_cast<uint32_t *>(p); p += 4 * 32 * K;
  m.mli = reinterpret_cast<uint32_t *>(p); p += 4 * 32 * K;
  m.stage = base;
  m.stage_cap = 4 * (ncap + 4) + 8 * 32 * K;
  m.text = p;
  m.ncap = ncap;
  return m;
}

// Output-space allocator: a warp claims chunks of the temporary id buffer so that the
// global cursor is touched once per ~1000 ids instead of once per sentence.
struct OutChunk {
  unsigned long long pos;
  uint32_t left;
};
constexpr uint32_t kOutChunkIds = 1024;

// K2 for one warp.  Fills wm.bp[1..n].
__device__ __forceinline__ void viterbi_warp(const KModel &M, const HotTrie &H, const WarpMem &wm, uint32_t n,
                                             uint32_t lane) {
  const uint8_t *text = wm.text;
  const uint32_t K = M.match_slots;
  float best0 = 0.f, best1 = 0.f;  // best_path_score of positions w+lane, w+32+lane
  uint32_t bp0 = 0, bp1 = 0;       // their back-pointers; 0 = starts_at == -1
  const uint32_t root = H.link(0);
  uint32_t *mval = wm.mval + lane * K;
  uint32_t *mli = wm.mli + lane * K;
  uint32_t w = 0;
  for (; w < n; w += 32) {
    // ---- phase A: 32 independent trie walks (unigram_model.cc:966-994) ----
    const uint32_t s = w + lane;
    uint32_t cnt = 0;
    bool active = false;
    if (s < n) {
      const uint32_t lead = text[s];
      active = !is_trail(lead);
      if (active) {
        uint32_t mblen = one_char_len(lead);
        if (mblen > n - s) mblen = n - s;
        bool has_single = false;
        uint32_t l = root;
        for (uint32_t k = s; k < n; ++k) {
          const uint32_t c = text[k];
          const uint32_t v = (l >> kLinkBaseShift) ^ c;
          l = H.link(v);
          if ((l & kLinkLabelMask) != c) break;
          const uint32_t kind = (l >> kLinkKindShift) & 3u;
          if (kind == kKindNormal || kind == kKindUserDefined) {
            const uint32_t plen = k + 1 - s;
            if (cnt < K) {
              mval[cnt] = kind == kKindNormal ? H.val(v) : kValUserDefined;
              mli[cnt] = (plen << 24) | v;
            }
            ++cnt;
            has_single |= plen == mblen;
          }
        }
        if (!has_single) {  // UNK edge, :995-1005
          if (cnt < K) { mval[cnt] = kValUnk; mli[cnt] = (mblen << 24) | kBpUnkIdx; }
          ++cnt;
        }
      }
    }
    __syncwarp();
    // ---- phase B: ordered fold in registers ----
    uint32_t amask = __ballot_sync(0xFFFFFFFFu, active);
    while (amask) {
      const int j = __ffs(amask) - 1;
      amask &= amask - 1;
      const float till_here = __shfl_sync(0xFFFFFFFFu, best0, j);
      const uint32_t cj = __shfl_sync(0xFFFFFFFFu, cnt, j);
      const uint32_t *jv = wm.mval + j * K;
      const uint32_t *jl = wm.mli + j * K;
      for (uint32_t m = 0; m < cj; ++m) {
        const uint32_t val = jv[m];  // broadcast reads
        const uint32_t li = jl[m];
        const uint32_t plen = li >> 24;
        const uint32_t erel = static_cast<uint32_t>(j) + plen;  // 1..63
        const bool hi = erel >= 32u;
        const float cur = hi ? best1 : best0;
        const bool unset = (hi ? bp1 : bp0) == 0u;
        float ns;
        bool better;
        if (val == kValUnk) {
          ns = __fadd_rn(M.unk_score, till_here);
          better = unset || ns > cur;
        } else {
          const double sc = val == kValUserDefined
                                ? static_cast<double>(__fmul_rn(static_cast<float>(plen), M.max_score)) - 0.1
                                : static_cast<double>(__uint_as_float(val));
          const double cand = sc + static_cast<double>(till_here);
          better = unset || cand > static_cast<double>(cur);
          ns = static_cast<float>(cand);
        }
        if (better && lane == (erel & 31u)) {
          if (hi) { best1 = ns; bp1 = li; } else { best0 = ns; bp0 = li; }
        }
      }
    }
    // positions [w, w+32) are final: spill their back-pointers, slide the window
    if (w + lane <= n) wm.bp[w + lane] = bp0;
    best0 = best1; bp0 = bp1;
    best1 = 0.f; bp1 = 0u;
    __syncwarp();
  }
  if (lane == 0 && w == n) wm.bp[n] = bp0;  // n is a multiple of 32: its entry was still in the window
  __syncwarp();
}

// Back-trace + id path of PopulateSentencePieceText for one warp.
__device__ __forceinline__ void finish_warp(const KModel &M, const KBatch &B, const WarpMem &wm, uint32_t sent,
                                            uint32_t n, uint32_t lane, OutChunk *oc) {
  // back-trace (unigram_model.cc:1010-1018): lane 0 compacts the path's back-pointers
  // to the top of bp[] (slot n - t for the t-th piece from the end; slot >= position).
  uint32_t n_tok = 0;
  if (lane == 0) {
    uint32_t e = n;
    while (e > 0) {
      const uint32_t v = wm.bp[e];
      const uint32_t bl = v >> 24;
      if (bl == 0 || bl > e) { atomicOr(B.status + 1, 1u); n_tok = 0; break; }
      wm.bp[n - n_tok] = v;
      e -= bl;
      ++n_tok;
    }
  }
  n_tok = __shfl_sync(0xFFFFFFFFu, n_tok, 0);
  __syncwarp();
  const uint32_t *tok = wm.bp + (n - n_tok + 1);
  const bool bf = M.flags & kFlagByteFallback;
  // pass 1: number of output ids
  uint32_t count = 0;
  {
    uint32_t carry_unk = 0;
    for (uint32_t k0 = 0; k0 < n_tok; k0 += 32) {
      const uint32_t k = k0 + lane;
      const uint32_t v = k < n_tok ? tok[k] : 0u;
      const bool isunk = k < n_tok && (v & 0xFFFFFFu) == kBpUnkIdx;
      const uint32_t um = __ballot_sync(0xFFFFFFFFu, isunk);
      uint32_t c = 0;
      if (k < n_tok) {
        if (bf) c = isunk ? (v >> 24) : 1u;
        else c = !(isunk && (lane ? ((um >> (lane - 1)) & 1u) : carry_unk));
      }
      count += __reduce_add_sync(0xFFFFFFFFu, c);
      carry_unk = um >> 31;
    }
  }
  // claim output space
  if (count > oc->left) {
    unsigned long long p = 0;
    const uint32_t want = count > kOutChunkIds ? count : kOutChunkIds;
    if (lane == 0) {
      p = atomicAdd(B.cursor, static_cast<unsigned long long>(want));
      if (p + want > B.tmp_cap) atomicOr(B.status + 2, 1u);
    }
    oc->pos = __shfl_sync(0xFFFFFFFFu, p, 0);
    oc->left = want;
    if (oc->pos + want > B.tmp_cap) oc->left = 0;
  }
  const unsigned long long pos = oc->pos;
  const bool room = count <= oc->left;
  if (lane == 0) {
    B.sent_start[sent] = pos;
    B.sent_count[sent] = room ? count : 0u;
  }
  if (!room) return;  // overflow flagged: the host retries with a larger buffer
  oc->pos += count;
  oc->left -= count;
  // pass 2: write ids
  uint32_t rank_base = 0, end_base = 0, carry_unk = 0;
  for (uint32_t k0 = 0; k0 < n_tok; k0 += 32) {
    const uint32_t k = k0 + lane;
    const uint32_t v = k < n_tok ? tok[k] : 0u;
    const uint32_t plen = v >> 24;
    const uint32_t idx = v & 0xFFFFFFu;
    const bool isunk = k < n_tok && idx == kBpUnkIdx;
    const uint32_t um = __ballot_sync(0xFFFFFFFFu, isunk);
    uint32_t c = 0;
    if (k < n_tok) {
      if (bf) c = isunk ? plen : 1u;
      else c = !(isunk && (lane ? ((um >> (lane - 1)) & 1u) : carry_unk));
    }
    // inclusive scans of output counts and of piece lengths
    uint32_t ic = c, il = plen;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t tc = __shfl_up_sync(0xFFFFFFFFu, ic, d);
      const uint32_t tl = __shfl_up_sync(0xFFFFFFFFu, il, d);
      if (lane >= static_cast<uint32_t>(d)) { ic += tc; il += tl; }
    }
    if (k < n_tok) {
      const uint32_t rank = rank_base + ic - c;
      if (isunk) {
        if (bf) {
          const uint32_t start = end_base + il - plen;
          for (uint32_t i = 0; i < plen; ++i) B.tmp_ids[pos + rank + i] = __ldg(M.byte_to_id + wm.text[start + i]);
        } else if (c) {
          B.tmp_ids[pos + rank] = M.unk_id;
        }
      } else {
        B.tmp_ids[pos + rank] = __ldg(M.trie_id + idx);
      }
    }
    rank_base += __shfl_sync(0xFFFFFFFFu, ic, 31);
    end_base += __shfl_sync(0xFFFFFFFFu, il, 31);
    carry_unk = um >> 31;
  }
}

// Persistent kernel: each warp claims 4 consecutive sentences at a time.
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) encode_unigram_warp_kernel(const KModel M, const KBatch B) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
  uint32_t *s_link = reinterpret_cast<uint32_t *>(smem + 16);
  uint32_t *s_val = s_link + M.hot_link;
  uint8_t *warps = reinterpret_cast<uint8_t *>(s_val + M.hot_val);
  stage_hot_trie(M, mbar, s_link, s_val);
  HotTrie H{s_link, s_val, M.trie_link, M.trie_val, M.hot_link, M.hot_val};
  const uint32_t lane = threadIdx.x & 31;
  const Tile<32> T;
  const WarpMem wm = carve_warp(warps + static_cast<size_t>(threadIdx.x >> 5) * B.tile_bytes, B.ncap, M.match_slots);
  TileMem tm{};  // view for the shared normalizer
  tm.text = wm.text;
  tm.ncap = wm.ncap;
  OutChunk oc{0ull, 0u};
  constexpr uint32_t kClaim = 4;
  for (;;) {
    uint32_t first = 0;
    if (lane == 0) first = atomicAdd(B.work_counter, kClaim);
    first = __shfl_sync(0xFFFFFFFFu, first, 0);
    const uint32_t work_n = B.sub_list ? B.sub_n : B.n;
    if (first >= work_n) break;
    const uint32_t last = first + kClaim < work_n ? first + kClaim : work_n;
    for (uint32_t wi = first; wi < last; ++wi) {
      const uint32_t sent = B.sub_list ? B.sub_list[2 * wi] : wi;
      const unsigned long long off = B.offsets[sent];
      const unsigned long long len64 = B.offsets[sent + 1] - off;
      bool fits = len64 + 32ull <= wm.stage_cap;
      uint32_t need = 0;
      if (fits) {
        const uint32_t len = static_cast<uint32_t>(len64);
        const uint8_t *g = B.bytes + off;
        const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(g) & 15u);
        const uint4 *ga = reinterpret_cast<const uint4 *>(g - mis);
        const uint32_t nvec = (mis + len + 15u) >> 4;
        uint4 *sa = reinterpret_cast<uint4 *>(wm.stage);
        for (uint32_t v = lane; v < nvec; v += 32) sa[v] = __ldg(ga + v);
        __syncwarp();
        const NormResult nr = normalize_tile<32, false>(M, T, wm.stage + mis, len, tm);
        const uint32_t n = nr.n;
        if (n > wm.ncap) {
          fits = false;
          need = n;
        } else if (n == 0) {
          if (lane == 0) { B.sent_start[sent] = 0; B.sent_count[sent] = 0; }
        } else {
          viterbi_warp(M, H, wm, n, lane);
          finish_warp(M, B, wm, sent, n, lane, &oc);
        }
      }
      if (!fits && lane == 0) {
        const uint32_t slot = atomicAdd(B.status, 1u);
        B.deferred[2 * slot] = sent;
        B.deferred[2 * slot + 1] = need;
      }
      __syncwarp();
    }
  }
}

}  // namespace spm_b200
#endif
