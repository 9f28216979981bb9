// nbest_kernel.cuh -- K5: exact n-best segmentation (config 5: SampleEncode nbest_size > 1,
// NBestEncode) on the GPU, one sentence per lane.
//
// Reference: unigram::Model::NBestEncode (src/unigram_model.cc:695-721) =
//   Lattice::SetSentence (:113-146)  character positions, BOS / EOS nodes
//   Model::PopulateNodes (:547-596)  one node per (start char, piece), UNK node last
//   Lattice::Viterbi     (:161-198)  float backtrace scores (the A* heuristic)
//   Lattice::NBest       (:345-509)  backward A* over a std::priority_queue<Hypothesis*>
// The ORDER of equal-priority results is whatever libstdc++'s binary heap yields
// (37 % of sentences have tied candidates), so the heap discipline of
// std::push_heap / std::pop_heap (bits/stl_heap.h: sift-up stops at equal keys;
// __adjust_heap walks the hole to a leaf preferring the right child unless it is smaller,
// then pushes the saved last element back up) and the agenda shrink at 10,000 entries are
// reproduced step for step.  The search is sequential per sentence, so each lane runs it
// for its own sentence with its lattice, hypothesis pool and heap in a per-lane slab in HBM.
#ifndef SPM_B200_NBEST_KERNEL_CUH_
#define SPM_B200_NBEST_KERNEL_CUH_

#include "lane_kernel.cuh"

namespace spm_b200 {

struct NbestGeom {
  uint32_t cap;        // normalized bytes per sentence
  uint32_t node_cap;   // lattice nodes per sentence
  uint32_t hyp_cap;    // hypotheses per sentence
  uint32_t heap_cap;   // agenda entries (>= 10000 + fan-out)
};
__host__ __device__ inline unsigned long long nbest_lane_bytes(const NbestGeom &g) {
  unsigned long long b = 0;
  b += 2ull * (g.cap + 4);                       // surf u16
  b += 4ull * (g.cap + 4) * 2;                   // begin_off, end_off u32
  b += g.node_cap * (2ull * 4 + 4ull * 3 + 4);   // npos,nend,nbb,nbe u16; nid,nscore,nbt; end_list u32
  b += g.hyp_cap * (4ull * 3 + 2);               // next u32, fx, gx, node u16
  b += 8ull * g.heap_cap;                        // agenda entries {fx bits, hypothesis}
  return (b + 15ull) & ~15ull;
}

struct NbestOut {
  int32_t *tmp_ids;
  unsigned long long tmp_cap;
  unsigned long long *cursor;
  unsigned long long *cand_start;  // [n * nbest]
  uint32_t *cand_count;            // [n * nbest]
  float *cand_score;               // [n * nbest]
  uint32_t *n_cands;               // [n]
  uint32_t *status;                // [1] error, [2] overflow, [3] capacity exceeded (unsupported)
};

__global__ void __launch_bounds__(512, 1) nbest_lane_kernel(const KModel M, const KBatch B, const NbestOut O,
                                                             uint8_t *text_slabs, uint8_t *scratch, const NbestGeom G,
                                                             uint32_t nbest) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t *s_tab = reinterpret_cast<uint32_t *>(smem);
  for (uint32_t i = threadIdx.x; i < kLaneTableBytes / 4; i += blockDim.x) {
    uint32_t v;
    if (i < 8) v = M.cm_lead[i];
    else if (i < 8 + 1024) v = M.cm_pair[i - 8];
    else if (i < 8 + 1024 + 128) v = static_cast<uint32_t>(M.cm_solo[i - 8 - 1024]);
    else {
      const uint32_t wq = i - (8 + 1024 + 128);
      v = ~((M.flags & kFlagHasCharsmap) ? M.cm_lead[wq] : 0u);
      if (wq == 1) v &= ~1u;
    }
    s_tab[i] = v;
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  LaneCtx c;
  {
    uint8_t *slab = text_slabs + static_cast<size_t>(warp_global) * lane_slab_bytes(G.cap);
    c.text_w = reinterpret_cast<uint32_t *>(slab) + lane;
    c.log = nullptr; c.rs = nullptr; c.rb = nullptr;
    c.s_lead = s_tab; c.s_pair = s_tab + 8;
    c.s_solo = reinterpret_cast<const int32_t *>(s_tab + 8 + 1024);
    c.s_plain = s_tab + 8 + 1024 + 128;
  }
  // per-lane scratch
  uint8_t *sp = scratch + (static_cast<size_t>(warp_global) * 32 + lane) * nbest_lane_bytes(G);
  uint32_t *begin_off = reinterpret_cast<uint32_t *>(sp); sp += 4ull * (G.cap + 4);
  uint32_t *end_off = reinterpret_cast<uint32_t *>(sp); sp += 4ull * (G.cap + 4);
  int32_t *nid = reinterpret_cast<int32_t *>(sp); sp += 4ull * G.node_cap;
  float *nscore = reinterpret_cast<float *>(sp); sp += 4ull * G.node_cap;
  float *nbt = reinterpret_cast<float *>(sp); sp += 4ull * G.node_cap;
  uint32_t *end_list = reinterpret_cast<uint32_t *>(sp); sp += 4ull * G.node_cap;
  uint32_t *hnext = reinterpret_cast<uint32_t *>(sp); sp += 4ull * G.hyp_cap;   // 0xFFFFFFFF = null
  float *hfx = reinterpret_cast<float *>(sp); sp += 4ull * G.hyp_cap;
  float *hgx = reinterpret_cast<float *>(sp); sp += 4ull * G.hyp_cap;
  uint2 *heap = reinterpret_cast<uint2 *>(sp); sp += 8ull * G.heap_cap;  // {fx bits, hypothesis index}: one load per level
  uint16_t *surf = reinterpret_cast<uint16_t *>(sp); sp += 2ull * (G.cap + 4);
  uint16_t *npos = reinterpret_cast<uint16_t *>(sp); sp += 2ull * G.node_cap;
  uint16_t *nend = reinterpret_cast<uint16_t *>(sp); sp += 2ull * G.node_cap;
  uint16_t *nbb = reinterpret_cast<uint16_t *>(sp); sp += 2ull * G.node_cap;
  uint16_t *nbe = reinterpret_cast<uint16_t *>(sp); sp += 2ull * G.node_cap;
  uint16_t *hnode = reinterpret_cast<uint16_t *>(sp);

  const uint2 *node2 = M.trie_node2;
  const uint32_t root = __ldg(&node2[0]).x;
  const bool bf = M.flags & kFlagByteFallback;
  auto text_byte = [&](uint32_t k) -> uint32_t {
    return (c.text_w[static_cast<size_t>(k >> 2) * 32] >> ((k & 3u) * 8u)) & 0xFFu;
  };
  // std::push_heap / std::pop_heap on (fx, hypothesis) entries keyed by fx (comp: a.fx < b.fx)
  auto heap_push = [&](uint32_t &hn, uint32_t v) {
    uint32_t hole = hn++;
    const float fv = hfx[v];
    while (hole > 0) {
      const uint32_t parent = (hole - 1) >> 1;
      const uint2 pe = heap[parent];
      if (!(__uint_as_float(pe.x) < fv)) break;  // equal keys do not move up
      heap[hole] = pe;
      hole = parent;
    }
    heap[hole] = make_uint2(__float_as_uint(fv), v);
  };
  auto heap_pop = [&](uint32_t &hn) -> uint32_t {
    const uint32_t top = heap[0].y;
    const uint32_t len = --hn;
    if (len == 0) return top;
    const uint2 value = heap[len];
    const float fv = __uint_as_float(value.x);
    uint32_t hole = 0, child = 0;
    while (child < (len - 1) / 2) {  // __adjust_heap
      child = 2 * (child + 1);
      uint2 ce = heap[child];
      const uint2 le = heap[child - 1];
      if (__uint_as_float(ce.x) < __uint_as_float(le.x)) { child--; ce = le; }
      heap[hole] = ce;
      hole = child;
    }
    if ((len & 1u) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      heap[hole] = heap[child - 1];
      hole = child - 1;
    }
    while (hole > 0) {  // __push_heap of the saved last element
      const uint32_t parent = (hole - 1) >> 1;
      const uint2 pe = heap[parent];
      if (!(__uint_as_float(pe.x) < fv)) break;
      heap[hole] = pe;
      hole = parent;
    }
    heap[hole] = value;
    return top;
  };

  for (;;) {
    uint32_t first = 0;
    if (lane == 0) first = atomicAdd(B.work_counter, 32u);
    first = __shfl_sync(0xFFFFFFFFu, first, 0);
    if (first >= B.n) break;
    const uint32_t sent = first + lane;
    if (sent < B.n) {
      const unsigned long long off = B.offsets[sent];
      const unsigned long long len64 = B.offsets[sent + 1] - off;
      uint32_t n = 0;
      bool too_big = len64 > 4ull * G.cap;
      if (!too_big) {
        n = lane_normalize(M, B.bytes + off, static_cast<uint32_t>(len64), c, G.cap);
        if (n == 0xFFFFFFFFu) { too_big = true; n = 0; }
      }
      const size_t cbase = static_cast<size_t>(sent) * nbest;
      uint32_t K = 0;
      if (too_big) {
        atomicOr(O.status + 3, 1u);
      } else if (n == 0) {
        // NBestEncode of an empty normalized string: one empty candidate, score 0 (:697-699)
        O.cand_start[cbase] = 0; O.cand_count[cbase] = 0; O.cand_score[cbase] = 0.f;
        K = 1;
      } else {
        // ---- Lattice::SetSentence ----
        uint32_t L = 0;
        for (uint32_t p = 0; p < n;) {
          uint32_t mb = one_char_len(text_byte(p));
          if (mb > n - p) mb = n - p;
          surf[L++] = static_cast<uint16_t>(p);
          p += mb;
        }
        surf[L] = static_cast<uint16_t>(n);
        // nodes 0 = BOS, 1 = EOS
        uint32_t nn = 2;
        npos[0] = 0; nend[0] = 0; nid[0] = -1; nscore[0] = 0.f; nbt[0] = 0.f; nbb[0] = 0; nbe[0] = 0;
        npos[1] = static_cast<uint16_t>(L); nend[1] = static_cast<uint16_t>(L); nid[1] = -1; nscore[1] = 0.f; nbt[1] = 0.f;
        nbb[1] = static_cast<uint16_t>(n); nbe[1] = static_cast<uint16_t>(n);
        bool overflow = false;
        // ---- Model::PopulateNodes ----
        for (uint32_t bp = 0; bp < L && !overflow; ++bp) {
          begin_off[bp] = nn;
          bool has_single = false;
          uint32_t l = root;
          uint32_t clen = 0;  // characters completed so far
          for (uint32_t kpos = surf[bp]; kpos < n; ++kpos) {
            const uint32_t ch = text_byte(kpos);
            const uint32_t v = (l >> kLinkBaseShift) ^ ch;
            l = __ldg(&node2[v]).x;
            if ((l & kLinkLabelMask) != ch) break;
            if (kpos + 1 == surf[bp + clen + 1]) ++clen;
            const uint32_t kind = (l >> kLinkKindShift) & 3u;
            if (kind == kKindNone || kind == kKindUnused) continue;
            // get_chars_length (:548-552): characters whose start lies before the piece's end
            const uint32_t length = (kpos + 1 == surf[bp + clen]) ? clen : clen + 1;
            if (nn >= G.node_cap) { overflow = true; break; }
            npos[nn] = static_cast<uint16_t>(bp);
            nend[nn] = static_cast<uint16_t>(bp + length);
            nid[nn] = __ldg(M.trie_id + v);
            nscore[nn] = kind == kKindUserDefined
                             ? static_cast<float>(static_cast<double>(__fmul_rn(static_cast<float>(length), M.max_score)) - 0.1)
                             : __uint_as_float(__ldg(M.trie_val + v));
            nbt[nn] = 0.f;
            nbb[nn] = surf[bp];
            nbe[nn] = surf[bp + length];
            ++nn;
            has_single |= length == 1;
          }
          if (!has_single && !overflow) {
            if (nn >= G.node_cap) { overflow = true; break; }
            npos[nn] = static_cast<uint16_t>(bp); nend[nn] = static_cast<uint16_t>(bp + 1);
            nid[nn] = M.unk_id; nscore[nn] = M.unk_score; nbt[nn] = 0.f;
            nbb[nn] = surf[bp]; nbe[nn] = surf[bp + 1];
            ++nn;
          }
        }
        begin_off[L] = nn;
        // end_nodes lists (insertion order): BOS first at position 0
        if (!overflow) {
          for (uint32_t p = 0; p <= L + 1; ++p) end_off[p] = 0;
          end_off[0 + 1] += 1;  // BOS
          for (uint32_t i = 2; i < nn; ++i) end_off[nend[i] + 1] += 1;
          for (uint32_t p = 0; p <= L; ++p) end_off[p + 1] += end_off[p];
          // fill using a running cursor kept in begin positions of end_list (stable)
          // (second pass: place each node at the next free slot of its end position)
          // use nbt[] as scratch? no: keep a cursor array in the tail of `heap` (unused until A*)
          for (uint32_t p = 0; p <= L; ++p) heap[p].x = end_off[p];
          end_list[heap[0].x++] = 0;
          for (uint32_t i = 2; i < nn; ++i) end_list[heap[nend[i]].x++] = i;
          // ---- Lattice::Viterbi: backtrace scores ----
          for (uint32_t pos = 0; pos <= L; ++pos) {
            const uint32_t rb = pos < L ? begin_off[pos] : 1u, re = pos < L ? begin_off[pos + 1] : 2u;
            for (uint32_t r = rb; r < re; ++r) {
              float best = 0.f;
              bool have_best = false;
              for (uint32_t q = end_off[pos]; q < end_off[pos + 1]; ++q) {
                const float sc = __fadd_rn(nbt[end_list[q]], nscore[r]);
                if (!have_best || sc > best) { best = sc; have_best = true; }
              }
              nbt[r] = best;
            }
          }
          // ---- Lattice::NBest: backward A* ----
          uint32_t pn = 0, hn = 0;
          hnode[0] = 1; hnext[0] = 0xFFFFFFFFu;  // EOS, next = null
          hgx[0] = 0.f;
          hfx[0] = nbt[1];
          pn = 1;
          heap_push(hn, 0);
          const uint32_t shrink_to = nbest * 10 < 512 ? nbest * 10 : 512;
          while (hn && !overflow) {
            const uint32_t top = heap_pop(hn);
            const uint32_t node = hnode[top];
            if (node == 0) {  // reached BOS: one result
              // pass 1: count ids
              uint32_t cnt = 0;
              bool prev_unk = false;
              for (uint32_t h = hnext[top]; hnext[h] != 0xFFFFFFFFu; h = hnext[h]) {
                const uint32_t nd = hnode[h];
                const bool isunk = nid[nd] == M.unk_id;
                if (bf) cnt += isunk ? static_cast<uint32_t>(nbe[nd] - nbb[nd]) : 1u;
                else cnt += !(isunk && prev_unk);
                prev_unk = isunk;
              }
              const unsigned long long pos = atomicAdd(O.cursor, static_cast<unsigned long long>(cnt));
              O.cand_start[cbase + K] = pos;
              O.cand_count[cbase + K] = cnt;
              O.cand_score[cbase + K] = hfx[top];
              if (pos + cnt > O.tmp_cap) {
                atomicOr(O.status + 2, 1u);
                O.cand_count[cbase + K] = 0;
              } else {
                uint32_t w = 0;
                prev_unk = false;
                for (uint32_t h = hnext[top]; hnext[h] != 0xFFFFFFFFu; h = hnext[h]) {
                  const uint32_t nd = hnode[h];
                  const bool isunk = nid[nd] == M.unk_id;
                  if (isunk) {
                    if (bf) {
                      for (uint32_t k = nbb[nd]; k < nbe[nd]; ++k) O.tmp_ids[pos + (w++)] = __ldg(M.byte_to_id + text_byte(k));
                    } else if (!prev_unk) {
                      O.tmp_ids[pos + (w++)] = M.unk_id;
                    }
                  } else {
                    O.tmp_ids[pos + (w++)] = nid[nd];
                  }
                  prev_unk = isunk;
                }
              }
              if (++K == nbest) break;
              continue;
            }
            // expand: one hypothesis per node ending where `node` begins, in end_nodes order
            const uint32_t p0 = npos[node];
            const float top_gx = hgx[top];
            for (uint32_t q = end_off[p0]; q < end_off[p0 + 1]; ++q) {
              const uint32_t ln = end_list[q];
              if (pn >= G.hyp_cap || hn + 1 >= G.heap_cap - 512u) { overflow = true; break; }
              hnode[pn] = static_cast<uint16_t>(ln);
              hnext[pn] = top;
              hgx[pn] = __fadd_rn(nscore[ln], top_gx);
              hfx[pn] = __fadd_rn(nbt[ln], top_gx);
              heap_push(hn, pn);
              ++pn;
            }
            if (hn >= 10000u && !overflow) {  // agenda shrink (:481-505): keep the best `shrink_to`
              // popped in descending order and re-pushed in that order: the heap array becomes that list
              // (stash them in the unused tail of the hypothesis fx array? no -- use the heap's own tail)
              uint2 *keep = heap + (G.heap_cap - shrink_to);
              for (uint32_t i = 0; i < shrink_to; ++i) keep[i].y = heap_pop(hn);
              hn = 0;
              for (uint32_t i = 0; i < shrink_to; ++i) heap_push(hn, keep[i].y);
            }
          }
        }
        if (overflow) atomicOr(O.status + 3, 1u);
      }
      O.n_cands[sent] = K;
      for (uint32_t k = K; k < nbest; ++k) { O.cand_count[cbase + k] = 0; O.cand_start[cbase + k] = 0; O.cand_score[cbase + k] = 0.f; }
    }
    __syncwarp();
  }
}

// picked candidate per sentence -> (start, count) for the shared scan + gather
__global__ void __launch_bounds__(256) pick_candidates_kernel(const uint32_t *picks, uint32_t n, uint32_t nbest,
                                                              const unsigned long long *cand_start,
                                                              const uint32_t *cand_count, unsigned long long *sent_start,
                                                              uint32_t *sent_count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t c = static_cast<size_t>(i) * nbest + picks[i];
  sent_start[i] = cand_start[c];
  sent_count[i] = cand_count[c];
}

}  // namespace spm_b200
#endif
