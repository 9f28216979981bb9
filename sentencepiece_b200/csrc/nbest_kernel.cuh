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
  uint32_t node_cap;   // lattice nodes per sentence (<= 65535: hypotheses hold 16-bit node indices)
  uint32_t hyp_cap;    // hypotheses per sentence
  uint32_t heap_cap;   // agenda entries (>= 10000 + fan-out + 512 for the shrink list)
};
// per-lane scratch in HBM, 16-byte records so that one load / store moves a whole hypothesis or node:
//   hyp  [hyp_cap]     uint4 {next, gx bits, node | ids_so_far << 16, begin char | unk << 16}
//   heap [heap_cap+2]  uint2 {fx bits, hypothesis}; entry i lives in slot i + 1, which puts the children
//                      (2k+1, 2k+2) of any entry in one aligned 16-byte pair (one load per level of a pop)
//   node [node_cap]    uint4 {id, score bits, backtrace bits, begin char | end char << 16}   creation order
//   elist[node_cap]    uint4 {node | begin char << 16, backtrace bits, score bits, ids | unk << 31}   sorted by end char
//   end_off[cap+4] u32, maxbt[cap+4] f32 (reused as the sort cursors), surf[cap+4] u16
__host__ __device__ inline unsigned long long nbest_lane_bytes(const NbestGeom &g) {
  unsigned long long b = 0;
  b += 16ull * g.hyp_cap;
  b += 8ull * (g.heap_cap + 2);
  b += 32ull * g.node_cap;
  b += 4ull * (g.cap + 4) * 2;
  b += 2ull * (g.cap + 4);
  return (b + 15ull) & ~15ull;
}
// agenda entries kept in shared memory per lane (the top levels of the binary heap); odd, so that a pair of
// children never straddles the shared / global boundary
constexpr int kNbestTop = 15;
template <int TOP>
__host__ __device__ constexpr uint32_t nbest_smem_bytes(uint32_t warps) {
  return kLaneTableBytes + warps * 32u * 8u * TOP;
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

template <int TOP, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) nbest_lane_kernel(const KModel M, const KBatch B, const NbestOut O,
                                                                 uint8_t *text_slabs, uint8_t *scratch, const NbestGeom G,
                                                                 uint32_t nbest) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t *s_tab = reinterpret_cast<uint32_t *>(smem);
  fill_lane_tables(M, s_tab);
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  LaneCtx c;
  c.pol = slab_policy(B.slab_l2);
  {
    uint8_t *slab = text_slabs + static_cast<size_t>(warp_global) * lane_slab_bytes(G.cap);
    c.text_w = reinterpret_cast<uint32_t *>(slab) + lane;
    c.log = nullptr; c.rs = nullptr; c.rb = nullptr;
    c.s_lead = s_tab; c.s_pair = s_tab + 8;
    c.s_solo = reinterpret_cast<const int32_t *>(s_tab + 8 + 1024);
    c.s_plain = s_tab + 8 + 1024 + 128;
    c.s_plainsp = c.s_plain + 4;
  }
  // agenda top: [warp][entry][lane]
  uint2 *s_top = reinterpret_cast<uint2 *>(smem + kLaneTableBytes) + static_cast<size_t>(threadIdx.x >> 5) * (32 * TOP) + lane;
  // per-lane scratch
  uint8_t *sp = scratch + (static_cast<size_t>(warp_global) * 32 + lane) * nbest_lane_bytes(G);
  uint4 *hyp = reinterpret_cast<uint4 *>(sp); sp += 16ull * G.hyp_cap;
  uint2 *heap = reinterpret_cast<uint2 *>(sp); sp += 8ull * (G.heap_cap + 2);   // slot = entry + 1
  uint4 *node = reinterpret_cast<uint4 *>(sp); sp += 16ull * G.node_cap;
  uint4 *elist = reinterpret_cast<uint4 *>(sp); sp += 16ull * G.node_cap;
  uint32_t *end_off = reinterpret_cast<uint32_t *>(sp); sp += 4ull * (G.cap + 4);
  float *maxbt = reinterpret_cast<float *>(sp); sp += 4ull * (G.cap + 4);
  uint16_t *surf = reinterpret_cast<uint16_t *>(sp);
  uint32_t *sort_cur = reinterpret_cast<uint32_t *>(maxbt);  // maxbt is dead once the nodes exist

  const uint2 *node2 = M.trie_node2;
  const uint32_t root = __ldg(&node2[0]).x;
  const bool bf = M.flags & kFlagByteFallback;
  auto text_byte = [&](uint32_t k) -> uint32_t {
    return (c.text_w[static_cast<size_t>(k >> 2) * 32] >> ((k & 3u) * 8u)) & 0xFFu;
  };
  auto hget = [&](uint32_t i) -> uint2 { return i < TOP ? s_top[i * 32] : heap[i + 1]; };
  auto hset = [&](uint32_t i, uint2 e) {
    if (i < TOP) s_top[i * 32] = e;
    else heap[i + 1] = e;
  };
  // std::push_heap / std::pop_heap on {fx, hypothesis} entries keyed by fx (comp: a.fx < b.fx)
  auto heap_push = [&](uint32_t &hn, uint2 e) {
    uint32_t hole = hn++;
    const float fv = __uint_as_float(e.x);
    while (hole > 0) {
      const uint32_t parent = (hole - 1) >> 1;
      const uint2 pe = hget(parent);
      if (!(__uint_as_float(pe.x) < fv)) break;  // equal keys do not move up
      hset(hole, pe);
      hole = parent;
    }
    hset(hole, e);
  };
  auto heap_pop = [&](uint32_t &hn) -> uint2 {
    const uint2 top = hget(0);
    const uint32_t len = --hn;
    if (len == 0) return top;
    const uint2 value = hget(len);
    const float fv = __uint_as_float(value.x);
    uint32_t hole = 0, child = 0;
    while (child < (len - 1) / 2) {  // __adjust_heap: move the larger child up (the right one on ties)
      child = 2 * (child + 1);
      uint2 le, re;
      if (child < TOP) {
        re = s_top[child * 32];
        le = s_top[(child - 1) * 32];
      } else {
        const uint4 pr = *reinterpret_cast<const uint4 *>(heap + child);  // slots child, child + 1
        le = make_uint2(pr.x, pr.y);
        re = make_uint2(pr.z, pr.w);
      }
      if (__uint_as_float(re.x) < __uint_as_float(le.x)) { child--; re = le; }
      hset(hole, re);
      hole = child;
    }
    if ((len & 1u) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      hset(hole, hget(child - 1));
      hole = child - 1;
    }
    while (hole > 0) {  // __push_heap of the saved last element
      const uint32_t parent = (hole - 1) >> 1;
      const uint2 pe = hget(parent);
      if (!(__uint_as_float(pe.x) < fv)) break;
      hset(hole, pe);
      hole = parent;
    }
    hset(hole, value);
    return top;
  };

  for (;;) {
    uint32_t first = 0;
    if (lane == 0) first = atomicAdd(B.work_counter, 32u);
    first = __shfl_sync(0xFFFFFFFFu, first, 0);
    if (first >= B.n) break;
    if (first + lane < B.n) {
      const uint32_t sent = B.order ? B.order[first + lane] : first + lane;
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
          surf[L] = static_cast<uint16_t>(p);
          maxbt[L] = -INFINITY;
          ++L;
          p += mb;
        }
        surf[L] = static_cast<uint16_t>(n);
        maxbt[L] = -INFINITY;
        maxbt[0] = 0.f;  // BOS
        // nodes 0 = BOS, 1 = EOS (EOS's backtrace score is filled in below)
        uint32_t nn = 2;
        node[0] = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
        bool overflow = false;
        // ---- Model::PopulateNodes, with Lattice::Viterbi's backtrace scores folded in ----
        // backtrace(r) = max over nodes q ending where r begins of fl(backtrace(q) + score(r)); fl(x + s) is
        // monotone in x, so it equals fl(max_q backtrace(q) + score(r)), and every q is complete before r is made.
        for (uint32_t bp = 0; bp < L && !overflow; ++bp) {
          const float in_bt = maxbt[bp];
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
            const float sc = kind == kKindUserDefined
                                 ? static_cast<float>(static_cast<double>(__fmul_rn(static_cast<float>(length), M.max_score)) - 0.1)
                                 : __uint_as_float(__ldg(M.trie_val + v));
            const float bt = __fadd_rn(in_bt, sc);
            node[nn] = make_uint4(static_cast<uint32_t>(__ldg(M.trie_id + v)), __float_as_uint(sc), __float_as_uint(bt),
                                  bp | ((bp + length) << 16));
            maxbt[bp + length] = fmaxf(maxbt[bp + length], bt);
            ++nn;
            has_single |= length == 1;
          }
          if (!has_single && !overflow) {
            if (nn >= G.node_cap) { overflow = true; break; }
            const float bt = __fadd_rn(in_bt, M.unk_score);
            node[nn] = make_uint4(static_cast<uint32_t>(M.unk_id), __float_as_uint(M.unk_score), __float_as_uint(bt),
                                  bp | ((bp + 1) << 16));
            maxbt[bp + 1] = fmaxf(maxbt[bp + 1], bt);
            ++nn;
          }
        }
        if (!overflow) {
          const float eos_bt = __fadd_rn(maxbt[L], 0.f);
          // end_nodes lists in insertion order (stable counting sort by end character); BOS is the list of 0
          for (uint32_t p = 0; p <= L + 1; ++p) end_off[p] = 0;
          end_off[0 + 1] += 1;
          for (uint32_t i = 2; i < nn; ++i) end_off[(node[i].w >> 16) + 1] += 1;
          for (uint32_t p = 0; p <= L; ++p) end_off[p + 1] += end_off[p];
          for (uint32_t p = 0; p <= L; ++p) sort_cur[p] = end_off[p];
          elist[sort_cur[0]++] = make_uint4(0u, 0u, 0u, 0u);
          for (uint32_t i = 2; i < nn; ++i) {
            const uint4 nd = node[i];
            const uint32_t b = nd.w & 0xFFFFu, e = nd.w >> 16;
            const bool isunk = static_cast<int32_t>(nd.x) == M.unk_id;
            const uint32_t contrib = (isunk && bf) ? static_cast<uint32_t>(surf[e] - surf[b]) : 1u;
            elist[sort_cur[e]++] = make_uint4(i | (b << 16), nd.z, nd.y, contrib | (isunk ? 0x80000000u : 0u));
          }
          // ---- Lattice::NBest: backward A* ----
          uint32_t pn = 1, hn = 0;
          hyp[0] = make_uint4(0xFFFFFFFFu, 0u, 1u, L);  // EOS: next = null, gx = 0, no ids yet
          heap_push(hn, make_uint2(__float_as_uint(eos_bt), 0u));
          const uint32_t shrink_to = nbest * 10 < 512 ? nbest * 10 : 512;
          while (hn && !overflow) {
            const uint2 te = heap_pop(hn);
            const uint4 th = hyp[te.y];
            if ((th.z & 0xFFFFu) == 0) {  // reached BOS: one result
              const uint32_t cnt = th.z >> 16;
              const unsigned long long pos = atomicAdd(O.cursor, static_cast<unsigned long long>(cnt));
              O.cand_start[cbase + K] = pos;
              O.cand_count[cbase + K] = cnt;
              O.cand_score[cbase + K] = __uint_as_float(te.x);
              if (pos + cnt > O.tmp_cap) {
                atomicOr(O.status + 2, 1u);
                O.cand_count[cbase + K] = 0;
              } else {
                uint32_t w = 0;
                bool prev_unk = false;
                for (uint4 h = hyp[th.x]; h.x != 0xFFFFFFFFu; h = hyp[h.x]) {
                  const uint4 nd = node[h.z & 0xFFFFu];
                  const bool isunk = static_cast<int32_t>(nd.x) == M.unk_id;
                  if (isunk) {
                    if (bf) {
                      for (uint32_t k = surf[nd.w & 0xFFFFu]; k < surf[nd.w >> 16]; ++k)
                        O.tmp_ids[pos + (w++)] = __ldg(M.byte_to_id + text_byte(k));
                    } else if (!prev_unk) {
                      O.tmp_ids[pos + (w++)] = M.unk_id;
                    }
                  } else {
                    O.tmp_ids[pos + (w++)] = static_cast<int32_t>(nd.x);
                  }
                  prev_unk = isunk;
                }
                if (w != cnt) atomicOr(O.status + 1, 1u);
              }
              if (++K == nbest) break;
              continue;
            }
            // expand: one hypothesis per node ending where this one begins, in end_nodes order
            const uint32_t p0 = th.w & 0xFFFFu;
            const bool top_unk = (th.w >> 16) & 1u;
            const float top_gx = __uint_as_float(th.y);
            const uint32_t top_cnt = th.z >> 16;
            uint32_t q = end_off[p0];
            const uint32_t qe = end_off[p0 + 1];
            if (pn + (qe - q) > G.hyp_cap || hn + (qe - q) + 1 >= G.heap_cap - 512u) { overflow = true; break; }
            uint4 rec = elist[q];
            for (; q < qe; ++q) {
              const uint4 nxt = elist[q + 1 < qe ? q + 1 : q];  // issued ahead of the agenda update below
              const bool isunk = rec.w >> 31;
              uint32_t cn = top_cnt + (rec.w & 0x7FFFFFFFu);
              if (!bf && isunk && top_unk) --cn;
              hyp[pn] = make_uint4(te.y, __float_as_uint(__fadd_rn(__uint_as_float(rec.z), top_gx)),
                                   (rec.x & 0xFFFFu) | (cn << 16), (rec.x >> 16) | (isunk ? 0x10000u : 0u));
              heap_push(hn, make_uint2(__float_as_uint(__fadd_rn(__uint_as_float(rec.y), top_gx)), pn));
              ++pn;
              rec = nxt;
            }
            if (hn >= 10000u) {  // agenda shrink (:481-505): pop the best `shrink_to`, clear, push them back in that order
              uint2 *keep = heap + (G.heap_cap + 1 - shrink_to);
              for (uint32_t i = 0; i < shrink_to; ++i) keep[i] = heap_pop(hn);
              hn = 0;
              for (uint32_t i = 0; i < shrink_to; ++i) heap_push(hn, keep[i]);
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
