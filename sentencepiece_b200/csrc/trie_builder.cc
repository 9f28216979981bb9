// trie_builder.cc -- see trie_builder.h.
#include "trie_builder.h"

#include <algorithm>
#include <cstring>
#include <queue>

namespace spm_b200 {
namespace {

struct Node {
  uint32_t first_child = 0;  // index into children_ (filled after sorting)
  uint32_t n_children = 0;
  int32_t key = -1;          // index into keys, -1 if none
  float weight = 0.f;
  uint32_t unit = 0;
};

}  // namespace

bool BuildDeviceTrie(const std::vector<TrieKey> &keys, int vocab_size, DeviceTrie *out, std::string *err) {
  *out = DeviceTrie();
  // ---- 1. plain trie over sorted keys ------------------------------------
  std::vector<uint32_t> order(keys.size());
  for (uint32_t i = 0; i < keys.size(); ++i) order[i] = i;
  auto key_less = [&](uint32_t a, uint32_t b) {
    const TrieKey &x = keys[a], &y = keys[b];
    const int c = memcmp(x.data, y.data, std::min(x.len, y.len));
    return c != 0 ? c < 0 : x.len < y.len;
  };
  std::sort(order.begin(), order.end(), key_less);
  for (const TrieKey &k : keys) {
    if (k.len == 0) { *err = "piece must not be empty."; return false; }
    if (memchr(k.data, 0, k.len)) { *err = "pieces containing NUL bytes are not supported."; return false; }
    out->max_key_len = std::max(out->max_key_len, k.len);
  }
  // Build level by level from the sorted key list: nodes[] with child ranges.
  struct Edge { uint32_t parent; uint8_t label; uint32_t child; };
  std::vector<Node> nodes(1);
  std::vector<uint8_t> node_label(1, 0);
  std::vector<uint32_t> node_parent(1, 0);
  std::vector<std::vector<std::pair<uint8_t, uint32_t>>> kids(1);
  {
    // incremental insertion following the sorted order keeps `path` = nodes of the previous key
    std::vector<uint32_t> path;  // path[d] = node after d+1 bytes of the previous key
    const TrieKey *prev = nullptr;
    for (uint32_t oi : order) {
      const TrieKey &k = keys[oi];
      uint32_t common = 0;
      if (prev) {
        const uint32_t m = std::min(prev->len, k.len);
        while (common < m && prev->data[common] == k.data[common]) ++common;
        if (common == k.len && prev->len == k.len) { *err = "piece is already defined."; return false; }
      }
      path.resize(common);
      uint32_t cur = common ? path[common - 1] : 0;
      for (uint32_t d = common; d < k.len; ++d) {
        const uint32_t nn = static_cast<uint32_t>(nodes.size());
        nodes.emplace_back();
        node_label.push_back(static_cast<uint8_t>(k.data[d]));
        node_parent.push_back(cur);
        kids.emplace_back();
        kids[cur].emplace_back(static_cast<uint8_t>(k.data[d]), nn);
        path.push_back(nn);
        cur = nn;
      }
      nodes[cur].key = static_cast<int32_t>(oi);
      prev = &k;
    }
  }
  out->num_nodes = static_cast<uint32_t>(nodes.size());
  // subtree weights (children were created after parents -> reverse index order is bottom-up)
  for (uint32_t n = 0; n < nodes.size(); ++n)
    if (nodes[n].key >= 0) nodes[n].weight = keys[nodes[n].key].weight;
  for (uint32_t n = static_cast<uint32_t>(nodes.size()) - 1; n > 0; --n) nodes[node_parent[n]].weight += nodes[n].weight;

  // trie_results_size_: the maximum number of keys that are prefixes of one key
  // (unigram_model.cc:635-644)
  {
    std::vector<uint32_t> depth_keys(nodes.size(), 0);
    for (uint32_t n = 1; n < nodes.size(); ++n) {
      depth_keys[n] = depth_keys[node_parent[n]] + (nodes[n].key >= 0 ? 1 : 0);
      out->max_matches_per_start = std::max(out->max_matches_per_start, depth_keys[n]);
    }
  }

  // ---- 2. double-array allocation, hottest parents first ------------------
  std::vector<uint32_t> &link = out->link;
  std::vector<uint8_t> used, used_base;
  auto add_block = [&]() {
    link.resize(link.size() + 256, kLinkInvalidLabel);
    used.resize(used.size() + 256, 0);
    used_base.resize(used_base.size() + 256, 0);
  };
  add_block();
  used[0] = 1;  // root
  used_base[0] = 1;  // base 0 is reserved: leaves/unused units carry base 0 and must never see children
  nodes[0].unit = 0;
  std::vector<uint32_t> unit_node(256, 0xFFFFFFFFu);
  unit_node[0] = 0;
  constexpr uint32_t kOpenBlocks = 16;
  uint32_t first_open = 0;

  auto cmp = [&](uint32_t a, uint32_t b) {
    if (nodes[a].weight != nodes[b].weight) return nodes[a].weight < nodes[b].weight;
    return a > b;  // deterministic: older node first
  };
  std::priority_queue<uint32_t, std::vector<uint32_t>, decltype(cmp)> pq(cmp);
  pq.push(0);
  std::vector<uint32_t> base_of(nodes.size(), 0);
  while (!pq.empty()) {
    const uint32_t n = pq.top();
    pq.pop();
    auto &ch = kids[n];
    if (ch.empty()) continue;
    // children were appended in sorted key order, so labels are ascending already
    uint32_t base = 0xFFFFFFFFu;
    const uint32_t n_blocks = static_cast<uint32_t>(link.size() / 256);
    for (uint32_t blk = first_open; blk < n_blocks && base == 0xFFFFFFFFu; ++blk) {
      for (uint32_t u = blk * 256; u < blk * 256 + 256; ++u) {
        if (used[u]) continue;
        const uint32_t b = u ^ ch[0].first;
        if (used_base[b]) continue;
        bool ok = true;
        for (size_t i = 1; i < ch.size(); ++i)
          if (used[b ^ ch[i].first]) { ok = false; break; }
        if (ok) { base = b; break; }
      }
    }
    if (base == 0xFFFFFFFFu) {
      const uint32_t blk = n_blocks;
      add_block();
      unit_node.resize(link.size(), 0xFFFFFFFFu);
      base = (blk * 256) ^ ch[0].first;
      if (link.size() / 256 - first_open > kOpenBlocks) first_open = static_cast<uint32_t>(link.size() / 256) - kOpenBlocks;
    }
    if (link.size() > kMaxTrieUnits) { *err = "vocabulary too large for the device trie (unit limit)."; return false; }
    used_base[base] = 1;
    base_of[n] = base;
    for (auto &c : ch) {
      const uint32_t u = base ^ c.first;
      used[u] = 1;
      nodes[c.second].unit = u;
      unit_node[u] = c.second;
      pq.push(c.second);
    }
  }
  // ---- 3. emit -----------------------------------------------------------
  out->cmask.assign(link.size(), 0u);
  for (uint32_t n = 0; n < nodes.size(); ++n) {
    uint32_t m = 0;
    for (const auto &c : kids[n]) m |= 1u << (c.first & 31u);
    out->cmask[nodes[n].unit] = m;
  }
  out->val.assign(link.size(), 0xFFFFFFFFu);
  out->id.assign(link.size(), -1);
  out->unit_of_id.assign(static_cast<size_t>(vocab_size), 0xFFFFFFFFu);
  for (uint32_t n = 0; n < nodes.size(); ++n) {
    const uint32_t u = nodes[n].unit;
    uint32_t kind = kKindNone;
    if (nodes[n].key >= 0) {
      const TrieKey &k = keys[nodes[n].key];
      kind = k.kind;
      memcpy(&out->val[u], &k.score, 4);
      out->id[u] = k.id;
      if (k.id >= 0 && k.id < vocab_size) out->unit_of_id[k.id] = u;
    }
    const uint32_t label9 = n == 0 ? kLinkInvalidLabel : node_label[n];
    link[u] = (base_of[n] << kLinkBaseShift) | (kind << kLinkKindShift) | label9;
  }
  return true;
}

}  // namespace spm_b200
