/* oracle/spm_oracle.c -- TEST INFRASTRUCTURE ONLY (see spm_oracle.h).
 *
 * Plain-C restatement of the reference's encode hot path, written from the
 * reference's behaviour, each function citing the file:line it follows
 * (paths relative to /root/reference).  Data structures are deliberately the
 * simplest thing that is obviously right (a hashed trie, arrays, a binary heap);
 * no attempt is made to be fast.
 *
 * Parity status: PINNED.  The CPU tests check this file
 *   (a) against the reference's own known-answer vectors for the path
 *       (src/normalizer_test.cc:37-357, src/unigram_model_test.cc:782-928,
 *        src/bpe_model_test.cc:49-250, src/sentencepiece_processor_test.cc:186-303,
 *        src/util_test.cc:127-226), restated in tests/test_oracle_kat.py,
 *   (b) against outputs of the reference itself: the UNMODIFIED reference compiled
 *       into oracle/_ref (ctypes shim oracle/ref_shim.cc) on seeded corpora
 *       (tests/test_oracle_golden.py, test_oracle_nbest.py, test_oracle_lattice.py,
 *       test_oracle_decode.py), and the committed golden id dumps under tests/golden/
 *       produced by tools/make_golden.py with that same compiled reference.
 */
#include "spm_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ utf-8 -- */

/* string_util::OneCharLen, src/util.h:151-153 */
static inline size_t one_char_len(unsigned char c) {
  static const unsigned char T[16] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 3, 4};
  return T[c >> 4];
}
/* IsTrailByte, src/util.h:157 */
static inline int is_trail(unsigned char c) { return (c & 0xC0) == 0x80; }
/* IsValidCodepoint, src/util.h:159-161 */
static inline int valid_cp(uint32_t c) { return c < 0xD800 || (c >= 0xE000 && c <= 0x10FFFF); }

/* string_util::DecodeUTF8, src/util.cc:51-84.  Returns the code point or
 * 0xFFFD (kUnicodeError) with *mblen = 1 for malformed input. */
static uint32_t decode_utf8(const unsigned char *b, size_t len, size_t *mblen) {
  if (b[0] < 0x80) { *mblen = 1; return b[0]; }
  if (len >= 2 && (b[0] & 0xE0) == 0xC0) {
    const uint32_t cp = ((uint32_t)(b[0] & 0x1F) << 6) | (b[1] & 0x3F);
    if (is_trail(b[1]) && cp >= 0x80 && valid_cp(cp)) { *mblen = 2; return cp; }
  } else if (len >= 3 && (b[0] & 0xF0) == 0xE0) {
    const uint32_t cp = ((uint32_t)(b[0] & 0x0F) << 12) | ((uint32_t)(b[1] & 0x3F) << 6) | (b[2] & 0x3F);
    if (is_trail(b[1]) && is_trail(b[2]) && cp >= 0x800 && valid_cp(cp)) { *mblen = 3; return cp; }
  } else if (len >= 4 && (b[0] & 0xF8) == 0xF0) {
    const uint32_t cp = ((uint32_t)(b[0] & 0x07) << 18) | ((uint32_t)(b[1] & 0x3F) << 12) |
                        ((uint32_t)(b[2] & 0x3F) << 6) | (b[3] & 0x3F);
    if (is_trail(b[1]) && is_trail(b[2]) && is_trail(b[3]) && cp >= 0x10000 && valid_cp(cp)) {
      *mblen = 4;
      return cp;
    }
  }
  *mblen = 1;
  return 0xFFFD;
}
/* IsValidDecodeUTF8, src/util.h:173-176: a literal U+FFFD (3 bytes) is valid */
static int is_valid_decode_utf8(const unsigned char *b, size_t len, size_t *mblen) {
  const uint32_t c = decode_utf8(b, len, mblen);
  return c != 0xFFFD || *mblen == 3;
}

/* ------------------------------------------------------------ hashed trie -- */

typedef struct {
  uint32_t *keys;   /* ((node << 8) | label) + 1 ; 0 = empty */
  int32_t *child;
  uint32_t cap;     /* power of two */
  uint32_t used;
  int32_t *value;   /* per node: key value or -1 */
  uint32_t n_nodes, node_cap;
} htrie;

static void ht_init(htrie *t) {
  memset(t, 0, sizeof *t);
  t->cap = 1024;
  t->keys = calloc(t->cap, sizeof(uint32_t));
  t->child = malloc(t->cap * sizeof(int32_t));
  t->node_cap = 1024;
  t->value = malloc(t->node_cap * sizeof(int32_t));
  t->value[0] = -1;
  t->n_nodes = 1;
}
static void ht_free(htrie *t) { free(t->keys); free(t->child); free(t->value); }
static inline uint32_t ht_hash(uint32_t k) { k *= 0x9E3779B1u; return k ^ (k >> 15); }
static int32_t ht_child(const htrie *t, uint32_t node, unsigned char label) {
  const uint32_t k = ((node << 8) | label) + 1;
  uint32_t h = ht_hash(k) & (t->cap - 1);
  while (t->keys[h]) {
    if (t->keys[h] == k) return t->child[h];
    h = (h + 1) & (t->cap - 1);
  }
  return -1;
}
static void ht_put(htrie *t, uint32_t k, int32_t c) {
  uint32_t h = ht_hash(k) & (t->cap - 1);
  while (t->keys[h]) h = (h + 1) & (t->cap - 1);
  t->keys[h] = k;
  t->child[h] = c;
  t->used++;
}
static void ht_grow(htrie *t) {
  uint32_t *ok = t->keys;
  int32_t *oc = t->child;
  const uint32_t ocap = t->cap;
  t->cap *= 2;
  t->keys = calloc(t->cap, sizeof(uint32_t));
  t->child = malloc(t->cap * sizeof(int32_t));
  t->used = 0;
  for (uint32_t i = 0; i < ocap; ++i) if (ok[i]) ht_put(t, ok[i], oc[i]);
  free(ok); free(oc);
}
/* returns 0 if inserted, 1 if the key already existed */
static int ht_insert(htrie *t, const char *s, size_t len, int32_t value) {
  uint32_t node = 0;
  for (size_t i = 0; i < len; ++i) {
    int32_t c = ht_child(t, node, (unsigned char)s[i]);
    if (c < 0) {
      if (t->used * 2 >= t->cap) ht_grow(t);
      if (t->n_nodes == t->node_cap) {
        t->node_cap *= 2;
        t->value = realloc(t->value, t->node_cap * sizeof(int32_t));
      }
      c = (int32_t)t->n_nodes++;
      t->value[c] = -1;
      ht_put(t, ((node << 8) | (unsigned char)s[i]) + 1, c);
    }
    node = (uint32_t)c;
  }
  if (t->value[node] >= 0) return 1;
  t->value[node] = value;
  return 0;
}
static int32_t ht_exact(const htrie *t, const char *s, size_t len) {
  uint32_t node = 0;
  for (size_t i = 0; i < len; ++i) {
    const int32_t c = ht_child(t, node, (unsigned char)s[i]);
    if (c < 0) return -1;
    node = (uint32_t)c;
  }
  return t->value[node];
}

/* ------------------------------------------------------------------ model -- */

struct oracle_model {
  int32_t model_type, vocab_size;
  char *piece_bytes;
  uint32_t *piece_off;
  float *scores;
  uint8_t *types;
  uint8_t byte_fallback, add_dummy_prefix, remove_extra_whitespaces, escape_whitespaces,
      treat_whitespace_as_suffix;
  /* ModelInterface::InitializePieces, src/model_interface.cc:63-151 */
  htrie pieces;    /* pieces_: NORMAL | USER_DEFINED | UNUSED -> id */
  htrie reserved;  /* reserved_id_map_: CONTROL | UNKNOWN | BYTE -> id */
  htrie user;      /* PrefixMatcher over USER_DEFINED, src/normalizer.cc:311-322 */
  int has_user;
  int32_t unk_id;
  float min_score, max_score;
  int32_t byte_to_id[256];
  /* precompiled charsmap, src/normalizer.cc:274-309 */
  uint8_t *charsmap;
  const uint32_t *cm_units;
  size_t cm_nunits;
  const char *cm_targets;
  size_t cm_targets_len;
  char *unk_surface; /* Decode: TrainerSpec.unk_surface */
  size_t unk_surface_len;
};

void oracle_free(void *p) { free(p); }
float oracle_min_score(const oracle_model *m) { return m->min_score; }
float oracle_max_score(const oracle_model *m) { return m->max_score; }
int32_t oracle_unk_id(const oracle_model *m) { return m->unk_id; }

/* ModelInterface::PieceToId, src/model_interface.cc:51-61 */
static int32_t piece_to_id(const oracle_model *m, const char *s, size_t len) {
  int32_t id = ht_exact(&m->reserved, s, len);
  if (id >= 0) return id;
  id = ht_exact(&m->pieces, s, len);
  if (id >= 0) return id;
  return m->unk_id;
}

static void set_err(char *err, size_t errlen, const char *msg) {
  if (err && errlen) snprintf(err, errlen, "%s", msg);
}

oracle_model *oracle_create(const oracle_model_desc *d, char *err, size_t errlen) {
  oracle_model *m = calloc(1, sizeof *m);
  m->model_type = d->model_type;
  m->vocab_size = d->vocab_size;
  const uint32_t total = d->piece_off[d->vocab_size];
  m->piece_bytes = malloc(total + 1);
  memcpy(m->piece_bytes, d->piece_bytes, total);
  m->piece_off = malloc(sizeof(uint32_t) * (size_t)(d->vocab_size + 1));
  memcpy(m->piece_off, d->piece_off, sizeof(uint32_t) * (size_t)(d->vocab_size + 1));
  m->scores = malloc(sizeof(float) * (size_t)d->vocab_size);
  memcpy(m->scores, d->scores, sizeof(float) * (size_t)d->vocab_size);
  m->types = malloc((size_t)d->vocab_size);
  memcpy(m->types, d->types, (size_t)d->vocab_size);
  m->byte_fallback = d->byte_fallback;
  m->add_dummy_prefix = d->add_dummy_prefix;
  m->remove_extra_whitespaces = d->remove_extra_whitespaces;
  m->escape_whitespaces = d->escape_whitespaces;
  m->treat_whitespace_as_suffix = d->treat_whitespace_as_suffix;
  ht_init(&m->pieces);
  ht_init(&m->reserved);
  ht_init(&m->user);
  m->unk_id = -1;
  int byte_found[256] = {0};
  /* src/model_interface.cc:88-133 */
  for (int32_t i = 0; i < d->vocab_size; ++i) {
    const char *p = m->piece_bytes + m->piece_off[i];
    const size_t l = m->piece_off[i + 1] - m->piece_off[i];
    const uint8_t t = m->types[i];
    if (l == 0) { set_err(err, errlen, "piece must not be empty."); goto fail; }
    const int normal = (t == ORACLE_NORMAL || t == ORACLE_USER_DEFINED || t == ORACLE_UNUSED);
    if (ht_insert(normal ? &m->pieces : &m->reserved, p, l, i)) {
      set_err(err, errlen, "piece is already defined.");
      goto fail;
    }
    if (t == ORACLE_USER_DEFINED) { ht_insert(&m->user, p, l, i); m->has_user = 1; }
    if (t == ORACLE_UNKNOWN) {
      if (m->unk_id >= 0) { set_err(err, errlen, "unk is already defined."); goto fail; }
      m->unk_id = i;
    }
    if (t == ORACLE_BYTE) {
      if (!m->byte_fallback) { set_err(err, errlen, "byte piece found although byte_fallback is false."); goto fail; }
      unsigned v;
      char tail;
      /* PieceToByte: exact "<0x%02X>" form, src/model_interface.cc:210-229 */
      if (l == 6 && sscanf(p, "<0x%02X%c", &v, &tail) == 2 && tail == '>' && v < 256) {
        char canon[8];
        snprintf(canon, sizeof canon, "<0x%02X>", v);
        if (memcmp(canon, p, 6) != 0) { set_err(err, errlen, "byte piece is invalid."); goto fail; }
        byte_found[v] = 1;
      } else { set_err(err, errlen, "byte piece is invalid."); goto fail; }
    }
  }
  if (m->unk_id < 0) { set_err(err, errlen, "unk is not defined."); goto fail; }
  if (m->byte_fallback)
    for (int b = 0; b < 256; ++b)
      if (!byte_found[b]) { set_err(err, errlen, "there are not 256 byte pieces although byte_fallback is true."); goto fail; }
  /* unigram::Model ctor, src/unigram_model.cc:657-664 (NB: max starts at FLT_MIN) */
  m->min_score = FLT_MAX;
  m->max_score = FLT_MIN;
  for (int32_t i = 0; i < d->vocab_size; ++i)
    if (m->types[i] == ORACLE_NORMAL) {
      if (m->scores[i] < m->min_score) m->min_score = m->scores[i];
      if (m->scores[i] > m->max_score) m->max_score = m->scores[i];
    }
  /* PieceToId(ByteToPiece(b)), src/sentencepiece_processor.cc:587-588 */
  for (int b = 0; b < 256; ++b) {
    char bp[8];
    snprintf(bp, sizeof bp, "<0x%02X>", b);
    m->byte_to_id[b] = piece_to_id(m, bp, 6);
  }
  /* Normalizer::Init / DecodePrecompiledCharsMap, src/normalizer.cc:47-69,274-309 */
  if (d->charsmap_len) {
    uint32_t trie_bytes = 0;
    if (d->charsmap_len <= 4) { set_err(err, errlen, "Blob for normalization rule is broken."); goto fail; }
    memcpy(&trie_bytes, d->charsmap, 4);
    if (trie_bytes >= d->charsmap_len) { set_err(err, errlen, "Trie data size exceeds the input blob size."); goto fail; }
    /* keep a 4-byte-aligned private copy */
    m->charsmap = malloc(d->charsmap_len + 8);
    memcpy(m->charsmap, d->charsmap, d->charsmap_len);
    m->charsmap[d->charsmap_len] = 0;
    m->cm_units = (const uint32_t *)(m->charsmap + 4);
    m->cm_nunits = trie_bytes / 4;
    m->cm_targets = (const char *)m->charsmap + 4 + trie_bytes;
    m->cm_targets_len = d->charsmap_len - 4 - trie_bytes;
  }
  return m;
fail:
  oracle_destroy(m);
  return NULL;
}

void oracle_destroy(oracle_model *m) {
  if (!m) return;
  ht_free(&m->pieces); ht_free(&m->reserved); ht_free(&m->user);
  free(m->piece_bytes); free(m->piece_off); free(m->scores); free(m->types); free(m->charsmap);
  free(m->unk_surface);
  free(m);
}

void oracle_set_types(oracle_model *m, const uint8_t *types) {
  memcpy(m->types, types, (size_t)m->vocab_size);
}

/* ------------------------------------------------------ darts (charsmap) -- */

/* Darts::DoubleArrayUnit, third_party/darts_clone/darts.h:50-80 */
static inline uint32_t da_offset(uint32_t u) { return (u >> 10) << ((u & (1u << 9)) >> 6); }
static inline uint32_t da_label(uint32_t u) { return u & ((1u << 31) | 0xFF); }
static inline int da_has_leaf(uint32_t u) { return (u >> 8) & 1; }
static inline uint32_t da_value(uint32_t u) { return u & ((1u << 31) - 1); }

/* Longest match of commonPrefixSearch (darts.h:469-513) as used by
 * NormalizePrefix (normalizer.cc:215-228): returns the longest key length (0 if
 * none) and its value. */
static size_t charsmap_longest(const oracle_model *m, const unsigned char *key, size_t len, uint32_t *value) {
  if (!m->cm_units) return 0;
  size_t longest = 0;
  uint32_t node = 0;
  uint32_t unit = m->cm_units[node];
  node ^= da_offset(unit);
  for (size_t i = 0; i < len; ++i) {
    node ^= key[i];
    if (node >= m->cm_nunits) break; /* the reference would read out of bounds; blobs are well formed */
    unit = m->cm_units[node];
    if (da_label(unit) != key[i]) break;
    node ^= da_offset(unit);
    if (da_has_leaf(unit)) {
      longest = i + 1; /* results come in increasing length; the last one wins (normalizer.cc:223-228) */
      *value = da_value(m->cm_units[node]);
    }
  }
  return longest;
}

/* PrefixMatcher::PrefixMatch, src/normalizer.cc:324-346 */
static size_t prefix_match(const oracle_model *m, const unsigned char *w, size_t len, int *found) {
  size_t mblen = 0;
  *found = 0;
  if (m->has_user) {
    uint32_t node = 0;
    for (size_t i = 0; i < len; ++i) {
      const int32_t c = ht_child(&m->user, node, w[i]);
      if (c < 0) break;
      node = (uint32_t)c;
      if (m->user.value[node] >= 0) { mblen = i + 1; *found = 1; }
    }
  }
  if (!*found) {
    const size_t l = one_char_len(w[0]);
    mblen = len < l ? len : l;
  }
  return mblen;
}

/* --------------------------------------------------------------- normalize -- */

typedef struct { char *s; uint64_t *map; size_t n, cap; } nbuf;
static void nb_push(nbuf *b, char c, uint64_t consumed) {
  if (b->n == b->cap) {
    b->cap = b->cap ? b->cap * 2 : 256;
    b->s = realloc(b->s, b->cap + 1);
    b->map = realloc(b->map, sizeof(uint64_t) * (b->cap + 2));
  }
  b->s[b->n] = c;
  b->map[b->n] = consumed;
  b->n++;
}

/* Normalizer::NormalizePrefix, src/normalizer.cc:195-253.  Returns the consumed
 * byte count; *sp / *sp_len is the replacement string. */
static size_t normalize_prefix(const oracle_model *m, const unsigned char *in, size_t len, const char **sp,
                               size_t *sp_len) {
  int found = 0;
  const size_t ml = prefix_match(m, in, len, &found);
  if (found) { *sp = (const char *)in; *sp_len = ml; return ml; }
  uint32_t value = 0;
  const size_t longest = charsmap_longest(m, in, len, &value);
  if (longest == 0) {
    size_t l = 0;
    if (!is_valid_decode_utf8(in, len, &l)) {
      *sp = "\xEF\xBF\xBD"; *sp_len = 3; return 1;
    }
    *sp = (const char *)in; *sp_len = l; return l;
  }
  *sp = m->cm_targets + value;
  *sp_len = strlen(*sp); /* NUL-delimited, normalizer.cc:247-249 */
  return longest;
}

static void add_ws(const oracle_model *m, nbuf *b, uint64_t consumed) {
  if (m->escape_whitespaces) {
    nb_push(b, (char)0xE2, consumed); nb_push(b, (char)0x96, consumed); nb_push(b, (char)0x81, consumed);
  } else {
    nb_push(b, ' ', consumed);
  }
}

int oracle_normalize(const oracle_model *m, const char *in_, size_t len, char **out, size_t *out_len,
                     uint64_t **n2o, size_t *n2o_len) {
  const unsigned char *in = (const unsigned char *)in_;
  nbuf b = {0};
  *out = NULL; *out_len = 0; *n2o = NULL; *n2o_len = 0;
  if (len == 0) return 0;
  uint64_t consumed = 0;
  const char *sp; size_t spl;
  /* heading spaces, normalizer.cc:86-95 */
  if (m->remove_extra_whitespaces) {
    while (len) {
      const size_t c = normalize_prefix(m, in, len, &sp, &spl);
      if (!(spl == 1 && sp[0] == ' ')) break;
      in += c; len -= c; consumed += c;
    }
  }
  if (len == 0) return 0;
  if (!m->treat_whitespace_as_suffix && m->add_dummy_prefix) add_ws(m, &b, consumed);
  int is_prev_space = m->remove_extra_whitespaces;
  while (len) {
    const size_t c = normalize_prefix(m, in, len, &sp, &spl);
    while (is_prev_space && spl && sp[0] == ' ') { ++sp; --spl; }
    if (spl) {
      for (size_t n = 0; n < spl; ++n) {
        if (m->escape_whitespaces && sp[n] == ' ') {
          nb_push(&b, (char)0xE2, consumed); nb_push(&b, (char)0x96, consumed); nb_push(&b, (char)0x81, consumed);
        } else {
          nb_push(&b, sp[n], consumed);
        }
      }
      is_prev_space = sp[spl - 1] == ' ';
    }
    consumed += c; in += c; len -= c;
    if (!m->remove_extra_whitespaces) is_prev_space = 0;
  }
  /* trailing spaces, normalizer.cc:166-176 */
  if (m->remove_extra_whitespaces) {
    const char *space = m->escape_whitespaces ? "\xE2\x96\x81" : " ";
    const size_t sl = m->escape_whitespaces ? 3 : 1;
    while (b.n >= sl && memcmp(b.s + b.n - sl, space, sl) == 0) {
      const size_t length = b.n - sl;
      consumed = b.map[length];
      b.n = length;
    }
  }
  if (m->treat_whitespace_as_suffix && m->add_dummy_prefix) add_ws(m, &b, consumed);
  if (b.n == 0 && b.s == NULL) { /* nothing was ever pushed */
    b.s = malloc(1); b.map = malloc(sizeof(uint64_t) * 2);
  }
  b.map[b.n] = consumed; /* normalizer.cc:181 (capacity reserved by nb_push) */
  *out = b.s; *out_len = b.n; *n2o = b.map; *n2o_len = b.n + 1;
  return 0;
}

/* ------------------------------------------------------------------ unigram -- */

typedef struct { int32_t id; float best_path_score; int32_t starts_at; } best_node;

/* unigram::Model::EncodeOptimized, src/unigram_model.cc:889-1020 */
static int unigram_encode(const oracle_model *m, const unsigned char *norm, size_t size, int32_t **ids_out,
                          uint32_t **ends_out, size_t *n_out) {
  *ids_out = NULL; *ends_out = NULL; *n_out = 0;
  if (size == 0) return 0;
  const float unk_score = m->min_score - 10.0f; /* kUnkPenalty, unigram_model.cc:955 */
  best_node *best = malloc(sizeof(best_node) * (size + 1));
  for (size_t i = 0; i <= size; ++i) { best[i].id = -1; best[i].best_path_score = 0; best[i].starts_at = -1; }
  size_t starts_at = 0;
  while (starts_at < size) {
    uint32_t node = 0;
    size_t key_pos = starts_at;
    const float till_here = best[starts_at].best_path_score;
    int has_single_node = 0;
    size_t mblen = one_char_len(norm[starts_at]);
    if (mblen > size - starts_at) mblen = size - starts_at;
    while (key_pos < size) {
      const int32_t c = ht_child(&m->pieces, node, norm[key_pos]);
      if (c < 0) break;               /* traverse() == -2 */
      node = (uint32_t)c;
      ++key_pos;
      const int32_t ret = m->pieces.value[node];
      if (ret < 0) continue;          /* traverse() == -1 */
      if (m->types[ret] == ORACLE_UNUSED) continue;
      best_node *t = &best[key_pos];
      const size_t length = key_pos - starts_at;
      /* Q1: `score` has type double (common type of double and float), the
       * candidate is double, the comparison is double > (double)float and the
       * store truncates to float.  unigram_model.cc:979-989 */
      const double score = m->types[ret] == ORACLE_USER_DEFINED
                               ? ((double)((float)length * m->max_score) - 0.1)
                               : (double)m->scores[ret];
      const double cand = score + (double)till_here;
      if (t->starts_at == -1 || cand > (double)t->best_path_score) {
        t->best_path_score = (float)cand;
        t->starts_at = (int32_t)starts_at;
        t->id = ret;
      }
      if (!has_single_node && length == mblen) has_single_node = 1;
    }
    if (!has_single_node) {
      best_node *t = &best[starts_at + mblen];
      const float cand = unk_score + till_here; /* float + float, unigram_model.cc:997-998 */
      if (t->starts_at == -1 || cand > t->best_path_score) {
        t->best_path_score = cand;
        t->starts_at = (int32_t)starts_at;
        t->id = m->unk_id;
      }
    }
    starts_at += mblen;
  }
  /* backtrack, unigram_model.cc:1010-1018 */
  size_t cnt = 0;
  for (size_t e = size; e > 0; e = (size_t)best[e].starts_at) ++cnt;
  int32_t *ids = malloc(sizeof(int32_t) * (cnt ? cnt : 1));
  uint32_t *ends = malloc(sizeof(uint32_t) * (cnt ? cnt : 1));
  size_t k = cnt;
  for (size_t e = size; e > 0; e = (size_t)best[e].starts_at) {
    --k;
    ids[k] = best[e].id;
    ends[k] = (uint32_t)e;
  }
  free(best);
  *ids_out = ids; *ends_out = ends; *n_out = cnt;
  return 0;
}

/* ---------------------------------------------------------------------- bpe -- */

typedef struct { int32_t left, right; float score; size_t size; } sym_pair;
typedef struct { int32_t prev, next; int freeze; size_t off, len; } symbol;
typedef struct { sym_pair *a; size_t n, cap; } agenda_t;

/* SymbolPairComparator, src/bpe_model.cc:51-57: "less" for a max-heap */
static inline int pair_less(const sym_pair *h1, const sym_pair *h2) {
  return h1->score < h2->score || (h1->score == h2->score && h1->left > h2->left);
}
static void agenda_push(agenda_t *q, sym_pair p) {
  if (q->n == q->cap) { q->cap = q->cap ? q->cap * 2 : 256; q->a = realloc(q->a, q->cap * sizeof(sym_pair)); }
  size_t i = q->n++;
  q->a[i] = p;
  while (i > 0) {
    const size_t parent = (i - 1) / 2;
    if (!pair_less(&q->a[parent], &q->a[i])) break;
    sym_pair t = q->a[parent]; q->a[parent] = q->a[i]; q->a[i] = t;
    i = parent;
  }
}
static sym_pair agenda_pop(agenda_t *q) {
  sym_pair top = q->a[0];
  q->a[0] = q->a[--q->n];
  size_t i = 0;
  for (;;) {
    size_t l = 2 * i + 1, r = l + 1, b = i;
    if (l < q->n && pair_less(&q->a[b], &q->a[l])) b = l;
    if (r < q->n && pair_less(&q->a[b], &q->a[r])) b = r;
    if (b == i) break;
    sym_pair t = q->a[b]; q->a[b] = q->a[i]; q->a[i] = t;
    i = b;
  }
  return top;
}

/* rev_merge, src/bpe_model.cc:72-76,102-106: keyed by the merged string's
 * CONTENT; a later insertion with the same content overwrites. */
typedef struct { size_t off, len, loff, llen, roff, rlen; } revm;
typedef struct { revm *a; size_t n, cap; } revm_tab;

static void revm_set(revm_tab *t, const unsigned char *base, revm e) {
  for (size_t i = 0; i < t->n; ++i)
    if (t->a[i].len == e.len && memcmp(base + t->a[i].off, base + e.off, e.len) == 0) { t->a[i] = e; return; }
  if (t->n == t->cap) { t->cap = t->cap ? t->cap * 2 : 16; t->a = realloc(t->a, t->cap * sizeof(revm)); }
  t->a[t->n++] = e;
}
static const revm *revm_find(const revm_tab *t, const unsigned char *base, size_t off, size_t len) {
  for (size_t i = 0; i < t->n; ++i)
    if (t->a[i].len == len && memcmp(base + t->a[i].off, base + off, len) == 0) return &t->a[i];
  return NULL;
}

typedef struct { int32_t *ids; uint32_t *ends; size_t n, cap; } outv;
static void out_push(outv *o, int32_t id, uint32_t end) {
  if (o->n == o->cap) {
    o->cap = o->cap ? o->cap * 2 : 64;
    o->ids = realloc(o->ids, o->cap * sizeof(int32_t));
    o->ends = realloc(o->ends, o->cap * sizeof(uint32_t));
  }
  o->ids[o->n] = id; o->ends[o->n] = end; o->n++;
}

/* resegment, src/bpe_model.cc:175-193 */
static void resegment(const oracle_model *m, const unsigned char *base, const revm_tab *rm, size_t off, size_t len,
                      outv *o) {
  const int32_t id = piece_to_id(m, (const char *)base + off, len);
  if (id == -1 || m->types[id] != ORACLE_UNUSED) { out_push(o, id, (uint32_t)(off + len)); return; }
  const revm *p = revm_find(rm, base, off, len);
  if (!p) { out_push(o, id, (uint32_t)(off + len)); return; }
  /* NB: the reference recurses on the string_views recorded in rev_merge, which
   * may point at ANOTHER occurrence of the same content; only the content
   * matters for ids, but the piece boundaries are taken relative to the current
   * occurrence here (lengths are what PopulateSentencePieceText consumes). */
  resegment(m, base, rm, off, p->llen, o);
  resegment(m, base, rm, off + p->llen, p->rlen, o);
}

/* bpe::Model::SampleEncode with alpha = 0, src/bpe_model.cc:38-203 */
static int bpe_encode(const oracle_model *m, const unsigned char *norm, size_t size, int32_t **ids_out,
                      uint32_t **ends_out, size_t *n_out) {
  *ids_out = NULL; *ends_out = NULL; *n_out = 0;
  if (size == 0) return 0;
  symbol *sym = malloc(sizeof(symbol) * size);
  size_t nsym = 0;
  /* split into characters; user-defined symbols are frozen, bpe_model.cc:110-120 */
  {
    size_t pos = 0;
    while (pos < size) {
      int found = 0;
      const size_t mblen = prefix_match(m, norm + pos, size - pos, &found);
      symbol s;
      s.freeze = found;
      s.off = pos; s.len = mblen;
      s.prev = nsym == 0 ? -1 : (int32_t)nsym - 1;
      pos += mblen;
      s.next = pos >= size ? -1 : (int32_t)nsym + 1;
      sym[nsym++] = s;
    }
  }
  agenda_t q = {0};
  revm_tab rm = {0};
#define MAYBE_ADD(L, R)                                                                       \
  do {                                                                                        \
    const int32_t l_ = (L), r_ = (R);                                                         \
    if (l_ == -1 || r_ == -1 || sym[l_].freeze || sym[r_].freeze) break;                      \
    const size_t plen = sym[l_].len + sym[r_].len;                                            \
    const int32_t pid = ht_exact(&m->pieces, (const char *)norm + sym[l_].off, plen);         \
    if (pid < 0) break;                                                                       \
    sym_pair h = {l_, r_, m->scores[pid], plen};                                              \
    agenda_push(&q, h);                                                                       \
    if (m->types[pid] == ORACLE_UNUSED) {                                                     \
      revm e = {sym[l_].off, plen, sym[l_].off, sym[l_].len, sym[r_].off, sym[r_].len};       \
      revm_set(&rm, norm, e);                                                                 \
    }                                                                                         \
  } while (0)
  for (size_t i = 1; i < nsym; ++i) MAYBE_ADD((int32_t)i - 1, (int32_t)i);
  while (q.n) {
    const sym_pair top = agenda_pop(&q);
    /* stale entry check, bpe_model.cc:147-151 */
    if (sym[top.left].len == 0 || sym[top.right].len == 0 ||
        sym[top.left].len + sym[top.right].len != top.size)
      continue;
    sym[top.left].len += sym[top.right].len;
    sym[top.left].next = sym[top.right].next;
    if (sym[top.right].next >= 0) sym[sym[top.right].next].prev = top.left;
    sym[top.right].len = 0;
    MAYBE_ADD(sym[top.left].prev, top.left);
    MAYBE_ADD(top.left, sym[top.left].next);
  }
#undef MAYBE_ADD
  outv o = {0};
  for (int32_t i = 0; i != -1; i = sym[i].next) resegment(m, norm, &rm, sym[i].off, sym[i].len, &o);
  free(sym); free(q.a); free(rm.a);
  *ids_out = o.ids; *ends_out = o.ends; *n_out = o.n;
  return 0;
}

int oracle_model_encode(const oracle_model *m, const char *norm, size_t len, int32_t **ids, uint32_t **ends,
                        size_t *n) {
  if (m->model_type == ORACLE_UNIGRAM) return unigram_encode(m, (const unsigned char *)norm, len, ids, ends, n);
  if (m->model_type == ORACLE_BPE) return bpe_encode(m, (const unsigned char *)norm, len, ids, ends, n);
  return 2;
}

/* ----------------------------------------------------- processor: id path -- */

/* SentencePieceProcessor::Encode + PopulateSentencePieceText id path,
 * src/sentencepiece_processor.cc:547-651 */
int oracle_encode(const oracle_model *m, const char *in, size_t len, int32_t **ids_out, uint32_t **tok_end_out,
                  size_t *n_out) {
  char *norm; size_t nlen; uint64_t *n2o; size_t n2o_len;
  *ids_out = NULL; *tok_end_out = NULL; *n_out = 0;
  if (oracle_normalize(m, in, len, &norm, &nlen, &n2o, &n2o_len)) return 1;
  int32_t *ids; uint32_t *ends; size_t n;
  int rc = oracle_model_encode(m, norm ? norm : "", nlen, &ids, &ends, &n);
  if (rc) { free(norm); free(n2o); return rc; }
  outv o = {0};
  size_t consumed = 0;
  int is_prev_unk = 0;
  for (size_t k = 0; k < n; ++k) {
    const int32_t id = ids[k];
    const size_t wlen = ends[k] - (k ? ends[k - 1] : 0);
    if (wlen == 0) { rc = 3; break; } /* "Empty piece is not allowed." :557 */
    const int is_unk = (id == m->unk_id);                    /* IsUnknown, model_interface.h:207-210 */
    const int is_ctrl = id >= 0 && m->types[id] == ORACLE_CONTROL;
    if (is_ctrl) {
      /* zero-width control piece: `consumed` does not advance (:561-567); the
       * final consumed == size check then fails.  Encoders never emit these. */
      rc = 4;
      break;
    }
    if (is_unk && m->byte_fallback) {
      for (size_t i = 0; i < wlen; ++i)  /* one <0xXX> piece per byte, :581-603 */
        out_push(&o, m->byte_to_id[(unsigned char)norm[consumed + i]], (uint32_t)(consumed + i + 1));
    } else if (is_prev_unk && is_unk) {
      o.ends[o.n - 1] = (uint32_t)(consumed + wlen);         /* merge unk run, :609-613 */
    } else {
      out_push(&o, id, (uint32_t)(consumed + wlen));
    }
    consumed += wlen;
    is_prev_unk = is_unk;
  }
  if (!rc && consumed != nlen) rc = 5; /* "all normalized characters are not consumed." :628 */
  free(ids); free(ends); free(norm); free(n2o);
  if (rc) { free(o.ids); free(o.ends); return rc; }
  *ids_out = o.ids; *tok_end_out = o.ends; *n_out = o.n;
  return 0;
}

int oracle_encode_batch(const oracle_model *m, const char *bytes, const uint64_t *offs, size_t n, int32_t **ids_out,
                        uint64_t *id_offsets) {
  size_t cap = 1024, total = 0;
  int32_t *all = malloc(cap * sizeof(int32_t));
  for (size_t i = 0; i < n; ++i) {
    int32_t *ids; uint32_t *te; size_t k;
    const int rc = oracle_encode(m, bytes + offs[i], (size_t)(offs[i + 1] - offs[i]), &ids, &te, &k);
    if (rc) { free(all); return (int)(i + 1); }
    id_offsets[i] = total;
    if (total + k > cap) { while (total + k > cap) cap *= 2; all = realloc(all, cap * sizeof(int32_t)); }
    if (k) memcpy(all + total, ids, k * sizeof(int32_t));
    total += k;
    free(ids); free(te);
  }
  id_offsets[n] = total;
  *ids_out = all;
  return 0;
}

/* =========================================================== n-best + sampling == */

typedef struct {
  int32_t pos, length;     /* in Unicode characters (Lattice::Node, unigram_model.h:38-49) */
  int32_t id;
  float score, backtrace_score;
  uint32_t bbeg, bend;     /* byte span in the normalized text */
} lnode;
typedef struct { int32_t *a; size_t n, cap; } ivec;
static void iv_push(ivec *v, int32_t x) {
  if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 8; v->a = realloc(v->a, v->cap * sizeof(int32_t)); }
  v->a[v->n++] = x;
}
typedef struct { int32_t node, next; float fx, gx; } hyp_t;

/* std::push_heap / std::pop_heap on indices ordered by fx with comp(a,b) = a.fx < b.fx
 * (/usr/include/c++/13/bits/stl_heap.h:135-148, 224-267) */
static void heap_push(int32_t *h, size_t *n, const hyp_t *pool, int32_t v) {
  size_t hole = (*n)++;
  while (hole > 0) {
    const size_t parent = (hole - 1) / 2;
    if (!(pool[h[parent]].fx < pool[v].fx)) break; /* equal keys do not move up */
    h[hole] = h[parent];
    hole = parent;
  }
  h[hole] = v;
}
static int32_t heap_pop(int32_t *h, size_t *n, const hyp_t *pool) {
  /* __pop_heap: result = first; value = last; __adjust_heap(first, 0, len-1, value) */
  const int32_t top = h[0];
  const size_t len = --(*n);
  if (len == 0) return top;
  const int32_t value = h[len];
  size_t hole = 0, child = 0;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (pool[h[child]].fx < pool[h[child - 1]].fx) child--;
    h[hole] = h[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    h[hole] = h[child - 1];
    hole = child - 1;
  }
  /* __push_heap(first, hole, 0, value) */
  while (hole > 0) {
    const size_t parent = (hole - 1) / 2;
    if (!(pool[h[parent]].fx < pool[value].fx)) break;
    h[hole] = h[parent];
    hole = parent;
  }
  h[hole] = value;
  return top;
}

/* id path of PopulateSentencePieceText over (byte_len, id) pieces of one candidate */
static void emit_candidate(const oracle_model *m, const unsigned char *norm, const uint32_t *plen, const int32_t *pid,
                           size_t np, outv *o) {
  size_t consumed = 0;
  int prev_unk = 0;
  for (size_t k = 0; k < np; ++k) {
    const int is_unk = pid[k] == m->unk_id;
    if (is_unk && m->byte_fallback) {
      for (uint32_t i = 0; i < plen[k]; ++i) out_push(o, m->byte_to_id[norm[consumed + i]], (uint32_t)(consumed + i + 1));
    } else if (prev_unk && is_unk) {
      o->ends[o->n - 1] = (uint32_t)(consumed + plen[k]);
    } else {
      out_push(o, pid[k], (uint32_t)(consumed + plen[k]));
    }
    consumed += plen[k];
    prev_unk = is_unk;
  }
}

int oracle_nbest_encode(const oracle_model *m, const char *in, size_t len, int nbest_size, int32_t **ids_out,
                        uint32_t **cand_off_out, float **scores_out, size_t *k_out) {
  char *norm_; size_t nlen; uint64_t *n2o; size_t n2o_len;
  *ids_out = NULL; *cand_off_out = NULL; *scores_out = NULL; *k_out = 0;
  if (m->model_type != ORACLE_UNIGRAM) return 2; /* IsNBestEncodeAvailable, sentencepiece_processor.cc:662 */
  if (oracle_normalize(m, in, len, &norm_, &nlen, &n2o, &n2o_len)) return 1;
  free(n2o);
  const unsigned char *norm = (const unsigned char *)norm_;
  outv all = {0};
  uint32_t *cand_off = NULL; float *scores = NULL; size_t K = 0;
  /* Model::NBestEncode, unigram_model.cc:695-721 */
  if (nlen == 0) {                       /* {{{}, 0.0}} */
    cand_off = malloc(2 * sizeof(uint32_t)); scores = malloc(sizeof(float));
    cand_off[0] = cand_off[1] = 0; scores[0] = 0.f; K = 1;
    goto done;
  }
  if (nbest_size > 1024) nbest_size = 1024;
  if (nbest_size < 1) nbest_size = 1;
  if (nbest_size <= 1) {                 /* Encode(normalized), score 0.0 */
    int32_t *ids; uint32_t *ends; size_t n;
    unigram_encode(m, norm, nlen, &ids, &ends, &n);
    uint32_t *pl = malloc(sizeof(uint32_t) * (n ? n : 1));
    for (size_t i = 0; i < n; ++i) pl[i] = ends[i] - (i ? ends[i - 1] : 0);
    emit_candidate(m, norm, pl, ids, n, &all);
    cand_off = malloc(2 * sizeof(uint32_t)); scores = malloc(sizeof(float));
    cand_off[0] = 0; cand_off[1] = (uint32_t)all.n; scores[0] = 0.f; K = 1;
    free(ids); free(ends); free(pl);
    goto done;
  }
  {
    /* Lattice::SetSentence, :113-146: character positions */
    uint32_t *surf = malloc(sizeof(uint32_t) * (nlen + 2));
    int L = 0;
    for (size_t p = 0; p < nlen;) {
      size_t mb = one_char_len(norm[p]);
      if (mb > nlen - p) mb = nlen - p;
      surf[L++] = (uint32_t)p;
      p += mb;
    }
    surf[L] = (uint32_t)nlen;
    size_t ncap = 64, nn = 0;
    lnode *nodes = malloc(ncap * sizeof(lnode));
    ivec *begin_nodes = calloc((size_t)L + 1, sizeof(ivec)), *end_nodes = calloc((size_t)L + 1, sizeof(ivec));
#define NEW_NODE() (nn == ncap ? (nodes = realloc(nodes, (ncap *= 2) * sizeof(lnode)), &nodes[nn++]) : &nodes[nn++])
    { lnode *bos = NEW_NODE(); memset(bos, 0, sizeof *bos); bos->id = -1; bos->pos = 0; iv_push(&end_nodes[0], 0); }
    { lnode *eos = NEW_NODE(); memset(eos, 0, sizeof *eos); eos->id = -1; eos->pos = L; iv_push(&begin_nodes[L], 1); }
    /* Model::PopulateNodes, :547-596 */
    const float unk_score = m->min_score - 10.0f;
    for (int bp = 0; bp < L; ++bp) {
      int has_single = 0;
      uint32_t node = 0;
      for (size_t kpos = surf[bp]; kpos < nlen; ++kpos) {     /* commonPrefixSearch: results by length */
        const int32_t c = ht_child(&m->pieces, node, norm[kpos]);
        if (c < 0) break;
        node = (uint32_t)c;
        const int32_t id = m->pieces.value[node];
        if (id < 0) continue;
        const uint32_t endb = (uint32_t)kpos + 1;
        int length = 0;                                       /* get_chars_length, :548-552 */
        { int pos = bp; while (surf[pos] < endb) ++pos; length = pos - bp; }
        if (m->types[id] == ORACLE_UNUSED) continue;
        lnode *nd = NEW_NODE();
        memset(nd, 0, sizeof *nd);
        nd->pos = bp; nd->length = length; nd->id = id;
        nd->bbeg = surf[bp]; nd->bend = surf[bp + length];
        nd->score = m->types[id] == ORACLE_USER_DEFINED ? (float)((double)((float)length * m->max_score) - 0.1)
                                                         : m->scores[id];
        iv_push(&begin_nodes[bp], (int32_t)(nn - 1));
        iv_push(&end_nodes[bp + length], (int32_t)(nn - 1));
        if (!has_single && length == 1) has_single = 1;
      }
      if (!has_single) {
        lnode *nd = NEW_NODE();
        memset(nd, 0, sizeof *nd);
        nd->pos = bp; nd->length = 1; nd->id = m->unk_id; nd->score = unk_score;
        nd->bbeg = surf[bp]; nd->bend = surf[bp + 1];
        iv_push(&begin_nodes[bp], (int32_t)(nn - 1));
        iv_push(&end_nodes[bp + 1], (int32_t)(nn - 1));
      }
    }
    /* Lattice::Viterbi, :161-198 (only the backtrace scores are needed) */
    for (int pos = 0; pos <= L; ++pos)
      for (size_t r = 0; r < begin_nodes[pos].n; ++r) {
        lnode *rn = &nodes[begin_nodes[pos].a[r]];
        float best = 0.f; int have = 0;
        for (size_t q = 0; q < end_nodes[pos].n; ++q) {
          const float sc = nodes[end_nodes[pos].a[q]].backtrace_score + rn->score;
          if (!have || sc > best) { best = sc; have = 1; }
        }
        rn->backtrace_score = best;
      }
    /* Lattice::NBest, :345-509 (sample == false) */
    size_t pcap = 1024, pn = 0, hcap = 1024, hn = 0;
    hyp_t *pool = malloc(pcap * sizeof(hyp_t));
    int32_t *heap = malloc(hcap * sizeof(int32_t));
    pool[pn].node = 1; pool[pn].next = -1; pool[pn].gx = 0.f; pool[pn].fx = nodes[1].backtrace_score; pn++;
    heap_push(heap, &hn, pool, 0);
    cand_off = malloc(sizeof(uint32_t) * ((size_t)nbest_size + 1));
    scores = malloc(sizeof(float) * (size_t)nbest_size);
    cand_off[0] = 0;
    uint32_t *pl = malloc(sizeof(uint32_t) * ((size_t)L + 1));
    int32_t *pi = malloc(sizeof(int32_t) * ((size_t)L + 1));
    while (hn) {
      const int32_t top = heap_pop(heap, &hn, pool);
      const int32_t node = pool[top].node;
      if (node == 0) {                                         /* reached BOS */
        size_t np = 0;
        for (int32_t h = pool[top].next; pool[h].next != -1; h = pool[h].next) {
          pl[np] = nodes[pool[h].node].bend - nodes[pool[h].node].bbeg;
          pi[np] = nodes[pool[h].node].id;
          ++np;
        }
        emit_candidate(m, norm, pl, pi, np, &all);
        scores[K] = pool[top].fx;
        cand_off[++K] = (uint32_t)all.n;
        if (K == (size_t)nbest_size) break;
        continue;
      }
      const ivec *en = &end_nodes[nodes[node].pos];
      for (size_t q = 0; q < en->n; ++q) {
        const lnode *ln = &nodes[en->a[q]];
        if (pn == pcap) { pcap *= 2; pool = realloc(pool, pcap * sizeof(hyp_t)); }
        pool[pn].node = en->a[q];
        pool[pn].gx = ln->score + pool[top].gx;
        pool[pn].fx = ln->backtrace_score + pool[top].gx;
        pool[pn].next = top;
        if (hn == hcap) { hcap *= 2; heap = realloc(heap, hcap * sizeof(int32_t)); }
        heap_push(heap, &hn, pool, (int32_t)pn);
        pn++;
      }
      if (hn >= 10000) {                                       /* agenda shrink, :481-505 */
        int size = nbest_size * 10 < 512 ? nbest_size * 10 : 512;
        int32_t *keep = malloc(sizeof(int32_t) * (size_t)size);
        for (int i = 0; i < size; ++i) keep[i] = heap_pop(heap, &hn, pool);
        /* pushing in descending order into an empty heap leaves them in that order */
        hn = 0;
        for (int i = 0; i < size; ++i) heap_push(heap, &hn, pool, keep[i]);
        free(keep);
      }
    }
    free(pl); free(pi); free(pool); free(heap);
    for (int i = 0; i <= L; ++i) { free(begin_nodes[i].a); free(end_nodes[i].a); }
    free(begin_nodes); free(end_nodes); free(nodes); free(surf);
#undef NEW_NODE
  }
done:
  free(norm_);
  free(all.ends);
  *ids_out = all.ids ? all.ids : malloc(4);
  *cand_off_out = cand_off; *scores_out = scores; *k_out = K;
  return 0;
}

/* std::mt19937 (MT19937, 32-bit) */
void oracle_mt_seed(oracle_mt19937 *g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}
static uint32_t mt_next(oracle_mt19937 *g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      const uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7FFFFFFFu);
      g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11; y ^= (y << 7) & 0x9D2C5680u; y ^= (y << 15) & 0xEFC60000u; y ^= y >> 18;
  return y;
}

/* sentencepiece_processor.cc:703-718 + libstdc++ discrete_distribution (bits/random.tcc):
 * probabilities normalised by their sum, cumulative sums with the last forced to 1.0,
 * p = generate_canonical<double,53>(mt) = (x0 + x1 * 2^32) / 2^64 in double arithmetic,
 * result = lower_bound(cp, p). */
int oracle_sample_pick(oracle_mt19937 *g, const float *scores, size_t k, float alpha) {
  if (k < 2) return 0; /* _M_initialize clears the table: operator() returns 0 without drawing */
  double *lp = malloc(sizeof(double) * k), *cp = malloc(sizeof(double) * k);
  for (size_t i = 0; i < k; ++i) lp[i] = (double)(alpha * scores[i]); /* float product widened */
  double Z = lp[0];                                                   /* log_domain::LogSum, util.cc:278-294 */
  for (size_t i = 1; i < k; ++i) {
    double xa = Z, xb = lp[i];
    if (xa > xb) { const double t = xa; xa = xb; xb = t; }
    Z = xb + log1p(exp(xa - xb));
  }
  double sum = 0.0;
  for (size_t i = 0; i < k; ++i) { lp[i] = exp(lp[i] - Z); sum += lp[i]; }
  double run = 0.0;
  for (size_t i = 0; i < k; ++i) { run += lp[i] / sum; cp[i] = run; }
  cp[k - 1] = 1.0;
  const double x0 = (double)mt_next(g), x1 = (double)mt_next(g);
  double p = (x0 + x1 * 4294967296.0) / 18446744073709551616.0;
  if (p >= 1.0) p = nextafter(1.0, 0.0);
  size_t lo = 0, hi = k;                                              /* std::lower_bound */
  while (lo < hi) { const size_t mid = lo + (hi - lo) / 2; if (cp[mid] < p) lo = mid + 1; else hi = mid; }
  free(lp); free(cp);
  return (int)lo;
}

static int sample_lattice_batch(const oracle_model *m, const char *bytes, const uint64_t *offs, size_t n, float alpha,
                                uint32_t seed, int32_t **ids_out, uint64_t *id_offsets);
int oracle_sample_encode_batch(const oracle_model *m, const char *bytes, const uint64_t *offs, size_t n, int nbest_size,
                               float alpha, uint32_t seed, int32_t **ids_out, uint64_t *id_offsets) {
  if (nbest_size < 0) return sample_lattice_batch(m, bytes, offs, n, alpha, seed, ids_out, id_offsets);  /* :689-693 */
  oracle_mt19937 g;
  oracle_mt_seed(&g, seed);
  size_t cap = 1024, total = 0;
  int32_t *all = malloc(cap * sizeof(int32_t));
  for (size_t i = 0; i < n; ++i) {
    int32_t *ids; uint32_t *co; float *sc; size_t k;
    if (oracle_nbest_encode(m, bytes + offs[i], (size_t)(offs[i + 1] - offs[i]), nbest_size, &ids, &co, &sc, &k)) {
      free(all);
      return (int)(i + 1);
    }
    const int pick = oracle_sample_pick(&g, sc, k, alpha);
    const size_t cnt = co[pick + 1] - co[pick];
    id_offsets[i] = total;
    if (total + cnt > cap) { while (total + cnt > cap) cap *= 2; all = realloc(all, cap * sizeof(int32_t)); }
    if (cnt) memcpy(all + total, ids + co[pick], cnt * sizeof(int32_t));
    total += cnt;
    free(ids); free(co); free(sc);
  }
  id_offsets[n] = total;
  *ids_out = all;
  return 0;
}


/* =================================================== full-lattice operations ==
 * SampleEncode(nbest_size < 0) = forward-filtering / backward-sampling (unigram_model.cc:511-542,
 * dispatch sentencepiece_processor.cc:689-693), SampleEncodeAndScore with wor == false (:741-855) and
 * CalculateEntropy (:266-291, :857-864), all over Lattice::ForwardAlgorithm (:200-217) and
 * LogSumExp (:47-59). */
typedef struct {
  int L;                 /* characters */
  uint32_t *surf;        /* [L+1] byte offset of every character (+ the end) */
  lnode *nodes;          /* 0 = BOS, 1 = EOS, then PopulateNodes order */
  size_t nn;
  ivec *begin_nodes, *end_nodes;  /* [L+1] */
} lattice_t;

static void lattice_free(lattice_t *t) {
  for (int i = 0; i <= t->L; ++i) { free(t->begin_nodes[i].a); free(t->end_nodes[i].a); }
  free(t->begin_nodes); free(t->end_nodes); free(t->nodes); free(t->surf);
}

/* Lattice::SetSentence (:113-146) + Model::PopulateNodes (:547-596); nlen > 0 */
static void lattice_build(const oracle_model *m, const unsigned char *norm, size_t nlen, lattice_t *t) {
  uint32_t *surf = malloc(sizeof(uint32_t) * (nlen + 2));
  int L = 0;
  for (size_t p = 0; p < nlen;) {
    size_t mb = one_char_len(norm[p]);
    if (mb > nlen - p) mb = nlen - p;
    surf[L++] = (uint32_t)p;
    p += mb;
  }
  surf[L] = (uint32_t)nlen;
  size_t ncap = 64, nn = 0;
  lnode *nodes = malloc(ncap * sizeof(lnode));
  ivec *begin_nodes = calloc((size_t)L + 1, sizeof(ivec)), *end_nodes = calloc((size_t)L + 1, sizeof(ivec));
#define NEW_NODE() (nn == ncap ? (nodes = realloc(nodes, (ncap *= 2) * sizeof(lnode)), &nodes[nn++]) : &nodes[nn++])
  { lnode *bos = NEW_NODE(); memset(bos, 0, sizeof *bos); bos->id = -1; bos->pos = 0; iv_push(&end_nodes[0], 0); }
  { lnode *eos = NEW_NODE(); memset(eos, 0, sizeof *eos); eos->id = -1; eos->pos = L; iv_push(&begin_nodes[L], 1); }
  const float unk_score = m->min_score - 10.0f;
  for (int bp = 0; bp < L; ++bp) {
    int has_single = 0;
    uint32_t node = 0;
    for (size_t kpos = surf[bp]; kpos < nlen; ++kpos) {
      const int32_t c = ht_child(&m->pieces, node, norm[kpos]);
      if (c < 0) break;
      node = (uint32_t)c;
      const int32_t id = m->pieces.value[node];
      if (id < 0) continue;
      const uint32_t endb = (uint32_t)kpos + 1;
      int length = 0;
      { int pos = bp; while (surf[pos] < endb) ++pos; length = pos - bp; }
      if (m->types[id] == ORACLE_UNUSED) continue;
      lnode *nd = NEW_NODE();
      memset(nd, 0, sizeof *nd);
      nd->pos = bp; nd->length = length; nd->id = id;
      nd->bbeg = surf[bp]; nd->bend = surf[bp + length];
      nd->score = m->types[id] == ORACLE_USER_DEFINED ? (float)((double)((float)length * m->max_score) - 0.1)
                                                       : m->scores[id];
      iv_push(&begin_nodes[bp], (int32_t)(nn - 1));
      iv_push(&end_nodes[bp + length], (int32_t)(nn - 1));
      if (!has_single && length == 1) has_single = 1;
    }
    if (!has_single) {
      lnode *nd = NEW_NODE();
      memset(nd, 0, sizeof *nd);
      nd->pos = bp; nd->length = 1; nd->id = m->unk_id; nd->score = unk_score;
      nd->bbeg = surf[bp]; nd->bend = surf[bp + 1];
      iv_push(&begin_nodes[bp], (int32_t)(nn - 1));
      iv_push(&end_nodes[bp + 1], (int32_t)(nn - 1));
    }
  }
#undef NEW_NODE
  t->L = L; t->surf = surf; t->nodes = nodes; t->nn = nn; t->begin_nodes = begin_nodes; t->end_nodes = end_nodes;
}

/* unigram_model.cc:47-59 */
static float log_sum_exp(float x, float y, int init_mode) {
  if (init_mode) return y;
  const float vmin = x < y ? x : y;   /* std::min / std::max */
  const float vmax = x < y ? y : x;
  if (vmax > vmin + 50.f) return vmax;
  return (float)(vmax + log(exp((double)(vmin - vmax)) + 1.0));
}

/* Lattice::ForwardAlgorithm, :200-217; alpha is indexed by node */
static float *lattice_forward(const lattice_t *t, float inv_theta) {
  float *alpha = calloc(t->nn, sizeof(float));
  for (int pos = 0; pos <= t->L; ++pos)
    for (size_t r = 0; r < t->begin_nodes[pos].n; ++r) {
      const int32_t rn = t->begin_nodes[pos].a[r];
      for (size_t q = 0; q < t->end_nodes[pos].n; ++q) {
        const int32_t ln = t->end_nodes[pos].a[q];
        const float prod = inv_theta * t->nodes[ln].score;   /* float product, then float sum */
        alpha[rn] = log_sum_exp(alpha[rn], prod + alpha[ln], q == 0);
      }
    }
  return alpha;
}

/* libstdc++ std::discrete_distribution<int>(probs.begin(), probs.end()) + operator()(mt) on float probabilities */
static int discrete_pick(oracle_mt19937 *g, const float *probs, size_t k) {
  if (k < 2) return 0;
  double *cp = malloc(sizeof(double) * k);
  double sum = 0.0;
  for (size_t i = 0; i < k; ++i) sum += (double)probs[i];
  double run = 0.0;
  for (size_t i = 0; i < k; ++i) { run += (double)probs[i] / sum; cp[i] = run; }
  cp[k - 1] = 1.0;
  const double x0 = (double)mt_next(g), x1 = (double)mt_next(g);
  double p = (x0 + x1 * 4294967296.0) / 18446744073709551616.0;
  if (p >= 1.0) p = nextafter(1.0, 0.0);
  size_t lo = 0, hi = k;
  while (lo < hi) { const size_t mid = lo + (hi - lo) / 2; if (cp[mid] < p) lo = mid + 1; else hi = mid; }
  free(cp);
  return (int)lo;
}

/* Lattice::Sample, :511-542: node indices of the sampled path, left to right; returns the count */
static size_t lattice_sample(const lattice_t *t, const float *alpha, float inv_theta, oracle_mt19937 *g, int32_t *path) {
  size_t np = 0, pcap = 16;
  float *probs = malloc(sizeof(float) * pcap);
  float Z = alpha[1];
  int32_t node = 1;  /* EOS */
  for (;;) {
    const ivec *en = &t->end_nodes[t->nodes[node].pos];
    if (en->n > pcap) { pcap = en->n; probs = realloc(probs, sizeof(float) * pcap); }
    for (size_t q = 0; q < en->n; ++q) {
      const lnode *ln = &t->nodes[en->a[q]];
      const float prod = inv_theta * ln->score;
      const float arg = alpha[en->a[q]] + prod - Z;          /* float expression, widened for exp */
      probs[q] = (float)exp((double)arg);                     /* std::vector<float> probs */
    }
    node = en->a[discrete_pick(g, probs, en->n)];
    if (node == 0) break;
    Z = alpha[node];
    path[np++] = node;
  }
  free(probs);
  for (size_t i = 0; i < np / 2; ++i) { const int32_t x = path[i]; path[i] = path[np - 1 - i]; path[np - 1 - i] = x; }
  return np;
}

static void emit_path(const oracle_model *m, const unsigned char *norm, const lattice_t *t, const int32_t *path, size_t np,
                      outv *o) {
  uint32_t *pl = malloc(sizeof(uint32_t) * (np ? np : 1));
  int32_t *pi = malloc(sizeof(int32_t) * (np ? np : 1));
  for (size_t i = 0; i < np; ++i) {
    pl[i] = t->nodes[path[i]].bend - t->nodes[path[i]].bbeg;
    pi[i] = t->nodes[path[i]].id;
  }
  emit_candidate(m, norm, pl, pi, np, o);
  free(pl); free(pi);
}

/* SampleEncode(input, nbest_size < 0, alpha) over a packed batch in order on one generator. */
static int sample_lattice_batch(const oracle_model *m, const char *bytes, const uint64_t *offs, size_t n, float alpha,
                                uint32_t seed, int32_t **ids_out, uint64_t *id_offsets) {
  if (m->model_type != ORACLE_UNIGRAM) return -1;
  oracle_mt19937 g;
  oracle_mt_seed(&g, seed);
  outv all = {0};
  for (size_t i = 0; i < n; ++i) {
    char *norm_; size_t nlen; uint64_t *n2o; size_t n2o_len;
    id_offsets[i] = all.n;
    if (oracle_normalize(m, bytes + offs[i], (size_t)(offs[i + 1] - offs[i]), &norm_, &nlen, &n2o, &n2o_len)) { free(all.ids); free(all.ends); return (int)(i + 1); }
    free(n2o);
    if (nlen) {
      lattice_t t;
      lattice_build(m, (const unsigned char *)norm_, nlen, &t);
      float *a = lattice_forward(&t, alpha);
      int32_t *path = malloc(sizeof(int32_t) * ((size_t)t.L + 1));
      const size_t np = lattice_sample(&t, a, alpha, &g, path);
      emit_path(m, (const unsigned char *)norm_, &t, path, np, &all);
      free(path); free(a);
      lattice_free(&t);
    }
    free(norm_);
  }
  id_offsets[n] = all.n;
  free(all.ends);
  *ids_out = all.ids ? all.ids : malloc(4);
  return 0;
}

/* SentencePieceProcessor::CalculateEntropy (sentencepiece_processor.cc:747-760 -> unigram_model.cc:266-291,857-864) */
int oracle_entropy_batch(const oracle_model *m, const char *bytes, const uint64_t *offs, size_t n, float inv_theta,
                         float *entropy) {
  if (m->model_type != ORACLE_UNIGRAM) return -1;
  for (size_t i = 0; i < n; ++i) {
    char *norm_; size_t nlen; uint64_t *n2o; size_t n2o_len;
    if (oracle_normalize(m, bytes + offs[i], (size_t)(offs[i + 1] - offs[i]), &norm_, &nlen, &n2o, &n2o_len)) return (int)(i + 1);
    free(n2o);
    entropy[i] = -0.0f;  /* an empty lattice: H[EOS] = 0, returned negated */
    if (nlen) {
      lattice_t t;
      lattice_build(m, (const unsigned char *)norm_, nlen, &t);
      float *a = lattice_forward(&t, inv_theta);
      float *H = calloc(t.nn, sizeof(float));
      for (int pos = 0; pos <= t.L; ++pos)
        for (size_t r = 0; r < t.begin_nodes[pos].n; ++r) {
          const int32_t rn = t.begin_nodes[pos].a[r];
          for (size_t q = 0; q < t.end_nodes[pos].n; ++q) {
            const int32_t ln = t.end_nodes[pos].a[q];
            const float prod = inv_theta * t.nodes[ln].score;
            const float tp = prod + a[ln] - a[rn];
            H[rn] += expf(tp) * (H[ln] + tp);                 /* std::exp(float) */
          }
        }
      entropy[i] = -H[1];
      free(H); free(a);
      lattice_free(&t);
    }
    free(norm_);
  }
  return 0;
}

/* SampleEncodeAndScore(input, samples, alpha, wor = false, include_best = false) over a packed batch in order on
 * one generator (unigram_model.cc:741-855): `samples` independent Lattice::Sample draws per sentence, each scored
 * sum(alpha * node score) - log Z.  Outputs: ids of all samples packed, cand_off[n*samples+1], scores[n*samples]. */
int oracle_sample_score_batch(const oracle_model *m, const char *bytes, const uint64_t *offs, size_t n, int samples,
                              float inv_theta, uint32_t seed, int32_t **ids_out, uint64_t *cand_off, float *scores) {
  if (m->model_type != ORACLE_UNIGRAM || samples < 1) return -1;
  oracle_mt19937 g;
  oracle_mt_seed(&g, seed);
  outv all = {0};
  size_t c = 0;
  for (size_t i = 0; i < n; ++i) {
    char *norm_; size_t nlen; uint64_t *n2o; size_t n2o_len;
    if (oracle_normalize(m, bytes + offs[i], (size_t)(offs[i + 1] - offs[i]), &norm_, &nlen, &n2o, &n2o_len)) { free(all.ids); free(all.ends); return (int)(i + 1); }
    free(n2o);
    if (!nlen) {  /* the reference returns no result at all ("SampleEncodeAndScore returns empty result."): empty slots */
      for (int k = 0; k < samples; ++k) { cand_off[c] = all.n; scores[c] = 0.f; ++c; }
      free(norm_);
      continue;
    }
    lattice_t t;
    lattice_build(m, (const unsigned char *)norm_, nlen, &t);
    float *a = lattice_forward(&t, inv_theta);
    const float marginal = a[1];
    int32_t *path = malloc(sizeof(int32_t) * ((size_t)t.L + 1));
    for (int k = 0; k < samples; ++k) {
      /* every draw rebuilds the lattice and its alpha in the reference (:838-851); same values */
      const size_t np = lattice_sample(&t, a, inv_theta, &g, path);
      float score = 0.f;
      for (size_t j = 0; j < np; ++j) score += inv_theta * t.nodes[path[j]].score;
      cand_off[c] = all.n;
      emit_path(m, (const unsigned char *)norm_, &t, path, np, &all);
      scores[c] = score - marginal;
      ++c;
    }
    free(path); free(a);
    lattice_free(&t);
    free(norm_);
  }
  cand_off[c] = all.n;
  free(all.ends);
  *ids_out = all.ids ? all.ids : malloc(4);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * SentencePieceProcessor::Decode(const std::vector<int>&, SentencePieceText*) -- text only
 * (src/sentencepiece_processor.cc:765-925).
 * ---------------------------------------------------------------------------------------------- */
void oracle_set_unk_surface(oracle_model *m, const char *s, size_t len) {
  free(m->unk_surface);
  m->unk_surface = (char *)malloc(len + 1);
  memcpy(m->unk_surface, s, len);
  m->unk_surface[len] = 0;
  m->unk_surface_len = strlen(m->unk_surface); /* the reference takes c_str() (:772-773) */
}

/* PieceToByte, src/model_interface.cc:214-230: exactly "<0xXX>" with upper-case hex digits */
static int piece_to_byte(const char *p, size_t len) {
  if (len != 6 || p[0] != '<' || p[1] != '0' || p[2] != 'x' || p[5] != '>') return -1;
  int v = 0;
  for (int i = 3; i < 5; ++i) {
    const char ch = p[i];
    int d;
    if (ch >= '0' && ch <= '9') d = ch - '0';
    else if (ch >= 'A' && ch <= 'F') d = ch - 'A' + 10;
    else return -1;
    v = v * 16 + d;
  }
  return v;
}

typedef struct { char *p; size_t n, cap; } obuf;
static void obuf_put(obuf *b, const char *s, size_t len) {
  if (b->n + len + 1 > b->cap) {
    b->cap = (b->n + len + 1) * 2 + 64;
    b->p = (char *)realloc(b->p, b->cap);
  }
  memcpy(b->p + b->n, s, len);
  b->n += len;
}

int oracle_decode_ids(const oracle_model *m, const int32_t *ids, size_t n, char **text_out, size_t *text_len) {
  static const char kDefaultUnk[] = " \xE2\x81\x87 ";    /* kDefaultUnknownSymbol, :52 */
  static const char kReplacement[] = "\xEF\xBF\xBD";     /* kReplacementCharacter, :55 */
  static const char kSpace[] = "\xE2\x96\x81";           /* kSpaceSymbol */
  const char *unk_surface = m->unk_surface ? m->unk_surface : kDefaultUnk;
  const size_t unk_len = m->unk_surface ? m->unk_surface_len : sizeof(kDefaultUnk) - 1;
  obuf text = {0, 0, 0};
  obuf_put(&text, "", 0);
  for (size_t i = 0; i < n; ++i)
    if (ids[i] < 0 || ids[i] >= m->vocab_size) { free(text.p); return 1; }   /* :915-918 */
  unsigned char *bytes = (unsigned char *)malloc(n + 1);
  size_t nbytes = 0;        /* pending run of BYTE pieces (:836-876, flushed before the next other piece) */
  int is_bos_ws = 1, bos_ws_seen = 0;  /* :879-880 */
  int rc = 0;
  for (size_t i = 0; i <= n && !rc; ++i) {
    const int is_byte = i < n && m->types[ids[i]] == ORACLE_BYTE;
    if (is_byte) {
      const int32_t id = ids[i];
      const int b = piece_to_byte(m->piece_bytes + m->piece_off[id], m->piece_off[id + 1] - m->piece_off[id]);
      if (b < 0) { rc = 2; break; }
      bytes[nbytes++] = (unsigned char)b;
      continue;
    }
    /* ProcessBytePieces: one Unicode character at a time; an invalid byte becomes U+FFFD */
    size_t off = 0;
    while (off < nbytes) {
      size_t consumed = 0;
      if (!is_valid_decode_utf8(bytes + off, nbytes - off, &consumed)) {
        obuf_put(&text, kReplacement, 3);
        consumed = 1;
      } else {
        obuf_put(&text, (const char *)bytes + off, consumed);
      }
      off += consumed;
    }
    nbytes = 0;
    if (i == n) break;
    if (bos_ws_seen || text.n != 0) is_bos_ws = 0;   /* :887 */
    const int32_t id = ids[i];
    const char *piece = m->piece_bytes + m->piece_off[id];
    size_t plen = m->piece_off[id + 1] - m->piece_off[id];
    const uint8_t type = m->types[id];
    /* DecodeSentencePiece (:779-812) */
    if (type == ORACLE_CONTROL) { bos_ws_seen = 0; continue; }
    if (type == ORACLE_UNKNOWN) { obuf_put(&text, unk_surface, unk_len); bos_ws_seen = 0; continue; }
    int has_bos_ws = 0;
    if (is_bos_ws && (m->add_dummy_prefix || m->remove_extra_whitespaces)) {
      if (plen >= 3 && memcmp(piece, kSpace, 3) == 0) { piece += 3; plen -= 3; has_bos_ws = 1; }
      if (m->remove_extra_whitespaces) has_bos_ws = 0;
    }
    for (size_t k = 0; k < plen;) {   /* StrReplaceAll(piece, {{kSpaceSymbol, " "}}) */
      if (k + 3 <= plen && memcmp(piece + k, kSpace, 3) == 0) { obuf_put(&text, " ", 1); k += 3; }
      else { obuf_put(&text, piece + k, 1); k += 1; }
    }
    bos_ws_seen = has_bos_ws;
  }
  free(bytes);
  if (rc) { free(text.p); return rc; }
  text.p[text.n] = 0;
  *text_out = text.p;
  *text_len = text.n;
  return 0;
}
