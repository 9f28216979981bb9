/* tools/corpus_gen.c -- seeded synthetic corpora for bench.py and the parity tests.
 *
 * BASELINE.json's configs 2-5 are quoted on synthetic corpora ("1M synthetic
 * ~128-byte English sentences", "mixed CJK/emoji synthetic corpus"); SURVEY.md
 * section 8(d) fixes their shape.  This generator produces them deterministically
 * from two small committed frequency tables (tests/golden/en_wordlist.tsv,
 * tests/golden/ja_charlist.tsv) so that the dev container and the GPU box see
 * byte-identical inputs without shipping 128 MB files.
 *
 * Sentence i of a corpus depends only on (kind, seed, i): ranks of a multi-GPU
 * run generate their own shard, and any prefix/sub-range of a corpus is itself
 * reproducible.
 *
 *   kind 0  "en"    English-like: Zipf(1.0) over ~150k word forms (botchan word
 *                   types + affixed / compounded derivations), punctuation,
 *                   digits, capitalisation noise; target length U[96,160] bytes.
 *   kind 1  "mixed" ASCII + full-width / compatibility forms (NFKC rules), kana,
 *                   kanji, half-width katakana + voiced marks, combining marks,
 *                   emoji (4-byte, mostly unknown -> byte fallback), U+3000 and
 *                   multi-space runs, tabs, and a sprinkle of malformed UTF-8.
 *
 * This is data plumbing, not part of the encode engine.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_WORD 48

typedef struct {
  char (*w)[MAX_WORD];
  uint8_t *len;
  uint32_t n;
} wordtab;

typedef struct {
  wordtab vocab;       /* ranked: base words by frequency, then derived forms */
  uint32_t n_base;
  char (*ja)[8];       /* UTF-8 chars */
  uint8_t *ja_len;
  double *ja_cdf;
  uint32_t n_ja;
} corpus_gen;

static inline uint64_t splitmix64(uint64_t *s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline double u01(uint64_t *s) { return (splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t below(uint64_t *s, uint32_t n) { return (uint32_t)(splitmix64(s) % n); }

/* exp/log without libm dependency surprises: use the C library (deterministic on
 * one libm; both boxes share the image). */
#include <math.h>

static const char *PREFIX[] = {"un", "re", "pre", "dis", "over", "under", "out", "mis", "non",
                               "inter", "sub", "super", "anti", "de", "co", "semi", "fore", "counter"};
static const char *SUFFIX[] = {"s", "ed", "ing", "er", "est", "ly", "ness", "ment", "tion", "able",
                               "ful", "less", "ish", "ize", "ized", "izing", "ism", "ist", "ity",
                               "ous", "ive", "al", "ship", "hood", "ward", "wise"};
#define NPRE (sizeof(PREFIX) / sizeof(PREFIX[0]))
#define NSUF (sizeof(SUFFIX) / sizeof(SUFFIX[0]))

static void tab_push(wordtab *t, uint32_t *cap, const char *s, size_t l) {
  if (l >= MAX_WORD) l = MAX_WORD - 1;
  if (t->n == *cap) {
    *cap = *cap ? *cap * 2 : 4096;
    t->w = realloc(t->w, (size_t)*cap * MAX_WORD);
    t->len = realloc(t->len, *cap);
  }
  memcpy(t->w[t->n], s, l);
  t->w[t->n][l] = 0;
  t->len[t->n] = (uint8_t)l;
  t->n++;
}

corpus_gen *corpus_gen_create(const char *en_path, const char *ja_path) {
  corpus_gen *g = calloc(1, sizeof(*g));
  uint32_t cap = 0;
  FILE *f = fopen(en_path, "rb");
  if (!f) { free(g); return NULL; }
  char line[256];
  while (fgets(line, sizeof line, f)) {
    char *tab = strchr(line, '\t');
    if (!tab) continue;
    tab_push(&g->vocab, &cap, line, (size_t)(tab - line));
  }
  fclose(f);
  g->n_base = g->vocab.n;
  /* ~150k derived forms, order fixed by a constant-seeded PRNG */
  uint64_t s = 0x5eed5eed12345678ull;
  const uint32_t n_derived = 150000;
  char buf[3 * MAX_WORD];
  for (uint32_t i = 0; i < n_derived; ++i) {
    /* bias the stems towards frequent base words */
    uint32_t a = (uint32_t)(exp(u01(&s) * log((double)g->n_base)) - 1.0);
    if (a >= g->n_base) a = g->n_base - 1;
    const uint32_t mode = below(&s, 10);
    size_t l = 0;
    if (mode < 4) { /* stem + suffix */
      l = (size_t)snprintf(buf, sizeof buf, "%s%s", g->vocab.w[a], SUFFIX[below(&s, NSUF)]);
    } else if (mode < 6) { /* prefix + stem */
      l = (size_t)snprintf(buf, sizeof buf, "%s%s", PREFIX[below(&s, NPRE)], g->vocab.w[a]);
    } else if (mode < 7) { /* prefix + stem + suffix */
      l = (size_t)snprintf(buf, sizeof buf, "%s%s%s", PREFIX[below(&s, NPRE)], g->vocab.w[a],
                           SUFFIX[below(&s, NSUF)]);
    } else { /* compound */
      uint32_t b = (uint32_t)(exp(u01(&s) * log((double)g->n_base)) - 1.0);
      if (b >= g->n_base) b = g->n_base - 1;
      l = (size_t)snprintf(buf, sizeof buf, "%s%s%s", g->vocab.w[a], below(&s, 4) == 0 ? "-" : "",
                           g->vocab.w[b]);
    }
    /* derived forms are lower-cased at the front ("I" + "ness" etc.) */
    if (buf[0] >= 'A' && buf[0] <= 'Z') buf[0] = (char)(buf[0] - 'A' + 'a');
    tab_push(&g->vocab, &cap, buf, l);
  }
  if (ja_path) {
    f = fopen(ja_path, "rb");
    if (f) {
      uint32_t jcap = 4096;
      g->ja = malloc((size_t)jcap * 8);
      g->ja_len = malloc(jcap);
      g->ja_cdf = malloc(sizeof(double) * jcap);
      double tot = 0;
      while (fgets(line, sizeof line, f) && g->n_ja < jcap) {
        char *tab = strchr(line, '\t');
        if (!tab || tab - line > 4 || tab == line) continue;
        memcpy(g->ja[g->n_ja], line, (size_t)(tab - line));
        g->ja_len[g->n_ja] = (uint8_t)(tab - line);
        tot += atof(tab + 1);
        g->ja_cdf[g->n_ja] = tot;
        g->n_ja++;
      }
      for (uint32_t i = 0; i < g->n_ja; ++i) g->ja_cdf[i] /= tot;
      fclose(f);
    }
  }
  return g;
}

void corpus_gen_destroy(corpus_gen *g) {
  if (!g) return;
  free(g->vocab.w); free(g->vocab.len); free(g->ja); free(g->ja_len); free(g->ja_cdf); free(g);
}

static inline uint32_t zipf_rank(corpus_gen *g, uint64_t *s) {
  /* Zipf(1.0): P(rank<=r) ~ ln(r+1)/ln(V+1)  ->  r = (V+1)^u - 1 */
  uint32_t r = (uint32_t)(exp(u01(s) * log((double)g->vocab.n + 1.0)) - 1.0);
  return r >= g->vocab.n ? g->vocab.n - 1 : r;
}

static size_t put_utf8(char *o, uint32_t c) {
  if (c < 0x80) { o[0] = (char)c; return 1; }
  if (c < 0x800) { o[0] = (char)(0xC0 | (c >> 6)); o[1] = (char)(0x80 | (c & 0x3F)); return 2; }
  if (c < 0x10000) {
    o[0] = (char)(0xE0 | (c >> 12)); o[1] = (char)(0x80 | ((c >> 6) & 0x3F)); o[2] = (char)(0x80 | (c & 0x3F));
    return 3;
  }
  o[0] = (char)(0xF0 | (c >> 18)); o[1] = (char)(0x80 | ((c >> 12) & 0x3F));
  o[2] = (char)(0x80 | ((c >> 6) & 0x3F)); o[3] = (char)(0x80 | (c & 0x3F));
  return 4;
}

static size_t gen_en(corpus_gen *g, uint64_t s, char *o) {
  const size_t target = 96 + below(&s, 65);
  size_t n = 0;
  int first = 1;
  while (n < target) {
    if (!first) {
      o[n++] = ' ';
      if (below(&s, 200) == 0) o[n++] = ' '; /* stray double space */
    }
    const uint32_t kind = below(&s, 100);
    if (kind < 3) { /* number */
      const uint32_t digits = 1 + below(&s, 4);
      for (uint32_t d = 0; d < digits; ++d) o[n++] = (char)('0' + below(&s, 10));
      if (below(&s, 4) == 0) { o[n++] = '.'; o[n++] = (char)('0' + below(&s, 10)); }
    } else {
      const uint32_t r = zipf_rank(g, &s);
      const size_t l = g->vocab.len[r];
      const int quote = below(&s, 60) == 0;
      if (quote) o[n++] = '"';
      memcpy(o + n, g->vocab.w[r], l);
      if ((first || below(&s, 33) == 0) && o[n] >= 'a' && o[n] <= 'z') o[n] = (char)(o[n] - 'a' + 'A');
      if (below(&s, 150) == 0) /* SHOUTING */
        for (size_t k = 0; k < l; ++k) if (o[n + k] >= 'a' && o[n + k] <= 'z') o[n + k] = (char)(o[n + k] - 32);
      n += l;
      if (quote) o[n++] = '"';
    }
    const uint32_t p = below(&s, 100);
    if (p < 8) o[n++] = ',';
    else if (p < 9) o[n++] = ';';
    else if (p < 10) o[n++] = ':';
    first = 0;
  }
  const uint32_t e = below(&s, 100);
  if (e < 80) o[n++] = '.';
  else if (e < 88) o[n++] = '?';
  else if (e < 94) o[n++] = '!';
  return n;
}

static const uint32_t COMPAT[] = {0x337F /*㍿*/, 0x2460 /*①*/, 0x2461, 0x3231 /*㈱*/, 0x33A1 /*㎡*/,
                                  0xFB01 /*ﬁ*/, 0x2122 /*™*/, 0x00BD /*½*/, 0x2167 /*Ⅷ*/, 0x3392 /*㎒*/,
                                  0xFDFA /* long arabic ligature */, 0x2026 /*…*/, 0x00A0 /*nbsp*/,
                                  0x200B /*zwsp*/, 0x2581 /* literal ▁ */, 0xFFFD, 0x00E9, 0x00FC};
#define NCOMPAT (sizeof(COMPAT) / sizeof(COMPAT[0]))

static size_t gen_mixed(corpus_gen *g, uint64_t s, char *o) {
  const size_t target = 64 + below(&s, 129);
  size_t n = 0;
  if (below(&s, 20) == 0) { o[n++] = ' '; if (below(&s, 2)) o[n++] = ' '; }  /* leading spaces */
  if (below(&s, 40) == 0) n += put_utf8(o + n, 0x3000);
  while (n < target) {
    const uint32_t kind = below(&s, 100);
    if (kind < 30) { /* english word */
      const uint32_t r = zipf_rank(g, &s);
      memcpy(o + n, g->vocab.w[r], g->vocab.len[r]);
      n += g->vocab.len[r];
    } else if (kind < 60 && g->n_ja) { /* kana / kanji run */
      const uint32_t len = 2 + below(&s, 11);
      for (uint32_t k = 0; k < len; ++k) {
        const double u = u01(&s);
        uint32_t lo = 0, hi = g->n_ja - 1;
        while (lo < hi) { uint32_t mid = (lo + hi) / 2; if (g->ja_cdf[mid] < u) lo = mid + 1; else hi = mid; }
        memcpy(o + n, g->ja[lo], g->ja_len[lo]);
        n += g->ja_len[lo];
      }
    } else if (kind < 68) { /* full-width latin word */
      const uint32_t r = zipf_rank(g, &s);
      for (size_t k = 0; k < g->vocab.len[r]; ++k) {
        const unsigned char c = (unsigned char)g->vocab.w[r][k];
        if (c > 0x20 && c < 0x7F) n += put_utf8(o + n, 0xFF00u + (c - 0x20));
        else o[n++] = (char)c;
      }
    } else if (kind < 74) { /* half-width katakana, sometimes + (semi)voiced mark */
      const uint32_t len = 1 + below(&s, 5);
      for (uint32_t k = 0; k < len; ++k) {
        n += put_utf8(o + n, 0xFF66u + below(&s, 0x38));
        const uint32_t m = below(&s, 6);
        if (m == 0) n += put_utf8(o + n, 0xFF9E);
        else if (m == 1) n += put_utf8(o + n, 0xFF9F);
      }
    } else if (kind < 80) { /* compatibility / odd chars */
      n += put_utf8(o + n, COMPAT[below(&s, NCOMPAT)]);
    } else if (kind < 86) { /* emoji run */
      const uint32_t len = 1 + below(&s, 3);
      for (uint32_t k = 0; k < len; ++k)
        n += put_utf8(o + n, below(&s, 2) ? 0x1F600u + below(&s, 0x50) : 0x1F300u + below(&s, 0x100));
    } else if (kind < 89) { /* base letter + combining mark (composition rules) */
      o[n++] = (char)("aeiounAEO"[below(&s, 9)]);
      n += put_utf8(o + n, 0x0300u + below(&s, 5));
    } else if (kind < 92) { /* digits */
      const uint32_t len = 1 + below(&s, 5);
      for (uint32_t k = 0; k < len; ++k) o[n++] = (char)('0' + below(&s, 10));
    } else if (kind < 94) { /* hangul syllables */
      const uint32_t len = 1 + below(&s, 4);
      for (uint32_t k = 0; k < len; ++k) n += put_utf8(o + n, 0xAC00u + below(&s, 11172));
    } else if (kind < 96) { /* malformed UTF-8 */
      const uint32_t m = below(&s, 5);
      if (m == 0) o[n++] = (char)(0x80 + below(&s, 0x40));          /* stray continuation */
      else if (m == 1) { o[n++] = (char)0xE3; o[n++] = (char)0x81; } /* truncated 3-byte */
      else if (m == 2) { o[n++] = (char)0xC0; o[n++] = (char)0xAF; } /* overlong */
      else if (m == 3) { o[n++] = (char)0xED; o[n++] = (char)0xA0; o[n++] = (char)0x80; } /* surrogate */
      else { o[n++] = (char)0xF8 + (char)below(&s, 8); }             /* invalid lead */
    } else if (kind < 98) { /* control chars the charsmap maps to "" or " " */
      o[n++] = (char)(below(&s, 2) ? '\t' : (1 + below(&s, 8)));
    } else { /* punctuation cluster */
      static const char P[] = ".,!?;:()[]-_/\\'\"@#$%&*+=<>~";
      o[n++] = P[below(&s, sizeof(P) - 1)];
    }
    /* separator */
    const uint32_t sep = below(&s, 100);
    if (sep < 55) o[n++] = ' ';
    else if (sep < 60) { o[n++] = ' '; o[n++] = ' '; if (below(&s, 3) == 0) o[n++] = ' '; }
    else if (sep < 64) n += put_utf8(o + n, 0x3000);
    /* else: no separator (CJK style) */
  }
  if (below(&s, 10) == 0) { o[n++] = ' '; if (below(&s, 2)) n += put_utf8(o + n, 0x3000); } /* trailing */
  return n;
}

/* Upper bound on one sentence's bytes for buffer sizing. */
uint64_t corpus_gen_max_sentence_bytes(int kind) { return kind == 0 ? 320 : 512; }

/* Fills `buf` with sentences [first, first+n) of corpus (kind, seed); offsets[n+1].
 * Returns total bytes written, or (uint64_t)-1 if `cap` is too small. */
uint64_t corpus_gen_fill(corpus_gen *g, int kind, uint64_t seed, uint64_t first, uint64_t n,
                         char *buf, uint64_t cap, uint64_t *offsets) {
  uint64_t pos = 0;
  const uint64_t maxs = corpus_gen_max_sentence_bytes(kind);
  for (uint64_t i = 0; i < n; ++i) {
    if (pos + maxs > cap) return (uint64_t)-1;
    /* hash (kind, seed, index) into an unrelated start state: plain state offsets
     * would make sentence i+1's stream a one-step shift of sentence i's */
    uint64_t t = (first + i) * 0xD1342543DE82EF95ull ^ (seed + 0x632BE59BD9B4E019ull) * 0xA0761D6478BD642Full ^
                 ((uint64_t)kind << 56);
    uint64_t s = splitmix64(&t);
    s ^= splitmix64(&t) >> 7;
    offsets[i] = pos;
    pos += kind == 0 ? gen_en(g, s, buf + pos) : gen_mixed(g, s, buf + pos);
  }
  offsets[n] = pos;
  return pos;
}
