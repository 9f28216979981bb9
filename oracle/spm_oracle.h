/* oracle/spm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's batched-encode hot path
 * (SURVEY.md section 8a rows a1-a6, a9-a12).  Used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline leg as the CHECKER of
 * the CUDA engine.  Nothing under sentencepiece_b200/ may include, link or load
 * this; the product fails loudly when its CUDA library is missing instead of
 * falling back here.
 *
 * Parity status: PINNED -- see oracle/spm_oracle.c header.
 */
#ifndef SPM_ORACLE_H_
#define SPM_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ModelProto::SentencePiece::Type, src/sentencepiece_model.proto:296-304 */
enum { ORACLE_NORMAL = 1, ORACLE_UNKNOWN = 2, ORACLE_CONTROL = 3, ORACLE_USER_DEFINED = 4,
       ORACLE_UNUSED = 5, ORACLE_BYTE = 6 };
/* TrainerSpec::ModelType, src/sentencepiece_model.proto:48-53 */
enum { ORACLE_UNIGRAM = 1, ORACLE_BPE = 2 };

typedef struct {
  int32_t model_type;
  int32_t vocab_size;
  const char *piece_bytes;       /* concatenated piece strings */
  const uint32_t *piece_off;     /* [vocab_size+1] */
  const float *scores;           /* [vocab_size] */
  const uint8_t *types;          /* [vocab_size] */
  uint8_t byte_fallback, add_dummy_prefix, remove_extra_whitespaces, escape_whitespaces,
      treat_whitespace_as_suffix;
  const uint8_t *charsmap;       /* NormalizerSpec.precompiled_charsmap verbatim */
  size_t charsmap_len;
} oracle_model_desc;

typedef struct oracle_model oracle_model;

oracle_model *oracle_create(const oracle_model_desc *d, char *err, size_t errlen);
void oracle_destroy(oracle_model *m);
/* live piece types (SetVocabulary / ResetVocabulary): sentencepiece_processor.cc:301-340 */
void oracle_set_types(oracle_model *m, const uint8_t *types);
float oracle_min_score(const oracle_model *m);
float oracle_max_score(const oracle_model *m);
int32_t oracle_unk_id(const oracle_model *m);

/* Normalizer::Normalize (src/normalizer.cc:71-186).  *out / *n2o are malloc'ed
 * (free with oracle_free).  n2o has out_len+1 entries unless out_len == 0, in
 * which case it has 0 entries (the reference returns both empty). */
int oracle_normalize(const oracle_model *m, const char *in, size_t len, char **out, size_t *out_len,
                     uint64_t **n2o, size_t *n2o_len);

/* Model::Encode on already-normalized text (unigram: EncodeOptimized
 * src/unigram_model.cc:889-1020; BPE: SampleEncode alpha=0 src/bpe_model.cc:38-203).
 * Outputs malloc'ed arrays: piece end offsets (exclusive, in `norm`) and ids. */
int oracle_model_encode(const oracle_model *m, const char *norm, size_t len, int32_t **ids,
                        uint32_t **ends, size_t *n);

/* SentencePieceProcessor::Encode id path: Normalize -> model Encode ->
 * PopulateSentencePieceText (src/sentencepiece_processor.cc:547-651): unk-run
 * merging, byte-fallback expansion.  Outputs (malloc'ed):
 *   ids[n]       token ids
 *   tok_end[n]   exclusive end offset of each output token in the normalized text
 *                (tokens partition the normalized text; a byte-fallback piece covers
 *                 exactly one byte)
 * Returns 0, or non-zero when the reference would return a non-OK Status. */
int oracle_encode(const oracle_model *m, const char *in, size_t len, int32_t **ids,
                  uint32_t **tok_end, size_t *n);

/* Batch convenience over a packed buffer; id_offsets[n+1]; *ids malloc'ed. */
int oracle_encode_batch(const oracle_model *m, const char *bytes, const uint64_t *offs, size_t n,
                        int32_t **ids, uint64_t *id_offsets);

/* ---- config 5: n-best + sampling ----
 * SentencePieceProcessor::NBestEncode (src/sentencepiece_processor.cc:653-676) =
 * Normalize -> unigram::Model::NBestEncode (src/unigram_model.cc:695-721: Lattice::SetSentence
 * :113-146, Model::PopulateNodes :547-596, Lattice::Viterbi :161-198, Lattice::NBest :345-509 with
 * libstdc++'s heap order) -> PopulateSentencePieceText per candidate.
 * Outputs (malloc'ed): ids of all candidates packed, cand_off[k+1], scores[k]; *k = candidates. */
int oracle_nbest_encode(const oracle_model *m, const char *in, size_t len, int nbest_size, int32_t **ids,
                        uint32_t **cand_off, float **scores, size_t *k);

/* std::mt19937 + std::discrete_distribution<int> exactly as libstdc++ runs them in
 * SentencePieceProcessor::SampleEncode (src/sentencepiece_processor.cc:699-719). */
typedef struct { uint32_t mt[624]; int idx; } oracle_mt19937;
void oracle_mt_seed(oracle_mt19937 *g, uint32_t seed);
/* index drawn for candidate scores[k] and `alpha`; consumes two 32-bit draws iff k >= 2 */
int oracle_sample_pick(oracle_mt19937 *g, const float *scores, size_t k, float alpha);

/* SampleEncode(input, nbest_size > 1, alpha) over a packed batch in order on one generator. */
int oracle_sample_encode_batch(const oracle_model *m, const char *bytes, const uint64_t *offs, size_t n, int nbest_size,
                               float alpha, uint32_t seed, int32_t **ids, uint64_t *id_offsets);

/* ---- next row: full-lattice operations of unigram models.
 * oracle_sample_encode_batch with nbest_size < 0 is SampleEncode's forward-filtering / backward-sampling branch
 * (sentencepiece_processor.cc:689-693 -> unigram_model.cc:511-542).  CalculateEntropy per sentence
 * (sentencepiece_processor.cc:747-760 -> unigram_model.cc:266-291) and SampleEncodeAndScore(wor = false,
 * include_best = false) (unigram_model.cc:741-855): `samples` draws per sentence on one generator across the batch,
 * candidate c of sentence i = ids[cand_off[i*samples+c] .. cand_off[i*samples+c+1]) with score scores[i*samples+c]. */
int oracle_entropy_batch(const oracle_model *m, const char *bytes, const uint64_t *offs, size_t n, float inv_theta,
                         float *entropy);
int oracle_sample_score_batch(const oracle_model *m, const char *bytes, const uint64_t *offs, size_t n, int samples,
                              float inv_theta, uint32_t seed, int32_t **ids, uint64_t *cand_off, float *scores);

/* ---- next row: SentencePieceProcessor::Decode(ids) -> text
 * (src/sentencepiece_processor.cc:765-925): IdToPiece, CONTROL pieces invisible, UNKNOWN -> unk_surface, the first
 * U+2581 stripped while the text is still empty (add_dummy_prefix / remove_extra_whitespaces), U+2581 -> ' ', runs of
 * BYTE pieces reassembled into UTF-8 with U+FFFD for every structurally invalid byte.  No denormalizer, no
 * decode_extra_options.  *text is malloc'ed.  Returns 0, 1 for an id out of range (kOutOfRange), 2 for a BYTE piece
 * that is not "<0xXX>". */
void oracle_set_unk_surface(oracle_model *m, const char *s, size_t len); /* TrainerSpec.unk_surface, default " \xE2\x81\x87 " */
int oracle_decode_ids(const oracle_model *m, const int32_t *ids, size_t n, char **text, size_t *text_len);

void oracle_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
