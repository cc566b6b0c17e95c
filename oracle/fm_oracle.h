/*
 * fm_oracle.h -- CPU ORACLE for the sview-fmindex hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's (baku4/sview-fmindex, Rust) blob
 * builder, blob loader and count/locate algorithm.  It exists so that the CUDA path
 * can be checked bit-for-bit on identical blobs and patterns.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call it.  The product (sview_fmindex_b200/libsvfm.so) never links or loads it.
 *
 * Parity status: the reference is Rust and no Rust toolchain exists in this image, so
 * the reference itself cannot be run here.  The oracle is pinned by the reference's
 * only hard-coded golden vectors (sview-fmindex/src/tests/readme/mod.rs:15,33,37,44),
 * by an independent brute-force matcher over the encoded text (stand-in for the
 * `fm-index 0.1` crate used at src/tests/result_answer/other_crate.rs:7-19) and by the
 * reference's invariance properties (src/tests/config_invariance/mod.rs:93-107,
 * src/tests/text_encoders_consistency/mod.rs:86-106).  Blob BYTE parity with the real
 * Rust builder is unpinned (no reference test or fixture checks bytes).
 *
 * All citations are relative to /root/reference/sview-fmindex/src/.
 */
#ifndef FM_ORACLE_H
#define FM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The (P, B, E) type triple of FmIndex<'a, P, B, E> (lib.rs:15).  It is NOT stored in
 * the blob; the caller supplies it exactly as the Rust caller picks the generic args. */
typedef struct ora_type {
    uint32_t pos_bits;  /* Position: 32 | 64            (text_length.rs:45,87)          */
    uint32_t planes;    /* BlockN:   2..6               (components/bwm/blocks/mod.rs)  */
    uint32_t vec_bits;  /* Vector:   32 | 64 | 128      (blocks/vector.rs:35-79)        */
    uint32_t encoder;   /* 0 = PassThrough, 1 = EncodingTable (text_encoder/)           */
} ora_type;

enum {
    ORA_OK = 0,
    /* LoadError (load_from_blob.rs:16-24) */
    ORA_ERR_INVALID_FORMAT = 1,
    ORA_ERR_BLOB_SIZE = 2,         /* detail[0] = expected total, detail[1] = actual total */
    /* BuildError (builder/mod.rs:37-57) */
    ORA_ERR_SYMBOL_COUNT_OVER = 10, /* detail[0] = MAX_SYMBOL, detail[1] = symbol_count */
    ORA_ERR_TEXT_LENGTH = 11,
    ORA_ERR_INVALID_BLOB_SIZE = 12,
    ORA_ERR_NOT_ALIGNED = 13,
    ORA_ERR_INVALID_CONFIG = 14,
    /* restatement-only */
    ORA_ERR_BAD_TYPE = 20,
    ORA_ERR_EMPTY_PATTERN = 21,     /* the reference panics (count_array.rs:211) */
    ORA_ERR_TOO_LARGE = 22,         /* oracle suffix sorter is limited to n < 2^31-1 */
    ORA_ERR_NOMEM = 23
} ;

/* Resolved byte offsets of every section of a blob. */
typedef struct ora_layout {
    uint64_t align;                 /* B::ALIGN_SIZE: 8, or 16 for u128 vectors */
    uint64_t off_encoder;           /* EncodingTable (256 B) or PassThrough (0 B) */
    uint64_t off_count_header, off_sa_header, off_bwm_header;
    uint64_t header_size;
    uint64_t off_count_array, off_kmer_multiplier, off_kmer_count_table;
    uint64_t off_suffix_array, off_sentinel_index, off_rank_checkpoints, off_blocks;
    uint64_t total_size;
    /* header fields */
    uint32_t symbol_count, kmer_size, count_array_len, kmer_multiplier_len;
    uint64_t kmer_count_table_len;
    uint32_t sampling_ratio;
    uint64_t suffix_array_len;
    uint64_t rank_checkpoints_len, blocks_len;
} ora_layout;

/* EncodingTable::from_symbols / from_symbols_with_wildcard (encoding_table.rs:15-34).
 * symbols: n_groups byte strings, concatenated in `bytes`, group g = bytes[offs[g]..offs[g+1]).
 * Returns the symbol_count (encoding_table.rs:35-37). */
uint32_t ora_encoding_table(const uint8_t* bytes, const uint32_t* offs, uint32_t n_groups,
                            int with_wildcard, uint8_t table_out[256]);

/* FmIndexBuilder::new + set_*_config + blob_size (builder/mod.rs:63-181).
 * kmer_size: 1 = LookupTableConfig::None, >=2 = KmerSize(k).  sampling_ratio: 1 = Uncompressed. */
int ora_builder_layout(ora_type t, uint64_t text_len, uint32_t symbol_count, uint32_t kmer_size,
                       uint32_t sampling_ratio, ora_layout* out, uint64_t detail[2]);
/* LookupTableConfig::MaxMemory (lookup_table_config.rs:41-53). */
uint32_t ora_kmer_size_for_max_memory(uint32_t pos_bits, uint32_t symbol_count, uint64_t max_bytes);

/* FmIndexBuilder::build (builder/mod.rs:187-264).  `text` is consumed (encoded and BWT'd in a
 * private copy; the caller's buffer is left untouched).  table256 may be NULL for PassThrough. */
int ora_build(ora_type t, const uint8_t* text, uint64_t text_len, uint32_t symbol_count,
              const uint8_t* table256, uint32_t kmer_size, uint32_t sampling_ratio,
              uint8_t* blob, uint64_t blob_len, uint64_t detail[2]);

typedef struct ora_index ora_index;

/* FmIndex::load (load_from_blob.rs:28-85).  Borrows `blob` (like the Rust lifetime 'a). */
int ora_load(const uint8_t* blob, uint64_t blob_len, ora_type t, ora_index** out, uint64_t detail[2]);
void ora_free(ora_index* ix);
const ora_layout* ora_index_layout(const ora_index* ix);
uint64_t ora_text_len(const ora_index* ix);

/* FmIndex::count / count_rev_iter (locate/with_slice.rs:5-8, with_rev_iter.rs:5-9).
 * reversed != 0: `pat` holds the pattern back-to-front and is consumed like the rev iterator. */
int ora_count(const ora_index* ix, const uint8_t* pat, uint64_t len, int reversed, uint64_t* count);
/* FmIndex::locate_to_buffer semantics (with_slice.rs:15-18): appends at out[*n_out..], SA-row order.
 * out holds positions widened to u64.  If more than cap entries are needed returns the needed
 * total in *n_out and writes nothing beyond cap. */
int ora_locate(const ora_index* ix, const uint8_t* pat, uint64_t len, int reversed,
               uint64_t* out, uint64_t cap, uint64_t* n_out);
/* (sp, ep) of get_pos_range (with_slice.rs:21-33) -- exposed so tests can compare the SA interval. */
int ora_pos_range(const ora_index* ix, const uint8_t* pat, uint64_t len, int reversed,
                  uint64_t* sp, uint64_t* ep);

/* Pattern-parallel drivers for the CPU baseline: one reference call per pattern (get_pos_range, then
 * write_locations_to_buffer for locate), patterns split contiguously over `threads` pthreads.
 * Fixed-length patterns (pats[i*len .. (i+1)*len)).
 * ora_count_batch: counts_out u64[n].
 * ora_locate_batch: ONE backward search per pattern like FmIndex::locate; counts_out (nullable) gets
 * ep-sp; if pos_out != NULL the locations of pattern i are written in SA-row order at
 * pos_out[out_offs[i]..] (u32 or u64 per pos_out_bits; out_offs = exclusive prefix sums of the counts);
 * checksum_out (nullable) = sum_i sum_loc (loc+1)*(2i+1) mod 2^64, an order-independent digest. */
int ora_count_batch(const ora_index* ix, const uint8_t* pats, uint64_t n, uint64_t len,
                    uint64_t* counts_out, int threads);
int ora_locate_batch(const ora_index* ix, const uint8_t* pats, uint64_t n, uint64_t len,
                     uint64_t* counts_out, const uint64_t* out_offs, void* pos_out, uint32_t pos_out_bits,
                     uint64_t* checksum_out, int threads);

/* Exposed for tests of the builder internals. */
int ora_suffix_array(const uint8_t* text_with_sentinel, int32_t n_with_sentinel, int32_t alphabet,
                     int32_t* sa_out);

#ifdef __cplusplus
}
#endif
#endif
