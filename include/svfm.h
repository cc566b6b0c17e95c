/*
 * svfm.h -- C ABI of the B200-native batched FM-index search engine (libsvfm.so).
 *
 * Drop-in boundary for ONE path of baku4/sview-fmindex: `FmIndex::load` + backward-search `count` +
 * SA-sampled `locate`, over many patterns at once, on a blob produced by the reference's
 * `FmIndexBuilder`.  The reference (pure Rust) has no FFI/plugin interface; the boundary is its public
 * API, so every entry point below names the Rust item it stands in for.  Citations are relative to
 * the reference's `sview-fmindex/src/`.
 *
 * Conventions: plain pointers and sizes, `int` status codes, nothing aborts or throws across the ABI.
 * Host pointers unless a parameter is named d_* (device pointer on the index's device).
 * There is NO CPU fallback: every call needs a CUDA device and fails with SVFM_ERR_CUDA without one.
 */
#ifndef SVFM_H
#define SVFM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------- */
enum {
    SVFM_OK = 0,
    SVFM_ERR_INVALID_FORMAT = 1,   /* LoadError::InvalidFormat          load_from_blob.rs:18-19,31-33 */
    SVFM_ERR_BLOB_SIZE = 2,        /* LoadError::MismatchedBlobSize(expected, actual)
                                      load_from_blob.rs:22-23,46-58; err_detail = {expected, actual}  */
    SVFM_ERR_SYMBOL_COUNT_OVER = 10, /* BuildError::SymbolCountOver(max, got)  builder/mod.rs:39-40,71-73 */
    SVFM_ERR_TEXT_LENGTH = 11,     /* BuildError::UnmatchedTextLength    builder/mod.rs:42-43 */
    SVFM_ERR_INVALID_BLOB_SIZE = 12, /* BuildError::InvalidBlobSize(expected, got) builder/mod.rs:45-46,205-209 */
    SVFM_ERR_NOT_ALIGNED = 13,     /* BuildError::NotAlignedBlob(required, offset) builder/mod.rs:48-49,198-203 */
    SVFM_ERR_INVALID_CONFIG = 14,  /* BuildError::InvalidConfig          builder/mod.rs:51-52 */
    SVFM_ERR_BAD_TYPE = 20,        /* (P, B, E) triple outside u32|u64 x Block2..6 x u32|u64|u128 */
    SVFM_ERR_EMPTY_PATTERN = 21,   /* the reference panics on an empty pattern (count_array.rs:211,257) */
    SVFM_ERR_TOO_LARGE = 22,       /* GPU builder: text_len must be < 2^32 - 1 */
    SVFM_ERR_NOMEM = 23,
    SVFM_ERR_BAD_SYMBOL = 24,      /* PassThrough pattern byte >= symbol_count: the reference reads a wrong
                                      checkpoint row silently (pass_through.rs:10-12, bwm/mod.rs:207-208) */
    SVFM_ERR_CAPACITY = 25,        /* caller buffer too small; *total holds the needed element count */
    SVFM_ERR_BAD_ARG = 26,
    SVFM_ERR_CUDA = 30             /* any CUDA failure, including "no device"; see svfm_last_error() */
};

/* ---- the (P, B, E) type triple -------------------------------------------------------------
 * `FmIndex<'a, P: Position, B: Block, E: TextEncoder>` (lib.rs:15-28).  The blob does not record it
 * (magic is "FI00"+4 zero bytes, components/magic_number.rs:12-27); the caller supplies it, exactly as
 * the Rust caller picks the generic arguments. */
typedef struct svfm_type {
    uint32_t pos_bits; /* Position: 32 | 64                 text_length.rs:45,87 */
    uint32_t planes;   /* Block2..Block6: 2..6              components/bwm/blocks/mod.rs:13-33 */
    uint32_t vec_bits; /* Vector: 32 | 64 | 128             components/bwm/blocks/vector.rs:35-79 */
    uint32_t encoder;  /* 0 PassThrough | 1 EncodingTable   components/text_encoder/text_encoders/ */
} svfm_type;

typedef struct svfm_info {
    svfm_type type;
    int32_t device;
    uint32_t symbol_count;     /* CountArrayHeader.symbol_count        count_array.rs:10-18 */
    uint32_t kmer_size;        /* CountArrayHeader.lookup_table_kmer_size */
    uint32_t sampling_ratio;   /* SuffixArrayHeader.sampling_ratio     suffix_array/mod.rs:12-18 */
    uint64_t text_len;         /* = count_array[symbol_count]          count_array.rs:117,125 */
    uint64_t suffix_array_len;
    uint64_t blocks_len;       /* BwmHeader.blocks_len                 bwm/mod.rs:9-16 */
    uint64_t sentinel_index;
    uint64_t blob_len;
    uint64_t header_size;
    uint64_t off_suffix_array, off_rank_checkpoints, off_blocks; /* byte offsets inside the blob */
} svfm_info;

typedef struct svfm_index svfm_index;     /* device-resident FmIndex (owns a byte-for-byte copy of the blob) */
typedef struct svfm_session svfm_session; /* one CUDA stream + scratch arena; one per concurrent caller */

/* flags of the batch calls */
enum {
    SVFM_REVERSED = 1, /* patterns are stored back-to-front: count_rev_iter / locate_rev_iter
                          (locate/with_rev_iter.rs:5-38, count_array.rs:235-274) */
    SVFM_SORTED = 2,   /* locate: ascending positions per pattern.  Default is the reference's SA-row
                          order (locate/mod.rs:19; "The locations may not be in order", README.md:77) */
    SVFM_OFFS32 = 4    /* locate: out_offs is uint32_t[n+1] instead of uint64_t[n+1] (half the bytes of a typical
                          result on the way back to the host); SVFM_ERR_TOO_LARGE when *total >= 2^32 */
};

/* ---- load ---------------------------------------------------------------------------------
 * `FmIndex::<P,B,E>::load(blob: &'a [u8]) -> Result<Self, LoadError>` (load_from_blob.rs:28-85).
 * Parses the five headers with the reference's alignment rules (components/mod.rs:1-23), checks the
 * magic/version and `body.len() == sum of aligned body sizes`, then uploads the blob to `device`
 * unchanged.  The handle owns the device copy: the host blob may be released after the call (the
 * Rust wrapper keeps its `&'a [u8]` for `blob()`, reference_to_source_blob.rs:9).
 * Blobs shorter than their headers make the reference panic (components/mod.rs:19): here
 * SVFM_ERR_INVALID_FORMAT. */
int svfm_load(const uint8_t* blob, size_t blob_len, svfm_type t, int device, svfm_index** out,
              uint64_t err_detail[2]);
/* Same, from a blob that already sits in device memory on `device` (device-to-device copy). */
int svfm_load_device(const uint8_t* d_blob, size_t blob_len, svfm_type t, int device, svfm_index** out,
                     uint64_t err_detail[2]);
void svfm_free(svfm_index* ix);
int svfm_index_info(const svfm_index* ix, svfm_info* out);
/* Device memory held by the handle, in bytes: out[0] the blob copy, out[1] the extended k-mer table, out[2] the
 * interleaved occ copy, out[3] scratch arenas of the idle sessions / upload staging (grow-only until svfm_free),
 * out[4] the packed text copy, out[5] the expanded suffix array, out[6] the sweep occ copy, out[7] reserved (0). */
int svfm_index_memory(svfm_index* ix, uint64_t out[8]);
/* Host-only part of load: validate + report sizes without touching a device (LoadError paths). */
int svfm_check_blob(const uint8_t* blob, size_t blob_len, svfm_type t, svfm_info* out, uint64_t err_detail[2]);

/* ---- batched count / locate (host buffers) -------------------------------------------------
 * Pattern i is pats[offs[i] .. offs[i+1]) or, when offs == NULL, pats[i*fixed_len .. (i+1)*fixed_len).
 * Results for pattern i are identical to `count(p_i)` / `locate(p_i)` of the reference
 * (locate/with_slice.rs:5-18); input order is preserved.  An empty pattern anywhere in the batch
 * fails the call with SVFM_ERR_EMPTY_PATTERN before any work.
 * Buffers may be pageable or pinned (svfm_host_alloc); pinned makes the copies asynchronous.
 * Calls on one index may run concurrently from several host threads. */

/* `count` x n.  counts_out: P[n] (uint32_t or uint64_t per svfm_type.pos_bits). */
int svfm_count_batch(svfm_index* ix, const uint8_t* pats, const uint64_t* offs, uint64_t n,
                     uint32_t fixed_len, uint32_t flags, void* counts_out);

/* `locate` x n, CSR result: out_offs[n+1] (exclusive prefix sums of the counts, so
 * count_i = out_offs[i+1]-out_offs[i]) and positions P[*total]; pattern i owns
 * positions[out_offs[i] .. out_offs[i+1]).  If *total > capacity nothing is written to `positions`,
 * out_offs and *total are still valid and the call returns SVFM_ERR_CAPACITY. */
int svfm_locate_batch(svfm_index* ix, const uint8_t* pats, const uint64_t* offs, uint64_t n,
                      uint32_t fixed_len, uint32_t flags, void* out_offs /* u64[n+1]; u32[n+1] with SVFM_OFFS32 */,
                      void* positions, uint64_t capacity, uint64_t* total);
/* Same, positions allocated by the library (like the Vec<P> `locate` returns; pinned host memory for results of
 * 1 MiB and more, plain memory below); release with svfm_free_positions only. */
int svfm_locate_batch_alloc(svfm_index* ix, const uint8_t* pats, const uint64_t* offs, uint64_t n,
                            uint32_t fixed_len, uint32_t flags, void* out_offs, void** positions,
                            uint64_t* total);
void svfm_free_positions(void* positions);

/* ---- packed fixed-length batches (fewer bytes over PCIe; the calls above stay the drop-in ones) -----------------
 * Pattern i occupies bytes [i*bpp, (i+1)*bpp) of `packed`, bpp = ceil(len*bits/8); symbol j of the pattern (the value
 * `TextEncoder::idx_of` returns for its byte, encoding_table.rs:9-11) sits in bits [j*bits, (j+1)*bits) of that little-
 * endian bit string; bits in 1..8 and every symbol index < 2^bits (so 2 bits carry ACGT even when the table also has an
 * N class; a pattern with a symbol index >= symbol_count fails with SVFM_ERR_BAD_SYMBOL).  20 bp at 2 bits = 5 bytes
 * instead of 20.  Results are those of the unpacked calls on the same patterns.  SVFM_REVERSED is not supported.
 * svfm_pack_patterns is a multi-threaded host helper producing this layout from byte patterns (table256 = the
 * EncodingTable bytes, NULL = bytes are symbol indices already); SVFM_ERR_BAD_SYMBOL when a symbol does not fit `bits`. */
int svfm_count_batch_packed(svfm_index* ix, const uint8_t* packed, uint64_t n, uint32_t len, uint32_t bits,
                            uint32_t flags, void* counts_out);
int svfm_locate_batch_packed(svfm_index* ix, const uint8_t* packed, uint64_t n, uint32_t len, uint32_t bits,
                             uint32_t flags, void* out_offs, void* positions, uint64_t capacity, uint64_t* total);
int svfm_pack_patterns(const uint8_t* pats, uint64_t n, uint32_t len, const uint8_t* table256, uint32_t bits,
                       uint8_t* packed_out);

/* Single-pattern conveniences = the batch calls with n = 1.
 * `count(&self, pattern:&[u8]) -> P`, `locate_to_buffer(&self, pattern, &mut Vec<P>)` (appends). */
int svfm_count(svfm_index* ix, const uint8_t* pattern, uint64_t len, uint32_t flags, uint64_t* count);
int svfm_locate(svfm_index* ix, const uint8_t* pattern, uint64_t len, uint32_t flags, void* positions,
                uint64_t capacity, uint64_t* total);

/* ---- device-resident batches (no host copies; what `value` in bench.py times) ----------------
 * All d_* pointers live on the index's device.  Work is enqueued on the session's stream;
 * svfm_session_sync waits for it.  d_counts_out: P[n].  Locate: d_out_offs u64[n+1]; positions are
 * written to the session's arena, *d_positions stays valid until the session's next locate call. */
int svfm_session_create(svfm_index* ix, svfm_session** out);
void svfm_session_destroy(svfm_session* s);
int svfm_session_sync(svfm_session* s);
void* svfm_session_stream(svfm_session* s); /* cudaStream_t */
int svfm_count_batch_device(svfm_session* s, const uint8_t* d_pats, const uint64_t* d_offs, uint64_t n,
                            uint32_t fixed_len, uint32_t flags, void* d_counts_out);
int svfm_locate_batch_device(svfm_session* s, const uint8_t* d_pats, const uint64_t* d_offs, uint64_t n,
                             uint32_t fixed_len, uint32_t flags, void* d_out_offs /* u64[n+1]; u32[n+1] with SVFM_OFFS32 */,
                             void** d_positions, uint64_t* total);

/* ---- per-phase device timing of a session (CUDA events on the session's stream) --------------------
 * Phases of the batch pipelines; ms[] = summed device time of the phase's kernels since the last reset,
 * launches[] = kernels launched in the phase.  Reading synchronises the session's stream. */
enum {
    SVFM_PHASE_PRESORT = 0, /* pattern encoding/packing + locality sort of the batch */
    SVFM_PHASE_SEARCH = 1,  /* backward-search kernels (search_kernel / sweep_round_kernel) */
    SVFM_PHASE_SCAN = 2,    /* exclusive prefix sum of the counts -> CSR offsets */
    SVFM_PHASE_LOCATE = 3,  /* LF-walk + sampled-SA lookup kernel */
    SVFM_PHASE_SEGSORT = 4, /* optional per-pattern sort of the positions (SVFM_SORTED) */
    SVFM_PHASE_OTHER = 5,   /* back to the caller's order: radix sort by pattern index, CSR offsets */
    SVFM_PHASE_MAX = 8
};
int svfm_session_set_timing(svfm_session* s, int enabled);
int svfm_session_get_timing(svfm_session* s, double ms[SVFM_PHASE_MAX], uint64_t launches[SVFM_PHASE_MAX], int reset);

/* ---- misc -------------------------------------------------------------------------------------- */
void* svfm_host_alloc(size_t bytes);  /* pinned host memory (cudaHostAlloc) or NULL */
void svfm_host_free(void* p);
/* Process-wide tuning knobs; results never depend on any of them.
 * SVFM_TUNE_SORT_MIN : batches with at least this many patterns (that do not qualify for the sweep search) are
 *                      locality-sorted by their trailing symbols before the search kernel (0 = always,
 *                      UINT64_MAX = never; default SVFM_TUNE_AUTO = never when the index has an extended k-mer table
 *                      -- nothing is left to share after the lookup -- else 131072; env SVFM_SORT_MIN).
 * SVFM_TUNE_CHUNK    : the host-buffer entry points cut a batch into chunks of about this many patterns and
 *                      pipeline upload / kernels / download (0 = one chunk; default SVFM_TUNE_AUTO = 4 Mi; env SVFM_CHUNK).
 * SVFM_TUNE_SWEEP_MIN: fixed-length batches with at least this many patterns use the sweep search -- the batch is
 *                      kept sorted by SA position and moves through the index as streams (default SVFM_TUNE_AUTO = the
 *                      measured break-even with the plain search kernel on a 1 Gbp index: 48 Mi patterns for locate and
 *                      16 Mi for count when the index has its packed text copy, expanded suffix array and a 2^30-entry
 *                      extended table (24 Mi / 8 Mi with a 2^28-entry one), else 5 Mi with a 2^24-entry extended table
 *                      and 10 Mi with a 2^28-entry one; env SVFM_SWEEP_MIN).
 * SVFM_TUNE_EXT_BITS : indexes loaded from now on get an extended k-mer table of at most 2^value entries (and at most
 *                      two per text symbol), derived from the blob at load (0 = none; default SVFM_TUNE_AUTO = 30, i.e.
 *                      8 GiB for u32 positions, when that is under 1/8 of the free device memory, else 28 = 2 GiB when
 *                      under 1/16, else 24 = 128 MiB; env SVFM_EXT_BITS).
 * SVFM_TUNE_WORKERS  : host threads / compute streams per host-buffer call (default 2; env SVFM_WORKERS).
 * SVFM_TUNE_ILV      : indexes loaded from now on also get an interleaved copy of the occ data -- block q and checkpoint
 *                      row q in one aligned 32/64/128-byte slot -- which the gather-bound kernels read instead of the two
 *                      blob sections (1 = build, the default; 0 = search the blob in place only; env SVFM_ILV).
 * SVFM_TUNE_BUCKET_SORTBACK : how a reordered batch gets back into the caller's order.  1 (default): locate writes its
 *                      records straight into buckets of 8192 pattern indices and one kernel finishes every bucket in shared
 *                      memory; `count` groups by the top index bits and scatters.  0: radix sorts by pattern index
 *                      (env SVFM_BUCKET_SORTBACK).
 * SVFM_TUNE_SMALL_MAX : host batches of at most this many patterns (and at most 256 KiB of pattern bytes) run as ONE kernel
 *                      launch that searches and locates, with patterns and results in mapped pinned host memory -- one
 *                      launch and one synchronisation per call instead of six runtime calls; a pattern with more than 8
 *                      occurrences sends its batch through the general pipeline (default and maximum 4096; 0 = never;
 *                      env SVFM_SMALL_MAX).
 * SVFM_TUNE_TEXT     : indexes loaded from now on also get a packed copy of the indexed text (1, 2, 4 or 8 bits per symbol),
 *                      recovered from the blob itself at load.  The search kernel then finishes a pattern whose SA interval
 *                      is down to a few rows with many symbols left (150 bp reads after ~16 symbols) by locating the
 *                      candidates and comparing the rest of the pattern with the text -- a few sectors instead of one
 *                      random line per remaining symbol (1 = build, the default; 0 = index only; env SVFM_TEXT).
 * SVFM_TUNE_L2_PERSIST: indexes loaded from now on whose extended k-mer table takes at most this many bytes (default 64 MiB;
 *                      0 = never) mark it persisting in L2 for all their kernels (cudaAccessPolicyWindow).  The default
 *                      2 GiB table is never pinned; this serves deployments that cap SVFM_TUNE_EXT_BITS (env SVFM_L2_PERSIST).
 * SVFM_TUNE_OWN_RADIX: the sweep search sorts its items by table index with this library's radix_pass_kernel (1) or with
 *                      cub::DeviceRadixSort (0, the default: measured 0.84 against 1.36 ms per pass; env SVFM_OWN_RADIX).
 * SVFM_TUNE_FULL_SA  : indexes loaded from now on whose blob samples the suffix array (ratio > 1) also get the EXPANDED
 *                      suffix array -- the text position of every SA row (4 bytes per text symbol below 2^32 symbols, else
 *                      8), derived from the blob at load by walking every row once -- when it takes at most 1/8 of the
 *                      device memory still free.  `locate` then reads one entry per row instead of LF-walking to a sampled
 *                      row (suffix_array/mod.rs:100-105 trades memory for that walk; a B200 has the memory).  1 (default) /
 *                      0 = never (env SVFM_FULL_SA).  Results never depend on it.
 * SVFM_TUNE_SWEEP_OCC: indexes loaded from now on with 64-bit vectors, at most three planes and at most four occurring
 *                      symbols also get the sweep occ copy -- block q's planes and 16-bit checkpoint deltas in ONE 32-byte
 *                      sector, so that a rank query of the sweep rounds is one 256-bit load (1 default / 0; env
 *                      SVFM_SWEEP_OCC).  Results never depend on it. */
enum { SVFM_TUNE_SORT_MIN = 0, SVFM_TUNE_CHUNK = 1, SVFM_TUNE_SWEEP_MIN = 2, SVFM_TUNE_EXT_BITS = 3, SVFM_TUNE_WORKERS = 4,
       SVFM_TUNE_ILV = 5, SVFM_TUNE_BUCKET_SORTBACK = 6, SVFM_TUNE_SMALL_MAX = 7, SVFM_TUNE_TEXT = 8, SVFM_TUNE_L2_PERSIST = 9, SVFM_TUNE_OWN_RADIX = 10,
       SVFM_TUNE_FULL_SA = 11, SVFM_TUNE_SWEEP_OCC = 12 };
#define SVFM_TUNE_AUTO 0xfffffffffffffffeull
int svfm_set_tuning(int key, uint64_t value);
const char* svfm_last_error(void);    /* thread-local text of the last SVFM_ERR_CUDA */
uint64_t svfm_launch_count(void);     /* kernels launched by this library since process start */
const char* svfm_version(void);

/* ---- index construction on the GPU (not part of the reference hot path) ----
 * `FmIndexBuilder::<P,B,E>::new(text_len, symbol_count, encoder)` + `set_lookup_table_config(KmerSize(k))`
 * + `set_suffix_array_config(Compressed(r))` + `blob_size()` + `build(text, blob)`
 * (builder/mod.rs:63-264).  Produces the same bytes as the reference builder: the blob is a pure
 * function of (text, encoder, P, B, k, r) because the suffix array with a unique smallest sentinel is
 * unique.  kmer_size 1 = LookupTableConfig::None; sampling_ratio 1 = SuffixArrayConfig::Uncompressed.
 * table256 = EncodingTable bytes (NULL for PassThrough). */
int svfm_blob_size(svfm_type t, uint64_t text_len, uint32_t symbol_count, uint32_t kmer_size,
                   uint32_t sampling_ratio, uint64_t* blob_size, uint64_t err_detail[2]);
int svfm_build(svfm_type t, const uint8_t* text, uint64_t text_len, uint32_t symbol_count,
               const uint8_t* table256, uint32_t kmer_size, uint32_t sampling_ratio, int device,
               uint8_t* blob_out, uint64_t blob_len, uint64_t err_detail[2]);
/* d_text and d_blob_out in device memory on `device`; d_text is left unchanged. */
int svfm_build_device(svfm_type t, const uint8_t* d_text, uint64_t text_len, uint32_t symbol_count,
                      const uint8_t* table256, uint32_t kmer_size, uint32_t sampling_ratio, int device,
                      uint8_t* d_blob_out, uint64_t blob_len, uint64_t err_detail[2]);

#ifdef __cplusplus
}
#endif
#endif /* SVFM_H */
