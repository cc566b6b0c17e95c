/*
 * fm_oracle.c -- CPU ORACLE for the sview-fmindex hot path.  TEST INFRASTRUCTURE ONLY
 * (see fm_oracle.h for who may call it and for the parity status).
 *
 * Restates, in plain C, the reference's blob framing, builder, loader and query path.
 * Citations are relative to /root/reference/sview-fmindex/src/.
 */
#include "fm_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Blob framing: Aligned::aligned_size (components/mod.rs:1-8), Header::aligned_size
 * (components/mod.rs:10-23), builder header/body sizes (builder/mod.rs:160-181).
 * ---------------------------------------------------------------------------------------- */
static uint64_t align_up(uint64_t raw, uint64_t a) {
    uint64_t rem = raw % a;
    return rem == 0 ? raw : raw + (a - rem);
}

static int type_ok(ora_type t) {
    if (t.pos_bits != 32 && t.pos_bits != 64) return 0;
    if (t.planes < 2 || t.planes > 6) return 0;
    if (t.vec_bits != 32 && t.vec_bits != 64 && t.vec_bits != 128) return 0;
    if (t.encoder > 1) return 0;
    return 1;
}

/* Fill every offset from the header fields already stored in L. */
static void layout_offsets(ora_type t, ora_layout* L) {
    uint64_t A = t.vec_bits == 128 ? 16 : 8; /* vector.rs:37,52,67 */
    uint64_t P = t.pos_bits / 8;
    uint64_t block_bytes = (uint64_t)t.planes * t.vec_bits / 8;
    L->align = A;
    uint64_t o = align_up(8, A);                               /* MagicNumber([u8;8]) */
    L->off_encoder = o;
    o += align_up(t.encoder ? 256 : 0, A);                     /* EncodingTable | PassThrough ZST */
    L->off_count_header = o;
    o += align_up(24, A);                                      /* CountArrayHeader, count_array.rs:10-18 */
    L->off_sa_header = o;
    o += align_up(16, A);                                      /* SuffixArrayHeader, suffix_array/mod.rs:12-18 */
    L->off_bwm_header = o;
    o += align_up(24, A);                                      /* BwmHeader, bwm/mod.rs:9-16 */
    L->header_size = o;
    L->off_count_array = o;
    o += align_up((uint64_t)L->count_array_len * P, A);
    L->off_kmer_multiplier = o;
    o += align_up((uint64_t)L->kmer_multiplier_len * 8, A);    /* usize = 8 B */
    L->off_kmer_count_table = o;
    o += align_up(L->kmer_count_table_len * P, A);
    L->off_suffix_array = o;
    o += align_up(L->suffix_array_len * P, A);
    L->off_sentinel_index = o;
    o += align_up(P, A);
    L->off_rank_checkpoints = o;
    o += align_up(L->rank_checkpoints_len * P, A);
    L->off_blocks = o;
    o += align_up(L->blocks_len * block_bytes, A);
    L->total_size = o;
}

static uint32_t max_symbol(uint32_t planes) { return 1u << planes; } /* block2.rs:15 .. block6.rs:15 */

uint32_t ora_kmer_size_for_max_memory(uint32_t pos_bits, uint32_t symbol_count, uint64_t max_bytes) {
    /* lookup_table_config.rs:41-53 */
    uint64_t swsc = (uint64_t)symbol_count + 1;
    uint32_t kmer_size = 2;
    for (;;) {
        uint64_t sz = pos_bits / 8;
        for (uint32_t e = 0; e < kmer_size; e++) sz *= swsc;
        if (sz <= max_bytes) kmer_size += 1; else break;
    }
    return kmer_size - 1;
}

int ora_builder_layout(ora_type t, uint64_t text_len, uint32_t symbol_count, uint32_t kmer_size,
                       uint32_t sampling_ratio, ora_layout* L, uint64_t detail[2]) {
    if (!type_ok(t)) return ORA_ERR_BAD_TYPE;
    if (symbol_count > max_symbol(t.planes)) {             /* builder/mod.rs:71-73 */
        if (detail) { detail[0] = max_symbol(t.planes); detail[1] = symbol_count; }
        return ORA_ERR_SYMBOL_COUNT_OVER;
    }
    if (kmer_size < 1 || sampling_ratio < 1) return ORA_ERR_INVALID_CONFIG;
    memset(L, 0, sizeof(*L));
    /* CountArrayHeader::new (count_array.rs:58-77) */
    L->symbol_count = symbol_count;
    L->kmer_size = kmer_size;
    L->count_array_len = symbol_count + 1;
    L->kmer_multiplier_len = kmer_size;
    {
        /* `(symbol_with_sentinel_count).pow(k) as u64` is computed in u32 in the reference */
        uint32_t p = 1;
        for (uint32_t e = 0; e < kmer_size; e++) p *= (symbol_count + 1);
        L->kmer_count_table_len = p;
    }
    /* SuffixArrayHeader::new (suffix_array/mod.rs:43-56) */
    L->sampling_ratio = sampling_ratio;
    L->suffix_array_len = text_len / sampling_ratio + (text_len % sampling_ratio ? 1 : 0);
    /* BwmHeader::new (bwm/mod.rs:71-89) */
    L->blocks_len = text_len / t.vec_bits + 1;
    L->rank_checkpoints_len = L->blocks_len * symbol_count;
    layout_offsets(t, L);
    return ORA_OK;
}

uint32_t ora_encoding_table(const uint8_t* bytes, const uint32_t* offs, uint32_t n_groups,
                            int with_wildcard, uint8_t table_out[256]) {
    /* encoding_table.rs:15-34 */
    uint32_t symbol_count = n_groups + (with_wildcard ? 1u : 0u);
    memset(table_out, (int)(uint8_t)(symbol_count - 1), 256);
    for (uint32_t g = 0; g < n_groups; g++)
        for (uint32_t i = offs[g]; i < offs[g + 1]; i++) table_out[bytes[i]] = (uint8_t)g;
    /* symbol_count() = max + 1 (encoding_table.rs:35-37) */
    uint32_t mx = 0;
    for (int i = 0; i < 256; i++) if (table_out[i] > mx) mx = table_out[i];
    return mx + 1;
}

/* ------------------------------------------------------------------------------------------
 * Suffix sorting (SA-IS, Nong/Zhang/Chan 2009; written for this oracle).  The reference sorts
 * with a rust-bio SA-IS copy or libdivsufsort (components/suffix_array/burrow_wheeler_transform/);
 * any correct sorter produces the same SA because the appended sentinel is unique and smallest.
 * ---------------------------------------------------------------------------------------- */
#define CHR(i) (cs == 4 ? ((const int32_t*)s)[i] : (int32_t)((const uint8_t*)s)[i])
#define TGET(i) ((t[(i) >> 3] >> ((i) & 7)) & 1)
#define TSET(i, b) (t[(i) >> 3] = (uint8_t)((b) ? (t[(i) >> 3] | (1u << ((i) & 7))) : (t[(i) >> 3] & ~(1u << ((i) & 7)))))
#define IS_LMS(i) ((i) > 0 && TGET(i) && !TGET((i) - 1))

static void get_buckets(const void* s, int32_t* bkt, int32_t n, int32_t K, int cs, int end) {
    for (int32_t i = 0; i <= K; i++) bkt[i] = 0;
    for (int32_t i = 0; i < n; i++) bkt[CHR(i)]++;
    int32_t sum = 0;
    for (int32_t i = 0; i <= K; i++) { sum += bkt[i]; bkt[i] = end ? sum : sum - bkt[i]; }
}

static void induce_l(const uint8_t* t, int32_t* SA, const void* s, int32_t* bkt, int32_t n, int32_t K, int cs) {
    get_buckets(s, bkt, n, K, cs, 0);
    for (int32_t i = 0; i < n; i++) {
        int32_t j = SA[i] - 1;
        if (j >= 0 && !TGET(j)) SA[bkt[CHR(j)]++] = j;
    }
}

static void induce_s(const uint8_t* t, int32_t* SA, const void* s, int32_t* bkt, int32_t n, int32_t K, int cs) {
    get_buckets(s, bkt, n, K, cs, 1);
    for (int32_t i = n - 1; i >= 0; i--) {
        int32_t j = SA[i] - 1;
        if (j >= 0 && TGET(j)) SA[--bkt[CHR(j)]] = j;
    }
}

/* s[n-1] must be the unique smallest symbol (0); symbols in [0, K]. */
static int sais_rec(const void* s, int32_t* SA, int32_t n, int32_t K, int cs) {
    uint8_t* t = (uint8_t*)calloc((size_t)n / 8 + 1, 1);
    int32_t* bkt = (int32_t*)malloc(sizeof(int32_t) * ((size_t)K + 1));
    if (!t || !bkt) { free(t); free(bkt); return ORA_ERR_NOMEM; }
    /* classify: S-type = 1, L-type = 0 */
    TSET(n - 2 >= 0 ? n - 2 : 0, 0);
    TSET(n - 1, 1);
    for (int32_t i = n - 3; i >= 0; i--)
        TSET(i, (CHR(i) < CHR(i + 1) || (CHR(i) == CHR(i + 1) && TGET(i + 1) == 1)) ? 1 : 0);
    /* stage 1: sort LMS substrings */
    get_buckets(s, bkt, n, K, cs, 1);
    for (int32_t i = 0; i < n; i++) SA[i] = -1;
    for (int32_t i = 1; i < n; i++) if (IS_LMS(i)) SA[--bkt[CHR(i)]] = i;
    induce_l(t, SA, s, bkt, n, K, cs);
    induce_s(t, SA, s, bkt, n, K, cs);
    free(bkt);
    /* compact sorted LMS substrings into SA[0..n1) */
    int32_t n1 = 0;
    for (int32_t i = 0; i < n; i++) if (IS_LMS(SA[i])) SA[n1++] = SA[i];
    for (int32_t i = n1; i < n; i++) SA[i] = -1;
    /* name them */
    int32_t name = 0, prev = -1;
    for (int32_t i = 0; i < n1; i++) {
        int32_t pos = SA[i];
        int diff = 0;
        for (int32_t d = 0; d < n; d++) {
            if (prev == -1 || CHR(pos + d) != CHR(prev + d) || TGET(pos + d) != TGET(prev + d)) { diff = 1; break; }
            else if (d > 0 && (IS_LMS(pos + d) || IS_LMS(prev + d))) break;
        }
        if (diff) { name++; prev = pos; }
        SA[n1 + pos / 2] = name - 1;
    }
    for (int32_t i = n - 1, j = n - 1; i >= n1; i--) if (SA[i] >= 0) SA[j--] = SA[i];
    /* stage 2: solve the reduced problem */
    int32_t* SA1 = SA;
    int32_t* s1 = SA + n - n1;
    if (name < n1) {
        int rc = sais_rec(s1, SA1, n1, name - 1, 4);
        if (rc) { free(t); return rc; }
    } else {
        for (int32_t i = 0; i < n1; i++) SA1[s1[i]] = i;
    }
    /* stage 3: induce the result */
    bkt = (int32_t*)malloc(sizeof(int32_t) * ((size_t)K + 1));
    if (!bkt) { free(t); return ORA_ERR_NOMEM; }
    get_buckets(s, bkt, n, K, cs, 1);
    for (int32_t i = 1, j = 0; i < n; i++) if (IS_LMS(i)) s1[j++] = i;
    for (int32_t i = 0; i < n1; i++) SA1[i] = s1[SA1[i]];
    for (int32_t i = n1; i < n; i++) SA[i] = -1;
    for (int32_t i = n1 - 1; i >= 0; i--) {
        int32_t j = SA[i];
        SA[i] = -1;
        SA[--bkt[CHR(j)]] = j;
    }
    induce_l(t, SA, s, bkt, n, K, cs);
    induce_s(t, SA, s, bkt, n, K, cs);
    free(bkt);
    free(t);
    return ORA_OK;
}

int ora_suffix_array(const uint8_t* text_with_sentinel, int32_t n, int32_t alphabet, int32_t* sa_out) {
    if (n <= 0) return ORA_OK;
    if (n == 1) { sa_out[0] = 0; return ORA_OK; }
    return sais_rec(text_with_sentinel, sa_out, n, alphabet, 1);
}

/* ------------------------------------------------------------------------------------------
 * 30 monomorphised copies of the query + build-body code.
 * ---------------------------------------------------------------------------------------- */
typedef unsigned __int128 u128_t;

#define POS_T uint32_t
#define VEC_T uint32_t
#define VBITS 32
#define NPL 2
#define SFX p32_n2_v32
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 3
#define SFX p32_n3_v32
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 4
#define SFX p32_n4_v32
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 5
#define SFX p32_n5_v32
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 6
#define SFX p32_n6_v32
#include "fm_query.inc"
#undef NPL
#undef SFX
#undef VEC_T
#undef VBITS
#define VEC_T uint64_t
#define VBITS 64
#define NPL 2
#define SFX p32_n2_v64
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 3
#define SFX p32_n3_v64
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 4
#define SFX p32_n4_v64
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 5
#define SFX p32_n5_v64
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 6
#define SFX p32_n6_v64
#include "fm_query.inc"
#undef NPL
#undef SFX
#undef VEC_T
#undef VBITS
#define VEC_T u128_t
#define VBITS 128
#define NPL 2
#define SFX p32_n2_v128
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 3
#define SFX p32_n3_v128
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 4
#define SFX p32_n4_v128
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 5
#define SFX p32_n5_v128
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 6
#define SFX p32_n6_v128
#include "fm_query.inc"
#undef NPL
#undef SFX
#undef VEC_T
#undef VBITS
#undef POS_T

#define POS_T uint64_t
#define VEC_T uint32_t
#define VBITS 32
#define NPL 2
#define SFX p64_n2_v32
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 3
#define SFX p64_n3_v32
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 4
#define SFX p64_n4_v32
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 5
#define SFX p64_n5_v32
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 6
#define SFX p64_n6_v32
#include "fm_query.inc"
#undef NPL
#undef SFX
#undef VEC_T
#undef VBITS
#define VEC_T uint64_t
#define VBITS 64
#define NPL 2
#define SFX p64_n2_v64
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 3
#define SFX p64_n3_v64
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 4
#define SFX p64_n4_v64
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 5
#define SFX p64_n5_v64
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 6
#define SFX p64_n6_v64
#include "fm_query.inc"
#undef NPL
#undef SFX
#undef VEC_T
#undef VBITS
#define VEC_T u128_t
#define VBITS 128
#define NPL 2
#define SFX p64_n2_v128
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 3
#define SFX p64_n3_v128
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 4
#define SFX p64_n4_v128
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 5
#define SFX p64_n5_v128
#include "fm_query.inc"
#undef NPL
#undef SFX
#define NPL 6
#define SFX p64_n6_v128
#include "fm_query.inc"
#undef NPL
#undef SFX
#undef VEC_T
#undef VBITS
#undef POS_T

typedef struct ora_vtable {
    void (*pos_range)(const void* view, const uint8_t* pat, uint64_t len, int reversed, uint64_t* sp, uint64_t* ep);
    uint64_t (*locate_row)(const void* view, uint64_t row);
    void* (*make_view)(const uint8_t* blob, const ora_layout* L, int has_table);
    void (*encode_bwm_body)(const uint8_t* bwt, uint64_t n, uint32_t S, uint64_t sentinel_index, uint8_t* blob, const ora_layout* L);
    void (*count_and_encode_text)(uint8_t* text, uint64_t n, const uint8_t* table, uint8_t* blob, const ora_layout* L);
    void (*write_sampled_sa)(const int32_t* sa_full, uint64_t n, uint32_t ratio, uint8_t* blob, const ora_layout* L);
} ora_vtable;

#define VT(sfx) { pos_range_any_##sfx, locate_row_any_##sfx, make_view_##sfx, encode_bwm_body_##sfx, \
                  count_and_encode_text_##sfx, write_sampled_sa_##sfx }
#define VT_ROW(p, v) VT(p##_n2_##v), VT(p##_n3_##v), VT(p##_n4_##v), VT(p##_n5_##v), VT(p##_n6_##v)
/* [pos: 0=u32 1=u64][vec: 0=32 1=64 2=128][planes-2] */
static const ora_vtable VTABLES[2][3][5] = {
    { { VT_ROW(p32, v32) }, { VT_ROW(p32, v64) }, { VT_ROW(p32, v128) } },
    { { VT_ROW(p64, v32) }, { VT_ROW(p64, v64) }, { VT_ROW(p64, v128) } },
};

static const ora_vtable* vtable_of(ora_type t) {
    int pi = t.pos_bits == 64 ? 1 : 0;
    int vi = t.vec_bits == 32 ? 0 : (t.vec_bits == 64 ? 1 : 2);
    return &VTABLES[pi][vi][t.planes - 2];
}

/* ------------------------------------------------------------------------------------------
 * FmIndexBuilder::build (builder/mod.rs:187-264)
 * ---------------------------------------------------------------------------------------- */
int ora_build(ora_type t, const uint8_t* text_in, uint64_t text_len, uint32_t symbol_count,
              const uint8_t* table256, uint32_t kmer_size, uint32_t sampling_ratio,
              uint8_t* blob, uint64_t blob_len, uint64_t detail[2]) {
    ora_layout L;
    int rc = ora_builder_layout(t, text_len, symbol_count, kmer_size, sampling_ratio, &L, detail);
    if (rc) return rc;
    if (t.encoder && !table256) return ORA_ERR_BAD_TYPE;
    if (((uintptr_t)blob) % L.align != 0) {                 /* builder/mod.rs:198-203 */
        if (detail) { detail[0] = L.align; detail[1] = ((uintptr_t)blob) % L.align; }
        return ORA_ERR_NOT_ALIGNED;
    }
    if (blob_len != L.total_size) {                         /* builder/mod.rs:205-209 */
        if (detail) { detail[0] = L.total_size; detail[1] = blob_len; }
        return ORA_ERR_INVALID_BLOB_SIZE;
    }
    if (text_len >= 0x7ffffffeull) return ORA_ERR_TOO_LARGE;
    const ora_vtable* vt = vtable_of(t);
    memset(blob, 0, (size_t)blob_len);

    /* 1) headers (builder/mod.rs:211-231) */
    static const uint8_t MAGIC[8] = { 'F', 'I', '0', '0', 0, 0, 0, 0 }; /* magic_number.rs:3-27 */
    memcpy(blob, MAGIC, 8);
    if (t.encoder) memcpy(blob + L.off_encoder, table256, 256);
    {
        uint8_t* h = blob + L.off_count_header;
        memcpy(h + 0, &L.symbol_count, 4);
        memcpy(h + 4, &L.kmer_size, 4);
        memcpy(h + 8, &L.count_array_len, 4);
        memcpy(h + 12, &L.kmer_multiplier_len, 4);
        memcpy(h + 16, &L.kmer_count_table_len, 8);
        h = blob + L.off_sa_header;
        memcpy(h + 0, &L.sampling_ratio, 4);
        memcpy(h + 8, &L.suffix_array_len, 8);
        h = blob + L.off_bwm_header;
        memcpy(h + 0, &L.symbol_count, 4);
        memcpy(h + 8, &L.rank_checkpoints_len, 8);
        memcpy(h + 16, &L.blocks_len, 8);
    }

    /* 2) bodies */
    uint8_t* text = (uint8_t*)malloc((size_t)text_len + 1);
    int32_t* sa = (int32_t*)malloc(sizeof(int32_t) * ((size_t)text_len + 1));
    uint8_t* bwt = (uint8_t*)malloc((size_t)text_len + 1);
    if (!text || !sa || !bwt) { free(text); free(sa); free(bwt); return ORA_ERR_NOMEM; }
    memcpy(text, text_in, (size_t)text_len);
    vt->count_and_encode_text(text, text_len, t.encoder ? table256 : NULL, blob, &L);

    /* get_compressed_suffix_array_and_pidx_while_bwt (crate_bio_manual/mod.rs:8-25):
     * push sentinel 0, SA over n+1 rows, bwt[i] = s[sa[i]-1] (sa[i]==0 -> the sentinel),
     * pidx = row whose BWT char is the sentinel; remove it; drop SA row 0. */
    text[text_len] = 0;
    int32_t n1 = (int32_t)text_len + 1;
    rc = ora_suffix_array(text, n1, (int32_t)symbol_count, sa);
    if (rc) { free(text); free(sa); free(bwt); return rc; }
    uint64_t pidx = 0, w = 0;
    for (int32_t i = 0; i < n1; i++) {
        if (sa[i] == 0) { pidx = (uint64_t)i; continue; }
        bwt[w++] = text[sa[i] - 1];
    }
    vt->write_sampled_sa(sa, text_len, sampling_ratio, blob, &L);
    vt->encode_bwm_body(bwt, text_len, symbol_count, pidx, blob, &L);
    free(text); free(sa); free(bwt);
    return ORA_OK;
}

/* ------------------------------------------------------------------------------------------
 * FmIndex::load (load_from_blob.rs:28-85)
 * ---------------------------------------------------------------------------------------- */
struct ora_index {
    ora_type type;
    ora_layout layout;
    const uint8_t* blob;
    const ora_vtable* vt;
    void* view;
    uint64_t text_len;
};

int ora_load(const uint8_t* blob, uint64_t blob_len, ora_type t, ora_index** out, uint64_t detail[2]) {
    if (!type_ok(t)) return ORA_ERR_BAD_TYPE;
    ora_layout L;
    memset(&L, 0, sizeof(L));
    /* header offsets do not depend on header contents */
    layout_offsets(t, &L);
    /* the reference unwrap()-panics on a blob shorter than its headers (components/mod.rs:19);
     * the restatement reports InvalidFormat instead */
    if (blob_len < L.header_size) return ORA_ERR_INVALID_FORMAT;
    if (((uintptr_t)blob) % L.align != 0) return ORA_ERR_INVALID_FORMAT; /* zerocopy ref_from_bytes().unwrap() */
    /* MagicNumber::is_valid && is_supported_version (magic_number.rs:38-49) */
    if (!(blob[0] == 'F' && blob[1] == 'I' && blob[2] == '0' && blob[3] == '0')) return ORA_ERR_INVALID_FORMAT;
    const uint8_t* h = blob + L.off_count_header;
    memcpy(&L.symbol_count, h + 0, 4);
    memcpy(&L.kmer_size, h + 4, 4);
    memcpy(&L.count_array_len, h + 8, 4);
    memcpy(&L.kmer_multiplier_len, h + 12, 4);
    memcpy(&L.kmer_count_table_len, h + 16, 8);
    h = blob + L.off_sa_header;
    memcpy(&L.sampling_ratio, h + 0, 4);
    memcpy(&L.suffix_array_len, h + 8, 8);
    h = blob + L.off_bwm_header;
    uint32_t bwm_symbol_count;
    memcpy(&bwm_symbol_count, h + 0, 4);
    memcpy(&L.rank_checkpoints_len, h + 8, 8);
    memcpy(&L.blocks_len, h + 16, 8);
    layout_offsets(t, &L);
    if (L.total_size != blob_len) {                         /* load_from_blob.rs:46-58 */
        if (detail) { detail[0] = L.total_size; detail[1] = blob_len; }
        return ORA_ERR_BLOB_SIZE;
    }
    ora_index* ix = (ora_index*)calloc(1, sizeof(ora_index));
    if (!ix) return ORA_ERR_NOMEM;
    ix->type = t;
    ix->layout = L;
    ix->blob = blob;
    ix->vt = vtable_of(t);
    /* BwmView takes its row stride from the BWM header's own symbol_count (bwm/mod.rs:158) */
    ora_layout Lv = L;
    Lv.symbol_count = bwm_symbol_count;
    ix->view = ix->vt->make_view(blob, &Lv, (int)t.encoder);
    if (!ix->view) { free(ix); return ORA_ERR_NOMEM; }
    /* text length is not stored; it equals count_array[S] (count_array.rs:117,125) */
    if (t.pos_bits == 32) ix->text_len = ((const uint32_t*)(blob + L.off_count_array))[L.count_array_len - 1];
    else ix->text_len = ((const uint64_t*)(blob + L.off_count_array))[L.count_array_len - 1];
    *out = ix;
    return ORA_OK;
}

void ora_free(ora_index* ix) {
    if (!ix) return;
    free(ix->view);
    free(ix);
}

const ora_layout* ora_index_layout(const ora_index* ix) { return &ix->layout; }
uint64_t ora_text_len(const ora_index* ix) { return ix->text_len; }

/* ------------------------------------------------------------------------------------------
 * count / locate entry points (locate/with_slice.rs:5-18, locate/with_rev_iter.rs:5-18)
 * ---------------------------------------------------------------------------------------- */
int ora_pos_range(const ora_index* ix, const uint8_t* pat, uint64_t len, int reversed, uint64_t* sp, uint64_t* ep) {
    if (len == 0) return ORA_ERR_EMPTY_PATTERN;
    ix->vt->pos_range(ix->view, pat, len, reversed, sp, ep);
    return ORA_OK;
}

int ora_count(const ora_index* ix, const uint8_t* pat, uint64_t len, int reversed, uint64_t* count) {
    uint64_t sp, ep;
    int rc = ora_pos_range(ix, pat, len, reversed, &sp, &ep);
    if (rc) return rc;
    *count = ep - sp;
    return ORA_OK;
}

int ora_locate(const ora_index* ix, const uint8_t* pat, uint64_t len, int reversed,
               uint64_t* out, uint64_t cap, uint64_t* n_out) {
    uint64_t sp, ep;
    int rc = ora_pos_range(ix, pat, len, reversed, &sp, &ep);
    if (rc) return rc;
    uint64_t w = *n_out;
    /* Position::as_vec_in_range(from, to): empty when from >= to (text_length.rs:83) */
    for (uint64_t pos = sp; pos < ep; pos++) {
        if (w < cap) out[w] = ix->vt->locate_row(ix->view, pos);
        w++;
    }
    *n_out = w;
    return ORA_OK;
}

/* ------------------------------------------------------------------------------------------
 * Pattern-parallel drivers (CPU baseline): reference algorithm, one call per pattern.
 * ---------------------------------------------------------------------------------------- */
typedef struct batch_job {
    const ora_index* ix;
    const uint8_t* pats;
    uint64_t lo, hi, len;
    uint64_t* counts;
    const uint64_t* offs;
    void* pos_out;
    uint32_t pos_out_bits;
    uint64_t checksum;
} batch_job;

static void* count_worker(void* arg) {
    batch_job* j = (batch_job*)arg;
    for (uint64_t i = j->lo; i < j->hi; i++) {
        uint64_t sp, ep;
        j->ix->vt->pos_range(j->ix->view, j->pats + i * j->len, j->len, 0, &sp, &ep);
        j->counts[i] = ep - sp;
    }
    return NULL;
}

/* mix used by the order-independent checksum: sum over patterns i and their locations loc of
 * (loc + 1) * (2 i + 1) mod 2^64 (the GPU side computes the same quantity). */
static void* locate_worker(void* arg) {
    batch_job* j = (batch_job*)arg;
    uint64_t acc = 0;
    for (uint64_t i = j->lo; i < j->hi; i++) {
        uint64_t sp, ep;
        j->ix->vt->pos_range(j->ix->view, j->pats + i * j->len, j->len, 0, &sp, &ep);
        if (j->counts) j->counts[i] = ep - sp;
        uint64_t w = j->offs ? j->offs[i] : 0;
        for (uint64_t pos = sp; pos < ep; pos++, w++) {
            uint64_t loc = j->ix->vt->locate_row(j->ix->view, pos);
            acc += (loc + 1) * (2 * i + 1);
            if (j->pos_out) {
                if (j->pos_out_bits == 64) ((uint64_t*)j->pos_out)[w] = loc;
                else ((uint32_t*)j->pos_out)[w] = (uint32_t)loc;
            }
        }
    }
    j->checksum = acc;
    return NULL;
}

static int run_batch(batch_job proto, uint64_t n, int threads, void* (*fn)(void*), uint64_t* checksum) {
    if (checksum) *checksum = 0;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n && n > 0) threads = (int)n;
    if (n == 0) return ORA_OK;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    batch_job* jobs = (batch_job*)malloc(sizeof(batch_job) * (size_t)threads);
    if (!th || !jobs) { free(th); free(jobs); return ORA_ERR_NOMEM; }
    for (int k = 0; k < threads; k++) {
        jobs[k] = proto;
        jobs[k].lo = n * (uint64_t)k / (uint64_t)threads;
        jobs[k].hi = n * (uint64_t)(k + 1) / (uint64_t)threads;
    }
    for (int k = 1; k < threads; k++) pthread_create(&th[k], NULL, fn, &jobs[k]);
    fn(&jobs[0]);
    for (int k = 1; k < threads; k++) pthread_join(th[k], NULL);
    if (checksum) for (int k = 0; k < threads; k++) *checksum += jobs[k].checksum;
    free(th); free(jobs);
    return ORA_OK;
}

int ora_count_batch(const ora_index* ix, const uint8_t* pats, uint64_t n, uint64_t len,
                    uint64_t* counts_out, int threads) {
    if (len == 0) return ORA_ERR_EMPTY_PATTERN;
    batch_job j;
    memset(&j, 0, sizeof(j));
    j.ix = ix; j.pats = pats; j.len = len; j.counts = counts_out;
    return run_batch(j, n, threads, count_worker, NULL);
}

int ora_locate_batch(const ora_index* ix, const uint8_t* pats, uint64_t n, uint64_t len,
                     uint64_t* counts_out, const uint64_t* out_offs, void* pos_out, uint32_t pos_out_bits,
                     uint64_t* checksum_out, int threads) {
    if (len == 0) return ORA_ERR_EMPTY_PATTERN;
    if (pos_out && !out_offs) return ORA_ERR_BAD_TYPE;
    batch_job j;
    memset(&j, 0, sizeof(j));
    j.ix = ix; j.pats = pats; j.len = len; j.counts = counts_out; j.offs = out_offs;
    j.pos_out = pos_out; j.pos_out_bits = pos_out_bits;
    return run_batch(j, n, threads, locate_worker, checksum_out);
}
