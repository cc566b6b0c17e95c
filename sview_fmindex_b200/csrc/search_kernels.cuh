// search_kernels.cuh -- hand-written CUDA (sm_100a) for the hot path: backward-search count and
// SA-sampled locate over a batch of patterns.  HBM-bound integer work: no tensor cores; TMA only as bulk copies that
// stage pattern bytes in shared memory (pack_sweep_kernel).
// Citations are relative to the reference's sview-fmindex/src/.
#pragma once
#include "device_index.cuh"

namespace svfm {

enum : int { ERRBIT_BAD_SYMBOL = 1, ERRBIT_EMPTY_PATTERN = 2 };

struct PatternBatch {
    const uint8_t* pats;   // concatenated pattern bytes
    const uint64_t* offs;  // n+1 offsets, or NULL for fixed-length patterns
    uint64_t n;
    uint32_t fixed_len;
    uint32_t reversed;     // patterns stored back-to-front (rev-iter twins, locate/with_rev_iter.rs)
    uint32_t preencoded;   // bytes are symbol indices already (packed entry points): the encoding table is skipped
    // packed_bits > 0: `pats` still holds the caller's PACKED patterns (packed_bpp bytes each, see unpack_patterns_kernel) and
    // fixed_len is the number of symbols; search_kernel unpacks them while staging a CTA's patterns in shared memory, so the
    // one-byte-per-symbol copy never exists in HBM.  Only the plain kernel reads this form (svfm_api.cu decides).
    uint32_t packed_bits = 0, packed_bpp = 0;
};

// Packed fixed-length patterns (svfm_*_batch_packed): symbol indices, `bits` bits each, first symbol in the lowest bits
// of the first byte, every pattern padded to a whole number of bytes (bpp).  Expands them to one byte per symbol.
static __global__ void __launch_bounds__(256)
unpack_patterns_kernel(const uint8_t* __restrict__ packed, uint64_t n, uint32_t len, uint32_t bits, uint32_t bpp,
                       uint8_t* __restrict__ out) {
    const uint64_t total = n * (uint64_t)len;
    const uint32_t mask = (1u << bits) - 1u;
    // one thread produces 4 consecutive output bytes (the output buffer is padded to a multiple of 4)
    for (uint64_t o = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; o < total; o += (uint64_t)gridDim.x * blockDim.x * 4) {
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const uint64_t at = o + b;
            if (at < total) {
                const uint64_t i = at / len;
                const uint32_t bit = (uint32_t)(at - i * len) * bits;
                const uint8_t* p = packed + i * (uint64_t)bpp + (bit >> 3);
                uint32_t v = p[0];
                if ((bit & 7u) + bits > 8u) v |= (uint32_t)p[1] << 8;   // bits <= 8: a symbol spans at most two bytes
                word |= ((v >> (bit & 7u)) & mask) << (8 * b);
            }
        }
        *reinterpret_cast<uint32_t*>(out + o) = word;
    }
}

constexpr int SEARCH_THREADS = 256;
constexpr uint32_t HEAVY_ROWS = 1024;  // patterns with more SA rows than this are located row-parallel
constexpr uint32_t SB_SHIFT = 12;      // bucketed sort-back: a bucket = 2^12 consecutive pattern indices (see SbRec below)
constexpr uint32_t SB_BUCKET = 1u << SB_SHIFT;

// Bucketed sort-back, sizing side (filled by the search kernels; see SbRec below).
struct SbOut {
    uint32_t* hist;             // per bucket: SA rows of its patterns (NULL: no bucketed sort-back for this batch)
    unsigned long long* total;  // all rows of the batch (64 bits: tells the host when the 32-bit counters have wrapped)
};

// One backward-search step for both range ends: FmIndex::next_pos_range (locate/mod.rs:39-45) =
// count_array[s] + get_next_rank(pos, s) (bwm/mod.rs:197-215) for pos in {sp, ep}; c = count_array[sym].
// Straight-line code: the block and checkpoint word of sp are always fetched, those of ep only when ep falls into another
// block (predicated loads); the match mask and prefix popcount then run for both ends unconditionally.  The round-1
// version branched into a one-block and a two-block path; inside a warp some lane almost always needs the second block
// (a 4-row interval straddles a 64-row block boundary with probability 1/16, so 86 % of all warps), and every warp then
// executed both paths: 98 instructions per warp and step, against ~60 here (ncu source view, sweep_round_kernel).
template <class P, int NPL, int VBITS, bool ILV = false>
__device__ __forceinline__ void backward_step(const DevIndex<P>& ix, uint32_t sym, P c, P& sp, P& ep) {
    using B = Block<NPL, VBITS>;
    using Q = typename QuotientOf<P>::type;   // block number: 32 bits are enough for 32-bit positions (compares, moves)
    Q q0, q1;
    uint32_t r0, r1;
    rank_addr<P, VBITS>(ix, sp, q0, r0);
    rank_addr<P, VBITS>(ix, ep, q1, r1);
    B b0, b1;
    P ck0, ck1 = 0;
    if constexpr (ILV) {
        const uint8_t* e0 = ix.ilv + (uint64_t)q0 * ix.ilv_stride;
        ck0 = ld_gather<P>(reinterpret_cast<const P*>(e0 + ix.ilv_ck_off) + sym);
        b0.load_aligned(e0);
    } else {
        ck0 = ld_gather<P>(ix.rank_checkpoints + ((uint64_t)q0 * ix.symbol_count + sym));
        b0.load(ix.blocks, q0);
    }
    // The second block goes into registers of its own (cleared, not copied from b0): a predicated load into a copy of b0
    // has to wait for b0's load to land first, which serialised the two fetches (ncu: a quarter of the round kernel's
    // stall samples sat on that copy).
    const bool two = q1 != q0;
    b1.clear();
    if (two) {
        if constexpr (ILV) {
            const uint8_t* e1 = ix.ilv + (uint64_t)q1 * ix.ilv_stride;
            ck1 = ld_gather<P>(reinterpret_cast<const P*>(e1 + ix.ilv_ck_off) + sym);
            b1.load_aligned(e1);
        } else {
            ck1 = ld_gather<P>(ix.rank_checkpoints + ((uint64_t)q1 * ix.symbol_count + sym));
            b1.load(ix.blocks, q1);
        }
    }
    typename B::W flip[NPL];
    B::flips(sym, flip);
    typename B::W m0[VecTraits<VBITS>::WORDS], m1[VecTraits<VBITS>::WORDS];
    b0.match_flips(flip, m0);
    b1.match_flips(flip, m1);
    sp = (P)(c + ck0 + (P)B::prefix_count(m0, r0));
#pragma unroll
    for (int k = 0; k < VecTraits<VBITS>::WORDS; k++) m1[k] = two ? m1[k] : m0[k];
    ep = (P)(c + (two ? ck1 : ck0) + (P)B::prefix_count(m1, r1));
    SVFM_ASSERT(sp <= ep);   // ranks are monotone (with_slice.rs:27: the loop condition relies on it)
}

// ---- sweep occ copy -------------------------------------------------------------------------------------------------------
// The sweep rounds are bound by the number of L1 requests and sectors they touch, not by DRAM (ncu, round 0 of a 10^8-pattern
// batch: l1tex throughput 86 %, 14 sector lookups per item -- a rank query on the blob is a checkpoint word plus three plane
// words from two arrays, for each end of the interval).  For indexes with 64-bit vectors, at most three planes and at most
// four occurring symbols (DNA) the engine keeps a third form of the occ data: entry q = ONE 32-byte sector holding the planes
// of block q (words 0 .. NPL-1) and, in word NPL, the checkpoint counts of the occurring symbols as 16-bit deltas against
// the checkpoint row of the block's SUPERBLOCK (SWP_SUPER consecutive blocks; that row is read from the blob in place --
// one row per 32 768 text positions, a few hundred KB in all, resident in L1/L2).  A rank query is then one 256-bit load.
// Derived from the blob, bytes only: rank = superblock row + delta + popcount, the same sum bwm/mod.rs:197-215 forms.
constexpr uint32_t SWP_SUPER_SHIFT = 9;                       // 512 blocks of 64 positions: deltas stay below 2^15
constexpr uint32_t SWP_SUPER_MASK = (1u << SWP_SUPER_SHIFT) - 1u;
struct SwpEntry { unsigned long long w[4]; };
__device__ __forceinline__ SwpEntry swp_load(const uint8_t* entry) {
    SwpEntry e;
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(e.w[0]), "=l"(e.w[1]), "=l"(e.w[2]), "=l"(e.w[3]) : "l"(entry));
    return e;
}
struct SwpSyms { uint8_t present[4]; uint32_t s_eff; };
template <class P>
__global__ void __launch_bounds__(256)
swp_build_kernel(const unsigned long long* __restrict__ blocks, uint32_t npl, const P* __restrict__ ck, uint32_t S, SwpSyms syms,
                 uint64_t blocks_len, ulonglong4* __restrict__ out) {
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < blocks_len; q += (uint64_t)gridDim.x * blockDim.x) {
        unsigned long long w[4] = {0, 0, 0, 0};
        for (uint32_t v = 0; v < npl; v++) w[v] = blocks[q * npl + v];
        const uint64_t sup = q & ~(uint64_t)SWP_SUPER_MASK;
        unsigned long long d = 0;
        for (uint32_t j = 0; j < syms.s_eff; j++) {
            const uint64_t delta = (uint64_t)(ck[q * S + syms.present[j]] - ck[sup * S + syms.present[j]]);
            SVFM_ASSERT(delta < 65536);
            d |= delta << (16 * j);
        }
        w[npl] = d;
        out[q] = make_ulonglong4(w[0], w[1], w[2], w[3]);
    }
}

// backward_step on the sweep occ copy: `code` = rank of the symbol among the occurring ones, sym = its symbol index.
template <class P, int NPL>
__device__ __forceinline__ void backward_step_swp(const DevIndex<P>& ix, uint32_t code, uint32_t sym, P c, P& sp, P& ep) {
    using B = Block<NPL, 64>;
    using Q = typename QuotientOf<P>::type;
    Q q0, q1;
    uint32_t r0, r1;
    rank_addr<P, 64>(ix, sp, q0, r0);
    rank_addr<P, 64>(ix, ep, q1, r1);
    const SwpEntry e0 = swp_load(ix.swp + (uint64_t)q0 * 32);
    const P sup0 = ld_gather<P>(ix.rank_checkpoints + ((uint64_t)(q0 & ~(Q)SWP_SUPER_MASK) * ix.symbol_count + sym));
    const bool two = q1 != q0;
    const bool two_sup = ((q0 ^ q1) >> SWP_SUPER_SHIFT) != 0;
    SwpEntry e1;
    e1.w[0] = e1.w[1] = e1.w[2] = e1.w[3] = 0;
    P sup1 = 0;
    if (two) e1 = swp_load(ix.swp + (uint64_t)q1 * 32);
    if (two_sup) sup1 = ld_gather<P>(ix.rank_checkpoints + ((uint64_t)(q1 & ~(Q)SWP_SUPER_MASK) * ix.symbol_count + sym));
    unsigned long long flip[NPL];
#pragma unroll
    for (int v = 0; v < NPL; v++) flip[v] = ((sym >> v) & 1u) ? 0ull : ~0ull;
    unsigned long long m0 = e0.w[0] ^ flip[0], m1 = e1.w[0] ^ flip[0];
#pragma unroll
    for (int v = 1; v < NPL; v++) { m0 &= e0.w[v] ^ flip[v]; m1 &= e1.w[v] ^ flip[v]; }
    const uint32_t sh = 16u * code;
    const uint32_t d0 = (uint32_t)(e0.w[NPL] >> sh) & 0xffffu, d1 = (uint32_t)(e1.w[NPL] >> sh) & 0xffffu;
    sp = (P)(c + sup0 + (P)d0 + (P)popc_top((uint64_t)m0, r0));
    m1 = two ? m1 : m0;
    ep = (P)(c + (two_sup ? sup1 : sup0) + (P)(two ? d1 : d0) + (P)popc_top((uint64_t)m1, r1));
    SVFM_ASSERT(sp <= ep);
    (void)sizeof(B);
}

// Locality key of a pattern = its trailing symbols packed `bits` per symbol from the top of a u64, the
// LAST symbol most significant (backward search consumes the pattern from its end, so patterns that share
// a suffix walk the same checkpoint rows and blocks for as many steps as the shared suffix is long).
// The key doubles as the encoded pattern: the search kernel takes the last min(len, 64/bits) symbols from
// it and never touches the pattern bytes again unless the pattern is longer.  Also validates the batch.
static __global__ void __launch_bounds__(SEARCH_THREADS)
pack_keys_kernel(const uint8_t* __restrict__ table, uint32_t S, const PatternBatch pb, uint32_t bits,
                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int* __restrict__ err) {
    __shared__ uint8_t s_table[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_table[i] = (table && !pb.preencoded) ? table[i] : (uint8_t)i;
    __syncthreads();
    const uint32_t m = 64u / bits;
    int errbits = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pb.n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t base, len;
        if (pb.offs) { base = pb.offs[i]; len = pb.offs[i + 1] - base; }
        else { base = i * (uint64_t)pb.fixed_len; len = pb.fixed_len; }
        const uint8_t* p = pb.pats + base;
        if (len == 0) errbits |= ERRBIT_EMPTY_PATTERN;
        const uint32_t take = len < m ? (uint32_t)len : m;
        uint64_t key = 0;
        for (uint32_t j = 0; j < take; j++) {
            const uint64_t fwd = len - 1 - j;  // j-th symbol from the end
            uint32_t s = s_table[__ldg(p + (pb.reversed ? (len - 1 - fwd) : fwd))];
            if (s >= S) { errbits |= ERRBIT_BAD_SYMBOL; s = S - 1; }
            key |= (uint64_t)s << (64u - bits * (j + 1));
        }
        keys[i] = key;
        vals[i] = (uint32_t)i;
    }
    if (errbits) atomicOr(err, errbits);
}

template <class P>
struct SearchIO {
    const uint64_t* keys;   // packed keys in work order, or NULL (symbols come from the pattern bytes)
    const uint32_t* idx;    // pattern index of each work item, or NULL (identity)
    uint32_t bits;          // bits per symbol in the key
    P* sp_out;              // work order, nullable
    P* cnt_out;             // work order, nullable
    unsigned long long* heavy_seen;  // nullable
    int* err;
    SbOut sb;
    uint8_t* resolved_out;  // work order, nullable: 1 = sp_out holds the TEXT POSITION of the single occurrence (text verification)
};

// FmIndex::write_locations_to_buffer (locate/mod.rs:14-37) for ONE SA row: LF-walk to the nearest
// sampled row (BwmView::get_pre_rank_and_symidx, bwm/mod.rs:217-236), then
// SuffixArrayView::get_location_of (suffix_array/mod.rs:100-105).
// One LF step from SA row `pos` (not the row of the sentinel): BwmView::get_pre_rank_and_symidx (bwm/mod.rs:217-236) followed
// by count_array[s] + rank, i.e. the row of the suffix that starts one text position earlier.
template <class P, int NPL, int VBITS, bool ILV>
__device__ __forceinline__ P lf_step(const DevIndex<P>& ix, const P* __restrict__ s_count, P pos) {
    uint64_t q;
    uint32_t rem;
    rank_addr<P, VBITS>(ix, pos, q, rem);
    Block<NPL, VBITS> b;
    P ck;
    uint32_t s;
    if constexpr (ILV) {
        const uint8_t* e = ix.ilv + q * ix.ilv_stride;  // block and checkpoint row arrive with one fetch
        b.load_aligned(e);
        if (sizeof(P) == 4 && ix.symbol_count <= 8 && (ix.ilv_ck_off & 7u) == 0) {   // (row end stays inside the slot: DESIGN.md)
            // Which checkpoint word is needed depends on the symbol inside the block: load the whole row (at most 8
            // words) together with the block instead of after it -- one memory round trip per LF step instead of two.
            const unsigned long long* rw = reinterpret_cast<const unsigned long long*>(e + ix.ilv_ck_off);   // 8-byte aligned
            const uint32_t pairs = (ix.symbol_count + 1) >> 1;
            unsigned long long r2[4] = {0, 0, 0, 0};
#pragma unroll
            for (uint32_t i = 0; i < 4; i++) if (i < pairs) r2[i] = __ldg(rw + i);
            s = b.symidx_of(rem);
            const unsigned long long pair = s < 4 ? (s < 2 ? r2[0] : r2[1]) : (s < 6 ? r2[2] : r2[3]);
            ck = (P)((s & 1u) ? (uint32_t)(pair >> 32) : (uint32_t)pair);
        } else {
            s = b.symidx_of(rem);
            ck = ld_gather<P>(reinterpret_cast<const P*>(e + ix.ilv_ck_off) + s);
        }
    } else {
        b.load(ix.blocks, q);
        s = b.symidx_of(rem);
        ck = ld_gather<P>(ix.rank_checkpoints + q * ix.symbol_count + s);
    }
    SVFM_ASSERT(s < ix.symbol_count);
    return (P)(s_count[s] + ck + (P)b.remain_count(rem, s));  // remain_count(0, .) == 0
}

// Is SA row `pos` sampled (pos % sampling_ratio == 0)?  quot = pos / sampling_ratio.
template <class P>
__device__ __forceinline__ bool sa_sampled(const DevIndex<P>& ix, P pos, uint64_t& quot) {
    if (ix.ratio_mask != 0xffffffffu) {
        quot = (uint64_t)pos >> ix.ratio_shift;
        return ((uint32_t)pos & ix.ratio_mask) == 0;
    }
    quot = (uint64_t)pos / ix.sampling_ratio;
    return quot * ix.sampling_ratio == (uint64_t)pos;
}

// ---- expanded suffix array ----------------------------------------------------------------------------------------------
// The blob samples the suffix array (every r-th row) because the reference trades CPU memory for LF-walk time
// (suffix_array/mod.rs:43-70, :100-105).  A B200 has 180 GB: at load the engine walks every row ONCE (fsa_build_kernel:
// locate_row on the sampled array) and keeps the text position of all rows next to the blob -- 4 bytes per text symbol,
// built only when that is a small share of the free device memory (SVFM_TUNE_FULL_SA).  `locate` then costs one read per
// row instead of 1 + (r-1)/2 .. r dependent random cache lines, and a batch in SA order reads the array as a stream.
// Results are identical by construction: the entries ARE what locate/mod.rs:21-33 computes for that row.
template <class P>
__device__ __forceinline__ P fsa_get(const DevIndex<P>& ix, P row) {
    if (sizeof(P) == 4 || ix.fsa32) return (P)__ldg(reinterpret_cast<const uint32_t*>(ix.fsa) + row);
    return (P)__ldg(reinterpret_cast<const unsigned long long*>(ix.fsa) + row);
}

template <class P, int NPL, int VBITS, bool ILV>
__device__ __forceinline__ P locate_row(const DevIndex<P>& ix, const P* __restrict__ s_count, P pos) {
    if (ix.fsa) return fsa_get<P>(ix, pos);
    P offset = 0;
    for (;;) {
        uint64_t quot;
        if (sa_sampled<P>(ix, pos, quot)) return (P)(ld_gather<P>(ix.suffix_array + quot) + offset);
        if (pos == (P)(ix.sentinel_index - 1)) return offset;  // None arm, locate/mod.rs:27-30
        pos = lf_step<P, NPL, VBITS, ILV>(ix, s_count, pos);
        offset += 1;
    }
}

// ---- text verification ---------------------------------------------------------------------------------------------
// Once the SA interval of a pattern's suffix has shrunk to a handful of rows while many symbols are still to come (150 bp
// reads: after ~16 of 150 symbols on a 1 Gbp text), walking the remaining symbols through the index costs one random
// cache line per symbol.  The engine keeps a packed copy of the TEXT next to the blob (derived from the blob at load:
// every SA row's text position comes from locate_row, its BWT symbol precedes that position) and finishes such patterns
// by locating each candidate row and comparing the unread prefix of the pattern against the text in front of it: two or
// three sectors instead of a hundred lines.  At most one candidate can then survive in the common case, and the pattern
// is RESOLVED: count 0, or count 1 with the text position known (no SA row is needed any more).  When two or more
// candidates survive -- a true repeat of the whole pattern -- the result must come in SA-row order of the full pattern
// (locate/mod.rs:19), which only the index knows, so the backward search simply continues.  Results are identical either
// way; tests run with the text copy on and off.
constexpr uint32_t VERIFY_MAX_ROWS = 4;

__device__ __forceinline__ uint32_t text_symbol(const uint32_t* __restrict__ text, uint32_t bits, uint64_t i) {
    const uint64_t bit = i * bits;
    return (__ldg(text + (bit >> 5)) >> (uint32_t)(bit & 31u)) & ((1u << bits) - 1u);
}

// FmIndex::get_pos_range (locate/with_slice.rs:21-33) for ONE non-empty pattern p[0, len) (stored back to front when
// `reversed`).  With in_key > 0 the last in_key symbols come out of `key` (see pack_keys_kernel).
// Seeding: patterns at least ext_m symbols long start from the extended k-mer table (one lookup resolves the last
// ext_m symbols -- the device twin of the reference's kLTS, count_array.rs:203-233, with a longer k chosen for
// HBM instead of for a CPU cache); shorter ones use the blob's own kLTS.
template <class P, int NPL, int VBITS, bool ILV>
__device__ __forceinline__ void search_pattern(const DevIndex<P>& ix, const uint8_t* p, uint64_t len, bool reversed,
                                               uint64_t key, uint32_t in_key, uint32_t bits, const uint8_t* s_table,
                                               const uint8_t* s_rank, const uint8_t* s_lut, const P* s_count, int& errbits,
                                               P& sp, P& ep, bool& resolved) {
    const uint32_t S = ix.symbol_count;
    const uint32_t k = ix.kmer_size;
    const uint64_t sym_mask = (1ull << bits) - 1;
    uint64_t pi = 0;
    sp = 0;
    ep = 0;
    // logical (forward) symbol j of the pattern
    auto sym_at = [&](uint64_t j) -> uint32_t {
        const uint64_t from_end = len - 1 - j;
        if (from_end < in_key) return (uint32_t)((key >> (64u - bits * ((uint32_t)from_end + 1))) & sym_mask);
        uint32_t s = s_table[p[reversed ? from_end : j]];   // plain load: p may point into the staged copy in shared memory
        if (s >= S) { errbits |= ERRBIT_BAD_SYMBOL; s = S - 1; }
        return s;
    };
    if (ix.ext && len >= ix.ext_m) {
        // extended table: index = trailing ext_m symbols as base-s_eff digits of their rank among the
        // symbols that occur in the text, first symbol most significant
        uint64_t e = 0;
        bool absent = false;
        for (uint32_t j = 0; j < ix.ext_m; j++) {
            const uint32_t r = s_rank[sym_at(len - ix.ext_m + j)];
            absent |= (r == 0xffu);
            e = e * ix.s_eff + r;
        }
        if (!absent) {
            const P* q = ix.ext + 2 * e;
            sp = __ldg(q);
            ep = (P)(sp + __ldg(q + 1));
        }
        pi = len - ix.ext_m;
    } else if (len < k) {
        // CountArrayView::get_initial_pos_range_and_idx_of_pattern (count_array.rs:203-233)
        uint64_t start = 0;
        for (uint64_t j = 0; j < len; j++) start += (uint64_t)(sym_at(j) + 1) * __ldg(ix.kmer_multiplier + j);
        const uint64_t end = start + __ldg(ix.kmer_multiplier + (len - 1)) - 1;
        sp = __ldg(ix.kmer_count_table + (start - 1));
        ep = __ldg(ix.kmer_count_table + end);
        pi = 0;
    } else {
        uint64_t start = 0;
        for (uint32_t j = 0; j < k; j++) start += (uint64_t)(sym_at(len - k + j) + 1) * __ldg(ix.kmer_multiplier + j);
        sp = __ldg(ix.kmer_count_table + (start - 1));
        ep = __ldg(ix.kmer_count_table + start);
        pi = len - k;
    }
    // LF mapping (with_slice.rs:27-31): stops as soon as the interval is empty
    resolved = false;
    bool may_verify = ix.text != nullptr;
    // accesses one candidate costs: its LF walk to a sampled row ((r-1)/2 steps on average), the SA word, ~2 text sectors
    const uint64_t verify_cost = (uint64_t)(ix.sampling_ratio - 1) / 2 + 3;
    uint32_t stalled = 0;
    for (;;) {
        // ---- backward steps, until the pattern is finished or text verification should take over.
        // One more backward step costs ~2 random lines and divides the candidates by the alphabet size on non-repetitive
        // text, so the search keeps stepping until ONE row is left; a small interval that has stopped shrinking (a repeat)
        // is verified as it is.  Not worth it when fewer steps are left than it costs.
        // (The lanes of a warp leave this loop after different numbers of steps and WAIT for each other behind it, so that
        // the long comparison below runs once per warp -- inside one loop the lanes entered it in different iterations and
        // the warp executed it four times over: ncu, 150 bp reads, 8.6 active lanes per instruction.)
        uint64_t rows = 0;
        bool verify = false;
        while (sp < ep && pi > 0) {
            rows = (uint64_t)(ep - sp);
            verify = may_verify && rows <= VERIFY_MAX_ROWS && (rows == 1 || stalled >= 2) && 2 * pi > rows * verify_cost;
            if (verify) break;
            pi -= 1;
            const uint32_t sy = sym_at(pi);
            backward_step<P, NPL, VBITS, ILV>(ix, sy, s_count[sy], sp, ep);
            stalled = (uint64_t)(ep - sp) == rows ? stalled + 1 : 0;
        }
        if (!verify) return;
        // ---- text verification (see above)
        uint64_t t0[VERIFY_MAX_ROWS];
        uint32_t alive = 0;
#pragma unroll
        for (uint32_t c = 0; c < VERIFY_MAX_ROWS; c++) {
            t0[c] = 0;
            if (c < rows) {
                const P pos = locate_row<P, NPL, VBITS, ILV>(ix, s_count, (P)(sp + c));
                SVFM_ASSERT((uint64_t)pos < (uint64_t)ix.count_array[ix.symbol_count]);   // a text position
                if ((uint64_t)pos >= pi) { t0[c] = (uint64_t)pos - pi; alive |= 1u << c; }   // else: it would start before the text
            }
        }
        // pattern[0, pi) against the text in front of every candidate, one 32-bit word of packed text per comparison;
        // the pattern's word is built once and shared by the candidates
        const uint32_t tb = ix.text_bits, per = 32u / tb;
        const bool plain = in_key == 0 && !reversed;   // the usual case: symbols straight from the pattern bytes
        for (uint64_t j0 = 0; j0 < pi && alive; j0 += per) {
            const uint32_t m = pi - j0 < per ? (uint32_t)(pi - j0) : per;
            uint32_t code = 0, flags = 0;
            if (plain) {
                const uint8_t* q = p + j0;
                for (uint32_t k2 = 0; k2 < m; k2++) {
                    const uint32_t v = s_lut[q[k2]];               // rank | 0x80 never occurs | 0x40 PassThrough byte >= S
                    flags |= v;
                    code |= (v & ((1u << tb) - 1u)) << (k2 * tb);
                }
            } else {
                for (uint32_t k2 = 0; k2 < m; k2++) {
                    const uint32_t rk = s_rank[sym_at(j0 + k2)];
                    flags |= rk & 0x80u;
                    code |= (rk & ((1u << tb) - 1u)) << (k2 * tb);
                }
            }
            if (flags & 0x40u) errbits |= ERRBIT_BAD_SYMBOL;
            if (flags & 0x80u) { alive = 0; break; }               // a symbol that never occurs in the text
            const uint32_t mask = m * tb >= 32u ? 0xffffffffu : (1u << (m * tb)) - 1u;
#pragma unroll
            for (uint32_t c = 0; c < VERIFY_MAX_ROWS; c++) {
                if (alive & (1u << c)) {
                    const uint64_t bit = (t0[c] + j0) * tb;
                    const uint32_t* tw = ix.text + (bit >> 5);
                    const uint32_t have = __funnelshift_r(__ldg(tw), __ldg(tw + 1), (uint32_t)(bit & 31u));   // array is padded
                    if ((have ^ code) & mask) alive &= ~(1u << c);
                }
            }
        }
        const uint32_t ok = (uint32_t)__popc(alive);
        if (ok <= 1) {
            P okpos = 0;
#pragma unroll
            for (uint32_t c = 0; c < VERIFY_MAX_ROWS; c++) if (alive & (1u << c)) okpos = (P)t0[c];
            sp = okpos;
            ep = (P)(okpos + ok);
            resolved = ok == 1;
            return;
        }
        may_verify = false;   // a repeat of the whole pattern: the index decides the order of its rows
    }
}

// One pattern per thread, grid-stride.  Work item w handles pattern idx[w] (idx == NULL: identity).  Outputs in WORK
// order, coalesced.  When the work order is the caller's order, a CTA first copies the bytes of its 256 patterns -- one
// contiguous range -- into shared memory with coalesced word loads (stage_bytes of dynamic shared memory; ranges that do
// not fit are read in place): a thread reading its own pattern byte by byte from global memory costs one L1 wavefront per
// lane and byte, which made 150 bp reads LSU-bound (21 ms per 10^7 reads) once text verification had removed the LF walks.
template <class P, int NPL, int VBITS, bool ILV>
__global__ void __launch_bounds__(SEARCH_THREADS)
search_kernel(const DevIndex<P> ix, const PatternBatch pb, const SearchIO<P> io, uint32_t stage_bytes) {
    extern __shared__ __align__(16) uint8_t s_stage[];
    __shared__ uint8_t s_table[256];
    __shared__ uint8_t s_rank[64];
    __shared__ uint8_t s_lut[256];   // byte -> rank among the occurring symbols | 0x80 never occurs | 0x40 PassThrough byte >= S
    __shared__ P s_count[65];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const uint32_t t = (ix.table && !pb.preencoded) ? ix.table[i] : (uint32_t)i;
        s_table[i] = (uint8_t)t;
        const uint32_t sidx = t >= ix.symbol_count ? ix.symbol_count - 1 : t;
        const uint32_t rk = ix.sym_rank[sidx];
        s_lut[i] = (uint8_t)((rk == 0xffu ? 0x80u : (rk & 0x3fu)) | (t >= ix.symbol_count ? 0x40u : 0u));
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_rank[i] = ix.sym_rank[i];
    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    __syncthreads();

    const uint32_t bits = io.bits;
    const uint32_t in_key = io.keys ? 64u / bits : 0u;  // symbols (from the end) available in the key
    const bool may_stage = stage_bytes >= 64 && !io.idx;
    int errbits = 0;
    unsigned long long rows = 0;
    for (uint64_t wb = (uint64_t)blockIdx.x * blockDim.x; wb < pb.n; wb += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t w = wb + threadIdx.x;
        // ---- stage the byte range of work items [wb, we) when it fits
        const uint64_t we = wb + blockDim.x < pb.n ? wb + blockDim.x : pb.n;
        uint64_t b0 = 0, b1 = 0;
        bool staged = false;
        uint32_t skew = 0;
        if (may_stage) {
            b0 = pb.offs ? pb.offs[wb] : wb * (uint64_t)pb.fixed_len;
            b1 = pb.offs ? pb.offs[we] : we * (uint64_t)pb.fixed_len;
            skew = (uint32_t)(reinterpret_cast<uintptr_t>(pb.pats + b0) & 3u);   // keep the word alignment of the source
            staged = b1 >= b0 && b1 - b0 + skew <= (uint64_t)stage_bytes;
        }
        if (pb.packed_bits) {   // fixed-length packed patterns: the stage holds them unpacked (host: the CTA's patterns always fit)
            b0 = wb * (uint64_t)pb.fixed_len;
            b1 = we * (uint64_t)pb.fixed_len;
            skew = 0;
            staged = true;
            SVFM_ASSERT(may_stage && b1 - b0 <= (uint64_t)stage_bytes);
            __syncthreads();
            const uint32_t nbytes = (uint32_t)(b1 - b0), len = pb.fixed_len, bits_p = pb.packed_bits, mask = (1u << bits_p) - 1u;
            const uint8_t* src = pb.pats + wb * (uint64_t)pb.packed_bpp;
            for (uint32_t o = threadIdx.x * 4u; o < nbytes; o += blockDim.x * 4u) {
                uint32_t i = o / len, j = o - i * len, word = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if (o + b < nbytes) {
                        const uint32_t bit = j * bits_p;
                        const uint8_t* q = src + (uint64_t)i * pb.packed_bpp + (bit >> 3);
                        uint32_t v = __ldg(q);
                        if ((bit & 7u) + bits_p > 8u) v |= (uint32_t)__ldg(q + 1) << 8;   // bits <= 8: a symbol spans at most two bytes
                        word |= ((v >> (bit & 7u)) & mask) << (8 * b);
                        if (++j == len) { j = 0; i++; }
                    }
                }
                *reinterpret_cast<uint32_t*>(s_stage + o) = word;   // the stage is 16-byte aligned and padded to a multiple of 4
            }
            __syncthreads();
        } else {
        if (may_stage) __syncthreads();   // the previous range is no longer read
        if (staged) {
            const uint8_t* src = pb.pats + b0;
            const uint32_t nbytes = (uint32_t)(b1 - b0);
            const uint32_t head = nbytes < ((4u - skew) & 3u) ? nbytes : ((4u - skew) & 3u);
            const uint32_t words = (nbytes - head) >> 2;
            const uint32_t tail0 = head + (words << 2);
            if (threadIdx.x < head) s_stage[skew + threadIdx.x] = src[threadIdx.x];
            const uint32_t* srcw = reinterpret_cast<const uint32_t*>(src + head);
            uint32_t* dstw = reinterpret_cast<uint32_t*>(s_stage + skew + head);
            for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) dstw[i] = __ldg(srcw + i);
            if (threadIdx.x < nbytes - tail0) s_stage[skew + tail0 + threadIdx.x] = src[tail0 + threadIdx.x];
        }
        if (may_stage) __syncthreads();
        }
        if (w >= pb.n) continue;
        const uint64_t i = io.idx ? (uint64_t)io.idx[w] : w;
        const uint64_t key = io.keys ? io.keys[w] : 0ull;
        P sp = 0, ep = 0;
        uint64_t base, len;
        if (pb.offs) { base = pb.offs[i]; len = pb.offs[i + 1] - base; }
        else { base = i * (uint64_t)pb.fixed_len; len = pb.fixed_len; }
        bool resolved = false;
        const uint8_t* p = staged ? s_stage + skew + (base - b0) : pb.pats + base;
        if (len == 0) errbits |= ERRBIT_EMPTY_PATTERN;
        else search_pattern<P, NPL, VBITS, ILV>(ix, p, len, pb.reversed != 0, key, in_key, bits, s_table, s_rank, s_lut, s_count, errbits, sp, ep, resolved);
        const P cnt = (P)(ep - sp);
        if (io.resolved_out) io.resolved_out[w] = resolved ? 1 : 0;
        if (io.sp_out) io.sp_out[w] = sp;
        if (io.cnt_out) io.cnt_out[w] = cnt;
        if (io.heavy_seen && (uint64_t)cnt > HEAVY_ROWS) atomicAdd(io.heavy_seen, 1ull);  // rare: sizes the heavy list
        if (io.sb.hist && cnt != 0) {
            atomicAdd(io.sb.hist + (i >> SB_SHIFT), (uint32_t)cnt);
            rows += (unsigned long long)cnt;
        }
    }
    if (io.sb.hist) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, d);
        if ((threadIdx.x & 31) == 0 && rows) atomicAdd(io.sb.total, rows);
    }
    if (errbits) atomicOr(io.err, errbits);
}

// ---- interleaved occ copy (built once per index at load) ----------------------------------
// entry q of `dst` (stride bytes) = block q (block_bytes) at offset 0, checkpoint row q (row_bytes) at offset ck_off.
// Pure byte movement in 4-byte words; every size involved is a multiple of 4.
static __global__ void ilv_build_kernel(const uint32_t* __restrict__ blocks, uint32_t block_words, const uint32_t* __restrict__ rows,
                                        uint32_t row_words, uint32_t* __restrict__ dst, uint32_t stride_words, uint32_t ck_off_words,
                                        uint64_t entries) {
    const uint64_t total = entries * stride_words;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t q = t / stride_words;
        const uint32_t w = (uint32_t)(t - q * stride_words);
        uint32_t v = 0;
        if (w < block_words) v = blocks[q * block_words + w];
        else if (w >= ck_off_words && w < ck_off_words + row_words) v = rows[q * row_words + (w - ck_off_words)];
        dst[t] = v;
    }
}

// ---- packed text copy (built once per index at load; "text verification" above) -----------------------------------------
// Row r of the suffix array: its suffix starts at text position locate_row(r), and the BWT symbol of the row is the text
// symbol right before it.  One thread per row; the symbol (as its rank among the occurring symbols) is OR-ed into the
// zeroed packed array.
// Expanded suffix array: fsa[row] = text position of SA row `row`, for every row (ix.fsa must be NULL here: the walk uses
// the blob's sampled array).
template <class P, int NPL, int VBITS, bool ILV, class E>
__global__ void __launch_bounds__(256)
fsa_build_kernel(const DevIndex<P> ix, uint64_t n, E* __restrict__ fsa) {
    __shared__ P s_count[65];
    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    __syncthreads();
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x)
        fsa[r] = (E)locate_row<P, NPL, VBITS, ILV>(ix, s_count, (P)r);
}

template <class P, int NPL, int VBITS, bool ILV>
__global__ void __launch_bounds__(256)
text_build_kernel(const DevIndex<P> ix, uint64_t n, uint32_t bits, uint32_t* __restrict__ text) {
    __shared__ P s_count[65];
    __shared__ uint8_t s_rank[64];
    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_rank[i] = ix.sym_rank[i];
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // the last text symbol precedes the sentinel: it is the BWT symbol of the `$` row, which the sentinel-free row
        // numbering skips -- stored BWT index 0 (bwm/mod.rs:107-108, crate_bio_manual/mod.rs:14-24)
        Block<NPL, VBITS> b;
        if constexpr (ILV) b.load_aligned(ix.ilv);
        else b.load(ix.blocks, (uint64_t)0);
        const uint64_t bit = (n - 1) * bits;
        atomicOr(text + (bit >> 5), (uint32_t)s_rank[b.symidx_of(0)] << (uint32_t)(bit & 31u));
    }
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        if ((P)r == (P)(ix.sentinel_index - 1)) continue;   // the row whose BWT symbol is the sentinel: nothing precedes text[0]
        uint64_t q;
        uint32_t rem;
        rank_addr<P, VBITS>(ix, (P)r, q, rem);
        Block<NPL, VBITS> b;
        if constexpr (ILV) b.load_aligned(ix.ilv + q * ix.ilv_stride);
        else b.load(ix.blocks, q);
        const uint32_t sym = b.symidx_of(rem);
        const P pos = locate_row<P, NPL, VBITS, ILV>(ix, s_count, (P)r);
        if (pos == 0) continue;
        const uint64_t bit = ((uint64_t)pos - 1) * bits;
        atomicOr(text + (bit >> 5), (uint32_t)s_rank[sym] << (uint32_t)(bit & 31u));
    }
}

// ---- extended k-mer table (built once per index at load) --------------------------------------------------------
// ext[e] = (sp, count) of the SA interval of the e-th string of length m over the symbols that occur in the text,
// in lexicographic order (e = base-s_eff number, first symbol most significant).  Level 1 comes from the count
// array, level j+1 from level j by one backward step per entry (locate/mod.rs:39-45): entry c*n_j + i of level j+1
// is the interval of (c followed by string i).  Entries of one level are visited in SA order, so every level is a
// sequential sweep over the checkpoint/block arrays.
template <class P>
__global__ void ext_level1_kernel(const DevIndex<P> ix, P* __restrict__ out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ix.s_eff) return;
    const uint32_t sym = ix.present[c];
    const P lo = ix.count_array[sym], hi = ix.count_array[sym + 1];
    out[2 * c] = lo;
    out[2 * c + 1] = (P)(hi - lo);
}

template <class P, int NPL, int VBITS>
__global__ void __launch_bounds__(SEARCH_THREADS)
ext_expand_kernel(const DevIndex<P> ix, const P* __restrict__ in, uint64_t n_in, P* __restrict__ out) {
    __shared__ P s_count[65];
    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    __syncthreads();
    const uint64_t n_out = n_in * ix.s_eff;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_out; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t c = t / n_in, i = t - c * n_in;
        P sp = in[2 * i];
        P ep = (P)(sp + in[2 * i + 1]);
        if (sp < ep) { const uint32_t sy = ix.present[c]; backward_step<P, NPL, VBITS>(ix, sy, s_count[sy], sp, ep); }
        out[2 * t] = sp;
        out[2 * t + 1] = (P)(ep - sp);
    }
}

// ---- sweep search of a dense fixed-length batch ------------------------------------------------------------------
// The whole batch moves through the index in SA order.  Every pattern becomes an item
//   prefix : index of its trailing m symbols in the extended table (0xffffffff: contains a symbol that never occurs)
//   rest   : the other len-m symbols as ranks among the occurring symbols, `bits` = ceil(log2 s_eff) each, the one
//            consumed first in the lowest bits
//   idx    : the caller's pattern index.
// Items are radix-sorted by prefix, i.e. by the SA interval of their m-symbol suffix; one streaming pass over the
// table seeds (sp, count).  Then rounds of T backward steps (with_slice.rs:27-31) + a stable radix partition by the
// T symbols just consumed (last one most significant): LF-mapping keeps the relative order of rows inside a symbol
// class, so after the partition the batch is again sorted by sp, and the next round reads checkpoints and blocks
// as a handful of monotone streams -- every index line comes from DRAM at most once per step, and DRAM sees
// streams instead of random sectors (measured on B200: 55 G random sectors/s = 1.76 TB/s vs 6.5 TB/s streaming).
template <class R>
struct SweepPay { R rest; uint32_t idx; };

// ---- TMA bulk copies (cp.async.bulk, global -> shared, completion on an mbarrier) --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}

constexpr uint32_t RADIX_BINS = 256;   // radix_pass_kernel: 8-bit digits
constexpr int PACK_TILE = 256;    // patterns per tile (= threads per CTA)
constexpr int PACK_STAGES = 4;    // tiles in flight per CTA

// Encode + pack one fixed-length batch into sweep items.  The pattern bytes stream through shared memory in tiles of
// PACK_TILE patterns, PACK_STAGES deep: one thread issues a TMA bulk copy per tile (cp.async.bulk, completion on an
// mbarrier), every thread then packs its own pattern out of the staged tile.  Unaligned batches and the ragged last
// tile take plain loads.
template <class R>
__global__ void __launch_bounds__(PACK_TILE)
pack_sweep_kernel(const uint8_t* __restrict__ table, const DevSyms syms, const PatternBatch pb, uint32_t bits, uint32_t m,
                  uint32_t* __restrict__ prefix, SweepPay<R>* __restrict__ pay, uint32_t digit_bits, uint32_t n_rounds,
                  uint32_t presort_shift, uint32_t presort_passes, uint32_t* __restrict__ hist, int* __restrict__ err) {
    extern __shared__ __align__(128) uint8_t s_tiles[];  // PACK_STAGES x tile_bytes | round histograms
    __shared__ __align__(8) uint64_t s_bar[PACK_STAGES];
    __shared__ uint16_t s_lut[256];                       // byte -> symbol index | rank << 8
    const uint32_t len = pb.fixed_len;
    const uint32_t S = syms.symbol_count;
    const uint32_t tile_bytes = PACK_TILE * len;
    const uint32_t stage_stride = (tile_bytes + 127u) & ~127u;
    // digit histograms of the partition rounds (sweep_round_kernel): round r sorts on bits [r*digit_bits, +digit_bits) of rest;
    // behind them the 256-bin histograms of the presort passes (radix_pass_kernel): pass q sorts on bits
    // [presort_shift + 8q, +8) of the table index
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_tiles + (size_t)PACK_STAGES * stage_stride);
    const uint32_t nb = 1u << digit_bits;
    const uint32_t n_hist = n_rounds * nb + presort_passes * RADIX_BINS;
    uint32_t* s_phist = s_hist + n_rounds * nb;
    for (uint32_t i = threadIdx.x; i < n_hist; i += blockDim.x) s_hist[i] = 0;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t sidx = (table && !pb.preencoded) ? table[i] : (uint32_t)i;
        const uint32_t bad = sidx >= S ? 0x8000u : 0u;     // PassThrough byte >= S
        if (sidx >= S) sidx = S - 1;
        s_lut[i] = (uint16_t)(sidx | ((uint32_t)(syms.sym_rank[sidx] & 0x7fu) << 8) | (syms.sym_rank[sidx] == 0xffu ? 0x4000u : 0u) | bad);
    }
    if (threadIdx.x == 0) {
        for (int st = 0; st < PACK_STAGES; st++) mbar_init(&s_bar[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t n_tiles = (pb.n + PACK_TILE - 1) / PACK_TILE;
    const uint64_t full_tiles = pb.n / PACK_TILE;
    const bool tma_ok = (reinterpret_cast<uintptr_t>(pb.pats) & 15u) == 0;  // tile_bytes = 256 * len is a multiple of 16
    int errbits = 0;

    // The kernel is issue-bound (ncu, round 1: 86 % of the issue slots busy), so the common case -- forward patterns whose
    // length is a multiple of four -- takes a lean loop: one word load per four symbols, one table lookup and one
    // multiply-add per symbol (the stored bytes run from the symbol consumed last to the one consumed first, so both the
    // remaining symbols and the table index accumulate Horner-style).
    auto pack_one = [&](const uint8_t* p, uint64_t i) {
        uint32_t e = 0, flags = 0;
        R rest = 0;
        if (!pb.reversed) {
            // stored bytes [0, n_rest) are the remaining symbols (Horner: the byte consumed first, right before the table's
            // m symbols, ends up lowest), [n_rest, len) the table index.  Two plain loops over the staged bytes: one byte load,
            // one table lookup and one shift-or / multiply-add per symbol (the kernel is issue-bound; the earlier single loop
            // with a per-symbol branch on the byte's role cost 20 instructions per symbol).
            const uint32_t n_rest = len - m;
            uint32_t f = 0;
#pragma unroll 4
            for (; f < n_rest; f++) {
                const uint32_t v = s_lut[p[f]];
                flags |= v;
                rest = (R)((rest << bits) | (R)((v >> 8) & 0x3fu));
            }
#pragma unroll 4
            for (; f < len; f++) {
                const uint32_t v = s_lut[p[f]];
                flags |= v;
                e = e * syms.s_eff + ((v >> 8) & 0x3fu);
            }
        } else {
            uint32_t mult = 1;
            for (uint32_t f = 0; f < len; f++) {  // f-th stored byte
                const uint32_t v = s_lut[p[f]];
                flags |= v;
                const uint32_t j = pb.reversed ? f : len - 1 - f;  // j-th symbol from the end of the pattern
                const uint32_t r = (v >> 8) & 0x3fu;
                if (j < m) {
                    if (pb.reversed) { e += r * mult; mult *= syms.s_eff; }  // j ascends: least significant digit first
                    else e = e * syms.s_eff + r;                             // j descends: Horner
                } else {
                    rest |= (R)r << (bits * (j - m));
                }
            }
        }
        if (flags & 0x8000u) errbits |= ERRBIT_BAD_SYMBOL;
        // 0x4000: a symbol that never occurs in the text -- count 0, wherever it sits in the pattern
        const uint32_t pfx = (flags & 0x4000u) ? 0xffffffffu : e;
        prefix[i] = pfx;
        for (uint32_t r = 0; r < n_rounds; r++) atomicAdd(&s_hist[r * nb + (uint32_t)((rest >> (r * digit_bits)) & (R)(nb - 1))], 1u);
        for (uint32_t q = 0; q < presort_passes; q++) atomicAdd(&s_phist[q * RADIX_BINS + ((pfx >> (presort_shift + 8u * q)) & (RADIX_BINS - 1u))], 1u);
        SweepPay<R> o;
        o.rest = rest;
        o.idx = (uint32_t)i;
        pay[i] = o;
    };

    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const uint64_t my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto tile_of = [&](uint64_t it) { return (uint64_t)blockIdx.x + it * gridDim.x; };
    auto issue = [&](uint64_t it) {  // thread 0 only
        const uint64_t t = tile_of(it);
        if (it < my_tiles && t < full_tiles && tma_ok) {
            uint64_t* bar = &s_bar[it % PACK_STAGES];
            mbar_expect_tx(bar, tile_bytes);
            bulk_g2s(s_tiles + (it % PACK_STAGES) * (uint64_t)stage_stride, pb.pats + t * (uint64_t)tile_bytes, tile_bytes, bar);
        }
    };
    if (threadIdx.x == 0)
        for (uint64_t it = 0; it + 1 < PACK_STAGES; it++) issue(it);
    for (uint64_t it = 0; it < my_tiles; it++) {
        const uint64_t t = tile_of(it);
        const uint64_t i0 = t * PACK_TILE;
        uint8_t* stage = s_tiles + (it % PACK_STAGES) * (uint64_t)stage_stride;
        // every thread is past iteration it-1 here (barrier at its end), so stage (it-1) % STAGES may be refilled
        if (threadIdx.x == 0) issue(it + PACK_STAGES - 1);
        if (t < full_tiles && tma_ok) {
            mbar_wait(&s_bar[it % PACK_STAGES], (uint32_t)((it / PACK_STAGES) & 1));
        } else {
            const uint64_t cnt = pb.n - i0 < PACK_TILE ? pb.n - i0 : PACK_TILE;
            const uint8_t* src = pb.pats + i0 * len;
            for (uint64_t o = threadIdx.x; o < cnt * len; o += PACK_TILE) stage[o] = __ldg(src + o);
            __syncthreads();
        }
        if (i0 + threadIdx.x < pb.n) pack_one(stage + (uint64_t)threadIdx.x * len, i0 + threadIdx.x);
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < n_hist; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
    if (errbits) atomicOr(err, errbits);
}

// Look-back descriptors are single words (flag | count) and carry no other data, so relaxed accesses are enough; GPU scope
// keeps the polls inside the device's L2 (ld.volatile compiles to a SYSTEM-scope strong load on sm_100: measured slower).
#ifndef SVFM_DESC_GPU
#define SVFM_DESC_GPU 1
#endif
__device__ __forceinline__ uint32_t ld_desc(const uint32_t* p) {
    uint32_t v;
#if SVFM_DESC_GPU
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#else
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#endif
    return v;
}
__device__ __forceinline__ void st_desc(uint32_t* p, uint32_t v) {
#if SVFM_DESC_GPU
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
constexpr uint32_t DESC_AGG = 1u << 30, DESC_PREFIX = 2u << 30, DESC_MASK = (1u << 30) - 1;

// Decoupled look-back for one digit: the sum of the digit's counts over the tiles [0, tile) = the aggregates of the
// predecessors back to the nearest one that has published its inclusive prefix.  LB_WINDOW descriptors are requested per
// trip before any of them is consumed, so a walk of k tiles costs ~k / LB_WINDOW L2 round trips instead of k.
#ifndef SVFM_LB_WINDOW
#define SVFM_LB_WINDOW 4
#endif
constexpr int LB_WINDOW = SVFM_LB_WINDOW;
__device__ __forceinline__ uint32_t lookback_excl(const uint32_t* desc, uint64_t tile, uint32_t nbins, uint32_t b) {
    uint32_t excl = 0;
    uint64_t t = tile;   // the descriptors of the tiles below t are still to be read
    while (t > 0) {
        uint32_t v[LB_WINDOW];
#pragma unroll
        for (int l = 0; l < LB_WINDOW; l++) v[l] = (uint64_t)l < t ? ld_desc(desc + (t - 1 - l) * nbins + b) : DESC_PREFIX;
        bool done = false;
#pragma unroll
        for (int l = 0; l < LB_WINDOW; l++) {
            if (!done) {
                while ((v[l] >> 30) == 0) v[l] = ld_desc(desc + (t - 1 - l) * nbins + b);
                excl += v[l] & DESC_MASK;
                done = (v[l] >> 30) == 2u;
            }
        }
        if (done) break;
        t = t > (uint64_t)LB_WINDOW ? t - LB_WINDOW : 0;
    }
    return excl;
}

// ---- one stable LSD radix pass over the packed items (table index -> payload), 8-bit digit ---------------------------------
// The sort of the sweep items by table index runs on this kernel (round 1 used cub::DeviceRadixSort): the same building
// blocks as the partition half of sweep_round_kernel -- tiles handed out by an atomic counter, stable rank inside the digit
// by warp match + per-warp counters, exclusive prefix over all earlier tiles by decoupled look-back (one descriptor word per
// tile and digit; one thread per digit walks back), shared-memory exchange so that every digit's run leaves with coalesced
// stores.  The batch histogram of the digit comes from pack_sweep_kernel.  Stable, so that two passes sort 16 bits.
constexpr int RADIX_THREADS = 256;
constexpr int RADIX_ITEMS = 8;
constexpr int RADIX_TILE = RADIX_THREADS * RADIX_ITEMS;
constexpr int RADIX_WARPS = RADIX_THREADS / 32;
template <class R>
__global__ void __launch_bounds__(RADIX_THREADS)
radix_pass_kernel(const uint32_t* __restrict__ key_in, const SweepPay<R>* __restrict__ pay_in, uint32_t* __restrict__ key_out,
                  SweepPay<R>* __restrict__ pay_out, uint64_t n, uint32_t shift, const uint32_t* __restrict__ hist,
                  uint32_t* desc, uint32_t* tile_counter) {
    using Pay = SweepPay<R>;
    extern __shared__ __align__(16) uint8_t s_dyn[];
    Pay* s_pay = reinterpret_cast<Pay*>(s_dyn);                                  // RADIX_TILE
    uint32_t* s_key = reinterpret_cast<uint32_t*>(s_pay + RADIX_TILE);          // RADIX_TILE
    uint32_t* s_whist = s_key + RADIX_TILE;                                       // RADIX_WARPS x RADIX_BINS
    __shared__ uint32_t s_binstart[RADIX_BINS], s_tilebin[RADIX_BINS], s_gbase[RADIX_BINS], s_scan[RADIX_WARPS], s_tile;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    {   // global start of every digit: exclusive scan of the batch histogram (one bin per thread)
        const uint32_t v = hist[threadIdx.x];
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(full, incl, d);
            if ((int)lane >= d) incl += o;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        uint32_t off = 0;
        for (uint32_t w = 0; w < warp; w++) off += s_scan[w];
        s_binstart[threadIdx.x] = off + incl - v;
    }
    __syncthreads();
    const uint64_t n_tiles = (n + RADIX_TILE - 1) / RADIX_TILE;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const uint64_t tile = s_tile;
        __syncthreads();
        if (tile >= n_tiles) break;
        const uint64_t base = tile * RADIX_TILE + (uint64_t)warp * (32 * RADIX_ITEMS) + lane;
        uint32_t key[RADIX_ITEMS], rank[RADIX_ITEMS];
        Pay pay[RADIX_ITEMS];
#pragma unroll
        for (int k = 0; k < RADIX_ITEMS; k++) {
            const uint64_t w = base + (uint64_t)k * 32;
            key[k] = 0;
            if (w < n) { key[k] = key_in[w]; pay[k] = pay_in[w]; }
        }
        for (uint32_t i = threadIdx.x; i < RADIX_WARPS * RADIX_BINS; i += RADIX_THREADS) s_whist[i] = 0;
        __syncthreads();
        uint32_t* my_hist = s_whist + warp * RADIX_BINS;
#pragma unroll
        for (int k = 0; k < RADIX_ITEMS; k++) {
            const bool valid = base + (uint64_t)k * 32 < n;
            const uint32_t dig = valid ? ((key[k] >> shift) & (RADIX_BINS - 1u)) : 0xffffffffu;
            const unsigned peers = __match_any_sync(full, dig);
            const int leader = __ffs(peers) - 1;
            uint32_t before = 0;
            if ((int)lane == leader && valid) {
                before = my_hist[dig];
                my_hist[dig] = before + __popc(peers);
            }
            before = __shfl_sync(full, before, leader);
            rank[k] = before + __popc(peers & ((1u << lane) - 1u));
            __syncwarp();
        }
        __syncthreads();
        {   // one digit per thread: offsets of the warps inside the tile, tile total, look-back, position of the digit's run
            const uint32_t b = threadIdx.x;
            uint32_t acc = 0;
#pragma unroll
            for (int w = 0; w < RADIX_WARPS; w++) {
                const uint32_t t = s_whist[w * RADIX_BINS + b];
                s_whist[w * RADIX_BINS + b] = acc;
                acc += t;
            }
            st_desc(desc + tile * RADIX_BINS + b, (tile > 0 ? DESC_AGG : DESC_PREFIX) | acc);
            uint32_t excl = 0;
            if (tile > 0) {
                excl = lookback_excl(desc, tile, RADIX_BINS, b);
                st_desc(desc + tile * RADIX_BINS + b, DESC_PREFIX | (excl + acc));
            }
            // exclusive scan of the tile totals over the digits: where the digit's run starts inside the exchange buffer
            uint32_t incl = acc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(full, incl, d);
                if ((int)lane >= d) incl += o;
            }
            if (lane == 31) s_scan[warp] = incl;
            __syncthreads();
            uint32_t run = incl - acc;
            for (uint32_t w = 0; w < warp; w++) run += s_scan[w];
            s_tilebin[b] = run;
            s_gbase[b] = s_binstart[b] + excl - run;   // global position = gbase[digit] + slot in the buffer (mod 2^32: n < 2^30)
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RADIX_ITEMS; k++) {
            if (base + (uint64_t)k * 32 < n) {
                const uint32_t dig = (key[k] >> shift) & (RADIX_BINS - 1u);
                const uint32_t slot = s_tilebin[dig] + my_hist[dig] + rank[k];
                SVFM_ASSERT(slot < (uint32_t)RADIX_TILE);
                s_key[slot] = key[k];
                s_pay[slot] = pay[k];
            }
        }
        __syncthreads();
        const uint32_t tile_n = (uint32_t)(n - tile * RADIX_TILE < (uint64_t)RADIX_TILE ? n - tile * RADIX_TILE : (uint64_t)RADIX_TILE);
        for (uint32_t j = threadIdx.x; j < tile_n; j += RADIX_THREADS) {
            const uint32_t kk = s_key[j];
            const uint32_t g = s_gbase[(kk >> shift) & (RADIX_BINS - 1u)] + j;
            SVFM_ASSERT((uint64_t)g < n);
            key_out[g] = kk;
            pay_out[g] = s_pay[j];
        }
        __syncthreads();
    }
}

// State of one pattern between rounds.  16 bytes for u32 positions and <= 32 bits of remaining symbols (one
// LDG.128 / STG.128 per item), 24 or 32 bytes otherwise.
template <class P, class R>
struct alignas((2 * sizeof(P) + sizeof(R) + 4) % 16 == 0 ? 16 : 8) SweepItem {
    P sp;
    P cnt;
    R rest;
    uint32_t idx;
};

#ifndef SVFM_ROUND_THREADS
#define SVFM_ROUND_THREADS 256
#endif
constexpr int ROUND_THREADS = SVFM_ROUND_THREADS;
#ifndef SVFM_ROUND_ITEMS
#define SVFM_ROUND_ITEMS 4
#endif
#ifndef SVFM_ROUND_MIN_CTAS
#define SVFM_ROUND_MIN_CTAS 5
#endif
#ifndef SVFM_ROUND_LBWARP
#define SVFM_ROUND_LBWARP 1
#endif
// ROUND_LBWARP: the last warp of the CTA carries no items; it computes the tile's digit totals, publishes them and walks the
// look-back WHILE the other warps run their backward steps (the look-back is a chain of dependent L2 round trips; with all
// warps waiting for it behind a barrier it cost a quarter of every tile's time: ncu, barrier stall 5.4 per issue slot).
constexpr bool ROUND_LBWARP = SVFM_ROUND_LBWARP != 0;
constexpr int ROUND_ITEMS = SVFM_ROUND_ITEMS;               // items per thread
constexpr int ROUND_WARPS = ROUND_THREADS / 32 - (ROUND_LBWARP ? 1 : 0);   // warps that carry items
constexpr int ROUND_TILE = ROUND_WARPS * 32 * ROUND_ITEMS;  // items per tile
constexpr int ROUND_MAX_BINS = 512;
enum : int { PART_NONE = 0, PART_SYMBOLS = 1, PART_INDEX = 2 };

template <class P, class R>
struct SweepRoundIO {
    // FIRST round: items as radix-sorted by prefix
    const uint32_t* prefix;
    const SweepPay<R>* pay;
    // later rounds: items as partitioned by the previous round
    const SweepItem<P, R>* items_in;
    // outputs (each nullable), written at the partitioned position (PART) or in place
    SweepItem<P, R>* items_out;
    P* sp_out;
    P* cnt_out;
    uint32_t* idx_out;
    // PART_SYMBOLS only
    const uint32_t* hist;      // digit histogram of this round over the whole batch (pack_sweep_kernel), nbins entries
    uint32_t* desc;            // tiles x nbins look-back descriptors, zeroed before the launch
    // PART_INDEX only: partition digit = idx >> idx_shift, bin b starts at min(n, b << idx_shift)
    uint32_t idx_shift;
    // PART_INDEX, and PART_SYMBOLS in arrival order (see below): nbins counters, zeroed before the launch
    uint32_t* bin_cursor;
    uint32_t* tile_counter;    // zeroed before the launch
    unsigned long long* heavy_seen;
    SbOut sb;                  // last round of `locate`: bucket sizes of the bucketed sort-back
};


// One round of the sweep search, fused with the radix partition that follows it:
//   load a tile of items (FIRST: seed them from the extended table) -> `steps` backward steps each
//   (with_slice.rs:27-31) -> partition digit -> stable rank of every item inside its digit (warp match + per-warp
//   counters) -> position of the tile's items inside every digit's run -> items leave through a shared-memory exchange
//   so that every digit's run is written with coalesced stores.
// PART_SYMBOLS: digit = the symbols just consumed, last one most significant; exclusive prefix of the digit counts over all
//   earlier tiles by decoupled look-back (one descriptor word per tile and digit: flag | count).  The digit depends on the
//   pattern alone, so a tile publishes its counts BEFORE its backward steps, and (ROUND_LBWARP) a warp without items
//   walks the look-back while the other warps run theirs; the walk requests LB_WINDOW descriptors per trip (lookback_excl).
//   History (B200, 10^8 20-mers, two rounds): every thread waiting behind a barrier for one thread per digit walking back
//   one descriptor at a time 5.31 ms; a block-wide window of 8 predecessors with two extra barriers per window 5.83 ms.
//   With io.bin_cursor set, PART_SYMBOLS skips the look-back: a tile reserves its space inside every digit's run with one
//   atomic add per digit, so the tiles of a run follow each other in ARRIVAL order instead of tile order.  Results never
//   depend on the order of the work items; the runs stay sorted up to the few hundred tiles in flight at any time (each
//   tile's share is sorted in itself), which is all the locality of the next round needs.
// PART_INDEX (last round of `count`): digit = top bits of the caller's pattern index; order inside a digit does not
//   matter (scatter_counts_kernel places by index), so a tile reserves its space with one atomic add per digit.
// Tiles are handed out by an atomic counter, so a tile's predecessors are always running or done.
// PART_NONE: items are written back in place.
template <class P, int NPL, int VBITS, class R, bool FIRST, int PART, bool SWP = false>
__global__ void __launch_bounds__(ROUND_THREADS, SVFM_ROUND_MIN_CTAS)
sweep_round_kernel(const DevIndex<P> ix, uint64_t n, uint32_t bits, uint32_t shift, uint32_t steps, uint32_t nbins,
                   const SweepRoundIO<P, R> io) {
    using Item = SweepItem<P, R>;
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ P s_count[65];
    __shared__ uint8_t s_present[64];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_scan[ROUND_THREADS / 32];
    __shared__ uint64_t s_wsum[ROUND_THREADS / 32];
    // dynamic: items[ROUND_TILE] | gbase[nbins] (u64) | binstart[nbins] (u64) | whist[ROUND_WARPS][nbins] | tilebin[nbins]
    Item* s_items = reinterpret_cast<Item*>(s_dyn);
    uint64_t* s_gbase = reinterpret_cast<uint64_t*>(s_dyn + sizeof(Item) * ROUND_TILE);
    uint64_t* s_binstart = s_gbase + nbins;
    uint32_t* s_whist = reinterpret_cast<uint32_t*>(s_binstart + nbins);
    uint32_t* s_tilebin = s_whist + ROUND_WARPS * nbins;
    unsigned long long rows = 0;

    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_present[i] = ix.present[i];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    if constexpr (PART == PART_SYMBOLS) {
        // global start of every digit = exclusive scan of the batch histogram (nbins <= 512: two values per thread)
        uint32_t v[ROUND_MAX_BINS / ROUND_THREADS];
        uint64_t sum = 0;
#pragma unroll
        for (int i = 0; i < ROUND_MAX_BINS / ROUND_THREADS; i++) {
            const uint32_t b = threadIdx.x * (ROUND_MAX_BINS / ROUND_THREADS) + i;
            v[i] = b < nbins ? io.hist[b] : 0u;
            sum += v[i];
        }
        uint64_t incl = sum;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const uint64_t o = __shfl_up_sync(full, incl, dlt);
            if ((int)lane >= dlt) incl += o;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint64_t run = incl - sum;
        for (uint32_t w = 0; w < warp; w++) run += s_wsum[w];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ROUND_MAX_BINS / ROUND_THREADS; i++) {
            const uint32_t b = threadIdx.x * (ROUND_MAX_BINS / ROUND_THREADS) + i;
            if (b < nbins) s_binstart[b] = run;
            run += v[i];
        }
    } else if constexpr (PART == PART_INDEX) {
        for (uint32_t b = threadIdx.x; b < nbins; b += ROUND_THREADS) {
            const uint64_t st = (uint64_t)b << io.idx_shift;
            s_binstart[b] = st < n ? st : n;
        }
    }
    __syncthreads();
    const R sym_mask = (R)((1ull << bits) - 1);
    const R digit_mask = (R)(nbins - 1);
    const uint64_t n_tiles = (n + ROUND_TILE - 1) / ROUND_TILE;

    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(io.tile_counter, 1u);
        __syncthreads();
        const uint64_t tile = s_tile;
        __syncthreads();  // everyone has read s_tile before thread 0 can overwrite it
        if (tile >= n_tiles) break;
        const bool worker = warp < (uint32_t)ROUND_WARPS;
        const uint64_t base = worker ? tile * ROUND_TILE + (uint64_t)warp * (32 * ROUND_ITEMS) + lane : n;   // look-back warp: no items
        P sp[ROUND_ITEMS], cnt[ROUND_ITEMS];
        R rest[ROUND_ITEMS];
        uint32_t idx[ROUND_ITEMS];
        // ---- load (warp-striped: row k of a warp is 32 consecutive items).  FIRST: all table indices first, then all
        // table entries (one 8/16-byte load each), so that a thread has its four dependent lookups in flight together.
        if constexpr (FIRST) {
            uint32_t e[ROUND_ITEMS];
#pragma unroll
            for (int k = 0; k < ROUND_ITEMS; k++) {
                const uint64_t w = base + (uint64_t)k * 32;
                e[k] = 0xffffffffu; rest[k] = 0; idx[k] = 0;
                if (w < n) {
                    e[k] = io.prefix[w];
                    const SweepPay<R> v = io.pay[w];
                    rest[k] = v.rest;
                    idx[k] = v.idx;
                }
            }
#pragma unroll
            for (int k = 0; k < ROUND_ITEMS; k++) {
                sp[k] = 0; cnt[k] = 0;
                if (e[k] != 0xffffffffu) {
                    if constexpr (sizeof(P) == 4) {
                        const uint2 q = __ldg(reinterpret_cast<const uint2*>(ix.ext) + e[k]);
                        sp[k] = q.x; cnt[k] = q.y;
                    } else {
                        const ulonglong2 q = __ldg(reinterpret_cast<const ulonglong2*>(ix.ext) + e[k]);
                        sp[k] = (P)q.x; cnt[k] = (P)q.y;
                    }
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < ROUND_ITEMS; k++) {
                const uint64_t w = base + (uint64_t)k * 32;
                sp[k] = 0; cnt[k] = 0; rest[k] = 0; idx[k] = 0;
                if (w < n) {
                    const Item v = io.items_in[w];
                    sp[k] = v.sp; cnt[k] = v.cnt; rest[k] = v.rest; idx[k] = v.idx;
                }
            }
        }
        // ---- stable rank inside the tile.  The digit depends on the pattern alone, so the tile's digit counts are
        // published BEFORE the backward steps: by the time this tile looks back, its predecessors have long done so too.
        uint32_t rank[ROUND_ITEMS], dig[ROUND_ITEMS];
        uint32_t* my_hist = s_whist + warp * nbins;
        if constexpr (PART != PART_NONE) {
            for (uint32_t i = threadIdx.x; i < ROUND_WARPS * nbins; i += ROUND_THREADS) s_whist[i] = 0;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < ROUND_ITEMS; k++) {
                const bool valid = base + (uint64_t)k * 32 < n;
                if constexpr (PART == PART_INDEX) dig[k] = valid ? (idx[k] >> io.idx_shift) : 0xffffffffu;
                else dig[k] = valid ? (uint32_t)((rest[k] >> shift) & digit_mask) : 0xffffffffu;
                const unsigned peers = __match_any_sync(full, dig[k]);
                const int leader = __ffs(peers) - 1;
                uint32_t before = 0;
                if ((int)lane == leader && valid) {
                    before = my_hist[dig[k]];
                    my_hist[dig[k]] = before + __popc(peers);
                }
                before = __shfl_sync(full, before, leader);
                rank[k] = before + __popc(peers & ((1u << lane) - 1u));
                __syncwarp();
            }
            __syncthreads();
        }
        // per digit: offsets of the warps inside the tile, tile total -> descriptor (aggregate) or atomic reservation
        auto aggregate = [&](uint32_t b0, uint32_t stride) {
            for (uint32_t b = b0; b < nbins; b += stride) {
                uint32_t acc = 0;
#pragma unroll
                for (int w = 0; w < ROUND_WARPS; w++) {
                    const uint32_t t = s_whist[w * nbins + b];
                    s_whist[w * nbins + b] = acc;
                    acc += t;
                }
                s_tilebin[b] = acc;
                if (PART == PART_SYMBOLS && !io.bin_cursor)
                    st_desc(io.desc + tile * nbins + b, (tile > 0 ? DESC_AGG : DESC_PREFIX) | acc);
                else
                    s_gbase[b] = s_binstart[b] + (acc ? atomicAdd(io.bin_cursor + b, acc) : 0u);
            }
        };
        // look-back over the earlier tiles (decoupled: aggregate or inclusive prefix per tile and digit), one thread per digit
        auto lookback = [&](uint32_t b0, uint32_t stride) {
            for (uint32_t b = b0; b < nbins; b += stride) {
                uint32_t excl = 0;
                if (tile > 0) {
                    excl = lookback_excl(io.desc, tile, nbins, b);
                    SVFM_ASSERT((uint64_t)excl <= n);
                    st_desc(io.desc + tile * nbins + b, DESC_PREFIX | (excl + s_tilebin[b]));
                }
                s_gbase[b] = s_binstart[b] + excl;
            }
        };
        const bool lb_needed = PART == PART_SYMBOLS && !io.bin_cursor;
        if constexpr (PART != PART_NONE && !ROUND_LBWARP) aggregate(threadIdx.x, ROUND_THREADS);
        if (PART != PART_NONE && ROUND_LBWARP && !worker) {
            aggregate(lane, 32);
            if (lb_needed) lookback(lane, 32);
        } else {
        // ---- backward steps
#pragma unroll
        for (int k = 0; k < ROUND_ITEMS; k++) {
            if (cnt[k] != 0) {
                P ep = (P)(sp[k] + cnt[k]);
                R left = (R)(rest[k] >> shift);
                for (uint32_t t = 0; t < steps && sp[k] < ep; t++, left >>= bits) {
                    const uint32_t code = (uint32_t)(left & sym_mask);
                    const uint32_t sy = s_present[code];
                    if constexpr (SWP) backward_step_swp<P, NPL>(ix, code, sy, s_count[sy], sp[k], ep);
                    else backward_step<P, NPL, VBITS>(ix, sy, s_count[sy], sp[k], ep);
                }
                cnt[k] = (P)(ep - sp[k]);
            }
            if (io.heavy_seen && (uint64_t)cnt[k] > HEAVY_ROWS) atomicAdd(io.heavy_seen, 1ull);
            if (io.sb.hist && cnt[k] != 0) {
                atomicAdd(io.sb.hist + (idx[k] >> SB_SHIFT), (uint32_t)cnt[k]);
                rows += (unsigned long long)cnt[k];
            }
        }
        }
        if constexpr (PART == PART_NONE) {
#pragma unroll
            for (int k = 0; k < ROUND_ITEMS; k++) {
                const uint64_t w = base + (uint64_t)k * 32;
                if (w < n) {
                    if (io.items_out) { Item v; v.sp = sp[k]; v.cnt = cnt[k]; v.rest = rest[k]; v.idx = idx[k]; io.items_out[w] = v; }
                    if (io.sp_out) io.sp_out[w] = sp[k];
                    if (io.cnt_out) io.cnt_out[w] = cnt[k];
                    if (io.idx_out) io.idx_out[w] = idx[k];
                }
            }
        } else {
        if constexpr (!ROUND_LBWARP) { if (lb_needed) lookback(threadIdx.x, ROUND_THREADS); }
        __syncthreads();
        // exclusive scan of the tile totals over the digits (position of every digit's run inside the exchange buffer)
        {
            uint32_t v[ROUND_MAX_BINS / ROUND_THREADS], sum = 0;
#pragma unroll
            for (int i = 0; i < ROUND_MAX_BINS / ROUND_THREADS; i++) {
                const uint32_t b = threadIdx.x * (ROUND_MAX_BINS / ROUND_THREADS) + i;
                v[i] = b < nbins ? s_tilebin[b] : 0u;
                sum += v[i];
            }
            uint32_t incl = sum;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
                const uint32_t o = __shfl_up_sync(full, incl, dlt);
                if ((int)lane >= dlt) incl += o;
            }
            if (lane == 31) s_scan[warp] = incl;
            __syncthreads();
            uint32_t warp_off = 0;
            for (uint32_t w = 0; w < warp; w++) warp_off += s_scan[w];
            uint32_t run = warp_off + incl - sum;
#pragma unroll
            for (int i = 0; i < ROUND_MAX_BINS / ROUND_THREADS; i++) {
                const uint32_t b = threadIdx.x * (ROUND_MAX_BINS / ROUND_THREADS) + i;
                if (b < nbins) { s_tilebin[b] = run; s_gbase[b] -= run; }  // global position = gbase[digit] + slot in the buffer
                run += v[i];
            }
        }
        __syncthreads();
        // ---- exchange through shared memory: slot = run start of the digit + warps before me + rank
#pragma unroll
        for (int k = 0; k < ROUND_ITEMS; k++) {
            if (dig[k] != 0xffffffffu) {
                const uint32_t slot = s_tilebin[dig[k]] + my_hist[dig[k]] + rank[k];
                SVFM_ASSERT(dig[k] < nbins && slot < (uint32_t)ROUND_TILE);
                Item v; v.sp = sp[k]; v.cnt = cnt[k]; v.rest = rest[k]; v.idx = idx[k];
                s_items[slot] = v;
            }
        }
        __syncthreads();
        const uint32_t tile_n = (uint32_t)(n - tile * ROUND_TILE < (uint64_t)ROUND_TILE ? n - tile * ROUND_TILE : (uint64_t)ROUND_TILE);
        for (uint32_t j = threadIdx.x; j < tile_n; j += ROUND_THREADS) {
            const Item v = s_items[j];
            uint32_t d;
            if constexpr (PART == PART_INDEX) d = v.idx >> io.idx_shift;
            else d = (uint32_t)((v.rest >> shift) & digit_mask);
            const uint64_t g = s_gbase[d] + j;
            SVFM_ASSERT(d < nbins && g < n);   // every digit's run stays inside the batch: histogram + look-back / reservation agree
            if (io.items_out) io.items_out[g] = v;
            if (io.sp_out) io.sp_out[g] = v.sp;
            if (io.cnt_out) io.cnt_out[g] = v.cnt;
            if (io.idx_out) io.idx_out[g] = v.idx;
        }
        __syncthreads();
        }  // PART
    }
    if (io.sb.hist) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) rows += __shfl_xor_sync(full, rows, d);
        if (lane == 0 && rows) atomicAdd(io.sb.total, rows);
    }
}

// ---- last round of the sweep search in `locate` mode: resolve instead of step (opt-in, SVFM_SWEEP_RESOLVE=1) ------------
// (Measured slower than the plain last round on SA-ordered batches -- see engine.cuh -- and therefore off by default.)
// After the first rounds almost every item of a typical batch is down to ONE SA row (20-mers on 1 Gbp: 94 % after 17
// symbols).  Locating that row costs the same LF walk + SA read that `locate` would pay after the last round anyway, and
// with the row's text position in hand the symbols that are still unread can be checked against the packed text copy (one
// more sector) instead of being walked through the index and partitioned once more: such an item is RESOLVED here (count 0
// or 1, sp = the text position; see "text verification").  Items with more than one row take their remaining backward steps
// on the spot (with_slice.rs:27-31) and are located row by row later, as before.  The items arrive sorted by SA position
// (they were partitioned by the previous round), so the LF walks and SA reads of this kernel stream like those of
// locate_warp_kernel did; outputs stay in that order.
template <class P, int NPL, int VBITS, class R, bool ILV>
__global__ void __launch_bounds__(256)
sweep_resolve_kernel(const DevIndex<P> ix, const SweepItem<P, R>* __restrict__ items, uint64_t n, uint32_t bits, uint32_t shift,
                     uint32_t steps, P* __restrict__ sp_out, P* __restrict__ cnt_out, uint32_t* __restrict__ idx_out,
                     uint8_t* __restrict__ resolved_out, unsigned long long* heavy_seen, SbOut sb) {
    __shared__ P s_count[65];
    __shared__ uint8_t s_present[64];
    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_present[i] = ix.present[i];
    __syncthreads();
    const R sym_mask = (R)((1ull << bits) - 1);
    unsigned long long rows = 0;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += (uint64_t)gridDim.x * blockDim.x) {
        const SweepItem<P, R> it = items[w];
        P sp = it.sp, cnt = it.cnt;
        const R left = (R)(it.rest >> shift);   // rank of the t-th symbol still to be consumed: (left >> bits*t) & mask
        bool resolved = false;
        if (cnt == 1) {
            const P pos = locate_row<P, NPL, VBITS, ILV>(ix, s_count, sp);
            bool same = (uint64_t)pos >= steps;   // else the pattern would start before the text does
            for (uint32_t t = 0; t < steps && same; t++)
                same = text_symbol(ix.text, ix.text_bits, (uint64_t)pos - 1 - t) == (uint32_t)((left >> (bits * t)) & sym_mask);
            sp = same ? (P)(pos - (P)steps) : (P)0;
            cnt = same ? 1 : 0;
            resolved = same;
        } else if (cnt != 0) {
            P ep = (P)(sp + cnt);
            R l2 = left;
            for (uint32_t t = 0; t < steps && sp < ep; t++, l2 >>= bits) {
                const uint32_t sy = s_present[(uint32_t)(l2 & sym_mask)];
                backward_step<P, NPL, VBITS>(ix, sy, s_count[sy], sp, ep);
            }
            cnt = (P)(ep - sp);
        }
        sp_out[w] = sp;
        cnt_out[w] = cnt;
        idx_out[w] = it.idx;
        resolved_out[w] = resolved ? 1 : 0;
        if (heavy_seen && (uint64_t)cnt > HEAVY_ROWS) atomicAdd(heavy_seen, 1ull);
        if (sb.hist && cnt != 0) {
            atomicAdd(sb.hist + (it.idx >> SB_SHIFT), (uint32_t)cnt);
            rows += (unsigned long long)cnt;
        }
    }
    if (sb.hist) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, d);
        if ((threadIdx.x & 31) == 0 && rows) atomicAdd(sb.total, rows);
    }
}

// ---- small batches (a few thousand patterns at most; every single-pattern call) -------------------------------------
// ONE launch does search and locate: a thread takes one pattern, runs the backward search and walks up to `slots_per` of
// its SA rows into a fixed slot array; the host turns counts + slots into the CSR result (or, when some pattern has more
// rows than slots, runs the general pipeline).  Patterns, counts, slots and the status word live in PINNED HOST memory
// mapped into the device address space, so the call is one kernel launch and one stream synchronisation -- no copy calls,
// no memsets: what a latency-bound single `count` / `locate` needs (round 1: 43 / 78 us per call, six runtime calls each).
struct SmallOut {
    void* cnt;        // P[n]
    void* slots;      // P[n * slots_per], NULL for count
    uint32_t slots_per;
    uint8_t* status;  // u8[n]: error bits of every pattern
};
template <class P, int NPL, int VBITS, bool ILV>
__global__ void __launch_bounds__(128)
small_batch_kernel(const DevIndex<P> ix, const PatternBatch pb, const SmallOut out) {
    __shared__ uint8_t s_table[256];
    __shared__ uint8_t s_rank[64];
    __shared__ uint8_t s_lut[256];   // byte -> rank among the occurring symbols | 0x80 never occurs | 0x40 PassThrough byte >= S
    __shared__ P s_count[65];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const uint32_t t = (ix.table && !pb.preencoded) ? ix.table[i] : (uint32_t)i;
        s_table[i] = (uint8_t)t;
        const uint32_t sidx = t >= ix.symbol_count ? ix.symbol_count - 1 : t;
        const uint32_t rk = ix.sym_rank[sidx];
        s_lut[i] = (uint8_t)((rk == 0xffu ? 0x80u : (rk & 0x3fu)) | (t >= ix.symbol_count ? 0x40u : 0u));
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_rank[i] = ix.sym_rank[i];
    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pb.n) return;
    int errbits = 0;
    P sp = 0, ep = 0;
    uint64_t base, len;
    if (pb.offs) { base = pb.offs[i]; len = pb.offs[i + 1] - base; }
    else { base = i * (uint64_t)pb.fixed_len; len = pb.fixed_len; }
    bool resolved = false;
    if (len == 0) errbits |= ERRBIT_EMPTY_PATTERN;
    else search_pattern<P, NPL, VBITS, ILV>(ix, pb.pats + base, len, pb.reversed != 0, 0ull, 0u, 1u, s_table, s_rank, s_lut, s_count, errbits, sp, ep, resolved);
    const P cnt = (P)(ep - sp);
    reinterpret_cast<P*>(out.cnt)[i] = cnt;
    if (out.slots) {
        P* slot = reinterpret_cast<P*>(out.slots) + i * out.slots_per;
        const uint32_t rows = (uint64_t)cnt < out.slots_per ? (uint32_t)cnt : out.slots_per;
        if (resolved) slot[0] = sp;   // text verification: the position itself
        else for (uint32_t j = 0; j < rows; j++) slot[j] = locate_row<P, NPL, VBITS, ILV>(ix, s_count, (P)(sp + j));
    }
    out.status[i] = (uint8_t)errbits;
}

constexpr int LOCATE_THREADS = 256;
#ifndef SVFM_LOCATE_MIN_CTAS
#define SVFM_LOCATE_MIN_CTAS 1
#endif

template <class P>
struct HeavyList {
    P* sp;                    // first SA row of the pattern
    P* cnt;                   // number of rows
    uint64_t* obase;          // where its positions start in the output
    uint32_t* pat;            // its pattern index (record mode)
    unsigned long long* n;    // entries appended so far
    uint64_t capacity;
};

// ---- bucketed sort-back (locate results -> the caller's CSR order without a radix sort) ---------------------------
// The locate kernels emit (pattern index, position) records straight into BUCKETS of 2^SB_SHIFT consecutive pattern
// indices.  The bucket sizes are known before locate starts (SbOut: the search kernels sum the counts per bucket), so a
// pattern reserves its cnt consecutive slots with ONE atomic add on its bucket's cursor and its rows land there in
// SA-row order (locate/mod.rs:19).  A bucket ends up holding every record of 4096 consecutive patterns, each pattern's
// records contiguous; sb_place_kernel finishes inside shared memory.  Writes go to ~n/4096 write fronts that advance
// monotonically, which L2 merges into full lines; nothing is sorted.  (Reserving the slots in the search kernel instead
// -- no atomic in locate -- was measured: 6.7 ms instead of 4.6 ms, because a bucket's records then arrive in an order
// unrelated to their addresses and every 8-byte store becomes a partial DRAM sector write.)
template <class P> struct SbRec;
template <> struct alignas(8) SbRec<uint32_t> { uint32_t pos; uint32_t idx; };
template <> struct alignas(16) SbRec<uint64_t> { uint64_t pos; uint32_t idx; uint32_t pad; };

template <class P>
struct BucketOut {
    SbRec<P>* recs;                // total records, bucket b at [base[b], base[b+1])
    unsigned long long* cursor;    // per bucket: next free record slot (starts at base[b])
};

// Locate for patterns in work order: work item w has SA rows sp_work[w] .. +cnt_work[w].
// CSR mode (BUCKET = false): the position of row sp+j goes to positions[offs[w] + j] (offs = exclusive prefix sums of
// cnt_work), i.e. in SA-row order inside a pattern (the reference's order, locate/mod.rs:19); with rec_key != NULL also
// rec_key[offs[w] + j] = idx[w] (records for the radix sort-back, SVFM_SORTED only).
// Bucket mode (BUCKET = true): see above; offs / positions / rec_key are unused.
// One warp owns 32 consecutive work items and spreads ALL their rows over its lanes (warp prefix sum of the
// counts, then each lane finds the owner of its row with a shuffle binary search), so a pattern with many
// rows does not serialise one lane.  Patterns with more than HEAVY_ROWS rows are deferred to
// locate_rows_kernel through the heavy list.
template <class P, int NPL, int VBITS, bool ILV, bool BUCKET>
__global__ void __launch_bounds__(LOCATE_THREADS, SVFM_LOCATE_MIN_CTAS)
locate_warp_kernel(const DevIndex<P> ix, const uint32_t* __restrict__ idx, const P* __restrict__ sp_work,
                   const P* __restrict__ cnt_work, const uint64_t* __restrict__ offs, uint64_t n,
                   P* __restrict__ positions, uint32_t* __restrict__ rec_key, HeavyList<P> heavy, BucketOut<P> bk,
                   const uint8_t* __restrict__ resolved) {
    __shared__ P s_count[65];
    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    __syncthreads();
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint64_t warp_id = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t groups = (n + 31) / 32;
    auto emit = [&](uint64_t at, P pos, uint32_t pattern) {
        if constexpr (BUCKET) {
            SbRec<P> r;
            r.pos = pos;
            r.idx = pattern;
            if constexpr (sizeof(P) == 8) r.pad = 0;
            bk.recs[at] = r;
        } else {
            positions[at] = pos;
            if (rec_key) rec_key[at] = pattern;
        }
    };
    // staged group (two steps ahead of the lanes): raw loads issued (ST_RAW), then slots reserved (ST_READY)
    enum : int { ST_EMPTY = 0, ST_RAW = 1, ST_READY = 2 };
    int st = ST_EMPTY;
    uint64_t g_next = warp_id;
    uint32_t s_c = 0, s_pat = 0;
    P s_sp = 0, s_cw = 0;
    uint64_t s_obase = 0;
    bool s_res = false;
    auto stage_load = [&]() {   // issue the loads of group g_next (nothing is consumed here)
        st = ST_EMPTY;
        if (g_next < groups) {
            const uint64_t w = g_next * 32 + lane;
            s_cw = 0; s_sp = 0; s_pat = (uint32_t)w; s_res = false;
            if (w < n) {
                s_cw = cnt_work[w];
                s_sp = sp_work[w];
                if (idx) s_pat = idx[w];
                if (resolved) s_res = resolved[w] != 0;
            }
            g_next += n_warps;
            st = ST_RAW;
        }
    };
    auto stage_reserve = [&]() {   // consume the raw loads: reserve the output slots, defer heavy patterns
        s_c = 0;
        if (s_cw != 0) {
            if constexpr (BUCKET) s_obase = atomicAdd(bk.cursor + (s_pat >> SB_SHIFT), (unsigned long long)s_cw);
            else s_obase = offs[(g_next - n_warps) * 32 + lane];
            if ((uint64_t)s_cw > HEAVY_ROWS) {
                const unsigned long long h = atomicAdd(heavy.n, 1ull);
                if (h < heavy.capacity) { heavy.sp[h] = s_sp; heavy.cnt[h] = s_cw; heavy.obase[h] = s_obase; heavy.pat[h] = s_pat; }
            } else {
                s_c = (uint32_t)s_cw;
            }
        }
        st = ST_READY;
    };
    // current group: item `lane` of the group has rows [b_incl - b_c, b_incl) of the group's T rows
    uint32_t b_incl = 0, b_c = 0, b_pat = 0, T = 0, next = 0;
    P b_sp = 0;
    uint64_t b_obase = 0;
    bool b_res = false;
    // the row this lane is walking
    bool active = false, res = false;
    P pos = 0, offset = 0;
    uint64_t at = 0;
    uint32_t pat = 0;
    stage_load();
    for (;;) {
        if (st == ST_RAW) stage_reserve();
        // ---- hand rows to the lanes that have none
        unsigned fm = __ballot_sync(full, !active);
        while (fm != 0 && next >= T && st == ST_READY) {   // the current group is used up: install the staged one
            b_c = s_c; b_sp = s_sp; b_pat = s_pat; b_obase = s_obase; b_res = s_res;
            b_incl = b_c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(full, b_incl, d);
                if ((int)lane >= d) b_incl += v;
            }
            T = __shfl_sync(full, b_incl, 31);
            next = 0;
            stage_load();
            if (T == 0 && st == ST_RAW) stage_reserve();   // an empty group: keep going
        }
        if (fm != 0 && next < T) {
            const uint32_t r = next + (uint32_t)__popc(fm & lt_mask);
            const bool take = !active && r < T;
            const uint32_t rr = r < T ? r : T - 1;
            int o = 0;   // owner = first lane whose inclusive sum exceeds rr
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t v = __shfl_sync(full, b_incl, o + step - 1);
                if (v <= rr) o += step;
            }
            o &= 31;
            const uint32_t e = __shfl_sync(full, b_incl - b_c, o);
            const P osp = __shfl_sync(full, b_sp, o);
            const uint64_t oo = __shfl_sync(full, b_obase, o);
            const uint32_t opat = __shfl_sync(full, b_pat, o);
            const bool ores = __shfl_sync(full, (int)b_res, o) != 0;
            if (take) {
                const uint32_t j = rr - e;
                pos = ores ? osp : (P)(osp + (P)j);
                at = oo + j;
                pat = opat;
                res = ores;
                offset = 0;
                active = true;
            }
            const uint32_t nf = (uint32_t)__popc(fm);
            next += nf < T - next ? nf : T - next;
        }
        if (!__any_sync(full, active)) {
            if (st == ST_EMPTY && next >= T) break;
            continue;
        }
        // ---- one memory round trip per lane: the sampled SA entry, or one LF step (locate/mod.rs:21-33)
        // (the SA load is issued before the LF branch and consumed after it, so that both kinds of lanes wait together)
        {
            uint64_t quot = 0;
            const bool direct = ix.fsa != nullptr;   // expanded suffix array: every row is one read
            const bool sampled = active && !res && (direct || sa_sampled<P>(ix, pos, quot));
            const bool none = active && !res && !sampled && pos == (P)(ix.sentinel_index - 1);   // None arm, locate/mod.rs:27-30
            const bool walk = active && !res && !sampled && !none;
            P value = res ? pos : offset;   // text verification left the text position in sp; None arm: the offset alone
            if (sampled) value = direct ? fsa_get<P>(ix, pos) : ld_gather<P>(ix.suffix_array + quot);
            if (walk) {
                pos = lf_step<P, NPL, VBITS, ILV>(ix, s_count, pos);
                offset += 1;
            }
            if (sampled) value = (P)(value + offset);
            if (active && !walk) {
                emit(at, value, pat);
                active = false;
            }
        }
    }
}

// Locate with the expanded suffix array: every row is ONE read, so there is nothing to balance between lanes -- a thread
// takes LOCATE_DIRECT_ITEMS work items (consecutive threads take consecutive items: coalesced loads of sp / cnt / idx, and
// in SA order -- the sweep search's last partition -- neighbouring lanes read neighbouring array entries), reserves each
// item's output slots, and copies its rows' entries.  Items with more than 32 rows are spread over the warp, more than
// HEAVY_ROWS go to the heavy list as in locate_warp_kernel.  Same outputs as locate_warp_kernel.
#ifndef SVFM_LOCATE_DIRECT_ITEMS
#define SVFM_LOCATE_DIRECT_ITEMS 4
#endif
constexpr int LOCATE_DIRECT_ITEMS = SVFM_LOCATE_DIRECT_ITEMS;
template <class P, bool BUCKET>
__global__ void __launch_bounds__(LOCATE_THREADS)
locate_direct_kernel(const DevIndex<P> ix, const uint32_t* __restrict__ idx, const P* __restrict__ sp_work,
                     const P* __restrict__ cnt_work, const uint64_t* __restrict__ offs, uint64_t n,
                     P* __restrict__ positions, uint32_t* __restrict__ rec_key, HeavyList<P> heavy, BucketOut<P> bk,
                     const uint8_t* __restrict__ resolved) {
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    auto emit = [&](uint64_t at, P pos, uint32_t pattern) {
        if constexpr (BUCKET) {
            SbRec<P> r;
            r.pos = pos;
            r.idx = pattern;
            if constexpr (sizeof(P) == 8) r.pad = 0;
            bk.recs[at] = r;
        } else {
            positions[at] = pos;
            if (rec_key) rec_key[at] = pattern;
        }
    };
    const uint64_t per_cta = (uint64_t)LOCATE_THREADS * LOCATE_DIRECT_ITEMS;
    const uint64_t n_tiles = (n + per_cta - 1) / per_cta;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {   // whole tiles: every lane of a warp stays in the loop
        P sp[LOCATE_DIRECT_ITEMS], cw[LOCATE_DIRECT_ITEMS];
        uint32_t pat[LOCATE_DIRECT_ITEMS];
        bool res[LOCATE_DIRECT_ITEMS];
        uint64_t obase[LOCATE_DIRECT_ITEMS];
#pragma unroll
        for (int k = 0; k < LOCATE_DIRECT_ITEMS; k++) {
            const uint64_t w = tile * per_cta + (uint64_t)k * LOCATE_THREADS + threadIdx.x;
            cw[k] = 0; sp[k] = 0; pat[k] = (uint32_t)w; res[k] = false;
            if (w < n) {
                cw[k] = cnt_work[w];
                sp[k] = sp_work[w];
                if (idx) pat[k] = idx[w];
                if (resolved) res[k] = resolved[w] != 0;
            }
        }
#pragma unroll
        for (int k = 0; k < LOCATE_DIRECT_ITEMS; k++) {
            obase[k] = 0;
            if (cw[k] != 0) {
                if constexpr (BUCKET) obase[k] = atomicAdd(bk.cursor + (pat[k] >> SB_SHIFT), (unsigned long long)cw[k]);
                else obase[k] = offs[tile * per_cta + (uint64_t)k * LOCATE_THREADS + threadIdx.x];
                if ((uint64_t)cw[k] > HEAVY_ROWS) {
                    const unsigned long long h = atomicAdd(heavy.n, 1ull);
                    if (h < heavy.capacity) { heavy.sp[h] = sp[k]; heavy.cnt[h] = cw[k]; heavy.obase[h] = obase[k]; heavy.pat[h] = pat[k]; }
                    cw[k] = 0;
                }
            }
        }
        // the first row of every item: all reads in flight before the first store
        P first[LOCATE_DIRECT_ITEMS];
#pragma unroll
        for (int k = 0; k < LOCATE_DIRECT_ITEMS; k++) first[k] = (cw[k] != 0 && !res[k]) ? fsa_get<P>(ix, sp[k]) : sp[k];
#pragma unroll
        for (int k = 0; k < LOCATE_DIRECT_ITEMS; k++) {
            if (cw[k] != 0) emit(obase[k], first[k], pat[k]);
            // further rows: up to 32 by the owner, more by the whole warp
            const bool wide = (uint64_t)cw[k] > 32;
            if (cw[k] > 1 && !wide)
                for (uint32_t j = 1; j < (uint32_t)cw[k]; j++) emit(obase[k] + j, fsa_get<P>(ix, (P)(sp[k] + (P)j)), pat[k]);
            unsigned wm = __ballot_sync(full, wide);
            while (wm) {
                const int o = __ffs(wm) - 1;
                wm &= wm - 1;
                const uint32_t oc = (uint32_t)__shfl_sync(full, (uint32_t)cw[k], o);   // <= HEAVY_ROWS
                const P osp = __shfl_sync(full, sp[k], o);
                const uint64_t oo = __shfl_sync(full, obase[k], o);
                const uint32_t opat = __shfl_sync(full, pat[k], o);
                for (uint32_t j = 1 + lane; j < oc; j += 32) emit(oo + j, fsa_get<P>(ix, (P)(osp + (P)j)), opat);
            }
        }
    }
}

// Row-parallel locate for the heavy list: one thread per SA row; row t belongs to the entry h with
// offs[h] <= t < offs[h+1].  The per-block window of candidate entries is found once with two binary
// searches; each thread then searches only that window.
template <class P, int NPL, int VBITS, bool ILV, bool BUCKET>
__global__ void __launch_bounds__(LOCATE_THREADS)
locate_rows_kernel(const DevIndex<P> ix, const P* __restrict__ sp, const uint64_t* __restrict__ offs,
                   const uint64_t* __restrict__ obase, const uint32_t* __restrict__ pat, uint64_t n, uint64_t total,
                   P* __restrict__ positions, uint32_t* __restrict__ rec_key, SbRec<P>* __restrict__ recs) {
    __shared__ P s_count[65];
    __shared__ uint64_t s_win[2];
    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    const uint64_t t0 = (uint64_t)blockIdx.x * LOCATE_THREADS;
    if (threadIdx.x < 2) {
        // largest i in [0, n) with offs[i] <= target
        uint64_t target = threadIdx.x == 0 ? t0 : (t0 + LOCATE_THREADS - 1 < total ? t0 + LOCATE_THREADS - 1 : total - 1);
        uint64_t lo = 0, hi = n;  // invariant: offs[lo] <= target < offs[hi]
        while (hi - lo > 1) {
            uint64_t mid = lo + ((hi - lo) >> 1);
            if (__ldg(offs + mid) <= target) lo = mid; else hi = mid;
        }
        s_win[threadIdx.x] = lo;
    }
    __syncthreads();
    const uint64_t t = t0 + threadIdx.x;
    if (t >= total) return;
    uint64_t lo = s_win[0], hi = s_win[1] + 1;
    while (hi - lo > 1) {
        uint64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(offs + mid) <= t) lo = mid; else hi = mid;
    }
    const uint64_t j = t - __ldg(offs + lo);
    const P row = (P)(__ldg(sp + lo) + (P)j);
    const uint64_t dst = __ldg(obase + lo) + j;
    const P pos = locate_row<P, NPL, VBITS, ILV>(ix, s_count, row);
    if constexpr (BUCKET) {
        SbRec<P> r;
        r.pos = pos;
        r.idx = __ldg(pat + lo);
        if constexpr (sizeof(P) == 8) r.pad = 0;
        recs[dst] = r;
    } else {
        positions[dst] = pos;
        if (rec_key) rec_key[dst] = __ldg(pat + lo);
    }
}

// Bucket bases: exclusive prefix sums of the per-bucket record counts (one CTA; nb <= a few 10^5).
// base[0..nb], base[nb] = total records; cursor[b] = base[b].
constexpr int SB_SCAN_THREADS = 1024;
static __global__ void __launch_bounds__(SB_SCAN_THREADS)
sb_scan_kernel(const uint32_t* __restrict__ hist, uint64_t nb, uint64_t* __restrict__ base, unsigned long long* __restrict__ cursor) {
    __shared__ uint64_t s_warp[SB_SCAN_THREADS / 32];
    __shared__ uint64_t s_carry;
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t b0 = 0; b0 < nb; b0 += SB_SCAN_THREADS) {
        const uint64_t b = b0 + threadIdx.x;
        const uint64_t v = b < nb ? (uint64_t)hist[b] : 0ull;
        uint64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t o = __shfl_up_sync(full, incl, d);
            if ((int)lane >= d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint64_t off = s_carry;
        for (uint32_t w = 0; w < warp; w++) off += s_warp[w];
        const uint64_t excl = off + incl - v;
        if (b < nb) { base[b] = excl; cursor[b] = excl; }
        __syncthreads();
        if (threadIdx.x == SB_SCAN_THREADS - 1) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) base[nb] = s_carry;
}

// One CTA per bucket: CSR offsets of the bucket's patterns and the final place of every record.
// Every pattern's records are contiguous inside the bucket and in SA-row order (see BucketOut), so the run boundaries
// give first[k] / end[k] of pattern k with plain shared-memory stores -- no atomics, no dependence on the order in which
// the patterns arrived.  out_offs[i] = base[bucket] + (exclusive prefix sum of the run lengths in pattern order);
// record p of pattern k goes to positions[out_offs[k] + (p - first[k])], through a shared-memory staging buffer when the
// bucket fits it, so that the global stores are coalesced.
// The per-pattern arrays are padded by one word per 32 (index k lives at k + k/32): a thread then owns 32 consecutive
// patterns without bank conflicts, and the scan needs one pass instead of one barrier pair per 256 patterns.
// A bucket must hold fewer than 2^32 records (the host falls back to the radix sort-back above 2^32 in total).
constexpr int SB_PLACE_THREADS = 256;
constexpr uint32_t SB_PAD = SB_BUCKET + SB_BUCKET / 32;
// Staging the placed positions in shared memory (coalesced stores) was measured SLOWER than storing them straight into the
// bucket's 16 KB output window, which L2 merges into full lines: the staging buffer cut the occupancy from 6 to 3 CTAs per
// SM (0.99 against 0.84 ms per 10^8 records).  SVFM_SB_STAGE > 0 brings it back for buckets of up to that many records.
#ifndef SVFM_SB_STAGE
#define SVFM_SB_STAGE 0
#endif
constexpr uint32_t SB_STAGE = SVFM_SB_STAGE;
constexpr size_t SB_PLACE_SMEM = (2 * (size_t)SB_PAD) * 4;  // + SB_STAGE * sizeof(P)
__device__ __forceinline__ uint32_t sb_pad(uint32_t k) { return k + (k >> 5); }
template <class P>
__global__ void __launch_bounds__(SB_PLACE_THREADS)
sb_place_kernel(const SbRec<P>* __restrict__ recs, const uint64_t* __restrict__ base, uint64_t n, uint64_t nb,
                void* __restrict__ out_offs, int offs32, P* __restrict__ positions) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    uint32_t* s_first = reinterpret_cast<uint32_t*>(s_dyn);          // SB_PAD
    uint32_t* s_off = s_first + SB_PAD;                               // SB_PAD: end[k], then exclusive offset of k
    P* s_pos = reinterpret_cast<P*>(s_off + SB_PAD);                  // SB_STAGE
    __shared__ uint32_t s_warp[SB_PLACE_THREADS / 32];
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t b = blockIdx.x;
    const uint64_t r0 = base[b];
    const uint32_t m = (uint32_t)(base[b + 1] - r0);                  // records of this bucket
    const uint64_t p0 = b << SB_SHIFT;                                // first pattern of this bucket
    const uint32_t np = (uint32_t)(n - p0 < (uint64_t)SB_BUCKET ? n - p0 : (uint64_t)SB_BUCKET);
    for (uint32_t k = threadIdx.x; k < SB_PAD; k += SB_PLACE_THREADS) { s_first[k] = 0; s_off[k] = 0; }
    __syncthreads();
    const SbRec<P>* rb = recs + r0;
    // ---- run boundaries
    for (uint32_t q0 = 0; q0 < m; q0 += SB_PLACE_THREADS) {
        const uint32_t p = q0 + threadIdx.x;
        const bool valid = p < m;
        const uint32_t k = valid ? (rb[p].idx & (SB_BUCKET - 1)) : 0xffffffffu;
        uint32_t k_prev = __shfl_up_sync(full, k, 1);
        uint32_t k_next = __shfl_down_sync(full, k, 1);
        if (lane == 0) k_prev = (valid && p > 0) ? (rb[p - 1].idx & (SB_BUCKET - 1)) : 0xffffffffu;
        if (lane == 31) k_next = (valid && p + 1 < m) ? (rb[p + 1].idx & (SB_BUCKET - 1)) : 0xffffffffu;
        if (valid) {
            if (k != k_prev) s_first[sb_pad(k)] = p;
            if (k != k_next) s_off[sb_pad(k)] = p + 1;   // end of the run
        }
    }
    __syncthreads();
    // ---- exclusive prefix sums of the run lengths, in pattern order: thread t owns patterns [t*PER, (t+1)*PER)
    constexpr uint32_t PER = SB_BUCKET / SB_PLACE_THREADS;
    static_assert(PER <= 32 && 32 % PER == 0, "a thread's patterns stay inside one padding group");
    {
        const uint32_t a0 = sb_pad(threadIdx.x * PER);
        uint32_t sum = 0;
#pragma unroll
        for (uint32_t i = 0; i < PER; i++) sum += s_off[a0 + i] - s_first[a0 + i];   // patterns without records: 0 - 0
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(full, incl, d);
            if ((int)lane >= d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
        for (uint32_t w = 0; w < warp; w++) run += s_warp[w];
#pragma unroll
        for (uint32_t i = 0; i < PER; i++) {
            const uint32_t v = s_off[a0 + i] - s_first[a0 + i];
            s_off[a0 + i] = run;
            run += v;
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < np; k += SB_PLACE_THREADS) {
        const uint64_t o = r0 + s_off[sb_pad(k)];
        if (offs32) reinterpret_cast<uint32_t*>(out_offs)[p0 + k] = (uint32_t)o;
        else reinterpret_cast<uint64_t*>(out_offs)[p0 + k] = o;
    }
    if (b + 1 == nb && threadIdx.x == 0) {
        if (offs32) reinterpret_cast<uint32_t*>(out_offs)[n] = (uint32_t)base[nb];
        else reinterpret_cast<uint64_t*>(out_offs)[n] = base[nb];
    }
    // ---- placement
    P* out = positions + r0;
    const bool staged = m <= SB_STAGE;
    for (uint32_t p = threadIdx.x; p < m; p += SB_PLACE_THREADS) {
        const SbRec<P> r = rb[p];
        const uint32_t k = sb_pad(r.idx & (SB_BUCKET - 1));
        const uint32_t slot = s_off[k] + (p - s_first[k]);
        SVFM_ASSERT((r.idx >> SB_SHIFT) == b && p >= s_first[k] && slot < m);   // the record is in its own bucket, inside its pattern's run
        if (staged) s_pos[slot] = r.pos;
        else out[slot] = r.pos;
    }
    if (staged) {
        __syncthreads();
        for (uint32_t p = threadIdx.x; p < m; p += SB_PLACE_THREADS) out[p] = s_pos[p];
    }
}

// counts_out[idx[w]] = cnt[w]: the work items arrive grouped by the top bits of idx (the last sweep round of `count`
// partitions by them), so the scattered 4/8-byte stores of the CTAs running at any time fall into a window of a few MB
// that L2 turns into full-line writes.
template <class P>
__global__ void __launch_bounds__(256)
scatter_counts_kernel(const uint32_t* __restrict__ idx, const P* __restrict__ cnt, uint64_t n, P* __restrict__ counts_out) {
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += (uint64_t)gridDim.x * blockDim.x)
        counts_out[idx[w]] = cnt[w];
}

// CSR offsets from the pattern indices of the records once they are sorted by pattern: first[k] = index of
// the first record of pattern k (patterns without records keep the 0xff.. fill); out_offs is then the
// reverse running minimum of `first` with out_offs[n] = total (done with a device scan by the caller).
static __global__ void run_starts_kernel(const uint32_t* __restrict__ sorted_key, uint64_t total, uint64_t* __restrict__ first) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t k = sorted_key[t];
        if (t == 0 || sorted_key[t - 1] != k) first[k] = t;
    }
}

}  // namespace svfm
