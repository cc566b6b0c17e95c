// search_kernels.cuh -- hand-written CUDA (sm_100a) for the hot path: backward-search count and
// SA-sampled locate over a batch of patterns.  HBM-bound integer work: no tensor cores, no TMA.
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
};

constexpr int SEARCH_THREADS = 256;
constexpr uint32_t HEAVY_ROWS = 1024;  // patterns with more SA rows than this are located row-parallel

// One backward-search step for both range ends: FmIndex::next_pos_range (locate/mod.rs:39-45) =
// count_array[s] + get_next_rank(pos, s) (bwm/mod.rs:197-215) for pos in {sp, ep}.
// All four gathers (2 checkpoint words + 2 blocks) are issued before the first use; when both ends
// fall into the same block (the common case once the interval is short) the block and the
// checkpoint word are fetched once.
template <class P, int NPL, int VBITS>
__device__ __forceinline__ void backward_step(const DevIndex<P>& ix, const P* __restrict__ s_count, uint32_t sym,
                                              P& sp, P& ep) {
    uint64_t q0, q1;
    uint32_t r0, r1;
    rank_addr<P, VBITS>(ix, sp, q0, r0);
    rank_addr<P, VBITS>(ix, ep, q1, r1);
    const P c = s_count[sym];
    Block<NPL, VBITS> b0;
    typename Block<NPL, VBITS>::W m[VecTraits<VBITS>::WORDS];
    if (q0 == q1) {
        const P ck = ld_gather<P>(ix.rank_checkpoints + q0 * ix.symbol_count + sym);
        b0.load(ix.blocks, q0);
        b0.match_mask(sym, m);
        sp = c + ck + (P)Block<NPL, VBITS>::prefix_count(m, r0);
        ep = c + ck + (P)Block<NPL, VBITS>::prefix_count(m, r1);
    } else {
        Block<NPL, VBITS> b1;
        const P ck0 = ld_gather<P>(ix.rank_checkpoints + q0 * ix.symbol_count + sym);
        const P ck1 = ld_gather<P>(ix.rank_checkpoints + q1 * ix.symbol_count + sym);
        b0.load(ix.blocks, q0);
        b1.load(ix.blocks, q1);
        b0.match_mask(sym, m);
        sp = c + ck0 + (P)Block<NPL, VBITS>::prefix_count(m, r0);
        b1.match_mask(sym, m);
        ep = c + ck1 + (P)Block<NPL, VBITS>::prefix_count(m, r1);
    }
}

// Bit position of the symbol that sits `from_end` symbols before the end of the pattern inside a packed key.
// m1 == 0: layout A, the LAST symbol in the top bits (reverse-lexicographic sort order).
// m1  > 0: layout C, the trailing m1 symbols in the top bits in FORWARD order (sorting on them orders the batch by
//          the SA interval of that m1-symbol suffix), the rest of the pattern below them.
__host__ __device__ __forceinline__ uint32_t key_shift(uint32_t from_end, uint32_t bits, uint32_t m1) {
    return from_end < m1 ? 64u - bits * (m1 - from_end) : 64u - bits * (from_end + 1);
}

// Locality key of a pattern = its trailing symbols packed `bits` per symbol from the top of a u64, the
// LAST symbol most significant (backward search consumes the pattern from its end, so patterns that share
// a suffix walk the same checkpoint rows and blocks for as many steps as the shared suffix is long).
// The key doubles as the encoded pattern: the search kernel takes the last min(len, 64/bits) symbols from
// it and never touches the pattern bytes again unless the pattern is longer.  Also validates the batch.
__global__ void __launch_bounds__(SEARCH_THREADS)
pack_keys_kernel(const uint8_t* __restrict__ table, uint32_t S, const PatternBatch pb, uint32_t bits, uint32_t m1,
                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int* __restrict__ err) {
    __shared__ uint8_t s_table[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_table[i] = table ? table[i] : (uint8_t)i;
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
            key |= (uint64_t)s << key_shift(j, bits, m1);
        }
        keys[i] = key;
        vals[i] = (uint32_t)i;
    }
    if (errbits) atomicOr(err, errbits);
}

// State of a pattern between the two phases of a large batch (see SearchIO): moved as the VALUE of the
// radix sort that re-orders the batch by its current SA position.
template <class P>
struct Item {
    uint64_t key;   // packed trailing symbols (pack_keys_kernel)
    uint32_t idx;   // the caller's pattern index
    uint32_t pi;    // symbols still to consume (forward index of the next one + 1); 0 = finished
    P cnt;          // ep - sp
};

template <class P>
struct SearchIO {
    const uint64_t* keys;   // packed keys in work order, or NULL (symbols come from the pattern bytes)
    const uint32_t* idx;    // pattern index of each work item, or NULL (identity)
    uint32_t bits;          // bits per symbol in the key
    uint32_t max_steps;     // backward steps to run in this launch (0xffffffff = to the end)
    const P* sp_in;         // resume: current sp of each work item (sorted ascending); NULL = start from the kLTS seed
    const Item<P>* items_in;
    P* sp_out;              // work order, nullable
    P* cnt_out;             // work order, nullable
    Item<P>* items_out;     // nullable: state for the next phase
    uint32_t* idx_out;      // nullable: pattern index per work item (resume launches)
    unsigned long long* heavy_seen;  // nullable
    int* err;
};

// FmIndex::get_pos_range (locate/with_slice.rs:21-33) for one pattern per thread, grid-stride.
// Work item w handles pattern idx[w] (idx == NULL: identity).  With keys != NULL the last 64/bits symbols of
// the pattern come out of keys[w] (see pack_keys_kernel).  Outputs in WORK order, coalesced.
//
// Large batches run it twice.  Phase 1 (work order = locality sort by trailing symbols) runs the kLTS seed and
// the first max_steps backward steps: neighbouring lanes share rows/blocks there.  Then the batch is radix-sorted
// by the current sp, and phase 2 (RESUME) finishes the search in SA order: LF-mapping keeps the relative order
// of rows inside a symbol class, so from then on the rows a window of neighbouring work items touches stay
// confined to a few narrow, slowly advancing windows of the checkpoint/block arrays -- every line is fetched from
// DRAM once per step instead of once per pattern (ncu, round 1: the single-phase kernel read 155 GB for 100 M
// 20-mers because its last ~8 steps were random sector pairs with >= 64-byte line fills).
template <class P, int NPL, int VBITS, bool RESUME>
__global__ void __launch_bounds__(SEARCH_THREADS)
search_kernel(const DevIndex<P> ix, const PatternBatch pb, const SearchIO<P> io) {
    __shared__ uint8_t s_table[256];
    __shared__ P s_count[65];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_table[i] = ix.table ? ix.table[i] : (uint8_t)i;
    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    __syncthreads();

    const uint32_t S = ix.symbol_count;
    const uint32_t k = ix.kmer_size;
    const uint32_t bits = io.bits;
    const bool have_keys = RESUME || io.keys != nullptr;
    const uint32_t in_key = have_keys ? 64u / bits : 0u;  // symbols (from the end) available in the key
    const uint64_t sym_mask = (1ull << bits) - 1;
    int errbits = 0;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < pb.n; w += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t i, key, pi = 0;
        P sp = 0, ep = 0;
        if (RESUME) {
            const Item<P> it = io.items_in[w];
            i = it.idx;
            key = it.key;
            pi = it.pi;
            sp = io.sp_in[w];
            ep = (P)(sp + it.cnt);
        } else {
            i = io.idx ? (uint64_t)io.idx[w] : w;
            key = io.keys ? io.keys[w] : 0ull;
        }
        uint64_t base, len;
        if (pb.offs) { base = pb.offs[i]; len = pb.offs[i + 1] - base; }
        else { base = i * (uint64_t)pb.fixed_len; len = pb.fixed_len; }
        const uint8_t* p = pb.pats + base;
        if (len == 0) {
            errbits |= ERRBIT_EMPTY_PATTERN;
        } else {
            // logical (forward) symbol j of the pattern
            auto sym_at = [&](uint64_t j) -> uint32_t {
                const uint64_t from_end = len - 1 - j;
                if (from_end < in_key) return (uint32_t)((key >> key_shift((uint32_t)from_end, bits, 0)) & sym_mask);
                uint32_t s = s_table[__ldg(p + (pb.reversed ? from_end : j))];
                if (s >= S) { errbits |= ERRBIT_BAD_SYMBOL; s = S - 1; }
                return s;
            };
            if (!RESUME) {
                // CountArrayView::get_initial_pos_range_and_idx_of_pattern (count_array.rs:203-233)
                if (len < k) {
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
            }
            // LF mapping (with_slice.rs:27-31): stops as soon as the interval is empty
            uint32_t steps = io.max_steps;
            while (sp < ep && pi > 0 && steps > 0) {
                pi -= 1;
                steps -= 1;
                backward_step<P, NPL, VBITS>(ix, s_count, sym_at(pi), sp, ep);
            }
            if (!(sp < ep)) pi = 0;  // empty interval: finished
        }
        const P cnt = (P)(ep - sp);
        if (io.sp_out) io.sp_out[w] = sp;
        if (io.cnt_out) io.cnt_out[w] = cnt;
        if (io.idx_out) io.idx_out[w] = (uint32_t)i;
        if (io.items_out) {
            Item<P> it;
            it.key = key;
            it.idx = (uint32_t)i;
            it.pi = (uint32_t)pi;
            it.cnt = cnt;
            io.items_out[w] = it;
        }
        if (io.heavy_seen && (uint64_t)cnt > HEAVY_ROWS) atomicAdd(io.heavy_seen, 1ull);  // rare: sizes the heavy list
    }
    if (errbits) atomicOr(io.err, errbits);
}

// ---- streaming search of a dense fixed-length batch ---------------------------------------------------------
// The batch is sorted by the forward order of its trailing m1 symbols (layout C keys), i.e. by the SA interval of
// that suffix.  seed_kernel resolves the suffix ONCE per run of equal suffixes: a warp owns 32 consecutive work
// items, the first lane of every run (and lane 0) runs the kLTS seed (count_array.rs:203-233) and the first
// m1-k backward steps (with_slice.rs:27-31), and the result is broadcast to the run with shuffles.  These early
// steps touch at most S^(k+j) distinct rows, which stay L2-resident.  step_kernel then runs the remaining steps
// for every work item, one launch per group of steps, state (sp, count) in global memory: because LF-mapping
// keeps the relative order of rows inside a symbol class, the rows touched by a window of neighbouring work items
// stay confined to a few narrow windows of the checkpoint/block arrays that advance monotonically -- every line
// comes from DRAM once per step, and repeats are L2 hits (measured on B200: 290 G random sector hits/s in L2
// versus 55 G/s from HBM).
template <class P, int NPL, int VBITS>
__global__ void __launch_bounds__(SEARCH_THREADS)
seed_kernel(const DevIndex<P> ix, const uint64_t* __restrict__ keys, uint64_t n, uint32_t bits, uint32_t m1,
            P* __restrict__ sp_out, P* __restrict__ cnt_out, unsigned long long* __restrict__ heavy_seen) {
    __shared__ P s_count[65];
    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    __syncthreads();
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t k = ix.kmer_size;
    const uint64_t sym_mask = (1ull << bits) - 1;
    const uint32_t top_shift = 64u - bits * m1;
    const uint64_t warp_id = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t w0 = warp_id * 32; w0 < n; w0 += n_warps * 32) {
        const uint64_t w = w0 + lane;
        const bool valid = w < n;
        const uint64_t key = valid ? keys[w] : 0ull;
        const uint64_t prefix = key >> top_shift;
        const uint64_t prev = __shfl_up_sync(full, prefix, 1);
        const bool leader = valid && (lane == 0 || prefix != prev);
        P sp = 0, ep = 0;
        if (leader) {
            uint64_t start = 0;
            for (uint32_t t = 0; t < k; t++) {  // forward index len-k+t  <->  from_end k-1-t
                const uint32_t sym = (uint32_t)((key >> key_shift(k - 1 - t, bits, m1)) & sym_mask);
                start += (uint64_t)(sym + 1) * __ldg(ix.kmer_multiplier + t);
            }
            sp = __ldg(ix.kmer_count_table + (start - 1));
            ep = __ldg(ix.kmer_count_table + start);
            for (uint32_t j = k; j < m1 && sp < ep; j++)
                backward_step<P, NPL, VBITS>(ix, s_count, (uint32_t)((key >> key_shift(j, bits, m1)) & sym_mask), sp, ep);
        }
        const unsigned leaders = __ballot_sync(full, leader);
        const unsigned upto = leaders & (0xffffffffu >> (31u - lane));
        const int src = upto ? 31 - __clz(upto) : 0;
        sp = __shfl_sync(full, sp, src);
        ep = __shfl_sync(full, ep, src);
        if (valid) {
            const P cnt = (P)(ep - sp);
            sp_out[w] = sp;
            cnt_out[w] = cnt;
            if (heavy_seen && (uint64_t)cnt > HEAVY_ROWS) atomicAdd(heavy_seen, 1ull);
        }
    }
}

// `steps` backward steps (symbols from_end = j_first, j_first+1, ...) for every live work item.
template <class P, int NPL, int VBITS>
__global__ void __launch_bounds__(SEARCH_THREADS)
step_kernel(const DevIndex<P> ix, const uint64_t* __restrict__ keys, uint64_t n, uint32_t bits, uint32_t m1,
            uint32_t j_first, uint32_t steps, P* __restrict__ sp_io, P* __restrict__ cnt_io,
            unsigned long long* __restrict__ heavy_seen) {
    __shared__ P s_count[65];
    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    __syncthreads();
    const uint64_t sym_mask = (1ull << bits) - 1;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += (uint64_t)gridDim.x * blockDim.x) {
        P cnt = cnt_io[w];
        if (cnt != 0) {
            P sp = sp_io[w];
            P ep = (P)(sp + cnt);
            const uint64_t key = keys[w];
            for (uint32_t t = 0; t < steps && sp < ep; t++)
                backward_step<P, NPL, VBITS>(ix, s_count, (uint32_t)((key >> key_shift(j_first + t, bits, m1)) & sym_mask), sp, ep);
            cnt = (P)(ep - sp);
            sp_io[w] = sp;
            cnt_io[w] = cnt;
        }
        if (heavy_seen && (uint64_t)cnt > HEAVY_ROWS) atomicAdd(heavy_seen, 1ull);
    }
}

// FmIndex::write_locations_to_buffer (locate/mod.rs:14-37) for ONE SA row: LF-walk to the nearest
// sampled row (BwmView::get_pre_rank_and_symidx, bwm/mod.rs:217-236), then
// SuffixArrayView::get_location_of (suffix_array/mod.rs:100-105).
template <class P, int NPL, int VBITS>
__device__ __forceinline__ P locate_row(const DevIndex<P>& ix, const P* __restrict__ s_count, P pos) {
    P offset = 0;
    for (;;) {
        // pos % sampling_ratio != 0
        uint64_t quot;
        bool sampled;
        if (ix.ratio_mask != 0xffffffffu) {
            sampled = ((uint32_t)pos & ix.ratio_mask) == 0;
            quot = (uint64_t)pos >> ix.ratio_shift;
        } else {
            quot = (uint64_t)pos / ix.sampling_ratio;
            sampled = (quot * ix.sampling_ratio == (uint64_t)pos);
        }
        if (sampled) return (P)(ld_gather<P>(ix.suffix_array + quot) + offset);
        if (pos == (P)(ix.sentinel_index - 1)) return offset;  // None arm, locate/mod.rs:27-30
        uint64_t q;
        uint32_t rem;
        rank_addr<P, VBITS>(ix, pos, q, rem);
        Block<NPL, VBITS> b;
        b.load(ix.blocks, q);
        const uint32_t s = b.symidx_of(rem);
        const P ck = ld_gather<P>(ix.rank_checkpoints + q * ix.symbol_count + s);
        pos = (P)(s_count[s] + ck + (P)b.remain_count(rem, s));  // remain_count(0, .) == 0
        offset += 1;
    }
}

constexpr int LOCATE_THREADS = 256;

template <class P>
struct HeavyList {
    P* sp;                    // first SA row of the pattern
    P* cnt;                   // number of rows
    uint64_t* obase;          // where its positions start in the output
    uint32_t* pat;            // its pattern index (record mode)
    unsigned long long* n;    // entries appended so far
    uint64_t capacity;
};

// Locate for patterns in work order: work item w has SA rows sp_work[w] .. +cnt_work[w] and writes the
// position of row sp+j at positions[offs[w] + j] (offs = exclusive prefix sums of cnt_work), i.e. in SA-row
// order inside a pattern (the reference's order, locate/mod.rs:19).  In record mode (rec_key != NULL) it also
// writes rec_key[offs[w] + j] = idx[w], the caller's pattern index, so that a stable sort by rec_key brings
// the positions into the caller's CSR order without any random scatter.
// One warp owns 32 consecutive work items and spreads ALL their rows over its lanes (warp prefix sum of the
// counts, then each lane finds the owner of its row with a shuffle binary search), so a pattern with many
// rows does not serialise one lane.  Patterns with more than HEAVY_ROWS rows are deferred to
// locate_rows_kernel through the heavy list.
template <class P, int NPL, int VBITS>
__global__ void __launch_bounds__(LOCATE_THREADS)
locate_warp_kernel(const DevIndex<P> ix, const uint32_t* __restrict__ idx, const P* __restrict__ sp_work,
                   const P* __restrict__ cnt_work, const uint64_t* __restrict__ offs, uint64_t n,
                   P* __restrict__ positions, uint32_t* __restrict__ rec_key, HeavyList<P> heavy) {
    __shared__ P s_count[65];
    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    __syncthreads();
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp_id = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t w0 = warp_id * 32; w0 < n; w0 += n_warps * 32) {
        const uint64_t w = w0 + lane;
        uint32_t c = 0;
        P sp = 0;
        uint64_t obase = 0;
        uint32_t pat = 0;
        if (w < n) {
            const P cw = cnt_work[w];
            if (cw != 0) {
                sp = sp_work[w];
                obase = offs[w];
                pat = idx ? idx[w] : (uint32_t)w;
                if ((uint64_t)cw > HEAVY_ROWS) {
                    const unsigned long long h = atomicAdd(heavy.n, 1ull);
                    if (h < heavy.capacity) { heavy.sp[h] = sp; heavy.cnt[h] = cw; heavy.obase[h] = obase; heavy.pat[h] = pat; }
                } else {
                    c = (uint32_t)cw;
                }
            }
        }
        if (__all_sync(full, c <= 1u)) {
            // common case: at most one row per pattern, no redistribution needed
            if (c) {
                positions[obase] = locate_row<P, NPL, VBITS>(ix, s_count, sp);
                if (rec_key) rec_key[obase] = pat;
            }
            continue;
        }
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(full, incl, d);
            if ((int)lane >= d) incl += v;
        }
        const uint32_t excl = incl - c;
        const uint32_t total = __shfl_sync(full, incl, 31);
        for (uint32_t r0 = 0; r0 < total; r0 += 32) {
            const uint32_t r = r0 + lane;
            // owner = first lane whose inclusive sum exceeds r
            int o = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t v = __shfl_sync(full, incl, o + step - 1);
                if (v <= r) o += step;
            }
            o &= 31;
            const uint32_t e = __shfl_sync(full, excl, o);
            const P osp = __shfl_sync(full, sp, o);
            const uint64_t oo = __shfl_sync(full, obase, o);
            const uint32_t opat = __shfl_sync(full, pat, o);
            if (r < total) {
                const uint32_t j = r - e;
                positions[oo + j] = locate_row<P, NPL, VBITS>(ix, s_count, (P)(osp + (P)j));
                if (rec_key) rec_key[oo + j] = opat;
            }
        }
    }
}

// Row-parallel locate for the heavy list: one thread per SA row; row t belongs to the entry h with
// offs[h] <= t < offs[h+1].  The per-block window of candidate entries is found once with two binary
// searches; each thread then searches only that window.
template <class P, int NPL, int VBITS>
__global__ void __launch_bounds__(LOCATE_THREADS)
locate_rows_kernel(const DevIndex<P> ix, const P* __restrict__ sp, const uint64_t* __restrict__ offs,
                   const uint64_t* __restrict__ obase, const uint32_t* __restrict__ pat, uint64_t n, uint64_t total,
                   P* __restrict__ positions, uint32_t* __restrict__ rec_key) {
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
    positions[dst] = locate_row<P, NPL, VBITS>(ix, s_count, row);
    if (rec_key) rec_key[dst] = __ldg(pat + lo);
}

// CSR offsets from the pattern indices of the records once they are sorted by pattern: first[k] = index of
// the first record of pattern k (patterns without records keep the 0xff.. fill); out_offs is then the
// reverse running minimum of `first` with out_offs[n] = total (done with a device scan by the caller).
__global__ void run_starts_kernel(const uint32_t* __restrict__ sorted_key, uint64_t total, uint64_t* __restrict__ first) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t k = sorted_key[t];
        if (t == 0 || sorted_key[t - 1] != k) first[k] = t;
    }
}

}  // namespace svfm
