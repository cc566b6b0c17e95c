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

// FmIndex::get_pos_range (locate/with_slice.rs:21-33) for one pattern per thread, grid-stride.
// Writes sp_out[i] (may be NULL) and cnt_out[i] = ep - sp.
template <class P, int NPL, int VBITS>
__global__ void __launch_bounds__(SEARCH_THREADS)
search_kernel(const DevIndex<P> ix, const PatternBatch pb, P* __restrict__ sp_out, P* __restrict__ cnt_out,
              int* __restrict__ err) {
    __shared__ uint8_t s_table[256];
    __shared__ P s_count[65];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_table[i] = ix.table ? ix.table[i] : (uint8_t)i;
    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    __syncthreads();

    const uint32_t S = ix.symbol_count;
    const uint32_t k = ix.kmer_size;
    int errbits = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pb.n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t base, len;
        if (pb.offs) { base = pb.offs[i]; len = pb.offs[i + 1] - base; }
        else { base = i * (uint64_t)pb.fixed_len; len = pb.fixed_len; }
        const uint8_t* p = pb.pats + base;
        P sp = 0, ep = 0;
        if (len == 0) {
            errbits |= ERRBIT_EMPTY_PATTERN;
        } else {
            // logical (forward) symbol j of the pattern
            auto sym_at = [&](uint64_t j) -> uint32_t {
                uint32_t s = s_table[__ldg(p + (pb.reversed ? (len - 1 - j) : j))];
                if (s >= S) { errbits |= ERRBIT_BAD_SYMBOL; s = S - 1; }
                return s;
            };
            // CountArrayView::get_initial_pos_range_and_idx_of_pattern (count_array.rs:203-233)
            uint64_t idx;
            if (len < k) {
                uint64_t start = 0;
                for (uint64_t j = 0; j < len; j++) start += (uint64_t)(sym_at(j) + 1) * __ldg(ix.kmer_multiplier + j);
                const uint64_t end = start + __ldg(ix.kmer_multiplier + (len - 1)) - 1;
                sp = __ldg(ix.kmer_count_table + (start - 1));
                ep = __ldg(ix.kmer_count_table + end);
                idx = 0;
            } else {
                uint64_t start = 0;
                for (uint32_t j = 0; j < k; j++) start += (uint64_t)(sym_at(len - k + j) + 1) * __ldg(ix.kmer_multiplier + j);
                sp = __ldg(ix.kmer_count_table + (start - 1));
                ep = __ldg(ix.kmer_count_table + start);
                idx = len - k;
            }
            // LF mapping (with_slice.rs:27-31): stops as soon as the interval is empty
            while (sp < ep && idx > 0) {
                idx -= 1;
                backward_step<P, NPL, VBITS>(ix, s_count, sym_at(idx), sp, ep);
            }
        }
        if (sp_out) sp_out[i] = sp;
        cnt_out[i] = (P)(ep - sp);
    }
    if (errbits) atomicOr(err, errbits);
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

// One thread per output slot t in [0, total): slot t belongs to the pattern i with
// out_offs[i] <= t < out_offs[i+1] and is SA row sp[i] + (t - out_offs[i]), which reproduces the
// reference's output order (for pos in sp..ep, locate/mod.rs:19).  The per-block window of candidate
// patterns is found once with two binary searches; each thread then searches only that window.
template <class P, int NPL, int VBITS>
__global__ void __launch_bounds__(LOCATE_THREADS)
locate_kernel(const DevIndex<P> ix, const P* __restrict__ sp, const uint64_t* __restrict__ out_offs, uint64_t n,
              uint64_t total, P* __restrict__ positions) {
    __shared__ P s_count[65];
    __shared__ uint64_t s_win[2];
    for (int i = threadIdx.x; i <= (int)ix.symbol_count; i += blockDim.x) s_count[i] = ix.count_array[i];
    const uint64_t t0 = (uint64_t)blockIdx.x * LOCATE_THREADS;
    if (threadIdx.x < 2) {
        // largest i in [0, n) with out_offs[i] <= target
        uint64_t target = threadIdx.x == 0 ? t0 : (t0 + LOCATE_THREADS - 1 < total ? t0 + LOCATE_THREADS - 1 : total - 1);
        uint64_t lo = 0, hi = n;  // invariant: out_offs[lo] <= target < out_offs[hi]
        while (hi - lo > 1) {
            uint64_t mid = lo + ((hi - lo) >> 1);
            if (__ldg(out_offs + mid) <= target) lo = mid; else hi = mid;
        }
        s_win[threadIdx.x] = lo;
    }
    __syncthreads();
    const uint64_t t = t0 + threadIdx.x;
    if (t >= total) return;
    uint64_t lo = s_win[0], hi = s_win[1] + 1;
    while (hi - lo > 1) {
        uint64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(out_offs + mid) <= t) lo = mid; else hi = mid;
    }
    const P row = (P)(__ldg(sp + lo) + (P)(t - __ldg(out_offs + lo)));
    positions[t] = locate_row<P, NPL, VBITS>(ix, s_count, row);
}

}  // namespace svfm
