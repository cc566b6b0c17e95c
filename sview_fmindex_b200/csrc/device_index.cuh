// device_index.cuh -- device-side view of the uploaded blob and the rank primitives.
// The blob sits in HBM byte-for-byte; these are typed pointers into it (the device twin of the
// reference's CountArrayView / SuffixArrayView / BwmView, components/*.rs).
// Citations are relative to the reference's sview-fmindex/src/.
#pragma once
#include <cassert>
#include <cstdint>
#include <cuda_runtime.h>

// Bounds / protocol assertions inside the kernels, compiled in with -DSVFM_DEBUG_CHECKS (tools/build_variant.sh checks
// "-DSVFM_DEBUG_CHECKS"; tests/sanitize_case.py runs every kernel of the hot path under that build).  compute-sanitizer
// is closed on the GPU pool this was developed on, so these device-side asserts are the tool-independent evidence that
// the look-back protocol, the partition arithmetic and the bucket reservations stay inside their buffers.
#ifdef SVFM_DEBUG_CHECKS
#define SVFM_ASSERT(x) assert(x)
#else
#define SVFM_ASSERT(x) ((void)0)
#endif

namespace svfm {

template <class P>
struct DevIndex {
    // CountArrayView (components/count_array.rs:21-29)
    const P* count_array;             // P[S+1]
    const uint64_t* kmer_multiplier;  // usize[k]
    const P* kmer_count_table;        // P[(S+1)^k]
    // SuffixArrayView (components/suffix_array/mod.rs:21-26)
    const P* suffix_array;
    // BwmView (components/bwm/mod.rs:19-26)
    const P* rank_checkpoints;        // row-major [block][symbol], row stride = symbol_count
    const void* blocks;               // BlockN<V>[blocks_len]
    const uint8_t* table;             // EncodingTable bytes, or NULL for PassThrough
    P sentinel_index;
    uint32_t symbol_count;            // S (row stride; bwm/mod.rs:158)
    uint32_t kmer_size;               // k
    uint32_t sampling_ratio;          // r
    uint32_t ratio_mask;              // r-1 if r is a power of two, else 0xffffffff
    uint32_t ratio_shift;             // log2(r) if power of two
    // extended k-mer table derived from the blob at load (search_kernels.cuh): P[2 * s_eff^ext_m] (sp, count)
    const P* ext;                     // NULL = not built
    uint32_t ext_m;                   // symbols resolved by one lookup
    uint32_t s_eff;                   // symbols that occur in the text
    uint8_t sym_rank[64];             // symbol index -> rank among the occurring symbols, 0xff = never occurs
    uint8_t present[64];              // rank -> symbol index
    // interleaved occ copy derived from the blob at load (opt-out SVFM_TUNE_ILV): entry q =
    // { block q | checkpoint row q } in ONE aligned slot of ilv_stride bytes, so that a rank query costs one DRAM fetch
    // instead of two or three.  Used by the gather-bound kernels (search_kernel, locate); NULL = not built.
    const uint8_t* ilv;
    uint32_t ilv_stride;              // bytes per entry (32, 64, 128 or a multiple of 128)
    uint32_t ilv_ck_off;              // byte offset of the checkpoint row inside an entry
    // packed copy of the indexed text, derived from the blob at load (search_kernels.cuh, "text verification"): symbol i =
    // rank among the occurring symbols, text_bits (1, 2, 4 or 8) bits each, little-endian inside 32-bit words; NULL = not built
    const uint32_t* text;
    uint32_t text_bits;
    // expanded suffix array derived from the blob at load (search_kernels.cuh, "expanded suffix array"): the text position
    // of EVERY SA row, 32-bit entries when the text is shorter than 2^32 symbols (fsa32), else P; NULL = not built
    const void* fsa;
    uint32_t fsa32;
    // compact occ copy for the sweep rounds, derived from the blob at load (search_kernels.cuh, "sweep occ copy"): one
    // 32-byte sector per block = its planes + 16-bit checkpoint deltas of the occurring symbols; NULL = not built
    const uint8_t* swp;
};

// the symbol maps alone (kernels that do not touch the index arrays)
struct DevSyms {
    uint32_t symbol_count;
    uint32_t s_eff;
    uint8_t sym_rank[64];
};

// ---- loads --------------------------------------------------------------------------------------
// Index gathers go through the read-only path and allocate in L1: the planes of one block are fetched by
// separate instructions that hit the same 1-2 sectors, and neighbouring lanes/warps of a locality-sorted
// batch share sectors (ncu, round 1: without L1 allocation every plane load was a separate L2 request).
__device__ __forceinline__ uint32_t ld_gather_u32(const uint32_t* p) { return __ldg(p); }
__device__ __forceinline__ uint64_t ld_gather_u64(const uint64_t* p) {
    return __ldg(reinterpret_cast<const unsigned long long*>(p));
}
template <class P> __device__ __forceinline__ P ld_gather(const P* p);
template <> __device__ __forceinline__ uint32_t ld_gather<uint32_t>(const uint32_t* p) { return ld_gather_u32(p); }
template <> __device__ __forceinline__ uint64_t ld_gather<uint64_t>(const uint64_t* p) { return ld_gather_u64(p); }

// ---- one occ block held in registers ---------------------------------------------------------------
// BlockN<V>([V; N]) (blocks/block2.rs:7 .. block6.rs:7).  Symbol j of the block is bit VBITS-1-j of
// every plane (vectorize shifts left once per symbol, block3.rs:20-35).  A plane is kept as WORDS
// machine words ordered from the most significant (first symbols) to the least:
//   u32  -> 1 x 32-bit word      u64 -> 1 x 64-bit word
//   u128 -> 2 x 64-bit words; little-endian in memory, so word 0 (symbols 0..63) is at byte offset 8.
template <int VBITS> struct VecTraits;
template <> struct VecTraits<32>  { using W = uint32_t; static constexpr int WORDS = 1; static constexpr int WBITS = 32; static constexpr int LOG2 = 5; };
template <> struct VecTraits<64>  { using W = uint64_t; static constexpr int WORDS = 1; static constexpr int WBITS = 64; static constexpr int LOG2 = 6; };
template <> struct VecTraits<128> { using W = uint64_t; static constexpr int WORDS = 2; static constexpr int WBITS = 64; static constexpr int LOG2 = 7; };

__device__ __forceinline__ int popc_w(uint32_t x) { return __popc(x); }
__device__ __forceinline__ int popc_w(uint64_t x) { return __popcll(x); }

// popcount of the `rem` most significant bits of x, rem in [0, width]: x >> (width - rem) with PTX shift semantics
// (a shift by the full width gives 0, so rem == 0 needs no branch -- in C++ that shift would be undefined).
__device__ __forceinline__ uint32_t popc_top(uint32_t x, uint32_t rem) {
    uint32_t r;
    asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(32u - rem));
    return (uint32_t)__popc(r);
}
__device__ __forceinline__ uint32_t popc_top(uint64_t x, uint32_t rem) {
    unsigned long long r;
    asm("shr.u64 %0, %1, %2;" : "=l"(r) : "l"((unsigned long long)x), "r"(64u - rem));
    return (uint32_t)__popcll(r);
}

template <int NPL, int VBITS>
struct Block {
    using T = VecTraits<VBITS>;
    using W = typename T::W;
    W w[NPL][T::WORDS];

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int v = 0; v < NPL; v++)
#pragma unroll
            for (int k = 0; k < T::WORDS; k++) w[v][k] = 0;
    }

    // In-place load from the blob's `blocks` section.  Blocks are only 8-byte aligned there (an odd word count puts every
    // other block on an odd word), so 64-bit words are fetched one by one; 16-byte window loads + a conditional shift were
    // measured slower once the sweep search made neighbouring lanes share blocks (round 1: 14.17 -> 13.64 ms of
    // sweep_round_kernel time per 10^8 patterns).  Even word counts (Block2/4/6<u64>, u128 vectors) load 16 bytes at a time:
    // the blocks section is shifted onto a 32-byte boundary at load (svfm_load).
    template <class Q>
    __device__ __forceinline__ void load(const void* blocks, Q q) {
        constexpr int NW = NPL * T::WORDS;
        if constexpr (sizeof(W) == 8 && (NW & 1)) {
            const unsigned long long* base = reinterpret_cast<const unsigned long long*>(blocks) + q * (uint64_t)NW;
#pragma unroll
            for (int i = 0; i < NW; i++) {
                const W word = __ldg(base + i);
                if (T::WORDS == 1) w[i][0] = word;
                else w[i / 2][(i & 1) ? 0 : T::WORDS - 1] = word;
            }
        } else if constexpr (sizeof(W) == 8) {
            const ulonglong2* p = reinterpret_cast<const ulonglong2*>(reinterpret_cast<const W*>(blocks) + q * (uint64_t)NW);
#pragma unroll
            for (int c = 0; c < NW / 2; c++) {
                const ulonglong2 v = __ldg(p + c);
                const W pair[2] = {v.x, v.y};
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int i = 2 * c + h;
                    if (T::WORDS == 1) w[i][0] = pair[h];
                    else w[i / 2][(i & 1) ? 0 : T::WORDS - 1] = pair[h];  // u128: low half first in memory
                }
            }
        } else {
            const W* base = reinterpret_cast<const W*>(blocks) + q * (uint64_t)NW;
#pragma unroll
            for (int v = 0; v < NPL; v++) w[v][0] = ld_gather<W>(base + v);
        }
    }

    // Same block from the interleaved copy: the entry starts on a 16-byte boundary, so no window shifting.
    __device__ __forceinline__ void load_aligned(const uint8_t* entry) {
        constexpr int NW = NPL * T::WORDS;
        if constexpr (sizeof(W) == 8) {
            const ulonglong2* p = reinterpret_cast<const ulonglong2*>(entry);
            unsigned long long raw[NW + 1];
#pragma unroll
            for (int c = 0; c < NW / 2; c++) {
                const ulonglong2 v = __ldg(p + c);
                raw[2 * c] = v.x;
                raw[2 * c + 1] = v.y;
            }
            if (NW & 1) raw[NW - 1] = __ldg(reinterpret_cast<const unsigned long long*>(entry) + (NW - 1));
#pragma unroll
            for (int i = 0; i < NW; i++) {
                if (T::WORDS == 1) w[i][0] = raw[i];
                else w[i / 2][(i & 1) ? 0 : T::WORDS - 1] = raw[i];  // u128: low half first in memory
            }
        } else {
            const W* base = reinterpret_cast<const W*>(entry);
#pragma unroll
            for (int v = 0; v < NPL; v++) w[v][0] = ld_gather<W>(base + v);
        }
    }

    // AND over planes of (bit v of symidx ? plane : !plane): the `match symidx` of get_remain_count_of
    // (block2.rs:39-44, block3.rs:43-52, block4.rs:47-64, block5.rs:51-84, block6.rs:55-120).
    __device__ __forceinline__ void match_mask(uint32_t symidx, W (&m)[T::WORDS]) const {
#pragma unroll
        for (int k = 0; k < T::WORDS; k++) m[k] = ~(W)0;
#pragma unroll
        for (int v = 0; v < NPL; v++) {
            const W flip = ((symidx >> v) & 1u) ? (W)0 : ~(W)0;
#pragma unroll
            for (int k = 0; k < T::WORDS; k++) m[k] &= (w[v][k] ^ flip);
        }
    }

    // The same in two halves, so that several blocks share the per-symbol part: flip[v] = all ones where plane v must be 0.
    __device__ __forceinline__ static void flips(uint32_t symidx, W (&flip)[NPL]) {
#pragma unroll
        for (int v = 0; v < NPL; v++) flip[v] = ((symidx >> v) & 1u) ? (W)0 : ~(W)0;
    }
    __device__ __forceinline__ void match_flips(const W (&flip)[NPL], W (&m)[T::WORDS]) const {
#pragma unroll
        for (int k = 0; k < T::WORDS; k++) m[k] = w[0][k] ^ flip[0];
#pragma unroll
        for (int v = 1; v < NPL; v++) {
#pragma unroll
            for (int k = 0; k < T::WORDS; k++) m[k] &= (w[v][k] ^ flip[v]);
        }
    }

    // popcount of the first `rem` symbols of a match mask; rem in [0, VBITS) (0 -> 0).
    __device__ __forceinline__ static uint32_t prefix_count(const W (&m)[T::WORDS], uint32_t rem) {
        if (T::WORDS == 1) {
            // count_bits >>= BLOCK_LEN - rem; count_ones()   (block3.rs:53-54); rem == 0 never shifts
            return popc_top(m[0], rem);
        } else {
            uint32_t r0 = rem < 64u ? rem : 64u;  // symbols taken from word 0
            uint32_t r1 = rem - r0;               // symbols taken from word 1
            return popc_top(m[0], r0) + popc_top(m[1], r1);
        }
    }

    // get_remain_count_of(rem, symidx) (block3.rs:42-55)
    __device__ __forceinline__ uint32_t remain_count(uint32_t rem, uint32_t symidx) const {
        W m[T::WORDS];
        match_mask(symidx, m);
        return prefix_count(m, rem);
    }

    // get_symidx_of(rem) (block3.rs:57-63): bit (VBITS-1-rem) of every plane, rem in [0, VBITS)
    __device__ __forceinline__ uint32_t symidx_of(uint32_t rem) const {
        uint32_t s = 0;
        const int k = (T::WORDS == 1) ? 0 : (int)(rem >> 6);
        const uint32_t sh = T::WBITS - 1 - (rem & (T::WBITS - 1));
#pragma unroll
        for (int v = 0; v < NPL; v++) {
            W word = (T::WORDS == 1) ? w[v][0] : (k ? w[v][T::WORDS - 1] : w[v][0]);
            s |= (uint32_t)((word >> sh) & 1u) << v;
        }
        return s;
    }
};

// Block numbers: 32 bits for 32-bit positions, 64 bits otherwise (offsets derived from them are computed in 64 bits).
template <class P> struct QuotientOf { using type = uint64_t; };
template <> struct QuotientOf<uint32_t> { using type = uint32_t; };

// BwmView::get_next_rank (components/bwm/mod.rs:197-215) split in two so that the caller can issue
// the loads of several ranks before consuming any of them.
template <class P, int VBITS, class Q>
__device__ __forceinline__ void rank_addr(const DevIndex<P>& ix, P pos, Q& q, uint32_t& rem) {
    if (pos < ix.sentinel_index) pos += 1;                  // bwm/mod.rs:202-204
    q = (Q)(pos >> VecTraits<VBITS>::LOG2);                 // div_rem_with_u32(BLOCK_LEN), text_length.rs:78
    rem = (uint32_t)pos & (uint32_t)(VBITS - 1);
}

}  // namespace svfm
