// engine.cuh -- internals shared by the translation units of libsvfm.so: the index / session handles, the per-type
// kernel launchers (templates over <Position, planes, Vector bits>) and the table through which svfm_api.cu reaches
// them.  The 30 (P, BlockN, Vector) instantiations are spread over inst_p*_v*.cu so that they compile in parallel.
// Citations are relative to the reference's sview-fmindex/src/.
#pragma once
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "blob_layout.h"
#include "common.cuh"
#include "search_kernels.cuh"

struct svfm_uploader;

struct svfm_index {
    svfm_type type;
    svfm::Layout L;
    int device = 0;
    uint8_t* d_alloc = nullptr;  // cudaMalloc base
    uint8_t* d_blob = nullptr;   // d_alloc + pad: the blob, byte-for-byte
    uint64_t blob_len = 0;
    uint64_t text_len = 0;
    uint64_t sentinel_index = 0;
    uint32_t symbols_present = 0;  // symbols with at least one occurrence in the text (from count_array)
    uint8_t sym_rank[64];          // symbol index -> rank among the occurring symbols (0xff: never occurs)
    uint8_t present[64];           // rank -> symbol index
    uint32_t ext_m = 0;            // extended k-mer table: symbols resolved per lookup (0 = no table)
    uint64_t ext_entries = 0;
    void* d_ext = nullptr;         // P[2 * ext_entries]
    uint8_t* d_ilv = nullptr;      // interleaved occ copy (blocks_len entries of ilv_stride bytes), or NULL
    uint32_t ilv_stride = 0, ilv_ck_off = 0;
    uint64_t l2_window_bytes = 0;  // > 0: the extended table is small enough to sit under an L2 persisting access window
    uint32_t* d_text = nullptr;    // packed text copy (text verification), or NULL
    uint32_t text_bits = 0;
    uint64_t text_bytes = 0;
    void* d_fsa = nullptr;         // expanded suffix array (text position of every SA row), or NULL
    uint32_t fsa32 = 0;            // its entries are 32 bits wide
    uint64_t fsa_bytes = 0;
    uint8_t* d_swp = nullptr;      // sweep occ copy (blocks_len entries of 32 bytes), or NULL
    std::mutex pool_mu;
    std::vector<svfm_session*> pool;  // idle sessions for the host-buffer entry points
    std::vector<svfm_uploader*> up_pool;  // idle uploaders
};

// Upload side of the host-buffer entry points: ONE stream keeps the host->device copy engine busy with the chunks
// of a batch back to back, into one device buffer; the worker sessions wait on a per-chunk event.
struct svfm_uploader {
    cudaStream_t stream = nullptr;
    svfm::DeviceBuffer pats, offs;
    std::vector<cudaEvent_t> ev;
};

struct svfm_session {
    svfm_index* ix = nullptr;
    cudaStream_t stream = nullptr;
    // host-buffer locate calls: the results of a chunk leave on copy_stream while `stream` already runs the search of the
    // worker's next chunk; ev_done = the chunk's kernels are finished, ev_copied = its downloads are (the next chunk waits
    // for it right before its first kernel that writes out_offs / positions)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_done = nullptr, ev_copied = nullptr;
    bool copy_pending = false;
    svfm::DeviceBuffer pats, offs, unpacked, sp, cnt, counts_out, woffs, out_offs, offs64, positions, positions_alt, cub_temp;
    svfm::DeviceBuffer keys0, keys1, vals0, vals1;          // locality sort (u64 packed pattern, u32 pattern index)
    svfm::DeviceBuffer pay0, pay1, items0, items1, sweep_hist, sweep_desc;  // sweep search: items moving through the partitions
    svfm::DeviceBuffer rec_key, rec_key_alt, first;          // radix sort-back of (pattern index -> position) records (SVFM_SORTED)
    svfm::DeviceBuffer sb_hist, sb_base, sb_cursor, sb_recs; // bucketed sort-back
    svfm::DeviceBuffer resolved;                             // text verification: per work item, sp holds the position
    svfm::DeviceBuffer heavy_sp, heavy_cnt, heavy_obase, heavy_pat, heavy_offs;
    unsigned long long* d_counters = nullptr;                // [0] heavy patterns seen by search, [1] heavy list length,
                                                             // [2] SA rows of the batch (bucketed sort-back)
    int* d_err = nullptr;
    uint64_t* h_pinned = nullptr;  // [0] = total, [1] = err bits
    uint64_t reserve_n = 0;        // host-buffer calls: size the scratch for at least this many patterns (the call's largest chunk),
                                   // so that every worker session grows its buffers once instead of whenever it meets a larger chunk
    uint8_t* h_small = nullptr;    // small-batch path: mapped pinned arena (patterns | offsets | counts | slots | status)
    // per-phase timing (svfm_session_set_timing)
    bool timing = false;
    struct Span { int phase; cudaEvent_t e0, e1; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> free_events;
    double phase_ms[SVFM_PHASE_MAX] = {0};
    uint64_t phase_launches[SVFM_PHASE_MAX] = {0};
};

namespace svfm {

static inline uint64_t rsv(const svfm_session* s, uint64_t n) { return n > s->reserve_n ? n : s->reserve_n; }

constexpr int MAX_DYNAMIC_SMEM = 200 * 1024;  // opt-in dynamic shared memory per CTA (B200: up to 227 KB)

template <class P>
static DevIndex<P> make_dev_index(const svfm_index* ix) {
    const Layout& L = ix->L;
    DevIndex<P> d;
    d.count_array = reinterpret_cast<const P*>(ix->d_blob + L.off_count_array);
    d.kmer_multiplier = reinterpret_cast<const uint64_t*>(ix->d_blob + L.off_kmer_multiplier);
    d.kmer_count_table = reinterpret_cast<const P*>(ix->d_blob + L.off_kmer_count_table);
    d.suffix_array = reinterpret_cast<const P*>(ix->d_blob + L.off_suffix_array);
    d.rank_checkpoints = reinterpret_cast<const P*>(ix->d_blob + L.off_rank_checkpoints);
    d.blocks = ix->d_blob + L.off_blocks;
    d.table = ix->type.encoder ? ix->d_blob + L.off_encoder : nullptr;
    d.sentinel_index = (P)ix->sentinel_index;
    d.symbol_count = L.bwm_symbol_count;
    d.kmer_size = L.kmer_size;
    d.sampling_ratio = L.sampling_ratio;
    const uint32_t r = L.sampling_ratio;
    if ((r & (r - 1)) == 0) {
        d.ratio_mask = r - 1;
        d.ratio_shift = 0;
        while ((1u << d.ratio_shift) < r) d.ratio_shift++;
    } else {
        d.ratio_mask = 0xffffffffu;
        d.ratio_shift = 0;
    }
    d.ext = reinterpret_cast<const P*>(ix->d_ext);
    d.ext_m = ix->ext_m;
    d.s_eff = ix->symbols_present;
    std::memcpy(d.sym_rank, ix->sym_rank, 64);
    std::memcpy(d.present, ix->present, 64);
    d.ilv = ix->d_ilv;
    d.ilv_stride = ix->ilv_stride;
    d.ilv_ck_off = ix->ilv_ck_off;
    d.text = ix->d_text;
    d.text_bits = ix->text_bits;
    d.fsa = ix->d_fsa;
    d.fsa32 = ix->fsa32;
    d.swp = ix->d_swp;
    return d;
}


// RAII span: records CUDA events around the kernels of one phase when timing is on.
struct PhaseTimer {
    svfm_session* s;
    int phase;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    PhaseTimer(svfm_session* s_, int phase_, uint64_t launches) : s(s_), phase(phase_) {
        s->phase_launches[phase] += launches;
        g_launches += launches;
        if (!s->timing) return;
        auto get = [&]() {
            cudaEvent_t e;
            if (!s->free_events.empty()) { e = s->free_events.back(); s->free_events.pop_back(); }
            else cudaEventCreate(&e);
            return e;
        };
        e0 = get();
        e1 = get();
        cudaEventRecord(e0, s->stream);
    }
    ~PhaseTimer() {
        if (!e0) return;
        cudaEventRecord(e1, s->stream);
        s->spans.push_back({phase, e0, e1});
    }
};

// Grid for a grid-stride kernel: a multiple of the CTAs that can be resident at once (SMs x occupancy);
// fewer when the work does not fill the machine.  Measured on B200 (1 Gbp index, 100M patterns): 4 resident
// waves run the search kernel 13% faster than exactly one (shorter per-thread item chains, less tail).
template <class K>
static int resident_grid(K kernel, uint64_t work_items, int threads, int device) {
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    uint64_t blocks = (work_items + threads - 1) / threads;
    static const double mult = [] { const char* e = std::getenv("SVFM_GRID_MULT"); return e ? atof(e) : 4.0; }();
    const uint64_t cap = (uint64_t)((double)sms * per_sm * mult);
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    return (int)blocks;
}

static int grid_for(uint64_t work_items, int threads, int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    uint64_t blocks = (work_items + threads - 1) / threads;
    const uint64_t cap = (uint64_t)sms * 8;  // a whole number of waves of resident CTAs; grid-stride beyond
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    return (int)blocks;
}


// Batch plans.  Everything that has to change order moves through radix sorts (streaming passes), never through
// random scatters: on B200 one random 32 B sector access costs as much HBM time as streaming ~120 bytes.
//   sweep  : dense fixed-length batch -> items sorted by SA interval, rounds of backward steps + radix partition
//            (search_kernels.cuh, "sweep search")
//   sorted : locality sort by trailing symbols, one search kernel (variable-length or long patterns)
//   neither: search kernel in the caller's order (small batches)
struct SortPlan {
    bool sorted = false;
    uint32_t bits = 0;     // bits per symbol in the packed key
    int begin_bit = 0, end_bit = 64;
    bool sweep = false;
    uint32_t m = 0;            // sweep: trailing symbols resolved by the extended table
    int prefix_bits = 0;       // sweep: significant bits of the table index
    uint32_t steps_per_round = 1;
    bool rest64 = false;       // sweep: the other symbols need a 64-bit word
};


static int bits_for(uint64_t n) {  // smallest b with 2^b >= n
    int b = 0;
    while (b < 63 && (1ull << b) < n) b++;
    return b;
}


constexpr uint32_t SEARCH_STAGE_MAX = 44 * 1024;

// One launch of the search kernel: keys/idx (or NULL) in, sp/cnt out.
template <class P, int NPL, int VBITS>
static int run_search(svfm_session* s, const PatternBatch& pb, const uint64_t* keys, const uint32_t* idx, uint32_t bits,
                      void* d_sp_work, void* d_cnt_work, const SbOut& sb, uint8_t* d_resolved) {
    DevIndex<P> dix = make_dev_index<P>(s->ix);
    if (!d_resolved && d_sp_work) dix.text = nullptr;   // locate without a flag array: rows only (count may always verify)
    const bool ilv = dix.ilv != nullptr;
    // dynamic shared memory for the staged pattern bytes of one CTA (search_kernels.cuh): up to ~44 KB, i.e. patterns of up
    // to 176 bytes in a fixed-length batch; less when the patterns are short (occupancy)
    uint32_t stage = 0;
    if (!idx) {
        const uint64_t want = pb.offs ? (uint64_t)SEARCH_STAGE_MAX : (uint64_t)SEARCH_THREADS * pb.fixed_len + 4;
        stage = (uint32_t)(want < (uint64_t)SEARCH_STAGE_MAX ? ((want + 15) & ~(uint64_t)15) : (uint64_t)SEARCH_STAGE_MAX);
    }
    SearchIO<P> io{};
    io.keys = keys;
    io.idx = idx;
    io.bits = bits;
    io.sp_out = (P*)d_sp_work;
    io.cnt_out = (P*)d_cnt_work;
    io.heavy_seen = s->d_counters;
    io.err = s->d_err;
    io.sb = sb;
    io.resolved_out = d_resolved;
    auto launch = [&](auto kernel) -> int {
        int sms = 148, per_sm = 1;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->ix->device);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SEARCH_THREADS, stage) != cudaSuccess || per_sm < 1) per_sm = 1;
        uint64_t blocks = (pb.n + SEARCH_THREADS - 1) / SEARCH_THREADS;
        const uint64_t cap = (uint64_t)sms * per_sm * 4;   // 4 resident waves (measured in round 1: 13 % faster than one)
        if (blocks > cap) blocks = cap;
        if (blocks == 0) blocks = 1;
        kernel<<<(unsigned)blocks, SEARCH_THREADS, stage, s->stream>>>(dix, pb, io, stage);
        SVFM_CUDA(cudaGetLastError());
        return SVFM_OK;
    };
    PhaseTimer pt(s, SVFM_PHASE_SEARCH, 1);
    return ilv ? launch(search_kernel<P, NPL, VBITS, true>) : launch(search_kernel<P, NPL, VBITS, false>);
}

// Small batches: one launch, inputs and outputs in mapped pinned host memory (search_kernels.cuh, small_batch_kernel).
constexpr uint64_t SMALL_MAX_PATTERNS = 4096;
constexpr uint64_t SMALL_MAX_BYTES = 256 << 10;
constexpr uint32_t SMALL_SLOTS = 8;
template <class P, int NPL, int VBITS>
static int run_small(svfm_session* s, const PatternBatch& pb, const SmallOut& out) {
    const DevIndex<P> dix = make_dev_index<P>(s->ix);
    const unsigned grid = (unsigned)((pb.n + 127) / 128);
    PhaseTimer pt(s, SVFM_PHASE_SEARCH, 1);
    if (dix.ilv) small_batch_kernel<P, NPL, VBITS, true><<<grid, 128, 0, s->stream>>>(dix, pb, out);
    else small_batch_kernel<P, NPL, VBITS, false><<<grid, 128, 0, s->stream>>>(dix, pb, out);
    SVFM_CUDA(cudaGetLastError());
    return SVFM_OK;
}

// Expanded suffix array (search_kernels.cuh), built once per index after the interleaved occ copy and before the text copy
// (whose construction then reads it instead of walking every row again).  Optional: skipped when it would take more than
// 1/8 of the device memory still free.
extern std::atomic<uint64_t> g_full_sa, g_text;   // svfm_api.cu (SVFM_TUNE_FULL_SA, SVFM_TUNE_TEXT)
template <class P, int NPL, int VBITS>
static int run_build_fsa(svfm_index* ix) {
    const uint64_t n = ix->text_len;
    if (!g_full_sa.load() || n == 0 || ix->L.sampling_ratio <= 1) return SVFM_OK;   // ratio 1: the blob's array is complete already
    const bool e32 = sizeof(P) == 4 || n < 0xffffffffull;
    const uint64_t bytes = (n + 1) * (e32 ? 4 : 8);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || bytes > free_b / 8) { (void)cudaGetLastError(); return SVFM_OK; }
    void* d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) { (void)cudaGetLastError(); return SVFM_OK; }   // optional structure
    const DevIndex<P> dix = make_dev_index<P>(ix);   // ix->d_fsa is still NULL: the walk uses the sampled array
    const int grid = grid_for(n, 256, ix->device) * 4;
    if (e32) {
        if (dix.ilv) fsa_build_kernel<P, NPL, VBITS, true, uint32_t><<<grid, 256>>>(dix, n, (uint32_t*)d);
        else fsa_build_kernel<P, NPL, VBITS, false, uint32_t><<<grid, 256>>>(dix, n, (uint32_t*)d);
    } else {
        if (dix.ilv) fsa_build_kernel<P, NPL, VBITS, true, unsigned long long><<<grid, 256>>>(dix, n, (unsigned long long*)d);
        else fsa_build_kernel<P, NPL, VBITS, false, unsigned long long><<<grid, 256>>>(dix, n, (unsigned long long*)d);
    }
    g_launches++;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(d); SVFM_CUDA(e); }
    ix->d_fsa = d;
    ix->fsa32 = e32 ? 1 : 0;
    ix->fsa_bytes = bytes;
    return SVFM_OK;
}

// Packed text copy for the text verification (search_kernels.cuh), built once per index after the expanded suffix array.
template <class P, int NPL, int VBITS>
static int run_build_text(svfm_index* ix) {
    int rc = run_build_fsa<P, NPL, VBITS>(ix);
    if (rc) return rc;
    if (!g_text.load()) return SVFM_OK;
    const uint64_t s_eff = ix->symbols_present;
    const uint64_t n = ix->text_len;
    if (s_eff < 1 || n == 0) return SVFM_OK;
    const uint32_t bits = s_eff <= 2 ? 1 : s_eff <= 4 ? 2 : s_eff <= 16 ? 4 : 8;
    const uint64_t bytes = ((n * bits + 31) / 32) * 4 + 16;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || bytes > free_b / 8) { (void)cudaGetLastError(); return SVFM_OK; }
    uint32_t* d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) { (void)cudaGetLastError(); return SVFM_OK; }   // optional structure
    cudaError_t e = cudaMemset(d, 0, bytes);
    if (e == cudaSuccess) {
        const DevIndex<P> dix = make_dev_index<P>(ix);
        const int grid = grid_for(n, 256, ix->device) * 4;
        if (dix.ilv) text_build_kernel<P, NPL, VBITS, true><<<grid, 256>>>(dix, n, bits, d);
        else text_build_kernel<P, NPL, VBITS, false><<<grid, 256>>>(dix, n, bits, d);
        g_launches++;
        e = cudaDeviceSynchronize();
    }
    if (e != cudaSuccess) { cudaFree(d); SVFM_CUDA(e); }
    ix->d_text = d;
    ix->text_bits = bits;
    ix->text_bytes = bytes;
    return SVFM_OK;
}

// Extended k-mer table, built once per index right after the upload (search_kernels.cuh).
template <class P, int NPL, int VBITS>
static int run_build_ext(svfm_index* ix, uint64_t ext_bits) {
    const uint64_t s_eff = ix->symbols_present;
    uint64_t budget = 1ull << (ext_bits > 32 ? 32 : ext_bits);
    if (ext_bits == 0 || s_eff < 2 || s_eff > 64) return SVFM_OK;
    // at most two entries per text symbol (an interval of 0.5 rows on average: most lookups of patterns that occur end on a
    // single row, and longer keys would only add empty entries); SVFM_EXT_PER_TEXT overrides the factor
    static const double per_text = [] { const char* e = std::getenv("SVFM_EXT_PER_TEXT"); return e ? atof(e) : 2.0; }();
    const uint64_t cap_text = (uint64_t)((double)ix->text_len * per_text);
    const uint64_t by_text = cap_text > s_eff ? cap_text : s_eff;
    if (budget > by_text) budget = by_text;
    uint32_t m = 1;
    uint64_t entries = s_eff;
    while (entries * s_eff <= budget && m < 31) { entries *= s_eff; m++; }
    if (entries > 0xfffffff0ull) return SVFM_OK;
    DevIndex<P> dix = make_dev_index<P>(ix);
    P *a = nullptr, *b = nullptr;
    // The table is optional (the search kernels fall back to the blob's own k-mer table): when the allocation fails --
    // another load on the same GPU may have taken the memory since cudaMemGetInfo -- retry with 2^24 entries, then go without.
    for (;;) {
        cudaError_t e = cudaMalloc(&a, entries * 2 * sizeof(P));
        if (e == cudaSuccess && m > 1) {
            e = cudaMalloc(&b, entries / s_eff * 2 * sizeof(P));
            if (e != cudaSuccess) { cudaFree(a); a = nullptr; }
        }
        if (e == cudaSuccess) break;
        (void)cudaGetLastError();
        if (e != cudaErrorMemoryAllocation) SVFM_CUDA(e);
        if (entries <= (1ull << 24)) return SVFM_OK;  // no table
        while (entries > (1ull << 24) && m > 1) { entries /= s_eff; m--; }
    }
    // levels alternate between the two buffers so that level m lands in `a`
    P* cur = (m % 2 == 1) ? a : b;
    P* other = (m % 2 == 1) ? b : a;
    ext_level1_kernel<P><<<1, 64>>>(dix, cur);
    g_launches++;
    uint64_t n_in = s_eff;
    for (uint32_t j = 1; j < m; j++) {
        const int grid = resident_grid(ext_expand_kernel<P, NPL, VBITS>, n_in * s_eff, SEARCH_THREADS, ix->device);
        ext_expand_kernel<P, NPL, VBITS><<<grid, SEARCH_THREADS>>>(dix, cur, n_in, other);
        g_launches++;
        std::swap(cur, other);
        n_in *= s_eff;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (b) cudaFree(b);
    if (e != cudaSuccess) { cudaFree(a); SVFM_CUDA(e); }
    ix->d_ext = a;
    ix->ext_m = m;
    ix->ext_entries = entries;
    return SVFM_OK;
}

// Output of the type-independent front of the sweep search (svfm_api.cu): items packed and radix-sorted by table index,
// digit histograms of every round followed by one zeroed tile counter per round.
constexpr int SWEEP_INDEX_BITS = 6;  // PART_INDEX: the last round of `count` groups its items by the top 6 bits of the pattern index
struct SweepPre {
    const uint32_t* prefix;
    const void* pay;   // SweepPay<R>[n]
    uint32_t* hist;    // [rounds][1 << (bits * steps_per_round)] digit histograms of the rounds
    uint32_t* counters; // [rounds] zeroed tile counters, then 64 zeroed bin cursors (PART_INDEX)
};
template <class R>
int run_sweep_presort(svfm_session* s, const PatternBatch& pb, const SortPlan& plan, uint32_t rounds, SweepPre* out);
extern template int run_sweep_presort<uint32_t>(svfm_session*, const PatternBatch&, const SortPlan&, uint32_t, SweepPre*);
extern template int run_sweep_presort<uint64_t>(svfm_session*, const PatternBatch&, const SortPlan&, uint32_t, SweepPre*);

// Sweep search (search_kernels.cuh): pack -> radix sort by table index -> rounds of [seed/resume + T backward steps +
// stable radix partition by the consumed symbols], each round ONE kernel.  Leaves sp/cnt/idx of every item in work order.
// final_mode: what the LAST round does with its items -- PART_NONE: written in place; PART_SYMBOLS (locate): partitioned
// once more, so that the LF walks and SA reads run in SA order; PART_INDEX (count): grouped by the top bits of the
// caller's pattern index, ready for scatter_counts_kernel.  sb (sb.hist nullable): the last round also reserves the record
// slots of the bucketed sort-back.
template <class P, int NPL, int VBITS, class R>
static int run_search_sweep_r(svfm_session* s, const PatternBatch& pb, const SortPlan& plan, int final_mode,
                              void* d_sp_work, void* d_cnt_work, const uint32_t** idx_out, const SbOut& sb, uint8_t* d_resolved) {
    const svfm_index* ix = s->ix;
    const DevIndex<P> dix = make_dev_index<P>(ix);
    const uint64_t n = pb.n;
    const uint32_t len = pb.fixed_len, m = plan.m, bits = plan.bits, T = plan.steps_per_round;
    const uint32_t remaining = len - m;
    const uint32_t rounds = remaining ? (remaining + T - 1) / T : 1;
    const uint32_t digit_bits = bits * T, nb_max = 1u << digit_bits;
    using Pay = SweepPay<R>;
    using Item = SweepItem<P, R>;
    const uint64_t n_tiles = (n + ROUND_TILE - 1) / ROUND_TILE;
    int rc;
    const uint64_t rn = rsv(s, n);
    if ((rc = s->vals0.reserve(rn * 4)) || (rc = s->sweep_desc.reserve((rn + ROUND_TILE - 1) / ROUND_TILE * nb_max * 4))) return rc;
    const bool partitions = rounds > 1;
    if (partitions && ((rc = s->items0.reserve(rn * sizeof(Item))) || (rc = s->items1.reserve(rn * sizeof(Item))))) return rc;
    Item* items[2] = {(Item*)s->items0.ptr, (Item*)s->items1.ptr};
    SweepPre pre{};
    if ((rc = run_sweep_presort<R>(s, pb, plan, rounds, &pre))) return rc;  // pack + radix sort by table index (svfm_api.cu)
    const uint32_t* prefix = pre.prefix;
    const Pay* pay = (const Pay*)pre.pay;
    uint32_t* hist = pre.hist;
    uint32_t* tile_counters = pre.counters;                     // one per round
    uint32_t* bin_cursor = tile_counters + rounds;               // SWEEP_INDEX_BINS counters (PART_INDEX), zeroed with the rest
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix->device);
    uint32_t* idx_work = (uint32_t*)s->vals0.ptr;
    int cur = 0;  // items[cur] holds the current state from round 1 on
    // locate mode with a text copy: the last round can resolve instead of stepping + partitioning (sweep_resolve_kernel).
    // OFF by default -- measured on B200, 10^8 20-mers: resolve 5.96 ms + locate 2.94 ms against last round 2.85 ms + locate
    // 4.88 ms: the text check adds one random sector per item (1.8 ms at the random-access rate), more than the skipped
    // backward steps and partition cost once the batch is SA-ordered.  SVFM_SWEEP_RESOLVE=1 turns it on.
    static const bool resolve_env = [] { const char* e = std::getenv("SVFM_SWEEP_RESOLVE"); return e ? atoi(e) != 0 : false; }();
    const bool resolve_last = resolve_env && final_mode == PART_SYMBOLS && rounds >= 2 && dix.text != nullptr && d_resolved != nullptr;
    for (uint32_t r = 0; r < rounds; r++) {
        const uint32_t first = r * T;
        const uint32_t steps = remaining - first < T ? remaining - first : T;
        const bool last = r + 1 == rounds;
        if (last && resolve_last) {
            PhaseTimer pt(s, SVFM_PHASE_SEARCH, 1);
            auto launch = [&](auto kernel) -> int {
                const int grid = resident_grid(kernel, n, 256, ix->device);
                kernel<<<grid, 256, 0, s->stream>>>(dix, items[cur], n, bits, bits * first, steps, (P*)d_sp_work, (P*)d_cnt_work, idx_work,
                                                    d_resolved, s->d_counters, sb);
                SVFM_CUDA(cudaGetLastError());
                return SVFM_OK;
            };
            rc = dix.ilv ? launch(sweep_resolve_kernel<P, NPL, VBITS, R, true>) : launch(sweep_resolve_kernel<P, NPL, VBITS, R, false>);
            if (rc) return rc;
            break;
        }
        int part = PART_SYMBOLS;
        if (last) part = final_mode == PART_INDEX ? PART_INDEX : (final_mode == PART_SYMBOLS && steps > 0 ? PART_SYMBOLS : PART_NONE);
        // the last partition digit may be narrower than digit_bits: the histogram was taken on digit_bits bits, whose
        // upper bits are then zero, so the wide digit sorts identically
        uint32_t nbins = nb_max;
        SweepRoundIO<P, R> io{};
        if (r == 0) { io.prefix = prefix; io.pay = pay; }
        else io.items_in = items[cur];
        if (last) {
            io.sp_out = (P*)d_sp_work;
            io.cnt_out = (P*)d_cnt_work;
            io.idx_out = idx_work;
            io.heavy_seen = s->d_counters;
            io.sb = sb;
        } else {
            io.items_out = items[r == 0 ? 0 : cur ^ 1];
        }
        io.hist = hist + (uint64_t)r * nb_max;
        io.desc = (uint32_t*)s->sweep_desc.ptr;
        io.tile_counter = tile_counters + r;
        if (part == PART_INDEX) {
            const int nbits = bits_for(n);
            io.idx_shift = nbits > SWEEP_INDEX_BITS ? (uint32_t)(nbits - SWEEP_INDEX_BITS) : 0u;
            io.bin_cursor = bin_cursor;
            nbins = 1u << SWEEP_INDEX_BITS;
        }
        static const bool arrival_order = [] { const char* e = std::getenv("SVFM_ROUND_ATOMIC"); return e ? atoi(e) != 0 : false; }();
        PhaseTimer pt(s, SVFM_PHASE_SEARCH, 1);
        if (part == PART_SYMBOLS && arrival_order) {
            io.bin_cursor = io.desc;  // nbins counters instead of tiles x nbins look-back descriptors
            SVFM_CUDA(cudaMemsetAsync(io.desc, 0, (size_t)nbins * 4, s->stream));
        } else if (part == PART_SYMBOLS) {
            SVFM_CUDA(cudaMemsetAsync(io.desc, 0, n_tiles * nbins * 4, s->stream));
        }
        const size_t smem = part != PART_NONE ? sizeof(Item) * ROUND_TILE + (size_t)nbins * 16 + ((size_t)ROUND_WARPS * nbins + nbins) * 4 +
                                                    0 : 0;
        auto launch = [&](auto kernel) -> int {
            // always the same value: concurrent sessions launch this kernel with different sizes, and the attribute is
            // per function, not per launch
            SVFM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYNAMIC_SMEM));
            if (smem > (size_t)MAX_DYNAMIC_SMEM) return SVFM_ERR_TOO_LARGE;
            int per_sm = 1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ROUND_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
            uint64_t grid = n_tiles < (uint64_t)sms * per_sm ? n_tiles : (uint64_t)sms * per_sm;
            kernel<<<(unsigned)grid, ROUND_THREADS, smem, s->stream>>>(dix, n, bits, bits * first, steps, nbins, io);
            SVFM_CUDA(cudaGetLastError());
            return SVFM_OK;
        };
        constexpr bool SWP_OK = VBITS == 64 && NPL <= 3;   // the sweep occ copy exists for these block shapes only
        bool launched = false;
        if constexpr (SWP_OK) {
            if (dix.swp) {
                launched = true;
                if (r == 0) {
                    if (part == PART_SYMBOLS) rc = launch(sweep_round_kernel<P, NPL, VBITS, R, true, PART_SYMBOLS, true>);
                    else if (part == PART_INDEX) rc = launch(sweep_round_kernel<P, NPL, VBITS, R, true, PART_INDEX, true>);
                    else rc = launch(sweep_round_kernel<P, NPL, VBITS, R, true, PART_NONE, true>);
                } else {
                    if (part == PART_SYMBOLS) rc = launch(sweep_round_kernel<P, NPL, VBITS, R, false, PART_SYMBOLS, true>);
                    else if (part == PART_INDEX) rc = launch(sweep_round_kernel<P, NPL, VBITS, R, false, PART_INDEX, true>);
                    else rc = launch(sweep_round_kernel<P, NPL, VBITS, R, false, PART_NONE, true>);
                }
            }
        }
        if (launched) {
        } else if (r == 0) {
            if (part == PART_SYMBOLS) rc = launch(sweep_round_kernel<P, NPL, VBITS, R, true, PART_SYMBOLS>);
            else if (part == PART_INDEX) rc = launch(sweep_round_kernel<P, NPL, VBITS, R, true, PART_INDEX>);
            else rc = launch(sweep_round_kernel<P, NPL, VBITS, R, true, PART_NONE>);
        } else {
            if (part == PART_SYMBOLS) rc = launch(sweep_round_kernel<P, NPL, VBITS, R, false, PART_SYMBOLS>);
            else if (part == PART_INDEX) rc = launch(sweep_round_kernel<P, NPL, VBITS, R, false, PART_INDEX>);
            else rc = launch(sweep_round_kernel<P, NPL, VBITS, R, false, PART_NONE>);
        }
        if (rc) return rc;
        if (r > 0 && !last) cur ^= 1;
    }
    *idx_out = idx_work;
    return SVFM_OK;
}

template <class P, int NPL, int VBITS>
static int run_search_sweep(svfm_session* s, const PatternBatch& pb, const SortPlan& plan, int final_mode, void* d_sp_work,
                            void* d_cnt_work, const uint32_t** idx_out, const SbOut& sb, uint8_t* d_resolved) {
    if (plan.rest64) return run_search_sweep_r<P, NPL, VBITS, uint64_t>(s, pb, plan, final_mode, d_sp_work, d_cnt_work, idx_out, sb, d_resolved);
    return run_search_sweep_r<P, NPL, VBITS, uint32_t>(s, pb, plan, final_mode, d_sp_work, d_cnt_work, idx_out, sb, d_resolved);
}

template <class P>
struct WidenCounts {
    const P* cnt;
    uint64_t n;
    __host__ __device__ uint64_t operator()(uint64_t i) const { return i < n ? (uint64_t)cnt[i] : 0ull; }
};

template <class P>
static int run_scan(svfm_session* s, uint64_t n, const void* d_cnt, uint64_t* d_out_offs) {
    // out_offs[0..n] = exclusive prefix sums of the counts, widened to u64 (out_offs[n] = total)
    WidenCounts<P> f{(const P*)d_cnt, n};
    auto in = thrust::make_transform_iterator(thrust::counting_iterator<uint64_t>(0), f);
    size_t temp = 0;
    SVFM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, temp, in, d_out_offs, n + 1, s->stream));
    int rc = s->cub_temp.reserve(temp);
    if (rc) return rc;
    PhaseTimer pt(s, SVFM_PHASE_SCAN, 2);
    SVFM_CUDA(cub::DeviceScan::ExclusiveSum(s->cub_temp.ptr, temp, in, d_out_offs, n + 1, s->stream));
    return SVFM_OK;
}


// LF-walk + sampled-SA lookup for every SA row of every pattern (SVFM_PHASE_LOCATE).
// d_recs != NULL: bucket mode (search_kernels.cuh, "bucketed sort-back") -- records go to d_recs through d_cursor and
// d_offs / d_positions / d_rec_key are unused.
template <class P, int NPL, int VBITS>
static int run_locate(svfm_session* s, uint64_t n, const uint32_t* idx, const void* d_sp_work, const void* d_cnt_work,
                      const uint64_t* d_offs, uint64_t total, uint64_t heavy_seen, void* d_positions, uint32_t* d_rec_key,
                      void* d_recs, unsigned long long* d_cursor, const uint8_t* d_resolved) {
    if (total == 0) return SVFM_OK;
    const DevIndex<P> dix = make_dev_index<P>(s->ix);
    HeavyList<P> heavy{nullptr, nullptr, nullptr, nullptr, s->d_counters + 1, 0};
    BucketOut<P> bk{(SbRec<P>*)d_recs, d_cursor};
    const bool bucket = d_recs != nullptr;
    int rc;
    if (heavy_seen) {
        if ((rc = s->heavy_sp.reserve(heavy_seen * sizeof(P))) || (rc = s->heavy_cnt.reserve((heavy_seen + 1) * sizeof(P))) ||
            (rc = s->heavy_obase.reserve(heavy_seen * sizeof(uint64_t))) || (rc = s->heavy_pat.reserve(heavy_seen * 4)) ||
            (rc = s->heavy_offs.reserve((heavy_seen + 1) * sizeof(uint64_t))))
            return rc;
        heavy.sp = (P*)s->heavy_sp.ptr;
        heavy.cnt = (P*)s->heavy_cnt.ptr;
        heavy.obase = (uint64_t*)s->heavy_obase.ptr;
        heavy.pat = (uint32_t*)s->heavy_pat.ptr;
        heavy.capacity = heavy_seen;
    }
    if (dix.fsa) {   // expanded suffix array: one read per row, no walk (search_kernels.cuh, locate_direct_kernel)
        PhaseTimer pt(s, SVFM_PHASE_LOCATE, 1);
        auto launch = [&](auto kernel) {
            // many short CTAs rather than a few resident waves: measured 2.48 / 2.27 / 2.19 ms per 10^8 rows at 2 / 4 / 8 waves
            const uint64_t tiles = (n + (uint64_t)LOCATE_THREADS * LOCATE_DIRECT_ITEMS - 1) / ((uint64_t)LOCATE_THREADS * LOCATE_DIRECT_ITEMS);
            const unsigned grid = (unsigned)(tiles < (1ull << 20) ? (tiles ? tiles : 1) : (1ull << 20));
            kernel<<<grid, LOCATE_THREADS, 0, s->stream>>>(dix, idx, (const P*)d_sp_work, (const P*)d_cnt_work, d_offs, n,
                                                           (P*)d_positions, d_rec_key, heavy, bk, d_resolved);
        };
        if (bucket) launch(locate_direct_kernel<P, true>);
        else launch(locate_direct_kernel<P, false>);
        SVFM_CUDA(cudaGetLastError());
    } else {
        PhaseTimer pt(s, SVFM_PHASE_LOCATE, 1);
        auto launch = [&](auto kernel) {
            const int grid = resident_grid(kernel, n, LOCATE_THREADS, s->ix->device);
            kernel<<<grid, LOCATE_THREADS, 0, s->stream>>>(dix, idx, (const P*)d_sp_work, (const P*)d_cnt_work, d_offs, n,
                                                           (P*)d_positions, d_rec_key, heavy, bk, d_resolved);
        };
        if (dix.ilv) {
            if (bucket) launch(locate_warp_kernel<P, NPL, VBITS, true, true>);
            else launch(locate_warp_kernel<P, NPL, VBITS, true, false>);
        } else {
            if (bucket) launch(locate_warp_kernel<P, NPL, VBITS, false, true>);
            else launch(locate_warp_kernel<P, NPL, VBITS, false, false>);
        }
        SVFM_CUDA(cudaGetLastError());
    }
    if (!heavy_seen) return SVFM_OK;
    // patterns with more than HEAVY_ROWS rows: one thread per row
    if ((rc = run_scan<P>(s, heavy_seen, heavy.cnt, (uint64_t*)s->heavy_offs.ptr))) return rc;
    SVFM_CUDA(cudaMemcpyAsync(&s->h_pinned[2], (uint64_t*)s->heavy_offs.ptr + heavy_seen, sizeof(uint64_t),
                              cudaMemcpyDeviceToHost, s->stream));
    SVFM_CUDA(cudaStreamSynchronize(s->stream));
    const uint64_t heavy_total = s->h_pinned[2];
    const uint64_t blocks = (heavy_total + LOCATE_THREADS - 1) / LOCATE_THREADS;
    if (blocks > 0x7fffffffull) return SVFM_ERR_TOO_LARGE;
    if (blocks) {
        PhaseTimer pt(s, SVFM_PHASE_LOCATE, 1);
        auto launch = [&](auto kernel) {
            kernel<<<(unsigned)blocks, LOCATE_THREADS, 0, s->stream>>>(dix, heavy.sp, (const uint64_t*)s->heavy_offs.ptr, heavy.obase,
                                                                      heavy.pat, heavy_seen, heavy_total, (P*)d_positions, d_rec_key,
                                                                      (SbRec<P>*)d_recs);
        };
        if (dix.ilv) {
            if (bucket) launch(locate_rows_kernel<P, NPL, VBITS, true, true>);
            else launch(locate_rows_kernel<P, NPL, VBITS, true, false>);
        } else {
            if (bucket) launch(locate_rows_kernel<P, NPL, VBITS, false, true>);
            else launch(locate_rows_kernel<P, NPL, VBITS, false, false>);
        }
        SVFM_CUDA(cudaGetLastError());
    }
    return SVFM_OK;
}


// ---- per-(Position, Vector) entry points; `planes` picks Block2..Block6 ------------------------------------------
struct TypeOps {
    int (*search)(uint32_t planes, svfm_session* s, const PatternBatch& pb, const uint64_t* keys, const uint32_t* idx, uint32_t bits,
                  void* d_sp_work, void* d_cnt_work, const SbOut& sb, uint8_t* d_resolved);
    int (*build_ext)(uint32_t planes, svfm_index* ix, uint64_t ext_bits);
    int (*build_text)(uint32_t planes, svfm_index* ix);
    int (*search_sweep)(uint32_t planes, svfm_session* s, const PatternBatch& pb, const SortPlan& plan, int final_mode,
                        void* d_sp_work, void* d_cnt_work, const uint32_t** idx_out, const SbOut& sb, uint8_t* d_resolved);
    int (*small)(uint32_t planes, svfm_session* s, const PatternBatch& pb, const SmallOut& out);
    int (*locate)(uint32_t planes, svfm_session* s, uint64_t n, const uint32_t* idx, const void* d_sp_work, const void* d_cnt_work,
                  const uint64_t* d_offs, uint64_t total, uint64_t heavy_seen, void* d_positions, uint32_t* d_rec_key,
                  void* d_recs, unsigned long long* d_cursor, const uint8_t* d_resolved);
};

#define SVFM_PLANES_SWITCH(P, VB, FN, ...)          \
    switch (planes) {                               \
        case 2: return FN<P, 2, VB>(__VA_ARGS__);   \
        case 3: return FN<P, 3, VB>(__VA_ARGS__);   \
        case 4: return FN<P, 4, VB>(__VA_ARGS__);   \
        case 5: return FN<P, 5, VB>(__VA_ARGS__);   \
        case 6: return FN<P, 6, VB>(__VA_ARGS__);   \
        default: return SVFM_ERR_BAD_TYPE;          \
    }

// One translation unit per (P, VB): defines `const TypeOps NAME`.
#define SVFM_DEFINE_TYPE_OPS(NAME, P, VB)                                                                                          \
    static int NAME##_search(uint32_t planes, svfm_session* s, const PatternBatch& pb, const uint64_t* keys, const uint32_t* idx,   \
                             uint32_t bits, void* d_sp_work, void* d_cnt_work, const SbOut& sb, uint8_t* d_resolved) {  \
        SVFM_PLANES_SWITCH(P, VB, run_search, s, pb, keys, idx, bits, d_sp_work, d_cnt_work, sb, d_resolved)                   \
    }                                                                                                                               \
    static int NAME##_build_ext(uint32_t planes, svfm_index* ix, uint64_t ext_bits) {                                               \
        SVFM_PLANES_SWITCH(P, VB, run_build_ext, ix, ext_bits)                                                                      \
    }                                                                                                                               \
    static int NAME##_build_text(uint32_t planes, svfm_index* ix) {                                                                 \
        SVFM_PLANES_SWITCH(P, VB, run_build_text, ix)                                                                               \
    }                                                                                                                               \
    static int NAME##_search_sweep(uint32_t planes, svfm_session* s, const PatternBatch& pb, const SortPlan& plan, int final_mode,  \
                                   void* d_sp_work, void* d_cnt_work, const uint32_t** idx_out, const SbOut& sb,        \
                                   uint8_t* d_resolved) {                                                                           \
        SVFM_PLANES_SWITCH(P, VB, run_search_sweep, s, pb, plan, final_mode, d_sp_work, d_cnt_work, idx_out, sb, d_resolved)   \
    }                                                                                                                               \
    static int NAME##_small(uint32_t planes, svfm_session* s, const PatternBatch& pb, const SmallOut& out) {                        \
        SVFM_PLANES_SWITCH(P, VB, run_small, s, pb, out)                                                                            \
    }                                                                                                                               \
    static int NAME##_locate(uint32_t planes, svfm_session* s, uint64_t n, const uint32_t* idx, const void* d_sp_work,              \
                             const void* d_cnt_work, const uint64_t* d_offs, uint64_t total, uint64_t heavy_seen,                   \
                             void* d_positions, uint32_t* d_rec_key, void* d_recs, unsigned long long* d_cursor,                    \
                             const uint8_t* d_resolved) {                                                                           \
        SVFM_PLANES_SWITCH(P, VB, run_locate, s, n, idx, d_sp_work, d_cnt_work, d_offs, total, heavy_seen, d_positions, d_rec_key,  \
                           d_recs, d_cursor, d_resolved)                                                                            \
    }                                                                                                                               \
    extern const TypeOps NAME;                                                                                                      \
    const TypeOps NAME = {NAME##_search, NAME##_build_ext, NAME##_build_text, NAME##_search_sweep, NAME##_small, NAME##_locate};

extern const TypeOps ops_p32_v32, ops_p32_v64, ops_p32_v128, ops_p64_v32, ops_p64_v64, ops_p64_v128;

}  // namespace svfm
