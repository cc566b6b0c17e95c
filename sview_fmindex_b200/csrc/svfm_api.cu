// svfm_api.cu -- C ABI (include/svfm.h) of the B200-native batched FM-index search engine.
// Index handle, sessions (stream + scratch arena), host/device batch entry points.
// Citations are relative to the reference's sview-fmindex/src/.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/reverse_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "engine.cuh"

namespace svfm {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

// ---------------------------------------------------------------------------------------------
// index handle
// ---------------------------------------------------------------------------------------------
static int finish_load(svfm_index* ix) {
    // text length is not stored in a header: it equals count_array[S] (count_array.rs:117,125)
    const Layout& L = ix->L;
    uint64_t v = 0;
    const uint64_t P = ix->type.pos_bits / 8;
    SVFM_CUDA(cudaMemcpy(&v, ix->d_blob + L.off_count_array + (uint64_t)(L.count_array_len - 1) * P, P, cudaMemcpyDeviceToHost));
    ix->text_len = v;
    {
        uint8_t ca[65 * 8] = {0};
        SVFM_CUDA(cudaMemcpy(ca, ix->d_blob + L.off_count_array, (uint64_t)L.count_array_len * P, cudaMemcpyDeviceToHost));
        uint64_t prev = 0;
        ix->symbols_present = 0;
        std::memset(ix->sym_rank, 0xff, 64);
        std::memset(ix->present, 0, 64);
        for (uint32_t i = 1; i < L.count_array_len && i <= 64; i++) {
            uint64_t c = 0;
            std::memcpy(&c, ca + i * P, P);
            if (c > prev) {  // symbol i-1 occurs count_array[i] - count_array[i-1] times
                ix->sym_rank[i - 1] = (uint8_t)ix->symbols_present;
                ix->present[ix->symbols_present] = (uint8_t)(i - 1);
                ix->symbols_present++;
            }
            prev = c;
        }
    }
    v = 0;
    SVFM_CUDA(cudaMemcpy(&v, ix->d_blob + L.off_sentinel_index, P, cudaMemcpyDeviceToHost));
    ix->sentinel_index = v;
    // the device kernels index blocks[pos / BLOCK_LEN] for pos <= text_len: make sure the header agrees
    if (ix->text_len / ix->type.vec_bits + 1 != L.blocks_len || ix->sentinel_index > ix->text_len ||
        ix->sentinel_index == 0)
        return SVFM_ERR_INVALID_FORMAT;
    const uint64_t r = L.sampling_ratio;
    if (ix->text_len / r + (ix->text_len % r ? 1 : 0) != L.suffix_array_len) return SVFM_ERR_INVALID_FORMAT;
    return SVFM_OK;
}

static int build_ext_table(svfm_index* ix);
static int build_ilv_table(svfm_index* ix);
static int build_text_copy(svfm_index* ix);
static int build_swp_table(svfm_index* ix);

static int load_common(const uint8_t* blob, size_t blob_len, svfm_type t, int device, bool src_on_device,
                       svfm_index** out, uint64_t err_detail[2]) {
    if (!out) return SVFM_ERR_BAD_ARG;
    *out = nullptr;
    if (!type_ok(t)) return SVFM_ERR_BAD_TYPE;
    if (!blob) return SVFM_ERR_BAD_ARG;
    if (src_on_device) SVFM_CUDA(cudaSetDevice(device));
    Layout L;
    int rc = parse_blob(
        [&](uint8_t* dst, uint64_t off, uint64_t n) {
            if (src_on_device) return cudaMemcpy(dst, blob + off, n, cudaMemcpyDeviceToHost) == cudaSuccess;
            std::memcpy(dst, blob + off, n);
            return true;
        },
        blob_len, t, L, err_detail);
    if (rc) return rc;
    SVFM_CUDA(cudaSetDevice(device));
    // (cudaLimitMaxL2FetchGranularity = 32 was tried here: on B200 it reads back but L2 still fills 64-byte sector
    // pairs -- tools/l2gran.cu -- so the library leaves the device limits alone.)
    svfm_index* ix = new svfm_index();
    ix->type = t;
    ix->L = L;
    ix->device = device;
    ix->blob_len = blob_len;
    // Shift the device copy so that the blocks section starts on a 32-byte boundary (sections are only
    // 8/16-byte aligned inside the blob): 16- and 32-byte blocks then never straddle a 32 B sector.
    const uint64_t pad = (32 - (L.off_blocks % 32)) % 32;
    cudaError_t e = cudaMalloc(&ix->d_alloc, blob_len + pad + 64);
    if (e != cudaSuccess) {
        g_last_error = std::string("cudaMalloc(blob): ") + cudaGetErrorString(e);
        delete ix;
        return e == cudaErrorMemoryAllocation ? SVFM_ERR_NOMEM : SVFM_ERR_CUDA;
    }
    ix->d_blob = ix->d_alloc + pad;
    e = cudaMemcpy(ix->d_blob, blob, blob_len, src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        g_last_error = std::string("cudaMemcpy(blob): ") + cudaGetErrorString(e);
        cudaFree(ix->d_alloc);
        delete ix;
        return SVFM_ERR_CUDA;
    }
    rc = finish_load(ix);
    if (rc == SVFM_OK) rc = build_ext_table(ix);
    if (rc == SVFM_OK) rc = build_ilv_table(ix);
    if (rc == SVFM_OK) rc = build_swp_table(ix);
    if (rc == SVFM_OK) rc = build_text_copy(ix);
    if (rc) {
        if (ix->d_swp) cudaFree(ix->d_swp);
        if (ix->d_fsa) cudaFree(ix->d_fsa);
        if (ix->d_text) cudaFree(ix->d_text);
        if (ix->d_ilv) cudaFree(ix->d_ilv);
        if (ix->d_ext) cudaFree(ix->d_ext);
        cudaFree(ix->d_alloc);
        delete ix;
        return rc;
    }
    *out = ix;
    return SVFM_OK;
}

// ---------------------------------------------------------------------------------------------
// sessions
// ---------------------------------------------------------------------------------------------
static void session_delete(svfm_session* s);
static int session_new(svfm_index* ix, svfm_session** out) {
    SVFM_CUDA(cudaSetDevice(ix->device));
    svfm_session* s = new svfm_session();
    s->ix = ix;
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_copied, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_err, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_counters, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaHostAlloc(&s->h_pinned, 8 * sizeof(uint64_t), cudaHostAllocDefault);
    if (e == cudaSuccess && ix->l2_window_bytes) {
        cudaStreamAttrValue v{};
        v.accessPolicyWindow.base_ptr = ix->d_ext;
        v.accessPolicyWindow.num_bytes = ix->l2_window_bytes;
        v.accessPolicyWindow.hitRatio = 1.0f;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (cudaStreamSetAttribute(s->stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) (void)cudaGetLastError();
    }
    if (e != cudaSuccess) {
        g_last_error = std::string("session_new: ") + cudaGetErrorString(e);
        (void)cudaGetLastError();
        session_delete(s);  // frees whatever was created before the failure (every member is null-checked)
        return e == cudaErrorMemoryAllocation ? SVFM_ERR_NOMEM : SVFM_ERR_CUDA;
    }
    *out = s;
    return SVFM_OK;
}

static void collect_spans(svfm_session* s);

static void session_delete(svfm_session* s) {
    if (!s) return;
    cudaSetDevice(s->ix->device);
    if (s->stream) { cudaStreamSynchronize(s->stream); cudaStreamDestroy(s->stream); }
    if (s->copy_stream) { cudaStreamSynchronize(s->copy_stream); cudaStreamDestroy(s->copy_stream); }
    if (s->ev_done) cudaEventDestroy(s->ev_done);
    if (s->ev_copied) cudaEventDestroy(s->ev_copied);
    if (s->d_err) cudaFree(s->d_err);
    if (s->d_counters) cudaFree(s->d_counters);
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    if (s->h_small) cudaFreeHost(s->h_small);
    collect_spans(s);
    for (cudaEvent_t e : s->free_events) cudaEventDestroy(e);
    delete s;
}

static void collect_spans(svfm_session* s) {
    for (auto& sp : s->spans) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess) s->phase_ms[sp.phase] += ms;
        s->free_events.push_back(sp.e0);
        s->free_events.push_back(sp.e1);
    }
    s->spans.clear();
}

struct SessionLease {
    svfm_index* ix;
    svfm_session* s = nullptr;
    explicit SessionLease(svfm_index* i) : ix(i) {}
    int acquire() {
        {
            std::lock_guard<std::mutex> g(ix->pool_mu);
            if (!ix->pool.empty()) { s = ix->pool.back(); ix->pool.pop_back(); }
        }
        if (s) return cudaSetDevice(ix->device) == cudaSuccess ? SVFM_OK : SVFM_ERR_CUDA;
        return session_new(ix, &s);
    }
    ~SessionLease() {
        if (s) { std::lock_guard<std::mutex> g(ix->pool_mu); ix->pool.push_back(s); }
    }
};

static std::atomic<uint64_t> g_sort_min{[] {
    const char* e = std::getenv("SVFM_SORT_MIN");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)SVFM_TUNE_AUTO;
}()};
// AUTO: with the extended table the first m symbols cost one lookup and every later step is a private row, so a
// locality sort has nothing left to share (measured: 0.16 ms unsorted vs 0.32 ms sorted for 3*10^5 20-mers, 176 vs 164 M
// reads/s for 150 bp reads); without the table the early steps still profit from it.
static uint64_t sort_min_patterns(const svfm_index* ix) {
    const uint64_t v = g_sort_min.load();
    if (v != (uint64_t)SVFM_TUNE_AUTO) return v;
    return ix->ext_m ? ~(uint64_t)0 : (uint64_t)(1u << 17);
}
static std::atomic<uint64_t> g_sweep_min{[] {
    const char* e = std::getenv("SVFM_SWEEP_MIN");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)SVFM_TUNE_AUTO;
}()};
// AUTO: the measured break-even with the plain search kernel on a 1 Gbp index (B200, 20-mers).  Without the row-derived
// structures: about 5.5 M patterns with a 2^24-entry extended table (8 steps left of a 20-mer), about 10 M with a 2^28-entry
// one (6 steps left).  With the packed text copy AND the expanded suffix array the plain kernel finishes a pattern with one
// array read and one text comparison as soon as a single row is left, and `locate` costs it one more read: count+locate
// 4.2 / 8.3 / 16.7 / 33 M patterns: 0.69 / 1.31 / 2.55 / 4.99 ms against 1.27 / 1.70 / 2.88 / 4.30 ms for the sweep search
// (break-even ~24 M); count alone 7.3-7.5 G patterns/s against 5.8 / 7.4 / 9.1 / 10.5 (break-even ~8 M).
static uint64_t sweep_min_patterns(const svfm_index* ix, bool for_locate = true) {
    const uint64_t v = g_sweep_min.load();
    if (v != (uint64_t)SVFM_TUNE_AUTO) return v;
    if (ix->d_text && ix->d_fsa) {
        // (2^28-entry table: break-even ~24 M / ~8 M; with the default 2^30-entry table the plain kernel runs 10^8 patterns in
        // 11.6 ms against 10.6 ms, 8.3 M in 1.06 ms against 1.6 ms)
        const bool big = ix->ext_entries > (1ull << 28);
        return for_locate ? (uint64_t)((big ? 48u : 24u) << 20) : (uint64_t)((big ? 16u : 8u) << 20);
    }
    return ix->ext_entries >= (1ull << 26) ? (uint64_t)(10u << 20) : (uint64_t)(5u << 20);
}
static std::atomic<uint64_t> g_ext_bits{[] {  // extended table: at most 2^bits entries (0 = no table)
    const char* e = std::getenv("SVFM_EXT_BITS");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)SVFM_TUNE_AUTO;
}()};
// AUTO: HBM capacity is what a B200 has plenty of -- 2^30 entries (8 GiB with u32 positions: 15 DNA symbols per lookup on
// a >= 0.54 Gbp text) when that is at most 1/8 of the device memory still free after the blob upload, else 2^28 (2 GiB, 14
// symbols) when that is at most 1/16, else 2^24 (128 MiB).  Measured on B200, 1 Gbp DNA, 10^8 20-mers, count+locate: 2^28 ->
// 2^30 entries takes the sweep search from 11.05 to 10.63 ms and the plain kernel from 14.1 to 11.6 ms.
static uint64_t ext_bits_for(const svfm_index* ix) {
    const uint64_t v = g_ext_bits.load();
    if (v != (uint64_t)SVFM_TUNE_AUTO) return v;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { (void)cudaGetLastError(); return 24; }
    const uint64_t pair = 2 * (uint64_t)(ix->type.pos_bits / 8);
    if ((pair << 30) <= free_b / 8) return 30;
    return (pair << 28) <= free_b / 16 ? 28 : 24;
}
static std::atomic<uint64_t> g_ilv{[] {  // build the interleaved occ copy at load
    const char* e = std::getenv("SVFM_ILV");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)1;
}()};
static std::atomic<uint64_t> g_bucket_sortback{[] {  // 1: bucketed sort-back (search_kernels.cuh), 0: radix sort-back
    const char* e = std::getenv("SVFM_BUCKET_SORTBACK");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)1;
}()};
static std::atomic<uint64_t> g_own_radix{[] {  // sweep presort: 1 = radix_pass_kernel, 0 = cub::DeviceRadixSort
    const char* e = std::getenv("SVFM_OWN_RADIX");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)0;
}()};
static std::atomic<uint64_t> g_sweep_final_sort{[] {  // locate: partition once more after the last round
    const char* e = std::getenv("SVFM_SWEEP_FINAL_SORT");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)1;
}()};

static SortPlan plan_sort(const svfm_index* ix, uint64_t n, const PatternBatch& pb, bool for_locate) {
    SortPlan p;
    const uint32_t S = ix->L.symbol_count;
    p.bits = (uint32_t)bits_for(S);
    if (p.bits == 0) p.bits = 1;
    if (n > 0xffffffffull) return p;
    // sweep items carry the symbols beyond the table's m as RANKS among the symbols that occur in the text (a pattern
    // with any other symbol has count 0 before the search starts), ceil(log2 s_eff) bits each
    uint32_t rank_bits = (uint32_t)bits_for(ix->symbols_present);
    if (rank_bits == 0) rank_bits = 1;
    if (!pb.offs && ix->ext_m && n >= sweep_min_patterns(ix, for_locate) && n < (1ull << 30) /* look-back descriptors: 30-bit counts */ &&
        pb.fixed_len >= ix->ext_m && (uint64_t)(pb.fixed_len - ix->ext_m) * rank_bits <= 64) {
        p.sweep = true;
        p.bits = rank_bits;
        p.m = ix->ext_m;
        p.prefix_bits = bits_for(ix->ext_entries) < 1 ? 1 : bits_for(ix->ext_entries);
        p.rest64 = (uint64_t)(pb.fixed_len - ix->ext_m) * p.bits > 32;
        // steps per round: the partition digit (bits * steps) has at most ROUND_MAX_BINS values; more steps per round
        // save partitions but interleave more symbol classes inside a round
        static const uint32_t t_env = [] { const char* e = std::getenv("SVFM_SWEEP_STEPS"); return e ? (uint32_t)atoi(e) : 0u; }();
        uint32_t t = t_env ? t_env : 6u / p.bits;
        while (t > 1 && p.bits * t > 9) t--;
        p.steps_per_round = t < 1 ? 1 : t;
        // With the packed text copy the generic kernel finishes a pattern by text verification as soon as its interval is
        // down to a few rows, whatever is left of it; the sweep search still walks every symbol, one partition round per
        // steps_per_round symbols.  Measured on B200: the sweep wins while it needs at most 2 rounds (20-mers on 1 Gbp DNA:
        // 2 rounds), the generic kernel beyond (32-mers on 3.1 Gbp: 6 rounds, 12-mer proteins: 6 rounds of one step).
        static const uint32_t max_rounds_env = [] { const char* e = std::getenv("SVFM_SWEEP_MAX_ROUNDS"); return e ? (uint32_t)atoi(e) : 2u; }();
        const uint32_t rounds = (pb.fixed_len - ix->ext_m + p.steps_per_round - 1) / p.steps_per_round;
        if (ix->d_text && sweep_min_patterns(ix, for_locate) != 0 && rounds > max_rounds_env) { p.sweep = false; p.bits = (uint32_t)bits_for(S) ? (uint32_t)bits_for(S) : 1u; }
        else return p;
    }
    if (n < sort_min_patterns(ix)) return p;
    // Sort on as many trailing symbols as it takes to tell the occ blocks apart: log_{S_eff}(blocks) + 1
    // symbols, S_eff = symbols that actually occur in the text (count_array).
    const double s_eff = ix->symbols_present > 1 ? (double)ix->symbols_present : 2.0;
    uint32_t m = 1;
    double reach = s_eff;
    while (reach < (double)ix->L.blocks_len && m < 64u / p.bits) { reach *= s_eff; m++; }
    m = m + 1 < 64u / p.bits ? m + 1 : 64u / p.bits;
    p.begin_bit = 64 - (int)(m * p.bits);
    p.end_bit = 64;
    p.sorted = true;
    return p;
}

static int run_presort(svfm_session* s, const PatternBatch& pb, const SortPlan& plan, const uint64_t** keys_out,
                       const uint32_t** idx_out) {
    int rc;
    const uint64_t rn = rsv(s, pb.n);
    if ((rc = s->keys0.reserve(rn * 8)) || (rc = s->keys1.reserve(rn * 8)) || (rc = s->vals0.reserve(rn * 4)) ||
        (rc = s->vals1.reserve(rn * 4)))
        return rc;
    cub::DoubleBuffer<uint64_t> keys((uint64_t*)s->keys0.ptr, (uint64_t*)s->keys1.ptr);
    cub::DoubleBuffer<uint32_t> vals((uint32_t*)s->vals0.ptr, (uint32_t*)s->vals1.ptr);
    size_t temp = 0;
    SVFM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp, keys, vals, (int64_t)pb.n, plan.begin_bit, plan.end_bit, s->stream));
    if ((rc = s->cub_temp.reserve(temp))) return rc;
    const int passes = (plan.end_bit - plan.begin_bit + 7) / 8;
    PhaseTimer pt(s, SVFM_PHASE_PRESORT, 2 + passes);
    const svfm_index* ix = s->ix;
    const uint8_t* table = ix->type.encoder ? ix->d_blob + ix->L.off_encoder : nullptr;
    pack_keys_kernel<<<resident_grid(pack_keys_kernel, pb.n, SEARCH_THREADS, ix->device), SEARCH_THREADS, 0, s->stream>>>(
        table, ix->L.symbol_count, pb, plan.bits, keys.Current(), vals.Current(), s->d_err);
    SVFM_CUDA(cudaGetLastError());
    SVFM_CUDA(cub::DeviceRadixSort::SortPairs(s->cub_temp.ptr, temp, keys, vals, (int64_t)pb.n, plan.begin_bit, plan.end_bit, s->stream));
    *keys_out = keys.Current();
    *idx_out = vals.Current();
    return SVFM_OK;
}

// Back to the caller's pattern order (count path): stable radix sort of (pattern index -> count) pairs; the
// indices are a permutation of 0..n-1, so the sorted values ARE the counts in the caller's order.
template <class P>
static int run_sortback_counts(svfm_session* s, uint64_t n, const uint32_t* idx, const void* d_cnt_work, void* d_counts_out) {
    int rc;
    if ((rc = s->rec_key_alt.reserve(n * 4))) return rc;
    size_t temp = 0;
    const int end_bit = bits_for(n) < 1 ? 1 : bits_for(n);
    SVFM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp, idx, (uint32_t*)s->rec_key_alt.ptr, (const P*)d_cnt_work,
                                              (P*)d_counts_out, (int64_t)n, 0, end_bit, s->stream));
    if ((rc = s->cub_temp.reserve(temp))) return rc;
    PhaseTimer pt(s, SVFM_PHASE_OTHER, 1 + (end_bit + 7) / 8);
    SVFM_CUDA(cub::DeviceRadixSort::SortPairs(s->cub_temp.ptr, temp, idx, (uint32_t*)s->rec_key_alt.ptr, (const P*)d_cnt_work,
                                              (P*)d_counts_out, (int64_t)n, 0, end_bit, s->stream));
    return SVFM_OK;
}

struct MinOp {
    __host__ __device__ uint64_t operator()(uint64_t a, uint64_t b) const { return a < b ? a : b; }
};

// Records (pattern index, position) -> the caller's CSR order.  Optional first stage (SVFM_SORTED): stable
// sort by position, so that positions end up ascending inside every pattern.  Second stage: stable sort by
// pattern index (LSD radix sort keeps SA-row order, or the ascending order of stage one, inside a pattern).
// With want_offs the CSR offsets are rebuilt from the sorted pattern indices (run starts + reverse running
// minimum); in direct mode they already exist.
template <class P>
static int run_sortback_records(svfm_session* s, uint64_t n, uint64_t total, bool by_position, bool want_offs,
                                uint64_t* d_out_offs, void** d_positions_inout) {
    int rc;
    if (total > 0) {
        if ((rc = s->positions_alt.reserve((total + 1) * sizeof(P))) || (rc = s->rec_key_alt.reserve((total + 1) * 4))) return rc;
        cub::DoubleBuffer<uint32_t> pat((uint32_t*)s->rec_key.ptr, (uint32_t*)s->rec_key_alt.ptr);
        cub::DoubleBuffer<P> pos((P*)s->positions.ptr, (P*)s->positions_alt.ptr);
        const int pos_bits = bits_for(s->ix->text_len + 1) < 1 ? 1 : bits_for(s->ix->text_len + 1);
        const int pat_bits = bits_for(n) < 1 ? 1 : bits_for(n);
        size_t t1 = 0, t2 = 0;
        SVFM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t1, pos, pat, (int64_t)total, 0, pos_bits, s->stream));
        SVFM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t2, pat, pos, (int64_t)total, 0, pat_bits, s->stream));
        if ((rc = s->cub_temp.reserve(t1 > t2 ? t1 : t2))) return rc;
        if (by_position) {
            PhaseTimer pt(s, SVFM_PHASE_SEGSORT, 1 + (pos_bits + 7) / 8);
            SVFM_CUDA(cub::DeviceRadixSort::SortPairs(s->cub_temp.ptr, t1, pos, pat, (int64_t)total, 0, pos_bits, s->stream));
        }
        {
            PhaseTimer pt(s, SVFM_PHASE_OTHER, 1 + (pat_bits + 7) / 8);
            SVFM_CUDA(cub::DeviceRadixSort::SortPairs(s->cub_temp.ptr, t2, pat, pos, (int64_t)total, 0, pat_bits, s->stream));
        }
        if (pos.Current() != (P*)s->positions.ptr) std::swap(s->positions, s->positions_alt);
        if (pat.Current() != (uint32_t*)s->rec_key.ptr) std::swap(s->rec_key, s->rec_key_alt);
        *d_positions_inout = s->positions.ptr;
    }
    if (!want_offs) return SVFM_OK;
    if ((rc = s->first.reserve((n + 1) * 8))) return rc;
    uint64_t* first = (uint64_t*)s->first.ptr;
    auto rin = thrust::make_reverse_iterator(first + n + 1);
    auto rout = thrust::make_reverse_iterator(d_out_offs + n + 1);
    size_t temp = 0;
    SVFM_CUDA(cub::DeviceScan::InclusiveScan(nullptr, temp, rin, rout, MinOp(), (int64_t)(n + 1), s->stream));
    if ((rc = s->cub_temp.reserve(temp))) return rc;
    PhaseTimer pt(s, SVFM_PHASE_OTHER, 4);
    SVFM_CUDA(cudaMemsetAsync(first, 0xff, n * 8, s->stream));
    s->h_pinned[4] = total;
    SVFM_CUDA(cudaMemcpyAsync(first + n, &s->h_pinned[4], 8, cudaMemcpyHostToDevice, s->stream));  // first[n] = total
    if (total)
        run_starts_kernel<<<grid_for(total, 256, s->ix->device), 256, 0, s->stream>>>((const uint32_t*)s->rec_key.ptr, total, first);
    SVFM_CUDA(cudaGetLastError());
    SVFM_CUDA(cub::DeviceScan::InclusiveScan(s->cub_temp.ptr, temp, rin, rout, MinOp(), (int64_t)(n + 1), s->stream));
    return SVFM_OK;
}

// Front of the sweep search, independent of the index type: pack_sweep_kernel + radix sort by table index
// (radix_pass_kernel, search_kernels.cuh).
template <class R>
int run_sweep_presort(svfm_session* s, const PatternBatch& pb, const SortPlan& plan, uint32_t rounds, SweepPre* out) {
    const svfm_index* ix = s->ix;
    using Pay = SweepPay<R>;
    const uint64_t n = pb.n;
    const uint32_t len = pb.fixed_len, bits = plan.bits;
    const uint32_t digit_bits = bits * plan.steps_per_round, nb_max = 1u << digit_bits;
    // Only the top 16 bits of the table index are sorted (two radix passes).  Items that differ in the lower bits only sit
    // within 1/65536 of the SA (a DNA table on 1 Gbp: 4096-16384 entries, ~240 occ blocks = 10 KB of index data, which
    // the CTAs working on that stretch share through L1/L2), and correctness never depends on the order.  Measured on B200,
    // 10^8 20-mers: 28 bits (4 passes) and 24 bits (3 passes) give the same round times; 16 bits make the first round
    // 0.6 ms slower and the sort 0.85 ms shorter (14.65 against 14.87 ms per batch); SVFM_PRESORT_BITS overrides.
    static const int sort_bits = [] { const char* e = std::getenv("SVFM_PRESORT_BITS"); return e ? atoi(e) : 16; }();
    const uint32_t sort_begin = plan.prefix_bits > sort_bits ? (uint32_t)(plan.prefix_bits - sort_bits) : 0u;
    // Who sorts: cub::DeviceRadixSort (onesweep) by default, this repo's radix_pass_kernel with SVFM_OWN_RADIX=1.  Measured
    // on B200, 10^8 items of 12 bytes, 8-bit digit: 0.84 ms per pass (cub) against 1.36 ms (radix_pass_kernel: a third of its
    // warp samples wait in the look-back, 10.6 polls per tile and digit; profiles/r2_radix_pass_ncu.txt) -- the library
    // kernel stays on the path until the own one is at least as fast.
    const bool own_radix = g_own_radix.load() != 0;
    const uint32_t passes = own_radix ? ((uint32_t)plan.prefix_bits - sort_begin + 7) / 8 : 0u;
    // histogram words: [rounds x nb_max] round digits | [passes x 256] presort digits | [rounds] round tile counters |
    // [64] PART_INDEX bin cursors | [passes] presort tile counters
    const uint64_t off_phist = (uint64_t)rounds * nb_max;
    const uint64_t off_counters = off_phist + (uint64_t)passes * RADIX_BINS;
    const uint64_t off_ptile = off_counters + rounds + 64;
    const uint64_t hist_words = off_ptile + passes + 8;
    const uint64_t rn = rsv(s, n);
    const uint64_t radix_tiles = (rn + RADIX_TILE - 1) / RADIX_TILE;
    int rc;
    if ((rc = s->keys0.reserve(rn * 4)) || (rc = s->keys1.reserve(rn * 4)) || (rc = s->pay0.reserve(rn * sizeof(Pay))) ||
        (rc = s->pay1.reserve(rn * sizeof(Pay))) || (rc = s->sweep_hist.reserve(hist_words * 4)) ||
        (rc = s->sweep_desc.reserve(radix_tiles * RADIX_BINS * 4)))
        return rc;
    uint32_t* key[2] = {(uint32_t*)s->keys0.ptr, (uint32_t*)s->keys1.ptr};
    Pay* pay[2] = {(Pay*)s->pay0.ptr, (Pay*)s->pay1.ptr};
    uint32_t* hist = (uint32_t*)s->sweep_hist.ptr;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix->device);
    cub::DoubleBuffer<uint32_t> ckey(key[0], key[1]);
    cub::DoubleBuffer<Pay> cpay(pay[0], pay[1]);
    size_t t1 = 0;
    if (!own_radix) {
        SVFM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t1, ckey, cpay, (int64_t)n, (int)sort_begin, plan.prefix_bits, s->stream));
        if ((rc = s->cub_temp.reserve(t1))) return rc;
    }
    PhaseTimer pt(s, SVFM_PHASE_PRESORT, own_radix ? 1 + passes : 2 + (plan.prefix_bits - sort_begin + 7) / 8);
    SVFM_CUDA(cudaMemsetAsync(hist, 0, hist_words * 4, s->stream));
    const uint8_t* table = ix->type.encoder ? ix->d_blob + ix->L.off_encoder : nullptr;
    DevSyms syms;
    syms.symbol_count = ix->L.symbol_count;
    syms.s_eff = ix->symbols_present;
    std::memcpy(syms.sym_rank, ix->sym_rank, 64);
    const size_t smem = (size_t)PACK_STAGES * (((size_t)PACK_TILE * len + 127) & ~(size_t)127) +
                        ((size_t)rounds * nb_max + (size_t)passes * RADIX_BINS) * 4;
    // always the same value (concurrent sessions pack batches of different pattern lengths; the attribute is per function)
    SVFM_CUDA(cudaFuncSetAttribute(pack_sweep_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYNAMIC_SMEM));
    if (smem > (size_t)MAX_DYNAMIC_SMEM) return SVFM_ERR_TOO_LARGE;
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pack_sweep_kernel<R>, PACK_TILE, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    uint64_t grid = (n + PACK_TILE - 1) / PACK_TILE;
    if (grid > (uint64_t)sms * per_sm) grid = (uint64_t)sms * per_sm;
    pack_sweep_kernel<R><<<(unsigned)grid, PACK_TILE, smem, s->stream>>>(table, syms, pb, bits, plan.m, key[0], pay[0], digit_bits, rounds,
                                                                        sort_begin, passes, hist, s->d_err);
    SVFM_CUDA(cudaGetLastError());
    if (!own_radix) {
        SVFM_CUDA(cub::DeviceRadixSort::SortPairs(s->cub_temp.ptr, t1, ckey, cpay, (int64_t)n, (int)sort_begin, plan.prefix_bits, s->stream));
        out->prefix = ckey.Current();
        out->pay = cpay.Current();
        out->hist = hist;
        out->counters = hist + off_counters;
        return SVFM_OK;
    }
    // stable LSD passes, 8 bits each
    const size_t rsmem = (size_t)RADIX_TILE * (sizeof(Pay) + 4) + (size_t)RADIX_WARPS * RADIX_BINS * 4;
    SVFM_CUDA(cudaFuncSetAttribute(radix_pass_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYNAMIC_SMEM));
    int rper = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rper, radix_pass_kernel<R>, RADIX_THREADS, rsmem) != cudaSuccess || rper < 1) rper = 1;
    const uint64_t n_tiles = (n + RADIX_TILE - 1) / RADIX_TILE;
    const uint64_t rgrid = n_tiles < (uint64_t)sms * rper ? n_tiles : (uint64_t)sms * rper;
    int cur = 0;
    for (uint32_t q = 0; q < passes; q++) {
        SVFM_CUDA(cudaMemsetAsync(s->sweep_desc.ptr, 0, n_tiles * RADIX_BINS * 4, s->stream));
        radix_pass_kernel<R><<<(unsigned)rgrid, RADIX_THREADS, rsmem, s->stream>>>(key[cur], pay[cur], key[cur ^ 1], pay[cur ^ 1], n,
                                                                                 sort_begin + 8 * q, hist + off_phist + (uint64_t)q * RADIX_BINS,
                                                                                 (uint32_t*)s->sweep_desc.ptr, hist + off_ptile + q);
        SVFM_CUDA(cudaGetLastError());
        cur ^= 1;
    }
    out->prefix = key[cur];
    out->pay = pay[cur];
    out->hist = hist;
    out->counters = hist + off_counters;
    return SVFM_OK;
}
template int run_sweep_presort<uint32_t>(svfm_session*, const PatternBatch&, const SortPlan&, uint32_t, SweepPre*);
template int run_sweep_presort<uint64_t>(svfm_session*, const PatternBatch&, const SortPlan&, uint32_t, SweepPre*);

// ---- dispatch: kernel-launching stages go through the per-(P, Vector) tables (inst_p*_v*.cu); the stages that only
// depend on the position width (radix sorts, scans) are instantiated here ------------------------------------------
static const TypeOps* type_ops(const svfm_type& t) {
    const bool p32 = t.pos_bits == 32;
    switch (t.vec_bits) {
        case 32: return p32 ? &ops_p32_v32 : &ops_p64_v32;
        case 64: return p32 ? &ops_p32_v64 : &ops_p64_v64;
        case 128: return p32 ? &ops_p32_v128 : &ops_p64_v128;
        default: return nullptr;
    }
}
#define SVFM_BY_POS(FN, ...) (s->ix->type.pos_bits == 32 ? FN<uint32_t>(__VA_ARGS__) : FN<uint64_t>(__VA_ARGS__))

static int dispatch_search(svfm_session* s, const PatternBatch& pb, const uint64_t* keys, const uint32_t* idx, uint32_t bits,
                           void* d_sp_work, void* d_cnt_work, const SbOut& sb = SbOut{}, uint8_t* d_resolved = nullptr) {
    const TypeOps* ops = type_ops(s->ix->type);
    return ops ? ops->search(s->ix->type.planes, s, pb, keys, idx, bits, d_sp_work, d_cnt_work, sb, d_resolved) : SVFM_ERR_BAD_TYPE;
}
static int dispatch_search_sweep(svfm_session* s, const PatternBatch& pb, const SortPlan& plan, int final_mode, void* d_sp_work,
                                 void* d_cnt_work, const uint32_t** idx_out, const SbOut& sb = SbOut{}, uint8_t* d_resolved = nullptr) {
    const TypeOps* ops = type_ops(s->ix->type);
    return ops ? ops->search_sweep(s->ix->type.planes, s, pb, plan, final_mode, d_sp_work, d_cnt_work, idx_out, sb, d_resolved)
               : SVFM_ERR_BAD_TYPE;
}
// Interleaved occ copy: { block q | checkpoint row q } per aligned slot.  Derived from the blob, bytes only.
static int build_ilv_table(svfm_index* ix) {
    if (!g_ilv.load()) return SVFM_OK;
    const Layout& L = ix->L;
    const uint64_t P = ix->type.pos_bits / 8;
    const uint64_t block_bytes = (uint64_t)ix->type.planes * ix->type.vec_bits / 8;
    const uint64_t row_bytes = (uint64_t)L.bwm_symbol_count * P;
    const uint64_t ck_off = (block_bytes + P - 1) / P * P;
    const uint64_t raw = ck_off + row_bytes;
    uint64_t stride = raw <= 32 ? 32 : raw <= 64 ? 64 : (raw + 127) / 128 * 128;
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, L.blocks_len * stride);
    if (e != cudaSuccess) {  // not enough memory for the copy: keep searching the blob in place
        (void)cudaGetLastError();
        return SVFM_OK;
    }
    ilv_build_kernel<<<grid_for(L.blocks_len * (stride / 4), 256, ix->device), 256>>>(
        reinterpret_cast<const uint32_t*>(ix->d_blob + L.off_blocks), (uint32_t)(block_bytes / 4),
        reinterpret_cast<const uint32_t*>(ix->d_blob + L.off_rank_checkpoints), (uint32_t)(row_bytes / 4), (uint32_t*)d,
        (uint32_t)(stride / 4), (uint32_t)(ck_off / 4), L.blocks_len);
    g_launches++;
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(d); SVFM_CUDA(e); }
    ix->d_ilv = (uint8_t*)d;
    ix->ilv_stride = (uint32_t)stride;
    ix->ilv_ck_off = (uint32_t)ck_off;
    return SVFM_OK;
}
// Sweep occ copy: one 32-byte sector per block (search_kernels.cuh).  Derived from the blob, bytes only; optional.
static std::atomic<uint64_t> g_sweep_occ{[] {
    const char* e = std::getenv("SVFM_SWEEP_OCC");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)1;
}()};
static int build_swp_table(svfm_index* ix) {
    if (!g_sweep_occ.load()) return SVFM_OK;
    const Layout& L = ix->L;
    const uint32_t npl = ix->type.planes, s_eff = ix->symbols_present;
    if (ix->type.vec_bits != 64 || npl > 3 || s_eff < 1 || s_eff > 4) return SVFM_OK;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || L.blocks_len * 32 > free_b / 8) { (void)cudaGetLastError(); return SVFM_OK; }
    void* d = nullptr;
    if (cudaMalloc(&d, L.blocks_len * 32) != cudaSuccess) { (void)cudaGetLastError(); return SVFM_OK; }
    SwpSyms syms{};
    syms.s_eff = s_eff;
    for (uint32_t j = 0; j < s_eff; j++) syms.present[j] = ix->present[j];
    const int grid = grid_for(L.blocks_len, 256, ix->device);
    const unsigned long long* blocks = reinterpret_cast<const unsigned long long*>(ix->d_blob + L.off_blocks);
    if (ix->type.pos_bits == 32)
        swp_build_kernel<uint32_t><<<grid, 256>>>(blocks, npl, reinterpret_cast<const uint32_t*>(ix->d_blob + L.off_rank_checkpoints),
                                                  L.bwm_symbol_count, syms, L.blocks_len, (ulonglong4*)d);
    else
        swp_build_kernel<uint64_t><<<grid, 256>>>(blocks, npl, reinterpret_cast<const uint64_t*>(ix->d_blob + L.off_rank_checkpoints),
                                                  L.bwm_symbol_count, syms, L.blocks_len, (ulonglong4*)d);
    g_launches++;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(d); SVFM_CUDA(e); }
    ix->d_swp = (uint8_t*)d;
    return SVFM_OK;
}
static std::atomic<uint64_t> g_l2_persist{[] {  // extended tables up to this many bytes get an L2 persisting window (0 = never)
    const char* e = std::getenv("SVFM_L2_PERSIST");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)(64u << 20);
}()};
static int build_ext_table(svfm_index* ix) {
    const TypeOps* ops = type_ops(ix->type);
    if (!ops) return SVFM_ERR_BAD_TYPE;
    int rc = ops->build_ext(ix->type.planes, ix, ext_bits_for(ix));
    if (rc) return rc;
    // north_star: "the k-mer table is staged in shared memory or an L2-persisting window".  The default table (2^30 entries,
    // 8 GiB) is far larger than L2 and is read once per pattern, so nothing is pinned for it; a table of at most 64 MiB
    // (SVFM_TUNE_EXT_BITS <= 23 for 32-bit positions: memory-constrained deployments) is marked persisting in L2 for every
    // kernel of every session of this index (cudaAccessPolicyWindow, set per stream in session_new).
    const uint64_t bytes = ix->d_ext ? ix->ext_entries * 2 * (ix->type.pos_bits / 8) : 0;
    int max_persist = 0, max_window = 0;
    if (bytes && bytes <= g_l2_persist.load() &&
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ix->device) == cudaSuccess &&
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ix->device) == cudaSuccess &&
        bytes <= (uint64_t)max_persist && bytes <= (uint64_t)max_window) {
        size_t cur = 0;
        if (cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize) == cudaSuccess && cur < bytes &&
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes) != cudaSuccess)
            (void)cudaGetLastError();
        else
            ix->l2_window_bytes = bytes;
    }
    (void)cudaGetLastError();
    return SVFM_OK;
}
std::atomic<uint64_t> g_text{[] {  // build the packed text copy at load (text verification, search_kernels.cuh)
    const char* e = std::getenv("SVFM_TEXT");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)1;
}()};
std::atomic<uint64_t> g_full_sa{[] {  // build the expanded suffix array at load (search_kernels.cuh)
    const char* e = std::getenv("SVFM_FULL_SA");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)1;
}()};
// Row-derived structures: the expanded suffix array, then the packed text copy (engine.cuh, run_build_text).
static int build_text_copy(svfm_index* ix) {
    if (!g_text.load() && !g_full_sa.load()) return SVFM_OK;
    const TypeOps* ops = type_ops(ix->type);
    return ops ? ops->build_text(ix->type.planes, ix) : SVFM_ERR_BAD_TYPE;
}
static int dispatch_locate(svfm_session* s, uint64_t n, const uint32_t* idx, const void* d_sp_work, const void* d_cnt_work,
                           const uint64_t* d_offs, uint64_t total, uint64_t heavy_seen, void* d_positions, uint32_t* d_rec_key,
                           void* d_recs = nullptr, unsigned long long* d_cursor = nullptr, const uint8_t* d_resolved = nullptr) {
    const TypeOps* ops = type_ops(s->ix->type);
    return ops ? ops->locate(s->ix->type.planes, s, n, idx, d_sp_work, d_cnt_work, d_offs, total, heavy_seen, d_positions, d_rec_key,
                             d_recs, d_cursor, d_resolved)
               : SVFM_ERR_BAD_TYPE;
}
// Bucketed sort-back, second half (search_kernels.cuh): CSR offsets + final place of every record, one CTA per bucket.
template <class P>
static int run_sb_place(svfm_session* s, uint64_t n, uint64_t nb, void* d_out_offs, bool offs32, void* d_positions) {
    const size_t smem = SB_PLACE_SMEM + (size_t)SB_STAGE * sizeof(P);
    SVFM_CUDA(cudaFuncSetAttribute(sb_place_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PhaseTimer pt(s, SVFM_PHASE_OTHER, 1);
    sb_place_kernel<P><<<(unsigned)nb, SB_PLACE_THREADS, smem, s->stream>>>((const SbRec<P>*)s->sb_recs.ptr, (const uint64_t*)s->sb_base.ptr,
                                                                           n, nb, d_out_offs, offs32 ? 1 : 0, (P*)d_positions);
    SVFM_CUDA(cudaGetLastError());
    return SVFM_OK;
}
template <class P>
static int run_scatter_counts(svfm_session* s, uint64_t n, const uint32_t* idx, const void* d_cnt_work, void* d_counts_out) {
    PhaseTimer pt(s, SVFM_PHASE_OTHER, 1);
    scatter_counts_kernel<P><<<grid_for(n, 256, s->ix->device), 256, 0, s->stream>>>(idx, (const P*)d_cnt_work, n, (P*)d_counts_out);
    SVFM_CUDA(cudaGetLastError());
    return SVFM_OK;
}
__global__ void narrow_offs_kernel(const uint64_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = (uint32_t)in[i];
}
static int dispatch_scan(svfm_session* s, uint64_t n, const void* d_cnt, uint64_t* d_out_offs) {
    return SVFM_BY_POS(run_scan, s, n, d_cnt, d_out_offs);
}
static int dispatch_sortback_counts(svfm_session* s, uint64_t n, const uint32_t* idx, const void* d_cnt_work, void* d_counts_out) {
    return SVFM_BY_POS(run_sortback_counts, s, n, idx, d_cnt_work, d_counts_out);
}
static int dispatch_sortback_records(svfm_session* s, uint64_t n, uint64_t total, bool by_position, bool want_offs,
                                     uint64_t* d_out_offs, void** d_positions) {
    return SVFM_BY_POS(run_sortback_records, s, n, total, by_position, want_offs, d_out_offs, d_positions);
}

static int err_from_bits(int bits) {
    if (bits & ERRBIT_EMPTY_PATTERN) return SVFM_ERR_EMPTY_PATTERN;
    if (bits & ERRBIT_BAD_SYMBOL) return SVFM_ERR_BAD_SYMBOL;
    return SVFM_OK;
}

// Device-resident count: small batches run the search kernel in the caller's order; dense fixed-length batches take the
// sweep search, whose last round groups the items by the top bits of the pattern index so that one scatter pass puts the
// counts back in the caller's order; the locality-sorted generic path sorts them back.  Leaves the error bits in s->d_err.
static int count_device(svfm_session* s, const PatternBatch& pb, void* d_counts_out) {
    SVFM_CUDA(cudaMemsetAsync(s->d_err, 0, sizeof(int), s->stream));
    SVFM_CUDA(cudaMemsetAsync(s->d_counters, 0, 4 * sizeof(unsigned long long), s->stream));
    if (pb.n == 0) return SVFM_OK;
    const SortPlan plan = plan_sort(s->ix, pb.n, pb, false);
    if (!plan.sorted && !plan.sweep) return dispatch_search(s, pb, nullptr, nullptr, 1, nullptr, d_counts_out);
    const uint64_t P = s->ix->type.pos_bits / 8;
    const uint64_t* keys = nullptr;
    const uint32_t* idx = nullptr;
    int rc;
    if ((rc = s->cnt.reserve((rsv(s, pb.n) + 1) * P))) return rc;
    if (plan.sweep) {
        if ((rc = s->sp.reserve((rsv(s, pb.n) + 1) * P))) return rc;
        const bool by_index = g_bucket_sortback.load() != 0;
        if ((rc = dispatch_search_sweep(s, pb, plan, by_index ? PART_INDEX : PART_NONE, s->sp.ptr, s->cnt.ptr, &idx))) return rc;
        if (by_index) return SVFM_BY_POS(run_scatter_counts, s, pb.n, idx, s->cnt.ptr, d_counts_out);
    } else {
        if ((rc = run_presort(s, pb, plan, &keys, &idx))) return rc;
        if ((rc = dispatch_search(s, pb, keys, idx, plan.bits, nullptr, s->cnt.ptr))) return rc;
    }
    return dispatch_sortback_counts(s, pb.n, idx, s->cnt.ptr, d_counts_out);
}

// Device-resident locate pipeline.
//   small batch : search -> scan -> (sync: total) -> LF-walk straight into CSR order [-> sort by position]
//   large batch : reorder (sweep search, or locality sort + search kernel); the search also sums the counts per bucket of
//                 8192 pattern indices -> bucket bases -> (sync: total) -> LF-walk, records straight into their bucket
//                 -> sb_place_kernel: CSR offsets + positions in the caller's order          ("bucketed sort-back")
//   SVFM_SORTED on a large batch, or 2^32 and more occurrences: scan (work order) -> LF-walk into records -> radix sort
//                 by position -> stable radix sort by pattern index -> CSR offsets
// d_out_offs: u64[n+1], or u32[n+1] with SVFM_OFFS32 (SVFM_ERR_TOO_LARGE when the total does not fit).
// The downloads of the session's previous chunk (locate_host) read out_offs / positions: every kernel that writes them
// is ordered behind the copies.  Called right before the first such kernel, i.e. AFTER the search kernels are enqueued,
// so that the search of this chunk overlaps the downloads of the last one.
static int order_behind_pending_copy(svfm_session* s, bool host_wait = false) {
    if (!s->copy_pending) return SVFM_OK;
    if (host_wait) SVFM_CUDA(cudaEventSynchronize(s->ev_copied));
    else SVFM_CUDA(cudaStreamWaitEvent(s->stream, s->ev_copied, 0));
    s->copy_pending = false;
    return SVFM_OK;
}

static int locate_device(svfm_session* s, const PatternBatch& pb, uint32_t flags, void* d_out_offs,
                         void** d_positions, uint64_t* total_out) {
    const uint64_t P = s->ix->type.pos_bits / 8;
    int rc;
    SVFM_CUDA(cudaMemsetAsync(s->d_err, 0, sizeof(int), s->stream));
    SVFM_CUDA(cudaMemsetAsync(s->d_counters, 0, 4 * sizeof(unsigned long long), s->stream));
    const uint64_t rn = rsv(s, pb.n);
    if ((rc = s->sp.reserve((rn + 1) * P))) return rc;
    if ((rc = s->cnt.reserve((rn + 1) * P))) return rc;
    const SortPlan plan = plan_sort(s->ix, pb.n, pb, true);
    const uint64_t* keys = nullptr;
    const uint32_t* idx = nullptr;
    const bool offs32 = (flags & SVFM_OFFS32) != 0;
    const bool by_position = (flags & SVFM_SORTED) != 0;
    const bool reordered = plan.sorted || plan.sweep;
    bool bucket = reordered && !by_position && g_bucket_sortback.load() != 0;
    const uint64_t nb = (pb.n + SB_BUCKET - 1) >> SB_SHIFT;
    SbOut sb{};
    uint8_t* d_resolved = nullptr;
    if (bucket) {
        const uint64_t rnb = (rn + SB_BUCKET - 1) >> SB_SHIFT;
        if ((rc = s->sb_hist.reserve(rnb * 4)) || (rc = s->sb_base.reserve((rnb + 1) * 8)) || (rc = s->sb_cursor.reserve(rnb * 8))) return rc;
        sb.hist = (uint32_t*)s->sb_hist.ptr;
        sb.total = s->d_counters + 2;
        SVFM_CUDA(cudaMemsetAsync(sb.hist, 0, nb * 4, s->stream));
    }
    if (plan.sweep) {
        uint8_t* res = nullptr;
        if (s->ix->d_text) {
            if ((rc = s->resolved.reserve(rn))) return rc;
            res = (uint8_t*)s->resolved.ptr;
            // all zero unless the last round resolves (then every flag is written): locate reads the array either way
            SVFM_CUDA(cudaMemsetAsync(res, 0, pb.n, s->stream));
            d_resolved = res;
        }
        if ((rc = dispatch_search_sweep(s, pb, plan, g_sweep_final_sort.load() != 0 ? PART_SYMBOLS : PART_NONE, s->sp.ptr, s->cnt.ptr,
                                        &idx, sb, res)))
            return rc;
    } else {
        if (plan.sorted && (rc = run_presort(s, pb, plan, &keys, &idx))) return rc;
        if (s->ix->d_text) {  // text verification: the search may hand back text positions instead of SA rows
            if ((rc = s->resolved.reserve(rn))) return rc;
            d_resolved = (uint8_t*)s->resolved.ptr;
        }
        if ((rc = dispatch_search(s, pb, keys, idx, plan.sorted ? plan.bits : 1, s->sp.ptr, s->cnt.ptr, sb, d_resolved))) return rc;
    }
    if (bucket) {
        {
            PhaseTimer pt(s, SVFM_PHASE_SCAN, 1);
            sb_scan_kernel<<<1, SB_SCAN_THREADS, 0, s->stream>>>(sb.hist, nb, (uint64_t*)s->sb_base.ptr, (unsigned long long*)s->sb_cursor.ptr);
            SVFM_CUDA(cudaGetLastError());
        }
        SVFM_CUDA(cudaMemcpyAsync(&s->h_pinned[0], s->d_counters + 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
        SVFM_CUDA(cudaMemcpyAsync(&s->h_pinned[1], s->d_err, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        SVFM_CUDA(cudaMemcpyAsync(&s->h_pinned[3], s->d_counters, sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
        SVFM_CUDA(cudaStreamSynchronize(s->stream));
        if ((rc = err_from_bits((int)(s->h_pinned[1] & 0xffffffffu)))) return rc;
        const uint64_t total = s->h_pinned[0];
        const uint64_t heavy_seen = s->h_pinned[3];
        *total_out = total;
        if (total < 0xffffffffull) {  // the 32-bit bucket counters did not wrap
            const uint64_t rec_bytes = P == 4 ? 8 : 16;
            if ((std::max(total, s->reserve_n) + 1) * P > s->positions.cap && (rc = order_behind_pending_copy(s, true))) return rc;  // the buffer is about to move
            if ((rc = s->positions.reserve((std::max(total, s->reserve_n) + 1) * P)) ||
                (rc = s->sb_recs.reserve((std::max(total, s->reserve_n) + 1) * rec_bytes)))
                return rc;
            *d_positions = s->positions.ptr;
            if ((rc = dispatch_locate(s, pb.n, idx, s->sp.ptr, s->cnt.ptr, nullptr, total, heavy_seen, nullptr, nullptr, s->sb_recs.ptr,
                                      (unsigned long long*)s->sb_cursor.ptr, d_resolved)))
                return rc;
            if ((rc = order_behind_pending_copy(s))) return rc;
            return SVFM_BY_POS(run_sb_place, s, pb.n, nb, d_out_offs, offs32, s->positions.ptr);
        }
        bucket = false;  // 2^32 records and more: radix sort-back
    }
    uint64_t* offs_final = (uint64_t*)d_out_offs;  // u64 CSR offsets in the caller's order
    if (offs32) {
        if ((rc = s->offs64.reserve((rn + 1) * 8))) return rc;
        offs_final = (uint64_t*)s->offs64.ptr;
    }
    uint64_t* offs_work = offs_final;  // small batch: work order == caller order
    if (reordered) {
        if ((rc = s->woffs.reserve((rn + 1) * 8))) return rc;
        offs_work = (uint64_t*)s->woffs.ptr;
    }
    if ((rc = order_behind_pending_copy(s))) return rc;   // the scan may write out_offs itself
    if ((rc = dispatch_scan(s, pb.n, s->cnt.ptr, offs_work))) return rc;
    SVFM_CUDA(cudaMemcpyAsync(&s->h_pinned[0], offs_work + pb.n, sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
    SVFM_CUDA(cudaMemcpyAsync(&s->h_pinned[1], s->d_err, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    SVFM_CUDA(cudaMemcpyAsync(&s->h_pinned[3], s->d_counters, sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
    SVFM_CUDA(cudaStreamSynchronize(s->stream));
    if ((rc = err_from_bits((int)(s->h_pinned[1] & 0xffffffffu)))) return rc;
    const uint64_t total = s->h_pinned[0];
    const uint64_t heavy_seen = s->h_pinned[3];
    *total_out = total;
    if (offs32 && total > 0xffffffffull) return SVFM_ERR_TOO_LARGE;
    if ((rc = s->positions.reserve((std::max(total, s->reserve_n) + 1) * P))) return rc;   // (no copy is pending any more)
    const bool records = reordered || by_position;
    if (records && (rc = s->rec_key.reserve((total + 1) * 4))) return rc;
    *d_positions = s->positions.ptr;
    if ((rc = dispatch_locate(s, pb.n, idx, s->sp.ptr, s->cnt.ptr, offs_work, total, heavy_seen, s->positions.ptr,
                              records ? (uint32_t*)s->rec_key.ptr : nullptr, nullptr, nullptr, d_resolved)))
        return rc;
    if (records && (rc = dispatch_sortback_records(s, pb.n, total, by_position, reordered, offs_final, d_positions))) return rc;
    if (offs32) {
        narrow_offs_kernel<<<grid_for(pb.n + 1, 256, s->ix->device), 256, 0, s->stream>>>(offs_final, pb.n + 1, (uint32_t*)d_out_offs);
        g_launches++;
        SVFM_CUDA(cudaGetLastError());
    }
    return SVFM_OK;
}

// Host-side validation shared by the host-buffer entry points.
static int check_host_patterns(const uint8_t* pats, const uint64_t* offs, uint64_t n, uint32_t fixed_len,
                               uint64_t* total_bytes) {
    if (n == 0) { *total_bytes = 0; return SVFM_OK; }
    if (!pats) return SVFM_ERR_BAD_ARG;
    if (!offs) {
        if (fixed_len == 0) return SVFM_ERR_EMPTY_PATTERN;
        *total_bytes = n * (uint64_t)fixed_len;
        return SVFM_OK;
    }
    for (uint64_t i = 0; i < n; i++) {
        if (offs[i + 1] < offs[i]) return SVFM_ERR_BAD_ARG;
        if (offs[i + 1] == offs[i]) return SVFM_ERR_EMPTY_PATTERN;
    }
    *total_bytes = offs[n];
    return SVFM_OK;
}

// ---------------------------------------------------------------------------------------------
// Host-buffer entry points.  A large batch is cut into chunks that flow through up to HOST_WORKERS
// sessions (one CUDA stream each, driven by one host thread each), so that the upload of chunk i+1, the
// kernels of chunk i and the download of chunk i-1 overlap on the copy engines and the SMs.
// ---------------------------------------------------------------------------------------------
static std::atomic<uint64_t> g_chunk_patterns{[] {
    const char* e = std::getenv("SVFM_CHUNK");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)SVFM_TUNE_AUTO;
}()};
static std::atomic<uint64_t> g_host_workers{[] {
    // 2: each worker overlaps the downloads of its last chunk with the search of its next one (copy stream), so two of them
    // keep the copy engines and the SMs busy; a third only adds contention between kernels (measured 4.77 / 4.98 / 4.60 G
    // patterns/s end to end with 3 / 2 / 4 workers, 10^8 packed 20-mers)
    const char* e = std::getenv("SVFM_WORKERS");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)2;
}()};
static const bool g_trace = std::getenv("SVFM_TRACE") != nullptr;
static double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// Result buffers handed to the caller (svfm_locate_batch_alloc): pinned memory for large results (asynchronous
// device-to-host copies), plain aligned memory below 1 MiB -- cudaHostAlloc/cudaFreeHost cost tens of microseconds,
// which dominates a single-pattern `locate`.  A 64-byte header in front of the data records which one it is.
struct ResultHeader { uint64_t magic, pinned; uint8_t pad[48]; };
static_assert(sizeof(ResultHeader) == 64, "result header");
constexpr uint64_t RESULT_MAGIC = 0x5356464d52534c54ull;  // "SVFMRSLT"
static void* result_alloc(uint64_t bytes) {
    void* raw = nullptr;
    const bool pinned = bytes >= (1u << 20);
    if (pinned) {
        if (cudaHostAlloc(&raw, bytes + sizeof(ResultHeader), cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    } else if (posix_memalign(&raw, 64, bytes + sizeof(ResultHeader)) != 0) {
        return nullptr;
    }
    ResultHeader* h = static_cast<ResultHeader*>(raw);
    h->magic = RESULT_MAGIC;
    h->pinned = pinned ? 1 : 0;
    return h + 1;
}
static void result_free(void* p) {
    if (!p) return;
    ResultHeader* h = static_cast<ResultHeader*>(p) - 1;
    if (h->magic != RESULT_MAGIC) return;  // not ours: leave it alone rather than corrupt the heap
    h->magic = 0;
    if (h->pinned) cudaFreeHost(h); else std::free(h);
}

template <class T>
__global__ void add_base_kernel(T* __restrict__ v, uint64_t n, T base) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) v[i] += base;
}

struct ChunkPlan {
    uint64_t n = 0, chunks = 0;
    std::vector<uint64_t> bounds;  // chunks + 1 pattern indices
    uint64_t begin(uint64_t c) const { return bounds[c]; }
    uint64_t end(uint64_t c) const { return bounds[c + 1]; }
    uint64_t largest() const { uint64_t m = 0; for (uint64_t c = 0; c < chunks; c++) m = std::max(m, end(c) - begin(c)); return m; }
};

// Even chunks of about SVFM_TUNE_CHUNK patterns, every boundary a multiple of 256 patterns (chunk copies stay 16-byte
// aligned for the TMA staging).  The LAST chunk is halved repeatedly (down to ~1 Mi patterns): the upload stream is the
// bottleneck of a host batch, so what remains after the last byte has arrived -- kernels + download of the final
// chunk -- should be small.
// AUTO: 4 Mi patterns.  Measured on B200, 10^8 20-mers, count+locate, pinned buffers (end-to-end G patterns/s, 2 workers, results
// leaving on the sessions' copy streams): 2-bit packed patterns 4.66 / 5.12 / 5.07 / 4.98 with 2 / 4 / 6 / 8 Mi chunks; byte
// patterns 2.35-2.41 whatever the chunk size (the upload bounds that call: 2 GB at PCIe speed).  Chunks this size take the
// plain search kernel (see sweep_min_patterns), whose time per pattern hardly depends on the batch size, and small chunks
// keep the head (first upload + first kernels) and the tail (last download) of the pipeline short.
static ChunkPlan plan_chunks(uint64_t n, uint64_t bytes_per_pattern = 0) {
    ChunkPlan p;
    p.n = n;
    uint64_t c = g_chunk_patterns.load();
    if (c == (uint64_t)SVFM_TUNE_AUTO) c = 4u << 20;
    (void)bytes_per_pattern;
    if (c == 0) c = n;
    uint64_t k = (n + c - 1) / c;
    if (k == 0) k = 1;
    uint64_t chunk = (n + k - 1) / k;
    if (k > 1) chunk = (chunk + 255) & ~(uint64_t)255;
    if (chunk == 0) chunk = 1;
    p.bounds.push_back(0);
    for (uint64_t at = chunk; at < n; at += chunk) p.bounds.push_back(at);
    static const bool taper = [] { const char* e = std::getenv("SVFM_TAPER"); return e ? atoi(e) != 0 : true; }();
    if (k > 1 && taper) {
        uint64_t a = p.bounds.back();
        while (n - a > (2u << 20)) {
            const uint64_t half = (((n - a) / 2) + 255) & ~(uint64_t)255;
            a += half;
            p.bounds.push_back(a);
        }
    }
    p.bounds.push_back(n);
    p.chunks = p.bounds.size() - 1;
    return p;
}

static std::atomic<uint64_t> g_upload_budget{[] {  // device bytes of pattern staging per host call
    const char* e = std::getenv("SVFM_UPLOAD_BUDGET");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)8 << 30;
}()};

struct UploaderLease {
    svfm_index* ix;
    svfm_uploader* u = nullptr;
    explicit UploaderLease(svfm_index* i) : ix(i) {}
    int acquire() {
        {
            std::lock_guard<std::mutex> g(ix->pool_mu);
            if (!ix->up_pool.empty()) { u = ix->up_pool.back(); ix->up_pool.pop_back(); }
        }
        SVFM_CUDA(cudaSetDevice(ix->device));
        if (u) return SVFM_OK;
        u = new svfm_uploader();
        cudaError_t e = cudaStreamCreateWithFlags(&u->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete u; u = nullptr; SVFM_CUDA(e); }
        return SVFM_OK;
    }
    ~UploaderLease() {
        if (u) { std::lock_guard<std::mutex> g(ix->pool_mu); ix->up_pool.push_back(u); }
    }
};

// One group of consecutive chunks [g0, g1) of a host batch whose pattern bytes fit the staging budget.
struct UploadGroup {
    svfm_uploader* u = nullptr;
    const uint8_t* pats = nullptr;
    const uint64_t* offs = nullptr;
    uint32_t fixed_len = 0, flags = 0;
    uint64_t g0 = 0, g1 = 0, first_pattern = 0, first_byte = 0;
    std::mutex mu;
    std::condition_variable cv;
    uint64_t issued = 0;  // chunks of this group whose copies are enqueued (events recorded)
    int rc = SVFM_OK;
};

static uint64_t chunk_bytes(const ChunkPlan& cp, const uint64_t* offs, uint32_t fixed_len, uint64_t c) {
    const uint64_t a = cp.begin(c), b = cp.end(c);
    return offs ? offs[b] - offs[a] : (b - a) * (uint64_t)fixed_len;
}

// Enqueue the copies of every chunk of the group on the uploader's stream (called by one host thread while the
// workers already consume the first chunks).
static void upload_group(const ChunkPlan& cp, UploadGroup& g) {
    svfm_uploader* u = g.u;
    for (uint64_t c = g.g0; c < g.g1; c++) {
        const uint64_t a = cp.begin(c), b = cp.end(c);
        cudaError_t e;
        if (!g.offs) {
            e = cudaMemcpyAsync((uint8_t*)u->pats.ptr + (a - g.first_pattern) * (uint64_t)g.fixed_len, g.pats + a * (uint64_t)g.fixed_len,
                                (b - a) * (uint64_t)g.fixed_len, cudaMemcpyHostToDevice, u->stream);
        } else {
            e = cudaMemcpyAsync((uint8_t*)u->pats.ptr + (g.offs[a] - g.first_byte), g.pats + g.offs[a], g.offs[b] - g.offs[a],
                                cudaMemcpyHostToDevice, u->stream);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync((uint64_t*)u->offs.ptr + (a - g.first_pattern), g.offs + a, (b - a + 1) * sizeof(uint64_t),
                                    cudaMemcpyHostToDevice, u->stream);
        }
        if (e == cudaSuccess) e = cudaEventRecord(u->ev[c - g.g0], u->stream);
        std::lock_guard<std::mutex> lk(g.mu);
        if (e != cudaSuccess) {
            g_last_error = std::string("upload: ") + cudaGetErrorString(e);
            g.rc = SVFM_ERR_CUDA;
            g.issued = g.g1 - g.g0;  // nobody may wait forever
            g.cv.notify_all();
            return;
        }
        g.issued = c - g.g0 + 1;
        g.cv.notify_all();
    }
}

// Worker side: make the session's stream wait for chunk c of the group and describe it as a PatternBatch.
static int await_chunk(svfm_session* s, const ChunkPlan& cp, UploadGroup& g, uint64_t c, PatternBatch& pb) {
    {
        std::unique_lock<std::mutex> lk(g.mu);
        g.cv.wait(lk, [&] { return g.issued > c - g.g0; });
        if (g.rc) return g.rc;
    }
    SVFM_CUDA(cudaStreamWaitEvent(s->stream, g.u->ev[c - g.g0], 0));
    const uint64_t a = cp.begin(c), b = cp.end(c);
    pb.n = b - a;
    pb.fixed_len = g.fixed_len;
    pb.reversed = (g.flags & SVFM_REVERSED) ? 1u : 0u;
    pb.preencoded = 0;
    if (!g.offs) {
        pb.pats = (const uint8_t*)g.u->pats.ptr + (a - g.first_pattern) * (uint64_t)g.fixed_len;
        pb.offs = nullptr;
    } else {
        pb.pats = (const uint8_t*)g.u->pats.ptr - g.first_byte;  // offsets stay absolute
        pb.offs = (const uint64_t*)g.u->offs.ptr + (a - g.first_pattern);
    }
    return SVFM_OK;
}

// Next group of chunks starting at g0: reserve its staging and events.
static int begin_group(svfm_uploader* u, const ChunkPlan& cp, const uint8_t* pats, const uint64_t* offs, uint32_t fixed_len,
                       uint32_t flags, uint64_t g0, UploadGroup& g) {
    const uint64_t budget = g_upload_budget.load();
    uint64_t g1 = g0, bytes = 0;
    while (g1 < cp.chunks) {
        const uint64_t cb = chunk_bytes(cp, offs, fixed_len, g1);
        if (g1 > g0 && bytes + cb > budget) break;
        bytes += cb;
        g1++;
    }
    g.u = u;
    g.pats = pats;
    g.offs = offs;
    g.fixed_len = fixed_len;
    g.flags = flags;
    g.g0 = g0;
    g.g1 = g1;
    g.first_pattern = cp.begin(g0);
    g.first_byte = offs ? offs[g.first_pattern] : g.first_pattern * (uint64_t)fixed_len;
    g.issued = 0;
    g.rc = SVFM_OK;
    int rc;
    if ((rc = u->pats.reserve(bytes + 256))) return rc;
    if (offs && (rc = u->offs.reserve((cp.end(g1 - 1) - g.first_pattern + 1) * sizeof(uint64_t)))) return rc;
    while (u->ev.size() < g1 - g0) {
        cudaEvent_t e;
        SVFM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        u->ev.push_back(e);
    }
    return SVFM_OK;
}

// per_chunk(session, c) for c in [c0, c1) over up to `workers` sessions; `prologue` runs on the calling thread before it
// turns into a worker itself.
template <class Fn, class Pro>
static int run_workers(svfm_index* ix, uint64_t c0, uint64_t c1, uint64_t reserve_n, Fn&& per_chunk, Pro&& prologue) {
    const uint64_t chunks = c1 - c0;
    const uint64_t hw = g_host_workers.load() ? g_host_workers.load() : 1;
    const int workers = (int)(chunks < hw ? chunks : hw);
    std::atomic<uint64_t> next{c0};
    std::atomic<int> first_err{SVFM_OK};
    std::string err_text;
    std::mutex err_mu;
    auto body = [&]() {
        SessionLease lease(ix);
        int rc = lease.acquire();
        if (lease.s) lease.s->reserve_n = reserve_n;
        for (;;) {
            const uint64_t c = next.fetch_add(1);
            if (c >= c1) break;
            const bool run = (rc == SVFM_OK && first_err.load() == SVFM_OK);
            if (run) rc = per_chunk(lease.s, c);
            if (rc != SVFM_OK) {
                int expected = SVFM_OK;
                if (first_err.compare_exchange_strong(expected, rc)) {
                    std::lock_guard<std::mutex> g(err_mu);
                    err_text = g_last_error;
                }
            }
            if (!run || rc != SVFM_OK) per_chunk(nullptr, c);  // publish "nothing": nobody may wait on this chunk forever
        }
        if (lease.s) {
            cudaStreamSynchronize(lease.s->stream);
            cudaStreamSynchronize(lease.s->copy_stream);   // deferred downloads of the last chunk (locate_host)
            lease.s->copy_pending = false;
            lease.s->reserve_n = 0;
        }
    };
    if (workers <= 1) {
        prologue();
        body();
    } else {
        std::vector<std::thread> th;
        for (int t = 1; t < workers; t++) th.emplace_back(body);
        prologue();
        body();
        for (auto& t : th) t.join();
    }
    if (first_err.load() != SVFM_OK) g_last_error = err_text;
    return first_err.load();
}

// Single-chunk batches (every single-pattern call among them) skip the upload stream and the worker threads: the
// copies go on the leased session's own stream.
static int upload_direct(svfm_session* s, const uint8_t* pats, const uint64_t* offs, uint64_t n, uint32_t fixed_len, uint32_t flags,
                         PatternBatch& pb) {
    int rc;
    pb.n = n;
    pb.fixed_len = fixed_len;
    pb.reversed = (flags & SVFM_REVERSED) ? 1u : 0u;
    pb.preencoded = 0;
    pb.offs = nullptr;
    const uint64_t bytes = offs ? offs[n] - offs[0] : n * (uint64_t)fixed_len;
    if ((rc = s->pats.reserve(bytes + 256))) return rc;
    SVFM_CUDA(cudaMemcpyAsync(s->pats.ptr, pats + (offs ? offs[0] : 0), bytes, cudaMemcpyHostToDevice, s->stream));
    pb.pats = (const uint8_t*)s->pats.ptr - (offs ? offs[0] : 0);  // offsets stay absolute
    if (offs) {
        if ((rc = s->offs.reserve((n + 1) * sizeof(uint64_t)))) return rc;
        SVFM_CUDA(cudaMemcpyAsync(s->offs.ptr, offs, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
        pb.offs = (const uint64_t*)s->offs.ptr;
    }
    return SVFM_OK;
}

// ---- small batches: one kernel launch + one synchronisation per call (search_kernels.cuh, small_batch_kernel) --------
// Arena layout (mapped pinned host memory, allocated on a session's first small call):
//   [0, 256 KiB) pattern bytes | offsets u64[4097] | counts P[4096] | slots P[4096 * 8] | status u8[4096]
constexpr uint64_t SMALL_OFF_OFFS = SMALL_MAX_BYTES;
constexpr uint64_t SMALL_OFF_CNT = SMALL_OFF_OFFS + (SMALL_MAX_PATTERNS + 1) * 8;
constexpr uint64_t SMALL_OFF_SLOTS = SMALL_OFF_CNT + SMALL_MAX_PATTERNS * 8;
constexpr uint64_t SMALL_OFF_STATUS = SMALL_OFF_SLOTS + SMALL_MAX_PATTERNS * SMALL_SLOTS * 8;
constexpr uint64_t SMALL_ARENA = SMALL_OFF_STATUS + SMALL_MAX_PATTERNS;
static std::atomic<uint64_t> g_small_max{[] {  // batches up to this many patterns take the one-launch path (0 = never)
    const char* e = std::getenv("SVFM_SMALL_MAX");
    return e ? (uint64_t)std::strtoull(e, nullptr, 10) : SMALL_MAX_PATTERNS;
}()};
static bool small_eligible(uint64_t n, uint64_t bytes, uint32_t flags) {
    const uint64_t lim = std::min<uint64_t>(g_small_max.load(), SMALL_MAX_PATTERNS);
    return n > 0 && n <= lim && bytes <= SMALL_MAX_BYTES;
}
// Runs the kernel and waits.  want_slots: locate.  Returns the arena through *arena.
static int small_run(svfm_session* s, const uint8_t* pats, const uint64_t* offs, uint64_t n, uint32_t fixed_len, uint32_t flags,
                     bool want_slots, bool preencoded, uint8_t** arena) {
    if (!s->h_small) {
        SVFM_CUDA(cudaHostAlloc((void**)&s->h_small, SMALL_ARENA, cudaHostAllocMapped));
    }
    uint8_t* a = s->h_small;
    const uint64_t first = offs ? offs[0] : 0;
    const uint64_t bytes = offs ? offs[n] - first : n * (uint64_t)fixed_len;
    std::memcpy(a, pats + first, bytes);
    PatternBatch pb;
    pb.pats = a - first;  // offsets stay absolute (the mapped arena has the same address on host and device under UVA)
    pb.offs = nullptr;
    if (offs) {
        std::memcpy(a + SMALL_OFF_OFFS, offs, (n + 1) * 8);
        pb.offs = reinterpret_cast<const uint64_t*>(a + SMALL_OFF_OFFS);
    }
    pb.n = n;
    pb.fixed_len = fixed_len;
    pb.reversed = (flags & SVFM_REVERSED) ? 1u : 0u;
    pb.preencoded = preencoded ? 1u : 0u;
    SmallOut out;
    out.cnt = a + SMALL_OFF_CNT;
    out.slots = want_slots ? a + SMALL_OFF_SLOTS : nullptr;
    out.slots_per = SMALL_SLOTS;
    out.status = a + SMALL_OFF_STATUS;
    const TypeOps* ops = type_ops(s->ix->type);
    if (!ops) return SVFM_ERR_BAD_TYPE;
    int rc = ops->small(s->ix->type.planes, s, pb, out);
    if (rc) return rc;
    SVFM_CUDA(cudaStreamSynchronize(s->stream));
    int bits = 0;
    for (uint64_t i = 0; i < n; i++) bits |= a[SMALL_OFF_STATUS + i];
    *arena = a;
    return err_from_bits(bits);
}

// Packed entry points: the host buffer holds `bpp` bytes per pattern (what the upload machinery sees as fixed_len); each
// chunk is expanded on the device to one symbol index per byte before the usual pipeline runs on it.
struct PackedSpec {
    uint32_t bits = 0;  // 0: not packed
    uint32_t len = 0;   // symbols per pattern
};
// Packed host patterns of one chunk.  When the chunk is going to take the plain search kernel anyway (no sweep search, no
// locality sort) and a CTA's patterns fit its stage, the kernel unpacks them itself while staging (PatternBatch::packed_bits)
// and nothing is expanded in HBM; otherwise unpack_patterns_kernel writes the one-byte-per-symbol form first.
static const bool g_packed_direct = [] { const char* e = std::getenv("SVFM_PACKED_DIRECT"); return e ? atoi(e) != 0 : true; }();
static int unpack_chunk(svfm_session* s, const PackedSpec& spec, PatternBatch& pb, bool for_locate) {
    if (!spec.bits) return SVFM_OK;
    const svfm_index* ix = s->ix;
    if (g_packed_direct && ix->ext_m != 0 && pb.n < sweep_min_patterns(ix, for_locate) && pb.n < sort_min_patterns(ix) &&
        (uint64_t)SEARCH_THREADS * spec.len + 4 <= (uint64_t)SEARCH_STAGE_MAX) {
        pb.packed_bits = spec.bits;
        pb.packed_bpp = pb.fixed_len;
        pb.fixed_len = spec.len;
        pb.preencoded = 1;
        return SVFM_OK;
    }
    const uint64_t bytes = pb.n * (uint64_t)spec.len;
    int rc;
    if ((rc = s->unpacked.reserve(rsv(s, pb.n) * (uint64_t)spec.len + 256))) return rc;
    PhaseTimer pt(s, SVFM_PHASE_PRESORT, 1);
    unpack_patterns_kernel<<<grid_for((bytes + 3) / 4, 256, s->ix->device), 256, 0, s->stream>>>(pb.pats, pb.n, spec.len, spec.bits, pb.fixed_len,
                                                                                             (uint8_t*)s->unpacked.ptr);
    SVFM_CUDA(cudaGetLastError());
    pb.pats = (const uint8_t*)s->unpacked.ptr;
    pb.fixed_len = spec.len;
    pb.preencoded = 1;
    return SVFM_OK;
}

static int count_host(svfm_index* ix, const uint8_t* pats, const uint64_t* offs, uint64_t n, uint32_t fixed_len,
                      uint32_t flags, void* counts_out, const PackedSpec& spec = PackedSpec()) {
    const uint64_t P = ix->type.pos_bits / 8;
    const ChunkPlan cp = plan_chunks(n, offs ? 0 : fixed_len);
    if (cp.chunks == 1) {
        SessionLease lease(ix);
        int r = lease.acquire();
        if (r) return r;
        svfm_session* s = lease.s;
        if (!spec.bits && small_eligible(n, offs ? offs[n] - offs[0] : n * (uint64_t)fixed_len, flags)) {
            uint8_t* a = nullptr;
            if ((r = small_run(s, pats, offs, n, fixed_len, flags, false, false, &a))) return r;
            std::memcpy(counts_out, a + SMALL_OFF_CNT, n * P);
            return SVFM_OK;
        }
        PatternBatch pb;
        if ((r = upload_direct(s, pats, offs, n, fixed_len, flags, pb))) return r;
        if ((r = unpack_chunk(s, spec, pb, false))) return r;
        if ((r = s->counts_out.reserve(n * P))) return r;
        if ((r = count_device(s, pb, s->counts_out.ptr))) return r;
        SVFM_CUDA(cudaMemcpyAsync(counts_out, s->counts_out.ptr, n * P, cudaMemcpyDeviceToHost, s->stream));
        SVFM_CUDA(cudaMemcpyAsync(&s->h_pinned[1], s->d_err, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        SVFM_CUDA(cudaStreamSynchronize(s->stream));
        return err_from_bits((int)(s->h_pinned[1] & 0xffffffffu));
    }
    UploaderLease up(ix);
    int rc = up.acquire();
    if (rc) return rc;
    for (uint64_t g0 = 0; g0 < cp.chunks;) {
        UploadGroup g;
        if ((rc = begin_group(up.u, cp, pats, offs, fixed_len, flags, g0, g))) return rc;
        rc = run_workers(ix, g.g0, g.g1, cp.largest(), [&](svfm_session* s, uint64_t c) -> int {
            if (!s) return SVFM_OK;
            const uint64_t a = cp.begin(c), b = cp.end(c);
            PatternBatch pb;
            int r;
            if ((r = await_chunk(s, cp, g, c, pb))) return r;
            if ((r = unpack_chunk(s, spec, pb, false))) return r;
            if ((r = s->counts_out.reserve(rsv(s, b - a) * P))) return r;  // not s->cnt: count_device uses that in work order
            if ((r = count_device(s, pb, s->counts_out.ptr))) return r;
            SVFM_CUDA(cudaMemcpyAsync((uint8_t*)counts_out + a * P, s->counts_out.ptr, (b - a) * P, cudaMemcpyDeviceToHost, s->stream));
            SVFM_CUDA(cudaMemcpyAsync(&s->h_pinned[1], s->d_err, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
            SVFM_CUDA(cudaStreamSynchronize(s->stream));
            return err_from_bits((int)(s->h_pinned[1] & 0xffffffffu));
        }, [&] { upload_group(cp, g); });
        cudaStreamSynchronize(up.u->stream);
        if (rc) return rc;
        g0 = g.g1;
    }
    return SVFM_OK;
}

static int locate_host(svfm_index* ix, const uint8_t* pats, const uint64_t* offs, uint64_t n, uint32_t fixed_len,
                       uint32_t flags, void* out_offs, void* positions, uint64_t capacity, void** alloc_out,
                       uint64_t* total_out, const PackedSpec& spec = PackedSpec()) {
    if (!ix || !out_offs || !total_out) return SVFM_ERR_BAD_ARG;
    *total_out = 0;
    if (alloc_out) *alloc_out = nullptr;
    uint64_t bytes = 0;
    int rc = check_host_patterns(pats, offs, n, fixed_len, &bytes);
    if (rc) return rc;
    const bool o32 = (flags & SVFM_OFFS32) != 0;   // out_offs is u32[n+1]
    const size_t OW = o32 ? sizeof(uint32_t) : sizeof(uint64_t);
    if (o32) static_cast<uint32_t*>(out_offs)[0] = 0; else static_cast<uint64_t*>(out_offs)[0] = 0;
    if (n == 0) return SVFM_OK;
    const uint64_t P = ix->type.pos_bits / 8;
    const ChunkPlan cp = plan_chunks(n, offs ? 0 : fixed_len);
    if (cp.chunks == 1) {  // no upload stream, no worker threads
        SessionLease lease(ix);
        if ((rc = lease.acquire())) return rc;
        svfm_session* s = lease.s;
        if (!spec.bits && small_eligible(n, bytes, flags)) {
            uint8_t* a = nullptr;
            if ((rc = small_run(s, pats, offs, n, fixed_len, flags, true, false, &a))) return rc;
            // counts + fixed slots -> CSR; a pattern with more rows than slots sends the batch through the general pipeline
            uint64_t total = 0;
            bool fits_slots = true;
            auto count_of = [&](uint64_t i) -> uint64_t {
                return P == 4 ? (uint64_t)reinterpret_cast<const uint32_t*>(a + SMALL_OFF_CNT)[i] : reinterpret_cast<const uint64_t*>(a + SMALL_OFF_CNT)[i];
            };
            for (uint64_t i = 0; i < n; i++) {
                const uint64_t c = count_of(i);
                fits_slots &= c <= SMALL_SLOTS;
                total += c;
            }
            if (fits_slots) {
                if (o32 && total > 0xffffffffull) return SVFM_ERR_TOO_LARGE;
                *total_out = total;
                uint64_t at = 0;
                for (uint64_t i = 0; i <= n; i++) {
                    if (o32) static_cast<uint32_t*>(out_offs)[i] = (uint32_t)at; else static_cast<uint64_t*>(out_offs)[i] = at;
                    if (i < n) at += count_of(i);
                }
                void* dst = positions;
                if (alloc_out && total && !(dst = result_alloc(total * P))) return SVFM_ERR_NOMEM;
                const bool fits = alloc_out || total <= capacity;
                if (fits && total) {
                    uint8_t* w = static_cast<uint8_t*>(dst);
                    for (uint64_t i = 0; i < n; i++) {
                        const uint64_t c = count_of(i);
                        if (!c) continue;
                        std::memcpy(w, a + SMALL_OFF_SLOTS + i * SMALL_SLOTS * P, c * P);
                        if (flags & SVFM_SORTED) {
                            if (P == 4) std::sort(reinterpret_cast<uint32_t*>(w), reinterpret_cast<uint32_t*>(w) + c);
                            else std::sort(reinterpret_cast<uint64_t*>(w), reinterpret_cast<uint64_t*>(w) + c);
                        }
                        w += c * P;
                    }
                }
                if (alloc_out) *alloc_out = total ? dst : nullptr;
                return fits ? SVFM_OK : SVFM_ERR_CAPACITY;
            }
        }
        PatternBatch pb;
        if ((rc = upload_direct(s, pats, offs, n, fixed_len, flags, pb))) return rc;
        if ((rc = unpack_chunk(s, spec, pb, true))) return rc;
        if ((rc = s->out_offs.reserve((n + 1) * sizeof(uint64_t)))) return rc;
        void* d_positions = nullptr;
        uint64_t total = 0;
        if ((rc = locate_device(s, pb, flags, s->out_offs.ptr, &d_positions, &total))) return rc;
        *total_out = total;
        SVFM_CUDA(cudaMemcpyAsync(out_offs, s->out_offs.ptr, (n + 1) * OW, cudaMemcpyDeviceToHost, s->stream));
        void* dst = positions;
        if (alloc_out && total) {
            if (!(dst = result_alloc(total * P))) { cudaStreamSynchronize(s->stream); return SVFM_ERR_NOMEM; }
        }
        const bool fits = alloc_out || total <= capacity;
        cudaError_t e = (total && fits) ? cudaMemcpyAsync(dst, d_positions, total * P, cudaMemcpyDeviceToHost, s->stream) : cudaSuccess;
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) {
            if (alloc_out && total) result_free(dst);
            SVFM_CUDA(e);
        }
        if (alloc_out) *alloc_out = total ? dst : nullptr;
        return fits ? SVFM_OK : SVFM_ERR_CAPACITY;
    }
    // chunk totals are published in completion order; a chunk's output base is the sum of all earlier totals
    std::mutex mu;
    std::condition_variable cv;
    std::vector<uint64_t> totals(cp.chunks, 0);
    std::vector<char> known(cp.chunks, 0);
    std::vector<void*> chunk_bufs(alloc_out ? cp.chunks : 0, nullptr);  // alloc mode: per-chunk pinned staging
    std::atomic<bool> overflow{false};
    const double t_begin = g_trace ? now_ms() : 0;
    auto publish = [&](uint64_t c, uint64_t t) {
        std::lock_guard<std::mutex> g(mu);
        if (!known[c]) { totals[c] = t; known[c] = 1; }
        cv.notify_all();
    };
    UploaderLease up(ix);
    if ((rc = up.acquire())) return rc;
    for (uint64_t g0 = 0; g0 < cp.chunks && rc == SVFM_OK;) {
    UploadGroup g;
    if ((rc = begin_group(up.u, cp, pats, offs, fixed_len, flags, g0, g))) break;
    g0 = g.g1;
    rc = run_workers(ix, g.g0, g.g1, cp.largest(), [&](svfm_session* s, uint64_t c) -> int {
        if (!s) { publish(c, 0); return SVFM_OK; }
        const uint64_t a = cp.begin(c), b = cp.end(c), m = b - a;
        PatternBatch pb;
        int r;
        const double t_0 = g_trace ? now_ms() : 0;
        if ((r = await_chunk(s, cp, g, c, pb))) return r;
        if (g_trace) { cudaStreamSynchronize(s->stream); }
        const double t_1 = g_trace ? now_ms() : 0;
        if ((r = unpack_chunk(s, spec, pb, true))) return r;
        if ((r = s->out_offs.reserve((rsv(s, m) + 1) * sizeof(uint64_t)))) return r;
        void* d_positions = nullptr;
        uint64_t total = 0;
        if ((r = locate_device(s, pb, flags, s->out_offs.ptr, &d_positions, &total))) return r;
        if (g_trace) { cudaStreamSynchronize(s->stream); }
        const double t_2 = g_trace ? now_ms() : 0;
        publish(c, total);
        uint64_t base = 0;
        {
            std::unique_lock<std::mutex> g(mu);
            cv.wait(g, [&] { for (uint64_t j = 0; j < c; j++) if (!known[j]) return false; return true; });
            for (uint64_t j = 0; j < c; j++) base += totals[j];
        }
        const bool last = (c + 1 == cp.chunks);
        const uint64_t n_offs = m + (last ? 1 : 0);
        if (o32 && base + total > 0xffffffffull) { g_last_error = "SVFM_OFFS32: 2^32 occurrences or more"; return SVFM_ERR_TOO_LARGE; }
        if (base) {
            if (o32) add_base_kernel<uint32_t><<<grid_for(n_offs, 256, ix->device), 256, 0, s->stream>>>((uint32_t*)s->out_offs.ptr, n_offs, (uint32_t)base);
            else add_base_kernel<uint64_t><<<grid_for(n_offs, 256, ix->device), 256, 0, s->stream>>>((uint64_t*)s->out_offs.ptr, n_offs, base);
            g_launches++;
            SVFM_CUDA(cudaGetLastError());
        }
        // downloads on the session's copy stream, behind this chunk's kernels; the worker goes on to its next chunk
        SVFM_CUDA(cudaEventRecord(s->ev_done, s->stream));
        SVFM_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev_done, 0));
        SVFM_CUDA(cudaMemcpyAsync((uint8_t*)out_offs + a * OW, s->out_offs.ptr, n_offs * OW, cudaMemcpyDeviceToHost, s->copy_stream));
        if (total) {
            if (alloc_out) {
                void* buf = result_alloc(total * P);
                if (!buf) { cudaStreamSynchronize(s->copy_stream); g_last_error = "result allocation failed"; return SVFM_ERR_NOMEM; }
                chunk_bufs[c] = buf;
                SVFM_CUDA(cudaMemcpyAsync(buf, d_positions, total * P, cudaMemcpyDeviceToHost, s->copy_stream));
            } else if (base + total <= capacity) {
                SVFM_CUDA(cudaMemcpyAsync((uint8_t*)positions + base * P, d_positions, total * P, cudaMemcpyDeviceToHost, s->copy_stream));
            } else {
                overflow.store(true);
            }
        }
        SVFM_CUDA(cudaEventRecord(s->ev_copied, s->copy_stream));
        s->copy_pending = true;
        if (g_trace) SVFM_CUDA(cudaStreamSynchronize(s->copy_stream));
        if (g_trace)
            std::fprintf(stderr, "[svfm trace] chunk %llu n=%llu: start %.2f  h2d %.2f  kernels %.2f  d2h %.2f ms\n",
                         (unsigned long long)c, (unsigned long long)m, t_0 - t_begin, t_1 - t_0, t_2 - t_1, now_ms() - t_2);
        return SVFM_OK;
    }, [&] { upload_group(cp, g); });
    cudaStreamSynchronize(up.u->stream);
    }
    uint64_t total = 0;
    for (uint64_t c = 0; c < cp.chunks; c++) total += totals[c];
    *total_out = total;
    if (alloc_out) {
        if (rc == SVFM_OK && total) {
            void* dst = nullptr;
            if (cp.chunks == 1) {
                dst = chunk_bufs[0];
                chunk_bufs[0] = nullptr;
            } else {
                if (!(dst = result_alloc(total * P))) rc = SVFM_ERR_NOMEM;
                uint64_t at = 0;
                for (uint64_t c = 0; dst && c < cp.chunks; c++) {
                    if (totals[c]) std::memcpy((uint8_t*)dst + at * P, chunk_bufs[c], totals[c] * P);
                    at += totals[c];
                }
            }
            *alloc_out = dst;
        }
        for (void* b : chunk_bufs) result_free(b);
    }
    if (rc) return rc;
    if (overflow.load() || (!alloc_out && total > capacity)) return SVFM_ERR_CAPACITY;
    return SVFM_OK;
}

}  // namespace svfm

using namespace svfm;

// ---------------------------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------------------------
extern "C" {

int svfm_load(const uint8_t* blob, size_t blob_len, svfm_type t, int device, svfm_index** out, uint64_t err_detail[2]) {
    return load_common(blob, blob_len, t, device, false, out, err_detail);
}

int svfm_load_device(const uint8_t* d_blob, size_t blob_len, svfm_type t, int device, svfm_index** out, uint64_t err_detail[2]) {
    return load_common(d_blob, blob_len, t, device, true, out, err_detail);
}

int svfm_check_blob(const uint8_t* blob, size_t blob_len, svfm_type t, svfm_info* out, uint64_t err_detail[2]) {
    if (!blob) return SVFM_ERR_BAD_ARG;
    Layout L;
    int rc = parse_blob([&](uint8_t* dst, uint64_t off, uint64_t n) { std::memcpy(dst, blob + off, n); return true; },
                        blob_len, t, L, err_detail);
    if (rc) return rc;
    if (out) {
        std::memset(out, 0, sizeof(*out));
        out->type = t;
        out->device = -1;
        out->symbol_count = L.symbol_count;
        out->kmer_size = L.kmer_size;
        out->sampling_ratio = L.sampling_ratio;
        const uint64_t P = t.pos_bits / 8;
        std::memcpy(&out->text_len, blob + L.off_count_array + (uint64_t)(L.count_array_len - 1) * P, P);
        std::memcpy(&out->sentinel_index, blob + L.off_sentinel_index, P);
        out->suffix_array_len = L.suffix_array_len;
        out->blocks_len = L.blocks_len;
        out->blob_len = blob_len;
        out->header_size = L.header_size;
        out->off_suffix_array = L.off_suffix_array;
        out->off_rank_checkpoints = L.off_rank_checkpoints;
        out->off_blocks = L.off_blocks;
    }
    return SVFM_OK;
}

void svfm_free(svfm_index* ix) {
    if (!ix) return;
    cudaSetDevice(ix->device);
    for (svfm_session* s : ix->pool) session_delete(s);
    ix->pool.clear();
    for (svfm_uploader* u : ix->up_pool) {
        if (u->stream) { cudaStreamSynchronize(u->stream); cudaStreamDestroy(u->stream); }
        for (cudaEvent_t e : u->ev) cudaEventDestroy(e);
        delete u;
    }
    ix->up_pool.clear();
    if (ix->d_swp) cudaFree(ix->d_swp);
    if (ix->d_fsa) cudaFree(ix->d_fsa);
    if (ix->d_text) cudaFree(ix->d_text);
    if (ix->d_ilv) cudaFree(ix->d_ilv);
    if (ix->d_ext) cudaFree(ix->d_ext);
    if (ix->d_alloc) cudaFree(ix->d_alloc);
    delete ix;
}

int svfm_index_info(const svfm_index* ix, svfm_info* out) {
    if (!ix || !out) return SVFM_ERR_BAD_ARG;
    std::memset(out, 0, sizeof(*out));
    out->type = ix->type;
    out->device = ix->device;
    out->symbol_count = ix->L.symbol_count;
    out->kmer_size = ix->L.kmer_size;
    out->sampling_ratio = ix->L.sampling_ratio;
    out->text_len = ix->text_len;
    out->suffix_array_len = ix->L.suffix_array_len;
    out->blocks_len = ix->L.blocks_len;
    out->sentinel_index = ix->sentinel_index;
    out->blob_len = ix->blob_len;
    out->header_size = ix->L.header_size;
    out->off_suffix_array = ix->L.off_suffix_array;
    out->off_rank_checkpoints = ix->L.off_rank_checkpoints;
    out->off_blocks = ix->L.off_blocks;
    return SVFM_OK;
}

int svfm_index_memory(svfm_index* ix, uint64_t out[8]) {
    if (!ix || !out) return SVFM_ERR_BAD_ARG;
    out[0] = ix->blob_len;
    out[1] = ix->d_ext ? ix->ext_entries * 2 * (ix->type.pos_bits / 8) : 0;
    out[2] = ix->d_ilv ? ix->L.blocks_len * (uint64_t)ix->ilv_stride : 0;
    out[4] = ix->d_text ? ix->text_bytes : 0;
    out[5] = ix->d_fsa ? ix->fsa_bytes : 0;
    out[6] = ix->d_swp ? ix->L.blocks_len * 32 : 0;
    out[7] = 0;
    uint64_t scratch = 0;
    std::lock_guard<std::mutex> g(ix->pool_mu);
    for (const svfm_session* s : ix->pool)
        for (const DeviceBuffer* b : {&s->pats, &s->offs, &s->unpacked, &s->sp, &s->cnt, &s->counts_out, &s->woffs, &s->out_offs, &s->offs64,
                                      &s->positions, &s->positions_alt, &s->cub_temp, &s->keys0, &s->keys1, &s->vals0, &s->vals1, &s->pay0,
                                      &s->pay1, &s->items0, &s->items1, &s->sweep_hist, &s->sweep_desc, &s->rec_key, &s->rec_key_alt, &s->first,
                                      &s->sb_hist, &s->sb_base, &s->sb_cursor, &s->sb_recs, &s->resolved,
                                      &s->heavy_sp, &s->heavy_cnt, &s->heavy_obase, &s->heavy_pat, &s->heavy_offs})
            scratch += b->cap;
    for (const svfm_uploader* u : ix->up_pool) scratch += u->pats.cap + u->offs.cap;
    out[3] = scratch;
    return SVFM_OK;
}

int svfm_count_batch(svfm_index* ix, const uint8_t* pats, const uint64_t* offs, uint64_t n, uint32_t fixed_len,
                     uint32_t flags, void* counts_out) {
    if (!ix || (n && !counts_out)) return SVFM_ERR_BAD_ARG;
    uint64_t bytes = 0;
    int rc = check_host_patterns(pats, offs, n, fixed_len, &bytes);
    if (rc) return rc;
    if (n == 0) return SVFM_OK;
    return count_host(ix, pats, offs, n, fixed_len, flags, counts_out);
}

int svfm_locate_batch(svfm_index* ix, const uint8_t* pats, const uint64_t* offs, uint64_t n, uint32_t fixed_len,
                      uint32_t flags, void* out_offs, void* positions, uint64_t capacity, uint64_t* total) {
    if (capacity && !positions) return SVFM_ERR_BAD_ARG;
    return locate_host(ix, pats, offs, n, fixed_len, flags, out_offs, positions, capacity, nullptr, total);
}

int svfm_locate_batch_alloc(svfm_index* ix, const uint8_t* pats, const uint64_t* offs, uint64_t n, uint32_t fixed_len,
                            uint32_t flags, void* out_offs, void** positions, uint64_t* total) {
    if (!positions) return SVFM_ERR_BAD_ARG;
    return locate_host(ix, pats, offs, n, fixed_len, flags, out_offs, nullptr, 0, positions, total);
}

static int packed_spec(uint32_t len, uint32_t bits, uint32_t flags, PackedSpec* spec, uint32_t* bpp) {
    if (len == 0) return SVFM_ERR_EMPTY_PATTERN;
    if (bits < 1 || bits > 8 || (flags & SVFM_REVERSED)) return SVFM_ERR_BAD_ARG;
    const uint64_t b = ((uint64_t)len * bits + 7) / 8;
    if (b > 0xffffffffull) return SVFM_ERR_BAD_ARG;
    spec->bits = bits;
    spec->len = len;
    *bpp = (uint32_t)b;
    return SVFM_OK;
}

int svfm_count_batch_packed(svfm_index* ix, const uint8_t* packed, uint64_t n, uint32_t len, uint32_t bits, uint32_t flags,
                            void* counts_out) {
    if (!ix || (n && (!counts_out || !packed))) return SVFM_ERR_BAD_ARG;
    PackedSpec spec;
    uint32_t bpp = 0;
    int rc = packed_spec(len, bits, flags, &spec, &bpp);
    if (rc) return rc;
    if (n == 0) return SVFM_OK;
    return count_host(ix, packed, nullptr, n, bpp, flags, counts_out, spec);
}

int svfm_locate_batch_packed(svfm_index* ix, const uint8_t* packed, uint64_t n, uint32_t len, uint32_t bits, uint32_t flags,
                             void* out_offs, void* positions, uint64_t capacity, uint64_t* total) {
    if (capacity && !positions) return SVFM_ERR_BAD_ARG;
    PackedSpec spec;
    uint32_t bpp = 0;
    int rc = packed_spec(len, bits, flags, &spec, &bpp);
    if (rc) return rc;
    return locate_host(ix, packed, nullptr, n, bpp, flags, out_offs, positions, capacity, nullptr, total, spec);
}

int svfm_pack_patterns(const uint8_t* pats, uint64_t n, uint32_t len, const uint8_t* table256, uint32_t bits, uint8_t* packed_out) {
    if ((n && (!pats || !packed_out)) || bits < 1 || bits > 8 || len == 0) return SVFM_ERR_BAD_ARG;
    const uint64_t bpp = ((uint64_t)len * bits + 7) / 8;
    const uint32_t limit = 1u << bits;
    unsigned hw = std::thread::hardware_concurrency();
    const unsigned nt = (unsigned)std::min<uint64_t>(hw ? hw : 1, (n + 65535) / 65536 ? (n + 65535) / 65536 : 1);
    std::atomic<int> bad{0};
    auto work = [&](uint64_t a, uint64_t b) {
        for (uint64_t i = a; i < b; i++) {
            const uint8_t* p = pats + i * (uint64_t)len;
            uint8_t* o = packed_out + i * bpp;
            std::memset(o, 0, bpp);
            for (uint32_t j = 0; j < len; j++) {
                const uint32_t s = table256 ? table256[p[j]] : p[j];
                if (s >= limit) { bad.store(1); continue; }
                const uint32_t bit = j * bits;
                o[bit >> 3] |= (uint8_t)(s << (bit & 7u));
                if ((bit & 7u) + bits > 8u) o[(bit >> 3) + 1] |= (uint8_t)(s >> (8u - (bit & 7u)));
            }
        }
    };
    std::vector<std::thread> th;
    const uint64_t per = (n + nt - 1) / nt;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work, std::min(n, t * per), std::min(n, (t + 1) * per));
    work(0, std::min(n, per));
    for (auto& t : th) t.join();
    return bad.load() ? SVFM_ERR_BAD_SYMBOL : SVFM_OK;
}

void svfm_free_positions(void* positions) {
    result_free(positions);
}

int svfm_count(svfm_index* ix, const uint8_t* pattern, uint64_t len, uint32_t flags, uint64_t* count) {
    if (!ix || !count) return SVFM_ERR_BAD_ARG;
    if (len == 0) return SVFM_ERR_EMPTY_PATTERN;
    if (len > 0xffffffffull) return SVFM_ERR_BAD_ARG;
    uint64_t c = 0;  // P-sized write into a zeroed u64 (little-endian)
    int rc = svfm_count_batch(ix, pattern, nullptr, 1, (uint32_t)len, flags, &c);
    if (rc) return rc;
    *count = c;
    return SVFM_OK;
}

int svfm_locate(svfm_index* ix, const uint8_t* pattern, uint64_t len, uint32_t flags, void* positions,
                uint64_t capacity, uint64_t* total) {
    if (!ix || !total) return SVFM_ERR_BAD_ARG;
    if (len == 0) return SVFM_ERR_EMPTY_PATTERN;
    if (len > 0xffffffffull) return SVFM_ERR_BAD_ARG;
    uint64_t offs2[2];
    return svfm_locate_batch(ix, pattern, nullptr, 1, (uint32_t)len, flags & ~(uint32_t)SVFM_OFFS32, offs2, positions, capacity, total);
}

int svfm_session_create(svfm_index* ix, svfm_session** out) {
    if (!ix || !out) return SVFM_ERR_BAD_ARG;
    return session_new(ix, out);
}

void svfm_session_destroy(svfm_session* s) { session_delete(s); }

int svfm_session_sync(svfm_session* s) {
    if (!s) return SVFM_ERR_BAD_ARG;
    SVFM_CUDA(cudaStreamSynchronize(s->stream));
    // status of the device-side pattern checks of the last batch enqueued on this session
    const int rc = err_from_bits((int)(s->h_pinned[1] & 0xffffffffu));
    s->h_pinned[1] = 0;
    return rc;
}

void* svfm_session_stream(svfm_session* s) { return s ? (void*)s->stream : nullptr; }

int svfm_session_set_timing(svfm_session* s, int enabled) {
    if (!s) return SVFM_ERR_BAD_ARG;
    s->timing = enabled != 0;
    return SVFM_OK;
}

int svfm_session_get_timing(svfm_session* s, double ms[SVFM_PHASE_MAX], uint64_t launches[SVFM_PHASE_MAX], int reset) {
    if (!s) return SVFM_ERR_BAD_ARG;
    SVFM_CUDA(cudaStreamSynchronize(s->stream));
    collect_spans(s);
    for (int i = 0; i < SVFM_PHASE_MAX; i++) {
        if (ms) ms[i] = s->phase_ms[i];
        if (launches) launches[i] = s->phase_launches[i];
        if (reset) { s->phase_ms[i] = 0; s->phase_launches[i] = 0; }
    }
    return SVFM_OK;
}

int svfm_count_batch_device(svfm_session* s, const uint8_t* d_pats, const uint64_t* d_offs, uint64_t n,
                            uint32_t fixed_len, uint32_t flags, void* d_counts_out) {
    if (!s || (n && (!d_pats || !d_counts_out))) return SVFM_ERR_BAD_ARG;
    if (!d_offs && fixed_len == 0 && n) return SVFM_ERR_EMPTY_PATTERN;
    SVFM_CUDA(cudaSetDevice(s->ix->device));
    PatternBatch pb{d_pats, d_offs, n, fixed_len, (flags & SVFM_REVERSED) ? 1u : 0u, 0u};
    int rc = count_device(s, pb, d_counts_out);
    if (rc) return rc;
    // the error word is read back asynchronously; svfm_session_sync + the next call report it
    SVFM_CUDA(cudaMemcpyAsync(&s->h_pinned[1], s->d_err, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    return SVFM_OK;
}

int svfm_locate_batch_device(svfm_session* s, const uint8_t* d_pats, const uint64_t* d_offs, uint64_t n,
                             uint32_t fixed_len, uint32_t flags, void* d_out_offs, void** d_positions,
                             uint64_t* total) {
    if (!s || !d_out_offs || !d_positions || !total || (n && !d_pats)) return SVFM_ERR_BAD_ARG;
    if (!d_offs && fixed_len == 0 && n) return SVFM_ERR_EMPTY_PATTERN;
    SVFM_CUDA(cudaSetDevice(s->ix->device));
    *total = 0;
    *d_positions = nullptr;
    if (n == 0) {
        SVFM_CUDA(cudaMemsetAsync(d_out_offs, 0, (flags & SVFM_OFFS32) ? sizeof(uint32_t) : sizeof(uint64_t), s->stream));
        return SVFM_OK;
    }
    PatternBatch pb{d_pats, d_offs, n, fixed_len, (flags & SVFM_REVERSED) ? 1u : 0u, 0u};
    return locate_device(s, pb, flags, d_out_offs, d_positions, total);
}

void* svfm_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void svfm_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int svfm_set_tuning(int key, uint64_t value) {
    switch (key) {
        case SVFM_TUNE_SORT_MIN: g_sort_min.store(value); return SVFM_OK;
        case SVFM_TUNE_CHUNK: g_chunk_patterns.store(value); return SVFM_OK;
        case SVFM_TUNE_SWEEP_MIN: g_sweep_min.store(value); return SVFM_OK;
        case SVFM_TUNE_EXT_BITS: g_ext_bits.store(value); return SVFM_OK;
        case SVFM_TUNE_WORKERS: g_host_workers.store(value); return SVFM_OK;
        case SVFM_TUNE_ILV: g_ilv.store(value); return SVFM_OK;
        case SVFM_TUNE_BUCKET_SORTBACK: g_bucket_sortback.store(value); return SVFM_OK;
        case SVFM_TUNE_SMALL_MAX: g_small_max.store(value); return SVFM_OK;
        case SVFM_TUNE_TEXT: g_text.store(value); return SVFM_OK;
        case SVFM_TUNE_FULL_SA: g_full_sa.store(value); return SVFM_OK;
        case SVFM_TUNE_SWEEP_OCC: g_sweep_occ.store(value); return SVFM_OK;
        case SVFM_TUNE_L2_PERSIST: g_l2_persist.store(value); return SVFM_OK;
        case SVFM_TUNE_OWN_RADIX: g_own_radix.store(value); return SVFM_OK;
        default: return SVFM_ERR_BAD_ARG;
    }
}

const char* svfm_last_error(void) { return g_last_error.c_str(); }
uint64_t svfm_launch_count(void) { return g_launches.load(); }
const char* svfm_version(void) { return "sview-fmindex-b200 0.1.0 (sm_100a)"; }

}  // extern "C"
