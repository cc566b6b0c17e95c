// builder.cu -- index construction on the GPU ("next" row, not the hot path).
//
// Produces, in device memory, the same bytes as the reference's FmIndexBuilder::build
// (builder/mod.rs:187-264): headers, count array + kLTS (count_array.rs:78-137), sampled suffix array
// (suffix_array/mod.rs:57-70 with the sentinel/pidx convention of
// suffix_array/burrow_wheeler_transform/crate_bio_manual/mod.rs:8-25) and rank checkpoints + bit-plane
// blocks (bwm/mod.rs:91-143, blocks/block3.rs:19-40).  The blob is a pure function of
// (text, encoder, P, B, k, r): the suffix array with a unique smallest sentinel is unique, so any correct
// suffix sorter yields identical bytes.  Suffix sorting here = one radix sort on packed symbol prefixes +
// prefix doubling on the (rare) ties; it exists so that 1-3 Gbp benchmark indexes can be built inside a
// GPU job in seconds instead of minutes of host SA-IS.  Limit: text_len < 2^32 - 1.
// Citations are relative to the reference's sview-fmindex/src/.
#include <cstring>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "blob_layout.h"
#include "common.cuh"

namespace svfm {
namespace {

constexpr int TPB = 256;

inline unsigned grid1d(uint64_t n, int tpb = TPB) { return (unsigned)((n + tpb - 1) / tpb); }

// ---- text encoding + counting ----------------------------------------------------------------------
// text[i] -> symidx + 1 (sentinel will be 0 for sorting), count_array.rs:113-116.
__global__ void encode_kernel(const uint8_t* __restrict__ text, uint64_t n, const uint8_t* __restrict__ table,
                              uint8_t* __restrict__ enc, uint32_t S, int* __restrict__ bad) {
    __shared__ uint8_t s_table[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_table[i] = table ? table[i] : (uint8_t)i;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t s = s_table[text[i]];
        if (s >= S) { *bad = 1; s = S - 1; }
        enc[i] = (uint8_t)(s + 1);
    }
}

// Histogram of k-mer table indices (count_array.rs:118-123): index_i = sum_j enc[i+j] * (S+1)^(k-1-j),
// symbols past the end of the text count as 0.  With k == 1 this is the symbol histogram.
__global__ void kmer_hist_kernel(const uint8_t* __restrict__ enc, uint64_t n, uint32_t k, uint32_t swsc,
                                 uint64_t table_len, unsigned long long* __restrict__ hist, int use_smem) {
    extern __shared__ uint32_t s_hist[];
    if (use_smem) {
        for (uint64_t i = threadIdx.x; i < table_len; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t idx = 0;
        for (uint32_t j = 0; j < k; j++) idx = idx * swsc + (i + j < n ? enc[i + j] : 0);
        if (use_smem) atomicAdd(&s_hist[idx], 1u);
        else atomicAdd(&hist[idx], 1ull);
    }
    if (use_smem) {
        __syncthreads();
        for (uint64_t i = threadIdx.x; i < table_len; i += blockDim.x)
            if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
    }
}

template <class P>
__global__ void narrow_kernel(const unsigned long long* __restrict__ in, P* __restrict__ out, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = (P)in[i];
}

// ---- suffix sorting ----------------------------------------------------------------------------------
// key_i = the first m symbols of suffix i packed MSB-first, b bits each; symbols past the end are 0, so a
// shorter suffix sorts before any longer one with the same prefix (the sentinel is the smallest symbol).
__global__ void pack_keys_kernel(const uint8_t* __restrict__ enc, uint64_t n, uint32_t b, uint32_t m,
                                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t key = 0;
        for (uint32_t j = 0; j < m; j++) key = (key << b) | (uint64_t)(i + j < n ? __ldg(enc + i + j) : 0);
        keys[i] = key;
        vals[i] = (uint32_t)i;
    }
}

struct HeadOrZero {
    const uint64_t* keys;
    __host__ __device__ uint32_t operator()(uint64_t i) const { return (i == 0 || keys[i] != keys[i - 1]) ? (uint32_t)i : 0u; }
};
struct MaxOp {
    __host__ __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

// isa1[sa[i]] = rank[i] + 1 (0 is reserved for "past the end of the text")
__global__ void scatter_isa_kernel(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ rank, uint64_t n,
                                   uint32_t* __restrict__ isa1) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        isa1[sa[i]] = rank[i] + 1;
}

// position i is tied when its rank group has more than one member
struct TiedAt {
    const uint32_t* rank;
    uint64_t n;
    __host__ __device__ bool operator()(uint64_t i) const {
        const bool head = rank[i] == (uint32_t)i;
        const bool next_head = (i + 1 == n) || rank[i + 1] == (uint32_t)(i + 1);
        return !(head && next_head);
    }
};

// key2 = (group rank << 32) | rank1 of the suffix h symbols further on (0 past the end)
__global__ void tie_keys_kernel(const uint32_t* __restrict__ tied_pos, uint64_t nt, const uint32_t* __restrict__ sa,
                                const uint32_t* __restrict__ rank, const uint32_t* __restrict__ isa1, uint64_t n,
                                uint64_t h, uint64_t* __restrict__ key2, uint32_t* __restrict__ val2) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t pos = tied_pos[t];
        const uint32_t suf = sa[pos];
        const uint64_t nxt = (uint64_t)suf + h;
        const uint32_t r1 = nxt < n ? isa1[nxt] : 0u;
        key2[t] = ((uint64_t)rank[pos] << 32) | r1;
        val2[t] = suf;
    }
}

struct Head2OrZero {
    const uint64_t* key2;
    const uint32_t* tied_pos;
    __host__ __device__ uint32_t operator()(uint64_t t) const {
        return (t == 0 || key2[t] != key2[t - 1]) ? tied_pos[t] : 0u;
    }
};

// write the refined order back: sorted element t goes to SA position tied_pos[t]
__global__ void tie_writeback_kernel(const uint32_t* __restrict__ tied_pos, uint64_t nt, const uint32_t* __restrict__ val2,
                                     const uint32_t* __restrict__ newrank, uint32_t* __restrict__ sa,
                                     uint32_t* __restrict__ rank, uint32_t* __restrict__ isa1) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t pos = tied_pos[t];
        sa[pos] = val2[t];
        rank[pos] = newrank[t];
        isa1[val2[t]] = newrank[t] + 1;
    }
}

struct StillTied {
    const uint32_t* newrank;
    const uint32_t* tied_pos;
    uint64_t nt;
    __host__ __device__ bool operator()(uint64_t t) const {
        const bool head = newrank[t] == tied_pos[t];
        const bool next_head = (t + 1 == nt) || newrank[t + 1] == tied_pos[t + 1];
        return !(head && next_head);
    }
};
struct PickPos {
    const uint32_t* tied_pos;
    __host__ __device__ uint32_t operator()(uint64_t t) const { return tied_pos[t]; }
};

// ---- BWT, sampled SA, blocks, checkpoints --------------------------------------------------------------
// sa = suffix array WITHOUT the sentinel row (n rows).  The full BWT has n+1 rows (row 0 = the sentinel
// suffix); pidx = row whose BWT symbol is the sentinel = (row i with sa[i] == 0) + 1; the stored BWT drops
// that row (crate_bio_manual/mod.rs:14-21).  stored[m] for m in [0, n):
//   m == 0      -> enc[n-1];   1 <= m < pidx -> enc[sa[m-1] - 1];   m >= pidx -> enc[sa[m] - 1]
__global__ void find_pidx_kernel(const uint32_t* __restrict__ sa, uint64_t n, unsigned long long* __restrict__ pidx) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (sa[i] == 0) *pidx = i + 1;
}

__global__ void bwt_kernel(const uint32_t* __restrict__ sa, const uint8_t* __restrict__ enc, uint64_t n,
                           const unsigned long long* __restrict__ pidx_p, uint8_t* __restrict__ bwt) {
    const uint64_t pidx = *pidx_p;
    for (uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; m < n; m += (uint64_t)gridDim.x * blockDim.x) {
        uint8_t c;
        if (m == 0) c = enc[n - 1];
        else {
            const uint32_t s = m < pidx ? sa[m - 1] : sa[m];
            c = __ldg(enc + (s - 1));
        }
        bwt[m] = c;
    }
}

// every ratio-th row of the sentinel-free SA (suffix_array/mod.rs:67, crate_bio_manual/mod.rs:23)
template <class P>
__global__ void sample_sa_kernel(const uint32_t* __restrict__ sa, uint64_t n, uint32_t ratio, P* __restrict__ out,
                                 uint64_t out_len) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < out_len; j += (uint64_t)gridDim.x * blockDim.x)
        out[j] = (P)sa[j * ratio];
}

// Block::vectorize (blocks/block3.rs:19-36) for one block per thread: symbol j of the chunk is bit
// VBITS-1-j of each plane (a partial last chunk ends up left-shifted, bwm/mod.rs:139-142), and the
// per-block symbol counts go to this block's own checkpoint row (turned into prefix sums afterwards).
template <class P, class W, int WORDS>
__global__ void blocks_kernel(const uint8_t* __restrict__ bwt, uint64_t n, uint32_t S, uint32_t planes, uint32_t vbits,
                              uint64_t blocks_len, W* __restrict__ blocks, P* __restrict__ ck) {
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < blocks_len; q += (uint64_t)gridDim.x * blockDim.x) {
        W pl[6][WORDS];
        for (uint32_t v = 0; v < 6; v++)
            for (int k = 0; k < WORDS; k++) pl[v][k] = 0;
        P* row = ck + q * S;
        for (uint32_t s = 0; s < S; s++) row[s] = 0;
        const uint64_t start = q * vbits;
        const uint32_t wbits = vbits / WORDS;
        for (uint32_t j = 0; j < vbits; j++) {
            if (start + j >= n) break;
            const uint32_t symidx = (uint32_t)bwt[start + j] - 1u;
            row[symidx] += 1;
            const int k = j / wbits;  // word 0 = first symbols
            const W bit = (W)1 << (wbits - 1 - (j % wbits));
            for (uint32_t v = 0; v < planes; v++)
                if ((symidx >> v) & 1u) pl[v][k] |= bit;
        }
        W* dst = blocks + q * (uint64_t)(planes * WORDS);
        for (uint32_t v = 0; v < planes; v++) {
            if (WORDS == 1) dst[v] = pl[v][0];
            else { dst[v * 2 + 0] = pl[v][1]; dst[v * 2 + 1] = pl[v][0]; }  // u128 little-endian: low half first
        }
    }
}

// Exclusive prefix sums down every checkpoint column (bwm/mod.rs:126-131,137): three phases over chunks
// of CK_CHUNK rows; thread s of a CTA owns symbol s.
constexpr int CK_CHUNK = 512;

template <class P>
__global__ void ck_chunk_totals_kernel(const P* __restrict__ ck, uint64_t rows, uint32_t S, unsigned long long* __restrict__ tot) {
    const uint64_t c = blockIdx.x;
    const uint32_t s = threadIdx.x;
    if (s >= S) return;
    const uint64_t r0 = c * CK_CHUNK, r1 = r0 + CK_CHUNK < rows ? r0 + CK_CHUNK : rows;
    unsigned long long acc = 0;
    for (uint64_t r = r0; r < r1; r++) acc += ck[r * S + s];
    tot[c * S + s] = acc;
}

__global__ void ck_scan_totals_kernel(unsigned long long* __restrict__ tot, uint64_t chunks, uint32_t S) {
    const uint32_t s = threadIdx.x;
    if (s >= S) return;
    unsigned long long acc = 0;
    for (uint64_t c = 0; c < chunks; c++) {
        const unsigned long long v = tot[c * S + s];
        tot[c * S + s] = acc;
        acc += v;
    }
}

template <class P>
__global__ void ck_apply_kernel(P* __restrict__ ck, uint64_t rows, uint32_t S, const unsigned long long* __restrict__ tot) {
    const uint64_t c = blockIdx.x;
    const uint32_t s = threadIdx.x;
    if (s >= S) return;
    const uint64_t r0 = c * CK_CHUNK, r1 = r0 + CK_CHUNK < rows ? r0 + CK_CHUNK : rows;
    unsigned long long acc = tot[c * S + s];
    for (uint64_t r = r0; r < r1; r++) {
        const P v = ck[r * S + s];
        ck[r * S + s] = (P)acc;
        acc += v;
    }
}

struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) cudaFree(p); }
    template <class T>
    int alloc(T** out, uint64_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T) + 256);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            g_last_error = std::string("cudaMalloc(builder scratch): ") + cudaGetErrorString(e);
            return e == cudaErrorMemoryAllocation ? SVFM_ERR_NOMEM : SVFM_ERR_CUDA;
        }
        ptrs.push_back(p);
        *out = (T*)p;
        return SVFM_OK;
    }
    void free_one(void* p) {
        for (auto& q : ptrs) if (q == p) { cudaFree(q); q = ptrs.back(); ptrs.pop_back(); return; }
    }
};

#define SVFM_TRY(expr) do { int rc_ = (expr); if (rc_) return rc_; } while (0)

// Suffix array of enc[0..n) (symbols 1..S, implicit smallest sentinel at n) into d_sa (n rows).
int suffix_sort(const uint8_t* d_enc, uint64_t n, uint32_t S, uint32_t* d_sa, Scratch& sc, cudaStream_t st) {
    uint32_t b = 1;
    while ((1u << b) < S + 1) b++;
    const uint32_t m = 64 / b;
    uint64_t *k0, *k1;
    uint32_t* v1;
    SVFM_TRY(sc.alloc(&k0, n));
    SVFM_TRY(sc.alloc(&k1, n));
    SVFM_TRY(sc.alloc(&v1, n));
    const unsigned g = grid1d(n) > 148u * 32u ? 148u * 32u : grid1d(n);
    pack_keys_kernel<<<g, TPB, 0, st>>>(d_enc, n, b, m, k0, d_sa);
    g_launches++;
    cub::DoubleBuffer<uint64_t> keys(k0, k1);
    cub::DoubleBuffer<uint32_t> vals(d_sa, v1);
    size_t temp = 0;
    SVFM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp, keys, vals, (int64_t)n, 0, (int)(m * b), st));
    uint8_t* d_temp;
    SVFM_TRY(sc.alloc(&d_temp, temp));
    SVFM_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, temp, keys, vals, (int64_t)n, 0, (int)(m * b), st));
    g_launches += 8;
    if (vals.Current() != d_sa) SVFM_CUDA(cudaMemcpyAsync(d_sa, vals.Current(), n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    // rank[i] = index of the head of i's group of equal keys; isa1 = inverse + 1
    uint32_t* rank = reinterpret_cast<uint32_t*>(keys.Alternate());      // 8n bytes: rank (4n) + isa1 (4n)
    uint32_t* isa1 = rank + n;
    {
        HeadOrZero f{keys.Current()};
        auto in = thrust::make_transform_iterator(thrust::counting_iterator<uint64_t>(0), f);
        size_t t2 = 0;
        SVFM_CUDA(cub::DeviceScan::InclusiveScan(nullptr, t2, in, rank, MaxOp(), (int64_t)n, st));
        if (t2 > temp) { SVFM_TRY(sc.alloc(&d_temp, t2)); temp = t2; }
        SVFM_CUDA(cub::DeviceScan::InclusiveScan(d_temp, t2, in, rank, MaxOp(), (int64_t)n, st));
        g_launches += 2;
    }
    // tied positions
    uint32_t* tied = v1;  // n entries available
    unsigned long long* d_nt;
    SVFM_TRY(sc.alloc(&d_nt, 1));
    {
        TiedAt f{rank, n};
        auto flags = thrust::make_transform_iterator(thrust::counting_iterator<uint64_t>(0), f);
        auto idx = thrust::counting_iterator<uint32_t>(0);
        size_t t2 = 0;
        SVFM_CUDA(cub::DeviceSelect::Flagged(nullptr, t2, idx, flags, tied, d_nt, (int64_t)n, st));
        if (t2 > temp) { SVFM_TRY(sc.alloc(&d_temp, t2)); temp = t2; }
        SVFM_CUDA(cub::DeviceSelect::Flagged(d_temp, t2, idx, flags, tied, d_nt, (int64_t)n, st));
        g_launches += 2;
    }
    unsigned long long nt = 0;
    SVFM_CUDA(cudaMemcpyAsync(&nt, d_nt, sizeof(nt), cudaMemcpyDeviceToHost, st));
    SVFM_CUDA(cudaStreamSynchronize(st));
    if (nt == 0) return SVFM_OK;
    scatter_isa_kernel<<<g, TPB, 0, st>>>(d_sa, rank, n, isa1);
    g_launches++;
    // prefix doubling on the tied positions only (Larsson-Sadakane style): after a round with offset h the
    // ranks order suffixes by their first 2h symbols.
    uint64_t *key2 = keys.Current(), *key2b;  // the sorted prefix keys are no longer needed
    uint32_t *val2, *val2b, *newrank, *tied_next;
    SVFM_TRY(sc.alloc(&key2b, nt));
    SVFM_TRY(sc.alloc(&val2, nt));
    SVFM_TRY(sc.alloc(&val2b, nt));
    SVFM_TRY(sc.alloc(&newrank, nt));
    SVFM_TRY(sc.alloc(&tied_next, nt));
    for (uint64_t h = m; nt > 0; h *= 2) {
        const unsigned gt = grid1d(nt) > 148u * 32u ? 148u * 32u : grid1d(nt);
        tie_keys_kernel<<<gt, TPB, 0, st>>>(tied, nt, d_sa, rank, isa1, n, h, key2, val2);
        g_launches++;
        cub::DoubleBuffer<uint64_t> kk(key2, key2b);
        cub::DoubleBuffer<uint32_t> vv(val2, val2b);
        size_t t2 = 0;
        SVFM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t2, kk, vv, (int64_t)nt, 0, 64, st));
        if (t2 > temp) { SVFM_TRY(sc.alloc(&d_temp, t2)); temp = t2; }
        SVFM_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, t2, kk, vv, (int64_t)nt, 0, 64, st));
        g_launches += 8;
        {
            Head2OrZero f{kk.Current(), tied};
            auto in = thrust::make_transform_iterator(thrust::counting_iterator<uint64_t>(0), f);
            SVFM_CUDA(cub::DeviceScan::InclusiveScan(nullptr, t2, in, newrank, MaxOp(), (int64_t)nt, st));
            if (t2 > temp) { SVFM_TRY(sc.alloc(&d_temp, t2)); temp = t2; }
            SVFM_CUDA(cub::DeviceScan::InclusiveScan(d_temp, t2, in, newrank, MaxOp(), (int64_t)nt, st));
            g_launches += 2;
        }
        tie_writeback_kernel<<<gt, TPB, 0, st>>>(tied, nt, vv.Current(), newrank, d_sa, rank, isa1);
        g_launches++;
        {
            StillTied f{newrank, tied, nt};
            auto flags = thrust::make_transform_iterator(thrust::counting_iterator<uint64_t>(0), f);
            auto src = thrust::make_transform_iterator(thrust::counting_iterator<uint64_t>(0), PickPos{tied});
            SVFM_CUDA(cub::DeviceSelect::Flagged(nullptr, t2, src, flags, tied_next, d_nt, (int64_t)nt, st));
            if (t2 > temp) { SVFM_TRY(sc.alloc(&d_temp, t2)); temp = t2; }
            SVFM_CUDA(cub::DeviceSelect::Flagged(d_temp, t2, src, flags, tied_next, d_nt, (int64_t)nt, st));
            g_launches += 2;
        }
        SVFM_CUDA(cudaMemcpyAsync(&nt, d_nt, sizeof(nt), cudaMemcpyDeviceToHost, st));
        SVFM_CUDA(cudaStreamSynchronize(st));
        std::swap(tied, tied_next);
        if (h > n) break;
    }
    return SVFM_OK;
}

template <class P>
int build_typed(const svfm_type& t, const Layout& L, const uint8_t* d_text, uint64_t n, uint32_t S,
                const uint8_t* table256, uint8_t* d_blob, cudaStream_t st) {
    Scratch sc;
    const uint32_t k = L.kmer_size;
    const uint32_t swsc = S + 1;
    SVFM_CUDA(cudaMemsetAsync(d_blob, 0, L.total_size, st));
    // 1) headers (builder/mod.rs:211-231)
    {
        std::vector<uint8_t> h(L.header_size, 0);
        const uint8_t magic[8] = {'F', 'I', '0', '0', 0, 0, 0, 0};
        std::memcpy(h.data(), magic, 8);
        if (t.encoder) std::memcpy(h.data() + L.off_encoder, table256, 256);
        uint8_t* p = h.data() + L.off_count_header;
        std::memcpy(p + 0, &L.symbol_count, 4);
        std::memcpy(p + 4, &L.kmer_size, 4);
        std::memcpy(p + 8, &L.count_array_len, 4);
        std::memcpy(p + 12, &L.kmer_multiplier_len, 4);
        std::memcpy(p + 16, &L.kmer_count_table_len, 8);
        p = h.data() + L.off_sa_header;
        std::memcpy(p + 0, &L.sampling_ratio, 4);
        std::memcpy(p + 8, &L.suffix_array_len, 8);
        p = h.data() + L.off_bwm_header;
        std::memcpy(p + 0, &L.symbol_count, 4);
        std::memcpy(p + 8, &L.rank_checkpoints_len, 8);
        std::memcpy(p + 16, &L.blocks_len, 8);
        SVFM_CUDA(cudaMemcpyAsync(d_blob, h.data(), h.size(), cudaMemcpyHostToDevice, st));
        SVFM_CUDA(cudaStreamSynchronize(st));
    }
    // 2) encode the text, count symbols and k-mers (count_array.rs:78-137)
    uint8_t *d_enc, *d_table = nullptr;
    int* d_bad;
    SVFM_TRY(sc.alloc(&d_enc, n + 64));
    SVFM_TRY(sc.alloc(&d_bad, 1));
    SVFM_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    SVFM_CUDA(cudaMemsetAsync(d_enc + n, 0, 64, st));
    if (t.encoder) {
        SVFM_TRY(sc.alloc(&d_table, 256));
        SVFM_CUDA(cudaMemcpyAsync(d_table, table256, 256, cudaMemcpyHostToDevice, st));
    }
    const unsigned g = grid1d(n) > 148u * 32u ? 148u * 32u : grid1d(n);
    encode_kernel<<<g, TPB, 0, st>>>(d_text, n, d_table, d_enc, S, d_bad);
    g_launches++;
    {
        // count_array: cumulative symbol counts; kmer_multiplier; kmer_count_table: inclusive prefix sums
        unsigned long long* d_hist;
        const uint64_t tl = L.kmer_count_table_len;
        SVFM_TRY(sc.alloc(&d_hist, tl + swsc));
        SVFM_CUDA(cudaMemsetAsync(d_hist, 0, (tl + swsc) * sizeof(unsigned long long), st));
        const int use_smem = tl * 4 <= 40 * 1024;
        kmer_hist_kernel<<<use_smem ? 148 * 4 : g, TPB, use_smem ? tl * 4 : 0, st>>>(d_enc, n, k, swsc, tl, d_hist, use_smem);
        kmer_hist_kernel<<<148 * 4, TPB, swsc * 4, st>>>(d_enc, n, 1, swsc, swsc, d_hist + tl, 1);
        g_launches += 2;
        size_t t2 = 0;
        SVFM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, t2, d_hist, d_hist, (int64_t)tl, st));
        uint8_t* d_temp;
        SVFM_TRY(sc.alloc(&d_temp, t2));
        SVFM_CUDA(cub::DeviceScan::InclusiveSum(d_temp, t2, d_hist, d_hist, (int64_t)tl, st));
        narrow_kernel<P><<<grid1d(tl), TPB, 0, st>>>(d_hist, reinterpret_cast<P*>(d_blob + L.off_kmer_count_table), tl);
        g_launches += 3;
        std::vector<unsigned long long> sym(swsc);
        SVFM_CUDA(cudaMemcpyAsync(sym.data(), d_hist + tl, swsc * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        SVFM_CUDA(cudaStreamSynchronize(st));
        // enc value v = symidx + 1, so sym[v] = occurrences of symidx v-1 = count_array[symidx + 1] before
        // accumulation (count_array.rs:117,125)
        std::vector<P> ca(swsc);
        unsigned long long acc = 0;
        for (uint32_t i = 0; i < swsc; i++) { acc += sym[i]; ca[i] = (P)acc; }
        SVFM_CUDA(cudaMemcpyAsync(d_blob + L.off_count_array, ca.data(), swsc * sizeof(P), cudaMemcpyHostToDevice, st));
        std::vector<uint64_t> mult(k);
        for (uint32_t j = 0; j < k; j++) { uint64_t p = 1; for (uint32_t e = 0; e < k - 1 - j; e++) p *= swsc; mult[j] = p; }
        SVFM_CUDA(cudaMemcpyAsync(d_blob + L.off_kmer_multiplier, mult.data(), k * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        SVFM_CUDA(cudaStreamSynchronize(st));
        sc.free_one(d_hist);
        sc.free_one(d_temp);
    }
    int bad = 0;
    SVFM_CUDA(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) return SVFM_ERR_BAD_SYMBOL;
    // 3) suffix array (suffix_array/mod.rs:57-70)
    uint32_t* d_sa;
    SVFM_TRY(sc.alloc(&d_sa, n));
    {
        Scratch sort_scratch;
        SVFM_TRY(suffix_sort(d_enc, n, S, d_sa, sort_scratch, st));
        SVFM_CUDA(cudaStreamSynchronize(st));
    }
    // 4) pidx, sampled SA, BWT, blocks + checkpoints (bwm/mod.rs:91-143)
    unsigned long long* d_pidx;
    uint8_t* d_bwt;
    SVFM_TRY(sc.alloc(&d_pidx, 1));
    SVFM_TRY(sc.alloc(&d_bwt, n));
    find_pidx_kernel<<<g, TPB, 0, st>>>(d_sa, n, d_pidx);
    sample_sa_kernel<P><<<grid1d(L.suffix_array_len) > 148u * 32u ? 148u * 32u : grid1d(L.suffix_array_len), TPB, 0, st>>>(
        d_sa, n, L.sampling_ratio, reinterpret_cast<P*>(d_blob + L.off_suffix_array), L.suffix_array_len);
    bwt_kernel<<<g, TPB, 0, st>>>(d_sa, d_enc, n, d_pidx, d_bwt);
    g_launches += 3;
    unsigned long long pidx = 0;
    SVFM_CUDA(cudaMemcpyAsync(&pidx, d_pidx, sizeof(pidx), cudaMemcpyDeviceToHost, st));
    SVFM_CUDA(cudaStreamSynchronize(st));
    const P pidx_p = (P)pidx;
    SVFM_CUDA(cudaMemcpyAsync(d_blob + L.off_sentinel_index, &pidx_p, sizeof(P), cudaMemcpyHostToDevice, st));
    P* d_ck = reinterpret_cast<P*>(d_blob + L.off_rank_checkpoints);
    const unsigned gb = grid1d(L.blocks_len, 128);
    if (t.vec_bits == 32)
        blocks_kernel<P, uint32_t, 1><<<gb, 128, 0, st>>>(d_bwt, n, S, t.planes, 32, L.blocks_len, reinterpret_cast<uint32_t*>(d_blob + L.off_blocks), d_ck);
    else if (t.vec_bits == 64)
        blocks_kernel<P, uint64_t, 1><<<gb, 128, 0, st>>>(d_bwt, n, S, t.planes, 64, L.blocks_len, reinterpret_cast<uint64_t*>(d_blob + L.off_blocks), d_ck);
    else
        blocks_kernel<P, uint64_t, 2><<<gb, 128, 0, st>>>(d_bwt, n, S, t.planes, 128, L.blocks_len, reinterpret_cast<uint64_t*>(d_blob + L.off_blocks), d_ck);
    g_launches++;
    const uint64_t chunks = (L.blocks_len + CK_CHUNK - 1) / CK_CHUNK;
    unsigned long long* d_tot;
    SVFM_TRY(sc.alloc(&d_tot, chunks * S));
    ck_chunk_totals_kernel<P><<<(unsigned)chunks, 64, 0, st>>>(d_ck, L.blocks_len, S, d_tot);
    ck_scan_totals_kernel<<<1, 64, 0, st>>>(d_tot, chunks, S);
    ck_apply_kernel<P><<<(unsigned)chunks, 64, 0, st>>>(d_ck, L.blocks_len, S, d_tot);
    g_launches += 3;
    SVFM_CUDA(cudaGetLastError());
    SVFM_CUDA(cudaStreamSynchronize(st));
    return SVFM_OK;
}

int build_device_impl(svfm_type t, const uint8_t* d_text, uint64_t n, uint32_t S, const uint8_t* table256, uint32_t k,
                      uint32_t r, int device, uint8_t* d_blob, uint64_t blob_len, uint64_t detail[2]) {
    Layout L;
    int rc = builder_layout(t, n, S, k, r, L, detail);
    if (rc) return rc;
    if (t.encoder && !table256) return SVFM_ERR_BAD_ARG;
    if (!d_text || !d_blob) return SVFM_ERR_BAD_ARG;
    if (n == 0) return SVFM_ERR_TEXT_LENGTH;
    if (blob_len != L.total_size) {  // builder/mod.rs:205-209
        if (detail) { detail[0] = L.total_size; detail[1] = blob_len; }
        return SVFM_ERR_INVALID_BLOB_SIZE;
    }
    if (((uintptr_t)d_blob) % L.align != 0) {  // builder/mod.rs:198-203
        if (detail) { detail[0] = L.align; detail[1] = ((uintptr_t)d_blob) % L.align; }
        return SVFM_ERR_NOT_ALIGNED;
    }
    if (n >= 0xffffffffull) return SVFM_ERR_TOO_LARGE;
    SVFM_CUDA(cudaSetDevice(device));
    cudaStream_t st;
    SVFM_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    rc = t.pos_bits == 32 ? build_typed<uint32_t>(t, L, d_text, n, S, table256, d_blob, st)
                          : build_typed<uint64_t>(t, L, d_text, n, S, table256, d_blob, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

}  // namespace
}  // namespace svfm

using namespace svfm;

extern "C" {

int svfm_blob_size(svfm_type t, uint64_t text_len, uint32_t symbol_count, uint32_t kmer_size,
                   uint32_t sampling_ratio, uint64_t* blob_size, uint64_t err_detail[2]) {
    if (!blob_size) return SVFM_ERR_BAD_ARG;
    Layout L;
    int rc = builder_layout(t, text_len, symbol_count, kmer_size, sampling_ratio, L, err_detail);
    if (rc) return rc;
    *blob_size = L.total_size;
    return SVFM_OK;
}

int svfm_build_device(svfm_type t, const uint8_t* d_text, uint64_t text_len, uint32_t symbol_count,
                      const uint8_t* table256, uint32_t kmer_size, uint32_t sampling_ratio, int device,
                      uint8_t* d_blob_out, uint64_t blob_len, uint64_t err_detail[2]) {
    return build_device_impl(t, d_text, text_len, symbol_count, table256, kmer_size, sampling_ratio, device,
                             d_blob_out, blob_len, err_detail);
}

int svfm_build(svfm_type t, const uint8_t* text, uint64_t text_len, uint32_t symbol_count, const uint8_t* table256,
               uint32_t kmer_size, uint32_t sampling_ratio, int device, uint8_t* blob_out, uint64_t blob_len,
               uint64_t err_detail[2]) {
    Layout L;
    int rc = builder_layout(t, text_len, symbol_count, kmer_size, sampling_ratio, L, err_detail);
    if (rc) return rc;
    if (!text || !blob_out) return SVFM_ERR_BAD_ARG;
    if (blob_len != L.total_size) {
        if (err_detail) { err_detail[0] = L.total_size; err_detail[1] = blob_len; }
        return SVFM_ERR_INVALID_BLOB_SIZE;
    }
    if (((uintptr_t)blob_out) % L.align != 0) {
        if (err_detail) { err_detail[0] = L.align; err_detail[1] = ((uintptr_t)blob_out) % L.align; }
        return SVFM_ERR_NOT_ALIGNED;
    }
    SVFM_CUDA(cudaSetDevice(device));
    uint8_t *d_text = nullptr, *d_blob = nullptr;
    SVFM_CUDA(cudaMalloc(&d_text, text_len + 64));
    cudaError_t e = cudaMalloc(&d_blob, blob_len + 64);
    if (e != cudaSuccess) { cudaFree(d_text); g_last_error = cudaGetErrorString(e); return SVFM_ERR_NOMEM; }
    e = cudaMemcpy(d_text, text, text_len, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = build_device_impl(t, d_text, text_len, symbol_count, table256, kmer_size, sampling_ratio, device, d_blob,
                               blob_len, err_detail);
        if (rc == SVFM_OK) e = cudaMemcpy(blob_out, d_blob, blob_len, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_text);
    cudaFree(d_blob);
    if (e != cudaSuccess) { g_last_error = std::string("svfm_build copy: ") + cudaGetErrorString(e); return SVFM_ERR_CUDA; }
    return rc;
}

}  // extern "C"
