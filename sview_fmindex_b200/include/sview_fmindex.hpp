// sview_fmindex.hpp -- C++ host-side mirror of the reference's public API for the hot path, over the
// C ABI of libsvfm.so (include/svfm.h).  Header-only.
//
// The reference is a Rust crate (baku4/sview-fmindex); this image has no Rust toolchain, so the host side
// above the C ABI is written in C++ with the reference's names, argument meaning and error behaviour
// (citations are relative to the reference's sview-fmindex/src/):
//
//   FmIndex<'a, P, B, E>::load(blob) -> Result<Self, LoadError>        load_from_blob.rs:28
//   count / locate / locate_to_buffer                                  locate/with_slice.rs:5-18
//   count_rev_iter / locate_rev_iter / locate_rev_iter_to_buffer       locate/with_rev_iter.rs:5-18
//   blob()                                                             reference_to_source_blob.rs:9
//   LoadError::{InvalidFormat, MismatchedBlobSize(expected, actual)}   load_from_blob.rs:16-24
//   FmIndexBuilder<P, B, E>::new / set_*_config / blob_size / build    builder/mod.rs:63-264
//   BuildError                                                         builder/mod.rs:37-57
//   blocks::Block2..Block6<V>, Position = u32 | u64                    components/bwm/blocks/, text_length.rs
//   text_encoders::{EncodingTable, PassThrough}                        components/text_encoder/text_encoders/
//
// plus the batched entry points (count_batch / locate_batch) this engine adds.  Rust's Result<T, E> becomes
// a C++ exception (LoadError / BuildError / SvfmError); Rust's panic on an empty pattern becomes SvfmError
// with code SVFM_ERR_EMPTY_PATTERN.  Everything runs on the GPU: there is no CPU fallback.
#pragma once
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/svfm.h"

namespace sview_fmindex {

// ---- errors ----------------------------------------------------------------------------------------
struct SvfmError : std::runtime_error {
    int code;
    uint64_t detail[2];
    SvfmError(int c, const uint64_t d[2], const std::string& what)
        : std::runtime_error(what + " (svfm code " + std::to_string(c) + ")"), code(c), detail{d ? d[0] : 0, d ? d[1] : 0} {}
};
struct LoadError : SvfmError {  // load_from_blob.rs:16-24
    using SvfmError::SvfmError;
    bool is_invalid_format() const { return code == SVFM_ERR_INVALID_FORMAT; }
    bool is_mismatched_blob_size() const { return code == SVFM_ERR_BLOB_SIZE; }  // detail = {expected, actual}
};
struct BuildError : SvfmError {  // builder/mod.rs:37-57
    using SvfmError::SvfmError;
};

inline void check(int rc, const uint64_t detail[2] = nullptr) {
    if (rc == SVFM_OK) return;
    std::string msg = rc == SVFM_ERR_CUDA ? std::string("CUDA: ") + svfm_last_error() : "svfm call failed";
    if (rc == SVFM_ERR_INVALID_FORMAT || rc == SVFM_ERR_BLOB_SIZE)
        throw LoadError(rc, detail, rc == SVFM_ERR_INVALID_FORMAT
                                        ? "Invalid FM-index format. The data does not appear to be a valid FM-index blob."
                                        : "Mismatched blob size");
    if (rc >= SVFM_ERR_SYMBOL_COUNT_OVER && rc <= SVFM_ERR_INVALID_CONFIG) throw BuildError(rc, detail, "build error");
    throw SvfmError(rc, detail, msg);
}

// ---- the type parameters ----------------------------------------------------------------------------
namespace blocks {
template <class V, int N>
struct BlockN {
    static_assert(std::is_same<V, uint32_t>::value || std::is_same<V, uint64_t>::value ||
                      std::is_same<V, unsigned __int128>::value,
                  "Vector is u32, u64 or u128 (blocks/vector.rs:35-79)");
    static constexpr uint32_t PLANES = N;
    static constexpr uint32_t BLOCK_LEN = sizeof(V) * 8;
    static constexpr uint32_t MAX_SYMBOL = 1u << N;  // block2.rs:15 .. block6.rs:15
};
template <class V> using Block2 = BlockN<V, 2>;
template <class V> using Block3 = BlockN<V, 3>;
template <class V> using Block4 = BlockN<V, 4>;
template <class V> using Block5 = BlockN<V, 5>;
template <class V> using Block6 = BlockN<V, 6>;
}  // namespace blocks

namespace text_encoders {
// EncodingTable([u8; 256]) (encoding_table.rs:7-38): the last symbol is the wildcard.
struct EncodingTable {
    static constexpr uint32_t KIND = 1;
    uint8_t table[256];
    static EncodingTable from_symbols(const std::vector<std::string>& symbols) {
        EncodingTable t;
        std::memset(t.table, (int)(uint8_t)(symbols.size() - 1), 256);
        for (size_t i = 0; i < symbols.size(); i++)
            for (unsigned char c : symbols[i]) t.table[c] = (uint8_t)i;
        return t;
    }
    static EncodingTable from_symbols_with_wildcard(const std::vector<std::string>& symbols) {
        EncodingTable t;
        std::memset(t.table, (int)(uint8_t)symbols.size(), 256);
        for (size_t i = 0; i < symbols.size(); i++)
            for (unsigned char c : symbols[i]) t.table[c] = (uint8_t)i;
        return t;
    }
    uint32_t symbol_count() const {
        uint32_t mx = 0;
        for (int i = 0; i < 256; i++) mx = table[i] > mx ? table[i] : mx;
        return mx + 1;
    }
    uint8_t idx_of(uint8_t sym) const { return table[sym]; }
    const uint8_t* bytes() const { return table; }
};
// PassThrough (pass_through.rs:6-13): zero-sized, identity.
struct PassThrough {
    static constexpr uint32_t KIND = 0;
    uint8_t idx_of(uint8_t sym) const { return sym; }
    const uint8_t* bytes() const { return nullptr; }
};
}  // namespace text_encoders

template <class P, class B, class E>
inline svfm_type type_of() {
    static_assert(std::is_same<P, uint32_t>::value || std::is_same<P, uint64_t>::value, "Position is u32 or u64");
    return svfm_type{(uint32_t)sizeof(P) * 8, B::PLANES, B::BLOCK_LEN, E::KIND};
}

// CSR result of a batched locate: pattern i owns positions[offsets[i] .. offsets[i+1]).
template <class P>
struct LocateBatch {
    std::vector<uint64_t> offsets;
    std::vector<P> positions;
};

// ---- FmIndex -----------------------------------------------------------------------------------------
template <class P, class B, class E>
class FmIndex {
  public:
    // FmIndex::load(blob) (load_from_blob.rs:28-85).  The device copy is owned by the handle; the view keeps
    // the host pointer only for blob() (the Rust struct borrows it for 'a).
    static FmIndex load(const uint8_t* blob, size_t len, int device = 0) {
        FmIndex ix;
        uint64_t detail[2] = {0, 0};
        check(svfm_load(blob, len, type_of<P, B, E>(), device, &ix.h_, detail), detail);
        ix.blob_ = blob;
        ix.blob_len_ = len;
        return ix;
    }
    static FmIndex load(const std::vector<uint8_t>& blob, int device = 0) { return load(blob.data(), blob.size(), device); }

    FmIndex(FmIndex&& o) noexcept : h_(o.h_), blob_(o.blob_), blob_len_(o.blob_len_) { o.h_ = nullptr; }
    FmIndex& operator=(FmIndex&& o) noexcept {
        if (this != &o) { release(); h_ = o.h_; blob_ = o.blob_; blob_len_ = o.blob_len_; o.h_ = nullptr; }
        return *this;
    }
    FmIndex(const FmIndex&) = delete;
    FmIndex& operator=(const FmIndex&) = delete;
    ~FmIndex() { release(); }

    std::pair<const uint8_t*, size_t> blob() const { return {blob_, blob_len_}; }
    svfm_index* handle() const { return h_; }

    // count(&self, pattern: &[u8]) -> P (locate/with_slice.rs:5-8)
    P count(const uint8_t* pattern, size_t len) const { return count_flags(pattern, len, 0); }
    P count(const std::string& pattern) const { return count((const uint8_t*)pattern.data(), pattern.size()); }
    // locate(&self, pattern) -> Vec<P>, SA-row order (locate/with_slice.rs:10-13; README.md:77)
    std::vector<P> locate(const uint8_t* pattern, size_t len) const {
        std::vector<P> out;
        locate_flags(pattern, len, 0, out);
        return out;
    }
    std::vector<P> locate(const std::string& pattern) const { return locate((const uint8_t*)pattern.data(), pattern.size()); }
    // locate_to_buffer(&self, pattern, &mut Vec<P>): appends, does not clear (locate/with_slice.rs:15-18)
    void locate_to_buffer(const uint8_t* pattern, size_t len, std::vector<P>& buffer) const { locate_flags(pattern, len, 0, buffer); }
    void locate_to_buffer(const std::string& pattern, std::vector<P>& buffer) const {
        locate_flags((const uint8_t*)pattern.data(), pattern.size(), 0, buffer);
    }
    // rev-iterator twins (locate/with_rev_iter.rs:5-18): any iterator range yielding the pattern back to front
    template <class It>
    P count_rev_iter(It first, It last) const {
        std::vector<uint8_t> rev(first, last);
        return count_flags(rev.data(), rev.size(), SVFM_REVERSED);
    }
    template <class It>
    std::vector<P> locate_rev_iter(It first, It last) const {
        std::vector<uint8_t> rev(first, last);
        std::vector<P> out;
        locate_flags(rev.data(), rev.size(), SVFM_REVERSED, out);
        return out;
    }
    template <class It>
    void locate_rev_iter_to_buffer(It first, It last, std::vector<P>& buffer) const {
        std::vector<uint8_t> rev(first, last);
        locate_flags(rev.data(), rev.size(), SVFM_REVERSED, buffer);
    }

    // ---- batched entry points (new) ----
    // fixed-length patterns, n * len bytes
    std::vector<P> count_batch(const uint8_t* pats, uint64_t n, uint32_t len) const {
        std::vector<P> out(n);
        check(svfm_count_batch(h_, pats, nullptr, n, len, 0, out.data()));
        return out;
    }
    std::vector<P> count_batch(const std::vector<std::string>& patterns) const {
        std::vector<uint8_t> data;
        std::vector<uint64_t> offs;
        pack(patterns, data, offs);
        std::vector<P> out(patterns.size());
        check(svfm_count_batch(h_, data.data(), offs.data(), patterns.size(), 0, 0, out.data()));
        return out;
    }
    LocateBatch<P> locate_batch(const uint8_t* pats, const uint64_t* offs, uint64_t n, uint32_t fixed_len,
                                bool sorted = false) const {
        LocateBatch<P> r;
        r.offsets.assign(n + 1, 0);
        void* pos = nullptr;
        uint64_t total = 0;
        check(svfm_locate_batch_alloc(h_, pats, offs, n, fixed_len, sorted ? SVFM_SORTED : 0, r.offsets.data(), &pos, &total));
        r.positions.assign((const P*)pos, (const P*)pos + total);
        svfm_free_positions(pos);
        return r;
    }
    LocateBatch<P> locate_batch(const std::vector<std::string>& patterns, bool sorted = false) const {
        std::vector<uint8_t> data;
        std::vector<uint64_t> offs;
        pack(patterns, data, offs);
        return locate_batch(data.data(), offs.data(), patterns.size(), 0, sorted);
    }

  private:
    FmIndex() = default;
    void release() {
        if (h_) svfm_free(h_);
        h_ = nullptr;
    }
    static void pack(const std::vector<std::string>& patterns, std::vector<uint8_t>& data, std::vector<uint64_t>& offs) {
        offs.assign(1, 0);
        for (const auto& p : patterns) {
            data.insert(data.end(), p.begin(), p.end());
            offs.push_back(data.size());
        }
        if (data.empty()) data.push_back(0);
    }
    P count_flags(const uint8_t* pattern, size_t len, uint32_t flags) const {
        uint64_t c = 0;
        check(svfm_count(h_, pattern, len, flags, &c));
        return (P)c;
    }
    void locate_flags(const uint8_t* pattern, size_t len, uint32_t flags, std::vector<P>& buffer) const {
        if (len == 0) check(SVFM_ERR_EMPTY_PATTERN);
        uint64_t offs[2] = {0, 0};
        void* pos = nullptr;
        uint64_t total = 0;
        check(svfm_locate_batch_alloc(h_, pattern, nullptr, 1, (uint32_t)len, flags, offs, &pos, &total));
        buffer.insert(buffer.end(), (const P*)pos, (const P*)pos + total);
        svfm_free_positions(pos);
    }
    svfm_index* h_ = nullptr;
    const uint8_t* blob_ = nullptr;
    size_t blob_len_ = 0;
};

// ---- FmIndexBuilder (GPU suffix sort) -----------------------------------------
namespace build_config {
struct SuffixArrayConfig {  // suffix_array_config.rs:5-34
    uint32_t ratio = 1;
    static SuffixArrayConfig Uncompressed() { return {1}; }
    static SuffixArrayConfig Compressed(uint32_t r) { return {r < 2 ? 0u : r}; }  // 0 = InvalidConfig
};
struct LookupTableConfig {  // lookup_table_config.rs:6-53
    uint32_t kmer = 1;
    uint64_t max_memory = 0;
    bool by_memory = false;
    static LookupTableConfig None() { return {}; }
    static LookupTableConfig KmerSize(uint32_t k) { LookupTableConfig c; c.kmer = k < 2 ? 0u : k; return c; }
    static LookupTableConfig MaxMemory(uint64_t bytes) { LookupTableConfig c; c.max_memory = bytes; c.by_memory = true; return c; }
};
}  // namespace build_config

template <class P, class B, class E>
class FmIndexBuilder {
  public:
    FmIndexBuilder(size_t text_len, uint32_t symbol_count, const E& encoder, int device = 0)
        : text_len_(text_len), symbol_count_(symbol_count), encoder_(encoder), device_(device) {
        blob_size();  // SymbolCountOver check of FmIndexBuilder::new (builder/mod.rs:71-73)
    }
    FmIndexBuilder& set_lookup_table_config(build_config::LookupTableConfig c) {
        if (c.by_memory) {
            uint64_t swsc = (uint64_t)symbol_count_ + 1;
            uint32_t k = 2;
            for (;;) {  // lookup_table_config.rs:41-53, plus: (S+1)^k must fit the u32 the reference stores it in
                unsigned __int128 entries = 1;
                for (uint32_t e = 0; e < k; e++) entries *= swsc;
                if (entries * sizeof(P) <= c.max_memory && entries <= (unsigned __int128)0xffffffffu) k++; else break;
            }
            kmer_ = k - 1;
        } else {
            if (c.kmer == 0) check(SVFM_ERR_INVALID_CONFIG);  // "K-mer size must be at least 2"
            kmer_ = c.kmer;
        }
        blob_size();
        return *this;
    }
    FmIndexBuilder& set_suffix_array_config(build_config::SuffixArrayConfig c) {
        if (c.ratio == 0) check(SVFM_ERR_INVALID_CONFIG);  // "Sampling ratio ... must be at least 2"
        ratio_ = c.ratio;
        blob_size();
        return *this;
    }
    size_t blob_size() const {
        uint64_t size = 0, detail[2] = {0, 0};
        check(svfm_blob_size(type_of<P, B, E>(), text_len_, symbol_count_, kmer_, ratio_, &size, detail), detail);
        return (size_t)size;
    }
    // build(&self, text: Vec<u8>, blob: &mut [u8]) (builder/mod.rs:187-264); suffix sort on the GPU
    void build(const std::vector<uint8_t>& text, uint8_t* blob, size_t blob_len) const {
        uint64_t detail[2] = {0, 0};
        if (text.size() != text_len_) {
            detail[0] = text_len_;
            detail[1] = text.size();
            check(SVFM_ERR_TEXT_LENGTH, detail);
        }
        check(svfm_build(type_of<P, B, E>(), text.data(), text.size(), symbol_count_, encoder_.bytes(), kmer_, ratio_,
                         device_, blob, blob_len, detail), detail);
    }

  private:
    size_t text_len_;
    uint32_t symbol_count_;
    E encoder_;
    int device_;
    uint32_t kmer_ = 1, ratio_ = 1;
};

}  // namespace sview_fmindex
