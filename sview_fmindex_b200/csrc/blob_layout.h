// blob_layout.h -- host-side parsing of the reference's blob framing (no CUDA in this header).
// Restates Aligned::aligned_size / Header::aligned_size (components/mod.rs:1-23), the five headers
// (magic_number.rs:12-27, encoding_table.rs:7, count_array.rs:10-18, suffix_array/mod.rs:12-18,
// bwm/mod.rs:9-16), the body section sizes (count_array.rs:31-51, suffix_array/mod.rs:28-35,
// bwm/mod.rs:43-63) and the checks of FmIndex::load (load_from_blob.rs:28-58).
// Citations are relative to the reference's sview-fmindex/src/.
#pragma once
#include <cstdint>
#include <cstring>

#include "../../include/svfm.h"

namespace svfm {

struct Layout {
    uint64_t align = 8;
    uint64_t off_encoder = 0, off_count_header = 0, off_sa_header = 0, off_bwm_header = 0, header_size = 0;
    uint64_t off_count_array = 0, off_kmer_multiplier = 0, off_kmer_count_table = 0;
    uint64_t off_suffix_array = 0, off_sentinel_index = 0, off_rank_checkpoints = 0, off_blocks = 0;
    uint64_t total_size = 0;
    // header fields
    uint32_t symbol_count = 0, kmer_size = 0, count_array_len = 0, kmer_multiplier_len = 0;
    uint64_t kmer_count_table_len = 0;
    uint32_t sampling_ratio = 0;
    uint64_t suffix_array_len = 0;
    uint32_t bwm_symbol_count = 0;
    uint64_t rank_checkpoints_len = 0, blocks_len = 0;
};

inline uint64_t align_up(uint64_t raw, uint64_t a) {
    uint64_t rem = raw % a;
    return rem == 0 ? raw : raw + (a - rem);
}

inline bool type_ok(const svfm_type& t) {
    return (t.pos_bits == 32 || t.pos_bits == 64) && t.planes >= 2 && t.planes <= 6 &&
           (t.vec_bits == 32 || t.vec_bits == 64 || t.vec_bits == 128) && t.encoder <= 1;
}

inline uint64_t block_bytes(const svfm_type& t) { return (uint64_t)t.planes * t.vec_bits / 8; }

// Section offsets from the header fields in L (all sections rounded up to B::ALIGN_SIZE).
inline void resolve_offsets(const svfm_type& t, Layout& L) {
    const uint64_t A = t.vec_bits == 128 ? 16 : 8;  // vector.rs:37,52,67
    const uint64_t P = t.pos_bits / 8;
    L.align = A;
    uint64_t o = align_up(8, A);
    L.off_encoder = o;
    o += align_up(t.encoder ? 256 : 0, A);
    L.off_count_header = o;
    o += align_up(24, A);
    L.off_sa_header = o;
    o += align_up(16, A);
    L.off_bwm_header = o;
    o += align_up(24, A);
    L.header_size = o;
    L.off_count_array = o;
    o += align_up((uint64_t)L.count_array_len * P, A);
    L.off_kmer_multiplier = o;
    o += align_up((uint64_t)L.kmer_multiplier_len * 8, A);  // Vec<usize>, 8 B on 64-bit targets
    L.off_kmer_count_table = o;
    o += align_up(L.kmer_count_table_len * P, A);
    L.off_suffix_array = o;
    o += align_up(L.suffix_array_len * P, A);
    L.off_sentinel_index = o;
    o += align_up(P, A);
    L.off_rank_checkpoints = o;
    o += align_up(L.rank_checkpoints_len * P, A);
    L.off_blocks = o;
    o += align_up(L.blocks_len * block_bytes(t), A);
    L.total_size = o;
}

// FmIndexBuilder::new + generate_headers (builder/mod.rs:63-135).
inline int builder_layout(const svfm_type& t, uint64_t text_len, uint32_t symbol_count, uint32_t kmer_size,
                          uint32_t sampling_ratio, Layout& L, uint64_t detail[2]) {
    if (!type_ok(t)) return SVFM_ERR_BAD_TYPE;
    const uint32_t max_symbol = 1u << t.planes;  // block2.rs:15 .. block6.rs:15
    if (symbol_count > max_symbol) {
        if (detail) { detail[0] = max_symbol; detail[1] = symbol_count; }
        return SVFM_ERR_SYMBOL_COUNT_OVER;
    }
    if (symbol_count < 1 || kmer_size < 1 || sampling_ratio < 1) return SVFM_ERR_INVALID_CONFIG;
    L = Layout();
    L.symbol_count = symbol_count;
    L.bwm_symbol_count = symbol_count;
    L.kmer_size = kmer_size;
    L.count_array_len = symbol_count + 1;
    L.kmer_multiplier_len = kmer_size;
    // (S+1)^k: the reference computes it with u32::pow (count_array.rs:70), which panics on overflow in debug builds and
    // wraps in release builds (the table is then too short for the indices the search computes, and the reference panics
    // on the bounds check).  Neither is acceptable across a C ABI, and the device histogram is indexed with the true
    // value: anything that does not fit the reference's u32 is an invalid configuration.
    unsigned __int128 p = 1;
    for (uint32_t e = 0; e < kmer_size; e++) {
        p *= (symbol_count + 1);
        if (p > (unsigned __int128)0xffffffffu) {
            if (detail) { detail[0] = 0xffffffffu; detail[1] = kmer_size; }
            return SVFM_ERR_INVALID_CONFIG;
        }
    }
    L.kmer_count_table_len = (uint64_t)p;
    L.sampling_ratio = sampling_ratio;
    L.suffix_array_len = text_len / sampling_ratio + (text_len % sampling_ratio ? 1 : 0);
    L.blocks_len = text_len / t.vec_bits + 1;
    L.rank_checkpoints_len = L.blocks_len * symbol_count;
    resolve_offsets(t, L);
    return SVFM_OK;
}

// Header part of FmIndex::load: `read` copies `n` bytes at blob offset `off` into dst (host memcpy or
// cudaMemcpy D2H, so the same code validates host and device blobs).
template <class ReadFn>
inline int parse_blob(ReadFn&& read, uint64_t blob_len, const svfm_type& t, Layout& L, uint64_t detail[2]) {
    if (!type_ok(t)) return SVFM_ERR_BAD_TYPE;
    L = Layout();
    resolve_offsets(t, L);  // header offsets do not depend on header contents
    if (blob_len < L.header_size) return SVFM_ERR_INVALID_FORMAT;
    uint8_t hdr[8 + 16 + 256 + 32 + 16 + 32];
    if (L.header_size > sizeof(hdr)) return SVFM_ERR_INVALID_FORMAT;
    if (!read(hdr, 0, L.header_size)) return SVFM_ERR_CUDA;
    // MagicNumber::is_valid && is_supported_version (magic_number.rs:38-49)
    if (!(hdr[0] == 'F' && hdr[1] == 'I' && hdr[2] == '0' && hdr[3] == '0')) return SVFM_ERR_INVALID_FORMAT;
    const uint8_t* h = hdr + L.off_count_header;
    std::memcpy(&L.symbol_count, h + 0, 4);
    std::memcpy(&L.kmer_size, h + 4, 4);
    std::memcpy(&L.count_array_len, h + 8, 4);
    std::memcpy(&L.kmer_multiplier_len, h + 12, 4);
    std::memcpy(&L.kmer_count_table_len, h + 16, 8);
    h = hdr + L.off_sa_header;
    std::memcpy(&L.sampling_ratio, h + 0, 4);
    std::memcpy(&L.suffix_array_len, h + 8, 8);
    h = hdr + L.off_bwm_header;
    std::memcpy(&L.bwm_symbol_count, h + 0, 4);
    std::memcpy(&L.rank_checkpoints_len, h + 8, 8);
    std::memcpy(&L.blocks_len, h + 16, 8);
    resolve_offsets(t, L);
    if (L.total_size != blob_len) {  // load_from_blob.rs:46-58
        if (detail) { detail[0] = L.total_size; detail[1] = blob_len; }
        return SVFM_ERR_BLOB_SIZE;
    }
    // Sanity the reference never needs (it indexes out of bounds / panics instead): a blob whose header
    // fields are inconsistent cannot be searched safely on a device.
    if (L.symbol_count == 0 || L.symbol_count > 64 || L.bwm_symbol_count != L.symbol_count ||
        L.count_array_len != L.symbol_count + 1 || L.kmer_size == 0 || L.kmer_multiplier_len != L.kmer_size ||
        L.sampling_ratio == 0 || L.blocks_len == 0 || L.rank_checkpoints_len != L.blocks_len * L.symbol_count)
        return SVFM_ERR_INVALID_FORMAT;
    {
        unsigned __int128 p = 1;
        for (uint32_t e = 0; e < L.kmer_size; e++) { p *= (L.symbol_count + 1); if (p > ((unsigned __int128)1 << 40)) break; }
        if (p != L.kmer_count_table_len) return SVFM_ERR_INVALID_FORMAT;
    }
    return SVFM_OK;
}

}  // namespace svfm
