// builder.cu -- index construction on the GPU.
#include "blob_layout.h"
#include "common.cuh"

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

int svfm_build(svfm_type, const uint8_t*, uint64_t, uint32_t, const uint8_t*, uint32_t, uint32_t, int, uint8_t*,
               uint64_t, uint64_t*) {
    g_last_error = "svfm_build: not implemented yet";
    return SVFM_ERR_CUDA;
}

int svfm_build_device(svfm_type, const uint8_t*, uint64_t, uint32_t, const uint8_t*, uint32_t, uint32_t, int,
                      uint8_t*, uint64_t, uint64_t*) {
    g_last_error = "svfm_build_device: not implemented yet";
    return SVFM_ERR_CUDA;
}

}  // extern "C"
