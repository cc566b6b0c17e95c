// C++ twin of the reference's README doctest (sview-fmindex/src/tests/readme/mod.rs:3-46) through the
// header-only facade sview_fmindex_b200/include/sview_fmindex.hpp over libsvfm.so.  Exit code 0 = pass.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../sview_fmindex_b200/include/sview_fmindex.hpp"

using namespace sview_fmindex;
using blocks::Block2;
using text_encoders::EncodingTable;

#define REQUIRE(cond)                                                        \
    do {                                                                     \
        if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

int main() {
    // (1) Define symbols to use
    auto encoding_table = EncodingTable::from_symbols({"Aa", "Cc", "Gg", "Tt"});
    uint32_t symbol_count = encoding_table.symbol_count();  // 4
    REQUIRE(symbol_count == 4);

    // (2) Build index
    std::string t = "CTCCGTACACCTGTTTCGTATCGGAXXYYZZ";
    std::vector<uint8_t> text(t.begin(), t.end());
    FmIndexBuilder<uint32_t, Block2<uint64_t>, EncodingTable> builder(text.size(), symbol_count, encoding_table);
    size_t blob_size = builder.blob_size();
    REQUIRE(blob_size == 552);
    void* mem = nullptr;
    REQUIRE(posix_memalign(&mem, 64, blob_size) == 0);
    uint8_t* blob = (uint8_t*)mem;
    builder.build(text, blob, blob_size);
    auto fm_index = FmIndex<uint32_t, Block2<uint64_t>, EncodingTable>::load(blob, blob_size);

    // (3) Match with pattern
    REQUIRE(fm_index.count("TA") == 2);
    auto locations = fm_index.locate("TA");
    std::sort(locations.begin(), locations.end());  // The locations may not be in order.
    REQUIRE((locations == std::vector<uint32_t>{5, 18}));
    locations = fm_index.locate("UNDEF");  // the last symbol is treated as wild card
    std::sort(locations.begin(), locations.end());
    REQUIRE((locations == std::vector<uint32_t>{25, 26}));

    // rev-iter twins and locate_to_buffer (appends)
    std::string rev = "AT";
    REQUIRE(fm_index.count_rev_iter(rev.begin(), rev.end()) == 2);
    std::vector<uint32_t> buf{7};
    fm_index.locate_to_buffer("TA", buf);
    REQUIRE(buf.size() == 3 && buf[0] == 7);

    // batched entry points
    auto counts = fm_index.count_batch({"TA", "UNDEF", "GGGG", "C"});
    REQUIRE((counts == std::vector<uint32_t>{2, 2, 0, 8}));
    auto lb = fm_index.locate_batch({"TA", "GGGG", "XXXXX"}, /*sorted=*/true);
    REQUIRE((lb.offsets == std::vector<uint64_t>{0, 2, 2, 4}));
    REQUIRE((lb.positions == std::vector<uint32_t>{5, 18, 25, 26}));

    // LoadError paths
    blob[0] = 'X';
    try {
        FmIndex<uint32_t, Block2<uint64_t>, EncodingTable>::load(blob, blob_size);
        REQUIRE(false);
    } catch (const LoadError& e) { REQUIRE(e.is_invalid_format()); }
    blob[0] = 'F';
    try {
        FmIndex<uint32_t, Block2<uint64_t>, EncodingTable>::load(blob, blob_size - 8);
        REQUIRE(false);
    } catch (const LoadError& e) { REQUIRE(e.is_mismatched_blob_size() && e.detail[1] == blob_size - 8); }
    try {
        fm_index.count("");
        REQUIRE(false);
    } catch (const SvfmError& e) { REQUIRE(e.code == SVFM_ERR_EMPTY_PATTERN); }
    std::free(mem);
    std::printf("readme_example: ok\n");
    return 0;
}
