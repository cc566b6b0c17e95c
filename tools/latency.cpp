// Single-call latency of the C ABI (include/svfm.h): svfm_count / svfm_locate on one 20-symbol pattern per call, and small
// batches, on a 100 Mbp index built on the GPU.   Build + run (GPU box):
//   g++ -O2 -std=c++17 tools/latency.cpp -Iinclude -Lsview_fmindex_b200 -lsvfm -Wl,-rpath,$PWD/sview_fmindex_b200 -o tools/latency.bin
//   SVFM_SMALL_MAX=0 tools/latency.bin     # the general pipeline, for comparison
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "svfm.h"

static uint64_t splitmix(uint64_t& x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 100000000ull;
    std::vector<uint8_t> text(n);
    uint64_t seed = 42;
    for (uint64_t i = 0; i < n; i++) text[i] = "ACGT"[splitmix(seed) >> 62];
    uint8_t table[256];
    for (int i = 0; i < 256; i++) table[i] = 4;
    const char* sym = "AaCcGgTtNn";
    for (int i = 0; i < 10; i++) table[(uint8_t)sym[i]] = (uint8_t)(i / 2);
    svfm_type t{32, 3, 64, 1};
    uint64_t size = 0, detail[2];
    if (svfm_blob_size(t, n, 5, 3, 2, &size, detail)) return 1;
    std::vector<uint8_t> raw(size + 64);
    uint8_t* blob = raw.data() + ((64 - ((uintptr_t)raw.data() & 63)) & 63);
    if (svfm_build(t, text.data(), n, 5, table, 3, 2, 0, blob, size, detail)) { fprintf(stderr, "build failed: %s\n", svfm_last_error()); return 1; }
    svfm_index* ix = nullptr;
    if (svfm_load(blob, size, t, 0, &ix, detail)) { fprintf(stderr, "load failed: %s\n", svfm_last_error()); return 1; }
    const int reps = 5000;
    std::vector<uint64_t> starts(reps);
    for (auto& s : starts) s = splitmix(seed) % (n - 20);
    uint64_t cnt = 0, total = 0, sink = 0;
    uint32_t pos[4096];
    for (int i = 0; i < 200; i++) { svfm_count(ix, &text[starts[i]], 20, 0, &cnt); svfm_locate(ix, &text[starts[i]], 20, 0, pos, 4096, &total); }
    double t0 = now_us();
    for (int i = 0; i < reps; i++) { svfm_count(ix, &text[starts[i]], 20, 0, &cnt); sink += cnt; }
    double t1 = now_us();
    for (int i = 0; i < reps; i++) { svfm_locate(ix, &text[starts[i]], 20, 0, pos, 4096, &total); sink += pos[0]; }
    double t2 = now_us();
    printf("svfm_count  %.1f us/call\nsvfm_locate %.1f us/call   (%d calls each, 20-symbol patterns cut from a %llu bp text)\n",
           (t1 - t0) / reps, (t2 - t1) / reps, reps, (unsigned long long)n);
    for (uint64_t m : {16ull, 256ull, 4096ull, 100000ull}) {
        std::vector<uint8_t> pats(m * 20);
        for (uint64_t i = 0; i < m; i++) { const uint64_t s = splitmix(seed) % (n - 20); for (int j = 0; j < 20; j++) pats[i * 20 + j] = text[s + j]; }
        std::vector<uint32_t> counts(m), out(m * 4 + 64);
        std::vector<uint64_t> offs(m + 1);
        const int r2 = m > 10000 ? 50 : 500;
        svfm_count_batch(ix, pats.data(), nullptr, m, 20, 0, counts.data());
        svfm_locate_batch(ix, pats.data(), nullptr, m, 20, 0, offs.data(), out.data(), out.size(), &total);
        t0 = now_us();
        for (int i = 0; i < r2; i++) svfm_count_batch(ix, pats.data(), nullptr, m, 20, 0, counts.data());
        t1 = now_us();
        for (int i = 0; i < r2; i++) svfm_locate_batch(ix, pats.data(), nullptr, m, 20, 0, offs.data(), out.data(), out.size(), &total);
        t2 = now_us();
        printf("batch of %6llu: count %8.1f us (%7.2f M/s)   locate %8.1f us (%7.2f M/s)\n", (unsigned long long)m, (t1 - t0) / r2,
               m * r2 / (t1 - t0), (t2 - t1) / r2, m * r2 / (t2 - t1));
    }
    svfm_free(ix);
    return (int)(sink & 0);
}
