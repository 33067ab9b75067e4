// Exhaustive check of blamm_format_score (the fast "%g" of the occurrence writer) against snprintf("%g") over EVERY float
// between 1e-5 and 1e7 in magnitude (both signs; the fast path covers [1e-4, 1e6), the rest exercises the hand-over to
// std::to_chars), on all host threads.
//   g++ -O2 -std=c++17 tools/format_check.cpp -Iinclude -Lblamm_b200/lib -lblammhost -Wl,-rpath,$PWD/blamm_b200/lib -lpthread -o /tmp/format_check
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <thread>
#include <vector>
#include "blamm_host.h"

int main(int argc, char** argv)
{
    const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1u;
    float lo = 1e-5f, hi = 1e7f;
    uint32_t blo, bhi; memcpy(&blo, &lo, 4); memcpy(&bhi, &hi, 4);
    const unsigned T = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint64_t> bad{0}, done{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; t++) th.emplace_back([&, t] {
        char a[40], b[40];
        uint64_t n = 0;
        for (uint64_t bits = (uint64_t)blo + t * stride; bits <= bhi; bits += (uint64_t)T * stride) {
            for (uint32_t sign = 0; sign < 2; sign++) {
                const uint32_t u = (uint32_t)bits | (sign << 31);
                float v; memcpy(&v, &u, 4);
                const int la = blamm_format_score(v, a);
                const int lb = snprintf(b, sizeof b, "%g", (double)v);
                if (la != lb || memcmp(a, b, (size_t)la) != 0) {
                    if (bad.fetch_add(1) < 10) { a[la] = 0; printf("MISMATCH bits %08x: got '%s' want '%s'\n", u, a, b); }
                }
                n++;
            }
        }
        done += n;
    });
    for (auto& x : th) x.join();
    printf("%llu floats checked, %llu mismatches\n", (unsigned long long)done.load(), (unsigned long long)bad.load());
    return bad ? 1 : 0;
}
