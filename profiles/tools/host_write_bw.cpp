// How fast can T host threads fill a (pre-touched) host buffer?  Bounds the host-side expansion of
// compact records into the caller's raw_records array.  g++ -O2 -pthread -mavx2 host_write_bw.cpp
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
    size_t bytes = (argc > 1 ? atoll(argv[1]) : 4096) << 20;
    unsigned hw = std::thread::hardware_concurrency();
    printf("hardware_concurrency %u\n", hw);
    uint8_t *buf = (uint8_t *)aligned_alloc(4096, bytes);
    double t0 = now();
    memset(buf, 1, bytes);
    printf("first touch 1 thread: %.2f GB/s\n", bytes / (now() - t0) / 1e9);
    for (int mode = 0; mode < 2; mode++)
        for (unsigned T : {1u, 2u, 4u, 8u, 16u, 32u}) {
            if (T > 2 * hw) break;
            double best = 0;
            for (int rep = 0; rep < 3; rep++) {
                std::vector<std::thread> th;
                double t1 = now();
                for (unsigned k = 0; k < T; k++)
                    th.emplace_back([=] {
                        size_t a = bytes / T * k, b = bytes / T * (k + 1);
                        if (mode == 0) memset(buf + a, rep + 2, b - a);
                        else {
                            __m256i v = _mm256_set1_epi16((short)(16000 + rep));
                            for (size_t o = a; o + 32 <= b; o += 32) _mm256_stream_si256((__m256i *)(buf + o), v);
                            _mm_sfence();
                        }
                    });
                for (auto &t : th) t.join();
                double gbs = bytes / (now() - t1) / 1e9;
                if (gbs > best) best = gbs;
            }
            printf("%s threads %2u: %.1f GB/s\n", mode ? "stream" : "memset", T, best);
        }
    return 0;
}
