// host_copy_bw.cpp -- how fast can the host cores of a GPU box copy memory?  (The pageable-memory path of the host entry
// points stages every byte through page-locked slots with CPU copies; this is its ceiling.)
//   g++ -O2 -pthread -mavx2 host_copy_bw.cpp -o host_copy_bw && ./host_copy_bw
#include <immintrin.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void copy_nt(char* dst, const char* src, size_t n) {
    for (size_t i = 0; i + 128 <= n; i += 128) {
        __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
        __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64)), d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
        _mm256_stream_si256((__m256i*)(dst + i), a);
        _mm256_stream_si256((__m256i*)(dst + i + 32), b);
        _mm256_stream_si256((__m256i*)(dst + i + 64), c);
        _mm256_stream_si256((__m256i*)(dst + i + 96), d);
    }
    _mm_sfence();
}

int main() {
    const size_t total = size_t(2) << 30;
    char* src = (char*)aligned_alloc(4096, total);
    char* dst = (char*)aligned_alloc(4096, total);
    memset(src, 1, total);
    memset(dst, 2, total);
    unsigned hw = std::thread::hardware_concurrency();
    for (int nt : {1, 2, 4, 8, 16, 32, 64}) {
        if ((unsigned)nt > hw) break;
        for (int mode = 0; mode < 2; mode++) {
            double best = 0;
            for (int rep = 0; rep < 3; rep++) {
                auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> th;
                for (int t = 0; t < nt; t++)
                    th.emplace_back([=] {
                        size_t per = total / nt / 4096 * 4096, off = per * t;
                        if (mode) copy_nt(dst + off, src + off, per);
                        else memcpy(dst + off, src + off, per);
                    });
                for (auto& x : th) x.join();
                double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                double gbs = total / s / 1e9;
                if (gbs > best) best = gbs;
            }
            printf("%2d threads  %-9s %7.1f GB/s\n", nt, mode ? "streaming" : "memcpy", best);
        }
    }
    return 0;
}
