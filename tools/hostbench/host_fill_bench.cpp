// Host-side micro-benchmark behind the end-to-end path of kmer_count (no GPU): what does it cost to build a fresh
// std::vector-sized array of 16-byte entries from SoA arrays with T threads taking parts of P entries (ListBuilder in
// hysortk_b200/cxx/hysortk.cpp), with and without transparent huge pages, and to copy a pageable buffer into a ring of
// staging buffers in pieces (stage_input in csrc/engine.cu)?   g++ -O3 -pthread host_fill_bench.cpp -o host_fill_bench
#include <sys/mman.h>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>
struct E { uint64_t k, c; };
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
    const size_t n = 6200000;
    std::vector<uint64_t> w(n); std::vector<uint32_t> c(n);
    for (size_t i = 0; i < n; ++i) { w[i] = i * 77; c[i] = i & 63; }
    const unsigned hw = std::thread::hardware_concurrency();
    printf("hardware threads %u\n", hw);
    for (int thp = 0; thp < 2; ++thp)
        for (unsigned T : {2u, 4u, 8u, 16u, 32u}) {
            if (T > 2 * hw) continue;
            for (size_t P : {size_t(65536), size_t(200000), size_t(800000)}) {
                double best = 1e9, bestfree = 0;
                for (int rep = 0; rep < 4; ++rep) {
                    std::allocator<E> al;
                    const double t0 = now();
                    E *b = al.allocate(n + 64);
                    if (thp) {
                        const uintptr_t huge = uintptr_t(2) << 20, a0 = ((uintptr_t)b + huge - 1) & ~(huge - 1), a1 = ((uintptr_t)b + n * 16) & ~(huge - 1);
                        if (a1 > a0) madvise((void *)a0, a1 - a0, MADV_HUGEPAGE);
                    }
                    std::atomic<size_t> next{0};
                    std::vector<std::thread> th;
                    for (unsigned t = 0; t < T; ++t) th.emplace_back([&] {
                        for (size_t f = next.fetch_add(P); f < n; f = next.fetch_add(P)) {
                            const size_t e = f + P < n ? f + P : n;
                            for (size_t i = f; i < e; ++i) { b[i].k = w[i]; b[i].c = c[i]; }
                        }
                    });
                    for (auto &x : th) x.join();
                    const double t1 = now();
                    al.deallocate(b, n + 64);
                    const double t2 = now();
                    if (t1 - t0 < best) { best = t1 - t0; bestfree = t2 - t1; }
                }
                printf("fill thp=%d threads=%2u part=%7zu: %.2f ms (free %.2f ms)\n", thp, T, P, best * 1e3, bestfree * 1e3);
            }
        }
    // warm destination (memory recycled by the allocator): the floor
    {
        std::vector<E> keep(n);
        for (unsigned T : {4u, 8u, 16u}) {
            double best = 1e9;
            for (int rep = 0; rep < 4; ++rep) {
                const double t0 = now();
                std::atomic<size_t> next{0};
                std::vector<std::thread> th;
                for (unsigned t = 0; t < T; ++t) th.emplace_back([&] {
                    for (size_t f = next.fetch_add(65536); f < n; f = next.fetch_add(65536)) {
                        const size_t e = f + 65536 < n ? f + 65536 : n;
                        for (size_t i = f; i < e; ++i) { keep[i].k = w[i]; keep[i].c = c[i]; }
                    }
                });
                for (auto &x : th) x.join();
                best = std::min(best, now() - t0);
            }
            printf("fill warm threads=%2u: %.2f ms\n", T, best * 1e3);
        }
    }
    // staging copy: 37.5 MB pageable -> ring of 1 MiB slots
    {
        const size_t nb = 37500000, piece = 1 << 20;
        std::vector<uint8_t> src(nb, 1), ring(64 * piece);
        for (unsigned T : {1u, 2u, 4u, 8u, 16u}) {
            double best = 1e9;
            for (int rep = 0; rep < 5; ++rep) {
                const double t0 = now();
                std::atomic<size_t> next{0};
                std::vector<std::thread> th;
                for (unsigned t = 0; t < T; ++t) th.emplace_back([&] {
                    for (size_t j = next.fetch_add(1); j * piece < nb; j = next.fetch_add(1))
                        memcpy(ring.data() + (j % 64) * piece, src.data() + j * piece, std::min(piece, nb - j * piece));
                });
                for (auto &x : th) x.join();
                best = std::min(best, now() - t0);
            }
            printf("staging copy threads=%2u: %.2f ms (%.1f GB/s)\n", T, best * 1e3, nb / best / 1e9);
        }
    }
    return 0;
}
