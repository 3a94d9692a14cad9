// Does a host-to-device copy run slower when the page-locked source was just written by CPU threads?  (The staging path of
// hsk_count copies a pageable DnaBuffer into page-locked ring slots and sends them right away.)
//   nvcc -O2 -o h2d_after_write h2d_after_write.cu && ./h2d_after_write
#include <cuda_runtime.h>
#include <emmintrin.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <thread>
#include <vector>
// copy with non-temporal stores: the destination lines go to memory instead of staying dirty in the writers' caches
static void stream_copy(unsigned char *dst, const unsigned char *src, size_t n)
{
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 15)) { dst[i] = src[i]; ++i; }
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i *)(src + i)), b = _mm_loadu_si128((const __m128i *)(src + i + 16));
        const __m128i c = _mm_loadu_si128((const __m128i *)(src + i + 32)), d = _mm_loadu_si128((const __m128i *)(src + i + 48));
        _mm_stream_si128((__m128i *)(dst + i), a); _mm_stream_si128((__m128i *)(dst + i + 16), b);
        _mm_stream_si128((__m128i *)(dst + i + 32), c); _mm_stream_si128((__m128i *)(dst + i + 48), d);
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
    const size_t nb = 37500000, nchunk = 8, chunk = (nb + nchunk - 1) / nchunk;
    unsigned char *pin, *dev;
    std::vector<unsigned char> src(nb, 3);
    cudaHostAlloc(&pin, nb, cudaHostAllocDefault);
    cudaMalloc(&dev, nb);
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    memset(pin, 1, nb);
    auto copy = [&](const char *what, bool chunked) {
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0, s);
            if (chunked) for (size_t o = 0; o < nb; o += chunk) cudaMemcpyAsync(dev + o, pin + o, std::min(chunk, nb - o), cudaMemcpyHostToDevice, s);
            else cudaMemcpyAsync(dev, pin, nb, cudaMemcpyHostToDevice, s);
            cudaEventRecord(e1, s);
            cudaStreamSynchronize(s);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            best = std::min(best, ms);
        }
        printf("%-60s %.3f ms  %.1f GB/s\n", what, best, nb / best / 1e6);
    };
    copy("idle source, one copy", false);
    copy("idle source, 8 copies", true);
    for (unsigned T : {1u, 4u, 16u}) {
        float best = 1e9; double bestcpu = 0;
        for (int rep = 0; rep < 5; ++rep) {
            const double t0 = now();
            std::vector<std::thread> th;
            for (unsigned t = 0; t < T; ++t) th.emplace_back([&, t] { const size_t lo = nb * t / T, hi = nb * (t + 1) / T; memcpy(pin + lo, src.data() + lo, hi - lo); });
            for (auto &x : th) x.join();
            const double t1 = now();
            cudaEventRecord(e0, s);
            for (size_t o = 0; o < nb; o += chunk) cudaMemcpyAsync(dev + o, pin + o, std::min(chunk, nb - o), cudaMemcpyHostToDevice, s);
            cudaEventRecord(e1, s);
            cudaStreamSynchronize(s);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) { best = ms; bestcpu = t1 - t0; }
        }
        printf("source just written by %2u threads (%.2f ms), 8 copies:        %.3f ms  %.1f GB/s\n", T, bestcpu * 1e3, best, nb / best / 1e6);
    }
    for (unsigned T : {4u, 16u}) {
        float best = 1e9; double bestcpu = 0;
        for (int rep = 0; rep < 5; ++rep) {
            const double t0 = now();
            std::vector<std::thread> th;
            for (unsigned t = 0; t < T; ++t) th.emplace_back([&, t] { const size_t lo = nb * t / T, hi = nb * (t + 1) / T; stream_copy(pin + lo, src.data() + lo, hi - lo); });
            for (auto &x : th) x.join();
            const double t1 = now();
            cudaEventRecord(e0, s);
            for (size_t o = 0; o < nb; o += chunk) cudaMemcpyAsync(dev + o, pin + o, std::min(chunk, nb - o), cudaMemcpyHostToDevice, s);
            cudaEventRecord(e1, s);
            cudaStreamSynchronize(s);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) { best = ms; bestcpu = t1 - t0; }
        }
        printf("source just written by %2u threads, streaming stores (%.2f ms): %.3f ms  %.1f GB/s\n", T, bestcpu * 1e3, best, nb / best / 1e6);
    }
    // copies issued while other threads keep writing another page-locked buffer
    {
        unsigned char *pin2; cudaHostAlloc(&pin2, nb, cudaHostAllocDefault);
        std::vector<std::thread> th; volatile bool stop = false;
        for (unsigned t = 0; t < 8; ++t) th.emplace_back([&, t] { while (!stop) { const size_t lo = nb * t / 8, hi = nb * (t + 1) / 8; memcpy(pin2 + lo, src.data() + lo, hi - lo); } });
        copy("idle source, 8 copies, 8 threads writing elsewhere", true);
        stop = true; for (auto &x : th) x.join();
    }
    return 0;
}
