// h2d_ceiling.cu -- the staging ceiling of this box: concurrent pinned host-to-device copies on 1, 2, 4, 8 GPUs.
// bench.py's end-to-end numbers include the H2D copy of every batch (272 MB per step and GPU for BASELINE cfg 2), so
// their multi-GPU scaling is bounded by what the host can feed; this measures that bound with nothing else running.
// One thread per device, each with its own pinned buffer (plain cudaHostAlloc, or write-combined with `wc`), a start
// barrier, `reps` back-to-back cudaMemcpyAsync of `mb` MB; aggregate = all bytes / slowest thread's wall time.
//   nvcc -O2 -o tools/h2d_ceiling tools/h2d_ceiling.cu ; tools/h2d_ceiling [mb=256] [reps=10]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

static double run(int ndev, size_t bytes, int reps, unsigned flags, bool d2h) {
    std::vector<std::thread> th;
    std::vector<double> secs(ndev, 0.0);
    std::atomic<int> ready{0}, go{0};
    for (int d = 0; d < ndev; d++)
        th.emplace_back([&, d] {
            cudaSetDevice(d);
            void *h = nullptr, *g = nullptr;
            cudaStream_t st;
            cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
            if (cudaHostAlloc(&h, bytes, flags) != cudaSuccess || cudaMalloc(&g, bytes) != cudaSuccess) {
                fprintf(stderr, "alloc failed on device %d\n", d);
                exit(1);
            }
            memset(h, d + 1, bytes);
            for (int i = 0; i < 2; i++) cudaMemcpyAsync(d2h ? h : g, d2h ? g : h, bytes, d2h ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice, st);
            cudaStreamSynchronize(st);
            ready++;
            while (!go.load()) std::this_thread::yield();
            const auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < reps; i++) cudaMemcpyAsync(d2h ? h : g, d2h ? g : h, bytes, d2h ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice, st);
            cudaStreamSynchronize(st);
            secs[d] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            cudaFreeHost(h);
            cudaFree(g);
            cudaStreamDestroy(st);
        });
    while (ready.load() < ndev) std::this_thread::yield();
    go = 1;
    for (auto &t : th) t.join();
    double worst = 0;
    for (double s : secs) worst = s > worst ? s : worst;
    return (double)bytes * reps * ndev / worst / 1e9;
}

int main(int argc, char **argv) {
    const size_t mb = argc > 1 ? (size_t)atoll(argv[1]) : 256;
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    for (int n = 1; n <= count; n *= 2) {
        const double plain = run(n, mb << 20, reps, cudaHostAllocDefault, false);
        const double wc = run(n, mb << 20, reps, cudaHostAllocWriteCombined, false);
        const double portable = run(n, mb << 20, reps, cudaHostAllocPortable, false);
        const double d2h = run(n, mb << 20, reps, cudaHostAllocDefault, true);
        printf("{\"gpus\": %d, \"mb_per_copy\": %zu, \"reps\": %d, \"h2d_gbs\": %.1f, \"h2d_wc_gbs\": %.1f, \"h2d_portable_gbs\": %.1f, "
               "\"d2h_gbs\": %.1f, \"h2d_gbs_per_gpu\": %.1f}\n",
               n, mb, reps, plain, wc, portable, d2h, plain / n);
        fflush(stdout);
    }
    return 0;
}
