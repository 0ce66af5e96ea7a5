// tools/tmpfs_write_bench.cpp -- how fast can T threads fill a fresh file (the SAM sink's problem)?
//   usage: tmpfs_write_bench <path> <mode> <threads> <GiB>
//   mode 0: ftruncate + shared mapping, threads memcpy 8 MiB chunks (what SamSink does)
//        1: fallocate the whole file first, then as 0        2: pwrite of 8 MiB chunks at disjoint offsets
//        3: fallocate running ahead on its own thread         4: MADV_POPULATE_WRITE running ahead on its own thread
// Diagnostic only; built by hand (g++ -O2 -pthread), not part of the library.
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <thread>
#include <vector>
#include <chrono>
#include <atomic>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
    const char *path = argv[1];
    const int mode = atoi(argv[2]);
    const int T = atoi(argv[3]);
    const size_t GB = (size_t)atoi(argv[4]);
    const size_t N = GB << 30, CH = 8u << 20;
    std::vector<char> src(CH, 'A');
    unlink(path);
    int fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
    double t0 = now();
    if (mode == 0 || mode == 1 || mode == 3 || mode == 4) { if (ftruncate(fd, N)) return 1; }
    char *m = nullptr;
    if (mode == 0 || mode == 1 || mode == 3 || mode == 4) m = (char *)mmap(nullptr, N, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    std::atomic<size_t> next{0}, falloc{0};
    std::vector<std::thread> th;
    double tf = 0;
    if (mode == 1) {   // fallocate everything first (single thread), then memcpy
        double a = now();
        if (fallocate(fd, 0, 0, N)) perror("fallocate");
        tf = now() - a;
    }
    std::thread fa;
    if (mode == 3) {   // fallocate ahead on its own thread, chunk by chunk
        fa = std::thread([&] { for (size_t o = 0; o < N; o += CH) { if (fallocate(fd, 0, o, CH)) perror("fallocate"); falloc.store(o + CH); } });
    }
    if (mode == 4) {   // MADV_POPULATE_WRITE ahead on its own thread
        fa = std::thread([&] { for (size_t o = 0; o < N; o += CH) { if (madvise(m + o, CH, 23 /*MADV_POPULATE_WRITE*/)) perror("madvise"); falloc.store(o + CH); } });
    }
    for (int t = 0; t < T; ++t)
        th.emplace_back([&] {
            for (;;) {
                size_t o = next.fetch_add(CH);
                if (o >= N) break;
                if (mode == 3 || mode == 4) while (falloc.load() < o + CH) std::this_thread::yield();
                if (mode == 2) { if (pwrite(fd, src.data(), CH, o) != (ssize_t)CH) perror("pwrite"); }
                else memcpy(m + o, src.data(), CH);
            }
        });
    for (auto &x : th) x.join();
    if (fa.joinable()) fa.join();
    double dt = now() - t0;
    printf("mode %d threads %d: %.2f s  %.2f GB/s  (fallocate %.2f s)\n", mode, T, dt, GB / dt, tf);
    close(fd);
    unlink(path);
}
