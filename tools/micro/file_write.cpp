// Microbenchmark behind the occurrence writer (cli.cpp: emitText): T threads put N GB of text into ONE new file in tmpfs --
// mode 0: memcpy into a shared mapping (what the CLI does), 1: MADV_POPULATE_WRITE per piece first, 2: pwrite per piece.
// usage: file_write [GB] [threads] [mode] [path]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
int main(int argc, char** argv) {
    const size_t gb = argc > 1 ? atoi(argv[1]) : 2; const int T = argc > 2 ? atoi(argv[2]) : 8; const int mode = argc > 3 ? atoi(argv[3]) : 0;
    const char* path = argc > 4 ? argv[4] : "/dev/shm/wtest.bin";
    const size_t n = gb << 30;
    char* src = (char*)malloc(n); memset(src, 'x', n);
    int fd = open(path, O_CREAT | O_TRUNC | O_RDWR, 0644);
    auto t0 = std::chrono::steady_clock::now();
    if (ftruncate(fd, n)) return 1;
    char* m = (char*)mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    std::vector<std::thread> th;
    const size_t piece = 16u << 20; const size_t np = n / piece;
    for (int t = 0; t < T; t++) th.emplace_back([&, t] {
        for (size_t i = t; i < np; i += T) {
            if (mode == 1) madvise(m + i * piece, piece, MADV_POPULATE_WRITE);
            if (mode == 2) { if (pwrite(fd, src + i * piece, piece, i * piece) < 0) perror("pwrite"); continue; }
            memcpy(m + i * piece, src + i * piece, piece);
        }
    });
    for (auto& x : th) x.join();
    munmap(m, n);
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("mode %d threads %d: %.2f s = %.2f GB/s\n", mode, T, s, gb * 1.073741824 / s);
    close(fd); unlink(path); return 0;
}
