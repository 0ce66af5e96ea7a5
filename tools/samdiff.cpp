// samdiff a.sam b.sam -- order-independent record comparison of two SAM files (header lines: @SQ compared in order,
// @PG ignored).  Every record line is reduced to a 128-bit hash; the two sorted hash multisets are merged.
// Prints one JSON object.  Used by tools/cli_scale.py to compare the drop-in's SAM file with the reference's at the
// full BASELINE size (20 M records), where a Python dict comparison would need tens of GB.
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

struct H { uint64_t a, b; bool operator<(const H &o) const { return a != o.a ? a < o.a : b < o.b; } bool operator==(const H &o) const { return a == o.a && b == o.b; } };

static inline uint64_t mix(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33; return h; }

static H hash_line(const char *p, size_t n) {
    uint64_t a = 0x9e3779b97f4a7c15ull ^ n, b = 0xc2b2ae3d27d4eb4full + n;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        a = mix(a ^ w);
        b = (b + w) * 0x100000001b3ull;
        b ^= b >> 29;
    }
    uint64_t w = 0;
    memcpy(&w, p + i, n - i);
    a = mix(a ^ w ^ 0xabcdefull);
    b = mix(b + w);
    return H{a, b};
}

struct File {
    const char *p = nullptr;
    size_t n = 0;
    std::vector<std::string> sq;
    std::vector<H> recs;
};

static bool load(const char *path, File &f) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat sb;
    fstat(fd, &sb);
    f.n = (size_t)sb.st_size;
    f.p = f.n ? (const char *)mmap(nullptr, f.n, PROT_READ, MAP_PRIVATE, fd, 0) : "";
    close(fd);
    if (f.n && f.p == MAP_FAILED) return false;
    // header
    size_t o = 0;
    while (o < f.n && f.p[o] == '@') {
        const char *e = (const char *)memchr(f.p + o, '\n', f.n - o);
        size_t len = e ? (size_t)(e - (f.p + o)) : f.n - o;
        if (len >= 3 && memcmp(f.p + o, "@PG", 3) != 0) f.sq.emplace_back(f.p + o, len);
        o += len + 1;
    }
    // records, in parallel over byte ranges aligned to line starts
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::vector<H>> parts(nt);
    std::vector<std::thread> th;
    const size_t body = f.n > o ? f.n - o : 0;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t]() {
            size_t lo = o + body * t / nt, hi = o + body * (t + 1) / nt;
            if (t > 0) { while (lo < f.n && f.p[lo - 1] != '\n') ++lo; }
            if (t + 1 < nt) { while (hi < f.n && f.p[hi - 1] != '\n') ++hi; } else hi = f.n;
            std::vector<H> v;
            while (lo < hi) {
                const char *e = (const char *)memchr(f.p + lo, '\n', f.n - lo);
                size_t len = e ? (size_t)(e - (f.p + lo)) : f.n - lo;
                if (len) v.push_back(hash_line(f.p + lo, len));
                lo += len + 1;
            }
            parts[t].swap(v);
        });
    for (auto &x : th) x.join();
    for (auto &v : parts) f.recs.insert(f.recs.end(), v.begin(), v.end());
    std::sort(f.recs.begin(), f.recs.end());
    return true;
}

int main(int argc, char **argv) {
    if (argc != 3) { fprintf(stderr, "usage: samdiff a.sam b.sam\n"); return 2; }
    File a, b;
    if (!load(argv[1], a) || !load(argv[2], b)) { fprintf(stderr, "cannot read input\n"); return 2; }
    size_t i = 0, j = 0, same = 0;
    while (i < a.recs.size() && j < b.recs.size()) {
        if (a.recs[i] == b.recs[j]) { ++same; ++i; ++j; }
        else if (a.recs[i] < b.recs[j]) ++i;
        else ++j;
    }
    printf("{\"records_a\": %zu, \"records_b\": %zu, \"identical\": %zu, \"pct\": %.6f, \"header_equal\": %s}\n", a.recs.size(),
           b.recs.size(), same, a.recs.empty() ? 0.0 : 100.0 * same / std::max(a.recs.size(), b.recs.size()),
           a.sq == b.sq ? "true" : "false");
    return 0;
}
