// samdiff a.sam b.sam -- order-independent record comparison of two SAM files (header lines: @SQ compared in order,
// @PG ignored).  Every record line is reduced to a 128-bit hash; the two sorted hash multisets are merged.
// Prints one JSON object.  Used by tools/cli_scale.py and bench.py to compare the drop-in's SAM file with the reference's
// at the full BASELINE size (20 M records), where a Python dict comparison would need tens of GB.  Records whose QNAME
// has the form <group>.<rest> are also counted per group ("groups": bench.py maps several workloads in one run).
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct H { uint64_t a, b; uint32_t g; uint32_t len; uint64_t off; bool operator<(const H &o) const { return a != o.a ? a < o.a : b < o.b; } bool operator==(const H &o) const { return a == o.a && b == o.b; } };

static std::mutex g_mu;
static std::vector<std::string> g_groups{std::string()};   // group names in order of first appearance; [0] = "" (no group);
                                                           // at most 16 groups are told apart, the rest counts under ""
static uint32_t group_of(const char *p, size_t n) {   // QNAME up to the first '.', "" when there is none
    size_t k = 0;
    while (k < n && p[k] != '\t' && p[k] != '.') ++k;
    std::string name = (k < n && p[k] == '.') ? std::string(p, k) : std::string();
    std::lock_guard<std::mutex> lk(g_mu);
    for (size_t i = 0; i < g_groups.size(); ++i) if (g_groups[i] == name) return (uint32_t)i;
    if (g_groups.size() >= 16) return 0;
    g_groups.push_back(name);
    return (uint32_t)g_groups.size() - 1;
}

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
    return H{a, b, 0, 0, 0};
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
            uint32_t cur = 0;
            size_t curlen = (size_t)-1;   // length of the current group's name (records of a group are contiguous)
            const char *curp = nullptr;
            while (lo < hi) {
                const char *e = (const char *)memchr(f.p + lo, '\n', f.n - lo);
                size_t len = e ? (size_t)(e - (f.p + lo)) : f.n - lo;
                if (len) {
                    H h = hash_line(f.p + lo, len);
                    const char *q = f.p + lo;
                    size_t k = 0;
                    while (k < len && q[k] != '\t' && q[k] != '.') ++k;
                    const size_t gl = (k < len && q[k] == '.') ? k : 0;
                    if (curlen != gl || (gl && memcmp(curp, q, gl) != 0)) { cur = group_of(q, len); curlen = gl; curp = q; }
                    h.g = cur;
                    h.len = (uint32_t)len;
                    h.off = (uint64_t)lo;
                    v.push_back(h);
                }
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
    if (argc != 3 && argc != 4) { fprintf(stderr, "usage: samdiff a.sam b.sam [file for up to 400 unmatched records]\n"); return 2; }
    FILE *fd = argc == 4 ? fopen(argv[3], "w") : nullptr;
    size_t shown = 0;
    auto show = [&](const File &f, const H &h, char tag) {
        if (!fd || shown >= 400) return;
        ++shown;
        fprintf(fd, "%c\t%.*s\n", tag, (int)std::min<uint32_t>(h.len, 120), f.p + h.off);   // up to the CIGAR; not the bases
    };
    File a, b;
    if (!load(argv[1], a) || !load(argv[2], b)) { fprintf(stderr, "cannot read input\n"); return 2; }
    size_t i = 0, j = 0, same = 0;
    std::vector<size_t> ga(g_groups.size(), 0), gb(g_groups.size(), 0), gs(g_groups.size(), 0);
    for (auto &h : a.recs) ++ga[h.g];
    for (auto &h : b.recs) ++gb[h.g];
    while (i < a.recs.size() && j < b.recs.size()) {
        if (a.recs[i] == b.recs[j]) { ++same; ++gs[a.recs[i].g]; ++i; ++j; }
        else if (a.recs[i] < b.recs[j]) { show(a, a.recs[i], 'a'); ++i; }
        else { show(b, b.recs[j], 'b'); ++j; }
    }
    for (; i < a.recs.size(); ++i) show(a, a.recs[i], 'a');
    for (; j < b.recs.size(); ++j) show(b, b.recs[j], 'b');
    if (fd) fclose(fd);
    printf("{\"records_a\": %zu, \"records_b\": %zu, \"identical\": %zu, \"pct\": %.6f, \"header_equal\": %s, \"groups\": {", a.recs.size(),
           b.recs.size(), same, a.recs.empty() ? 0.0 : 100.0 * same / std::max(a.recs.size(), b.recs.size()),
           a.sq == b.sq ? "true" : "false");
    for (size_t g = 0; g < g_groups.size(); ++g)
        printf("%s\"%s\": {\"records_a\": %zu, \"records_b\": %zu, \"identical\": %zu, \"pct\": %.6f}", g ? ", " : "", g_groups[g].c_str(),
               ga[g], gb[g], gs[g], std::max(ga[g], gb[g]) ? 100.0 * gs[g] / std::max(ga[g], gb[g]) : 0.0);
    printf("}}\n");
    return 0;
}
