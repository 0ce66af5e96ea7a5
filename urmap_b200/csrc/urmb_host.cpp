// urmb_host.cpp -- host side of the drop-in: URMAP's command line, FASTQ/FASTA readers, SAM writer and the
// byte-identical CPU -make_ufi, all above the C-ABI of include/urmb.h.  The per-read search itself only ever
// runs in the CUDA kernels behind urmb_submit/urmb_wait; there is no CPU mapping path in this file.
//
//   urmap_b200 -make_ufi ref.fa -output ref.ufi [-wordlength 24] [-maxix 32] [-slots N] [-load_factor 0.6] [-veryfast] [-gpu_build]
//   urmap_b200 -map reads.fq[.gz] -ufi ref.ufi -samout out.sam [-threads N] [-veryfast] [-gpus G] [-batch N]
//   urmap_b200 -map2 R1.fq -reverse R2.fq -ufi ref.ufi -samout out.sam [-threads N] [-veryfast] [-minq 10]
//
// Reference behaviour restated here (file:line under /root/reference/src):
//   option spellings / errors   cmdline.cpp:148-269, myopts.h, getcmd.cpp:6-26, myutils.cpp:915-960
//   FASTQ records               fastqseqsource.cpp:9-116, linereader.cpp:54-99
//   FASTA -> SeqData            fastaseqsource.cpp:26-112, ufindex.cpp:462-510 (upper-case, 32 x '-' pads)
//   index construction          ufindex.cpp:83-408, 945-1000; ufindexio.cpp:15-49,117-179; prime.cpp:11
//   SAM text                    setsam.cpp:12-207, output1.cpp:8-30, output2.cpp:18-132, state1.cpp:129-145,736-752
//   CIGAR                       cigar.cpp:4-41,141-199, state1.cpp:717-734
//   end-of-run summary          state1.cpp:593-632
#include <ctype.h>
#include <omp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/urmb.h"

extern "C" int urmb_build_index_device(const void *d_seq, uint64_t seq_data_size, uint64_t slot_count,
                                       uint32_t word_length, uint32_t max_ix, void *d_blob, uint64_t *stats);
extern "C" const char *urmb_build_last_error();

#define URMB_VERSION "0.1"

static std::vector<std::string> g_argv;
static bool g_quiet = false;
static FILE *g_log = nullptr;

[[noreturn]] static void Die(const char *fmt, ...) {  // myutils.cpp:915-960
    char msg[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg, sizeof msg, fmt, ap);
    va_end(ap);
    for (FILE *f : {stderr, g_log}) {
        if (!f) continue;
        fprintf(f, "\n");
        for (auto &a : g_argv) fprintf(f, "%s ", a.c_str());
        fprintf(f, "\n\n---Fatal error---\n%s\n", msg);
        fflush(f);
    }
    exit(1);
}

static void Progress(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    if (!g_quiet) {
        va_list ap2;
        va_copy(ap2, ap);
        vfprintf(stderr, fmt, ap2);
        va_end(ap2);
    }
    if (g_log) vfprintf(g_log, fmt, ap);
    va_end(ap);
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------------------------------------
// options
// ------------------------------------------------------------------------------------------------
struct Opts {
    std::string make_ufi, map, map2, reverse, ufi, samout, output, log, slots;
    unsigned threads = 0, wordlength = 24, maxix = 32, minq = 10, gpus = 1, batch = 262144;
    double load_factor = 0.6;
    bool veryfast = false, quiet = false, gpu_build = false, version = false;
    bool set_maxix = false, set_wordlength = false, set_threads = false;
};

static Opts ParseCmdLine(int argc, char **argv) {
    Opts o;
    for (int i = 0; i < argc; ++i) g_argv.push_back(argv[i]);
    auto bad = [&](const std::string &why) {
        fprintf(stderr, "\nInvalid command line\n%s\n\n", why.c_str());  // cmdline.cpp:28-38
        exit(1);
    };
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a.size() < 2 || a[0] != '-') bad("Expected -option_name, got '" + a + "'");
        std::string name = a.substr(a[1] == '-' ? 2 : 1);
        auto val = [&]() -> std::string {
            if (i + 1 >= argc) bad("Missing value for option -" + name);
            return std::string(argv[++i]);
        };
        if (name == "make_ufi") o.make_ufi = val();
        else if (name == "map") o.map = val();
        else if (name == "map2") o.map2 = val();
        else if (name == "reverse") o.reverse = val();
        else if (name == "ufi") o.ufi = val();
        else if (name == "samout") o.samout = val();
        else if (name == "output") o.output = val();
        else if (name == "log") o.log = val();
        else if (name == "slots") o.slots = val();
        else if (name == "threads") { o.threads = (unsigned)strtoul(val().c_str(), nullptr, 10); o.set_threads = true; }
        else if (name == "wordlength") { o.wordlength = (unsigned)strtoul(val().c_str(), nullptr, 10); o.set_wordlength = true; }
        else if (name == "maxix") { o.maxix = (unsigned)strtoul(val().c_str(), nullptr, 10); o.set_maxix = true; }
        else if (name == "minq") o.minq = (unsigned)strtoul(val().c_str(), nullptr, 10);
        else if (name == "load_factor") o.load_factor = atof(val().c_str());
        else if (name == "gpus") o.gpus = (unsigned)strtoul(val().c_str(), nullptr, 10);
        else if (name == "batch") o.batch = (unsigned)strtoul(val().c_str(), nullptr, 10);
        else if (name == "veryfast") o.veryfast = true;
        else if (name == "quiet") o.quiet = true;
        else if (name == "gpu_build") o.gpu_build = true;
        else if (name == "version") o.version = true;
        else bad("Unknown option " + name);
    }
    int ncmd = (!o.make_ufi.empty()) + (!o.map.empty()) + (!o.map2.empty()) + (o.version ? 1 : 0);
    if (ncmd == 0) bad("No command specified");       // getcmd.cpp:6-11
    if (ncmd > 1) bad("Two commands specified");
    return o;
}

// ------------------------------------------------------------------------------------------------
// line reader (plain or .gz), linereader.cpp:54-99: CR dropped, last line may lack LF
// ------------------------------------------------------------------------------------------------
class LineReader {
   public:
    explicit LineReader(const std::string &path) : path_(path) {
        gz_ = gzopen(path.c_str(), "rb");  // transparently reads uncompressed files too
        if (!gz_) Die("Cannot open %s", path.c_str());
        gzbuffer(gz_, 1 << 20);
        buf_.resize(32u << 20);
    }
    ~LineReader() { if (gz_) gzclose(gz_); }
    // returns false at EOF with nothing read; line excludes the terminator
    bool ReadLine(const char *&p, size_t &n) {
        line_.clear();
        bool any = false;
        for (;;) {
            if (off_ >= len_) {
                if (eof_) break;
                int r = gzread(gz_, buf_.data(), (unsigned)buf_.size());
                if (r < 0) Die("Read error in %s", path_.c_str());
                len_ = (size_t)r;
                off_ = 0;
                if (r == 0) { eof_ = true; break; }
            }
            const char *b = buf_.data() + off_;
            const char *e = (const char *)memchr(b, '\n', len_ - off_);
            size_t m = e ? (size_t)(e - b) : len_ - off_;
            if (line_.empty() && e) {  // fast path: whole line inside the buffer
                ++line_nr_;
                off_ += m + 1;
                if (memchr(b, '\r', m) == nullptr) { p = b; n = m; return true; }
                append_nocr(b, m);
                p = line_.data(); n = line_.size();
                return true;
            }
            append_nocr(b, m);
            any = true;
            off_ += m + (e ? 1 : 0);
            if (e) { ++line_nr_; p = line_.data(); n = line_.size(); return true; }
        }
        if (!any && line_.empty()) return false;
        ++line_nr_;
        p = line_.data(); n = line_.size();
        return n > 0;
    }
    unsigned line_nr() const { return line_nr_; }
    const std::string &path() const { return path_; }

   private:
    void append_nocr(const char *b, size_t m) {
        for (size_t i = 0; i < m; ++i) if (b[i] != '\r') line_.push_back(b[i]);
    }
    std::string path_;
    gzFile gz_ = nullptr;
    std::vector<char> buf_;
    size_t off_ = 0, len_ = 0;
    bool eof_ = false;
    std::string line_;
    unsigned line_nr_ = 0;
};

// ------------------------------------------------------------------------------------------------
// FASTQ batches
// ------------------------------------------------------------------------------------------------
struct HostBatch {
    std::vector<uint8_t> seqs, quals, labels;
    std::vector<uint32_t> offs{0}, loffs{0};
    uint32_t n = 0;
    void clear() { seqs.clear(); quals.clear(); labels.clear(); offs.assign(1, 0); loffs.assign(1, 0); n = 0; }
};

class FastqReader {  // FASTQSeqSource::GetNextLo, fastqseqsource.cpp:9-116
   public:
    explicit FastqReader(const std::string &path) : lr_(path) {}
    // appends up to max_reads records; returns number read
    uint32_t Fill(HostBatch &b, uint32_t max_reads) {
        uint32_t got = 0;
        const char *p;
        size_t n;
        while (got < max_reads) {
            if (!lr_.ReadLine(p, n)) break;
            if (n == 0) {  // empty lines are only allowed at EOF
                unsigned ln = lr_.line_nr();
                while (lr_.ReadLine(p, n))
                    if (n != 0) Die("Empty line nr %u in FASTQ file '%s'", ln, lr_.path().c_str());
                break;
            }
            if (p[0] != '@') Die("Bad line %u in FASTQ file '%s': expected '@'", lr_.line_nr(), lr_.path().c_str());
            b.labels.insert(b.labels.end(), p + 1, p + n);
            b.loffs.push_back((uint32_t)b.labels.size());
            if (!lr_.ReadLine(p, n)) Die("Unexpected end-of-file in FASTQ file %s", lr_.path().c_str());
            const size_t L = n;
            for (size_t i = 0; i < L; ++i) {
                unsigned char c = (unsigned char)p[i];
                if (!isalpha(c)) {
                    if (isprint(c)) Die("Invalid sequence letter '%c' in FASTQ, line %u file %s", c, lr_.line_nr(), lr_.path().c_str());
                    Die("Non-printing byte 0x%02x in FASTQ sequence line %u file %s", c, lr_.line_nr(), lr_.path().c_str());
                }
            }
            b.seqs.insert(b.seqs.end(), p, p + L);
            b.offs.push_back((uint32_t)b.seqs.size());
            lr_.ReadLine(p, n);  // '+' line, contents ignored
            if (!lr_.ReadLine(p, n)) Die("Unexpected end-of-file in FASTQ file %s", lr_.path().c_str());
            if (n != L) Die("Bad FASTQ record: %u bases, %u quals line %u file %s", (unsigned)L, (unsigned)n, lr_.line_nr(), lr_.path().c_str());
            b.quals.insert(b.quals.end(), p, p + n);
            ++b.n;
            ++got;
        }
        return got;
    }

   private:
    LineReader lr_;
};

// ------------------------------------------------------------------------------------------------
// alphabet helpers (alpha.cpp:1309, 3005, 3525)
// ------------------------------------------------------------------------------------------------
static uint8_t g_Letter[256], g_CompLetter[256], g_CompChar[256];
static void InitAlpha() {
    memset(g_Letter, 0xFF, 256);
    memset(g_CompLetter, 0xFF, 256);
    memset(g_CompChar, '?', 256);
    const char *s = "ACGTU";
    const uint8_t v[5] = {0, 1, 2, 3, 3};
    for (int i = 0; i < 5; ++i) {
        g_Letter[(uint8_t)s[i]] = v[i];
        g_Letter[(uint8_t)(s[i] | 0x20)] = v[i];
        g_CompLetter[(uint8_t)s[i]] = 3 - v[i];
        if (s[i] != 'U') g_CompLetter[(uint8_t)(s[i] | 0x20)] = 3 - v[i];
    }
    const char *from = "ABCDGHKMNRSTUVWXY", *to = "TVGHCDMKNYSAABWXR";
    for (int i = 0; from[i]; ++i) {
        g_CompChar[(uint8_t)from[i]] = (uint8_t)to[i];
        if (from[i] != 'U') g_CompChar[(uint8_t)(from[i] | 0x20)] = (uint8_t)(to[i] | 0x20);
    }
}

// ------------------------------------------------------------------------------------------------
// SAM
// ------------------------------------------------------------------------------------------------
struct Contigs {
    std::vector<std::string> labels;
    std::vector<uint32_t> lengths, offsets;
    // UFIndex::PosToCoordL, ufindex.cpp:729-755
    uint32_t PosToCoordL(uint32_t Pos, int &idx, uint32_t &L) const {
        int64_t Lo = 0, Hi = (int64_t)labels.size() - 1;
        while (Lo <= Hi) {
            int64_t k = (Lo + Hi) / 2;
            uint32_t Off = offsets[k], Len = lengths[k];
            if (Pos >= Off && Pos < Off + Len) { idx = (int)k; L = Len; return Pos - Off; }
            if (Pos > Off) Lo = k + 1; else Hi = k - 1;
        }
        idx = -1;
        L = 0;
        return UINT32_MAX;
    }
};

static inline void put_u(std::string &o, uint32_t v) { char t[16]; int n = snprintf(t, sizeof t, "%u", v); o.append(t, n); }
static inline void put_i(std::string &o, int v) { char t[16]; int n = snprintf(t, sizeof t, "%d", v); o.append(t, n); }

// PathToCIGAR (cigar.cpp:4-41, D<->I swapped) + CIGAROpsFixDanglingMs (cigar.cpp:141-199).  The reference's
// second fix-up block cannot fire once the first has (it would need a length that is both <= 2 and > 4).
static void RunsToCigar(const uint16_t *runs, unsigned nruns, unsigned QL, std::string &o) {
    if (nruns == 0) { put_u(o, QL); o.push_back('M'); return; }
    char ops[512];
    unsigned lens[512];
    unsigned N = 0;
    for (unsigned i = 0; i < nruns && N < 512; ++i) {
        unsigned op = runs[i] & 3, len = runs[i] >> 2;
        char c = op == 0 ? 'M' : (op == 1 ? 'I' : 'D');
        if (N && ops[N - 1] == c) lens[N - 1] += len;
        else { ops[N] = c; lens[N] = len; ++N; }
    }
    unsigned first = 0, last = N;
    if (N >= 3) {
        if (ops[0] == 'M' && lens[0] <= 2 && lens[1] > 4 && ops[2] == 'M') { lens[2] += lens[0]; first = 1; }
        else if (ops[N - 1] == 'M' && lens[N - 1] <= 2 && lens[N - 2] > 4 && ops[N - 3] == 'M') { lens[N - 3] += lens[N - 1]; last = N - 1; }
    }
    for (unsigned i = first; i < last; ++i) { put_u(o, lens[i]); o.push_back(ops[i]); }
}

static void AppendQName(std::string &o, const uint8_t *Label, unsigned n) {  // setsam.cpp:32-44
    if (n > 2 && Label[n - 2] == '/' && (Label[n - 1] == '1' || Label[n - 1] == '2')) n -= 2;
    for (unsigned i = 0; i < n; ++i) {
        char c = (char)Label[i];
        if (c == ' ' || c == '\t') break;
        o.push_back(c);
    }
}

struct Mapped { int idx; uint32_t coord; };

static Mapped SetMappedPos(const Contigs &C, const urmb_result &r, unsigned QL) {  // state1.cpp:129-145
    Mapped m{-1, UINT32_MAX};
    if (!(r.flags & 2)) return m;
    uint32_t TL;
    int idx;
    uint32_t c = C.PosToCoordL(r.db_pos, idx, TL);
    if (c + QL > TL) return m;
    m.idx = idx;
    m.coord = c;
    return m;
}

static void SamUnmapped(std::string &o, uint32_t aFlags, const uint8_t *Label, unsigned LabelLen, const uint8_t *Seq,
                        const uint8_t *Qual, unsigned QL) {  // setsam.cpp:12-73
    uint32_t Flags = 0x04;
    if (aFlags & 0x01) Flags |= 0x01;
    if (aFlags & 0x40) Flags |= 0x40; else if (aFlags & 0x80) Flags |= 0x80;
    if (aFlags & 0x08) Flags |= 0x08; else if (aFlags & 0x20) Flags |= 0x20;
    AppendQName(o, Label, LabelLen);
    o.push_back('\t');
    put_u(o, Flags);
    o += "\t*\t0\t0\t*\t*\t0\t0\t";
    o.append((const char *)Seq, QL);
    o.push_back('\t');
    o.append((const char *)Qual, QL);
    o.push_back('\n');
}

static void SamRecord(const Contigs &C, std::string &o, uint32_t Flags, const Mapped &self, const urmb_result &r,
                      const uint16_t *runs, int MateIdx, uint32_t MatePos, int TLEN, const uint8_t *Label,
                      unsigned LabelLen, const uint8_t *Seq, const uint8_t *Qual, unsigned QL) {  // setsam.cpp:75-207
    if (self.idx < 0) { SamUnmapped(o, Flags, Label, LabelLen, Seq, Qual, QL); return; }
    const bool Plus = (r.flags & 1) != 0;
    AppendQName(o, Label, LabelLen);
    o.push_back('\t');
    put_u(o, Flags);
    o.push_back('\t');
    o += C.labels[self.idx];
    o.push_back('\t');
    put_u(o, self.coord + 1);
    o.push_back('\t');
    put_u(o, r.mapq);
    o.push_back('\t');
    RunsToCigar(runs + r.path_off, r.path_runs, QL, o);
    o.push_back('\t');
    if (MateIdx < 0 || C.labels[MateIdx].empty() || C.labels[MateIdx] == "*") o.push_back('*');
    else if (C.labels[MateIdx] == C.labels[self.idx]) o.push_back('=');
    else o += C.labels[MateIdx];
    o.push_back('\t');
    if (MatePos == 0 || MatePos == UINT32_MAX) o.push_back('0'); else put_u(o, MatePos + 1);
    o.push_back('\t');
    put_i(o, TLEN);
    o.push_back('\t');
    if (Plus) o.append((const char *)Seq, QL);
    else for (unsigned i = 0; i < QL; ++i) o.push_back((char)g_CompChar[Seq[QL - 1 - i]]);
    o.push_back('\t');
    if (Plus) o.append((const char *)Qual, QL);
    else for (unsigned i = 1; i <= QL; ++i) o.push_back((char)Qual[QL - i]);
    o.push_back('\n');
}

static uint32_t GetPairedFlags(bool First, bool RevComp, bool MateRevComp, bool MateUnmapped) {  // output2.cpp:18-36
    uint32_t Flags = First ? 0x41 : 0x81;
    if (RevComp) Flags |= 0x10;
    if (MateUnmapped) Flags |= 0x08; else if (MateRevComp) Flags |= 0x20;
    return Flags;
}

struct HitCounters { uint64_t query = 0, accept = 0, reject = 0, nohit = 0; };

static inline void UpdateHitStats(HitCounters &hc, bool has_top, unsigned mapq, unsigned minq) {  // output1.cpp:20-30
    ++hc.query;
    if (!has_top) ++hc.nohit;
    else if (mapq >= minq) ++hc.accept;
    else ++hc.reject;
}

// formats reads [lo,hi) of a finished batch
static void FormatSE(const Contigs &C, const HostBatch &b, const urmb_result *res, const uint16_t *runs, uint32_t lo,
                     uint32_t hi, unsigned minq, std::string &o, HitCounters &hc) {
    for (uint32_t i = lo; i < hi; ++i) {  // State1::Output1: SetSAM(0, "*", UINT32_MAX, 0)
        const unsigned QL = b.offs[i + 1] - b.offs[i];
        Mapped m = SetMappedPos(C, res[i], QL);
        SamRecord(C, o, 0, m, res[i], runs, -1, UINT32_MAX, 0, b.labels.data() + b.loffs[i], b.loffs[i + 1] - b.loffs[i],
                  b.seqs.data() + b.offs[i], b.quals.data() + b.offs[i], QL);
        UpdateHitStats(hc, m.idx >= 0, m.idx >= 0 ? res[i].mapq : 0, minq);
    }
}

static void FormatPE(const Contigs &C, const HostBatch &b1, const HostBatch &b2, const urmb_result *r1,
                     const urmb_result *r2, const uint16_t *runs, uint32_t lo, uint32_t hi, unsigned minq, std::string &o,
                     HitCounters &hc) {
    for (uint32_t i = lo; i < hi; ++i) {  // State2::SetSAM2, output2.cpp:71-132
        const unsigned L1 = b1.offs[i + 1] - b1.offs[i], L2 = b2.offs[i + 1] - b2.offs[i];
        Mapped m1 = SetMappedPos(C, r1[i], L1), m2 = SetMappedPos(C, r2[i], L2);
        const bool Mapped1 = m1.idx >= 0, Mapped2 = m2.idx >= 0;
        const bool Plus1 = Mapped1 && (r1[i].flags & 1), Plus2 = Mapped2 && (r2[i].flags & 1);
        const bool StrandsConsistent = Mapped1 && Mapped2 && (Plus1 != Plus2);
        int TLEN1 = 0, TLEN2 = 0;
        bool CorrectlyPaired = false;
        if (Mapped1 && Mapped2) {
            if (m1.coord <= m2.coord) {
                TLEN1 = int(m2.coord + L2) - int(m1.coord);
                if (TLEN1 > 0 && TLEN1 < 1000 && StrandsConsistent) CorrectlyPaired = true;
                if (TLEN1 > 1000) TLEN1 = 0;
                TLEN2 = -TLEN1;
            } else {
                TLEN2 = int(m1.coord + L1) - int(m2.coord);
                if (TLEN2 > 0 && TLEN2 < 1000 && StrandsConsistent) CorrectlyPaired = true;
                if (TLEN2 > 1000) TLEN2 = 0;
                TLEN1 = -TLEN2;
            }
        }
        const bool RevComp1 = Mapped1 && !(r1[i].flags & 1), RevComp2 = Mapped2 && !(r2[i].flags & 1);
        uint32_t Flags1 = GetPairedFlags(true, RevComp1, RevComp2, !Mapped2);
        uint32_t Flags2 = GetPairedFlags(false, RevComp2, RevComp1, !Mapped1);
        if (CorrectlyPaired) { Flags1 |= 0x02; Flags2 |= 0x02; }
        SamRecord(C, o, Flags1, m1, r1[i], runs, m2.idx, m2.coord, TLEN1, b1.labels.data() + b1.loffs[i],
                  b1.loffs[i + 1] - b1.loffs[i], b1.seqs.data() + b1.offs[i], b1.quals.data() + b1.offs[i], L1);
        SamRecord(C, o, Flags2, m2, r2[i], runs, m1.idx, m1.coord, TLEN2, b2.labels.data() + b2.loffs[i],
                  b2.loffs[i + 1] - b2.loffs[i], b2.seqs.data() + b2.offs[i], b2.quals.data() + b2.offs[i], L2);
        UpdateHitStats(hc, Mapped1, Mapped1 ? r1[i].mapq : 0, minq);
        UpdateHitStats(hc, Mapped2, Mapped2 ? r2[i].mapq : 0, minq);
    }
}

// ------------------------------------------------------------------------------------------------
// -map / -map2
// ------------------------------------------------------------------------------------------------
static std::string Commas(uint64_t v) {
    std::string s = std::to_string(v), o;
    for (size_t i = 0; i < s.size(); ++i) {
        o.push_back(s[i]);
        size_t rem = s.size() - 1 - i;
        if (rem && rem % 3 == 0) o.push_back(',');
    }
    return o;
}

struct InFlight {
    std::unique_ptr<HostBatch> b1, b2;
    int gpu = 0, slot = 0;
};

static int CmdMap(const Opts &o, bool paired) {
    if (o.ufi.empty()) Die("-ufi required");
    if (paired && o.reverse.empty()) Die("-reverse required");  // map2.cpp:42
    const double t_start = now_s();
    urmb_index_host *hix = nullptr;
    if (urmb_index_load_host(o.ufi.c_str(), &hix) != 0) Die("%s", urmb_last_error(nullptr));
    urmb_index_desc d;
    uint32_t ncontig = 0;
    urmb_index_info(hix, &d, &ncontig);
    Contigs C;
    for (uint32_t i = 0; i < ncontig; ++i) {
        urmb_contig c;
        urmb_index_contig(hix, i, &c);
        C.labels.push_back(c.label);
        C.lengths.push_back(c.length);
        C.offsets.push_back(c.offset);
    }
    if (o.veryfast && d.max_ix > 3) fprintf(stderr, "\nWARNING: index not optimal for -veryfast\n\n");  // map.cpp:49
    urmb_params p;
    p.method = (!paired && o.veryfast) ? 7 : 6;       // map.cpp:34-37; map2 always uses method 6 (map2.cpp:15)
    p.pe_method = (paired && o.veryfast) ? 5 : 4;     // map2.cpp:46-48
    p.band_radius = -1;
    p.minq = (int)o.minq;
    const int ngpu = (int)std::max(1u, o.gpus);
    std::vector<urmb_ctx *> ctxs(ngpu, nullptr);
    for (int g = 0; g < ngpu; ++g)
        if (urmb_ctx_create(g, &p, &ctxs[g]) != 0) Die("GPU %d: %s", g, urmb_last_error(nullptr));
    if (urmb_index_broadcast(ctxs.data(), ngpu, hix) != 0) Die("index upload: %s", urmb_last_error(ctxs[0]));
    const double t_loaded = now_s();
    Progress("Index %s loaded into %d GPU(s) in %.1f s\n", o.ufi.c_str(), ngpu, t_loaded - t_start);

    FILE *fsam = nullptr;
    if (!o.samout.empty()) {
        fsam = fopen(o.samout.c_str(), "wb");
        if (!fsam) Die("Cannot create %s", o.samout.c_str());
        setvbuf(fsam, nullptr, _IOFBF, 16 << 20);
        for (uint32_t i = 0; i < ncontig; ++i) fprintf(fsam, "@SQ\tSN:%s\tLN:%u\n", C.labels[i].c_str(), C.lengths[i]);
        fprintf(fsam, "@PG\tID:urmap\tPN:urmap\tVN:%s\tCL:", URMB_VERSION "-b200");  // state1.cpp:736-752
        for (auto &a : g_argv) fprintf(fsam, "%s ", a.c_str());
        fprintf(fsam, "\n");
    }
    int nthreads = o.set_threads ? (int)o.threads : std::min(omp_get_num_procs(), 32);
    if (nthreads < 1) nthreads = 1;
    omp_set_num_threads(nthreads);

    FastqReader rd1(paired ? o.map2 : o.map);
    std::unique_ptr<FastqReader> rd2;
    if (paired) rd2.reset(new FastqReader(o.reverse));

    // reader thread(s) -> bounded queue of batches
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::pair<std::unique_ptr<HostBatch>, std::unique_ptr<HostBatch>>> q;
    bool done = false;
    const size_t qcap = 4;
    std::thread reader([&]() {
        for (;;) {
            std::unique_ptr<HostBatch> a(new HostBatch), b;
            uint32_t n1 = 0, n2 = 0;
            if (paired) {
                b.reset(new HostBatch);
                std::thread t2([&]() { n2 = rd2->Fill(*b, o.batch); });
                n1 = rd1.Fill(*a, o.batch);
                t2.join();
                if (n1 != n2) Die("Premature end of file in FASTQ%c", n1 > n2 ? '2' : '1');  // map2.cpp:31
            } else {
                n1 = rd1.Fill(*a, o.batch);
            }
            std::unique_lock<std::mutex> lk(mu);
            if (n1 == 0) { done = true; cv.notify_all(); return; }
            cv.wait(lk, [&]() { return q.size() < qcap; });
            q.emplace_back(std::move(a), std::move(b));
            cv.notify_all();
        }
    });

    HitCounters total;
    std::deque<InFlight> fly;
    std::vector<std::string> outs(nthreads);
    std::vector<HitCounters> hcs(nthreads);
    auto finish = [&](InFlight &f) {
        const urmb_result *r1, *r2;
        const uint16_t *runs;
        uint32_t used;
        if (urmb_wait(ctxs[f.gpu], f.slot, &r1, &r2, &runs, &used) != 0) Die("GPU %d: %s", f.gpu, urmb_last_error(ctxs[f.gpu]));
        const uint32_t n = f.b1->n;
        for (auto &s : outs) s.clear();
        for (auto &h : hcs) h = HitCounters();
#pragma omp parallel num_threads(nthreads)
        {
            int t = omp_get_thread_num(), nt = omp_get_num_threads();
            uint32_t lo = (uint32_t)((uint64_t)n * t / nt), hi = (uint32_t)((uint64_t)n * (t + 1) / nt);
            if (paired) FormatPE(C, *f.b1, *f.b2, r1, r2, runs, lo, hi, o.minq, outs[t], hcs[t]);
            else FormatSE(C, *f.b1, r1, runs, lo, hi, o.minq, outs[t], hcs[t]);
        }
        for (int t = 0; t < nthreads; ++t) {
            if (fsam) fwrite(outs[t].data(), 1, outs[t].size(), fsam);
            total.query += hcs[t].query; total.accept += hcs[t].accept; total.reject += hcs[t].reject; total.nohit += hcs[t].nohit;
        }
    };
    uint64_t k = 0;
    const size_t max_fly = (size_t)ngpu * URMB_SLOTS;
    for (;;) {
        std::pair<std::unique_ptr<HostBatch>, std::unique_ptr<HostBatch>> item;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&]() { return !q.empty() || done; });
            if (q.empty()) break;
            item = std::move(q.front());
            q.pop_front();
            cv.notify_all();
        }
        if (fly.size() == max_fly) { finish(fly.front()); fly.pop_front(); }
        InFlight f;
        f.gpu = (int)(k % ngpu);
        f.slot = (int)((k / ngpu) % URMB_SLOTS);
        f.b1 = std::move(item.first);
        f.b2 = std::move(item.second);
        urmb_batch u1{f.b1->n, f.b1->seqs.data(), f.b1->offs.data()}, u2{0, nullptr, nullptr};
        if (paired) u2 = urmb_batch{f.b2->n, f.b2->seqs.data(), f.b2->offs.data()};
        if (urmb_submit(ctxs[f.gpu], f.slot, &u1, paired ? &u2 : nullptr) != 0) Die("GPU %d: %s", f.gpu, urmb_last_error(ctxs[f.gpu]));
        fly.push_back(std::move(f));
        ++k;
    }
    while (!fly.empty()) { finish(fly.front()); fly.pop_front(); }
    reader.join();
    if (fsam) fclose(fsam);
    const double t_end = now_s();
    const double secs = t_end - t_loaded;
    auto pct = [&](uint64_t x) { return total.query ? 100.0 * x / total.query : 0.0; };
    Progress("\n%16.1f  Seconds to load index\n%16.1f  Seconds in mapper\n", t_loaded - t_start, secs);  // state1.cpp:593-632
    Progress("%16s  Reads (%llu)\n", Commas(total.query).c_str(), (unsigned long long)total.query);
    Progress("%16.0f  Reads/sec. (%d GPUs, %d host threads)\n", secs > 0 ? total.query / secs : 0.0, ngpu, nthreads);
    Progress("%16s  Mapped Q>=%u (%.1f%%)\n", Commas(total.accept).c_str(), o.minq, pct(total.accept));
    Progress("%16s  Mapped Q< %u (%.1f%%)\n", Commas(total.reject).c_str(), o.minq, pct(total.reject));
    Progress("%16s  Unmapped (%.1f%%)\n\n", Commas(total.nohit).c_str(), pct(total.nohit));
    for (auto c : ctxs) urmb_ctx_destroy(c);
    urmb_index_free_host(hix);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// -make_ufi (CPU, byte-identical to the reference)
// ------------------------------------------------------------------------------------------------
static bool IsPrime64(uint64_t n) {
    if (n < 2) return false;
    if (n % 2 == 0) return n == 2;
    for (uint64_t dd = 3; dd * dd <= n; dd += 2) if (n % dd == 0) return false;
    return true;
}
static uint64_t GetPrime(uint64_t n) {  // prime.cpp:11; primes.h = first prime >= x, x = 100, x <- x*100/95
    uint64_t x = 100;
    for (int i = 0; i < 410; ++i) {
        uint64_t p = x;
        while (!IsPrime64(p)) ++p;
        if (p >= n) return p;
        x = x * 100 / 95;
    }
    Die("GetPrime(%.3g) overflow", (double)n);
}

static inline uint64_t murmur64(uint64_t h) {
    h ^= (h >> 33); h *= 0xff51afd7ed558ccdULL; h ^= (h >> 33); h *= 0xc4ceb9fe1a85ec53ULL; h ^= (h >> 33);
    return h;
}

struct UfiBuilder {
    static const uint8_t T_FREE = 0, T_END = 127, T_MY = 128, T_PLUS1 = 254, T_BOTH1 = 255, T_LONG_MINE = 253, T_LONG_OTHER = 125;
    uint32_t W = 24, MaxIx = 32;
    uint64_t SlotCount = 0, ShiftMask = 0;
    std::vector<uint8_t> Blob, Seq, CntP, CntM;
    std::vector<std::string> Labels;
    std::vector<uint32_t> Lengths, Offsets;
    uint32_t Truncated = 0;

    uint8_t Tally(uint64_t s) const { return Blob[5 * s]; }
    uint32_t Pos(uint64_t s) const { uint32_t v; memcpy(&v, &Blob[5 * s + 1], 4); return v; }
    void SetTally(uint64_t s, uint8_t t) { Blob[5 * s] = t; }
    void SetPos(uint64_t s, uint32_t p) { memcpy(&Blob[5 * s + 1], &p, 4); }
    void SetNext(uint64_t s, uint8_t nx) { SetTally(s, (uint8_t)((Tally(s) & T_MY) | nx)); }
    uint64_t Slot(uint64_t word) const { return murmur64(word & ShiftMask) % SlotCount; }

    void ReadFasta(const std::string &path) {  // fastaseqsource.cpp:26-112 (trunclabels on, gaps stripped) + ufindex.cpp:462-510
        LineReader lr(path);
        const char *p;
        size_t n;
        std::vector<std::vector<uint8_t>> seqs;
        std::vector<std::string> labs;
        bool have = false;
        while (lr.ReadLine(p, n)) {
            if (n > 0 && p[0] == '>') {
                std::string lab;
                for (size_t i = 1; i < n && !isspace((unsigned char)p[i]); ++i) lab.push_back(p[i]);
                labs.push_back(lab);
                seqs.emplace_back();
                have = true;
                continue;
            }
            if (!have) {
                if (n == 0) continue;
                Die("Bad FASTA file %s, expected '>' in line %u", path.c_str(), lr.line_nr());
            }
            auto &s = seqs.back();
            for (size_t i = 0; i < n; ++i) {
                unsigned char c = (unsigned char)p[i];
                if (isspace(c) || c == '-' || c == '.') continue;
                if (!isalpha(c)) continue;  // BadByte: counted and skipped
                s.push_back((uint8_t)toupper(c));
            }
        }
        uint64_t size = 0;
        std::vector<size_t> keep;
        for (size_t i = 0; i < seqs.size(); ++i)
            if (!seqs[i].empty()) keep.push_back(i);
            else fprintf(stderr, "\nWARNING: Empty sequence in FASTA file %s, label >%s\n\n", path.c_str(), labs[i].c_str());
        for (size_t k = 0; k < keep.size(); ++k) {
            size_t i = keep[k];
            Labels.push_back(labs[i]);
            Lengths.push_back((uint32_t)seqs[i].size());
            Offsets.push_back((uint32_t)size);
            size += seqs[i].size();
            if (k + 1 != keep.size()) size += 32;  // PADGAP
        }
        Seq.resize(size);
        uint64_t off = 0;
        for (size_t k = 0; k < keep.size(); ++k) {
            auto &s = seqs[keep[k]];
            memcpy(&Seq[off], s.data(), s.size());
            off += s.size();
            if (k + 1 != keep.size()) { memset(&Seq[off], '-', 32); off += 32; }
            std::vector<uint8_t>().swap(s);
        }
    }

    template <class F> void ForEachPlusWord(F f) const {  // rolling word of ufindex.cpp:105-146
        uint64_t Word = 0;
        uint32_t K = 0;
        for (uint64_t p = 0; p < Seq.size(); ++p) {
            uint8_t L = g_Letter[Seq[p]];
            if (L == 0xFF) { K = 0; Word = 0; continue; }
            if (K < W) ++K;
            Word = (Word << 2) | L;
            if (K == W) f(Slot(Word), (uint32_t)(p - (W - 1)));
        }
    }

    uint64_t FindEndOfList(uint64_t s) const {  // ufindex.cpp:945-985
        uint64_t s2 = s;
        for (;;) {
            uint8_t T = Tally(s2);
            uint32_t P = Pos(s2);
            if (T == T_PLUS1 || T == T_BOTH1 || T == T_END) return s2;
            if (T == T_LONG_MINE || T == T_LONG_OTHER) {
                uint64_t a = (s2 + (P & 0xffff)) % SlotCount;
                s2 = (a + (P >> 16)) % SlotCount;
            } else
                s2 = (s2 + (T & 127)) % SlotCount;
        }
    }
    unsigned FindFreeSlot(uint64_t s) const {  // ufindex.cpp:987-1000
        for (unsigned i = 1; i < 0xffff; ++i) {
            uint64_t s2 = (s + i) % SlotCount;
            uint8_t n = CntP[s2];
            if (n > 0 && n <= MaxIx) continue;
            if (Tally(s2) == T_FREE) return i;
        }
        return UINT32_MAX;
    }
    void TruncateSlot(uint64_t s) {  // ufindex.cpp:153-192
        ++Truncated;
        uint64_t s2 = s;
        for (;;) {
            uint8_t T = Tally(s2);
            uint32_t P = Pos(s2);
            SetTally(s2, T_FREE);
            SetPos(s2, UINT32_MAX);
            if (T == T_PLUS1 || T == T_BOTH1 || T == T_END) return;
            if (T == T_LONG_MINE || T == T_LONG_OTHER) {
                uint64_t a = (s2 + (P & 0xffff)) % SlotCount;
                s2 = (a + (P >> 16)) % SlotCount;
            } else
                s2 = (s2 + (T & 127)) % SlotCount;
        }
    }
    void UpdateSlot(uint64_t s, uint32_t pos) {  // ufindex.cpp:194-322
        uint8_t n = CntP[s], m = CntM[s];
        if (n > MaxIx || m > MaxIx) return;
        if (Tally(s) == T_FREE) {
            SetPos(s, pos);
            SetTally(s, (n == 1 && m == 0) ? T_BOTH1 : T_PLUS1);
            return;
        }
        uint64_t eol = FindEndOfList(s);
        unsigned step = FindFreeSlot(eol);
        if (step == UINT32_MAX) { TruncateSlot(s); return; }
        uint64_t fs = (eol + step) % SlotCount;
        if (step > 124) {
            unsigned step2 = FindFreeSlot(fs);
            if (step2 == UINT32_MAX) { TruncateSlot(s); return; }
            uint32_t eolpos = Pos(eol);
            uint64_t fs2 = (fs + step2) % SlotCount;
            SetNext(eol, eol == s ? (uint8_t)(T_LONG_MINE & 127) : T_LONG_OTHER);
            SetPos(eol, step | (step2 << 16));
            SetTally(fs, T_LONG_OTHER);
            SetPos(fs, eolpos);
            SetTally(fs2, T_END);
            SetPos(fs2, pos);
            return;
        }
        SetNext(eol, (uint8_t)step);
        SetTally(fs, T_END);
        SetPos(fs, pos);
    }

    void MakeIndex() {  // ufindex.cpp:83-151
        Blob.resize(5 * SlotCount);
        for (uint64_t s = 0; s < SlotCount; ++s) { SetTally(s, T_FREE); SetPos(s, UINT32_MAX); }
        CntP.assign(SlotCount, 0);
        CntM.assign(SlotCount, 0);
        ForEachPlusWord([&](uint64_t s, uint32_t) { if (CntP[s] < 255) ++CntP[s]; });
        {  // CountSlots_Minus, ufindex.cpp:373-408: walk backwards through the complement letters
            uint64_t Word = 0;
            uint32_t K = 0;
            for (uint64_t p = Seq.size(); p-- > 0;) {
                uint8_t L = g_CompLetter[Seq[p]];
                if (L == 0xFF) { K = 0; Word = 0; continue; }
                if (K < W) ++K;
                Word = (Word << 2) | L;
                if (K == W) { uint64_t s = Slot(Word); if (CntM[s] < 255) ++CntM[s]; }
            }
        }
        ForEachPlusWord([&](uint64_t s, uint32_t pos) { UpdateSlot(s, pos); });
    }

    void ToFile(const std::string &path) const {  // ufindexio.cpp:15-49
        FILE *f = fopen(path.c_str(), "wb");
        if (!f) Die("Cannot create %s", path.c_str());
        auto w32 = [&](uint32_t v) { fwrite(&v, 4, 1, f); };
        w32(0x55464931u); w32(W); w32(MaxIx); w32((uint32_t)Seq.size());
        fwrite(&SlotCount, 8, 1, f);
        w32((uint32_t)Labels.size());
        for (size_t i = 0; i < Labels.size(); ++i) {
            w32(Lengths[i]); w32(Offsets[i]); w32((uint32_t)Labels[i].size());
            fwrite(Labels[i].data(), 1, Labels[i].size(), f);
        }
        w32(0x55464932u);
        fwrite(Blob.data(), 1, Blob.size(), f);
        w32(0x55464933u);
        fwrite(Seq.data(), 1, Seq.size(), f);
        w32(0x55464935u);
        if (fclose(f) != 0) Die("Write error %s", path.c_str());
    }
};

extern "C" int urmb_host_gpu_build(const uint8_t *seq, uint64_t n, uint64_t slots, uint32_t W, uint32_t maxix, uint8_t *blob);

static int CmdMakeUfi(const Opts &o) {  // cmd_make_ufi, ufindexio.cpp:117-179
    UfiBuilder B;
    B.W = o.set_wordlength ? o.wordlength : 24;
    B.MaxIx = o.veryfast ? 3 : 32;
    if (o.set_maxix) B.MaxIx = o.maxix;
    FILE *f = fopen(o.make_ufi.c_str(), "rb");
    if (!f) Die("Cannot open %s", o.make_ufi.c_str());
    fseeko(f, 0, SEEK_END);
    int64_t GenomeSize = ftello(f);
    fclose(f);
    if (!o.slots.empty()) B.SlotCount = strtoull(o.slots.c_str(), nullptr, 10);
    else B.SlotCount = GetPrime((uint64_t)(int64_t)(GenomeSize / o.load_factor));
    if (GenomeSize > (int64_t)UINT32_MAX - 100000) Die("Genome too big (%lld)", (long long)GenomeSize);
    B.ShiftMask = B.W >= 32 ? ~0ull : ((1ull << (2 * B.W)) - 1);
    Progress("\n  Genome size  %lld\n        Slots  %llu\n  Load factor  %.2f\n  Word length  %u\n   Max abund.  %u\n\n",
             (long long)GenomeSize, (unsigned long long)B.SlotCount, GenomeSize / (double)B.SlotCount, B.W, B.MaxIx);
    double t0 = now_s();
    B.ReadFasta(o.make_ufi);
    Progress("Read %zu sequences, %zu bases (%.1f s)\n", B.Labels.size(), B.Seq.size(), now_s() - t0);
    t0 = now_s();
    if (o.gpu_build) {
        B.Blob.resize(5 * B.SlotCount);
        if (urmb_host_gpu_build(B.Seq.data(), B.Seq.size(), B.SlotCount, B.W, B.MaxIx, B.Blob.data()) != 0)
            Die("GPU index build failed: %s", urmb_build_last_error());
        Progress("Index built on the GPU (functionally equivalent layout) in %.1f s\n", now_s() - t0);
    } else {
        B.MakeIndex();
        Progress("Index built in %.1f s\n%u slots truncated\n", now_s() - t0, B.Truncated);
    }
    B.ToFile(o.output);
    return 0;
}

int main(int argc, char **argv) {
    InitAlpha();
    Opts o = ParseCmdLine(argc, argv);
    g_quiet = o.quiet;
    if (!o.log.empty()) g_log = fopen(o.log.c_str(), "w");
    if (o.version) { printf("urmap_b200 v%s (B200-native drop-in for urmap -map/-map2)\n", URMB_VERSION); return 0; }
    if (!o.make_ufi.empty()) return CmdMakeUfi(o);
    if (!o.map.empty()) return CmdMap(o, false);
    return CmdMap(o, true);
}
