// urmb_kernels.cu -- hand-written sm_100a kernels of the URMAP mapping hot path.
//
//   probe_kernel   : SetSlotsVec (state1.cpp:396) + GetBlob (ufindex.h:184) for every k-mer of every
//                    read on both strands: hash, Barrett-reduce, gather the 5-byte slot record from the
//                    HBM-resident UFI table.  Pure function of (read, index); HBM-gather bound.
//   search_kernel  : one warp per read (SE, Search_Lo search1m6.cpp:35) or per pair (PE, Search4/5
//                    search2m4.cpp:15 / search2m5.cpp:9).  The order-dependent state machine is replayed
//                    exactly; the primitives inside it are warp-cooperative:
//                      extend  (ExtendPen extendpen.cpp:9, ExtendScan extendscan.cpp:51): 32 bases per
//                              ballot into a mismatch bitmask, then replay over set bits only;
//                      row walk (GetRow_Blob ufindex.cpp:883): positions held one per lane;
//                      viterbi (State1::Viterbi viterbi.cpp:11 + TraceBackBitMem): 32-row blocks, lane =
//                              row, anti-diagonal wavefront over columns with shuffles, fp32 arithmetic
//                              identical to the reference, trace bits in shared memory (flank DP) or in
//                              a per-warp HBM scratch (mate-rescue DP).
// No tensor cores: nothing here is a dense contraction.
#include <cuda_runtime.h>
#include <stdint.h>

#include "urmb_internal.h"

namespace urmb {

#define FULL 0xffffffffu

// Kernel launch / dynamic shared memory spelled through macros so that tests/emu can compile this very
// file as plain C++ (a lock-step warp emulator used for debugging only; never part of the product).
#ifndef URMB_EMU
#define URMB_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#define URMB_LAUNCH(kern, grid, block, smem, stream, ...) \
    kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#endif

// ---- tally encoding (ufindex.h:23-36) ----
constexpr uint8_t T_FREE = 0, T_END = 127, T_MY_BIT = 128, T_PLUS1 = 254, T_BOTH1 = 255;
constexpr uint8_t T_LONG_MINE = 253, T_LONG_OTHER = 125, T_NEXT_MASK = 127;
constexpr uint32_t POS_INVALID_WORD = 0xFFFFFFFEu;  // probe output for k-mers containing a non-ACGTU letter
// ---- trace bits (tracebit.h:4-7) ----
constexpr uint8_t TB_DM = 1, TB_IM = 2, TB_MD = 4, TB_MI = 8;
// ---- constants (state1.h:12-19) ----
constexpr int SECONDARY_HIT_MAX_DELTA = 12;
constexpr uint32_t PRIME_STRIDE = 27;
constexpr uint32_t SCANK = 4;
constexpr int MAX_TL = 1000;
constexpr uint32_t BRN = 2;
#define NEG_INF (-9e9f)  // MINUS_INFINITY, mx.h:12

// g_CharToLetterNucleo semantics (alpha.cpp:1309): A/a=0 C/c=1 G/g=2 T/t/U/u=3 else 0xFF
__device__ __forceinline__ uint32_t letter_of(uint32_t c) {
    uint32_t u = c & 0xDFu;
    uint32_t r = 0xFFu;
    if (u == 'A') r = 0;
    else if (u == 'C') r = 1;
    else if (u == 'G') r = 2;
    else if (u == 'T' || u == 'U') r = 3;
    return r;
}

// g_CharToCompChar semantics (alpha.cpp:3005): IUPAC-aware, case-preserving, 'u' and unknown -> '?'
__device__ __forceinline__ uint32_t compchar_of(uint32_t c) {
    uint32_t u = c & 0xDFu;
    if (u < 'A' || u > 'Y' || (c & 0xC0u) != 0x40u) return '?';
    // index by letter A..Y
    const char *tbl = "TVGH??CD??M?KN???YSAABWXR";
    uint32_t o = (uint32_t)(unsigned char)tbl[u - 'A'];
    if (o == '?') return '?';
    if (c & 0x20u) {
        if (u == 'U') return '?';
        o |= 0x20u;
    }
    return o;
}

__device__ __forceinline__ uint64_t murmur64(uint64_t h) {  // ufindex.h:50
    h ^= (h >> 33);
    h *= 0xff51afd7ed558ccdULL;
    h ^= (h >> 33);
    h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= (h >> 33);
    return h;
}

// h % slot_count with a precomputed floor(2^64/p): q_est in {q-1, q} so one correction suffices.
__device__ __forceinline__ uint64_t mod_slots(uint64_t h, uint64_t p, uint64_t magic) {
    uint64_t q = __umul64hi(h, magic);
    uint64_t r = h - q * p;
    if (r >= p) r -= p;
    return r;
}

__device__ __forceinline__ uint64_t add_mod(uint64_t a, uint64_t b, uint64_t p) {
    uint64_t s = a + b;  // a < p < 2^63, b small
    if (s >= p) s %= p;
    return s;
}

// Index and genome bytes are touched once per candidate at random addresses: keep them out of L1 so that the
// per-warp search state (local/global scratch) stays resident there.
__device__ __forceinline__ uint32_t ldg_stream_u32(const uint32_t *p) {
#ifdef URMB_EMU
    return *p;
#else
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#endif
}
__device__ __forceinline__ uint32_t ldg_stream_u8(const uint8_t *p) {
#ifdef URMB_EMU
    return *p;
#else
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#endif
}

// 5-byte record at byte offset 5*slot: two aligned 32-bit loads (the table is padded).
template <bool STREAM = true>
__device__ __forceinline__ void load_blob(const uint8_t *blob, uint64_t slot, uint32_t &tally, uint32_t &pos) {
    uint64_t a = 5ull * slot;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(blob + (a & ~3ull));
    uint32_t w0 = STREAM ? ldg_stream_u32(w) : __ldg(w), w1 = STREAM ? ldg_stream_u32(w + 1) : __ldg(w + 1);
    uint32_t sh = (uint32_t)(a & 3ull) * 8u;   // tally at bit sh, pos at bits sh+8 .. sh+40 (<= 64)
    uint64_t v = (((uint64_t)w1 << 32) | w0) >> sh;
    tally = (uint32_t)v & 0xFFu;
    pos = (uint32_t)(v >> 8);
}

// =====================================================================================
// probe kernel
// =====================================================================================
// One warp per read. Letters of both strands are staged in shared memory (1 B/base), then
// lane = k-mer start: build the 2W-bit word, hash, reduce, gather.
__global__ void __launch_bounds__(256) probe_kernel(DevIndex ix, DevBatch b, DevProbe pr) {
    URMB_DYN_SMEM(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    uint8_t *sl = smem + (size_t)warp * 2 * b.seqcap;  // [2][seqcap] letters fwd / rc
    const uint32_t W = ix.word_len;
    for (uint32_t r = blockIdx.x * wpb + warp; r < b.n_reads; r += gridDim.x * wpb) {
        const uint32_t off = b.offs[r], L = b.offs[r + 1] - off;
        for (uint32_t i = lane; i < L; i += 32) {
            uint32_t c = b.seqs[off + i];
            sl[i] = (uint8_t)letter_of(c);
            sl[b.seqcap + (L - 1 - i)] = (uint8_t)letter_of(compchar_of(c));
        }
        __syncwarp();
        const uint32_t QWC = (L >= W) ? L - W + 1 : 0;
        const size_t base = (size_t)r * 2 * b.qcap;
        for (uint32_t s = 0; s < 2; ++s) {
            const uint8_t *let = sl + s * b.seqcap;
            for (uint32_t q = lane; q < b.qcap; q += 32) {
                uint32_t tally = T_FREE, pos = POS_INVALID_WORD;
                uint64_t slot = ~0ull;
                if (q < QWC) {
                    uint64_t word = 0;
                    uint32_t bad = 0;
                    for (uint32_t t = 0; t < W; ++t) {
                        uint32_t l = let[q + t];
                        bad |= l & 0x80u;
                        word = (word << 2) | (l & 3u);
                    }
                    if (!bad) {
                        slot = mod_slots(murmur64(word & ix.shift_mask), ix.slot_count, ix.magic);
                        load_blob<false>(ix.blob, slot, tally, pos);
                    }
                }
                pr.tally[base + s * b.qcap + q] = (uint8_t)tally;
                pr.pos[base + s * b.qcap + q] = pos;
                pr.slot[base + s * b.qcap + q] = slot;
            }
        }
        __syncwarp();
    }
}

// =====================================================================================
// search kernel: per-warp environment and per-mate state
// =====================================================================================
struct Env {
    DevIndex ix;
    DevParams P;
    WarpScratch *ws;
    uint8_t *s_tb;      // flank-DP trace bits (shared), rows x tb_stride + column-LB array
    uint8_t *s_win;     // flank-DP genome window (shared)
    uint32_t tb_stride;
    uint32_t tb_rows;
    int lane;
};

struct Mate {
    const uint8_t *q;       // shared: read bytes
    const uint8_t *rc;      // shared: reverse complement bytes
    const uint8_t *tally;   // shared: [2][qcap]
    const uint32_t *pos;    // shared: [2][qcap]
    const uint64_t *slots;  // global: [2][qcap]
    uint32_t *alive;        // shared: [2][8] bit q set = the BOTH1 candidate at (strand, q) survives the prefilter
    MateScratch *g;
    uint32_t QL, QWC, qcap;
    int HitCount, HSPCount, Top, MaxPenalty, Best, Second, BestHSP;
    uint32_t Mapq;
    int nPend[2];
    int overflow;
};

__device__ __forceinline__ const uint8_t *mate_seq(const Mate &m, bool Plus) { return Plus ? m.q : m.rc; }

// ---- run-length path helpers -----------------------------------------------------------
__device__ __forceinline__ void runs_append(uint16_t *runs, int &n, uint32_t op, uint32_t len, int cap, int &ovf,
                                            int lane) {
    // all lanes track n; lane 0 writes
    if (len == 0) return;
    if (n > 0) {
        uint16_t last = runs[n - 1];
        if ((last & 3u) == op && (last >> 2) + len <= 16383u) {
            __syncwarp();   // every lane has read `last` before lane 0 overwrites it
            if (lane == 0) runs[n - 1] = (uint16_t)(((uint32_t)(last >> 2) + len) << 2 | op);
            __syncwarp();
            return;
        }
    }
    if (n >= cap) { ovf = 1; return; }
    if (lane == 0) runs[n] = (uint16_t)((len << 2) | op);
    __syncwarp();
    ++n;
}

// ---- hits / HSPs -----------------------------------------------------------------------
__device__ bool overlaps_hit(const Env &E, const Mate &m, uint32_t DBStartPos) {  // state1.cpp:230 (strand ignored)
    const uint32_t key = DBStartPos >> 6;
    for (int base = 0; base < m.HitCount; base += 32) {
        int h = base + E.lane;
        bool f = (h < m.HitCount) && ((m.g->hit_pos[h] >> 6) == key);
        if (__any_sync(FULL, f)) return true;
    }
    return false;
}

__device__ int overlaps_hsp(const Env &E, const Mate &m, uint32_t StartPosQ, uint32_t StartPosDB) {  // state1.cpp:241
    const int64_t diag = (int64_t)StartPosDB - (int64_t)StartPosQ;
    for (int base = 0; base < m.HSPCount; base += 32) {
        int h = base + E.lane;
        bool f = (h < m.HSPCount) && ((int64_t)m.g->hsp_dbstart[h] - (int64_t)m.g->hsp_qstart[h] == diag);
        uint32_t bal = __ballot_sync(FULL, f);
        if (bal) return base + __ffs(bal) - 1;
    }
    return -1;
}

// State1::AddHitX, state1.cpp:508-551. runs == nullptr / nruns == 0 => empty path.
__device__ int add_hit(const Env &E, Mate &m, uint32_t StartPosDB, bool Plus, int Score, const uint16_t *runs,
                       int nruns) {
    if (Score < 10) return -1;
    if (overlaps_hit(E, m, StartPosDB)) return -1;
    int Pen = (int)m.QL - Score;
    int MaxPen = Pen - 2 * E.P.MM;
    if (MaxPen < m.MaxPenalty) m.MaxPenalty = MaxPen;
    int idx = m.HitCount;
    if (idx >= kHitCap) { m.overflow = 1; return -1; }
    if (nruns > kRunCap) { m.overflow = 1; nruns = kRunCap; }
    if (E.lane == 0) {
        m.g->hit_pos[idx] = StartPosDB;
        m.g->hit_score[idx] = (int16_t)Score;
        m.g->hit_plus[idx] = Plus ? 1 : 0;
        m.g->hit_nruns[idx] = (uint8_t)nruns;
    }
    for (int i = E.lane; i < nruns; i += 32) m.g->hit_runs[idx][i] = runs[i];
    __syncwarp();
    if (Score > m.Best) {
        m.Second = m.Best;
        m.Best = Score;
        m.Top = idx;
    } else if (Score == m.Best)
        m.Second = Score;
    else {
        if (Score < m.Best - SECONDARY_HIT_MAX_DELTA) return -1;
        if (Score > m.Second) m.Second = Score;
    }
    ++m.HitCount;
    return idx;
}

__device__ __forceinline__ void hsp_store(const Env &E, Mate &m, int k, uint32_t qs, uint32_t dbs, bool Plus,
                                          uint32_t len, int Score) {
    __syncwarp();   // callers read the old record before it is replaced
    if (E.lane == 0) {
        m.g->hsp_qstart[k] = (uint16_t)qs;
        m.g->hsp_dbstart[k] = dbs;
        m.g->hsp_len[k] = (uint16_t)len;
        m.g->hsp_score[k] = (int16_t)Score;
        m.g->hsp_flags[k] = Plus ? 1 : 0;  // aligned = false
    }
    __syncwarp();
}

// State1::AddHSPX, state1.cpp:553-591
__device__ void add_hsp(const Env &E, Mate &m, uint32_t qs, uint32_t dbs, bool Plus, uint32_t len, int Score) {
    if (Score < m.Best - 4) return;
    int k = overlaps_hsp(E, m, qs, dbs);
    if (k >= 0) {
        if (Score > (int)m.g->hsp_score[k]) hsp_store(E, m, k, qs, dbs, Plus, len, Score);
        return;
    }
    if (m.HSPCount >= kHspCap) { m.overflow = 1; return; }
    hsp_store(E, m, m.HSPCount, qs, dbs, Plus, len, Score);
    ++m.HSPCount;
    if (Score > m.BestHSP) m.BestHSP = Score;
}

// State1::AddHSPScan, extendscan.cpp:8-49
__device__ int add_hsp_scan(const Env &E, Mate &m, uint32_t qs, uint32_t dbs, bool Plus, uint32_t len, int Score) {
    int k = overlaps_hsp(E, m, qs, dbs);
    if (k >= 0) {
        if (Score > (int)m.g->hsp_score[k]) hsp_store(E, m, k, qs, dbs, Plus, len, Score);
        return k;
    }
    if (m.HSPCount >= kHspCap) { m.overflow = 1; return -1; }
    k = m.HSPCount++;
    hsp_store(E, m, k, qs, dbs, Plus, len, Score);
    if (Score > m.BestHSP) m.BestHSP = Score;
    return k;
}

// ---- gapless x-drop extension ----------------------------------------------------------
struct ExtOut {
    int Best, Start, End;
    bool fail;
};

// Mismatch bitmask: lane k keeps the word for read positions [32k, 32k+32).
__device__ __forceinline__ uint32_t mm_word(uint32_t mymask, int k) { return __shfl_sync(FULL, mymask, k); }

__device__ __forceinline__ int mm_next(uint32_t mymask, int p, int nw, int QL) {  // first mismatch >= p, or QL
    int k = p >> 5;
    uint32_t w = mm_word(mymask, k) & (0xFFFFFFFFu << (p & 31));
    while (w == 0 && ++k < nw) w = mm_word(mymask, k);
    return w ? (k << 5) + __ffs(w) - 1 : QL;
}
__device__ __forceinline__ int mm_prev(uint32_t mymask, int p) {  // last mismatch <= p, or -1
    int k = p >> 5;
    uint32_t w = mm_word(mymask, k) & (0xFFFFFFFFu >> (31 - (p & 31)));
    while (w == 0 && --k >= 0) w = mm_word(mymask, k);
    return w ? (k << 5) + 31 - __clz(w) : -1;
}

// Common body of ExtendPen (extendpen.cpp:21-79) and ExtendScan (extendscan.cpp:64-133).
__device__ ExtOut extend_core(const Env &E, const Mate &m, uint32_t SeedPosQ, uint32_t DBLo, bool Plus,
                              bool LeftCountsPen) {
    const uint8_t *Qs = mate_seq(m, Plus);
    const uint8_t *T = E.ix.seq + DBLo;
    const int QL = (int)m.QL, nw = (QL + 31) >> 5, W = (int)E.ix.word_len;
    uint32_t mymask = 0;
    for (int k = 0; k < nw; ++k) {
        int idx = (k << 5) + E.lane;
        bool mis = (idx < QL) && ((uint32_t)Qs[idx] != ldg_stream_u8(T + idx));
        uint32_t w = __ballot_sync(FULL, mis);
        if (E.lane == k) mymask = w;
    }
    const int MM = E.P.MM, XD = E.P.XDROP, MaxPen = m.MaxPenalty;
    ExtOut o;
    o.fail = false;
    int Pen = 0, Score = W, Best = 0;
    int End = (int)SeedPosQ + W - 1;
    int p = End + 1;
    while (p < QL) {
        int n = mm_next(mymask, p, nw, QL);
        int run = n - p;
        if (run > 0) {
            Score += run;
            if (Score > Best) { Best = Score; End = n - 1; }
        }
        if (n >= QL) break;
        Pen -= MM;
        if (Pen > MaxPen) { o.fail = true; break; }
        Score += MM;
        if (Best - Score > XD) break;
        p = n + 1;
    }
    int Start = (int)SeedPosQ;
    if (!o.fail) {
        p = Start - 1;
        while (p >= 0) {
            int n = mm_prev(mymask, p);
            int run = p - n;
            if (run > 0) {
                Score += run;
                if (Score > Best) { Best = Score; Start = n + 1; }
            }
            if (n < 0) break;
            if (LeftCountsPen) Pen -= MM;
            if (Pen > MaxPen) { o.fail = true; break; }
            Score += MM;
            if (Best - Score > XD) break;
            p = n - 1;
        }
    }
    o.Best = Best;
    o.Start = Start;
    o.End = End;
    return o;
}


// ---- state-independent candidate prefilter ---------------------------------------------------
// ExtendPen returns -1 WITHOUT touching any state whenever the unconstrained gapless extension (no penalty
// bound) is neither full-length nor reaches MIN_HSP_SCORE: the penalty bound and the hit-overlap test can only
// turn more calls into -1 (extendpen.cpp:11-17,45,71).  About 90 % of all candidates are hash-collision
// artefacts of the key-less table that die within a few bases of the seed, so each LANE tests one candidate
// against <= 16 bases per side; only candidates that are still alive go through the exact warp-wide path, in
// the reference's order.  32 independent genome gathers are in flight per warp instead of one.
__device__ __forceinline__ uint64_t load8_global(const uint8_t *p) {  // 8 bytes at an arbitrary address (padded buffer)
    const uint32_t *w = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3) * 8;
    uint32_t w0 = ldg_stream_u32(w), w1 = ldg_stream_u32(w + 1), w2 = ldg_stream_u32(w + 2);
    uint64_t lo = ((uint64_t)w1 << 32) | w0;
    return sh ? ((lo >> sh) | ((uint64_t)w2 << (64 - sh))) : lo;
}
__device__ __forceinline__ uint64_t load8_shared(const uint8_t *p) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3) * 8;
    uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    uint64_t lo = ((uint64_t)w1 << 32) | w0;
    return sh ? ((lo >> sh) | ((uint64_t)w2 << (64 - sh))) : lo;
}

constexpr int kPrefilterWin = 16;  // bases examined on each side of the seed

// Lane-local. true = must take the exact path; false = ExtendPen would certainly return -1 with no side effect.
#ifdef URMB_EMU
static unsigned long long g_pf_calls = 0, g_pf_alive = 0, g_ext_calls = 0, g_ext_core = 0;
#define EMU_COUNT(x) (++(x))
#else
#define EMU_COUNT(x)
#endif
__device__ bool prefilter_alive_impl(const Env &E, const Mate &m, uint32_t SeedPosQ, uint32_t SeedPosDB, bool Plus);
__device__ bool prefilter_alive(const Env &E, const Mate &m, uint32_t SeedPosQ, uint32_t SeedPosDB, bool Plus) {
    const bool a = prefilter_alive_impl(E, m, SeedPosQ, SeedPosDB, Plus);
    EMU_COUNT(g_pf_calls);
    if (a) EMU_COUNT(g_pf_alive);
    return a;
}
__device__ bool prefilter_alive_impl(const Env &E, const Mate &m, uint32_t SeedPosQ, uint32_t SeedPosDB, bool Plus) {
    if (SeedPosDB < SeedPosQ) return false;  // extendpen.cpp:11
    const uint32_t DBLo = SeedPosDB - SeedPosQ;
    const uint8_t *Qs = mate_seq(m, Plus);
    const uint8_t *T = E.ix.seq + DBLo;
    const int QL = (int)m.QL, W = (int)E.ix.word_len, MM = E.P.MM, XD = E.P.XDROP;
    int Score = W, Best = 0;
    int End = (int)SeedPosQ + W - 1, Start = (int)SeedPosQ;
    // all four 8-byte genome chunks (2 per side) are requested before any is consumed: one memory round trip
    const int pr0 = End + 1, pl0 = Start - 1;
    const int llo1 = max(0, pl0 - 7), llo2 = max(0, pl0 - 15);
    uint64_t xr[2], xl[2];
    {
        uint64_t tr0 = load8_global(T + pr0), tr1 = load8_global(T + pr0 + 8);
        uint64_t tl0 = load8_global(T + llo1), tl1 = load8_global(T + llo2);
        xr[0] = load8_shared(Qs + pr0) ^ tr0;
        xr[1] = load8_shared(Qs + pr0 + 8) ^ tr1;
        xl[0] = load8_shared(Qs + llo1) ^ tl0;
        xl[1] = load8_shared(Qs + llo2) ^ tl1;
    }
    {   // right scan, extendpen.cpp:29-52 without the penalty bound
        int p = pr0;
        const int lim = min(QL, p + kPrefilterWin);
        bool term = false;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint64_t x = xr[c];
            const int n = min(8, lim - p);
            for (int j = 0; j < n && !term; ++j, ++p) {
                if (((x >> (8 * j)) & 0xFFu) == 0) {
                    ++Score;
                    if (Score > Best) { Best = Score; End = p; }
                } else {
                    Score += MM;
                    if (Best - Score > XD) term = true;
                }
            }
        }
        if (!term && p < QL) return true;
    }
    {   // left scan, extendpen.cpp:55-78
        int p = pl0;
        bool term = false;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int lo = (c == 0) ? llo1 : llo2;
            const int hi = (c == 0) ? pl0 : pl0 - 8;     // highest position of this chunk
            const uint64_t x = xl[c];
            for (int q = hi; q >= lo && q >= 0 && !term; --q, --p) {
                if (((x >> (8 * (q - lo))) & 0xFFu) == 0) {
                    ++Score;
                    if (Score > Best) { Best = Score; Start = q; }
                } else {
                    Score += MM;
                    if (Best - Score > XD) term = true;
                }
            }
        }
        if (!term && p >= 0) return true;
    }
    if (Start == 0 && End == QL - 1) return true;
    const int MinHSPScore = (int)((double)(E.P.MIN_HSP_PCT * QL) / 100.0);
    return Best >= MinHSPScore;
}

// Prefilter every BOTH1 candidate of a mate (both strands) in parallel and publish the survivors as bitmasks.
__device__ void build_alive_table(const Env &E, Mate &m) {
    if (!(E.P.flags & 1u)) {   // switch off: every candidate takes the exact path
        for (uint32_t r = E.lane; r < 16; r += 32) m.alive[r] = 0xFFFFFFFFu;
        __syncwarp();
        return;
    }
    for (int s = 0; s < 2; ++s) {
        for (uint32_t q0 = 0; q0 < 256; q0 += 32) {
            const uint32_t q = q0 + E.lane;
            bool a = false;
            if (q < m.QWC && m.tally[s * m.qcap + q] == T_BOTH1) a = prefilter_alive(E, m, q, m.pos[s * m.qcap + q], s == 0);
            const uint32_t w = __ballot_sync(FULL, a);
            if (E.lane == 0) m.alive[s * 8 + (q0 >> 5)] = w;
            if (q0 + 32 >= m.QWC) {
                for (uint32_t r = (q0 >> 5) + 1 + E.lane; r < 8; r += 32) m.alive[s * 8 + r] = 0;
                break;
            }
        }
    }
    __syncwarp();
}

// State1::ExtendPen, extendpen.cpp:9-95. +score: full-length hit; -2: HSP saved; -1 otherwise.
__device__ __noinline__ int extend_pen(const Env &E, Mate &m, uint32_t SeedPosQ, uint32_t SeedPosDB, bool Plus) {
    if (E.lane == 0) EMU_COUNT(g_ext_calls);
    if (SeedPosDB < SeedPosQ) return -1;
    {   // BOTH1 candidate on its own strand that the prefilter proved dead: -1 with no side effect
        const uint32_t si = Plus ? 0u : 1u;
        if (SeedPosQ < m.QWC && m.tally[si * m.qcap + SeedPosQ] == T_BOTH1 && m.pos[si * m.qcap + SeedPosQ] == SeedPosDB &&
            !((m.alive[si * 8 + (SeedPosQ >> 5)] >> (SeedPosQ & 31)) & 1u))
            return -1;
    }
    const uint32_t DBLo = SeedPosDB - SeedPosQ;
    if (overlaps_hit(E, m, DBLo)) return -1;
    if (E.lane == 0) EMU_COUNT(g_ext_core);
    ExtOut o = extend_core(E, m, SeedPosQ, DBLo, Plus, true);
    if (o.fail) return -1;
    const int MinHSPScore = (int)((double)(E.P.MIN_HSP_PCT * (int)m.QL) / 100.0);
    if (o.Start == 0 && o.End == (int)m.QL - 1) {
        add_hit(E, m, DBLo, Plus, o.Best, nullptr, 0);
        return o.Best;
    }
    if (o.Best >= MinHSPScore) {
        add_hsp(E, m, (uint32_t)o.Start, DBLo + (uint32_t)o.Start, Plus, (uint32_t)(o.End - o.Start + 1), o.Best);
        return -2;
    }
    return -1;
}

// ---- banded Viterbi --------------------------------------------------------------------
// DiagBox::GetRange_j, diagbox.h:150-170
__device__ __forceinline__ void range_j(uint32_t LA, uint32_t LB, uint32_t dlo, uint32_t dhi, uint32_t i, uint32_t &Sj,
                                        uint32_t &Ej) {
    Sj = (dlo + i >= LA) ? dlo + i - LA : 0;
    if (Sj >= LB) Sj = LB - 1;
    Ej = (dhi + i + 1 >= LA) ? dhi + i + 1 - LA : 0;
    if (Ej > LB) Ej = LB;
}

// Trace-bit store. BIG: full (LA+1) x (LB+1) byte matrix in per-warp HBM scratch.
// !BIG: band-relative rows in shared memory: column index c = j - i + K, plus a separate column-LB array.
template <bool BIG>
struct TBStore {
    uint8_t *base;
    uint8_t *collb;   // !BIG only: column LB, one byte per row
    uint8_t *rowla;   // !BIG only: row LA (last-row insert chain), one byte per cell (written by many lanes)
    uint32_t stride;  // BIG: LB+1 bytes ; !BIG: cells per row (even), two cells per byte
    int K;            // !BIG: LA + 1 - dlo
    uint32_t LA, LB;
    __device__ __forceinline__ void put(uint32_t i, uint32_t j, uint8_t v) const {
        if (BIG) {
            base[(size_t)i * stride + j] = v;
        } else {
            if (j == LB) {
                if (i < LA) collb[i] = v;
                return;
            }
            if (i == LA) {
                int c = (int)j - (int)(LA - 1) + K;
                if (c >= 0 && c < (int)stride) rowla[c] = v;
                return;
            }
            int c = (int)j - (int)i + K;
            if (c >= 0 && c < (int)stride) {   // a row is only ever written by its own lane: plain read-modify-write
                uint8_t *b = base + i * (stride >> 1) + (c >> 1);
                const uint32_t sh = (c & 1) * 4;
                *b = (uint8_t)((*b & (0xF0u >> sh)) | ((v & 0xFu) << sh));
            }
        }
    }
    __device__ __forceinline__ uint8_t get(uint32_t i, uint32_t j) const {
        if (BIG) return base[(size_t)i * stride + j];
        if (j == LB) return (i < LA) ? collb[i] : 0;
        if (i == LA) {
            int c = (int)j - (int)(LA - 1) + K;
            return (c >= 0 && c < (int)stride) ? rowla[c] : 0;
        }
        int c = (int)j - (int)i + K;
        if (c < 0 || c >= (int)stride) return 0;
        return (base[i * (stride >> 1) + (c >> 1)] >> ((c & 1) * 4)) & 0xFu;
    }
};

// State1::Viterbi (viterbi.cpp:11-261) + TraceBackBitMem (tracebackbitmem.cpp:8-75).
// A: read part (shared). B: genome window (shared for the flank DP, global for the rescue DP).
// Returns the score; the path is left REVERSED as RLE runs in E.ws->runs_a (n_rev runs).
template <bool BIG>
__device__ __noinline__ float viterbi_warp(const Env &E, const uint8_t *A, uint32_t LA, const uint8_t *B, uint32_t LB,
                                           bool Left, bool Right, int &n_rev, int &ovf) {
    const int lane = E.lane;
    uint16_t *rev = E.ws->runs_a;
    n_rev = 0;
    const float GO = (float)E.P.GO, GE = (float)E.P.GE, MMs = (float)E.P.MM;
    if (LA == 0 || LB == 0) {  // viterbi.cpp:14-36 (never reached from AlignHSP/Scan; same unsigned arithmetic)
        if (LA == 0 && LB == 0) return 0.0f;
        if (LA == 0) {
            runs_append(rev, n_rev, 2, LB, kRunCap, ovf, lane);
            return (float)((uint32_t)E.P.GO + (LB - 1) * (uint32_t)E.P.GE);
        }
        runs_append(rev, n_rev, 1, LA, kRunCap, ovf, lane);
        return (float)((uint32_t)E.P.GO + (LA - 1) * (uint32_t)E.P.GE);
    }
    uint32_t dlo = min(LA, LB), dhi = max(LA, LB);
    if (dlo > E.P.R) dlo -= E.P.R; else dlo = 1;
    dhi += E.P.R;
    if (dhi > LA + LB - 1) dhi = LA + LB - 1;

    TBStore<BIG> tb;
    tb.LA = LA;
    tb.LB = LB;
    if (BIG) {
        tb.base = E.ws->tb;
        tb.collb = nullptr;
        tb.rowla = nullptr;
        tb.stride = LB + 1;
        tb.K = 0;
    } else {
        tb.base = E.s_tb;
        tb.stride = E.tb_stride;
        tb.collb = E.s_tb + (size_t)E.tb_rows * (E.tb_stride >> 1);
        tb.rowla = tb.collb + E.tb_rows;
        tb.K = (int)LA + 1 - (int)dlo;
    }
    float *rowM = E.ws->rowM, *rowD = E.ws->rowD;

    for (uint32_t i0 = 0; i0 < LA; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool rowact = i < LA;
        uint32_t Sj = 0, Ej = 0;
        if (rowact) range_j(LA, LB, dlo, dhi, i, Sj, Ej);
        uint32_t jbase, tmpE;
        range_j(LA, LB, dlo, dhi, i0, jbase, tmpE);
        const uint32_t ilast = min(i0 + 31, LA - 1);
        uint32_t lS, lE;
        range_j(LA, LB, dlo, dhi, ilast, lS, lE);
        const uint32_t lastcol = (lE == LB) ? LB : lE - 1;  // includes the virtual column LB
        const int nsteps = (int)(lastcol - jbase) + (int)(ilast - i0) + 1;
        uint32_t pS = 0, pE = 0;
        if (i0 > 0) range_j(LA, LB, dlo, dhi, i0 - 1, pS, pE);
        const uint32_t a = rowact ? A[i] : 0x100u;
        float outM = NEG_INF, outD = NEG_INF, Mdiag = NEG_INF, I0 = NEG_INF;
        const float openA = (Left && i == 0) ? 0.0f : GO, extA = (Left && i == 0) ? 0.0f : GE;
        const bool lastrow = rowact && (i == ilast);
        for (int s = 0; s < nsteps; ++s) {
            float upM = __shfl_up_sync(FULL, outM, 1), upD = __shfl_up_sync(FULL, outD, 1);
            const int js = (int)jbase + s - lane;
            const uint32_t j = (uint32_t)js;
            if (lane == 0) {
                if (i0 == 0) {
                    upM = NEG_INF;
                    upD = NEG_INF;
                } else {
                    bool inprev = (j >= pS && j < pE);
                    upM = inprev ? rowM[j] : NEG_INF;
                    upD = (inprev || (j == LB && pE == LB)) ? rowD[j] : NEG_INF;
                }
            }
            const bool incol = rowact && js >= (int)Sj && js < (int)Ej;
            const bool vcol = rowact && j == LB && Ej == LB && js >= 0;
            float myM = NEG_INF, myD = NEG_INF;
            if (incol) {
                const float M0 = (j == 0) ? ((i == 0) ? 0.0f : NEG_INF) : Mdiag;
                const uint32_t bch = BIG ? (uint32_t)__ldg(B + j) : (uint32_t)B[j];
                uint8_t bits = 0;
                float xM = M0;
                if (upD > xM) { xM = upD; bits = TB_DM; }
                if (I0 > xM) { xM = I0; bits = TB_IM; }
                myM = xM + ((a == bch) ? 1.0f : MMs);
                const bool freeB = (j == 0) && Left;
                const float md = M0 + (freeB ? 0.0f : GO);
                float d = upD + (freeB ? 0.0f : GE);
                if (md >= d) { d = md; bits |= TB_MD; }
                myD = d;
                const float mi = M0 + openA;
                I0 += extA;
                if (mi >= I0) { I0 = mi; bits |= TB_MI; }
                tb.put(i, j, bits);
                if (j == Sj && Sj > 0) tb.put(i, Sj - 1, TB_IM);
            } else if (vcol) {  // viterbi.cpp:187-200, end of Drow[]
                const float md = Mdiag + GO;
                float d = upD + GE;
                uint8_t t = 0;
                if (md >= d) { d = md; t = TB_MD; }
                myD = d;
                tb.put(i, LB, t);
            }
            Mdiag = upM;
            outM = myM;
            outD = myD;
            if (lastrow) {
                if (incol) { rowM[j] = myM; rowD[j] = myD; }
                else if (vcol) rowD[j] = myD;
            }
        }
        if (rowact && Ej < LB) tb.put(i, LB, TB_MD);  // -inf >= -inf in the reference (viterbi.cpp:194)
        __syncwarp();
    }

    // last row of DPI, viterbi.cpp:207-236
    uint32_t Sj, Ej;
    range_j(LA, LB, dlo, dhi, LA - 1, Sj, Ej);
    const float gop = Right ? 0.0f : GO, gex = Right ? 0.0f : GE;
    float I1 = NEG_INF;
    {
        // chunks of 32 columns: lane loads its Mrow[j-1], then a sequential max-plus chain via shuffles
        for (uint32_t c0 = Sj; c0 < Ej; c0 += 32) {
            uint32_t j = c0 + lane;
            float mprev = NEG_INF;
            if (j < Ej && j > Sj) mprev = rowM[j - 1];
            uint32_t n = min(32u, Ej - c0);
            uint8_t myt = 0;
            for (uint32_t t = 0; t < n; ++t) {
                float mp = __shfl_sync(FULL, mprev, t);
                float mi = mp + gop;
                I1 += gex;
                bool take = mi > I1;
                if (take) I1 = mi;
                if ((uint32_t)lane == t) myt = take ? TB_MI : 0;
            }
            if (j < Ej) tb.put(LA, j, myt);
        }
    }
    __syncwarp();
    float Score = rowM[LB - 1];
    int State = 0;  // 0 M, 1 D, 2 I
    const float FinalD = rowD[LB];
    if (FinalD > Score) { Score = FinalD; State = 1; }
    if (I1 > Score) { Score = I1; State = 2; }

    // traceback (uniform across lanes)
    uint32_t ti = LA, tj = LB;
    uint32_t curop = (uint32_t)State, curlen = 0;
    for (;;) {
        if (ti == 0 && tj == 0) break;
        if ((uint32_t)State == curop) ++curlen;
        else {
            runs_append(rev, n_rev, curop, curlen, kRunCap, ovf, lane);
            curop = (uint32_t)State;
            curlen = 1;
        }
        uint8_t t;
        if (State == 0) {
            if (ti == 0 || tj == 0) break;
            t = tb.get(ti - 1, tj - 1);
            State = (t & TB_DM) ? 1 : ((t & TB_IM) ? 2 : 0);
            --ti; --tj;
        } else if (State == 1) {
            if (ti == 0) break;
            t = tb.get(ti - 1, tj);
            State = (t & TB_MD) ? 0 : 1;
            --ti;
        } else {
            if (tj == 0) break;
            t = tb.get(ti, tj - 1);
            State = (t & TB_MI) ? 0 : 2;
            --tj;
        }
    }
    runs_append(rev, n_rev, curop, curlen, kRunCap, ovf, lane);
    return Score;
}

// Flank DP dispatch: the shared-memory trace store holds bands up to tb_stride-2 wide (always true for
// LB = LA + 2R (+1)); a window clipped by the end of the genome can be wider and then uses the HBM store.
__device__ float flank_viterbi(const Env &E, const uint8_t *A, uint32_t LA, uint32_t TLo, uint32_t LB, bool Left,
                               bool Right, int &n_rev, int &ovf) {
    uint32_t dlo = min(LA, LB), dhi = max(LA, LB);
    if (dlo > E.P.R) dlo -= E.P.R; else dlo = 1;
    dhi += E.P.R;
    if (LA + LB >= 1 && dhi > LA + LB - 1) dhi = LA + LB - 1;
    const bool fits = (LA + 1 <= E.tb_rows) && (dhi - dlo + 3 <= E.tb_stride) && (LA > 0) && (LB > 0);
    if (fits) return viterbi_warp<false>(E, A, LA, E.s_win, LB, Left, Right, n_rev, ovf);
    return viterbi_warp<true>(E, A, LA, E.ix.seq + TLo, LB, Left, Right, n_rev, ovf);
}

// State1::AlignHSP, alignhsp.cpp:60-172
__device__ __noinline__ int align_hsp(const Env &E, Mate &m, int HSPIndex) {
    const uint8_t fl = m.g->hsp_flags[HSPIndex];
    if (fl & 2) return -1;
    __syncwarp();   // all lanes have read the flags
    if (E.lane == 0) m.g->hsp_flags[HSPIndex] = fl | 2;
    __syncwarp();
    const uint32_t StartPosQ = m.g->hsp_qstart[HSPIndex], StartPosDB = m.g->hsp_dbstart[HSPIndex];
    const uint32_t HSPLength = m.g->hsp_len[HSPIndex];
    const int HSPScore = m.g->hsp_score[HSPIndex];
    const bool Plus = (fl & 1) != 0;
    int TotalPen = (int)HSPLength - HSPScore;
    int TotalScore = HSPScore;
    if (TotalPen > m.MaxPenalty) return -1;
    const uint32_t QL = m.QL, TL = E.ix.seq_size;
    uint32_t CombinedTLo = StartPosDB;
    const uint8_t *Qs = mate_seq(m, Plus);
    uint16_t *path = E.ws->runs_p;
    int np = 0, ovf = 0;
    const int pcap = 3 * kRunCap;
    if (StartPosQ > 0) {
        if (StartPosDB < StartPosQ) return -1;
        const uint32_t LeftQL = StartPosQ;
        const uint32_t LeftTHi = StartPosDB - 1;
        const uint32_t LeftTL = LeftQL + BRN * E.P.R;
        if (LeftTL >= LeftTHi) return -1;
        const uint32_t LeftTLo = LeftTHi - LeftTL + 1;
        bool dash = false;
        for (uint32_t k = E.lane; k < LeftTL; k += 32) {
            uint8_t c = __ldg(E.ix.seq + LeftTLo + k);
            E.s_win[k] = c;
            dash |= (c == '-');
        }
        if (__any_sync(FULL, dash)) return -1;
        __syncwarp();
        int nrev = 0;
        int LeftScore = (int)flank_viterbi(E, Qs, LeftQL, LeftTLo, LeftTL, true, false, nrev, ovf);
        // forward path = reverse(rev); TrimLeftIs (pathinfo.cpp:168): leading I's = last reversed run
        const uint16_t *rev = E.ws->runs_a;
        uint32_t LeftICount = 0;
        if (nrev > 0 && (rev[nrev - 1] & 3u) == 2u) {
            LeftICount = rev[nrev - 1] >> 2;
            --nrev;
        }
        for (int k = nrev - 1; k >= 0; --k) runs_append(path, np, rev[k] & 3u, rev[k] >> 2, pcap, ovf, E.lane);
        CombinedTLo = LeftTLo + LeftICount;
        const int AllGapScore = E.P.GO + ((int)LeftQL - 1) * E.P.GE;
        if (AllGapScore > LeftScore) LeftScore = AllGapScore;
        TotalScore += LeftScore;
        TotalPen += (int)LeftQL - LeftScore;
        if (TotalPen > m.MaxPenalty) return -1;
    }
    runs_append(path, np, 0, HSPLength, pcap, ovf, E.lane);
    const uint32_t RightQLo = StartPosQ + HSPLength;
    if (RightQLo < QL) {
        const uint32_t RightQL = QL - RightQLo;
        const uint32_t RightTLo = StartPosDB + HSPLength;
        uint32_t RightTHi = RightTLo + RightQL + BRN * E.P.R;
        if (RightTHi >= TL) RightTHi = TL - 1;
        if (RightTHi < RightTLo) return -1;
        const uint32_t RightTL = RightTHi - RightTLo + 1;
        bool dash = false;
        for (uint32_t k = E.lane; k < RightTL; k += 32) {
            uint8_t c = __ldg(E.ix.seq + RightTLo + k);
            E.s_win[k] = c;
            dash |= (c == '-');
        }
        if (__any_sync(FULL, dash)) return -1;
        __syncwarp();
        int nrev = 0;
        int RightScore = (int)flank_viterbi(E, Qs + RightQLo, RightQL, RightTLo, RightTL, false, true, nrev, ovf);
        const uint16_t *rev = E.ws->runs_a;
        // TrimRightIs (pathinfo.cpp:187): trailing I's = first reversed run; index 0 of the path is never removed
        int first = 0;
        uint32_t keepI = 0;
        if (nrev > 0 && (rev[0] & 3u) == 2u) {
            if (nrev == 1) keepI = 1;  // whole path is I's: one survives
            first = 1;
        }
        for (int k = nrev - 1; k >= first; --k) runs_append(path, np, rev[k] & 3u, rev[k] >> 2, pcap, ovf, E.lane);
        if (keepI) runs_append(path, np, 2, 1, pcap, ovf, E.lane);
        const int AllGapScore = E.P.GO + ((int)RightQL - 1) * E.P.GE;
        if (AllGapScore > RightScore) RightScore = AllGapScore;
        TotalScore += RightScore;
        TotalPen += (int)RightQL - RightScore;
        if (TotalPen > m.MaxPenalty) return -1;
    }
    if (ovf) m.overflow = 1;
    return add_hit(E, m, CombinedTLo, Plus, TotalScore, path, np);
}

// State1::ExtendScan, extendscan.cpp:51-187 (returns hit index or -1)
__device__ __noinline__ int extend_scan(const Env &E, Mate &m, uint32_t SeedPosQ, uint32_t SeedPosDB, bool Plus) {
    if (SeedPosDB < SeedPosQ) return -1;
    const uint32_t DBLo = SeedPosDB - SeedPosQ;
    ExtOut o = extend_core(E, m, SeedPosQ, DBLo, Plus, false);
    if (o.fail) return -1;
    const int MinHSPScore = (int)E.ix.word_len * 2;
    if (o.Start == 0 && o.End == (int)m.QL - 1) return add_hit(E, m, DBLo, Plus, o.Best, nullptr, 0);
    if (o.Best < MinHSPScore) return -1;
    int k = add_hsp_scan(E, m, (uint32_t)o.Start, DBLo + (uint32_t)o.Start, Plus, (uint32_t)(o.End - o.Start + 1), o.Best);
    if (k < 0) return -1;
    return align_hsp(E, m, k);
}

// State1::CalcMAPQ6, search1m6.cpp:9-33 (fp64, same operation order)
__device__ uint32_t calc_mapq6(const Mate &m) {
    if (m.HitCount == 0) return 0;
    if (m.Best <= 0) return 0;
    double BestPossible = (double)m.QL;
    double Second = (double)m.Second;
    if (Second < BestPossible / 2.0) {
        Second = BestPossible / 2.0;
        if ((double)m.Best <= Second) return 0;
    }
    double Fract = (double)m.Best / BestPossible;
    double Drop = (double)m.Best - Second;
    if (Drop > 40) Drop = 40;
    double v = __dmul_rn(__dmul_rn(Drop, Fract), Fract);
    uint32_t mapq = (uint32_t)v;
    if (mapq > 40) mapq = 40;
    return mapq;
}

// UFIndex::GetRow_Blob, ufindex.cpp:883-943. Positions are left one per lane in `mypos`.
__device__ __noinline__ uint32_t get_row(const Env &E, uint64_t Slot, uint32_t Tally, uint32_t Pos0, uint32_t &mypos) {
    mypos = 0;
    uint32_t T = Tally;
    if ((T & T_MY_BIT) == 0) return 0;
    uint64_t Slot2 = Slot;
    uint32_t Pos = Pos0;
    uint32_t K = 0;
    const uint64_t SC = E.ix.slot_count;
    for (;;) {
        if (K > 0) load_blob(E.ix.blob, Slot2, T, Pos);
        if ((uint32_t)E.lane == K) mypos = Pos;
        ++K;
        if (K == E.ix.max_ix) return K;
        if (T == T_PLUS1 || T == T_BOTH1) return 1;
        if (T == T_END) return K;
        if (T == T_LONG_MINE || T == T_LONG_OTHER) {
            uint32_t StepA = Pos & 0xffffu, StepB = Pos >> 16;
            uint64_t SlotA = add_mod(Slot2, StepA, SC);
            Slot2 = add_mod(SlotA, StepB, SC);
            uint32_t ta, pa;
            load_blob(E.ix.blob, SlotA, ta, pa);
            if ((uint32_t)E.lane == K - 1) mypos = pa;
        } else {
            Slot2 = add_mod(Slot2, T & T_NEXT_MASK, SC);
        }
    }
}

__device__ __forceinline__ uint32_t m_tally(const Mate &m, int strand /*0 plus,1 minus*/, uint32_t q) {
    return m.tally[strand * m.qcap + q];
}
__device__ __forceinline__ uint32_t m_pos(const Mate &m, int strand, uint32_t q) { return m.pos[strand * m.qcap + q]; }
__device__ __forceinline__ uint64_t m_slot(const Mate &m, int strand, uint32_t q) {
    return __ldg(m.slots + strand * m.qcap + q);
}

// Lane-local GetRow_Blob (ufindex.cpp:883-943) limited to what the "rows <= 2 now, longer rows later" logic
// needs: n = 0 (not mine), 1, 2, or 3 meaning "RowLength > 2"; p0/p1 = the first two positions.
__device__ void row_head3(const Env &E, uint64_t Slot, uint32_t Tally, uint32_t Pos0, uint32_t &n, uint32_t &p0,
                          uint32_t &p1) {
    n = 0; p0 = 0; p1 = 0;
    uint32_t T = Tally, Pos = Pos0, K = 0;
    if ((T & T_MY_BIT) == 0) return;
    uint64_t Slot2 = Slot;
    const uint64_t SC = E.ix.slot_count;
    for (;;) {
        if (K > 0) load_blob(E.ix.blob, Slot2, T, Pos);
        if (K == 0) p0 = Pos; else if (K == 1) p1 = Pos;
        ++K;
        if (K == E.ix.max_ix) { n = K; return; }
        if (T == T_PLUS1 || T == T_BOTH1) { n = 1; return; }
        if (T == T_END) { n = K; return; }
        if (K == 3) { n = 3; return; }
        if (T == T_LONG_MINE || T == T_LONG_OTHER) {
            uint32_t StepA = Pos & 0xffffu, StepB = Pos >> 16;
            uint64_t SlotA = add_mod(Slot2, StepA, SC);
            Slot2 = add_mod(SlotA, StepB, SC);
            uint32_t ta, pa;
            load_blob(E.ix.blob, SlotA, ta, pa);
            if (K == 1) p0 = pa; else if (K == 2) p1 = pa;
        } else {
            Slot2 = add_mod(Slot2, T & T_NEXT_MASK, SC);
        }
    }
}

// One item (QPos) per lane: walk the heads of 32 rows and prefilter their candidates in parallel, then visit the
// items in lane order exactly like the reference's loop body (search1m6.cpp:181-199, search1pepend.cpp:53-68):
// rows longer than 2 are deferred through `defer` (returns how many were deferred), the others are extended now.
template <class Defer>
__device__ void rows_short_stage(const Env &E, Mate &m, int s, bool valid, uint32_t QPos, Defer defer) {
    uint32_t n = 0, p0 = 0, p1 = 0;
    if (valid) row_head3(E, m_slot(m, s, QPos), m_tally(m, s, QPos), m_pos(m, s, QPos), n, p0, p1);
    if (n > E.ix.max_ix) n = E.ix.max_ix;
    const bool pf = (E.P.flags & 2u) != 0;
    const bool a0 = valid && n >= 1 && n <= 2 && (!pf || prefilter_alive(E, m, QPos, p0, s == 0));
    const bool a1 = valid && n == 2 && (!pf || prefilter_alive(E, m, QPos, p1, s == 0));
    uint32_t vmask = __ballot_sync(FULL, valid);
    const uint32_t m0 = __ballot_sync(FULL, a0), m1 = __ballot_sync(FULL, a1), big = __ballot_sync(FULL, n > 2);
    while (vmask) {
        const int b = __ffs(vmask) - 1;
        vmask &= vmask - 1;
        const uint32_t qb = __shfl_sync(FULL, QPos, b);
        if (big >> b & 1u) { defer(qb); continue; }
        if (m0 >> b & 1u) extend_pen(E, m, qb, __shfl_sync(FULL, p0, b), s == 0);
        if (m1 >> b & 1u) extend_pen(E, m, qb, __shfl_sync(FULL, p1, b), s == 0);
    }
}

// A deferred (long) row: full GetRow_Blob, one position per lane, prefilter in parallel, extend survivors in order.
__device__ void row_long_stage(const Env &E, Mate &m, int s, uint32_t QPos) {
    uint32_t mypos;
    const uint32_t RowLength = get_row(E, m_slot(m, s, QPos), m_tally(m, s, QPos), m_pos(m, s, QPos), mypos);
    const bool a = ((uint32_t)E.lane < RowLength) && (!(E.P.flags & 2u) || prefilter_alive(E, m, QPos, mypos, s == 0));
    uint32_t am = __ballot_sync(FULL, a);
    while (am) {
        const int r = __ffs(am) - 1;
        am &= am - 1;
        extend_pen(E, m, QPos, __shfl_sync(FULL, mypos, r), s == 0);
    }
}

__device__ void reset_search(const Env &E, Mate &m) {
    m.HitCount = 0;
    m.HSPCount = 0;
    m.Top = -1;
    m.Best = 0;
    m.Second = 0;
    m.BestHSP = 0;
    m.Mapq = 0xFFFFFFFFu;
    m.MaxPenalty = E.P.MAXPEN;
    m.nPend[0] = m.nPend[1] = 0;
}

// State1::Search_Lo, search1m6.cpp:35-277
__device__ void search_lo(const Env &E, Mate &m) {
    const uint32_t W = E.ix.word_len;
    const int QL = (int)m.QL;
    if (m.QL < W) { m.Mapq = 0; return; }   // reference underflows: report no hit
    const uint32_t QWC = m.QWC;
    m.MaxPenalty = E.P.MAXPEN;
    const int MinScorePhase1 = QL + E.P.XP1 * E.P.MM;
    const int MinScorePhase3 = QL + E.P.XP3 * E.P.MM;
    const int MinScorePhase4 = QL + E.P.XP4 * E.P.MM;
    const int TermHSPScorePhase3 = (QL * E.P.TERM3_PCT) / 100;
    m.BestHSP = 0;
    // phase 1: BOTH1 seeds at stride W
    for (uint32_t QPos = 0; QPos < QWC; QPos += W) {
        for (int s = 0; s < 2; ++s) {
            if (m_tally(m, s, QPos) != T_BOTH1) continue;
            int Score = extend_pen(E, m, QPos, m_pos(m, s, QPos), s == 0);
            if (Score >= MinScorePhase1) { m.Mapq = calc_mapq6(m); return; }
        }
    }
    // phase 2: remaining BOTH1 seeds. Lanes scan 32 positions at a time, then visit them in order.
    for (uint32_t q0 = 0; q0 < QWC; q0 += 32) {
        uint32_t q = q0 + E.lane;
        bool okq = (q < QWC) && (q % W != 0);
        uint32_t bp = __ballot_sync(FULL, okq && m_tally(m, 0, q) == T_BOTH1);
        uint32_t bm = __ballot_sync(FULL, okq && m_tally(m, 1, q) == T_BOTH1);
        uint32_t any = bp | bm;
        while (any) {
            int bit = __ffs(any) - 1;
            any &= any - 1;
            uint32_t QPos = q0 + bit;
            for (int s = 0; s < 2; ++s) {
                if (!(((s == 0) ? bp : bm) >> bit & 1u)) continue;
                int Score = extend_pen(E, m, QPos, m_pos(m, s, QPos), s == 0);
                if (Score >= MinScorePhase1) { m.Mapq = calc_mapq6(m); return; }
            }
        }
    }
    // phase 3
    if (m.BestHSP > TermHSPScorePhase3) {
        for (int i = 0; i < m.HSPCount; ++i) align_hsp(E, m, i);
        if (m.Best >= MinScorePhase1) { m.Mapq = calc_mapq6(m); return; }
    }
    // phase 4: non-BOTH1 owned slots; rows <= 2 now, longer rows deferred
    int nTodo[2] = {0, 0};
    for (int s = 0; s < 2; ++s) {
        int nt = 0;
        for (uint32_t q0 = 0; q0 < QWC; q0 += 32) {
            const uint32_t q = q0 + E.lane;
            const uint32_t T = (q < QWC) ? m_tally(m, s, q) : 0;
            const bool cand = (T != T_FREE && T != T_BOTH1 && (T & T_MY_BIT));
            rows_short_stage(E, m, s, cand, q, [&](uint32_t qd) {
                if (E.lane == 0) m.g->todo[s][nt] = (uint8_t)qd;
                ++nt;
            });
        }
        nTodo[s] = nt;
    }
    __syncwarp();
    if (m.Best >= MinScorePhase3) { m.Mapq = calc_mapq6(m); return; }
    // phase 5
    for (int s = 0; s < 2; ++s)
        for (int t = 0; t < nTodo[s]; ++t) row_long_stage(E, m, s, m.g->todo[s][t]);
    if (m.Best >= MinScorePhase4) { m.Mapq = calc_mapq6(m); return; }
    // phase 6
    for (int i = 0; i < m.HSPCount; ++i) align_hsp(E, m, i);
    m.Mapq = calc_mapq6(m);
}

// ---- paired-end ------------------------------------------------------------------------
struct SeedIt {   // iterator state of GetFirst/NextBoth1Seed for one mate
    uint32_t k;    // UINT_MAX when exhausted
    uint32_t QPos, DBPos;
    bool Plus;
};

__device__ __forceinline__ void pend_push(const Env &E, Mate &m, int s, uint32_t QPos) {
    if (E.lane == 0) m.g->pend[s][m.nPend[s]] = (uint8_t)QPos;
    ++m.nPend[s];
}

// Shared scan loop of GetFirstBoth1Seed (getseed.cpp:9-54) and GetNextBoth1Seed (getseed.cpp:87-137).
// `first` disables the same-diagonal filter.
__device__ uint32_t seed_scan(const Env &E, Mate &m, uint32_t kstart, bool first, SeedIt &it) {
    const uint32_t QWC = m.QWC;
    for (uint32_t k = kstart; k < QWC; ++k) {
        uint32_t QPos = (k * PRIME_STRIDE) % QWC;
        if (first) it.QPos = QPos;
        for (int s = 0; s < 2; ++s) {
            uint32_t T = m_tally(m, s, QPos);
            uint32_t P = m_pos(m, s, QPos);
            if (T == T_FREE && P == POS_INVALID_WORD) continue;  // Slot == UINT64_MAX
            if ((T & T_MY_BIT) == 0) continue;
            if (T != T_BOTH1) { pend_push(E, m, s, QPos); continue; }
            if (!first && (P - QPos == it.DBPos - it.QPos)) continue;
            it.DBPos = P;
            it.QPos = QPos;
            it.Plus = (s == 0);
            return k;
        }
    }
    return 0xFFFFFFFFu;
}

__device__ void seed_first(const Env &E, Mate &m, SeedIt &it) {
    it.QPos = 0; it.DBPos = 0; it.Plus = false;
    it.k = seed_scan(E, m, 0, true, it);
}

__device__ void seed_next(const Env &E, Mate &m, SeedIt &it) {  // getseed.cpp:56-138
    const uint32_t QWC = m.QWC;
    if (it.Plus) {  // minus strand at the same k; a non-BOTH1 owned slot is NOT pushed to pending here
        uint32_t QPos = (it.k * PRIME_STRIDE) % QWC;
        uint32_t T = m_tally(m, 1, QPos), P = m_pos(m, 1, QPos);
        if (!(T == T_FREE && P == POS_INVALID_WORD) && T == T_BOTH1) {
            if (P - QPos != it.DBPos - it.QPos) {
                it.DBPos = P;
                it.QPos = QPos;
                it.Plus = false;
                return;
            } else
                pend_push(E, m, 1, QPos);
        }
    }
    it.k = seed_scan(E, m, it.k + 1, false, it);
}

// State1::SearchPE_Pending, search1pepend.cpp:9-130 (k is always UINT_MAX at the call sites)
__device__ void search_pe_pending(const Env &E, Mate &m) {
    const int QL = (int)m.QL;
    m.MaxPenalty = E.P.MAXPEN;
    const int MinScorePhase1 = QL + E.P.XP1 * E.P.MM;
    const int TermHSPScorePhase3 = (QL * E.P.TERM3_PCT) / 100;
    if (m.Best >= MinScorePhase1) { m.Mapq = calc_mapq6(m); return; }
    if (m.BestHSP >= TermHSPScorePhase3) {
        for (int i = 0; i < m.HSPCount; ++i) align_hsp(E, m, i);
        if (m.Best >= MinScorePhase1) { m.Mapq = calc_mapq6(m); return; }
    }
    __syncwarp();
    int n2[2] = {0, 0};
    for (int s = 0; s < 2; ++s) {   // pending round 1: 32 list entries at a time
        int nd = 0;
        for (int i0 = 0; i0 < m.nPend[s]; i0 += 32) {
            const int i = i0 + E.lane;
            const bool valid = i < m.nPend[s];
            const uint32_t QPos = valid ? m.g->pend[s][i] : 0;
            __syncwarp();   // the list is compacted in place below (nd <= i0 for every write)
            rows_short_stage(E, m, s, valid, QPos, [&](uint32_t qd) {
                if (E.lane == 0) m.g->pend[s][nd] = (uint8_t)qd;
                ++nd;
            });
            __syncwarp();
        }
        n2[s] = nd;
    }
    for (int s = 0; s < 2; ++s)     // pending round 2
        for (int i = 0; i < n2[s]; ++i) row_long_stage(E, m, s, m.g->pend[s][i]);
    const int B = max(m.Best, m.BestHSP) - 8;
    for (int i = 0; i < m.HSPCount; ++i) {
        if ((int)m.g->hsp_score[i] < B) continue;
        align_hsp(E, m, i);
    }
    m.Mapq = calc_mapq6(m);
}

// State1::ScanSlots, scanslots.cpp:7-62.  Lanes hash 32 window positions at a time; matches against
// the mate's slots at QPos = 0,27,54,81 are then visited in window order.
__device__ void scan_slots(const Env &E, Mate &m, uint32_t DBLo, uint32_t DBSegLength, bool Plus) {
    const uint32_t W = E.ix.word_len;
    if (m.QL <= W * 4) return;
    const int s = Plus ? 0 : 1;
    uint64_t qslot[SCANK];
    uint32_t qpos[SCANK];
    for (uint32_t k = 0; k < SCANK; ++k) {
        qpos[k] = (k * PRIME_STRIDE) % m.QWC;
        qslot[k] = m_slot(m, s, qpos[k]);  // ~0 for invalid words: never equals a real slot
    }
    const uint8_t *T = E.ix.seq + DBLo;
    if (DBSegLength < W) return;
    const uint32_t nwords = DBSegLength - W + 1;
    for (uint32_t p0 = 0; p0 < nwords; p0 += 32) {
        uint32_t p = p0 + E.lane;   // window word start
        uint32_t hitmask = 0;
        if (p < nwords) {
            uint64_t word = 0;
            uint32_t bad = 0;
            for (uint32_t t = 0; t < W; ++t) {
                uint32_t l = letter_of(__ldg(T + p + t));
                bad |= l & 0x80u;
                word = (word << 2) | (l & 3u);
            }
            if (!bad) {
                uint64_t slot = mod_slots(murmur64(word & E.ix.shift_mask), E.ix.slot_count, E.ix.magic);
                for (uint32_t k = 0; k < SCANK; ++k)
                    if (slot == qslot[k]) hitmask |= 1u << k;
            }
        }
        uint32_t any = __ballot_sync(FULL, hitmask != 0);
        while (any) {
            int bit = __ffs(any) - 1;
            any &= any - 1;
            uint32_t hm = __shfl_sync(FULL, hitmask, bit);
            for (uint32_t k = 0; k < SCANK; ++k)
                if (hm >> k & 1u) extend_scan(E, m, qpos[k], DBLo + p0 + bit, Plus);
        }
    }
}

// State1::Scan, scan.cpp:14-39
__device__ __noinline__ void scan_mate(const Env &E, Mate &m, uint32_t DBPos, uint32_t DBSegLength, bool Plus,
                                       bool DoVit) {
    const int SavedMaxPenalty = m.MaxPenalty;
    const int SavedHitCount = m.HitCount;
    m.MaxPenalty = 130;
    scan_slots(E, m, DBPos, DBSegLength, Plus);
    m.MaxPenalty = SavedMaxPenalty;
    if (m.HitCount > SavedHitCount) return;
    if (!DoVit) return;
    int nrev = 0, ovf = 0;
    float Score = viterbi_warp<true>(E, mate_seq(m, Plus), m.QL, E.ix.seq + DBPos, DBSegLength, true, true, nrev, ovf);
    if ((double)Score >= (double)m.QL / 3.0) {
        const uint16_t *rev = E.ws->runs_a;
        uint16_t *path = E.ws->runs_p;
        int np = 0;
        uint32_t LeftICount = 0;
        int last = nrev - 1;
        if (nrev > 0 && (rev[last] & 3u) == 2u) { LeftICount = rev[last] >> 2; --last; }
        int first = 0;
        // TrimRightIs on the already left-trimmed path: index 0 is never removed
        if (last >= 0 && (rev[0] & 3u) == 2u) first = (last == 0) ? 0 : 1;
        uint32_t keep1 = (last == 0 && (rev[0] & 3u) == 2u) ? 1 : 0;
        if (keep1) runs_append(path, np, 2, 1, 3 * kRunCap, ovf, E.lane);
        else for (int k = last; k >= first; --k) runs_append(path, np, rev[k] & 3u, rev[k] >> 2, 3 * kRunCap, ovf, E.lane);
        if (ovf) m.overflow = 1;
        add_hit(E, m, DBPos + LeftICount, Plus, (int)Score, path, np);
    }
}

struct PairState {
    int BestPairScore, SecondBestPairScore;
    int BestF, BestR;   // hit indexes of the best pair (-1: none)
    int PairCount;
};

// State2::FindPairs, state2.cpp:20-85 (only what AdjustTopHitsAndMapqs consumes is kept)
__device__ void find_pairs(const Env &E, const Mate &F, const Mate &R, PairState &ps) {
    const int QL2 = (int)((F.QL + R.QL) / 2);
    ps.BestPairScore = -1;
    ps.SecondBestPairScore = -1;
    ps.BestF = ps.BestR = -1;
    ps.PairCount = 0;
    for (int hf = 0; hf < F.HitCount; ++hf) {
        const int ScoreF = F.g->hit_score[hf];
        if (ScoreF < F.Second - 12) continue;
        const int64_t PosF = F.g->hit_pos[hf];
        const int PlusF = F.g->hit_plus[hf];
        for (int base = 0; base < R.HitCount; base += 32) {
            int hr = base + E.lane;
            bool ok = false;
            int ScoreR = 0;
            if (hr < R.HitCount) {
                ScoreR = R.g->hit_score[hr];
                int64_t PosR = R.g->hit_pos[hr];
                int64_t d = PosF - PosR;
                if (d < 0) d = -d;
                ok = (ScoreR >= R.Second - 12) && (d + QL2 <= 1000) && ((int)R.g->hit_plus[hr] != PlusF);
            }
            uint32_t bal = __ballot_sync(FULL, ok);
            while (bal) {
                int bit = __ffs(bal) - 1;
                bal &= bal - 1;
                int Total = ScoreF + __shfl_sync(FULL, ScoreR, bit);
                if (Total > ps.BestPairScore) {
                    ps.SecondBestPairScore = ps.BestPairScore;
                    ps.BestPairScore = Total;
                    ps.BestF = hf;
                    ps.BestR = base + bit;
                } else if (Total == ps.BestPairScore)
                    ps.SecondBestPairScore = ps.BestPairScore;
                else if (Total > ps.SecondBestPairScore)
                    ps.SecondBestPairScore = Total;
                ++ps.PairCount;
            }
        }
    }
}

// State2::ScanPair, state2.cpp:87-137 (quirk 6: both window extensions use the forward mate's length)
__device__ void scan_pair(const Env &E, Mate &F, Mate &R) {
    const int HitCountF = F.HitCount, HitCountR = R.HitCount;
    const bool DoVitF = ((int)F.Mapq >= 10), DoVitR = ((int)R.Mapq >= 10);
    const uint32_t QLx = F.QL;
    for (int h = 0; h < HitCountF; ++h) {
        if ((int)F.g->hit_score[h] < F.Second) continue;
        uint32_t DBPos = F.g->hit_pos[h];
        if (F.g->hit_plus[h]) scan_mate(E, R, DBPos, kScanSeg, false, DoVitF);
        else if (DBPos >= (uint32_t)kScanSeg) scan_mate(E, R, DBPos - kScanSeg, kScanSeg + 2 * QLx, true, DoVitF);
    }
    for (int h = 0; h < HitCountR; ++h) {
        if ((int)R.g->hit_score[h] < R.Second) continue;
        uint32_t DBPos = R.g->hit_pos[h];
        if (R.g->hit_plus[h]) scan_mate(E, F, DBPos, kScanSeg, false, DoVitR);
        else if (DBPos >= (uint32_t)kScanSeg) scan_mate(E, F, DBPos - kScanSeg, kScanSeg + 2 * QLx, true, DoVitR);
    }
}

// State2::ExtendBoth1Pair4/5, search2m4.cpp:189-208, search2m5.cpp:134-156
__device__ bool extend_both1_pair(const Env &E, Mate &F, Mate &R, uint32_t QPosf, uint32_t DBPosf, bool Plusf,
                                  uint32_t QPosr, uint32_t DBPosr, int TermPairScore) {
    int FwdScore = extend_pen(E, F, QPosf, DBPosf, Plusf);
    if (FwdScore <= 0) return false;
    int RevScore = extend_pen(E, R, QPosr, DBPosr, !Plusf);
    if (RevScore <= 0) return false;
    if (FwdScore + RevScore < TermPairScore) return false;
    F.Mapq = 40;
    R.Mapq = 40;
    return true;
}

__device__ __forceinline__ void seed_record(const Env &E, Mate &m, int n, const SeedIt &it, int &ovf) {
    if (n >= kSeedCap) { ovf = 1; return; }
    if (E.lane == 0) {
        m.g->seed_db[n] = it.DBPos;
        m.g->seed_q[n] = (uint8_t)it.QPos;
        m.g->seed_plus[n] = it.Plus ? 1 : 0;
    }
    __syncwarp();
}

// State2::Search4 / Search5, search2m4.cpp:15-187, search2m5.cpp:9-132
__device__ void search_pair(const Env &E, Mate &F, Mate &R) {
    reset_search(E, F);
    reset_search(E, R);
    const uint32_t W = E.ix.word_len;
    if (F.QL < W || R.QL < W) { F.Mapq = R.Mapq = 0; return; }
    const int QLf = (int)F.QL, QLr = (int)R.QL, QL2 = (QLf + QLr) / 2;
    const int TermPairScore = QLf + QLr + 5 * E.P.MM;
    int NB1f = 0, NB1r = 0, ovf = 0;
    SeedIt itf, itr;
    seed_first(E, F, itf);
    seed_first(E, R, itr);
    do {
        if (itf.k != 0xFFFFFFFFu) {
            seed_record(E, F, NB1f, itf, ovf);
            if (NB1f < kSeedCap) ++NB1f;
            for (int base = 0; base < NB1r; base += 32) {
                int i = base + E.lane;
                uint32_t dbr = (i < NB1r) ? R.g->seed_db[i] : 0;
                int64_t d = (int64_t)itf.DBPos - (int64_t)dbr;
                if (d < 0) d = -d;
                uint32_t bal = __ballot_sync(FULL, (i < NB1r) && (d + QL2 <= MAX_TL));
                while (bal) {
                    int bit = __ffs(bal) - 1;
                    bal &= bal - 1;
                    int idx = base + bit;
                    if (extend_both1_pair(E, F, R, itf.QPos, itf.DBPos, itf.Plus, R.g->seed_q[idx], R.g->seed_db[idx],
                                          TermPairScore))
                        return;
                }
            }
        }
        if (itr.k != 0xFFFFFFFFu) {
            seed_record(E, R, NB1r, itr, ovf);
            if (NB1r < kSeedCap) ++NB1r;
            for (int base = 0; base < NB1f; base += 32) {
                int i = base + E.lane;
                uint32_t dbf = (i < NB1f) ? F.g->seed_db[i] : 0;
                int64_t d = (int64_t)dbf - (int64_t)itr.DBPos;
                if (d < 0) d = -d;
                uint32_t bal = __ballot_sync(FULL, (i < NB1f) && (d + QL2 <= MAX_TL));
                while (bal) {
                    int bit = __ffs(bal) - 1;
                    bal &= bal - 1;
                    int idx = base + bit;
                    if (extend_both1_pair(E, F, R, F.g->seed_q[idx], F.g->seed_db[idx], !itr.Plus, itr.QPos, itr.DBPos,
                                          TermPairScore))
                        return;
                }
            }
        }
        if (itf.k != 0xFFFFFFFFu) seed_next(E, F, itf);
        if (itr.k != 0xFFFFFFFFu) seed_next(E, R, itr);
    } while (itf.k != 0xFFFFFFFFu || itr.k != 0xFFFFFFFFu);
    if (ovf) F.overflow = 1;

    for (int i = 0; i < NB1f; ++i) extend_pen(E, F, F.g->seed_q[i], F.g->seed_db[i], F.g->seed_plus[i] != 0);
    for (int i = 0; i < NB1r; ++i) extend_pen(E, R, R.g->seed_q[i], R.g->seed_db[i], R.g->seed_plus[i] != 0);

    if (E.P.pe_method == 5) {
        search_pe_pending(E, F);
        search_pe_pending(E, R);
        return;
    }
    const int TermF = (QLf * 9) / 10, TermR = (QLr * 9) / 10;
    if (F.Best >= TermF && R.Best >= TermR) {
        int64_t d = (int64_t)F.g->hit_pos[F.Top] - (int64_t)R.g->hit_pos[R.Top];
        if (d < 0) d = -d;
        if (d + QL2 <= MAX_TL) {
            F.Mapq = 40;
            R.Mapq = 40;
            return;
        }
    }
    search_pe_pending(E, F);
    search_pe_pending(E, R);
    PairState ps;
    find_pairs(E, F, R, ps);
    if (ps.PairCount == 0) {
        scan_pair(E, F, R);
        find_pairs(E, F, R, ps);
    }
    // State2::AdjustTopHitsAndMapqs, search2.cpp:8-57
    if (ps.PairCount == 0) {
        F.Mapq /= 2;
        R.Mapq /= 2;
        return;
    }
    double Fract = (double)ps.BestPairScore / (double)(QLf + QLr);
    double Drop = (double)(ps.BestPairScore - ps.SecondBestPairScore);
    if (Drop > 30) Drop = 30;
    uint32_t mapq = (uint32_t)__dmul_rn(__dmul_rn(Drop, Fract), Fract);
    if (mapq > 40) mapq = 40;
    if (mapq > F.Mapq) F.Mapq = mapq;
    if (mapq > R.Mapq) R.Mapq = mapq;
    if (ps.BestF >= 0) {
        F.Top = ps.BestF;
        R.Top = ps.BestR;
    }
}

// ---- result write-back -----------------------------------------------------------------
__device__ void write_result(const Env &E, const Mate &m, const DevOut &o, uint32_t r) {
    urmb_result res;
    res.db_pos = 0xFFFFFFFFu;
    res.path_off = 0;
    res.path_runs = 0;
    res.score = 0;
    res.best = (int16_t)m.Best;
    res.second = (int16_t)m.Second;
    res.mapq = (uint8_t)min(m.Mapq, 255u);
    res.flags = m.overflow ? 0x80 : 0;
    res.hit_count = (uint8_t)min(m.HitCount, 255);
    res.hsp_count = (uint8_t)min(m.HSPCount, 255);
    if (m.Top >= 0) {
        const int t = m.Top;
        res.db_pos = m.g->hit_pos[t];
        res.score = m.g->hit_score[t];
        res.flags |= (m.g->hit_plus[t] ? 1 : 0) | 2;
        const uint32_t n = m.g->hit_nruns[t];
        if (n) {
            uint32_t off = 0;
            if (E.lane == 0) off = atomicAdd(&o.counters[0], n);
            off = __shfl_sync(FULL, off, 0);
            if (off + n <= o.runs_cap) {
                for (uint32_t k = E.lane; k < n; k += 32) o.runs[off + k] = m.g->hit_runs[t][k];
                res.path_off = off;
                res.path_runs = (uint16_t)n;
            } else {
                res.flags |= 0x80;
            }
        }
    }
    if (E.lane == 0) {
        o.res[r] = res;
        if (res.flags & 0x80) atomicAdd(&o.counters[1], 1u);
    }
}

// Load one read into shared memory (bytes, reverse complement, probe results).
__device__ void load_mate(const Env &E, Mate &m, const DevBatch &b, const DevProbe &pr, uint32_t r, uint8_t *s_q,
                          uint8_t *s_rc, uint8_t *s_tally, uint32_t *s_pos, uint32_t *s_alive, MateScratch *g) {
    const uint32_t off = b.offs[r], L = b.offs[r + 1] - off;
    for (uint32_t i = E.lane; i < L; i += 32) {
        uint32_t c = b.seqs[off + i];
        s_q[i] = (uint8_t)c;
        s_rc[L - 1 - i] = (uint8_t)compchar_of(c);  // RevCompSeq, seqinfo.cpp:9
    }
    const size_t base = (size_t)r * 2 * b.qcap;
    for (uint32_t i = E.lane; i < 2 * b.qcap; i += 32) {
        s_tally[i] = pr.tally[base + i];
        s_pos[i] = pr.pos[base + i];
    }
    __syncwarp();
    m.q = s_q;
    m.rc = s_rc;
    m.tally = s_tally;
    m.pos = s_pos;
    m.slots = pr.slot + base;
    m.g = g;
    m.QL = L;
    m.QWC = (L >= E.ix.word_len) ? L - E.ix.word_len + 1 : 0;
    m.qcap = b.qcap;
    m.overflow = 0;
    m.alive = s_alive;
    build_alive_table(E, m);
}

__device__ __forceinline__ void search_body(const DevIndex &ix, const DevParams &P, const DevBatch &b, const DevProbe &pr,
                                            const DevOut &o, WarpScratch *scratch, uint32_t smem_per_warp,
                                            uint32_t tb_stride, uint32_t tb_rows) {
    URMB_DYN_SMEM(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int gw = blockIdx.x * wpb + warp;
    uint8_t *sw = smem + (size_t)warp * smem_per_warp;
    // layout: pos[nm][2][qcap] u32 | q[nm][seqcap] | rc[nm][seqcap] | tally[nm][2][qcap] | win | tb
    const int nm = b.paired ? 2 : 1;
    uint32_t *s_pos = reinterpret_cast<uint32_t *>(sw);
    uint8_t *p8 = sw + (size_t)nm * 2 * b.qcap * 4;
    uint8_t *s_q = p8;
    p8 += (size_t)nm * b.seqcap;
    uint8_t *s_rc = p8;
    p8 += (size_t)nm * b.seqcap;
    uint8_t *s_tally = p8;
    p8 += (size_t)nm * 2 * b.qcap;
    uint32_t *s_alive = reinterpret_cast<uint32_t *>(p8);
    p8 += (size_t)nm * 16 * 4;
    Env E;
    E.ix = ix;
    E.P = P;
    E.ws = scratch + gw;
    E.s_win = p8;
    p8 += b.seqcap + 64;
    E.s_tb = p8;
    E.tb_stride = tb_stride;
    E.tb_rows = tb_rows;
    E.lane = lane;

    for (;;) {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(&o.counters[2], 1u);
        u = __shfl_sync(FULL, u, 0);
        if (u >= b.n_units) break;
        if (!b.paired) {
            Mate m;
            load_mate(E, m, b, pr, u, s_q, s_rc, s_tally, s_pos, s_alive, &E.ws->m[0]);
            reset_search(E, m);   // State1::Search, search1.cpp:7-24
            search_lo(E, m);
            write_result(E, m, o, u);
        } else {
            Mate F, R;
            load_mate(E, F, b, pr, u, s_q, s_rc, s_tally, s_pos, s_alive, &E.ws->m[0]);
            load_mate(E, R, b, pr, b.n_units + u, s_q + b.seqcap, s_rc + b.seqcap, s_tally + 2 * b.qcap,
                      s_pos + 2 * b.qcap, s_alive + 16, &E.ws->m[1]);
            search_pair(E, F, R);
            write_result(E, F, o, u);
            write_result(E, R, o, b.n_units + u);
        }
        __syncwarp();
    }
}

// Two register budgets of the same body: MINB = 4 blocks/SM (<=128 regs) or 3 blocks/SM (<=168 regs).
template <int MINB>
__global__ void __launch_bounds__(128, MINB) search_kernel(DevIndex ix, DevParams P, DevBatch b, DevProbe pr, DevOut o,
                                                       WarpScratch *scratch, uint32_t smem_per_warp, uint32_t tb_stride,
                                                       uint32_t tb_rows) {
    search_body(ix, P, b, pr, o, scratch, smem_per_warp, tb_stride, tb_rows);
}

// =====================================================================================
// host-side launchers
// =====================================================================================
static inline uint32_t tb_stride_for(const DevParams &P) { return 4 * P.R + 6; }

size_t search_smem_per_warp(const DevBatch &b, const DevParams &P) {
    const int nm = b.paired ? 2 : 1;
    size_t s = (size_t)nm * 2 * b.qcap * 4 + (size_t)nm * b.seqcap * 2 + (size_t)nm * 2 * b.qcap + (size_t)nm * 64;
    s += b.seqcap + 64;
    const uint32_t rows = b.seqcap + 2;
    s += (size_t)rows * (tb_stride_for(P) / 2) + rows + tb_stride_for(P);   // nibble rows + column LB + row LA
    return (s + 15) & ~(size_t)15;
}

int max_search_warps(int sm_count) { return sm_count * 32; }

int launch_probe(const DevIndex &ix, const DevBatch &b, const DevProbe &pr, void *stream, int sm_count) {
    const int threads = 256;
    const size_t smem = (size_t)(threads / 32) * 2 * b.seqcap;
    int blocks = (int)((b.n_reads + 7) / 8);
    if (blocks > sm_count * 16) blocks = sm_count * 16;
    if (blocks < 1) blocks = 1;
    URMB_LAUNCH(probe_kernel, blocks, threads, smem, stream, ix, b, pr);
    return (int)cudaGetLastError();
}

int launch_search(const DevIndex &ix, const DevParams &P, const DevBatch &b, const DevProbe &pr, const DevOut &o,
                  WarpScratch *scratch, int n_scratch_warps, void *stream, int sm_count, int *warps_used) {
    const int wpb = 4;
    const size_t spw = search_smem_per_warp(b, P);
    const size_t smem = spw * wpb;
    const bool lowreg = !(P.flags & 8u);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(search_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(search_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done = true;
    }
    int per_sm = 0;
    if (lowreg) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, search_kernel<4>, wpb * 32, smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, search_kernel<3>, wpb * 32, smem);
    if (per_sm < 1) per_sm = 1;
    int blocks = sm_count * per_sm;
    if (blocks * wpb > n_scratch_warps) blocks = n_scratch_warps / wpb;
    const int need = (int)((b.n_units + wpb - 1) / wpb);
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    if (warps_used) *warps_used = blocks * wpb;
    const uint32_t rows = b.seqcap + 2;
    if (lowreg) {
        URMB_LAUNCH(search_kernel<4>, blocks, wpb * 32, smem, stream, ix, P, b, pr, o, scratch, (uint32_t)spw,
                    tb_stride_for(P), rows);
    } else {
        URMB_LAUNCH(search_kernel<3>, blocks, wpb * 32, smem, stream, ix, P, b, pr, o, scratch, (uint32_t)spw,
                    tb_stride_for(P), rows);
    }
    return (int)cudaGetLastError();
}

}  // namespace urmb
