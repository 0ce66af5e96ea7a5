// urmb_kernels.cu -- hand-written sm_100a kernels of the URMAP mapping hot path.
//
//   probe_kernel   : SetSlotsVec (state1.cpp:396) + GetBlob (ufindex.h:184) for every k-mer of every read on both
//                    strands (hash, Barrett-reduce, gather the 5-byte slot record from the HBM-resident UFI table),
//                    then the state-independent half of ExtendPen for every BOTH1 candidate.  HBM-gather bound.
//   search kernels : one warp per read / pair / saved mate.  The reference's order-dependent state machine
//                    (SE Search_Lo search1m6.cpp:35; PE Search4/5 search2m4.cpp:15 / search2m5.cpp:9) is replayed
//                    exactly, but it is cut at its phase borders into SMALL kernels that hand the per-mate state
//                    over through a pool in HBM -- one 160 KB kernel spent half its cycles on instruction fetch:
//                      PE: pair_kernel (seed pairing, every pair) -> align_kernel_a -> rows_kernel -> align_kernel_c
//                          (SearchPE_Pending, search1pepend.cpp:9, per saved mate) -> finish_kernel (FindPairs,
//                          AdjustTopHitsAndMapqs) per chunk; rescue_kernel (ScanPair) once per batch;
//                      SE: seed_kernel_se (phases 1-2) -> align_kernel_se3 -> rows_kernel_se -> align_kernel_se6.
//                    The primitives inside are warp-cooperative:
//                      extend  (ExtendPen extendpen.cpp:9, ExtendScan extendscan.cpp:51): one candidate per lane on
//                              the 2-bit packed genome, XOR -> mismatch bitmask -> walk over set bits only;
//                      row walk (GetRow_Blob ufindex.cpp:883): one dependent-gather chain per lane;
//                      viterbi (State1::Viterbi viterbi.cpp:11 + TraceBackBitMem): flank DP with the band across
//                              the lanes (max-plus prefix scan for the insert chain), mate-rescue DP as 32-row
//                              blocks with an anti-diagonal wavefront; fp32 arithmetic identical to the reference.
// No tensor cores: nothing here is a dense contraction.
#include <cuda_runtime.h>
#include <stdint.h>

#include "urmb_internal.h"

namespace URMB_NS {

#define FULL 0xffffffffu
// lane of the calling thread: from the special register, not from Env (every out-of-line function would otherwise start
// with a load of Env::lane from local memory)
#define URMB_LANE ((int)(threadIdx.x & 31u))

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
    const uint32_t u = c & 0xDFu;
    const bool ok = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T') | (u == 'U');
    return ok ? (((u >> 1) ^ (u >> 2)) & 3u) : 0xFFu;   // A 0, C 1, G 2, T/U 3
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
    uint64_t s = a + b;  // a < p < 2^63, b < 2^17: one or two subtractions replace the 64-bit modulo
    while (s >= p) s -= p;
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
// genome packing (derived data, built once per index on the device)
// =====================================================================================
// One thread per 32 bases: 2-bit codes, first base in the most significant bits (so a k-mer is a funnel
// shift of two words and "next mismatch to the right" is a count-leading-zeros), plus an exception bit
// for every byte that is not exactly 'A','C','G','T' (N runs, IUPAC codes, the '-' contig padding, the zero
// padding after the last contig).  Windows that touch an exception bit are compared byte by byte instead.
constexpr int kCoarseShift = 10;   // one coarse exception bit per 1024 bases
__global__ void __launch_bounds__(256) pack_genome_kernel(const uint8_t *seq, size_t n_bytes, size_t n_words,
                                                          uint64_t *seq2, uint32_t *seqx, uint32_t *seqc) {
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (size_t)gridDim.x * blockDim.x) {
        uint64_t code = 0;
        uint32_t exc = 0;
        for (uint32_t t = 0; t < 32; ++t) {
            const size_t g = w * 32 + t;
            const uint32_t c = (g < n_bytes) ? seq[g] : 0u;
            uint32_t l = 0;
            if (c == 'A') l = 0;
            else if (c == 'C') l = 1;
            else if (c == 'G') l = 2;
            else if (c == 'T') l = 3;
            else exc |= 1u << t;
            code |= (uint64_t)l << (62 - 2 * t);
        }
        seq2[w] = code;
        seqx[w] = exc;
        if (exc) {
            const size_t cb = (w * 32) >> kCoarseShift;
            atomicOr(&seqc[cb >> 5], 1u << (cb & 31));
        }
    }
}

// =====================================================================================
// read staging (shared by both kernels)
// =====================================================================================
// Per strand, in shared memory: the ASCII bytes (exact byte compare, DP), the 2-bit packing in the genome's
// bit order (8 words + 1 pad for 256 bases), and one "letter is not ACGTU" bit per base (LSB first).
constexpr int kPkWords = kMaxLen / 32 + 1;   // u64 words per strand
constexpr int kBadWords = kMaxLen / 32 + 1;  // u32 words per strand
struct ReadView {
    const uint8_t *q, *rc;    // [seqcap] read bytes / RevCompSeq bytes (seqinfo.cpp:9)
    const uint64_t *pk;       // [2][kPkWords]
    const uint32_t *bad;      // [2][kBadWords]
    uint32_t QL;
    bool slow;                // some byte is not upper-case ACGT: every extension takes the byte path
    bool hasbad;              // some letter is invalid for hashing (g_CharToLetterNucleo == 0xFF)
};
constexpr size_t kReadViewBytes = 2 * kPkWords * 8 + 2 * kBadWords * 4;   // + 2*seqcap bytes

// Warp-cooperative. s_q/s_rc: seqcap bytes each; s_pk: 2*kPkWords u64; s_bad: 2*kBadWords u32.
__device__ __noinline__ void stage_read(int lane, const uint8_t *src, uint32_t L, uint32_t seqcap, uint8_t *s_q,
                                        uint8_t *s_rc, uint64_t *s_pk, uint32_t *s_bad, ReadView &rv) {
    bool notacgt = false;
#pragma unroll 1
    for (uint32_t i = lane; i < L; i += 32) {
        const uint32_t c = src[i];
        s_q[i] = (uint8_t)c;
        s_rc[L - 1 - i] = (uint8_t)compchar_of(c);
        notacgt |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
    }
    for (int i = lane; i < 2 * kPkWords; i += 32) s_pk[i] = 0;
    for (int i = lane; i < 2 * kBadWords; i += 32) s_bad[i] = 0;
    __syncwarp();
    // lane = group of 8 bases (kMaxLen / 8 == 32 groups), both strands
    uint32_t anybad = 0;
    const uint32_t g = (uint32_t)lane;
    if (8 * g < L) {
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            const uint8_t *bytes = s ? s_rc : s_q;
            uint32_t code = 0, bad = 0;
#pragma unroll 1
            for (uint32_t t = 0; t < 8; ++t) {
                const uint32_t i = 8 * g + t;
                uint32_t l = (i < L) ? letter_of(bytes[i]) : 0u;
                if (l & 0x80u) { bad |= 1u << t; l = 0; }
                code = (code << 2) | (l & 3u);
            }
            reinterpret_cast<uint16_t *>(s_pk + s * kPkWords)[4 * (g >> 2) + (3 - (g & 3))] = (uint16_t)code;
            reinterpret_cast<uint8_t *>(s_bad + s * kBadWords)[g] = (uint8_t)bad;
            anybad |= bad;
        }
    }
    rv.q = s_q;
    rv.rc = s_rc;
    rv.pk = s_pk;
    rv.bad = s_bad;
    rv.QL = L;
    rv.hasbad = __any_sync(FULL, anybad != 0);
    rv.slow = __any_sync(FULL, notacgt);
    __syncwarp();
}

// Slot of the k-mer starting at q on strand s (State1::SetSlotsVec, state1.cpp:396-438): ~0 when a letter is invalid.
__device__ __forceinline__ uint64_t slot_of(const DevIndex &ix, const ReadView &rv, int s, uint32_t q) {
    const uint32_t W = ix.word_len;
    if (rv.hasbad) {
        const uint32_t *bw = rv.bad + s * kBadWords + (q >> 5);
        const uint64_t win = (((uint64_t)bw[1] << 32) | bw[0]) >> (q & 31);
        if (win & ((W >= 32) ? 0xFFFFFFFFull : ((1ull << W) - 1))) return ~0ull;
    }
    const uint64_t *pw = rv.pk + s * kPkWords + (q >> 5);
    const uint32_t off = 2 * (q & 31);
    const uint64_t hi = off ? ((pw[0] << off) | (pw[1] >> (64 - off))) : pw[0];
    const uint64_t word = hi >> (64 - 2 * W);
    return mod_slots(murmur64(word & ix.shift_mask), ix.slot_count, ix.magic);
}

// Out-of-line versions for the search kernels, which are instruction-fetch sensitive (one shared copy instead of an
// inlined one per call site); the probe kernel keeps the inlined forms in its hot loop.
__device__ __noinline__ void load_blob_s(const uint8_t *blob, uint64_t slot, uint32_t &tally, uint32_t &pos) {
    load_blob<true>(blob, slot, tally, pos);
}
__device__ __noinline__ uint64_t slot_of_s(const DevIndex &ix, const ReadView &rv, int s, uint32_t q) {
    return slot_of(ix, rv, s, q);
}

// =====================================================================================
// state-independent gapless extension
// =====================================================================================
// ExtendPen (extendpen.cpp:9-95) and ExtendScan (extendscan.cpp:51-187) walk right then left from the seed and
// consult the search state in exactly one place: "Pen > m_MaxPenalty -> return -1".  Pen only grows (by
// -MISMATCH per visited mismatch), so the bounded walk fails iff the unbounded walk's final penalty exceeds the
// bound.  The unbounded walk is a pure function of (read strand, seed, genome window): ONE LANE computes it for
// one candidate -- 32 candidates per warp in flight -- and the order-dependent part of the reference (overlap
// test, penalty bound, hit / HSP bookkeeping) is replayed afterwards from the packed result in O(1).
//   packed: Best[0:9] Start[9:17] End[17:25] visited-mismatches[25:32] (saturating; 127 * 3 > any bound)
constexpr uint32_t EXT_NONE = 0xFFFFFFFFu;   // ExtendPen returns -1 without looking at the genome (or not a candidate)
__device__ __forceinline__ uint32_t ext_pack(int Best, int Start, int End, int nmis) {
    return (uint32_t)Best | ((uint32_t)Start << 9) | ((uint32_t)End << 17) | ((uint32_t)min(nmis, 127) << 25);
}
__device__ __forceinline__ int ext_best(uint32_t x) { return (int)(x & 511u); }
__device__ __forceinline__ int ext_start(uint32_t x) { return (int)((x >> 9) & 255u); }
__device__ __forceinline__ int ext_end(uint32_t x) { return (int)((x >> 17) & 255u); }
__device__ __forceinline__ int ext_nmis(uint32_t x) { return (int)(x >> 25); }

// Exact byte-by-byte walk (reads with lower-case / IUPAC letters, windows touching N, '-' or the end padding).
__device__ __noinline__ uint32_t pure_ext_bytes(const uint8_t *Qs, const uint8_t *T, int QL, int W, int MM, int XD,
                                                uint32_t SeedPosQ, bool LeftCountsPen) {
    int Score = W, Best = 0, nmis = 0;
    int End = (int)SeedPosQ + W - 1;
    for (int p = End + 1; p < QL; ++p) {
        if ((uint32_t)Qs[p] == ldg_stream_u8(T + p)) {
            if (++Score > Best) { Best = Score; End = p; }
        } else {
            ++nmis;
            Score += MM;
            if (Best - Score > XD) break;
        }
    }
    int Start = (int)SeedPosQ;
    for (int p = Start - 1; p >= 0; --p) {
        if ((uint32_t)Qs[p] == ldg_stream_u8(T + p)) {
            if (++Score > Best) { Best = Score; Start = p; }
        } else {
            if (LeftCountsPen) ++nmis;
            Score += MM;
            if (Best - Score > XD) break;
        }
    }
    return ext_pack(Best, Start, End, nmis);
}

// Lane-local. Plus selects the read strand; the candidate is (SeedPosQ, SeedPosDB) on diagonal DBLo.
// PenBound: any bound known to be >= m_MaxPenalty at the time the reference would make this call; the walk stops
// as soon as the penalty exceeds it (the packed result then only says "fails the bound", which is all that is used).
// The mismatch flags of a candidate window, 16 bases per 32-bit word: base t of half-word h (read position 16 h + t) at
// bit 30 - 2 t.  NH = compiled number of half-words (10 covers reads up to 160 bases with half the code of 16: the
// search kernels are instruction-fetch sensitive).  Returns false when the window touches a non-ACGT genome byte.
template <int NH, bool NZ>
__device__ __forceinline__ bool ext_flags(const DevIndex &ix, const DevParams &P, const ReadView &rv, bool Plus, uint32_t DBLo,
                                          int nw, int nh, int QL, uint32_t *mm, uint32_t &nz) {
    constexpr int NW = NH / 2 + 1;   // 64-bit genome words that can be touched
    const uint64_t *g = ix.seq2 + (DBLo >> 5);
    const uint32_t sh = 2 * (DBLo & 31);
    const uint32_t *rp = reinterpret_cast<const uint32_t *>(rv.pk + (Plus ? 0 : kPkWords));
    // coarse exception bits of the (at most two) 1024-base blocks under words [DBLo>>5, (DBLo>>5)+nw]
    const uint32_t cb0 = DBLo >> kCoarseShift, cb1 = (((DBLo >> 5) + (uint32_t)nw) << 5) >> kCoarseShift;
    uint32_t exc = 1u;
    if (!(P.flags & 32u)) exc = ((__ldg(ix.seqc + (cb0 >> 5)) >> (cb0 & 31)) | (__ldg(ix.seqc + (cb1 >> 5)) >> (cb1 & 31))) & 1u;
    // the genome window as a stream of 32-bit pieces in base order: S[2k] = high half of word k, S[2k+1] = low half
    uint32_t S[2 * NW + 1];
#pragma unroll
    for (int k = 0; k < NW; ++k) {   // all loads are issued before the first one is consumed
        uint64_t v = 0;
        if (k <= nw) v = __ldg(g + k);
        S[2 * k] = (uint32_t)(v >> 32);
        S[2 * k + 1] = (uint32_t)v;
    }
    S[2 * NW] = 0;
    if (exc) {   // rare: look at the fine bits
        const uint32_t *x = ix.seqx + (DBLo >> 5);
        exc = 0;
        for (int k = 0; k <= nw; ++k) exc |= __ldg(x + k);
        if (exc) return false;
    }
    const bool odd = sh >= 32;        // the window starts in the low half of word 0
    const uint32_t s = sh & 31u;
#pragma unroll
    for (int j = 0; j < NH; ++j) {
        uint32_t d = 0;
        if (j < nh) {
            const uint32_t lo = odd ? S[j + 1] : S[j], hi = odd ? S[j + 2] : S[j + 1];
            const uint32_t a = __funnelshift_l(hi, lo, s);   // (lo << s) | (hi >> (32 - s)), s in [0, 31]
            d = a ^ rp[j ^ 1];                               // packed read: 64-bit words, high half first
            d = (d | (d >> 1)) & 0x55555555u;
        }
        mm[j] = d;
    }
    if (QL & 15) mm[nh - 1] &= 0xFFFFFFFFu << (32 - 2 * (QL & 15));
    nz = 0;   // bit j: half-word j holds a mismatch (the NZ walks jump over clean half-words without touching mm)
    if (NZ) {
#pragma unroll
        for (int j = 0; j < NH; ++j) nz |= (uint32_t)(mm[j] != 0) << j;
    }
    return true;
}

// NZ: the walks skip runs of clean half-words through a bit mask.  It pays where most candidates are true ones (the
// probe kernel: -1.0 ms per 2 M reads) and costs where most are hash collisions that stop after a few mismatches (the
// row stages: +1.4 ms), so each kernel instantiates the form that suits it (profiles/r02n).
template <bool NZ>
__device__ __noinline__ uint32_t pure_ext_t(const DevIndex &ix, const DevParams &P, const ReadView &rv, bool Plus,
                                            uint32_t SeedPosQ, uint32_t SeedPosDB, bool LeftCountsPen, int PenBound) {
    if (SeedPosDB < SeedPosQ) return EXT_NONE;   // extendpen.cpp:11
    const uint32_t DBLo = SeedPosDB - SeedPosQ;
    const int QL = (int)rv.QL, W = (int)ix.word_len, MM = P.MM, XD = P.XDROP;
    const int nw = (QL + 31) >> 5;   // <= 8 packed 64-bit words
    const int nh = (QL + 15) >> 4;   // <= 16 half-words of 16 bases
    const int maxmis = min(PenBound / -MM, 126);   // nmis > maxmis  <=>  nmis * -MM > PenBound
    uint32_t mm[16], nz = 0;
    bool slow = rv.slow;
    if (!slow) slow = (nh <= 10) ? !ext_flags<10, NZ>(ix, P, rv, Plus, DBLo, nw, nh, QL, mm, nz) : !ext_flags<16, NZ>(ix, P, rv, Plus, DBLo, nw, nh, QL, mm, nz);
    if (slow) return pure_ext_bytes(Plus ? rv.q : rv.rc, ix.seq + DBLo, QL, W, MM, XD, SeedPosQ, LeftCountsPen);

    // Both walks are single loops whose iterations either step to the next half-word or consume one mismatch, so that
    // lanes with many mismatches (bogus candidates) and lanes that cross many clean words (true candidates) take a
    // similar number of trips.
    int Score = W, Best = 0, nmis = 0;
    int End = (int)SeedPosQ + W - 1;
    {   // right walk, extendpen.cpp:29-52: mismatches in increasing position = decreasing bit index
        int p = End + 1;
        int h = p >> 4;
        uint32_t w = 0;
        if (h < nh) w = mm[h] & (0xFFFFFFFFu >> (2 * (p & 15)));
        bool stop = false;
        for (;;) {
            if (w == 0) {   // next half-word (NZ: with a mismatch) to the right
                if (NZ) {
                    const uint32_t rem = (h < 31) ? nz & (0xFFFFFFFEu << h) : 0u;
                    if (!rem) break;
                    h = __ffs(rem) - 1;
                } else if (++h >= nh) break;
                w = mm[h];
                continue;
            }
            const int hb = 31 - __clz(w);
            w ^= 1u << hb;
            const int pos = (h << 4) + ((30 - hb) >> 1);
            const int run = pos - p;
            if (run > 0) {
                Score += run;
                if (Score > Best) { Best = Score; End = pos - 1; }
            }
            ++nmis;
            Score += MM;
            p = pos + 1;
            if (Best - Score > XD || nmis > maxmis) { stop = true; break; }
        }
        if (!stop) {
            const int run = QL - p;
            if (run > 0) {
                Score += run;
                if (Score > Best) { Best = Score; End = QL - 1; }
            }
        }
    }
    if (nmis > maxmis) return ext_pack(0, 1, 0, 127);
    int Start = (int)SeedPosQ;
    {   // left walk, extendpen.cpp:55-78
        int p = Start - 1;
        int h = p >> 4;   // -1 when the seed starts the read
        uint32_t w = 0;
        if (h >= 0) w = mm[h] & (0xFFFFFFFFu << (30 - 2 * (p & 15)));
        bool stop = false;
        for (;;) {
            if (w == 0) {   // next half-word (NZ: with a mismatch) to the left
                if (NZ) {
                    const uint32_t rem = (h > 0) ? nz & ((1u << h) - 1u) : 0u;
                    if (!rem) break;
                    h = 31 - __clz(rem);
                } else if (--h < 0) break;
                w = mm[h];
                continue;
            }
            const int lb = __ffs(w) - 1;
            w &= w - 1;
            const int pos = (h << 4) + ((30 - lb) >> 1);
            const int run = p - pos;
            if (run > 0) {
                Score += run;
                if (Score > Best) { Best = Score; Start = pos + 1; }
            }
            if (LeftCountsPen) ++nmis;
            Score += MM;
            p = pos - 1;
            if (Best - Score > XD || nmis > maxmis) { stop = true; break; }
        }
        if (!stop) {
            const int run = p + 1;
            if (run > 0) {
                Score += run;
                if (Score > Best) { Best = Score; Start = 0; }
            }
        }
    }
    if (nmis > maxmis) return ext_pack(0, 1, 0, 127);
    return ext_pack(Best, Start, End, nmis);
}

__device__ __forceinline__ uint32_t pure_ext(const DevIndex &ix, const DevParams &P, const ReadView &rv, bool Plus,
                                             uint32_t SeedPosQ, uint32_t SeedPosDB, bool LeftCountsPen, int PenBound) {
    return pure_ext_t<false>(ix, P, rv, Plus, SeedPosQ, SeedPosDB, LeftCountsPen, PenBound);
}

// =====================================================================================
// probe + extend kernel
// =====================================================================================
// One warp per read.  Pass 1, lane = k-mer start: hash (SetSlotsVec), gather the 5-byte slot record from the
// HBM-resident table (GetBlob); BOTH1 slots are appended to a per-warp candidate list in shared memory.  Pass 2,
// lane = candidate: gather the candidate's packed genome window and run the pure extension above, 32 candidates at a
// time so that every lane has work.  Every chain is slot sector -> genome sectors; thousands of chains per SM are in
// flight: HBM-gather bound.
__host__ __device__ inline size_t probe_smem_per_warp(uint32_t qcap, uint32_t seqcap) {
    //     read view + bytes                  c_pos                 c_qs
    return ((kReadViewBytes + 2 * (size_t)seqcap + 2 * (size_t)qcap * 4 + 2 * (size_t)qcap * 2) + 15) & ~(size_t)15;
}
#ifndef URMB_PROBE_BATCH
#define URMB_PROBE_BATCH 4   // with the first look compiled in, 8 rounds unrolled cost more in instruction fetch than they hide (profiles/r06j)
#endif
#ifndef URMB_PROBE_LB
#define URMB_PROBE_LB 3
#endif
// The complete probe of one staged read: every k-mer of both strands, then the pure extension of every BOTH1 candidate.
// vsrc = the read's staged view in shared memory (packed strands + invalid-letter bits), copied out for the search kernels.
__device__ __forceinline__ void probe_read(const DevIndex &ix, const DevParams &P, const DevBatch &b, const DevProbe &pr, uint32_t r,
                                           const ReadView &rv, const uint8_t *vsrc, const uint8_t *s_rc, uint32_t *c_pos,
                                           uint16_t *c_qs, int lane) {
    const uint32_t W = ix.word_len, L = rv.QL;
    const uint32_t lt = (1u << lane) - 1u;
    {   // the staged read is kept for the search kernels (they would otherwise redo stage_read up to four times)
        uint32_t *vw = reinterpret_cast<uint32_t *>(pr.view + (size_t)r * pr.view_stride);
        const uint32_t *src32 = reinterpret_cast<const uint32_t *>(vsrc);
        for (uint32_t i = lane; i < kReadViewBytes / 4; i += 32) vw[i] = src32[i];
        if (lane == 0) vw[kReadViewBytes / 4] = (rv.slow ? 1u : 0u) | (rv.hasbad ? 2u : 0u);
        const uint32_t *rc32 = reinterpret_cast<const uint32_t *>(s_rc);
        for (uint32_t i = lane; i < b.seqcap / 4; i += 32) vw[kViewHdr / 4 + i] = rc32[i];
    }
    const uint32_t QWC = (L >= W) ? L - W + 1 : 0;
    const size_t base = (size_t)r * 2 * b.qcap;
    uint32_t nc = 0;
    // Slot records of URMB_PROBE_BATCH rounds of 32 k-mers are gathered together (two loads per round and lane in
    // flight) before the first one is looked at: the candidate list needs a ballot per round, which would otherwise put one trip to
    // HBM between consecutive rounds.
    const uint32_t rps = b.qcap / 32, nrounds = 2 * rps;   // rounds of 32 k-mers: the plus strand, then the minus strand
    for (uint32_t r0 = 0; r0 < nrounds; r0 += URMB_PROBE_BATCH) {
        uint32_t w0[URMB_PROBE_BATCH], w1[URMB_PROBE_BATCH], shp = 0, okm = 0;
#pragma unroll
        for (int k = 0; k < URMB_PROBE_BATCH; ++k) {
            const uint32_t rr = r0 + k, s = rr >= rps, q = (rr - s * rps) * 32 + lane;
            w0[k] = w1[k] = 0;
            if (rr < nrounds && q < QWC) {
#ifdef URMB_PROBE_SLOT_CALL
                const uint64_t slot = slot_of_s(ix, rv, (int)s, q);
#else
                const uint64_t slot = slot_of(ix, rv, (int)s, q);
#endif
                if (slot != ~0ull) {
                    const uint64_t a = 5ull * slot;   // record at byte offset 5*slot: two aligned words cover it
                    const uint32_t *w = reinterpret_cast<const uint32_t *>(ix.blob + (a & ~3ull));
                    w0[k] = __ldg(w);
                    w1[k] = __ldg(w + 1);
                    shp |= (uint32_t)(a & 3ull) << (2 * k);
                    okm |= 1u << k;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < URMB_PROBE_BATCH; ++k) {
            const uint32_t rr = r0 + k, s = rr >= rps, q = (rr - s * rps) * 32 + lane;
            if (rr >= nrounds) break;
            uint32_t tally = T_FREE, pos = POS_INVALID_WORD;
            if ((okm >> k) & 1u) {
                const uint64_t v = (((uint64_t)w1[k] << 32) | w0[k]) >> (((shp >> (2 * k)) & 3u) * 8u);
                tally = (uint32_t)v & 0xFFu;
                pos = (uint32_t)(v >> 8);
            }
            const bool cand = tally == T_BOTH1;
            const uint32_t bal = __ballot_sync(FULL, cand);
            if (cand) {
                const uint32_t i = nc + __popc(bal & lt);
                c_pos[i] = pos;
                c_qs[i] = (uint16_t)(q | (s << 15));
            } else {
                pr.ext[base + s * b.qcap + q] = EXT_NONE;
            }
            nc += __popc(bal);
            pr.tally[base + s * b.qcap + q] = (uint8_t)tally;
            pr.pos[base + s * b.qcap + q] = pos;
        }
    }
    __syncwarp();
    for (uint32_t i = lane; i < nc; i += 32) {
        const uint32_t qs = c_qs[i], q = qs & 0x7FFFu, s = qs >> 15;
        pr.ext[base + s * b.qcap + q] = pure_ext_t<true>(ix, P, rv, s == 0, q, c_pos[i], true, P.MAXPEN);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(256, URMB_PROBE_LB) probe_kernel(DevIndex ix, DevParams P, DevBatch b, DevProbe pr) {
    URMB_DYN_SMEM(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    uint8_t *sw = smem + (size_t)warp * probe_smem_per_warp(b.qcap, b.seqcap);
    uint64_t *s_pk = reinterpret_cast<uint64_t *>(sw);
    uint32_t *s_bad = reinterpret_cast<uint32_t *>(sw + 2 * kPkWords * 8);
    uint32_t *c_pos = reinterpret_cast<uint32_t *>(sw + kReadViewBytes);
    uint16_t *c_qs = reinterpret_cast<uint16_t *>(c_pos + 2 * b.qcap);
    uint8_t *s_q = reinterpret_cast<uint8_t *>(c_qs + 2 * b.qcap), *s_rc = s_q + b.seqcap;
    for (uint32_t r = blockIdx.x * wpb + warp; r < b.n_reads; r += gridDim.x * wpb) {
        const uint32_t off = b.offs[r], L = b.offs[r + 1] - off;
        ReadView rv;
        stage_read(lane, b.seqs + off, L, b.seqcap, s_q, s_rc, s_pk, s_bad, rv);
        probe_read(ix, P, b, pr, r, rv, sw, s_rc, c_pos, c_qs, lane);
    }
}

// ---- first look (paired input) -----------------------------------------------------------
// 42 % of the pairs of a typical run leave State2::Search4/5 inside the seed loop (search2m4.cpp:79-142): the first BOTH1
// seed of each mate lies on the true diagonal, both gapless extensions span the whole read and the scores add up to
// QLf + QLr - 15 or more.  The reference has then touched a handful of slots, not 2 x 2 x 127.  The first look replays the
// seed loop over the first kLookK steps of both seed iterators (k < 8: both strands of both mates = 32 slot probes, one per
// lane) for as long as that is a pure function of those probes:
//   * the iterator of either mate returns, within the first 16 visits, exactly the BOTH1 visits whose diagonal differs
//     from the previous BOTH1 visit (getseed.cpp:9-138, as build_seeds_pe);
//   * inside the loop seeds are only extended by ExtendBoth1Pair4/5, i.e. for seed pairs within the template length, and on
//     untouched search states ExtendPen is the pure function pure_ext_t computes: a call that returns <= 0 without adding a
//     hit or an HSP changes nothing, so the replay may go on; the first call that does add something either completes the
//     exit (both full length, sum >= the bound: MAPQ 40/40, one hit per mate -- the result is written here) or ends the
//     replay;
//   * the replay also ends when it would need a seed beyond the probed visits, or an extension on the strand that is not
//     the seed's own (search2m4.cpp:122 passes !Plusr for the forward mate).
// A pair whose replay ends that way takes the complete path as before (the 32 records are L2 hits then), so the first look
// only has to be sound, never complete.
constexpr int kLookK = 8;   // seed-iterator steps of the first look: 2 mates x 2 strands x kLookK = one probe per lane
__device__ __forceinline__ int nth_set_bit(uint32_t m, int n) {
#pragma unroll 1
    for (int i = 0; i < n; ++i) m &= m - 1;
    return __ffs(m) - 1;
}
__device__ __forceinline__ bool ext_is_noop_p(const DevParams &P, uint32_t x, int QL, int MaxPenalty) {
    if (x == EXT_NONE) return true;
    if (ext_nmis(x) * -P.MM > MaxPenalty) return true;
    if (ext_start(x) == 0 && ext_end(x) == QL - 1) return false;
    const int MinHSPScore = (P.MIN_HSP_PCT * QL) / 100;   // == int(PCT*QL/100.0), extendpen.cpp:22
    return ext_best(x) < MinHSPScore;
}
__device__ __noinline__ bool pair_first_look(const DevIndex &ix, const DevParams &P, const ReadView &rvF, const ReadView &rvR,
                                             urmb_result *resF, urmb_result *resR) {
    const int lane = URMB_LANE;
    const uint32_t W = ix.word_len;
    if (rvF.QL < W || rvR.QL < W) return false;
    const int QLf = (int)rvF.QL, QLr = (int)rvR.QL, QL2 = (QLf + QLr) / 2;
    // lane = mate << 4 | visit, visit v = 2 k + strand as the iterators go (plus, then minus at every k)
    static_assert(4 * kLookK == 32, "one probe per lane");
    const int mate = lane >> 4, v = lane & 15, sgn = v & 1;
    const uint32_t k = (uint32_t)v >> 1;
    const ReadView &rv = mate ? rvR : rvF;
    const int QL = (int)rv.QL;
    const uint32_t QWC = rv.QL - W + 1;
    const bool valid = k < QWC;
    const uint32_t QPos = valid ? (k * PRIME_STRIDE) % QWC : 0u;
    uint32_t tally = T_FREE, pos = 0;
    if (valid) {   // out-of-line forms: the probe kernels are instruction-fetch sensitive (profiles/r05i)
        const uint64_t slot = slot_of_s(ix, rv, sgn, QPos);
        if (slot != ~0ull) load_blob_s(ix.blob, slot, tally, pos);
    }
    const bool b1 = tally == T_BOTH1;
    const uint32_t diag = pos - QPos;
    const uint32_t half = mate ? 0xFFFF0000u : 0x0000FFFFu, lt = (1u << lane) - 1u;
    const uint32_t b1mask = __ballot_sync(FULL, b1);
    const uint32_t below = b1mask & lt & half;
    const uint32_t pd = __shfl_sync(FULL, diag, below ? 31 - __clz(below) : 0);
    const bool ret = b1 && (!below || diag != pd);
    const uint32_t retmask = __ballot_sync(FULL, ret);
    const uint32_t retF = retmask & 0xFFFFu, retR = retmask >> 16;
    const int nF = __popc(retF), nR = __popc(retR);
    if (nF == 0 || nR == 0) return false;
    // seeds of the other mate within the template length of this one (bit j: visit j of the other mate)
    uint32_t partners = 0;
    if (QL2 > MAX_TL) return false;
    const uint32_t maxd = (uint32_t)(MAX_TL - QL2);   // |DBPosf - DBPosr| + QL2 <= MAX_TL
#pragma unroll 1
    for (uint32_t rm = retmask; rm; rm &= rm - 1) {   // every seed of either mate (a handful)
        const int src = __ffs(rm) - 1;
        const uint32_t op = __shfl_sync(FULL, pos, src);
        const uint32_t d = pos > op ? pos - op : op - pos;
        if (((src ^ lane) & 16) && d <= maxd) partners |= 1u << (src & 15);
    }
    if (!ret) partners = 0;
    if (!__any_sync(FULL, partners != 0)) return false;
    // ExtendPen of a seed on its own strand against an untouched search state: 0 = returns <= 0 and changes nothing,
    // 1 = full-length hit stored (returns its score), 2 = anything else (an HSP is saved; a hit below AddHitX's floor)
    int cls = 0, score = 0;
    if (partners) {
        const uint32_t x = pure_ext_t<true>(ix, P, rv, sgn == 0, QPos, pos, true, P.MAXPEN);
        if (!ext_is_noop_p(P, x, QL, P.MAXPEN)) {
            score = ext_best(x);
            cls = (ext_start(x) == 0 && ext_end(x) == QL - 1 && score >= 10) ? 1 : 2;
        }
    }
    const int Term = QLf + QLr + 5 * P.MM;
    int exF = -1, exR = -1;
#pragma unroll 1
    for (int t = 0; t < 16; ++t) {
        if (t >= nF) return false;   // the next seed lies beyond the probed visits
        const int fl = nth_set_bit(retF, t), rl = (t < nR) ? 16 + nth_set_bit(retR, t) : 32;
        {   // F seed t against the R seeds recorded so far (search2m4.cpp:81-109): ExtendPen(F seed) comes first in every call
            const uint32_t pm = __shfl_sync(FULL, partners, fl) & retR & (uint32_t)((1ull << (rl - 16)) - 1ull);
            if (pm) {
                const int cf = __shfl_sync(FULL, cls, fl);
                if (cf == 2) return false;
                if (cf == 1) {
                    const int r0 = 16 + __ffs(pm) - 1;
                    // the R seed is extended on the strand opposite to F's: only its own strand's extension is at hand
                    if (((r0 ^ fl) & 1) == 0) return false;
                    if (__shfl_sync(FULL, cls, r0) != 1) return false;
                    if (__shfl_sync(FULL, score, fl) + __shfl_sync(FULL, score, r0) < Term) return false;
                    exF = fl; exR = r0;
                    break;
                }
            }
        }
        if (t >= nR) return false;
        {   // R seed t against the F seeds recorded so far, F seed t included (search2m4.cpp:110-137): the F seed is extended
            // on the strand opposite to R's
            uint32_t pm = __shfl_sync(FULL, partners, rl) & retF & ((2u << fl) - 1u);
            while (pm) {
                const int f0 = __ffs(pm) - 1;
                pm &= pm - 1;
                if (((f0 ^ rl) & 1) == 0) return false;
                const int cf = __shfl_sync(FULL, cls, f0);
                if (cf == 0) continue;
                if (cf == 2) return false;
                if (__shfl_sync(FULL, cls, rl) != 1) return false;
                if (__shfl_sync(FULL, score, f0) + __shfl_sync(FULL, score, rl) < Term) return false;
                exF = f0; exR = rl;
                break;
            }
            if (exF >= 0) break;
        }
    }
    if (exF < 0) return false;
    if (lane == exF || lane == exR) {   // one stored hit, m_BestScore = its score, m_SecondBestScore 0, MAPQ 40 (search2m4.cpp:203-205)
        urmb_result res;
        res.db_pos = diag;
        res.path_off = 0;
        res.path_runs = 0;
        res.score = (int16_t)score;
        res.best = (int16_t)score;
        res.second = 0;
        res.mapq = 40;
        res.flags = (uint8_t)((sgn == 0 ? 1 : 0) | 2);
        res.hit_count = 1;
        res.hsp_count = 0;
        *(mate ? resR : resF) = res;
    }
    return true;
}

__host__ __device__ inline size_t probe_pair_smem_per_warp(uint32_t qcap, uint32_t seqcap) {
    //     two read views + bytes                        c_pos                 c_qs
    return (2 * (kReadViewBytes + 2 * (size_t)seqcap) + 2 * (size_t)qcap * 4 + 2 * (size_t)qcap * 2 + 15) & ~(size_t)15;
}
// Paired input: one warp per pair -- stage both mates, first look, and the complete probe of both mates when that does not
// finish the pair.
__global__ void __launch_bounds__(256, URMB_PROBE_LB) probe_pair_kernel(DevIndex ix, DevParams P, DevBatch b, DevProbe pr, urmb_result *res,
                                                                        uint32_t *counters) {
    URMB_DYN_SMEM(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    uint8_t *sw = smem + (size_t)warp * probe_pair_smem_per_warp(b.qcap, b.seqcap);
    const size_t msz = kReadViewBytes + 2 * (size_t)b.seqcap;   // per mate: packed strands, invalid-letter bits, bytes, reverse complement
    uint32_t *c_pos = reinterpret_cast<uint32_t *>(sw + 2 * msz);
    uint16_t *c_qs = reinterpret_cast<uint16_t *>(c_pos + 2 * b.qcap);
    uint32_t nlook = 0;
    for (uint32_t u = blockIdx.x * wpb + warp; u < b.n_units; u += gridDim.x * wpb) {
        ReadView rv[2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const uint32_t r = m ? b.n_units + u : u;
            const uint32_t off = b.offs[r], L = b.offs[r + 1] - off;
            uint8_t *a = sw + m * msz;
            stage_read(lane, b.seqs + off, L, b.seqcap, a + kReadViewBytes, a + kReadViewBytes + b.seqcap, reinterpret_cast<uint64_t *>(a),
                       reinterpret_cast<uint32_t *>(a + 2 * kPkWords * 8), rv[m]);
        }
        // bits 10 / 11 (measurements): the first look is computed but not used / not computed
        const bool fin = !(P.flags & 2048u) && pair_first_look(ix, P, rv[0], rv[1], res + u, res + b.n_units + u) && !(P.flags & 1024u);
        if (lane == 0) pr.done[u] = fin ? 1 : 0;
        if (fin) {
            ++nlook;
            __syncwarp();
            continue;
        }
#pragma unroll 1
        for (int m = 0; m < 2; ++m) {
            const uint8_t *a = sw + m * msz;
            // the view of this mate by value: rv[] was handed to functions by reference and lives in local memory, and the
            // probe's inner loops consult the view for every k-mer
            ReadView v;
            v.q = a + kReadViewBytes;
            v.rc = a + kReadViewBytes + b.seqcap;
            v.pk = reinterpret_cast<const uint64_t *>(a);
            v.bad = reinterpret_cast<const uint32_t *>(a + 2 * kPkWords * 8);
            v.QL = m ? rv[1].QL : rv[0].QL;
            v.slow = m ? rv[1].slow : rv[0].slow;
            v.hasbad = m ? rv[1].hasbad : rv[0].hasbad;
            probe_read(ix, P, b, pr, m ? b.n_units + u : u, v, a, v.rc, c_pos, c_qs, lane);
        }
    }
    if (lane == 0 && nlook) atomicAdd(&counters[CT_FIRST_LOOK], nlook);
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
    int lane_;          // (kept for the emulator harness; the kernels read the lane from threadIdx: no load from the stack)
};

struct Mate {
    ReadView rv;            // shared: bytes, packed strands, invalid-letter bits
    const uint8_t *q;       // = rv.q
    const uint8_t *rc;      // = rv.rc
    const uint8_t *tally;   // global (probe output): [2][qcap]
    const uint32_t *pos;    // global: [2][qcap]
    const uint32_t *ext;    // global: [2][qcap]
    // ordered BOTH1 candidate list of the current search (shared, 2*qcap entries): the seed sequence of
    // GetFirst/NextBoth1Seed (PE) or the phase-1 + phase-2 visit order of Search_Lo (SE)
    uint32_t *sd_db;        // DBPos
    uint32_t *sd_ext;       // packed pure extension on the seed's own strand
    uint16_t *sd_qs;        // QPos | strand << 15 (strand 0 = plus)
    uint32_t *sd_dead;      // bit i: ExtendPen(seed i, own strand) is known to return <= 0 without side effects
    int nSeeds;
    MateScratch *g;
    uint32_t QL, QWC, qcap;
    int HitCount, HSPCount, Top, MaxPenalty, Best, Second, BestHSP;
    int nRuns;              // used part of g->runs_pool
    uint32_t Mapq;
    int nPend[2];
    int overflow;
};

__device__ __forceinline__ const uint8_t *mate_seq(const Mate &m, bool Plus) { return Plus ? m.q : m.rc; }

// ---- run-length path helpers -----------------------------------------------------------
__device__ __noinline__ void runs_append(uint16_t *runs, int &n, uint32_t op, uint32_t len, int cap, int &ovf,
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
        int h = base + URMB_LANE;
        bool f = (h < m.HitCount) && ((m.g->hit_pos[h] >> 6) == key);
        if (__any_sync(FULL, f)) return true;
    }
    return false;
}

// Lane-local form of OverlapsHit for candidate prechecks: the hit list only grows, so a candidate whose bucket is
// already taken will still get -1 from ExtendPen whenever the reference reaches it (extendpen.cpp:15-17).
__device__ __forceinline__ bool overlaps_hit_lane(const Mate &m, uint32_t DBStartPos) {
    const uint32_t key = DBStartPos >> 6;
    for (int h = 0; h < m.HitCount; ++h)
        if ((m.g->hit_pos[h] >> 6) == key) return true;
    return false;
}

__device__ int overlaps_hsp(const Env &E, const Mate &m, uint32_t StartPosQ, uint32_t StartPosDB) {  // state1.cpp:241
    const int64_t diag = (int64_t)StartPosDB - (int64_t)StartPosQ;
    for (int base = 0; base < m.HSPCount; base += 32) {
        int h = base + URMB_LANE;
        bool f = (h < m.HSPCount) && ((int64_t)m.g->hsp_dbstart[h] - (int64_t)m.g->hsp_qstart[h] == diag);
        uint32_t bal = __ballot_sync(FULL, f);
        if (bal) return base + __ffs(bal) - 1;
    }
    return -1;
}

// State1::AddHitX, state1.cpp:508-551. runs == nullptr / nruns == 0 => empty path.
__device__ __noinline__ int add_hit(const Env &E, Mate &m, uint32_t StartPosDB, bool Plus, int Score, const uint16_t *runs,
                       int nruns) {
    if (Score < 10) return -1;
    if (overlaps_hit(E, m, StartPosDB)) return -1;
    int Pen = (int)m.QL - Score;
    int MaxPen = Pen - 2 * E.P.MM;
    if (MaxPen < m.MaxPenalty) m.MaxPenalty = MaxPen;
    int idx = m.HitCount;
    if (idx >= kHitCap) { m.overflow |= 1; return -1; }
    if (nruns > kRunCap) { m.overflow |= 2; nruns = kRunCap; }
    if (m.nRuns + nruns > kRunPool) { m.overflow |= 4; nruns = 0; }
    if (URMB_LANE == 0) {
        m.g->hit_pos[idx] = StartPosDB;
        m.g->hit_score[idx] = (int16_t)Score;
        m.g->hit_plus[idx] = Plus ? 1 : 0;
        m.g->hit_nruns[idx] = (uint8_t)nruns;
        m.g->hit_roff[idx] = (uint16_t)m.nRuns;
    }
    for (int i = URMB_LANE; i < nruns; i += 32) m.g->runs_pool[m.nRuns + i] = runs[i];
    __syncwarp();
    if (Score > m.Best) {
        m.Second = m.Best;
        m.Best = Score;
        m.Top = idx;
    } else if (Score == m.Best)
        m.Second = Score;
    else {
        if (Score < m.Best - SECONDARY_HIT_MAX_DELTA) return -1;   // the record stays uncommitted
        if (Score > m.Second) m.Second = Score;
    }
    ++m.HitCount;
    m.nRuns += nruns;
    return idx;
}

__device__ __forceinline__ void hsp_store(const Env &E, Mate &m, int k, uint32_t qs, uint32_t dbs, bool Plus,
                                          uint32_t len, int Score) {
    __syncwarp();   // callers read the old record before it is replaced
    if (URMB_LANE == 0) {
        m.g->hsp_qstart[k] = (uint16_t)qs;
        m.g->hsp_dbstart[k] = dbs;
        m.g->hsp_len[k] = (uint16_t)len;
        m.g->hsp_score[k] = (int16_t)Score;
        m.g->hsp_flags[k] = Plus ? 1 : 0;  // aligned = false
    }
    __syncwarp();
}

// State1::AddHSPX, state1.cpp:553-591
__device__ __noinline__ void add_hsp(const Env &E, Mate &m, uint32_t qs, uint32_t dbs, bool Plus, uint32_t len, int Score) {
    if (Score < m.Best - 4) return;
    int k = overlaps_hsp(E, m, qs, dbs);
    if (k >= 0) {
        if (Score > (int)m.g->hsp_score[k]) hsp_store(E, m, k, qs, dbs, Plus, len, Score);
        return;
    }
    if (m.HSPCount >= kHspCap) { m.overflow |= 8; return; }
    hsp_store(E, m, m.HSPCount, qs, dbs, Plus, len, Score);
    ++m.HSPCount;
    if (Score > m.BestHSP) m.BestHSP = Score;
}

// State1::AddHSPScan, extendscan.cpp:8-49
__device__ __noinline__ int add_hsp_scan(const Env &E, Mate &m, uint32_t qs, uint32_t dbs, bool Plus, uint32_t len, int Score) {
    int k = overlaps_hsp(E, m, qs, dbs);
    if (k >= 0) {
        if (Score > (int)m.g->hsp_score[k]) hsp_store(E, m, k, qs, dbs, Plus, len, Score);
        return k;
    }
    if (m.HSPCount >= kHspCap) { m.overflow |= 8; return -1; }
    k = m.HSPCount++;
    hsp_store(E, m, k, qs, dbs, Plus, len, Score);
    if (Score > m.BestHSP) m.BestHSP = Score;
    return k;
}

// ---- order-dependent half of ExtendPen ---------------------------------------------------
// extendpen.cpp:11-17,45,71,80-94 replayed from the packed pure result. +score: full-length hit; -2: HSP saved;
// -1 otherwise.  *stored = the hit went into the hit list (then every later call on this diagonal bucket is -1).
__device__ __noinline__ int extend_apply(const Env &E, Mate &m, uint32_t SeedPosQ, uint32_t SeedPosDB, bool Plus,
                                         uint32_t x, bool &stored) {
    stored = false;
    if (x == EXT_NONE) return -1;
    const uint32_t DBLo = SeedPosDB - SeedPosQ;
    if (overlaps_hit(E, m, DBLo)) return -1;
    if (ext_nmis(x) * -E.P.MM > m.MaxPenalty) return -1;
    const int Best = ext_best(x), Start = ext_start(x), End = ext_end(x);
    if (Start == 0 && End == (int)m.QL - 1) {
        stored = add_hit(E, m, DBLo, Plus, Best, nullptr, 0) >= 0;
        return Best;
    }
    const int MinHSPScore = (E.P.MIN_HSP_PCT * (int)m.QL) / 100;   // == int(PCT*QL/100.0), extendpen.cpp:22
    if (Best >= MinHSPScore) {
        add_hsp(E, m, (uint32_t)Start, DBLo + (uint32_t)Start, Plus, (uint32_t)(End - Start + 1), Best);
        return -2;
    }
    return -1;
}

// A call that can only return -1 and change nothing, judged from the pure result and a penalty bound that is
// >= the bound at the time of the call (m_MaxPenalty never grows while BOTH1 seeds / rows are being extended).
__device__ __forceinline__ bool ext_is_noop(const Env &E, uint32_t x, int QL, int MaxPenalty) {
    if (x == EXT_NONE) return true;
    if (ext_nmis(x) * -E.P.MM > MaxPenalty) return true;
    if (ext_start(x) == 0 && ext_end(x) == QL - 1) return false;
    const int MinHSPScore = (E.P.MIN_HSP_PCT * QL) / 100;   // == int(PCT*QL/100.0), extendpen.cpp:22
    return ext_best(x) < MinHSPScore;
}

// ExtendPen for a candidate that is not in a seed list (uniform arguments): every lane computes the same pure result.
__device__ int extend_pen(const Env &E, Mate &m, uint32_t SeedPosQ, uint32_t SeedPosDB, bool Plus) {
    const uint32_t x = pure_ext(E.ix, E.P, m.rv, Plus, SeedPosQ, SeedPosDB, true, m.MaxPenalty);
    bool stored;
    return extend_apply(E, m, SeedPosQ, SeedPosDB, Plus, x, stored);
}

// ---- seed lists ----------------------------------------------------------------------------
__device__ __forceinline__ bool seed_dead(const Mate &m, int i) { return (m.sd_dead[i >> 5] >> (i & 31)) & 1u; }

// After seed appends: fetch the packed pure extensions of the seeds from the probe rows (one gather round for 32
// seeds, instead of a dependent load inside every round of the list builders) and mark the seeds whose extension is
// a no-op from the start.
__device__ __noinline__ void seeds_init_dead(const Env &E, Mate &m, bool fetch_db) {
    __syncwarp();
#pragma unroll 1
    for (int base = 0; base < m.nSeeds; base += 32) {
        const int i = base + URMB_LANE;
        uint32_t x = EXT_NONE;
        if (i < m.nSeeds) {
            const uint32_t qs = m.sd_qs[i];
            const uint32_t row = (qs >> 15) * m.qcap + (qs & 0x7FFFu);
            x = __ldg(m.ext + row);
            if (fetch_db) m.sd_db[i] = __ldg(m.pos + row);   // the single-end builder only looked at the tallies
            m.sd_ext[i] = x;
        }
        const bool d = (i >= m.nSeeds) || ext_is_noop(E, x, (int)m.QL, m.MaxPenalty);
        const uint32_t w = __ballot_sync(FULL, d);
        if (URMB_LANE == 0) m.sd_dead[base >> 5] = w;
    }
    __syncwarp();
}

// A hit was stored at HitDBLo (and m_MaxPenalty possibly lowered): every seed in the same 64-base bucket now fails
// OverlapsHit (state1.cpp:230, strand ignored), every seed over the new penalty bound fails the bound.
__device__ __noinline__ void seeds_kill(const Env &E, Mate &m, uint32_t HitDBLo) {
    const uint32_t key = HitDBLo >> 6;
#pragma unroll 1
    for (int base = 0; base < m.nSeeds; base += 32) {
        const int i = base + URMB_LANE;
        bool d = false;
        if (i < m.nSeeds) {
            const uint32_t x = m.sd_ext[i];
            d = (((m.sd_db[i] - (m.sd_qs[i] & 0x7FFFu)) >> 6) == key) || (ext_nmis(x) * -E.P.MM > m.MaxPenalty);
        }
        const uint32_t w = __ballot_sync(FULL, d);
        if (URMB_LANE == 0 && w) m.sd_dead[base >> 5] |= w;
    }
    __syncwarp();
}

// ExtendPen(seed i) on the seed's own strand, through the memo.
__device__ __noinline__ int apply_seed(const Env &E, Mate &m, int i) {
    if (seed_dead(m, i)) return -1;
    const uint32_t qs = m.sd_qs[i], db = m.sd_db[i];
    bool stored;
    const int r = extend_apply(E, m, qs & 0x7FFFu, db, (qs >> 15) == 0, m.sd_ext[i], stored);
    __syncwarp();
    if (stored) seeds_kill(E, m, db - (qs & 0x7FFFu));
    else if (r <= 0) {
        if (URMB_LANE == 0) m.sd_dead[i >> 5] |= 1u << (i & 31);
        __syncwarp();
    }
    return r;
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
    const int lane = URMB_LANE;
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

// Flank DP (AlignHSP): State1::Viterbi (viterbi.cpp:11-261) + TraceBackBitMem with the BAND across the lanes.
// Band coordinate c = j - (dlo + i - LA) in [0, BW), BW = dhi - dlo + 1 <= 64; lane l owns c = 2l and 2l+1.  Going
// from row i-1 to row i the band slides one column to the right, so for cell (i, j) at coordinate c:
//     M(i-1, j-1)  is the previous row's value at the SAME coordinate (a register, no shuffle);
//     D(i-1, j)    is the previous row's value at coordinate c+1 (own register or one shuffle);
//     the I chain  I0 <- max(I0 + ext, M0 + open) only depends on the previous row, i.e. it is a max-plus prefix
//                  scan over c: with U_c = I_out(c) - c*ext and V_c = M0(c) + open - c*ext, U_c = max(U_{c-1}, V_c);
//                  the reference's ">= favours open" tie is V_c >= U_{c-1}.  All finite values are small integers
//                  and NEG + k == NEG in fp32, so the scan reproduces the sequential floats bit for bit.
// Rows are sequential (LA <= 126 steps of ~100 instructions) and no lane idles in a wavefront ramp.
// Trace bits: one byte per lane per row (two nibbles) in shared memory, TB[i][LB] and row LA kept separately,
// TB[i][Startj-1] = IM (viterbi.cpp:119) answered on the fly.
#ifdef URMB_BAND_F32
__device__ __noinline__ float viterbi_band(const Env &E, const uint8_t *A, uint32_t LA, const uint8_t *B, uint32_t LB,
                                           bool Left, bool Right, int &n_rev, int &ovf) {
    const int lane = URMB_LANE;
    uint16_t *rev = E.ws->runs_a;
    n_rev = 0;
    const float GO = (float)E.P.GO, GE = (float)E.P.GE, MMs = (float)E.P.MM;
    uint32_t dlo = min(LA, LB), dhi = max(LA, LB);
    if (dlo > E.P.R) dlo -= E.P.R; else dlo = 1;
    dhi += E.P.R;
    if (dhi > LA + LB - 1) dhi = LA + LB - 1;
    const int BW = (int)(dhi - dlo) + 1;
    uint8_t *tb = E.s_tb;
    uint8_t *collb = tb + (size_t)E.tb_rows * 32;
    uint8_t *rowla = collb + E.tb_rows;
    const int c0 = 2 * lane, c1 = c0 + 1;
    float pM0 = NEG_INF, pM1 = NEG_INF, pD0 = NEG_INF, pD1 = NEG_INF;   // previous row at my coordinates
    float DLB = NEG_INF;                                                 // Drow[LB]
#pragma unroll 1
    for (uint32_t i = 0; i < LA; ++i) {
        const int j0 = (int)dlo + (int)i - (int)LA;
        const int ja = j0 + c0, jb = ja + 1;
        const bool ina = c0 < BW && ja >= 0 && ja < (int)LB, inb = c1 < BW && jb >= 0 && jb < (int)LB;
        const uint32_t a = A[i];
        const bool free0 = Left && i == 0;
        const float openA = free0 ? 0.0f : GO, nextA = free0 ? 0.0f : -GE;   // nextA = -extA >= 0
        const float M0a = (ja == 0) ? ((i == 0) ? 0.0f : NEG_INF) : pM0;
        const float M0b = (jb == 0) ? ((i == 0) ? 0.0f : NEG_INF) : pM1;
        float upDb = __shfl_down_sync(FULL, pD0, 1);
        if (lane == 31) upDb = NEG_INF;
        const float upDa = pD1;
        // Drow[LB] (viterbi.cpp:187-200): M0 after the row loop is the previous row's M at column Endj-1
        {
            float M0e = NEG_INF;
            const int e = (int)LB - j0;   // previous-row coordinate of column Endj-1 when the band is clipped at LB
            if (e >= 0 && e < BW) {
                const float t0 = __shfl_sync(FULL, pM0, e >> 1), t1 = __shfl_sync(FULL, pM1, e >> 1);
                M0e = (e & 1) ? t1 : t0;
            }
            const float md = M0e + GO;
            DLB += GE;
            uint8_t t = 0;
            if (md >= DLB) { DLB = md; t = TB_MD; }
            if (lane == 0) collb[i] = t;
        }
        // I chain: prefix max of V over the band
        const float Va = ina ? (M0a + openA) + (float)c0 * nextA : NEG_INF;
        const float Vb = inb ? (M0b + openA) + (float)c1 * nextA : NEG_INF;
        float inc = fmaxf(Va, Vb);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float t = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc = fmaxf(inc, t);
        }
        float Ua = __shfl_up_sync(FULL, inc, 1);   // U_{c0-1}
        if (lane == 0) Ua = NEG_INF;
        const float Ub = fmaxf(Ua, Va);            // U_{c1-1}
        const float Ia = Ua - (float)(c0 - 1) * nextA, Ib = Ub - (float)(c1 - 1) * nextA;
        float Ma = NEG_INF, Mb = NEG_INF, Da = NEG_INF, Db = NEG_INF;
        uint32_t bits = 0;
        if (ina) {
            uint32_t t = 0;
            float xM = M0a;
            if (upDa > xM) { xM = upDa; t = TB_DM; }
            if (Ia > xM) { xM = Ia; t = TB_IM; }
            Ma = xM + ((a == (uint32_t)B[ja]) ? 1.0f : MMs);
            const bool freeB = (ja == 0) && Left;
            const float md = M0a + (freeB ? 0.0f : GO);
            float d = upDa + (freeB ? 0.0f : GE);
            if (md >= d) { d = md; t |= TB_MD; }
            Da = d;
            if (Va >= Ua) t |= TB_MI;
            bits = t;
        }
        if (inb) {
            uint32_t t = 0;
            float xM = M0b;
            if (upDb > xM) { xM = upDb; t = TB_DM; }
            if (Ib > xM) { xM = Ib; t = TB_IM; }
            Mb = xM + ((a == (uint32_t)B[jb]) ? 1.0f : MMs);
            const float md = M0b + GO;   // jb >= 1: never the free column
            float d = upDb + GE;
            if (md >= d) { d = md; t |= TB_MD; }
            Db = d;
            if (Vb >= Ub) t |= TB_MI;
            bits |= t << 4;
        }
        tb[i * 32 + lane] = (uint8_t)bits;
        pM0 = Ma; pM1 = Mb; pD0 = Da; pD1 = Db;
    }
    // last row of DPI, viterbi.cpp:207-236 (strict >): chain over M(LA-1, j-1), j in [Startj, LB)
    const int j0f = (int)dlo - 1;   // column of coordinate 0 in row LA-1
    float I1;
    {
        const float gop = Right ? 0.0f : GO, ngex = Right ? 0.0f : -GE;
        const int ja = j0f + c0, jb = ja + 1;
        const bool ina = c0 < BW && ja >= 0 && ja < (int)LB, inb = c1 < BW && jb >= 0 && jb < (int)LB;
        float Mla = __shfl_up_sync(FULL, pM1, 1);   // M at coordinate c0-1
        if (lane == 0) Mla = NEG_INF;
        const float Mlb = pM0;
        const float Va = ina ? (Mla + gop) + (float)c0 * ngex : NEG_INF;
        const float Vb = inb ? (Mlb + gop) + (float)c1 * ngex : NEG_INF;
        float inc = fmaxf(Va, Vb);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float t = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc = fmaxf(inc, t);
        }
        float Ua = __shfl_up_sync(FULL, inc, 1);
        if (lane == 0) Ua = NEG_INF;
        const float Ub = fmaxf(Ua, Va);
        rowla[c0] = (ina && Va > Ua) ? TB_MI : 0;
        rowla[c1] = (inb && Vb > Ub) ? TB_MI : 0;
        const float tot = __shfl_sync(FULL, inc, 31);
        I1 = tot - (float)((int)LB - 1 - j0f) * ngex;
    }
    __syncwarp();
    float Score;
    {
        const int cf = (int)LB - 1 - j0f;   // coordinate of column LB-1 in the last row
        const float t0 = __shfl_sync(FULL, pM0, (cf >> 1) & 31), t1 = __shfl_sync(FULL, pM1, (cf >> 1) & 31);
        Score = (cf >= 0 && cf < BW) ? ((cf & 1) ? t1 : t0) : NEG_INF;
    }
    int State = 0;  // 0 M, 1 D, 2 I
    if (DLB > Score) { Score = DLB; State = 1; }
    if (I1 > Score) { Score = I1; State = 2; }

    // traceback (uniform across lanes), tracebackbitmem.cpp:22-73
    auto get = [&](uint32_t i, uint32_t j) -> uint32_t {
        if (i == LA) {
            const int c = (int)j - j0f;
            return (c >= 0 && c < BW) ? rowla[c] : 0u;
        }
        if (j == LB) return collb[i];
        const int jz = (int)dlo + (int)i - (int)LA;
        const int c = (int)j - jz;
        if (c == -1) return (jz > 0) ? (uint32_t)TB_IM : 0u;   // TB[i][Startj-1] = IM, viterbi.cpp:119
        if (c < 0 || c >= BW) return 0u;
        return ((uint32_t)tb[i * 32 + (c >> 1)] >> ((c & 1) * 4)) & 0xFu;
    };
    uint32_t ti = LA, tj = LB;
    uint32_t curop = (uint32_t)State, curlen = 0;
#pragma unroll 1
    for (;;) {
        if (ti == 0 && tj == 0) break;
        if ((uint32_t)State == curop) ++curlen;
        else {
            runs_append(rev, n_rev, curop, curlen, kRunCap, ovf, lane);
            curop = (uint32_t)State;
            curlen = 1;
        }
        uint32_t t;
        if (State == 0) {
            if (ti == 0 || tj == 0) break;
            t = get(ti - 1, tj - 1);
            State = (t & TB_DM) ? 1 : ((t & TB_IM) ? 2 : 0);
            --ti; --tj;
        } else if (State == 1) {
            if (ti == 0) break;
            t = get(ti - 1, tj);
            State = (t & TB_MD) ? 0 : 1;
            --ti;
        } else {
            if (tj == 0) break;
            t = get(ti, tj - 1);
            State = (t & TB_MI) ? 0 : 2;
            --tj;
        }
    }
    runs_append(rev, n_rev, curop, curlen, kRunCap, ovf, lane);
    return Score;
}
#else
// Integer form (default).  Every finite value of the reference's fp32 DP is a small integer, and MINUS_INFINITY only ever
// loses a comparison against a finite value, so 32-bit integers with NEG = -2^28 give the same scores and -- on every cell
// a traceback can visit, where at least one operand of each comparison is finite -- the same trace bits (two "minus
// infinities" compare differently here than in fp32, where NEG + k == NEG, but such cells are unreachable).  Against the
// fp32 form: no special cases for column 0 (row 0 is seeded through the previous-row registers), the Drow[LB] column is
// only computed on the rows whose band reaches it, the scan needs no lane predicate, and the traceback walks whole runs
// (32 cells of a diagonal / column / row per step) instead of one cell at a time.
__device__ __noinline__ float viterbi_band(const Env &E, const uint8_t *A, uint32_t LA, const uint8_t *B, uint32_t LB,
                                           bool Left, bool Right, int &n_rev, int &ovf) {
    const int lane = URMB_LANE;
    uint16_t *rev = E.ws->runs_a;
    n_rev = 0;
    constexpr int NEG = -(1 << 28);
    const int GO = E.P.GO, GE = E.P.GE, MMs = E.P.MM;
    uint32_t dlo = min(LA, LB), dhi = max(LA, LB);
    if (dlo > E.P.R) dlo -= E.P.R; else dlo = 1;
    dhi += E.P.R;
    if (dhi > LA + LB - 1) dhi = LA + LB - 1;
    const int BW = (int)(dhi - dlo) + 1;
    uint8_t *tb = E.s_tb;
    uint8_t *collb = tb + (size_t)E.tb_rows * 32;
    uint8_t *rowla = collb + E.tb_rows;
    const int c0 = 2 * lane, c1 = c0 + 1;
    // previous row at my coordinates; M0 of cell (0, 0) is 0 (viterbi.cpp:106-116): column 0 of row 0 has coordinate LA - dlo
    const int cz = (int)LA - (int)dlo;
    int pM0 = (c0 == cz) ? 0 : NEG, pM1 = (c1 == cz) ? 0 : NEG, pD0 = NEG, pD1 = NEG;
    int DLB = NEG;                                                 // Drow[LB]
#pragma unroll 1
    for (uint32_t i = 0; i < LA; ++i) {
        const int j0 = (int)dlo + (int)i - (int)LA;
        const int ja = j0 + c0, jb = ja + 1;
        const bool ina = c0 < BW && ja >= 0 && ja < (int)LB, inb = c1 < BW && jb >= 0 && jb < (int)LB;
        const int a = (int)A[i];
        const bool free0 = Left && i == 0;
        const int openA = free0 ? 0 : GO, nextA = free0 ? 0 : -GE;   // nextA = -extA >= 0
        int upDb = __shfl_down_sync(FULL, pD0, 1);
        if (lane == 31) upDb = NEG;
        const int upDa = pD1;
        // Drow[LB] (viterbi.cpp:187-200): M0 after the row loop is the previous row's M at column Endj-1; before the band
        // reaches column LB both operands are minus infinity and the reference's ">=" sets the bit
        {
            const int e = (int)LB - j0;   // previous-row coordinate of column Endj-1 when the band is clipped at LB
            uint8_t t = TB_MD;
            if (e >= 0 && e < BW) {       // uniform
                const int t0 = __shfl_sync(FULL, pM0, e >> 1), t1 = __shfl_sync(FULL, pM1, e >> 1);
                const int md = ((e & 1) ? t1 : t0) + GO;
                DLB += GE;
                t = 0;
                if (md >= DLB) { DLB = md; t = TB_MD; }
            }
            if (lane == 0) collb[i] = t;
        }
        // I chain: prefix max of V over the band
        const int Va = ina ? (pM0 + openA) + c0 * nextA : NEG;
        const int Vb = inb ? (pM1 + openA) + c1 * nextA : NEG;
        int inc = max(Va, Vb);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) inc = max(inc, __shfl_up_sync(FULL, inc, d));   // lanes below d get their own value back
        int Ua = __shfl_up_sync(FULL, inc, 1);   // U_{c0-1}
        if (lane == 0) Ua = NEG;
        const int Ub = max(Ua, Va);              // U_{c1-1}
        const int Ia = Ua - (c0 - 1) * nextA, Ib = Ub - (c1 - 1) * nextA;
        int Ma = NEG, Mb = NEG, Da = NEG, Db = NEG;
        uint32_t bits = 0;
        if (ina) {
            uint32_t t = 0;
            int xM = pM0;
            if (upDa > xM) { xM = upDa; t = TB_DM; }
            if (Ia > xM) { xM = Ia; t = TB_IM; }
            Ma = xM + ((a == (int)B[ja]) ? 1 : MMs);
            const bool freeB = (ja == 0) && Left;
            const int md = pM0 + (freeB ? 0 : GO);
            int d = upDa + (freeB ? 0 : GE);
            if (md >= d) { d = md; t |= TB_MD; }
            Da = d;
            if (Va >= Ua) t |= TB_MI;
            bits = t;
        }
        if (inb) {
            uint32_t t = 0;
            int xM = pM1;
            if (upDb > xM) { xM = upDb; t = TB_DM; }
            if (Ib > xM) { xM = Ib; t = TB_IM; }
            Mb = xM + ((a == (int)B[jb]) ? 1 : MMs);
            const bool freeB = (jb == 0) && Left;   // column 0 is the first cell of its row (OpenB / ExtB, viterbi.cpp:102-103)
            const int md = pM1 + (freeB ? 0 : GO);
            int d = upDb + (freeB ? 0 : GE);
            if (md >= d) { d = md; t |= TB_MD; }
            Db = d;
            if (Vb >= Ub) t |= TB_MI;
            bits |= t << 4;
        }
        tb[i * 32 + lane] = (uint8_t)bits;
        pM0 = Ma; pM1 = Mb; pD0 = Da; pD1 = Db;
    }
    // last row of DPI, viterbi.cpp:207-236 (strict >): chain over M(LA-1, j-1), j in [Startj, LB)
    const int j0f = (int)dlo - 1;   // column of coordinate 0 in row LA-1
    int I1;
    {
        const int gop = Right ? 0 : GO, ngex = Right ? 0 : -GE;
        const int ja = j0f + c0, jb = ja + 1;
        const bool ina = c0 < BW && ja >= 0 && ja < (int)LB, inb = c1 < BW && jb >= 0 && jb < (int)LB;
        int Mla = __shfl_up_sync(FULL, pM1, 1);   // M at coordinate c0-1
        if (lane == 0) Mla = NEG;
        const int Mlb = pM0;
        const int Va = ina ? (Mla + gop) + c0 * ngex : NEG;
        const int Vb = inb ? (Mlb + gop) + c1 * ngex : NEG;
        int inc = max(Va, Vb);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) inc = max(inc, __shfl_up_sync(FULL, inc, d));
        int Ua = __shfl_up_sync(FULL, inc, 1);
        if (lane == 0) Ua = NEG;
        const int Ub = max(Ua, Va);
        rowla[c0] = (ina && Va > Ua) ? TB_MI : 0;
        rowla[c1] = (inb && Vb > Ub) ? TB_MI : 0;
        const int tot = __shfl_sync(FULL, inc, 31);
        I1 = tot - ((int)LB - 1 - j0f) * ngex;
    }
    __syncwarp();
    int Score;
    {
        const int cf = (int)LB - 1 - j0f;   // coordinate of column LB-1 in the last row
        const int t0 = __shfl_sync(FULL, pM0, (cf >> 1) & 31), t1 = __shfl_sync(FULL, pM1, (cf >> 1) & 31);
        Score = (cf >= 0 && cf < BW) ? ((cf & 1) ? t1 : t0) : NEG;
    }
    int State = 0;  // 0 M, 1 D, 2 I
    if (DLB > Score) { Score = DLB; State = 1; }
    if (I1 > Score) { Score = I1; State = 2; }

    // traceback, tracebackbitmem.cpp:22-73.  The reference emits the state, then (unless a coordinate it needs is 0) reads
    // the trace bits of the cell it leaves and moves; here lane k looks at the cell k steps ahead along the current
    // direction, so a whole run of one state is emitted per iteration.  Everything stays uniform across the lanes.
    auto get = [&](int i, int j) -> uint32_t {
        if (i == (int)LA) {
            const int c = j - j0f;
            return (c >= 0 && c < BW) ? rowla[c] : 0u;
        }
        if (j == (int)LB) return collb[i];
        const int jz = (int)dlo + i - (int)LA;
        const int c = j - jz;
        if (c == -1) return (jz > 0) ? (uint32_t)TB_IM : 0u;   // TB[i][Startj-1] = IM, viterbi.cpp:119
        if (c < 0 || c >= BW) return 0u;
        return ((uint32_t)tb[i * 32 + (c >> 1)] >> ((c & 1) * 4)) & 0xFu;
    };
    int ti = (int)LA, tj = (int)LB;
    uint32_t curop = (uint32_t)State, curlen = 0;
#pragma unroll 1
    for (;;) {
        // position of lane k: k steps further along the run of State
        const int pi = ti - ((State == 2) ? 0 : lane), pj = tj - ((State == 1) ? 0 : lane);
        // 1 = the walk ends here without emitting, 2 = emits and ends, 3 = emits, changes state and moves, 0 = emits and moves
        uint32_t kind, nstate = (uint32_t)State;
        if (pi <= 0 && pj <= 0) kind = (pi == 0 && pj == 0) ? 1u : 4u;           // 4: beyond the end of the walk
        else if (pi < 0 || pj < 0) kind = 4u;
        else if (State == 0) {
            if (pi == 0 || pj == 0) kind = 2u;
            else {
                const uint32_t t = get(pi - 1, pj - 1);
                nstate = (t & TB_DM) ? 1u : ((t & TB_IM) ? 2u : 0u);
                kind = nstate != 0u ? 3u : 0u;
            }
        } else if (State == 1) {
            if (pi == 0) kind = 2u;
            else {
                const uint32_t t = get(pi - 1, pj);
                nstate = (t & TB_MD) ? 0u : 1u;
                kind = nstate != 1u ? 3u : 0u;
            }
        } else {
            if (pj == 0) kind = 2u;
            else {
                const uint32_t t = get(pi, pj - 1);
                nstate = (t & TB_MI) ? 0u : 2u;
                kind = nstate != 2u ? 3u : 0u;
            }
        }
        const uint32_t stop = __ballot_sync(FULL, kind != 0u);
        const int f = stop ? __ffs(stop) - 1 : 32;                 // first lane at which the run ends
        const uint32_t fk = __shfl_sync(FULL, kind, f & 31), fs = __shfl_sync(FULL, nstate, f & 31);
        const uint32_t emitted = (f == 32) ? 32u : (uint32_t)f + ((fk == 1u) ? 0u : 1u);
        if (emitted) {
            if ((uint32_t)State == curop) curlen += emitted;
            else {
                runs_append(rev, n_rev, curop, curlen, kRunCap, ovf, lane);
                curop = (uint32_t)State;
                curlen = emitted;
            }
        }
        if (f < 32 && fk != 3u) break;                             // the walk is over (kinds 1 and 2; 4 cannot come first)
        const int steps = (f == 32) ? 32 : f + 1;
        if (State != 2) ti -= steps;
        if (State != 1) tj -= steps;
        if (f < 32) State = (int)fs;
    }
    runs_append(rev, n_rev, curop, curlen, kRunCap, ovf, lane);
    return Score <= NEG / 2 ? NEG_INF : (float)Score;
}
#endif

// Full-window DP of the mate rescue (State1::Scan, scan.cpp:27: State1::Viterbi of the whole read against a window of
// 1024 (+ 2 QL) bases, Left = Right = true in the reference's call).  Same recurrence and traceback as viterbi_warp (32-row
// blocks swept as an anti-diagonal wavefront, lane = row), rebuilt around the memory system:
//   * 32-bit integers instead of fp32 (see viterbi_band for why scores and reachable trace bits are the same);
//   * the window is staged in shared memory once (it was one global load per cell);
//   * trace bits are packed, eight 4-bit cells per word, word index [row block][step / 8][lane]: one coalesced 128-byte
//     store per eight steps instead of 32 scattered byte stores per step (one 32-byte sector each);
//   * the previous block's last row (M, D) is fetched 32 columns at a time by the whole warp and handed to lane 0 through
//     shuffles (it was a dependent global load per step);
//   * the last-row insert chain is a max-plus prefix scan, 32 columns per step;
//   * the traceback walks whole runs (32 cells per step) as viterbi_band does.
// A: read strand (shared).  Bg: window in global memory.  Leaves the path REVERSED as RLE runs in E.ws->runs_a.
struct FullTB {
    const uint32_t *w;      // packed trace words
    const uint8_t *rowla;   // row LA (last-row insert chain), one byte per column
    uint32_t wpl;           // words per lane and row block
    uint32_t LA, LB, dlo, dhi;
    uint32_t SjL;           // Startj of the last row
};
__device__ __forceinline__ uint32_t full_tb_get(const FullTB &T, int i, int j) {
    if (i == (int)T.LA) return (j >= (int)T.SjL && j < (int)T.LB) ? (uint32_t)T.rowla[j] : 0u;
    uint32_t Sj, Ej;
    range_j(T.LA, T.LB, T.dlo, T.dhi, (uint32_t)i, Sj, Ej);
    const bool edge = (j == (int)T.LB);
    if (edge && Ej < T.LB) return TB_MD;                       // -inf >= -inf in the reference (viterbi.cpp:194)
    if (!edge) {
        if (j + 1 == (int)Sj) return TB_IM;                    // TB[i][Startj-1] = IM, viterbi.cpp:119
        if (j < (int)Sj || j >= (int)Ej) return 0u;
    }
    const uint32_t i0 = (uint32_t)i & ~31u, l = (uint32_t)i & 31u;
    uint32_t jb, te;
    range_j(T.LA, T.LB, T.dlo, T.dhi, i0, jb, te);
    const uint32_t st = (uint32_t)j - jb + l;
    return (T.w[((size_t)(i0 >> 5) * T.wpl + (st >> 3)) * 32 + l] >> (4 * (st & 7))) & 0xFu;
}

__device__ __noinline__ float viterbi_full(const Env &E, const uint8_t *A, uint32_t LA, const uint8_t *Bg, uint32_t LB, bool Left,
                                           bool Right, int &n_rev, int &ovf) {
    const int lane = URMB_LANE;
    uint16_t *rev = E.ws->runs_a;
    n_rev = 0;
    if (LA == 0 || LB == 0) return viterbi_warp<true>(E, A, LA, Bg, LB, Left, Right, n_rev, ovf);   // degenerate (never from Scan)
    constexpr int NEG = -(1 << 28);
    const int GO = E.P.GO, GE = E.P.GE, MMs = E.P.MM;
    uint32_t dlo = min(LA, LB), dhi = max(LA, LB);
    if (dlo > E.P.R) dlo -= E.P.R; else dlo = 1;
    dhi += E.P.R;
    if (dhi > LA + LB - 1) dhi = LA + LB - 1;
    // the window in shared memory when the flank-DP trace area is large enough for it, else read from global memory
    const uint8_t *B = Bg;
    if ((size_t)E.tb_rows * 32 >= (size_t)LB + 4) {
        uint8_t *sb = E.s_tb;
        for (uint32_t k = lane; k < LB; k += 32) sb[k] = __ldg(Bg + k);
        B = sb;
        __syncwarp();
    }
    int *rowM = reinterpret_cast<int *>(E.ws->rowM), *rowD = reinterpret_cast<int *>(E.ws->rowD);
    uint32_t *tbw = reinterpret_cast<uint32_t *>(E.ws->tb);
    const uint32_t wpl = (LB + 1 + 32 + 7) / 8 + 1;   // steps of a block <= LB + 1 + 31
    uint8_t *rowla = E.ws->tb + (size_t)((LA + 31) / 32) * wpl * 32 * 4;
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
        const int a = rowact ? (int)A[i] : 0x100;
        int outM = NEG, outD = NEG, Mdiag = NEG, I0 = NEG;
        const int openA = (Left && i == 0) ? 0 : GO, extA = (Left && i == 0) ? 0 : GE;
        const bool lastrow = rowact && (i == ilast);
        uint32_t *wrow = tbw + (size_t)(i0 >> 5) * wpl * 32 + lane;
        uint32_t acc = 0;
        int cM = NEG, cD = NEG;   // previous block's last row, 32 columns at a time (lane t: column jbase + s0 + t)
#pragma unroll 1
        for (int s = 0; s < nsteps; ++s) {
            if ((s & 31) == 0 && i0 > 0) {
                const uint32_t jj = jbase + (uint32_t)s + (uint32_t)lane;
                const bool inprev = (jj >= pS && jj < pE);
                cM = inprev ? rowM[jj] : NEG;
                cD = (inprev || (jj == LB && pE == LB)) ? rowD[jj] : NEG;
            }
            int upM = __shfl_up_sync(FULL, outM, 1), upD = __shfl_up_sync(FULL, outD, 1);
            const int bM = __shfl_sync(FULL, cM, s & 31), bD = __shfl_sync(FULL, cD, s & 31);
            if (lane == 0) { upM = bM; upD = bD; }
            const int js = (int)jbase + s - lane;
            const uint32_t j = (uint32_t)js;
            const bool incol = rowact && js >= (int)Sj && js < (int)Ej;
            const bool vcol = rowact && j == LB && Ej == LB && js >= 0;
            int myM = NEG, myD = NEG;
            uint32_t bits = 0;
            if (incol) {
                const int M0 = (j == 0) ? ((i == 0) ? 0 : NEG) : Mdiag;
                const int bch = (int)B[j];
                int xM = M0;
                if (upD > xM) { xM = upD; bits = TB_DM; }
                if (I0 > xM) { xM = I0; bits = TB_IM; }
                myM = xM + ((a == bch) ? 1 : MMs);
                const bool freeB = (j == 0) && Left;
                const int md = M0 + (freeB ? 0 : GO);
                int d = upD + (freeB ? 0 : GE);
                if (md >= d) { d = md; bits |= TB_MD; }
                myD = d;
                const int mi = M0 + openA;
                I0 += extA;
                if (mi >= I0) { I0 = mi; bits |= TB_MI; }
            } else if (vcol) {  // viterbi.cpp:187-200, end of Drow[]
                const int md = Mdiag + GO;
                int d = upD + GE;
                if (md >= d) { d = md; bits = TB_MD; }
                myD = d;
            }
            acc |= bits << (4 * (s & 7));
            if ((s & 7) == 7) {
                wrow[(size_t)(s >> 3) * 32] = acc;
                acc = 0;
            }
            Mdiag = upM;
            outM = myM;
            outD = myD;
            if (lastrow) {
                if (incol) { rowM[j] = myM; rowD[j] = myD; }
                else if (vcol) rowD[j] = myD;
            }
        }
        if (nsteps & 7) wrow[(size_t)((nsteps - 1) >> 3) * 32] = acc;
        __syncwarp();
    }

    // last row of DPI, viterbi.cpp:207-236 (strict >): I1 = max-plus chain over M(LA-1, j-1), j in [Startj, LB)
    uint32_t SjL, EjL;
    range_j(LA, LB, dlo, dhi, LA - 1, SjL, EjL);
    const int gop = Right ? 0 : GO, ngex = Right ? 0 : -GE;
    int Ucarry = NEG;   // running max of V_c = (M(LA-1, j-1) + gop) + c * ngex over c = j - SjL
    for (uint32_t c0 = 0; SjL + c0 < EjL; c0 += 32) {
        const uint32_t c = c0 + lane, j = SjL + c;
        const bool in = j < EjL;
        const int mp = (in && j > SjL) ? rowM[j - 1] : NEG;   // Mrow[Startj-1] = -inf, viterbi.cpp:212
        const int V = in ? (mp + gop) + (int)c * ngex : NEG;
        int inc = V;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) inc = max(inc, __shfl_up_sync(FULL, inc, d));
        int Uprev = __shfl_up_sync(FULL, inc, 1);
        if (lane == 0) Uprev = NEG;
        Uprev = max(Uprev, Ucarry);
        if (in) rowla[j] = (V > Uprev) ? TB_MI : 0;
        Ucarry = max(Ucarry, __shfl_sync(FULL, inc, 31));
    }
    const int I1 = (EjL > SjL) ? Ucarry - (int)(EjL - 1 - SjL) * ngex : NEG;
    __syncwarp();
    int Score = rowM[LB - 1];
    int State = 0;  // 0 M, 1 D, 2 I
    const int FinalD = rowD[LB];
    if (FinalD > Score) { Score = FinalD; State = 1; }
    if (I1 > Score) { Score = I1; State = 2; }

    FullTB T{tbw, rowla, wpl, LA, LB, dlo, dhi, SjL};
    int ti = (int)LA, tj = (int)LB;
    uint32_t curop = (uint32_t)State, curlen = 0;
#pragma unroll 1
    for (;;) {   // run-wise traceback, see viterbi_band
        const int pi = ti - ((State == 2) ? 0 : lane), pj = tj - ((State == 1) ? 0 : lane);
        uint32_t kind, nstate = (uint32_t)State;
        if (pi <= 0 && pj <= 0) kind = (pi == 0 && pj == 0) ? 1u : 4u;
        else if (pi < 0 || pj < 0) kind = 4u;
        else if (State == 0) {
            if (pi == 0 || pj == 0) kind = 2u;
            else {
                const uint32_t t = full_tb_get(T, pi - 1, pj - 1);
                nstate = (t & TB_DM) ? 1u : ((t & TB_IM) ? 2u : 0u);
                kind = nstate != 0u ? 3u : 0u;
            }
        } else if (State == 1) {
            if (pi == 0) kind = 2u;
            else {
                const uint32_t t = full_tb_get(T, pi - 1, pj);
                nstate = (t & TB_MD) ? 0u : 1u;
                kind = nstate != 1u ? 3u : 0u;
            }
        } else {
            if (pj == 0) kind = 2u;
            else {
                const uint32_t t = full_tb_get(T, pi, pj - 1);
                nstate = (t & TB_MI) ? 0u : 2u;
                kind = nstate != 2u ? 3u : 0u;
            }
        }
        const uint32_t stop = __ballot_sync(FULL, kind != 0u);
        const int f = stop ? __ffs(stop) - 1 : 32;
        const uint32_t fk = __shfl_sync(FULL, kind, f & 31), fs = __shfl_sync(FULL, nstate, f & 31);
        const uint32_t emitted = (f == 32) ? 32u : (uint32_t)f + ((fk == 1u) ? 0u : 1u);
        if (emitted) {
            if ((uint32_t)State == curop) curlen += emitted;
            else {
                runs_append(rev, n_rev, curop, curlen, kRunCap, ovf, lane);
                curop = (uint32_t)State;
                curlen = emitted;
            }
        }
        if (f < 32 && fk != 3u) break;
        const int steps = (f == 32) ? 32 : f + 1;
        if (State != 2) ti -= steps;
        if (State != 1) tj -= steps;
        if (f < 32) State = (int)fs;
    }
    runs_append(rev, n_rev, curop, curlen, kRunCap, ovf, lane);
    return Score <= NEG / 2 ? NEG_INF : (float)Score;
}

// Flank DP dispatch: bands up to 64 wide (always true for LB = LA + 2R (+1), R <= 12) run with the band across the
// lanes and trace bits in shared memory; anything else (never seen from AlignHSP) takes the row-block kernel.
__device__ __noinline__ float flank_viterbi(const Env &E, const uint8_t *A, uint32_t LA, uint32_t TLo, uint32_t LB, bool Left,
                               bool Right, int &n_rev, int &ovf) {
    uint32_t dlo = min(LA, LB), dhi = max(LA, LB);
    if (dlo > E.P.R) dlo -= E.P.R; else dlo = 1;
    dhi += E.P.R;
    if (LA + LB >= 1 && dhi > LA + LB - 1) dhi = LA + LB - 1;
    const bool fits = (LA + 1 <= E.tb_rows) && (dhi - dlo + 1 <= 64) && (LA > 0) && (LB > 0);
    if (fits) return viterbi_band(E, A, LA, E.s_win, LB, Left, Right, n_rev, ovf);
    return viterbi_warp<true>(E, A, LA, E.ix.seq + TLo, LB, Left, Right, n_rev, ovf);
}

// State1::AlignHSP, alignhsp.cpp:60-172
__device__ __noinline__ int align_hsp(const Env &E, Mate &m, int HSPIndex) {
    const uint8_t fl = m.g->hsp_flags[HSPIndex];
    if (fl & 2) return -1;
    __syncwarp();   // all lanes have read the flags
    if (URMB_LANE == 0) m.g->hsp_flags[HSPIndex] = fl | 2;
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
        for (uint32_t k = URMB_LANE; k < LeftTL; k += 32) {
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
        for (int k = nrev - 1; k >= 0; --k) runs_append(path, np, rev[k] & 3u, rev[k] >> 2, pcap, ovf, URMB_LANE);
        CombinedTLo = LeftTLo + LeftICount;
        const int AllGapScore = E.P.GO + ((int)LeftQL - 1) * E.P.GE;
        if (AllGapScore > LeftScore) LeftScore = AllGapScore;
        TotalScore += LeftScore;
        TotalPen += (int)LeftQL - LeftScore;
        if (TotalPen > m.MaxPenalty) return -1;
    }
    runs_append(path, np, 0, HSPLength, pcap, ovf, URMB_LANE);
    const uint32_t RightQLo = StartPosQ + HSPLength;
    // The start of the alignment is known once the left flank is done.  AddHitX (state1.cpp:508-551) drops a hit whose
    // 64-base bucket is taken before it looks at anything else, and nothing below has another effect, so the right-flank
    // DP of such a hit is skipped: same result (-1, no change of state), about half of the DP work of the candidates that
    // sit on a locus already found (tandem repeats in a scan window produce hundreds of them).
    if (RightQLo < QL && overlaps_hit(E, m, CombinedTLo)) return -1;
    if (RightQLo < QL) {
        const uint32_t RightQL = QL - RightQLo;
        const uint32_t RightTLo = StartPosDB + HSPLength;
        uint32_t RightTHi = RightTLo + RightQL + BRN * E.P.R;
        if (RightTHi >= TL) RightTHi = TL - 1;
        if (RightTHi < RightTLo) return -1;
        const uint32_t RightTL = RightTHi - RightTLo + 1;
        bool dash = false;
        for (uint32_t k = URMB_LANE; k < RightTL; k += 32) {
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
        for (int k = nrev - 1; k >= first; --k) runs_append(path, np, rev[k] & 3u, rev[k] >> 2, pcap, ovf, URMB_LANE);
        if (keepI) runs_append(path, np, 2, 1, pcap, ovf, URMB_LANE);
        const int AllGapScore = E.P.GO + ((int)RightQL - 1) * E.P.GE;
        if (AllGapScore > RightScore) RightScore = AllGapScore;
        TotalScore += RightScore;
        TotalPen += (int)RightQL - RightScore;
        if (TotalPen > m.MaxPenalty) return -1;
    }
    if (ovf) m.overflow |= 16;
    return add_hit(E, m, CombinedTLo, Plus, TotalScore, path, np);
}

// State1::ExtendScan, extendscan.cpp:51-187 (returns hit index or -1). Uniform arguments.
// The order-dependent half of ExtendScan from the packed pure result x (pure_ext with LeftCountsPen = false: the left walk
// adds no penalty, quirk 5; EXT_NONE when the diagonal starts before the genome).
__device__ __noinline__ int extend_scan_apply(const Env &E, Mate &m, uint32_t SeedPosQ, uint32_t SeedPosDB, bool Plus, uint32_t x) {
    if (x == EXT_NONE) return -1;
    const uint32_t DBLo = SeedPosDB - SeedPosQ;
    if (ext_nmis(x) * -E.P.MM > m.MaxPenalty) return -1;
    const int Best = ext_best(x), Start = ext_start(x), End = ext_end(x);
    const int MinHSPScore = (int)E.ix.word_len * 2;
    if (Start == 0 && End == (int)m.QL - 1) return add_hit(E, m, DBLo, Plus, Best, nullptr, 0);
    if (Best < MinHSPScore) return -1;
    int k = add_hsp_scan(E, m, (uint32_t)Start, DBLo + (uint32_t)Start, Plus, (uint32_t)(End - Start + 1), Best);
    if (k < 0) return -1;
    return align_hsp(E, m, k);
}

// State1::CalcMAPQ6, search1m6.cpp:9-33 (fp64, same operation order)
__device__ __noinline__ uint32_t calc_mapq6(const Mate &m) {
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
        if (K > 0) load_blob_s(E.ix.blob, Slot2, T, Pos);
        if ((uint32_t)URMB_LANE == K) mypos = Pos;
        ++K;
        if (K == E.ix.max_ix) return K;
        if (T == T_PLUS1 || T == T_BOTH1) return 1;
        if (T == T_END) return K;
        if (T == T_LONG_MINE || T == T_LONG_OTHER) {
            uint32_t StepA = Pos & 0xffffu, StepB = Pos >> 16;
            uint64_t SlotA = add_mod(Slot2, StepA, SC);
            Slot2 = add_mod(SlotA, StepB, SC);
            uint32_t ta, pa;
            load_blob_s(E.ix.blob, SlotA, ta, pa);
            if ((uint32_t)URMB_LANE == K - 1) mypos = pa;
        } else {
            Slot2 = add_mod(Slot2, T & T_NEXT_MASK, SC);
        }
    }
}

__device__ __forceinline__ uint32_t m_tally(const Mate &m, int strand /*0 plus,1 minus*/, uint32_t q) {
    return __ldg(m.tally + strand * m.qcap + q);
}
__device__ __forceinline__ uint32_t m_pos(const Mate &m, int strand, uint32_t q) { return __ldg(m.pos + strand * m.qcap + q); }
__device__ __forceinline__ uint32_t m_ext(const Mate &m, int strand, uint32_t q) { return __ldg(m.ext + strand * m.qcap + q); }
__device__ __forceinline__ uint64_t m_slot(const Env &E, const Mate &m, int strand, uint32_t q) {
    return slot_of_s(E.ix, m.rv, strand, q);
}

// Lane-local GetRow_Blob (ufindex.cpp:883-943) limited to what the "rows <= 2 now, longer rows later" logic
// needs: n = 0 (not mine), 1, 2, or 3 meaning "RowLength > 2"; p0/p1 = the first two positions.
__device__ __noinline__ void row_head3(const Env &E, uint64_t Slot, uint32_t Tally, uint32_t Pos0, uint32_t &n, uint32_t &p0,
                          uint32_t &p1) {
    n = 0; p0 = 0; p1 = 0;
    uint32_t T = Tally, Pos = Pos0, K = 0;
    if ((T & T_MY_BIT) == 0) return;
    uint64_t Slot2 = Slot;
    const uint64_t SC = E.ix.slot_count;
    for (;;) {
        if (K > 0) load_blob_s(E.ix.blob, Slot2, T, Pos);
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
            load_blob_s(E.ix.blob, SlotA, ta, pa);
            if (K == 1) p0 = pa; else if (K == 2) p1 = pa;
        } else {
            Slot2 = add_mod(Slot2, T & T_NEXT_MASK, SC);
        }
    }
}

// First round over the lists of owned non-BOTH1 slots of both strands (search1m6.cpp:181-199,
// search1pepend.cpp:53-68): rows of length <= 2 are extended now, longer rows are deferred through `def0` / `def1`.
// Three passes so that every lane has work:
//   1. the heads of all rows of both lists, one dependent-gather chain per lane;
//   2. the candidates of the short rows laid out flat, 32 pure extensions at a time;
//   3. the order-dependent bookkeeping, visiting the plus list and then the minus list in the reference's order.
// The hit-overlap precheck and the penalty bound of passes 1 and 2 use the state at entry: both only ever get
// stricter, so a candidate dropped here is still a no-op when the reference reaches it, and pass 3 re-applies the
// current state (extend_apply).  Taking the minus list through passes 1 and 2 together with the plus list halves the
// number of dependent-gather rounds per mate; either list may be empty.
// `def0` / `def1` may alias `list0` / `list1` (written in pass 3 only; pass 1 keeps its own copy of the lists).
__device__ __noinline__ void rows_short_round(const Env &E, Mate &m, const uint8_t *list0, int n0, const uint8_t *list1, int n1,
                                              uint8_t *def0, int &nd0, uint8_t *def1, int &nd1) {
    nd0 = nd1 = 0;
    const int n = n0 + n1;
    if (n <= 0) return;
    constexpr int CAP = 2 * kMaxLen;
    uint32_t *hp0 = reinterpret_cast<uint32_t *>(E.ws->rowM);   // [CAP] first position of the row
    uint32_t *hp1 = hp0 + CAP;                                  // [CAP] second position
    uint32_t *hx0 = reinterpret_cast<uint32_t *>(E.ws->rowD);   // [CAP] pure extension results
    uint32_t *hx1 = hx0 + CAP;
    uint8_t *hq = E.ws->tb, *hn = E.ws->tb + CAP;               // [CAP] QPos, row length (3 = longer than 2)
    uint16_t *flat = reinterpret_cast<uint16_t *>(E.ws->tb + 2 * CAP);   // [2 * CAP] entry | which << 15
    const uint32_t lt = (1u << URMB_LANE) - 1u;
    int total = 0;
#pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += 32) {   // pass 1
        const int i = i0 + URMB_LANE;
        const bool valid = i < n;
        const int s = (i < n0) ? 0 : 1;
        const uint32_t QPos = valid ? (s ? list1[i - n0] : list0[i]) : 0u;
        uint32_t rn = 0, p0 = 0, p1 = 0;
        if (valid) row_head3(E, m_slot(E, m, s, QPos), m_tally(m, s, QPos), m_pos(m, s, QPos), rn, p0, p1);
        if (rn > E.ix.max_ix) rn = E.ix.max_ix;
        const bool c0 = valid && rn >= 1 && rn <= 2 && p0 >= QPos && !overlaps_hit_lane(m, p0 - QPos);
        const bool c1 = valid && rn == 2 && p1 >= QPos && !overlaps_hit_lane(m, p1 - QPos);
        const uint32_t b0 = __ballot_sync(FULL, c0), b1 = __ballot_sync(FULL, c1);
        if (valid) {
            hq[i] = (uint8_t)QPos;
            hn[i] = (uint8_t)rn;
            hp0[i] = p0;
            hp1[i] = p1;
            hx0[i] = EXT_NONE;
            hx1[i] = EXT_NONE;
        }
        if (c0) flat[total + __popc(b0 & lt)] = (uint16_t)i;
        total += __popc(b0);
        if (c1) flat[total + __popc(b1 & lt)] = (uint16_t)(i | 0x8000);
        total += __popc(b1);
    }
    __syncwarp();
#pragma unroll 1
    for (int f0 = 0; f0 < total; f0 += 32) {   // pass 2
        const int f = f0 + URMB_LANE;
        if (f < total) {
            const uint32_t e = flat[f], i = e & 0x7FFFu;
            const uint32_t pz = (e & 0x8000u) ? hp1[i] : hp0[i];
            const uint32_t x = pure_ext(E.ix, E.P, m.rv, (int)i < n0, hq[i], pz, true, m.MaxPenalty);
            if (e & 0x8000u) hx1[i] = x; else hx0[i] = x;
        }
    }
    __syncwarp();
#pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += 32) {   // pass 3; a block of 32 entries may straddle the two lists
        const int i = i0 + URMB_LANE;
        const bool valid = i < n;
        const uint32_t QPos = valid ? hq[i] : 0u, rn = valid ? hn[i] : 0u;
        const uint32_t p0 = valid ? hp0[i] : 0u, p1 = valid ? hp1[i] : 0u;
        const uint32_t x0 = valid ? hx0[i] : EXT_NONE, x1 = valid ? hx1[i] : EXT_NONE;
        const bool a0 = !ext_is_noop(E, x0, (int)m.QL, m.MaxPenalty), a1 = !ext_is_noop(E, x1, (int)m.QL, m.MaxPenalty);
        uint32_t vmask = __ballot_sync(FULL, valid && (rn > 2 || a0 || a1));
        const uint32_t m0 = __ballot_sync(FULL, a0), m1 = __ballot_sync(FULL, a1), big = __ballot_sync(FULL, rn > 2);
        bool stored;
        while (vmask) {
            const int b = __ffs(vmask) - 1;
            vmask &= vmask - 1;
            const uint32_t qb = __shfl_sync(FULL, QPos, b);
            const bool plus = i0 + b < n0;
            if (big >> b & 1u) {
                if (plus) {
                    if (URMB_LANE == 0) def0[nd0] = (uint8_t)qb;
                    ++nd0;
                } else {
                    if (URMB_LANE == 0) def1[nd1] = (uint8_t)qb;
                    ++nd1;
                }
                continue;
            }
            const uint32_t pb0 = __shfl_sync(FULL, p0, b), pb1 = __shfl_sync(FULL, p1, b);
            const uint32_t xb0 = __shfl_sync(FULL, x0, b), xb1 = __shfl_sync(FULL, x1, b);
            if (m0 >> b & 1u) extend_apply(E, m, qb, pb0, plus, xb0, stored);
            if (m1 >> b & 1u) extend_apply(E, m, qb, pb1, plus, xb1, stored);
        }
    }
    __syncwarp();
}

// Lane-local GetRow_Blob (ufindex.cpp:883-943): the whole row into out[0 .. MaxIx); returns the row length.
__device__ uint32_t row_walk_lane(const Env &E, uint64_t Slot, uint32_t Tally, uint32_t Pos0, uint32_t *out) {
    uint32_t T = Tally;
    if ((T & T_MY_BIT) == 0) return 0;
    uint64_t Slot2 = Slot;
    uint32_t Pos = Pos0, K = 0;
    const uint64_t SC = E.ix.slot_count;
    for (;;) {
        if (K > 0) load_blob_s(E.ix.blob, Slot2, T, Pos);
        out[K] = Pos;
        ++K;
        if (K == E.ix.max_ix) return K;
        if (T == T_PLUS1 || T == T_BOTH1) return 1;
        if (T == T_END) return K;
        if (T == T_LONG_MINE || T == T_LONG_OTHER) {
            const uint32_t StepA = Pos & 0xffffu, StepB = Pos >> 16;
            const uint64_t SlotA = add_mod(Slot2, StepA, SC);
            Slot2 = add_mod(SlotA, StepB, SC);
            uint32_t ta, pa;
            load_blob_s(E.ix.blob, SlotA, ta, pa);
            out[K - 1] = pa;
        } else {
            Slot2 = add_mod(Slot2, T & T_NEXT_MASK, SC);
        }
    }
}

// Deferred (long) rows, search1m6.cpp:205-243 / search1pepend.cpp:89-110: the reference walks one row at a time and
// extends its positions in order.  Here 32 rows are walked at once (one dependent-gather chain per lane), their
// candidates are laid out row-major in per-warp scratch, the pure extensions run 32 candidates at a time, and the
// order-dependent bookkeeping visits the survivors in exactly the reference's order.
__device__ __noinline__ void rows_long_batch(const Env &E, Mate &m, const uint8_t *list0, int n0, const uint8_t *list1, int n1) {
    uint32_t *stage = reinterpret_cast<uint32_t *>(E.ws->rowM);   // [32 rows][32 positions]
    uint32_t *fpos = reinterpret_cast<uint32_t *>(E.ws->rowD);    // flat candidate positions (<= 1024)
    uint8_t *fq = E.ws->tb;                                       // flat candidate QPos
    const uint32_t mx = E.ix.max_ix > 32u ? E.ix.max_ix : 32u;    // row stride
    if (mx > 32u) {   // an index built with -maxix above 32 (ufindexio.cpp:135-136): the three arrays in the trace-bit area
        static_assert((size_t)kBigRows * kBigCols >= (size_t)32 * URMB_MAX_IX * 9, "trace-bit area holds the rows of -maxix URMB_MAX_IX");
        stage = reinterpret_cast<uint32_t *>(E.ws->tb);
        fpos = stage + 32u * mx;
        fq = reinterpret_cast<uint8_t *>(fpos + 32u * mx);
    }
    const int n = n0 + n1;   // the plus list, then the minus list (one may be empty)
#pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + URMB_LANE;
        const bool valid = i < n;
        const int s = (i < n0) ? 0 : 1;
        const uint32_t QPos = valid ? (s ? list1[i - n0] : list0[i]) : 0u;
        uint32_t len = 0;
        if (valid) len = row_walk_lane(E, m_slot(E, m, s, QPos), m_tally(m, s, QPos), m_pos(m, s, QPos), stage + mx * URMB_LANE);
        uint32_t off = len;   // exclusive prefix sum over lanes
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, off, d);
            if (URMB_LANE >= d) off += t;
        }
        const uint32_t total = __shfl_sync(FULL, off, 31);
        off -= len;
        // candidates of the plus rows come first: [0, nplus) is the plus strand
        const int lp = min(max(n0 - i0, 0), 32);   // lanes [0, lp) hold plus rows
        const uint32_t nplus = (lp >= 32) ? total : __shfl_sync(FULL, off, lp);
        for (uint32_t k = 0; k < len; ++k) {
            fpos[off + k] = stage[mx * URMB_LANE + k];
            fq[off + k] = (uint8_t)QPos;
        }
        __syncwarp();
#pragma unroll 1
        for (uint32_t f0 = 0; f0 < total; f0 += 32) {
            const uint32_t f = f0 + URMB_LANE;
            uint32_t x = EXT_NONE, q = 0, pz = 0;
            if (f < total) {
                q = fq[f];
                pz = fpos[f];
                if (pz >= q && !overlaps_hit_lane(m, pz - q)) x = pure_ext(E.ix, E.P, m.rv, f < nplus, q, pz, true, m.MaxPenalty);
            }
            uint32_t am = __ballot_sync(FULL, !ext_is_noop(E, x, (int)m.QL, m.MaxPenalty));
            bool stored;
            while (am) {
                const int r = __ffs(am) - 1;
                am &= am - 1;
                extend_apply(E, m, __shfl_sync(FULL, q, r), __shfl_sync(FULL, pz, r), f0 + (uint32_t)r < nplus,
                             __shfl_sync(FULL, x, r), stored);
            }
        }
        __syncwarp();
    }
}

__device__ void reset_search(const Env &E, Mate &m) {
    m.HitCount = 0;
    m.HSPCount = 0;
    m.Top = -1;
    m.Best = 0;
    m.Second = 0;
    m.BestHSP = 0;
    m.nRuns = 0;
    m.Mapq = 0xFFFFFFFFu;
    m.MaxPenalty = E.P.MAXPEN;
    m.nPend[0] = m.nPend[1] = 0;
}

// State1::Search_Lo, search1m6.cpp:35-277, cut at its phase borders so that the single-end search runs as small
// kernels (seeds / HSP alignment / rows / HSP alignment).  Each piece returns true when the search is over.
// Phases 1 and 2 (search1m6.cpp:48-131): BOTH1 seeds at stride W (plus then minus at each QPos), then all remaining
// QPos.  The visit order is laid out 32 visits at a time into the seed list together with the probe kernel's pure
// results; the order-dependent bookkeeping then runs over the candidates that can still do something.
__device__ __noinline__ bool se_phase12(const Env &E, Mate &m) {
    const uint32_t W = E.ix.word_len;
    const int QL = (int)m.QL;
    if (m.QL < W) { m.Mapq = 0; return true; }   // reference underflows: report no hit
    const uint32_t QWC = m.QWC;
    m.MaxPenalty = E.P.MAXPEN;
    const int MinScorePhase1 = QL + E.P.XP1 * E.P.MM;
    m.BestHSP = 0;
    const uint32_t n1 = (QWC + W - 1) / W;          // phase-1 QPos count
    const uint32_t nvis = 2 * QWC;
    m.nSeeds = 0;
#pragma unroll 1
    for (uint32_t v0 = 0; v0 < nvis; v0 += 32) {
        const uint32_t v = v0 + URMB_LANE;
        uint32_t q = 0;
        const int sgn = (int)(v & 1u);
        if (v < 2 * n1) q = (v >> 1) * W;
        else {
            const uint32_t vv = (v - 2 * n1) >> 1;
            q = (W > 1) ? (vv / (W - 1)) * W + vv % (W - 1) + 1 : QWC;
        }
        const bool c = (v < nvis) && (q < QWC) && (m_tally(m, sgn, q) == T_BOTH1);
        const uint32_t bal = __ballot_sync(FULL, c);
        if (c) m.sd_qs[m.nSeeds + __popc(bal & ((1u << URMB_LANE) - 1u))] = (uint16_t)(q | ((uint32_t)sgn << 15));
        m.nSeeds += __popc(bal);
    }
    seeds_init_dead(E, m, true);
    for (int w0 = 0; w0 < m.nSeeds; w0 += 32) {
        uint32_t live = ~m.sd_dead[w0 >> 5];
        while (live) {
            const int bit = __ffs(live) - 1;
            const int Score = apply_seed(E, m, w0 + bit);
            if (Score >= MinScorePhase1) { m.Mapq = calc_mapq6(m); return true; }
            live = ~m.sd_dead[w0 >> 5] & ((bit == 31) ? 0u : (0xFFFFFFFFu << (bit + 1)));
        }
    }
    return false;
}
__device__ __forceinline__ bool se_phase3_needed(const DevParams &P, int QL, int BestHSP) {
    return BestHSP > (QL * P.TERM3_PCT) / 100;
}
// Phase 3 (search1m6.cpp:133-147)
__device__ __noinline__ bool se_phase3(const Env &E, Mate &m) {
    const int QL = (int)m.QL;
    if (se_phase3_needed(E.P, QL, m.BestHSP)) {
        for (int i = 0; i < m.HSPCount; ++i) align_hsp(E, m, i);
        if (m.Best >= QL + E.P.XP1 * E.P.MM) { m.Mapq = calc_mapq6(m); return true; }
    }
    return false;
}
// Phases 4 and 5 (search1m6.cpp:149-245): non-BOTH1 owned slots; rows <= 2 now, longer rows deferred.  Two pieces so
// that the staged search can run them as two kernels: the deferred lists stay in m.g->todo, their lengths in m.nPend
// (the paired-end pending counters, unused by the single-end search).
__device__ __noinline__ bool se_phase4(const Env &E, Mate &m) {
    const int QL = (int)m.QL;
    const uint32_t QWC = m.QWC;
    int nls[2];
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        // the owned non-BOTH1 slots of this strand in QPos order (search1m6.cpp:170-203), then rows <= 2 / deferral
        uint8_t *lst = m.g->todo[s];
        int nl = 0;
#pragma unroll 1
        for (uint32_t q0 = 0; q0 < QWC; q0 += 32) {
            const uint32_t q = q0 + URMB_LANE;
            const uint32_t T = (q < QWC) ? m_tally(m, s, q) : 0;
            const bool cand = (T != T_FREE && T != T_BOTH1 && (T & T_MY_BIT));
            const uint32_t bal = __ballot_sync(FULL, cand);
            if (cand) lst[nl + __popc(bal & ((1u << URMB_LANE) - 1u))] = (uint8_t)q;
            nl += __popc(bal);
        }
        nls[s] = nl;
    }
    __syncwarp();
    int nt0 = 0, nt1 = 0;
    rows_short_round(E, m, m.g->todo[0], nls[0], m.g->todo[1], nls[1], m.g->todo[0], nt0, m.g->todo[1], nt1);
    m.nPend[0] = nt0;
    m.nPend[1] = nt1;
    __syncwarp();
    if (m.Best >= QL + E.P.XP3 * E.P.MM) { m.Mapq = calc_mapq6(m); return true; }
    return false;
}
__device__ __forceinline__ bool se_phase5_done(const DevParams &P, int QL, int Best) { return Best >= QL + P.XP4 * P.MM; }
__device__ __noinline__ bool se_phase5(const Env &E, Mate &m) {
    rows_long_batch(E, m, m.g->todo[0], m.nPend[0], m.g->todo[1], m.nPend[1]);
    if (se_phase5_done(E.P, (int)m.QL, m.Best)) { m.Mapq = calc_mapq6(m); return true; }
    return false;
}
__device__ bool se_phase45(const Env &E, Mate &m) {
    if (se_phase4(E, m)) return true;
    return se_phase5(E, m);
}
// Phase 6 (search1m6.cpp:247-276)
__device__ __noinline__ void se_phase6(const Env &E, Mate &m) {
    for (int i = 0; i < m.HSPCount; ++i) align_hsp(E, m, i);
    m.Mapq = calc_mapq6(m);
}

// ---- paired-end ------------------------------------------------------------------------
// GetFirstBoth1Seed / GetNextBoth1Seed (getseed.cpp:9-54,56-138) for the whole read at once.
// The iterator visits v = 2k+strand with QPos = (27k) % QWC and returns a BOTH1 slot unless its diagonal equals
// the diagonal of the previously RETURNED seed.  A skipped BOTH1 slot has, by construction, the diagonal of the
// last returned one, so "returned" == "diagonal differs from the previous BOTH1 visit": a run-length dedup that
// 32 lanes evaluate with one ballot and one shuffle per 32 visits.  Pending lists (state1.h:86-87) in visit order:
//   plus visit : owned non-BOTH1 slot;
//   minus visit: owned non-BOTH1 slot, except when the plus slot of the same k was just returned -- then only a
//                BOTH1 slot on the same diagonal is pushed (getseed.cpp:63-85).
// The lists are only consumed after the seed loop ran to completion, so building them up front is exact.
__device__ __noinline__ void build_seeds_pe(const Env &E, Mate &m) {
    const uint32_t QWC = m.QWC;
    m.nSeeds = 0;
    m.nPend[0] = m.nPend[1] = 0;
    bool have_prev = false;
    uint32_t prev_diag = 0;
    const uint32_t lt = (1u << URMB_LANE) - 1u;
    // (27 k) % QWC without a division per visit: 27 k < 2^13 and QWC < 2^8, so floor(n / QWC) = (n * ceil(2^32 / QWC)) >> 32 exactly
    const uint32_t qrec = QWC ? (uint32_t)((0x100000000ull + QWC - 1) / QWC) : 0u;
#pragma unroll 1
    for (uint32_t v0 = 0; v0 < 2 * QWC; v0 += 32) {
        const uint32_t v = v0 + URMB_LANE, k = v >> 1;
        const int sgn = (int)(v & 1u);
        const bool valid = k < QWC;
        const uint32_t n27 = k * PRIME_STRIDE;
        const uint32_t QPos = valid ? n27 - __umulhi(n27, qrec) * QWC : 0;
        const uint32_t T = valid ? m_tally(m, sgn, QPos) : 0u;
        const uint32_t Pz = valid ? m_pos(m, sgn, QPos) : 0u;
        const bool mine = (T & T_MY_BIT) != 0;          // invalid words carry T_FREE
        const bool b1 = mine && T == T_BOTH1;
        const uint32_t diag = Pz - QPos;
        const uint32_t b1mask = __ballot_sync(FULL, b1);
        const uint32_t below = b1mask & lt;
        const uint32_t pd = __shfl_sync(FULL, diag, below ? 31 - __clz(below) : 0);
        const bool hasp = below ? true : have_prev;
        const uint32_t pdiag = below ? pd : prev_diag;
        const bool ret = b1 && (!hasp || diag != pdiag);
        const uint32_t retmask = __ballot_sync(FULL, ret);
        const bool plus_ret = sgn && ((retmask >> ((URMB_LANE - 1) & 31)) & 1u);
        const bool pend = sgn ? (plus_ret ? (b1 && !ret) : (mine && !b1)) : (mine && !b1);
        const uint32_t pp = __ballot_sync(FULL, pend && !sgn), pm = __ballot_sync(FULL, pend && sgn);
        if (ret) {
            const int i = m.nSeeds + __popc(retmask & lt);
            m.sd_db[i] = Pz;
            m.sd_qs[i] = (uint16_t)(QPos | ((uint32_t)sgn << 15));
        }
        if (pend) {
            const int i = m.nPend[sgn] + __popc((sgn ? pm : pp) & lt);
            m.g->pend[sgn][i] = (uint8_t)QPos;
        }
        m.nSeeds += __popc(retmask);
        m.nPend[0] += __popc(pp);
        m.nPend[1] += __popc(pm);
        if (b1mask) {
            have_prev = true;
            prev_diag = __shfl_sync(FULL, diag, 31 - __clz(b1mask));
        }
    }
    seeds_init_dead(E, m, false);
}

// State1::SearchPE_Pending, search1pepend.cpp:9-130 (k is always UINT_MAX at the call sites)
// The function is cut at its two internal borders so that the paired-end second pass can run it as three small
// kernels (HSP alignment / pending rows / HSP alignment): pend_stage_a returns true when the search is over.
__device__ __forceinline__ bool pend_done_at_entry(const Env &E, const Mate &m) {   // search1pepend.cpp:27-31
    return m.Best >= (int)m.QL + E.P.XP1 * E.P.MM;
}
__device__ __noinline__ bool pend_stage_a(const Env &E, Mate &m) {
    const int QL = (int)m.QL;
    m.MaxPenalty = E.P.MAXPEN;
    const int MinScorePhase1 = QL + E.P.XP1 * E.P.MM;
    const int TermHSPScorePhase3 = (QL * E.P.TERM3_PCT) / 100;
    if (m.Best >= MinScorePhase1) { m.Mapq = calc_mapq6(m); return true; }
    if (m.BestHSP >= TermHSPScorePhase3) {
        for (int i = 0; i < m.HSPCount; ++i) align_hsp(E, m, i);
        if (m.Best >= MinScorePhase1) { m.Mapq = calc_mapq6(m); return true; }
    }
    __syncwarp();
    return false;
}
// Pending round 1 (rows <= 2 extended, longer rows deferred: the lists are compacted in place and nPend becomes the
// number of deferred rows) and pending round 2 (the deferred rows): two kernels in the staged search.
__device__ __noinline__ void pend_stage_b1(const Env &E, Mate &m) {
    int nd0 = 0, nd1 = 0;
    rows_short_round(E, m, m.g->pend[0], m.nPend[0], m.g->pend[1], m.nPend[1], m.g->pend[0], nd0, m.g->pend[1], nd1);
    m.nPend[0] = nd0;
    m.nPend[1] = nd1;
}
__device__ __noinline__ void pend_stage_b2(const Env &E, Mate &m) {
    rows_long_batch(E, m, m.g->pend[0], m.nPend[0], m.g->pend[1], m.nPend[1]);
}
__device__ void pend_stage_b(const Env &E, Mate &m) {
    pend_stage_b1(E, m);
    pend_stage_b2(E, m);
}
__device__ __noinline__ void pend_stage_c(const Env &E, Mate &m) {
    const int B = max(m.Best, m.BestHSP) - 8;
    for (int i = 0; i < m.HSPCount; ++i) {
        if ((int)m.g->hsp_score[i] < B) continue;
        align_hsp(E, m, i);
    }
    m.Mapq = calc_mapq6(m);
}
__device__ void search_pe_pending(const Env &E, Mate &m) {
    if (pend_stage_a(E, m)) return;
    pend_stage_b(E, m);
    pend_stage_c(E, m);
}

// State1::ScanSlots, scanslots.cpp:7-62.  Lanes hash 32 window positions at a time; matches against
// the mate's slots at QPos = 0,27,54,81 are then visited in window order.
__device__ __noinline__ void scan_slots(const Env &E, Mate &m, uint32_t DBLo, uint32_t DBSegLength, bool Plus) {
    const uint32_t W = E.ix.word_len;
    if (m.QL <= W * 4) return;
    const int s = Plus ? 0 : 1;
    uint64_t qslot[SCANK];
    uint32_t qpos[SCANK];
    for (uint32_t k = 0; k < SCANK; ++k) {
        qpos[k] = (k * PRIME_STRIDE) % m.QWC;
        qslot[k] = m_slot(E, m, s, qpos[k]);  // ~0 for invalid words: never equals a real slot
    }
    const uint8_t *T = E.ix.seq + DBLo;
    if (DBSegLength < W) return;
    const uint32_t nwords = DBSegLength - W + 1;
    // The window's k-mers come from the 2-bit packed genome (two 64-bit loads and a funnel shift instead of W byte loads
    // and letter tests) unless the window touches a byte that is not exactly ACGT (coarse exception bitmap): then the
    // bytes decide, letter by letter, as g_CharToLetterNucleo does (scanslots.cpp:27-45).
    bool bytes = true;   // some byte of the window is not exactly ACGT (or the bitmap is switched off)
    if (!(E.P.flags & 32u)) {
        bytes = false;
        for (uint32_t cb = DBLo >> kCoarseShift; cb <= (DBLo + DBSegLength - 1) >> kCoarseShift; ++cb)
            bytes = bytes || ((__ldg(E.ix.seqc + (cb >> 5)) >> (cb & 31)) & 1u);
    }
    const uint32_t wshift = 64 - 2 * W;
    for (uint32_t p0 = 0; p0 < nwords; p0 += 32) {
        uint32_t p = p0 + URMB_LANE;   // window word start
        uint32_t hitmask = 0;
        if (p < nwords) {
            uint64_t word = 0;
            uint32_t bad = 0;
            if (bytes) {
                for (uint32_t t = 0; t < W; ++t) {
                    uint32_t l = letter_of(__ldg(T + p + t));
                    bad |= l & 0x80u;
                    word = (word << 2) | (l & 3u);
                }
            } else {
                const uint32_t g = DBLo + p, off = 2 * (g & 31u);
                const uint64_t *pw = E.ix.seq2 + (g >> 5);
                const uint64_t a = __ldg(pw), b = __ldg(pw + 1);
                const uint64_t hi = off ? ((a << off) | (b >> (64 - off))) : a;
                word = hi >> wshift;
            }
            if (!bad) {
                uint64_t slot = mod_slots(murmur64(word & E.ix.shift_mask), E.ix.slot_count, E.ix.magic);
                for (uint32_t k = 0; k < SCANK; ++k)
                    if (slot == qslot[k]) hitmask |= 1u << k;
            }
        }
        uint32_t any = __ballot_sync(FULL, hitmask != 0);
        if (!any) continue;
        // The pure halves of this round's ExtendScan calls, one match per lane (a tandem-repeat window matches at every
        // period: the uniform form computed one extension per call on all 32 lanes).  The penalty bound only falls while
        // the window is scanned, so the bound of now is valid for every call of the round.
        uint32_t xs[SCANK];
#pragma unroll
        for (uint32_t k = 0; k < SCANK; ++k) {
            xs[k] = EXT_NONE;
            const bool mine = (hitmask >> k) & 1u;
            if (__any_sync(FULL, mine) && mine) xs[k] = pure_ext(E.ix, E.P, m.rv, Plus, qpos[k], DBLo + p, false, m.MaxPenalty);
        }
        while (any) {
            int bit = __ffs(any) - 1;
            any &= any - 1;
            uint32_t hm = __shfl_sync(FULL, hitmask, bit);
#pragma unroll
            for (uint32_t k = 0; k < SCANK; ++k) {
                const uint32_t x = __shfl_sync(FULL, xs[k], bit);
                if (hm >> k & 1u) extend_scan_apply(E, m, qpos[k], DBLo + p0 + bit, Plus, x);
            }
        }
    }
}

// State1::Scan, scan.cpp:14-39, in three pieces so that the full-window Viterbi can run elsewhere:
//   scan_mate_pre  : ScanSlots under the raised penalty bound; true when the DP has to run (no hit found, DoVit)
//   (the DP)       : viterbi_full(strand of the mate, window), a pure function of its arguments
//   scan_mate_post : the hit of a good enough DP (path trimmed as TrimLeftIs / TrimRightIs do)
__device__ __noinline__ bool scan_mate_pre(const Env &E, Mate &m, uint32_t DBPos, uint32_t DBSegLength, bool Plus, bool DoVit) {
    const int SavedMaxPenalty = m.MaxPenalty;
    const int SavedHitCount = m.HitCount;
    m.MaxPenalty = 130;
    scan_slots(E, m, DBPos, DBSegLength, Plus);
    m.MaxPenalty = SavedMaxPenalty;
    if (m.HitCount > SavedHitCount) return false;
    return DoVit;
}
__device__ __noinline__ void scan_mate_post(const Env &E, Mate &m, uint32_t DBPos, bool Plus, float Score, const uint16_t *rev,
                                            int nrev, int ovf) {
    if ((double)Score >= (double)m.QL / 3.0) {
        uint16_t *path = E.ws->runs_p;
        int np = 0;
        uint32_t LeftICount = 0;
        int last = nrev - 1;
        if (nrev > 0 && (rev[last] & 3u) == 2u) { LeftICount = rev[last] >> 2; --last; }
        int first = 0;
        // TrimRightIs on the already left-trimmed path: index 0 is never removed
        if (last >= 0 && (rev[0] & 3u) == 2u) first = (last == 0) ? 0 : 1;
        uint32_t keep1 = (last == 0 && (rev[0] & 3u) == 2u) ? 1 : 0;
        if (keep1) runs_append(path, np, 2, 1, 3 * kRunCap, ovf, URMB_LANE);
        else for (int k = last; k >= first; --k) runs_append(path, np, rev[k] & 3u, rev[k] >> 2, 3 * kRunCap, ovf, URMB_LANE);
        if (ovf) m.overflow |= 16;
        add_hit(E, m, DBPos + LeftICount, Plus, (int)Score, path, np);
    }
}
__device__ __noinline__ void scan_mate(const Env &E, Mate &m, uint32_t DBPos, uint32_t DBSegLength, bool Plus,
                                       bool DoVit) {
    if (!scan_mate_pre(E, m, DBPos, DBSegLength, Plus, DoVit)) return;
    int nrev = 0, ovf = 0;
    float Score = viterbi_full(E, mate_seq(m, Plus), m.QL, E.ix.seq + DBPos, DBSegLength, true, true, nrev, ovf);
    scan_mate_post(E, m, DBPos, Plus, Score, E.ws->runs_a, nrev, ovf);
}

struct PairState {
    int BestPairScore, SecondBestPairScore;
    int BestF, BestR;   // hit indexes of the best pair (-1: none)
    int SecF, SecR;     // hit indexes of m_SecondPairIndex (-1: UINT_MAX)
    int PairCount;
};

// State2::FindPairs, state2.cpp:20-85 (only what AdjustTopHitsAndMapqs consumes is kept)
__device__ __noinline__ void find_pairs(const Env &E, const Mate &F, const Mate &R, PairState &ps) {
    const int QL2 = (int)((F.QL + R.QL) / 2);
    ps.BestPairScore = -1;
    ps.SecondBestPairScore = -1;
    ps.BestF = ps.BestR = -1;
    ps.SecF = ps.SecR = -1;
    ps.PairCount = 0;
    for (int hf = 0; hf < F.HitCount; ++hf) {
        const int ScoreF = F.g->hit_score[hf];
        if (ScoreF < F.Second - 12) continue;
        const int64_t PosF = F.g->hit_pos[hf];
        const int PlusF = F.g->hit_plus[hf];
        for (int base = 0; base < R.HitCount; base += 32) {
            int hr = base + URMB_LANE;
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
                if (Total > ps.BestPairScore) {   // state2.cpp:61-67
                    ps.SecF = ps.BestF;
                    ps.SecR = ps.BestR;
                    ps.SecondBestPairScore = ps.BestPairScore;
                    ps.BestPairScore = Total;
                    ps.BestF = hf;
                    ps.BestR = base + bit;
                } else if (Total == ps.BestPairScore) {   // :68-72
                    ps.SecF = hf;
                    ps.SecR = base + bit;
                    ps.SecondBestPairScore = ps.BestPairScore;
                } else if (Total > ps.SecondBestPairScore) {   // :73-77: the second INDEX becomes the best pair's
                    ps.SecF = ps.BestF;
                    ps.SecR = ps.BestR;
                    ps.SecondBestPairScore = Total;
                }
                ++ps.PairCount;
            }
        }
    }
}

// State2::ScanPair, state2.cpp:87-137 (quirk 6: both window extensions use the forward mate's length)
__device__ __noinline__ void scan_pair(const Env &E, Mate &F, Mate &R) {
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

// State2::ScanPair as a resumable walk (see RescueHdr).  h holds the position; returns true when both loops are through,
// false when the walk stopped at a full-window DP whose request is then in h (INLINE: the DP runs here, never stops).
template <bool INLINE>
__device__ __noinline__ bool scan_pair_resume(const Env &E, Mate &F, Mate &R, RescuePos &h) {
    const uint32_t QLx = F.QL;
    for (; h.loop < 2; ++h.loop, h.h = 0) {
        Mate &src = h.loop ? R : F;
        Mate &dst = h.loop ? F : R;
        const int n = h.loop ? h.hcR : h.hcF;
        const bool DoVit = (h.loop ? h.dovR : h.dovF) != 0;
        while (h.h < n) {
            const int k = h.h++;
            if ((int)src.g->hit_score[k] < src.Second) continue;
            const uint32_t DBPos = src.g->hit_pos[k];
            uint32_t pos, len;
            bool plus;
            if (src.g->hit_plus[k]) { pos = DBPos; len = kScanSeg; plus = false; }
            else if (DBPos >= (uint32_t)kScanSeg) { pos = DBPos - kScanSeg; len = kScanSeg + 2 * QLx; plus = true; }
            else continue;
            if (!scan_mate_pre(E, dst, pos, len, plus, DoVit)) continue;
            if (INLINE) {
                int nrev = 0, ovf = 0;
                const float Score = viterbi_full(E, mate_seq(dst, plus), dst.QL, E.ix.seq + pos, len, true, true, nrev, ovf);
                scan_mate_post(E, dst, pos, plus, Score, E.ws->runs_a, nrev, ovf);
            } else {
                h.state = 1;
                h.dp_pos = pos;
                h.dp_len = len;
                h.dp_plus = plus ? 1 : 0;
                h.dp_mate = h.loop ? 0 : 1;
                return false;
            }
        }
    }
    return true;
}

// State2::AdjustTopHitsAndMapqs, search2.cpp:8-57
__device__ __noinline__ void adjust_pair(Mate &F, Mate &R, const PairState &ps) {
    if (ps.PairCount == 0) {
        F.Mapq /= 2;
        R.Mapq /= 2;
        return;
    }
    double Fract = (double)ps.BestPairScore / (double)(F.QL + R.QL);
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

// ExtendPen of seed i of mate m on strand Plus: through the memo when Plus is the seed's own strand, otherwise
// (search2m4.cpp:94-95,122 extend a stored seed on the strand dictated by the OTHER mate's seed) computed here.
// m_SecondHit of both mates (search2.cpp:49-56) for -tabbedout; the array is zero-filled before every launch, so pairs
// that never reach AdjustTopHitsAndMapqs keep "no second hit".  A pair that does reach it without a second pair is written
// as such: the big-capacity rerun searches a pair again whose first search -- over truncated lists -- may have left one
// (seen once in 10 M pairs: a tandem-repeat pair, profiles/r07i).
__device__ __forceinline__ void write_second(const Env &E, const Mate &F, const Mate &R, const PairState &ps, urmb_second *out,
                                             uint32_t u, uint32_t n_units) {
    if (!out || URMB_LANE != 0) return;
    urmb_second a, b;
    a.db_pos = b.db_pos = 0;
    a.score = b.score = 0;
    a.flags = b.flags = 0;
    a.pad = b.pad = 0;
    if (ps.PairCount != 0 && ps.SecF >= 0) {
        a.db_pos = F.g->hit_pos[ps.SecF];
        a.score = F.g->hit_score[ps.SecF];
        a.flags = (uint8_t)(2u | (F.g->hit_plus[ps.SecF] ? 1u : 0u));
        b.db_pos = R.g->hit_pos[ps.SecR];
        b.score = R.g->hit_score[ps.SecR];
        b.flags = (uint8_t)(2u | (R.g->hit_plus[ps.SecR] ? 1u : 0u));
    }
    out[u] = a;
    out[n_units + u] = b;
}

__device__ int apply_seed_on(const Env &E, Mate &m, int i, bool Plus) {
    const uint32_t qs = m.sd_qs[i];
    if (((qs >> 15) == 0) == Plus) return apply_seed(E, m, i);
    return extend_pen(E, m, qs & 0x7FFFu, m.sd_db[i], Plus);
}

// State2::ExtendBoth1Pair4/5, search2m4.cpp:189-208, search2m5.cpp:134-156
__device__ bool extend_both1_pair(const Env &E, Mate &F, Mate &R, int fi, int ri, bool Plusf, int TermPairScore,
                                  int &FwdScore) {
    FwdScore = apply_seed_on(E, F, fi, Plusf);
    if (FwdScore <= 0) return false;
    int RevScore = apply_seed_on(E, R, ri, !Plusf);
    if (RevScore <= 0) return false;
    if (FwdScore + RevScore < TermPairScore) return false;
    F.Mapq = 40;
    R.Mapq = 40;
    return true;
}

// ExtendPen every recorded seed in order (search2m4.cpp:139-153): only seeds that can still do something are visited.
__device__ __noinline__ void extend_all_seeds(const Env &E, Mate &m) {
    for (int w0 = 0; w0 < m.nSeeds; w0 += 32) {
        uint32_t live = ~m.sd_dead[w0 >> 5];
        while (live) {
            const int bit = __ffs(live) - 1;
            apply_seed(E, m, w0 + bit);
            live = ~m.sd_dead[w0 >> 5] & ((bit == 31) ? 0u : (0xFFFFFFFFu << (bit + 1)));
        }
    }
}

// State2::Search4 / Search5, search2m4.cpp:15-187, search2m5.cpp:9-132.
// FAST: only the part every pair goes through (seed pairing, extension of all BOTH1 seeds, the 90 % rule); returns
// false when the pair needs SearchPE_Pending / pair finding / mate rescue, which the second-pass kernel then runs
// from scratch on the compacted list of such pairs (the first part is cheap to redo and nothing has to be saved).
template <bool FAST>
__device__ __noinline__ bool search_pair(const Env &E, Mate &F, Mate &R, PairState *ps_out = nullptr) {
    reset_search(E, F);
    reset_search(E, R);
    const uint32_t W = E.ix.word_len;
    if (F.QL < W || R.QL < W) { F.Mapq = R.Mapq = 0; return true; }
    const int QLf = (int)F.QL, QLr = (int)R.QL, QL2 = (QLf + QLr) / 2;
    const int TermPairScore = QLf + QLr + 5 * E.P.MM;
    build_seeds_pe(E, F);
    build_seeds_pe(E, R);
    // Iteration t of the reference's do-while records F seed t then R seed t (while they last):
    //   (a) F seed t against R seeds [0, min(t, NR)): ExtendPen(F t) comes first in every ExtendBoth1Pair4 call, so
    //       once it is known to return <= 0 without side effects the remaining partners are no-ops;
    //   (b) R seed t against F seeds [0, min(t+1, NF)): partners whose own-strand memo is dead are no-ops.
    const int NF = F.nSeeds, NR = R.nSeeds;
#pragma unroll 1
    for (int t = 0; t < max(NF, NR); ++t) {
        if (t < NF && !seed_dead(F, t)) {
            const int nR = min(t, NR);
            const uint32_t dbf = F.sd_db[t];
            const bool Plusf = (F.sd_qs[t] >> 15) == 0;
            bool gone = false;
            for (int base = 0; base < nR && !gone; base += 32) {
                const int i = base + URMB_LANE;
                int64_t d = (int64_t)dbf - (int64_t)((i < nR) ? R.sd_db[i] : 0u);
                if (d < 0) d = -d;
                uint32_t bal = __ballot_sync(FULL, (i < nR) && (d + QL2 <= MAX_TL));
                while (bal) {
                    const int bit = __ffs(bal) - 1;
                    bal &= bal - 1;
                    int fs;
                    if (extend_both1_pair(E, F, R, t, base + bit, Plusf, TermPairScore, fs)) return true;
                    if (fs <= 0) { gone = true; break; }
                }
            }
        }
        if (t < NR) {
            const int nF = min(t + 1, NF);
            const uint32_t dbr = R.sd_db[t];
            const bool Plusr = (R.sd_qs[t] >> 15) == 0;
            for (int base = 0; base < nF; base += 32) {
                const int i = base + URMB_LANE;
                bool ok = false;
                if (i < nF) {
                    int64_t d = (int64_t)F.sd_db[i] - (int64_t)dbr;
                    if (d < 0) d = -d;
                    const bool own = ((F.sd_qs[i] >> 15) == 0) == !Plusr;
                    ok = (d + QL2 <= MAX_TL) && !(own && seed_dead(F, i));
                }
                uint32_t bal = __ballot_sync(FULL, ok);
                while (bal) {
                    const int bit = __ffs(bal) - 1;
                    bal &= bal - 1;
                    int fs;
                    if (extend_both1_pair(E, F, R, base + bit, t, !Plusr, TermPairScore, fs)) return true;
                }
            }
        }
    }

    extend_all_seeds(E, F);
    extend_all_seeds(E, R);

    if (E.P.pe_method == 5) {
        if (FAST) return false;
        search_pe_pending(E, F);
        search_pe_pending(E, R);
        return true;
    }
    const int TermF = (QLf * 9) / 10, TermR = (QLr * 9) / 10;
    if (F.Best >= TermF && R.Best >= TermR) {
        int64_t d = (int64_t)F.g->hit_pos[F.Top] - (int64_t)R.g->hit_pos[R.Top];
        if (d < 0) d = -d;
        if (d + QL2 <= MAX_TL) {
            F.Mapq = 40;
            R.Mapq = 40;
            return true;
        }
    }
    if (FAST) return false;
    search_pe_pending(E, F);
    search_pe_pending(E, R);
    PairState ps;
    find_pairs(E, F, R, ps);
    if (ps.PairCount == 0) {
        scan_pair(E, F, R);
        find_pairs(E, F, R, ps);
    }
    adjust_pair(F, R, ps);
    if (ps_out) *ps_out = ps;
    return true;
}

// ---- result write-back -----------------------------------------------------------------
__device__ __noinline__ void write_result(const Env &E, const Mate &m, const DevOut &o, uint32_t r) {
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
            if (URMB_LANE == 0) off = atomicAdd(&o.counters[CT_RUNS], n);
            off = __shfl_sync(FULL, off, 0);
            if (off + n <= o.runs_cap) {
                const uint16_t *src = m.g->runs_pool + m.g->hit_roff[t];
                for (uint32_t k = URMB_LANE; k < n; k += 32) o.runs[off + k] = src[k];
                res.path_off = off;
                res.path_runs = (uint16_t)n;
            } else {
                res.flags |= 0x80;
            }
        }
    }
#ifndef URMB_BIG
    if ((E.P.flags & 256u) && r % 5 == 0) res.flags |= 0x80;   // test hook (URMB_FLAGS bit 8): every fifth read takes the in-stream big-capacity rerun
#endif
    if (URMB_LANE == 0) {
        o.res[r] = res;
#ifdef URMB_BIG
        atomicAdd(&o.counters[CT_DBG_HSPS + (m.HSPCount <= 256 ? 0 : m.HSPCount <= 512 ? 1 : m.HSPCount <= 1024 ? 2 : 3)], 1u);
        atomicMax(&o.counters[CT_DBG_MAXHSP], (uint32_t)m.HSPCount);
#endif
        if (res.flags & 0x80) {
            atomicAdd(&o.counters[CT_OVERFLOW], 1u);
#ifndef URMB_BIG
            if (o.ovf_list) {
                const uint32_t k = atomicAdd(&o.counters[CT_OVF_LIST], 1u);
                if (k < o.ovf_cap) o.ovf_list[k] = r;
            }
#endif
            for (int k = 0; k < 5; ++k)   // which capacity: hits, runs of a path, run pool, HSPs, path assembly (statistics)
                if (m.overflow >> k & 1) atomicAdd(&o.counters[CT_DBG_OVF + k], 1u);
        }
    }
}

// Per-mate shared-memory footprint: read view (+ the seed lists of the first pass).
__host__ __device__ inline size_t mate_smem_bytes(uint32_t qcap, uint32_t seqcap, bool seeds) {
    size_t n = 2 * (size_t)seqcap + kReadViewBytes;   // bytes fwd+rc, packed strands, bad bits
    //            sd_db + sd_ext          sd_qs              sd_dead
    if (seeds) n += 2 * (size_t)qcap * 8 + 2 * (size_t)qcap * 2 + ((2 * (size_t)qcap + 31) / 32) * 4;
    return (n + 15) & ~(size_t)15;
}

// What a kernel keeps per warp in shared memory: nm mate areas | flank-DP window | flank-DP trace bits.
struct SmemPlan {
    uint32_t nm;       // mate areas
    uint32_t seeds;    // mate areas carry seed lists
    uint32_t dp;       // flank-DP window + trace bits
};
__host__ __device__ inline size_t smem_per_warp(const DevBatch &b, const SmemPlan &pl) {
    size_t s = (size_t)pl.nm * mate_smem_bytes(b.qcap, b.seqcap, pl.seeds != 0);
    if (pl.dp) {
        const uint32_t rows = b.seqcap + 2;
        s += b.seqcap + 64;                      // genome window
        s += (size_t)rows * 32 + rows + 66;      // band trace bytes (2 nibbles per lane per row) + column LB + row LA
    }
    return (s + 15) & ~(size_t)15;
}

// Stage one read: bytes, reverse complement, packed strands in shared memory; probe results stay in global.
__device__ __noinline__ void load_mate(const Env &E, Mate &m, const DevBatch &b, const DevProbe &pr, uint32_t r, uint8_t *sm,
                                       MateScratch *g, bool seeds) {
    const uint32_t off = __ldg(b.offs + r), L = __ldg(b.offs + r + 1) - off;
    uint64_t *s_pk = reinterpret_cast<uint64_t *>(sm);
    uint32_t *s_bad = reinterpret_cast<uint32_t *>(sm + 2 * kPkWords * 8);
    uint8_t *p8 = sm + kReadViewBytes;
    m.sd_db = m.sd_ext = m.sd_dead = nullptr;
    m.sd_qs = nullptr;
    if (seeds) {
        m.sd_db = reinterpret_cast<uint32_t *>(p8);
        p8 += 2 * (size_t)b.qcap * 4;
        m.sd_ext = reinterpret_cast<uint32_t *>(p8);
        p8 += 2 * (size_t)b.qcap * 4;
        m.sd_dead = reinterpret_cast<uint32_t *>(p8);
        p8 += ((2 * (size_t)b.qcap + 31) / 32) * 4;
        m.sd_qs = reinterpret_cast<uint16_t *>(p8);
        p8 += 2 * (size_t)b.qcap * 2;
    }
    uint8_t *s_q = p8, *s_rc = p8 + b.seqcap;
    {   // the probe kernel's staged read: packed strands + bad bits + flags + reverse complement; forward bytes as given.
        // Every global load is issued before the first shared store (the stores would otherwise fence the loads of the
        // next piece behind them: three dependent round trips to HBM instead of one after the offsets).
        static_assert(kReadViewBytes / 4 <= 64 && kMaxLen / 4 <= 64, "two rounds of 32 lanes per piece");
        const uint32_t *vw = reinterpret_cast<const uint32_t *>(pr.view + (size_t)r * pr.view_stride);
        const uint32_t lane = (uint32_t)URMB_LANE, nrc = b.seqcap / 4;
        const uint32_t v0 = __ldg(vw + lane);
        const uint32_t v1 = (lane + 32 < kReadViewBytes / 4) ? __ldg(vw + lane + 32) : 0u;
        const uint32_t fl = __ldg(vw + kReadViewBytes / 4);
        const uint32_t c0 = (lane < nrc) ? __ldg(vw + kViewHdr / 4 + lane) : 0u;
        const uint32_t c1 = (lane + 32 < nrc) ? __ldg(vw + kViewHdr / 4 + lane + 32) : 0u;
        // forward bytes: aligned words covering [off, off + L) (the batch buffer is padded), bytes placed below
        const uint32_t sk = off & 3u, nws = (sk + L + 3) >> 2;   // <= 65 words
        const uint32_t *sw32 = reinterpret_cast<const uint32_t *>(b.seqs + (off - sk));
        uint32_t w[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) w[k] = (lane + 32 * k < nws) ? __ldg(sw32 + lane + 32 * k) : 0u;
        uint32_t *dst32 = reinterpret_cast<uint32_t *>(sm);
        dst32[lane] = v0;
        if (lane + 32 < kReadViewBytes / 4) dst32[lane + 32] = v1;
        uint32_t *rc32 = reinterpret_cast<uint32_t *>(s_rc);
        if (lane < nrc) rc32[lane] = c0;
        if (lane + 32 < nrc) rc32[lane + 32] = c1;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t p0 = 4 * (lane + 32 * k);   // byte position of the word in the aligned stream
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const uint32_t p = p0 + t - sk;        // position in the read (wraps below zero for the skipped bytes)
                if (p < L) s_q[p] = (uint8_t)(w[k] >> (8 * t));
            }
        }
        m.rv.q = s_q;
        m.rv.rc = s_rc;
        m.rv.pk = s_pk;
        m.rv.bad = s_bad;
        m.rv.QL = L;
        m.rv.slow = (fl & 1u) != 0;
        m.rv.hasbad = (fl & 2u) != 0;
        __syncwarp();
    }
    const size_t base = (size_t)r * 2 * b.qcap;
    m.q = s_q;
    m.rc = s_rc;
    m.tally = pr.tally + base;
    m.pos = pr.pos + base;
    m.ext = pr.ext + base;
    m.nSeeds = 0;
    m.g = g;
    m.QL = L;
    m.QWC = (L >= E.ix.word_len) ? L - E.ix.word_len + 1 : 0;
    m.qcap = b.qcap;
    m.overflow = 0;
}

// ---- saved mate state (paired-end second pass) -----------------------------------------------------------
// First pass -> pool: the used part of the per-warp scratch and the scalars.  SearchPE_Pending's entry test
// (search1pepend.cpp:27-31) is evaluated here so that finished mates never enter a stage kernel.
template <bool PE>
__device__ __noinline__ void save_mate(const Env &E, Mate &m, MateSave *dst) {
    bool done = false;
    if (PE) {
        done = pend_done_at_entry(E, m);
        m.MaxPenalty = E.P.MAXPEN;   // search1pepend.cpp:15
        if (done) m.Mapq = calc_mapq6(m);
    }
    // scratch -> pool never overlap: with __restrict__ the loads of a whole round are issued before its stores
    const MateScratch *__restrict__ g = m.g;
    MateScratch *__restrict__ d = &dst->s;
#pragma unroll 1
    for (int i0 = 0; i0 < max(m.HitCount, m.HSPCount); i0 += 32) {
        const int i = i0 + URMB_LANE;
        const bool a = i < m.HitCount, c = i < m.HSPCount;
        uint32_t hp = 0, hd = 0;
        int16_t hs = 0, xs = 0;
        uint16_t hr = 0, xq = 0, xl = 0;
        uint8_t hl = 0, hn = 0, xf = 0;
        if (a) { hp = g->hit_pos[i]; hs = g->hit_score[i]; hl = g->hit_plus[i]; hn = g->hit_nruns[i]; hr = g->hit_roff[i]; }
        if (c) { hd = g->hsp_dbstart[i]; xq = g->hsp_qstart[i]; xl = g->hsp_len[i]; xs = g->hsp_score[i]; xf = g->hsp_flags[i]; }
        if (a) { d->hit_pos[i] = hp; d->hit_score[i] = hs; d->hit_plus[i] = hl; d->hit_nruns[i] = hn; d->hit_roff[i] = hr; }
        if (c) { d->hsp_dbstart[i] = hd; d->hsp_qstart[i] = xq; d->hsp_len[i] = xl; d->hsp_score[i] = xs; d->hsp_flags[i] = xf; }
    }
#pragma unroll 1
    for (int i = URMB_LANE; i < m.nRuns; i += 32) d->runs_pool[i] = g->runs_pool[i];
    if (!done)
#pragma unroll 1
        for (int i0 = 0; i0 < max(m.nPend[0], m.nPend[1]); i0 += 32) {
            const int i = i0 + URMB_LANE;
            uint8_t p0 = 0, p1 = 0;
            if (i < m.nPend[0]) p0 = g->pend[0][i];
            if (i < m.nPend[1]) p1 = g->pend[1][i];
            if (i < m.nPend[0]) d->pend[0][i] = p0;
            if (i < m.nPend[1]) d->pend[1][i] = p1;
        }
    if (URMB_LANE == 0) {
        MateHdr h;
        h.HitCount = m.HitCount; h.HSPCount = m.HSPCount; h.Top = m.Top; h.MaxPenalty = m.MaxPenalty;
        h.Best = m.Best; h.Second = m.Second; h.BestHSP = m.BestHSP; h.nRuns = m.nRuns;
        h.Mapq = m.Mapq; h.nPend[0] = m.nPend[0]; h.nPend[1] = m.nPend[1];
        h.overflow = m.overflow; h.done = done ? 1 : 0;
        h.pad[0] = h.pad[1] = h.pad[2] = 0;
        dst->h = h;
    }
    __syncwarp();
}

__device__ __forceinline__ void hdr_to_mate(const MateHdr &h, Mate &m) {
    m.HitCount = h.HitCount; m.HSPCount = h.HSPCount; m.Top = h.Top; m.MaxPenalty = h.MaxPenalty;
    m.Best = h.Best; m.Second = h.Second; m.BestHSP = h.BestHSP; m.nRuns = h.nRuns;
    m.Mapq = h.Mapq; m.nPend[0] = h.nPend[0]; m.nPend[1] = h.nPend[1];
    m.overflow = h.overflow;
}
__device__ __forceinline__ void mate_to_hdr(const Env &E, const Mate &m, MateSave *sv, bool done) {
    __syncwarp();
    if (URMB_LANE == 0) {
        MateHdr h;
        h.HitCount = m.HitCount; h.HSPCount = m.HSPCount; h.Top = m.Top; h.MaxPenalty = m.MaxPenalty;
        h.Best = m.Best; h.Second = m.Second; h.BestHSP = m.BestHSP; h.nRuns = m.nRuns;
        h.Mapq = m.Mapq; h.nPend[0] = m.nPend[0]; h.nPend[1] = m.nPend[1];
        h.overflow = m.overflow; h.done = done ? 1 : 0;
        h.pad[0] = h.pad[1] = h.pad[2] = 0;
        sv->h = h;
    }
}

// A mate of the second pass without its read (pair finishing only touches the hit list).
__device__ __forceinline__ void bare_mate(const Env &E, Mate &m, const DevBatch &b, uint32_t r, MateSave *sv) {
    const uint32_t L = b.offs[r + 1] - b.offs[r];
    m.q = m.rc = nullptr;
    m.tally = nullptr; m.pos = nullptr; m.ext = nullptr;
    m.sd_db = m.sd_ext = m.sd_dead = nullptr;
    m.sd_qs = nullptr;
    m.nSeeds = 0;
    m.g = &sv->s;
    m.QL = L;
    m.QWC = (L >= E.ix.word_len) ? L - E.ix.word_len + 1 : 0;
    m.qcap = b.qcap;
    hdr_to_mate(sv->h, m);
}

// Saved mate -> rescue pool: hit list with its path runs, HSP list, scalars (the pending lists are spent by then).
__device__ __noinline__ void copy_save(const Env &E, const MateSave *__restrict__ src, MateSave *__restrict__ dst) {
    const MateHdr h = src->h;
    const MateScratch *__restrict__ g = &src->s;
    MateScratch *__restrict__ d = &dst->s;
#pragma unroll 1
    for (int i0 = 0; i0 < max(h.HitCount, h.HSPCount); i0 += 32) {
        const int i = i0 + URMB_LANE;
        if (i < h.HitCount) {
            d->hit_pos[i] = g->hit_pos[i]; d->hit_score[i] = g->hit_score[i]; d->hit_plus[i] = g->hit_plus[i];
            d->hit_nruns[i] = g->hit_nruns[i]; d->hit_roff[i] = g->hit_roff[i];
        }
        if (i < h.HSPCount) {
            d->hsp_dbstart[i] = g->hsp_dbstart[i]; d->hsp_qstart[i] = g->hsp_qstart[i]; d->hsp_len[i] = g->hsp_len[i];
            d->hsp_score[i] = g->hsp_score[i]; d->hsp_flags[i] = g->hsp_flags[i];
        }
    }
#pragma unroll 1
    for (int i = URMB_LANE; i < h.nRuns; i += 32) d->runs_pool[i] = g->runs_pool[i];
    if (URMB_LANE == 0) dst->h = h;
    __syncwarp();
}

__device__ __forceinline__ void make_env(Env &E, const DevIndex &ix, const DevParams &P, const DevBatch &b,
                                         WarpScratch *scratch, const SmemPlan &pl, uint8_t *sw, int gw, int lane) {
    E.ix = ix;
    E.P = P;
    E.ws = scratch + gw;
    uint8_t *p8 = sw + (size_t)pl.nm * mate_smem_bytes(b.qcap, b.seqcap, pl.seeds != 0);
    E.s_win = p8;
    E.s_tb = p8 + b.seqcap + 64;
    E.tb_stride = 4 * P.R + 6;
    E.tb_rows = pl.dp ? b.seqcap + 2 : 0;
    E.lane_ = lane;
}

struct KArgs {   // one parameter block for every search kernel
    DevIndex ix;
    DevParams P;
    DevBatch b;
    DevProbe pr;
    DevOut o;
    WarpScratch *scratch;
    MateSave *pool;
    uint32_t unit_base, unit_count;   // first-pass kernels: the chunk of units this launch covers
    uint32_t spw;                     // shared bytes per warp
};

// MODE 3 (single-end, complete search in one kernel: the big-capacity rerun of urmb_big.cu, where a handful of reads must
// not cost a dozen dependent launches).
// MODE 0 (single-end, first pass): every read of the chunk, phases 1-2; reads that need more are saved to the pool.
// MODE 1 (paired, first pass): every pair of the chunk, the part
// every pair goes through; pairs that need more are saved to the pool.  MODE 2 (paired, mate rescue): complete search
// of the pairs listed in o.rescue.
template <int MODE>
__device__ __forceinline__ void search_body(const KArgs &A) {
    URMB_DYN_SMEM(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int gw = blockIdx.x * wpb + warp;
    uint8_t *sw = smem + (size_t)warp * A.spw;
    const DevBatch &b = A.b;
    const DevOut &o = A.o;
    const SmemPlan pl{b.paired ? 2u : 1u, 1u, MODE >= 2 ? 1u : 0u};
    const size_t msz = mate_smem_bytes(b.qcap, b.seqcap, true);
    Env E;
    make_env(E, A.ix, A.P, b, A.scratch, pl, sw, gw, lane);
    // units from a list: the pairs finish_kernel left to the legacy mate rescue (o.rescue), or -- big-capacity build -- the
    // units of the current rerun pass (o.ovf_units from CT_OVF_BASE on)
    const bool listed = (MODE == 2) || (MODE == 3 && A.unit_count == 0xFFFFFFFFu);
#ifdef URMB_BIG
    const uint32_t lbase = o.counters[CT_OVF_BASE];
    const uint32_t n_work = listed ? o.counters[CT_OVF_UNITS] - lbase : A.unit_count;
    const uint32_t *ulist = o.ovf_units + lbase;
    constexpr int kListHead = CT_OVF_HEAD;
#else
    const uint32_t n_work = listed ? o.counters[CT_RESCUE_LEGACY] : A.unit_count;
    const uint32_t *ulist = o.rescue;
    constexpr int kListHead = CT_RESCUE_HEAD;
#endif

    for (;;) {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(&o.counters[listed ? kListHead : CT_HEAD], 1u);
        u = __shfl_sync(FULL, u, 0);
        if (u >= n_work) break;
        u = listed ? ulist[u] : A.unit_base + u;
        if (MODE == 1 && A.pr.done && A.pr.done[u]) continue;   // finished by the probe kernel's first look
        if (MODE == 3) {   // single-end, the whole of Search_Lo in one kernel (search1m6.cpp:35-277)
            Mate m;
            load_mate(E, m, b, A.pr, u, sw, &E.ws->m[0], true);
            reset_search(E, m);
            if (!se_phase12(E, m) && !se_phase3(E, m) && !se_phase45(E, m)) se_phase6(E, m);
            write_result(E, m, o, u);
        } else if (MODE == 0) {
            Mate m;
            load_mate(E, m, b, A.pr, u, sw, &E.ws->m[0], true);
            reset_search(E, m);   // State1::Search, search1.cpp:7-24
            if (se_phase12(E, m)) write_result(E, m, o, u);
            else {
                uint32_t t = 0;
                if (lane == 0) {
                    t = atomicAdd(&o.counters[CT_TODO], 1u);
                    o.todo[t] = u;
                }
                t = __shfl_sync(FULL, t, 0);
                save_mate<false>(E, m, A.pool + t);
            }
        } else {
            Mate F, R;
            load_mate(E, F, b, A.pr, u, sw, &E.ws->m[0], true);
            load_mate(E, R, b, A.pr, b.n_units + u, sw + msz, &E.ws->m[1], true);
            PairState ps;
            ps.PairCount = 0;
            ps.SecF = ps.SecR = -1;
            const bool done = search_pair<MODE == 1>(E, F, R, MODE == 2 ? &ps : nullptr);
            if (done) {
                if (MODE == 2) write_second(E, F, R, ps, o.second, u, b.n_units);
                write_result(E, F, o, u);
                write_result(E, R, o, b.n_units + u);
            } else {   // MODE 1 only
                uint32_t t = 0;
                if (lane == 0) {
                    t = atomicAdd(&o.counters[CT_TODO], 1u);
                    o.todo[t] = u;
                }
                t = __shfl_sync(FULL, t, 0);
                save_mate<true>(E, F, A.pool + 2 * (size_t)t);
                save_mate<true>(E, R, A.pool + 2 * (size_t)t + 1);
            }
        }
        __syncwarp();
    }
}

// Second pass, one warp per saved MATE.  STAGE 0: SearchPE_Pending up to its first HSP alignments
// (search1pepend.cpp:15-51); STAGE 1: the pending rows (:53-110); STAGE 2: the final HSP alignments and MAPQ (:112-129).
// STAGE 3-5: the same for the single-end search (one saved state per read): phase 3, phases 4-5, phase 6 + result.
// STAGE 6: paired-end pending round 2 (the deferred long rows), STAGE 1 then only does round 1.
// STAGE 7: single-end phase 5 (the deferred long rows), STAGE 4 then only does phase 4.
template <int STAGE>
__device__ __forceinline__ void stage_body(const KArgs &A) {
    constexpr bool SE = (STAGE >= 3 && STAGE <= 5) || STAGE == 7;
    URMB_DYN_SMEM(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int gw = blockIdx.x * wpb + warp;
    uint8_t *sw = smem + (size_t)warp * A.spw;
    const DevBatch &b = A.b;
    const SmemPlan pl{1u, 0u, (STAGE == 1 || STAGE == 4 || STAGE == 6 || STAGE == 7) ? 0u : 1u};
    Env E;
    make_env(E, A.ix, A.P, b, A.scratch, pl, sw, gw, lane);
    const uint32_t n_work = (SE ? 1u : 2u) * A.o.counters[CT_TODO];
    for (;;) {
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(&A.o.counters[(STAGE == 6 || STAGE == 7) ? CT_STAGE_B2 : CT_STAGE_A + (SE ? STAGE - 3 : STAGE)], 1u);
        k = __shfl_sync(FULL, k, 0);
        if (k >= n_work) break;
        MateSave *sv = A.pool + k;
        const MateHdr h = sv->h;
        const uint32_t u = A.o.todo[SE ? k : k >> 1];
        const uint32_t r = (!SE && (k & 1u)) ? b.n_units + u : u;
        if (h.done && STAGE != 5) continue;
        if (STAGE == 6 && h.nPend[0] + h.nPend[1] == 0) continue;   // no deferred rows
        if (STAGE == 7 && h.nPend[0] + h.nPend[1] == 0 &&
            !se_phase5_done(A.P, (int)(b.offs[r + 1] - b.offs[r]), h.Best)) continue;   // phase 5 would do nothing
        if (STAGE == 0 || STAGE == 3) {   // nothing to align (save_mate already reset the paired-end penalty bound)
            const int QL = (int)(b.offs[r + 1] - b.offs[r]);
            if (STAGE == 0 && h.BestHSP < (QL * A.P.TERM3_PCT) / 100) continue;
            if (STAGE == 3 && !se_phase3_needed(A.P, QL, h.BestHSP)) continue;
        }
        Mate m;
        if (STAGE == 5 && h.done) {   // finished in an earlier stage: only the result record is left
            bare_mate(E, m, b, r, sv);
            write_result(E, m, A.o, u);
            continue;
        }
        load_mate(E, m, b, A.pr, r, sw, &sv->s, false);
        hdr_to_mate(h, m);
        bool done = false;
        if (STAGE == 0) done = pend_stage_a(E, m);
        else if (STAGE == 1) pend_stage_b1(E, m);
        else if (STAGE == 6) pend_stage_b2(E, m);
        else if (STAGE == 2) { pend_stage_c(E, m); done = true; }
        else if (STAGE == 3) done = se_phase3(E, m);
        else if (STAGE == 4) done = se_phase4(E, m);
        else if (STAGE == 7) done = se_phase5(E, m);
        else { se_phase6(E, m); done = true; }
        if (STAGE == 5) write_result(E, m, A.o, u);
        else mate_to_hdr(E, m, sv, done);
        __syncwarp();
    }
}

// Second pass, one warp per saved PAIR: State2::FindPairs + AdjustTopHitsAndMapqs (search2m4.cpp:177-186) and the
// result records; pairs without any pair of hits go to the mate-rescue kernel.
__device__ __forceinline__ void finish_body(const KArgs &A) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int gw = blockIdx.x * wpb + warp;
    const DevBatch &b = A.b;
    Env E;
    E.ix = A.ix;
    E.P = A.P;
    E.ws = A.scratch + gw;
    E.s_win = E.s_tb = nullptr;
    E.tb_stride = E.tb_rows = 0;
    E.lane_ = lane;
    const uint32_t n_work = A.o.counters[CT_TODO];
    if (gw == 0 && lane == 0) atomicAdd(&A.o.counters[CT_TODO_TOTAL], n_work);
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(&A.o.counters[CT_FINISH], 1u);
        t = __shfl_sync(FULL, t, 0);
        if (t >= n_work) break;
        const uint32_t u = A.o.todo[t];
        Mate F, R;
        bare_mate(E, F, b, u, A.pool + 2 * (size_t)t);
        bare_mate(E, R, b, b.n_units + u, A.pool + 2 * (size_t)t + 1);
        if (A.P.pe_method != 5) {
            PairState ps;
            find_pairs(E, F, R, ps);
            if (ps.PairCount == 0) {   // State2::ScanPair needed (search2m4.cpp:179-183)
                // the pair's saved states move to the rescue pool (this chunk's pool is about to be reused); when that
                // is full the pair is searched again from scratch by the legacy kernel
                uint32_t e = 0;
                if (lane == 0) e = atomicAdd(&A.o.counters[CT_RESCUE], 1u);
                e = __shfl_sync(FULL, e, 0);
                if (A.o.rpool && e < A.o.rescue_cap) {
                    RescueSave *rs = A.o.rpool + e;
                    copy_save(E, A.pool + 2 * (size_t)t, &rs->m[0]);
                    copy_save(E, A.pool + 2 * (size_t)t + 1, &rs->m[1]);
                    if (lane == 0) {
                        rs->h.unit = u;
                        rs->h.p.state = 0;
                    }
                } else if (lane == 0) {
                    A.o.rescue[atomicAdd(&A.o.counters[CT_RESCUE_LEGACY], 1u)] = u;
                }
                __syncwarp();
                continue;
            }
            adjust_pair(F, R, ps);
            write_second(E, F, R, ps, A.o.second, u, b.n_units);
        }
        write_result(E, F, A.o, u);
        write_result(E, R, A.o, b.n_units + u);
        __syncwarp();
    }
}

// Mate-rescue round `round`, one warp per pair of the round's list: State2::ScanPair continued from the saved states in the
// rescue pool (search2m4.cpp:179-186).  A pair that reaches a full-window DP is appended to the next list and stops
// (INLINE, the last round: the DP runs here); a pair that gets through both loops is finished: FindPairs,
// AdjustTopHitsAndMapqs, result records.
template <bool INLINE>
__device__ __forceinline__ void rescue_scan_body(const KArgs &A, int round) {
    URMB_DYN_SMEM(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int gw = blockIdx.x * wpb + warp;
    uint8_t *sw = smem + (size_t)warp * A.spw;
    const DevBatch &b = A.b;
    const DevOut &o = A.o;
    const SmemPlan pl{2u, 0u, 1u};
    const size_t msz = mate_smem_bytes(b.qcap, b.seqcap, false);
    Env E;
    make_env(E, A.ix, A.P, b, A.scratch, pl, sw, gw, lane);
    const uint32_t n_work = round == 0 ? min(o.counters[CT_RESCUE], o.rescue_cap) : o.counters[CT_RQ_COUNT + round];
    const uint32_t *list = o.rq[round & 1];
    uint32_t *next = o.rq[(round + 1) & 1];
    for (;;) {
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(&o.counters[CT_RQ_SCAN + round], 1u);
        k = __shfl_sync(FULL, k, 0);
        if (k >= n_work) break;
        const uint32_t e = round == 0 ? k : list[k];
        RescueSave *rs = o.rpool + e;
        RescuePos h = rs->h.p;
        const uint32_t u = rs->h.unit;
        Mate F, R;
        load_mate(E, F, b, A.pr, u, sw, &rs->m[0].s, false);
        load_mate(E, R, b, A.pr, b.n_units + u, sw + msz, &rs->m[1].s, false);
        hdr_to_mate(rs->m[0].h, F);
        hdr_to_mate(rs->m[1].h, R);
#ifndef URMB_EMU
        const long long t_in = clock64();
#endif
        int nwin = 0;
        if (h.state == 0) {   // State2::ScanPair entry, state2.cpp:87-98
            h.loop = 0;
            h.h = 0;
            h.hcF = F.HitCount;
            h.hcR = R.HitCount;
            h.dovF = ((int)F.Mapq >= 10) ? 1 : 0;
            h.dovR = ((int)R.Mapq >= 10) ? 1 : 0;
            // scan windows this pair starts with (statistics; the second loop's bound can still rise)
            for (int k = lane; k < max(F.HitCount, R.HitCount); k += 32)
                nwin += (k < F.HitCount && (int)F.g->hit_score[k] >= F.Second) + (k < R.HitCount && (int)R.g->hit_score[k] >= R.Second);
            for (int d = 16; d; d >>= 1) nwin += __shfl_xor_sync(FULL, nwin, d);
            if (lane == 0) atomicAdd(&o.counters[CT_DBG_WIN + (nwin <= 4 ? 0 : nwin <= 16 ? 1 : nwin <= 64 ? 2 : nwin <= 256 ? 3 : 4)], 1u);
        } else {              // the DP this pair stopped at has been run
            Mate &dst = h.dp_mate ? R : F;
            scan_mate_post(E, dst, h.dp_pos, h.dp_plus != 0, rs->h.dp_score, rs->h.dp_runs, rs->h.dp_nrev, rs->h.dp_ovf);
        }
        h.state = 0;
        if (scan_pair_resume<INLINE>(E, F, R, h)) {
            PairState ps;
            find_pairs(E, F, R, ps);
            adjust_pair(F, R, ps);
            write_second(E, F, R, ps, o.second, u, b.n_units);
            write_result(E, F, o, u);
            write_result(E, R, o, b.n_units + u);
        } else {
            mate_to_hdr(E, F, &rs->m[0], false);
            mate_to_hdr(E, R, &rs->m[1], false);
            if (lane == 0) {
                rs->h.p = h;
                next[atomicAdd(&o.counters[CT_RQ_COUNT + round + 1], 1u)] = e;
            }
        }
#ifndef URMB_EMU
        if (lane == 0) {
            const uint32_t dt = (uint32_t)((clock64() - t_in) >> 10);
            if (atomicMax(&o.counters[CT_DBG_MAXT], dt) < dt) o.counters[CT_DBG_MAXW] = (uint32_t)nwin | ((uint32_t)round << 24);
        }
#endif
        __syncwarp();
    }
}

// The full-window DPs of the pairs round `round` stopped at (list round + 1), one warp per DP: State1::Viterbi over the
// whole scan window (scan.cpp:27), result (score, reversed RLE path) left in the pair's RescueHdr.
__device__ __forceinline__ void rescue_dp_body(const KArgs &A, int round) {
    URMB_DYN_SMEM(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int gw = blockIdx.x * wpb + warp;
    uint8_t *sw = smem + (size_t)warp * A.spw;
    const DevBatch &b = A.b;
    const DevOut &o = A.o;
    const SmemPlan pl{1u, 0u, 1u};   // the flank-DP trace area holds the window
    Env E;
    make_env(E, A.ix, A.P, b, A.scratch, pl, sw, gw, lane);
    const uint32_t n_work = o.counters[CT_RQ_COUNT + round + 1];
    const uint32_t *list = o.rq[(round + 1) & 1];
    if (gw == 0 && lane == 0) atomicAdd(&o.counters[CT_RESCUE_DPS], n_work);
    for (;;) {
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(&o.counters[CT_RQ_DP + round + 1], 1u);
        k = __shfl_sync(FULL, k, 0);
        if (k >= n_work) break;
        RescueSave *rs = o.rpool + list[k];
        const uint32_t u = rs->h.unit;
        const uint32_t pos = rs->h.p.dp_pos, len = rs->h.p.dp_len;
        const bool plus = rs->h.p.dp_plus != 0;
        const int mate = rs->h.p.dp_mate;
        Mate m;
        load_mate(E, m, b, A.pr, mate ? b.n_units + u : u, sw, &rs->m[mate].s, false);
        int nrev = 0, ovf = 0;
        const float Score = viterbi_full(E, mate_seq(m, plus), m.QL, E.ix.seq + pos, len, true, true, nrev, ovf);
        __syncwarp();
        for (int i = lane; i < nrev; i += 32) rs->h.dp_runs[i] = E.ws->runs_a[i];
        if (lane == 0) {
            rs->h.dp_score = Score;
            rs->h.dp_nrev = nrev;
            rs->h.dp_ovf = ovf;
        }
        __syncwarp();
    }
}

// Resident blocks per SM each kernel is compiled for (register budget = 65536 / (128 * blocks)).
#ifndef URMB_LB_PAIR
#define URMB_LB_PAIR 5
#endif
#ifndef URMB_LB_ROWS
#define URMB_LB_ROWS 7
#endif
#ifndef URMB_LB_ALIGN
#define URMB_LB_ALIGN 6
#endif
__global__ void __launch_bounds__(128, URMB_LB_PAIR) seed_kernel_se(const __grid_constant__ KArgs A) { search_body<0>(A); }
__global__ void __launch_bounds__(128, URMB_LB_ALIGN) align_kernel_se3(const __grid_constant__ KArgs A) { stage_body<3>(A); }
__global__ void __launch_bounds__(128, URMB_LB_ROWS) rows_kernel_se(const __grid_constant__ KArgs A) { stage_body<4>(A); }
__global__ void __launch_bounds__(128, URMB_LB_ROWS) rows_long_kernel_se(const __grid_constant__ KArgs A) { stage_body<7>(A); }
__global__ void __launch_bounds__(128, URMB_LB_ALIGN) align_kernel_se6(const __grid_constant__ KArgs A) { stage_body<5>(A); }
__global__ void __launch_bounds__(128, URMB_LB_PAIR) pair_kernel(const __grid_constant__ KArgs A) { search_body<1>(A); }
__global__ void __launch_bounds__(128, 4) rescue_kernel(const __grid_constant__ KArgs A) { search_body<2>(A); }
__global__ void __launch_bounds__(128, 4) se_full_kernel(const __grid_constant__ KArgs A) { search_body<3>(A); }
__global__ void identity_list_kernel(uint32_t *list, uint32_t n, uint32_t *counters) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) list[i] = i;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counters[CT_OVF_UNITS] = n;
        counters[CT_OVF_BASE] = 0;
        counters[CT_OVF_HEAD] = 0;
        counters[CT_RESCUE_LEGACY] = n;   // (the fast build's listed mode)
    }
}
__global__ void __launch_bounds__(128, URMB_LB_ALIGN) align_kernel_a(const __grid_constant__ KArgs A) { stage_body<0>(A); }
__global__ void __launch_bounds__(128, URMB_LB_ROWS) rows_kernel(const __grid_constant__ KArgs A) { stage_body<1>(A); }
__global__ void __launch_bounds__(128, URMB_LB_ROWS) rows_long_kernel(const __grid_constant__ KArgs A) { stage_body<6>(A); }
__global__ void __launch_bounds__(128, URMB_LB_ALIGN) align_kernel_c(const __grid_constant__ KArgs A) { stage_body<2>(A); }
__global__ void __launch_bounds__(128, 4) finish_kernel(const __grid_constant__ KArgs A) { finish_body(A); }
__global__ void __launch_bounds__(128, 5) rescue_scan_kernel(const __grid_constant__ KArgs A, int round) { rescue_scan_body<false>(A, round); }
__global__ void __launch_bounds__(128, 4) rescue_last_kernel(const __grid_constant__ KArgs A, int round) { rescue_scan_body<true>(A, round); }
__global__ void __launch_bounds__(128, 6) rescue_dp_kernel(const __grid_constant__ KArgs A, int round) { rescue_dp_body(A, round); }

// =====================================================================================
// host-side launchers
// =====================================================================================
int max_search_warps(int sm_count) { return sm_count * 32; }

size_t packed_words(size_t n_bytes) { return n_bytes / 32 + 2; }
size_t coarse_words(size_t n_bytes) { return ((packed_words(n_bytes) * 32) >> kCoarseShift) / 32 + 2; }

int launch_pack_genome(const uint8_t *seq, size_t n_bytes, uint64_t *seq2, uint32_t *seqx, uint32_t *seqc, void *stream) {
    const size_t n_words = packed_words(n_bytes);
    size_t blocks = (n_words + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    URMB_LAUNCH(pack_genome_kernel, (int)blocks, 256, 0, stream, seq, n_bytes, n_words, seq2, seqx, seqc);
    return (int)cudaGetLastError();
}

int launch_probe(const DevIndex &ix, const DevParams &P, const DevBatch &b, const DevProbe &pr, void *stream, int sm_count,
                 const DevOut *o) {
    const int threads = 256;
    if (b.paired && pr.done && o && !(P.flags & (256u | 512u))) {   // one warp per pair, with the first look
        const size_t smem = (size_t)(threads / 32) * probe_pair_smem_per_warp(b.qcap, b.seqcap);
        cudaFuncSetAttribute(probe_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        int blocks = (int)((b.n_units + 7) / 8);
        if (blocks > sm_count * 16) blocks = sm_count * 16;
        if (blocks < 1) blocks = 1;
        URMB_LAUNCH(probe_pair_kernel, blocks, threads, smem, stream, ix, P, b, pr, o->res, o->counters);
        return (int)cudaGetLastError();
    }
    DevProbe q = pr;
    q.done = nullptr;
    if (b.paired && pr.done) {   // no first look: no pair is finished here
#ifndef URMB_EMU
        const cudaError_t me = cudaMemsetAsync(pr.done, 0, b.n_units, (cudaStream_t)stream);
        if (me != cudaSuccess) return (int)me;
#else
        memset(pr.done, 0, b.n_units);
#endif
    }
    const size_t smem = (size_t)(threads / 32) * probe_smem_per_warp(b.qcap, b.seqcap);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int blocks = (int)((b.n_reads + 7) / 8);
    if (blocks > sm_count * 16) blocks = sm_count * 16;
    if (blocks < 1) blocks = 1;
    URMB_LAUNCH(probe_kernel, blocks, threads, smem, stream, ix, P, b, q);
    return (int)cudaGetLastError();
}

// Persistent grid: SMs x resident blocks (bounded by the per-warp scratch), work claimed by atomicAdd.
template <class K, class... X>
static int launch_one(K kern, int klass, const LaunchTrace *tr, KArgs &A, const SmemPlan &pl, uint32_t max_items,
                      const SearchRes &R, void *stream, int sm_count, int *warps_used, X... extra) {
    const int wpb = 4;
    A.spw = (uint32_t)smem_per_warp(A.b, pl);
    const size_t smem = (size_t)A.spw * wpb;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, smem);
    if (per_sm < 1) per_sm = 1;
    int blocks = sm_count * per_sm;
    if (blocks * wpb > R.n_scratch_warps) blocks = R.n_scratch_warps / wpb;
    const int need = (int)((max_items + wpb - 1) / wpb);
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    if (warps_used) *warps_used = blocks * wpb;
    if (tr) tr->mark(tr->user, klass, 0);
    URMB_LAUNCH(kern, blocks, wpb * 32, smem, stream, A, extra...);
    if (tr) tr->mark(tr->user, klass, 1);
    return (int)cudaGetLastError();
}

static KArgs make_kargs(const DevIndex &ix, const DevParams &P, const DevBatch &b, const DevProbe &pr, const DevOut &o,
                        const SearchRes &R) {
    KArgs A;
    A.ix = ix; A.P = P; A.b = b; A.pr = pr; A.o = o;
    A.scratch = R.scratch;
    A.pool = R.pool;
    A.unit_base = 0;
    A.unit_count = b.n_units;
    A.spw = 0;
    return A;
}

// Returns the number of kernels launched or a negative cudaError.
//   single-end: per chunk  seed_kernel_se -> align_kernel_se3 -> rows_kernel_se -> rows_long_kernel_se -> align_kernel_se6.
//   paired-end: per chunk of R.pool_pairs pairs  pair_kernel -> align_kernel_a -> rows_kernel -> align_kernel_c ->
//               finish_kernel (each a small kernel over the saved mate states); launch_rescue does the rest.
int launch_search(const DevIndex &ix, const DevParams &P, const DevBatch &b, const DevProbe &pr, const DevOut &o,
                  const SearchRes &R, void *stream, int sm_count, int *warps_used, const LaunchTrace *tr) {
    KArgs A = make_kargs(ix, P, b, pr, o, R);
    int e;
#define URMB_TRY(call) do { e = (call); if (e) return -e; } while (0)
    if (!R.pool || R.pool_pairs == 0) return -1;   // cudaErrorInvalidValue
    int n = 0;
    if (!b.paired) {
        const uint32_t chunk = 2 * R.pool_pairs;   // one saved state per read
        for (uint32_t u0 = 0; u0 < b.n_units; u0 += chunk) {
            const uint32_t cnt = (b.n_units - u0 < chunk) ? b.n_units - u0 : chunk;
            A.unit_base = u0;
            A.unit_count = cnt;
#ifndef URMB_EMU
            URMB_TRY((int)cudaMemsetAsync(o.counters + CT_CHUNK0, 0, (CT_CHUNK_END - CT_CHUNK0) * sizeof(uint32_t), (cudaStream_t)stream));
#else
            for (int i = CT_CHUNK0; i < CT_CHUNK_END; ++i) o.counters[i] = 0;
#endif
            URMB_TRY(launch_one(seed_kernel_se, 1, tr, A, SmemPlan{1, 1, 0}, cnt, R, stream, sm_count, warps_used));
            URMB_TRY(launch_one(align_kernel_se3, 2, tr, A, SmemPlan{1, 0, 1}, cnt, R, stream, sm_count, nullptr));
            URMB_TRY(launch_one(rows_kernel_se, 3, tr, A, SmemPlan{1, 0, 0}, cnt, R, stream, sm_count, nullptr));
            URMB_TRY(launch_one(rows_long_kernel_se, 7, tr, A, SmemPlan{1, 0, 0}, cnt, R, stream, sm_count, nullptr));
            URMB_TRY(launch_one(align_kernel_se6, 4, tr, A, SmemPlan{1, 0, 1}, cnt, R, stream, sm_count, nullptr));
            n += 5;
        }
        return n;
    }
    for (uint32_t u0 = 0; u0 < b.n_units; u0 += R.pool_pairs) {
        const uint32_t cnt = (b.n_units - u0 < R.pool_pairs) ? b.n_units - u0 : R.pool_pairs;
        A.unit_base = u0;
        A.unit_count = cnt;
#ifndef URMB_EMU
        URMB_TRY((int)cudaMemsetAsync(o.counters + CT_CHUNK0, 0, (CT_CHUNK_END - CT_CHUNK0) * sizeof(uint32_t), (cudaStream_t)stream));
#else
        for (int i = CT_CHUNK0; i < CT_CHUNK_END; ++i) o.counters[i] = 0;
#endif
        URMB_TRY(launch_one(pair_kernel, 1, tr, A, SmemPlan{2, 1, 0}, cnt, R, stream, sm_count, warps_used));
        URMB_TRY(launch_one(align_kernel_a, 2, tr, A, SmemPlan{1, 0, 1}, 2 * cnt, R, stream, sm_count, nullptr));
        URMB_TRY(launch_one(rows_kernel, 3, tr, A, SmemPlan{1, 0, 0}, 2 * cnt, R, stream, sm_count, nullptr));
        URMB_TRY(launch_one(rows_long_kernel, 7, tr, A, SmemPlan{1, 0, 0}, 2 * cnt, R, stream, sm_count, nullptr));
        URMB_TRY(launch_one(align_kernel_c, 4, tr, A, SmemPlan{1, 0, 1}, 2 * cnt, R, stream, sm_count, nullptr));
        URMB_TRY(launch_one(finish_kernel, 5, tr, A, SmemPlan{0, 0, 0}, cnt, R, stream, sm_count, nullptr));
        n += 6;
    }
    return n;
}

// Every unit of the batch searched from scratch by ONE kernel (after the probe kernel): se_full_kernel, or for pairs the
// legacy mate-rescue kernel over a list of all pairs.  Two launches instead of a dozen: the big-capacity rerun of a few reads.
int launch_search_monolithic(const DevIndex &ix, const DevParams &P, const DevBatch &b, const DevProbe &pr, const DevOut &o,
                             const SearchRes &R, void *stream, int sm_count) {
    KArgs A = make_kargs(ix, P, b, pr, o, R);
    int e;
#define URMB_TRY(call) do { e = (call); if (e) return -e; } while (0)
    if (!b.paired) {
        URMB_TRY(launch_one(se_full_kernel, 1, nullptr, A, SmemPlan{1, 1, 1}, b.n_units, R, stream, sm_count, nullptr));
        return 1;
    }
    URMB_LAUNCH(identity_list_kernel, 1, 256, 0, stream, o.ovf_units ? o.ovf_units : o.rescue, b.n_units, o.counters);
    URMB_TRY((int)cudaGetLastError());
    URMB_TRY(launch_one(rescue_kernel, 9, nullptr, A, SmemPlan{2, 1, 1}, b.n_units, R, stream, sm_count, nullptr));
    return 2;
#undef URMB_TRY
}

// The reads recorded in o.ovf_list (over a per-mate capacity of the fast build) as a duplicate-free list of units in
// o.rescue, ready for a listed launch of rescue_kernel / se_full_kernel; the overflow counter starts again from zero (the
// rerun counts what is still over).  One block; the list is short.
__global__ void __launch_bounds__(256) overflow_list_kernel(DevOut o, uint32_t n_units, int paired) {
    // entries [start, n) of ovf_list are new since the last pass: their units (the smaller read index of a pair stands
    // for it; both mates of a pair are always written by the same kernel, hence listed in the same pass) become the
    // units [CT_OVF_BASE, CT_OVF_UNITS) of this pass; the reads leave the overflow count (the rerun counts what stays over)
    const uint32_t start = o.counters[CT_OVF_DONE];
    const uint32_t n = min(o.counters[CT_OVF_LIST], o.ovf_cap);
    const uint32_t base = o.counters[CT_OVF_UNITS];
    __shared__ uint32_t cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    for (uint32_t i = start + threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t r = o.ovf_list[i], u = (paired && r >= n_units) ? r - n_units : r;
        bool first = true;
        for (uint32_t j = start; j < n && first; ++j) {
            const uint32_t rj = o.ovf_list[j], uj = (paired && rj >= n_units) ? rj - n_units : rj;
            if (uj == u && (rj < r)) first = false;
        }
        if (first) o.ovf_units[base + atomicAdd(&cnt, 1u)] = u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        o.counters[CT_OVF_BASE] = base;
        o.counters[CT_OVF_UNITS] = base + cnt;
        o.counters[CT_OVF_HEAD] = 0;
        o.counters[CT_OVF_DONE] = n;
        if (n > start) atomicSub(&o.counters[CT_OVERFLOW], n - start);
    }
}

// The listed units searched again from scratch by one kernel (big-capacity build: urmb_big.cu calls this after the mate
// rescue of the batch, on the same stream).  Results, path runs and second hits are written in place.
int launch_overflow_rerun(const DevIndex &ix, const DevParams &P, const DevBatch &b, const DevProbe &pr, const DevOut &o,
                          const SearchRes &R, void *stream, int sm_count) {
    if (!o.ovf_list || !o.ovf_units || !o.ovf_cap) return 0;
    KArgs A = make_kargs(ix, P, b, pr, o, R);
    A.unit_count = 0xFFFFFFFFu;   // listed launch
    int e;
#define URMB_TRY(call) do { e = (call); if (e) return -e; } while (0)
    URMB_LAUNCH(overflow_list_kernel, 1, 256, 0, stream, o, b.n_units, b.paired);
    URMB_TRY((int)cudaGetLastError());
    const uint32_t items = b.n_units < o.ovf_cap ? b.n_units : o.ovf_cap;
    if (b.paired) URMB_TRY(launch_one(rescue_kernel, 11, nullptr, A, SmemPlan{2, 1, 1}, items, R, stream, sm_count, nullptr));
    else URMB_TRY(launch_one(se_full_kernel, 11, nullptr, A, SmemPlan{1, 1, 1}, items, R, stream, sm_count, nullptr));
    return 2;
#undef URMB_TRY
}

// Mate rescue of the pairs finish_kernel found without a pair of hits: kRescueRounds rounds of
// (rescue_scan_kernel, rescue_dp_kernel) over the rescue pool, a last round that finishes the stragglers in place, and the
// legacy kernel (complete search from scratch) for the pairs that did not fit into the pool.
int launch_rescue(const DevIndex &ix, const DevParams &P, const DevBatch &b, const DevProbe &pr, const DevOut &o,
                  const SearchRes &R, void *stream, int sm_count, const LaunchTrace *tr) {
#define URMB_TRY(call) do { e = (call); if (e) return -e; } while (0)
    if (!b.paired || P.pe_method == 5) return 0;
    KArgs A = make_kargs(ix, P, b, pr, o, R);
    int e, n = 0;
    if (o.rpool && o.rescue_cap) {
        const uint32_t cap = b.n_units < o.rescue_cap ? b.n_units : o.rescue_cap;
        const int rounds = P.rescue_rounds < 0 ? 0 : (P.rescue_rounds > kRescueRounds ? (int)kRescueRounds : P.rescue_rounds);
        for (int r = 0; r < rounds; ++r) {
            URMB_TRY(launch_one(rescue_scan_kernel, 6, tr, A, SmemPlan{2, 0, 1}, cap, R, stream, sm_count, nullptr, r));
            URMB_TRY(launch_one(rescue_dp_kernel, 8, tr, A, SmemPlan{1, 0, 1}, cap, R, stream, sm_count, nullptr, r));
        }
        URMB_TRY(launch_one(rescue_last_kernel, 10, tr, A, SmemPlan{2, 0, 1}, cap, R, stream, sm_count, nullptr, rounds));
        n += 2 * rounds + 1;
    }
    URMB_TRY(launch_one(rescue_kernel, 9, tr, A, SmemPlan{2, 1, 1}, o.rpool ? (b.n_units < (uint32_t)(4 * sm_count) ? b.n_units : (uint32_t)(4 * sm_count)) : b.n_units,
                        R, stream, sm_count, nullptr));
    return n + 1;
#undef URMB_TRY
}

}  // namespace URMB_NS
