// urmb_big.cu -- the search kernels compiled a second time, in namespace urmb_big, with per-mate capacities that no read
// of up to URMB_MAX_READ_LEN bases can exceed (urmb_internal.h).  urmb_wait maps the few reads that overflowed the
// capacities of the fast build again through these kernels, so that no result depends on a capacity: the reference's hit
// and HSP lists grow without bound (state1.cpp:190, AllocHits / AllocHSPs).  Same source, same semantics; only the
// array sizes of MateScratch differ.  The plain-C entry points below are what urmb_api.cu calls (the structures it
// passes are the cap-independent ones, identical in both namespaces).
#define URMB_BIG 1
#define URMB_NS urmb_big
#include "urmb_kernels.cu"

extern "C" size_t urmb_big_scratch_bytes() { return sizeof(urmb_big::WarpScratch); }
extern "C" size_t urmb_big_save_bytes() { return sizeof(urmb_big::MateSave); }

// Probe kernel + one kernel that searches every unit from scratch (launch_search_monolithic) on `stream`.
// ix / P / batch / probe / out point to the namespace-urmb structures of the same names; scratch holds n_scratch_warps
// x urmb_big_scratch_bytes(), pool 2 x pool_pairs x urmb_big_save_bytes().  Returns kernels launched or a negative cudaError.
extern "C" int urmb_big_map(const void *ix_, const void *P_, const void *batch_, const void *probe_, const void *out_,
                            void *scratch, int n_scratch_warps, void *pool, uint32_t pool_pairs, void *stream, int sm_count) {
    using namespace urmb_big;
    const DevIndex &ix = *reinterpret_cast<const DevIndex *>(ix_);
    const DevParams &P = *reinterpret_cast<const DevParams *>(P_);
    const DevBatch &b = *reinterpret_cast<const DevBatch *>(batch_);
    const DevProbe &pr = *reinterpret_cast<const DevProbe *>(probe_);
    DevOut o = *reinterpret_cast<const DevOut *>(out_);
    o.rpool = nullptr;
    o.rescue_cap = 0;
    o.rq[0] = o.rq[1] = nullptr;
    SearchRes R{reinterpret_cast<WarpScratch *>(scratch), n_scratch_warps, reinterpret_cast<MateSave *>(pool), pool_pairs};
    int e = launch_probe(ix, P, b, pr, stream, sm_count);
    if (e) return -e;
    const int n = launch_search_monolithic(ix, P, b, pr, o, R, stream, sm_count);
    if (n < 0) return n;
    return 1 + n;
}

// The reads the fast build recorded in out->ovf_list searched again, in place, by the big-capacity kernels: queued on
// `stream` behind the batch's mate rescue, no host round trip.  scratch: n_scratch_warps x urmb_big_scratch_bytes().
extern "C" int urmb_big_rerun_listed(const void *ix_, const void *P_, const void *batch_, const void *probe_, const void *out_,
                                     void *scratch, int n_scratch_warps, void *stream, int sm_count) {
    using namespace urmb_big;
    const DevIndex &ix = *reinterpret_cast<const DevIndex *>(ix_);
    const DevParams &P = *reinterpret_cast<const DevParams *>(P_);
    const DevBatch &b = *reinterpret_cast<const DevBatch *>(batch_);
    const DevProbe &pr = *reinterpret_cast<const DevProbe *>(probe_);
    DevOut o = *reinterpret_cast<const DevOut *>(out_);
    o.rpool = nullptr;
    o.rescue_cap = 0;
    o.rq[0] = o.rq[1] = nullptr;
    SearchRes R{reinterpret_cast<WarpScratch *>(scratch), n_scratch_warps, nullptr, 0};
    return launch_overflow_rerun(ix, P, b, pr, o, R, stream, sm_count);
}
