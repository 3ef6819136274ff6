// lcb_traverse.cuh -- warp-per-seed carving-path traversal for sm_100a.
//
// One warp evaluates ProcessVertex::Process (SibeliaZ-LCB/blocksfinder.h:228-310) for one seed:
// Path::Init (path.h:33-46), the forward/backward extension loops driven by MostPopularVertex
// (blocksfinder.h:708-768) and Path::PointPushBack/Front (path.h:430-602), Path::Score (path.h:604-628).
//
// Design (B200-first, not a translation):
//  * the reference's `used` bit per edge becomes a 32-bit EPOCH per edge: the index of the seed that
//    claimed it (0xFFFFFFFF = free).  A traversal sees an edge as used iff epoch < its threshold
//    (phase start for the speculative evaluation, its own seed index for a commit-time re-run), so
//    thousands of seeds of different phases run concurrently against ONE array with no snapshots;
//  * every epoch the traversal's outcome depends on is recorded as a read-set of index intervals, so a
//    later round can re-validate the result instead of recomputing it (lcb_device.cu);
//  * per-seed state (path vertex->distance hash, instance table in multiset order, vote table, best
//    snapshot) lives in shared memory and spills to a per-warp HBM scratch arena only for big paths;
//  * lanes are spent on the memory-latency-bound parts: look-ahead walks (one lane per depth), occurrence
//    lists (one lane per occurrence), `used` scans, ordered searches via ballot; control flow is warp-uniform.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lcb {

constexpr uint32_t kFree = 0xFFFFFFFFu;
constexpr int kNotSet = 0x7FFFFFFF;
constexpr unsigned kFull = 0xFFFFFFFFu;

// ---- capacities -----------------------------------------------------------------------------------
#ifndef LCB_INST_SMEM
#define LCB_INST_SMEM 32
#endif
#ifndef LCB_HASH_SMEM
#define LCB_HASH_SMEM 256
#endif
constexpr int kInstSmem = LCB_INST_SMEM; // instances kept in shared memory (<= 32: one lane per instance in the searches)
constexpr int kHashSmem = LCB_HASH_SMEM; // path hash slots in shared memory (half of them usable; a power of two >= 32)
static_assert(kInstSmem >= 8 && kInstSmem <= 32 && kHashSmem >= 32 && (kHashSmem & (kHashSmem - 1)) == 0, "shared-memory capacities");
constexpr int kVoteSmem = 128;     // vote table slots in shared memory
#ifndef LCB_TINY_ARENA
// per-warp arena; a seed that outgrows any of these is re-run in a big arena slot
constexpr int kInstMax = 4096;
constexpr int kHashMax = 65536;
constexpr int kPathMax = 32768;    // path vertices (half the hash slots)
constexpr int kVoteMax = 8192;
constexpr int kReadSetMax = 65536; // read-set intervals per traversal
#else
// test build (build.py --tiny-arena): ordinary fixtures outgrow the per-warp arena all the time, so the GPU suite
// exercises the big-slot re-run
constexpr int kInstMax = 64;
constexpr int kHashMax = 1024;
constexpr int kPathMax = 512;
constexpr int kVoteMax = 1024;
constexpr int kReadSetMax = 512;
#endif
// big arena slots (a few per device, taken on demand): the hard caps
constexpr int kBigInstMax = 32768;
constexpr int kBigHashMax = 1 << 20; // 20 hash bits (hash_of >> 12)
constexpr int kBigPathMax = 1 << 19;
constexpr int kBigVoteMax = 1 << 18;
constexpr int kBigReadSetMax = 1 << 20;
constexpr int kBigSlots = 64;
constexpr int kSmallInstMax = kInstMax, kSmallHashMax = kHashMax, kSmallPathMax = kPathMax, kSmallVoteMax = kVoteMax,
              kSmallReadSetMax = kReadSetMax;

struct Caps { // capacities of the arena a traversal currently runs in
    int inst, hash, path, vote, rs;
};
__host__ __device__ __forceinline__ Caps caps_of(bool big)
{
    return big ? Caps{kBigInstMax, kBigHashMax, kBigPathMax, kBigVoteMax, kBigReadSetMax}
               : Caps{kInstMax, kHashMax, kPathMax, kVoteMax, kReadSetMax};
}
struct Ctx;
__device__ __forceinline__ int cap_inst(const Ctx &c);
__device__ __forceinline__ int cap_hash(const Ctx &c);
__device__ __forceinline__ int cap_path(const Ctx &c);
__device__ __forceinline__ int cap_vote(const Ctx &c);
__device__ __forceinline__ int cap_rs(const Ctx &c);

struct Inst { // Path::Instance (path.h:53-181) with cached end-point data; 56 bytes
    int fg, bg;         // front_/back_ : global record index
    int fv, bv;         // strand-signed vertex id at front/back
    unsigned fbp, bbp;  // raw base-pair position at front/back
    int fdist, bdist;   // frontDistance_/backDistance_
    int key;            // compareIdx_ as a global index
    int rlo, rhi;       // read extent over epoch indices (rlo > rhi: empty)
    int clo, chi;       // chromosome bounds [clo, chi) of this instance
    unsigned flags;     // bit0 strand (+), bit1 frontFinished, bit2 backFinished
};
constexpr unsigned kPos = 1u, kFFin = 2u, kBFin = 4u;

struct Index { // device view of the junction index
    // rec[g] = {id, bp, first occurrence slot of |id|, (#occurrences << 16) | (next_ch << 8) | prev_rc}: one 16-byte
    // load per walk step yields the vertex, its position, where its occurrence list lives and both edge characters
    const int4 *rec;
    // occ[o] = {g | (stored id < 0 ? 1<<31 : 0), bp}: an occurrence's strand and position without touching rec[g]
    const int2 *occ;
    const uint32_t *vtx_off;
    const uint32_t *chr_off;
    int C, N, V;
};

__device__ __forceinline__ unsigned char rec_next_ch(const int4 &r) { return (unsigned char)((unsigned)r.w >> 8); }
__device__ __forceinline__ unsigned char rec_prev_rc(const int4 &r) { return (unsigned char)r.w; }
__device__ __forceinline__ unsigned rec_occ_count(const int4 &r) { return (unsigned)r.w >> 16; }

struct Params {
    int k, b, m, flank, depth;
};

struct WarpSmem { // ~5 KB per warp
    Inst inst[kInstSmem];
    int4 best[kInstSmem];
    int2 hash[kHashSmem];
    int2 vote[kVoteSmem];
    unsigned vlast[kVoteSmem];
    unsigned short ord[kInstSmem];
    unsigned short good[kInstSmem];
    // shadow copy of the path state at the best forward point (restored instead of Clear + Init + re-push)
    Inst s_inst[kInstSmem];
    int2 s_hash[kHashSmem];
    unsigned short s_ord[kInstSmem];
    unsigned short s_good[kInstSmem];
};

struct WarpArena { // HBM scratch of one traversal (per-warp arena or big slot): spill space + variable-length logs.
    // One base pointer and one flag live in registers; the arrays are base + a constant offset per arena kind.
    unsigned char *base;
    bool big;
    __device__ __forceinline__ size_t pick(size_t small_off, size_t big_off) const { return big ? big_off : small_off; }
#define LCB_AR_OFF(K)                                                                                                   \
    static constexpr size_t o_inst_##K = 0;                                                                             \
    static constexpr size_t o_best_##K = o_inst_##K + sizeof(Inst) * (size_t)k##K##InstMax;                             \
    static constexpr size_t o_hash_##K = o_best_##K + sizeof(int4) * (size_t)k##K##InstMax;                             \
    static constexpr size_t o_redge_##K = o_hash_##K + sizeof(int2) * (size_t)k##K##HashMax;                            \
    static constexpr size_t o_vote_##K = o_redge_##K + sizeof(int4) * (size_t)k##K##PathMax;                            \
    static constexpr size_t o_rs_##K = o_vote_##K + sizeof(int2) * (size_t)k##K##VoteMax;                               \
    static constexpr size_t o_vlast_##K = o_rs_##K + sizeof(int2) * (size_t)k##K##ReadSetMax;                           \
    static constexpr size_t o_hslot_##K = o_vlast_##K + sizeof(unsigned) * (size_t)k##K##VoteMax;                       \
    static constexpr size_t o_ord_##K = o_hslot_##K + sizeof(int) * (size_t)k##K##PathMax;                              \
    static constexpr size_t o_good_##K = o_ord_##K + sizeof(unsigned short) * (size_t)k##K##InstMax;                    \
    static constexpr size_t o_end_##K = o_good_##K + sizeof(unsigned short) * (size_t)k##K##InstMax;
    LCB_AR_OFF(Small)
    LCB_AR_OFF(Big)
#undef LCB_AR_OFF
    __device__ __forceinline__ Inst *inst() const { return (Inst *)(base + pick(o_inst_Small, o_inst_Big)); }            // cap_inst
    __device__ __forceinline__ int4 *best() const { return (int4 *)(base + pick(o_best_Small, o_best_Big)); }            // cap_inst
    __device__ __forceinline__ int2 *hash() const { return (int2 *)(base + pick(o_hash_Small, o_hash_Big)); }            // cap_hash
    // successful right edges {end vertex, length, source g | strand << 31, path distance}
    __device__ __forceinline__ int4 *redge() const { return (int4 *)(base + pick(o_redge_Small, o_redge_Big)); }         // cap_path
    __device__ __forceinline__ int2 *vote() const { return (int2 *)(base + pick(o_vote_Small, o_vote_Big)); }            // cap_vote
    // read-set intervals [lo, hi] over epoch indices
    __device__ __forceinline__ int2 *rs() const { return (int2 *)(base + pick(o_rs_Small, o_rs_Big)); }                  // cap_rs
    __device__ __forceinline__ unsigned *vlast() const { return (unsigned *)(base + pick(o_vlast_Small, o_vlast_Big)); } // cap_vote
    // slots occupied in the HBM hash (for O(path) clearing)
    __device__ __forceinline__ int *hslot() const { return (int *)(base + pick(o_hslot_Small, o_hslot_Big)); }           // cap_path
    __device__ __forceinline__ unsigned short *ord() const { return (unsigned short *)(base + pick(o_ord_Small, o_ord_Big)); }
    __device__ __forceinline__ unsigned short *good() const { return (unsigned short *)(base + pick(o_good_Small, o_good_Big)); }
};

struct Counters {
    unsigned long long walk, occ, scan, score;
    unsigned pushes, mpv_fast, mpv_mid, mpv_slow, push_par, push_ser; // developer diagnostics
};

struct Ctx { // warp-uniform traversal state (registers)
    Index ix;
    Params pr;
    const uint32_t *E;
    uint32_t thresh;
    int lane;
    // storage (shared or arena)
    Inst *inst;
    int4 *best;
    unsigned short *ord, *good;
    int2 *hash;
    int2 *vote;
    int icap, hmask, vcap;
    bool hbig;
    WarpSmem *sm;
    WarpArena ar;
    // path
    int origin, right_vertex, left_vertex;
    int right_flank, left_flank; // rightBodyFlank_, leftBodyFlank_
    int nright, nleft;           // successful pushes
    int ninst, ngood, nbest, hcount, nrs;
    int err; // 0 or LCB_ERR_CAPACITY | (which cap << 8): 1 path vertices, 2 read-set intervals, 3 instances, 4 vote table
    bool collect; // count walk/occurrence/scan/score steps (diagnostics; off on the timed path)
    bool vote_clean; // the shared-memory vote table is all-empty (mpv_mid leaves it so, the general path does not)
    // shadow state (WarpSmem::s_*)
    bool snap_valid, snap_hash;
    int snap_ninst, snap_ngood, snap_hcount, snap_right_flank, snap_right_vertex, snap_nright;
    Counters ct;
};

// one flag in registers instead of five capacities
__device__ __forceinline__ int cap_inst(const Ctx &c) { return c.ar.big ? kBigInstMax : kInstMax; }
__device__ __forceinline__ int cap_hash(const Ctx &c) { return c.ar.big ? kBigHashMax : kHashMax; }
__device__ __forceinline__ int cap_path(const Ctx &c) { return c.ar.big ? kBigPathMax : kPathMax; }
__device__ __forceinline__ int cap_vote(const Ctx &c) { return c.ar.big ? kBigVoteMax : kVoteMax; }
__device__ __forceinline__ int cap_rs(const Ctx &c) { return c.ar.big ? kBigReadSetMax : kReadSetMax; }

// bytes of one arena (per-warp or big slot)
__host__ __device__ __forceinline__ size_t arena_stride_of(bool big)
{
    const size_t s = big ? WarpArena::o_end_Big : WarpArena::o_end_Small;
    return (s + 255) & ~(size_t)255;
}

__device__ __forceinline__ void arena_bind(Ctx &c, unsigned char *p, bool big)
{
    c.ar.base = p;
    c.ar.big = big;
}

__device__ __forceinline__ int ffs_lane(unsigned m) { return __ffs((int)m) - 1; }

// position of the k-th (0-based) set bit of a 4-bit mask (cheap replacement of __fns for the walk scheduler)
__device__ __forceinline__ int kth_bit4(unsigned m, int k)
{
    int r = 0;
#pragma unroll
    for (int b = 0; b < 4; b++)
        if (((m >> b) & 1u) && __popc(m & ((1u << b) - 1u)) == k) r = b;
    return r;
}

__device__ __forceinline__ void prefetch_l1(const void *p)
{
#if defined(__CUDA_ARCH__) // (nothing to do where the header is compiled for the host: tests/trav_emu.cpp)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

__device__ __forceinline__ unsigned hash_of(int key) { return (unsigned)key * 2654435761u; }

// vertex -> path distance (DistanceKeeper, distancekeeper.h:9-41), open addressing, key 0 = empty
__device__ __forceinline__ int hash_find(const int2 *h, int mask, int key)
{
    unsigned s = (hash_of(key) >> 12) & (unsigned)mask;
    while (true) {
        int2 kv = h[s];
        if (kv.x == key) return kv.y;
        if (kv.x == 0) return kNotSet;
        s = (s + 1) & (unsigned)mask;
    }
}

__device__ __forceinline__ void chr_bounds(const Index &ix, int g, int &lo, int &hi)
{
    int a = 0, b = ix.C; // chr_off[a] <= g < chr_off[b]
    while (b - a > 1) {
        int mid = (a + b) >> 1;
        if ((int)__ldg(ix.chr_off + mid) <= g) a = mid;
        else b = mid;
    }
    lo = (int)__ldg(ix.chr_off + a);
    hi = (int)__ldg(ix.chr_off + a + 1);
}

__device__ __forceinline__ void ctx_reset_storage(Ctx &c)
{
    c.inst = c.sm->inst;
    c.best = c.sm->best;
    c.ord = c.sm->ord;
    c.good = c.sm->good;
    c.vote = c.sm->vote;
    c.icap = kInstSmem;
    c.vcap = kVoteSmem;
}

__device__ __forceinline__ void hash_clear(Ctx &c)
{
    if (c.hbig) { // only the slots the path occupied
        for (int i = c.lane; i < c.hcount; i += 32) c.ar.hash()[c.ar.hslot()[i]].x = 0;
    } else {
        for (int i = c.lane; i < kHashSmem; i += 32) c.hash[i] = make_int2(0, 0);
    }
    __syncwarp();
    c.hcount = 0;
}

// move the shared-memory path hash into the (all-zero) arena table; returns nothing, caller repoints
__device__ __noinline__ void hash_migrate(const int2 *src, int2 *dst, int *hslot, int lane, unsigned dmask)
{
    int n = 0;
    for (int base = 0; base < kHashSmem; base += 32) {
        int2 kv = src[base + lane];
        bool live = kv.x != 0;
        unsigned m = __ballot_sync(kFull, live);
        if (live) {
            unsigned s = (hash_of(kv.x) >> 12) & dmask;
            while (atomicCAS(&dst[s].x, 0, kv.x) != 0) s = (s + 1) & dmask;
            dst[s].y = kv.y;
            hslot[n + __popc(m & ((1u << lane) - 1))] = (int)s;
        }
        n += __popc(m);
    }
    __syncwarp();
}

// uniform insert (all lanes pass the same key); grows shared -> arena at half load
__device__ __forceinline__ void hash_insert(Ctx &c, int key, int val)
{
    if (!c.hbig && (c.hcount + 1) * 2 > kHashSmem) {
        hash_migrate(c.hash, c.ar.hash(), c.ar.hslot(), c.lane, (unsigned)(cap_hash(c) - 1));
        c.hash = c.ar.hash();
        c.hmask = cap_hash(c) - 1;
        c.hbig = true;
    }
    if (c.hbig && c.hcount + 1 > cap_path(c)) {
        c.err = LCB_ERR_CAPACITY | (1 << 8);
        return;
    }
    __syncwarp(); // every lane has finished probing the table (hash_find of the same key) before lane 0 changes it
    if (c.lane == 0) {
        unsigned s = (hash_of(key) >> 12) & (unsigned)c.hmask;
        while (c.hash[s].x != 0) s = (s + 1) & (unsigned)c.hmask;
        c.hash[s] = make_int2(key, val);
        if (c.hbig) c.ar.hslot()[c.hcount] = (int)s;
    }
    c.hcount++;
    __syncwarp();
}

__device__ __forceinline__ void rs_add(Ctx &c, int lo, int hi) // uniform
{
    if (c.nrs >= cap_rs(c)) {
        c.err = LCB_ERR_CAPACITY | (2 << 8);
        return;
    }
    if (c.lane == 0) c.ar.rs()[c.nrs] = make_int2(lo, hi);
    c.nrs++;
}

__device__ __forceinline__ void inst_extend_reads(Inst &I, int lo, int hi) // single writer
{
    if (lo < I.rlo) I.rlo = lo;
    if (hi > I.rhi) I.rhi = hi;
}

// spill the instance tables from shared memory to the arena
__device__ __noinline__ void inst_spill(const WarpSmem *sm, WarpArena ar, int ninst, int ngood, int nbest, int lane)
{
    for (int i = lane; i < ninst; i += 32) {
        ar.inst()[i] = sm->inst[i];
        ar.ord()[i] = sm->ord[i];
    }
    for (int i = lane; i < ngood; i += 32) ar.good()[i] = sm->good[i];
    for (int i = lane; i < nbest; i += 32) ar.best()[i] = sm->best[i];
    __syncwarp();
}

__device__ __forceinline__ void inst_grow(Ctx &c)
{
    if (c.icap != kInstSmem) {
        c.err = LCB_ERR_CAPACITY | (3 << 8);
        return;
    }
    inst_spill(c.sm, c.ar, c.ninst, c.ngood, c.nbest, c.lane);
    c.inst = c.ar.inst();
    c.ord = c.ar.ord();
    c.good = c.ar.good();
    c.best = c.ar.best();
    c.icap = cap_inst(c);
}

// position of the first instance (multiset order) whose key > `key`  (std::multiset::upper_bound)
// `ord` is sorted by key at all times: an instance changes its key only to the position of an occurrence that was found
// right next to it in this order (ChangeBack on a + instance / ChangeFront on a - instance take the multiset neighbour,
// path.h:446-470,515-540), so the order the reference's tree relies on is the order of this array.  Big tables are
// searched 32 ways per step (three dependent probes for 32768 instances) instead of front to back.
__device__ __forceinline__ int ord_upper_bound(const Ctx &c, int key)
{
    int lo = 0, hi = c.ninst; // every position < lo has key <= `key`, every position >= hi has key > `key`
    while (hi - lo > 64) {
        const int step = (hi - lo + 31) >> 5;
        const int i = lo + (c.lane + 1) * step - 1;
        const bool gt = i >= hi || c.inst[c.ord[i]].key > key;
        const unsigned m = __ballot_sync(kFull, gt);
        if (!m) return hi; // cannot happen (lane 31 probes hi - 1 or beyond), kept for safety
        const int f = ffs_lane(m);
        const int nhi = min(hi, lo + (f + 1) * step - 1);
        lo += f * step;
        hi = nhi;
    }
    for (int base = lo; base < hi; base += 32) {
        int i = base + c.lane;
        bool gt = i < hi && c.inst[c.ord[i]].key > key;
        unsigned m = __ballot_sync(kFull, gt);
        if (m) return base + ffs_lane(m);
    }
    return hi;
}
#ifdef LCB_CHECK_MPV
__device__ __forceinline__ int ord_upper_bound_linear(const Ctx &c, int key)
{
    for (int base = 0; base < c.ninst; base += 32) {
        int i = base + c.lane;
        bool gt = i < c.ninst && c.inst[c.ord[i]].key > key;
        unsigned m = __ballot_sync(kFull, gt);
        if (m) return base + ffs_lane(m);
    }
    return c.ninst;
}
#endif

// Instance(it, distance) + multiset insert + allInstance_.push_back   (path.h:82-91, :43, :492, :561)
__device__ __forceinline__ void inst_insert(Ctx &c, int at, int g, bool pos, int v, unsigned bp, int dist, int flag_idx,
                                         int clo, int chi)
{
    if (c.ninst >= c.icap) {
        inst_grow(c);
        if (c.err) return;
    }
    int id = c.ninst;
    for (int top = c.ninst - 1; top >= at; top -= 32) { // shift [at, ninst) up by one, high chunks first
        int i = top - c.lane;
        unsigned short x = 0;
        if (i >= at) x = c.ord[i];
        __syncwarp();
        if (i >= at) c.ord[i + 1] = x;
        __syncwarp();
    }
    if (c.lane == 0) {
        Inst I;
        I.fg = I.bg = g;
        I.fv = I.bv = v;
        I.fbp = I.bbp = bp;
        I.fdist = I.bdist = dist;
        I.key = g;
        I.rlo = flag_idx >= 0 ? flag_idx : 0x7FFFFFFF;
        I.rhi = flag_idx >= 0 ? flag_idx : -1;
        I.clo = clo;
        I.chi = chi;
        I.flags = pos ? kPos : 0u;
        c.inst[id] = I;
        c.ord[at] = (unsigned short)id;
    }
    c.ninst++;
    __syncwarp();
}

// Path::Clear (path.h:650-677): also flushes the per-instance read extents into the read-set log
__device__ __forceinline__ void path_clear(Ctx &c)
{
    for (int base = 0; base < c.ninst; base += 32) {
        int i = base + c.lane;
        int lo = 0, hi = -1;
        if (i < c.ninst) {
            lo = c.inst[i].rlo;
            hi = c.inst[i].rhi;
        }
        bool live = lo <= hi;
        unsigned m = __ballot_sync(kFull, live);
        int n = __popc(m);
        if (c.nrs + n > cap_rs(c)) {
            c.err = LCB_ERR_CAPACITY | (2 << 8);
            break;
        }
        if (live) c.ar.rs()[c.nrs + __popc(m & ((1u << c.lane) - 1))] = make_int2(lo, hi);
        c.nrs += n;
    }
    hash_clear(c);
    if (c.hbig) { // arena table is all-zero again; go back to shared memory
        c.hbig = false;
        c.hash = c.sm->hash;
        c.hmask = kHashSmem - 1;
        for (int i = c.lane; i < kHashSmem; i += 32) c.hash[i] = make_int2(0, 0);
    }
    // the best snapshot survives Clear(): keep it where it is, move the rest back to shared memory
    if (c.icap != kInstSmem && c.nbest <= kInstSmem) {
        for (int i = c.lane; i < c.nbest; i += 32) c.sm->best[i] = c.best[i];
        __syncwarp();
        ctx_reset_storage(c);
    } else if (c.icap != kInstSmem) {
        c.inst = c.ar.inst(); // stay in the arena
    }
    c.ninst = c.ngood = 0;
    c.nright = c.nleft = 0;
    __syncwarp();
}

// Shadow copy of everything Path holds after the current number of right pushes; only while the state is still in
// shared memory (small paths -- the common case).  Replaces blocksfinder.h:271-284 (Clear, Init, re-push best edges),
// which recomputes exactly this state.
__device__ __forceinline__ void snapshot_state(Ctx &c)
{
    if (c.icap != kInstSmem) {
        c.snap_valid = false;
        return;
    }
    WarpSmem *sm = c.sm;
    const int *src = (const int *)sm->inst;
    int *dst = (int *)sm->s_inst;
    const int words = c.ninst * (int)(sizeof(Inst) / sizeof(int));
    for (int i = c.lane; i < words; i += 32) dst[i] = src[i];
    // a path hash that outgrew shared memory is not copied: restore_state rebuilds it from the right-edge log
    c.snap_hash = !c.hbig;
    if (c.snap_hash)
        for (int i = c.lane; i < kHashSmem; i += 32) sm->s_hash[i] = sm->hash[i];
    if (c.lane < c.ninst) sm->s_ord[c.lane] = sm->ord[c.lane];
    if (c.lane < c.ngood) sm->s_good[c.lane] = sm->good[c.lane];
    c.snap_ninst = c.ninst, c.snap_ngood = c.ngood, c.snap_hcount = c.hcount;
    c.snap_right_flank = c.right_flank, c.snap_right_vertex = c.right_vertex, c.snap_nright = c.nright;
    c.snap_valid = true;
    __syncwarp();
}

// after path_clear(): bring the shadow state back (storage is in shared memory again, hash table zeroed)
__device__ __forceinline__ void restore_state(Ctx &c)
{
    WarpSmem *sm = c.sm;
    const int *src = (const int *)sm->s_inst;
    int *dst = (int *)sm->inst;
    const int words = c.snap_ninst * (int)(sizeof(Inst) / sizeof(int));
    for (int i = c.lane; i < words; i += 32) dst[i] = src[i];
    if (c.lane < c.snap_ninst) sm->ord[c.lane] = sm->s_ord[c.lane];
    if (c.lane < c.snap_ngood) sm->good[c.lane] = sm->s_good[c.lane];
    c.ninst = c.snap_ninst, c.ngood = c.snap_ngood;
    c.right_flank = c.snap_right_flank, c.right_vertex = c.snap_right_vertex, c.nright = c.snap_nright;
    c.left_flank = 0, c.left_vertex = c.origin, c.nleft = 0;
    if (c.snap_hash) {
        for (int i = c.lane; i < kHashSmem; i += 32) sm->hash[i] = sm->s_hash[i];
        c.hcount = c.snap_hcount;
    } else { // same insertion order as Init + the pushes it replaces: origin, then every successful right edge
        hash_insert(c, c.origin, 0);
        for (int i = 0; i < c.snap_nright && !c.err; i++) {
            const int4 e = c.ar.redge()[i];
            hash_insert(c, e.x, e.w);
        }
    }
    __syncwarp();
}

// per-lane occurrence record used by Init and PointPush*
struct Occ {
    int g, v_id, clo, chi, flag;
    unsigned bp;
    bool pos, used;
};

__device__ __forceinline__ Occ load_occurrence(const Ctx &c, unsigned o, int vertex)
{
    Occ r;
    const int2 oc = __ldg(c.ix.occ + o);
    r.g = oc.x & 0x7FFFFFFF;
    r.bp = (unsigned)oc.y;
    r.pos = (oc.x < 0) == (vertex < 0); // JunctionIterator::IsPositiveStrand: stored id == vertex (junctionstorage.h:408-411)
    r.v_id = 0;
    chr_bounds(c.ix, r.g, r.clo, r.chi);
    bool has = r.pos || r.g > r.clo; // IsUsed on the - strand at idx 0 is false (junctionstorage.h:277-282)
    r.flag = has ? (r.pos ? r.g : r.g - 1) : -1;
    r.used = has ? (__ldg(c.E + r.flag) < c.thresh) : false;
    return r;
}

// Path::Init (path.h:33-46)
__device__ __forceinline__ void path_init(Ctx &c, int vid, unsigned char ch)
{
    c.origin = c.right_vertex = c.left_vertex = vid;
    c.right_flank = c.left_flank = 0;
    hash_insert(c, vid, 0);
    int av = vid < 0 ? -vid : vid;
    unsigned o0 = __ldg(c.ix.vtx_off + av), o1 = __ldg(c.ix.vtx_off + av + 1);
    for (unsigned base = o0; base < o1; base += 32) {
        unsigned o = base + (unsigned)c.lane;
        Occ q;
        bool match = false;
        if (o < o1) {
            q = load_occurrence(c, o, vid);
            const int4 rr = __ldg(c.ix.rec + q.g);
            match = (q.pos ? rec_next_ch(rr) : rec_prev_rc(rr)) == ch; // seqIt.GetChar(), junctionstorage.h:234-243
        }
        unsigned mm = __ballot_sync(kFull, match);
        while (mm) { // occurrences in (chr, idx) order
            int src = ffs_lane(mm);
            mm &= mm - 1;
            int g = __shfl_sync(kFull, q.g, src);
            bool pos = __shfl_sync(kFull, (int)q.pos, src);
            bool used = __shfl_sync(kFull, (int)q.used, src);
            unsigned bp = __shfl_sync(kFull, q.bp, src);
            int flag = __shfl_sync(kFull, q.flag, src);
            int clo = __shfl_sync(kFull, q.clo, src), chi = __shfl_sync(kFull, q.chi, src);
            if (used) {
                rs_add(c, flag, flag); // outcome depends on this epoch although no instance is born
            } else {
                int at = ord_upper_bound(c, g);
                inst_insert(c, at, g, pos, vid, bp, 0, flag, clo, chi);
            }
            if (c.err) return;
        }
    }
}

// any epoch < thresh in [lo, hi]?   (the `used` scan of Path::Compatible, path.h:387-393)
__device__ __forceinline__ bool scan_used(Ctx &c, int lo, int hi)
{
    bool any = false;
    for (int base = lo; base <= hi && !any; base += 32) {
        int f = base + c.lane;
        bool u = f <= hi && __ldg(c.E + f) < c.thresh;
        any = __any_sync(kFull, u);
    }
    if (c.collect) c.ct.scan += (unsigned long long)(hi - lo + 1);
    return any;
}

// Path::PointPushBack / PointPushFront with their workers (path.h:430-602).  BACK: `v` = e.GetEndVertex(),
// FRONT: `v` = e.GetStartVertex().  e_ch_g/e_ch_pos locate the junction whose char is e.GetChar();
// e_other is e.GetEndVertex() for FRONT (the far-branch test `start1.GetVertexId() != e.GetEndVertex()`).
// Fast path of the PointPush* workers (straight-line code, the common case): all occurrences of the pushed vertex (<= 32) lie on distinct chromosomes,
// so they cannot see each other's effects inside this push; every lane evaluates one occurrence against the
// pre-push state (multiset neighbours, Within, Compatible incl. its epoch scan), then the effects are applied in
// occurrence order (goodInstance_ appends, new instances).  Returns false when the precondition does not hold.
__device__ __forceinline__ bool push_parallel(Ctx &c, const bool BACK, int v, int dist, unsigned o0, unsigned cnt, int e_ch_g,
                                              bool e_ch_pos, int e_other)
{
    const bool live = (unsigned)c.lane < cnt;
    Occ q;
    q.g = 0, q.bp = 0, q.pos = false, q.used = false, q.flag = -1, q.clo = -1 - c.lane, q.chi = 0, q.v_id = 0;
    if (live) q = load_occurrence(c, o0 + (unsigned)c.lane, v);
    { // an occurrence list is sorted by (chr, idx): two occurrences on one chromosome are neighbours in it
        const int prev_clo = __shfl_up_sync(kFull, q.clo, 1);
        if (__any_sync(kFull, live && c.lane > 0 && prev_clo == q.clo)) return false;
    }
    // multiset neighbours of every occurrence: lane p holds the p-th instance (id, key) of the multiset order, every lane
    // finds its upper_bound by shuffles instead of chasing ord[] -> inst[].key through shared memory
    const int n = c.ninst;
    int ub = n, hi_cand = -1, lo_cand = -1, hi_key = 0, lo_key = 0;
    if (n <= 32) {
        int ordreg = 0, keyreg = 0x7FFFFFFF;
        if (c.lane < n) {
            ordreg = c.ord[c.lane];
            keyreg = c.inst[ordreg].key;
        }
        for (int p = n - 1; p >= 0; p--) {
            const int kp = __shfl_sync(kFull, keyreg, p);
            if (kp > q.g) ub = p;
        }
        const int hs = min(ub, 31), ls = max(ub - 1, 0);
        hi_cand = __shfl_sync(kFull, ordreg, hs), hi_key = __shfl_sync(kFull, keyreg, hs);
        lo_cand = __shfl_sync(kFull, ordreg, ls), lo_key = __shfl_sync(kFull, keyreg, ls);
    } else if (live) {
        int lo = 0, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (c.inst[c.ord[mid]].key > q.g) hi = mid;
            else lo = mid + 1;
        }
        ub = lo;
        if (ub < n) hi_cand = c.ord[ub], hi_key = c.inst[hi_cand].key;
        if (ub > 0) lo_cand = c.ord[ub - 1], lo_key = c.inst[lo_cand].key;
    }
    int outcome = 0, cand = -1, scan_lo = 0, scan_hi = -1; // 0 skip, 1 extend, 2 new instance, 3 found used
    if (live) {
        int hi_id = -1, lo_id = -1;
        if (ub < n && hi_key < q.chi) hi_id = hi_cand;
        if (ub > 0 && lo_key >= q.clo) lo_id = lo_cand;
        bool within = false;
        if (hi_id >= 0) {
            int a = c.inst[hi_id].fg, b = c.inst[hi_id].bg;
            within = q.g >= min(a, b) && q.g <= max(a, b);
        }
        if (!within) {
            cand = (q.pos == BACK) ? lo_id : hi_id;
            bool extend = false;
            int cend_v = 0;
            if (cand >= 0) {
                const Inst &I = c.inst[cand];
                const bool cpos = (I.flags & kPos) != 0;
                const int cg = BACK ? I.bg : I.fg;
                const unsigned cbp = BACK ? I.bbp : I.fbp;
                const int cdist = BACK ? I.bdist : I.fdist;
                cend_v = BACK ? I.bv : I.fv;
                if (cpos == q.pos) {
                    long long rd = BACK ? (long long)q.bp - (long long)cbp : (long long)cbp - (long long)q.bp;
                    if (!q.pos) rd = -rd;
                    const long long ad = BACK ? (long long)dist - cdist : (long long)cdist - dist;
                    bool ok = rd >= 0;
                    if (ok && (rd > c.pr.b || ad > c.pr.b)) {
                        const int step = q.pos ? 1 : -1;
                        ok = BACK ? (q.g == cg + step) : (cg == q.g + step);
                        if (ok) {
                            const int4 ce = __ldg(c.ix.rec + e_ch_g), cs = __ldg(c.ix.rec + (BACK ? cg : q.g));
                            ok = (q.pos ? rec_next_ch(cs) : rec_prev_rc(cs)) == (e_ch_pos ? rec_next_ch(ce) : rec_prev_rc(ce));
                            if (!BACK) ok = ok && I.fv == e_other;
                        }
                    }
                    if (ok) {
                        scan_lo = min(cg, q.g);
                        scan_hi = max(cg, q.g) - 1;
                        for (int f = scan_lo; f <= scan_hi && ok; f++) ok = !(__ldg(c.E + f) < c.thresh);
                    }
                    extend = ok;
                }
            }
            outcome = (extend && cend_v != v) ? 1 : (!q.used ? 2 : 3);
        }
    }
    if (c.collect) c.ct.scan += (unsigned long long)__reduce_add_sync(kFull, scan_lo <= scan_hi ? (unsigned)(scan_hi - scan_lo + 1) : 0u);
    __syncwarp(); // every lane has finished reading the instance table (searches, neighbours) before any lane changes it
    // ---- apply.  Candidates of different lanes are different instances (different chromosomes).
    bool newly_good = false;
    if (live && cand >= 0 && scan_lo <= scan_hi) inst_extend_reads(c.inst[cand], scan_lo, scan_hi);
    if (outcome == 1) {
        Inst &I = c.inst[cand];
        const unsigned fin = BACK ? kBFin : kFFin;
        if (!(I.flags & fin)) {
            unsigned a = I.fbp > I.bbp ? I.fbp - I.bbp : I.bbp - I.fbp;
            const bool prev_good = (long long)a >= c.pr.m;
            if (BACK) {
                I.bg = q.g, I.bv = v, I.bbp = q.bp, I.bdist = dist;
                if (q.pos) I.key = q.g;
            } else {
                I.fg = q.g, I.fv = v, I.fbp = q.bp, I.fdist = dist;
                if (!q.pos) I.key = q.g;
            }
            a = I.fbp > I.bbp ? I.fbp - I.bbp : I.bbp - I.fbp;
            newly_good = !prev_good && (long long)a >= c.pr.m;
            if (q.flag >= 0) inst_extend_reads(I, q.flag, q.flag);
            if (q.used) I.flags |= fin;
        }
    }
    const unsigned lt = (1u << c.lane) - 1u;
    const unsigned gm = __ballot_sync(kFull, newly_good);
    if (newly_good) c.good[c.ngood + __popc(gm & lt)] = (unsigned short)cand;
    c.ngood += __popc(gm);
    const unsigned om = __ballot_sync(kFull, outcome == 3);
    if (om) {
        if (c.nrs + __popc(om) > cap_rs(c)) {
            c.err = LCB_ERR_CAPACITY | (2 << 8);
            return true;
        }
        if (outcome == 3) c.ar.rs()[c.nrs + __popc(om & lt)] = make_int2(q.flag, q.flag);
        c.nrs += __popc(om);
    }
    __syncwarp();
    unsigned nm = __ballot_sync(kFull, outcome == 2);
    while (nm) { // allInstance_ order == occurrence order
        const int src = ffs_lane(nm);
        nm &= nm - 1;
        const int g = __shfl_sync(kFull, q.g, src);
        const bool pos = __shfl_sync(kFull, (int)q.pos, src);
        const unsigned bp = __shfl_sync(kFull, q.bp, src);
        const int flag = __shfl_sync(kFull, q.flag, src);
        const int clo = __shfl_sync(kFull, q.clo, src), chi = __shfl_sync(kFull, q.chi, src);
        const int at = ord_upper_bound(c, g);
        inst_insert(c, at, g, pos, v, bp, dist, flag, clo, chi);
        if (c.err) return true;
    }
    return true;
}

// The PointPush* workers for a GROUP of occurrences of the pushed vertex: the longest prefix of [o0, o0 + cnt) (at most
// 32) whose members cannot see each other's effects inside this push.  Every lane evaluates one occurrence against the
// state before the group (multiset neighbours, Within, Compatible incl. its epoch scan), then the effects are applied in
// occurrence order (goodInstance_ appends, new instances).  Returns the number of occurrences consumed (>= 1); the caller
// goes on with the next group against the updated state.
// Two occurrences X < Y (the list is sorted by (chr, idx), i.e. by g) interact only if X changes or creates a multiset
// neighbour of Y: they fall into the same gap of the same chromosome's multiset, or X's upper neighbour is Y's lower one.
// Upper bounds are monotone in g, so it is enough to compare every occurrence with its predecessor; occurrences on
// different chromosomes (the common case) never interact.
__device__ __forceinline__ int push_group(Ctx &c, const bool BACK, int v, int dist, unsigned o0, unsigned cnt, int e_ch_g,
                                          bool e_ch_pos, int e_other)
{
    bool live = (unsigned)c.lane < cnt;
    Occ q;
    q.g = 0, q.bp = 0, q.pos = false, q.used = false, q.flag = -1, q.clo = -1 - c.lane, q.chi = 0, q.v_id = 0;
    if (live) q = load_occurrence(c, o0 + (unsigned)c.lane, v);
    // multiset neighbours of every occurrence: lane p holds the p-th instance (id, key) of the multiset order, every lane
    // finds its upper_bound by shuffles instead of chasing ord[] -> inst[].key through shared memory
    const int n = c.ninst;
    int ub = n, hi_cand = -1, lo_cand = -1, hi_key = 0, lo_key = 0;
    if (n <= 32) {
        int ordreg = 0, keyreg = 0x7FFFFFFF;
        if (c.lane < n) {
            ordreg = c.ord[c.lane];
            keyreg = c.inst[ordreg].key;
        }
        for (int p = n - 1; p >= 0; p--) {
            const int kp = __shfl_sync(kFull, keyreg, p);
            if (kp > q.g) ub = p;
        }
        const int hs = min(ub, 31), ls = max(ub - 1, 0);
        hi_cand = __shfl_sync(kFull, ordreg, hs), hi_key = __shfl_sync(kFull, keyreg, hs);
        lo_cand = __shfl_sync(kFull, ordreg, ls), lo_key = __shfl_sync(kFull, keyreg, ls);
    } else if (live) { // every lane bisects the order on its own (`ord` is sorted by key, see ord_upper_bound)
        int lo = 0, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (c.inst[c.ord[mid]].key > q.g) hi = mid;
            else lo = mid + 1;
        }
        ub = lo;
        if (ub < n) hi_cand = c.ord[ub], hi_key = c.inst[hi_cand].key;
        if (ub > 0) lo_cand = c.ord[ub - 1], lo_key = c.inst[lo_cand].key;
    }
#ifdef LCB_CHECK_MPV
    { // self-check build: the bisection / shuffle search against a front-to-back scan
        int ref = n;
        if (live)
            for (int p = n - 1; p >= 0; p--)
                if (c.inst[c.ord[p]].key > q.g) ref = p;
        if (__any_sync(kFull, live && ref != ub)) c.err = 91;
    }
#endif
    int hi_id = -1, lo_id = -1;
    if (live) {
        if (ub < n && hi_key < q.chi) hi_id = hi_cand;
        if (ub > 0 && lo_key >= q.clo) lo_id = lo_cand;
    }
    int take = (int)min(cnt, 32u);
    { // cut the group in front of the first occurrence that could see the effects of its predecessor
        const int prev_clo = __shfl_up_sync(kFull, q.clo, 1), prev_ub = __shfl_up_sync(kFull, ub, 1);
        const int prev_hi = __shfl_up_sync(kFull, hi_id, 1);
        const bool clash = live && c.lane > 0 && prev_clo == q.clo && (prev_ub == ub || (prev_hi >= 0 && prev_hi == lo_id));
        const unsigned cm = __ballot_sync(kFull, clash);
        if (cm) take = ffs_lane(cm);
        live = c.lane < take;
    }
    int outcome = 0, cand = -1, scan_lo = 0, scan_hi = -1; // 0 skip, 1 extend, 2 new instance, 3 found used
    if (live) {
        bool within = false;
        if (hi_id >= 0) {
            int a = c.inst[hi_id].fg, b = c.inst[hi_id].bg;
            within = q.g >= min(a, b) && q.g <= max(a, b);
        }
        if (!within) {
            cand = (q.pos == BACK) ? lo_id : hi_id;
            bool extend = false;
            int cend_v = 0;
            if (cand >= 0) {
                const Inst &I = c.inst[cand];
                const bool cpos = (I.flags & kPos) != 0;
                const int cg = BACK ? I.bg : I.fg;
                const unsigned cbp = BACK ? I.bbp : I.fbp;
                const int cdist = BACK ? I.bdist : I.fdist;
                cend_v = BACK ? I.bv : I.fv;
                if (cpos == q.pos) {
                    long long rd = BACK ? (long long)q.bp - (long long)cbp : (long long)cbp - (long long)q.bp;
                    if (!q.pos) rd = -rd;
                    const long long ad = BACK ? (long long)dist - cdist : (long long)cdist - dist;
                    bool ok = rd >= 0;
                    if (ok && (rd > c.pr.b || ad > c.pr.b)) {
                        const int step = q.pos ? 1 : -1;
                        ok = BACK ? (q.g == cg + step) : (cg == q.g + step);
                        if (ok) {
                            const int4 ce = __ldg(c.ix.rec + e_ch_g), cs = __ldg(c.ix.rec + (BACK ? cg : q.g));
                            ok = (q.pos ? rec_next_ch(cs) : rec_prev_rc(cs)) == (e_ch_pos ? rec_next_ch(ce) : rec_prev_rc(ce));
                            if (!BACK) ok = ok && I.fv == e_other;
                        }
                    }
                    if (ok) {
                        scan_lo = min(cg, q.g);
                        scan_hi = max(cg, q.g) - 1;
                        for (int f = scan_lo; f <= scan_hi && ok; f++) ok = !(__ldg(c.E + f) < c.thresh);
                    }
                    extend = ok;
                }
            }
            outcome = (extend && cend_v != v) ? 1 : (!q.used ? 2 : 3);
        }
    }
    if (c.collect) c.ct.scan += (unsigned long long)__reduce_add_sync(kFull, scan_lo <= scan_hi ? (unsigned)(scan_hi - scan_lo + 1) : 0u);
    __syncwarp(); // every lane has finished reading the instance table (searches, neighbours) before any lane changes it
    // ---- apply.  Candidates of different lanes are different instances (see the group rule above).
    bool newly_good = false;
    if (live && cand >= 0 && scan_lo <= scan_hi) inst_extend_reads(c.inst[cand], scan_lo, scan_hi);
    if (outcome == 1) {
        Inst &I = c.inst[cand];
        const unsigned fin = BACK ? kBFin : kFFin;
        if (!(I.flags & fin)) {
            unsigned a = I.fbp > I.bbp ? I.fbp - I.bbp : I.bbp - I.fbp;
            const bool prev_good = (long long)a >= c.pr.m;
            if (BACK) {
                I.bg = q.g, I.bv = v, I.bbp = q.bp, I.bdist = dist;
                if (q.pos) I.key = q.g;
            } else {
                I.fg = q.g, I.fv = v, I.fbp = q.bp, I.fdist = dist;
                if (!q.pos) I.key = q.g;
            }
            a = I.fbp > I.bbp ? I.fbp - I.bbp : I.bbp - I.fbp;
            newly_good = !prev_good && (long long)a >= c.pr.m;
            if (q.flag >= 0) inst_extend_reads(I, q.flag, q.flag);
            if (q.used) I.flags |= fin;
        }
    }
    const unsigned lt = (1u << c.lane) - 1u;
    const unsigned gm = __ballot_sync(kFull, newly_good);
    if (newly_good) c.good[c.ngood + __popc(gm & lt)] = (unsigned short)cand;
    c.ngood += __popc(gm);
    const unsigned om = __ballot_sync(kFull, outcome == 3);
    if (om) {
        if (c.nrs + __popc(om) > cap_rs(c)) {
            c.err = LCB_ERR_CAPACITY | (2 << 8);
            return take;
        }
        if (outcome == 3) c.ar.rs()[c.nrs + __popc(om & lt)] = make_int2(q.flag, q.flag);
        c.nrs += __popc(om);
    }
    __syncwarp();
    unsigned nm = __ballot_sync(kFull, outcome == 2);
    while (nm) { // allInstance_ order == occurrence order
        const int src = ffs_lane(nm);
        nm &= nm - 1;
        const int g = __shfl_sync(kFull, q.g, src);
        const bool pos = __shfl_sync(kFull, (int)q.pos, src);
        const unsigned bp = __shfl_sync(kFull, q.bp, src);
        const int flag = __shfl_sync(kFull, q.flag, src);
        const int clo = __shfl_sync(kFull, q.clo, src), chi = __shfl_sync(kFull, q.chi, src);
        const int at = ord_upper_bound(c, g);
#ifdef LCB_CHECK_MPV
        if (at != ord_upper_bound_linear(c, g)) c.err = 91;
#endif
        inst_insert(c, at, g, pos, v, bp, dist, flag, clo, chi);
        if (c.err) return take;
    }
    return take;
}

// occ_first/occ_count: the pushed vertex's occurrence range when the caller already holds it (rec[g].z/.w), else -1
__device__ __forceinline__ bool path_push(Ctx &c, const bool BACK, int v, int len, int e_ch_g, bool e_ch_pos, int e_other,
                                          int occ_first, int occ_count)
{
    if (hash_find(c.hash, c.hmask, v) != kNotSet) return false; // vertex already in the path
    const int dist = BACK ? c.right_flank + len : c.left_flank - len;
    hash_insert(c, v, dist);
    if (c.err) return true;
    unsigned o0, o1;
    if (occ_count >= 0) {
        o0 = (unsigned)occ_first, o1 = o0 + (unsigned)occ_count;
    } else {
        int av = v < 0 ? -v : v;
        o0 = __ldg(c.ix.vtx_off + av), o1 = __ldg(c.ix.vtx_off + av + 1);
    }
    if (c.collect) c.ct.occ += o1 - o0, c.ct.pushes++;
    bool handled = false;
    if (o1 - o0 <= 32u) {
        handled = push_parallel(c, BACK, v, dist, o0, o1 - o0, e_ch_g, e_ch_pos, e_other);
        if (c.err) return true;
        if (c.collect) {
            if (handled) c.ct.push_par++;
            else c.ct.push_ser++;
        }
    }
    // everything else (a chromosome that carries the vertex twice, more than 32 occurrences): conflict-free groups
    for (unsigned o = o0; o < o1 && !handled;) {
        const int took = push_group(c, BACK, v, dist, o, o1 - o, e_ch_g, e_ch_pos, e_other);
        if (c.err) return true;
        o += (unsigned)took;
    }
    if (BACK) {
        if (c.nright >= cap_path(c)) {
            c.err = LCB_ERR_CAPACITY | (1 << 8);
            return true;
        }
        if (c.lane == 0) c.ar.redge()[c.nright] = make_int4(v, len, e_ch_g | (e_ch_pos ? (int)0x80000000 : 0), dist);
        c.nright++;
        c.right_flank = dist;
        c.right_vertex = v;
    } else {
        c.nleft++;
        c.left_flank = dist;
        c.left_vertex = v;
    }
    __syncwarp();
    return true;
}

// Path::Score (path.h:604-628): any flank penalty >= maxFlankingSize makes the whole score -INT32_MAX.
// With flank <= 32767 every squared penalty fits 32 bits and sum(real) is split in 16-bit halves, so four 32-bit warp
// reductions are exact; larger -b values take 64-bit shuffles.
__device__ __forceinline__ long long path_score(Ctx &c)
{
    long long total = 0;
    bool bad = false;
    const bool small = c.pr.flank <= 32767;
    for (int base = 0; base < c.ngood; base += 32) {
        const int i = base + c.lane;
        unsigned real = 0;
        long long pen = 0;
        if (i < c.ngood) {
            const Inst &I = c.inst[c.good[i]];
            real = I.fbp > I.bbp ? I.fbp - I.bbp : I.bbp - I.fbp;
            const long long rp = (long long)c.right_flank - I.bdist;
            const long long lp = (long long)(-c.left_flank) + I.fdist;
            if (lp >= c.pr.flank || rp >= c.pr.flank) bad = true;
            else pen = rp + lp;
        }
        if (small) {
            const unsigned pen2 = (unsigned)(pen * pen);
            const unsigned lo = __reduce_add_sync(kFull, real & 0xFFFFu), hi = __reduce_add_sync(kFull, real >> 16);
            const unsigned pp = __reduce_add_sync(kFull, pen2 >> 5), pr = __reduce_add_sync(kFull, pen2 & 31u);
            total += (long long)lo + ((long long)hi << 16) - (((long long)pp << 5) + pr);
        } else {
            long long term = (long long)real - pen * pen;
#pragma unroll
            for (int d = 16; d; d >>= 1) term += __shfl_xor_sync(kFull, term, d);
            total += term;
        }
    }
    if (c.collect) c.ct.score += (unsigned long long)c.ngood;
    bad = __any_sync(kFull, bad);
    return bad ? -(long long)0x7FFFFFFF : total;
}

// bestInstance = copies of *goodInstance_[i] in list order (blocksfinder.h:818-825, :881-888)
__device__ __forceinline__ void snapshot_best(Ctx &c)
{
    for (int i = c.lane; i < c.ngood; i += 32) {
        const Inst &I = c.inst[c.good[i]];
        c.best[i] = make_int4(I.fg | ((I.flags & kPos) ? (int)0x80000000 : 0), I.bg, (int)I.fbp, (int)I.bbp);
    }
    c.nbest = c.ngood;
    __syncwarp();
}

struct Next { // result of MostPopularVertex
    int vid, og, d;
    bool opos;
};

// BlocksFinder::MostPopularVertex (blocksfinder.h:708-768), general path.  Up to four look-ahead walks share the warp per
// pass (one lane per (walk, depth)); walks that run out of lanes continue in the next pass.  The vote is accumulated
// order-independently in a small hash table (count += weight, last event = max (list position, depth)) and resolved in
// closed form: among the vertices whose final count is the maximum, the one whose last increment came from the
// smallest origin (strand, chr, idx) wins, earliest event on ties -- exactly where the reference's running arg-max
// with its `origin < ret.origin` tie-break ends (blocksfinder.h:733-741).
__device__ __forceinline__ Next most_popular_vertex(Ctx &c, bool forward, bool try_used)
{
    Next best;
    best.vid = 0, best.og = 0, best.d = 0, best.opos = false;
    const int start_vid = forward ? c.right_vertex : c.left_vertex;
    const bool use_good = c.ngood >= 2;
    const int n = use_good ? c.ngood : c.ninst;
    int2 *tab = c.sm->vote;
    unsigned *last = c.sm->vlast;
    int cap = kVoteSmem;
    c.vote_clean = false;
    for (int attempt = 0;; attempt++) {
        for (int i = c.lane; i < cap; i += 32) tab[i] = make_int2(0, 0), last[i] = 0u;
        __syncwarp();
        int distinct = 0;
        bool overflow = false;
        for (int lb = 0; lb < n && !overflow; lb += 32) {
            const int qi = lb + c.lane;
            int my_id = 0;
            bool elig = false;
            if (qi < n) {
                my_id = use_good ? (int)c.good[qi] : qi;
                elig = (forward ? c.inst[my_id].bv : c.inst[my_id].fv) == start_vid;
            }
            const unsigned em = __ballot_sync(kFull, elig);
            const int E = __popc(em);
            for (int gb = 0; gb < E && !overflow; gb += 4) {
                unsigned pend = E - gb >= 4 ? 0xFu : ((1u << (E - gb)) - 1u);
                int d0 = 0;
                while (pend && !overflow) {
                    if (distinct + 32 >= cap) { // a pass inserts at most 32 keys; probing needs one empty slot to terminate
                        overflow = true;
                        break;
                    }
                    const int K = __popc(pend), L = 32 / K;
                    const int k = c.lane / L, dd = c.lane % L + 1;
                    const bool lane_on = k < K;
                    const int wi = lane_on ? kth_bit4(pend, k) : 0;
                    // lane of the list entry that is the (gb + wi)-th eligible one: K ballots instead of __fns per lane
                    int src = 0;
                    {
                        const int my_rank = __popc(em & ((1u << c.lane) - 1u));
                        for (int kk = 0; kk < K; kk++) {
                            const int target = gb + kth_bit4(pend, kk);
                            const unsigned hit = __ballot_sync(kFull, elig && my_rank == target);
                            if (k == kk) src = ffs_lane(hit);
                        }
                    }
                    const int id = __shfl_sync(kFull, my_id, src & 31);
                    const unsigned qord = (unsigned)(lb + (src & 31));
                    const Inst &I = c.inst[id];
                    const bool pos = (I.flags & kPos) != 0;
                    const int og = forward ? I.bg : I.fg;
                    const unsigned obp = forward ? I.bbp : I.fbp;
                    const unsigned weight = (I.fbp > I.bbp ? I.fbp - I.bbp : I.bbp - I.fbp) + 1u;
                    const int clo = I.clo, chi = I.chi;
                    const int step = (forward == pos) ? 1 : -1;
                    const int d = d0 + dd;
                    const int g = og + step * d;
                    bool in_range = lane_on && g >= clo && g < chi; // it.Valid()
                    int vid = 0, flag = -1;
                    bool used = false, inpath = false;
                    if (in_range) {
                        const int4 rc = __ldg(c.ix.rec + g);
                        prefetch_l1(c.ix.occ + rc.z); // the push of this vertex starts with its occurrence list
                        vid = pos ? rc.x : -rc.x;
                        long long dp = (long long)(unsigned)rc.y - (long long)obp;
                        if (dp < 0) dp = -dp;
                        in_range = d < c.pr.depth || dp <= c.pr.b;
                        if (in_range) {
                            const bool has = pos || g > clo;
                            flag = has ? (pos ? g : g - 1) : -1;
                            if (has && !try_used) used = __ldg(c.E + flag) < c.thresh;
                            inpath = hash_find(c.hash, c.hmask, vid) != kNotSet;
                        }
                    }
                    const bool ok = in_range && !inpath && !used;
                    const unsigned seg = !lane_on ? 0u : (L == 32 ? kFull : (((1u << L) - 1u) << (k * L)));
                    const unsigned failm = __ballot_sync(kFull, lane_on && !ok);
                    const unsigned fail = failm & seg;
                    const int nok = fail ? ffs_lane(fail) - k * L : L;
                    const bool active = lane_on && dd - 1 < nok;
                    bool fresh = false;
                    if (active) { // count[vid] += weight; remember the last (list position, depth) that touched it
                        unsigned sl = (hash_of(vid) >> 12) & (unsigned)(cap - 1);
                        while (true) {
                            const int old = atomicCAS(&tab[sl].x, 0, vid);
                            if (old == 0 || old == vid) {
                                fresh = old == 0;
                                break;
                            }
                            sl = (sl + 1) & (unsigned)(cap - 1);
                        }
                        atomicAdd((unsigned *)&tab[sl].y, weight);
                        atomicMax(&last[sl], (qord << 16) | (unsigned)d); // list position < 65536; d <= -b + depth < 65536 (checked at create)
                    }
                    distinct += __popc(__ballot_sync(kFull, fresh));
                    { // loop-body executions and the epochs each walk depended on
                        const bool stop_in_body = lane_on && dd - 1 == nok && in_range;
                        if (c.collect && attempt == 0) c.ct.walk += (unsigned long long)__popc(__ballot_sync(kFull, active || stop_in_body));
                        const bool dep = flag >= 0 && !try_used && (active || (stop_in_body && !inpath));
                        for (int kk = 0; kk < K; kk++) {
                            const int lo = __reduce_min_sync(kFull, dep && k == kk ? flag : 0x7FFFFFFF);
                            const int hi = __reduce_max_sync(kFull, dep && k == kk ? flag : -1);
                            const int sid = __shfl_sync(kFull, id, kk * L);
                            if (lo <= hi && c.lane == 0) inst_extend_reads(c.inst[sid], lo, hi);
                        }
                    }
                    // walks that found their end leave; the others go on from depth d0 + L
                    unsigned next_pend = 0;
                    for (int kk = 0; kk < K; kk++) {
                        const unsigned segk = L == 32 ? kFull : (((1u << L) - 1u) << (kk * L));
                        if (!(failm & segk)) next_pend |= 1u << kth_bit4(pend, kk);
                    }
                    pend = next_pend;
                    d0 += L;
                    __syncwarp();
                }
            }
        }
        if (!overflow) break;
        if (cap >= cap_vote(c)) {
            c.err = LCB_ERR_CAPACITY | (4 << 8);
            return best;
        }
        // start over with the arena's table: 8192 slots first, then (big arena slots only) up to eight times more each time
        tab = c.ar.vote(), last = c.ar.vlast(), cap = cap == kVoteSmem ? kVoteMax : min(cap * 8, cap_vote(c));
    }
    __syncwarp();
    // ---- resolve
    unsigned M = 0;
    for (int base = 0; base < cap; base += 32) {
        const int2 e = tab[base + c.lane];
        M = max(M, __reduce_max_sync(kFull, e.x ? (unsigned)e.y : 0u));
    }
    if (M == 0) return best;
    unsigned best_okey = 0xFFFFFFFFu, best_ev = 0xFFFFFFFFu;
    for (int base = 0; base < cap; base += 32) {
        const int2 e = tab[base + c.lane];
        const bool cand = e.x != 0 && (unsigned)e.y == M;
        unsigned okey = 0xFFFFFFFFu, ev = 0xFFFFFFFFu;
        int og = 0;
        bool pos = false;
        if (cand) {
            ev = last[base + c.lane];
            const int q = (int)(ev >> 16);
            const Inst &I = c.inst[use_good ? (int)c.good[q] : q];
            pos = (I.flags & kPos) != 0;
            og = forward ? I.bg : I.fg;
            okey = (pos ? 0x80000000u : 0u) | (unsigned)og; // origin order: - strand first, then (chr, idx)
        }
        if (!__any_sync(kFull, cand)) continue;
        const unsigned kmin = __reduce_min_sync(kFull, okey);
        const unsigned emin = __reduce_min_sync(kFull, cand && okey == kmin ? ev : 0xFFFFFFFFu);
        if (kmin < best_okey || (kmin == best_okey && emin < best_ev)) {
            const int wl = ffs_lane(__ballot_sync(kFull, cand && okey == kmin && ev == emin));
            best_okey = kmin, best_ev = emin;
            best.vid = __shfl_sync(kFull, e.x, wl);
            best.og = __shfl_sync(kFull, og, wl);
            best.opos = __shfl_sync(kFull, (int)pos, wl) != 0;
            best.d = (int)(emin & 0xFFFFu);
        }
    }
    return best;
}


// MostPopularVertex for up to four walks of up to 8 * kMidTiers junctions each -- the common case on a handful of
// genomes, where the general path would spend two or three serial passes.  Every lane owns one depth of one walk in each
// tier, so all junction records and epochs of all walks are requested at once (one memory round trip), the vote is
// accumulated once in the shared-memory table, and the closed-form resolution (see most_popular_vertex) is evaluated by
// the lanes that own the items instead of by scanning the table; the lanes then empty the slots they used.
// Returns 0 when the general path must be taken (more walks, or a walk longer than its lanes).
constexpr int kMidTiers = 3;
__device__ __forceinline__ int mpv_mid(Ctx &c, bool forward, bool try_used, Next &best)
{
    const int start_vid = forward ? c.right_vertex : c.left_vertex;
    const bool use_good = c.ngood >= 2;
    const int n = use_good ? c.ngood : c.ninst;
    best.vid = 0, best.og = 0, best.d = 0, best.opos = false;
    if (n > 32) return 0;
    int my_id = 0;
    bool elig = false;
    if (c.lane < n) {
        my_id = use_good ? (int)c.good[c.lane] : c.lane;
        elig = (forward ? c.inst[my_id].bv : c.inst[my_id].fv) == start_vid;
    }
    const unsigned em = __ballot_sync(kFull, elig);
    const int E = __popc(em);
    if (E == 0) return 1;
    if (E > 4) return 0;
    const int k = c.lane >> 3, dd = c.lane & 7;
    const bool lane_on = k < E;
    int src = 0;
    {
        unsigned m = em;
        for (int kk = 0; kk < E; kk++) {
            const int l = ffs_lane(m);
            m &= m - 1;
            if (k == kk) src = l;
        }
    }
    const int id = __shfl_sync(kFull, my_id, src);
    const Inst &I = c.inst[id];
    const bool pos = (I.flags & kPos) != 0;
    const int og = forward ? I.bg : I.fg;
    const unsigned obp = forward ? I.bbp : I.fbp;
    const unsigned weight = (I.fbp > I.bbp ? I.fbp - I.bbp : I.bbp - I.fbp) + 1u;
    const int clo = I.clo, chi = I.chi;
    const int step = (forward == pos) ? 1 : -1;
    const unsigned seg = 0xFFu << (k * 8);
    int vid[kMidTiers], flag[kMidTiers];
    bool inr[kMidTiers], ok[kMidTiers], inpath[kMidTiers];
    {
        int4 rc[kMidTiers];
        uint32_t ep[kMidTiers];
#pragma unroll
        for (int t = 0; t < kMidTiers; t++) { // every load of every walk is in flight before the first one is used
            const int g = og + step * (t * 8 + dd + 1);
            inr[t] = lane_on && g >= clo && g < chi; // it.Valid()
            const bool has = pos || g > clo;
            flag[t] = (inr[t] && has) ? (pos ? g : g - 1) : -1;
            rc[t] = make_int4(0, 0, 0, 0);
            ep[t] = kFree;
            if (inr[t]) rc[t] = __ldg(c.ix.rec + g);
            if (flag[t] >= 0 && !try_used) ep[t] = __ldg(c.E + flag[t]);
        }
#pragma unroll
        for (int t = 0; t < kMidTiers; t++) {
            const int d = t * 8 + dd + 1;
            vid[t] = 0;
            inpath[t] = false;
            bool used = false;
            if (inr[t]) {
                if (t == 0) prefetch_l1(c.ix.occ + rc[t].z); // the push of this vertex starts with its occurrence list
                vid[t] = pos ? rc[t].x : -rc[t].x;
                long long dp = (long long)(unsigned)rc[t].y - (long long)obp;
                if (dp < 0) dp = -dp;
                inr[t] = d < c.pr.depth || dp <= c.pr.b;
            }
            if (inr[t]) {
                used = flag[t] >= 0 && ep[t] < c.thresh; // ep stays kFree under try_used
                inpath[t] = hash_find(c.hash, c.hmask, vid[t]) != kNotSet;
            } else {
                flag[t] = -1;
            }
            ok[t] = inr[t] && !inpath[t] && !used;
        }
    }
    // first junction of every walk that ends it
    int nok = 8 * kMidTiers;
#pragma unroll
    for (int t = kMidTiers - 1; t >= 0; t--) {
        const unsigned fail = __ballot_sync(kFull, lane_on && !ok[t]) & seg;
        if (fail) nok = t * 8 + ffs_lane(fail) - k * 8;
    }
    if (__any_sync(kFull, lane_on && nok == 8 * kMidTiers)) return 0; // a walk needs more depth than its lanes offer
    bool active[kMidTiers];
    int mylo = 0x7FFFFFFF, myhi = -1;
    unsigned steps = 0;
#pragma unroll
    for (int t = 0; t < kMidTiers; t++) {
        const int di = t * 8 + dd; // 0-based depth
        active[t] = lane_on && di < nok;
        const bool stop_in_body = lane_on && di == nok && inr[t];
        if (c.collect) steps += (unsigned)__popc(__ballot_sync(kFull, active[t] || stop_in_body));
        const bool dep = flag[t] >= 0 && !try_used && (active[t] || (stop_in_body && !inpath[t]));
        if (dep) mylo = min(mylo, flag[t]), myhi = max(myhi, flag[t]);
    }
    if (c.collect) c.ct.walk += steps;
    for (int s = 0; s < E; s++) { // the epochs each walk depended on
        const int lo = __reduce_min_sync(kFull, k == s ? mylo : 0x7FFFFFFF);
        const int hi = __reduce_max_sync(kFull, k == s ? myhi : -1);
        const int sid = __shfl_sync(kFull, id, s * 8);
        if (lo <= hi && c.lane == 0) inst_extend_reads(c.inst[sid], lo, hi);
    }
    __syncwarp();
    // ---- vote
    int2 *tab = c.sm->vote;
    unsigned *last = c.sm->vlast;
    if (!c.vote_clean) {
        for (int i = c.lane; i < kVoteSmem; i += 32) tab[i] = make_int2(0, 0), last[i] = 0u;
        c.vote_clean = true;
        __syncwarp();
    }
    unsigned sl[kMidTiers], ev[kMidTiers];
#pragma unroll
    for (int t = 0; t < kMidTiers; t++) {
        sl[t] = 0;
        ev[t] = ((unsigned)src << 20) | (unsigned)(t * 8 + dd + 1);
        if (active[t]) { // count[vid] += weight; remember the last (list position, depth) that touched it
            unsigned x = (hash_of(vid[t]) >> 12) & (unsigned)(kVoteSmem - 1);
            while (true) {
                const int old = atomicCAS(&tab[x].x, 0, vid[t]);
                if (old == 0 || old == vid[t]) break;
                x = (x + 1) & (unsigned)(kVoteSmem - 1);
            }
            sl[t] = x;
            atomicAdd((unsigned *)&tab[x].y, weight);
            atomicMax(&last[x], ev[t]);
        }
    }
    __syncwarp();
    // ---- resolve: among the vertices with the maximal final count, the one whose LAST increment came from the smallest
    // origin (- strand first, then (chr, idx)); that increment's item is the winner and also names the walk to follow
    unsigned total[kMidTiers];
    bool is_last[kMidTiers];
    unsigned mymax = 0;
#pragma unroll
    for (int t = 0; t < kMidTiers; t++) {
        total[t] = 0;
        is_last[t] = false;
        if (active[t]) {
            total[t] = (unsigned)tab[sl[t]].y;
            is_last[t] = last[sl[t]] == ev[t];
            mymax = max(mymax, total[t]);
        }
    }
    const unsigned M = __reduce_max_sync(kFull, mymax);
    __syncwarp();
#pragma unroll
    for (int t = 0; t < kMidTiers; t++)
        if (active[t]) tab[sl[t]] = make_int2(0, 0), last[sl[t]] = 0u; // leave the table empty
    __syncwarp();
    if (M == 0) return 1;
    const unsigned okey = (pos ? 0x80000000u : 0u) | (unsigned)og;
    unsigned myev = 0xFFFFFFFFu;
    int myt = -1;
#pragma unroll
    for (int t = 0; t < kMidTiers; t++)
        if (is_last[t] && total[t] == M && ev[t] < myev) myev = ev[t], myt = t;
    const unsigned kmin = __reduce_min_sync(kFull, myt >= 0 ? okey : 0xFFFFFFFFu);
    // candidates of one walk share okey; different walks with equal okey cannot exist (an origin is one junction+strand)
    const unsigned emin = __reduce_min_sync(kFull, (myt >= 0 && okey == kmin) ? myev : 0xFFFFFFFFu);
    const unsigned win = __ballot_sync(kFull, myt >= 0 && okey == kmin && myev == emin);
    if (!win) return 1;
    const int wl = ffs_lane(win);
    int wv = 0;
#pragma unroll
    for (int t = 0; t < kMidTiers; t++)
        if (myt == t) wv = vid[t];
    best.vid = __shfl_sync(kFull, wv, wl);
    best.og = __shfl_sync(kFull, og, wl);
    best.opos = __shfl_sync(kFull, (int)pos, wl) != 0;
    best.d = (int)(emin & 0xFFFFFu);
    return 1;
}

struct Staged { // per-lane copy of the look-ahead walks of the last MostPopularVertex (fast path)
    int vid, o0;
    unsigned bp, w;
    int base; // first lane of the winning walk (uniform)
    unsigned obp;
    bool valid;
};

// MostPopularVertex when <= 4 instances sit on the path end and every look-ahead fits its lane segment: all walks are
// fetched at once (one lane per (instance, depth)) and the vote is evaluated in closed form instead of being replayed:
// counts are order-independent sums; the running arg-max of the reference ends on -- among the vertices whose FINAL
// count is the maximum M -- the one whose last increment came from the smallest origin (strand, chr, idx), earliest
// such event on ties (proof sketch in DESIGN.md section 4).  Returns 0 when the general path must be taken.
__device__ __forceinline__ int mpv_fast(Ctx &c, bool forward, bool try_used, Next &best, Staged &sg)
{
    const int start_vid = forward ? c.right_vertex : c.left_vertex;
    const bool use_good = c.ngood >= 2;
    const int n = use_good ? c.ngood : c.ninst;
    best.vid = 0, best.og = 0, best.d = 0, best.opos = false;
    sg.valid = false;
    if (n > 32) return 0;
    int my_id = 0;
    bool elig = false;
    if (c.lane < n) {
        my_id = use_good ? (int)c.good[c.lane] : c.lane;
        elig = (forward ? c.inst[my_id].bv : c.inst[my_id].fv) == start_vid;
    }
    const unsigned em = __ballot_sync(kFull, elig);
    const int E = __popc(em);
    if (E == 0) return 1;
    if (E > 2) return 0; // three or more: the general path shares the warp between walks and continues them
    const int L = 32 / E; // 32 or 16 lanes (= depths) per instance
    const int slot = c.lane / L, d = c.lane % L + 1;
    const bool lane_on = slot < E;
    const int src = !lane_on ? 0 : (slot == 0 ? ffs_lane(em) : ffs_lane(em & (em - 1))); // E <= 2 here
    const int id = __shfl_sync(kFull, my_id, src);
    const Inst &I = c.inst[id];
    const bool pos = (I.flags & kPos) != 0;
    const int og = forward ? I.bg : I.fg;
    const unsigned obp = forward ? I.bbp : I.fbp;
    const unsigned weight = (I.fbp > I.bbp ? I.fbp - I.bbp : I.bbp - I.fbp) + 1u;
    const int clo = I.clo, chi = I.chi;
    const int step = (forward == pos) ? 1 : -1;
    const int g = og + step * d;
    bool in_range = lane_on && g >= clo && g < chi;
    int vid = 0, flag = -1, o0 = 0;
    unsigned bp = 0, w = 0;
    bool used = false, inpath = false;
    if (in_range) {
        const int4 rc = __ldg(c.ix.rec + g);
        prefetch_l1(c.ix.occ + rc.z);
        vid = pos ? rc.x : -rc.x;
        bp = (unsigned)rc.y, o0 = rc.z, w = (unsigned)rc.w;
        long long dp = (long long)bp - (long long)obp;
        if (dp < 0) dp = -dp;
        in_range = d < c.pr.depth || dp <= c.pr.b;
        if (in_range) {
            const bool has = pos || g > clo;
            flag = has ? (pos ? g : g - 1) : -1;
            if (has && !try_used) used = __ldg(c.E + flag) < c.thresh;
            inpath = hash_find(c.hash, c.hmask, vid) != kNotSet;
        }
    }
    const bool ok = in_range && !inpath && !used;
    const unsigned seg = !lane_on ? 0u : (L == 32 ? kFull : (((1u << L) - 1u) << (slot * L)));
    const unsigned fail = __ballot_sync(kFull, lane_on && !ok) & seg;
    const int nok = fail ? ffs_lane(fail) - slot * L : L;
    if (__any_sync(kFull, lane_on && nok == L)) return 0; // a walk needs more depth than its segment offers
    const bool active = lane_on && d - 1 < nok;
    { // loop-body executions (walk steps) and the epochs each walk depended on
        const bool stop_in_body = lane_on && d - 1 == nok && in_range;
        if (c.collect) c.ct.walk += (unsigned long long)__popc(__ballot_sync(kFull, active || stop_in_body));
        const bool dep = flag >= 0 && !try_used && (active || (lane_on && d - 1 == nok && in_range && !inpath));
        for (int s = 0; s < E; s++) {
            const int lo = __reduce_min_sync(kFull, dep && slot == s ? flag : 0x7FFFFFFF);
            const int hi = __reduce_max_sync(kFull, dep && slot == s ? flag : -1);
            const int sid = __shfl_sync(kFull, id, s * L);
            if (lo <= hi && c.lane == 0) inst_extend_reads(c.inst[sid], lo, hi);
        }
        __syncwarp();
    }
    const unsigned long long key = active ? (unsigned long long)(unsigned)vid : (0x100000000ull | (unsigned)c.lane);
    const unsigned peers = __match_any_sync(kFull, key);
    unsigned total = 0;
    for (int s = 0; s < E; s++) {
        const unsigned segs = L == 32 ? kFull : (((1u << L) - 1u) << (s * L));
        total += (unsigned)__popc(peers & segs) * __shfl_sync(kFull, weight, s * L);
    }
    const bool is_last = active && c.lane == 31 - __clz((int)peers);
    const unsigned M = __reduce_max_sync(kFull, active ? total : 0u);
    const bool cand = is_last && total == M;
    const unsigned okey = (pos ? 0x80000000u : 0u) | (unsigned)og; // origin order: - strand first, then (chr, idx)
    const unsigned kmin = __reduce_min_sync(kFull, cand ? okey : 0xFFFFFFFFu);
    const unsigned win = __ballot_sync(kFull, cand && okey == kmin);
    if (!win) return 1;
    const int wl = ffs_lane(win);
    best.vid = __shfl_sync(kFull, vid, wl);
    best.og = __shfl_sync(kFull, og, wl);
    best.opos = __shfl_sync(kFull, (int)pos, wl) != 0;
    best.d = __shfl_sync(kFull, d, wl);
    sg.vid = vid, sg.o0 = o0, sg.bp = bp, sg.w = w;
    sg.base = (wl / L) * L;
    sg.obp = __shfl_sync(kFull, obp, wl);
    sg.valid = true;
    return 1;
}

// ExtendPathForward / ExtendPathBackward (blocksfinder.h:770-895)
__device__ __forceinline__ bool extend_path(Ctx &c, const bool FORWARD, int &best_size, long long &best_score, long long &now_score)
{
    Next nx;
    Staged sg;
    nx.vid = 0;
    sg.valid = false;
    for (int attempt = 0; attempt < 2 && nx.vid == 0 && !c.err; attempt++) { // forward retries with tryUsed (:782-785)
        if (attempt == 1 && !FORWARD) break;
        if (mpv_fast(c, FORWARD, attempt == 1, nx, sg)) {
            if (c.collect) c.ct.mpv_fast++;
        } else if (sg.valid = false, mpv_mid(c, FORWARD, attempt == 1, nx)) {
            if (c.collect) c.ct.mpv_mid++;
#ifdef LCB_CHECK_MPV
            const Next chk = most_popular_vertex(c, FORWARD, attempt == 1);
            if (chk.vid != nx.vid || (nx.vid != 0 && (chk.og != nx.og || chk.d != nx.d || chk.opos != nx.opos))) c.err = 90;
#endif
        } else {
            if (c.collect) c.ct.mpv_slow++;
            nx = most_popular_vertex(c, FORWARD, attempt == 1);
        }
    }
    if (c.err || nx.vid == 0) return false;
    bool success = false;
    const int step = (FORWARD == nx.opos) ? 1 : -1;
    int prev_v = FORWARD ? c.right_vertex : c.left_vertex; // vertex of the origin junction
    unsigned prev_bp = 0;
    bool have_prev_bp = false;
    if (sg.valid) prev_bp = sg.obp, have_prev_bp = true;
    // `for (it = origin; it.GetVertexId() != next; ++it) push(it.Outgoing/IngoingEdge())`: stops at the FIRST junction
    // of the walk that carries the chosen vertex (blocksfinder.h:789, :852)
    for (int j0 = sg.valid ? 1 : 0; j0 <= nx.d; j0 += 32) {
        int4 rc = make_int4(0, 0, 0, 0);
        if (!sg.valid) { // general path: re-read the junctions og .. og+step*d (cache-hot)
            int j = j0 + c.lane;
            if (j <= nx.d) rc = __ldg(c.ix.rec + (nx.og + step * j));
        }
        const int cnt = sg.valid ? nx.d : min(32, nx.d - j0 + 1);
        bool done = false;
        for (int t = 0; t < cnt && !done; t++) {
            const int jj = j0 + t;
            int v, o_first, o_count;
            unsigned bp;
            if (sg.valid) {
                const int ln = sg.base + jj - 1;
                v = __shfl_sync(kFull, sg.vid, ln);
                bp = __shfl_sync(kFull, sg.bp, ln);
                o_first = __shfl_sync(kFull, sg.o0, ln);
                o_count = (int)(__shfl_sync(kFull, sg.w, ln) >> 16);
            } else {
                const int idv = __shfl_sync(kFull, rc.x, t);
                bp = (unsigned)__shfl_sync(kFull, rc.y, t);
                o_first = __shfl_sync(kFull, rc.z, t);
                o_count = (int)((unsigned)__shfl_sync(kFull, rc.w, t) >> 16);
                v = nx.opos ? idv : -idv;
            }
            if (jj > 0 && have_prev_bp) {
                const int len = (int)(bp > prev_bp ? bp - prev_bp : prev_bp - bp);
                const int g_prev = nx.og + step * (jj - 1), g_now = nx.og + step * jj;
                const bool ok = path_push(c, FORWARD, v, len, FORWARD ? g_prev : g_now, nx.opos, prev_v, o_first, o_count);
                if (c.err) return false;
                success = ok;
                if (ok) {
                    now_score = path_score(c);
                    if (now_score > best_score) {
                        best_score = now_score;
                        best_size = (FORWARD ? c.nright : c.nleft) + 1;
                        if (now_score > 0) snapshot_best(c);
                        if (FORWARD) snapshot_state(c);
                    }
                }
                done = v == nx.vid;
            }
            prev_v = v;
            prev_bp = bp;
            have_prev_bp = true;
        }
        if (done || sg.valid) break;
    }
    return success;
}

// ProcessVertex::Process (blocksfinder.h:228-310).  On return c.best[0..nbest) is bestInstance and
// c.ar.rs()[0..nrs) the read-set.
__device__ __forceinline__ void process_seed(Ctx &c, int vid, unsigned char ch)
{
    c.hbig = false;
    c.hash = c.sm->hash;
    c.hmask = kHashSmem - 1;
    ctx_reset_storage(c);
    c.ninst = c.ngood = c.nbest = c.hcount = c.nrs = 0;
    c.nright = c.nleft = 0;
    for (int i = c.lane; i < kHashSmem; i += 32) c.hash[i] = make_int2(0, 0);
    __syncwarp();
    long long best_score = 0, score = 0;
    int best_size[2] = {1, 1}; // bestLeftSize, bestRightSize
    const int min_run = c.pr.b * 2;
    c.snap_valid = false;
    for (int phase = 1; phase >= 0; phase--) { // 1: forward, 0: backward
        const bool forward = phase == 1;
        if (forward) {
            path_init(c, vid, ch);
            if (c.err) return;
            if (c.ninst == 0) break; // a seed without live instances cannot move (MostPopularVertex finds nothing)
            snapshot_state(c);       // bestRightSize == 1: the state right after Init
        } else {
            const int replay = best_size[1] - 1;
            path_clear(c);
            if (c.err) return;
            if (c.snap_valid && c.snap_nright == replay && c.icap == kInstSmem) {
                restore_state(c);
            } else { // big paths: re-play the best right part (blocksfinder.h:271-284)
                path_init(c, vid, ch);
                for (int i = 0; i < replay && !c.err; i++) {
                    int4 e = c.ar.redge()[i];
                    path_push(c, true, e.x, e.y, e.z & 0x7FFFFFFF, e.z < 0, 0, -1, -1);
                }
                if (c.err) return;
            }
        }
        while (true) {
            bool ret = true, positive = false;
            const int prev_len = c.right_flank - c.left_flank;
            while ((ret = extend_path(c, forward, best_size[phase], best_score, score)) &&
                   (c.right_flank - c.left_flank) - prev_len <= min_run) {
                if (forward) positive = positive || score > 0; // backward: empty body, stray ';' at blocksfinder.h:297
            }
            if (!forward) positive = positive || score > 0;
            if (c.err) return;
            if (!ret || !positive) break;
        }
    }
    path_clear(c);
}

} // namespace lcb
