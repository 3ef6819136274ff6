// lcb_lean.cuh -- the COMMON CASE of the carving-path traversal as a small kernel body.
//
// lcb_traverse.cuh evaluates ProcessVertex::Process (SibeliaZ-LCB/blocksfinder.h:228-310) for any seed: thousands of path
// instances, vertices that occur many times on one chromosome, votes over thousands of vertices, arenas in HBM.  That
// generality costs ~11.7 k SASS instructions and 128 registers, and the profile of the north-star input shows what it
// buys: 26 % of the stall samples are instruction-cache misses (16 warps per SM at 16 different places of a 186 KB body)
// and ~1650 warp instructions per extension step (profiles/ncu_k_traverse_r2_baseline.md).
//
// This header is the same algorithm restricted to what nearly every seed of a few-genome input needs:
//   * at most 32 path instances, kept in shared memory (13-word records: a conflict-free stride);
//   * at most kLHash/2 + kLPath2 path vertices (the first 128 in a shared-memory hash, the rest of a long path in a
//     per-warp table in HBM; both with an insertion log, so that Path::Clear and the return to the best forward state
//     undo insertions instead of copying or zeroing a table);
//   * every pushed vertex occurs at most 32 times, at most once per chromosome (the lanes evaluate all occurrences at
//     once against the pre-push state, as push_parallel does);
//   * look-ahead votes over at most ~160 distinct vertices (any number of walks, four at a time, 24 junctions of each
//     per pass);
//   * |path distances| < 2^30 (so 32-bit arithmetic is exact), at most 63 chromosomes (offsets cached in shared memory).
// A seed that needs more makes process_seed return kBail after leaving the shared state clean; the caller evaluates it
// with the general code (lcb_traverse.cuh), which is the specification of everything here: same results, same
// bestInstance order, same read-set policy.  The multiset order of the reference (std::multiset keyed by compareIdx_,
// path.h:82-133) is kept in REGISTERS: lane p holds the id and key of the p-th instance in that order.
//
// Reference lines are cited next to each piece; lcb_traverse.cuh carries the long explanations.
#pragma once
#include "lcb_traverse.cuh"

namespace lcb {
namespace lean {

constexpr int kLInst = 32;   // instances (one lane each in the searches)
#ifndef LCB_LEAN_HASH
#define LCB_LEAN_HASH 256
#endif
constexpr int kLHash = LCB_LEAN_HASH; // path hash slots (half of them usable)
constexpr int kLVote = 256;     // vote table slots (multi-pass votes)
constexpr int kLVoteFast = 128; // ... of which the single-pass vote uses the first 128 and resolves by scanning them
constexpr int kLTiers = 3;     // multi-pass vote: 8 * kLTiers junctions of every walk per pass
constexpr int kLFastTiers = 2; // single-pass vote: walks of at most 16 junctions (the others take the multi-pass body)
constexpr int kLChr = 64;    // chr_off entries cached per CTA (C + 1 <= kLChr)
constexpr int kLHash2 = 8192; // second-level path hash in HBM (per warp): the vertices beyond the first kLHash / 2
constexpr int kLPath2 = kLHash2 / 2;
constexpr int kOk = 0, kBail = 1, kRetry = 3;
// reasons for handing an evaluation back (diagnostics: lcb_stats.lean_bail_why)
constexpr int kWhyOccurrences = 0, kWhyPathLength = 1, kWhySameChromosome = 2, kWhyInstances = 3, kWhyReadSet = 4, kWhyWalkDepth = 5,
              kWhyVote = 6, kWhyDistance = 7, kWhyCount = 8;

struct LInst { // Path::Instance (path.h:53-181); compareIdx_ = (flags & kPos) ? bg : fg is derived, not stored
    int fg, bg;        // front_/back_ : global record index
    int fv, bv;        // strand-signed vertex id at front/back
    unsigned fbp, bbp; // raw base-pair position at front/back
    int fdist, bdist;  // frontDistance_/backDistance_
    int rlo, rhi;      // read extent over epoch indices (rlo > rhi: empty)
    int clo, chi;      // chromosome bounds [clo, chi)
    unsigned flags;    // kPos | kFFin | kBFin
};
static_assert(sizeof(LInst) == 52, "13-word records");

struct LeanSmem {
    LInst inst[kLInst];
    int2 hash[kLHash];    // vertex -> path distance (DistanceKeeper, distancekeeper.h:9-41); key 0 = empty
    int2 vote[kLVote];    // vertex -> weight sum
    unsigned vlast[kLVote]; // vertex -> last (list position << 20 | depth) that voted for it
    int4 best[kLInst];    // bestInstance
    unsigned short hslot[kLHash / 2]; // slot of the i-th inserted vertex (undo log)
    unsigned char good[kLInst], s_good[kLInst];
    unsigned char elist[kLInst]; // list positions of the instances that sit on the path end (one look-ahead walk each)
    int best_scratch[8];         // result of the out-of-line multi-pass vote (DeepOut)
    unsigned char used[kLVote];  // slots of the vote table that hold a key (so that resolving and emptying it costs what the
                                 // vote had, not what the table could hold)
#ifdef LCB_TMA_WINDOWS
    // A/B variant (profiles/ab_tma_graph_r2.md): the look-ahead windows of a group of four walks are staged by bulk
    // copies (cp.async.bulk + mbarrier) instead of being read by one load per lane
    alignas(16) int4 w_rec[4][8 * kLTiers];
    alignas(16) uint32_t w_E[4][32];
    alignas(8) unsigned long long mbar;
#endif
};

struct LCtx { // warp-uniform unless noted
    const int4 *rec;
    const int2 *occ;
    const uint32_t *vtx_off;
    const uint32_t *E;
    const uint32_t *chr_off_s; // shared-memory copy of chr_off[0..C]
    int C;
    int b, m, flank, depth;
    uint32_t thresh;
    int lane;
    LeanSmem *sm;
    int2 *rs;   // read-set log of this warp (HBM)
    int2 *hash2;           // second-level path hash (HBM, all-empty between evaluations), kLHash2 slots
    unsigned short *hslot2; // its insertion log, kLPath2 entries
    LInst *shadow; // instances at the best forward point (HBM, written at every improvement and read once: restored
                   // instead of Clear + Init + re-push, blocksfinder.h:271-284)
    int rs_cap; // its capacity in intervals
    int why;    // why the last evaluation was handed back (kWhy*)
    int deep_bias; // 0..8: how often the recent look-ahead votes needed the multi-pass body
#ifdef LCB_TMA_WINDOWS
    unsigned tma_phase; // parity of the mbarrier's current phase
#endif
    int origin, right_vertex, left_vertex;
    int right_flank, left_flank;
    int nright, nleft;
    int ninst, ngood, nbest, hcount, nrs;
    int ordreg, keyreg; // PER LANE: id and key of the lane-th instance in multiset order (lane >= ninst: key = INT_MAX)
    int last_clo, last_chi; // PER LANE: bounds of the chromosome this lane looked up last (collinear genomes: the same one next time)
    // state at the best forward point
    int snap_ninst, snap_ngood, snap_hcount, snap_right_flank, snap_right_vertex, snap_nright, snap_ordreg, snap_keyreg;
};

__device__ __forceinline__ unsigned lanemask_lt(int lane) { return (1u << lane) - 1u; }

// chromosome bounds of record g from the shared-memory offsets
__device__ __forceinline__ void chr_bounds_s(const LCtx &c, int g, int &lo, int &hi)
{
    int a = 0, b = c.C; // off[a] <= g < off[b]
    while (b - a > 1) {
        const int mid = (a + b) >> 1;
        if ((int)c.chr_off_s[mid] <= g) a = mid;
        else b = mid;
    }
    lo = (int)c.chr_off_s[a];
    hi = (int)c.chr_off_s[a + 1];
}

// second level (long paths only): a real function, so that its probe loop is not repeated at every call site
__device__ __noinline__ int hash_find2(const int2 *hash2, int key)
{
    unsigned s = (hash_of(key) >> 10) & (unsigned)(kLHash2 - 1);
    while (true) {
        const int2 kv = hash2[s];
        if (kv.x == key) return kv.y;
        if (kv.x == 0) return kNotSet;
        s = (s + 1) & (unsigned)(kLHash2 - 1);
    }
}

// The first kLHash / 2 vertices of a path live in shared memory, the rest (long paths only) in the warp's table in HBM.
__device__ __forceinline__ int hash_find(const LCtx &c, int key) // per-lane key
{
    unsigned s = (hash_of(key) >> 12) & (unsigned)(kLHash - 1);
    while (true) {
        const int2 kv = c.sm->hash[s];
        if (kv.x == key) return kv.y;
        if (kv.x == 0) break;
        s = (s + 1) & (unsigned)(kLHash - 1);
    }
    if (c.hcount <= kLHash / 2) return kNotSet;
    return hash_find2(c.hash2, key);
}

// uniform key, known to be absent; false: the path outgrew the table
__device__ __forceinline__ bool hash_insert(LCtx &c, int key, int val)
{
    __syncwarp(); // every lane has finished probing before lane 0 changes the table
    if (c.hcount < kLHash / 2) {
        unsigned s = (hash_of(key) >> 12) & (unsigned)(kLHash - 1);
        while (c.sm->hash[s].x != 0) s = (s + 1) & (unsigned)(kLHash - 1);
        if (c.lane == 0) {
            c.sm->hash[s] = make_int2(key, val);
            c.sm->hslot[c.hcount] = (unsigned short)s;
        }
    } else {
        if (c.hcount - kLHash / 2 >= kLPath2) return false;
        unsigned s = (hash_of(key) >> 10) & (unsigned)(kLHash2 - 1);
        while (c.hash2[s].x != 0) s = (s + 1) & (unsigned)(kLHash2 - 1);
        if (c.lane == 0) {
            c.hash2[s] = make_int2(key, val);
            c.hslot2[c.hcount - kLHash / 2] = (unsigned short)s;
        }
    }
    c.hcount++;
    __syncwarp();
    return true;
}

// forget the vertices inserted after the first `keep`: newest first is not needed, the slots simply become empty again
// (they were empty before their insertion and nothing was inserted behind them that is kept)
__device__ __noinline__ void hash_erase(LeanSmem *sm, int2 *hash2, const unsigned short *hslot2, int from, int to, int lane)
{
    for (int i = from + lane; i < to; i += 32) {
        if (i < kLHash / 2) sm->hash[sm->hslot[i]].x = 0;
        else hash2[hslot2[i - kLHash / 2]].x = 0;
    }
    __syncwarp();
}
__device__ __forceinline__ void hash_truncate(LCtx &c, int keep)
{
    if (keep < c.hcount) hash_erase(c.sm, c.hash2, c.hslot2, keep, c.hcount, c.lane);
    c.hcount = keep;
}

__device__ __forceinline__ bool rs_add(LCtx &c, int lo, int hi) // uniform
{
    if (c.nrs >= c.rs_cap) return false;
    if (c.lane == 0) c.rs[c.nrs] = make_int2(lo, hi);
    c.nrs++;
    return true;
}

// Path::Clear (path.h:650-677): flush the per-instance read extents into the read-set log, empty the hash
__device__ __forceinline__ bool path_clear(LCtx &c)
{
    int lo = 0, hi = -1;
    if (c.lane < c.ninst) lo = c.sm->inst[c.lane].rlo, hi = c.sm->inst[c.lane].rhi;
    const bool live = lo <= hi;
    const unsigned m = __ballot_sync(kFull, live);
    const int n = __popc(m);
    const bool fits = c.nrs + n <= c.rs_cap;
    if (fits) {
        if (live) c.rs[c.nrs + __popc(m & lanemask_lt(c.lane))] = make_int2(lo, hi);
        c.nrs += n;
    }
    hash_truncate(c, 0);
    c.ninst = c.ngood = 0;
    c.nright = c.nleft = 0;
    c.ordreg = 0, c.keyreg = 0x7FFFFFFF;
    return fits;
}

// leave the shared state clean after a bail-out in the middle of an evaluation
__device__ __forceinline__ void abandon(LCtx &c)
{
    hash_truncate(c, 0);
    c.ninst = c.ngood = c.nbest = 0;
}

__device__ __forceinline__ void snapshot_state(LCtx &c)
{
    LeanSmem *sm = c.sm;
    const int *src = (const int *)sm->inst;
    int *dst = (int *)c.shadow;
    const int words = c.ninst * (int)(sizeof(LInst) / sizeof(int));
    for (int i = c.lane; i < words; i += 32) dst[i] = src[i];
    if (c.lane < c.ngood) sm->s_good[c.lane] = sm->good[c.lane];
    c.snap_ninst = c.ninst, c.snap_ngood = c.ngood, c.snap_hcount = c.hcount;
    c.snap_right_flank = c.right_flank, c.snap_right_vertex = c.right_vertex, c.snap_nright = c.nright;
    c.snap_ordreg = c.ordreg, c.snap_keyreg = c.keyreg;
    __syncwarp();
}

// replaces Clear + Init + re-push of the best right part (blocksfinder.h:271-284): the read extents of the instances
// that are dropped or rolled back still count (their epochs were read), so they are flushed first
__device__ __forceinline__ bool restore_state(LCtx &c)
{
    LeanSmem *sm = c.sm;
    {
        int lo = 0, hi = -1;
        if (c.lane < c.ninst) lo = sm->inst[c.lane].rlo, hi = sm->inst[c.lane].rhi;
        const bool live = lo <= hi;
        const unsigned m = __ballot_sync(kFull, live);
        const int n = __popc(m);
        if (c.nrs + n > c.rs_cap) return false;
        if (live) c.rs[c.nrs + __popc(m & lanemask_lt(c.lane))] = make_int2(lo, hi);
        c.nrs += n;
    }
    __syncwarp();
    const int *src = (const int *)c.shadow;
    int *dst = (int *)sm->inst;
    const int words = c.snap_ninst * (int)(sizeof(LInst) / sizeof(int));
    for (int i = c.lane; i < words; i += 32) dst[i] = src[i];
    if (c.lane < c.snap_ngood) sm->good[c.lane] = sm->s_good[c.lane];
    hash_truncate(c, c.snap_hcount);
    c.ninst = c.snap_ninst, c.ngood = c.snap_ngood;
    c.right_flank = c.snap_right_flank, c.right_vertex = c.snap_right_vertex, c.nright = c.snap_nright;
    c.left_flank = 0, c.left_vertex = c.origin, c.nleft = 0;
    c.ordreg = c.snap_ordreg, c.keyreg = c.snap_keyreg;
    __syncwarp();
    return true;
}

struct LOcc { // per-lane occurrence
    int g, clo, chi, flag;
    unsigned bp;
    bool pos, used;
};

__device__ __forceinline__ LOcc load_occurrence(LCtx &c, unsigned o, int vertex)
{
    LOcc r;
    const int2 oc = __ldg(c.occ + o);
    r.g = oc.x & 0x7FFFFFFF;
    r.bp = (unsigned)oc.y;
    r.pos = (oc.x < 0) == (vertex < 0); // JunctionIterator::IsPositiveStrand (junctionstorage.h:408-411)
    if (r.g >= c.last_clo && r.g < c.last_chi) r.clo = c.last_clo, r.chi = c.last_chi;
    else chr_bounds_s(c, r.g, r.clo, r.chi);
    const bool has = r.pos || r.g > r.clo; // IsUsed on the - strand at idx 0 is false (junctionstorage.h:277-282)
    r.flag = has ? (r.pos ? r.g : r.g - 1) : -1;
    r.used = has ? (__ldg(c.E + r.flag) < c.thresh) : false;
    c.last_clo = r.clo, c.last_chi = r.chi;
    return r;
}

__device__ __forceinline__ void write_new_instance(LInst &I, const LOcc &q, int v, int dist)
{
    I.fg = I.bg = q.g;
    I.fv = I.bv = v;
    I.fbp = I.bbp = q.bp;
    I.fdist = I.bdist = dist;
    I.rlo = q.flag >= 0 ? q.flag : 0x7FFFFFFF;
    I.rhi = q.flag >= 0 ? q.flag : -1;
    I.clo = q.clo, I.chi = q.chi;
    I.flags = q.pos ? kPos : 0u;
}

// Path::Init (path.h:33-46): one instance per unused occurrence of the seed vertex whose edge character matches; the
// occurrence list is sorted by (chr, idx), so creation order == multiset order
__device__ __forceinline__ int path_init(LCtx &c, int vid, unsigned char ch)
{
    c.origin = c.right_vertex = c.left_vertex = vid;
    c.right_flank = c.left_flank = 0;
    c.ordreg = 0, c.keyreg = 0x7FFFFFFF;
    hash_insert(c, vid, 0);
    const int av = vid < 0 ? -vid : vid;
    const unsigned o0 = __ldg(c.vtx_off + av), o1 = __ldg(c.vtx_off + av + 1);
    if (o1 - o0 > 32u) return c.why = kWhyOccurrences, kBail;
    const bool live = (unsigned)c.lane < o1 - o0;
    LOcc q;
    q.g = 0, q.bp = 0, q.pos = false, q.used = false, q.flag = -1, q.clo = 0, q.chi = 0;
    bool match = false;
    if (live) {
        q = load_occurrence(c, o0 + (unsigned)c.lane, vid);
        const int4 rr = __ldg(c.rec + q.g);
        match = (q.pos ? rec_next_ch(rr) : rec_prev_rc(rr)) == ch; // seqIt.GetChar(), junctionstorage.h:234-243
    }
    const unsigned lt = lanemask_lt(c.lane);
    const unsigned um = __ballot_sync(kFull, match && q.used); // outcome depends on these epochs although no instance is born
    if (um) {
        if (c.nrs + __popc(um) > c.rs_cap) return c.why = kWhyReadSet, kBail;
        if (match && q.used) c.rs[c.nrs + __popc(um & lt)] = make_int2(q.flag, q.flag);
        c.nrs += __popc(um);
    }
    const unsigned nm = __ballot_sync(kFull, match && !q.used);
    if (match && !q.used) write_new_instance(c.sm->inst[__popc(nm & lt)], q, vid, 0);
    c.ninst = __popc(nm);
    __syncwarp();
    if (c.lane < c.ninst) c.ordreg = c.lane, c.keyreg = c.sm->inst[c.lane].fg;
    return kOk;
}

// Path::PointPushBack / PointPushFront with their workers (path.h:430-602) for a vertex whose occurrences (<= 32) lie on
// distinct chromosomes.  `back`: BACK push, v = e.GetEndVertex(); else FRONT, v = e.GetStartVertex().  e_ch_g/e_ch_pos
// locate the junction whose char is e.GetChar(); e_other is e.GetEndVertex() for FRONT.
// Returns 0 pushed, 1 vertex already in the path (nothing happened), 2 bail (state untouched except the hash entry).
__device__ __forceinline__ int path_push(LCtx &c, const bool back, int v, int len, int e_ch_g, bool e_ch_pos, int e_other,
                                         unsigned o0, unsigned cnt)
{
    if (hash_find(c, v) != kNotSet) return 1;
    const int dist = back ? c.right_flank + len : c.left_flank - len;
    if (cnt > 32u) return c.why = kWhyOccurrences, 2;
    if (dist >= (1 << 30) || dist <= -(1 << 30)) return c.why = kWhyDistance, 2;
    if (!hash_insert(c, v, dist)) return c.why = kWhyPathLength, 2;
    const bool live = (unsigned)c.lane < cnt;
    LOcc q;
    q.g = 0, q.bp = 0, q.pos = false, q.used = false, q.flag = -1, q.clo = -1 - c.lane, q.chi = 0;
    if (live) q = load_occurrence(c, o0 + (unsigned)c.lane, v);
    { // two occurrences on one chromosome are neighbours in the (chr, idx)-sorted list: not for this code
        const int prev_clo = __shfl_up_sync(kFull, q.clo, 1);
        if (__any_sync(kFull, live && c.lane > 0 && prev_clo == q.clo)) {
            hash_truncate(c, c.hcount - 1);
            return c.why = kWhySameChromosome, 2;
        }
    }
    // multiset neighbours: position of the first key > g, by one ballot per occurrence over the sorted key registers
    const int n = c.ninst;
    unsigned my_gt = 0;
    for (unsigned o = 0; o < cnt; o++) {
        const int go = __shfl_sync(kFull, q.g, (int)o);
        const unsigned gt = __ballot_sync(kFull, c.keyreg > go); // lanes >= n hold INT_MAX: first of them = n
        if ((unsigned)c.lane == o) my_gt = gt;
    }
    const int ub = min(my_gt ? ffs_lane(my_gt) : 32, n);
    const int hs = min(ub, 31), ls = max(ub - 1, 0);
    const int hi_cand = __shfl_sync(kFull, c.ordreg, hs), hi_key = __shfl_sync(kFull, c.keyreg, hs);
    const int lo_cand = __shfl_sync(kFull, c.ordreg, ls), lo_key = __shfl_sync(kFull, c.keyreg, ls);
    int outcome = 0, cand = -1, scan_lo = 0, scan_hi = -1; // 0 skip, 1 extend, 2 new instance, 3 found used
    if (live) {
        int hi_id = -1, lo_id = -1;
        if (ub < n && hi_key < q.chi) hi_id = hi_cand;
        if (ub > 0 && lo_key >= q.clo) lo_id = lo_cand;
        bool within = false;
        if (hi_id >= 0) {
            const int a = c.sm->inst[hi_id].fg, bb = c.sm->inst[hi_id].bg;
            within = q.g >= min(a, bb) && q.g <= max(a, bb);
        }
        if (!within) {
            cand = (q.pos == back) ? lo_id : hi_id;
            bool extend = false;
            int cend_v = 0;
            if (cand >= 0) {
                const LInst &I = c.sm->inst[cand];
                const bool cpos = (I.flags & kPos) != 0;
                const int cg = back ? I.bg : I.fg;
                const unsigned cbp = back ? I.bbp : I.fbp;
                const int cdist = back ? I.bdist : I.fdist;
                cend_v = back ? I.bv : I.fv;
                if (cpos == q.pos) { // Compatible (path.h:380-428): pure tests first, the used scan last
                    // realDiff = strand-aware (end.pos - start.pos): the larger operand must be x
                    const bool fwd = back == q.pos;
                    const unsigned x = fwd ? q.bp : cbp, y = fwd ? cbp : q.bp;
                    bool ok = x >= y;
                    const unsigned rd = x - y;
                    const int ad = back ? dist - cdist : cdist - dist;
                    if (ok && (rd > (unsigned)c.b || ad > c.b)) {
                        const int step = q.pos ? 1 : -1;
                        ok = back ? (q.g == cg + step) : (cg == q.g + step);
                        if (ok) {
                            const int4 ce = __ldg(c.rec + e_ch_g), cs = __ldg(c.rec + (back ? cg : q.g));
                            ok = (q.pos ? rec_next_ch(cs) : rec_prev_rc(cs)) == (e_ch_pos ? rec_next_ch(ce) : rec_prev_rc(ce));
                            if (!back) ok = ok && I.fv == e_other;
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
    const unsigned lt = lanemask_lt(c.lane);
    const unsigned nm = __ballot_sync(kFull, outcome == 2);
    if (c.ninst + __popc(nm) > kLInst) { // more instances than lanes: the general code's business
        hash_truncate(c, c.hcount - 1);
        return c.why = kWhyInstances, 2;
    }
    const unsigned om = __ballot_sync(kFull, outcome == 3);
    if (c.nrs + __popc(om) > c.rs_cap) {
        hash_truncate(c, c.hcount - 1);
        return c.why = kWhyReadSet, 2;
    }
    __syncwarp(); // every lane has finished reading the instance table before any lane changes it
    // ---- apply.  Candidates of different lanes are different instances (different chromosomes).
    bool newly_good = false, key_moves = false;
    if (live && cand >= 0) {
        LInst &I = c.sm->inst[cand];
        if (scan_lo <= scan_hi) {
            if (scan_lo < I.rlo) I.rlo = scan_lo;
            if (scan_hi > I.rhi) I.rhi = scan_hi;
        }
        const unsigned fin = back ? kBFin : kFFin;
        if (outcome == 1 && !(I.flags & fin)) {
            unsigned a = I.fbp > I.bbp ? I.fbp - I.bbp : I.bbp - I.fbp;
            const bool prev_good = a >= (unsigned)c.m;
            if (back) I.bg = q.g, I.bv = v, I.bbp = q.bp, I.bdist = dist;
            else I.fg = q.g, I.fv = v, I.fbp = q.bp, I.fdist = dist;
            key_moves = back == q.pos; // compareIdx_ follows the back of a + instance / the front of a - instance
            a = I.fbp > I.bbp ? I.fbp - I.bbp : I.bbp - I.fbp;
            newly_good = !prev_good && a >= (unsigned)c.m;
            if (q.flag >= 0) {
                if (q.flag < I.rlo) I.rlo = q.flag;
                if (q.flag > I.rhi) I.rhi = q.flag;
            }
            if (q.used) I.flags |= fin;
        }
    }
    const unsigned gm = __ballot_sync(kFull, newly_good);
    if (newly_good) c.sm->good[c.ngood + __popc(gm & lt)] = (unsigned char)cand;
    c.ngood += __popc(gm);
    if (om) {
        if (outcome == 3) c.rs[c.nrs + __popc(om & lt)] = make_int2(q.flag, q.flag);
        c.nrs += __popc(om);
    }
    // keys that moved: the lane that holds the instance in the order follows (the order itself cannot change, see
    // ord_upper_bound in lcb_traverse.cuh)
    unsigned km = __ballot_sync(kFull, key_moves);
    while (km) {
        const int src = ffs_lane(km);
        km &= km - 1;
        const int id = __shfl_sync(kFull, cand, src), key = __shfl_sync(kFull, q.g, src);
        if (c.ordreg == id && c.lane < c.ninst) c.keyreg = key;
    }
    // new instances: allInstance_ order == occurrence order; each goes behind the keys <= its own
    if (outcome == 2) write_new_instance(c.sm->inst[c.ninst + __popc(nm & lt)], q, v, dist);
    unsigned left = nm;
    int added = 0;
    while (left) {
        const int src = ffs_lane(left);
        left &= left - 1;
        const int at = __shfl_sync(kFull, ub, src) + added; // earlier new instances have smaller keys: they sit below
        const int key = __shfl_sync(kFull, q.g, src);
        const int up_ord = __shfl_up_sync(kFull, c.ordreg, 1), up_key = __shfl_up_sync(kFull, c.keyreg, 1);
        if (c.lane > at) c.ordreg = up_ord, c.keyreg = up_key;
        if (c.lane == at) c.ordreg = c.ninst + added, c.keyreg = key;
        added++;
    }
    c.ninst += added;
    if (back) {
        c.nright++;
        c.right_flank = dist;
        c.right_vertex = v;
    } else {
        c.nleft++;
        c.left_flank = dist;
        c.left_vertex = v;
    }
    __syncwarp();
    return 0;
}

// Path::Score (path.h:604-628).  Penalties below maxFlankingSize <= 32767 keep every square below 2^32; the sum of the
// real lengths is split in 16-bit halves: four exact 32-bit reductions.
__device__ __forceinline__ long long path_score(const LCtx &c)
{
    unsigned real = 0, pen2 = 0;
    bool bad = false;
    if (c.lane < c.ngood) {
        const LInst &I = c.sm->inst[c.sm->good[c.lane]];
        real = I.fbp > I.bbp ? I.fbp - I.bbp : I.bbp - I.fbp;
        const int rp = c.right_flank - I.bdist, lp = I.fdist - c.left_flank;
        if (lp >= c.flank || rp >= c.flank) bad = true;
        else {
            const unsigned pen = (unsigned)(rp + lp); // 0 <= pen < 2 * flank <= 65534
            pen2 = pen * pen;
        }
    }
    const unsigned lo = __reduce_add_sync(kFull, real & 0xFFFFu), hi = __reduce_add_sync(kFull, real >> 16);
    const unsigned pp = __reduce_add_sync(kFull, pen2 >> 5), pr = __reduce_add_sync(kFull, pen2 & 31u);
    bad = __any_sync(kFull, bad);
    const long long total = (long long)lo + ((long long)hi << 16) - (((long long)pp << 5) + pr);
    return bad ? -(long long)0x7FFFFFFF : total;
}

// bestInstance = copies of *goodInstance_[i] in list order (blocksfinder.h:818-825, :881-888)
__device__ __forceinline__ void snapshot_best(LCtx &c)
{
    if (c.lane < c.ngood) {
        const LInst &I = c.sm->inst[c.sm->good[c.lane]];
        c.sm->best[c.lane] = make_int4(I.fg | ((I.flags & kPos) ? (int)0x80000000 : 0), I.bg, (int)I.fbp, (int)I.bbp);
    }
    c.nbest = c.ngood;
    __syncwarp();
}


#if defined(LCB_TMA_WINDOWS) && defined(__CUDACC__)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "LCB_WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra LCB_DONE_%=;\n"
                 "bra LCB_WAIT_%=;\n"
                 "LCB_DONE_%=:\n"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
#endif

// BlocksFinder::MostPopularVertex (blocksfinder.h:708-768): four look-ahead walks at a time (8 lanes each, kLTiers depths
// per lane, all loads of a pass in flight at once; a walk longer than that goes on in further passes), votes in the
// shared-memory table, closed-form resolution over the slots the vote used (see most_popular_vertex in lcb_traverse.cuh
// for why the running arg-max has a closed form).
// Returns kBail when the vote has too many distinct vertices for the table (left empty).
__device__ __forceinline__ int most_popular_vertex_fast(LCtx &c, bool forward, bool try_used, Next &best)
{
    LeanSmem *sm = c.sm;
    best.vid = 0, best.og = 0, best.d = 0, best.opos = false;
    const int start_vid = forward ? c.right_vertex : c.left_vertex;
    const bool use_good = c.ngood >= 2;
    const int n = use_good ? c.ngood : c.ninst;
    int my_id = 0;
    bool elig = false;
    if (c.lane < n) {
        my_id = use_good ? (int)sm->good[c.lane] : c.lane;
        elig = (forward ? sm->inst[my_id].bv : sm->inst[my_id].fv) == start_vid;
    }
    const unsigned em = __ballot_sync(kFull, elig);
    const int E = __popc(em);
    if (E == 0) return kOk;
    if (elig) sm->elist[__popc(em & lanemask_lt(c.lane))] = (unsigned char)c.lane;
    __syncwarp();
    const int k = c.lane >> 3, dd = c.lane & 7;
    int distinct = 0;
    bool fail = false;
    for (int gb = 0; gb < E && !fail; gb += 4) {
        const bool lane_on = gb + k < E;
        const int q = lane_on ? (int)sm->elist[gb + k] : 0; // list position of my walk's instance
        const int id = use_good ? (int)sm->good[q] : q;
        const LInst &I = sm->inst[id];
        const bool pos = (I.flags & kPos) != 0;
        const int og = forward ? I.bg : I.fg;
        const unsigned obp = forward ? I.bbp : I.fbp;
        const unsigned weight = (I.fbp > I.bbp ? I.fbp - I.bbp : I.bbp - I.fbp) + 1u;
        const int clo = I.clo, chi = I.chi;
        const int step = (forward == pos) ? 1 : -1;
        const unsigned seg = 0xFFu << (k * 8);
        int vid[kLFastTiers], flag[kLFastTiers];
        bool inr[kLFastTiers], ok[kLFastTiers], inpath[kLFastTiers];
        {
            int4 rc[kLFastTiers];
            uint32_t ep[kLFastTiers];
#if defined(LCB_TMA_WINDOWS) && defined(__CUDACC__)
            // window of the walk: records [a, b) of [wa, wa + 24) inside the chromosome; its epochs: 32 entries from ea
            const int wa = step > 0 ? og + 1 : og - 8 * kLFastTiers;
            const int a = max(wa, clo), b = min(wa + 8 * kLFastTiers, chi);
            const int ea = max(a - 1, 0) & ~3;
            const bool leader = lane_on && dd == 0 && a < b;
            const unsigned my_bytes = leader ? (unsigned)(b - a) * 16u + (try_used ? 0u : 128u) : 0u;
            const unsigned all_bytes = __reduce_add_sync(kFull, my_bytes);
            if (all_bytes) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // earlier generic reads of the buffers before the async writes
                if (c.lane == 0) mbar_expect(&sm->mbar, all_bytes);
                __syncwarp();
                if (leader) {
                    bulk_g2s(&sm->w_rec[k][a - wa], c.rec + a, (unsigned)(b - a) * 16u, &sm->mbar);
                    if (!try_used) bulk_g2s(&sm->w_E[k][0], c.E + ea, 128u, &sm->mbar);
                }
                mbar_wait(&sm->mbar, c.tma_phase);
                c.tma_phase ^= 1u;
            }
#pragma unroll
            for (int t = 0; t < kLFastTiers; t++) {
                const int g = og + step * (t * 8 + dd + 1);
                inr[t] = lane_on && g >= clo && g < chi; // it.Valid()
                const bool has = pos || g > clo;
                flag[t] = (inr[t] && has) ? (pos ? g : g - 1) : -1;
                rc[t] = make_int4(0, 0, 0, 0);
                ep[t] = kFree;
                if (inr[t]) rc[t] = sm->w_rec[k][g - wa];
                if (flag[t] >= 0 && !try_used) ep[t] = sm->w_E[k][flag[t] - ea];
            }
#else
#pragma unroll
            for (int t = 0; t < kLFastTiers; t++) { // every load of every walk is in flight before the first one is used
                const int g = og + step * (t * 8 + dd + 1);
                inr[t] = lane_on && g >= clo && g < chi; // it.Valid()
                const bool has = pos || g > clo;
                flag[t] = (inr[t] && has) ? (pos ? g : g - 1) : -1;
                rc[t] = make_int4(0, 0, 0, 0);
                ep[t] = kFree;
                if (inr[t]) rc[t] = __ldg(c.rec + g);
                if (flag[t] >= 0 && !try_used) ep[t] = __ldg(c.E + flag[t]);
            }
#endif
#pragma unroll
            for (int t = 0; t < kLFastTiers; t++) {
                const int d = t * 8 + dd + 1;
                vid[t] = 0;
                inpath[t] = false;
                bool used = false;
                if (inr[t]) {
                    if (t == 0) prefetch_l1(c.occ + rc[t].z); // the push of this vertex starts with its occurrence list
                    vid[t] = pos ? rc[t].x : -rc[t].x;
                    const unsigned bp = (unsigned)rc[t].y;
                    const unsigned dp = bp > obp ? bp - obp : obp - bp;
                    inr[t] = d < c.depth || dp <= (unsigned)c.b;
                }
                if (inr[t]) {
                    used = flag[t] >= 0 && ep[t] < c.thresh; // ep stays kFree under try_used
                    inpath[t] = hash_find(c, vid[t]) != kNotSet;
                } else {
                    flag[t] = -1;
                }
                ok[t] = inr[t] && !inpath[t] && !used;
            }
        }
        // first junction of every walk that ends it
        int nok = 8 * kLFastTiers;
#pragma unroll
        for (int t = kLFastTiers - 1; t >= 0; t--) {
            const unsigned f = __ballot_sync(kFull, lane_on && !ok[t]) & seg;
            if (f) nok = t * 8 + ffs_lane(f) - k * 8;
        }
        if (__any_sync(kFull, lane_on && nok == 8 * kLFastTiers)) { // a walk needs more depth than its lanes offer
            fail = true;
            break;
        }
        int mylo = 0x7FFFFFFF, myhi = -1;
#pragma unroll
        for (int t = 0; t < kLFastTiers; t++) {
            const int di = t * 8 + dd; // 0-based depth
            const bool active = lane_on && di < nok;
            const bool stop_in_body = lane_on && di == nok && inr[t];
            const bool dep = flag[t] >= 0 && !try_used && (active || (stop_in_body && !inpath[t]));
            if (dep) mylo = min(mylo, flag[t]), myhi = max(myhi, flag[t]);
            if (active) { // count[vid] += weight; remember the last (list position, depth) that touched it
                unsigned x = (hash_of(vid[t]) >> 12) & (unsigned)(kLVoteFast - 1);
                while (true) {
                    const int old = atomicCAS(&sm->vote[x].x, 0, vid[t]);
                    if (old == 0 || old == vid[t]) {
                        if (old == 0) distinct++; // per-lane count, summed below
                        break;
                    }
                    x = (x + 1) & (unsigned)(kLVoteFast - 1);
                }
                atomicAdd((unsigned *)&sm->vote[x].y, weight);
                atomicMax(&sm->vlast[x], ((unsigned)q << 20) | (unsigned)(di + 1));
            }
        }
        // the epochs each walk depended on (its leader lane owns the instance record; walks have distinct instances)
        {
            int lo = mylo, hi = myhi;
#pragma unroll
            for (int s = 1; s < 8; s <<= 1) {
                lo = min(lo, __shfl_xor_sync(kFull, lo, s));
                hi = max(hi, __shfl_xor_sync(kFull, hi, s));
            }
            if (lane_on && dd == 0 && lo <= hi) {
                LInst &W = sm->inst[id];
                if (lo < W.rlo) W.rlo = lo;
                if (hi > W.rhi) W.rhi = hi;
            }
        }
        // the next group inserts up to 4 * 8 * kLFastTiers keys: go on only while its probing is certain to find empty slots
        if (gb + 4 < E) {
            const int total_distinct = (int)__reduce_add_sync(kFull, (unsigned)distinct);
            if (total_distinct + 4 * 8 * kLFastTiers > kLVoteFast - 1) fail = true;
        }
        __syncwarp();
    }
    // ---- resolve (and leave the table empty): among the vertices with the maximal final count, the one whose LAST
    // increment came from the smallest origin (- strand first, then (chr, idx)), earliest event on ties
    int2 e[kLVoteFast / 32];
    unsigned ev[kLVoteFast / 32];
    unsigned mymax = 0;
#pragma unroll
    for (int r = 0; r < kLVoteFast / 32; r++) {
        e[r] = sm->vote[r * 32 + c.lane];
        ev[r] = sm->vlast[r * 32 + c.lane];
        if (e[r].x != 0) {
            sm->vote[r * 32 + c.lane] = make_int2(0, 0);
            sm->vlast[r * 32 + c.lane] = 0u;
            mymax = max(mymax, (unsigned)e[r].y);
        }
    }
    __syncwarp();
    if (fail) return kRetry;
    const unsigned M = __reduce_max_sync(kFull, mymax);
    if (M == 0) return kOk;
    unsigned my_okey = 0xFFFFFFFFu, my_ev = 0xFFFFFFFFu;
    int my_vid = 0;
#pragma unroll
    for (int r = 0; r < kLVoteFast / 32; r++) {
        if (e[r].x != 0 && (unsigned)e[r].y == M) {
            const int q = (int)(ev[r] >> 20);
            const LInst &I = sm->inst[use_good ? (int)sm->good[q] : q];
            const unsigned okey = ((I.flags & kPos) ? 0x80000000u : 0u) | (unsigned)(forward ? I.bg : I.fg);
            if (okey < my_okey || (okey == my_okey && ev[r] < my_ev)) my_okey = okey, my_ev = ev[r], my_vid = e[r].x;
        }
    }
    const unsigned kmin = __reduce_min_sync(kFull, my_okey);
    const unsigned emin = __reduce_min_sync(kFull, my_okey == kmin ? my_ev : 0xFFFFFFFFu);
    // (okey, event) names one (walk, depth) item, hence one vertex: at most one lane matches
    const unsigned win = __ballot_sync(kFull, my_okey == kmin && my_ev == emin);
    const int wl = ffs_lane(win);
    best.vid = __shfl_sync(kFull, my_vid, wl);
    best.og = (int)(kmin & 0x7FFFFFFFu);
    best.opos = (kmin >> 31) != 0;
    best.d = (int)(emin & 0xFFFFFu);
    return kOk;
}

__device__ __forceinline__ int most_popular_vertex_deep(LCtx &c, bool forward, bool try_used, Next &best)
{
    LeanSmem *sm = c.sm;
    best.vid = 0, best.og = 0, best.d = 0, best.opos = false;
    const int start_vid = forward ? c.right_vertex : c.left_vertex;
    const bool use_good = c.ngood >= 2;
    const int n = use_good ? c.ngood : c.ninst;
    int my_id = 0;
    bool elig = false;
    if (c.lane < n) {
        my_id = use_good ? (int)sm->good[c.lane] : c.lane;
        elig = (forward ? sm->inst[my_id].bv : sm->inst[my_id].fv) == start_vid;
    }
    const unsigned em = __ballot_sync(kFull, elig);
    const int E = __popc(em);
    if (E == 0) return kOk;
    if (elig) sm->elist[__popc(em & lanemask_lt(c.lane))] = (unsigned char)c.lane;
    __syncwarp();
    const int k = c.lane >> 3, dd = c.lane & 7;
    bool fail = false;
    int nused = 0; // keys in the vote table (it is empty between two calls)
    for (int gb = 0; gb < E && !fail; gb += 4) {
        const bool lane_on = gb + k < E;
        const int q = lane_on ? (int)sm->elist[gb + k] : 0; // list position of my walk's instance
        const int id = use_good ? (int)sm->good[q] : q;
        const LInst &I = sm->inst[id];
        const bool pos = (I.flags & kPos) != 0;
        const int og = forward ? I.bg : I.fg;
        const unsigned obp = forward ? I.bbp : I.fbp;
        const unsigned weight = (I.fbp > I.bbp ? I.fbp - I.bbp : I.bbp - I.fbp) + 1u;
        const int clo = I.clo, chi = I.chi;
        const int step = (forward == pos) ? 1 : -1;
        const unsigned seg = 0xFFu << (k * 8);
        int mylo = 0x7FFFFFFF, myhi = -1;
        bool pending = lane_on; // my walk has not met its end yet
        // a pass covers 8 * kLTiers junctions of every pending walk; nearly always one pass is all there is
        for (int d0 = 0;; d0 += 8 * kLTiers) {
            // a pass inserts at most 4 * 8 * kLTiers keys: go on only while its probing is certain to find empty slots
            if (nused + 4 * 8 * kLTiers > kLVote - 1) {
                fail = true;
                c.why = kWhyVote;
                break;
            }
            int vid[kLTiers], flag[kLTiers];
            bool inr[kLTiers], ok[kLTiers], inpath[kLTiers];
            {
                int4 rc[kLTiers];
                uint32_t ep[kLTiers];
#if defined(LCB_TMA_WINDOWS) && defined(__CUDACC__)
                // window of the walk: records [a, b) of [wa, wa + 24) inside the chromosome; its epochs: 32 entries from ea
                const int wa = step > 0 ? og + d0 + 1 : og - d0 - 8 * kLTiers;
                const int a = max(wa, clo), b = min(wa + 8 * kLTiers, chi);
                const int ea = max(a - 1, 0) & ~3;
                const bool leader = pending && dd == 0 && a < b;
                const unsigned my_bytes = leader ? (unsigned)(b - a) * 16u + (try_used ? 0u : 128u) : 0u;
                const unsigned all_bytes = __reduce_add_sync(kFull, my_bytes);
                if (all_bytes) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // earlier generic reads of the buffers before the async writes
                    if (c.lane == 0) mbar_expect(&sm->mbar, all_bytes);
                    __syncwarp();
                    if (leader) {
                        bulk_g2s(&sm->w_rec[k][a - wa], c.rec + a, (unsigned)(b - a) * 16u, &sm->mbar);
                        if (!try_used) bulk_g2s(&sm->w_E[k][0], c.E + ea, 128u, &sm->mbar);
                    }
                    mbar_wait(&sm->mbar, c.tma_phase);
                    c.tma_phase ^= 1u;
                }
#pragma unroll
                for (int t = 0; t < kLTiers; t++) {
                    const int g = og + step * (d0 + t * 8 + dd + 1);
                    inr[t] = pending && g >= clo && g < chi; // it.Valid()
                    const bool has = pos || g > clo;
                    flag[t] = (inr[t] && has) ? (pos ? g : g - 1) : -1;
                    rc[t] = make_int4(0, 0, 0, 0);
                    ep[t] = kFree;
                    if (inr[t]) rc[t] = sm->w_rec[k][g - wa];
                    if (flag[t] >= 0 && !try_used) ep[t] = sm->w_E[k][flag[t] - ea];
                }
#else
#pragma unroll
                for (int t = 0; t < kLTiers; t++) { // every load of every walk is in flight before the first one is used
                    const int g = og + step * (d0 + t * 8 + dd + 1);
                    inr[t] = pending && g >= clo && g < chi; // it.Valid()
                    const bool has = pos || g > clo;
                    flag[t] = (inr[t] && has) ? (pos ? g : g - 1) : -1;
                    rc[t] = make_int4(0, 0, 0, 0);
                    ep[t] = kFree;
                    if (inr[t]) rc[t] = __ldg(c.rec + g);
                    if (flag[t] >= 0 && !try_used) ep[t] = __ldg(c.E + flag[t]);
                }
#endif
#pragma unroll
                for (int t = 0; t < kLTiers; t++) {
                    const int d = d0 + t * 8 + dd + 1;
                    vid[t] = 0;
                    inpath[t] = false;
                    bool used = false;
                    if (inr[t]) {
                        if (t == 0 && d0 == 0) prefetch_l1(c.occ + rc[t].z); // the push of this vertex starts with its occurrence list
                        vid[t] = pos ? rc[t].x : -rc[t].x;
                        const unsigned bp = (unsigned)rc[t].y;
                        const unsigned dp = bp > obp ? bp - obp : obp - bp;
                        inr[t] = d < c.depth || dp <= (unsigned)c.b;
                    }
                    if (inr[t]) {
                        used = flag[t] >= 0 && ep[t] < c.thresh; // ep stays kFree under try_used
                        inpath[t] = hash_find(c, vid[t]) != kNotSet;
                    } else {
                        flag[t] = -1;
                    }
                    ok[t] = inr[t] && !inpath[t] && !used;
                }
            }
            // first junction of this pass that ends the walk (8 * kLTiers: none, the walk goes on in the next pass)
            int nok = 8 * kLTiers;
#pragma unroll
            for (int t = kLTiers - 1; t >= 0; t--) {
                const unsigned f = __ballot_sync(kFull, pending && !ok[t]) & seg;
                if (f) nok = t * 8 + ffs_lane(f) - k * 8;
            }
#pragma unroll
            for (int t = 0; t < kLTiers; t++) {
                const int di = t * 8 + dd; // 0-based depth inside the pass
                const bool active = pending && di < nok;
                const bool stop_in_body = pending && di == nok && inr[t];
                const bool dep = flag[t] >= 0 && !try_used && (active || (stop_in_body && !inpath[t]));
                if (dep) mylo = min(mylo, flag[t]), myhi = max(myhi, flag[t]);
                bool fresh = false;
                unsigned x = 0;
                if (active) { // count[vid] += weight; remember the last (list position, depth) that touched it
                    x = (hash_of(vid[t]) >> 12) & (unsigned)(kLVote - 1);
                    while (true) {
                        const int old = atomicCAS(&sm->vote[x].x, 0, vid[t]);
                        fresh = old == 0;
                        if (old == 0 || old == vid[t]) break;
                        x = (x + 1) & (unsigned)(kLVote - 1);
                    }
                    atomicAdd((unsigned *)&sm->vote[x].y, weight);
                    atomicMax(&sm->vlast[x], ((unsigned)q << 20) | (unsigned)(d0 + di + 1));
                }
                const unsigned fm = __ballot_sync(kFull, fresh); // new keys: remember their slots
                if (fresh) sm->used[nused + __popc(fm & lanemask_lt(c.lane))] = (unsigned char)x;
                nused += __popc(fm);
            }
            pending = pending && nok == 8 * kLTiers;
            __syncwarp();
            if (!__any_sync(kFull, pending)) break;
        }
        // the epochs each walk depended on (its leader lane owns the instance record; walks have distinct instances)
        {
            int lo = mylo, hi = myhi;
#pragma unroll
            for (int s = 1; s < 8; s <<= 1) {
                lo = min(lo, __shfl_xor_sync(kFull, lo, s));
                hi = max(hi, __shfl_xor_sync(kFull, hi, s));
            }
            if (lane_on && dd == 0 && lo <= hi) {
                LInst &W = sm->inst[id];
                if (lo < W.rlo) W.rlo = lo;
                if (hi > W.rhi) W.rhi = hi;
            }
        }
        __syncwarp();
    }
    // ---- resolve over the slots that were used (and leave the table empty): among the vertices with the maximal final
    // count, the one whose LAST increment came from the smallest origin (- strand first, then (chr, idx)), earliest event
    // on ties
    __syncwarp();
    unsigned M = 0;
    for (int base = 0; base < nused; base += 32) {
        unsigned cnt = 0;
        if (base + c.lane < nused) cnt = (unsigned)sm->vote[sm->used[base + c.lane]].y;
        M = max(M, __reduce_max_sync(kFull, cnt));
    }
    unsigned my_okey = 0xFFFFFFFFu, my_ev = 0xFFFFFFFFu;
    int my_vid = 0;
    for (int base = 0; base < nused; base += 32) {
        if (base + c.lane < nused) {
            const int x = sm->used[base + c.lane];
            const int2 e = sm->vote[x];
            const unsigned ev = sm->vlast[x];
            sm->vote[x] = make_int2(0, 0);
            sm->vlast[x] = 0u;
            if (!fail && (unsigned)e.y == M) {
                const int q = (int)(ev >> 20);
                const LInst &I = sm->inst[use_good ? (int)sm->good[q] : q];
                const unsigned okey = ((I.flags & kPos) ? 0x80000000u : 0u) | (unsigned)(forward ? I.bg : I.fg);
                if (okey < my_okey || (okey == my_okey && ev < my_ev)) my_okey = okey, my_ev = ev, my_vid = e.x;
            }
        }
    }
    __syncwarp();
    if (fail) return kBail;
    if (M == 0) return kOk;
    const unsigned kmin = __reduce_min_sync(kFull, my_okey);
    const unsigned emin = __reduce_min_sync(kFull, my_okey == kmin ? my_ev : 0xFFFFFFFFu);
    // (okey, event) names one (walk, depth) item, hence one vertex: at most one lane matches
    const unsigned win = __ballot_sync(kFull, my_okey == kmin && my_ev == emin);
    const int wl = ffs_lane(win);
    best.vid = __shfl_sync(kFull, my_vid, wl);
    best.og = (int)(kmin & 0x7FFFFFFFu);
    best.opos = (kmin >> 31) != 0;
    best.d = (int)(emin & 0xFFFFFu);
    return kOk;
}

// The vote: the single-pass body first (every walk ends within 24 junctions, a table scan resolves it); the rare vote that
// does not fit is done again by the multi-pass body (deeper walks, bigger table).  Two bodies instead of one general one
// because the general one costs the common case 16 % (profiles/ab_tma_graph_r2.md).
// The multi-pass body as a real function (a cold one): it gets the few fields of the context it reads by value and
// hands its result back through memory, so that neither its instructions nor its registers sit in the hot path.
struct DeepOut {
    Next best;
    int status, why;
};
__device__ __noinline__ void most_popular_vertex_cold(LeanSmem *sm, const int4 *rec, const int2 *occ, const uint32_t *E, const int2 *hash2,
                                                      int hcount, uint32_t thresh, int b, int depth, int lane, int right_vertex,
                                                      int left_vertex, int ngood, int ninst, bool forward, bool try_used, DeepOut *out)
{
    LCtx c;
    c.sm = sm, c.rec = rec, c.occ = occ, c.E = E, c.hash2 = const_cast<int2 *>(hash2), c.hcount = hcount, c.thresh = thresh;
    c.b = b, c.depth = depth, c.lane = lane, c.right_vertex = right_vertex, c.left_vertex = left_vertex;
    c.ngood = ngood, c.ninst = ninst, c.why = 0;
#if defined(LCB_TMA_WINDOWS) && defined(__CUDACC__)
    c.tma_phase = 0; // (the bulk-copy variant is only measured on inputs that never come here)
#endif
    Next best;
    const int status = most_popular_vertex_deep(c, forward, try_used, best);
    if (lane == 0) out->best = best, out->status = status, out->why = c.why;
}

__device__ __forceinline__ int most_popular_vertex(LCtx &c, bool forward, bool try_used, Next &best)
{
    // On inputs whose walks are usually deeper than the single-pass body covers (many genomes, dense junctions) trying it
    // first only doubles the work: the warp remembers how its last votes went and goes straight to the multi-pass body
    // while most of them needed it.
    if (c.deep_bias < 4) {
        const int r = most_popular_vertex_fast(c, forward, try_used, best);
        if (r != kRetry) {
            c.deep_bias = max(c.deep_bias - 1, 0);
            return r;
        }
        c.deep_bias = min(c.deep_bias + 2, 8);
    } else {
        c.deep_bias--; // (an attempt with the fast body every now and then)
    }
    DeepOut *out = (DeepOut *)c.sm->best_scratch; // (shared memory: visible to the whole warp after the call)
    most_popular_vertex_cold(c.sm, c.rec, c.occ, c.E, c.hash2, c.hcount, c.thresh, c.b, c.depth, c.lane, c.right_vertex, c.left_vertex,
                             c.ngood, c.ninst, forward, try_used, out);
    __syncwarp();
    best = out->best;
    c.why = out->why;
    const int status = out->status;
    __syncwarp();
    return status;
}

// ExtendPathForward / ExtendPathBackward (blocksfinder.h:770-895).  Returns 0 failed, 1 success, 2 bail.
__device__ __forceinline__ int extend_path(LCtx &c, const bool forward, int &best_size, long long &best_score, long long &now_score)
{
    Next nx;
    nx.vid = 0;
#pragma unroll 1
    for (int attempt = 0; attempt < (forward ? 2 : 1) && nx.vid == 0; attempt++) // forward retries with tryUsed (:782-785)
        if (most_popular_vertex(c, forward, attempt == 1, nx) != kOk) return 2;
    if (nx.vid == 0) return 0;
    bool success = false;
    const int step = (forward == nx.opos) ? 1 : -1;
    // `for (it = origin; it.GetVertexId() != next; ++it) push(it.Outgoing/IngoingEdge())`: stops at the FIRST junction of
    // the walk that carries the chosen vertex (blocksfinder.h:789, :852); junctions og .. og + step * d, a warp's worth at a time
    int4 rc = make_int4(0, 0, 0, 0);
    if (c.lane <= nx.d) rc = __ldg(c.rec + (nx.og + step * c.lane));
    int prev_v = forward ? c.right_vertex : c.left_vertex; // vertex of the origin junction
    unsigned prev_bp = (unsigned)__shfl_sync(kFull, rc.y, 0);
    for (int j = 1; j <= nx.d; j++) {
        if ((j & 31) == 0) { // next 32 junctions of a deep walk
            rc = make_int4(0, 0, 0, 0);
            if (j + c.lane <= nx.d) rc = __ldg(c.rec + (nx.og + step * (j + c.lane)));
        }
        const int idv = __shfl_sync(kFull, rc.x, j & 31);
        const unsigned bp = (unsigned)__shfl_sync(kFull, rc.y, j & 31);
        const unsigned o_first = (unsigned)__shfl_sync(kFull, rc.z, j & 31);
        const unsigned o_count = (unsigned)__shfl_sync(kFull, rc.w, j & 31) >> 16;
        const int v = nx.opos ? idv : -idv;
        const int len = (int)(bp > prev_bp ? bp - prev_bp : prev_bp - bp);
        const int g_prev = nx.og + step * (j - 1), g_now = nx.og + step * j;
        const int r = path_push(c, forward, v, len, forward ? g_prev : g_now, nx.opos, prev_v, o_first, o_count);
        if (r == 2) return 2;
        success = r == 0;
        if (success) {
            now_score = path_score(c);
            if (now_score > best_score) {
                best_score = now_score;
                best_size = (forward ? c.nright : c.nleft) + 1;
                if (now_score > 0) snapshot_best(c);
                if (forward) snapshot_state(c);
            }
        }
        if (v == nx.vid) break;
        prev_v = v;
        prev_bp = bp;
    }
    return success ? 1 : 0;
}

// ProcessVertex::Process (blocksfinder.h:228-310).  On kOk: c.sm->best[0..nbest) is bestInstance, c.rs[0..nrs) the
// read-set.  On kBail nothing is published and the shared state is clean.
__device__ __forceinline__ int process_seed(LCtx &c, int vid, unsigned char ch)
{
    c.ninst = c.ngood = c.nbest = c.hcount = c.nrs = 0;
    c.nright = c.nleft = 0;
    long long best_score = 0, score = 0;
    int best_size[2] = {1, 1}; // bestLeftSize, bestRightSize
    const int min_run = c.b * 2;
    for (int phase = 1; phase >= 0; phase--) { // 1: forward, 0: backward
        const bool forward = phase == 1;
        if (forward) {
            if (path_init(c, vid, ch) != kOk) {
                abandon(c);
                return kBail;
            }
            if (c.ninst == 0) break; // a seed without live instances cannot move (MostPopularVertex finds nothing)
            snapshot_state(c);       // bestRightSize == 1: the state right after Init
        } else {
            // snapshots are taken at every improvement of the forward score, so the snapshot IS the state after
            // bestRightSize - 1 right pushes
            if (!restore_state(c)) {
                abandon(c);
                c.why = kWhyReadSet;
                return kBail;
            }
        }
        while (true) {
            int ret = 1;
            bool positive = false;
            const int prev_len = c.right_flank - c.left_flank;
            while ((ret = extend_path(c, forward, best_size[phase], best_score, score)) == 1 &&
                   (c.right_flank - c.left_flank) - prev_len <= min_run) {
                if (forward) positive = positive || score > 0; // backward: empty body, stray ';' at blocksfinder.h:297
            }
            if (ret == 2) {
                abandon(c);
                return kBail;
            }
            if (!forward) positive = positive || score > 0;
            if (ret == 0 || !positive) break;
        }
    }
    const int nb = c.nbest;
    if (!path_clear(c)) {
        abandon(c);
        c.why = kWhyReadSet;
        return kBail;
    }
    c.nbest = nb;
    return kOk;
}

} // namespace lean
} // namespace lcb
