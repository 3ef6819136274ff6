// lcb_device.cu -- device side of the B200 sibeliaz-lcb path and the C ABI around it (include/sibeliaz_lcb.h).
//
// What runs here replaces BlocksFinder::FindBlocks (SibeliaZ-LCB/blocksfinder.h:453-530):
//   * seed (bundle) enumeration + sort                         blocksfinder.h:461-503, :517
//   * the carving-path traversal per seed                      blocksfinder.h:228-310, path.h
//       k_traverse_lean (lcb_lean.cuh): the common case in a small body; k_traverse (lcb_traverse.cuh): everything
//   * the 256-seed phase / ordered-commit protocol             blocksfinder.h:334-433
//       rounds driven from the device (k_round_begin / k_round_end, tail of a round as a CUDA graph), change map +
//       lane-per-seed validation, peer-mailbox exchange between GPUs
//
// The commit protocol is sequential in the reference.  Here it is evaluated as a FIXPOINT over a window
// of W seeds (W a multiple of the phase size; all earlier seeds are final):
//
//   r0_i = Process(i, edge used  <=>  epoch < phase_start(i))                 "speculative" result
//   conf_i = |r0_i| > 1 and some edge of r0_i has epoch < i                   blocksfinder.h:375-398
//   r1_i = Process(i, edge used  <=>  epoch < i)          (only if conf_i)    blocksfinder.h:404-412
//   final_i = conf_i ? (|r1_i| > 1 ? r1_i : {}) : (|r0_i| > 1 ? r0_i : {})
//   epoch'[e] = min { i : e in edges(final_i) }  over the window, on top of the committed epochs
//
// final_i depends only on epochs < i, so the system has exactly one solution -- the reference's result --
// and Jacobi iteration reaches it: every round re-evaluates (all in parallel) only the seeds whose recorded
// read-set saw a changed epoch.  Block ids and the blocksInstance_ order are then a prefix sum in seed order.
//
// The window ROLLS: seeds [c0, c1) are active; every round admits new seeds at c1 (evaluated against the current
// epoch estimate, next to the re-evaluations of the unconverged ones, so the latency-bound tail rounds of earlier
// seeds are filled with fresh work) and commits the clean prefix [c0, first dirty seed): all seeds before the first
// dirty one are consistent with epochs that only they determine, hence final (same induction as above).
#include "sibeliaz_lcb.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "lcb_traverse.cuh"
#include "lcb_lean.cuh"
#include "graph_internal.h"
#include "lcb_internal.h"

#ifdef LCB_WITH_NCCL
#include <dlfcn.h>
#include <nccl.h> // types only: the library itself is dlopen'ed on first multi-GPU use (single-GPU hosts never load it)
#endif

using namespace lcb;

extern std::atomic<unsigned> lcb_host_thread_cap; // lcb_host.cpp

namespace {

#ifndef LCB_TRAVERSE_CTAS_PER_SM
#define LCB_TRAVERSE_CTAS_PER_SM 4 // 4 warps each: 16 resident warps per SM at 128 registers per thread (5 % faster than 3 x 168 on the throughput-bound 4x100 Mbp input)
#endif
constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;

struct Control { // device-resident round state, mirrored to pinned host memory once per round
    unsigned head;      // work-queue cursor of the running traversal launch (common-case kernel, or the only kernel)
    unsigned max_ns;    // longest single evaluation since the last reset (globaltimer ns, saturating)
    unsigned head_heavy; // work-queue cursor of the general kernel over win.list_heavy (beside the common-case kernel)
    unsigned head_drain; // ... and of the general kernel behind the common-case kernel (takes what is still in the list)
    unsigned n_heavy;   // entries of win.list_heavy: seeds known to need the general kernel (queued by the validation) + what
                        // the common-case kernel hands over while it runs
    unsigned heavy_done;  // entries of the list the general kernel processed in the last round
    unsigned lean_started; // the common-case kernel of this round is running (set by its first thread)
    unsigned lean_exited; // CTAs of the common-case kernel that have finished this round (the general kernel beside it stops
                          // waiting for hand-overs when all have)
    unsigned n0, n1;    // length of the work list (bit 31 of an item: commit-time re-run only); n1 unused
    unsigned dirty;     // seeds whose dependencies changed in the last validation
    unsigned first_dirty; // smallest such seed index (0xFFFFFFFF: none): everything before it is final
    unsigned err;       // first LCB_ERR_* raised by a kernel
    unsigned pool_overflow;
    unsigned long long inst_used, rs_used; // bump allocators
    unsigned long long ct_walk, ct_occ, ct_scan, ct_score;
    unsigned long long runs0, runs1;
    unsigned blocks_done, out_done; // emit: blocks / instances committed so far
    unsigned long long dbg[6];
    unsigned long long lean_runs, lean_bails; // evaluations finished by the common-case kernel / handed to the general one
    unsigned lean_why[lean::kWhyCount];       // ... by reason
    // ---- device-driven schedule (single GPU): the round loop's decisions are taken by k_round_begin / k_round_end, the host
    // only keeps launches queued and watches `Mirror` in pinned memory
    unsigned c0, c1;          // rolling active set [c0, c1): commit frontier, admission frontier
    unsigned admit_lo, admit, admit_n0; // seeds admitted at the beginning of this round; length of the work list before them
    unsigned delta, cap;      // admission rate (adapted every round), bound on the active set
    unsigned drain, hold;     // result pools half full / heavy re-evaluations occupy the warps: no admission
    unsigned n_seeds;
    unsigned parity;          // which epoch buffer is current
    unsigned done, halt;      // all seeds committed / the host must look (error, pool overflow)
    unsigned rounds, windows;
    unsigned commit_first, commit_n; // prefix committed by this round's emit kernels
    unsigned long long t_first, t_last; // %globaltimer bracket of this round's traversal kernels
    // where the rest of a round goes (diagnostics, lcb_stats.ms_tail): %globaltimer at the start of k_rebase, k_claim, k_diff,
    // k_validate, k_round_end, and sums of the differences (gap after the traversal, rebase, claim, diff, validation, commit)
    unsigned long long t_mark[5], tail_ns[6];
    unsigned x_entries, x_inst;         // results / instances this rank has stored into its peers' mailboxes this round
    unsigned x_tag_base;                // run number << 22: tags of different runs never compare equal
    unsigned big_runs;            // evaluations that outgrew the per-warp arena and ran in a big slot
    unsigned big_lock[kBigSlots]; // 1 = big arena slot taken (always released by its holder)
};

struct Window { // per-seed arrays of the active seeds, ring-indexed by j = seed & mask
    unsigned *res_off[2], *res_cnt[2]; // best instances of slot s in inst_pool
    unsigned *rs_off[2], *rs_cnt[2];   // read-set of slot s in rs_pool
    unsigned char *conf, *has1;
    unsigned char *heavy; // the seed needs the general kernel (it made the common-case kernel bail once)
    unsigned *blk, *out_off;
    unsigned *list0;
    unsigned *list_heavy;
    int4 *inst_pool;
    int2 *rs_pool;
    unsigned long long inst_cap, rs_cap;
    unsigned mask; // ring size - 1
};

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

#define CUDA_TRY(x)                                                                                       \
    do {                                                                                                  \
        cudaError_t e_ = (x);                                                                             \
        if (e_ != cudaSuccess) {                                                                          \
            ctx->error = std::string(#x) + ": " + cudaGetErrorString(e_);                                 \
            return LCB_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

#ifdef LCB_WITH_NCCL
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi *nccl_api()
{
    static NcclApi api = []() {
        NcclApi a;
        const char *cands[] = {getenv("LCB_NCCL_LIB"), "libnccl.so.2",
#ifdef LCB_NCCL_PATH
                               LCB_NCCL_PATH,
#endif
                               "libnccl.so"};
        void *h = nullptr;
        for (const char *c : cands)
            if (c && (h = dlopen(c, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!h) return a;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
        a.Broadcast = (decltype(a.Broadcast))dlsym(h, "ncclBroadcast");
        a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Broadcast && a.AllGather && a.GetErrorString;
        return a;
    }();
    return &api;
}
#define ncclGetUniqueId nccl_api()->GetUniqueId
#define ncclCommInitRank nccl_api()->CommInitRank
#define ncclCommDestroy nccl_api()->CommDestroy
#define ncclBroadcast nccl_api()->Broadcast
#define ncclAllGather nccl_api()->AllGather
#define ncclGetErrorString nccl_api()->GetErrorString
#define NCCL_TRY(x)                                                                                       \
    do {                                                                                                  \
        ncclResult_t r_ = (x);                                                                            \
        if (r_ != ncclSuccess) {                                                                          \
            ctx->error = std::string(#x) + ": " + ncclGetErrorString(r_);                                 \
            return LCB_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)
#endif

// edges of one instance as an epoch index range: Finalize's MarkUsed loop (blocksfinder.h:327-330)
__device__ __forceinline__ void inst_edges(const int4 &b, int &lo, int &hi)
{
    int fg = b.x & 0x7FFFFFFF, bg = b.y;
    lo = min(fg, bg);
    hi = max(fg, bg) - 1;
}


// ------------------------------------------------------------------------------------------------
// Multi-GPU: peer exchange over NVLink (one process per GPU, CUDA IPC).  The index and the epochs are replicated, the
// seeds of the active set are dealt round-robin (i % R); every rank needs every seed's RESULT (for the claims and the
// ordered commit) but only the owner needs its read-set.  So a traversal warp that publishes a result also stores it into
// every peer's MAILBOX (plain stores to peer memory, overlapped with the traversal of the other seeds); after the traversal
// a one-warp kernel stores the counts and a round tag, the peers wait for the tags and file the results into their own
// pools.  A second, 16-word exchange per round carries the owners' validation outcome (first dirty seed, overflow, timing)
// so that all ranks take the same decisions.  No collective library call and no host in the loop.
// Mailbox of a rank: R x 2 regions (source rank, round parity), each {XHdr, entries[ent_cap], instances[inst_cap]}.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 8;
struct XHdr {
    unsigned tag, n_entries, n_inst, ctl_tag; // tags: (run << 22) | round, written last
    unsigned ctl[12];                         // first_dirty, err, overflow, drain, ms bits, longest ns, n0, n_heavy
};
static_assert(sizeof(XHdr) == 64, "mailbox header");
struct XchDev { // by value to the kernels; R <= 1: no exchange
    int R, me;
    unsigned ent_cap, inst_cap;
    unsigned char *box[kMaxRanks]; // box[r]: rank r's mailbox in this process' address space (box[me]: the local one)
};
__host__ __device__ __forceinline__ size_t xch_region_bytes(unsigned ent_cap, unsigned inst_cap)
{
    return sizeof(XHdr) + (size_t)ent_cap * 16 + (size_t)inst_cap * 16;
}
__device__ __forceinline__ unsigned char *xch_region(const XchDev &x, int owner, int src, unsigned parity)
{
    return x.box[owner] + ((size_t)src * 2 + parity) * xch_region_bytes(x.ent_cap, x.inst_cap);
}

// a freshly published result goes to every peer: entry {item, conflict flag, instances, offset} + the instances
__device__ __noinline__ void x_send_peers(const XchDev &x, Control *ctl, unsigned i, int slot, bool conf, int nbest, const int4 *best, int lane)
{
    unsigned e = 0, io = 0;
    if (lane == 0) {
        e = atomicAdd(&ctl->x_entries, 1u);
        io = atomicAdd(&ctl->x_inst, (unsigned)nbest);
    }
    e = __shfl_sync(kFull, e, 0);
    io = __shfl_sync(kFull, io, 0);
    if (e >= x.ent_cap || io + (unsigned)nbest > x.inst_cap) { // handled like a full result pool: all ranks restart the active set
        if (lane == 0) atomicExch(&ctl->pool_overflow, 1u);
        return;
    }
    const unsigned parity = ctl->rounds & 1u;
    for (int r = 0; r < x.R; r++) {
        if (r == x.me) continue;
        unsigned char *reg = xch_region(x, r, x.me, parity);
        if (lane == 0) ((uint4 *)(reg + sizeof(XHdr)))[e] = make_uint4(i | (slot ? 0x80000000u : 0u), conf ? 1u : 0u, (unsigned)nbest, io);
        int4 *dst = (int4 *)(reg + sizeof(XHdr) + (size_t)x.ent_cap * 16) + io;
        for (int t = lane; t < nbest; t += 32) dst[t] = best[t];
    }
}

__device__ __forceinline__ void x_send(const XchDev &x, Control *ctl, unsigned i, int slot, bool conf, int nbest, const int4 *best, int lane)
{
    if (x.R > 1) x_send_peers(x, ctl, i, slot, conf, nbest, best, lane);
}

// after the traversal kernels of a round: counts, then the tag, into every peer's mailbox
__global__ void k_xflag(XchDev x, Control *ctl)
{
    if (ctl->done | ctl->halt) return;
    const int r = threadIdx.x;
    if (r >= x.R || r == x.me) return;
    XHdr *h = (XHdr *)xch_region(x, r, x.me, ctl->rounds & 1u);
    h->n_entries = min(ctl->x_entries, x.ent_cap);
    h->n_inst = ctl->x_inst;
    __threadfence_system();
    *(volatile unsigned *)&h->tag = ctl->x_tag_base | ctl->rounds;
    __threadfence_system();
}

// wait until every peer's results of this round are in the local mailbox (lane r watches peer r)
__global__ void k_xwait(XchDev x, Control *ctl)
{
    if (ctl->done | ctl->halt) return;
    const int r = threadIdx.x;
    if (r >= x.R || r == x.me) return;
    const volatile XHdr *h = (const volatile XHdr *)xch_region(x, x.me, r, ctl->rounds & 1u);
    const unsigned want = ctl->x_tag_base | ctl->rounds;
    const unsigned long long t0 = global_ns();
    while (h->tag != want) {
        __nanosleep(200);
        if (global_ns() - t0 > 60000000000ull) { // a peer that has not shown up for a minute is gone
            atomicCAS(&ctl->err, 0u, (unsigned)LCB_ERR_CUDA);
            break;
        }
    }
}

// file the peers' results of this round into the local pools and per-seed state (one warp per entry)
__global__ void k_xapply(XchDev x, Window win, Control *ctl)
{
    if (ctl->done | ctl->halt) return;
    const int lane = threadIdx.x & 31;
    const unsigned warps = gridDim.x * (blockDim.x >> 5), warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned parity = ctl->rounds & 1u;
    for (int src = 0; src < x.R; src++) {
        if (src == x.me) continue;
        const unsigned char *reg = xch_region(x, x.me, src, parity);
        const XHdr *h = (const XHdr *)reg;
        const uint4 *ent = (const uint4 *)(reg + sizeof(XHdr));
        const int4 *inst = (const int4 *)(reg + sizeof(XHdr) + (size_t)x.ent_cap * 16);
        const unsigned n = h->n_entries;
        for (unsigned e = warp; e < n; e += warps) {
            const uint4 en = ent[e];
            const unsigned i = en.x & 0x7FFFFFFFu, j = i & win.mask, cnt = en.z;
            const int slot = (int)(en.x >> 31);
            unsigned long long io = 0;
            if (lane == 0) io = atomicAdd(&ctl->inst_used, (unsigned long long)cnt);
            io = __shfl_sync(kFull, io, 0);
            if (io + cnt > win.inst_cap) {
                if (lane == 0) {
                    atomicExch(&ctl->pool_overflow, 1u);
                    win.res_cnt[slot][j] = 0;
                }
                continue;
            }
            for (unsigned t = lane; t < cnt; t += 32) win.inst_pool[io + t] = inst[en.w + t];
            if (lane == 0) {
                win.res_off[slot][j] = (unsigned)io;
                win.res_cnt[slot][j] = cnt;
                if (slot == 0) win.conf[j] = (unsigned char)en.y;
            }
        }
    }
}

// after the validation: this rank's outcome of the round to every peer
__global__ void k_csend(XchDev x, Control *ctl, unsigned long long inst_cap, unsigned long long rs_cap)
{
    if (ctl->done | ctl->halt) return;
    const int r = threadIdx.x;
    if (r >= x.R || r == x.me) return;
    XHdr *h = (XHdr *)xch_region(x, r, x.me, ctl->rounds & 1u);
    const float ms = ctl->t_last > ctl->t_first ? (float)(ctl->t_last - ctl->t_first) * 1e-6f : 0.f;
    h->ctl[0] = ctl->first_dirty;
    h->ctl[1] = ctl->err;
    h->ctl[2] = ctl->pool_overflow;
    h->ctl[3] = (ctl->inst_used * 2 > inst_cap || ctl->rs_used * 2 > rs_cap) ? 1u : 0u;
    h->ctl[4] = __float_as_uint(ms);
    h->ctl[5] = ctl->max_ns;
    h->ctl[6] = ctl->n0;
    h->ctl[7] = ctl->n_heavy;
    __threadfence_system();
    *(volatile unsigned *)&h->ctl_tag = ctl->x_tag_base | ctl->rounds;
    __threadfence_system();
}

// ------------------------------------------------------------------------------------------------
// traversal kernel: persistent warps pull (seed, slot) items from a list
// ------------------------------------------------------------------------------------------------
// `heavy_mode` 0: `list` holds plain work items.  1: the work is win.list_heavy (entries are item + 1, 0 = not written yet;
// consumed entries are zeroed), complete when the kernel starts.  2: the same list while the common-case kernel is still
// appending to it from the other stream: a warp that finds the list exhausted waits until all `lean_total` CTAs of that
// kernel have exited (MINB = 5: 96 registers, so that a CTA of this kernel fits beside five of the other).
template <bool COLLECT, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_traverse(Index ix, Params pr, const uint32_t *__restrict__ E,
                                                        const int *__restrict__ seed_vid,
                                                        const unsigned char *__restrict__ seed_ch,
                                                        unsigned phase, int force_slot, const unsigned *__restrict__ list,
                                                        const unsigned *__restrict__ n_ptr, unsigned *__restrict__ cursor, Window win,
                                                        Control *ctl, unsigned char *arena_base, size_t arena_stride,
                                                        unsigned char *big_base, int collect, int heavy_mode, unsigned lean_total, XchDev xch)
{
    __shared__ WarpSmem smem[kWarpsPerBlock];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t warp_global = (size_t)blockIdx.x * kWarpsPerBlock + wib;
    Ctx c;
    c.ix = ix;
    c.pr = pr;
    c.E = E;
    c.lane = lane;
    c.sm = &smem[wib];
    c.err = 0;
    c.collect = COLLECT;
    c.ct.walk = c.ct.occ = c.ct.scan = c.ct.score = 0;
    c.ct.pushes = c.ct.mpv_fast = c.ct.mpv_mid = c.ct.mpv_slow = c.ct.push_par = c.ct.push_ser = 0;
    c.vote_clean = false;
    arena_bind(c, arena_base + warp_global * arena_stride, false);
    int big_slot = -1; // >= 0: this warp holds that big arena slot
    const bool off = (ctl->halt | ctl->done) != 0; // (a round queued behind the one that ended the run, or stopped it)
    const unsigned n = off ? 0u : *n_ptr;
    const unsigned long long t_enter = global_ns();
    if (n && blockIdx.x == 0 && threadIdx.x == 0) atomicMin(&ctl->t_first, global_ns());
    unsigned done = 0, done1 = 0;
    unsigned long long longest = 0;
    while (!off) {
        unsigned idx = 0;
        if (lane == 0) idx = atomicAdd(cursor, 1u);
        idx = __shfl_sync(kFull, idx, 0);
        unsigned item = 0;
        if (heavy_mode == 0) {
            if (idx >= n) break;
            item = list[idx];
        } else if (heavy_mode == 1) { // behind the common-case kernel: the list is complete; consumed entries are zero
            if (idx >= n) break;
            if (lane == 0) item = atomicExch(&win.list_heavy[idx], 0u);
            item = __shfl_sync(kFull, item, 0);
            if (item == 0u) continue;
            item -= 1u;
        } else {
            if (lane == 0) {
                volatile unsigned *vl = win.list_heavy;
                const volatile unsigned *vn = &ctl->n_heavy, *vx = &ctl->lean_exited;
                while (true) {
                    if (idx < *vn) { // the entry exists; its writer may still be between the counter and the store
                        unsigned v;
                        while ((v = vl[idx]) == 0u) __nanosleep(100);
                        vl[idx] = 0u;
                        item = v;
                        break;
                    }
                    if (*vx >= lean_total) {
                        __threadfence();
                        if (idx < *vn) continue; // appended just before the last producer left
                        break;
                    }
                    // Under a tool that serialises kernels (a profiler) the common-case kernel cannot start while this one
                    // runs: give up after a millisecond of nothing; the launch behind the join drains what is handed over later.
                    if (!*(const volatile unsigned *)&ctl->lean_started && global_ns() - t_enter > 1000000ull) break;
                    __nanosleep(1000);
                }
            }
            item = __shfl_sync(kFull, item, 0);
            if (item == 0u) break;
            item -= 1u;
        }
        const unsigned i = item & 0x7FFFFFFFu;
        const unsigned j = i & win.mask;
        int slot = force_slot >= 0 ? force_slot : (int)(item >> 31); // bit 31: commit-time re-run queued by the validation
        // at most two evaluations: the speculative one and, when its result runs into edges that earlier seeds of its
        // phase claimed, the commit-time re-run right behind it (blocksfinder.h:375-412) -- one call site, same warp
        while (true) {
            c.thresh = slot == 0 ? (i / phase) * phase : i;
            const unsigned long long t_begin = global_ns();
            process_seed(c, seed_vid[i], seed_ch[i]);
            if ((c.err & 0xFF) == LCB_ERR_CAPACITY && big_slot < 0) {
                // the seed outgrew the per-warp arena: leave that arena clean (spill hash all-empty), take one of the few big
                // slots (their holders wait for nothing, so waiting here cannot deadlock) and evaluate the seed again
                hash_clear(c);
                if (lane == 0) {
                    unsigned s = (unsigned)warp_global % kBigSlots;
                    while (atomicCAS(&ctl->big_lock[s], 0u, 1u) != 0u) {
                        s = (s + 1) % kBigSlots;
                        __nanosleep(200);
                    }
                    __threadfence();
                    big_slot = (int)s;
                    atomicAdd(&ctl->big_runs, 1u);
                }
                big_slot = __shfl_sync(kFull, big_slot, 0);
                arena_bind(c, big_base + (size_t)big_slot * arena_stride_of(true), true);
                c.err = 0;
                continue; // same call site, big arena
            }
            longest = max(longest, global_ns() - t_begin);
            if (c.err) break;
            // publish bestInstance and the read-set
            unsigned long long io = 0, ro = 0;
            if (lane == 0) {
                io = atomicAdd(&ctl->inst_used, (unsigned long long)c.nbest);
                ro = atomicAdd(&ctl->rs_used, (unsigned long long)c.nrs);
            }
            io = __shfl_sync(kFull, io, 0);
            ro = __shfl_sync(kFull, ro, 0);
            if (io + c.nbest > win.inst_cap || ro + c.nrs > win.rs_cap) {
                // the host abandons the active set and starts it again; until it notices, the kernels queued behind this
                // one must see a harmless (empty) entry for this seed
                if (lane == 0) {
                    atomicExch(&ctl->pool_overflow, 1u);
                    win.res_cnt[slot][j] = 0;
                    win.rs_cnt[slot][j] = 0;
                }
                break;
            }
            for (int t = lane; t < c.nbest; t += 32) win.inst_pool[io + t] = c.best[t];
            for (int t = lane; t < c.nrs; t += 32) win.rs_pool[ro + t] = c.ar.rs()[t];
            if (lane == 0) {
                win.res_off[slot][j] = (unsigned)io;
                win.res_cnt[slot][j] = (unsigned)c.nbest;
                win.rs_off[slot][j] = (unsigned)ro;
                win.rs_cnt[slot][j] = (unsigned)c.nrs;
            }
            if (slot != 0) {
                x_send(xch, ctl, i, 1, false, c.nbest, c.best, lane);
                done1++;
                break;
            }
            done++;
            bool conf = false; // commit-time conflict test of the fresh result: any of its edges claimed by a seed < i?
            if (c.nbest > 1)
                for (int t = 0; t < c.nbest && !conf; t++) {
                    int lo, hi;
                    inst_edges(c.best[t], lo, hi);
                    for (int base = lo; base <= hi && !conf; base += 32) {
                        const int f = base + lane;
                        conf = __any_sync(kFull, f <= hi && __ldg(E + f) < i);
                    }
                }
            if (lane == 0) {
                win.conf[j] = conf;
                win.has1[j] = conf;
            }
            x_send(xch, ctl, i, 0, conf, c.nbest, c.best, lane);
            if (!conf) break;
            slot = 1;
        }
        if (big_slot >= 0) { // results are published (or the run failed): hand the big slot back
            if (c.err) hash_clear(c);
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                atomicExch(&ctl->big_lock[big_slot], 0u);
            }
            big_slot = -1;
            arena_bind(c, arena_base + warp_global * arena_stride, false);
        }
        if (c.err) {
            if (lane == 0) atomicCAS(&ctl->err, 0u, (unsigned)c.err);
            break;
        }
    }
    if (xch.R > 1) __threadfence_system(); // this warp's stores to the peers' mailboxes
    if (lane == 0) {
        if (longest) atomicMax(&ctl->max_ns, (unsigned)min(longest, 0xFFFFFFFFull));
        if (done) atomicAdd(&ctl->runs0, (unsigned long long)done);
        if (done1) atomicAdd(&ctl->runs1, (unsigned long long)done1);
        if (done + done1) atomicMax(&ctl->t_last, global_ns());
        if (collect) {
            atomicAdd(&ctl->ct_walk, c.ct.walk);
            atomicAdd(&ctl->ct_occ, c.ct.occ);
            atomicAdd(&ctl->ct_scan, c.ct.scan);
            atomicAdd(&ctl->ct_score, c.ct.score);
            atomicAdd(&ctl->dbg[0], (unsigned long long)c.ct.pushes);
            atomicAdd(&ctl->dbg[1], (unsigned long long)c.ct.mpv_fast);
            atomicAdd(&ctl->dbg[2], (unsigned long long)c.ct.mpv_slow);
            atomicAdd(&ctl->dbg[3], (unsigned long long)c.ct.push_par);
            atomicAdd(&ctl->dbg[4], (unsigned long long)c.ct.push_ser);
            atomicAdd(&ctl->dbg[5], (unsigned long long)c.ct.mpv_mid);
        }
    }
}


// ------------------------------------------------------------------------------------------------
// common-case traversal kernel (lcb_lean.cuh): same work list, same publication; what it cannot evaluate goes to
// `win.list_heavy` for the general kernel, which runs right behind it in the same round
// ------------------------------------------------------------------------------------------------
#ifndef LCB_LEAN_CTAS_PER_SM
#define LCB_LEAN_CTAS_PER_SM 6
#endif
constexpr int kLeanRs = 4096; // read-set intervals per evaluation in the common-case kernel's log (more: general kernel)

__global__ void __launch_bounds__(kThreads, LCB_LEAN_CTAS_PER_SM) k_traverse_lean(Index ix, Params pr, const uint32_t *__restrict__ E,
                                                        const int *__restrict__ seed_vid,
                                                        const unsigned char *__restrict__ seed_ch, unsigned phase,
                                                        const unsigned *__restrict__ list, const unsigned *__restrict__ n_ptr,
                                                        Window win, Control *ctl, int2 *__restrict__ rs_base, lean::LInst *__restrict__ shadow_base,
                                                        int2 *__restrict__ hash2_base, unsigned short *__restrict__ hslot2_base, XchDev xch)
{
    __shared__ lean::LeanSmem smem[kWarpsPerBlock];
    __shared__ uint32_t chr_off_s[lean::kLChr];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int i = threadIdx.x; i <= ix.C; i += blockDim.x) chr_off_s[i] = ix.chr_off[i];
    { // the hash and the vote table start empty and every evaluation leaves them so
        lean::LeanSmem *sm = &smem[wib];
        for (int i = lane; i < lean::kLHash; i += 32) sm->hash[i] = make_int2(0, 0);
        for (int i = lane; i < lean::kLVote; i += 32) sm->vote[i] = make_int2(0, 0), sm->vlast[i] = 0u;
#ifdef LCB_TMA_WINDOWS
        if (lane == 0) lean::mbar_init(&sm->mbar);
#endif
    }
    __syncthreads();
    const size_t warp_global = (size_t)blockIdx.x * kWarpsPerBlock + wib;
    lean::LCtx c;
    c.rec = ix.rec, c.occ = ix.occ, c.vtx_off = ix.vtx_off, c.E = E;
    c.chr_off_s = chr_off_s, c.C = ix.C;
    c.b = pr.b, c.m = pr.m, c.flank = pr.flank, c.depth = pr.depth;
    c.lane = lane;
    c.sm = &smem[wib];
    c.rs = rs_base + warp_global * (size_t)kLeanRs;
    c.rs_cap = kLeanRs;
    c.shadow = shadow_base + warp_global * (size_t)lean::kLInst;
    c.hash2 = hash2_base + warp_global * (size_t)lean::kLHash2;
    c.hslot2 = hslot2_base + warp_global * (size_t)lean::kLPath2;
    c.last_clo = 0, c.last_chi = 0;
    c.why = 0;
    c.deep_bias = 0;
#ifdef LCB_TMA_WINDOWS
    c.tma_phase = 0;
#endif
    const unsigned n = (ctl->halt | ctl->done) ? 0u : *n_ptr; // (a round queued behind the one that ended the run, or stopped it)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *(volatile unsigned *)&ctl->lean_started = 1u;
        if (n) atomicMin(&ctl->t_first, global_ns());
    }
    unsigned done = 0, done1 = 0, bails = 0;
    unsigned long long longest = 0;
    while (true) {
        unsigned idx = 0;
        if (lane == 0) idx = atomicAdd(&ctl->head, 1u);
        idx = __shfl_sync(kFull, idx, 0);
        if (idx >= n) break;
        const unsigned item = list[idx];
        const unsigned i = item & 0x7FFFFFFFu;
        const unsigned j = i & win.mask;
        int slot = (int)(item >> 31); // bit 31: commit-time re-run queued by the validation
        if (win.heavy[j]) { // known to need the general kernel
            if (lane == 0) win.list_heavy[atomicAdd(&ctl->n_heavy, 1u)] = item + 1u;
            continue;
        }
        while (true) {
            c.thresh = slot == 0 ? (i / phase) * phase : i;
            const unsigned long long t_begin = global_ns();
            const int r = lean::process_seed(c, seed_vid[i], seed_ch[i]);
            if (r != lean::kOk) { // the rest of this item (the evaluation that failed and what follows it) is the general kernel's
                if (lane == 0) {
                    win.heavy[j] = 1;
                    win.list_heavy[atomicAdd(&ctl->n_heavy, 1u)] = (i | (slot ? 0x80000000u : 0u)) + 1u;
                    atomicAdd(&ctl->lean_why[c.why], 1u);
                }
                bails++;
                break;
            }
            longest = max(longest, global_ns() - t_begin);
            unsigned long long io = 0, ro = 0;
            if (lane == 0) {
                io = atomicAdd(&ctl->inst_used, (unsigned long long)c.nbest);
                ro = atomicAdd(&ctl->rs_used, (unsigned long long)c.nrs);
            }
            io = __shfl_sync(kFull, io, 0);
            ro = __shfl_sync(kFull, ro, 0);
            if (io + c.nbest > win.inst_cap || ro + c.nrs > win.rs_cap) {
                if (lane == 0) {
                    atomicExch(&ctl->pool_overflow, 1u);
                    win.res_cnt[slot][j] = 0;
                    win.rs_cnt[slot][j] = 0;
                }
                break;
            }
            if (lane < c.nbest) win.inst_pool[io + lane] = c.sm->best[lane];
            for (int t = lane; t < c.nrs; t += 32) win.rs_pool[ro + t] = c.rs[t];
            if (lane == 0) {
                win.res_off[slot][j] = (unsigned)io;
                win.res_cnt[slot][j] = (unsigned)c.nbest;
                win.rs_off[slot][j] = (unsigned)ro;
                win.rs_cnt[slot][j] = (unsigned)c.nrs;
            }
            if (slot != 0) {
                x_send(xch, ctl, i, 1, false, c.nbest, c.sm->best, lane);
                done1++;
                break;
            }
            done++;
            bool conf = false; // commit-time conflict test of the fresh result: any of its edges claimed by a seed < i?
            if (c.nbest > 1)
                for (int t = 0; t < c.nbest && !conf; t++) {
                    int lo, hi;
                    inst_edges(c.sm->best[t], lo, hi);
                    for (int base = lo; base <= hi && !conf; base += 32) {
                        const int f = base + lane;
                        conf = __any_sync(kFull, f <= hi && __ldg(E + f) < i);
                    }
                }
            if (lane == 0) {
                win.conf[j] = conf;
                win.has1[j] = conf;
            }
            x_send(xch, ctl, i, 0, conf, c.nbest, c.sm->best, lane);
            if (!conf) break;
            slot = 1;
        }
    }
    if (xch.R > 1) __threadfence_system(); // this warp's stores to the peers' mailboxes
    if (lane == 0) {
        if (longest) atomicMax(&ctl->max_ns, (unsigned)min(longest, 0xFFFFFFFFull));
        if (done) atomicAdd(&ctl->runs0, (unsigned long long)done);
        if (done1) atomicAdd(&ctl->runs1, (unsigned long long)done1);
        if (done + done1) atomicAdd(&ctl->lean_runs, (unsigned long long)(done + done1));
        if (done + done1 + bails) atomicMax(&ctl->t_last, global_ns());
        if (bails) atomicAdd(&ctl->lean_bails, (unsigned long long)bails);
    }
    __syncthreads(); // every hand-over of this CTA is in the list before the CTA counts as gone
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&ctl->lean_exited, 1u);
    }
}

// Change map of a round (k_diff): per 64 epoch entries, the smallest seed index that appears in an entry that differs between
// the current and the new epochs (min over the old and the new value; 0xFFFFFFFF: nothing changed).  A seed sees an edge as
// used iff its epoch is below the seed's limit, so a change matters to it only if one of the two values is below that limit:
// changes made by LATER seeds -- nearly all of a round's, the new claims come from the newly admitted seeds -- are invisible.
// The validation reads these words (under 1 MB in all, cache-resident) instead of both epoch arrays.
constexpr int kDiffShift = 6;
__device__ __forceinline__ bool range_dirty(const uint32_t *__restrict__ diff, int lo, int hi, uint32_t limit)
{
    if (!diff) return true;
    for (int b = lo >> kDiffShift; b <= (hi >> kDiffShift); b++)
        if (diff[b] < limit) return true;
    return false;
}

// does any edge of result `slot 0` of seed j carry an epoch < limit?   (warp-wide); `was`: the answer against the current epochs
__device__ __forceinline__ bool result_conflicts(const Window &win, unsigned j, const uint32_t *E, uint32_t limit,
                                                 int lane, const uint32_t *__restrict__ diff, bool was)
{
    const unsigned cnt = win.res_cnt[0][j], off = win.res_off[0][j];
    if (diff) { // unchanged epochs under every edge of the result: unchanged answer
        bool any = false;
        for (unsigned t = 0; t < cnt && !any; t++) {
            int lo, hi;
            inst_edges(win.inst_pool[off + t], lo, hi);
            for (int base = (lo >> kDiffShift) + lane; base <= (hi >> kDiffShift) && !any; base += 32) any = diff[base] < limit;
            any = __any_sync(kFull, any);
        }
        if (!any) return was;
    }
    bool hit = false;
    for (unsigned t = 0; t < cnt && !hit; t++) {
        int lo, hi;
        inst_edges(win.inst_pool[off + t], lo, hi);
        for (int base = lo; base <= hi && !hit; base += 32) {
            int f = base + lane;
            hit = __any_sync(kFull, f <= hi && E[f] < limit);
        }
    }
    return hit;
}

__device__ __forceinline__ void final_result(const Window &win, unsigned j, unsigned &off, unsigned &cnt)
{
    const int s = win.conf[j] ? 1 : 0;
    cnt = win.res_cnt[s][j];
    off = win.res_off[s][j];
    if (cnt <= 1) cnt = 0; // `if (instance.size() > 1)` blocksfinder.h:375,408
}

// start of a round's new epochs: the committed claims only (claims of seeds < c0 are final and minimal)
// `sched` (device-driven schedule): the frontier comes from the control block, and thread 0 resets the validation counters
// (this kernel runs between the traversal, which reads n0, and the validation, which rebuilds it)
__global__ void k_rebase(uint32_t *dst, const uint32_t *src, size_t n, uint32_t c0, Control *sched) // dst may alias src
{
    if (sched) {
        if (sched->done | sched->halt) return;
        c0 = sched->c0;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            sched->t_mark[0] = global_ns();
            sched->n0 = 0, sched->n1 = 0, sched->dirty = 0, sched->first_dirty = 0xFFFFFFFFu;
            sched->heavy_done = sched->n_heavy;        // what the general kernel had to do in this round
            sched->head_heavy = 0, sched->n_heavy = 0; // the validation queues the known-heavy seeds of the next round
        }
    }
    for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < n; f += (size_t)gridDim.x * blockDim.x) {
        const uint32_t e = src[f];
        dst[f] = e < c0 ? e : kFree;
    }
}

// epoch'[e] = min(seed index) over the final results of the active seeds [lo, hi)
__global__ void k_claim(uint32_t *__restrict__ Enew, unsigned lo_seed, unsigned hi_seed, Window win, const Control *sched)
{
    if (sched) {
        if (sched->done | sched->halt) return;
        lo_seed = sched->c0, hi_seed = sched->c1;
        if (blockIdx.x == 0 && threadIdx.x == 0) const_cast<Control *>(sched)->t_mark[1] = global_ns();
    }
    const int lane = threadIdx.x & 31;
    const unsigned warps = gridDim.x * (blockDim.x >> 5), warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // most seeds have no block to claim: one lane per seed finds the few that do, then the warp writes their edges
    for (unsigned base = lo_seed + warp * 32u; base < hi_seed; base += warps * 32u) {
        const unsigned i = base + (unsigned)lane;
        unsigned off = 0, cnt = 0;
        if (i < hi_seed) final_result(win, i & win.mask, off, cnt);
        unsigned m = __ballot_sync(kFull, cnt != 0);
        while (m) {
            const int src = ffs_lane(m);
            m &= m - 1;
            const unsigned o = __shfl_sync(kFull, off, src), c = __shfl_sync(kFull, cnt, src), seed = base + (unsigned)src;
            for (unsigned t = 0; t < c; t++) {
                int lo, hi;
                inst_edges(win.inst_pool[o + t], lo, hi);
                for (int f = lo + lane; f <= hi; f += 32) atomicMin(&Enew[f], seed);
            }
        }
    }
}

// did any epoch in the read-set change its meaning (< limit) between Ecur and Enew?
__device__ __forceinline__ bool readset_changed(const int2 *rs, unsigned cnt, const uint32_t *Ecur,
                                                const uint32_t *Enew, uint32_t limit, int lane, const uint32_t *__restrict__ diff)
{
    bool changed = false;
    for (unsigned base = 0; base < cnt && !changed; base += 32) {
        bool ch = false;
        if (base + lane < cnt) {
            int2 iv = rs[base + lane];
            if (range_dirty(diff, iv.x, iv.y, limit))
                for (int f = iv.x; f <= iv.y && !ch; f++) ch = (Ecur[f] < limit) != (Enew[f] < limit);
        }
        changed = __any_sync(kFull, ch);
    }
    return changed;
}

// diff[b] = min over the entries of [64 b, 64 b + 64) that differ between the two arrays of min(old, new)   (n is padded to
// a multiple of 128: whole warps of uint4, two blocks per warp)
__global__ void k_diff(const uint4 *__restrict__ Ecur, const uint4 *__restrict__ Enew, size_t n4, uint32_t *__restrict__ diff,
                       const Control *sched)
{
    if (sched && (sched->done | sched->halt)) return;
    if (sched && blockIdx.x == 0 && threadIdx.x == 0) const_cast<Control *>(sched)->t_mark[2] = global_ns();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += stride) { // n4 is a multiple of 32: whole warps
        const uint4 a = Ecur[t], b = Enew[t];
        uint32_t m = kFree;
        if (a.x != b.x) m = min(m, min(a.x, b.x));
        if (a.y != b.y) m = min(m, min(a.y, b.y));
        if (a.z != b.z) m = min(m, min(a.z, b.z));
        if (a.w != b.w) m = min(m, min(a.w, b.w));
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) m = min(m, __shfl_xor_sync(kFull, m, d)); // the 16 lanes of a block
        if ((threadIdx.x & 15) == 0) diff[t >> 4] = m;
    }
}

// one seed against the new epochs, by a whole warp: is its result still what an evaluation against them would give?
// The conflict status of the speculative result depends on replicated data only (results + epochs) and is recomputed by
// every rank for every seed; the read-sets live with the seed's owner (i % R == me), who alone decides about re-evaluations.
__device__ __forceinline__ void validate_seed(unsigned i, const uint32_t *__restrict__ Ecur, const uint32_t *__restrict__ Enew, unsigned phase,
                                              const Window &win, Control *ctl, const uint32_t *__restrict__ diff, int lane, bool own)
{
    const unsigned j = i & win.mask;
    const uint32_t T = (i / phase) * phase;
    const bool rs0 = own && readset_changed(win.rs_pool + win.rs_off[0][j], win.rs_cnt[0][j], Ecur, Enew, T, lane, diff);
    const bool was = win.conf[j] != 0;
    bool conf = false;
    if (win.res_cnt[0][j] > 1) conf = result_conflicts(win, j, Enew, i, lane, diff, was);
    if (lane == 0) win.conf[j] = conf;
    if (!own) return;
    unsigned dirty = 0;
    if (rs0) { // the speculative result itself is stale: evaluate again (and with it, if need be, the commit-time re-run)
        if (lane == 0) {
            if (win.heavy[j]) win.list_heavy[atomicAdd(&ctl->n_heavy, 1u)] = i + 1u;
            else win.list0[atomicAdd(&ctl->n0, 1u)] = i;
            win.has1[j] = 0;
        }
        dirty = 1;
    } else {
        if (conf != was) dirty = 1;
        bool rerun = false;
        if (conf) {
            if (!win.has1[j]) rerun = true;
            else rerun = readset_changed(win.rs_pool + win.rs_off[1][j], win.rs_cnt[1][j], Ecur, Enew, i, lane, diff);
        }
        if (lane == 0) {
            if (rerun) {
                win.has1[j] = 1;
                // evaluated together with the speculative work
                if (win.heavy[j]) win.list_heavy[atomicAdd(&ctl->n_heavy, 1u)] = (i | 0x80000000u) + 1u;
                else win.list0[atomicAdd(&ctl->n0, 1u)] = i | 0x80000000u;
            }
            if (!conf) win.has1[j] = 0;
        }
        if (rerun) dirty = 1;
    }
    if (lane == 0 && dirty) {
        atomicAdd(&ctl->dirty, 1u);
        atomicMin(&ctl->first_dirty, i);
    }
}

// Re-validate every seed of the active set against the new epochs; build next round's work lists.  With the change map
// (`diff`) one LANE per seed first looks whether any epoch the seed depends on changed at all -- its read-set intervals, the
// edges of its result (conflict status), the read-set of its commit-time re-run; almost always nothing did, and the seed costs
// a few cached byte loads.  The others get the full check by the whole warp.
__global__ void k_validate(const uint32_t *__restrict__ Ecur, const uint32_t *__restrict__ Enew, unsigned lo_seed,
                           unsigned hi_seed, unsigned phase, Window win, Control *ctl, int sched, const uint32_t *__restrict__ diff,
                           unsigned R, unsigned me)
{
    if (sched) {
        if (ctl->done | ctl->halt) return;
        lo_seed = ctl->c0, hi_seed = ctl->c1;
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->t_mark[3] = global_ns();
    }
    const int lane = threadIdx.x & 31;
    const unsigned warps = gridDim.x * (blockDim.x >> 5), warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (!diff) { // no change map (LCB_NO_DIFF=1): every seed gets the full check
        for (unsigned i = lo_seed + warp; i < hi_seed; i += warps) validate_seed(i, Ecur, Enew, phase, win, ctl, diff, lane, i % R == me);
        return;
    }
    constexpr unsigned kBatch = 16, kLaneIntervals = 64; // a lane looks at read-sets up to this size by itself, 16 loads in flight
    for (unsigned base = lo_seed + warp * 32u; base < hi_seed; base += warps * 32u) {
        const unsigned i = base + (unsigned)lane;
        bool need = false;
        if (i < hi_seed) {
            const unsigned j = i & win.mask;
            const bool own = i % R == me;
            const unsigned cnt0 = win.rs_cnt[0][j], rc = win.res_cnt[0][j];
            const bool c1 = own && win.conf[j] != 0;
            const unsigned cnt1 = c1 ? win.rs_cnt[1][j] : 0u;
            if (cnt0 > kLaneIntervals || cnt1 > kLaneIntervals || rc > 8u || (c1 && !win.has1[j])) {
                need = true; // big read-sets go to the whole warp right away (32 intervals at a time)
            } else {
                for (int slot = 0; slot < 2 && !need; slot++) {
                    const unsigned cnt = slot ? cnt1 : cnt0;
                    const uint32_t limit = slot ? i : (i / phase) * phase; // commit-time re-run / speculative evaluation
                    if (!cnt) continue;
                    const int2 *rs = win.rs_pool + win.rs_off[slot][j];
                    for (unsigned b0 = 0; b0 < cnt && !need; b0 += kBatch) {
                        int2 iv[kBatch];
#pragma unroll
                        for (unsigned t = 0; t < kBatch; t++) iv[t] = b0 + t < cnt ? rs[b0 + t] : make_int2(0, -1);
#pragma unroll
                        for (unsigned t = 0; t < kBatch; t++)
                            if (iv[t].x <= iv[t].y) need = need || range_dirty(diff, iv[t].x, iv[t].y, limit);
                    }
                }
                if (!need && rc > 1) {
                    const unsigned ro = win.res_off[0][j];
                    for (unsigned t = 0; t < rc && !need; t++) {
                        int lo, hi;
                        inst_edges(win.inst_pool[ro + t], lo, hi);
                        need = range_dirty(diff, lo, hi, i); // conflict test: edges claimed by a seed < i
                    }
                }
            }
        }
        unsigned m = __ballot_sync(kFull, need);
        while (m) {
            const int src = ffs_lane(m);
            m &= m - 1;
            validate_seed(base + (unsigned)src, Ecur, Enew, phase, win, ctl, diff, lane, (base + (unsigned)src) % R == me);
        }
    }
}

// admission of the seeds [lo, lo + n): fresh per-seed state; this rank's share (i % R == me) joins work list 0
__global__ void k_admit(unsigned lo, unsigned n, unsigned R, unsigned me, unsigned n0_before, Window win, Control *ctl, int sched)
{
    if (sched) { // k_round_begin has decided what to admit and has already set the new length of the work list
        if (ctl->done | ctl->halt) return;
        lo = ctl->admit_lo, n = ctl->admit, n0_before = ctl->admit_n0;
    }
    const unsigned first_own = lo + (me + R - lo % R) % R; // smallest i >= lo with i % R == me
    if (!sched && blockIdx.x == 0 && threadIdx.x == 0) ctl->n0 = n0_before + (lo + n > first_own ? (lo + n - first_own + R - 1) / R : 0u);
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const unsigned i = lo + t, j = i & win.mask;
        win.res_cnt[0][j] = win.res_cnt[1][j] = 0;
        win.rs_cnt[0][j] = win.rs_cnt[1][j] = 0;
        win.conf[j] = 0;
        win.has1[j] = 0;
        win.heavy[j] = 0;
        if (i % R == me) win.list0[n0_before + (i - first_own) / R] = i;
    }
}

struct SchedParams { // admission thresholds of the round loop (developer knobs, see lcb_find_blocks)
    float grow_below, min_round_ms, shrink_above, heavy_ms;
    unsigned phase, window_max, warps;
    unsigned long long inst_cap, rs_cap;
};

// start of a round: admission decision (what the host loop of lcb_find_blocks decides between two rounds)
__global__ void k_round_begin(Control *ctl, unsigned R, unsigned me)
{
    if (ctl->t_mark[4]) { // commit kernels of the round before (also after the last round)
        const unsigned long long now = global_ns();
        if (now > ctl->t_mark[4]) ctl->tail_ns[5] += now - ctl->t_mark[4];
        ctl->t_mark[4] = 0;
    }
    if (ctl->done | ctl->halt) return;
    if (ctl->c0 == ctl->c1) { // nothing active: no pool entry is referenced any more
        ctl->inst_used = 0, ctl->rs_used = 0;
        ctl->drain = 0;
        ctl->n0 = 0;
        ctl->windows++;
    }
    const unsigned c0 = ctl->c0, c1 = ctl->c1, S = ctl->n_seeds, cap = ctl->cap;
    unsigned admit = 0;
    if (!ctl->drain && !(ctl->hold && c1 > c0) && c1 < S && c1 - c0 < cap) admit = min(min(ctl->delta, S - c1), cap - (c1 - c0));
    ctl->admit_lo = c1, ctl->admit = admit, ctl->admit_n0 = ctl->n0;
    {
        const unsigned first_own = c1 + (me + R - c1 % R) % R; // smallest i >= c1 with i % R == me
        ctl->n0 += c1 + admit > first_own ? (c1 + admit - first_own + R - 1) / R : 0u;
    }
    ctl->c1 = c1 + admit;
    ctl->rounds++;
    ctl->head = 0, ctl->max_ns = 0, ctl->head_drain = 0, ctl->lean_exited = 0, ctl->lean_started = 0;
    ctl->t_first = ~0ull, ctl->t_last = 0;
    ctl->x_entries = 0, ctl->x_inst = 0;
}

struct Mirror { // pinned host memory the device writes at the end of every round
    volatile unsigned rounds_done, done, halt, c0;
    volatile unsigned n_heavy; // seeds already known to need the general kernel in the next round
    volatile unsigned heavy_last; // what the general kernel had to do in the last round (known + handed over)
};

// end of a round: admission rate for the next one, commit frontier, epoch buffer swap
__global__ void k_round_end(Control *ctl, SchedParams sp, Mirror *mirror, XchDev x)
{
    if (ctl->done | ctl->halt) { // a round queued behind the last one: nothing to commit (the emit kernels follow unconditionally)
        ctl->commit_n = 0;
        return;
    }
    if (ctl->inst_used * 2 > sp.inst_cap || ctl->rs_used * 2 > sp.rs_cap) ctl->drain = 1;
    {
        const unsigned long long now = global_ns();
        ctl->t_mark[4] = now;
        if (ctl->t_last && ctl->t_mark[0] > ctl->t_last) ctl->tail_ns[0] += ctl->t_mark[0] - ctl->t_last;
        for (int q = 0; q < 3; q++)
            if (ctl->t_mark[q + 1] > ctl->t_mark[q]) ctl->tail_ns[q + 1] += ctl->t_mark[q + 1] - ctl->t_mark[q];
        if (now > ctl->t_mark[3]) ctl->tail_ns[4] += now - ctl->t_mark[3];
    }
    float ms = ctl->t_last > ctl->t_first ? (float)(ctl->t_last - ctl->t_first) * 1e-6f : 0.f;
    unsigned n0_all = ctl->n0;
    if (x.R > 1) { // every rank combines the same R outcomes: same decisions everywhere
        const unsigned want = ctl->x_tag_base | ctl->rounds;
        for (int r = 0; r < x.R; r++) {
            if (r == x.me) continue;
            const volatile XHdr *h = (const volatile XHdr *)xch_region(x, x.me, r, ctl->rounds & 1u);
            const unsigned long long t0 = global_ns();
            while (h->ctl_tag != want) {
                __nanosleep(200);
                if (global_ns() - t0 > 60000000000ull) {
                    ctl->err = ctl->err ? ctl->err : (unsigned)LCB_ERR_CUDA;
                    break;
                }
            }
            ctl->first_dirty = min(ctl->first_dirty, h->ctl[0]);
            if (h->ctl[1] && !ctl->err) ctl->err = h->ctl[1];
            ctl->pool_overflow |= h->ctl[2];
            ctl->drain |= h->ctl[3];
            ms = fmaxf(ms, __uint_as_float(h->ctl[4]));
            ctl->max_ns = max(ctl->max_ns, h->ctl[5]);
            n0_all = max(n0_all, h->ctl[6]);
        }
    }
    // Admission rate: a launch lasts max(longest evaluation, work / resident warps).  While it is latency-bound more fresh
    // seeds are free; when the fresh work dominates, far-ahead speculation only adds re-evaluations.
    const float longest_ms = (float)ctl->max_ns * 1e-6f;
    const unsigned delta = ctl->delta;
    unsigned next_delta = delta, hold = 0;
    if (ms < sp.grow_below * longest_ms || ms < sp.min_round_ms) next_delta = (unsigned)min((unsigned long long)sp.window_max, 2ull * delta);
    else if (ms > sp.shrink_above * longest_ms && ms > 2 * sp.min_round_ms) next_delta = max(sp.phase, delta / 2 / sp.phase * sp.phase);
    if (longest_ms > sp.heavy_ms) { // heavy evaluations: fresh seeds are free only while warps are left over
        if (n0_all + sp.phase / (unsigned)max(x.R, 1) > sp.warps) hold = 1;
        else next_delta = min(next_delta, max(sp.phase, (sp.warps - n0_all) * (unsigned)max(x.R, 1) / sp.phase * sp.phase));
    }
    ctl->hold = hold;
    if (ctl->err | ctl->pool_overflow) { // the host handles both (error report / restart with half the active set)
        ctl->halt = 1;
        ctl->commit_n = 0;
    } else {
        const unsigned fd = min(ctl->first_dirty, ctl->c1);
        ctl->commit_first = ctl->c0;
        ctl->commit_n = fd > ctl->c0 ? fd - ctl->c0 : 0u;
        if (fd > ctl->c0) ctl->c0 = fd;
        ctl->parity ^= 1u;
        ctl->delta = min(next_delta, ctl->cap);
        if (ctl->c0 >= ctl->n_seeds) ctl->done = 1;
    }
    mirror->c0 = ctl->c0;
    mirror->n_heavy = ctl->n_heavy;
    mirror->heavy_last = ctl->heavy_done;
    mirror->done = ctl->done;
    mirror->halt = ctl->halt;
    __threadfence_system();
    mirror->rounds_done = ctl->rounds;
    __threadfence_system();
}


// Finalize (blocksfinder.h:312-332) for a converged window: block ids and output offsets are prefix
// sums in seed order (one block, W <= 65536)
// per-seed size of the final result (0 for seeds this rank does not own); summed across ranks before the scan
__global__ void k_final_counts(unsigned lo, unsigned n, Window win, unsigned *cnt_out, const Control *sched)
{
    if (sched) lo = sched->commit_first, n = (sched->halt ? 0u : sched->commit_n); // (commit_n of the round that set `done` still counts)
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        unsigned off, cnt;
        final_result(win, (lo + t) & win.mask, off, cnt);
        cnt_out[t] = cnt;
    }
}

__global__ void __launch_bounds__(1024) k_emit_scan(unsigned first, unsigned n, Window win, Control *ctl,
                                                     const unsigned *__restrict__ counts, int sched)
{
    // one CTA; warp w owns the contiguous segment [w * seg, (w + 1) * seg) and walks it 32 seeds at a time (coalesced loads,
    // shuffle scans): pass 1 totals per warp, a scan of the 32 totals, pass 2 the offsets
    __shared__ unsigned wb[32], wo[32];
    if (sched) {
        first = ctl->commit_first, n = (ctl->halt ? 0u : ctl->commit_n);
        if (n == 0) return;
    }
    const unsigned blocks_before = ctl->blocks_done, out_before = ctl->out_done; // updated after the barriers
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned seg = ((n + 31) / 32 + 31) / 32 * 32; // seeds per warp, a multiple of 32
    const unsigned lo = min(n, w * seg), hi = min(n, lo + seg);
    unsigned nb = 0, no = 0;
    for (unsigned t = lo + lane; t < hi; t += 32) {
        const unsigned cnt = counts[t];
        nb += cnt ? 1u : 0u;
        no += cnt;
    }
    nb = __reduce_add_sync(kFull, nb), no = __reduce_add_sync(kFull, no);
    if (lane == 0) wb[w] = nb, wo[w] = no;
    __syncthreads();
    if (w == 0) { // exclusive scan of the warp totals
        const unsigned vb = wb[lane], vo = wo[lane];
        unsigned ib = vb, io = vo;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned tb = __shfl_up_sync(kFull, ib, d), to = __shfl_up_sync(kFull, io, d);
            if ((int)lane >= d) ib += tb, io += to;
        }
        wb[lane] = ib - vb, wo[lane] = io - vo;
        if (lane == 31) {
            ctl->blocks_done = blocks_before + ib;
            ctl->out_done = out_before + io;
        }
    }
    __syncthreads();
    unsigned b = blocks_before + wb[w], o = out_before + wo[w]; // blocks / instances before this warp's segment
    for (unsigned base = lo; base < hi; base += 32) {
        const unsigned t = base + lane;
        const unsigned cnt = t < hi ? counts[t] : 0u;
        const unsigned has = cnt ? 1u : 0u;
        unsigned ib = has, io = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned tb = __shfl_up_sync(kFull, ib, d), to = __shfl_up_sync(kFull, io, d);
            if ((int)lane >= d) ib += tb, io += to;
        }
        if (t < hi) {
            const unsigned j = (first + t) & win.mask;
            win.out_off[j] = o + io - cnt;
            win.blk[j] = cnt ? b + ib : 0u;
        }
        b += __shfl_sync(kFull, ib, 31), o += __shfl_sync(kFull, io, 31);
    }
}

__global__ void k_emit_write(Index ix, int k, unsigned first, unsigned n, Window win, lcb_block_instance *out, Control *sched)
{
    if (sched) first = sched->commit_first, n = (sched->halt ? 0u : sched->commit_n);
    for (unsigned t_ = blockIdx.x * blockDim.x + threadIdx.x; t_ < n; t_ += gridDim.x * blockDim.x) {
    const unsigned j = (first + t_) & win.mask;
    unsigned off, cnt;
    final_result(win, j, off, cnt);
    for (unsigned t = 0; t < cnt; t++) {
        int4 b = win.inst_pool[off + t];
        const bool pos = b.x < 0;
        const int fg = b.x & 0x7FFFFFFF;
        int a = 0, e = ix.C;
        while (e - a > 1) {
            int mid = (a + e) >> 1;
            if ((int)ix.chr_off[mid] <= fg) a = mid;
            else e = mid;
        }
        lcb_block_instance r;
        r.chr = (uint32_t)a;
        if (pos) { // blocksfinder.h:320
            r.id = (int32_t)win.blk[j];
            r.start = (uint32_t)b.z;
            r.end = (uint32_t)b.w + (uint32_t)k;
        } else { // blocksfinder.h:324
            r.id = -(int32_t)win.blk[j];
            r.start = (uint32_t)b.w;
            r.end = (uint32_t)b.z + (uint32_t)k;
        }
        out[win.out_off[j] + t] = r;
    }
    }
}

// ------------------------------------------------------------------------------------------------
// seed enumeration (blocksfinder.h:461-503): one thread per signed vertex
// ------------------------------------------------------------------------------------------------
struct SeedArrays {
    int *vid;
    unsigned char *ch;
    unsigned *count;
    unsigned long long *rank;
    unsigned *res_pos, *res_chr;
};

template <bool FILL>
__global__ void k_seed_enum(Index ix, unsigned *__restrict__ per_vertex, const unsigned *__restrict__ offset, SeedArrays out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = 2LL * ix.V;
    if (t >= total) return;
    // t -> v in (-V, V): t = v + V
    const int v = (int)(t - ix.V);
    const int av = v < 0 ? -v : v;
    if (av == 0 || t == 0) { // v = -V does not exist (reference loops v in [-V+1, V-1]); vertex 0 is unused
        if (!FILL) per_vertex[t] = 0;
        return;
    }
    const unsigned o0 = ix.vtx_off[av], o1 = ix.vtx_off[av + 1];
    unsigned emitted = 0;
    unsigned w = FILL ? offset[t] : 0;
    for (unsigned a = o0; a < o1; a++) {
        const int2 oa = ix.occ[a];
        const int ga = oa.x & 0x7FFFFFFF;
        const bool pa = (oa.x < 0) == (v < 0);
        const int4 ra = ix.rec[ga];
        const unsigned char ca = pa ? rec_next_ch(ra) : rec_prev_rc(ra);
        bool first = true;
        for (unsigned b = o0; b < a && first; b++) {
            const int2 ob = ix.occ[b];
            const int4 rb = ix.rec[ob.x & 0x7FFFFFFF];
            const unsigned char cb = ((ob.x < 0) == (v < 0)) ? rec_next_ch(rb) : rec_prev_rc(rb);
            first = cb != ca;
        }
        if (!first) continue;
        unsigned count = 0;
        bool good = false;
        unsigned long long rank = 0, base = 1;
        unsigned long long best = ~0ULL; // (pos, chr) packed for the lexicographic min
        for (unsigned b = a; b < o1; b++) {
            const int2 ob = ix.occ[b];
            const int gb = ob.x & 0x7FFFFFFF;
            const int4 rb = ix.rec[gb];
            const bool pb = (ob.x < 0) == (v < 0);
            const unsigned char cb = pb ? rec_next_ch(rb) : rec_prev_rc(rb);
            if (cb != ca) continue;
            count++;
            int lo = 0, hi = ix.C;
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if ((int)ix.chr_off[mid] <= gb) lo = mid;
                else hi = mid;
            }
            rank += (unsigned long long)lo * base;
            base *= 31ULL;
            if (pb) {
                good = true;
                unsigned long long key = ((unsigned long long)(unsigned)rb.y << 32) | (unsigned)lo;
                best = key < best ? key : best;
            }
        }
        if (count > 1 && good) {
            if (FILL) {
                out.vid[w] = v;
                out.ch[w] = ca;
                out.count[w] = count;
                out.rank[w] = rank;
                out.res_pos[w] = (unsigned)(best >> 32);
                out.res_chr[w] = (unsigned)best;
                w++;
            }
            emitted++;
        }
    }
    if (!FILL) per_vertex[t] = emitted;
}

#include "device_prims.cuh" // scan, stable LSD radix sort, gather, iota (shared with graph_device.cu)

// ------------------------------------------------------------------------------------------------
// fused pipeline: JunctionStorage::Init (junctionstorage.h:572-650) on the device, from the junction records a
// resident graph (graph_device.cu) left there.  Same arrays as the host loader builds (lcb_host.cpp), same order.
// ------------------------------------------------------------------------------------------------
__global__ void k_fill_u32(unsigned *p, size_t n, unsigned v)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_ix_count(const int32_t *__restrict__ id, unsigned n, unsigned *__restrict__ cnt)
{
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) atomicAdd(&cnt[id[j] < 0 ? -id[j] : id[j]], 1u);
}
// abundance filter, strict < (junctionstorage.h:610)
__global__ void k_ix_keep(const int32_t *__restrict__ id, unsigned n, const unsigned *__restrict__ cnt, unsigned a, unsigned *__restrict__ keep)
{
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) keep[j] = cnt[id[j] < 0 ? -id[j] : id[j]] < a ? 1u : 0u;
}
__global__ void k_ix_vsize(const unsigned *__restrict__ cnt, unsigned V, unsigned a, unsigned *__restrict__ sz) // sz[V], sz[V+1] = 0
{
    const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V + 2) sz[v] = (v < V && cnt[v] < a) ? cnt[v] : 0u;
}
__global__ void k_ix_compact(const int32_t *__restrict__ id, const uint32_t *__restrict__ chr, const uint32_t *__restrict__ pos,
                             const unsigned *__restrict__ keep, const unsigned *__restrict__ gidx, unsigned n, int32_t *__restrict__ kid,
                             uint32_t *__restrict__ kbp, uint32_t *__restrict__ kchr, unsigned *__restrict__ key)
{
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || !keep[j]) return;
    const unsigned g = gidx[j];
    kid[g] = id[j], kbp[g] = pos[j], kchr[g] = chr[j];
    key[g] = (unsigned)(id[j] < 0 ? -id[j] : id[j]);
}
// rec[g]: id, position, occurrence list of |id|, and the two characters the traversal needs (junctionstorage.h:641-642)
__global__ void k_ix_rec(const int32_t *__restrict__ kid, const uint32_t *__restrict__ kbp, const uint32_t *__restrict__ kchr, unsigned N,
                         const uint32_t *__restrict__ vtx_off, const unsigned *__restrict__ cnt, const uint8_t *__restrict__ text,
                         const uint64_t *__restrict__ goff, int k, int4 *__restrict__ rec, unsigned *__restrict__ too_many)
{
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    const int id = kid[g];
    const unsigned a = (unsigned)(id < 0 ? -id : id), c = cnt[a];
    if (c > 65535u) *too_many = 1u;
    const uint64_t start = goff[kchr[g]], len = goff[kchr[g] + 1] - start - 1;
    const uint64_t p = kbp[g];
    auto upper = [](unsigned ch) { return (ch >= 'a' && ch <= 'z') ? ch - 32u : ch; };
    const unsigned next = p + (uint64_t)k < len ? upper(text[start + p + (uint64_t)k]) : 0u;
    unsigned prev = 'N';
    if (p > 0) {
        const unsigned ch = upper(text[start + p - 1]);
        prev = ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : ch == 'T' ? 'A' : 'N';
    }
    rec[g] = make_int4(id, (int)kbp[g], (int)vtx_off[a], (int)((c << 16) | (next << 8) | prev));
}
__global__ void k_ix_occ(const unsigned *__restrict__ perm, const int32_t *__restrict__ kid, const uint32_t *__restrict__ kbp, unsigned N,
                         int2 *__restrict__ occ)
{
    const unsigned o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= N) return;
    const unsigned g = perm[o];
    occ[o] = make_int2((int)(g | (kid[g] < 0 ? 0x80000000u : 0u)), (int)kbp[g]);
}
// device record layout from the arrays of an lcb_index_view (what lcb_index_pack does on the host):
//   rec[g] = {id, bp, first occurrence slot of |id|, (#occurrences << 16) | (next_ch << 8) | prev_rc}
//   occ[o] = {g | (stored id < 0 ? 1 << 31 : 0), bp} in occ_g order;  vtx_off as 32-bit (+ one padding entry)
__global__ void k_pack_view(const int32_t *__restrict__ pos_id, const uint32_t *__restrict__ pos_bp, const unsigned char *__restrict__ next_ch,
                            const unsigned char *__restrict__ prev_rc, const long long *__restrict__ occ_g,
                            const long long *__restrict__ vtx_off, size_t N, size_t nV, int4 *__restrict__ rec, int2 *__restrict__ occ,
                            uint32_t *__restrict__ vtx_off32, unsigned *__restrict__ too_many)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < N) {
        const int32_t id = pos_id[g];
        const long long a = id < 0 ? -(long long)id : (long long)id;
        const long long o0 = vtx_off[a], cnt = vtx_off[a + 1] - o0;
        if (cnt > 65535) *too_many = 1u;
        rec[g] = make_int4(id, (int)pos_bp[g], (int)o0, (int)(((unsigned)cnt << 16) | ((unsigned)next_ch[g] << 8) | (unsigned)prev_rc[g]));
        const long long og = occ_g[g]; // occurrence slot g of the CSR (not record g)
        occ[g] = make_int2((int)((unsigned)og | (pos_id[og] < 0 ? 0x80000000u : 0u)), (int)pos_bp[og]);
    }
    if (g < nV) vtx_off32[g] = (uint32_t)vtx_off[g];
    if (g == nV) vtx_off32[g] = (uint32_t)vtx_off[nV - 1]; // padding entry (V + 1 repeats V)
}

// chr_off[c] = first g of chromosome c (pre-filled with N: chromosomes without records own an empty range)
__global__ void k_ix_chroff(const uint32_t *__restrict__ kchr, unsigned N, uint32_t *__restrict__ chr_off)
{
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    const unsigned c1 = kchr[g];
    const unsigned c0 = g ? kchr[g - 1] + 1 : 0u;
    for (unsigned c = c0; c <= c1; c++) chr_off[c] = g;
}

} // namespace

// =================================================================================================
// host side
// =================================================================================================
struct lcb_ctx {
    std::string error;
    lcb_params prm;
    int device = 0, sms = 0;
    cudaStream_t stream = nullptr;
    // index
    Index ix{};
    int4 *d_rec = nullptr;
    int2 *d_occ = nullptr;
    uint32_t *d_vtx_off = nullptr, *d_chr_off = nullptr;
    uint32_t *d_E[2] = {nullptr, nullptr};
    // seeds
    uint64_t n_seeds = 0;
    bool seeds_ready = false;
    int *d_seed_vid = nullptr;
    unsigned char *d_seed_ch = nullptr;
    unsigned *d_seed_count = nullptr, *d_seed_res_pos = nullptr, *d_seed_res_chr = nullptr;
    unsigned long long *d_seed_rank = nullptr;
    unsigned max_seed_count = 0; // count of the first (= most abundant) bundle
    // window
    Window win{};
    unsigned wmax = 0;
    Control *d_ctl = nullptr, *h_ctl = nullptr;
    unsigned char *d_arena = nullptr;
    size_t arena_stride = 0;
    size_t d_big = 0;
    int grid_traverse = 0;
    int grid_lean = 0;         // 0: the common-case kernel is not used (see use_lean in create_end)
    int2 *d_lean_rs = nullptr; // its per-warp read-set logs
    lean::LInst *d_lean_shadow = nullptr; // ... and shadow copies of the instance table
    int2 *d_lean_hash2 = nullptr;         // ... and second-level path hashes (all-empty between evaluations)
    unsigned short *d_lean_hslot2 = nullptr;
    uint32_t *d_diff = nullptr; // change map of a round: one word per 64 epoch entries (k_diff; null: LCB_NO_DIFF=1)
    lcb_block_instance *d_out = nullptr;
    lcb_stats st{};
    int rank = 0, n_ranks = 1;
    XchDev xch{}; // R <= 1: no peers
#ifdef LCB_WITH_NCCL
    ncclComm_t comm = nullptr;
#endif
    unsigned *d_counts = nullptr;
    std::vector<void *> allocs, seed_allocs;
    std::vector<size_t> alloc_bytes, seed_alloc_bytes;
    bool arena_dirty = false;
    cudaEvent_t ev_step0 = nullptr, ev_step1 = nullptr;
    bool step_timed = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // device-driven round loop (single GPU)
    cudaStream_t stream2 = nullptr;                     // side stream: the general kernel beside the common-case kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    Mirror *h_mirror = nullptr;                         // pinned; written by k_round_end
    cudaGraphExec_t tail_graph[2] = {nullptr, nullptr}; // the non-traversal part of a round + the next round's admission, per epoch-buffer parity
    std::vector<cudaEvent_t> ev_ring;                   // event pairs around the traversal launches of the rounds in flight
};

namespace {

// Process-wide cache of device / pinned blocks: contexts are created and destroyed per call by hosts that hand
// over HOST arrays each time (bench.py's e2e leg, the CLI), and cudaMalloc/cudaFree of multi-GB scratch would
// otherwise dominate.  Blocks return to the cache in lcb_destroy and are really freed by lcb_trim_cache().
struct CachedBlock {
    void *p;
    size_t bytes;
    int device;   // -1: pinned host memory
    bool zeroed;  // arena invariant: every traversal leaves its spill hash all-zero
};
std::mutex g_cache_mu;
std::vector<CachedBlock> g_cache;
std::vector<std::pair<void *, size_t>> g_results; // page-locked result buffers handed out by lcb_find_blocks
#ifdef LCB_WITH_NCCL
ncclComm_t g_comm = nullptr; // reused by every context of this process (lcb_comm_init), freed by lcb_trim_cache
int g_comm_dev = -1, g_comm_rank = -1, g_comm_size = 0;
#endif
// peer mailboxes of this process (set up once per communicator by lcb_comm_init, shared by its contexts)
struct XchHost {
    bool ready = false;
    XchDev dev{};
    void *local = nullptr;
    unsigned run_seq = 0; // runs (lcb_find_blocks calls) so far: every rank counts the same
};
XchHost g_xch;
constexpr unsigned kXchEntries = 1u << 18, kXchInst = 1u << 20; // per (source rank, parity): 4 MB of entries, 16 MB of instances

// `want_zeroed`: the caller needs the arena invariant (null: any block will do).  A block tagged zeroed is handed
// only to such callers while an untagged one of the right size exists; *want_zeroed tells whether the invariant
// holds for the block returned (false: the caller clears it).
cudaError_t cached_alloc(void **p, size_t bytes, int device, bool *want_zeroed)
{
    bytes = (std::max<size_t>(bytes, 1) + 511) & ~(size_t)511;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        int best = -1;
        auto better = [&](const CachedBlock &a, const CachedBlock &b) { // is a the better choice than b?
            const bool wa = a.zeroed == (want_zeroed != nullptr), wb = b.zeroed == (want_zeroed != nullptr);
            if (wa != wb) return wa;
            return a.bytes < b.bytes;
        };
        for (size_t i = 0; i < g_cache.size(); i++)
            if (g_cache[i].device == device && g_cache[i].bytes >= bytes && g_cache[i].bytes <= bytes + bytes / 4 + 4096 &&
                (best < 0 || better(g_cache[i], g_cache[(size_t)best])))
                best = (int)i;
        if (best >= 0) {
            *p = g_cache[(size_t)best].p;
            if (want_zeroed) *want_zeroed = g_cache[(size_t)best].zeroed;
            g_cache.erase(g_cache.begin() + best);
            return cudaSuccess;
        }
    }
    if (want_zeroed) *want_zeroed = false;
    cudaError_t e = device < 0 ? cudaMallocHost(p, bytes) : cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation && lcb_cache_device_bytes(device) > 0) {
        // blocks parked by earlier contexts of other sizes are reclaimable memory: give them back and try once more
        cudaGetLastError();
        lcb_cache_trim_device(device);
        e = device < 0 ? cudaMallocHost(p, bytes) : cudaMalloc(p, bytes);
    }
    return e;
}

void cached_free(void *p, size_t bytes, int device, bool zeroed = false)
{
    bytes = (std::max<size_t>(bytes, 1) + 511) & ~(size_t)511;
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache.push_back(CachedBlock{p, bytes, device, zeroed});
}

} // namespace

size_t lcb_cache_device_bytes(int device)
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    size_t n = 0;
    for (const auto &b : g_cache)
        if (b.device == device) n += b.bytes;
    return n;
}

void lcb_cache_trim_device(int device)
{
    std::vector<CachedBlock> mine;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (size_t i = 0; i < g_cache.size();)
            if (g_cache[i].device == device) {
                mine.push_back(g_cache[i]);
                g_cache.erase(g_cache.begin() + (long)i);
            } else {
                i++;
            }
    }
    int cur = 0;
    if (device >= 0) cudaGetDevice(&cur), cudaSetDevice(device);
    for (auto &b : mine) {
        if (device < 0) cudaFreeHost(b.p);
        else cudaFree(b.p);
    }
    if (device >= 0) cudaSetDevice(cur);
}

namespace {

constexpr unsigned long long kInstPoolCap = 16ull << 20, kRsPoolCap = 128ull << 20;

size_t arena_stride_bytes() { return arena_stride_of(false); }
// entries of an epoch array: N + 32 (the traversal reads a little past the end), padded to whole warps of uint4 for k_diff
size_t epoch_len(size_t N) { return (N + 32 + 127) / 128 * 128; }
// one allocation: a per-warp arena for every resident warp, then the big slots
size_t arena_total_bytes(size_t warps) { return arena_stride_of(false) * warps + arena_stride_of(true) * (size_t)kBigSlots; }

template <typename T>
int dev_alloc(lcb_ctx *ctx, T **p, size_t n, bool *was_cached = nullptr)
{
    void *q = nullptr;
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    CUDA_TRY(cached_alloc(&q, bytes, ctx->device, was_cached));
    ctx->allocs.push_back(q);
    ctx->alloc_bytes.push_back(bytes);
    *p = (T *)q;
    return LCB_OK;
}

int exclusive_scan(lcb_ctx *ctx, const unsigned *in, unsigned *out, size_t n, unsigned *d_total, unsigned *d_tile_sum)
{
    unsigned tiles = (unsigned)((n + kScanTile - 1) / kScanTile);
    k_scan_tiles<<<tiles, 256, 0, ctx->stream>>>(in, out, d_tile_sum, n);
    k_scan_sums<<<1, 32, 0, ctx->stream>>>(d_tile_sum, tiles, d_total);
    k_scan_add<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(out, d_tile_sum, n);
    ctx->st.kernel_launches += 3;
    CUDA_TRY(cudaGetLastError());
    return LCB_OK;
}

// stable sort pass(es) of `perm` by one key word; skips digits on which all keys agree
template <typename K>
int radix_sort_word(lcb_ctx *ctx, unsigned **perm, unsigned **tmp, const K *key, int bits, bool descending, unsigned n,
                    unsigned *d_hist, unsigned *d_flag)
{
    unsigned blocks = (n + kSortTile - 1) / kSortTile;
    for (int shift = 0; shift < bits; shift += 8) {
        CUDA_TRY(cudaMemsetAsync(d_flag, 0, sizeof(unsigned), ctx->stream));
        k_radix_hist<K><<<blocks, 256, 0, ctx->stream>>>(*perm, key, shift, n, d_hist, descending ? 1 : 0);
        k_is_uniform_digit<<<256, 256, 0, ctx->stream>>>(d_hist, blocks, n, d_flag);
        unsigned flag = 0;
        CUDA_TRY(cudaMemcpyAsync(&flag, d_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->st.kernel_launches += 2;
        if (flag) continue;
        // digit-major exclusive scan of hist gives every (digit, block) its output base
        size_t hn = (size_t)256 * blocks;
        unsigned *d_total = d_flag + 1;
        int rc = exclusive_scan(ctx, d_hist, d_hist, hn, d_total, d_hist + hn);
        if (rc) return rc;
        k_radix_scatter<K><<<blocks, 256, 0, ctx->stream>>>(*perm, *tmp, key, shift, n, d_hist, descending ? 1 : 0);
        ctx->st.kernel_launches += 1;
        std::swap(*perm, *tmp);
    }
    CUDA_TRY(cudaGetLastError());
    return LCB_OK;
}

} // namespace

extern "C" void lcb_default_params(lcb_params *p)
{
    memset(p, 0, sizeof *p);
    p->k = 25;
    p->max_branch = 200;
    p->min_block = 200;
    p->max_flank = 200;
    p->looking_depth = 8;
    p->phase_size = 256;
    p->window_init = 16384;
    p->window_max = 1 << 20;
    p->device = 0;
    p->collect_counters = 0;
}

// Creates the CUDA context on `device` and parks the index-independent scratch (arena, pools) in the block cache,
// so a host can overlap it with parsing its inputs (the CLI does).  Optional; lcb_create works without it.
extern "C" int lcb_warmup(int device)
{
    const bool trace = getenv("LCB_LOAD_TRACE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (trace) fprintf(stderr, "[warmup] %-24s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    };
    if (cudaSetDevice(device) != cudaSuccess) return LCB_ERR_CUDA;
    if (cudaFree(nullptr) != cudaSuccess) return LCB_ERR_CUDA;
    lap("context");
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_traverse<false, LCB_TRAVERSE_CTAS_PER_SM>, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    const size_t arena_bytes = arena_total_bytes((size_t)per_sm * (size_t)sms * kWarpsPerBlock);
    void *arena = nullptr, *ip = nullptr, *rp = nullptr;
    bool cached = false;
    if (cached_alloc(&arena, arena_bytes, device, &cached) != cudaSuccess) return LCB_ERR_CUDA;
    lap("arena malloc");
    if (!cached && cudaMemset(arena, 0, arena_bytes) != cudaSuccess) return LCB_ERR_CUDA;
    if (cached_alloc(&ip, sizeof(int4) * kInstPoolCap, device, nullptr) != cudaSuccess) return LCB_ERR_CUDA;
    if (cached_alloc(&rp, sizeof(int2) * kRsPoolCap, device, nullptr) != cudaSuccess) return LCB_ERR_CUDA;
    cudaDeviceSynchronize();
    lap("pools + memset");
    {
        // CUDA loads kernels lazily, on first use: touch all of them here, off the critical path
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, k_traverse<false, LCB_TRAVERSE_CTAS_PER_SM>);
        cudaFuncGetAttributes(&fa, k_traverse<false, 5>);
        cudaFuncGetAttributes(&fa, k_traverse_lean);
        cudaFuncGetAttributes(&fa, k_rebase);
        cudaFuncGetAttributes(&fa, k_claim);
        cudaFuncGetAttributes(&fa, k_validate);
        cudaFuncGetAttributes(&fa, k_diff);
        cudaFuncGetAttributes(&fa, k_admit);
        cudaFuncGetAttributes(&fa, k_round_begin);
        cudaFuncGetAttributes(&fa, k_round_end);
        cudaFuncGetAttributes(&fa, k_final_counts);
        cudaFuncGetAttributes(&fa, k_emit_scan);
        cudaFuncGetAttributes(&fa, k_emit_write);
        cudaFuncGetAttributes(&fa, k_seed_enum<false>);
        cudaFuncGetAttributes(&fa, k_seed_enum<true>);
        cudaFuncGetAttributes(&fa, k_scan_tiles);
        cudaFuncGetAttributes(&fa, k_scan_sums);
        cudaFuncGetAttributes(&fa, k_scan_add);
        cudaFuncGetAttributes(&fa, k_radix_hist<unsigned>);
        cudaFuncGetAttributes(&fa, k_radix_hist<unsigned long long>);
        cudaFuncGetAttributes(&fa, k_radix_scatter<unsigned>);
        cudaFuncGetAttributes(&fa, k_radix_scatter<unsigned long long>);
        cudaFuncGetAttributes(&fa, k_is_uniform_digit);
        cudaFuncGetAttributes(&fa, k_iota);
        cudaFuncGetAttributes(&fa, k_gather<int>);
        cudaFuncGetAttributes(&fa, k_gather<unsigned>);
        cudaFuncGetAttributes(&fa, k_gather<unsigned char>);
        cudaFuncGetAttributes(&fa, k_gather<unsigned long long>);
        if (getenv("LCB_WARM_GRAPH")) lcg::preload_kernels(); // set by hosts that will also find the junctions (--construct)
        lap("kernel preload");
    }
    cached_free(arena, arena_bytes, device, true);
    cached_free(ip, sizeof(int4) * kInstPoolCap, device);
    cached_free(rp, sizeof(int2) * kRsPoolCap, device);
    return LCB_OK;
}

extern "C" const char *lcb_last_error(lcb_ctx *ctx) { return ctx ? ctx->error.c_str() : "null context"; }

extern "C" void lcb_destroy(lcb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (size_t i = 0; i < ctx->allocs.size(); i++) {
        // the arena goes back tagged "spill hash all-empty" unless a failed traversal left it in an unknown state
        const bool arena = ctx->allocs[i] == (void *)ctx->d_arena;
        cached_free(ctx->allocs[i], ctx->alloc_bytes[i], ctx->device, arena && !ctx->arena_dirty);
    }
    for (size_t i = 0; i < ctx->seed_allocs.size(); i++) cached_free(ctx->seed_allocs[i], ctx->seed_alloc_bytes[i], ctx->device);
    if (ctx->ev_step0) cudaEventDestroy(ctx->ev_step0);
    if (ctx->ev_step1) cudaEventDestroy(ctx->ev_step1);
    if (ctx->h_ctl) cached_free(ctx->h_ctl, sizeof(Control), -1);
    if (ctx->h_mirror) cached_free(ctx->h_mirror, sizeof(Mirror), -1);
    for (auto &g : ctx->tail_graph)
        if (g) cudaGraphExecDestroy(g);
    for (auto &e : ctx->ev_ring) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2), cudaStreamDestroy(ctx->stream2);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

namespace {

struct CreateTrace { // LCB_LOAD_TRACE=1: stage timer of lcb_create*
    bool on = getenv("LCB_LOAD_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void operator()(const char *what) const
    {
        if (on) fprintf(stderr, "[create] %-24s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
};

// parameters, device, stream, events
int create_begin(lcb_ctx *ctx, const lcb_params *params, const CreateTrace &lap)
{
    ctx->prm = *params;
    lcb_params &p = ctx->prm;
    if (p.phase_size <= 0) p.phase_size = 256;
    if (p.looking_depth <= 0) p.looking_depth = 8;
    if (p.window_init <= 0) p.window_init = 16384;
    if (p.window_max <= 0) p.window_max = 1 << 20;
    p.window_max = std::min(p.window_max, 1 << 20);
    p.window_max = std::max(p.phase_size, p.window_max / p.phase_size * p.phase_size);
    p.window_init = std::max(p.phase_size, std::min(p.window_init, p.window_max) / p.phase_size * p.phase_size);
    if (p.k <= 0 || p.max_branch < 0 || p.min_block < 0) {
        ctx->error = "bad parameters";
        return LCB_ERR_ARG;
    }
    if ((long long)p.max_branch + p.looking_depth >= 65536) { // a look-ahead walk has at most -b + depth junctions: 16 bits in the vote
        ctx->error = "max_branch + looking_depth must stay below 65536";
        return LCB_ERR_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        ctx->error = "no CUDA device: this library has no CPU fallback";
        return LCB_ERR_CUDA;
    }
    ctx->device = p.device;
    CUDA_TRY(cudaSetDevice(ctx->device));
    {
        // cudaGetDeviceProperties costs milliseconds per call: query the three attributes that matter instead
        int major = 0, sms = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, ctx->device));
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        if (major < 10) {
            ctx->error = "device " + std::to_string(ctx->device) + " is not sm_100-class";
            return LCB_ERR_CUDA;
        }
        ctx->sms = sms;
    }
    lap("device");
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    {
        int lo_prio = 0, hi_prio = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        CUDA_TRY(cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, hi_prio));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreate(&ctx->ev0));
    CUDA_TRY(cudaEventCreate(&ctx->ev1));
    CUDA_TRY(cudaEventCreate(&ctx->ev_step0));
    CUDA_TRY(cudaEventCreate(&ctx->ev_step1));
    lap("stream + events");
    return LCB_OK;
}

// window state, pools, arena: everything that does not depend on where the index came from
int create_end(lcb_ctx *ctx, const CreateTrace &lap)
{
    lcb_params &p = ctx->prm;
    const int64_t N = ctx->ix.N;
    // ---- window state, pools, arena ----
    int rc;
    unsigned W = 256; // ring of per-seed state: a power of two >= the largest active set
    while (W < (unsigned)p.window_max) W <<= 1;
    ctx->wmax = W;
    ctx->win.mask = W - 1;
    for (int s = 0; s < 2; s++) {
        if ((rc = dev_alloc(ctx, &ctx->win.res_off[s], W))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->win.res_cnt[s], W))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->win.rs_off[s], W))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->win.rs_cnt[s], W))) return rc;
    }
    if ((rc = dev_alloc(ctx, &ctx->win.conf, W))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->win.has1, W))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->win.blk, W))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->win.out_off, W))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->win.list0, W))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->win.list_heavy, W))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->win.heavy, W))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_counts, W))) return rc;
    ctx->win.inst_cap = kInstPoolCap;
    ctx->win.rs_cap = kRsPoolCap;
    if (const char *e = getenv("LCB_TEST_POOL_ENTRIES")) { // testing aid: tiny result pools force the window-halving retry
        unsigned long long v = strtoull(e, nullptr, 10);
        if (v >= 1024) ctx->win.inst_cap = std::min(ctx->win.inst_cap, v), ctx->win.rs_cap = std::min(ctx->win.rs_cap, v);
    }
    lap("ring arrays");
    if ((rc = dev_alloc(ctx, &ctx->win.inst_pool, (size_t)ctx->win.inst_cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->win.rs_pool, (size_t)ctx->win.rs_cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_ctl, 1))) return rc;
    CUDA_TRY(cached_alloc((void **)&ctx->h_ctl, sizeof(Control), -1, nullptr));
    CUDA_TRY(cudaMemsetAsync(ctx->d_ctl, 0, sizeof(Control), ctx->stream));
    lap("pools + control");
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_traverse<false, LCB_TRAVERSE_CTAS_PER_SM>, kThreads, 0));
    if (per_sm < 1) per_sm = 1;
    ctx->grid_traverse = per_sm * ctx->sms;
    // The common-case kernel (lcb_lean.cuh) takes every work item first when its preconditions hold for the whole run;
    // LCB_NO_LEAN=1 keeps the general kernel alone (A/B runs, and the step counters live only there).
    ctx->grid_lean = 0;
    if (ctx->ix.C + 1 <= lean::kLChr && p.max_flank <= 32767 && !p.collect_counters && !getenv("LCB_NO_LEAN")) {
        int lean_per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lean_per_sm, k_traverse_lean, kThreads, 0));
        ctx->grid_lean = std::max(lean_per_sm, 1) * ctx->sms;
        if ((rc = dev_alloc(ctx, &ctx->d_lean_rs, (size_t)ctx->grid_lean * kWarpsPerBlock * kLeanRs))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_lean_shadow, (size_t)ctx->grid_lean * kWarpsPerBlock * lean::kLInst))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_lean_hash2, (size_t)ctx->grid_lean * kWarpsPerBlock * lean::kLHash2))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_lean_hslot2, (size_t)ctx->grid_lean * kWarpsPerBlock * lean::kLPath2))) return rc;
        CUDA_TRY(cudaMemsetAsync(ctx->d_lean_hash2, 0, sizeof(int2) * (size_t)ctx->grid_lean * kWarpsPerBlock * lean::kLHash2, ctx->stream));
    }
    lap("occupancy query");
    ctx->arena_stride = arena_stride_bytes();
    const size_t arena_bytes = arena_total_bytes((size_t)ctx->grid_traverse * kWarpsPerBlock);
    ctx->d_big = ctx->arena_stride * (size_t)ctx->grid_traverse * kWarpsPerBlock; // offset of the big slots inside the arena
    bool arena_cached = false;
    if ((rc = dev_alloc(ctx, &ctx->d_arena, arena_bytes, &arena_cached))) return rc;
    if (!arena_cached) CUDA_TRY(cudaMemsetAsync(ctx->d_arena, 0, arena_bytes, ctx->stream)); // the spill hash must start all-empty
    if ((rc = dev_alloc(ctx, &ctx->d_out, (size_t)N + 1))) return rc;
    if (!getenv("LCB_NO_DIFF") && (rc = dev_alloc(ctx, &ctx->d_diff, epoch_len((size_t)N) / 64 + 2))) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    lap("window state + arena");
    return LCB_OK;
}

} // namespace

namespace {

// pack + upload the index from host arrays (the H2D part of lcb_create)
int create_upload(lcb_ctx *ctx, const lcb_index_view *v, const CreateTrace &lap)
{
    if (v->n_records < 0 || v->n_records >= (int64_t)0x7FFFFFF0 || v->n_vertices >= (int64_t)0x3FFFFFF0 || v->n_chr < 0) {
        ctx->error = "index too large for 32-bit device indices";
        return LCB_ERR_ARG;
    }
    const int64_t N = v->n_records, V = v->n_vertices;
    const int C = v->n_chr;
    auto t0 = std::chrono::steady_clock::now();
    // ---- pack + upload the index (8 B + 2 B per record, u32 CSR) ----
    if (v->packed_rec && v->packed_occ) {
        // the host already holds the device layout (lcb_index_pack): straight copies, no pinned staging to allocate
        const size_t b_rec = sizeof(int4) * (size_t)N, b_occ = sizeof(int2) * (size_t)N;
        std::vector<uint32_t> vo((size_t)V + 2), co((size_t)C + 1);
        for (int64_t i = 0; i <= V; i++) vo[(size_t)i] = (uint32_t)v->vtx_off[i];
        vo[(size_t)V + 1] = vo[(size_t)V];
        for (int i = 0; i <= C; i++) co[(size_t)i] = (uint32_t)v->chr_off[i];
        int rc;
        if ((rc = dev_alloc(ctx, &ctx->d_rec, (size_t)N))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_occ, (size_t)N))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_vtx_off, (size_t)V + 2))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_chr_off, (size_t)C + 1))) return rc;
        for (int e = 0; e < 2; e++)
            if ((rc = dev_alloc(ctx, &ctx->d_E[e], epoch_len((size_t)N)))) return rc;
        lap("index alloc");
        CUDA_TRY(cudaMemcpyAsync(ctx->d_rec, v->packed_rec, b_rec, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(ctx->d_occ, v->packed_occ, b_occ, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(ctx->d_vtx_off, vo.data(), vo.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(ctx->d_chr_off, co.data(), co.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->st.h2d_bytes = (uint64_t)(b_rec + b_occ + (vo.size() + co.size()) * sizeof(uint32_t));
    } else {
        // The view's arrays go to the device as they are (18 bytes per record instead of the 24 packed ones), staged through
        // ONE page-locked buffer by the host threads, chunk by chunk with the DMA copies right behind them; the device record
        // layout is built there (k_pack_view).  Packing on the host first cost 2-3x as long for the same result.
        const size_t nN = (size_t)N, nV = (size_t)V + 1;
        auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
        const size_t o_id = 0, o_bp = o_id + up(4 * nN), o_nc = o_bp + up(4 * nN), o_pr = o_nc + up(nN), o_og = o_pr + up(nN),
                     o_vo = o_og + up(8 * nN), o_end = o_vo + up(8 * nV);
        unsigned char *stage = nullptr;
        CUDA_TRY(cached_alloc((void **)&stage, o_end, -1, nullptr));
        struct Unpin {
            unsigned char *p;
            size_t n;
            ~Unpin() { cached_free(p, n, -1); }
        } unpin{stage, o_end};
        int rc;
        unsigned char *d_raw = nullptr;
        unsigned *d_flag = nullptr;
        if ((rc = dev_alloc(ctx, &d_raw, o_end))) return rc;
        if ((rc = dev_alloc(ctx, &d_flag, 4))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_rec, nN))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_occ, nN))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_vtx_off, nV + 1))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_chr_off, (size_t)C + 1))) return rc;
        for (int e = 0; e < 2; e++)
            if ((rc = dev_alloc(ctx, &ctx->d_E[e], epoch_len(nN)))) return rc;
        CUDA_TRY(cudaMemsetAsync(d_flag, 0, sizeof(unsigned), ctx->stream));
        unsigned T = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        if (const unsigned cap = lcb_host_thread_cap.load()) T = std::min(T, cap); // lcb_set_host_threads (the CLI's -t)
        struct Part {
            const void *src;
            size_t off, elem, count;
        };
        const Part parts[6] = {{v->pos_id, o_id, 4, nN}, {v->pos_bp, o_bp, 4, nN}, {v->next_ch, o_nc, 1, nN},
                               {v->prev_rc, o_pr, 1, nN}, {v->occ_g, o_og, 8, nN},  {v->vtx_off, o_vo, 8, nV}};
        const size_t chunk = 2u << 20; // elements of every array per round of copies
        const size_t longest = std::max(nN, nV);
        for (size_t c0 = 0; c0 < longest; c0 += chunk) {
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < T; t++)
                pool.emplace_back([&, t]() {
                    for (const Part &p : parts) {
                        if (c0 >= p.count) continue;
                        const size_t n = std::min(chunk, p.count - c0), a = n * t / T, b = n * (t + 1) / T;
                        memcpy(stage + p.off + (c0 + a) * p.elem, (const unsigned char *)p.src + (c0 + a) * p.elem, (b - a) * p.elem);
                    }
                });
            for (auto &th : pool) th.join();
            for (const Part &p : parts) {
                if (c0 >= p.count) continue;
                const size_t n = std::min(chunk, p.count - c0);
                CUDA_TRY(cudaMemcpyAsync(d_raw + p.off + c0 * p.elem, stage + p.off + c0 * p.elem, n * p.elem, cudaMemcpyHostToDevice, ctx->stream));
            }
        }
        lap("stage + copies queued");
        std::vector<uint32_t> co((size_t)C + 1);
        for (int i = 0; i <= C; i++) co[(size_t)i] = (uint32_t)v->chr_off[i];
        CUDA_TRY(cudaMemcpyAsync(ctx->d_chr_off, co.data(), co.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        k_pack_view<<<(unsigned)((std::max(nN, nV + 1) + 255) / 256), 256, 0, ctx->stream>>>(
            (const int32_t *)(d_raw + o_id), (const uint32_t *)(d_raw + o_bp), d_raw + o_nc, d_raw + o_pr, (const long long *)(d_raw + o_og),
            (const long long *)(d_raw + o_vo), nN, nV, ctx->d_rec, ctx->d_occ, ctx->d_vtx_off, d_flag);
        unsigned too_many = 0;
        CUDA_TRY(cudaMemcpyAsync(&too_many, d_flag, sizeof too_many, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaGetLastError());
        if (too_many) {
            ctx->error = "a junction occurs more than 65535 times: lower the abundance threshold (-a)";
            return LCB_ERR_ARG;
        }
        { // the raw copy is not needed any more: back to the block cache now, not at destroy
            for (size_t q = 0; q < ctx->allocs.size(); q++)
                if (ctx->allocs[q] == (void *)d_raw) {
                    cached_free(ctx->allocs[q], ctx->alloc_bytes[q], ctx->device);
                    ctx->allocs.erase(ctx->allocs.begin() + (long)q);
                    ctx->alloc_bytes.erase(ctx->alloc_bytes.begin() + (long)q);
                    break;
                }
        }
        const size_t b_rec = 18 * nN, b_occ = 0, b_vo = 8 * nV, b_co = sizeof(uint32_t) * ((size_t)C + 1);
        ctx->st.h2d_bytes = (uint64_t)(b_rec + b_occ + b_vo + b_co);
    }
    ctx->st.ms_h2d = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    lap("index alloc + h2d");
    ctx->ix.rec = ctx->d_rec;
    ctx->ix.vtx_off = ctx->d_vtx_off;
    ctx->ix.occ = ctx->d_occ;
    ctx->ix.chr_off = ctx->d_chr_off;
    ctx->ix.C = C;
    ctx->ix.N = (int)N;
    ctx->ix.V = (int)V;
    ctx->st.n_records = (uint64_t)N;
    ctx->st.n_vertices = (uint64_t)V;
    return LCB_OK;
}

} // namespace

extern "C" int lcb_create(const lcb_index_view *v, const lcb_params *params, lcb_ctx **out)
{
    if (!v || !params || !out) return LCB_ERR_ARG;
    lcb_ctx *ctx = new lcb_ctx;
    *out = ctx; // returned even on failure so that lcb_last_error works; caller destroys it
    CreateTrace lap;
    int rc;
    if ((rc = create_begin(ctx, params, lap))) return rc;
    if ((rc = create_upload(ctx, v, lap))) return rc;
    return create_end(ctx, lap);
}

extern "C" int lcb_create_shared(const lcb_index_view *v, const lcb_params *params, int rank, int n_ranks, const void *id_bytes, lcb_ctx **out)
{
    if (!params || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks || (rank == 0 && !v)) return LCB_ERR_ARG;
    if (n_ranks == 1) return lcb_create(v, params, out);
    lcb_ctx *ctx = new lcb_ctx;
    *out = ctx;
#ifdef LCB_WITH_NCCL
    CreateTrace lap;
    int rc;
    if ((rc = create_begin(ctx, params, lap))) return rc;
    if ((rc = lcb_comm_init(ctx, rank, n_ranks, id_bytes))) return rc;
    lap("communicator");
    // rank 0 uploads (one PCIe copy), everybody else receives over NVLink; a header first so that the receivers can allocate
    long long hdr[4] = {0, 0, 0, 0}; // N, V, C, rank 0's status
    if (rank == 0) {
        hdr[3] = create_upload(ctx, v, lap);
        hdr[0] = ctx->ix.N, hdr[1] = ctx->ix.V, hdr[2] = ctx->ix.C;
    }
    long long *d_hdr = nullptr;
    if ((rc = dev_alloc(ctx, &d_hdr, 4))) return rc;
    CUDA_TRY(cudaMemcpyAsync(d_hdr, hdr, sizeof hdr, cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(ncclBroadcast(d_hdr, d_hdr, sizeof hdr, ncclUint8, 0, ctx->comm, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(hdr, d_hdr, sizeof hdr, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (hdr[3]) {
        if (rank) ctx->error = "rank 0 could not upload the index";
        return (int)hdr[3];
    }
    const size_t N = (size_t)hdr[0], V = (size_t)hdr[1];
    const int C = (int)hdr[2];
    if (rank) {
        auto t0 = std::chrono::steady_clock::now();
        if ((rc = dev_alloc(ctx, &ctx->d_rec, N))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_occ, N))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_vtx_off, V + 2))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_chr_off, (size_t)C + 1))) return rc;
        for (int e = 0; e < 2; e++)
            if ((rc = dev_alloc(ctx, &ctx->d_E[e], epoch_len(N)))) return rc;
        ctx->ix.rec = ctx->d_rec, ctx->ix.occ = ctx->d_occ, ctx->ix.vtx_off = ctx->d_vtx_off, ctx->ix.chr_off = ctx->d_chr_off;
        ctx->ix.C = C, ctx->ix.N = (int)N, ctx->ix.V = (int)V;
        ctx->st.n_records = N, ctx->st.n_vertices = V;
        ctx->st.h2d_bytes = 0;
        ctx->st.ms_h2d = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    NCCL_TRY(ncclBroadcast(ctx->d_rec, ctx->d_rec, sizeof(int4) * N, ncclUint8, 0, ctx->comm, ctx->stream));
    NCCL_TRY(ncclBroadcast(ctx->d_occ, ctx->d_occ, sizeof(int2) * N, ncclUint8, 0, ctx->comm, ctx->stream));
    NCCL_TRY(ncclBroadcast(ctx->d_vtx_off, ctx->d_vtx_off, sizeof(uint32_t) * (V + 2), ncclUint8, 0, ctx->comm, ctx->stream));
    NCCL_TRY(ncclBroadcast(ctx->d_chr_off, ctx->d_chr_off, sizeof(uint32_t) * ((size_t)C + 1), ncclUint8, 0, ctx->comm, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    lap("index broadcast");
    return create_end(ctx, lap);
#else
    (void)id_bytes;
    ctx->error = "multi-GPU support is not compiled in (NCCL not found at build time)";
    return LCB_ERR_STATE;
#endif
}

extern "C" int lcb_create_from_graph(const lcg_graph *graph, lcb_index *index, int abundance, const lcb_params *params, lcb_ctx **out)
{
    if (!graph || !index || !params || !out || abundance < 0) return LCB_ERR_ARG;
    lcb_ctx *ctx = new lcb_ctx;
    *out = ctx;
    const lcg::Resident *res = lcg::resident_of(graph);
    if (!res) {
        ctx->error = "the graph holds no device-resident data (build it with lcg_build_resident)";
        return LCB_ERR_STATE;
    }
    if (res->n_junctions >= 0x7FFFFFF0ull || res->n_vertices >= 0x3FFFFFF0ull) {
        ctx->error = "index too large for 32-bit device indices";
        return LCB_ERR_ARG;
    }
    if (params->device != res->device || params->k != res->k) {
        ctx->error = "the graph was built on another device or with another k";
        return LCB_ERR_ARG;
    }
    CreateTrace lap;
    int rc = create_begin(ctx, params, lap);
    if (rc) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    const unsigned nj = (unsigned)res->n_junctions, V = (unsigned)res->n_vertices;
    const int C = nj ? (int)res->last_chr + 1 : 0;
    {
        std::string e;
        if ((rc = lcb_index_set_chromosomes(index, C, res->k, e))) {
            ctx->error = e;
            return rc;
        }
    }
    // temporaries of this function go back to the block cache on every exit path
    struct Temps {
        lcb_ctx *c;
        std::vector<std::pair<void *, size_t>> v;
        ~Temps()
        {
            cudaStreamSynchronize(c->stream);
            for (auto &b : v) cached_free(b.first, b.second, c->device);
        }
    } temps{ctx, {}};
    auto talloc = [&](auto **p, size_t n) -> int {
        void *q = nullptr;
        const size_t bytes = std::max<size_t>(n, 1) * sizeof(**p);
        CUDA_TRY(cached_alloc(&q, bytes, ctx->device, nullptr));
        temps.v.emplace_back(q, bytes);
        *p = (std::remove_reference_t<decltype(**p)> *)q;
        return LCB_OK;
    };
    unsigned *d_cnt = nullptr, *d_keep = nullptr, *d_gidx = nullptr, *d_tile = nullptr, *d_small = nullptr, *d_sz = nullptr;
    if ((rc = talloc(&d_cnt, (size_t)V + 2))) return rc;
    if ((rc = talloc(&d_keep, nj))) return rc;
    if ((rc = talloc(&d_gidx, nj))) return rc;
    if ((rc = talloc(&d_tile, std::max<size_t>(nj, (size_t)V + 2) / kScanTile + 8))) return rc;
    if ((rc = talloc(&d_small, 8))) return rc;
    if ((rc = talloc(&d_sz, (size_t)V + 2))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_vtx_off, (size_t)V + 2))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_chr_off, (size_t)C + 1))) return rc;
    CUDA_TRY(cudaMemsetAsync(d_cnt, 0, ((size_t)V + 2) * sizeof(unsigned), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(d_small, 0, 8 * sizeof(unsigned), ctx->stream));
    const unsigned jb = (nj + 255) / 256, vb = (V + 2 + 255) / 256;
    unsigned N = 0;
    if (nj) {
        k_ix_count<<<jb, 256, 0, ctx->stream>>>(res->d_id, nj, d_cnt);
        k_ix_keep<<<jb, 256, 0, ctx->stream>>>(res->d_id, nj, d_cnt, (unsigned)abundance, d_keep);
        if ((rc = exclusive_scan(ctx, d_keep, d_gidx, nj, d_small, d_tile))) return rc;
        CUDA_TRY(cudaMemcpyAsync(&N, d_small, sizeof N, cudaMemcpyDeviceToHost, ctx->stream));
    }
    k_ix_vsize<<<vb, 256, 0, ctx->stream>>>(d_cnt, V, (unsigned)abundance, d_sz);
    if ((rc = exclusive_scan(ctx, d_sz, ctx->d_vtx_off, (size_t)V + 2, d_small + 1, d_tile))) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaGetLastError());
    lap("count + filter");
    int32_t *d_kid = nullptr;
    uint32_t *d_kbp = nullptr, *d_kchr = nullptr;
    unsigned *d_key = nullptr, *d_perm = nullptr, *d_tmp = nullptr, *d_hist = nullptr;
    if ((rc = talloc(&d_kid, N))) return rc;
    if ((rc = talloc(&d_kbp, N))) return rc;
    if ((rc = talloc(&d_kchr, N))) return rc;
    if ((rc = talloc(&d_key, N))) return rc;
    if ((rc = talloc(&d_perm, N))) return rc;
    if ((rc = talloc(&d_tmp, N))) return rc;
    const unsigned sblocks = (N + kSortTile - 1) / kSortTile;
    if ((rc = talloc(&d_hist, (size_t)256 * sblocks + ((size_t)256 * sblocks) / kScanTile + 8))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_rec, (size_t)N))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_occ, (size_t)N))) return rc;
    for (int e = 0; e < 2; e++)
        if ((rc = dev_alloc(ctx, &ctx->d_E[e], epoch_len((size_t)N)))) return rc;
    k_fill_u32<<<(unsigned)((C + 1 + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_chr_off, (size_t)C + 1, N);
    if (N) {
        const unsigned nbk = (N + 255) / 256;
        k_ix_compact<<<jb, 256, 0, ctx->stream>>>(res->d_id, res->d_chr, res->d_pos, d_keep, d_gidx, nj, d_kid, d_kbp, d_kchr, d_key);
        k_ix_rec<<<nbk, 256, 0, ctx->stream>>>(d_kid, d_kbp, d_kchr, N, ctx->d_vtx_off, d_cnt, res->d_text, res->d_goff, res->k, ctx->d_rec, d_small + 2);
        k_ix_chroff<<<nbk, 256, 0, ctx->stream>>>(d_kchr, N, ctx->d_chr_off);
        // occurrence lists: stable sort of g by |id| leaves every list in (chr, idx) order (junctionstorage.h:646-649)
        k_iota<<<nbk, 256, 0, ctx->stream>>>(d_perm, N);
        int bits = 8;
        while (bits < 32 && (V >> bits)) bits += 8;
        if ((rc = radix_sort_word<unsigned>(ctx, &d_perm, &d_tmp, d_key, bits, false, N, d_hist, d_small + 4))) return rc;
        k_ix_occ<<<nbk, 256, 0, ctx->stream>>>(d_perm, d_kid, d_kbp, N, ctx->d_occ);
    }
    unsigned flags[8] = {0};
    CUDA_TRY(cudaMemcpyAsync(flags, d_small, sizeof flags, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaGetLastError());
    if (flags[2]) {
        ctx->error = "a junction occurs more than 65535 times: lower the abundance threshold (-a)";
        return LCB_ERR_ARG;
    }
    lap("device index build");
    ctx->st.ms_h2d = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    ctx->st.h2d_bytes = 0;
    ctx->ix.rec = ctx->d_rec;
    ctx->ix.vtx_off = ctx->d_vtx_off;
    ctx->ix.occ = ctx->d_occ;
    ctx->ix.chr_off = ctx->d_chr_off;
    ctx->ix.C = C;
    ctx->ix.N = (int)N;
    ctx->ix.V = (int)V;
    ctx->st.n_records = N;
    ctx->st.n_vertices = V;
    return create_end(ctx, lap);
}

extern "C" int lcb_comm_unique_id(void *id_bytes)
{
#ifdef LCB_WITH_NCCL
    if (!nccl_api()->ok) return LCB_ERR_STATE;
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return LCB_ERR_CUDA;
    static_assert(sizeof(id) <= LCB_COMM_ID_BYTES, "id size");
    memset(id_bytes, 0, LCB_COMM_ID_BYTES);
    memcpy(id_bytes, &id, sizeof id);
    return LCB_OK;
#else
    (void)id_bytes;
    return LCB_ERR_STATE;
#endif
}

extern "C" int lcb_comm_init(lcb_ctx *ctx, int rank, int n_ranks, const void *id_bytes)
{
    if (!ctx || n_ranks < 1 || rank < 0 || rank >= n_ranks) return LCB_ERR_ARG;
    if (n_ranks == 1) {
        ctx->rank = 0, ctx->n_ranks = 1;
        return LCB_OK;
    }
#ifdef LCB_WITH_NCCL
    if (!id_bytes) return LCB_ERR_ARG;
    if (!nccl_api()->ok) {
        ctx->error = "libnccl.so.2 could not be loaded (set LCB_NCCL_LIB)";
        return LCB_ERR_STATE;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    {
        // communicators are expensive (seconds) and independent of the index: keep one per (device, rank, size)
        std::lock_guard<std::mutex> lk(g_cache_mu);
        if (g_comm && g_comm_dev == ctx->device && g_comm_rank == rank && g_comm_size == n_ranks && g_xch.ready) {
            ctx->comm = g_comm;
            ctx->rank = rank, ctx->n_ranks = n_ranks;
            ctx->xch = g_xch.dev;
            return LCB_OK;
        }
    }
    if (n_ranks > kMaxRanks) {
        ctx->error = "at most 8 ranks (one node)";
        return LCB_ERR_ARG;
    }
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    ncclComm_t comm = nullptr;
    ncclResult_t r = ncclCommInitRank(&comm, n_ranks, id, rank);
    if (r != ncclSuccess) {
        ctx->error = std::string("ncclCommInitRank: ") + ncclGetErrorString(r);
        return LCB_ERR_CUDA;
    }
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        if (g_comm) ncclCommDestroy(g_comm);
        g_comm = comm, g_comm_dev = ctx->device, g_comm_rank = rank, g_comm_size = n_ranks;
    }
    ctx->comm = comm;
    ctx->rank = rank, ctx->n_ranks = n_ranks;
    // ---- peer mailboxes: one allocation per rank, opened by every peer through CUDA IPC (handles travel once over NCCL)
    {
        if (g_xch.local) cudaFree(g_xch.local); // (a communicator of another shape was replaced)
        g_xch = XchHost{};
        const size_t bytes = xch_region_bytes(kXchEntries, kXchInst) * 2 * (size_t)n_ranks;
        CUDA_TRY(cudaMalloc(&g_xch.local, bytes));
        CUDA_TRY(cudaMemset(g_xch.local, 0, bytes));
        cudaIpcMemHandle_t mine;
        CUDA_TRY(cudaIpcGetMemHandle(&mine, g_xch.local));
        cudaIpcMemHandle_t *d_all = nullptr;
        CUDA_TRY(cudaMalloc((void **)&d_all, sizeof(cudaIpcMemHandle_t) * (size_t)n_ranks));
        CUDA_TRY(cudaMemcpy(d_all + rank, &mine, sizeof mine, cudaMemcpyHostToDevice));
        NCCL_TRY(ncclAllGather(d_all + rank, d_all, sizeof mine, ncclUint8, comm, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        std::vector<cudaIpcMemHandle_t> all((size_t)n_ranks);
        CUDA_TRY(cudaMemcpy(all.data(), d_all, sizeof mine * (size_t)n_ranks, cudaMemcpyDeviceToHost));
        cudaFree(d_all);
        g_xch.dev.R = n_ranks, g_xch.dev.me = rank;
        g_xch.dev.ent_cap = kXchEntries, g_xch.dev.inst_cap = kXchInst;
        for (int r = 0; r < n_ranks; r++) {
            if (r == rank) {
                g_xch.dev.box[r] = (unsigned char *)g_xch.local;
                continue;
            }
            void *p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                ctx->error = std::string("cudaIpcOpenMemHandle (peer mailbox of rank ") + std::to_string(r) + "): " + cudaGetErrorString(e);
                return LCB_ERR_CUDA;
            }
            g_xch.dev.box[r] = (unsigned char *)p;
        }
        g_xch.ready = true;
        ctx->xch = g_xch.dev;
    }
    return LCB_OK;
#else
    ctx->error = "multi-GPU support is not compiled in (NCCL not found at build time)";
    return LCB_ERR_STATE;
#endif
}

extern "C" int lcb_enumerate_seeds(lcb_ctx *ctx, uint64_t *n_seeds)
{
    if (!ctx) return LCB_ERR_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (ctx->seeds_ready) {
        if (n_seeds) *n_seeds = ctx->n_seeds;
        return LCB_OK;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < ctx->seed_allocs.size(); i++) cached_free(ctx->seed_allocs[i], ctx->seed_alloc_bytes[i], ctx->device);
    ctx->seed_allocs.clear();
    ctx->seed_alloc_bytes.clear();
    std::vector<void *> keep;
    std::vector<size_t> keep_b;
    keep.swap(ctx->allocs); // everything allocated below is seed-scoped
    keep_b.swap(ctx->alloc_bytes);
    struct Restore {
        lcb_ctx *c;
        std::vector<void *> &k;
        std::vector<size_t> &kb;
        ~Restore()
        {
            c->seed_allocs.insert(c->seed_allocs.end(), c->allocs.begin(), c->allocs.end());
            c->seed_alloc_bytes.insert(c->seed_alloc_bytes.end(), c->alloc_bytes.begin(), c->alloc_bytes.end());
            c->allocs.swap(k);
            c->alloc_bytes.swap(kb);
        }
    } restore{ctx, keep, keep_b};
    ctx->st.kernel_launches = 0;
    ctx->st.traverse_launches = 0;
    CUDA_TRY(cudaEventRecord(ctx->ev_step0, ctx->stream));
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    const size_t slots = 2 * (size_t)ctx->ix.V;
    unsigned *d_cnt = nullptr, *d_off = nullptr, *d_tile = nullptr, *d_total = nullptr;
    int rc;
    if ((rc = dev_alloc(ctx, &d_cnt, slots + 1))) return rc;
    if ((rc = dev_alloc(ctx, &d_off, slots + 1))) return rc;
    if ((rc = dev_alloc(ctx, &d_tile, slots / kScanTile + 2))) return rc;
    if ((rc = dev_alloc(ctx, &d_total, 4))) return rc;
    const unsigned blocks = (unsigned)((slots + 255) / 256);
    SeedArrays none{};
    if (slots) k_seed_enum<false><<<blocks, 256, 0, ctx->stream>>>(ctx->ix, d_cnt, nullptr, none);
    ctx->st.kernel_launches += 1;
    unsigned total = 0;
    if (slots) {
        if ((rc = exclusive_scan(ctx, d_cnt, d_off, slots, d_total, d_tile))) return rc;
        CUDA_TRY(cudaMemcpyAsync(&total, d_total, sizeof total, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const unsigned S = total;
    ctx->n_seeds = S;
    SeedArrays raw{}, srt{};
    if ((rc = dev_alloc(ctx, &raw.vid, S))) return rc;
    if ((rc = dev_alloc(ctx, &raw.ch, S))) return rc;
    if ((rc = dev_alloc(ctx, &raw.count, S))) return rc;
    if ((rc = dev_alloc(ctx, &raw.rank, S))) return rc;
    if ((rc = dev_alloc(ctx, &raw.res_pos, S))) return rc;
    if ((rc = dev_alloc(ctx, &raw.res_chr, S))) return rc;
    if ((rc = dev_alloc(ctx, &srt.vid, S))) return rc;
    if ((rc = dev_alloc(ctx, &srt.ch, S))) return rc;
    if ((rc = dev_alloc(ctx, &srt.count, S))) return rc;
    if ((rc = dev_alloc(ctx, &srt.rank, S))) return rc;
    if ((rc = dev_alloc(ctx, &srt.res_pos, S))) return rc;
    if ((rc = dev_alloc(ctx, &srt.res_chr, S))) return rc;
    if (S) {
        k_seed_enum<true><<<blocks, 256, 0, ctx->stream>>>(ctx->ix, nullptr, d_off, raw);
        ctx->st.kernel_launches += 1;
        // Bundle::operator< (blocksfinder.h:195-208): count desc, rank asc, resolve=(pos, chr) asc -- a total order,
        // so a stable LSD radix sort from the least significant key reproduces std::sort's result exactly
        unsigned *perm = nullptr, *tmp = nullptr, *d_hist = nullptr, *d_flag = nullptr;
        const unsigned sblocks = (S + kSortTile - 1) / kSortTile;
        if ((rc = dev_alloc(ctx, &perm, S))) return rc;
        if ((rc = dev_alloc(ctx, &tmp, S))) return rc;
        if ((rc = dev_alloc(ctx, &d_hist, (size_t)256 * sblocks + (256 * sblocks) / kScanTile + 8))) return rc;
        if ((rc = dev_alloc(ctx, &d_flag, 4))) return rc;
        k_iota<<<(S + 255) / 256, 256, 0, ctx->stream>>>(perm, S);
        if ((rc = radix_sort_word<unsigned>(ctx, &perm, &tmp, raw.res_chr, 32, false, S, d_hist, d_flag))) return rc;
        if ((rc = radix_sort_word<unsigned>(ctx, &perm, &tmp, raw.res_pos, 32, false, S, d_hist, d_flag))) return rc;
        if ((rc = radix_sort_word<unsigned long long>(ctx, &perm, &tmp, raw.rank, 64, false, S, d_hist, d_flag))) return rc;
        if ((rc = radix_sort_word<unsigned>(ctx, &perm, &tmp, raw.count, 32, true, S, d_hist, d_flag))) return rc;
        const unsigned gb = (S + 255) / 256;
        k_gather<<<gb, 256, 0, ctx->stream>>>(perm, raw.vid, srt.vid, S);
        k_gather<<<gb, 256, 0, ctx->stream>>>(perm, raw.ch, srt.ch, S);
        k_gather<<<gb, 256, 0, ctx->stream>>>(perm, raw.count, srt.count, S);
        k_gather<<<gb, 256, 0, ctx->stream>>>(perm, raw.rank, srt.rank, S);
        k_gather<<<gb, 256, 0, ctx->stream>>>(perm, raw.res_pos, srt.res_pos, S);
        k_gather<<<gb, 256, 0, ctx->stream>>>(perm, raw.res_chr, srt.res_chr, S);
        ctx->st.kernel_launches += 7;
    }
    ctx->d_seed_vid = srt.vid;
    ctx->d_seed_ch = srt.ch;
    ctx->d_seed_count = srt.count;
    ctx->d_seed_rank = srt.rank;
    ctx->d_seed_res_pos = srt.res_pos;
    ctx->d_seed_res_chr = srt.res_chr;
    ctx->max_seed_count = 0;
    if (S) CUDA_TRY(cudaMemcpyAsync(&ctx->max_seed_count, srt.count, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream)); // sorted: count desc
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->st.ms_enumerate = ms;
    ctx->st.n_seeds = S;
    ctx->seeds_ready = true;
    if (n_seeds) *n_seeds = S;
    return LCB_OK;
}

extern "C" int lcb_get_seeds(lcb_ctx *ctx, int64_t *vid, uint8_t *ch, uint64_t *count, uint64_t *rank, uint64_t *res_pos,
                             uint64_t *res_chr)
{
    if (!ctx || !ctx->seeds_ready) return LCB_ERR_STATE;
    const size_t S = ctx->n_seeds;
    std::vector<int> hv(S);
    std::vector<unsigned> hu(S);
    std::vector<unsigned long long> hr(S);
    if (vid) {
        CUDA_TRY(cudaMemcpy(hv.data(), ctx->d_seed_vid, S * sizeof(int), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < S; i++) vid[i] = hv[i];
    }
    if (ch) CUDA_TRY(cudaMemcpy(ch, ctx->d_seed_ch, S, cudaMemcpyDeviceToHost));
    if (count) {
        CUDA_TRY(cudaMemcpy(hu.data(), ctx->d_seed_count, S * sizeof(unsigned), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < S; i++) count[i] = hu[i];
    }
    if (rank) {
        CUDA_TRY(cudaMemcpy(hr.data(), ctx->d_seed_rank, S * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < S; i++) rank[i] = hr[i];
    }
    if (res_pos) {
        CUDA_TRY(cudaMemcpy(hu.data(), ctx->d_seed_res_pos, S * sizeof(unsigned), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < S; i++) res_pos[i] = hu[i];
    }
    if (res_chr) {
        CUDA_TRY(cudaMemcpy(hu.data(), ctx->d_seed_res_chr, S * sizeof(unsigned), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < S; i++) res_chr[i] = hu[i];
    }
    return LCB_OK;
}

namespace {

constexpr int kSideCtas = 32; // CTAs of the general kernel that run beside the common-case kernel (one per SM on 32 SMs)

// One round's traversal.  `host_reset`: the host-driven loop resets the cursors here (the device-driven one does it in
// k_round_begin / k_rebase).  `beside`: the general kernel runs BESIDE the common-case kernel on the side stream (few CTAs
// at 96 registers, polling the hand-over list) instead of behind it: a heavy seed's evaluation lasts about as long as a
// whole round of light ones, and behind the common-case kernel it would add that to every round that has one.
int launch_traverse(lcb_ctx *ctx, const uint32_t *E, int slot, const unsigned *list, const unsigned *n_ptr, bool host_reset, bool beside)
{
    Params pr{ctx->prm.k, ctx->prm.max_branch, ctx->prm.min_block, ctx->prm.max_flank, ctx->prm.looking_depth};
    const unsigned phase = (unsigned)ctx->prm.phase_size;
    if (host_reset) {
        CUDA_TRY(cudaMemsetAsync(&ctx->d_ctl->head, 0, 4 * sizeof(unsigned), ctx->stream)); // head, max_ns, head_heavy, head_drain
        CUDA_TRY(cudaMemsetAsync(&ctx->d_ctl->lean_started, 0, 2 * sizeof(unsigned), ctx->stream)); // lean_started, lean_exited
    }
    if (ctx->grid_lean > 0 && slot < 0) {
        const int side = beside ? std::min(kSideCtas, ctx->grid_lean / 2) : 0;
        const unsigned lean_grid = (unsigned)(ctx->grid_lean - side);
        if (side) { // fork: the side stream's kernel is queued first so that its CTAs are placed first
            CUDA_TRY(cudaEventRecord(ctx->ev_fork, ctx->stream));
            CUDA_TRY(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
            k_traverse<false, 5><<<side, kThreads, 0, ctx->stream2>>>(ctx->ix, pr, E, ctx->d_seed_vid, ctx->d_seed_ch, phase, slot,
                                                                       ctx->win.list_heavy, &ctx->d_ctl->n_heavy, &ctx->d_ctl->head_heavy,
                                                                       ctx->win, ctx->d_ctl, ctx->d_arena, ctx->arena_stride,
                                                                       ctx->d_arena + ctx->d_big, 0, 2, lean_grid, ctx->xch);
            CUDA_TRY(cudaEventRecord(ctx->ev_join, ctx->stream2));
        }
        k_traverse_lean<<<lean_grid, kThreads, 0, ctx->stream>>>(ctx->ix, pr, E, ctx->d_seed_vid, ctx->d_seed_ch, phase, list, n_ptr,
                                                                 ctx->win, ctx->d_ctl, ctx->d_lean_rs, ctx->d_lean_shadow, ctx->d_lean_hash2, ctx->d_lean_hslot2, ctx->xch);
        if (side) CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
        // behind it: the list is complete (after a kernel beside: whatever that one left, normally nothing)
        k_traverse<false, LCB_TRAVERSE_CTAS_PER_SM><<<side ? std::max(ctx->sms, 1) : ctx->grid_traverse, kThreads, 0, ctx->stream>>>(
                ctx->ix, pr, E, ctx->d_seed_vid, ctx->d_seed_ch, phase, slot, ctx->win.list_heavy, &ctx->d_ctl->n_heavy,
                &ctx->d_ctl->head_drain, ctx->win, ctx->d_ctl, ctx->d_arena, ctx->arena_stride, ctx->d_arena + ctx->d_big, 0, 1, 0u, ctx->xch);
        ctx->st.kernel_launches += side ? 3 : 2;
        ctx->st.traverse_launches++;
        return LCB_OK;
    }
    if (ctx->prm.collect_counters)
        k_traverse<true, LCB_TRAVERSE_CTAS_PER_SM><<<ctx->grid_traverse, kThreads, 0, ctx->stream>>>(
            ctx->ix, pr, E, ctx->d_seed_vid, ctx->d_seed_ch, phase, slot, list, n_ptr, &ctx->d_ctl->head, ctx->win, ctx->d_ctl, ctx->d_arena,
            ctx->arena_stride, ctx->d_arena + ctx->d_big, ctx->prm.collect_counters, 0, 0u, ctx->xch);
    else
        k_traverse<false, LCB_TRAVERSE_CTAS_PER_SM><<<ctx->grid_traverse, kThreads, 0, ctx->stream>>>(
            ctx->ix, pr, E, ctx->d_seed_vid, ctx->d_seed_ch, phase, slot, list, n_ptr, &ctx->d_ctl->head, ctx->win, ctx->d_ctl, ctx->d_arena,
            ctx->arena_stride, ctx->d_arena + ctx->d_big, 0, 0, 0u, ctx->xch);
    ctx->st.kernel_launches++;
    ctx->st.traverse_launches++;
    return LCB_OK;
}

int fetch_control(lcb_ctx *ctx)
{
    CUDA_TRY(cudaMemcpyAsync(ctx->h_ctl, ctx->d_ctl, sizeof(Control), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaGetLastError());
    return LCB_OK;
}

} // namespace


namespace {

constexpr int kRoundsInFlight = 4; // rounds the host keeps queued ahead of the device
constexpr int kEvRing = 16;        // event pairs (> kRoundsInFlight)

// the non-traversal part of round r (epochs, validation, schedule, commit) followed by the admission of round r + 1
int enqueue_tail(lcb_ctx *ctx, int parity, const SchedParams &sp)
{
    const size_t N = (size_t)ctx->ix.N;
    uint32_t *Ecur = ctx->d_E[parity], *Enew = ctx->d_E[parity ^ 1];
    const unsigned vgrid = (unsigned)ctx->sms * 8;
    const unsigned egrid = (unsigned)std::min<size_t>((N + 1023) / 1024 + 1, (size_t)ctx->sms * 16);
    const unsigned sgrid = (unsigned)ctx->sms * 4; // grid-stride kernels over admitted / committed seeds
    cudaStream_t st = ctx->stream;
    const unsigned R = (unsigned)ctx->n_ranks, me = (unsigned)ctx->rank;
    if (R > 1) { // the peers' results of this round: tags out, tags in, results filed
        k_xflag<<<1, 32, 0, st>>>(ctx->xch, ctx->d_ctl);
        k_xwait<<<1, 32, 0, st>>>(ctx->xch, ctx->d_ctl);
        k_xapply<<<vgrid, 256, 0, st>>>(ctx->xch, ctx->win, ctx->d_ctl);
    }
    k_rebase<<<egrid, 256, 0, st>>>(Enew, Ecur, N, 0u, ctx->d_ctl);
    k_claim<<<vgrid, 256, 0, st>>>(Enew, 0u, 0u, ctx->win, ctx->d_ctl);
    if (ctx->d_diff) k_diff<<<egrid, 256, 0, st>>>((const uint4 *)Ecur, (const uint4 *)Enew, epoch_len(N) / 4, ctx->d_diff, ctx->d_ctl);
    k_validate<<<vgrid, 256, 0, st>>>(Ecur, Enew, 0u, 0u, (unsigned)ctx->prm.phase_size, ctx->win, ctx->d_ctl, 1, ctx->d_diff, R, me);
    if (R > 1) k_csend<<<1, 32, 0, st>>>(ctx->xch, ctx->d_ctl, ctx->win.inst_cap, ctx->win.rs_cap);
    k_round_end<<<1, 1, 0, st>>>(ctx->d_ctl, sp, ctx->h_mirror, ctx->xch);
    k_final_counts<<<sgrid, 256, 0, st>>>(0u, 0u, ctx->win, ctx->d_counts, ctx->d_ctl);
    k_emit_scan<<<1, 1024, 0, st>>>(0u, 0u, ctx->win, ctx->d_ctl, ctx->d_counts, 1);
    k_emit_write<<<sgrid, 256, 0, st>>>(ctx->ix, ctx->prm.k, 0u, 0u, ctx->win, ctx->d_out, ctx->d_ctl);
    k_round_begin<<<1, 1, 0, st>>>(ctx->d_ctl, R, me);
    k_admit<<<sgrid, 256, 0, st>>>(0u, 0u, R, me, 0u, ctx->win, ctx->d_ctl, 1);
    CUDA_TRY(cudaGetLastError());
    return LCB_OK;
}
constexpr unsigned kTailKernels = 10;

// Single-GPU round loop driven from the device: every decision of a round (admission, commit frontier, buffer swap) is taken
// by k_round_begin / k_round_end from the control block, so the host never waits for a round: it keeps kRoundsInFlight rounds
// queued (two traversal launches + one graph launch each) and watches a few words the device writes to pinned memory.
int find_blocks_device_loop(lcb_ctx *ctx, float &trav_ms)
{
    const unsigned S = (unsigned)ctx->n_seeds;
    const unsigned phase = (unsigned)ctx->prm.phase_size;
    const size_t N = (size_t)ctx->ix.N;
    const bool trace_rounds = getenv("LCB_TRACE_ROUNDS") != nullptr;
    const bool use_graph = !getenv("LCB_NO_GRAPH");
    SchedParams sp;
    sp.grow_below = getenv("LCB_GROW_BELOW") ? (float)atof(getenv("LCB_GROW_BELOW")) : 1.2f;
    // Rounds shorter than this always grow: a round costs ~0.25 ms outside the traversal (epochs, validation, commit) however few
    // seeds it has.  Sweep on the 4 x 100 Mbp input: 0.6 / 1.0 / 1.5 / 2.5 ms -> 150 / 133 / 123 / 117 ms per pass
    // (profiles/sweep_round_length_r2.log); configs[1] does not care (40 ms throughout).
    sp.min_round_ms = getenv("LCB_MIN_ROUND_MS") ? (float)atof(getenv("LCB_MIN_ROUND_MS")) : 2.5f;
    sp.shrink_above = getenv("LCB_SHRINK_ABOVE") ? (float)atof(getenv("LCB_SHRINK_ABOVE")) : 2.0f;
    sp.heavy_ms = getenv("LCB_HEAVY_MS") ? (float)atof(getenv("LCB_HEAVY_MS")) : 20.0f;
    sp.phase = phase;
    sp.window_max = (unsigned)ctx->prm.window_max;
    sp.warps = (unsigned)(ctx->grid_traverse * kWarpsPerBlock);
    sp.inst_cap = ctx->win.inst_cap, sp.rs_cap = ctx->win.rs_cap;
    if (!ctx->h_mirror) CUDA_TRY(cached_alloc((void **)&ctx->h_mirror, sizeof(Mirror), -1, nullptr));
    if (ctx->ev_ring.empty()) {
        ctx->ev_ring.resize(2 * kEvRing);
        for (auto &e : ctx->ev_ring) CUDA_TRY(cudaEventCreate(&e));
    }
    Mirror *mir = ctx->h_mirror;
    mir->rounds_done = 0, mir->done = 0, mir->halt = 0, mir->c0 = 0, mir->n_heavy = 0;
    mir->heavy_last = 0xFFFFFFFFu; // nothing known yet: the general kernel behind the common-case kernel
    const bool allow_beside = !getenv("LCB_HEAVY_BEHIND"); // A/B switch: the general kernel always behind the common-case kernel
    // ---- initial control block
    Control &h = *ctx->h_ctl;
    memset(&h, 0, sizeof(Control));
    h.first_dirty = 0xFFFFFFFFu;
    h.n_seeds = S;
    h.x_tag_base = (++g_xch.run_seq & 0x3FFu) << 22; // (every rank runs lcb_find_blocks the same number of times)
    h.cap = (unsigned)ctx->prm.window_max;
    h.delta = (unsigned)ctx->prm.window_init;
    if (ctx->max_seed_count >= 32) // see lcb_find_blocks: repeat-rich inputs start with one wave of warps
        h.delta = std::min(h.delta, std::max(phase, sp.warps / phase * phase));
    if (S == 0) h.done = 1;
    CUDA_TRY(cudaMemcpyAsync(ctx->d_ctl, &h, sizeof(Control), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    // ---- the tail of a round as a graph (one per parity); captured once per context
    if (use_graph && !ctx->tail_graph[0]) {
        for (int par = 0; par < 2; par++) {
            cudaGraph_t g = nullptr;
            CUDA_TRY(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            const int rc = enqueue_tail(ctx, par, sp);
            cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
            if (rc) return rc;
            CUDA_TRY(e);
            CUDA_TRY(cudaGraphInstantiate(&ctx->tail_graph[par], g, 0));
            cudaGraphDestroy(g);
        }
    }
    const unsigned sgrid = (unsigned)ctx->sms * 4;
    k_round_begin<<<1, 1, 0, ctx->stream>>>(ctx->d_ctl, (unsigned)ctx->n_ranks, (unsigned)ctx->rank);
    k_admit<<<sgrid, 256, 0, ctx->stream>>>(0u, 0u, (unsigned)ctx->n_ranks, (unsigned)ctx->rank, 0u, ctx->win, ctx->d_ctl, 1);
    ctx->st.kernel_launches += 2;
    unsigned launched = 0, timed = 0;
    int parity = 0, rc;
    auto collect_times = [&](unsigned upto) { // rounds < upto are complete: read their event pairs
        for (; timed < upto; timed++) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ctx->ev_ring[2 * (timed % kEvRing)], ctx->ev_ring[2 * (timed % kEvRing) + 1]) == cudaSuccess) trav_ms += ms;
        }
    };
    while (true) {
        // keep the queue full
        while (!mir->done && !mir->halt && launched - mir->rounds_done < (unsigned)(trace_rounds ? 1 : kRoundsInFlight)) {
            collect_times(std::min((unsigned)mir->rounds_done, launched));
            CUDA_TRY(cudaEventRecord(ctx->ev_ring[2 * (launched % kEvRing)], ctx->stream));
            // beside while the known-heavy seeds fit the side CTAs' warps; otherwise the general kernel gets the whole GPU behind
            const bool beside = allow_beside && std::max((unsigned)mir->n_heavy, (unsigned)mir->heavy_last) <= (unsigned)(kSideCtas * kWarpsPerBlock) / 2;
            if ((rc = launch_traverse(ctx, ctx->d_E[parity], -1, ctx->win.list0, &ctx->d_ctl->n0, false, beside))) return rc;
            CUDA_TRY(cudaEventRecord(ctx->ev_ring[2 * (launched % kEvRing) + 1], ctx->stream));
            if (use_graph) CUDA_TRY(cudaGraphLaunch(ctx->tail_graph[parity], ctx->stream));
            else if ((rc = enqueue_tail(ctx, parity, sp))) return rc;
            ctx->st.kernel_launches += kTailKernels;
            launched++;
            parity ^= 1;
            if (trace_rounds) {
                if ((rc = fetch_control(ctx))) return rc;
                const Control &t = *ctx->h_ctl;
                fprintf(stderr, "[round] %u active [%u,%u) next admit %u traverse=%.3f ms longest=%.3f ms n0=%u dirty=%u first=%u delta=%u heavy=%u pools %.1f%% %.1f%%\n",
                        launched, t.c0, t.c1, t.admit, (t.t_last > t.t_first ? (double)(t.t_last - t.t_first) * 1e-6 : 0.0), t.max_ns * 1e-6, t.n0, t.dirty,
                        t.first_dirty, t.delta, t.n_heavy, 100.0 * t.inst_used / ctx->win.inst_cap, 100.0 * t.rs_used / ctx->win.rs_cap);
            }
        }
        if (mir->done || mir->halt) {
            CUDA_TRY(cudaStreamSynchronize(ctx->stream)); // the rounds queued behind the last one are no-ops
            collect_times(launched);
            if ((rc = fetch_control(ctx))) return rc;
            Control &c = *ctx->h_ctl;
            if (c.err) {
                ctx->arena_dirty = true;
                static const char *const what[] = {"a per-seed buffer", "path vertices (cap 524288)", "read-set intervals (cap 1048576)",
                                                   "path instances (cap 32768)", "look-ahead vote table (cap 262144 vertices)"};
                ctx->error = std::string("a per-seed device buffer overflowed its hard cap: ") + what[std::min(c.err >> 8, 4u)];
                return (int)(c.err & 0xFF);
            }
            if (c.done) break;
            // halt without error: a result pool ran full.  Forget the active set (committed seeds are final), halve it, go on.
            if (c.c1 - c.c0 <= phase) {
                ctx->error = "result pools overflowed at the minimum active set (one phase)";
                return LCB_ERR_CAPACITY;
            }
            c.cap = std::max(phase, (c.c1 - c.c0) / 2 / phase * phase);
            c.delta = std::min(c.delta, c.cap);
            c.c1 = c.c0;
            c.n0 = c.n1 = c.dirty = 0;
            c.n_heavy = c.head_heavy = 0;
            CUDA_TRY(cudaMemsetAsync(ctx->win.list_heavy, 0, sizeof(unsigned) * ctx->wmax, ctx->stream));
            c.pool_overflow = 0, c.halt = 0;
            c.first_dirty = 0xFFFFFFFFu;
            // the round that overflowed did not swap the buffers: its `current` epochs are those of c.parity
            parity = (int)c.parity;
            CUDA_TRY(cudaMemcpyAsync(ctx->d_ctl, &c, sizeof(Control), cudaMemcpyHostToDevice, ctx->stream));
            k_rebase<<<(unsigned)std::min<size_t>((N + 1023) / 1024 + 1, (size_t)ctx->sms * 16), 256, 0, ctx->stream>>>(ctx->d_E[parity], ctx->d_E[parity], N, c.c0, nullptr); // drop the abandoned seeds' claims
            k_round_begin<<<1, 1, 0, ctx->stream>>>(ctx->d_ctl, (unsigned)ctx->n_ranks, (unsigned)ctx->rank);
            k_admit<<<sgrid, 256, 0, ctx->stream>>>(0u, 0u, (unsigned)ctx->n_ranks, (unsigned)ctx->rank, 0u, ctx->win, ctx->d_ctl, 1);
            ctx->st.kernel_launches += 3;
            ctx->st.pool_restarts++;
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            mir->halt = 0;
            mir->rounds_done = launched = timed = 0; // round counters restart with the control block's (rounds keeps counting on the device)
            {
                // keep the device's round counter and the host's in step: the mirror reports ctl->rounds
                launched = timed = c.rounds;
                mir->rounds_done = c.rounds;
            }
            continue;
        }
        // queue full: wait for the device to finish a round (pinned-memory poll, no stream synchronisation)
        const unsigned seen = mir->rounds_done;
        for (unsigned spins = 1; mir->rounds_done == seen && !mir->done && !mir->halt; spins++) {
            std::this_thread::yield();
            if ((spins & 0x3FFu) == 0) { // now and then: is the device still alive?  (a faulting kernel never reports a round)
                const cudaError_t q = cudaStreamQuery(ctx->stream);
                if (q != cudaSuccess && q != cudaErrorNotReady) {
                    ctx->arena_dirty = true;
                    ctx->error = std::string("the traversal rounds stopped on the device: ") + cudaGetErrorString(q);
                    return LCB_ERR_CUDA;
                }
                if (q == cudaSuccess && mir->rounds_done == seen && !mir->done && !mir->halt) {
                    // everything queued has run, yet no round was reported: cannot happen (k_round_end reports every round)
                    ctx->arena_dirty = true;
                    ctx->error = "the traversal rounds ended without a report from the device";
                    return LCB_ERR_CUDA;
                }
            }
        }
    }
    ctx->st.rounds = ctx->h_ctl->rounds;
    ctx->st.windows = ctx->h_ctl->windows;
    return LCB_OK;
}

} // namespace

extern "C" int lcb_find_blocks(lcb_ctx *ctx, lcb_block_instance **out, uint64_t *n_out, lcb_stats *stats)
{
    if (!ctx || !out || !n_out) return LCB_ERR_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc;
    ctx->step_timed = !ctx->seeds_ready;
    if (!ctx->seeds_ready && (rc = lcb_enumerate_seeds(ctx, nullptr))) return rc;
    auto t_begin = std::chrono::steady_clock::now();
    if (ctx->n_seeds >= 0x7FFFFFF0ull) {
        ctx->error = "too many seeds for 31-bit work items";
        return LCB_ERR_ARG;
    }
    const unsigned S = (unsigned)ctx->n_seeds;
    const unsigned phase = (unsigned)ctx->prm.phase_size;
    const size_t N = (size_t)ctx->ix.N;
    uint32_t *Ecur = ctx->d_E[0], *Enew = ctx->d_E[1];
    CUDA_TRY(cudaMemsetAsync(Ecur, 0xFF, epoch_len(N) * sizeof(uint32_t), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(Enew, 0xFF, epoch_len(N) * sizeof(uint32_t), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(ctx->d_ctl, 0, sizeof(Control), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(ctx->win.list_heavy, 0, sizeof(unsigned) * ctx->wmax, ctx->stream)); // 0 = entry not written
    memset(ctx->h_ctl, 0, sizeof(Control));
    const unsigned vgrid = (unsigned)ctx->sms * 8;
    const unsigned egrid = (unsigned)std::min<size_t>((N + 1023) / 1024 + 1, (size_t)ctx->sms * 16);
    float trav_ms = 0;
    ctx->st.windows = ctx->st.rounds = ctx->st.pool_restarts = 0;
    const bool trace_rounds = getenv("LCB_TRACE_ROUNDS") != nullptr;
    // admission thresholds (developer knobs): round time relative to its longest single evaluation
    const double grow_below = getenv("LCB_GROW_BELOW") ? atof(getenv("LCB_GROW_BELOW")) : 1.2;
    const double min_round_ms = getenv("LCB_MIN_ROUND_MS") ? atof(getenv("LCB_MIN_ROUND_MS")) : 2.5; // rounds shorter than this always grow
    const double shrink_above = getenv("LCB_SHRINK_ABOVE") ? atof(getenv("LCB_SHRINK_ABOVE")) : 2.0;
    const unsigned R = (unsigned)ctx->n_ranks, me = (unsigned)ctx->rank;
    // rolling active set [c0, c1): c0 = commit frontier, c1 = admission frontier
    unsigned c0 = 0, c1 = 0;
    unsigned n0 = 0;                                     // this rank's work list 0 as left by the last validation
    unsigned delta = (unsigned)ctx->prm.window_init;     // seeds admitted per round (adapted)
    unsigned cap = (unsigned)ctx->prm.window_max;        // bound on the active set (halved by pool overflows)
    // Bundles are sorted by abundance.  When the first ones have dozens of occurrences (repeat-rich input, small k) their
    // speculative evaluations against the still empty epochs are the most expensive of the whole run and nearly all of
    // them will be thrown away: admit no more than one wave of resident warps to begin with (the rule below takes over).
    if (ctx->max_seed_count >= 32)
        delta = std::min(delta, std::max(phase, (unsigned)(ctx->grid_traverse * kWarpsPerBlock) / phase * phase));
    bool drain = false;                                  // result pools half full: stop admitting until the set is empty
    bool hold = false;                                   // heavy re-evaluations keep every warp busy: admit nothing this round
    const double heavy_ms = getenv("LCB_HEAVY_MS") ? atof(getenv("LCB_HEAVY_MS")) : 20.0;
    // The round loop runs from the device (find_blocks_device_loop); with several ranks the results travel through the peers'
    // mailboxes inside it.  LCB_HOST_LOOP=1 (one rank only, the A/B switch): the loop below, where the host takes every
    // round's decisions after a stream synchronisation.
    if (R > 1 && !ctx->xch.R) {
        ctx->error = "lcb_comm_init must run on every rank before lcb_find_blocks";
        return LCB_ERR_STATE;
    }
    const bool device_loop = R > 1 || !getenv("LCB_HOST_LOOP");
    if (device_loop) {
        if ((rc = find_blocks_device_loop(ctx, trav_ms))) return rc;
        c0 = S;
    }
    while (c0 < S) {
        // ---- admission
        if (c0 == c1) { // nothing active: no pool entry is referenced any more
            CUDA_TRY(cudaMemsetAsync(&ctx->d_ctl->inst_used, 0, 2 * sizeof(unsigned long long), ctx->stream));
            drain = false;
            n0 = 0;
            ctx->st.windows++;
        }
        unsigned admit = 0;
        if (!drain && !(hold && c1 > c0) && c1 < S && c1 - c0 < cap) admit = std::min(std::min(delta, S - c1), cap - (c1 - c0));
        if (admit) {
            k_admit<<<(admit + 255) / 256, 256, 0, ctx->stream>>>(c1, admit, R, me, n0, ctx->win, ctx->d_ctl, 0);
            ctx->st.kernel_launches++;
            c1 += admit;
        }
        ctx->st.rounds++;
        // A. speculative evaluations (new + invalidated seeds), each followed at once by its commit-time re-run when the
        //    fresh result conflicts; plus the commit-time re-runs the last validation queued (tagged items)
        CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
        if ((rc = launch_traverse(ctx, Ecur, -1, ctx->win.list0, &ctx->d_ctl->n0, true, false))) return rc;
        CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
        // D. new epochs: committed claims + the active seeds' current final results
        k_rebase<<<egrid, 256, 0, ctx->stream>>>(Enew, Ecur, N, c0, nullptr);
        k_claim<<<vgrid, 256, 0, ctx->stream>>>(Enew, c0, c1, ctx->win, nullptr);
        // E. validation + next work lists
        CUDA_TRY(cudaMemsetAsync(&ctx->d_ctl->n0, 0, 3 * sizeof(unsigned), ctx->stream));        // n0, n1, dirty
        CUDA_TRY(cudaMemsetAsync(&ctx->d_ctl->n_heavy, 0, sizeof(unsigned), ctx->stream));       // (the validation queues the known-heavy seeds)
        CUDA_TRY(cudaMemsetAsync(&ctx->d_ctl->first_dirty, 0xFF, sizeof(unsigned), ctx->stream));
        if (ctx->d_diff) k_diff<<<egrid, 256, 0, ctx->stream>>>((const uint4 *)Ecur, (const uint4 *)Enew, epoch_len(N) / 4, ctx->d_diff, nullptr);
        k_validate<<<vgrid, 256, 0, ctx->stream>>>(Ecur, Enew, c0, c1, phase, ctx->win, ctx->d_ctl, 0, ctx->d_diff, R, me);
        ctx->st.kernel_launches += 4;
        if ((rc = fetch_control(ctx))) return rc;
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        trav_ms += ms;
        Control &h = *ctx->h_ctl;
        if (h.inst_used * 2 > ctx->win.inst_cap || h.rs_used * 2 > ctx->win.rs_cap) drain = true;
        // Admission rate: a launch lasts max(longest evaluation, work / resident warps).  While it is latency-bound
        // more fresh seeds are free; when the fresh work dominates, far-ahead speculation only adds re-evaluations.
        unsigned next_delta = delta;
        {
            const double longest_ms = h.max_ns * 1e-6;
            if (ms < grow_below * longest_ms || ms < min_round_ms) next_delta = (unsigned)std::min<unsigned long long>((unsigned long long)ctx->prm.window_max, 2ull * delta);
            else if (ms > shrink_above * longest_ms && ms > 2 * min_round_ms) next_delta = std::max(phase, delta / 2 / phase * phase);
            // Heavy evaluations (repeat-rich input): a warp that re-evaluates an invalidated seed is busy for about as long as
            // the whole round, so fresh seeds are free only while warps are left over
            if (longest_ms > heavy_ms) {
                const unsigned warps = (unsigned)(ctx->grid_traverse * kWarpsPerBlock);
                if (h.n0 + phase / R > warps) next_delta |= 0x80000000u; // hold
                else next_delta = std::min(next_delta, std::max(phase, (warps - h.n0) * R / phase * phase));
            }
        }
        unsigned first_dirty = h.first_dirty;
        hold = (next_delta >> 31) != 0;
        next_delta &= 0x7FFFFFFFu;
        if (trace_rounds)
            fprintf(stderr, "[round] %llu active [%u,%u) admitted %u traverse=%.3f ms longest=%.3f ms next: n0=%u dirty=%u first=%u delta=%u pools %.1f%% %.1f%%\n",
                    (unsigned long long)ctx->st.rounds, c0, c1, admit, ms, h.max_ns * 1e-6, h.n0, h.dirty, first_dirty, next_delta,
                    100.0 * h.inst_used / ctx->win.inst_cap, 100.0 * h.rs_used / ctx->win.rs_cap);
        if (h.err) {
            ctx->arena_dirty = true;
            // (the per-warp arena's smaller caps are not errors: such seeds are re-run in a big arena slot)
            static const char *const what[] = {"a per-seed buffer", "path vertices (cap 524288)", "read-set intervals (cap 1048576)",
                                               "path instances (cap 32768)", "look-ahead vote table (cap 262144 vertices)"};
            const unsigned w = std::min(h.err >> 8, 4u);
            ctx->error = std::string("a per-seed device buffer overflowed its hard cap: ") + what[w];
            return (int)(h.err & 0xFF);
        }
        if (h.pool_overflow) {
            // a result pool ran full: forget the active set (committed seeds are final), halve it and start again
            if (c1 - c0 <= phase) {
                ctx->error = "result pools overflowed at the minimum active set (one phase)";
                return LCB_ERR_CAPACITY;
            }
            cap = std::max(phase, (c1 - c0) / 2 / phase * phase);
            delta = std::min(delta, cap);
            c1 = c0;
            CUDA_TRY(cudaMemsetAsync(&ctx->d_ctl->n0, 0, 3 * sizeof(unsigned), ctx->stream));
            CUDA_TRY(cudaMemsetAsync(&ctx->d_ctl->n_heavy, 0, sizeof(unsigned), ctx->stream));
            CUDA_TRY(cudaMemsetAsync(ctx->win.list_heavy, 0, sizeof(unsigned) * ctx->wmax, ctx->stream));
            CUDA_TRY(cudaMemsetAsync(&ctx->d_ctl->pool_overflow, 0, sizeof(unsigned), ctx->stream));
            k_rebase<<<egrid, 256, 0, ctx->stream>>>(Ecur, Ecur, N, c0, nullptr); // drop the abandoned seeds' claims
            ctx->st.kernel_launches++;
            ctx->st.pool_restarts++;
            continue;
        }
        // F. commit the clean prefix: Finalize in seed order (block ids, blocksInstance_ records)
        const unsigned fd = std::min(first_dirty, c1);
        if (fd > c0) {
            const unsigned n = fd - c0;
            k_final_counts<<<(n + 255) / 256, 256, 0, ctx->stream>>>(c0, n, ctx->win, ctx->d_counts, nullptr);
            k_emit_scan<<<1, 1024, 0, ctx->stream>>>(c0, n, ctx->win, ctx->d_ctl, ctx->d_counts, 0);
            k_emit_write<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->ix, ctx->prm.k, c0, n, ctx->win, ctx->d_out, nullptr);
            ctx->st.kernel_launches += 3;
            c0 = fd;
        }
        std::swap(Ecur, Enew); // the new epochs become current
        n0 = h.n0;
        delta = std::min(next_delta, cap);
    }
    if ((rc = fetch_control(ctx))) return rc; // totals of the last commit
    const unsigned out_done = ctx->h_ctl->out_done, blocks_done = ctx->h_ctl->blocks_done;
    // ---- results ----
    auto t_d2h = std::chrono::steady_clock::now();
    // the result goes to page-locked memory from the block cache (no first-touch page faults, a true DMA copy);
    // lcb_free_blocks hands it back
    lcb_block_instance *host = nullptr;
    const size_t host_bytes = sizeof(lcb_block_instance) * std::max<size_t>(out_done, 1);
    if (cached_alloc((void **)&host, host_bytes, -1, nullptr) != cudaSuccess || !host) {
        cudaGetLastError();
        ctx->error = "out of host memory";
        return LCB_ERR_ARG;
    }
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        g_results.emplace_back((void *)host, host_bytes);
    }
    CUDA_TRY(cudaEventRecord(ctx->ev_step1, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(host, ctx->d_out, sizeof(lcb_block_instance) * out_done, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->st.ms_step_device = 0;
    if (ctx->step_timed) {
        float sm = 0;
        cudaEventElapsedTime(&sm, ctx->ev_step0, ctx->ev_step1);
        ctx->st.ms_step_device = sm;
    }
    auto t_end = std::chrono::steady_clock::now();
    ctx->st.ms_d2h = std::chrono::duration<double, std::milli>(t_end - t_d2h).count();
    ctx->st.d2h_bytes = sizeof(lcb_block_instance) * (uint64_t)out_done;
    ctx->st.ms_find = std::chrono::duration<double, std::milli>(t_end - t_begin).count();
    ctx->st.ms_traverse_kernels = trav_ms;
    ctx->st.n_block_instances = out_done;
    ctx->st.n_blocks = blocks_done;
    ctx->st.traversals_first = ctx->h_ctl->runs0;
    ctx->st.big_arena_runs = ctx->h_ctl->big_runs;
    ctx->st.lean_runs = ctx->h_ctl->lean_runs;
    ctx->st.lean_bails = ctx->h_ctl->lean_bails;
    for (int q = 0; q < 6; q++) ctx->st.ms_tail[q] = (double)ctx->h_ctl->tail_ns[q] * 1e-6;
    for (int w = 0; w < 8; w++) ctx->st.lean_bail_why[w] = ctx->h_ctl->lean_why[w];
    ctx->st.traversals_rerun = ctx->h_ctl->runs1;
    ctx->st.t_walk = ctx->h_ctl->ct_walk;
    ctx->st.t_occ = ctx->h_ctl->ct_occ;
    ctx->st.t_scan = ctx->h_ctl->ct_scan;
    ctx->st.t_score = ctx->h_ctl->ct_score;
    *out = host;
    *n_out = out_done;
    if (stats) *stats = ctx->st;
    return LCB_OK;
}

extern "C" void lcb_free_blocks(lcb_block_instance *p)
{
    if (!p) return;
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (size_t i = 0; i < g_results.size(); i++)
            if (g_results[i].first == (void *)p) {
                bytes = g_results[i].second;
                g_results.erase(g_results.begin() + (long)i);
                break;
            }
    }
    if (bytes) cached_free(p, bytes, -1); // page-locked result buffer: back to the block cache
    else free(p);
}

extern "C" void lcb_trim_cache(void)
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (auto &b : g_cache) {
        if (b.device < 0) cudaFreeHost(b.p);
        else {
            cudaSetDevice(b.device);
            cudaFree(b.p);
        }
    }
    g_cache.clear();
#ifdef LCB_WITH_NCCL
    if (g_comm) ncclCommDestroy(g_comm);
    g_comm = nullptr;
#endif
    if (g_xch.ready) {
        for (int r = 0; r < g_xch.dev.R; r++)
            if (r != g_xch.dev.me && g_xch.dev.box[r]) cudaIpcCloseMemHandle(g_xch.dev.box[r]);
        if (g_xch.local) cudaFree(g_xch.local);
        g_xch = XchHost{};
    }
}

extern "C" int lcb_reset_seeds(lcb_ctx *ctx)
{
    if (!ctx) return LCB_ERR_ARG;
    ctx->seeds_ready = false;
    return LCB_OK;
}

extern "C" int lcb_get_stats(lcb_ctx *ctx, lcb_stats *stats)
{
    if (!ctx || !stats) return LCB_ERR_ARG;
    *stats = ctx->st;
    return LCB_OK;
}
