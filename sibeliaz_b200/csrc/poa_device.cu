// poa_device.cu -- the alignment stage on the GPU (include/sibeliaz_align.h): kernel k_poa (one warp per block, the
// block's partial order graph and score matrix in a per-warp arena in HBM, see poa_core.cuh), the driver that deals the
// blocks to arenas of three sizes, and the chunk-file front end that writes alignment.maf.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "host_common.h"
#include "lcb_internal.h"
#include "poa_core.cuh"
#include "sibeliaz_align.h"

namespace {

constexpr int kPoaWarps = 4; // warps per CTA
constexpr uint8_t kPending = 0, kDone = 4, kPoolFull = 3; // block status (1, 2: poa::Work::err)

struct Queue {
    unsigned head;
    unsigned long long rows_used;
    unsigned long long cells;
};

// Persistent warps pull blocks from `list`; a warp's arena is arena_base + warp * arena_stride and is big enough for every
// block of the list at `level` (the host sorted them that way).
__global__ void __launch_bounds__(kPoaWarps * 32) k_poa(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ copy_off,
                                                       const uint32_t *__restrict__ block_off, const uint32_t *__restrict__ list,
                                                       unsigned n_list, int level, uint8_t *arena_base, unsigned long long arena_stride,
                                                       poa::Params pr, Queue *q, uint8_t *rows, unsigned long long rows_cap,
                                                       unsigned long long *row_off, uint32_t *block_cols, uint8_t *status)
{
    __shared__ poa::Work ws[kPoaWarps];
    __shared__ unsigned long long offs[kPoaWarps];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t warp = (size_t)blockIdx.x * kPoaWarps + wib;
    poa::Work &w = ws[wib];
    unsigned long long cells = 0;
    while (true) {
        unsigned idx = 0;
        if (lane == 0) idx = atomicAdd(&q->head, 1u);
        idx = __shfl_sync(0xFFFFFFFFu, idx, 0);
        if (idx >= n_list) break;
        const uint32_t b = list[idx];
        const uint32_t c0 = block_off[b], c1 = block_off[b + 1];
        unsigned long long mx = 0;
        for (uint32_t c = c0 + (uint32_t)lane; c < c1; c += 32) mx = max(mx, (unsigned long long)(copy_off[c + 1] - copy_off[c]));
        for (int d = 16; d; d >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
        const unsigned long long sum = copy_off[c1] - copy_off[c0];
        if (lane == 0) poa::poa_bind(w, arena_base + warp * arena_stride, poa::poa_caps_for(sum, mx, level), c1 - c0);
        __syncwarp();
        poa::run_block(w, pr, seq, copy_off, c0, c1, lane, 32);
        if (w.err) {
            if (lane == 0) status[b] = (uint8_t)w.err;
            __syncwarp();
            continue;
        }
        const unsigned long long need = (unsigned long long)(c1 - c0) * w.n_columns;
        if (lane == 0) offs[wib] = atomicAdd(&q->rows_used, need);
        __syncwarp();
        const unsigned long long off = offs[wib];
        if (off + need > rows_cap) { // the host enlarges the pool and runs the block again
            if (lane == 0) status[b] = kPoolFull;
            __syncwarp();
            continue;
        }
        for (uint32_t k = 0; k < c1 - c0; k++) {
            poa::write_row(w, k, rows + off + (unsigned long long)k * w.n_columns, lane, 32);
            if (lane == 0) row_off[c0 + k] = off + (unsigned long long)k * w.n_columns;
        }
        if (lane == 0) {
            block_cols[b] = w.n_columns;
            status[b] = kDone;
            cells += (unsigned long long)w.n_nodes * (mx + 1); // order of magnitude of the work, for the statistics
        }
        __syncwarp();
    }
    if (lane == 0 && cells) atomicAdd(&q->cells, cells);
}

// Long blocks: one block per CTA (poa_core.cuh, run_block_cta); CTAs pull blocks from the same kind of queue, the arena of
// CTA c is arena_base + c * arena_stride.
constexpr int kCtaThreads = 1024;
__global__ void __launch_bounds__(kCtaThreads) k_poa_cta(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ copy_off,
                                                         const uint32_t *__restrict__ block_off, const uint32_t *__restrict__ list,
                                                         unsigned n_list, int level, uint8_t *arena_base, unsigned long long arena_stride,
                                                         poa::Params pr, Queue *q, uint8_t *rows, unsigned long long rows_cap,
                                                         unsigned long long *row_off, uint32_t *block_cols, uint8_t *status)
{
#if defined(__CUDA_ARCH__) // (the team code of poa_core.cuh exists for the device pass only)
    __shared__ poa::Work w;
    __shared__ int32_t seg[64];
    __shared__ unsigned s_idx;
    __shared__ unsigned long long s_off, s_mx;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    unsigned long long cells = 0;
    while (true) {
        if (tid == 0) s_idx = atomicAdd(&q->head, 1u), s_mx = 0;
        __syncthreads();
        const unsigned idx = s_idx;
        if (idx >= n_list) break;
        const uint32_t b = list[idx];
        const uint32_t c0 = block_off[b], c1 = block_off[b + 1];
        unsigned long long mx = 0;
        for (uint32_t c = c0 + (uint32_t)tid; c < c1; c += (uint32_t)nthreads) mx = max(mx, (unsigned long long)(copy_off[c + 1] - copy_off[c]));
        if (mx) atomicMax(&s_mx, mx);
        __syncthreads();
        mx = s_mx;
        const unsigned long long sum = copy_off[c1] - copy_off[c0];
        if (tid == 0) poa::poa_bind(w, arena_base + (size_t)blockIdx.x * arena_stride, poa::poa_caps_for(sum, mx, level), c1 - c0);
        __syncthreads();
        poa::run_block_cta(w, pr, seq, copy_off, c0, c1, tid, nthreads, seg);
        if (w.err) {
            if (tid == 0) status[b] = (uint8_t)w.err;
            __syncthreads();
            continue;
        }
        const unsigned long long need = (unsigned long long)(c1 - c0) * w.n_columns;
        if (tid == 0) s_off = atomicAdd(&q->rows_used, need);
        __syncthreads();
        const unsigned long long off = s_off;
        if (off + need > rows_cap) {
            if (tid == 0) status[b] = kPoolFull;
            __syncthreads();
            continue;
        }
        for (uint32_t k = 0; k < c1 - c0; k++) {
            poa::write_row_cta(w, k, rows + off + (unsigned long long)k * w.n_columns, tid, nthreads);
            if (tid == 0) row_off[c0 + k] = off + (unsigned long long)k * w.n_columns;
        }
        if (tid == 0) {
            block_cols[b] = w.n_columns;
            status[b] = kDone;
            cells += (unsigned long long)w.n_nodes * (mx + 1);
        }
        __syncthreads();
    }
    if (tid == 0 && cells) atomicAdd(&q->cells, cells);
#endif
}

struct Fail {
    int code;
    std::string msg;
};

#define CU(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) throw Fail{LCA_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)}; \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    void alloc(size_t count)
    {
        release();
        n = std::max<size_t>(count, 1);
        CU(cudaMalloc((void **)&p, n * sizeof(T)));
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
    }
    ~DevBuf() { release(); }
};

} // namespace

struct lca_result {
    std::vector<uint8_t> rows;
    std::vector<uint64_t> row_off; // n_copies + 1 (offsets of the rows in `rows`, made contiguous on the host)
    std::vector<uint32_t> cols;    // per block
    lca_stats st{};
};

extern "C" void lca_default_params(lca_params *p)
{
    if (!p) return;
    p->match = 5, p->mismatch = -4, p->gap = -8, p->device = 0;
}

namespace {

void align_impl(const uint8_t *seq, const uint64_t *copy_off, uint64_t n_copies, const uint32_t *block_off, uint32_t n_blocks,
                const lca_params &prm, lca_result &res)
{
    const auto t_begin = std::chrono::steady_clock::now();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw Fail{LCA_ERR_CUDA, "no CUDA device: this library has no CPU fallback"};
    CU(cudaSetDevice(prm.device));
    int major = 0, sms = 0;
    CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, prm.device));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, prm.device));
    if (major < 10) throw Fail{LCA_ERR_CUDA, "device " + std::to_string(prm.device) + " is not sm_100-class"};
    if (block_off[n_blocks] != n_copies) throw Fail{LCA_ERR_ARG, "block_off[n_blocks] != n_copies"};
    const uint64_t n_bases = copy_off[n_copies];
    res.st.n_blocks = n_blocks, res.st.n_copies = n_copies, res.st.n_bases = n_bases;
    res.row_off.assign(n_copies + 1, 0);
    res.cols.assign(n_blocks, 0);
    if (n_blocks == 0) return;

    // ---- per block: sizes, arena need per level
    std::vector<uint64_t> sum(n_blocks), mx(n_blocks);
    uint64_t pool_guess = 0;
    for (uint32_t b = 0; b < n_blocks; b++) {
        uint64_t m = 0;
        for (uint32_t c = block_off[b]; c < block_off[b + 1]; c++) m = std::max<uint64_t>(m, copy_off[c + 1] - copy_off[c]);
        mx[b] = m, sum[b] = copy_off[block_off[b + 1]] - copy_off[block_off[b]];
        // the worst-case arena (level 2) counts 2 * 31 * characters + ... entries in 32 bits (poa_caps_for): 66 M characters per block
        if (sum[b] >= 66000000ull) throw Fail{LCA_ERR_CAPACITY, "block " + std::to_string(b) + " has more than 66 M characters in its copies"};
        pool_guess += (uint64_t)(block_off[b + 1] - block_off[b]) * std::min<uint64_t>(sum[b], m + m / 4 + 64);
    }
    auto need = [&](uint32_t b, int level) { return poa::poa_arena_bytes(poa::poa_caps_for(sum[b], mx[b], level), block_off[b + 1] - block_off[b]) + 256; };

    // ---- device copies of the input
    cudaStream_t stream;
    CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    struct StreamGuard {
        cudaStream_t s;
        ~StreamGuard() { cudaStreamDestroy(s); }
    } sguard{stream};
    DevBuf<uint8_t> d_seq, d_status, d_rows;
    DevBuf<uint64_t> d_copy_off;
    DevBuf<unsigned long long> d_row_off;
    DevBuf<uint32_t> d_block_off, d_cols, d_list;
    DevBuf<Queue> d_q;
    d_seq.alloc(n_bases), d_copy_off.alloc(n_copies + 1), d_block_off.alloc((size_t)n_blocks + 1);
    d_status.alloc(n_blocks), d_cols.alloc(n_blocks), d_row_off.alloc(n_copies + 1), d_list.alloc(n_blocks), d_q.alloc(1);
    CU(cudaMemcpyAsync(d_seq.p, seq, n_bases, cudaMemcpyHostToDevice, stream));
    CU(cudaMemcpyAsync(d_copy_off.p, copy_off, (n_copies + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
    CU(cudaMemcpyAsync(d_block_off.p, block_off, ((size_t)n_blocks + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    CU(cudaMemsetAsync(d_status.p, kPending, n_blocks, stream));
    CU(cudaMemsetAsync(d_q.p, 0, sizeof(Queue), stream));
    res.st.h2d_bytes = n_bases + (n_copies + 1) * 8 + ((size_t)n_blocks + 1) * 4;
    uint64_t rows_cap = pool_guess + 4096;
    d_rows.alloc(rows_cap);

    size_t free_b = 0, total_b = 0;
    lcb_cache_trim_device(prm.device); // scratch parked by an LCB context of this process is invisible to cudaMemGetInfo
    CU(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t budget = (uint64_t)(free_b * 0.85); // for the arenas of one launch
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_poa, kPoaWarps * 32, 0));
    const uint64_t resident_warps = (uint64_t)std::max(per_sm, 1) * sms * kPoaWarps;

    cudaEvent_t ev0, ev1;
    CU(cudaEventCreate(&ev0));
    CU(cudaEventCreate(&ev1));
    struct EvGuard {
        cudaEvent_t a, b;
        ~EvGuard() { cudaEventDestroy(a), cudaEventDestroy(b); }
    } eguard{ev0, ev1};
    const poa::Params pr{prm.match, prm.mismatch, prm.gap};
    std::vector<uint8_t> status(n_blocks, kPending);
    float ms_kernels = 0;

    // Blocks whose longest copy has at least this many characters are given a whole CTA (run_block_cta) instead of a warp.
    // Default threshold 2048 (LCA_CTA_ROWS=<n> overrides, LCA_CTA_ROWS=0 keeps every block on a warp): with it all 1350 blocks of
    // the examples take 8.4 s on one B200 instead of 108 s, byte-identical (profiles/align_examples_full_r2.log).
    uint64_t cta_from = 2048;
    if (const char *e = getenv("LCA_CTA_ROWS")) {
        const uint64_t v = strtoull(e, nullptr, 10);
        cta_from = v ? std::max<uint64_t>(v, 256) : ~0ull;
    }
    auto launch_cta = [&](const std::vector<uint32_t> &list, int level, uint64_t stride) {
        const unsigned ctas = (unsigned)std::min<uint64_t>(std::min<uint64_t>(list.size(), (uint64_t)sms), std::max<uint64_t>(budget / stride, 1));
        DevBuf<uint8_t> arena;
        arena.alloc((size_t)ctas * stride);
        CU(cudaMemcpyAsync(d_list.p, list.data(), list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
        CU(cudaMemsetAsync(&d_q.p->head, 0, sizeof(unsigned), stream));
        CU(cudaEventRecord(ev0, stream));
        k_poa_cta<<<ctas, kCtaThreads, 0, stream>>>(d_seq.p, d_copy_off.p, d_block_off.p, d_list.p, (unsigned)list.size(), level, arena.p,
                                                    (unsigned long long)stride, pr, d_q.p, d_rows.p, (unsigned long long)rows_cap,
                                                    d_row_off.p, d_cols.p, d_status.p);
        CU(cudaEventRecord(ev1, stream));
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(status.data(), d_status.p, n_blocks, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, ev0, ev1);
        ms_kernels += ms;
        res.st.kernel_launches++;
    };
    // one launch over `list` (block ids) at `level` with arenas of `stride` bytes
    auto launch = [&](const std::vector<uint32_t> &list, int level, uint64_t stride) {
        if (!list.empty() && mx[list[0]] >= cta_from) return launch_cta(list, level, stride);
        uint64_t warps = std::min<uint64_t>(std::min<uint64_t>(list.size(), resident_warps), std::max<uint64_t>(budget / stride, 1));
        // never more arenas than the budget holds: whole CTAs, or one partial CTA
        const unsigned ctas = warps >= (uint64_t)kPoaWarps ? (unsigned)(warps / kPoaWarps) : 1u;
        const unsigned threads = warps >= (uint64_t)kPoaWarps ? kPoaWarps * 32u : (unsigned)warps * 32u;
        DevBuf<uint8_t> arena;
        arena.alloc((size_t)ctas * (threads / 32) * stride);
        CU(cudaMemcpyAsync(d_list.p, list.data(), list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
        CU(cudaMemsetAsync(&d_q.p->head, 0, sizeof(unsigned), stream));
        CU(cudaEventRecord(ev0, stream));
        k_poa<<<ctas, threads, 0, stream>>>(d_seq.p, d_copy_off.p, d_block_off.p, d_list.p, (unsigned)list.size(), level, arena.p,
                                                   (unsigned long long)stride, pr, d_q.p, d_rows.p, (unsigned long long)rows_cap,
                                                   d_row_off.p, d_cols.p, d_status.p);
        CU(cudaEventRecord(ev1, stream));
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(status.data(), d_status.p, n_blocks, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, ev0, ev1);
        ms_kernels += ms;
        res.st.kernel_launches++;
    };

    // ---- level by level; inside a level size classes (powers of 4), so that a few long blocks do not dictate the arena
    // of thousands of short ones
    std::vector<uint32_t> todo(n_blocks);
    for (uint32_t b = 0; b < n_blocks; b++) todo[b] = b;
    for (int level = 0; level < 3 && !todo.empty(); level++) {
        std::vector<uint32_t> next;
        std::sort(todo.begin(), todo.end(), [&](uint32_t a, uint32_t b2) { return need(a, level) < need(b2, level); });
        size_t at = 0;
        uint64_t class_cap = 1ull << 20;
        while (at < todo.size()) {
            while (need(todo[at], level) > class_cap) class_cap *= 4;
            size_t end = at;
            const bool long_class = mx[todo[at]] >= cta_from; // a launch is either all-warp or all-CTA
            while (end < todo.size() && need(todo[end], level) <= class_cap && (mx[todo[end]] >= cta_from) == long_class) end++;
            std::vector<uint32_t> list(todo.begin() + (long)at, todo.begin() + (long)end);
            const uint64_t stride = (need(list.back(), level) + 255) & ~255ull;
            if (stride > budget) { // not even one arena of this size: try the next level only if it is smaller (it is not)
                throw Fail{LCA_ERR_CAPACITY, "block " + std::to_string(list.back()) + " (" + std::to_string(sum[list.back()]) +
                                                 " characters, longest copy " + std::to_string(mx[list.back()]) + ") needs " +
                                                 std::to_string(stride >> 20) + " MB of device memory at level " + std::to_string(level)};
            }
            while (!list.empty()) {
                launch(list, level, stride);
                std::vector<uint32_t> again;
                bool pool_full = false;
                for (uint32_t b : list) {
                    if (status[b] == kDone) res.st.blocks_level[level]++;
                    else if (status[b] == 1) next.push_back(b);
                    else if (status[b] == kPoolFull) again.push_back(b), pool_full = true;
                    else throw Fail{LCA_ERR_INTERNAL, "block " + std::to_string(b) + ": alignment kernel failed (status " + std::to_string(status[b]) + ")"};
                }
                if (pool_full) { // enlarge the row pool (what is written stays valid: offsets do not move) and run them again
                    Queue hq;
                    CU(cudaMemcpy(&hq, d_q.p, sizeof hq, cudaMemcpyDeviceToHost));
                    uint64_t extra = 4096;
                    for (uint32_t b : again) extra += (uint64_t)(block_off[b + 1] - block_off[b]) * sum[b];
                    const uint64_t new_cap = hq.rows_used + extra;
                    DevBuf<uint8_t> bigger;
                    bigger.alloc(new_cap);
                    CU(cudaMemcpy(bigger.p, d_rows.p, (size_t)std::min<uint64_t>(hq.rows_used, rows_cap), cudaMemcpyDeviceToDevice));
                    std::swap(bigger.p, d_rows.p);
                    std::swap(bigger.n, d_rows.n);
                    rows_cap = new_cap;
                }
                list.swap(again);
            }
            at = end;
        }
        todo.swap(next);
    }
    if (!todo.empty()) throw Fail{LCA_ERR_CAPACITY, "block " + std::to_string(todo[0]) + " does not fit the worst-case arena"};

    // ---- results: rows made contiguous in copy order
    Queue hq;
    CU(cudaMemcpy(&hq, d_q.p, sizeof hq, cudaMemcpyDeviceToHost));
    res.st.cells = hq.cells;
    std::vector<uint8_t> pool((size_t)hq.rows_used);
    std::vector<unsigned long long> roff(n_copies + 1);
    if (hq.rows_used) CU(cudaMemcpy(pool.data(), d_rows.p, (size_t)hq.rows_used, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(roff.data(), d_row_off.p, n_copies * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(res.cols.data(), d_cols.p, (size_t)n_blocks * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    res.st.d2h_bytes = hq.rows_used + n_copies * 8 + (uint64_t)n_blocks * 4;
    res.rows.resize((size_t)hq.rows_used);
    uint64_t o = 0;
    for (uint32_t b = 0; b < n_blocks; b++)
        for (uint32_t c = block_off[b]; c < block_off[b + 1]; c++) {
            res.row_off[c] = o;
            memcpy(res.rows.data() + o, pool.data() + roff[c], res.cols[b]);
            o += res.cols[b];
        }
    res.row_off[n_copies] = o;
    res.rows.resize((size_t)o);
    res.st.ms_kernels = ms_kernels;
    res.st.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
}

void set_err(char *err, size_t errlen, const std::string &m)
{
    if (err && errlen) snprintf(err, errlen, "%s", m.c_str());
}

} // namespace

extern "C" int lca_align(const uint8_t *seq, const uint64_t *copy_off, uint64_t n_copies, const uint32_t *block_off, uint32_t n_blocks,
                         const lca_params *params, lca_result **out, char *err, size_t errlen)
{
    if (!copy_off || !block_off || !params || !out || (!seq && n_copies && copy_off[n_copies])) return LCA_ERR_ARG;
    lca_result *res = new lca_result;
    try {
        align_impl(seq, copy_off, n_copies, block_off, n_blocks, *params, *res);
    } catch (Fail &f) {
        set_err(err, errlen, f.msg);
        delete res;
        return f.code;
    } catch (std::exception &e) {
        set_err(err, errlen, e.what());
        delete res;
        return LCA_ERR_INTERNAL;
    }
    *out = res;
    return LCA_OK;
}

extern "C" const uint8_t *lca_rows(const lca_result *r, const uint64_t **row_off)
{
    if (!r) return nullptr;
    if (row_off) *row_off = r->row_off.data();
    return r->rows.data();
}
extern "C" uint32_t lca_block_columns(const lca_result *r, uint32_t b) { return r && b < r->cols.size() ? r->cols[b] : 0; }
extern "C" void lca_get_stats(const lca_result *r, lca_stats *st)
{
    if (r && st) *st = r->st;
}
extern "C" void lca_free(lca_result *r) { delete r; }

extern "C" int lca_align_chunk_files(const char *const *files, int n_files, const char *cmd, const char *out_maf, const lca_params *params,
                                     lca_stats *stats, char *err, size_t errlen)
{
    if (!files || n_files < 0 || !out_maf || !params) return LCA_ERR_ARG;
    try {
        // `find ... -name "*.msa" -print0 | sort -z` under LC_ALL=C (sibeliaz:129-130): byte order of the paths
        std::vector<std::string> paths(files, files + n_files);
        std::sort(paths.begin(), paths.end());
        std::vector<uint8_t> seq;
        std::vector<uint64_t> copy_off{0};
        std::vector<uint32_t> block_off{0};
        std::vector<std::string> header;
        for (const std::string &path : paths) {
            MappedFile mf;
            if (!mf.open(path.c_str())) throw Fail{LCA_ERR_IO, "Cannot open file " + path};
            const char *p = (const char *)mf.data, *end = p + mf.size;
            while (p < end) { // one block per line; '@' stands for a newline (sibeliaz:89)
                const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
                if (!eol) eol = end;
                bool open = false;
                const char *t = p;
                while (t < eol) {
                    const char *at = (const char *)memchr(t, '@', (size_t)(eol - t));
                    if (!at) at = eol;
                    if (at > t) {
                        if (*t == '>') { // `cut -d' ' -f2-`, ';' -> ' ', "s " in front (sibeliaz:79)
                            if (open) copy_off.push_back(seq.size());
                            const char *sp = (const char *)memchr(t, ' ', (size_t)(at - t));
                            std::string h = sp ? std::string(sp + 1, at) : std::string(t, at);
                            for (char &ch : h)
                                if (ch == ';') ch = ' ';
                            header.push_back("s " + h);
                            open = true;
                        } else if (open) {
                            seq.insert(seq.end(), (const uint8_t *)t, (const uint8_t *)at);
                        }
                    }
                    t = at + 1;
                }
                if (open) {
                    copy_off.push_back(seq.size());
                    block_off.push_back((uint32_t)(copy_off.size() - 1));
                }
                p = eol + 1;
            }
        }
        lca_result res;
        align_impl(seq.data(), copy_off.data(), copy_off.size() - 1, block_off.data(), (uint32_t)(block_off.size() - 1), *params, res);
        std::string text = "##maf version=1\n# sibeliaz v1.2.7 \n# cmd=" + std::string(cmd ? cmd : "") + "\n";
        text.reserve(text.size() + res.rows.size() + header.size() * 64);
        for (uint32_t b = 0; b + 1 < block_off.size(); b++) {
            text += "\na\n";
            for (uint32_t c = block_off[b]; c < block_off[b + 1]; c++) {
                text += header[c];
                text += ' ';
                text.append((const char *)res.rows.data() + res.row_off[c], res.cols[b]);
                text += '\n';
            }
        }
        FILE *f = fopen(out_maf, "wb");
        if (!f) throw Fail{LCA_ERR_IO, std::string("Cannot open file ") + out_maf};
        const bool ok = fwrite(text.data(), 1, text.size(), f) == text.size();
        if (fclose(f) != 0 || !ok) throw Fail{LCA_ERR_IO, std::string("Cannot write ") + out_maf};
        if (stats) *stats = res.st;
    } catch (Fail &f) {
        set_err(err, errlen, f.msg);
        return f.code;
    } catch (Failure &f) {
        set_err(err, errlen, f.what());
        return f.code;
    } catch (std::exception &e) {
        set_err(err, errlen, e.what());
        return LCA_ERR_INTERNAL;
    }
    return LCA_OK;
}
