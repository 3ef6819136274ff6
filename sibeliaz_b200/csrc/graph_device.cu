// graph_device.cu -- device pipeline of the B200-native junction finder (include/sibeliaz_graph.h): what TwoPaCo's
// VertexEnumeratorImpl computes with a Bloom filter, two passes over 1024 mutex-guarded hash sets and temporary files
// (TwoPaCo/src/graphconstructor/vertexenumerator.h:122-466), done exactly with ONE open-addressing k-mer table in HBM.
//
// Text layout: G = 'N' rec0 'N' rec1 'N' ... 'N', packed to 2 bits per base + 1 "not definite" bit per base, so that the
// k-mer starting at any position is two 64-bit loads and a funnel shift (k <= 31; a wider k-mer, up to k = 255, is up to
// eight words and its table slot names it by the position of one occurrence: graph_kmer.cuh).  One thread per position:
//
//   k_pack        bytes -> 2-bit codes + N mask                                         streaming, 1.4 B per base
//   k_edges       every definite k-mer: canonical key, find-or-insert, OR the neighbour characters it shows (with the
//                 A/T dummies next to 'N', vertexenumerator.h:1046-1058) into the vertex' mask   <- the HBM-bound kernel:
//                 one random 8-byte key probe + one 8-byte mask word per position
//   k_candidates  positions with more than one in- or out-edge (vertexenumerator.h:630-660): flag them and OR their
//                 canonical (prev, next) pair into the vertex (what CandidateFinalFilteringWorker's hash sets collect)
//   k_decide      per table slot: bifurcation <=> >= 2 candidate occurrences whose pairs differ, or agree on an 'N'
//                 (vertexenumerator.h:760-790); collect the bifurcation k-mers
//   radix sort    ids = 1 + rank of the canonical k-mer (deterministic; bifurcationstorage.h:65 sorts too)
//   k_assign_ids, k_mark_stubs, compaction of the flagged positions (genome order), k_emit_ids, k_final_*
//                 -> the junction records (record, position, signed id; stubs numbered) in genome order, on the device
//
// The rules with all their citations are spelled out in the header of include/sibeliaz_graph.h's companion document
// (DESIGN.md section 9); the tests compare this pipeline byte for byte with a CPU restatement of them.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "graph_internal.h"
#include "lcb_internal.h"

namespace {

#include "device_prims.cuh"

#include "graph_kmer.cuh"

__global__ void k_table_init(Slot *slot, uint64_t n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) *(ulonglong2 *)&slot[i] = make_ulonglong2(kEmpty, 0ULL);
}

// ---- bytes -> 2-bit codes + N mask: one thread per 32 bases ------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack(const uint8_t *__restrict__ text, uint64_t words, uint64_t *__restrict__ bits,
                                              uint32_t *__restrict__ nm)
{
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= words) return;
    const uint4 *src = (const uint4 *)(text + w * 32);
    uint64_t b = 0;
    uint32_t m = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint4 v = src[h];
        const uint32_t part[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const unsigned c = (part[j >> 2] >> (8 * (j & 3))) & 0xDFu; // letters to upper case
            unsigned code = 4;
            if (c == 'A') code = 0;
            else if (c == 'C') code = 1;
            else if (c == 'G') code = 2;
            else if (c == 'T') code = 3;
            const int i = h * 16 + j;
            if (code < 4) b |= (uint64_t)code << (2 * i);
            else m |= 1u << i;
        }
    }
    bits[w] = b;
    nm[w] = m;
}

// ---- the table-building pass and the candidate pass: one thread per position (bodies in graph_kmer.cuh) ----------------
template <int W>
__global__ void __launch_bounds__(256) k_edges(Text t, Table tb, unsigned long long *n_kmers)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    const bool live = edges_at<W>(t, tb, p);
    const unsigned m = __ballot_sync(0xFFFFFFFFu, live);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_kmers, (unsigned long long)__popc(m));
}

template <int W>
__global__ void __launch_bounds__(256) k_candidates(Text t, Table tb, uint8_t *__restrict__ flag, unsigned *__restrict__ count)
{
    candidate_at<W>(t, tb, flag, count, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1);
}

template <bool FILL>
__global__ void __launch_bounds__(256) k_decide(Table tb, const unsigned *__restrict__ count, unsigned long long abundance,
                                                unsigned long long *counters /* [0] distinct, [1] bifurcations */,
                                                uint64_t *__restrict__ bif_keys)
{
    // grid-stride over the slots; the counting pass keeps its sums in registers (one atomic pair per block), the filling
    // pass reserves output space once per warp and iteration
    __shared__ unsigned long long su[8], sb[8];
    const unsigned lane = threadIdx.x & 31;
    unsigned long long n_used = 0, n_bif = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t slots = tb.mask + 1, rounds = (slots + stride - 1) / stride;
    for (uint64_t it = 0; it < rounds; it++) {
        const uint64_t s = it * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        bool used = false, bif = false;
        uint64_t key = kEmpty;
        if (s < slots) {
            const ulonglong2 v = *(const ulonglong2 *)&tb.slot[s];
            key = v.x;
            used = key != kEmpty;
            if (used) bif = is_bifurcation(v.y) && (!count || (unsigned long long)count[s] <= abundance);
        }
        if (!FILL) {
            n_used += used, n_bif += bif;
        } else {
            const unsigned mb = __ballot_sync(0xFFFFFFFFu, bif);
            if (mb) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&counters[1], (unsigned long long)__popc(mb));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (bif) bif_keys[base + __popc(mb & ((1u << lane) - 1))] = key;
            }
        }
    }
    if (!FILL) {
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            n_used += __shfl_xor_sync(0xFFFFFFFFu, n_used, d);
            n_bif += __shfl_xor_sync(0xFFFFFFFFu, n_bif, d);
        }
        if (lane == 0) su[threadIdx.x >> 5] = n_used, sb[threadIdx.x >> 5] = n_bif;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long a = 0, c = 0;
            for (int w = 0; w < 8; w++) a += su[w], c += sb[w];
            if (a) atomicAdd(&counters[0], a);
            if (c) atomicAdd(&counters[1], c);
        }
    }
}

__global__ void k_assign_ids(Table tb, const uint64_t *__restrict__ sorted, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t s = find(tb, sorted[i]);
    atomicOr(&tb.slot[s].info, (unsigned long long)(i + 1) << kIdShift);
}

// ---- ordered compaction of the flagged positions: 256 threads x 16 flags per block ---------------------------------
constexpr int kFlagTile = 4096;
__device__ __forceinline__ unsigned flags16(const uint8_t *flag, uint64_t base, uint64_t n)
{
    unsigned m = 0; // bit j: flag[base + j]
    if (base + 16 <= n) {
        const uint4 v = *(const uint4 *)(flag + base);
        const uint32_t part[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 16; j++)
            if ((part[j >> 2] >> (8 * (j & 3))) & 0xFFu) m |= 1u << j;
    } else {
        for (int j = 0; j < 16 && base + j < n; j++)
            if (flag[base + j]) m |= 1u << j;
    }
    return m;
}
__global__ void __launch_bounds__(256) k_flag_count(const uint8_t *__restrict__ flag, uint64_t n, unsigned *__restrict__ block_count)
{
    __shared__ unsigned s[8];
    const uint64_t base = (uint64_t)blockIdx.x * kFlagTile + (uint64_t)threadIdx.x * 16;
    unsigned c = base < n ? __popc(flags16(flag, base, n)) : 0;
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int w = 0; w < 8; w++) tot += s[w];
        block_count[blockIdx.x] = tot;
    }
}
__global__ void __launch_bounds__(256) k_flag_write(const uint8_t *__restrict__ flag, uint64_t n, const unsigned *__restrict__ block_off,
                                                    uint64_t *__restrict__ out)
{
    __shared__ unsigned s[256];
    const uint64_t base = (uint64_t)blockIdx.x * kFlagTile + (uint64_t)threadIdx.x * 16;
    const unsigned m = base < n ? flags16(flag, base, n) : 0;
    const unsigned c = __popc(m);
    s[threadIdx.x] = c;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        const unsigned x = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
        __syncthreads();
        s[threadIdx.x] += x;
        __syncthreads();
    }
    uint64_t w = (uint64_t)block_off[blockIdx.x] + s[threadIdx.x] - c;
    for (unsigned mm = m; mm; mm &= mm - 1) out[w++] = base + (uint64_t)(__ffs((int)mm) - 1);
}

// first and last k-mer of every record that yields a task (vertexenumerator.h:913-920); bit 1 of the position's flag
__global__ void k_mark_stubs(const uint64_t *__restrict__ goff, int n_records, int k, uint8_t *__restrict__ flag)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_records) return;
    const uint64_t len = goff[r + 1] - goff[r] - 1;
    if (len < (uint64_t)k) return; // shorter records yield no task (vertexenumerator.h:1176)
    flag[goff[r]] |= 2;
    flag[goff[r] + len - (uint64_t)k] |= 2; // the same byte when len == k; no other thread touches these two
}

template <int W>
__global__ void k_emit_ids(Text t, Table tb, const uint8_t *__restrict__ flag, const uint64_t *__restrict__ pos, unsigned n,
                           int32_t *__restrict__ id)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t p = pos[i];
    id[i] = id_at<W>(t, tb, flag[p], p);
}

// ---- k > 31: the bifurcation k-mers' words for the ranking sort, and the ids from the sorted order ------------------
template <int W>
__global__ void k_canon_words(Text t, const uint64_t *__restrict__ bif_keys, unsigned n, uint64_t *__restrict__ words)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) canon_words_at<W>(t, bif_keys, n, i, words);
}
template <int W>
__global__ void k_assign_ids_wide(Text t, Table tb, const uint64_t *__restrict__ bif_keys, const unsigned *__restrict__ perm, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) assign_id_at<W>(t, tb, bif_keys[perm[i]], i);
}

// ---- final list: entries with id != 0 in order, stubs numbered in order, position -> (record, offset) ----------------
constexpr int kFinalTile = 2048; // 256 threads x 8 entries
__global__ void __launch_bounds__(256) k_final_count(const int32_t *__restrict__ id, unsigned n, unsigned *__restrict__ kept, unsigned *__restrict__ stubs)
{
    __shared__ unsigned sk[8], ss[8];
    const unsigned base = blockIdx.x * kFinalTile + threadIdx.x * 8;
    unsigned ck = 0, cs = 0;
    for (int j = 0; j < 8; j++)
        if (base + j < n) {
            const int32_t v = id[base + j];
            ck += v != 0;
            cs += v == kStub;
        }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        ck += __shfl_xor_sync(0xFFFFFFFFu, ck, d);
        cs += __shfl_xor_sync(0xFFFFFFFFu, cs, d);
    }
    if ((threadIdx.x & 31) == 0) sk[threadIdx.x >> 5] = ck, ss[threadIdx.x >> 5] = cs;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned a = 0, b = 0;
        for (int w = 0; w < 8; w++) a += sk[w], b += ss[w];
        kept[blockIdx.x] = a;
        stubs[blockIdx.x] = b;
    }
}
__global__ void __launch_bounds__(256) k_final_write(const int32_t *__restrict__ id, const uint64_t *__restrict__ pos, unsigned n,
                                                     const unsigned *__restrict__ kept_off, const unsigned *__restrict__ stub_off,
                                                     const uint64_t *__restrict__ goff, int n_records, int32_t stub_base,
                                                     uint32_t *__restrict__ out_chr, uint32_t *__restrict__ out_pos, int32_t *__restrict__ out_id)
{
    __shared__ unsigned sk[256], ss[256];
    const unsigned base = blockIdx.x * kFinalTile + threadIdx.x * 8;
    int32_t v[8];
    unsigned ck = 0, cs = 0;
    for (int j = 0; j < 8; j++) {
        v[j] = base + j < n ? id[base + j] : 0;
        ck += v[j] != 0;
        cs += v[j] == kStub;
    }
    sk[threadIdx.x] = ck, ss[threadIdx.x] = cs;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        const unsigned a = threadIdx.x >= d ? sk[threadIdx.x - d] : 0, b = threadIdx.x >= d ? ss[threadIdx.x - d] : 0;
        __syncthreads();
        sk[threadIdx.x] += a, ss[threadIdx.x] += b;
        __syncthreads();
    }
    unsigned w = kept_off[blockIdx.x] + sk[threadIdx.x] - ck, sn = stub_off[blockIdx.x] + ss[threadIdx.x] - cs;
    for (int j = 0; j < 8; j++) {
        if (v[j] == 0) continue;
        const uint64_t p = pos[base + j];
        int lo = 0, hi = n_records; // goff[lo] <= p < goff[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (goff[mid] <= p) lo = mid;
            else hi = mid;
        }
        out_chr[w] = (uint32_t)lo;
        out_pos[w] = (uint32_t)(p - goff[lo]);
        out_id[w] = v[j] == kStub ? stub_base + (int32_t)sn++ : v[j];
        w++;
    }
}

// ---- host helpers ----------------------------------------------------------------------------------------------------
struct Scope { // frees everything on every exit path
    std::vector<void *> dev;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t l2_fetch_before = 0; // != 0: the device's L2 fetch granularity was changed for this stage (LCG_L2_FETCH)
    ~Scope()
    {
        if (l2_fetch_before) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, l2_fetch_before);
        for (void *p : dev) cudaFree(p);
        for (cudaEvent_t e : ev)
            if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
};

#define CU(x)                                                                    \
    do {                                                                         \
        cudaError_t e_ = (x);                                                    \
        if (e_ != cudaSuccess) {                                                 \
            err = std::string(#x) + ": " + cudaGetErrorString(e_);               \
            return LCG_ERR_CUDA;                                                 \
        }                                                                        \
    } while (0)

template <typename T>
int dalloc(Scope &sc, T **p, size_t n, std::string &err)
{
    void *q = nullptr;
    CU(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
    sc.dev.push_back(q);
    *p = (T *)q;
    return LCG_OK;
}

int exclusive_scan_u32(Scope &sc, const unsigned *in, unsigned *out, size_t n, unsigned *d_total, unsigned *d_tile, uint64_t &launches,
                       std::string &err)
{
    const unsigned tiles = (unsigned)((n + kScanTile - 1) / kScanTile);
    k_scan_tiles<<<tiles, 256, 0, sc.stream>>>(in, out, d_tile, n);
    k_scan_sums<<<1, 32, 0, sc.stream>>>(d_tile, tiles, d_total);
    k_scan_add<<<(unsigned)((n + 255) / 256), 256, 0, sc.stream>>>(out, d_tile, n);
    launches += 3;
    CU(cudaGetLastError());
    return LCG_OK;
}

// k-mer width in 64-bit words -> the matching instantiation (W = 1: k <= 31)
#define LCG_FOR_WIDTH(width, ...)                         \
    switch (width) {                                      \
    case 1: { constexpr int W = 1; __VA_ARGS__; } break;  \
    default: LCG_FOR_WIDE(width, __VA_ARGS__)             \
    }
#define LCG_FOR_WIDE(width, ...)                          \
    switch (width) {                                      \
    case 2: { constexpr int W = 2; __VA_ARGS__; } break;  \
    case 3: { constexpr int W = 3; __VA_ARGS__; } break;  \
    case 4: { constexpr int W = 4; __VA_ARGS__; } break;  \
    case 5: { constexpr int W = 5; __VA_ARGS__; } break;  \
    case 6: { constexpr int W = 6; __VA_ARGS__; } break;  \
    case 7: { constexpr int W = 7; __VA_ARGS__; } break;  \
    default: { constexpr int W = 8; __VA_ARGS__; } break; \
    }

} // namespace

namespace lcg {

int run_device(const DeviceInput &in, DeviceOutput &out, std::string &err)
{
    const auto t_begin = std::chrono::steady_clock::now();
    auto ms_since = [](std::chrono::steady_clock::time_point a) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
    };
    const int k = in.k;
    if (k < 1 || k > kMaxK) {
        err = "k must be between 1 and " + std::to_string(kMaxK);
        return LCG_ERR_ARG;
    }
    const int width = (2 * k + 63) / 64; // words of a k-mer
    const bool trace = getenv("LCG_TRACE") != nullptr;
    cudaStream_t trace_stream = nullptr;
    auto lap = [&](const char *what) { // developer aid: synchronising stage timer
        if (!trace) return;
        if (trace_stream) cudaStreamSynchronize(trace_stream);
        fprintf(stderr, "[graph] %-26s %9.2f ms\n", what, ms_since(t_begin));
    };
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        err = "no CUDA device: this library has no CPU fallback";
        return LCG_ERR_CUDA;
    }
    CU(cudaSetDevice(in.device));
    {
        int major = 0;
        CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, in.device));
        if (major < 10) {
            err = "device " + std::to_string(in.device) + " is not sm_100-class";
            return LCG_ERR_CUDA;
        }
    }
    Scope sc;
    CU(cudaStreamCreateWithFlags(&sc.stream, cudaStreamNonBlocking));
    for (auto &e : sc.ev) CU(cudaEventCreate(&e));
    trace_stream = sc.stream;
    if (const char *e = getenv("LCG_L2_FETCH")) {
        // A/B switch: the table passes read one 16-byte slot per position at a random address; the granularity at which L2
        // fetches a missing line from HBM (32 / 64 / 128 bytes, a hint) decides how many bytes each of them costs.  Restored
        // on exit: the stages after this one walk consecutive records.
        size_t before = 0;
        if (cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity) == cudaSuccess && before &&
            cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e)) == cudaSuccess) {
            sc.l2_fetch_before = before;
            size_t now = 0;
            cudaDeviceGetLimit(&now, cudaLimitMaxL2FetchGranularity);
            if (trace) fprintf(stderr, "[graph] L2 fetch granularity %zu -> %zu bytes\n", before, now);
        } else {
            cudaGetLastError();
        }
    }
    lap("context + stream");
    lcg_stats &st = out.st;
    st.n_records = (uint64_t)in.n_records;
    // ---- layout of G
    std::vector<uint64_t> goff((size_t)in.n_records + 1);
    uint64_t g = 1;
    for (int r = 0; r < in.n_records; r++) {
        goff[(size_t)r] = g;
        g += in.len[r] + 1;
        st.n_bases += in.len[r];
    }
    goff[(size_t)in.n_records] = g;
    const uint64_t G = g;                                // positions; G[0] and G[G-1] are 'N'
    if (width > 1 && G >= kRepMask) { // a wide slot names its k-mer by a 40-bit text position
        err = "more than 2^40 characters with k > 31";
        return LCG_ERR_ARG;
    }
    const uint64_t words = (G + 31) / 32 + 2;            // + padding words read by windows near the end
    const uint64_t padded = words * 32;
    // ---- k-mer table: a power of two >= 2 x positions
    uint64_t cap = 1 << 16;
    while (cap < 2 * G) cap <<= 1;
    st.table_slots = cap;
    const bool finite_abundance = in.abundance != UINT64_MAX;
    {
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const double need0 = (double)cap * 16.0 + (double)padded * 2.375 + (64 << 20);
        if (need0 + (double)cap * 4.0 > (double)free_b && lcb_cache_device_bytes(in.device) > 0) {
            lcb_cache_trim_device(in.device); // scratch parked by the LCB stage (lcb_warmup / lcb_destroy) is reclaimable
            CU(cudaMemGetInfo(&free_b, &total_b));
        }
        const double need = (double)cap * (16.0 + (finite_abundance ? 4.0 : 0.0)) + (double)padded * (1.0 + 1.0 + 0.25 + 0.125) + (64 << 20);
        if (need > (double)free_b) {
            err = "the k-mer table (" + std::to_string((unsigned long long)(need / (1 << 20))) + " MiB) does not fit the device (" +
                  std::to_string((unsigned long long)(free_b >> 20)) + " MiB free)";
            return LCG_ERR_MEMORY;
        }
    }
    uint8_t *d_text = nullptr, *d_flag = nullptr;
    uint64_t *d_bits = nullptr;
    Slot *d_slot = nullptr;
    uint32_t *d_nm = nullptr;
    unsigned long long *d_ctr = nullptr;
    unsigned *d_cnt = nullptr;
    int rc;
    if ((rc = dalloc(sc, &d_text, padded, err))) return rc;
    if ((rc = dalloc(sc, &d_bits, words, err))) return rc;
    if ((rc = dalloc(sc, &d_nm, words, err))) return rc;
    if ((rc = dalloc(sc, &d_flag, padded, err))) return rc;
    if ((rc = dalloc(sc, &d_slot, cap, err))) return rc;
    if ((rc = dalloc(sc, &d_ctr, 4, err))) return rc;
    if (finite_abundance && (rc = dalloc(sc, &d_cnt, cap, err))) return rc;
    uint64_t *d_goff = nullptr;
    if ((rc = dalloc(sc, &d_goff, goff.size(), err))) return rc;
    CU(cudaMemcpyAsync(d_goff, goff.data(), goff.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, sc.stream));
    lap("allocations");
    // ---- sequences -> device (separators and padding stay 'N')
    const auto t_h2d = std::chrono::steady_clock::now();
    CU(cudaMemsetAsync(d_text, 'N', padded, sc.stream));
    CU(cudaStreamSynchronize(sc.stream));
    for (int r = 0; r < in.n_records; r++)
        if (in.len[r]) CU(cudaMemcpyAsync(d_text + goff[(size_t)r], in.seq[r], in.len[r], cudaMemcpyHostToDevice, sc.stream));
    CU(cudaStreamSynchronize(sc.stream));
    st.ms_h2d = ms_since(t_h2d);
    lap("h2d");
    // ---- device pipeline
    CU(cudaEventRecord(sc.ev[0], sc.stream));
    k_table_init<<<(unsigned)((cap + 255) / 256), 256, 0, sc.stream>>>(d_slot, cap);
    CU(cudaMemsetAsync(d_flag, 0, padded, sc.stream));
    CU(cudaMemsetAsync(d_ctr, 0, 4 * sizeof(unsigned long long), sc.stream));
    if (d_cnt) CU(cudaMemsetAsync(d_cnt, 0, cap * sizeof(unsigned), sc.stream));
    k_pack<<<(unsigned)((words + 255) / 256), 256, 0, sc.stream>>>(d_text, words, d_bits, d_nm);
    const Text text{d_bits, d_nm, G, k};
    const Table tb{d_slot, cap - 1};
    const unsigned pos_blocks = (unsigned)((G + 255) / 256);
    lap("init + pack");
    CU(cudaEventRecord(sc.ev[1], sc.stream));
    LCG_FOR_WIDTH(width, k_edges<W><<<pos_blocks, 256, 0, sc.stream>>>(text, tb, d_ctr + 2));
    CU(cudaEventRecord(sc.ev[2], sc.stream));
    lap("k_edges");
    LCG_FOR_WIDTH(width, k_candidates<W><<<pos_blocks, 256, 0, sc.stream>>>(text, tb, d_flag, d_cnt));
    if (in.n_records) k_mark_stubs<<<(in.n_records + 255) / 256, 256, 0, sc.stream>>>(d_goff, in.n_records, k, d_flag);
    lap("k_candidates + stubs");
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, in.device);
    const unsigned slot_blocks = (unsigned)std::min<uint64_t>((cap + 255) / 256, (uint64_t)sms * 16);
    k_decide<false><<<slot_blocks, 256, 0, sc.stream>>>(tb, d_cnt, (unsigned long long)in.abundance, d_ctr, nullptr);
    st.kernel_launches += 6;
    unsigned long long h_ctr[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(h_ctr, d_ctr, sizeof h_ctr, cudaMemcpyDeviceToHost, sc.stream));
    CU(cudaStreamSynchronize(sc.stream));
    CU(cudaGetLastError());
    st.n_distinct = h_ctr[0];
    st.n_bifurcations = h_ctr[1];
    st.n_kmers = h_ctr[2];
    if (h_ctr[1] >= (1ULL << 29)) {
        err = "more than 2^29 junction vertices";
        return LCG_ERR_ARG;
    }
    const unsigned nb = (unsigned)h_ctr[1];
    lap("k_decide (count)");
    // ---- ids: sort the bifurcation k-mers
    uint64_t *d_bif = nullptr, *d_sorted = nullptr;
    if ((rc = dalloc(sc, &d_bif, nb, err))) return rc;
    if ((rc = dalloc(sc, &d_sorted, nb, err))) return rc;
    if (nb) {
        CU(cudaMemsetAsync(d_ctr + 1, 0, sizeof(unsigned long long), sc.stream));
        k_decide<true><<<slot_blocks, 256, 0, sc.stream>>>(tb, d_cnt, (unsigned long long)in.abundance, d_ctr, d_bif);
        unsigned *perm = nullptr, *tmp = nullptr, *d_hist = nullptr, *d_small = nullptr;
        const unsigned sblocks = (nb + kSortTile - 1) / kSortTile;
        const size_t hn = (size_t)256 * sblocks;
        if ((rc = dalloc(sc, &perm, nb, err))) return rc;
        if ((rc = dalloc(sc, &tmp, nb, err))) return rc;
        if ((rc = dalloc(sc, &d_hist, hn + hn / kScanTile + 8, err))) return rc;
        if ((rc = dalloc(sc, &d_small, 8, err))) return rc;
        k_iota<<<(nb + 255) / 256, 256, 0, sc.stream>>>(perm, nb);
        st.kernel_launches += 2;
        // stable LSD passes over the 2k key bits; a wide k-mer's words are written out first (d_bif then holds slot keys:
        // fingerprint | position of the representative) and sorted word by word, least significant first
        uint64_t *d_words = nullptr;
        if (width > 1) {
            if ((rc = dalloc(sc, &d_words, (size_t)width * nb, err))) return rc;
            LCG_FOR_WIDE(width, k_canon_words<W><<<(nb + 255) / 256, 256, 0, sc.stream>>>(text, d_bif, nb, d_words));
            st.kernel_launches += 1;
        }
        for (int w = 0; w < width; w++) {
            const uint64_t *keys = width > 1 ? d_words + (size_t)w * nb : d_bif;
            const int bits = std::min(64, 2 * k - 64 * w);
            for (int shift = 0; shift < bits; shift += 8) {
                k_radix_hist<uint64_t><<<sblocks, 256, 0, sc.stream>>>(perm, keys, shift, nb, d_hist, 0);
                if ((rc = exclusive_scan_u32(sc, d_hist, d_hist, hn, d_small, d_hist + hn, st.kernel_launches, err))) return rc;
                k_radix_scatter<uint64_t><<<sblocks, 256, 0, sc.stream>>>(perm, tmp, keys, shift, nb, d_hist, 0);
                std::swap(perm, tmp);
                st.kernel_launches += 2;
            }
        }
        if (width == 1) {
            k_gather<<<(nb + 255) / 256, 256, 0, sc.stream>>>(perm, d_bif, d_sorted, nb);
            k_assign_ids<<<(nb + 255) / 256, 256, 0, sc.stream>>>(tb, d_sorted, nb);
            st.kernel_launches += 2;
        } else {
            LCG_FOR_WIDE(width, k_assign_ids_wide<W><<<(nb + 255) / 256, 256, 0, sc.stream>>>(text, tb, d_bif, perm, nb));
            st.kernel_launches += 1;
        }
    }
    lap("collect + sort + ids");
    // ---- flagged positions in genome order
    const unsigned fblocks = (unsigned)((G + kFlagTile - 1) / kFlagTile);
    unsigned *d_bc = nullptr, *d_bo = nullptr, *d_tile = nullptr, *d_total = nullptr;
    if ((rc = dalloc(sc, &d_bc, fblocks, err))) return rc;
    if ((rc = dalloc(sc, &d_bo, fblocks, err))) return rc;
    if ((rc = dalloc(sc, &d_tile, fblocks / kScanTile + 8, err))) return rc;
    if ((rc = dalloc(sc, &d_total, 4, err))) return rc;
    k_flag_count<<<fblocks, 256, 0, sc.stream>>>(d_flag, G, d_bc);
    st.kernel_launches += 1;
    if ((rc = exclusive_scan_u32(sc, d_bc, d_bo, fblocks, d_total, d_tile, st.kernel_launches, err))) return rc;
    unsigned nc = 0;
    CU(cudaMemcpyAsync(&nc, d_total, sizeof nc, cudaMemcpyDeviceToHost, sc.stream));
    CU(cudaStreamSynchronize(sc.stream));
    CU(cudaGetLastError());
    st.n_candidates = nc;
    lap("flag count + scan");
    uint64_t *d_pos = nullptr;
    int32_t *d_id = nullptr;
    if ((rc = dalloc(sc, &d_pos, nc, err))) return rc;
    if ((rc = dalloc(sc, &d_id, nc, err))) return rc;
    unsigned nj = 0, ns = 0;
    uint32_t *d_jchr = nullptr, *d_jpos = nullptr;
    int32_t *d_jid = nullptr;
    if (nc) {
        k_flag_write<<<fblocks, 256, 0, sc.stream>>>(d_flag, G, d_bo, d_pos);
        LCG_FOR_WIDTH(width, k_emit_ids<W><<<(nc + 255) / 256, 256, 0, sc.stream>>>(text, tb, d_flag, d_pos, nc, d_id));
        const unsigned qblocks = (nc + kFinalTile - 1) / kFinalTile;
        unsigned *d_kc = nullptr, *d_sc = nullptr, *d_ko = nullptr, *d_so = nullptr, *d_tile2 = nullptr, *d_tot2 = nullptr;
        if ((rc = dalloc(sc, &d_kc, qblocks, err))) return rc;
        if ((rc = dalloc(sc, &d_sc, qblocks, err))) return rc;
        if ((rc = dalloc(sc, &d_ko, qblocks, err))) return rc;
        if ((rc = dalloc(sc, &d_so, qblocks, err))) return rc;
        if ((rc = dalloc(sc, &d_tile2, qblocks / kScanTile + 8, err))) return rc;
        if ((rc = dalloc(sc, &d_tot2, 4, err))) return rc;
        k_final_count<<<qblocks, 256, 0, sc.stream>>>(d_id, nc, d_kc, d_sc);
        st.kernel_launches += 3;
        if ((rc = exclusive_scan_u32(sc, d_kc, d_ko, qblocks, d_tot2, d_tile2, st.kernel_launches, err))) return rc;
        if ((rc = exclusive_scan_u32(sc, d_sc, d_so, qblocks, d_tot2 + 1, d_tile2, st.kernel_launches, err))) return rc;
        unsigned tot[2] = {0, 0};
        CU(cudaMemcpyAsync(tot, d_tot2, sizeof tot, cudaMemcpyDeviceToHost, sc.stream));
        CU(cudaStreamSynchronize(sc.stream));
        CU(cudaGetLastError());
        nj = tot[0], ns = tot[1];
        if ((uint64_t)nb + 42 + ns >= 0x7FFFFFFFull) {
            err = "vertex ids exceed 31 bits";
            return LCG_ERR_ARG;
        }
        if ((rc = dalloc(sc, &d_jchr, nj, err))) return rc;
        if ((rc = dalloc(sc, &d_jpos, nj, err))) return rc;
        if ((rc = dalloc(sc, &d_jid, nj, err))) return rc;
        if (nj) {
            k_final_write<<<qblocks, 256, 0, sc.stream>>>(d_id, d_pos, nc, d_ko, d_so, d_goff, in.n_records, (int32_t)(nb + 42), d_jchr, d_jpos, d_jid);
            st.kernel_launches += 1;
        }
    }
    CU(cudaEventRecord(sc.ev[3], sc.stream));
    CU(cudaStreamSynchronize(sc.stream));
    CU(cudaGetLastError());
    lap("emit + final list");
    float ms = 0;
    cudaEventElapsedTime(&ms, sc.ev[0], sc.ev[3]);
    st.ms_device = ms;
    cudaEventElapsedTime(&ms, sc.ev[1], sc.ev[2]);
    st.ms_edges = ms;
    st.n_junctions = nj;
    // ---- results
    const auto t_d2h = std::chrono::steady_clock::now();
    out.n = nj; // plain arrays: a std::vector would zero-fill 180 MB on 4x100 Mbp before the copy overwrites it
    uint32_t last_chr = 0;
    if (in.keep_on_device) { // the fused pipeline reads the records where they are; a host copy is made on demand (download)
        if (nj) CU(cudaMemcpyAsync(&last_chr, d_jchr + (nj - 1), sizeof last_chr, cudaMemcpyDeviceToHost, sc.stream));
        CU(cudaStreamSynchronize(sc.stream));
    } else {
        out.chr.reset(new uint32_t[std::max<size_t>(nj, 1)]);
        out.pos.reset(new uint32_t[std::max<size_t>(nj, 1)]);
        out.id.reset(new int32_t[std::max<size_t>(nj, 1)]);
    }
    if (nj && !in.keep_on_device) {
        CU(cudaMemcpyAsync(out.chr.get(), d_jchr, (size_t)nj * sizeof(uint32_t), cudaMemcpyDeviceToHost, sc.stream));
        CU(cudaMemcpyAsync(out.pos.get(), d_jpos, (size_t)nj * sizeof(uint32_t), cudaMemcpyDeviceToHost, sc.stream));
        CU(cudaMemcpyAsync(out.id.get(), d_jid, (size_t)nj * sizeof(int32_t), cudaMemcpyDeviceToHost, sc.stream));
        CU(cudaStreamSynchronize(sc.stream));
    }
    st.ms_d2h = ms_since(t_d2h);
    lap("d2h");
    if (in.keep_on_device) { // ownership of these blocks moves to the caller
        Resident &r = out.resident;
        r.device = in.device, r.k = k, r.n_records = in.n_records;
        r.d_text = d_text, r.d_goff = d_goff, r.d_chr = d_jchr, r.d_pos = d_jpos, r.d_id = d_jid;
        r.n_junctions = nj;
        r.n_vertices = (ns ? (uint64_t)nb + 42 + ns - 1 : (uint64_t)nb) + 1;
        r.last_chr = last_chr;
        for (void *keep : {(void *)d_text, (void *)d_goff, (void *)d_jchr, (void *)d_jpos, (void *)d_jid})
            sc.dev.erase(std::remove(sc.dev.begin(), sc.dev.end(), keep), sc.dev.end());
    }
    st.ms_total = ms_since(t_begin);
    return LCG_OK;
}

int download(const Resident &r, uint32_t *chr, uint32_t *pos, int32_t *id, std::string &err)
{
    CU(cudaSetDevice(r.device));
    const size_t n = (size_t)r.n_junctions;
    if (n && chr) CU(cudaMemcpy(chr, r.d_chr, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (n && pos) CU(cudaMemcpy(pos, r.d_pos, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (n && id) CU(cudaMemcpy(id, r.d_id, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return LCG_OK;
}

void preload_kernels()
{
    // CUDA loads kernels lazily, on first use: a host that knows it will build a graph touches them off the critical path
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_table_init);
    cudaFuncGetAttributes(&fa, k_pack);
    cudaFuncGetAttributes(&fa, k_edges<1>); // the one-word instantiations: what the wrapper's defaults (k = 15, 25) use
    cudaFuncGetAttributes(&fa, k_candidates<1>);
    cudaFuncGetAttributes(&fa, k_mark_stubs);
    cudaFuncGetAttributes(&fa, k_decide<false>);
    cudaFuncGetAttributes(&fa, k_decide<true>);
    cudaFuncGetAttributes(&fa, k_assign_ids);
    cudaFuncGetAttributes(&fa, k_flag_count);
    cudaFuncGetAttributes(&fa, k_flag_write);
    cudaFuncGetAttributes(&fa, k_emit_ids<1>);
    cudaFuncGetAttributes(&fa, k_final_count);
    cudaFuncGetAttributes(&fa, k_final_write);
    cudaFuncGetAttributes(&fa, k_iota);
    cudaFuncGetAttributes(&fa, k_scan_tiles);
    cudaFuncGetAttributes(&fa, k_scan_sums);
    cudaFuncGetAttributes(&fa, k_scan_add);
    cudaFuncGetAttributes(&fa, k_radix_hist<uint64_t>);
    cudaFuncGetAttributes(&fa, k_radix_scatter<uint64_t>);
    cudaFuncGetAttributes(&fa, k_gather<uint64_t>);
}

void free_resident(Resident &r)
{
    if (!r.d_text && !r.d_goff && !r.d_chr) return;
    cudaSetDevice(r.device);
    for (void *p : {(void *)r.d_text, (void *)r.d_goff, (void *)r.d_chr, (void *)r.d_pos, (void *)r.d_id})
        if (p) cudaFree(p);
    r = Resident();
}

} // namespace lcg
