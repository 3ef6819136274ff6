// graph_emu.cpp -- TEST INFRASTRUCTURE: the junction finder's per-position device code (sibeliaz_b200/csrc/graph_kmer.cuh),
// compiled for the host exactly as written and driven pass by pass the way graph_device.cu's run_device launches it:
// one call per text position, from several free-running host threads (atomics are real atomics), so the table protocol
// (claim by compare-and-swap, k-mers wider than a word compared through their representative's text position) and all
// the index arithmetic are what runs on the GPU.  Its junction file must equal the CPU restatement's byte for byte
// (tests/test_graph_emulation.py).  The passes that do not depend on k (packing, compaction of the flagged positions, the
// final list) are done plainly here; they are covered by the GPU tests.
//
//   graph_emu <k> <threads> <abundance|0> <out.dbg> <fasta>...
#include <algorithm>
#include <cctype>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#define __device__
#define __host__
#define __forceinline__ inline
static inline uint64_t __brevll(uint64_t x)
{
    uint64_t r = 0;
    for (int i = 0; i < 64; i++) r |= ((x >> i) & 1ULL) << (63 - i);
    return r;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long val)
{
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

namespace {
#include "../sibeliaz_b200/csrc/graph_kmer.cuh"

struct Input {
    std::vector<std::string> rec;
    int k, threads;
    uint64_t abundance;
};

template <class F>
void parallel_positions(int threads, uint64_t first, uint64_t last, F f) // f(p) for first <= p < last, interleaved over the threads
{
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([=]() {
            for (uint64_t p = first + (uint64_t)t; p < last; p += (uint64_t)threads) f(p);
        });
    for (auto &th : pool) th.join();
}

struct Junction {
    uint32_t chr, pos;
    int64_t id;
};

template <int W>
std::vector<Junction> run(const Input &in)
{
    const int k = in.k;
    // ---- layout of G and packing (run_device + k_pack)
    const int R = (int)in.rec.size();
    std::vector<uint64_t> goff((size_t)R + 1);
    uint64_t g = 1;
    for (int r = 0; r < R; r++) {
        goff[(size_t)r] = g;
        g += in.rec[(size_t)r].size() + 1;
    }
    goff[(size_t)R] = g;
    const uint64_t G = g, words = (G + 31) / 32 + 2, padded = words * 32;
    std::string text(padded, 'N');
    for (int r = 0; r < R; r++) memcpy(&text[goff[(size_t)r]], in.rec[(size_t)r].data(), in.rec[(size_t)r].size());
    std::vector<uint64_t> bits(words, 0);
    std::vector<uint32_t> nm(words, 0);
    for (uint64_t i = 0; i < padded; i++) {
        const unsigned c = (unsigned char)text[i] & 0xDFu;
        const int code = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
        if (code < 4) bits[i >> 5] |= (uint64_t)code << (2 * (i & 31));
        else nm[i >> 5] |= 1u << (i & 31);
    }
    uint64_t cap = 1 << 16;
    while (cap < 2 * G) cap <<= 1;
    std::vector<Slot> slot(cap, Slot{kEmpty, 0ULL});
    std::vector<uint8_t> flag(padded, 0);
    const bool finite = in.abundance != UINT64_MAX;
    std::vector<unsigned> count(finite ? cap : 0, 0);
    const Text t{bits.data(), nm.data(), G, k};
    const Table tb{slot.data(), cap - 1};
    const uint64_t p_end = ((G + 255) / 256) * 256 + 1; // the launch covers whole blocks: positions past the text do nothing
    // ---- passes
    parallel_positions(in.threads, 1, p_end, [&](uint64_t p) { edges_at<W>(t, tb, p); });
    parallel_positions(in.threads, 1, p_end, [&](uint64_t p) { candidate_at<W>(t, tb, flag.data(), finite ? count.data() : nullptr, p); });
    for (int r = 0; r < R; r++) { // k_mark_stubs
        const uint64_t len = goff[(size_t)r + 1] - goff[(size_t)r] - 1;
        if (len < (uint64_t)k) continue;
        flag[goff[(size_t)r]] |= 2;
        flag[goff[(size_t)r] + len - (uint64_t)k] |= 2;
    }
    std::vector<uint64_t> bif; // k_decide
    for (uint64_t s = 0; s < cap; s++)
        if (slot[s].key != kEmpty && is_bifurcation(slot[s].info) && (!finite || (unsigned long long)count[s] <= in.abundance)) bif.push_back(slot[s].key);
    const unsigned nb = (unsigned)bif.size();
    std::vector<unsigned> perm(nb);
    std::iota(perm.begin(), perm.end(), 0u);
    if constexpr (W == 1) {
        std::sort(bif.begin(), bif.end());
        for (unsigned i = 0; i < nb; i++) atomicOr(&slot[find(tb, bif[i])].info, (unsigned long long)(i + 1) << kIdShift); // k_assign_ids
    } else {
        std::vector<uint64_t> wordsv((size_t)W * nb);
        for (unsigned i = 0; i < nb; i++) canon_words_at<W>(t, bif.data(), nb, i, wordsv.data());
        for (int w = 0; w < W; w++) // what the LSD radix passes amount to: stable, least significant word first
            std::stable_sort(perm.begin(), perm.end(), [&](unsigned a, unsigned b) { return wordsv[(size_t)w * nb + a] < wordsv[(size_t)w * nb + b]; });
        parallel_positions(in.threads, 0, nb, [&](uint64_t i) { assign_id_at<W>(t, tb, bif[perm[i]], (unsigned)i); });
    }
    // ---- flagged positions in genome order, ids, final list (k_flag_*, k_emit_ids, k_final_*)
    std::vector<Junction> out;
    int32_t stub = (int32_t)(nb + 42);
    for (uint64_t p = 0; p < G; p++) {
        if (!flag[p]) continue;
        const int32_t id = id_at<W>(t, tb, flag[p], p);
        if (id == 0) continue;
        const int r = (int)(std::upper_bound(goff.begin(), goff.end(), p) - goff.begin()) - 1;
        out.push_back(Junction{(uint32_t)r, (uint32_t)(p - goff[(size_t)r]), id == kStub ? (int64_t)stub++ : (int64_t)id});
    }
    return out;
}

bool read_fasta(const char *path, std::vector<std::string> &records)
{
    std::ifstream in(path);
    if (!in) return false;
    std::string line;
    bool open = false;
    while (std::getline(in, line)) {
        if (!line.empty() && line[0] == '>') {
            records.emplace_back();
            open = true;
            continue;
        }
        if (!open) continue;
        for (char c : line)
            if (!isspace((unsigned char)c)) records.back().push_back(c); // case is folded by the packing, as on the device
    }
    return true;
}

} // namespace

int main(int argc, char **argv)
{
    if (argc < 6) {
        fprintf(stderr, "usage: graph_emu <k> <threads> <abundance|0> <out.dbg> <fasta>...\n");
        return 2;
    }
    Input in;
    in.k = atoi(argv[1]);
    in.threads = std::max(1, atoi(argv[2]));
    in.abundance = strtoull(argv[3], nullptr, 10);
    if (!in.abundance) in.abundance = UINT64_MAX;
    for (int i = 5; i < argc; i++)
        if (!read_fasta(argv[i], in.rec)) {
            fprintf(stderr, "cannot read %s\n", argv[i]);
            return 1;
        }
    if (in.k < 1 || in.k > kMaxK || in.k % 2 == 0) {
        fprintf(stderr, "k must be odd and at most %d\n", kMaxK);
        return 1;
    }
    std::vector<Junction> js;
    switch ((2 * in.k + 63) / 64) {
    case 1: js = run<1>(in); break;
    case 2: js = run<2>(in); break;
    case 3: js = run<3>(in); break;
    case 4: js = run<4>(in); break;
    case 5: js = run<5>(in); break;
    case 6: js = run<6>(in); break;
    case 7: js = run<7>(in); break;
    default: js = run<8>(in); break;
    }
    FILE *f = fopen(argv[4], "wb");
    if (!f) return 1;
    uint32_t now = 0;
    auto put = [f](uint32_t pos, int64_t id) {
        fwrite(&pos, 4, 1, f);
        fwrite(&id, 8, 1, f);
    };
    for (const Junction &j : js) {
        for (; j.chr > now; ++now) put(0xFFFFFFFFu, INT64_MAX);
        put(j.pos, j.id);
    }
    fclose(f);
    printf("%zu junction records\n", js.size());
    return 0;
}
