/*
 * jacobi_proto.cpp -- EXPERIMENT (test infrastructure): CPU prototype of the speculative
 * round-based ("Jacobi") evaluation that the CUDA path uses, built on the oracle's Process().
 * It answers: how many rounds does a window of W seeds need, and how much extra traversal work
 * does speculation cost, while producing exactly the sequential result?
 *
 *   final_i(E) = r0 = Process(i, used := E < 256*floor(i/256));
 *                if |r0| > 1 and marks(r0) hit E < i:  r1 = Process(i, used := E < i); r1 if |r1|>1 else {}
 *   E'[f]      = min { i : f in marks(final_i) }                      (fixpoint == sequential result)
 *
 * usage: jacobi_proto graph k b m a W fasta...
 */
#define LCBO_EPOCH
#include "../lcb_oracle.cpp"
#include <chrono>

typedef std::vector<std::pair<int64_t, int64_t>> Intervals;

static void Compress(std::vector<int64_t> &log, Intervals &out)
{
    out.clear();
    std::sort(log.begin(), log.end());
    for (size_t i = 0; i < log.size();) {
        size_t j = i;
        while (j + 1 < log.size() && log[j + 1] <= log[j] + 1) j++;
        out.push_back({log[i], log[j]});
        i = j + 1;
    }
    log.clear();
}

static void Marks(const Inst &a, int64_t &lo, int64_t &hi) { lo = std::min(a.fg, a.bg); hi = std::max(a.fg, a.bg) - 1; }

int main(int argc, char **argv)
{
    if (argc < 8) return 2;
    char err[512];
    int k = atoi(argv[2]), b = atoi(argv[3]), m = atoi(argv[4]), a = atoi(argv[5]);
    int64_t W = atoll(argv[6]);
    lcbo *L = lcbo_load(argv[1], argv + 7, argc - 7, k, a, err, sizeof err);
    if (!L) { fprintf(stderr, "%s\n", err); return 1; }
    int64_t S = lcbo_enumerate_seeds(L);
    L->min_block = m; L->max_branch = b; L->max_flank = b; L->looking_depth = 8;
    L->distance.assign((size_t)L->V * 2 + 2, INT_MAX);
    L->count.assign((size_t)L->V * 2 + 2, 0);
    L->order.assign(L->C, std::vector<int>());
    const uint32_t INF = 0xFFFFFFFFu;
    std::vector<uint32_t> Ebase(L->N, INF), Enew;
    std::vector<uint32_t> Ecur = Ebase;
    L->epoch = Ecur.data();
    std::vector<int64_t> log;
    L->readlog = &log;
    uint64_t runs0 = 0, runs1 = 0, total_rounds = 0, windows = 0, max_rounds = 0, rs_intervals = 0, rs_len = 0, val_reads = 0;
    std::vector<Block> blocks;
    int64_t blocks_found = 0;
    for (int64_t w0 = 0; w0 < S; w0 += W) {
        int64_t w1 = std::min(S, w0 + W), n = w1 - w0;
        std::vector<std::vector<Inst>> r0(n), r1(n);
        std::vector<Intervals> R0(n), R1(n);
        std::vector<char> need0(n, 1), need1(n, 0), has1(n, 0), conf(n, 0);
        Ecur = Ebase; // E_cur
        L->epoch = Ecur.data();
        int rounds = 0;
        while (true) {
            rounds++;
            for (int64_t j = 0; j < n; j++) {
                int64_t i = w0 + j;
                if (need0[j]) {
                    L->thresh = (uint32_t)(i / 256 * 256);
                    L->Process(L->seed[i], r0[j]);
                    Compress(log, R0[j]);
                    need0[j] = 0;
                    runs0++;
                }
                bool c = false;
                if (r0[j].size() > 1)
                    for (auto &in : r0[j]) {
                        int64_t lo, hi; Marks(in, lo, hi);
                        for (int64_t f = lo; f <= hi && !c; f++) c = Ecur[f] < (uint32_t)i;
                    }
                conf[j] = c;
                if (c && (!has1[j] || need1[j])) {
                    L->thresh = (uint32_t)i;
                    L->Process(L->seed[i], r1[j]);
                    Compress(log, R1[j]);
                    has1[j] = 1; need1[j] = 0;
                    runs1++;
                }
                if (!c) { has1[j] = 0; need1[j] = 0; }
            }
            Enew = Ebase;
            for (int64_t j = 0; j < n; j++) {
                const std::vector<Inst> &fin = conf[j] ? r1[j] : r0[j];
                if (fin.size() > 1)
                    for (auto &in : fin) {
                        int64_t lo, hi; Marks(in, lo, hi);
                        for (int64_t f = lo; f <= hi; f++) Enew[f] = std::min(Enew[f], (uint32_t)(w0 + j));
                    }
            }
            int64_t dirty = 0;
            for (int64_t j = 0; j < n; j++) {
                uint32_t i = (uint32_t)(w0 + j), T = i / 256 * 256;
                bool d0 = false;
                for (auto &iv : R0[j]) {
                    for (int64_t f = iv.first; f <= iv.second && !d0; f++) { val_reads++; d0 = (Ecur[f] < T) != (Enew[f] < T); }
                    if (d0) break;
                }
                if (d0) { need0[j] = 1; has1[j] = 0; dirty++; continue; }
                bool c = false;
                if (r0[j].size() > 1)
                    for (auto &in : r0[j]) {
                        int64_t lo, hi; Marks(in, lo, hi);
                        for (int64_t f = lo; f <= hi && !c; f++) c = Enew[f] < i;
                    }
                if (c != (bool)conf[j]) dirty++;
                if (c && has1[j]) {
                    bool d1 = false;
                    for (auto &iv : R1[j]) {
                        for (int64_t f = iv.first; f <= iv.second && !d1; f++) { val_reads++; d1 = (Ecur[f] < i) != (Enew[f] < i); }
                        if (d1) break;
                    }
                    if (d1) { need1[j] = 1; dirty++; }
                }
            }
            if (rounds <= 12 || dirty == 0) fprintf(stderr, "  window %lld round %d dirty %lld\n", (long long)(w0 / W), rounds, (long long)dirty);
            Ecur.swap(Enew);
            L->epoch = Ecur.data();
            if (!dirty) break;
        }
        total_rounds += rounds; windows++; max_rounds = std::max<uint64_t>(max_rounds, rounds);
        Ebase = Ecur;
        for (int64_t j = 0; j < n; j++) {
            for (auto &iv : R0[j]) { rs_intervals++; rs_len += iv.second - iv.first + 1; }
            const std::vector<Inst> &fin = conf[j] ? r1[j] : r0[j];
            if (fin.size() > 1) {
                int64_t cur = ++blocks_found;
                for (auto &in : fin) {
                    if (in.pos) blocks.push_back(Block{(int)cur, (size_t)L->Position(in.fg, true), (size_t)(L->Position(in.bg, true) + k), (size_t)in.chr});
                    else blocks.push_back(Block{(int)-cur, (size_t)(L->Position(in.bg, false) - k), (size_t)L->Position(in.fg, false), (size_t)in.chr});
                }
            }
        }
    }
    printf("S %lld W %lld windows %llu rounds total %llu max %llu | runs r0 %llu r1 %llu | R0 intervals %llu len %llu | validation reads %llu | T_walk %llu T_occ %llu T_scan %llu\n",
           (long long)S, (long long)W, (unsigned long long)windows, (unsigned long long)total_rounds, (unsigned long long)max_rounds,
           (unsigned long long)runs0, (unsigned long long)runs1, (unsigned long long)rs_intervals, (unsigned long long)rs_len,
           (unsigned long long)val_reads, (unsigned long long)L->ctr[0], (unsigned long long)L->ctr[1], (unsigned long long)L->ctr[2]);
    // compare with the sequential oracle
    L->readlog = nullptr;
    lcbo *Q = lcbo_load(argv[1], argv + 7, argc - 7, k, a, err, sizeof err);
    // the sequential path needs the non-epoch build; instead emulate: epoch INF + marks via thresholds is not available,
    // so just dump blocks for an external diff.
    FILE *f = fopen("/tmp/w/jacobi_blocks.txt", "w");
    for (auto &bk : blocks) fprintf(f, "%d %zu %zu %zu\n", bk.id, bk.chr, bk.start, bk.end);
    fclose(f);
    lcbo_free(Q);
    lcbo_free(L);
    return 0;
}
