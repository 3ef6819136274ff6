/*
 * sched_sim.cpp -- EXPERIMENT (test infrastructure): CPU simulation of round schedulers for the speculative
 * fixpoint evaluation used by the CUDA path, with a simple device cost model, so that
 * scheduling policies can be compared without GPU time.  Built on the oracle's epoch-mode Process().
 *
 * Policies (all produce the sequential result; the program checks that against the plain window scheme):
 *   window : admit W seeds when the active set is empty (W doubles from w_init to w_max) -- what the device did
 *   rolling: admit `delta` new seeds every round on top of the unconverged ones, commit the clean prefix
 *   eager  : evaluate the commit-time re-run (slot 1) together with slot 0 whenever the round's list is short
 *
 * Cost model: one evaluation costs  c_base + c_push * pushes  cycles on one warp; a launch over a list lasts
 *   max(longest evaluation, sum / warps) ; a round = launch A (+ launch C when re-runs were found) + overhead.
 *
 * usage: sched_sim graph k b m a policy(window|rolling) w_init w_max eager(0|1) fasta...
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

struct SeedState {
    std::vector<Inst> r0, r1;
    Intervals R0, R1;
    char need0 = 1, need1 = 0, has1 = 0, conf = 0;
};

int main(int argc, char **argv)
{
    if (argc < 11) return 2;
    char err[512];
    int k = atoi(argv[2]), b = atoi(argv[3]), m = atoi(argv[4]), a = atoi(argv[5]);
    const bool rolling = !strcmp(argv[6], "rolling");
    int64_t w_init = atoll(argv[7]), w_max = atoll(argv[8]);
    const int eager = atoi(argv[9]);
    lcbo *L = lcbo_load(argv[1], argv + 10, argc - 10, k, a, err, sizeof err);
    if (!L) { fprintf(stderr, "%s\n", err); return 1; }
    int64_t S = lcbo_enumerate_seeds(L);
    lcbo_epoch_prepare(L, m, b, b, 8);
    const uint32_t INF = 0xFFFFFFFFu;
    std::vector<uint32_t> Ebase(L->N, INF), Enew, Ecur = Ebase;
    L->epoch = Ecur.data();
    std::vector<int64_t> log;
    L->readlog = &log;
    const double c_base = getenv("C_BASE") ? atof(getenv("C_BASE")) : 6000, c_push = getenv("C_PUSH") ? atof(getenv("C_PUSH")) : 12000;
    const double warps = 1776, hz = 1.965e9, overhead_ms = getenv("C_OVH") ? atof(getenv("C_OVH")) : 0.08;
    const int64_t eager_max = getenv("EAGER_MAX") ? atoll(getenv("EAGER_MAX")) : 4096;
    const int64_t cap = getenv("CAP") ? atoll(getenv("CAP")) : w_max;
    std::vector<SeedState> st(S);
    int64_t c0 = 0, c1 = 0, W = w_init, delta = w_init;
    double total_ms = 0;
    uint64_t rounds = 0, evals0 = 0, evals1 = 0, blocks_found = 0;
    FILE *fout = fopen(getenv("BLOCKS_OUT") ? getenv("BLOCKS_OUT") : "/tmp/sched_blocks.txt", "w");
    const bool verbose = getenv("VERBOSE") != nullptr;
    auto eval = [&](int64_t i, int slot, double &cost) {
        uint64_t p0 = L->ctr[7];
        L->thresh = slot == 0 ? (uint32_t)(i / 256 * 256) : (uint32_t)i;
        L->Process(L->seed[i], slot == 0 ? st[i].r0 : st[i].r1);
        Compress(log, slot == 0 ? st[i].R0 : st[i].R1);
        cost = c_base + c_push * (double)(L->ctr[7] - p0);
        (slot == 0 ? evals0 : evals1)++;
    };
    double prev_rate = 0;
    int hold = 0;
    int64_t window_start = 0;
    double window_ms = 0;
    while (c0 < S) {
        // ---- admission
        int64_t admit = 0;
        if (!rolling) {
            if (c0 == c1) admit = std::min(W, S - c1), window_start = c1, window_ms = 0;
        } else {
            admit = std::min(std::min(delta, S - c1), std::max<int64_t>(0, cap - (c1 - c0)));
        }
        c1 += admit;
        rounds++;
        // ---- launch A
        double sumA = 0, maxA = 0, sumC = 0, maxC = 0;
        int64_t nA = 0, nC = 0;
        std::vector<int64_t> fresh;
        int64_t n_list = 0;
        for (int64_t i = c0; i < c1; i++) n_list += st[i].need0 || st[i].need1;
        const bool do_eager = eager && n_list <= eager_max;
        for (int64_t i = c0; i < c1; i++) {
            SeedState &s = st[i];
            double c;
            if (s.need0) {
                eval(i, 0, c), sumA += c, maxA = std::max(maxA, c), nA++;
                s.need0 = 0, s.has1 = 0, s.need1 = 0;
                fresh.push_back(i);
                if (do_eager) { eval(i, 1, c), sumA += c, maxA = std::max(maxA, c), nA++, s.has1 = 1; }
            } else if (s.need1) {
                eval(i, 1, c), sumA += c, maxA = std::max(maxA, c), nA++;
                s.need1 = 0, s.has1 = 1;
            }
        }
        // ---- B + C
        for (int64_t i : fresh) {
            SeedState &s = st[i];
            bool c = false;
            if (s.r0.size() > 1)
                for (auto &in : s.r0) {
                    int64_t lo, hi; Marks(in, lo, hi);
                    for (int64_t f = lo; f <= hi && !c; f++) c = Ecur[f] < (uint32_t)i;
                }
            s.conf = c;
            if (c && !s.has1) {
                double cc;
                eval(i, 1, cc), sumC += cc, maxC = std::max(maxC, cc), nC++;
                s.has1 = 1;
            }
            if (!c) s.has1 = 0;
        }
        // ---- D: claims
        Enew = Ebase;
        for (int64_t i = c0; i < c1; i++) {
            const std::vector<Inst> &fin = st[i].conf ? st[i].r1 : st[i].r0;
            if (fin.size() > 1)
                for (auto &in : fin) {
                    int64_t lo, hi; Marks(in, lo, hi);
                    for (int64_t f = lo; f <= hi; f++) Enew[f] = std::min(Enew[f], (uint32_t)i);
                }
        }
        // ---- E: validation
        int64_t dirty = 0, fd = c1;
        for (int64_t i = c0; i < c1; i++) {
            SeedState &s = st[i];
            uint32_t T = (uint32_t)(i / 256 * 256);
            bool d0 = false;
            for (auto &iv : s.R0) {
                for (int64_t f = iv.first; f <= iv.second && !d0; f++) d0 = (Ecur[f] < T) != (Enew[f] < T);
                if (d0) break;
            }
            bool isdirty = false;
            if (d0) { s.need0 = 1; s.has1 = 0; isdirty = true; }
            else {
                bool c = false;
                if (s.r0.size() > 1)
                    for (auto &in : s.r0) {
                        int64_t lo, hi; Marks(in, lo, hi);
                        for (int64_t f = lo; f <= hi && !c; f++) c = Enew[f] < (uint32_t)i;
                    }
                if (c != (bool)s.conf) isdirty = true;
                bool rerun = false;
                if (c) {
                    if (!s.has1) rerun = true;
                    else {
                        for (auto &iv : s.R1) {
                            for (int64_t f = iv.first; f <= iv.second && !rerun; f++) rerun = (Ecur[f] < (uint32_t)i) != (Enew[f] < (uint32_t)i);
                            if (rerun) break;
                        }
                    }
                }
                s.conf = c;
                if (rerun) { s.need1 = 1; isdirty = true; }
                if (!c) s.has1 = 0;
            }
            if (isdirty) { dirty++; fd = std::min(fd, i); }
        }
        // ---- F: commit the clean prefix
        for (int64_t i = c0; i < fd; i++) {
            const std::vector<Inst> &fin = st[i].conf ? st[i].r1 : st[i].r0;
            if (fin.size() > 1) {
                int64_t cur = ++blocks_found;
                for (auto &in : fin) {
                    int64_t lo, hi; Marks(in, lo, hi);
                    for (int64_t f = lo; f <= hi; f++) Ebase[f] = std::min(Ebase[f], (uint32_t)i);
                    if (in.pos) fprintf(fout, "%lld %d %lld %lld\n", (long long)cur, in.chr, (long long)L->Position(in.fg, true), (long long)(L->Position(in.bg, true) + k));
                    else fprintf(fout, "%lld %d %lld %lld\n", -(long long)cur, in.chr, (long long)(L->Position(in.bg, false) - k), (long long)L->Position(in.fg, false));
                }
            }
            st[i] = SeedState(); // free memory
        }
        const int64_t committed = fd - c0;
        c0 = fd;
        Ecur.swap(Enew);
        L->epoch = Ecur.data();
        const double tA = nA ? std::max(maxA, sumA / warps) / hz * 1e3 : 0, tC = nC ? std::max(maxC, sumC / warps) / hz * 1e3 : 0;
        const double t_round = tA + tC + overhead_ms;
        total_ms += t_round;
        window_ms += t_round;
        if (verbose)
            fprintf(stderr, "round %llu active [%lld,%lld) admit %lld A: n %lld %.3f ms (max %.3f) C: n %lld %.3f ms dirty %lld committed %lld total %.2f\n",
                    (unsigned long long)rounds, (long long)c0, (long long)c1, (long long)admit, (long long)nA, tA, maxA / hz * 1e3, (long long)nC, tC,
                    (long long)dirty, (long long)committed, total_ms);
        if (!rolling) {
            if (c0 == c1) { // window converged: the device's adaptation rule (lcb_device.cu)
                const int64_t n = c1 - window_start;
                const double rate = n / std::max(window_ms, 1e-3);
                if (hold > 0) hold--;
                else if (n == W && prev_rate > 0 && rate < 0.7 * prev_rate && W > 256) W = std::max<int64_t>(256, W / 2 / 256 * 256), hold = 3;
                else if (n == W) W = std::min(w_max, W * 2);
                prev_rate = rate;
            }
        } else {
            // grow admission while the round is latency-bound, shrink when throughput-bound
            const double longest = std::max(maxA, maxC) / hz * 1e3;
            if (tA < 1.3 * longest || tA < 0.3) delta = std::min<int64_t>(w_max, delta * 2);
            else if (tA > 2.5 * longest) delta = std::max<int64_t>(256, delta / 2);
        }
    }
    fclose(fout);
    printf("policy %s eager %d w_init %lld w_max %lld cap %lld: rounds %llu evals0 %llu evals1 %llu blocks %llu  model time %.2f ms\n", argv[6], eager,
           (long long)w_init, (long long)w_max, (long long)cap, (unsigned long long)rounds, (unsigned long long)evals0, (unsigned long long)evals1,
           (unsigned long long)blocks_found, total_ms);
    lcbo_free(L);
    return 0;
}
