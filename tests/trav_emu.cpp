// trav_emu.cpp -- TEST INFRASTRUCTURE: ProcessVertex::Process exactly as the traversal kernel runs it
// (sibeliaz_b200/csrc/lcb_traverse.cuh: process_seed with mpv_fast / mpv_mid / the general vote, push_parallel / push_group,
// path_score, the shadow state, ...) executed on the CPU by 32 host threads in lockstep (tests/cuda_emu.h), one evaluation
// after the other against a given epoch array, so that the device code is checked against the oracle without a GPU.
//   trav_emu <input.bin> <output.txt> [--lean]
// --lean: the common-case body (lcb_lean.cuh) first, the general code only for the evaluations it hands back (as the two
// traversal kernels do on the device); prints how many were handed back.
// input.bin (little endian): int64 N, V, C, S; int32 k, b, m, flank, depth; int4 rec[N]; int2 occ[N]; uint32 vtx_off[V+2];
//   uint32 chr_off[C+1]; uint32 E[N+32]; then S x {int32 vid; uint32 ch; uint32 thresh}
// output.txt: per evaluation one line "n  fg|pos<<62 bg  ..." in bestInstance order (the format of lcbo_epoch_process)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "cuda_emu.h"
#define LCB_ERR_CAPACITY 5
#include "../sibeliaz_b200/csrc/lcb_traverse.cuh"
#include "../sibeliaz_b200/csrc/lcb_lean.cuh"

template <class T>
static void rd(FILE *f, T *p, size_t n)
{
    if (n && fread(p, sizeof(T), n, f) != n) {
        fprintf(stderr, "short read\n");
        exit(2);
    }
}

int main(int argc, char **argv)
{
    if (argc < 3) return 1;
    const bool use_lean = argc > 3 && std::string(argv[3]) == "--lean";
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 1;
    int64_t hdr[4];
    int32_t prm[5];
    rd(f, hdr, 4), rd(f, prm, 5);
    const int64_t N = hdr[0], V = hdr[1], C = hdr[2], S = hdr[3];
    std::vector<int4> rec((size_t)N);
    std::vector<int2> occ((size_t)N);
    std::vector<uint32_t> vtx_off((size_t)V + 2), chr_off((size_t)C + 1), E((size_t)N + 32);
    rd(f, rec.data(), rec.size()), rd(f, occ.data(), occ.size()), rd(f, vtx_off.data(), vtx_off.size());
    rd(f, chr_off.data(), chr_off.size()), rd(f, E.data(), E.size());
    struct Job {
        int32_t vid;
        uint32_t ch, thresh;
    };
    std::vector<Job> jobs((size_t)S);
    rd(f, jobs.data(), jobs.size());
    fclose(f);

    lcb::Index ix;
    ix.rec = rec.data(), ix.occ = occ.data(), ix.vtx_off = vtx_off.data(), ix.chr_off = chr_off.data();
    ix.C = (int)C, ix.N = (int)N, ix.V = (int)V;
    lcb::Params pr{prm[0], prm[1], prm[2], prm[3], prm[4]};
    auto sm = std::make_unique<lcb::WarpSmem>(); // the warp's shared memory
    auto lsm = std::make_unique<lcb::lean::LeanSmem>();
    memset(lsm.get(), 0, sizeof(lcb::lean::LeanSmem)); // k_traverse_lean zeroes the hash and the vote table once per warp
    std::vector<int2> lean_rs((size_t)lcb::kReadSetMax);
    std::vector<lcb::lean::LInst> lean_shadow((size_t)lcb::lean::kLInst);
    std::vector<int2> lean_hash2((size_t)lcb::lean::kLHash2, int2{0, 0});
    std::vector<unsigned short> lean_hslot2((size_t)lcb::lean::kLPath2);
    long long bails = 0;
    std::vector<unsigned char> arena(lcb::arena_stride_of(false) + 256, 0); // spill arena: the hash part must start all-empty
    unsigned char *abase = (unsigned char *)(((uintptr_t)arena.data() + 255) & ~(uintptr_t)255);
    std::barrier<> wb(32);
    emu::warp_bar[0] = &wb;
    std::vector<std::vector<long long>> results((size_t)S);
    std::vector<int> errs((size_t)S, 0);
    std::vector<std::thread> lanes;
    for (int l = 0; l < 32; l++)
        lanes.emplace_back([&, l]() {
            emu::lane = l, emu::warp = 0;
            lcb::Ctx c; // per-lane registers on the device
            c.ix = ix, c.pr = pr, c.E = E.data(), c.lane = l, c.sm = sm.get();
            c.err = 0, c.collect = false;
            c.ct.walk = c.ct.occ = c.ct.scan = c.ct.score = 0;
            c.ct.pushes = c.ct.mpv_fast = c.ct.mpv_mid = c.ct.mpv_slow = c.ct.push_par = c.ct.push_ser = 0;
            c.vote_clean = false;
            lcb::arena_bind(c, abase, false);
            lcb::lean::LCtx lc;
            lc.rec = ix.rec, lc.occ = ix.occ, lc.vtx_off = ix.vtx_off, lc.E = E.data(), lc.chr_off_s = ix.chr_off, lc.C = ix.C;
            lc.b = pr.b, lc.m = pr.m, lc.flank = pr.flank, lc.depth = pr.depth;
            lc.lane = l, lc.sm = lsm.get(), lc.rs = lean_rs.data(), lc.rs_cap = (int)lean_rs.size();
            lc.last_clo = lc.last_chi = 0, lc.why = 0, lc.deep_bias = 0;
            lc.shadow = lean_shadow.data();
            lc.hash2 = lean_hash2.data(), lc.hslot2 = lean_hslot2.data();
            for (int64_t s = 0; s < S; s++) {
                c.thresh = jobs[(size_t)s].thresh;
                c.err = 0;
                if (use_lean) {
                    lc.thresh = jobs[(size_t)s].thresh;
                    const int r = lcb::lean::process_seed(lc, jobs[(size_t)s].vid, (unsigned char)jobs[(size_t)s].ch);
                    __syncwarp();
                    if (r == lcb::lean::kOk) {
                        if (l == 0) {
                            errs[(size_t)s] = 0;
                            for (int t = 0; t < lc.nbest; t++) {
                                const int4 b = lsm->best[t];
                                results[(size_t)s].push_back((long long)(b.x & 0x7FFFFFFF) | (b.x < 0 ? 1ll << 62 : 0));
                                results[(size_t)s].push_back(b.y);
                            }
                        }
                        __syncwarp();
                        continue;
                    }
                    if (l == 0) bails++;
                }
                lcb::process_seed(c, jobs[(size_t)s].vid, (unsigned char)jobs[(size_t)s].ch);
                __syncwarp();
                if (l == 0) {
                    errs[(size_t)s] = c.err;
                    if (!c.err)
                        for (int t = 0; t < c.nbest; t++) {
                            const int4 b = c.best[t];
                            results[(size_t)s].push_back((long long)(b.x & 0x7FFFFFFF) | (b.x < 0 ? 1ll << 62 : 0));
                            results[(size_t)s].push_back(b.y);
                        }
                }
                __syncwarp();
                if (c.err) { // leave the arena clean for the next evaluation (k_traverse does the same before a big-slot re-run)
                    lcb::hash_clear(c);
                    __syncwarp();
                }
            }
        });
    for (auto &t : lanes) t.join();
    if (use_lean) fprintf(stderr, "lean: %lld of %lld evaluations handed back to the general code\n", bails, (long long)S);
    FILE *o = fopen(argv[2], "w");
    for (int64_t s = 0; s < S; s++) {
        if (errs[(size_t)s]) {
            fprintf(o, "err %d\n", errs[(size_t)s]);
            continue;
        }
        fprintf(o, "%zu", results[(size_t)s].size() / 2);
        for (long long v : results[(size_t)s]) fprintf(o, " %lld", v);
        fprintf(o, "\n");
    }
    fclose(o);
    return 0;
}
