// poa_warp_emu.cpp -- TEST INFRASTRUCTURE: runs the WARP-LEVEL code of the product's POA core (the row kernel with its
// shuffles and max-scans, sibeliaz_b200/csrc/poa_core.cuh under POA_WARP_EMULATION) on the CPU: 32 host threads are the
// lanes of one warp and meet in a barrier at every __shfl_*_sync / __syncwarp, which is the lockstep the hardware gives.
// A missing synchronisation in the device code shows up here as a data race between free-running threads (wrong output
// or a failing run), an index slip in the scan as a wrong MSA.  Not a product path.
//   poa_warp_emu --chunk <file.tmp> [--cta <warps>]     MAF paragraphs, as oracle/poa_oracle prints them
// --cta: the one-block-per-CTA variant for long blocks (run_block_cta: <warps> x 32 threads share every row)
#include <barrier>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace emu {
constexpr int kMaxWarps = 32;
std::barrier<> *warp_bar[kMaxWarps]; // one per warp: __shfl_*_sync, __syncwarp
std::barrier<> *cta_bar;             // __syncthreads
int32_t slot[kMaxWarps][32];
thread_local int lane, warp;
} // namespace emu

inline int32_t __shfl_up_sync(unsigned, int32_t v, int d)
{
    emu::slot[emu::warp][emu::lane] = v;
    emu::warp_bar[emu::warp]->arrive_and_wait();
    const int32_t r = emu::lane >= d ? emu::slot[emu::warp][emu::lane - d] : v;
    emu::warp_bar[emu::warp]->arrive_and_wait();
    return r;
}
inline int32_t __shfl_sync(unsigned, int32_t v, int src)
{
    emu::slot[emu::warp][emu::lane] = v;
    emu::warp_bar[emu::warp]->arrive_and_wait();
    const int32_t r = emu::slot[emu::warp][src & 31];
    emu::warp_bar[emu::warp]->arrive_and_wait();
    return r;
}
inline void __syncwarp() { emu::warp_bar[emu::warp]->arrive_and_wait(); }
inline void __syncthreads() { emu::cta_bar->arrive_and_wait(); }

#define POA_WARP_EMULATION 1
#include "../sibeliaz_b200/csrc/poa_core.cuh"

int main(int argc, char **argv)
{
    std::string chunk;
    poa::Params pr{5, -4, -8};
    int cta_warps = 0, first_level = 2; // --level 0|1: start at an optimistic arena level and climb the driver's retry ladder
    for (int i = 1; i < argc; i++) {
        if (std::string(argv[i]) == "--chunk" && i + 1 < argc) chunk = argv[++i];
        else if (std::string(argv[i]) == "--cta" && i + 1 < argc) cta_warps = atoi(argv[++i]);
        else if (std::string(argv[i]) == "--level" && i + 1 < argc) first_level = atoi(argv[++i]);
    }
    std::ifstream in(chunk);
    if (chunk.empty() || !in || cta_warps < 0 || cta_warps > emu::kMaxWarps) return 1;
    const int nwarps = cta_warps ? cta_warps : 1, nthreads = nwarps * 32;
    std::vector<std::unique_ptr<std::barrier<>>> wb;
    for (int v = 0; v < nwarps; v++) wb.emplace_back(new std::barrier<>(32)), emu::warp_bar[v] = wb.back().get();
    std::barrier<> cb(nthreads);
    emu::cta_bar = &cb;
    int32_t seg[64];
    std::string line;
    while (std::getline(in, line)) {
        std::vector<std::string> header;
        std::vector<uint8_t> seq;
        std::vector<uint64_t> off{0};
        size_t p = 0;
        bool open = false;
        while (p < line.size()) {
            size_t q = line.find('@', p);
            if (q == std::string::npos) q = line.size();
            std::string tok = line.substr(p, q - p);
            p = q + 1;
            if (tok.empty()) continue;
            if (tok[0] == '>') {
                if (open) off.push_back(seq.size());
                size_t sp = tok.find(' ');
                std::string h = sp == std::string::npos ? tok : tok.substr(sp + 1);
                for (char &ch : h)
                    if (ch == ';') ch = ' ';
                header.push_back("s " + h);
                open = true;
            } else if (open) {
                seq.insert(seq.end(), tok.begin(), tok.end());
            }
        }
        if (open) off.push_back(seq.size());
        const uint32_t copies = (uint32_t)off.size() - 1;
        if (!copies) continue;
        uint64_t sum = seq.size(), mx = 0;
        for (uint32_t c = 0; c < copies; c++) mx = std::max<uint64_t>(mx, off[c + 1] - off[c]);
        poa::Work w; // shared by the 32 lanes, like the kernel's shared-memory copy
        std::vector<std::string> rows(copies);
        for (int level = first_level;; level++) { // a block that outgrows its arena reports err = 1 and runs again one level up
        poa::Caps caps = poa::poa_caps_for(sum, mx, level);
        std::vector<uint8_t> arena(poa::poa_arena_bytes(caps, copies) + 64, 0xCD);
        poa::poa_bind(w, arena.data(), caps, copies);
        std::vector<std::thread> lanes;
        for (int t = 0; t < nthreads; t++)
            lanes.emplace_back([&, t]() {
                emu::lane = t & 31, emu::warp = t >> 5;
                if (cta_warps) {
                    poa::run_block_cta(w, pr, seq.data(), off.data(), 0, copies, t, nthreads, seg);
                    if (w.err) return;
                    for (uint32_t k = 0; k < copies; k++) {
                        if (t == 0) rows[k].assign(w.n_columns, '?');
                        __syncthreads();
                        poa::write_row_cta(w, k, (uint8_t *)&rows[k][0], t, nthreads);
                    }
                    return;
                }
                poa::run_block(w, pr, seq.data(), off.data(), 0, copies, t, 32);
                if (w.err) return; // uniform: every lane sees the same flag after the last barrier
                for (uint32_t k = 0; k < copies; k++) {
                    if (t == 0) rows[k].assign(w.n_columns, '?');
                    __syncwarp();
                    poa::write_row(w, k, (uint8_t *)&rows[k][0], t, 32);
                }
            });
        for (auto &t : lanes) t.join();
        if (w.err == 1 && level < 2) {
            fprintf(stderr, "block retried at level %d\n", level + 1);
            continue;
        }
        if (w.err) {
            fprintf(stderr, "poa core failed: err %d\n", w.err);
            return 2;
        }
        break;
        }
        std::cout << "\na\n";
        for (uint32_t k = 0; k < copies; k++) std::cout << header[k] << ' ' << rows[k] << "\n";
    }
    return 0;
}
