// poa_warp_emu.cpp -- TEST INFRASTRUCTURE: runs the WARP-LEVEL code of the product's POA core (the row kernel with its
// shuffles and max-scans, sibeliaz_b200/csrc/poa_core.cuh under POA_WARP_EMULATION) on the CPU: 32 host threads are the
// lanes of one warp and meet in a barrier at every __shfl_*_sync / __syncwarp, which is the lockstep the hardware gives.
// A missing synchronisation in the device code shows up here as a data race between free-running threads (wrong output
// or a failing run), an index slip in the scan as a wrong MSA.  Not a product path.
//   poa_warp_emu --chunk <file.tmp>     MAF paragraphs, as oracle/poa_oracle prints them
#include <barrier>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

namespace emu {
std::barrier<> *bar;
int32_t slot[32];
thread_local int lane;
} // namespace emu

inline int32_t __shfl_up_sync(unsigned, int32_t v, int d)
{
    emu::slot[emu::lane] = v;
    emu::bar->arrive_and_wait();
    const int32_t r = emu::lane >= d ? emu::slot[emu::lane - d] : v;
    emu::bar->arrive_and_wait();
    return r;
}
inline int32_t __shfl_sync(unsigned, int32_t v, int src)
{
    emu::slot[emu::lane] = v;
    emu::bar->arrive_and_wait();
    const int32_t r = emu::slot[src & 31];
    emu::bar->arrive_and_wait();
    return r;
}
inline void __syncwarp() { emu::bar->arrive_and_wait(); }

#define POA_WARP_EMULATION 1
#include "../sibeliaz_b200/csrc/poa_core.cuh"

int main(int argc, char **argv)
{
    std::string chunk;
    poa::Params pr{5, -4, -8};
    for (int i = 1; i < argc; i++)
        if (std::string(argv[i]) == "--chunk" && i + 1 < argc) chunk = argv[++i];
    std::ifstream in(chunk);
    if (chunk.empty() || !in) return 1;
    std::barrier<> bar(32);
    emu::bar = &bar;
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
        poa::Caps caps = poa::poa_caps_for(sum, mx, 2);
        std::vector<uint8_t> arena(poa::poa_arena_bytes(caps, copies) + 64, 0xCD);
        poa::poa_bind(w, arena.data(), caps, copies);
        std::vector<std::string> rows(copies);
        std::vector<std::thread> lanes;
        for (int l = 0; l < 32; l++)
            lanes.emplace_back([&, l]() {
                emu::lane = l;
                poa::run_block(w, pr, seq.data(), off.data(), 0, copies, l, 32);
                if (w.err) return; // uniform: every lane sees the same flag after the last barrier
                for (uint32_t k = 0; k < copies; k++) {
                    if (l == 0) rows[k].assign(w.n_columns, '?');
                    __syncwarp();
                    poa::write_row(w, k, (uint8_t *)&rows[k][0], l, 32);
                }
            });
        for (auto &t : lanes) t.join();
        if (w.err) {
            fprintf(stderr, "poa core failed: err %d\n", w.err);
            return 2;
        }
        std::cout << "\na\n";
        for (uint32_t k = 0; k < copies; k++) std::cout << header[k] << ' ' << rows[k] << "\n";
    }
    return 0;
}
