// poa_core_host.cpp -- TEST INFRASTRUCTURE: the product's POA core (sibeliaz_b200/csrc/poa_core.cuh) compiled for the HOST,
// where a "warp" is one thread, so that the very code the kernel runs (graph update, topological sort, traceback, MSA; the
// row recurrence in its scalar form) is checked against the oracle without a GPU.  Not a product path: nothing links this.
//   poa_core_host --chunk <file.tmp> [--level 0|1|2]    MAF paragraphs, as oracle/poa_oracle prints them
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../sibeliaz_b200/csrc/poa_core.cuh"

int main(int argc, char **argv)
{
    std::string chunk;
    int level = 0;
    poa::Params pr{5, -4, -8};
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--chunk" && i + 1 < argc) chunk = argv[++i];
        else if (a == "--level" && i + 1 < argc) level = atoi(argv[++i]);
    }
    std::ifstream in(chunk);
    if (chunk.empty() || !in) return 1;
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
        poa::Work w;
        std::vector<uint8_t> arena;
        for (int lv = level;; lv++) { // the retry ladder of the device driver
            poa::Caps caps = poa::poa_caps_for(sum, mx, lv);
            arena.assign(poa::poa_arena_bytes(caps, copies) + 64, 0xCD);
            poa::poa_bind(w, arena.data(), caps, copies);
            poa::run_block(w, pr, seq.data(), off.data(), 0, copies, 0, 1);
            if (w.err != 1 || lv >= 2) break;
            fprintf(stderr, "block retried at level %d\n", lv + 1);
        }
        if (w.err) {
            fprintf(stderr, "poa core failed: err %d\n", w.err);
            return 2;
        }
        std::cout << "\na\n";
        std::string row(w.n_columns, '?');
        for (uint32_t k = 0; k < copies; k++) {
            poa::write_row(w, k, (uint8_t *)&row[0], 0, 1);
            std::cout << header[k] << ' ' << row << "\n";
        }
    }
    return 0;
}
