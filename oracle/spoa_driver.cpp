// spoa_driver.cpp -- TEST INFRASTRUCTURE (checker of the alignment stage).
//
// A small front end over the UNMODIFIED spoa library of the reference (spoa/src/{alignment_engine,graph,
// sisd_alignment_engine,simd_alignment_engine_dispatcher}.cpp, compiled where they lie by oracle/Makefile).  The
// reference's own front end (spoa/src/main.cpp) cannot be built here: it needs bioparser and biosoup, which the reference
// pulls through meson wraps / CMake FetchContent (spoa/subprojects/*.wrap) and which are absent from /root/reference.  This
// file restates exactly what main.cpp does for the one command line the pipeline uses,
//     spoa <block.fa> -l 1 -r 1 -e -8          (SibeliaZ-LCB/sibeliaz:66)
// i.e. AlignmentEngine::Create(kNW, m=5, n=-4, g=-8, e=-8, q=-10, c=-4) (main.cpp:209-214,258-260: linear gaps because
// g >= e), Prealloc(max_len, 4) (:270-278), then per sequence in file order Align + AddAlignment (:282-320), then
// GenerateMultipleSequenceAlignment(false) printed as ">name\nrow\n" (:330-338).
//
//   spoa-ref <fasta> [-m -n -g -e -q -c -l <v>] [-r 1]      one block, same stdout as `spoa`
//   spoa-ref --chunk <file.tmp> [options]                  every line of an LCB chunk file is one block
//       ("> hdr;start;len;strand;chrLen@SEQ@" per copy, blocksfinder.h:533-582); prints the MAF paragraphs the wrapper's
//       align()/block_align() append to <chunk>.msa (sibeliaz:64-100)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "spoa/spoa.hpp"

namespace {

struct Options {
    int m = 5, n = -4, g = -8, e = -6, q = -10, c = -4, algorithm = 0, result = 0;
};

// the rows spoa prints for one block (without the ">name" lines); empty on failure, as the wrapper sees it.
// One engine serves every block of a chunk: it keeps no state between Align calls besides its (growing) buffers.
std::vector<std::string> Msa(const std::vector<std::string> &seq, const Options &o)
{
    static std::unique_ptr<spoa::AlignmentEngine> engine;
    if (!engine) engine = spoa::AlignmentEngine::Create(static_cast<spoa::AlignmentType>(o.algorithm), o.m, o.n, o.g, o.e, o.q, o.c);
    size_t max_len = 0;
    for (const auto &s : seq) max_len = std::max(max_len, s.size());
    engine->Prealloc(max_len, 4);
    spoa::Graph graph{};
    for (const auto &s : seq) {
        std::int32_t score = 0;
        spoa::Alignment alignment = engine->Align(s, graph, &score);
        graph.AddAlignment(alignment, s);
    }
    return graph.GenerateMultipleSequenceAlignment(false);
}

// FASTA as bioparser reads it: name = header up to the first blank, sequence lines concatenated
void ReadFasta(std::istream &in, std::vector<std::string> &name, std::vector<std::string> &seq)
{
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '>') {
            size_t e = 1;
            while (e < line.size() && line[e] != ' ' && line[e] != '\t') e++;
            name.push_back(line.substr(1, e - 1));
            seq.emplace_back();
        } else if (!seq.empty()) {
            seq.back() += line;
        }
    }
}

} // namespace

int main(int argc, char **argv)
{
    Options o;
    std::string file, chunk;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() { return i + 1 < argc ? atoi(argv[++i]) : 0; };
        if (a == "-m") o.m = val();
        else if (a == "-n") o.n = val();
        else if (a == "-g") o.g = val();
        else if (a == "-e") o.e = val();
        else if (a == "-q") o.q = val();
        else if (a == "-c") o.c = val();
        else if (a == "-l") o.algorithm = val();
        else if (a == "-r") o.result = val();
        else if (a == "--chunk" && i + 1 < argc) chunk = argv[++i];
        else file = a;
    }
    try {
        if (!chunk.empty()) {
            std::ifstream in(chunk);
            if (!in) {
                fprintf(stderr, "cannot open %s\n", chunk.c_str());
                return 1;
            }
            std::string line;
            while (std::getline(in, line)) { // one block per line; '@' stands for a newline (sibeliaz:89)
                std::vector<std::string> header, seq;
                size_t p = 0;
                while (p < line.size()) {
                    size_t q = line.find('@', p);
                    if (q == std::string::npos) q = line.size();
                    std::string tok = line.substr(p, q - p);
                    p = q + 1;
                    if (tok.empty()) continue;
                    if (tok[0] == '>') {
                        // `cut -d' ' -f2-` of the header line, then ';' -> ' ', then "s " in front (sibeliaz:79)
                        size_t sp = tok.find(' ');
                        std::string h = sp == std::string::npos ? tok : tok.substr(sp + 1);
                        for (char &ch : h)
                            if (ch == ';') ch = ' ';
                        header.push_back("s " + h);
                        seq.emplace_back();
                    } else if (!seq.empty()) {
                        seq.back() += tok;
                    }
                }
                if (seq.empty()) continue;
                std::vector<std::string> rows;
                try {
                    rows = Msa(seq, o);
                } catch (std::exception &) { // the wrapper drops a block whose spoa run printed nothing (sibeliaz:68-72)
                    continue;
                }
                if (rows.empty()) continue;
                std::cout << "\na\n";
                for (size_t i = 0; i < rows.size() && i < header.size(); i++) std::cout << header[i] << ' ' << rows[i] << "\n";
            }
            return 0;
        }
        if (file.empty()) {
            fprintf(stderr, "[spoa::] error: missing input file!\n");
            return 1;
        }
        std::ifstream in(file);
        if (!in) {
            fprintf(stderr, "cannot open %s\n", file.c_str());
            return 1;
        }
        std::vector<std::string> name, seq;
        ReadFasta(in, name, seq);
        if (o.result != 1) {
            fprintf(stderr, "spoa-ref: only -r 1 (multiple sequence alignment) is restated\n");
            return 1;
        }
        std::vector<std::string> rows = Msa(seq, o);
        for (size_t i = 0; i < rows.size(); i++) std::cout << ">" << name[i] << "\n" << rows[i] << "\n";
    } catch (std::exception &ex) {
        std::cerr << ex.what() << std::endl;
        return 1;
    }
    return 0;
}
