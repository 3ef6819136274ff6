// poa_oracle.cpp -- TEST INFRASTRUCTURE: CPU restatement of the alignment stage after the LCB path (SURVEY.md
// section 8f row 3): per block, `spoa <block.fa> -l 1 -r 1 -e -8` (SibeliaZ-LCB/sibeliaz:66) -- partial order alignment
// with global (Needleman-Wunsch) alignment of every copy against the growing graph, linear gaps, MSA output.
//
// Restated from the reference's spoa (spoa/src/graph.cpp, spoa/src/sisd_alignment_engine.cpp), every function citing the
// lines it follows, in flat arrays (the layout a device kernel will use): no pointers, nodes / edges / aligned sets by index.
// Pinned: byte-identical output to oracle/_ref/spoa-ref (the unmodified reference library) on every block of the examples
// (tests/test_alignment_oracle.py), which in turn reproduces the shipped golden alignment.maf.
//
//   poa_oracle --chunk <file.tmp> [-m 5 -n -4 -g -8]     MAF paragraphs of every block of an LCB chunk file
//
// Only what the pipeline's command line reaches is restated: type kNW, subtype kLinear (g >= e: main.cpp:209-214 +
// alignment_engine.cpp:57-63), unit weights (FASTA input), result mode 1 (MSA without consensus).
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

namespace {

constexpr int32_t kNegativeInfinity = INT32_MIN + 1024; // sisd_alignment_engine.cpp:13-14

struct Graph { // spoa::Graph (graph.hpp), index-based
    int num_codes = 0;
    int coder[256];
    char decoder[256];
    std::vector<int> code;                      // Node::code
    std::vector<std::vector<int>> in_tail;      // Node::inedges -> Edge::tail, insertion order (the DP's precedence)
    std::vector<std::vector<int>> out_head;     // Node::outedges -> Edge::head
    std::vector<std::vector<int>> aligned;      // Node::aligned_nodes, insertion order
    std::vector<std::vector<int>> seq_path;     // sequences_[s] followed through Successor(s): the nodes of copy s in order
    std::vector<int> rank_to_node;
    Graph() { std::fill(coder, coder + 256, -1), std::fill(decoder, decoder + 256, (char)-1); }

    int AddNode(int c) // graph.cpp:78-81
    {
        code.push_back(c);
        in_tail.emplace_back(), out_head.emplace_back(), aligned.emplace_back();
        return (int)code.size() - 1;
    }
    void AddEdge(int tail, int head) // graph.cpp:83-93 (labels and weights only matter to the consensus, not to the MSA)
    {
        for (int h : out_head[tail])
            if (h == head) return;
        out_head[tail].push_back(head);
        in_tail[head].push_back(tail);
    }
    // graph.cpp:95-112: a chain of new nodes for sequence[begin, end); returns its first node or -1
    int AddSequence(const std::string &s, uint32_t begin, uint32_t end, std::vector<int> &path)
    {
        if (begin == end) return -1;
        int prev = -1, first = -1;
        for (uint32_t i = begin; i < end; i++) {
            int curr = AddNode(coder[(unsigned char)s[i]]);
            if (first < 0) first = curr;
            if (prev >= 0) AddEdge(prev, curr);
            prev = curr;
            path.push_back(curr);
        }
        return first;
    }

    // graph.cpp:156-246
    void AddAlignment(const std::vector<std::pair<int32_t, int32_t>> &alignment, const std::string &s)
    {
        const uint32_t len = (uint32_t)s.size();
        if (len == 0) return;
        for (uint32_t i = 0; i < len; i++) {
            unsigned char ch = (unsigned char)s[i];
            if (coder[ch] == -1) coder[ch] = num_codes, decoder[num_codes++] = (char)ch;
        }
        std::vector<int> path;
        if (alignment.empty()) {
            AddSequence(s, 0, len, path);
            seq_path.push_back(path);
            TopologicalSort();
            return;
        }
        std::vector<uint32_t> valid;
        for (const auto &it : alignment)
            if (it.second != -1) valid.push_back((uint32_t)it.second);
        // add unaligned bases
        std::vector<int> head_path, tail_path;
        int begin = AddSequence(s, 0, valid.front(), head_path);
        int prev = begin >= 0 ? (int)code.size() - 1 : -1;
        int last = AddSequence(s, valid.back() + 1, len, tail_path);
        path = head_path;
        // add aligned bases
        for (const auto &it : alignment) {
            if (it.second == -1) continue;
            const int c = coder[(unsigned char)s[it.second]];
            int curr = -1;
            if (it.first == -1) {
                curr = AddNode(c);
            } else {
                const int jt = it.first;
                if (code[jt] == c) {
                    curr = jt;
                } else {
                    for (int kt : aligned[jt])
                        if (code[kt] == c) {
                            curr = kt;
                            break;
                        }
                    if (curr < 0) {
                        curr = AddNode(c);
                        for (int kt : aligned[jt]) {
                            aligned[kt].push_back(curr);
                            aligned[curr].push_back(kt);
                        }
                        aligned[jt].push_back(curr);
                        aligned[curr].push_back(jt);
                    }
                }
            }
            if (begin < 0) begin = curr;
            if (prev >= 0) AddEdge(prev, curr);
            prev = curr;
            path.push_back(curr);
        }
        if (last >= 0) AddEdge(prev, last);
        path.insert(path.end(), tail_path.begin(), tail_path.end());
        seq_path.push_back(path);
        TopologicalSort();
    }

    // graph.cpp:248-301: depth-first, predecessors first, a node's aligned set right behind it
    void TopologicalSort()
    {
        rank_to_node.clear();
        const size_t n = code.size();
        std::vector<uint8_t> marks(n, 0), ignored(n, 0);
        std::vector<int> stack;
        for (size_t s = 0; s < n; s++) {
            if (marks[s] != 0) continue;
            stack.push_back((int)s);
            while (!stack.empty()) {
                const int curr = stack.back();
                bool is_valid = true;
                if (marks[curr] != 2) {
                    for (int t : in_tail[curr])
                        if (marks[t] != 2) stack.push_back(t), is_valid = false;
                    if (!ignored[curr])
                        for (int a : aligned[curr])
                            if (marks[a] != 2) stack.push_back(a), ignored[a] = 1, is_valid = false;
                    if (is_valid) {
                        marks[curr] = 2;
                        if (!ignored[curr]) {
                            rank_to_node.push_back(curr);
                            for (int a : aligned[curr]) rank_to_node.push_back(a);
                        }
                    } else {
                        marks[curr] = 1;
                    }
                }
                if (is_valid) stack.pop_back();
            }
        }
    }

    // graph.cpp:319-357 (+ :303-317): one column per rank, aligned nodes share it
    std::vector<std::string> Msa() const
    {
        std::vector<uint32_t> column(code.size());
        uint32_t j = 0;
        for (uint32_t i = 0; i < rank_to_node.size(); ++i, ++j) {
            const int it = rank_to_node[i];
            column[it] = j;
            for (int a : aligned[it]) column[a] = j, ++i;
        }
        std::vector<std::string> dst;
        for (const auto &path : seq_path) {
            std::string row(j, '-');
            for (int node : path) row[column[node]] = decoder[code[node]];
            dst.push_back(row);
        }
        return dst;
    }
};

struct Engine { // SisdAlignmentEngine, type kNW, subtype kLinear
    int m = 5, n = -4, g = -8;
    std::vector<int32_t> H, profile;
    std::vector<uint32_t> rank;

    // sisd_alignment_engine.cpp:259-293 (Align), :118-257 (Initialize), :295-456 (Linear)
    std::vector<std::pair<int32_t, int32_t>> Align(const std::string &s, const Graph &G)
    {
        std::vector<std::pair<int32_t, int32_t>> alignment;
        const uint32_t len = (uint32_t)s.size();
        const size_t nodes = G.code.size();
        if (nodes == 0 || len == 0) return alignment;
        const uint64_t W = (uint64_t)len + 1, Hh = nodes + 1;
        H.resize(W * Hh);
        profile.resize((size_t)G.num_codes * W);
        rank.resize(nodes);
        for (int c = 0; c < G.num_codes; c++) { // :125-132
            profile[(size_t)c * W] = 0;
            for (uint32_t j = 0; j < len; j++) profile[(size_t)c * W + j + 1] = G.decoder[c] == s[j] ? m : n;
        }
        for (uint32_t i = 0; i < G.rank_to_node.size(); i++) rank[G.rank_to_node[i]] = i; // :134-137
        H[0] = 0;                                                                         // :176-178
        for (uint64_t j = 1; j < W; j++) H[j] = (int32_t)j * g;                           // :213-216
        for (uint64_t i = 1; i < Hh; i++) {                                               // :217-225
            const auto &tails = G.in_tail[G.rank_to_node[i - 1]];
            int32_t penalty = tails.empty() ? 0 : kNegativeInfinity;
            for (int t : tails) penalty = std::max(penalty, H[((uint64_t)rank[t] + 1) * W]);
            H[i * W] = penalty + g;
        }
        int32_t max_score = kNegativeInfinity;
        uint32_t max_i = 0, max_j = 0;
        for (int it : G.rank_to_node) { // :318-364
            const int32_t *prof = &profile[(size_t)G.code[it] * W];
            const uint32_t i = rank[it] + 1;
            const auto &tails = G.in_tail[it];
            uint32_t pred_i = tails.empty() ? 0 : rank[tails[0]] + 1;
            int32_t *row = &H[(uint64_t)i * W];
            const int32_t *pred = &H[(uint64_t)pred_i * W];
            for (uint64_t j = 1; j < W; j++) row[j] = std::max(pred[j - 1] + prof[j], pred[j] + g);
            for (size_t p = 1; p < tails.size(); p++) {
                pred = &H[((uint64_t)rank[tails[p]] + 1) * W];
                for (uint64_t j = 1; j < W; j++) row[j] = std::max(pred[j - 1] + prof[j], std::max(row[j], pred[j] + g));
            }
            for (uint64_t j = 1; j < W; j++) {
                row[j] = std::max(row[j - 1] + g, row[j]);
                if (G.out_head[it].empty() && j == W - 1 && max_score < row[j]) max_score = row[j], max_i = i, max_j = (uint32_t)j;
            }
        }
        if (max_i == 0 && max_j == 0) return alignment; // :366-368
        // backtrack: diagonal through the in-edges in their order, then vertical, then horizontal (:374-452)
        uint32_t i = max_i, j = max_j, prev_i = 0, prev_j = 0;
        while (!(i == 0 && j == 0)) {
            const int32_t Hij = H[(uint64_t)i * W + j];
            bool found = false;
            if (i != 0 && j != 0) {
                const int it = G.rank_to_node[i - 1];
                const int32_t match = profile[(size_t)G.code[it] * W + j];
                const auto &tails = G.in_tail[it];
                const uint32_t p0 = tails.empty() ? 0 : rank[tails[0]] + 1;
                if (Hij == H[(uint64_t)p0 * W + (j - 1)] + match) {
                    prev_i = p0, prev_j = j - 1, found = true;
                } else {
                    for (size_t p = 1; p < tails.size(); p++) {
                        const uint32_t pi = rank[tails[p]] + 1;
                        if (Hij == H[(uint64_t)pi * W + (j - 1)] + match) {
                            prev_i = pi, prev_j = j - 1, found = true;
                            break;
                        }
                    }
                }
            }
            if (!found && i != 0) {
                const auto &tails = G.in_tail[G.rank_to_node[i - 1]];
                const uint32_t p0 = tails.empty() ? 0 : rank[tails[0]] + 1;
                if (Hij == H[(uint64_t)p0 * W + j] + g) {
                    prev_i = p0, prev_j = j, found = true;
                } else {
                    for (size_t p = 1; p < tails.size(); p++) {
                        const uint32_t pi = rank[tails[p]] + 1;
                        if (Hij == H[(uint64_t)pi * W + j] + g) {
                            prev_i = pi, prev_j = j, found = true;
                            break;
                        }
                    }
                }
            }
            if (!found && Hij == H[(uint64_t)i * W + j - 1] + g) prev_i = i, prev_j = j - 1, found = true;
            alignment.emplace_back(i == prev_i ? -1 : G.rank_to_node[i - 1], j == prev_j ? -1 : (int32_t)j - 1);
            i = prev_i, j = prev_j;
        }
        std::reverse(alignment.begin(), alignment.end());
        return alignment;
    }
};

} // namespace

int main(int argc, char **argv)
{
    Engine engine;
    std::string chunk;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() { return i + 1 < argc ? atoi(argv[++i]) : 0; };
        if (a == "-m") engine.m = val();
        else if (a == "-n") engine.n = val();
        else if (a == "-g") engine.g = val();
        else if (a == "--chunk" && i + 1 < argc) chunk = argv[++i];
    }
    std::ifstream in(chunk);
    if (chunk.empty() || !in) {
        fprintf(stderr, "usage: %s --chunk <file.tmp> [-m 5 -n -4 -g -8]\n", argv[0]);
        return 1;
    }
    std::string line;
    while (std::getline(in, line)) { // one block per line, '@' = newline (sibeliaz:89); header -> "s name start len strand size"
        std::vector<std::string> header, seq;
        size_t p = 0;
        while (p < line.size()) {
            size_t q = line.find('@', p);
            if (q == std::string::npos) q = line.size();
            std::string tok = line.substr(p, q - p);
            p = q + 1;
            if (tok.empty()) continue;
            if (tok[0] == '>') {
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
        Graph G;
        for (const auto &s : seq) G.AddAlignment(engine.Align(s, G), s); // main.cpp:282-320
        std::vector<std::string> rows = G.Msa();
        if (rows.empty()) continue;
        std::cout << "\na\n";
        for (size_t i = 0; i < rows.size() && i < header.size(); i++) std::cout << header[i] << ' ' << rows[i] << "\n";
    }
    return 0;
}
