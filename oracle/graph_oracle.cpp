/*
 * graph_oracle.cpp -- TEST INFRASTRUCTURE: CPU restatement of what TwoPaCo's graph constructor computes (SURVEY.md
 * section 8f row 1, the step before the sibeliaz-lcb hot path): the junction positions of the compacted de Bruijn graph
 * of a set of FASTA records and a consistent (vertex id, strand) label for each, in the junction-file wire format of
 * TwoPaCo/src/common/junctionapi.h:106-136.  Only tests/, __graft_entry__.smoke() and bench.py may load it.
 *
 * The reference reaches its result through a Bloom filter (candidates) followed by an exact hash-set pass; this file
 * restates the RESULT of those passes in the limit of no Bloom false positives (every rule cites the code it follows):
 *
 *   S' = 'N' + record (non-ACGT -> 'N') + 'N'                         vertexenumerator.h:1150-1190 (DistributeTasks)
 *   edge set E (strand-symmetric (k+1)-mers):                          :1033-1060 (FilterFillerWorker)
 *       for every definite k-mer X at pos: X.next if next is definite, else the two dummies X.A and X.T;
 *       if prev is not definite also the dummies A.X and T.X
 *   candidate(pos)  <=>  X definite and (in > 1 or out > 1), where in = 2 if prev is not definite else |{c : c.X in E}|,
 *       out likewise                                                   :630-660 (CandidateCheckingWorker)
 *   X is a bifurcation  <=>  it has >= 2 candidate occurrences and, in canonical orientation, their (prev, next) pairs
 *       are not all equal, or they are equal but prev (or next) is 'N' :760-790 (CandidateFinalFilteringWorker),
 *       candidateoccurence.h:26-50, and count <= abundance             :1228-1256 (TrueBifurcations)
 *   a junction is written at pos <=> candidate(pos) and X is a bifurcation; the first and the last k-mer of a record get
 *       a unique "stub" id when they are not junctions otherwise       :905-925 (EdgeConstructionWorker)
 *
 * Labels.  The reference orients every bifurcation k-mer by comparing two rolling-hash values whose character table
 * is seeded from /dev/urandom (ngramhashing/mersennetwister.h:242-262, candidateoccurence.h:34), sorts the stored
 * orientations (bifurcationstorage.h:65) and uses the 1-based rank as the id: ids and signs change from run to run.
 * What is invariant -- and all that sibeliaz-lcb depends on -- is the partition of the junction positions into vertices
 * and the relative strand inside each vertex.  This oracle therefore uses a fixed rule (orientation = the smaller of the
 * k-mer and its reverse complement with A<C<G<T, id = 1 + rank of that k-mer, stubs numbered in genome order from
 * #bifurcations + 42 as in :393), and gro_canonicalize() maps any junction file to the label-free normal form
 * (vertices renumbered by first appearance, first appearance positive) in which two files can be compared byte for
 * byte.  tests/test_graph_oracle.py pins this restatement against the compiled reference twopaco in that normal form.
 *
 * Limits: k odd.  k <= 31: one 64-bit word per k-mer; larger k: the k-mer as a string (same order: 'A' < 'C' < 'G' < 'T').
 */
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

inline int Code(char c)
{
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    }
    return 4; // not definite
}

// FASTA records as TwoPaCo's parser yields them (streamfastaparser.cpp:28-92): header up to the first whitespace is
// irrelevant here; sequence characters are upper-cased, whitespace skipped
bool ReadFasta(const char *path, std::vector<std::string> &records, std::string &err)
{
    std::ifstream in(path);
    if (!in) {
        err = std::string("Can't open file ") + path;
        return false;
    }
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
            if (!isspace((unsigned char)c)) records.back().push_back((char)toupper((unsigned char)c));
    }
    return true;
}

struct VertexInfo {
    uint8_t in = 0, out = 0; // canonical-orientation neighbour characters seen in E (bit c)
    uint32_t pairs = 0;      // (prev, next) pairs of the candidate occurrences, 5 x 5 bits
    uint32_t cand = 0;       // candidate occurrences (saturating)
    int64_t id = 0;
};

struct Junction {
    uint32_t chr, pos;
    int64_t id;
};

// Key = uint64_t: 2 bits per base, first base most significant (k <= 31).  Key = std::string: the bases themselves.
template <class Key>
struct Builder {
    int k;
    uint64_t mask;
    std::vector<std::string> rec;
    std::unordered_map<Key, VertexInfo> vtx;
    std::vector<Junction> out;

    // calls f(pos in S', forward k-mer, reverse-complement k-mer, prev code, next code) for every definite k-mer of S'
    template <class F>
    void ForEachKmer(const std::string &s, F f) const
    {
        const size_t L = s.size(); // S' = N + s + N, S'[i] = s[i-1]
        auto at = [&](size_t i) { return (i == 0 || i == L + 1) ? 4 : Code(s[i - 1]); };
        if (L + 2 < (size_t)k + 2) return;
        uint64_t fw = 0, rc = 0;
        int definite = 0;
        for (size_t i = 1; i <= L; i++) { // S'[i] enters the window ending at i
            const int c = at(i);
            if (c < 4) {
                fw = ((fw << 2) | (uint64_t)c) & mask;
                rc = (rc >> 2) | ((uint64_t)(3 - c) << ((2 * (k - 1)) & 63));
                definite++;
            } else {
                fw = rc = 0;
                definite = 0;
            }
            if (i >= (size_t)k && definite >= k) {
                const size_t pos = i - k + 1; // window S'[pos, pos + k)
                Emit(s, pos, fw, rc, at(pos - 1), at(pos + k), f, (Key *)nullptr);
            }
        }
    }
    template <class F>
    static void Emit(const std::string &, size_t pos, uint64_t fw, uint64_t rc, int prev, int next, F &f, uint64_t *)
    {
        f(pos, fw, rc, prev, next);
    }
    template <class F>
    void Emit(const std::string &s, size_t pos, uint64_t, uint64_t, int prev, int next, F &f, std::string *) const
    {
        const std::string fw = s.substr(pos - 1, (size_t)k); // S'[pos, pos + k) = s[pos - 1, pos - 1 + k)
        std::string rc(fw.rbegin(), fw.rend());
        for (char &c : rc) c = "TGCA"[Code(c)];
        f(pos, fw, rc, prev, next);
    }

    void Run(uint64_t abundance)
    {
        mask = k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1); // (unused by the string keys)
        // ---- E as per-vertex neighbour masks in canonical orientation
        for (const std::string &s : rec)
            ForEachKmer(s, [&](size_t, const Key &fw, const Key &rc, int prev, int next) {
                const bool fwd = fw < rc;
                VertexInfo &v = vtx[fwd ? fw : rc];
                const uint8_t dummy = (1u << 0) | (1u << 3); // A and T: closed under complement
                uint8_t in = prev < 4 ? (uint8_t)(1u << prev) : dummy, outm = next < 4 ? (uint8_t)(1u << next) : dummy;
                if (fwd) {
                    v.in |= in, v.out |= outm;
                } else { // reverse occurrence: its in-neighbours are the canonical k-mer's out-neighbours, complemented
                    auto comp = [](uint8_t m) { return (uint8_t)(((m & 1) << 3) | ((m & 2) << 1) | ((m & 4) >> 1) | ((m & 8) >> 3)); };
                    v.out |= comp(in), v.in |= comp(outm);
                }
            });
        // ---- candidates and the pairs they show
        auto candidate = [&](const VertexInfo &v, bool fwd, int prev, int next) {
            const int in = prev < 4 ? __builtin_popcount(fwd ? v.in : v.out) : 2;
            const int outc = next < 4 ? __builtin_popcount(fwd ? v.out : v.in) : 2;
            return in > 1 || outc > 1;
        };
        for (const std::string &s : rec)
            ForEachKmer(s, [&](size_t, const Key &fw, const Key &rc, int prev, int next) {
                const bool fwd = fw < rc;
                VertexInfo &v = vtx[fwd ? fw : rc];
                if (!candidate(v, fwd, prev, next)) return;
                const int cp = fwd ? prev : (next < 4 ? 3 - next : 4), cn = fwd ? next : (prev < 4 ? 3 - prev : 4);
                v.pairs |= 1u << (cp * 5 + cn);
                if (v.cand < 0xFFFFFFFFu) v.cand++;
            });
        // ---- bifurcations, ids = 1 + rank of the canonical k-mer
        std::vector<Key> keys;
        for (auto &kv : vtx) {
            VertexInfo &v = kv.second;
            bool bif = false;
            if (v.cand >= 2) {
                if (__builtin_popcount(v.pairs) >= 2) bif = true;
                else {
                    const int p = __builtin_ctz(v.pairs);
                    bif = p / 5 == 4 || p % 5 == 4; // the shared prev (or next) is 'N': unknown twice
                }
            }
            if (bif && (uint64_t)v.cand <= abundance) keys.push_back(kv.first);
            else v.cand = 0;
        }
        std::sort(keys.begin(), keys.end());
        for (size_t i = 0; i < keys.size(); i++) vtx[keys[i]].id = (int64_t)i + 1;
        int64_t stub = (int64_t)keys.size() + 42;
        // ---- junction records in genome order
        for (size_t c = 0; c < rec.size(); c++) {
            const std::string &s = rec[c];
            const size_t L = s.size();
            if (L < (size_t)k) continue;
            std::vector<Junction> here;
            ForEachKmer(s, [&](size_t pos, const Key &fw, const Key &rc, int prev, int next) {
                const bool fwd = fw < rc;
                const VertexInfo &v = vtx[fwd ? fw : rc];
                if (v.id && candidate(v, fwd, prev, next)) here.push_back(Junction{(uint32_t)c, (uint32_t)(pos - 1), fwd ? v.id : -v.id});
            });
            const uint32_t first = 0, last = (uint32_t)(L - k);
            const bool has_first = !here.empty() && here.front().pos == first, has_last = !here.empty() && here.back().pos == last;
            if (!has_first) out.push_back(Junction{(uint32_t)c, first, stub++});
            for (const Junction &j : here) out.push_back(j);
            if (!has_last && last != first) out.push_back(Junction{(uint32_t)c, last, stub++});
        }
    }
};

void WriteFile(const std::vector<Junction> &js, FILE *f)
{
    uint32_t now = 0;
    auto put = [f](uint32_t pos, int64_t id) {
        fwrite(&pos, 4, 1, f);
        fwrite(&id, 8, 1, f);
    };
    for (const Junction &j : js) {
        for (; j.chr > now; ++now) put(0xFFFFFFFFu, INT64_MAX); // junctionapi.h:117-123
        put(j.pos, j.id);
    }
}

bool ReadFile(const char *path, std::vector<Junction> &js)
{
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    uint32_t chr = 0;
    while (true) {
        uint32_t pos;
        int64_t id;
        if (fread(&pos, 4, 1, f) != 1 || fread(&id, 8, 1, f) != 1) break;
        if (pos == 0xFFFFFFFFu || id == INT64_MAX) { // junctionapi.h:80-98
            chr++;
            continue;
        }
        js.push_back(Junction{chr, pos, id});
    }
    fclose(f);
    return true;
}

} // namespace

extern "C" {

/* Builds the junction file of the given FASTA files at vertex size k.  Returns the number of junction records, -1 on error. */
int64_t gro_build(const char *const *fastas, int n, int k, uint64_t abundance, const char *out_path, char *err, int errlen)
{
    if (k < 1 || k % 2 == 0) {
        snprintf(err, errlen, "k must be odd");
        return -1;
    }
    std::vector<std::string> rec;
    std::string e;
    for (int i = 0; i < n; i++)
        if (!ReadFasta(fastas[i], rec, e)) {
            snprintf(err, errlen, "%s", e.c_str());
            return -1;
        }
    for (std::string &s : rec)
        for (char &c : s)
            if (Code(c) == 4) c = 'N';
    std::vector<Junction> js;
    // GRO_STRING_KEYS: the string-keyed restatement for a small k too (a test compares the two)
    if (k <= 31 && !getenv("GRO_STRING_KEYS")) {
        Builder<uint64_t> b;
        b.k = k;
        b.rec.swap(rec);
        b.Run(abundance);
        js.swap(b.out);
    } else {
        Builder<std::string> b;
        b.k = k;
        b.rec.swap(rec);
        b.Run(abundance);
        js.swap(b.out);
    }
    FILE *f = fopen(out_path, "wb");
    if (!f) {
        snprintf(err, errlen, "Can't create the output file");
        return -1;
    }
    WriteFile(js, f);
    fclose(f);
    return (int64_t)js.size();
}

/* Label-free normal form of a junction file: vertices renumbered 1, 2, ... by first appearance, the first appearance of
 * every vertex positive.  Returns the number of records written to out_path, -1 on error. */
int64_t gro_canonicalize(const char *in_path, const char *out_path)
{
    std::vector<Junction> js;
    if (!ReadFile(in_path, js)) return -1;
    std::unordered_map<int64_t, std::pair<int64_t, bool>> label; // |id| -> (new id, flip)
    for (Junction &j : js) {
        const int64_t a = j.id < 0 ? -j.id : j.id;
        auto it = label.find(a);
        if (it == label.end()) it = label.emplace(a, std::make_pair((int64_t)label.size() + 1, j.id < 0)).first;
        const bool neg = (j.id < 0) != it->second.second;
        j.id = neg ? -it->second.first : it->second.first;
    }
    FILE *f = fopen(out_path, "wb");
    if (!f) return -1;
    WriteFile(js, f);
    fclose(f);
    return (int64_t)js.size();
}
}

#ifdef GRAPH_ORACLE_MAIN
int main(int argc, char **argv)
{
    if (argc >= 4 && !strcmp(argv[1], "canon")) return gro_canonicalize(argv[2], argv[3]) < 0;
    if (argc < 5) {
        fprintf(stderr, "usage: graph_oracle build k out.dbg fasta...  |  graph_oracle canon in.dbg out.dbg\n");
        return 2;
    }
    char err[256];
    int64_t n = gro_build(argv + 4, argc - 4, atoi(argv[2]), UINT64_MAX, argv[3], err, sizeof err);
    if (n < 0) {
        fprintf(stderr, "error: %s\n", err);
        return 1;
    }
    printf("%lld junction records\n", (long long)n);
    return 0;
}
#endif
