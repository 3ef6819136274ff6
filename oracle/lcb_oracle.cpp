/*
 * lcb_oracle.cpp -- CPU restatement of the sibeliaz-lcb hot path (JunctionStorage + BlocksFinder).
 *
 * TEST INFRASTRUCTURE ONLY (see lcb_oracle.h).  Written from the behaviour of the reference
 * (/root/reference/SibeliaZ-LCB, v1.2.7); every function cites the file:line it restates.  The data
 * layout is the structure-of-arrays index of SURVEY.md section 7, not the reference's AoS vectors,
 * and the control flow is sequential: one phase of `phase_size` seeds is evaluated against the
 * `used` flags as they stand at the phase start, then committed in seed order -- which is what the
 * reference's OpenMP region computes for any thread count (blocksfinder.h:343-431).
 *
 * C++ rather than C on purpose: the output stage is only reproducible byte for byte if the very
 * same libstdc++ std::sort (unstable introsort) runs on the very same sequences as in the reference
 * (blocksfinder.h:103,662; blocksfinder.cpp:146).
 *
 * Pinned by tests/test_oracle_pin.py against the reference's golden GFF and against oracle/_ref.
 */
#include "lcb_oracle.h"

#include <algorithm>
#include <cerrno>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <sys/types.h>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// dnachar.cpp:52-85 -- complement table: A<->T, C<->G, everything else -> 'N'; valid IUPAC set.
// ---------------------------------------------------------------------------------------------
inline uint8_t ReverseChar(uint8_t c)
{
    switch (c) {
    case 'A': return 'T';
    case 'T': return 'A';
    case 'C': return 'G';
    case 'G': return 'C';
    }
    return 'N';
}

inline bool IsValidChar(int c)
{
    static const char *valid = "ACGTURYKMSWBDHWNXV"; // dnachar.cpp:13
    return c > 0 && c < 128 && strchr(valid, c) != nullptr;
}

struct Seed { // BlocksFinder::Bundle, blocksfinder.h:182-209
    int64_t vid;
    uint8_t ch;
    uint64_t count, rank, res_pos, res_chr;
    bool operator<(const Seed &a) const
    {
        if (count != a.count) return count > a.count;
        if (rank != a.rank) return rank < a.rank;
        if (res_pos != a.res_pos) return res_pos < a.res_pos;
        return res_chr < a.res_chr;
    }
};

struct Inst { // Path::Instance, path.h:53-181
    int64_t fg, bg;       // front_/back_ as global position indices
    bool pos;             // strand of both iterators
    bool ffin, bfin;      // frontFinished_/backFinished_
    int64_t fdist, bdist; // frontDistance_/backDistance_
    int64_t key;          // compareIdx_ (as a global index: same order inside a chromosome)
    int32_t chr;
};

struct PEdge { // Edge, junctionstorage.h:21-114 (fields the path uses)
    int64_t sv, ev;
    uint8_t ch;
    int64_t len;
};

struct Point { // Path::Point, path.h:185-218
    PEdge e;
    int64_t start_dist;
};

struct Block { // BlockInstance, blocksfinder.h:29-51
    int id;
    size_t start, end, chr;
    int Sign() const { return id > 0 ? +1 : -1; }
    int BlockId() const { return abs(id); }
    size_t Length() const { return end - start; }
    bool operator<(const Block &o) const // blocksfinder.cpp:104-107
    {
        return std::make_pair(BlockId(), std::make_pair(chr, start)) <
               std::make_pair(o.BlockId(), std::make_pair(o.chr, o.start));
    }
};

bool CompareById(const Block &a, const Block &b) { return a.BlockId() < b.BlockId(); } // blocksfinder.cpp:32-35

} // namespace

struct lcbo {
    int64_t k = 0;
    // ---- index (JunctionStorage) ----
    int32_t C = 0;
    int64_t N = 0, V = 0;
    std::vector<int64_t> chr_off; // C+1
    std::vector<int32_t> pos_id;  // Position.id   junctionstorage.h:142
    std::vector<uint32_t> pos_bp; // Position.pos  :143
    std::vector<uint8_t> used;    // Position.used :144
    std::vector<uint8_t> next_ch; // Vertex.ch     :641
    std::vector<uint8_t> prev_rc; // Vertex.revCh  :642
    std::vector<uint32_t> pos_chr;
    std::vector<int64_t> vtx_off; // CSR over vertex_[v]
    std::vector<int64_t> occ_g;
    std::vector<std::string> seq, header;
    // ---- seeds ----
    std::vector<Seed> seed;
    // ---- path state (Path, path.h) ----
    int64_t max_branch = 0, min_block = 0, max_flank = 0, looking_depth = 8;
    std::vector<int> distance; // DistanceKeeper, distancekeeper.h:9-41
    std::vector<uint32_t> count;
    std::vector<int64_t> count_touched;
    std::vector<Inst> inst;               // pool; index == position in allInstance_
    std::vector<std::vector<int>> order;  // instance_[chr] in multiset order (ids into inst)
    std::vector<int> good;                // goodInstance_
    std::vector<Point> left_body, right_body;
    int64_t origin = 0, left_flank = 0, right_flank = 0;
    // ---- results ----
    std::vector<Block> blocks; // blocksInstance_
    int64_t blocks_found = 0;
    uint64_t ctr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#ifdef LCBO_EPOCH
    const uint32_t *epoch = nullptr; // caller-owned epoch array (see lcb_oracle_epoch API below)
    uint32_t thresh = 0;
    mutable std::vector<int64_t> *readlog = nullptr;
#endif

    // ----- iterator helpers: JunctionSequentialIterator, junctionstorage.h:158-396 -----
    int64_t Vertex(int64_t g, bool pos) const { return pos ? pos_id[g] : -(int64_t)pos_id[g]; }   // :171-174
    int64_t Position(int64_t g, bool pos) const { return pos ? pos_bp[g] : (int64_t)pos_bp[g] + k; } // :176-184
    bool Valid(int64_t g, int32_t chr) const { return g >= chr_off[chr] && g < chr_off[chr + 1]; }  // :265-268
    bool IsUsed(int64_t g, bool pos) const                                                          // :270-283
    {
#ifdef LCBO_EPOCH /* experiments/jacobi_proto.cpp: `used` seen through an epoch threshold, reads logged */
        if (!pos && g <= chr_off[pos_chr[g]]) return false;
        int64_t f = pos ? g : g - 1;
        if (readlog) readlog->push_back(f);
        return epoch[f] < thresh;
#else
        if (pos) return used[g] != 0;
        if (g > chr_off[pos_chr[g]]) return used[g - 1] != 0;
        return false;
#endif
    }
    void MarkUsed(int64_t g, bool pos) // :285-295
    {
        if (pos) used[g] = 1;
        else if (g > chr_off[pos_chr[g]]) used[g - 1] = 1;
    }
    uint8_t GetChar(int64_t g, bool pos) const { return pos ? next_ch[g] : prev_rc[g]; } // :234-243
    static int64_t Step(int64_t g, bool pos, int64_t by = 1) { return pos ? g + by : g - by; }      // :376-386
    PEdge OutgoingEdge(int64_t g, bool pos) const // :191-208
    {
        if (pos) return PEdge{pos_id[g], pos_id[g + 1], next_ch[g], (int64_t)pos_bp[g + 1] - (int64_t)pos_bp[g]};
        return PEdge{-(int64_t)pos_id[g], -(int64_t)pos_id[g - 1], prev_rc[g], (int64_t)pos_bp[g] - (int64_t)pos_bp[g - 1]};
    }
    PEdge IngoingEdge(int64_t g, bool pos) const // :210-227
    {
        if (pos) return PEdge{pos_id[g - 1], pos_id[g], next_ch[g - 1], (int64_t)pos_bp[g] - (int64_t)pos_bp[g - 1]};
        return PEdge{-(int64_t)pos_id[g + 1], -(int64_t)pos_id[g], prev_rc[g + 1], (int64_t)pos_bp[g + 1] - (int64_t)pos_bp[g]};
    }
    static bool IterLess(int64_t g1, bool p1, int64_t g2, bool p2) // :349-362 (strand, chr, idx)
    {
        if (p1 != p2) return p1 < p2;
        return g1 < g2;
    }

    // ----- DistanceKeeper -----
    bool DistIsSet(int64_t v) const { return distance[v + V] != INT_MAX; }
    void DistSet(int64_t v, int d) { distance[v + V] = d; }
    int DistGet(int64_t v) const { return distance[v + V]; }
    void DistUnset(int64_t v) { distance[v + V] = INT_MAX; }

    // ----- Path -----
    int64_t RealLength(const Inst &a) const { return llabs(Position(a.fg, a.pos) - Position(a.bg, a.pos)); } // path.h:165-168
    bool IsGood(const Inst &a) const { return RealLength(a) >= min_block; }                                   // :645-648
    static bool Within(const Inst &a, int64_t g) // :170-175
    {
        int64_t l = std::min(a.fg, a.bg), r = std::max(a.fg, a.bg);
        return g >= l && g <= r;
    }
    int UpperBound(const std::vector<int> &s, int64_t key) const // std::multiset::upper_bound on compareIdx_
    {
        int lo = 0, hi = (int)s.size();
        while (lo < hi) {
            int mid = (lo + hi) / 2;
            if (inst[s[mid]].key > key) hi = mid;
            else lo = mid + 1;
        }
        return lo;
    }
    void InsertInstance(int64_t g, bool pos, int64_t dist) // Instance ctor :82-91 + multiset insert + allInstance_.push_back
    {
        Inst a;
        a.fg = a.bg = g;
        a.pos = pos;
        a.ffin = a.bfin = false;
        a.fdist = a.bdist = dist;
        a.key = g;
        a.chr = (int32_t)pos_chr[g];
        int id = (int)inst.size();
        inst.push_back(a);
        std::vector<int> &s = order[a.chr];
        s.insert(s.begin() + UpperBound(s, g), id);
    }

    void PathInit(int64_t vid, uint8_t ch) // Path::Init, path.h:33-46
    {
        origin = vid;
        DistSet(vid, 0);
        left_flank = right_flank = 0;
        int64_t av = llabs(vid);
        for (int64_t o = vtx_off[av]; o < vtx_off[av + 1]; o++) {
            int64_t g = occ_g[o];
            bool pos = pos_id[g] == vid;
            if (!IsUsed(g, pos) && ch == GetChar(g, pos)) InsertInstance(g, pos, 0);
        }
    }

    uint64_t mx[3] = {0, 0, 0}; // diagnostics: largest path (vertices), instance table and good list seen at a Clear
    void PathClear() // Path::Clear, path.h:650-677
    {
        mx[0] = std::max<uint64_t>(mx[0], left_body.size() + right_body.size() + 1);
        mx[1] = std::max<uint64_t>(mx[1], inst.size());
        mx[2] = std::max<uint64_t>(mx[2], good.size());
        for (auto &pt : left_body) DistUnset(pt.e.sv);
        for (auto &pt : right_body) DistUnset(pt.e.ev);
        left_body.clear();
        right_body.clear();
        DistUnset(origin);
        for (auto &a : inst) order[a.chr].clear();
        inst.clear();
        good.clear();
    }

    int64_t LeftDistance() const { return -left_flank; }
    int64_t RightDistance() const { return right_flank; }
    int64_t MiddlePathLength() const { return LeftDistance() + RightDistance(); }
    int64_t RightVertex() const { return right_body.empty() ? origin : right_body.back().e.ev; } // :285-293
    int64_t LeftVertex() const { return left_body.empty() ? origin : left_body.back().e.sv; }    // :320-328

    bool Compatible(int64_t sg, bool spos, int64_t eg, bool epos, const PEdge &e) // path.h:380-428
    {
        if (spos != epos) return false;
#ifdef LCBO_EPOCH /* experiment: the pure distance tests first, so the flag scan is bounded by max_branch bp */
        {
            int64_t rd = Position(eg, epos) - Position(sg, spos);
            if (!spos) rd = -rd;
            if (rd < 0) return false;
            int64_t ad = (int64_t)DistGet(Vertex(eg, epos)) - (int64_t)DistGet(Vertex(sg, spos));
            int64_t n1 = Step(sg, spos);
            if ((rd > max_branch || ad > max_branch) &&
                (!Valid(n1, (int32_t)pos_chr[sg]) || GetChar(sg, spos) != e.ch || eg != n1 || Vertex(n1, spos) != e.ev))
                return false;
        }
#endif
        for (int64_t it = sg; it != eg; it = Step(it, spos)) {
            ctr[2]++;
            if (IsUsed(it, spos)) return false;
        }
        int64_t real_diff = Position(eg, epos) - Position(sg, spos);
        int64_t anc_diff = (int64_t)DistGet(Vertex(eg, epos)) - (int64_t)DistGet(Vertex(sg, spos));
        int64_t s1 = Step(sg, spos);
        int32_t chr = (int32_t)pos_chr[sg];
        if (spos) {
            if (real_diff < 0) return false;
            if ((real_diff > max_branch || anc_diff > max_branch) &&
                (!Valid(s1, chr) || GetChar(sg, spos) != e.ch || eg != s1 || Vertex(s1, spos) != e.ev))
                return false;
        } else {
            if (-real_diff < 0) return false;
            if ((-real_diff > max_branch || anc_diff > max_branch) &&
                (!Valid(s1, chr) || GetChar(sg, spos) != e.ch || eg != s1 || Vertex(s1, spos) != e.ev))
                return false;
        }
        return true;
    }

    bool PointPushBack(const PEdge &e) // path.h:568-584 + PointPushBackWorker :499-566
    {
        int64_t vertex = e.ev;
        if (DistIsSet(vertex)) return false;
        int64_t start_dist = right_flank;
        int64_t dist = start_dist + e.len;
        DistSet(vertex, (int)dist);
        int64_t av = llabs(vertex);
        for (int64_t o = vtx_off[av]; o < vtx_off[av + 1]; o++) {
            ctr[1]++;
            int64_t g = occ_g[o];
            bool pos = pos_id[g] == vertex;
            std::vector<int> &s = order[pos_chr[g]];
            int ub = UpperBound(s, g);
            if (ub != (int)s.size() && Within(inst[s[ub]], g)) continue;
            bool new_instance = true;
            int cand = -1;
            if (pos) {
                if (ub != 0) {
                    cand = s[ub - 1];
                    if (Compatible(inst[cand].bg, inst[cand].pos, g, pos, e)) new_instance = false;
                }
            } else {
                if (ub != (int)s.size()) {
                    cand = s[ub];
                    if (Compatible(inst[cand].bg, inst[cand].pos, g, pos, e)) new_instance = false;
                }
            }
            if (!new_instance && Vertex(inst[cand].bg, inst[cand].pos) != vertex) {
                Inst &a = inst[cand];
                if (!a.bfin) {
                    bool prev_good = IsGood(a);
                    a.bg = g; // ChangeBack, path.h:124-133
                    a.bdist = dist;
                    if (a.pos) a.key = g;
                    if (!prev_good && IsGood(a)) good.push_back(cand);
                    if (IsUsed(g, pos)) a.bfin = true;
                }
            } else if (!IsUsed(g, pos)) {
                InsertInstance(g, pos, dist);
            }
        }
        right_body.push_back(Point{e, start_dist});
        right_flank = start_dist + e.len;
        return true;
    }

    bool PointPushFront(const PEdge &e) // path.h:586-602 + PointPushFrontWorker :430-497
    {
        int64_t vertex = e.sv;
        if (DistIsSet(vertex)) return false;
        int64_t end_dist = left_flank;
        int64_t dist = end_dist - e.len;
        DistSet(vertex, (int)dist);
        int64_t av = llabs(vertex);
        for (int64_t o = vtx_off[av]; o < vtx_off[av + 1]; o++) {
            ctr[1]++;
            int64_t g = occ_g[o];
            bool pos = pos_id[g] == vertex;
            std::vector<int> &s = order[pos_chr[g]];
            int ub = UpperBound(s, g);
            if (ub != (int)s.size() && Within(inst[s[ub]], g)) continue;
            bool new_instance = true;
            int cand = -1;
            if (pos) {
                if (ub != (int)s.size()) {
                    cand = s[ub];
                    if (Compatible(g, pos, inst[cand].fg, inst[cand].pos, e)) new_instance = false;
                }
            } else {
                if (ub != 0) {
                    cand = s[ub - 1];
                    if (Compatible(g, pos, inst[cand].fg, inst[cand].pos, e)) new_instance = false;
                }
            }
            if (!new_instance && Vertex(inst[cand].fg, inst[cand].pos) != vertex) {
                Inst &a = inst[cand];
                if (!a.ffin) {
                    bool prev_good = IsGood(a);
                    a.fg = g; // ChangeFront, path.h:113-122
                    a.fdist = dist;
                    if (!a.pos) a.key = g;
                    if (!prev_good && IsGood(a)) good.push_back(cand);
                    if (IsUsed(g, pos)) a.ffin = true;
                }
            } else if (!IsUsed(g, pos)) {
                InsertInstance(g, pos, dist);
            }
        }
        left_body.push_back(Point{e, dist});
        left_flank = dist;
        return true;
    }

    int64_t Score() // Path::Score, path.h:604-628
    {
        int64_t ret = 0;
        for (int id : good) {
            ctr[3]++;
            const Inst &a = inst[id];
            int64_t score = RealLength(a);
            int64_t right_pen = RightDistance() - a.bdist;
            int64_t left_pen = LeftDistance() + a.fdist;
            if (left_pen >= max_flank || right_pen >= max_flank) {
                ret = -INT32_MAX;
                break;
            }
            score -= (right_pen + left_pen) * (right_pen + left_pen);
            ret += score;
        }
        return ret;
    }

    // BlocksFinder::MostPopularVertex, blocksfinder.h:708-768.  Returns the vertex (0 = none) and origin.
    int64_t MostPopularVertex(bool forward, bool try_used, int64_t &org_g, bool &org_pos)
    {
        ctr[6]++;
        int64_t best_vid = 0, best_count = 0;
        int64_t best_g = 0;
        bool best_pos = false;
        int64_t start_vid = forward ? RightVertex() : LeftVertex();
        bool use_good = good.size() >= 2;
        size_t n = use_good ? good.size() : inst.size();
        for (size_t q = 0; q < n; q++) {
            const Inst &a = inst[use_good ? good[q] : (int)q];
            int64_t og = forward ? a.bg : a.fg;
            int64_t now_vid = Vertex(og, a.pos);
            if (now_vid != start_vid) continue;
            int64_t weight = llabs(Position(a.fg, a.pos) - Position(a.bg, a.pos)) + 1;
            int64_t it = forward ? Step(og, a.pos) : Step(og, a.pos, -1);
            for (size_t d = 1; Valid(it, a.chr) && (d < (size_t)looking_depth ||
                                                    llabs(Position(it, a.pos) - Position(og, a.pos)) <= max_branch);
                 d++) {
                ctr[0]++;
                int64_t vid = Vertex(it, a.pos);
                if (!DistIsSet(vid) && (!IsUsed(it, a.pos) || try_used)) {
                    int64_t adj = vid + V;
                    if (count[adj] == 0) count_touched.push_back(adj);
                    count[adj] += (uint32_t)weight;
                    if ((int64_t)count[adj] > best_count ||
                        ((int64_t)count[adj] == best_count && IterLess(og, a.pos, best_g, best_pos))) {
                        best_g = og;
                        best_pos = a.pos;
                        best_count = count[adj];
                        best_vid = vid;
                    }
                } else {
                    break;
                }
                it = forward ? Step(it, a.pos) : Step(it, a.pos, -1);
            }
        }
        for (int64_t adj : count_touched) count[adj] = 0;
        count_touched.clear();
        org_g = best_g;
        org_pos = best_pos;
        return best_vid;
    }

    void SnapshotGood(std::vector<Inst> &best)
    {
        best.clear();
        for (int id : good) best.push_back(inst[id]);
    }

    bool ExtendPathForward(size_t &best_right_size, int64_t &best_score, int64_t &now_score,
                           std::vector<Inst> &best) // blocksfinder.h:770-832
    {
        bool success = false;
        int64_t og;
        bool opos;
        int64_t next = MostPopularVertex(true, false, og, opos);
        if (next == 0) next = MostPopularVertex(true, true, og, opos);
        if (next != 0) {
            for (int64_t it = og; Vertex(it, opos) != next; it = Step(it, opos)) {
                ctr[7]++;
                success = PointPushBack(OutgoingEdge(it, opos));
                if (success) {
                    now_score = Score();
                    if (now_score > best_score) {
                        best_score = now_score;
                        best_right_size = right_body.size() + 1;
                        if (now_score > 0) SnapshotGood(best);
                    }
                }
            }
        }
        return success;
    }

    bool ExtendPathBackward(size_t &best_left_size, int64_t &best_score, int64_t &now_score,
                            std::vector<Inst> &best) // blocksfinder.h:834-895
    {
        bool success = false;
        int64_t og;
        bool opos;
        int64_t next = MostPopularVertex(false, false, og, opos);
        if (next != 0) {
            for (int64_t it = og; Vertex(it, opos) != next; it = Step(it, opos, -1)) {
                ctr[7]++;
                success = PointPushFront(IngoingEdge(it, opos));
                if (success) {
                    now_score = Score();
                    if (now_score > best_score) {
                        best_score = now_score;
                        best_left_size = left_body.size() + 1;
                        if (now_score > 0) SnapshotGood(best);
                    }
                }
            }
        }
        return success;
    }

    void Process(const Seed &b, std::vector<Inst> &best) // ProcessVertex::Process, blocksfinder.h:228-310
    {
        ctr[4]++;
        int64_t score = 0; // uninitialised in the reference (:230); never read before being set
        best.clear();
        int64_t vid = b.vid;
        PathInit(vid, b.ch);
        int64_t best_score = 0;
        size_t best_right_size = right_body.size() + 1;
        size_t best_left_size = left_body.size() + 1;
        int64_t min_run = max_branch * 2;
        while (true) {
            bool ret = true, positive = false;
            int64_t prev_length = MiddlePathLength();
            while ((ret = ExtendPathForward(best_right_size, best_score, score, best)) &&
                   MiddlePathLength() - prev_length <= min_run) {
                positive = positive || (score > 0);
            }
            if (!ret || !positive) break;
        }
        std::vector<PEdge> best_edge;
        for (size_t i = 0; i + 1 < best_right_size; i++) best_edge.push_back(right_body[i].e);
        PathClear();
        PathInit(vid, b.ch);
        for (auto &e : best_edge) {
            ctr[7]++;
            PointPushBack(e);
        }
        while (true) {
            bool ret = true, positive = false;
            int64_t prev_length = MiddlePathLength();
            while ((ret = ExtendPathBackward(best_left_size, best_score, score, best)) &&
                   MiddlePathLength() - prev_length <= min_run)
                ; // stray ';' of blocksfinder.h:297 -- the braces below run once
            {
                positive = positive || (score > 0);
            }
            if (!ret || !positive) break;
        }
        PathClear();
    }

    void Finalize(const std::vector<Inst> &instance, std::set<size_t> &invalid_chr) // blocksfinder.h:312-332
    {
        int64_t current = ++blocks_found;
        for (const Inst &a : instance) {
            invalid_chr.insert(a.chr);
            if (a.pos)
                blocks.push_back(Block{(int)+current, (size_t)Position(a.fg, true), (size_t)(Position(a.bg, true) + k), (size_t)a.chr});
            else
                blocks.push_back(Block{(int)-current, (size_t)(Position(a.bg, false) - k), (size_t)Position(a.fg, false), (size_t)a.chr});
            for (int64_t it = a.fg; it != a.bg; it = Step(it, a.pos)) MarkUsed(it, a.pos);
        }
    }
};

// =================================================================================================
// Loading: TwoPaCo::JunctionPositionReader (common/junctionapi.h:80-98), StreamFastaParser
// (common/streamfastaparser.cpp:28-92), JunctionStorage::Init (junctionstorage.h:572-650).
// =================================================================================================
namespace {

void ReadFasta(const std::string &file, std::vector<std::string> &seq, std::vector<std::string> &header)
{
    FILE *f = fopen(file.c_str(), "rb");
    if (!f) throw std::runtime_error("Can't open file " + file);
    std::string data;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, n);
    fclose(f);
    size_t p = 0, sz = data.size();
    std::string current_header;
    while (p < sz) {
        if (data[p] != '>') // streamfastaparser.cpp:33-36
            throw std::runtime_error(std::string("The FASTA header should start with a '>', started with '") + data[p] + "'");
        p++;
        size_t e = p;
        while (e < sz && data[e] != '\n') e++;
        { // header = first whitespace-delimited token (ss >> currentHeader_, :45); kept if the line is blank
            size_t a = p;
            while (a < e && isspace((unsigned char)data[a])) a++;
            size_t b = a;
            while (b < e && !isspace((unsigned char)data[b])) b++;
            if (b > a && e < sz) current_header = data.substr(a, b - a);
        }
        p = e < sz ? e + 1 : e;
        header.push_back(current_header);
        seq.emplace_back();
        std::string &s = seq.back();
        while (p < sz && data[p] != '>') { // GetChar, :60-92
            unsigned char c = (unsigned char)data[p++];
            if (isspace(c)) continue;
            int u = toupper(c);
            if (!IsValidChar(u))
                throw std::runtime_error(std::string("Found an invalid character '") + (char)c + "' in sequence " + current_header);
            s.push_back((char)u);
        }
    }
}

} // namespace

extern "C" lcbo *lcbo_load(const char *graph, const char *const *fastas, int n_fastas, int k, int abundance,
                           char *err, int errlen)
{
    lcbo *L = new lcbo;
    try {
        L->k = k;
        std::vector<uint8_t> raw;
        {
            FILE *f = fopen(graph, "rb");
            if (!f) throw std::runtime_error("Can't read the input file"); // junctionapi.h:48
            uint8_t buf[1 << 16];
            size_t n;
            while ((n = fread(buf, 1, sizeof buf, f)) > 0) raw.insert(raw.end(), buf, buf + n);
            fclose(f);
        }
        size_t nrec = raw.size() / 12;
        // pass 1 (junctionstorage.h:576-594): abundance per |id|, number of chromosomes and vertices
        std::vector<uint32_t> rchr(nrec);
        std::vector<uint8_t> rsep(nrec);
        std::vector<size_t> abund;
        uint32_t chr = 0;
        int64_t maxabs = -1;
        uint32_t maxchr = 0;
        bool any = false;
        for (size_t i = 0; i < nrec; i++) {
            uint32_t pos;
            int64_t id;
            memcpy(&pos, &raw[i * 12], 4);
            memcpy(&id, &raw[i * 12 + 4], 8);
            if (pos == UINT32_MAX || id == INT64_MAX) { // separator, junctionapi.h:92
                rsep[i] = 1;
                chr++;
                continue;
            }
            rchr[i] = chr;
            any = true;
            maxchr = std::max(maxchr, chr);
            size_t a = (size_t)llabs(id);
            if ((int64_t)a > maxabs) maxabs = (int64_t)a;
            if (a >= abund.size()) abund.resize(a + 1, 0);
            ++abund[a];
        }
        L->C = any ? (int32_t)maxchr + 1 : 0;
        L->V = maxabs + 1;
        // pass 2 (:597-617): keep records with abundance < threshold
        L->chr_off.assign(L->C + 1, 0);
        std::vector<int64_t> vcount(L->V + 1, 0);
        for (size_t i = 0; i < nrec; i++) {
            if (rsep[i]) continue;
            uint32_t pos;
            int64_t id;
            memcpy(&pos, &raw[i * 12], 4);
            memcpy(&id, &raw[i * 12 + 4], 8);
            size_t a = (size_t)llabs(id);
            if (abund[a] < (size_t)abundance) {
                L->pos_id.push_back((int32_t)id);
                L->pos_bp.push_back(pos);
                L->pos_chr.push_back(rchr[i]);
                L->chr_off[rchr[i] + 1]++;
                vcount[a]++;
            }
        }
        L->N = (int64_t)L->pos_id.size();
        for (int32_t c = 0; c < L->C; c++) L->chr_off[c + 1] += L->chr_off[c];
        L->vtx_off.assign(L->V + 1, 0);
        for (int64_t v = 0; v < L->V; v++) L->vtx_off[v + 1] = L->vtx_off[v] + vcount[v];
        L->occ_g.resize(L->N);
        {
            std::vector<int64_t> cur(L->vtx_off.begin(), L->vtx_off.end() - 1);
            // records are already in (chr, idx) order, which is the order std::sort gives at :646-649
            for (int64_t g = 0; g < L->N; g++) L->occ_g[cur[llabs((int64_t)L->pos_id[g])]++] = g;
        }
        L->used.assign(L->N, 0);
        // sequences (:620-633)
        for (int i = 0; i < n_fastas; i++) ReadFasta(fastas[i], L->seq, L->header);
        if ((int32_t)L->seq.size() < L->C) throw std::runtime_error("fewer FASTA records than chromosomes in the graph");
        // ch / revCh (:635-644)
        L->next_ch.resize(L->N);
        L->prev_rc.resize(L->N);
        for (int64_t g = 0; g < L->N; g++) {
            const std::string &s = L->seq[L->pos_chr[g]];
            size_t p = L->pos_bp[g];
            L->next_ch[g] = p + k < s.size() ? (uint8_t)s[p + k] : 0; // std::string[size()] == '\0'
            L->prev_rc[g] = p > 0 ? ReverseChar((uint8_t)s[p - 1]) : 'N';
        }
    } catch (std::exception &e) {
        if (err && errlen > 0) snprintf(err, errlen, "%s", e.what());
        delete L;
        return nullptr;
    }
    return L;
}

extern "C" void lcbo_free(lcbo *L) { delete L; }
extern "C" int64_t lcbo_num_records(const lcbo *L) { return L->N; }
extern "C" int64_t lcbo_num_vertices(const lcbo *L) { return L->V; }
extern "C" int32_t lcbo_num_chr(const lcbo *L) { return L->C; }

extern "C" void lcbo_get_index(const lcbo *L, int64_t *chr_off, int32_t *pos_id, uint32_t *pos_bp, uint8_t *next_ch,
                               uint8_t *prev_rc, int64_t *vtx_off, int64_t *occ_g, int64_t *chr_len)
{
    if (chr_off) memcpy(chr_off, L->chr_off.data(), sizeof(int64_t) * (L->C + 1));
    if (pos_id) memcpy(pos_id, L->pos_id.data(), sizeof(int32_t) * L->N);
    if (pos_bp) memcpy(pos_bp, L->pos_bp.data(), sizeof(uint32_t) * L->N);
    if (next_ch) memcpy(next_ch, L->next_ch.data(), L->N);
    if (prev_rc) memcpy(prev_rc, L->prev_rc.data(), L->N);
    if (vtx_off) memcpy(vtx_off, L->vtx_off.data(), sizeof(int64_t) * (L->V + 1));
    if (occ_g) memcpy(occ_g, L->occ_g.data(), sizeof(int64_t) * L->N);
    if (chr_len)
        for (int32_t c = 0; c < L->C; c++) chr_len[c] = (int64_t)L->seq[c].size();
}

// =================================================================================================
// Seeds: BlocksFinder::FindBlocks bundle enumeration, blocksfinder.h:461-503, sort :517
// =================================================================================================
extern "C" int64_t lcbo_enumerate_seeds(lcbo *L)
{
    L->seed.clear();
    for (int64_t v = -L->V + 1; v < L->V; v++) {
        std::set<char> good;
        std::map<char, size_t> count;
        int64_t av = llabs(v);
        for (int64_t o = L->vtx_off[av]; o < L->vtx_off[av + 1]; o++) {
            int64_t g = L->occ_g[o];
            bool pos = L->pos_id[g] == v;
            char c = (char)L->GetChar(g, pos);
            if (pos) good.insert(c);
            count[c] += 1;
        }
        for (auto p : count) {
            if (p.second > 1 && good.count(p.first)) {
                Seed b{v, (uint8_t)p.first, p.second, 0, SIZE_MAX, SIZE_MAX};
                uint64_t base = 1;
                for (int64_t o = L->vtx_off[av]; o < L->vtx_off[av + 1]; o++) {
                    int64_t g = L->occ_g[o];
                    bool pos = L->pos_id[g] == v;
                    if ((char)L->GetChar(g, pos) == p.first) {
                        b.rank += (uint64_t)L->pos_chr[g] * base;
                        base *= 31;
                        if (pos) {
                            std::pair<size_t, size_t> r(L->pos_bp[g], L->pos_chr[g]);
                            if (r < std::make_pair((size_t)b.res_pos, (size_t)b.res_chr)) {
                                b.res_pos = r.first;
                                b.res_chr = r.second;
                            }
                        }
                    }
                }
                L->seed.push_back(b);
            }
        }
    }
    std::sort(L->seed.begin(), L->seed.end());
    return (int64_t)L->seed.size();
}

extern "C" void lcbo_get_seeds(const lcbo *L, int64_t *vid, uint8_t *ch, uint64_t *count, uint64_t *rank,
                               uint64_t *res_pos, uint64_t *res_chr)
{
    for (size_t i = 0; i < L->seed.size(); i++) {
        const Seed &s = L->seed[i];
        if (vid) vid[i] = s.vid;
        if (ch) ch[i] = s.ch;
        if (count) count[i] = s.count;
        if (rank) rank[i] = s.rank;
        if (res_pos) res_pos[i] = s.res_pos;
        if (res_chr) res_chr[i] = s.res_chr;
    }
}

// =================================================================================================
// Phases + ordered commit: ProcessVertex::operator(), blocksfinder.h:334-433
// =================================================================================================
extern "C" int64_t lcbo_find_blocks(lcbo *L, int min_block, int max_branch, int max_flank, int looking_depth,
                                    int phase_size)
{
    if (L->seed.empty()) lcbo_enumerate_seeds(L);
    L->min_block = min_block;
    L->max_branch = max_branch;
    L->max_flank = max_flank;
    L->looking_depth = looking_depth;
    L->distance.assign((size_t)L->V * 2 + 2, INT_MAX);
    L->count.assign((size_t)L->V * 2 + 2, 0);
    L->order.assign(L->C, std::vector<int>());
    std::fill(L->used.begin(), L->used.end(), 0);
    L->blocks.clear();
    L->blocks_found = 0;
    memset(L->ctr, 0, sizeof L->ctr);
    std::vector<std::vector<Inst>> result(phase_size);
    std::set<size_t> invalid_chr;
    size_t S = L->seed.size();
    for (size_t phase = 0; phase < S; phase += phase_size) {
        size_t limit = std::min(S, phase + (size_t)phase_size);
        for (size_t i = phase; i < limit; i++) L->Process(L->seed[i], result[i - phase]); // parallel part, :345-367
        for (size_t i = phase; i < limit; i++) {                                          // thread 0, :372-414
            std::vector<Inst> &instance = result[i - phase];
            if (instance.size() > 1) {
                bool is_good = true;
                for (const Inst &a : instance) {
                    if (invalid_chr.count(a.chr) == 0) continue;
                    for (int64_t it = a.fg; it != a.bg; it = lcbo::Step(it, a.pos)) {
                        if (L->IsUsed(it, a.pos)) {
                            is_good = false;
                            break;
                        }
                    }
                    if (!is_good) break;
                }
                if (is_good) {
                    L->Finalize(instance, invalid_chr);
                } else {
                    L->ctr[5]++;
                    L->Process(L->seed[i], instance);
                    if (instance.size() > 1) L->Finalize(instance, invalid_chr);
                }
            }
        }
        invalid_chr.clear();
    }
    if (getenv("LCBO_MAXIMA"))
        fprintf(stderr, "oracle maxima: path %llu vertices, %llu instances, %llu good instances\n",
                (unsigned long long)L->mx[0], (unsigned long long)L->mx[1], (unsigned long long)L->mx[2]);
    return (int64_t)L->blocks.size();
}

extern "C" void lcbo_get_blocks(const lcbo *L, int32_t *id, uint32_t *chr, uint64_t *start, uint64_t *end)
{
    for (size_t i = 0; i < L->blocks.size(); i++) {
        id[i] = L->blocks[i].id;
        chr[i] = (uint32_t)L->blocks[i].chr;
        start[i] = L->blocks[i].start;
        end[i] = L->blocks[i].end;
    }
}

extern "C" void lcbo_get_counters(const lcbo *L, uint64_t *c) { memcpy(c, L->ctr, sizeof L->ctr); }

// =================================================================================================
// Output: GenerateOutput blocksfinder.h:605-670; ListBlocksIndicesGFF blocksfinder.cpp:141-174;
// ListBlocksSequences blocksfinder.h:533-582; CalculateCoverage blocksfinder.cpp:109-124
// =================================================================================================
namespace {

struct SortByMultiplicity { // blocksfinder.h:584-603
    const std::vector<int> &multiplicity;
    bool operator()(const Block &a, const Block &b) const
    {
        int m1 = multiplicity[a.BlockId()], m2 = multiplicity[b.BlockId()];
        if (m1 != m2) return m1 > m2;
        return a.BlockId() < b.BlockId();
    }
};

template <class F>
void GroupBy(std::vector<Block> &store, F pred, std::vector<std::pair<size_t, size_t>> &out) // blocksfinder.h:100-110
{
    std::sort(store.begin(), store.end(), pred);
    for (size_t now = 0; now < store.size();) {
        size_t prev = now;
        for (; now < store.size() && !pred(store[prev], store[now]); now++)
            ;
        out.push_back(std::make_pair(prev, now));
    }
}

} // namespace

extern "C" int lcbo_generate_output(lcbo *L, const char *outdir, int gen_seq, int chunks, int min_block,
                                    int64_t *blocks_found_out, double *coverage_out, char *err, int errlen)
{
    try {
        std::vector<std::vector<bool>> covered(L->C);
        for (int32_t i = 0; i < L->C; i++) covered[i].assign(L->seq[i].size() + 1, false);
        int64_t trimmed_id = 1;
        std::vector<std::pair<size_t, size_t>> group;
        std::vector<Block> buffer, trimmed;
        std::vector<int> copies(L->blocks_found + 1, 0);
        std::vector<Block> bi = L->blocks; // keep L->blocks in commit order for lcbo_get_blocks
        for (auto &b : bi) copies[b.BlockId()]++;
        GroupBy(bi, SortByMultiplicity{copies}, group);
        for (auto g : group) {
            buffer.clear();
            for (size_t i = g.first; i < g.second; i++) {
                size_t chr = bi[i].chr, start = bi[i].start, end = bi[i].end;
                for (; covered[chr][start] && start < end; start++)
                    ;
                for (; covered[chr][end] && end > start; end--)
                    ;
                if (end - start >= (size_t)min_block) {
                    buffer.push_back(Block{(int)(bi[i].Sign() * trimmed_id), start, end, chr});
                    std::fill(covered[chr].begin() + start, covered[chr].begin() + end, true);
                }
            }
            if (buffer.size() > 1) {
                trimmed_id++;
                for (auto &b : buffer) trimmed.push_back(b);
            } else {
                for (auto &b : buffer) std::fill(covered[b.chr].begin() + b.start, covered[b.chr].begin() + b.end, false);
            }
        }
        size_t total = 0, total_block = 0;
        for (int32_t i = 0; i < L->C; i++) total += L->seq[i].size();
        for (auto &b : trimmed) total_block += b.Length();
        if (blocks_found_out) *blocks_found_out = trimmed_id - 1;
        if (coverage_out) *coverage_out = total ? double(total_block) / total : 0.0;
        std::sort(trimmed.begin(), trimmed.end());
        if (mkdir(outdir, 0755) != 0 && errno != EEXIST) throw std::runtime_error(std::string("Cannot create dir ") + outdir);
        {
            std::string fn = std::string(outdir) + "/blocks_coords.gff";
            std::ofstream out(fn.c_str());
            if (!out) throw std::runtime_error("Cannot open file " + fn);
            std::vector<Block> block(trimmed);
            std::sort(block.begin(), block.end(), CompareById);
            out << "##gff-version 3.1.26\n";
            for (int32_t i = 0; i < L->C; i++) out << "##sequence-region " << L->header[i] << " 1 " << L->seq[i].size() << "\n";
            for (auto &b : block)
                out << L->header[b.chr] << "\tSibeliaZ\tSO:0000856\t" << b.start + 1 << "\t" << b.end << "\t.\t"
                    << (b.id > 0 ? "+" : "-") << "\t.\tID=" << (size_t)b.BlockId() << "\n";
        }
        if (gen_seq) {
            std::vector<std::ofstream> chunk_out(chunks);
            for (int i = 0; i < chunks; i++) {
                std::string fn = std::string(outdir) + "/" + std::to_string(i) + ".tmp";
                chunk_out[i].open(fn.c_str());
                if (!chunk_out[i]) throw std::runtime_error("Cannot open file " + fn);
            }
            std::vector<Block> bl(trimmed);
            std::vector<std::pair<size_t, size_t>> grp;
            GroupBy(bl, CompareById, grp);
            size_t now_chunk = 0;
            for (auto &gr : grp) {
                std::ofstream &out = chunk_out[now_chunk];
                for (size_t b = gr.first; b < gr.second; b++) {
                    size_t length = bl[b].Length(), chr = bl[b].chr, chr_size = L->seq[chr].size();
                    out << "> " << L->header[chr] << ";";
                    if (bl[b].id > 0) {
                        out << bl[b].start << ";" << length << ";+;" << chr_size << '@';
                        out.write(L->seq[chr].data() + bl[b].start, (std::streamsize)length);
                    } else {
                        size_t start = chr_size - bl[b].end;
                        out << start << ";" << length << ";-;" << chr_size << '@';
                        for (size_t i = 0; i < length; i++) out << (char)ReverseChar((uint8_t)L->seq[chr][bl[b].end - 1 - i]);
                    }
                    out << '@';
                }
                out << "\n";
                now_chunk = (now_chunk + 1) % chunks;
            }
        }
    } catch (std::exception &e) {
        if (err && errlen > 0) snprintf(err, errlen, "%s", e.what());
        return 1;
    }
    return 0;
}


#ifdef LCBO_EPOCH
// -------------------------------------------------------------------------------------------------
// Epoch-threshold evaluation of one seed (test infrastructure for the speculative-round protocol used by the
// CUDA path): edge e is "used" iff epoch[e] < thresh.  Returns the best instances as
// (front g | strand bit 62, back g) pairs and the read-set as sorted, merged [lo, hi] intervals.
// -------------------------------------------------------------------------------------------------
extern "C" void lcbo_epoch_prepare(lcbo *L, int min_block, int max_branch, int max_flank, int looking_depth)
{
    if (L->seed.empty()) lcbo_enumerate_seeds(L);
    L->min_block = min_block;
    L->max_branch = max_branch;
    L->max_flank = max_flank;
    L->looking_depth = looking_depth;
    L->distance.assign((size_t)L->V * 2 + 2, INT_MAX);
    L->count.assign((size_t)L->V * 2 + 2, 0);
    L->order.assign(L->C, std::vector<int>());
}

extern "C" int lcbo_epoch_process(lcbo *L, int64_t seed_index, uint32_t thresh, const uint32_t *epoch, int64_t *inst_out,
                                  int inst_cap, int64_t *rs_out, int rs_cap, int *n_rs)
{
    std::vector<int64_t> log;
    std::vector<Inst> best;
    L->epoch = epoch;
    L->thresh = thresh;
    L->readlog = &log;
    L->Process(L->seed[(size_t)seed_index], best);
    L->readlog = nullptr;
    std::sort(log.begin(), log.end());
    int m = 0;
    for (size_t i = 0; i < log.size();) {
        size_t j = i;
        while (j + 1 < log.size() && log[j + 1] <= log[j] + 1) j++;
        if (m < rs_cap) {
            rs_out[2 * m] = log[i];
            rs_out[2 * m + 1] = log[j];
        }
        m++;
        i = j + 1;
    }
    *n_rs = m;
    int n = 0;
    for (const Inst &a : best) {
        if (n < inst_cap) {
            inst_out[2 * n] = a.fg | (a.pos ? (int64_t)1 << 62 : 0);
            inst_out[2 * n + 1] = a.bg;
        }
        n++;
    }
    return n;
}
#endif

#ifdef LCB_ORACLE_MAIN
// Minimal CLI used by tests and by hand: lcb_oracle <graph> <k> <b> <m> <a> <outdir> <noseq 0|1> <chunks> <fasta...>
int main(int argc, char **argv)
{
    if (argc < 10) {
        fprintf(stderr, "usage: %s graph k b m a outdir noseq chunks fasta...\n", argv[0]);
        return 2;
    }
    char err[512] = {0};
    int k = atoi(argv[2]), b = atoi(argv[3]), m = atoi(argv[4]), a = atoi(argv[5]);
    lcbo *L = lcbo_load(argv[1], argv + 9, argc - 9, k, a, err, sizeof err);
    if (!L) {
        fprintf(stderr, "error: %s\n", err);
        return 1;
    }
    int64_t S = lcbo_enumerate_seeds(L);
    int64_t nb = lcbo_find_blocks(L, m, b, b, 8, 256);
    int64_t found = 0;
    double cov = 0;
    if (lcbo_generate_output(L, argv[6], !atoi(argv[7]), atoi(argv[8]), m, &found, &cov, err, sizeof err)) {
        fprintf(stderr, "error: %s\n", err);
        return 1;
    }
    if (const char *dump = getenv("LCBO_DUMP_BLOCKS")) { // raw blocksInstance_ in commit order, for experiments
        FILE *f = fopen(dump, "w");
        for (auto &bk : L->blocks) fprintf(f, "%d %zu %zu %zu\n", bk.id, bk.chr, bk.start, bk.end);
        fclose(f);
    }
    uint64_t c[8];
    lcbo_get_counters(L, c);
    printf("records %lld vertices %lld seeds %lld instances %lld\n", (long long)lcbo_num_records(L),
           (long long)lcbo_num_vertices(L), (long long)S, (long long)nb);
    printf("T_walk %llu T_occ %llu T_scan %llu T_score %llu process %llu reruns %llu mpv %llu pushes %llu\n",
           (unsigned long long)c[0], (unsigned long long)c[1], (unsigned long long)c[2], (unsigned long long)c[3],
           (unsigned long long)c[4], (unsigned long long)c[5], (unsigned long long)c[6], (unsigned long long)c[7]);
    printf("Blocks found: %lld\nCoverage: %.2f\n", (long long)found, cov);
    lcbo_free(L);
    return 0;
}
#endif
