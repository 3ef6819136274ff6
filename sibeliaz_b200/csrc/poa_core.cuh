// poa_core.cuh -- partial order alignment of one block (the stage after the LCB path):
// what `spoa <block.fa> -l 1 -r 1 -e -8` computes (SibeliaZ-LCB/sibeliaz:66): every copy of the block is aligned globally
// (Needleman-Wunsch, linear gaps) against the partial order graph of the copies before it, merged into the graph, and
// the rows of the multiple sequence alignment are read off the final graph.  Reference: spoa/src/graph.cpp
// (AddAlignment :156-246, TopologicalSort :248-301, MSA :303-357) and spoa/src/sisd_alignment_engine.cpp
// (Initialize :118-257, Linear :295-456); every function below cites what it follows.
//
// One WARP works on one block.  The graph lives in flat index arrays inside a per-warp arena in HBM (no pointers, no
// allocation); the sequential parts (graph update, topological sort, traceback) are executed by lane 0, the dynamic
// programme -- where the time goes -- by the whole warp: one lane per column, 32 columns per step, the in-row gap
// recurrence H[j] = max(M[j], H[j-1] + g) as a max-scan of M[j] - j*g (exact in integers).
//
// The same source compiles for the host (no __CUDA_ARCH__): there a "warp" is one thread and the row kernel is a plain
// loop.  tests/ use that build to check this very code against the CPU checker without a GPU; the product path is the
// kernel (poa_device.cu).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define POA_HD __host__ __device__ __forceinline__
#else
#define POA_HD inline
#endif
// The warp-level code is compiled for the device -- and for the host under POA_WARP_EMULATION, where tests/poa_warp_emu.cpp
// supplies __shfl_up_sync / __shfl_sync / __syncwarp for 32 host threads in lockstep, so that the row kernel below is
// exercised exactly as written without a GPU.
#if defined(__CUDA_ARCH__) || defined(POA_WARP_EMULATION)
#define POA_WARP_CODE 1
#endif
#if defined(__CUDA_ARCH__)
#define POA_DEV __device__ __forceinline__
#else
#define POA_DEV inline
#endif

namespace poa {

constexpr int32_t kNegInf = INT32_MIN + 1024; // sisd_alignment_engine.cpp:13-14
constexpr int kMaxCodes = 32;                // distinct characters in one block (DNA: <= 5 + case)

struct Params {
    int m, n, g; // match, mismatch, gap (linear: -g and -e of the command line coincide)
};

// Capacities of one arena; the work of a block fits iff  nodes <= max_nodes, edges <= max_edges, ... (checked as it grows)
struct Caps {
    uint32_t max_nodes;   // graph nodes
    uint32_t max_edges;   // graph edges
    uint32_t max_aligned; // entries of all aligned_nodes lists together
    uint32_t max_path;    // sum of the copies' lengths (nodes of all sequence paths)
    uint32_t max_stack;   // DFS stack of the topological sort
    uint64_t max_cells;   // H: (nodes + 1) * (len + 1) of the largest alignment
    uint32_t max_align;   // alignment pairs: nodes + len
};

struct Work { // views into the arena (see poa_bind); all indices 32-bit
    // graph
    uint8_t *code;                       // [max_nodes]
    int32_t *in_first, *in_last;         // [max_nodes] edge lists in insertion order (-1: empty)
    int32_t *out_first, *out_last;       // [max_nodes]
    int32_t *al_first, *al_last;         // [max_nodes] aligned_nodes lists (entries in al_node / al_next)
    int32_t *edge_tail, *edge_head;      // [max_edges]
    int32_t *edge_next_in, *edge_next_out;
    int32_t *al_node, *al_next;          // [max_aligned]
    int32_t *path;                       // [max_path] the nodes of every copy, copy after copy
    uint32_t *path_off;                  // [copies + 1]
    // order
    int32_t *rank_to_node;               // [max_nodes]
    uint32_t *rank;                      // [max_nodes] node -> rank
    uint8_t *marks;                      // [max_nodes] bits 0-1 mark, bit 2 ignored
    int32_t *stack;                      // [max_stack]
    uint32_t *column;                    // [max_nodes] MSA column of a node
    // alignment
    int32_t *H;                          // [max_cells]
    int32_t *al_pairs;                   // [2 * max_align] (node, position) pairs, back to front
    // counts
    uint32_t n_nodes, n_edges, n_aligned, n_path, n_copies, n_codes, n_pairs, n_columns;
    uint8_t decoder[kMaxCodes];
    Caps cap;
    int err; // 0, or 1 = a capacity was exceeded (the caller retries the block in a bigger arena)
};

// ---- arena layout ------------------------------------------------------------------------------------------------
POA_HD uint64_t poa_align_up(uint64_t x) { return (x + 15) & ~(uint64_t)15; }

POA_HD uint64_t poa_arena_bytes(const Caps &c, uint32_t max_copies)
{
    uint64_t b = 0;
    b += poa_align_up(c.max_nodes);                          // code
    b += 6 * poa_align_up(4ull * c.max_nodes);               // in/out/al first+last
    b += 4 * poa_align_up(4ull * c.max_edges);               // edge arrays
    b += 2 * poa_align_up(4ull * c.max_aligned);             // aligned pool
    b += poa_align_up(4ull * c.max_path);                    // path
    b += poa_align_up(4ull * (max_copies + 1));              // path_off
    b += 3 * poa_align_up(4ull * c.max_nodes);               // rank_to_node, rank, column
    b += poa_align_up(c.max_nodes);                          // marks
    b += poa_align_up(4ull * c.max_stack);                   // stack
    b += poa_align_up(4ull * c.max_cells);                   // H
    b += poa_align_up(8ull * c.max_align);                   // alignment pairs
    return b;
}

POA_HD void poa_bind(Work &w, uint8_t *p, const Caps &c, uint32_t max_copies)
{
    w.cap = c;
    w.code = p, p += poa_align_up(c.max_nodes);
    w.in_first = (int32_t *)p, p += poa_align_up(4ull * c.max_nodes);
    w.in_last = (int32_t *)p, p += poa_align_up(4ull * c.max_nodes);
    w.out_first = (int32_t *)p, p += poa_align_up(4ull * c.max_nodes);
    w.out_last = (int32_t *)p, p += poa_align_up(4ull * c.max_nodes);
    w.al_first = (int32_t *)p, p += poa_align_up(4ull * c.max_nodes);
    w.al_last = (int32_t *)p, p += poa_align_up(4ull * c.max_nodes);
    w.edge_tail = (int32_t *)p, p += poa_align_up(4ull * c.max_edges);
    w.edge_head = (int32_t *)p, p += poa_align_up(4ull * c.max_edges);
    w.edge_next_in = (int32_t *)p, p += poa_align_up(4ull * c.max_edges);
    w.edge_next_out = (int32_t *)p, p += poa_align_up(4ull * c.max_edges);
    w.al_node = (int32_t *)p, p += poa_align_up(4ull * c.max_aligned);
    w.al_next = (int32_t *)p, p += poa_align_up(4ull * c.max_aligned);
    w.path = (int32_t *)p, p += poa_align_up(4ull * c.max_path);
    w.path_off = (uint32_t *)p, p += poa_align_up(4ull * (max_copies + 1));
    w.rank_to_node = (int32_t *)p, p += poa_align_up(4ull * c.max_nodes);
    w.rank = (uint32_t *)p, p += poa_align_up(4ull * c.max_nodes);
    w.column = (uint32_t *)p, p += poa_align_up(4ull * c.max_nodes);
    w.marks = p, p += poa_align_up(c.max_nodes);
    w.stack = (int32_t *)p, p += poa_align_up(4ull * c.max_stack);
    w.H = (int32_t *)p, p += poa_align_up(4ull * c.max_cells);
    w.al_pairs = (int32_t *)p;
}

// Capacities for a block whose copies have `sum_len` characters in total, the longest `max_len`.  Every character adds at
// most one node and one edge; an aligned set holds at most one node per character code.  Level 2 is certainly enough;
// levels 0 and 1 are optimistic about the two big items (copies of a block are similar, so the graph has about as many
// nodes as the longest copy has characters; DNA has few codes) -- a block that outgrows them reports err = 1 and is run
// again one level up.
POA_HD Caps poa_caps_for(uint64_t sum_len, uint64_t max_len, int level)
{
    Caps c;
    const uint64_t codes = level >= 2 ? (uint64_t)(kMaxCodes - 1) : 7;
    const uint64_t rows = level == 0 ? 2 * (max_len + 1) : (level == 1 ? 4 * (max_len + 1) : sum_len + 1);
    c.max_nodes = (uint32_t)sum_len + 1;
    c.max_edges = (uint32_t)sum_len + 1;
    c.max_aligned = (uint32_t)(sum_len * codes) + 1;
    c.max_path = (uint32_t)sum_len + 1;
    c.max_stack = (uint32_t)(3 * (sum_len + 1) + 2 * (uint64_t)c.max_aligned + 2);
    c.max_cells = (rows < sum_len + 1 ? rows : sum_len + 1) * (max_len + 1);
    c.max_align = (uint32_t)(sum_len + max_len + 2);
    return c;
}

// ---- graph (lane 0 / host) -------------------------------------------------------------------------------------------
POA_HD void graph_reset(Work &w)
{
    w.n_nodes = w.n_edges = w.n_aligned = w.n_path = w.n_copies = w.n_codes = 0;
    w.path_off[0] = 0;
    w.err = 0;
}

POA_HD int code_of(Work &w, uint8_t ch) // Graph::coder_ (graph.cpp:170-175): codes in order of first appearance
{
    for (uint32_t i = 0; i < w.n_codes; i++)
        if (w.decoder[i] == ch) return (int)i;
    if (w.n_codes >= (uint32_t)kMaxCodes) {
        w.err = 1;
        return 0;
    }
    w.decoder[w.n_codes] = ch;
    return (int)w.n_codes++;
}

POA_HD int add_node(Work &w, int c) // graph.cpp:78-81
{
    if (w.n_nodes >= w.cap.max_nodes) {
        w.err = 1;
        return 0;
    }
    const int id = (int)w.n_nodes++;
    w.code[id] = (uint8_t)c;
    w.in_first[id] = w.in_last[id] = w.out_first[id] = w.out_last[id] = w.al_first[id] = w.al_last[id] = -1;
    return id;
}

POA_HD void add_edge(Work &w, int tail, int head) // graph.cpp:83-93 (labels / weights serve the consensus only)
{
    for (int e = w.out_first[tail]; e >= 0; e = w.edge_next_out[e])
        if (w.edge_head[e] == head) return;
    if (w.n_edges >= w.cap.max_edges) {
        w.err = 1;
        return;
    }
    const int e = (int)w.n_edges++;
    w.edge_tail[e] = tail, w.edge_head[e] = head, w.edge_next_in[e] = w.edge_next_out[e] = -1;
    if (w.out_last[tail] < 0) w.out_first[tail] = e;
    else w.edge_next_out[w.out_last[tail]] = e;
    w.out_last[tail] = e;
    if (w.in_last[head] < 0) w.in_first[head] = e;
    else w.edge_next_in[w.in_last[head]] = e;
    w.in_last[head] = e;
}

POA_HD void aligned_push(Work &w, int node, int other) // node->aligned_nodes.emplace_back(other)
{
    if (w.n_aligned >= w.cap.max_aligned) {
        w.err = 1;
        return;
    }
    const int x = (int)w.n_aligned++;
    w.al_node[x] = other, w.al_next[x] = -1;
    if (w.al_last[node] < 0) w.al_first[node] = x;
    else w.al_next[w.al_last[node]] = x;
    w.al_last[node] = x;
}

// graph.cpp:95-112: a chain of new nodes for s[begin, end); appends them to path at `at`; returns the first node or -1
POA_HD int add_chain(Work &w, const uint8_t *s, uint32_t begin, uint32_t end, uint32_t at)
{
    if (begin == end) return -1;
    int prev = -1, first = -1;
    for (uint32_t i = begin; i < end && !w.err; i++) {
        const int curr = add_node(w, code_of(w, s[i]));
        if (first < 0) first = curr;
        if (prev >= 0) add_edge(w, prev, curr);
        prev = curr;
        w.path[at + (i - begin)] = curr;
    }
    return first;
}

// graph.cpp:248-301: depth first, predecessors before a node, a node's aligned set right behind it
POA_HD void topological_sort(Work &w)
{
    const uint32_t n = w.n_nodes;
    for (uint32_t i = 0; i < n; i++) w.marks[i] = 0;
    uint32_t nr = 0, sp = 0;
    for (uint32_t s = 0; s < n && !w.err; s++) {
        if ((w.marks[s] & 3) != 0) continue;
        w.stack[sp++] = (int32_t)s;
        while (sp) {
            const int curr = w.stack[sp - 1];
            bool valid = true;
            if ((w.marks[curr] & 3) != 2) {
                for (int e = w.in_first[curr]; e >= 0; e = w.edge_next_in[e]) {
                    const int t = w.edge_tail[e];
                    if ((w.marks[t] & 3) != 2) {
                        if (sp >= w.cap.max_stack) {
                            w.err = 1;
                            return;
                        }
                        w.stack[sp++] = t, valid = false;
                    }
                }
                if (!(w.marks[curr] & 4)) {
                    for (int x = w.al_first[curr]; x >= 0; x = w.al_next[x]) {
                        const int a = w.al_node[x];
                        if ((w.marks[a] & 3) != 2) {
                            if (sp >= w.cap.max_stack) {
                                w.err = 1;
                                return;
                            }
                            w.stack[sp++] = a, w.marks[a] |= 4, valid = false;
                        }
                    }
                }
                if (valid) {
                    w.marks[curr] = (uint8_t)((w.marks[curr] & 4) | 2);
                    if (!(w.marks[curr] & 4)) {
                        w.rank_to_node[nr++] = curr;
                        for (int x = w.al_first[curr]; x >= 0; x = w.al_next[x]) w.rank_to_node[nr++] = w.al_node[x];
                    }
                } else {
                    w.marks[curr] = (uint8_t)((w.marks[curr] & 4) | 1);
                }
            }
            if (valid) sp--;
        }
    }
    for (uint32_t i = 0; i < nr; i++) w.rank[w.rank_to_node[i]] = i; // sisd_alignment_engine.cpp:134-137
}

// graph.cpp:156-246.  `pairs`: n_pairs (node, position) pairs (-1 = gap) stored BACK TO FRONT as traceback() leaves them
// (the reference reverses the vector, sisd_alignment_engine.cpp:454); s: the copy
POA_HD void add_alignment(Work &w, const int32_t *pairs, uint32_t n_pairs, const uint8_t *s, uint32_t len)
{
    if (len == 0) {
        w.path_off[w.n_copies + 1] = w.n_path;
        w.n_copies++;
        return;
    }
    if (w.n_path + len > w.cap.max_path) {
        w.err = 1;
        return;
    }
    for (uint32_t i = 0; i < len; i++) code_of(w, s[i]); // codes in order of appearance in the copy (:170-175)
    const uint32_t base = w.n_path;
    if (n_pairs == 0) {
        add_chain(w, s, 0, len, base);
    } else {
        int32_t vfront = -1, vback = -1;
        for (uint32_t q = n_pairs; q-- > 0;)
            if (pairs[2 * q + 1] != -1) {
                if (vfront < 0) vfront = pairs[2 * q + 1];
                vback = pairs[2 * q + 1];
            }
        // add unaligned bases (:197-200)
        int begin = add_chain(w, s, 0, (uint32_t)vfront, base);
        int prev = begin >= 0 ? (int)w.n_nodes - 1 : -1;
        const int last = add_chain(w, s, (uint32_t)vback + 1, len, base + (uint32_t)vback + 1);
        // add aligned bases (:202-240)
        for (uint32_t q = n_pairs; q-- > 0 && !w.err;) {
            const int32_t node = pairs[2 * q], pos = pairs[2 * q + 1];
            if (pos == -1) continue;
            const int c = code_of(w, s[pos]);
            int curr = -1;
            if (node == -1) {
                curr = add_node(w, c);
            } else if (w.code[node] == c) {
                curr = node;
            } else {
                for (int x = w.al_first[node]; x >= 0; x = w.al_next[x])
                    if (w.code[w.al_node[x]] == c) {
                        curr = w.al_node[x];
                        break;
                    }
                if (curr < 0) {
                    curr = add_node(w, c);
                    for (int x = w.al_first[node]; x >= 0 && !w.err; x = w.al_next[x]) {
                        const int kt = w.al_node[x];
                        aligned_push(w, kt, curr);
                        aligned_push(w, curr, kt);
                    }
                    aligned_push(w, node, curr);
                    aligned_push(w, curr, node);
                }
            }
            if (begin < 0) begin = curr;
            if (prev >= 0) add_edge(w, prev, curr);
            prev = curr;
            w.path[base + (uint32_t)pos] = curr;
        }
        if (last >= 0) add_edge(w, prev, last);
    }
    w.n_path = base + len;
    w.path_off[w.n_copies + 1] = w.n_path;
    w.n_copies++;
    if (!w.err) topological_sort(w);
}

// graph.cpp:303-317: one column per rank, aligned nodes share it; returns the number of columns
POA_HD uint32_t msa_columns(Work &w)
{
    uint32_t j = 0;
    for (uint32_t i = 0; i < w.n_nodes; ++i, ++j) {
        const int it = w.rank_to_node[i];
        w.column[it] = j;
        for (int x = w.al_first[it]; x >= 0; x = w.al_next[x]) w.column[w.al_node[x]] = j, ++i;
    }
    return j;
}

// ---- dynamic programme ---------------------------------------------------------------------------------------------------
// Column 0 and row 0 (sisd_alignment_engine.cpp:176-178, :213-225).  Uniform; on the device the lanes stride the columns.
POA_HD void dp_init(Work &w, const Params &pr, uint32_t len, int lane, int lanes)
{
    const uint64_t W = (uint64_t)len + 1;
    for (uint64_t j = (uint64_t)lane; j < W; j += (uint64_t)lanes) w.H[j] = (int32_t)j * pr.g;
    if (lane == 0) {
        for (uint32_t i = 1; i <= w.n_nodes; i++) {
            const int it = w.rank_to_node[i - 1];
            int32_t penalty = w.in_first[it] < 0 ? 0 : kNegInf;
            for (int e = w.in_first[it]; e >= 0; e = w.edge_next_in[e]) {
                const int32_t v = w.H[((uint64_t)w.rank[w.edge_tail[e]] + 1) * W];
                penalty = v > penalty ? v : penalty;
            }
            w.H[(uint64_t)i * W] = penalty + pr.g;
        }
    }
#if defined(POA_WARP_CODE)
    __syncwarp();
#endif
}

#if defined(POA_WARP_CODE)
// One row of H (sisd_alignment_engine.cpp:318-352), the whole warp.  A step covers 4 x 32 columns: lane l owns columns
// base + 32 k + l + 1 (k = 0..3), so that the loads of four chunks are in flight together; the four max-scans then run one
// after the other because each needs the carry of the one before.  The first two predecessor rows are kept as pointers
// (almost every node has one or two in-edges), further ones are reached through the edge list.
POA_DEV void dp_row(Work &w, const Params &pr, const uint8_t *s, uint32_t len, uint32_t i, int lane)
{
    const uint64_t W = (uint64_t)len + 1;
    const int it = w.rank_to_node[i - 1];
    const uint8_t ch = w.decoder[w.code[it]];
    int32_t *row = w.H + (uint64_t)i * W;
    const int e0 = w.in_first[it];
    const int e1 = e0 >= 0 ? w.edge_next_in[e0] : -1;
    const int e2 = e1 >= 0 ? w.edge_next_in[e1] : -1;
    const int32_t *pred0 = e0 >= 0 ? w.H + ((uint64_t)w.rank[w.edge_tail[e0]] + 1) * W : w.H; // no predecessor: row 0 (:321-323)
    const int32_t *pred1 = e1 >= 0 ? w.H + ((uint64_t)w.rank[w.edge_tail[e1]] + 1) * W : nullptr;
    int32_t carry = row[0]; // H[i][base] - base * g with base = 0
    for (uint32_t base = 0; base < len; base += 128) {
        int32_t x[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t j = base + 32u * (uint32_t)k + (uint32_t)lane + 1;
            int32_t M = kNegInf;
            if (j <= len) {
                const int32_t sc = s[j - 1] == ch ? pr.m : pr.n;
                const int32_t a = pred0[j - 1] + sc, b = pred0[j] + pr.g;
                M = a > b ? a : b;
                if (pred1) {
                    const int32_t a1 = pred1[j - 1] + sc, b1 = pred1[j] + pr.g;
                    const int32_t v1 = a1 > b1 ? a1 : b1;
                    M = v1 > M ? v1 : M;
                    for (int e = e2; e >= 0; e = w.edge_next_in[e]) {
                        const int32_t *pred = w.H + ((uint64_t)w.rank[w.edge_tail[e]] + 1) * W;
                        const int32_t a2 = pred[j - 1] + sc, b2 = pred[j] + pr.g;
                        const int32_t v2 = a2 > b2 ? a2 : b2;
                        M = v2 > M ? v2 : M;
                    }
                }
                M -= (int32_t)j * pr.g;
            }
            x[k] = M;
        }
        // H[j] = max(M[j], H[j-1] + g)  <=>  H[j] - j g = max(M[j] - j g, H[j-1] - (j-1) g): inclusive max-scan + carry
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t j = base + 32u * (uint32_t)k + (uint32_t)lane + 1;
            int32_t v = x[k];
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t y = __shfl_up_sync(0xFFFFFFFFu, v, d);
                if (lane >= d) v = y > v ? y : v;
            }
            v = v > carry ? v : carry;
            if (j <= len) row[j] = v + (int32_t)j * pr.g;
            carry = __shfl_sync(0xFFFFFFFFu, v, 31);
        }
    }
    __syncwarp();
}
#else
inline void dp_row(Work &w, const Params &pr, const uint8_t *s, uint32_t len, uint32_t i, int)
{
    const uint64_t W = (uint64_t)len + 1;
    const int it = w.rank_to_node[i - 1];
    const uint8_t ch = w.decoder[w.code[it]];
    int32_t *row = w.H + (uint64_t)i * W;
    for (uint32_t j = 1; j <= len; j++) {
        const int32_t sc = s[j - 1] == ch ? pr.m : pr.n;
        int32_t M = kNegInf;
        if (w.in_first[it] < 0) {
            const int32_t a = w.H[j - 1] + sc, b = w.H[j] + pr.g;
            M = a > b ? a : b;
        } else {
            for (int e = w.in_first[it]; e >= 0; e = w.edge_next_in[e]) {
                const int32_t *pred = w.H + ((uint64_t)w.rank[w.edge_tail[e]] + 1) * W;
                const int32_t a = pred[j - 1] + sc, b = pred[j] + pr.g;
                const int32_t v = a > b ? a : b;
                M = v > M ? v : M;
            }
        }
        const int32_t h = row[j - 1] + pr.g;
        row[j] = h > M ? h : M;
    }
}
#endif

// The end of the global alignment: the first sink (in rank order) with the largest score in the last column (:353-356),
// then the backtrack (:374-452): diagonal through the in-edges in their order, then vertical, then horizontal.
// Lane 0 / host.  Returns the number of pairs written to w.al_pairs back to front (pair k at [2k], [2k+1]).
POA_HD uint32_t traceback(Work &w, const Params &pr, const uint8_t *s, uint32_t len)
{
    const uint64_t W = (uint64_t)len + 1;
    int32_t best = kNegInf;
    uint32_t i = 0, j = 0;
    for (uint32_t r = 1; r <= w.n_nodes; r++) {
        const int it = w.rank_to_node[r - 1];
        if (w.out_first[it] >= 0) continue;
        const int32_t v = w.H[(uint64_t)r * W + len];
        if (best < v) best = v, i = r, j = len;
    }
    if (i == 0 && j == 0) return 0;
    uint32_t n = 0, prev_i = 0, prev_j = 0;
    while (!(i == 0 && j == 0)) {
        const int32_t Hij = w.H[(uint64_t)i * W + j];
        bool found = false;
        if (i != 0 && j != 0) {
            const int it = w.rank_to_node[i - 1];
            const int32_t match = s[j - 1] == w.decoder[w.code[it]] ? pr.m : pr.n;
            if (w.in_first[it] < 0) {
                if (Hij == w.H[j - 1] + match) prev_i = 0, prev_j = j - 1, found = true;
            } else {
                for (int e = w.in_first[it]; e >= 0; e = w.edge_next_in[e]) {
                    const uint32_t pi = w.rank[w.edge_tail[e]] + 1;
                    if (Hij == w.H[(uint64_t)pi * W + (j - 1)] + match) {
                        prev_i = pi, prev_j = j - 1, found = true;
                        break;
                    }
                }
            }
        }
        if (!found && i != 0) {
            const int it = w.rank_to_node[i - 1];
            if (w.in_first[it] < 0) {
                if (Hij == w.H[j] + pr.g) prev_i = 0, prev_j = j, found = true;
            } else {
                for (int e = w.in_first[it]; e >= 0; e = w.edge_next_in[e]) {
                    const uint32_t pi = w.rank[w.edge_tail[e]] + 1;
                    if (Hij == w.H[(uint64_t)pi * W + j] + pr.g) {
                        prev_i = pi, prev_j = j, found = true;
                        break;
                    }
                }
            }
        }
        if (!found && j != 0 && Hij == w.H[(uint64_t)i * W + j - 1] + pr.g) prev_i = i, prev_j = j - 1, found = true;
        if (!found || n >= w.cap.max_align) { // cannot happen for a matrix this code filled; never loop forever
            w.err = 2;
            return 0;
        }
        w.al_pairs[2 * n] = i == prev_i ? -1 : w.rank_to_node[i - 1];
        w.al_pairs[2 * n + 1] = j == prev_j ? -1 : (int32_t)j - 1;
        n++;
        i = prev_i, j = prev_j;
    }
    return n;
}

// ---- one block -------------------------------------------------------------------------------------------------------------
POA_HD void poa_sync()
{
#if defined(POA_WARP_CODE)
    __syncwarp();
#endif
}

// main.cpp:282-320 for the copies [c0, c1) of one block: Align + AddAlignment per copy in file order, then the MSA columns.
// `w` is shared by the lanes of the warp (shared memory on the device); lane 0 runs the sequential parts.
// Leaves w.n_columns and w.column / w.path for the row writer; w.err != 0: the arena was too small (1) or a bug (2).
POA_HD void run_block(Work &w, const Params &pr, const uint8_t *seq, const uint64_t *copy_off, uint32_t c0, uint32_t c1, int lane,
                      int lanes)
{
    if (lane == 0) graph_reset(w);
    poa_sync();
    for (uint32_t c = c0; c < c1; c++) {
        const uint8_t *s = seq + copy_off[c];
        const uint32_t len = (uint32_t)(copy_off[c + 1] - copy_off[c]);
        const uint32_t nodes = w.n_nodes;
        if (nodes != 0 && len != 0) { // sisd_alignment_engine.cpp:268-270: otherwise the alignment is empty
            if ((uint64_t)(nodes + 1) * ((uint64_t)len + 1) > w.cap.max_cells || nodes + len + 2 > w.cap.max_align) {
                poa_sync(); // every lane has read w.err (end of the previous copy), w.n_nodes and w.cap before the flag is written
                if (lane == 0) w.err = 1;
                poa_sync();
                return;
            }
            dp_init(w, pr, len, lane, lanes);
            for (uint32_t i = 1; i <= nodes; i++) dp_row(w, pr, s, len, i, lane);
            if (lane == 0) w.n_pairs = traceback(w, pr, s, len);
        } else if (lane == 0) {
            w.n_pairs = 0;
        }
        poa_sync();
        if (lane == 0 && !w.err) add_alignment(w, w.al_pairs, w.n_pairs, s, len);
        poa_sync();
        if (w.err) return;
    }
    if (lane == 0) w.n_columns = msa_columns(w);
    poa_sync();
}

// graph.cpp:324-337: row of copy k (0-based inside the block) into out[0 .. n_columns)
POA_HD void write_row(const Work &w, uint32_t k, uint8_t *out, int lane, int lanes)
{
    for (uint32_t j = (uint32_t)lane; j < w.n_columns; j += (uint32_t)lanes) out[j] = '-';
    poa_sync();
    for (uint32_t t = w.path_off[k] + (uint32_t)lane; t < w.path_off[k + 1]; t += (uint32_t)lanes) {
        const int node = w.path[t];
        out[w.column[node]] = w.decoder[w.code[node]];
    }
    poa_sync();
}

// ---- long blocks: one block per CTA ------------------------------------------------------------------------------------------
// A row of a long block (tens of thousands of columns) is hundreds of independent 128-column pieces: here the warps of a
// whole CTA share one row.  A step covers nwarps x 128 columns; warp v scans its 128 columns without a carry (pass 1),
// publishes the maximum of its piece, and after one CTA barrier every warp folds the carry of the step and the maxima of
// the warps before it into its own columns (pass 2) -- max is associative, so this is the same H as the one-warp scan.
// Thread 0 of the CTA runs the sequential parts.  (Compiled as warp code: device, or host under POA_WARP_EMULATION.)
#if defined(POA_WARP_CODE)
POA_DEV void cta_sync() { __syncthreads(); }

// seg: 2 x 32 ints of shared memory (double-buffered by step parity, so one barrier per step is enough)
POA_DEV void dp_row_cta(Work &w, const Params &pr, const uint8_t *s, uint32_t len, uint32_t i, int lane, int warp, int nwarps,
                        int32_t *seg)
{
    const uint64_t W = (uint64_t)len + 1;
    const int it = w.rank_to_node[i - 1];
    const uint8_t ch = w.decoder[w.code[it]];
    int32_t *row = w.H + (uint64_t)i * W;
    const int e0 = w.in_first[it];
    const int e1 = e0 >= 0 ? w.edge_next_in[e0] : -1;
    const int e2 = e1 >= 0 ? w.edge_next_in[e1] : -1;
    const int32_t *pred0 = e0 >= 0 ? w.H + ((uint64_t)w.rank[w.edge_tail[e0]] + 1) * W : w.H;
    const int32_t *pred1 = e1 >= 0 ? w.H + ((uint64_t)w.rank[w.edge_tail[e1]] + 1) * W : nullptr;
    int32_t carry = row[0]; // of the step: H[i][base] - base * g
    const uint32_t span = 128u * (uint32_t)nwarps;
    int parity = 0;
    for (uint32_t base = 0; base < len; base += span, parity ^= 1) {
        const uint32_t wbase = base + 128u * (uint32_t)warp;
        int32_t x[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t j = wbase + 32u * (uint32_t)k + (uint32_t)lane + 1;
            int32_t M = kNegInf;
            if (j <= len) {
                const int32_t sc = s[j - 1] == ch ? pr.m : pr.n;
                const int32_t a = pred0[j - 1] + sc, b = pred0[j] + pr.g;
                M = a > b ? a : b;
                if (pred1) {
                    const int32_t a1 = pred1[j - 1] + sc, b1 = pred1[j] + pr.g;
                    const int32_t v1 = a1 > b1 ? a1 : b1;
                    M = v1 > M ? v1 : M;
                    for (int e = e2; e >= 0; e = w.edge_next_in[e]) {
                        const int32_t *pred = w.H + ((uint64_t)w.rank[w.edge_tail[e]] + 1) * W;
                        const int32_t a2 = pred[j - 1] + sc, b2 = pred[j] + pr.g;
                        const int32_t v2 = a2 > b2 ? a2 : b2;
                        M = v2 > M ? v2 : M;
                    }
                }
                M -= (int32_t)j * pr.g;
            }
            x[k] = M;
        }
        // pass 1: inclusive max-scan of the warp's 128 columns, no carry from outside the warp
        int32_t inner = kNegInf;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int32_t v = x[k];
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t y = __shfl_up_sync(0xFFFFFFFFu, v, d);
                if (lane >= d) v = y > v ? y : v;
            }
            v = v > inner ? v : inner;
            x[k] = v;
            inner = __shfl_sync(0xFFFFFFFFu, v, 31);
        }
        if (lane == 0) seg[parity * 32 + warp] = inner; // maximum of the warp's piece
        cta_sync();
        // pass 2: carry of the step + the pieces of the warps before this one; every warp also learns the step's carry-out
        int32_t before = kNegInf, all = kNegInf;
        for (int v = 0; v < nwarps; v++) {
            const int32_t t = seg[parity * 32 + v];
            if (v < warp) before = t > before ? t : before;
            all = t > all ? t : all;
        }
        const int32_t cin = before > carry ? before : carry;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t j = wbase + 32u * (uint32_t)k + (uint32_t)lane + 1;
            const int32_t v = x[k] > cin ? x[k] : cin;
            if (j <= len) row[j] = v + (int32_t)j * pr.g;
        }
        carry = all > carry ? all : carry;
    }
    cta_sync(); // the row is complete before any warp starts the next one
}

// run_block for a CTA: tid = thread index in the CTA, nthreads = CTA size (a multiple of 32, at most 1024)
POA_DEV void run_block_cta(Work &w, const Params &pr, const uint8_t *seq, const uint64_t *copy_off, uint32_t c0, uint32_t c1, int tid,
                          int nthreads, int32_t *seg)
{
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    if (tid == 0) graph_reset(w);
    cta_sync();
    for (uint32_t c = c0; c < c1; c++) {
        const uint8_t *s = seq + copy_off[c];
        const uint32_t len = (uint32_t)(copy_off[c + 1] - copy_off[c]);
        const uint32_t nodes = w.n_nodes;
        if (nodes != 0 && len != 0) {
            if ((uint64_t)(nodes + 1) * ((uint64_t)len + 1) > w.cap.max_cells || nodes + len + 2 > w.cap.max_align) {
                cta_sync(); // every thread has read w.n_nodes / w.cap before the flag is written
                if (tid == 0) w.err = 1;
                cta_sync();
                return;
            }
            { // dp_init: row 0 by all threads, column 0 by thread 0
                const uint64_t W = (uint64_t)len + 1;
                for (uint64_t j = (uint64_t)tid; j < W; j += (uint64_t)nthreads) w.H[j] = (int32_t)j * pr.g;
                cta_sync();
                if (tid == 0) {
                    for (uint32_t i = 1; i <= nodes; i++) {
                        const int it = w.rank_to_node[i - 1];
                        int32_t penalty = w.in_first[it] < 0 ? 0 : kNegInf;
                        for (int e = w.in_first[it]; e >= 0; e = w.edge_next_in[e]) {
                            const int32_t v = w.H[((uint64_t)w.rank[w.edge_tail[e]] + 1) * W];
                            penalty = v > penalty ? v : penalty;
                        }
                        w.H[(uint64_t)i * W] = penalty + pr.g;
                    }
                }
                cta_sync();
            }
            for (uint32_t i = 1; i <= nodes; i++) dp_row_cta(w, pr, s, len, i, lane, warp, nwarps, seg);
            if (tid == 0) w.n_pairs = traceback(w, pr, s, len);
        } else if (tid == 0) {
            w.n_pairs = 0;
        }
        cta_sync();
        if (tid == 0 && !w.err) add_alignment(w, w.al_pairs, w.n_pairs, s, len);
        cta_sync();
        if (w.err) return;
    }
    if (tid == 0) w.n_columns = msa_columns(w);
    cta_sync();
}

POA_DEV void write_row_cta(const Work &w, uint32_t k, uint8_t *out, int tid, int nthreads)
{
    for (uint32_t j = (uint32_t)tid; j < w.n_columns; j += (uint32_t)nthreads) out[j] = '-';
    cta_sync();
    for (uint32_t t = w.path_off[k] + (uint32_t)tid; t < w.path_off[k + 1]; t += (uint32_t)nthreads) {
        const int node = w.path[t];
        out[w.column[node]] = w.decoder[w.code[node]];
    }
    cta_sync();
}
#endif

} // namespace poa
