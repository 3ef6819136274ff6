// graph_kmer.cuh -- per-position device code of the junction finder (graph_device.cu): the k-mer table and what one
// thread does at one text position in each pass.  Two k-mer widths share it:
//
//   W = 1   k <= 31: a k-mer is one 64-bit word and IS the table key (the original path);
//   W >= 2  31 < k <= 255 (the reference's CAPACITY template, TwoPaCo/src/graphconstructor/vertexenumerator.cpp:20-58,
//           takes any k its build allows): the canonical k-mer is W words.  The 16-byte slot keeps its shape: the key
//           word holds a 24-bit fingerprint and the text position of the occurrence that claimed the slot (the
//           "representative"), so that claiming stays ONE 64-bit atomicCAS; whoever meets a slot with its fingerprint
//           compares its k-mer with the representative's, read from the packed text.
//
// Everything here is per thread (no warp collectives), so tests/graph_emu.cpp compiles this file for the host as written
// and runs the passes position by position; the test suite compares its junction file with the CPU restatement's.
// Include INSIDE an anonymous namespace, after <cstdint>.
#pragma once

constexpr uint64_t kEmpty = ~0ULL;
constexpr int kIdShift = 34;               // info: 0-3 in chars, 4-7 out chars, 8-32 pairs, 33 "a pair seen twice", 34.. id
constexpr uint64_t kMultiBit = 1ULL << 33;
constexpr int kMaxWideWords = 8;           // 2k <= 512 bits
constexpr int kMaxK = 255;
constexpr int kRepBits = 40;               // wide slots: bits 0-39 text position of the representative, 40-63 fingerprint
constexpr uint64_t kRepMask = (1ULL << kRepBits) - 1;

struct Text {
    const uint64_t *bits; // 32 bases per word, base i of a word at bits 2i
    const uint32_t *nm;   // 32 bases per word, bit i set: not definite
    uint64_t n;           // positions in G
    int k;
};

struct Slot { // key and vertex word side by side: one 32-byte sector per probe
    unsigned long long key, info;
};
struct Table {
    Slot *slot;
    uint64_t mask;
};

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

__device__ __forceinline__ unsigned comp4(unsigned m) // neighbour-character set under complement: bit c -> bit 3 - c
{
    return ((m & 1u) << 3) | ((m & 2u) << 1) | ((m & 4u) >> 1) | ((m & 8u) >> 3);
}

// 2-bit groups of a word in reverse order
__device__ __forceinline__ uint64_t revpairs64(uint64_t x)
{
    const uint64_t r = __brevll(x);
    return ((r & 0x5555555555555555ULL) << 1) | ((r >> 1) & 0x5555555555555555ULL);
}

template <int W>
struct Kmer { // canonical k-mer: the smaller of the k-mer and its reverse complement as numbers, first base most significant
    uint64_t w[W]; // little-endian words (w[0] least significant)
    uint64_t hash;
    bool fwd;      // the k-mer itself is the canonical one
    int prev, next; // 0-3, 4 = not definite
};
template <>
struct Kmer<1> {
    uint64_t key;
    bool fwd;
    int prev, next;
};

// ---- k <= 31 -----------------------------------------------------------------------------------------------------------
// k-mer at position p (1 <= p, p + k < n).  Returns false when it holds a non-definite character.
__device__ __forceinline__ bool load_kmer(const Text &t, uint64_t p, Kmer<1> &out)
{
    const int k = t.k;
    const uint64_t q = p - 1; // window [p - 1, p + k]: prev, k-mer, next  (k + 2 <= 33 bits of the N mask)
    const uint64_t w = q >> 5;
    const unsigned off = (unsigned)(q & 31);
    const uint64_t nmw = ((uint64_t)t.nm[w] | ((uint64_t)t.nm[w + 1] << 32)) >> off;
    if ((nmw >> 1) & ((1ULL << k) - 1)) return false;
    const bool prev_n = nmw & 1, next_n = (nmw >> (k + 1)) & 1;
    // 2-bit codes of [p - 1, p + k]: up to 66 bits -> the k-mer from two words, prev and next on their own
    const uint64_t pw = p >> 5;
    const unsigned po = 2u * (unsigned)(p & 31);
    uint64_t x = t.bits[pw] >> po;
    if (po) x |= t.bits[pw + 1] << (64 - po);
    const uint64_t kmask = (1ULL << (2 * k)) - 1;
    x &= kmask; // base i of the k-mer at bits 2i
    const uint64_t r = __brevll(x);
    const uint64_t fw = (((r & 0x5555555555555555ULL) << 1) | ((r >> 1) & 0x5555555555555555ULL)) >> (64 - 2 * k);
    const uint64_t rc = ~x & kmask; // reverse complement with ITS first base most significant
    out.fwd = fw < rc;
    out.key = out.fwd ? fw : rc;
    out.prev = prev_n ? 4 : (int)((t.bits[q >> 5] >> (2 * (q & 31))) & 3);
    const uint64_t e = p + (uint64_t)k;
    out.next = next_n ? 4 : (int)((t.bits[e >> 5] >> (2 * (e & 31))) & 3);
    return true;
}

__device__ __forceinline__ uint64_t find_or_insert(const Table &tb, uint64_t key)
{
    uint64_t s = mix64(key) & tb.mask;
    while (true) {
        const uint64_t cur = tb.slot[s].key;
        if (cur == key) return s;
        if (cur == kEmpty) {
            const unsigned long long old = atomicCAS(&tb.slot[s].key, (unsigned long long)kEmpty, (unsigned long long)key);
            if (old == kEmpty || old == key) return s;
        }
        s = (s + 1) & tb.mask;
    }
}

__device__ __forceinline__ uint64_t find(const Table &tb, uint64_t key) // the key is present
{
    uint64_t s = mix64(key) & tb.mask;
    while (tb.slot[s].key != key) s = (s + 1) & tb.mask;
    return s;
}

__device__ __forceinline__ uint64_t find_or_insert(const Table &tb, const Text &, const Kmer<1> &km, uint64_t) { return find_or_insert(tb, km.key); }
__device__ __forceinline__ uint64_t find(const Table &tb, const Text &, const Kmer<1> &km) { return find(tb, km.key); }

// ---- 31 < k <= 255 -----------------------------------------------------------------------------------------------------
// The window's 2-bit codes as a number with base i at bits 2i (little-endian words); bits from 2k on are cleared.
template <int W>
__device__ __forceinline__ void load_codes(const Text &t, uint64_t p, uint64_t (&x)[W])
{
    const uint64_t pw = p >> 5;
    const unsigned po = 2u * (unsigned)(p & 31);
    const unsigned top = (unsigned)(2 * t.k - 64 * (W - 1)); // bits of the top word: 2 .. 64
#pragma unroll
    for (int j = 0; j < W; j++) {
        uint64_t v = t.bits[pw + j] >> po;
        if (po) v |= t.bits[pw + j + 1] << (64 - po); // (p + k - 1) >> 5 >= pw + W - 1: inside the padded array
        x[j] = v;
    }
    if (top < 64) x[W - 1] &= (1ULL << top) - 1;
}

// x (base i at bits 2i) -> canonical value.  Its reverse complement with ITS first base most significant is ~x; the
// k-mer itself with its first base most significant is x with the k two-bit groups in reverse order.
template <int W>
__device__ __forceinline__ void canonical(const uint64_t (&x)[W], int k, Kmer<W> &out)
{
    const unsigned top = (unsigned)(2 * k - 64 * (W - 1));
    const unsigned s = 64u - top; // 0 .. 62: the reversal over 64 W bits sits s bits too high
    uint64_t r[W + 1];
#pragma unroll
    for (int j = 0; j < W; j++) r[j] = revpairs64(x[W - 1 - j]);
    r[W] = 0;
    uint64_t fw[W], rc[W];
#pragma unroll
    for (int j = 0; j < W; j++) {
        fw[j] = s ? (r[j] >> s) | (r[j + 1] << (64 - s)) : r[j];
        rc[j] = ~x[j];
    }
    if (top < 64) rc[W - 1] &= (1ULL << top) - 1;
    bool fwd = false, decided = false; // fw < rc, most significant word first (k odd: never equal)
#pragma unroll
    for (int j = W - 1; j >= 0; j--)
        if (!decided && fw[j] != rc[j]) {
            fwd = fw[j] < rc[j];
            decided = true;
        }
    out.fwd = fwd;
    uint64_t h = 0x9E3779B97F4A7C15ULL;
#pragma unroll
    for (int j = 0; j < W; j++) {
        out.w[j] = fwd ? fw[j] : rc[j];
        h = mix64(h ^ out.w[j]);
    }
    out.hash = h;
}

template <int W>
__device__ __forceinline__ bool load_kmer(const Text &t, uint64_t p, Kmer<W> &out)
{
    const int k = t.k;
    const uint64_t last = p + (uint64_t)k - 1; // a non-definite character in [p, last]?
    for (uint64_t w = p >> 5; w <= (last >> 5); w++) {
        uint32_t m = t.nm[w];
        if (w == (p >> 5)) m &= ~0u << (unsigned)(p & 31);
        if (w == (last >> 5)) m &= ~0u >> (31u - (unsigned)(last & 31));
        if (m) return false;
    }
    const uint64_t q = p - 1, e = last + 1;
    const bool prev_n = (t.nm[q >> 5] >> (q & 31)) & 1u, next_n = (t.nm[e >> 5] >> (e & 31)) & 1u;
    uint64_t x[W];
    load_codes<W>(t, p, x);
    canonical<W>(x, k, out);
    out.prev = prev_n ? 4 : (int)((t.bits[q >> 5] >> (2 * (q & 31))) & 3);
    out.next = next_n ? 4 : (int)((t.bits[e >> 5] >> (2 * (e & 31))) & 3);
    return true;
}

template <int W>
__device__ __forceinline__ unsigned long long wide_key(const Kmer<W> &km, uint64_t p) // fingerprint | position; never kEmpty (p < 2^40 - 1)
{
    return (unsigned long long)((km.hash >> kRepBits) << kRepBits) | (unsigned long long)p;
}

// does the slot key `cur` (not empty) stand for km's k-mer?
template <int W>
__device__ __forceinline__ bool same_kmer(const Text &t, unsigned long long cur, const Kmer<W> &km)
{
    if ((cur >> kRepBits) != (km.hash >> kRepBits)) return false;
    uint64_t x[W];
    load_codes<W>(t, cur & kRepMask, x); // the representative is a definite k-mer: whoever stored it had loaded it
    Kmer<W> rep;
    canonical<W>(x, t.k, rep);
    bool same = true;
#pragma unroll
    for (int j = 0; j < W; j++) same &= rep.w[j] == km.w[j];
    return same;
}

template <int W>
__device__ __forceinline__ uint64_t find_or_insert(const Table &tb, const Text &t, const Kmer<W> &km, uint64_t p)
{
    uint64_t s = km.hash & tb.mask;
    const unsigned long long mine = wide_key<W>(km, p);
    while (true) {
        unsigned long long cur = *(volatile unsigned long long *)&tb.slot[s].key;
        if (cur == kEmpty) {
            cur = atomicCAS(&tb.slot[s].key, (unsigned long long)kEmpty, mine);
            if (cur == kEmpty) return s; // claimed: this occurrence is the representative
        }
        if (same_kmer<W>(t, cur, km)) return s;
        s = (s + 1) & tb.mask;
    }
}

template <int W>
__device__ __forceinline__ uint64_t find(const Table &tb, const Text &t, const Kmer<W> &km) // the k-mer is present
{
    uint64_t s = km.hash & tb.mask;
    while (!same_kmer<W>(t, tb.slot[s].key, km)) s = (s + 1) & tb.mask;
    return s;
}

// ---- what one thread does at position p in each pass (any W) ------------------------------------------------------------
// table-building pass: find-or-insert the k-mer, OR the neighbour characters this occurrence shows into the vertex word
// (next to a non-definite character the two dummies A and T, vertexenumerator.h:1046-1058).  Returns "a definite k-mer".
template <int W>
__device__ __forceinline__ bool edges_at(const Text &t, const Table &tb, uint64_t p)
{
    if (p + (uint64_t)t.k >= t.n) return false;
    Kmer<W> km;
    if (!load_kmer(t, p, km)) return false;
    const unsigned in = km.prev < 4 ? 1u << km.prev : 9u, out = km.next < 4 ? 1u << km.next : 9u;
    const unsigned long long add = km.fwd ? (in | (out << 4)) : (comp4(out) | (comp4(in) << 4));
    const uint64_t s = find_or_insert(tb, t, km, p);
    if ((tb.slot[s].info & add) != add) atomicOr(&tb.slot[s].info, add);
    return true;
}

// candidate pass (vertexenumerator.h:630-660): positions with more than one in- or out-edge are flagged and OR their
// canonical (prev, next) pair into the vertex (what CandidateFinalFilteringWorker's hash sets collect)
template <int W>
__device__ __forceinline__ void candidate_at(const Text &t, const Table &tb, uint8_t *flag, unsigned *count, uint64_t p)
{
    if (p + (uint64_t)t.k >= t.n) return;
    Kmer<W> km;
    if (!load_kmer(t, p, km)) return;
    const uint64_t s = find(tb, t, km);
    const unsigned long long info = tb.slot[s].info;
    const unsigned vin = (unsigned)info & 15u, vout = ((unsigned)info >> 4) & 15u;
    const int in = km.prev < 4 ? __popc(km.fwd ? vin : vout) : 2;
    const int out = km.next < 4 ? __popc(km.fwd ? vout : vin) : 2;
    if (in <= 1 && out <= 1) return;
    flag[p] = 1;
    const int cp = km.fwd ? km.prev : (km.next < 4 ? 3 - km.next : 4), cn = km.fwd ? km.next : (km.prev < 4 ? 3 - km.prev : 4);
    const unsigned long long bit = 1ULL << (8 + cp * 5 + cn);
    const unsigned long long old = atomicOr(&tb.slot[s].info, bit);
    if ((old & bit) && !(old & kMultiBit)) atomicOr(&tb.slot[s].info, kMultiBit);
    if (count) atomicAdd(&count[s], 1u); // only with a finite abundance threshold (twopaco -a)
}

__device__ __forceinline__ bool is_bifurcation(unsigned long long info)
{
    const unsigned pairs = (unsigned)(info >> 8) & 0x1FFFFFFu;
    if (!pairs) return false;
    if (pairs & (pairs - 1)) return true; // two different (prev, next) pairs
    if (!(info & kMultiBit)) return false; // a single candidate occurrence
    const int p = __ffs((int)pairs) - 1;
    return p / 5 == 4 || p % 5 == 4; // the shared prev (or next) is 'N': unknown twice
}

constexpr int32_t kStub = INT32_MIN; // a first / last k-mer of a record that is not a junction: gets a unique id

// signed id of the flagged position p (flag bit 0: candidate, bit 1: first / last k-mer of a record)
template <int W>
__device__ __forceinline__ int32_t id_at(const Text &t, const Table &tb, unsigned f, uint64_t p)
{
    int32_t out = 0;
    Kmer<W> km;
    if ((f & 1u) && load_kmer(t, p, km)) {
        const long long v = (long long)(tb.slot[find(tb, t, km)].info >> kIdShift);
        out = (int32_t)(km.fwd ? v : -v);
    }
    if (out == 0 && (f & 2u)) out = kStub;
    return out;
}

// wide only: the words of the i-th bifurcation k-mer (for the sort that ranks them), word-major
template <int W>
__device__ __forceinline__ void canon_words_at(const Text &t, const uint64_t *bif_keys, unsigned n, unsigned i, uint64_t *words)
{
    uint64_t x[W];
    load_codes<W>(t, bif_keys[i] & kRepMask, x);
    Kmer<W> km;
    canonical<W>(x, t.k, km);
#pragma unroll
    for (int j = 0; j < W; j++) words[(size_t)j * n + i] = km.w[j];
}

// wide only: the i-th k-mer in sorted order gets id i + 1; its slot is the one whose key word is bif_keys[..] itself
template <int W>
__device__ __forceinline__ void assign_id_at(const Text &t, const Table &tb, unsigned long long key, unsigned rank)
{
    uint64_t x[W];
    load_codes<W>(t, key & kRepMask, x);
    Kmer<W> km;
    canonical<W>(x, t.k, km);
    uint64_t s = km.hash & tb.mask;
    while (tb.slot[s].key != key) s = (s + 1) & tb.mask;
    atomicOr(&tb.slot[s].info, (unsigned long long)(rank + 1) << kIdShift);
}
