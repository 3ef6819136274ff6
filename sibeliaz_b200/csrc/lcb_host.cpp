// lcb_host.cpp -- host front end of the B200 sibeliaz-lcb path: input parsing, SoA index build and
// the output stage.  Mirrors the behaviour (not the code) of the reference:
//   junction file   SibeliaZ-LCB/common/junctionapi.h:80-98
//   FASTA           SibeliaZ-LCB/common/streamfastaparser.cpp:28-92, common/dnachar.cpp:13,52-58
//   index           SibeliaZ-LCB/junctionstorage.h:572-650
//   output          SibeliaZ-LCB/blocksfinder.h:533-670, blocksfinder.cpp:109-174
// Everything here is host-only C++ (no CUDA); the device side lives in lcb_device.cu.
#include "sibeliaz_lcb.h"

#include "host_common.h"
#include "lcb_internal.h"

#include <algorithm>
#include <chrono>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <stdexcept>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

std::atomic<unsigned> lcb_host_thread_cap{0};

extern "C" void lcb_set_host_threads(int n) { lcb_host_thread_cap.store(n > 0 ? (unsigned)n : 0u); }

struct lcb_index {
    int k = 0;
    int32_t C = 0;
    int64_t N = 0, V = 0;
    std::vector<int64_t> chr_off, vtx_off, occ_g;
    std::vector<int32_t> pos_id;
    std::vector<uint32_t> pos_bp;
    std::vector<uint8_t> next_ch, prev_rc;
    std::vector<int32_t> packed_rec, packed_occ; // 4 / 2 words per record (lcb_index_pack)
    std::vector<const uint8_t *> seq_ptr;        // lcb_index_get_sequences
    std::vector<uint64_t> seq_len;
    FastaRecords fasta;
    std::string error;
};

extern "C" int lcb_index_load(const char *graph_file, const char *const *fasta_files, int n_fasta, int k, int abundance,
                              lcb_index **out, char *err, size_t errlen)
{
    if (!graph_file || !out || n_fasta < 0 || k <= 0) return LCB_ERR_ARG;
    lcb_index *ix = new lcb_index;
    try {
        ix->k = k;
        const unsigned T = WorkerCount();
        const bool trace = getenv("LCB_LOAD_TRACE") != nullptr;
        auto t_start = std::chrono::steady_clock::now();
        auto lap = [&](const char *what) {
            if (trace) fprintf(stderr, "[load] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
        };
        // ---- FASTA files on their own threads (each splits its bodies further), concurrently with the junction stream
        std::vector<FastaRecords> per_file((size_t)n_fasta);
        std::vector<std::string> fasta_err((size_t)n_fasta);
        std::vector<int> fasta_code((size_t)n_fasta, LCB_OK);
        const unsigned per_fasta_threads = std::max(1u, T / (unsigned)std::max(1, n_fasta));
        std::vector<std::thread> fasta_pool;
        for (int i = 0; i < n_fasta; i++)
            fasta_pool.emplace_back([&, i]() {
                try {
                    ParseFastaParallel(fasta_files[i], per_file[(size_t)i], per_fasta_threads);
                } catch (Failure &e) {
                    fasta_code[(size_t)i] = e.code;
                    fasta_err[(size_t)i] = e.what();
                } catch (std::exception &e) {
                    fasta_code[(size_t)i] = LCB_ERR_IO;
                    fasta_err[(size_t)i] = e.what();
                }
            });
        struct Joiner {
            std::vector<std::thread> &p;
            ~Joiner()
            {
                for (auto &t : p)
                    if (t.joinable()) t.join();
            }
        } joiner{fasta_pool};

        // ---- junction records: {u32 pos; i64 id} packed little-endian, 12 bytes; separators bump the chromosome
        MappedFile g;
        if (!g.open(graph_file)) throw Failure(LCB_ERR_IO, "Can't read the input file");
        const size_t nrec = g.size / 12;
        auto rd = [&g](size_t i, uint32_t &pos, int64_t &id) {
            memcpy(&pos, g.data + i * 12, 4);
            memcpy(&id, g.data + i * 12 + 4, 8);
        };
        auto chunk = [nrec](unsigned t, unsigned TT, size_t &lo, size_t &hi) {
            lo = nrec * t / TT;
            hi = nrec * (t + 1) / TT;
        };
        std::vector<size_t> seps(T + 1, 0), recs(T + 1, 0);
        std::vector<int64_t> tmax(T, -1);
        Parallel(T, [&](unsigned t, unsigned TT) {
            size_t lo, hi, ns = 0, nr = 0;
            int64_t mx = -1;
            chunk(t, TT, lo, hi);
            for (size_t i = lo; i < hi; i++) {
                uint32_t pos;
                int64_t id;
                rd(i, pos, id);
                if (pos == UINT32_MAX || id == INT64_MAX) ++ns;
                else {
                    ++nr;
                    int64_t a = id < 0 ? -id : id;
                    if (a > mx) mx = a;
                }
            }
            seps[t + 1] = ns, recs[t + 1] = nr, tmax[t] = mx;
        });
        lap("junction pass 1");
        int64_t max_abs = -1;
        for (unsigned t = 0; t < T; t++) {
            seps[t + 1] += seps[t];
            recs[t + 1] += recs[t];
            max_abs = std::max(max_abs, tmax[t]);
        }
        const size_t M = recs[T]; // records before the abundance filter
        ix->V = max_abs + 1;
        std::vector<int32_t> id_all(M);
        std::vector<uint32_t> bp_all(M), chr_all(M);
        std::vector<uint32_t> occ_count((size_t)ix->V + 1, 0);
        Parallel(T, [&](unsigned t, unsigned TT) {
            size_t lo, hi;
            chunk(t, TT, lo, hi);
            uint32_t chr = (uint32_t)seps[t];
            size_t w = recs[t];
            for (size_t i = lo; i < hi; i++) {
                uint32_t pos;
                int64_t id;
                rd(i, pos, id);
                if (pos == UINT32_MAX || id == INT64_MAX) {
                    ++chr;
                    continue;
                }
                id_all[w] = (int32_t)id; // int32 truncation as in Position/Vertex (junctionstorage.h:129,148)
                bp_all[w] = pos;
                chr_all[w] = chr;
                ++w;
                __atomic_fetch_add(&occ_count[(size_t)(id < 0 ? -id : id)], 1u, __ATOMIC_RELAXED);
            }
        });
        lap("junction pass 2");
        ix->C = M ? (int32_t)chr_all[M - 1] + 1 : 0;
        // ---- abundance filter (strict <, junctionstorage.h:610) + compaction
        auto rchunk = [M](unsigned t, unsigned TT, size_t &lo, size_t &hi) {
            lo = M * t / TT;
            hi = M * (t + 1) / TT;
        };
        auto vabs = [&id_all](size_t i) { return (size_t)(id_all[i] < 0 ? -(int64_t)id_all[i] : (int64_t)id_all[i]); };
        std::vector<size_t> kept(T + 1, 0);
        Parallel(T, [&](unsigned t, unsigned TT) {
            size_t lo, hi, n = 0;
            rchunk(t, TT, lo, hi);
            for (size_t i = lo; i < hi; i++) n += occ_count[vabs(i)] < (uint32_t)abundance;
            kept[t + 1] = n;
        });
        for (unsigned t = 0; t < T; t++) kept[t + 1] += kept[t];
        const size_t N = kept[T];
        ix->N = (int64_t)N;
        ix->pos_id.resize(N);
        ix->pos_bp.resize(N);
        std::vector<uint32_t> rec_chr(N);
        Parallel(T, [&](unsigned t, unsigned TT) {
            size_t lo, hi;
            rchunk(t, TT, lo, hi);
            size_t w = kept[t];
            for (size_t i = lo; i < hi; i++)
                if (occ_count[vabs(i)] < (uint32_t)abundance) {
                    ix->pos_id[w] = id_all[i];
                    ix->pos_bp[w] = bp_all[i];
                    rec_chr[w] = chr_all[i];
                    ++w;
                }
        });
        lap("filter + compaction");
        ix->chr_off.assign((size_t)ix->C + 1, 0);
        for (size_t gi = 0; gi < N; gi++) ++ix->chr_off[(size_t)rec_chr[gi] + 1];
        for (int32_t c = 0; c < ix->C; c++) ix->chr_off[(size_t)c + 1] += ix->chr_off[(size_t)c];
        // ---- CSR of occurrences: every thread owns a vertex range and scans the records in genome order, so each
        //      occurrence list comes out sorted by (chr, idx) like std::sort leaves it at junctionstorage.h:646-649
        ix->vtx_off.assign((size_t)ix->V + 1, 0);
        for (int64_t v = 0; v < ix->V; v++)
            ix->vtx_off[(size_t)v + 1] = ix->vtx_off[(size_t)v] + (occ_count[(size_t)v] < (uint32_t)abundance ? occ_count[(size_t)v] : 0);
        ix->occ_g.resize(N);
        {
            std::vector<int64_t> cursor(ix->vtx_off.begin(), ix->vtx_off.end());
            Parallel(T, [&](unsigned t, unsigned TT) {
                // balance by occurrences: vertex range [va, vb) holding ~N/T of them
                auto split = [&](unsigned q) -> int64_t {
                    int64_t target = (int64_t)(N * q / TT);
                    return (int64_t)(std::lower_bound(ix->vtx_off.begin(), ix->vtx_off.end(), target) - ix->vtx_off.begin());
                };
                int64_t va = t == 0 ? 0 : split(t), vb = t + 1 == TT ? ix->V + 1 : split(t + 1);
                for (size_t gi = 0; gi < N; gi++) {
                    int64_t a = ix->pos_id[gi] < 0 ? -(int64_t)ix->pos_id[gi] : (int64_t)ix->pos_id[gi];
                    if (a >= va && a < vb) ix->occ_g[(size_t)cursor[(size_t)a]++] = (int64_t)gi;
                }
            });
        }
        lap("CSR");
        // ---- sequences
        for (auto &th : fasta_pool) th.join();
        lap("FASTA joined");
        for (int i = 0; i < n_fasta; i++)
            if (fasta_code[(size_t)i] != LCB_OK) throw Failure(fasta_code[(size_t)i], fasta_err[(size_t)i]);
        for (auto &pf : per_file)
            for (size_t r = 0; r < pf.seq.size(); r++) {
                ix->fasta.name.push_back(std::move(pf.name[r]));
                ix->fasta.seq.push_back(std::move(pf.seq[r]));
            }
        if ((int64_t)ix->fasta.seq.size() < ix->C)
            throw Failure(LCB_ERR_FORMAT, "the graph refers to more sequences than the FASTA files contain");
        // ---- the two characters the traversal needs per junction (junctionstorage.h:641-642)
        ix->next_ch.resize(N);
        ix->prev_rc.resize(N);
        std::vector<int> bad(T, 0);
        Parallel(T, [&](unsigned t, unsigned TT) {
            for (size_t gi = N * t / TT; gi < N * (t + 1) / TT; gi++) {
                const std::string &s = ix->fasta.seq[rec_chr[gi]];
                size_t p = ix->pos_bp[gi];
                if (p + (size_t)k > s.size()) {
                    bad[t] = 1;
                    continue;
                }
                ix->next_ch[gi] = p + (size_t)k < s.size() ? (uint8_t)s[p + (size_t)k] : 0;
                ix->prev_rc[gi] = p > 0 ? Complement((uint8_t)s[p - 1]) : (uint8_t)'N';
            }
        });
        lap("junction chars");
        for (unsigned t = 0; t < T; t++)
            if (bad[t]) throw Failure(LCB_ERR_FORMAT, "junction position beyond the end of its sequence (wrong -k or FASTA?)");
    } catch (Failure &e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.what());
        int code = e.code;
        delete ix;
        return code;
    } catch (std::exception &e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.what());
        delete ix;
        return LCB_ERR_IO;
    }
    *out = ix;
    return LCB_OK;
}

extern "C" int lcb_index_load_fasta(const char *const *fasta_files, int n_fasta, int k, lcb_index **out, char *err, size_t errlen)
{
    if (!out || n_fasta < 0 || (n_fasta && !fasta_files) || k <= 0) return LCB_ERR_ARG;
    lcb_index *ix = new lcb_index;
    ix->k = k;
    std::vector<FastaRecords> per_file((size_t)n_fasta);
    std::vector<std::string> ferr((size_t)n_fasta);
    std::vector<int> fcode((size_t)n_fasta, LCB_OK);
    {
        const unsigned T = WorkerCount(), per = std::max(1u, T / (unsigned)std::max(1, n_fasta));
        std::vector<std::thread> pool;
        for (int i = 0; i < n_fasta; i++)
            pool.emplace_back([&, i]() {
                try {
                    ParseFastaParallel(fasta_files[i], per_file[(size_t)i], per);
                } catch (Failure &e) {
                    fcode[(size_t)i] = e.code;
                    ferr[(size_t)i] = e.what();
                } catch (std::exception &e) {
                    fcode[(size_t)i] = LCB_ERR_IO;
                    ferr[(size_t)i] = e.what();
                }
            });
        for (auto &t : pool) t.join();
    }
    for (int i = 0; i < n_fasta; i++)
        if (fcode[(size_t)i] != LCB_OK) {
            if (err && errlen) snprintf(err, errlen, "%s", ferr[(size_t)i].c_str());
            const int code = fcode[(size_t)i];
            delete ix;
            return code;
        }
    for (auto &pf : per_file)
        for (size_t r = 0; r < pf.seq.size(); r++) {
            ix->fasta.name.push_back(std::move(pf.name[r]));
            ix->fasta.seq.push_back(std::move(pf.seq[r]));
        }
    ix->chr_off.assign(1, 0);
    ix->vtx_off.assign(1, 0);
    *out = ix;
    return LCB_OK;
}

extern "C" int32_t lcb_index_get_sequences(const lcb_index *cix, const uint8_t *const **seq, const uint64_t **len)
{
    if (!cix) return 0;
    lcb_index *ix = const_cast<lcb_index *>(cix); // lazily built accessor arrays
    if (ix->seq_ptr.size() != ix->fasta.seq.size()) {
        ix->seq_ptr.clear();
        ix->seq_len.clear();
        for (const std::string &s : ix->fasta.seq) {
            ix->seq_ptr.push_back((const uint8_t *)s.data());
            ix->seq_len.push_back(s.size());
        }
    }
    if (seq) *seq = ix->seq_ptr.data();
    if (len) *len = ix->seq_len.data();
    return (int32_t)ix->fasta.seq.size();
}

int lcb_index_set_chromosomes(lcb_index *ix, int32_t n_chr, int k, std::string &err)
{
    if (!ix || n_chr < 0) return LCB_ERR_ARG;
    if ((size_t)n_chr > ix->fasta.seq.size()) {
        err = "the graph refers to more sequences than the FASTA files contain";
        return LCB_ERR_FORMAT;
    }
    ix->C = n_chr;
    ix->k = k;
    return LCB_OK;
}

extern "C" int lcb_index_pack(lcb_index *ix)
{
    if (!ix) return LCB_ERR_ARG;
    const size_t N = (size_t)ix->N;
    if (ix->packed_rec.size() == 4 * N && ix->packed_occ.size() == 2 * N) return LCB_OK;
    ix->packed_rec.resize(4 * N);
    ix->packed_occ.resize(2 * N);
    const unsigned T = WorkerCount();
    std::vector<int> too_many(T, 0);
    Parallel(T, [&](unsigned t, unsigned TT) {
        for (size_t g = N * t / TT; g < N * (t + 1) / TT; g++) {
            const int32_t id = ix->pos_id[g];
            const size_t a = (size_t)(id < 0 ? -(int64_t)id : (int64_t)id);
            const int64_t o0 = ix->vtx_off[a], cnt = ix->vtx_off[a + 1] - o0;
            if (cnt > 65535) too_many[t] = 1;
            int32_t *r = &ix->packed_rec[4 * g];
            r[0] = id, r[1] = (int32_t)ix->pos_bp[g], r[2] = (int32_t)o0;
            r[3] = (int32_t)(((uint32_t)cnt << 16) | ((uint32_t)ix->next_ch[g] << 8) | (uint32_t)ix->prev_rc[g]);
            const size_t og = (size_t)ix->occ_g[g]; // occurrence slot g of the CSR (not record g)
            int32_t *o = &ix->packed_occ[2 * g];
            o[0] = (int32_t)((uint32_t)og | (ix->pos_id[og] < 0 ? 0x80000000u : 0u)), o[1] = (int32_t)ix->pos_bp[og];
        }
    });
    for (unsigned t = 0; t < T; t++)
        if (too_many[t]) {
            ix->packed_rec.clear();
            ix->packed_occ.clear();
            return LCB_ERR_ARG; // lcb_create reports the reason (abundance above 65535)
        }
    return LCB_OK;
}

extern "C" int lcb_index_get_view(const lcb_index *ix, lcb_index_view *v)
{
    if (!ix || !v) return LCB_ERR_ARG;
    const bool packed = ix->N > 0 && ix->packed_rec.size() == 4 * (size_t)ix->N && ix->packed_occ.size() == 2 * (size_t)ix->N;
    v->packed_rec = packed ? ix->packed_rec.data() : nullptr;
    v->packed_occ = packed ? ix->packed_occ.data() : nullptr;
    v->n_chr = ix->C;
    v->n_records = ix->N;
    v->n_vertices = ix->V;
    v->chr_off = ix->chr_off.data();
    v->pos_id = ix->pos_id.data();
    v->pos_bp = ix->pos_bp.data();
    v->next_ch = ix->next_ch.data();
    v->prev_rc = ix->prev_rc.data();
    v->vtx_off = ix->vtx_off.data();
    v->occ_g = ix->occ_g.data();
    return LCB_OK;
}

extern "C" int32_t lcb_index_num_chr(const lcb_index *ix) { return ix ? ix->C : 0; }
extern "C" const char *lcb_index_chr_name(const lcb_index *ix, int32_t c)
{
    return (ix && c >= 0 && (size_t)c < ix->fasta.name.size()) ? ix->fasta.name[(size_t)c].c_str() : "";
}
extern "C" int64_t lcb_index_chr_length(const lcb_index *ix, int32_t c)
{
    return (ix && c >= 0 && (size_t)c < ix->fasta.seq.size()) ? (int64_t)ix->fasta.seq[(size_t)c].size() : -1;
}
extern "C" void lcb_index_free(lcb_index *ix) { delete ix; }

// =================================================================================================
// Output stage.  Byte-identical GFF needs the same comparison sequence fed to the same libstdc++
// std::sort (an unstable introsort) as the reference does at blocksfinder.h:103/623, :662 and
// blocksfinder.cpp:146 -- so the three sorts below are kept, on records in the same order.
// =================================================================================================
namespace {

struct OutBlock {
    int32_t id;
    uint32_t chr;
    uint64_t start, end;
    int32_t abs_id() const { return id < 0 ? -id : id; }
};

// per-base coverage flags (the reference's std::vector<bool> covered[chr], blocksfinder.h:607-611) on raw words
struct Bitmap {
    std::vector<uint64_t> w;
    void reset(size_t bits) { w.assign((bits + 63) / 64, 0); }
    bool test(uint64_t i) const { return (w[i >> 6] >> (i & 63)) & 1; }
    void fill(uint64_t lo, uint64_t hi, bool v) // [lo, hi)
    {
        if (lo >= hi) return;
        uint64_t a = lo >> 6, b = (hi - 1) >> 6;
        uint64_t ma = ~0ULL << (lo & 63), mb = ~0ULL >> (63 - ((hi - 1) & 63));
        if (a == b) {
            uint64_t m = ma & mb;
            w[a] = v ? (w[a] | m) : (w[a] & ~m);
            return;
        }
        w[a] = v ? (w[a] | ma) : (w[a] & ~ma);
        for (uint64_t i = a + 1; i < b; i++) w[i] = v ? ~0ULL : 0;
        w[b] = v ? (w[b] | mb) : (w[b] & ~mb);
    }
};

struct TextBuffer {
    std::string s;
    void put(const char *p, size_t n) { s.append(p, n); }
    void put(const std::string &x) { s.append(x); }
    void put(char c) { s.push_back(c); }
    void num(uint64_t v)
    {
        char tmp[24];
        int n = 0;
        do {
            tmp[n++] = (char)('0' + v % 10);
            v /= 10;
        } while (v);
        while (n) s.push_back(tmp[--n]);
    }
};

void WriteFile(const std::string &path, const std::string &data)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) throw Failure(LCB_ERR_IO, "Cannot open file " + path);
    size_t done = fwrite(data.data(), 1, data.size(), f);
    if (fclose(f) != 0 || done != data.size()) throw Failure(LCB_ERR_IO, "Cannot write file " + path);
}

} // namespace

extern "C" int lcb_write_output(const lcb_index *ix, const lcb_block_instance *blocks, uint64_t n, int min_block,
                                const char *out_dir, int gen_seq, int chunks, int64_t *blocks_found, double *coverage,
                                char *err, size_t errlen)
{
    if (!ix || (!blocks && n) || !out_dir) return LCB_ERR_ARG;
    try {
        if (gen_seq && chunks <= 0) throw Failure(LCB_ERR_ARG, "--chunks must be positive when block sequences are written");
        const bool trace = getenv("LCB_LOAD_TRACE") != nullptr;
        auto t_start = std::chrono::steady_clock::now();
        auto lap = [&](const char *what) {
            if (trace) fprintf(stderr, "[output] %-24s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
        };
        const int32_t C = ix->C;
        int32_t max_id = 0;
        for (uint64_t i = 0; i < n; i++) {
            if (blocks[i].chr >= (uint32_t)C) throw Failure(LCB_ERR_ARG, "block instance refers to an unknown sequence");
            max_id = std::max(max_id, blocks[i].id < 0 ? -blocks[i].id : blocks[i].id);
        }
        std::vector<int> copies((size_t)max_id + 1, 0);
        std::vector<OutBlock> inst(n);
        for (uint64_t i = 0; i < n; i++) {
            inst[i] = OutBlock{blocks[i].id, blocks[i].chr, blocks[i].start, blocks[i].end};
            ++copies[(size_t)inst[i].abs_id()];
        }
        // (1) blocksfinder.h:623 -- by (copies desc, id asc), unstable
        auto by_multiplicity = [&copies](const OutBlock &a, const OutBlock &b) {
            int ma = copies[(size_t)a.abs_id()], mb = copies[(size_t)b.abs_id()];
            if (ma != mb) return ma > mb;
            return a.abs_id() < b.abs_id();
        };
        lap("copy + count");
        {
            // std::sort's permutation is a function of the comparison outcomes only, so sorting 16-byte (key, index)
            // pairs whose integer order IS by_multiplicity yields the reference's order without chasing copies[] in
            // every comparison and without moving 24-byte records
            struct Keyed {
                uint64_t key;
                uint32_t idx;
            };
            std::vector<Keyed> keyed(n);
            for (uint64_t i = 0; i < n; i++) {
                const uint32_t a = (uint32_t)inst[i].abs_id();
                keyed[i] = Keyed{((uint64_t)(0xFFFFFFFFu - (uint32_t)copies[a]) << 32) | a, (uint32_t)i};
            }
            std::sort(keyed.begin(), keyed.end(), [](const Keyed &a, const Keyed &b) { return a.key < b.key; });
            std::vector<OutBlock> sorted(n);
            for (uint64_t i = 0; i < n; i++) sorted[i] = inst[keyed[i].idx];
            inst.swap(sorted);
        }
        lap("sort by multiplicity");
        // trimming against per-base coverage bitmaps, blocksfinder.h:607-656
        std::vector<Bitmap> covered((size_t)C);
        for (int32_t c = 0; c < C; c++) covered[(size_t)c].reset(ix->fasta.seq[(size_t)c].size() + 1);
        std::vector<OutBlock> kept, group;
        int64_t next_id = 1;
        for (size_t lo = 0; lo < inst.size();) {
            size_t hi = lo;
            while (hi < inst.size() && !by_multiplicity(inst[lo], inst[hi])) ++hi;
            group.clear();
            for (size_t i = lo; i < hi; i++) {
                if (i + 12 < inst.size()) { // the bitmaps are far larger than the caches: fetch the words of a later record now
                    const OutBlock &f = inst[i + 12];
                    const Bitmap &fc = covered[f.chr];
                    __builtin_prefetch(&fc.w[f.start >> 6], 1);
                    __builtin_prefetch(&fc.w[f.end >> 6], 1);
                }
                Bitmap &cov = covered[inst[i].chr];
                uint64_t s = inst[i].start, e = inst[i].end;
                while (cov.test(s) && s < e) ++s;
                while (cov.test(e) && e > s) --e;
                if (e - s >= (uint64_t)min_block) {
                    group.push_back(OutBlock{(int32_t)(inst[i].id > 0 ? next_id : -next_id), inst[i].chr, s, e});
                    cov.fill(s, e, true);
                }
            }
            if (group.size() > 1) {
                ++next_id;
                kept.insert(kept.end(), group.begin(), group.end());
            } else {
                for (const OutBlock &b : group) covered[b.chr].fill(b.start, b.end, false);
            }
            lo = hi;
        }
        lap("trim against coverage");
        uint64_t total = 0, in_blocks = 0;
        for (int32_t c = 0; c < C; c++) total += ix->fasta.seq[(size_t)c].size();
        for (const OutBlock &b : kept) in_blocks += b.end - b.start;
        if (blocks_found) *blocks_found = next_id - 1;
        if (coverage) *coverage = total ? double(in_blocks) / double(total) : 0.0;
        // (2) blocksfinder.h:662 -- BlockInstance::operator< = (|id|, chr, start), unstable
        std::sort(kept.begin(), kept.end(), [](const OutBlock &a, const OutBlock &b) {
            if (a.abs_id() != b.abs_id()) return a.abs_id() < b.abs_id();
            if (a.chr != b.chr) return a.chr < b.chr;
            return a.start < b.start;
        });
        lap("sort kept");
        if (mkdir(out_dir, 0755) != 0 && errno != EEXIST) throw Failure(LCB_ERR_IO, std::string("Cannot create dir ") + out_dir);
        auto by_id = [](const OutBlock &a, const OutBlock &b) { return a.abs_id() < b.abs_id(); };
        {
            // (3) blocksfinder.cpp:146 -- by |id| only, unstable, on a copy
            std::vector<OutBlock> rows(kept);
            std::sort(rows.begin(), rows.end(), by_id);
            TextBuffer t;
            t.put("##gff-version 3.1.26\n", 21);
            for (int32_t c = 0; c < C; c++) {
                t.put("##sequence-region ", 18);
                t.put(ix->fasta.name[(size_t)c]);
                t.put(" 1 ", 3);
                t.num(ix->fasta.seq[(size_t)c].size());
                t.put('\n');
            }
            // rows are formatted by several threads into private buffers and concatenated in order
            const unsigned T = std::max(1u, std::min(WorkerCount(), (unsigned)(rows.size() / 65536 + 1)));
            std::vector<TextBuffer> part(T);
            Parallel(T, [&](unsigned tt, unsigned TT) {
                TextBuffer &p = part[tt];
                const size_t lo = rows.size() * tt / TT, hi = rows.size() * (tt + 1) / TT;
                p.s.reserve((hi - lo) * 56 + 64);
                for (size_t i = lo; i < hi; i++) {
                    const OutBlock &b = rows[i];
                    p.put(ix->fasta.name[b.chr]);
                    p.put("\tSibeliaZ\tSO:0000856\t", 21);
                    p.num(b.start + 1);
                    p.put('\t');
                    p.num(b.end);
                    p.put("\t.\t", 3);
                    p.put(b.id > 0 ? '+' : '-');
                    p.put("\t.\tID=", 6);
                    p.num((uint64_t)b.abs_id());
                    p.put('\n');
                }
            });
            lap("sort rows + format");
            // the parts go to their offsets of the file from their own threads (no concatenated copy)
            std::vector<size_t> off(T + 1, t.s.size());
            for (unsigned i = 0; i < T; i++) off[i + 1] = off[i] + part[i].s.size();
            const std::string path = std::string(out_dir) + "/blocks_coords.gff";
            const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
            if (fd < 0) throw Failure(LCB_ERR_IO, "Cannot open file " + path);
            auto put_at = [fd](const std::string &data, size_t at) {
                size_t done = 0;
                while (done < data.size()) {
                    ssize_t w = ::pwrite(fd, data.data() + done, data.size() - done, (off_t)(at + done));
                    if (w < 0 && errno == EINTR) continue;
                    if (w <= 0) return false;
                    done += (size_t)w;
                }
                return true;
            };
            std::vector<int> ok(T, 1);
            bool head_ok = put_at(t.s, 0);
            Parallel(T, [&](unsigned tt, unsigned) { ok[tt] = put_at(part[tt].s, off[tt]) ? 1 : 0; });
            bool all_ok = head_ok && ::close(fd) == 0;
            for (unsigned i = 0; i < T; i++) all_ok = all_ok && ok[i];
            if (!all_ok) throw Failure(LCB_ERR_IO, "Cannot write file " + path);
            lap("write gff");
        }
        if (gen_seq) {
            // blocksfinder.h:533-582: one line per block, blocks dealt round-robin over `chunks` files
            std::vector<OutBlock> rows(kept);
            std::sort(rows.begin(), rows.end(), by_id);
            // block g goes to chunk file g % chunks; thread t formats and writes the files with (index % T) == t
            std::vector<std::pair<size_t, size_t>> groups;
            for (size_t lo = 0; lo < rows.size();) {
                size_t hi = lo;
                while (hi < rows.size() && !by_id(rows[lo], rows[hi])) ++hi;
                groups.emplace_back(lo, hi);
                lo = hi;
            }
            const unsigned T = std::max(1u, std::min<unsigned>(WorkerCount(), (unsigned)chunks));
            std::vector<std::string> failed(T);
            Parallel(T, [&](unsigned tt, unsigned TT) {
                try {
                    std::vector<TextBuffer> chunk;
                    std::vector<size_t> mine; // chunk indices of this thread
                    for (size_t c = tt; c < (size_t)chunks; c += TT) mine.push_back(c);
                    chunk.resize(mine.size());
                    for (size_t g = 0; g < groups.size(); g++) {
                        const size_t which = g % (size_t)chunks;
                        if (which % TT != tt) continue;
                        TextBuffer &t = chunk[which / TT];
                        for (size_t i = groups[g].first; i < groups[g].second; i++) {
                            const OutBlock &b = rows[i];
                            const std::string &s = ix->fasta.seq[b.chr];
                            uint64_t len = b.end - b.start;
                            t.put("> ", 2);
                            t.put(ix->fasta.name[b.chr]);
                            t.put(';');
                            if (b.id > 0) {
                                t.num(b.start);
                                t.put(';');
                                t.num(len);
                                t.put(";+;", 3);
                                t.num(s.size());
                                t.put('@');
                                t.put(s.data() + b.start, (size_t)len);
                            } else {
                                t.num(s.size() - b.end);
                                t.put(';');
                                t.num(len);
                                t.put(";-;", 3);
                                t.num(s.size());
                                t.put('@');
                                const size_t at = t.s.size();
                                t.s.resize(at + (size_t)len);
                                char *w = &t.s[at];
                                for (uint64_t j = 0; j < len; j++) w[j] = (char)Complement((uint8_t)s[b.end - 1 - j]);
                            }
                            t.put('@');
                        }
                        t.put('\n');
                    }
                    for (size_t k = 0; k < mine.size(); k++) WriteFile(std::string(out_dir) + "/" + std::to_string(mine[k]) + ".tmp", chunk[k].s);
                } catch (std::exception &e) {
                    failed[tt] = e.what();
                }
            });
            for (const std::string &f : failed)
                if (!f.empty()) throw Failure(LCB_ERR_IO, f);
        }
    } catch (Failure &e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.what());
        return e.code;
    } catch (std::exception &e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.what());
        return LCB_ERR_IO;
    }
    return LCB_OK;
}

extern "C" const char *lcb_version(void) { return "sibeliaz-lcb-b200 0.1 (output-compatible with SibeliaZ-LCB 1.2.7)"; }
