// graph_host.cpp -- host front end of the junction finder (include/sibeliaz_graph.h): FASTA parsing on the host threads,
// the device pipeline (graph_device.cu, which returns the finished junction records: what the reference's
// EdgeConstructionWorker writes, TwoPaCo/src/graphconstructor/vertexenumerator.h:856-993), and the junction file with its
// chromosome separators (TwoPaCo/src/common/junctionapi.h:117-131).  Host-only C++ (no CUDA here).
#include "sibeliaz_graph.h"

#include <chrono>

#include "graph_internal.h"
#include "host_common.h"

struct lcg_graph {
    uint64_t n = 0;
    std::unique_ptr<uint32_t[]> chr, pos;
    std::unique_ptr<int32_t[]> id; // 31 bits suffice (checked on the device side); widened at the ABI / in the file
    lcg_stats st{};
    lcg::Resident resident;
    ~lcg_graph() { lcg::free_resident(resident); }
};

namespace {
// a resident graph keeps its records on the device; the host copy appears when somebody asks for it
int EnsureHostCopy(const lcg_graph *cg, std::string &err)
{
    lcg_graph *g = const_cast<lcg_graph *>(cg);
    if (g->chr || !g->resident.d_text) return LCG_OK;
    const size_t n = std::max<size_t>((size_t)g->n, 1);
    g->chr.reset(new uint32_t[n]);
    g->pos.reset(new uint32_t[n]);
    g->id.reset(new int32_t[n]);
    return lcg::download(g->resident, g->chr.get(), g->pos.get(), g->id.get(), err);
}
} // namespace

const lcg::Resident *lcg::resident_of(const lcg_graph *g) { return g && g->resident.d_text ? &g->resident : nullptr; }

namespace {

void SetErr(char *err, size_t errlen, const std::string &m)
{
    if (err && errlen) snprintf(err, errlen, "%s", m.c_str());
}

int Build(const uint8_t *const *seq, const uint64_t *len, int n, int k, uint64_t abundance, int device, double ms_parse, bool keep,
          lcg_graph **out, char *err, size_t errlen)
{
    if (k < 1 || k % 2 == 0) {
        SetErr(err, errlen, "value of K must be odd");
        return LCG_ERR_ARG;
    }
    if (k > 255) { // (the reference's CAPACITY template stops at its build's MAX_CAPACITY, vertexenumerator.cpp:20-58)
        SetErr(err, errlen, "k > 255 is not supported by the GPU junction finder (a k-mer is at most eight 64-bit words)");
        return LCG_ERR_ARG;
    }
    for (int r = 0; r < n; r++)
        if (len[r] >= 0xFFFFFFFFull) { // positions are 32-bit in the wire format (junctionapi.h:32-33)
            SetErr(err, errlen, "a sequence is too long for 32-bit junction positions");
            return LCG_ERR_ARG;
        }
    const auto t0 = std::chrono::steady_clock::now();
    lcg::DeviceInput in{seq, len, n, k, abundance, device, keep};
    lcg::DeviceOutput dev;
    std::string e;
    const int rc = lcg::run_device(in, dev, e);
    if (rc) {
        SetErr(err, errlen, e);
        return rc;
    }
    lcg_graph *g = new lcg_graph;
    g->st = dev.st;
    g->st.ms_parse = ms_parse;
    g->n = dev.n;
    g->chr = std::move(dev.chr);
    g->pos = std::move(dev.pos);
    g->id = std::move(dev.id);
    g->st.n_junctions = g->n;
    g->resident = dev.resident;
    dev.resident = lcg::Resident();
    g->st.ms_total = ms_parse + std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = g;
    return LCG_OK;
}

} // namespace

extern "C" int lcg_build(const uint8_t *const *seq, const uint64_t *len, int n_records, int k, uint64_t abundance, int device,
                         lcg_graph **out, char *err, size_t errlen)
{
    if (!out || n_records < 0 || (n_records && (!seq || !len))) return LCG_ERR_ARG;
    return Build(seq, len, n_records, k, abundance, device, 0.0, false, out, err, errlen);
}

extern "C" int lcg_build_resident(const uint8_t *const *seq, const uint64_t *len, int n_records, int k, uint64_t abundance, int device,
                                  lcg_graph **out, char *err, size_t errlen)
{
    if (!out || n_records < 0 || (n_records && (!seq || !len))) return LCG_ERR_ARG;
    return Build(seq, len, n_records, k, abundance, device, 0.0, true, out, err, errlen);
}

extern "C" int lcg_build_from_fasta(const char *const *fasta_files, int n_files, int k, uint64_t abundance, int device,
                                    lcg_graph **out, char *err, size_t errlen)
{
    if (!out || n_files < 0 || (n_files && !fasta_files)) return LCG_ERR_ARG;
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<FastaRecords> per_file((size_t)n_files);
    std::vector<std::string> ferr((size_t)n_files);
    std::vector<int> fcode((size_t)n_files, LCG_OK);
    {
        const unsigned T = WorkerCount(), per = std::max(1u, T / (unsigned)std::max(1, n_files));
        std::vector<std::thread> pool;
        for (int i = 0; i < n_files; i++)
            pool.emplace_back([&, i]() {
                try {
                    ParseFastaParallel(fasta_files[i], per_file[(size_t)i], per);
                } catch (Failure &e) {
                    fcode[(size_t)i] = e.code == LCB_ERR_FORMAT ? LCG_ERR_FORMAT : LCG_ERR_IO;
                    ferr[(size_t)i] = e.what();
                } catch (std::exception &e) {
                    fcode[(size_t)i] = LCG_ERR_IO;
                    ferr[(size_t)i] = e.what();
                }
            });
        for (auto &t : pool) t.join();
    }
    for (int i = 0; i < n_files; i++)
        if (fcode[(size_t)i]) {
            SetErr(err, errlen, ferr[(size_t)i]);
            return fcode[(size_t)i];
        }
    std::vector<const uint8_t *> seq;
    std::vector<uint64_t> len;
    for (auto &pf : per_file)
        for (auto &s : pf.seq) {
            seq.push_back((const uint8_t *)s.data());
            len.push_back(s.size());
        }
    const double ms_parse = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return Build(seq.data(), len.data(), (int)seq.size(), k, abundance, device, ms_parse, false, out, err, errlen);
}

extern "C" uint64_t lcg_num_junctions(const lcg_graph *g) { return g ? g->n : 0; }

extern "C" int lcg_get_junctions(const lcg_graph *g, uint32_t *chr, uint32_t *pos, int64_t *id)
{
    if (!g) return LCG_ERR_ARG;
    {
        std::string e;
        const int rc = EnsureHostCopy(g, e);
        if (rc) return rc;
    }
    const size_t n = (size_t)g->n;
    if (chr) memcpy(chr, g->chr.get(), n * sizeof(uint32_t));
    if (pos) memcpy(pos, g->pos.get(), n * sizeof(uint32_t));
    if (id)
        for (size_t i = 0; i < n; i++) id[i] = g->id[i];
    return LCG_OK;
}

extern "C" int lcg_write_junction_file(const lcg_graph *g, const char *path, char *err, size_t errlen)
{
    if (!g || !path) return LCG_ERR_ARG;
    {
        std::string e;
        const int rc = EnsureHostCopy(g, e);
        if (rc) {
            if (err && errlen) snprintf(err, errlen, "%s", e.c_str());
            return rc;
        }
    }
    // 12-byte records {u32 pos; i64 id}; {0xFFFFFFFF, INT64_MAX} once per chromosome change (junctionapi.h:117-131)
    std::string buf;
    buf.reserve((size_t)g->n * 12 + 1024);
    uint32_t now = 0;
    auto rec = [&buf](uint32_t pos, int64_t id) {
        buf.append((const char *)&pos, 4);
        buf.append((const char *)&id, 8);
    };
    for (size_t i = 0; i < (size_t)g->n; i++) {
        for (; g->chr[i] > now; ++now) rec(0xFFFFFFFFu, INT64_MAX);
        rec(g->pos[i], (int64_t)g->id[i]);
    }
    FILE *f = fopen(path, "wb");
    if (!f) {
        SetErr(err, errlen, "Can't create the output file");
        return LCG_ERR_IO;
    }
    const size_t done = fwrite(buf.data(), 1, buf.size(), f);
    if (fclose(f) != 0 || done != buf.size()) {
        SetErr(err, errlen, "Can't write to the output file");
        return LCG_ERR_IO;
    }
    return LCG_OK;
}

extern "C" int lcg_get_stats(const lcg_graph *g, lcg_stats *st)
{
    if (!g || !st) return LCG_ERR_ARG;
    *st = g->st;
    return LCG_OK;
}

extern "C" void lcg_free(lcg_graph *g) { delete g; }
