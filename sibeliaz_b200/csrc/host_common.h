// host_common.h -- host-only helpers shared by the LCB front end (lcb_host.cpp) and the junction finder's front end
// (graph_host.cpp): error type, file mapping, FASTA reader (rules of SibeliaZ-LCB/common/streamfastaparser.cpp:28-92 and
// TwoPaCo/src/common/streamfastaparser.cpp, which are the same parser), thread helpers.  Everything sits in an anonymous
// namespace: each translation unit gets its own copy.
#pragma once
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdint>
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

#ifndef LCB_ERR_IO
#define LCB_ERR_IO 2
#define LCB_ERR_FORMAT 3
#endif

extern std::atomic<unsigned> lcb_host_thread_cap; // one per library (defined in lcb_host.cpp); 0 = no cap

namespace {

struct Failure : std::runtime_error {
    int code;
    Failure(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

// Read-only mapping of a whole file (falls back to read() for empty / special files).
struct MappedFile {
    const uint8_t *data = nullptr;
    size_t size = 0;
    std::vector<uint8_t> owned;
    bool mapped = false;
    bool open(const char *path)
    {
        int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
            void *p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p != MAP_FAILED) {
                madvise(p, (size_t)st.st_size, MADV_SEQUENTIAL);
                data = (const uint8_t *)p;
                size = (size_t)st.st_size;
                mapped = true;
                ::close(fd);
                return true;
            }
        }
        uint8_t buf[1 << 16];
        ssize_t n;
        while ((n = ::read(fd, buf, sizeof buf)) > 0) owned.insert(owned.end(), buf, buf + n);
        ::close(fd);
        data = owned.data();
        size = owned.size();
        return true;
    }
    ~MappedFile()
    {
        if (mapped) munmap((void *)data, size);
    }
};

// Character classes for the FASTA scanner: 0 = invalid, 1 = whitespace, 2 = '>', else upper-cased base.
struct FastaTable {
    uint8_t t[256];
    FastaTable()
    {
        memset(t, 0, sizeof t);
        const char *valid = "ACGTURYKMSWBDHWNXV"; // dnachar.cpp:13
        for (const char *p = valid; *p; ++p) {
            t[(uint8_t)*p] = (uint8_t)*p;
            t[(uint8_t)tolower(*p)] = (uint8_t)*p;
        }
        for (int c = 0; c < 256; c++)
            if (isspace(c)) t[c] = 1;
        t[(uint8_t)'>'] = 2;
    }
};
const FastaTable kFasta;

inline uint8_t Complement(uint8_t c) // dnachar.cpp:52-58
{
    switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    }
    return 'N';
}

struct FastaRecords {
    std::vector<std::string> name;
    std::vector<std::string> seq;
};

// Host threads the library may use for parsing, packing and output formatting: min(cores, 32) (the wrapper's own cap,
// SibeliaZ-LCB/sibeliaz:139) unless the caller set a lower bound with lcb_set_host_threads (the CLI passes -t).
inline unsigned WorkerCount()
{
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    unsigned n = std::min(hw, 32u);
    const unsigned cap = lcb_host_thread_cap.load(std::memory_order_relaxed);
    if (cap) n = std::min(n, cap);
    return n;
}

// runs fn(t, T) on T threads and joins
template <class F>
void Parallel(unsigned T, F fn)
{
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < T; t++) pool.emplace_back([&fn, t, T]() { fn(t, T); });
    fn(0, T);
    for (auto &th : pool) th.join();
}

// FASTA with the sequence bodies filtered in parallel: count kept characters per 4 MiB slice, then write
void ParseFastaParallel(const char *path, FastaRecords &out, unsigned T)
{
    MappedFile f;
    if (!f.open(path)) throw Failure(LCB_ERR_IO, std::string("Can't open file ") + path);
    const uint8_t *p = f.data, *end = f.data + f.size;
    std::string header;
    while (p < end) {
        if (*p != '>')
            throw Failure(LCB_ERR_FORMAT, std::string("The FASTA header should start with a '>', started with '") + (char)*p + "'");
        ++p;
        const uint8_t *nl = (const uint8_t *)memchr(p, '\n', (size_t)(end - p));
        const uint8_t *line_end = nl ? nl : end;
        if (nl) {
            const uint8_t *a = p;
            while (a < line_end && isspace(*a)) ++a;
            const uint8_t *b = a;
            while (b < line_end && !isspace(*b)) ++b;
            if (b > a) header.assign((const char *)a, (size_t)(b - a));
        }
        p = nl ? nl + 1 : end;
        const uint8_t *q = (const uint8_t *)memchr(p, '>', (size_t)(end - p));
        const uint8_t *rec_end = q ? q : end;
        const size_t body = (size_t)(rec_end - p);
        const size_t slice = 4u << 20;
        const size_t ns = (body + slice - 1) / slice;
        std::vector<size_t> kept(ns + 1, 0);
        std::vector<const uint8_t *> bad(ns, nullptr);
        const unsigned Tn = (unsigned)std::max<size_t>(1, std::min<size_t>(T, ns));
        Parallel(Tn, [&](unsigned t, unsigned TT) {
            for (size_t s = t; s < ns; s += TT) {
                const uint8_t *a = p + s * slice, *b = std::min(rec_end, a + slice);
                size_t n = 0;
                for (const uint8_t *r = a; r < b; ++r) {
                    uint8_t c = kFasta.t[*r];
                    n += c > 2;
                    if (c == 0 && !bad[s]) bad[s] = r;
                }
                kept[s + 1] = n;
            }
        });
        for (size_t s = 0; s < ns; s++)
            if (bad[s])
                throw Failure(LCB_ERR_FORMAT, std::string("Found an invalid character '") + (char)*bad[s] + "' in sequence " + header);
        for (size_t s = 0; s < ns; s++) kept[s + 1] += kept[s];
        out.name.push_back(header);
        out.seq.emplace_back();
        std::string &str = out.seq.back();
        str.resize(kept[ns]);
        char *base = str.empty() ? nullptr : &str[0];
        Parallel(Tn, [&](unsigned t, unsigned TT) {
            for (size_t s = t; s < ns; s += TT) {
                const uint8_t *a = p + s * slice, *b = std::min(rec_end, a + slice);
                char *w = base + kept[s];
                for (const uint8_t *r = a; r < b; ++r) {
                    uint8_t c = kFasta.t[*r];
                    if (c > 2) *w++ = (char)c;
                }
            }
        });
        p = rec_end;
    }
}

} // namespace
