// sibeliaz-lcb (B200): drop-in for the reference binary invoked at SibeliaZ-LCB/sibeliaz:146.
// Same flags, defaults, stdout lines and exit codes as SibeliaZ-LCB/sibeliaz.cpp:37-157; the work is done
// by libsibeliaz_lcb through its C ABI (include/sibeliaz_lcb.h).  -t bounds the host threads (parsing, packing, output
// formatting).  Additive flags: --gpu, --window, --stats, --construct (fused pipeline: the junctions are found on the GPU
// from the FASTA files, no --graph file), --detach (return as soon as the outputs are complete and let a worker process
// tear the CUDA context down; default is one process that exits when everything is released; --sync-exit, the old name of
// the default, is still accepted).
#include "cli_common.h"
#include "sibeliaz_graph.h"
#include "sibeliaz_lcb.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cerrno>
#include <cstring>
#include <string>
#include <sys/wait.h>
#include <thread>
#include <unistd.h>
#include <vector>

namespace {

struct Options {
    unsigned k = 25, b = 200, m = 200, t = 1, a = 150, chunks = 0;
    std::string graph, outdir;
    bool noseq = false, have_graph = false, stats = false, detach = false, construct = false;
    int gpu = 0, window = 0;
    std::vector<std::string> fasta;
};

void Usage(FILE *f)
{
    fprintf(f,
            "USAGE:\n   sibeliaz-lcb  [--chunks <integer>] [--noseq] [-o <directory name>] --graph <file name>\n"
            "                 [-a <integer>] [-t <integer>] [-m <integer>] [-b <integer>] [-k <oddc>]\n"
            "                 [--gpu <ordinal>] [--window <seeds>] [--stats] [--detach] [--construct] [--] [--version] [-h]\n"
            "                 <fasta files with genomes> ...\n\n"
            "   SibeliaZ-LCB, a program for construction of locally-collinear blocks from complete genomes\n"
            "   (B200-native implementation; flags and outputs follow SibeliaZ-LCB 1.2.7)\n");
}

using cli::ParseUnsigned;

// returns 0 ok, 1 error (message printed), 2 exit quietly with success (help/version)
int Parse(int argc, char **argv, Options &o)
{
    bool positional_only = false;
    for (int i = 1; i < argc; i++) {
        std::string arg = argv[i];
        std::string val;
        bool has_inline = false;
        if (!positional_only && arg.size() > 2 && arg[0] == '-' && arg[1] == '-') {
            size_t eq = arg.find('=');
            if (eq != std::string::npos) {
                val = arg.substr(eq + 1);
                arg = arg.substr(0, eq);
                has_inline = true;
            }
        }
        auto value = [&](const char *name) -> const char * {
            if (has_inline) return val.c_str();
            if (i + 1 >= argc) {
                fprintf(stderr, "error: Missing a value for this argument! for arg %s\n", name);
                return nullptr;
            }
            return argv[++i];
        };
        auto number = [&](const char *name, unsigned &dst) -> bool {
            const char *v = value(name);
            if (!v) return false;
            if (!ParseUnsigned(v, dst)) {
                fprintf(stderr, "error: Couldn't read argument value from string '%s' for arg %s\n", v, name);
                return false;
            }
            return true;
        };
        if (positional_only || arg.empty() || arg[0] != '-' || arg == "-") {
            o.fasta.push_back(argv[i]);
        } else if (arg == "--") {
            positional_only = true;
        } else if (arg == "-h" || arg == "--help") {
            Usage(stdout);
            return 2;
        } else if (arg == "--version") {
            printf("\nsibeliaz-lcb  version: 1.2.7 (%s)\n\n", lcb_version());
            return 2;
        } else if (arg == "-k" || arg == "--kvalue") {
            if (!number("-k (--kvalue)", o.k)) return 1;
            if (o.k % 2 != 1) {
                fprintf(stderr, "error: Value '%u' does not meet constraint: value of K must be odd for arg -k (--kvalue)\n", o.k);
                return 1;
            }
        } else if (arg == "-b" || arg == "--branchsize") {
            if (!number("-b (--branchsize)", o.b)) return 1;
        } else if (arg == "-m" || arg == "--blocksize") {
            if (!number("-m (--blocksize)", o.m)) return 1;
        } else if (arg == "-t" || arg == "--threads") {
            if (!number("-t (--threads)", o.t)) return 1;
        } else if (arg == "-a" || arg == "--abundance") {
            if (!number("-a (--abundance)", o.a)) return 1;
        } else if (arg == "--chunks") {
            if (!number("--chunks", o.chunks)) return 1;
        } else if (arg == "--graph") {
            const char *v = value("--graph");
            if (!v) return 1;
            o.graph = v;
            o.have_graph = true;
        } else if (arg == "-o" || arg == "--outdir") {
            const char *v = value("-o (--outdir)");
            if (!v) return 1;
            o.outdir = v;
        } else if (arg == "--noseq") {
            o.noseq = true;
        } else if (arg == "--stats") {
            o.stats = true;
        } else if (arg == "--sync-exit") {
            o.detach = false;
        } else if (arg == "--detach") {
            o.detach = true;
        } else if (arg == "--construct") {
            o.construct = true; // no junction file: find the junctions on the GPU too (the twopaco step, fused)
        } else if (arg == "--gpu") {
            unsigned g = 0;
            if (!number("--gpu", g)) return 1;
            o.gpu = (int)g;
        } else if (arg == "--window") {
            unsigned w = 0;
            if (!number("--window", w)) return 1;
            o.window = (int)w;
        } else {
            fprintf(stderr, "error: Couldn't find match for argument for arg %s\n", arg.c_str());
            return 1;
        }
    }
    if (!o.have_graph && !o.construct) {
        fprintf(stderr, "error: Required argument missing: graph for arg --graph\n");
        return 1;
    }
    if (o.fasta.empty()) {
        fprintf(stderr, "error: Required argument missing: filenames for arg (--filenames)\n");
        return 1;
    }
    return 0;
}

double Ms(std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b)
{
    return std::chrono::duration<double, std::milli>(b - a).count();
}

// The whole job.  `done(rc)` is called once the outputs are on disk (or the job has failed): everything after it is
// teardown of multi-GB host/device state that nobody has to wait for.
template <class Done>
int Run(const Options &o, Done done)
{
    char err[1024] = {0};
    auto t0 = std::chrono::steady_clock::now();
    printf("Loading the graph...\n");
    fflush(stdout);
    int device = o.gpu;
    lcb_set_host_threads((int)o.t); // -t: the caller's thread budget (sibeliaz.cpp:81-87)
    if (!getenv("CUDA_VISIBLE_DEVICES")) {
        // context creation touches every visible GPU of the box: show the driver only the one this job uses
        setenv("CUDA_VISIBLE_DEVICES", std::to_string(o.gpu).c_str(), 1);
        device = 0;
    }
    if (o.construct) setenv("LCB_WARM_GRAPH", "1", 1);
    double ms_warm = 0;
    std::thread warm([device, &ms_warm]() { // CUDA context + scratch while the files are parsed
        auto a = std::chrono::steady_clock::now();
        lcb_warmup(device);
        ms_warm = Ms(a, std::chrono::steady_clock::now());
    });
    std::vector<const char *> files;
    for (auto &f : o.fasta) files.push_back(f.c_str());
    lcb_index *index = nullptr;
    int load_rc = o.construct ? lcb_index_load_fasta(files.data(), (int)files.size(), (int)o.k, &index, err, sizeof err)
                              : lcb_index_load(o.graph.c_str(), files.data(), (int)files.size(), (int)o.k, (int)o.a, &index, err, sizeof err);
    auto t_parsed = std::chrono::steady_clock::now();
    int pack_rc = 0;
    if (!load_rc && !o.construct) pack_rc = lcb_index_pack(index); // device record layout, built while the context is still coming up
    auto t_packed = std::chrono::steady_clock::now();
    warm.join();
    if (load_rc) {
        fprintf(stderr, "error: %s\n", err);
        return done(1);
    }
    if (pack_rc) { // (the reference accepts any -a; here a vertex may occur at most 65535 times after the abundance filter)
        fprintf(stderr, "error: a junction occurs more than 65535 times: lower the abundance threshold (-a)\n");
        return done(1);
    }
    auto t1 = std::chrono::steady_clock::now();
    printf("Analyzing the graph...\n");
    fflush(stdout);
    lcb_index_view view;
    memset(&view, 0, sizeof view);
    if (!o.construct) lcb_index_get_view(index, &view);
    lcb_params p;
    lcb_default_params(&p);
    p.k = (int)o.k;
    p.max_branch = (int)o.b;
    p.max_flank = (int)o.b; // sibeliaz.cpp:136 passes maxBranchSize twice
    p.min_block = (int)o.m;
    p.device = device;
    if (o.window > 0) p.window_init = p.window_max = o.window;
    lcb_ctx *ctx = nullptr;
    int rc;
    lcg_graph *graph = nullptr;
    double ms_graph = 0;
    if (o.construct) {
        const uint8_t *const *seq = nullptr;
        const uint64_t *len = nullptr;
        const int32_t n_rec = lcb_index_get_sequences(index, &seq, &len);
        rc = lcg_build_resident(seq, len, n_rec, (int)o.k, UINT64_MAX, device, &graph, err, sizeof err);
        if (rc) {
            fprintf(stderr, "error: %s\n", err);
            return done(1);
        }
        ms_graph = Ms(t1, std::chrono::steady_clock::now());
        rc = lcb_create_from_graph(graph, index, (int)o.a, &p, &ctx);
    } else {
        rc = lcb_create(&view, &p, &ctx);
    }
    auto t_created = std::chrono::steady_clock::now();
    uint64_t n_seeds = 0;
    if (!rc) rc = lcb_enumerate_seeds(ctx, &n_seeds);
    lcb_block_instance *blocks = nullptr;
    uint64_t n_blocks = 0;
    lcb_stats st;
    memset(&st, 0, sizeof st);
    printf("[");
    fflush(stdout);
    if (!rc) rc = lcb_find_blocks(ctx, &blocks, &n_blocks, &st);
    if (rc) {
        fprintf(stderr, "error: %s\n", ctx ? lcb_last_error(ctx) : "cannot create the device context");
        return done(1);
    }
    { // the reference prints one dot per progressPortion_ seeds (blocksfinder.h:362-365,509-513)
        uint64_t portion = n_seeds / 50 ? n_seeds / 50 : 1;
        uint64_t dots = (n_seeds + portion - 1) / portion;
        for (uint64_t i = 0; i < dots; i++) putchar('.');
    }
    printf("]\n");
    auto t2 = std::chrono::steady_clock::now();
    printf("Generating the output...\n");
    fflush(stdout);
    int64_t found = 0;
    double coverage = 0;
    rc = lcb_write_output(index, blocks, n_blocks, (int)o.m, o.outdir.c_str(), !o.noseq, (int)o.chunks, &found, &coverage, err,
                          sizeof err);
    if (rc) {
        fprintf(stderr, "error: %s\n", err);
        return done(1);
    }
    printf("Blocks found: %lld\n", (long long)found);
    printf("Coverage: %.2f\n", coverage);
    auto t3 = std::chrono::steady_clock::now();
    if (o.stats) {
        fprintf(stderr,
                "{\"records\": %llu, \"vertices\": %llu, \"seeds\": %llu, \"block_instances\": %llu, \"windows\": %llu, "
                "\"rounds\": %llu, \"traversals_first\": %llu, \"traversals_rerun\": %llu, \"kernel_launches\": %llu, "
                "\"ms_parse\": %.3f, \"ms_pack\": %.3f, \"ms_graph\": %.3f, \"ms_warmup_thread\": %.3f, \"ms_load\": %.3f, \"ms_create\": %.3f, \"ms_create_enumerate_find\": %.3f, \"ms_enumerate\": %.3f, \"ms_find\": %.3f, "
                "\"ms_traverse_kernels\": %.3f, \"ms_output\": %.3f, \"ms_total\": %.3f, \"junctions_per_sec\": %.1f}\n",
                (unsigned long long)st.n_records, (unsigned long long)st.n_vertices, (unsigned long long)st.n_seeds,
                (unsigned long long)st.n_block_instances, (unsigned long long)st.windows, (unsigned long long)st.rounds,
                (unsigned long long)st.traversals_first, (unsigned long long)st.traversals_rerun,
                (unsigned long long)st.kernel_launches, Ms(t0, t_parsed), Ms(t_parsed, t_packed), ms_graph, ms_warm, Ms(t0, t1), Ms(t1, t_created), Ms(t1, t2), st.ms_enumerate, st.ms_find,
                st.ms_traverse_kernels, Ms(t2, t3), Ms(t0, t3),
                (st.ms_enumerate + st.ms_find) > 0 ? 1000.0 * (double)st.n_records / (st.ms_enumerate + st.ms_find) : 0.0);
    }
    return done(0);
}

} // namespace

int main(int argc, char **argv)
{
    Options o;
    int pr = Parse(argc, argv, o);
    if (pr == 2) return 0;
    if (pr) return 1;
    // Default: one process; it exits when the outputs are on disk and the CUDA context is released.
    // --detach: the job runs in a worker process.  When its outputs are complete it reports the exit status through a pipe
    // and this process returns at once; the worker then releases its CUDA context, device and pinned memory (hundreds of
    // milliseconds on a large GPU) without anybody waiting for it -- the GPU stays held for that long after the return.
    int report_fd = -1;
    if (o.detach) {
        int fds[2];
        if (pipe(fds) == 0) {
            fflush(stdout);
            fflush(stderr);
            const pid_t pid = fork(); // before any CUDA call: the worker initialises CUDA itself
            if (pid > 0) {
                close(fds[1]);
                unsigned char status = 1;
                ssize_t n;
                while ((n = read(fds[0], &status, 1)) < 0 && errno == EINTR) {}
                if (n == 1) return status;
                int ws = 0; // the worker died before reporting
                waitpid(pid, &ws, 0);
                return WIFEXITED(ws) && WEXITSTATUS(ws) ? WEXITSTATUS(ws) : 1;
            }
            if (pid == 0) {
                close(fds[0]);
                report_fd = fds[1];
            } else {
                close(fds[0]);
                close(fds[1]);
            }
        }
    }
    auto done = [report_fd](int rc) {
        fflush(stdout);
        fflush(stderr);
        if (report_fd >= 0) {
            const unsigned char status = (unsigned char)rc;
            ssize_t w = write(report_fd, &status, 1);
            (void)w;
            close(report_fd);
        }
        _exit(rc); // skip destructors: the OS and the driver reclaim everything
        return rc;
    };
    return Run(o, done);
}
