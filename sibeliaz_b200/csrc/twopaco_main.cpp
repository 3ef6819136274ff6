// twopaco (B200): drop-in for the graph-construction step of the sibeliaz wrapper (SibeliaZ-LCB/sibeliaz:145),
//   twopaco --tmpdir <dir> -t <threads> -k <k> --filtermemory <GB> -o <file> <fasta...>
// Same flags as TwoPaCo/src/graphconstructor/constructor.cpp:58-143; the Bloom-filter and threading knobs (-f,
// --filtermemory, -q, -r, -t, --tmpdir) are accepted and ignored: the junctions are found exactly, on the GPU, through
// libsibeliaz_lcb's C ABI (include/sibeliaz_graph.h).  Additive flags: --gpu <ordinal>, --stats.
#include "cli_common.h"
#include "sibeliaz_graph.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int main(int argc, char **argv)
{
    unsigned k = 25;
    uint64_t abundance = UINT64_MAX;
    std::string out = "de_bruijn.bin";
    std::vector<std::string> files;
    bool have_filter = false, stats = false;
    int gpu = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto value = [&](const char *name) -> const char * {
            if (i + 1 >= argc) {
                fprintf(stderr, "error: Missing a value for this argument! for arg %s\n", name);
                exit(1);
            }
            return argv[++i];
        };
        if (a == "-k" || a == "--kvalue") {
            const char *v = value("-k (--kvalue)");
            if (!cli::ParseUnsigned(v, k)) {
                cli::BadValue(v, "-k (--kvalue)");
                return 1;
            }
            if (k % 2 != 1) {
                fprintf(stderr, "error: Value '%u' does not meet constraint: value of K must be odd for arg -k (--kvalue)\n", k);
                return 1;
            }
        } else if (a == "-f" || a == "--filtersize" || a == "--filtermemory") {
            value(a.c_str());
            have_filter = true;
        } else if (a == "-q" || a == "--hashfnumber" || a == "-r" || a == "--rounds" || a == "-t" || a == "--threads" || a == "--tmpdir") {
            value(a.c_str());
        } else if (a == "-a" || a == "--abundance") {
            const char *v = value("-a (--abundance)");
            if (!cli::ParseU64(v, abundance)) {
                cli::BadValue(v, "-a (--abundance)");
                return 1;
            }
        } else if (a == "-o" || a == "--outfile") {
            out = value("-o (--outfile)");
        } else if (a == "--gpu") {
            const char *v = value("--gpu");
            if (!cli::ParseInt(v, gpu) || gpu < 0) {
                cli::BadValue(v, "--gpu");
                return 1;
            }
        } else if (a == "--stats") {
            stats = true;
        } else if (a == "--test") {
            fprintf(stderr, "error: the self-test lives in the test-suite of this implementation (pytest tests -m gpu -k graph)\n");
            return 1;
        } else if (a == "--version") {
            printf("\ntwopaco  version: 1.1.0 (B200 junction finder)\n\n");
            return 0;
        } else if (a == "-h" || a == "--help") {
            printf("USAGE:\n   twopaco  {-f <integer>|--filtermemory <float>} [-o <file name>] [--tmpdir <directory name>] [-a <integer>]\n"
                   "            [-t <integer>] [-r <integer>] [-q <integer>] [-k <oddc>] [--gpu <ordinal>] [--stats] <fasta files with genomes> ...\n");
            return 0;
        } else if (a == "--") {
            for (++i; i < argc; i++) files.push_back(argv[i]);
        } else if (!a.empty() && a[0] == '-' && a != "-") {
            fprintf(stderr, "error: Couldn't find match for argument for arg %s\n", a.c_str());
            return 1;
        } else {
            files.push_back(a);
        }
    }
    if (!have_filter) {
        fprintf(stderr, "error: One (and only one) of -f (--filtersize) or --filtermemory is required\n");
        return 1;
    }
    if (files.empty()) {
        fprintf(stderr, "error: Required argument missing: filenames\n");
        return 1;
    }
    std::vector<const char *> fa;
    for (auto &f : files) fa.push_back(f.c_str());
    char err[1024] = {0};
    lcg_graph *g = nullptr;
    int rc = lcg_build_from_fasta(fa.data(), (int)fa.size(), (int)k, abundance, gpu, &g, err, sizeof err);
    if (!rc) rc = lcg_write_junction_file(g, out.c_str(), err, sizeof err);
    if (rc) {
        fprintf(stderr, "error: %s\n", err);
        return 1;
    }
    lcg_stats st;
    lcg_get_stats(g, &st);
    printf("Vertex length = %u\nTrue junctions count = %llu\nTrue marks count: %llu\n", k, (unsigned long long)st.n_bifurcations,
           (unsigned long long)st.n_junctions);
    if (stats)
        fprintf(stderr,
                "{\"records\": %llu, \"bases\": %llu, \"kmers\": %llu, \"distinct\": %llu, \"candidates\": %llu, \"bifurcations\": %llu, "
                "\"junctions\": %llu, \"table_slots\": %llu, \"ms_parse\": %.3f, \"ms_h2d\": %.3f, \"ms_device\": %.3f, \"ms_edges\": %.3f, "
                "\"ms_d2h\": %.3f, \"ms_total\": %.3f}\n",
                (unsigned long long)st.n_records, (unsigned long long)st.n_bases, (unsigned long long)st.n_kmers,
                (unsigned long long)st.n_distinct, (unsigned long long)st.n_candidates, (unsigned long long)st.n_bifurcations,
                (unsigned long long)st.n_junctions, (unsigned long long)st.table_slots, st.ms_parse, st.ms_h2d, st.ms_device, st.ms_edges,
                st.ms_d2h, st.ms_total);
    lcg_free(g);
    return 0;
}
