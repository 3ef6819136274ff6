// sibeliaz-align -- the alignment stage of the pipeline as one command (include/sibeliaz_align.h):
//
//   sibeliaz-align --cmd "<the wrapper's argument string>" -o <outdir>/alignment.maf [--gpu <ordinal>] [--cleanup] [--stats]
//                  <outdir>/*.tmp
//
// replaces the body of global_alignment() in SibeliaZ-LCB/sibeliaz:118-134 (one `spoa` process per block behind
// xargs/bash, the per-chunk .msa files, their sorted concatenation): same alignment.maf, byte for byte.
// Exit status 0, or 1 with "error: <message>" on stderr like the other binaries of the pipeline.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>
#include <vector>

#include "cli_common.h"
#include "sibeliaz_align.h"

int main(int argc, char **argv)
{
    std::string cmd, out;
    std::vector<const char *> files;
    bool cleanup = false, stats = false;
    lca_params p;
    lca_default_params(&p);
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto value = [&](const char *name) -> const char * {
            if (i + 1 >= argc) {
                fprintf(stderr, "error: missing value for %s\n", name);
                exit(1);
            }
            return argv[++i];
        };
        auto integer = [&](const char *name, int &dst) {
            const char *v = value(name);
            if (!cli::ParseInt(v, dst)) {
                cli::BadValue(v, name);
                exit(1);
            }
        };
        if (a == "--cmd") cmd = value("--cmd");
        else if (a == "-o" || a == "--out") out = value("-o");
        else if (a == "--gpu") integer("--gpu", p.device);
        else if (a == "-m") integer("-m", p.match);
        else if (a == "-n") integer("-n", p.mismatch);
        else if (a == "-g" || a == "-e") integer("-g", p.gap);
        else if (a == "--cleanup") cleanup = true;
        else if (a == "--stats") stats = true;
        else if (a == "-h" || a == "--help") {
            printf("usage: sibeliaz-align --cmd <string> -o <alignment.maf> [--gpu n] [--cleanup] [--stats] <chunk.tmp>...\n");
            return 0;
        } else files.push_back(argv[i]);
    }
    if (out.empty()) {
        fprintf(stderr, "error: missing -o <alignment.maf>\n");
        return 1;
    }
    char err[1024] = {0};
    lca_stats st;
    const int rc = lca_align_chunk_files(files.data(), (int)files.size(), cmd.c_str(), out.c_str(), &p, &st, err, sizeof err);
    if (rc) {
        fprintf(stderr, "error: %s\n", err);
        return 1;
    }
    if (cleanup)
        for (const char *f : files) unlink(f); // the wrapper removes the chunk files once alignment.maf exists (sibeliaz:132)
    if (stats)
        fprintf(stderr, "{\"blocks\": %llu, \"copies\": %llu, \"bases\": %llu, \"cells\": %llu, \"ms_kernels\": %.3f, \"ms_total\": %.3f, "
                        "\"levels\": [%llu, %llu, %llu]}\n",
                (unsigned long long)st.n_blocks, (unsigned long long)st.n_copies, (unsigned long long)st.n_bases, (unsigned long long)st.cells,
                st.ms_kernels, st.ms_total, (unsigned long long)st.blocks_level[0], (unsigned long long)st.blocks_level[1],
                (unsigned long long)st.blocks_level[2]);
    return 0;
}
