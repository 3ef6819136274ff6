/* sibeliaz_align.h -- C ABI of the alignment stage after the LCB path.
 *
 * What it replaces: the wrapper's global_alignment() (SibeliaZ-LCB/sibeliaz:118-134), which runs
 *     spoa <block.fa> -l 1 -r 1 -e -8                                   (sibeliaz:66)
 * once per block of every <i>.tmp chunk file the LCB step wrote (blocksfinder.h:533-582) and pastes the rows behind the
 * block's headers into alignment.maf (sibeliaz:64-100).  spoa = partial order alignment: every copy of a block is aligned
 * globally (linear gaps) against the graph of the copies before it (spoa/src/sisd_alignment_engine.cpp:295-456), merged
 * (spoa/src/graph.cpp:156-246), and the MSA rows are read off the graph (graph.cpp:303-357).
 * Here: one warp per block on the GPU, thousands of blocks in flight; results are byte-identical to spoa's.
 *
 * No exception crosses this boundary; every function returns 0 or an LCA_ERR_* code and fills `err`.
 * There is no CPU fallback: without an sm_100-class device lca_align* return LCA_ERR_CUDA.
 */
#ifndef SIBELIAZ_ALIGN_H
#define SIBELIAZ_ALIGN_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LCA_OK 0
#define LCA_ERR_ARG 1
#define LCA_ERR_IO 2
#define LCA_ERR_FORMAT 3
#define LCA_ERR_CUDA 4
#define LCA_ERR_CAPACITY 5 /* one block needs more device memory than there is (the message names it) */
#define LCA_ERR_INTERNAL 6

typedef struct lca_params {
    int match;    /* spoa -m, default 5   (spoa/src/main.cpp:209)                                          */
    int mismatch; /* spoa -n, default -4                                                                    */
    int gap;      /* spoa -g = -e: the pipeline passes -e -8 next to the default -g -8, i.e. linear gaps   */
    int device;   /* CUDA ordinal                                                                           */
} lca_params;

typedef struct lca_stats {
    uint64_t n_blocks, n_copies, n_bases;
    uint64_t cells;            /* dynamic-programming cells filled ((nodes + 1) * (len + 1) per aligned copy)   */
    uint64_t blocks_level[3];  /* blocks finished in the optimistic / roomier / worst-case arena               */
    uint64_t kernel_launches;
    double ms_kernels;         /* CUDA-event time of the alignment kernels                                     */
    double ms_total;           /* host wall time of the call                                                   */
    uint64_t h2d_bytes, d2h_bytes;
} lca_stats;

typedef struct lca_result lca_result;

void lca_default_params(lca_params *p);

/* Blocks as flat arrays: copy c is seq[copy_off[c] .. copy_off[c+1]) (any bytes; compared as they are, like spoa does);
 * block b consists of the copies block_off[b] .. block_off[b+1]-1, in the order spoa would read them. */
int lca_align(const uint8_t *seq, const uint64_t *copy_off, uint64_t n_copies, const uint32_t *block_off, uint32_t n_blocks,
              const lca_params *params, lca_result **out, char *err, size_t errlen);

/* MSA row of copy c: rows[row_off[c] .. row_off[c+1]), '-' = gap; all rows of a block have lca_block_columns(b) bytes. */
const uint8_t *lca_rows(const lca_result *, const uint64_t **row_off);
uint32_t lca_block_columns(const lca_result *, uint32_t block);
void lca_get_stats(const lca_result *, lca_stats *);
void lca_free(lca_result *);

/* The whole stage: every line of every chunk file is one block ("> hdr;start;len;strand;size@SEQ@" per copy); writes
 * out_maf exactly as global_alignment() does: "##maf version=1", "# sibeliaz v1.2.7 ", "# cmd=<cmd>", then per chunk file
 * (in C-locale order of the paths) per block an empty line, "a", and "s hdr start len strand size row" lines. */
int lca_align_chunk_files(const char *const *files, int n_files, const char *cmd, const char *out_maf, const lca_params *params,
                          lca_stats *stats, char *err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif
