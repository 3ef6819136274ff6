/*
 * lcb_oracle.h -- C interface of the CPU restatement oracle for the sibeliaz-lcb hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The oracle is pinned (tests/test_oracle_pin.py) against
 *   (1) the reference's own golden  examples/sibeliaz_out/blocks_coords.gff  (k=25 defaults), and
 *   (2) outputs of the unmodified reference compiled into oracle/_ref/ (oracle/Makefile).
 *
 * Every function restates a piece of /root/reference/SibeliaZ-LCB; citations are in lcb_oracle.cpp.
 */
#ifndef LCB_ORACLE_H
#define LCB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lcbo lcbo;

/* JunctionStorage::Init (junctionstorage.h:572-650).  Returns NULL and fills err on failure. */
lcbo *lcbo_load(const char *graph, const char *const *fastas, int n_fastas, int k, int abundance,
                char *err, int errlen);
void lcbo_free(lcbo *);

int64_t lcbo_num_records(const lcbo *);  /* N: junction records kept by the abundance filter   */
int64_t lcbo_num_vertices(const lcbo *); /* V = vertex_.size() = max|id|+1                     */
int32_t lcbo_num_chr(const lcbo *);      /* C: chromosomes                                      */

/* Structure-of-arrays view of the index (SURVEY.md section 7 layout); caller allocates. */
void lcbo_get_index(const lcbo *, int64_t *chr_off /*C+1*/, int32_t *pos_id /*N*/, uint32_t *pos_bp /*N*/,
                    uint8_t *next_ch /*N*/, uint8_t *prev_rc /*N*/, int64_t *vtx_off /*V+1*/,
                    int64_t *occ_g /*N*/, int64_t *chr_len /*C*/);

/* Bundle enumeration + sort (blocksfinder.h:461-503,517).  Returns S. */
int64_t lcbo_enumerate_seeds(lcbo *);
void lcbo_get_seeds(const lcbo *, int64_t *vid, uint8_t *ch, uint64_t *count, uint64_t *rank,
                    uint64_t *res_pos, uint64_t *res_chr);

/* BlocksFinder::FindBlocks traversal + ordered commit (blocksfinder.h:228-433,505-527), executed
 * sequentially with exactly the reference's phase semantics.  Returns the number of block
 * instances (blocksInstance_.size()), in commit order. */
int64_t lcbo_find_blocks(lcbo *, int min_block, int max_branch, int max_flank, int looking_depth,
                         int phase_size);
void lcbo_get_blocks(const lcbo *, int32_t *id, uint32_t *chr, uint64_t *start, uint64_t *end);

/* counters[0..7] = T_walk, T_occ, T_scan, T_score, process_calls, reruns, mpv_calls, pushes
 * (SURVEY.md section 8d / Appendix B definitions). */
void lcbo_get_counters(const lcbo *, uint64_t *counters8);

/* BlocksFinder::GenerateOutput (blocksfinder.h:605-670, blocksfinder.cpp:141-174). */
int lcbo_generate_output(lcbo *, const char *outdir, int gen_seq, int chunks, int min_block,
                         int64_t *blocks_found, double *coverage, char *err, int errlen);

#ifdef __cplusplus
}
#endif
#endif
