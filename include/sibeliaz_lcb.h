/*
 * sibeliaz_lcb.h -- C ABI of the B200-native sibeliaz-lcb hot path.
 *
 * The reference (medvedevgroup/SibeliaZ v1.2.7) has no FFI for this path: its seams are the process
 * CLI (SibeliaZ-LCB/sibeliaz:146), the junction-file / GFF formats and the C++ call sequence in
 * SibeliaZ-LCB/sibeliaz.cpp:125-143
 *
 *     JunctionStorage storage(graph, fastas, k, threads, abundance, 0);          // :126-131
 *     BlocksFinder finder(storage, k);                                           // :134
 *     finder.FindBlocks(minBlock, maxBranch, maxBranch, 8, 0, threads, path);    // :135-141
 *     finder.GenerateOutput(outDir, !noSeq, chunks);                             // :143
 *
 * Each entry point below replaces one of those calls and says which.  Plain pointers and sizes only;
 * no exception crosses the boundary: every function returns LCB_OK (0) or an LCB_ERR_* code and
 * leaves a message retrievable with lcb_last_error()/lcb_index_last_error().
 *
 * There is no CPU fallback: lcb_create fails with LCB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SIBELIAZ_LCB_H
#define SIBELIAZ_LCB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCB_OK 0
#define LCB_ERR_ARG 1      /* bad argument / inconsistent index view                          */
#define LCB_ERR_IO 2       /* file could not be read or written (reference: runtime_error)    */
#define LCB_ERR_FORMAT 3   /* malformed FASTA / junction file                                  */
#define LCB_ERR_CUDA 4     /* CUDA / NCCL failure, or no usable device                         */
#define LCB_ERR_CAPACITY 5 /* a per-seed device buffer overflowed its hard cap (big arena slot) */
#define LCB_ERR_STATE 6    /* call sequence violated (e.g. find_blocks before create)          */

/* ------------------------------------------------------------------------------------------------
 * Host front end: replaces JunctionStorage::JunctionStorage / ::Init
 * (SibeliaZ-LCB/junctionstorage.h:572-656; wire format common/junctionapi.h:80-98; FASTA rules
 * common/streamfastaparser.cpp:28-92; complement table common/dnachar.cpp:52-85).
 * ---------------------------------------------------------------------------------------------- */
typedef struct lcb_index lcb_index;

/* Structure-of-arrays junction index (host pointers).  g in [0, n_records) enumerates the kept
 * junction records in genome order; chromosome c owns g in [chr_off[c], chr_off[c+1]). */
typedef struct lcb_index_view {
    int32_t n_chr;            /* C                                                              */
    int64_t n_records;        /* N: records with abundance < threshold (junctionstorage.h:610)   */
    int64_t n_vertices;       /* V = max|id| + 1 (junctionstorage.h:586-590)                     */
    const int64_t *chr_off;   /* [C+1]                                                          */
    const int32_t *pos_id;    /* [N] Position.id, signed junction id (junctionstorage.h:142)     */
    const uint32_t *pos_bp;   /* [N] Position.pos (junctionstorage.h:143)                        */
    const uint8_t *next_ch;   /* [N] seq[pos+k] raw byte, 0 at sequence end (:641)               */
    const uint8_t *prev_rc;   /* [N] complement(seq[pos-1]), 'N' at pos 0 (:642)                 */
    const int64_t *vtx_off;   /* [V+1] CSR over the occurrence lists vertex_[|id|] (:695)        */
    const int64_t *occ_g;     /* [N] occurrences of each vertex as g, sorted by (chr, idx) (:646)*/
    /* Optional (may be NULL): the same index already in the device record layout, e.g. built by
     * lcb_index_pack while the CUDA context is still being created; lcb_create then uploads these as they are.
     *   packed_rec[g] = int32x4 {id, bp, first occurrence slot of |id|, (#occurrences << 16) | (next_ch << 8) | prev_rc}
     *   packed_occ[o] = int32x2 {g | (stored id < 0 ? 1 << 31 : 0), bp}   in occ_g order */
    const void *packed_rec;   /* [N] 16-byte records                                             */
    const void *packed_occ;   /* [N] 8-byte records                                              */
} lcb_index_view;

int lcb_index_load(const char *graph_file, const char *const *fasta_files, int n_fasta, int k, int abundance,
                   lcb_index **out, char *err, size_t errlen);
/* Fused pipeline (no junction file): FASTA records only; the junction part of the index is then built ON THE DEVICE by
 * lcb_create_from_graph from a graph made by lcg_build_resident (include/sibeliaz_graph.h) out of these very records. */
int lcb_index_load_fasta(const char *const *fasta_files, int n_fasta, int k, lcb_index **out, char *err, size_t errlen);
/* Record r is seq[r][0 .. len[r]); the arrays live as long as the index.  Returns the number of records. */
int32_t lcb_index_get_sequences(const lcb_index *, const uint8_t *const **seq, const uint64_t **len);

/* Builds the device record layout on the host threads (optional; lcb_create packs by itself otherwise). */
int lcb_index_pack(lcb_index *);
int lcb_index_get_view(const lcb_index *, lcb_index_view *view);
int32_t lcb_index_num_chr(const lcb_index *);
const char *lcb_index_chr_name(const lcb_index *, int32_t chr);
int64_t lcb_index_chr_length(const lcb_index *, int32_t chr);
void lcb_index_free(lcb_index *);

/* ------------------------------------------------------------------------------------------------
 * Device context: replaces BlocksFinder::BlocksFinder + ::FindBlocks
 * (SibeliaZ-LCB/blocksfinder.h:213-217, :453-530 and everything it calls in path.h).
 * ---------------------------------------------------------------------------------------------- */
typedef struct lcb_ctx lcb_ctx;

typedef struct lcb_params {
    int32_t k;             /* -k (sibeliaz.cpp:49)                                              */
    int32_t max_branch;    /* -b: maxBranchSize (sibeliaz.cpp:135)                              */
    int32_t min_block;     /* -m: minBlockSize (sibeliaz.cpp:134)                               */
    int32_t max_flank;     /* maxFlankingSize; the CLI passes -b again (sibeliaz.cpp:136)       */
    int32_t looking_depth; /* hard-coded 8 in the reference (sibeliaz.cpp:137)                  */
    int32_t phase_size;    /* hard-coded 256 (blocksfinder.h:519); part of the output semantics */
    int32_t window_init;   /* seeds admitted per round at the start (multiple of phase_size); 0 = default */
    int32_t window_max;    /* bound on the active (speculated, uncommitted) seeds; 0 = default  */
    int32_t device;        /* CUDA device ordinal                                               */
    int32_t collect_counters; /* 1: also count walk / occurrence / scan / score steps on device */
} lcb_params;

typedef struct lcb_block_instance { /* BlockInstance(id, chr, start, end), blocksfinder.h:33    */
    int32_t id;                     /* +-block id in commit order, sign = strand (blocksfinder.h:320-324) */
    uint32_t chr;
    uint32_t start;
    uint32_t end;
} lcb_block_instance;

typedef struct lcb_stats {
    uint64_t n_records, n_vertices, n_seeds;
    uint64_t n_block_instances, n_blocks;  /* blocksInstance_.size(), blocksFound_               */
    uint64_t windows, rounds;              /* times the active set started from empty / evaluation rounds */
    uint64_t traversals_first, traversals_rerun; /* Process() evaluations: phase-snapshot / commit-time */
    uint64_t kernel_launches;              /* kernels of this library launched by find_blocks+enumerate */
    uint64_t t_walk, t_occ, t_scan, t_score; /* device-side step counts (collect_counters=1)      */
    double ms_enumerate;                   /* seed enumeration + sort, device time                */
    double ms_find;                        /* traversal + commit, host wall time                  */
    double ms_traverse_kernels;            /* sum of CUDA-event durations of the traversal kernel */
    uint64_t traverse_launches;
    double ms_h2d, ms_d2h;                 /* index upload / result download                      */
    double ms_step_device;                 /* CUDA-event time from the start of seed enumeration to the end of
                                              find_blocks on the library's stream (when find_blocks enumerates)  */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t pool_restarts;                /* active sets abandoned because a result pool ran full */
    uint64_t big_arena_runs;               /* evaluations that outgrew the per-warp scratch and were re-run in a big slot */
    uint64_t lean_runs, lean_bails;        /* evaluations finished by the common-case kernel / handed on to the general kernel */
    uint64_t lean_bail_why[8];             /* ... by reason: > 32 occurrences of a vertex, > 128 path vertices, two occurrences on one
                                              chromosome, > 32 instances, read-set log full, look-ahead walk > 24 junctions, vote table
                                              full, |path distance| >= 2^30 */
    double ms_tail[6];                     /* the rest of the rounds on the device (%globaltimer, summed): gap after the traversal,
                                              epochs of committed seeds, claims, change map, validation (+ peer control), commit */
} lcb_stats;

void lcb_default_params(lcb_params *p);

/* Optional: create the CUDA context on `device` and pre-allocate the index-independent scratch, e.g. on a
 * thread while the host parses its inputs. */
int lcb_warmup(int device);

/* Uploads the index (arrays are caller-owned and may be freed once this returns). */
int lcb_create(const lcb_index_view *index, const lcb_params *params, lcb_ctx **out);

/* Fused pipeline: the junction index (abundance filter, occurrence lists, edge characters: JunctionStorage::Init,
 * junctionstorage.h:572-650) is built on the device from a resident graph; `index` (from lcb_index_load_fasta on the same
 * records) learns the number of chromosomes the junctions span and serves lcb_write_output afterwards. */
struct lcg_graph;
int lcb_create_from_graph(const struct lcg_graph *graph, lcb_index *index, int abundance, const lcb_params *params, lcb_ctx **out);

/* Multi-GPU (optional): one process per GPU.  Rank 0 obtains an id with lcb_comm_unique_id, the
 * caller distributes the bytes (e.g. torch.distributed broadcast), every rank calls lcb_comm_init. */
#define LCB_COMM_ID_BYTES 128
int lcb_comm_unique_id(void *id_bytes);
int lcb_comm_init(lcb_ctx *, int rank, int n_ranks, const void *id_bytes);

/* Multi-GPU creation in one call: rank 0 uploads the index once and the other ranks receive it over NVLink (NCCL
 * broadcast) instead of pushing the same arrays over PCIe n_ranks times; `index` is read on rank 0 only (may be NULL
 * elsewhere).  Includes lcb_comm_init.  Collective: every rank calls it. */
int lcb_create_shared(const lcb_index_view *index, const lcb_params *params, int rank, int n_ranks, const void *id_bytes,
                      lcb_ctx **out);

/* Bundle enumeration + sort (blocksfinder.h:461-503,517). */
int lcb_enumerate_seeds(lcb_ctx *, uint64_t *n_seeds);
/* Parity hook: copies the sorted seed list (Bundle fields, blocksfinder.h:182-192); any pointer may be NULL. */
int lcb_get_seeds(lcb_ctx *, int64_t *vid, uint8_t *ch, uint64_t *count, uint64_t *rank, uint64_t *res_pos,
                  uint64_t *res_chr);

/* Forget the enumerated seeds so that the next lcb_find_blocks repeats the whole FindBlocks equivalent
 * (enumeration + sort + traversal + commit) on the resident index; used to time repeated passes. */
int lcb_reset_seeds(lcb_ctx *);

/* Traversal + ordered commit (blocksfinder.h:228-433).  *out is library-owned until lcb_free_blocks;
 * records are in the reference's commit order (that order feeds the output stage's unstable sorts). */
int lcb_find_blocks(lcb_ctx *, lcb_block_instance **out, uint64_t *n, lcb_stats *stats);
void lcb_free_blocks(lcb_block_instance *);
int lcb_get_stats(lcb_ctx *, lcb_stats *stats);
const char *lcb_last_error(lcb_ctx *);
void lcb_destroy(lcb_ctx *);

/* ------------------------------------------------------------------------------------------------
 * Output stage: replaces BlocksFinder::GenerateOutput (blocksfinder.h:605-670),
 * ListBlocksIndicesGFF (blocksfinder.cpp:141-174) and ListBlocksSequences (blocksfinder.h:533-582).
 * ---------------------------------------------------------------------------------------------- */
int lcb_write_output(const lcb_index *, const lcb_block_instance *blocks, uint64_t n, int min_block,
                     const char *out_dir, int gen_seq, int chunks, int64_t *blocks_found, double *coverage,
                     char *err, size_t errlen);

/* Upper bound on the host threads the library uses for parsing, packing and output formatting: the `threads`
 * argument of JunctionStorage / FindBlocks (sibeliaz.cpp:126-141, the CLI's -t).  n <= 0: min(cores, 32). */
void lcb_set_host_threads(int n);

/* Returns the process-wide cache of device scratch / pinned staging blocks to the driver. */
void lcb_trim_cache(void);

const char *lcb_version(void);

#ifdef __cplusplus
}
#endif
#endif
