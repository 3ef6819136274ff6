/*
 * sibeliaz_graph.h -- C ABI of the B200-native junction finder: the step that feeds the sibeliaz-lcb hot path
 * (SURVEY.md section 8f, row 1).  It replaces the `twopaco` invocation of the sibeliaz wrapper (SibeliaZ-LCB/sibeliaz:145)
 * and, behind it, TwoPaCo's VertexEnumeratorImpl (TwoPaCo/src/graphconstructor/vertexenumerator.h:122-466): from FASTA
 * records to the junction positions of the compacted de Bruijn graph with a consistent (vertex id, strand) label each,
 * in the wire format of TwoPaCo/src/common/junctionapi.h:106-136.
 *
 * What is computed is the reference's result in the limit of no Bloom-filter false positives (exact k-mer table in HBM,
 * no filter, no temporary files); vertex ids are deterministic here (1 + rank of the canonical k-mer), whereas the
 * reference's change from run to run (its hash seeds come from /dev/urandom) -- see oracle/graph_oracle.cpp for the
 * rules with their file:line citations and for the label-free normal form in which outputs are compared.
 *
 * Plain C types only; every call returns 0 on success or an LCG_ERR_* code and fills `err`.  No CPU fallback:
 * lcg_build* fail with LCG_ERR_CUDA without an sm_100-class device.
 */
#ifndef SIBELIAZ_GRAPH_H
#define SIBELIAZ_GRAPH_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LCG_OK 0
#define LCG_ERR_ARG 1
#define LCG_ERR_IO 2     /* reference: "Can't open file", "Can't create the output file"                 */
#define LCG_ERR_FORMAT 3 /* malformed FASTA (streamfastaparser.cpp:24,35,82)                              */
#define LCG_ERR_CUDA 4
#define LCG_ERR_MEMORY 5 /* the k-mer table does not fit the device                                       */

typedef struct lcg_graph lcg_graph;

typedef struct lcg_stats {
    uint64_t n_records;       /* FASTA records                                                            */
    uint64_t n_bases;         /* characters in them                                                       */
    uint64_t n_kmers;         /* definite k-mer positions                                                 */
    uint64_t n_distinct;      /* distinct canonical k-mers                                                */
    uint64_t n_candidates;    /* positions with more than one in- or out-edge                             */
    uint64_t n_bifurcations;  /* junction vertices (TwoPaCo: "True junctions count")                      */
    uint64_t n_junctions;     /* records written, stubs included (TwoPaCo: "True marks count")            */
    uint64_t table_slots;     /* capacity of the k-mer table                                              */
    uint64_t kernel_launches;
    double ms_parse;          /* FASTA -> sequences (host threads)                                        */
    double ms_h2d;            /* sequences -> device                                                      */
    double ms_device;         /* CUDA-event time from the first to the last kernel                        */
    double ms_edges;          /* of which: the table-building pass (the HBM-bound kernel)                 */
    double ms_d2h;            /* junction list -> host                                                    */
    double ms_total;
} lcg_stats;

/* FASTA files -> junction list.  k odd, 1 <= k <= 255 (a k-mer is one 64-bit word up to k = 31, two to eight words
 * beyond; TwoPaCo's CAPACITY template, vertexenumerator.cpp:20-58).  abundance: vertices with more candidate occurrences are dropped
 * (twopaco -a; pass UINT64_MAX for the wrapper's behaviour). */
int lcg_build_from_fasta(const char *const *fasta_files, int n_files, int k, uint64_t abundance, int device,
                         lcg_graph **out, char *err, size_t errlen);

/* Same from sequences already in memory: record r is seq[r][0 .. len[r]) (any case; non-ACGT counts as 'N'). */
int lcg_build(const uint8_t *const *seq, const uint64_t *len, int n_records, int k, uint64_t abundance, int device,
              lcg_graph **out, char *err, size_t errlen);

/* As lcg_build, and the record bytes and the junction records also STAY on the device, for lcb_create_from_graph
 * (include/sibeliaz_lcb.h): the fused pipeline, no junction file in between.  lcg_free releases them. */
int lcg_build_resident(const uint8_t *const *seq, const uint64_t *len, int n_records, int k, uint64_t abundance, int device,
                       lcg_graph **out, char *err, size_t errlen);

uint64_t lcg_num_junctions(const lcg_graph *);
/* Junction records in genome order (JunctionPosition: chr, pos, id; junctionapi.h:10-39); any pointer may be NULL. */
int lcg_get_junctions(const lcg_graph *, uint32_t *chr, uint32_t *pos, int64_t *id);
/* Writes the junction file sibeliaz-lcb reads (JunctionPositionWriter, junctionapi.h:106-136). */
int lcg_write_junction_file(const lcg_graph *, const char *path, char *err, size_t errlen);
int lcg_get_stats(const lcg_graph *, lcg_stats *);
void lcg_free(lcg_graph *);

#ifdef __cplusplus
}
#endif
#endif
