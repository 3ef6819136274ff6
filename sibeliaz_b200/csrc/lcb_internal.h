// lcb_internal.h -- seams inside libsibeliaz_lcb that are not part of the public C ABI.
#pragma once
#include <cstdint>
#include <string>

struct lcb_index;

// lcb_host.cpp: the fused pipeline tells a FASTA-only index how many chromosomes the junctions span (what
// lcb_index_load derives from the junction file, junctionstorage.h:577-583).  Fails when the FASTA holds fewer records.
int lcb_index_set_chromosomes(lcb_index *ix, int32_t n_chr, int k, std::string &err);
