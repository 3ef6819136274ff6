// lcb_internal.h -- seams inside libsibeliaz_lcb that are not part of the public C ABI.
#pragma once
#include <cstdint>
#include <string>

struct lcb_index;

// lcb_host.cpp: the fused pipeline tells a FASTA-only index how many chromosomes the junctions span (what
// lcb_index_load derives from the junction file, junctionstorage.h:577-583).  Fails when the FASTA holds fewer records.
int lcb_index_set_chromosomes(lcb_index *ix, int32_t n_chr, int k, std::string &err);

// lcb_device.cu: the process-wide cache of device scratch blocks (contexts return their allocations to it in lcb_destroy).
// The other stages size themselves from cudaMemGetInfo, which cannot see these blocks: they ask how much is parked and
// release it when they need the room.
size_t lcb_cache_device_bytes(int device);
void lcb_cache_trim_device(int device); // device < 0: the pinned host blocks
