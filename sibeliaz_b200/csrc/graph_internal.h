// graph_internal.h -- seam between the host front end (graph_host.cpp, no CUDA) and the device pipeline
// (graph_device.cu) of the junction finder.  Not part of the public C ABI (include/sibeliaz_graph.h).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "sibeliaz_graph.h"

namespace lcg {

// Global text layout on the device: G = 'N' rec0 'N' rec1 'N' ... recR-1 'N'  (one separator does duty as the trailing
// 'N' of a record and the leading 'N' of the next, vertexenumerator.h:1158,1187).  Record r starts at
// goff[r] = 1 + sum_{j<r} (len[j] + 1).
struct DeviceInput {
    const uint8_t *const *seq;
    const uint64_t *len;
    int n_records;
    int k;
    uint64_t abundance;
    int device;
    bool keep_on_device = false;
};

// What stays on the device after a build with keep_on_device (the fused pipeline hands it to lcb_create_from_graph):
// the record bytes in the layout G, and the junction records.  Owned by the lcg_graph; released by lcg_free.
struct Resident {
    int device = 0, k = 0, n_records = 0;
    uint8_t *d_text = nullptr;   // G, upper/lower case as given
    uint64_t *d_goff = nullptr;  // [n_records + 1]
    uint32_t *d_chr = nullptr, *d_pos = nullptr;
    int32_t *d_id = nullptr;
    uint64_t n_junctions = 0;
    uint64_t n_vertices = 0;     // max |id| + 1
    uint32_t last_chr = 0;       // record index of the last junction
};

struct DeviceOutput {
    // the junction records in genome order: record index, position in the record, signed vertex id (stubs included)
    uint64_t n = 0;
    std::unique_ptr<uint32_t[]> chr, pos;
    std::unique_ptr<int32_t[]> id;
    lcg_stats st{};
    Resident resident; // filled when DeviceInput::keep_on_device
};

int run_device(const DeviceInput &in, DeviceOutput &out, std::string &err);
void free_resident(Resident &r);
int download(const Resident &r, uint32_t *chr, uint32_t *pos, int32_t *id, std::string &err); // junction records -> host
void preload_kernels();
const Resident *resident_of(const lcg_graph *g); // nullptr when the graph was built without keep_on_device

} // namespace lcg
