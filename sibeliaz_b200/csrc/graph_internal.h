// graph_internal.h -- seam between the host front end (graph_host.cpp, no CUDA) and the device pipeline
// (graph_device.cu) of the junction finder.  Not part of the public C ABI (include/sibeliaz_graph.h).
#pragma once
#include <cstdint>
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
};

struct DeviceOutput {
    // every candidate position (more than one in- or out-edge), in genome order, as an index into G, with the signed
    // vertex id of its k-mer or 0 when the k-mer is not a bifurcation
    std::vector<uint64_t> pos;
    std::vector<int32_t> id;
    lcg_stats st{};
};

int run_device(const DeviceInput &in, DeviceOutput &out, std::string &err);

} // namespace lcg
