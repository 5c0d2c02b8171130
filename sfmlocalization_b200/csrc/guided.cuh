// guided.cuh -- guided matching between two descriptor tables (guided.cu), for callers inside the
// library: hulo_guided_match has the same table on both sides, the per-query engine matches the
// views of the resident map (I side) against the query image (J side).
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <vector>

#include "context.cuh"

namespace hulo {

struct GuidedGroups;

// One side of the pairs: a table of folded descriptor rows on the device, its segments (images),
// and the feature positions of its rows on both sides of the bus.
struct GuidedSide {
    const uint4 *rows = nullptr;      // device
    const uint64_t *seg = nullptr;    // host: n_seg + 1 row offsets
    size_t n_seg = 0;
    const double *h_xy = nullptr;     // host, row order (duplicate-position filter)
    const double2 *d_xy = nullptr;    // device, row order
    GuidedGroups *groups = nullptr;   // optional cache of the position groups (survives the call)
};

// rep[f] = first feature of the image at the same float position, member[f] = position shared
struct GuidedGroups {
    struct Seg { std::vector<int32_t> rep; std::vector<uint8_t> member; };
    std::map<uint32_t, Seg> segs;
    const Seg &of(const GuidedSide &side, uint32_t S);
};

// pairs: (segment of the I side, segment of the J side).  Everything else as hulo_guided_match.
int guided_match_sides(hulo_gpu *h, const GuidedSide &SI, const GuidedSide &SJ, const uint32_t *pairs, size_t n_pairs,
                       const double *F, const double *error_th, double dist_ratio, int dedup, uint64_t *pair_offsets,
                       uint32_t *out_i, uint32_t *out_j, size_t cap, size_t *n_out);

}  // namespace hulo
