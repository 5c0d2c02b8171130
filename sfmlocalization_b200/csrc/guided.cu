// guided.cu -- K1g: guided matching of the F-matrix geometric filter, batched over image pairs,
// sm_100a.
//
// Stands behind hulo::geometricMatch(..., bGuided_matching = true)
// (VisionLocalizeCommon/src/MatchUtils.cpp:372-420, -gm of ExtFeatAndMatch and of the
// localisation CLIs; the reconstruction drivers switch it on by default,
// PyVisionLocalizeCommon/src/hulo_param/ReconstructParam.py:70-71), i.e. OpenMVG 1.1's
// GeometricFilter_FMatrix_AC::Geometry_guided_matching -> geometry_aware::GuidedMatching with
// Regions (robust_estimation/guided_matching.hpp): for every feature i of image I, among the
// features j of image J whose squared distance to the epipolar line F x_i is below the robust
// precision, the nearest and second nearest descriptor; kept iff a second exists and
// best < ratio * second.  OpenMVG is third-party and not vendored; the test suite's CPU
// restatement is the checker.
//
// It is the brute-force matcher again, gated by geometry: one thread owns one feature of I
// (descriptor in registers, epipolar line in fp64), the features of J stream through shared
// memory in tiles (positions as float2 and double2, descriptors as folded 64-byte rows).  The
// gate is evaluated in fp32 against the normalised line with a 0.05 px margin (4 instructions per
// (i, j)); the ~1 % that pass are collected in a 32-bit mask per 32 rows and only those are
// re-tested in fp64 with exactly the expression tree of the CPU restatement (explicit fma order,
// IEEE division), so the strict comparison decides identically, and their Hamming distance is
// taken.  Best-2 as packed (distance << 22 | j) keys: lowest j wins ties like the strict < of
// distanceRatio::update.  Survivors leave through the order-preserving compaction of match_post.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "context.cuh"
#include "guided.cuh"
#include "match_post.cuh"

namespace hulo {
namespace {

constexpr int kGuidedThreads = 256;
constexpr int kGuidedTile = 128;          // rows of J per shared-memory tile

struct alignas(16) GuidedItem {
    uint32_t pair;       // position in the batch (F, threshold)
    uint32_t i_row0;     // first row of the tile (row in the table of the I side)
    uint32_t i_rows;     // rows in the tile (<= kGuidedThreads)
    uint32_t j_row0;     // first row of image J (row in the table of the J side)
    uint32_t j_rows;
    uint32_t pad;
    uint64_t out_slot0;  // val[out_slot0 + t] for row i_row0 + t
};

__global__ void __launch_bounds__(kGuidedThreads) guided_kernel(
    const uint4 *__restrict__ rows, const double2 *__restrict__ xy, const uint4 *__restrict__ rows_j,
    const double2 *__restrict__ xy_j, const double *__restrict__ Fs,
    const double *__restrict__ thr2, double ratio, const GuidedItem *__restrict__ items, uint32_t n_items,
    int32_t *__restrict__ val) {
    __shared__ float2 s_xyf[kGuidedTile];
    __shared__ double2 s_xyd[kGuidedTile];
    __shared__ __align__(16) uint4 s_desc[kGuidedTile * 4];
    const int tid = threadIdx.x;
    for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
        const GuidedItem item = items[it];
        const double *F = Fs + 9 * (size_t)item.pair;
        const double th = thr2[item.pair];
        const bool live = (uint32_t)tid < item.i_rows;
        uint32_t q[16];
        double l0 = 0, l1 = 0, l2 = 0, den = 1;
        float l0f = 0, l1f = 0, l2f = 0, gate = -1.0f;
        if (live) {
            const uint4 *src = rows + (size_t)(item.i_row0 + tid) * 4;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const uint4 t = __ldg(src + v);
                q[4 * v] = t.x; q[4 * v + 1] = t.y; q[4 * v + 2] = t.z; q[4 * v + 3] = t.w;
            }
            const double2 p = xy[item.i_row0 + tid];
            // epipolar line F x_i, same fma tree as the CPU restatement
            l0 = fma(F[0], p.x, fma(F[1], p.y, F[2]));
            l1 = fma(F[3], p.x, fma(F[4], p.y, F[5]));
            l2 = fma(F[6], p.x, fma(F[7], p.y, F[8]));
            den = fma(l0, l0, l1 * l1);
            const double inv = rsqrt(den);
            if (isfinite(inv) && th >= 0.0) {           // a degenerate line (den = 0 or NaN) matches nothing
                l0f = (float)(l0 * inv); l1f = (float)(l1 * inv); l2f = (float)(l2 * inv);
                gate = (float)(sqrt(th) + 0.05);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) q[k] = 0;
        }
        uint32_t best0 = 0xFFFFFFFFu, best1 = 0xFFFFFFFFu;
        for (uint32_t j0 = 0; j0 < item.j_rows; j0 += kGuidedTile) {
            const uint32_t nj = min((uint32_t)kGuidedTile, item.j_rows - j0);
            __syncthreads();                             // the previous tile has been consumed
            for (uint32_t e = tid; e < nj; e += kGuidedThreads) {
                const double2 p = xy_j[item.j_row0 + j0 + e];
                s_xyd[e] = p;
                s_xyf[e] = make_float2((float)p.x, (float)p.y);
            }
            // pad the last group of 32 with NaN positions (they fail every gate), so the gate loop
            // below always runs 32 rows with compile-time bit positions
            for (uint32_t e = nj + tid; e < ((nj + 31u) & ~31u); e += kGuidedThreads) s_xyf[e] = make_float2(NAN, NAN);
            for (uint32_t e = tid; e < nj * 4; e += kGuidedThreads)
                s_desc[e] = __ldg(rows_j + (size_t)(item.j_row0 + j0) * 4 + e);
            __syncthreads();
            for (uint32_t g0 = 0; g0 < nj; g0 += 32) {
                uint32_t mask = 0;
#pragma unroll
                for (uint32_t jj = 0; jj < 32; ++jj) {
                    const float2 p = s_xyf[g0 + jj];
                    const float t = fmaf(l0f, p.x, fmaf(l1f, p.y, l2f));
                    if (fabsf(t) < gate) mask |= 1u << jj;
                }
                while (mask) {
                    const uint32_t jj = (uint32_t)__ffs(mask) - 1u;
                    mask &= mask - 1u;
                    const uint32_t jl = g0 + jj;
                    const double2 p = s_xyd[jl];
                    const double d = fma(l0, p.x, fma(l1, p.y, l2));
                    const double err = d * d / den;
                    if (!(err < th)) continue;
                    uint32_t x[16];
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const uint4 w = s_desc[jl * 4 + v];
                        x[4 * v] = q[4 * v] ^ w.x; x[4 * v + 1] = q[4 * v + 1] ^ w.y;
                        x[4 * v + 2] = q[4 * v + 2] ^ w.z; x[4 * v + 3] = q[4 * v + 3] ^ w.w;
                    }
                    // the rows are stored folded (knn2.cuh): undo the fold of the difference
                    x[15] ^= x[11] ^ x[14];
                    int h = 0;
#pragma unroll
                    for (int i = 0; i < 5; ++i) x[3 * i + 2] ^= x[3 * i] ^ x[3 * i + 1];
#pragma unroll
                    for (int k = 0; k < 16; ++k) h += __popc(x[k]);
                    const uint32_t key = ((uint32_t)h << kKeyIdxBits) | (j0 + jl);
                    const uint32_t hi = max(best0, key);
                    best0 = min(best0, key);
                    best1 = min(best1, hi);
                }
            }
        }
        if (live) {
            int32_t out = -1;
            if (best1 != 0xFFFFFFFFu) {
                // distanceRatio::isValid: a second candidate exists and best < ratio * second (doubles)
                const double bd = (double)(best0 >> kKeyIdxBits), sbd = (double)(best1 >> kKeyIdxBits);
                if (bd < ratio * sbd) out = (int32_t)(best0 & kKeyIdxMask);
            }
            val[item.out_slot0 + tid] = out;
        }
    }
}

// Position groups of one image: rep[f] = the first feature of the image at the same float
// position as feature f, member[f] = 1 when that position is shared by several features.  A
// duplicate position 4-tuple needs two matches whose I-features share a position, so images without
// shared positions (the normal case) skip the duplicate filter altogether.  Open addressing over the
// 64 bits of (float x, float y).
void position_groups(const double *xy, size_t n, int32_t *rep, uint8_t *member) {
    size_t cap = 16;
    while (cap < 2 * n + 2) cap <<= 1;
    std::vector<int32_t> slot(cap, -1);
    std::vector<uint64_t> keys(n);
    for (size_t f = 0; f < n; ++f) {
        const float x = (float)xy[2 * f], y = (float)xy[2 * f + 1];
        uint32_t a, b;
        memcpy(&a, &x, 4); memcpy(&b, &y, 4);
        if (x == 0.0f) a = 0;                       // +0 and -0 compare equal as floats
        if (y == 0.0f) b = 0;
        const uint64_t k = ((uint64_t)a << 32) | b;
        keys[f] = k;
        size_t h = (size_t)((k * 0x9E3779B97F4A7C15ULL) >> 20) & (cap - 1);
        while (slot[h] >= 0 && keys[(size_t)slot[h]] != k) h = (h + 1) & (cap - 1);
        if (slot[h] < 0 || x != x || y != y) {      // NaN never equals anything
            if (slot[h] < 0) slot[h] = (int32_t)f;
            rep[f] = (int32_t)f;
            member[f] = 0;
        } else {
            rep[f] = slot[h];
            member[f] = 1;
            member[(size_t)slot[h]] = 1;
        }
    }
}

}  // namespace

// Position groups of the images of one side, built the first time an image appears in a pair.
const GuidedGroups::Seg &GuidedGroups::of(const GuidedSide &side, uint32_t S) {
    auto it = segs.find(S);
    if (it != segs.end()) return it->second;
    Seg &g = segs[S];
    const size_t a = (size_t)side.seg[S], n = (size_t)(side.seg[S + 1] - side.seg[S]);
    g.rep.resize(std::max<size_t>(n, 1));
    g.member.assign(std::max<size_t>(n, 1), 0);
    position_groups(side.h_xy + 2 * a, n, g.rep.data(), g.member.data());
    return g;
}

int guided_match_sides(hulo_gpu *h, const GuidedSide &SI, const GuidedSide &SJ, const uint32_t *pairs, size_t n_pairs,
                       const double *F, const double *error_th, double dist_ratio, int dedup, uint64_t *pair_offsets,
                       uint32_t *out_i, uint32_t *out_j, size_t cap, size_t *n_out) {
    const uint64_t max_batch_rows = 16u << 20;
    size_t total_out = 0;
    bool overflow = false;
    std::vector<GuidedItem> items;
    std::vector<uint64_t> row_off, h_seg_out;
    std::vector<uint32_t> hi, hj;
    GuidedGroups own_i, own_j;
    GuidedGroups &groups_i = SI.groups ? *SI.groups : own_i, &groups_j = SJ.groups ? *SJ.groups : own_j;
    std::vector<uint64_t> seen;
    size_t p0 = 0;
    while (p0 < n_pairs) {
        items.clear();
        row_off.assign(1, 0);
        size_t p1 = p0;
        while (p1 < n_pairs) {
            const uint32_t I = pairs[2 * p1], J = pairs[2 * p1 + 1];
            const uint64_t nI = SI.seg[I + 1] - SI.seg[I], nJ = SJ.seg[J + 1] - SJ.seg[J];
            if (p1 > p0 && row_off.back() + nI > max_batch_rows) break;
            for (uint64_t t0 = 0; t0 < nI; t0 += kGuidedThreads) {
                GuidedItem it{};
                it.pair = (uint32_t)(p1 - p0);
                it.i_row0 = (uint32_t)(SI.seg[I] + t0);
                it.i_rows = (uint32_t)std::min<uint64_t>(kGuidedThreads, nI - t0);
                it.j_row0 = (uint32_t)SJ.seg[J];
                it.j_rows = (uint32_t)nJ;
                it.out_slot0 = row_off.back() + t0;
                items.push_back(it);
            }
            row_off.push_back(row_off.back() + nI);
            ++p1;
        }
        const size_t bp = p1 - p0;
        const uint64_t n_rows = row_off.back();
        HULO_ARG(n_rows < (uint64_t)0x7fffffff, "batch too large");
        if (n_rows > 0) {
            const size_t n_blocks = (n_rows + kCompactBlockRows - 1) / kCompactBlockRows;
            // items: item list ; scratch0: val ; scratch1: row_off | F | thresholds ; scratch2: total,
            // seg_out_off, block counts ; scratch3: out_i | out_j
            HULO_CUDA(h->items.reserve(items.size() * sizeof(GuidedItem)));
            HULO_CUDA(h->scratch0.reserve(n_rows * sizeof(int32_t)));
            HULO_CUDA(h->scratch1.reserve((bp + 1) * sizeof(uint64_t) + bp * 10 * sizeof(double)));
            HULO_CUDA(h->scratch2.reserve((n_blocks + 2) * sizeof(uint32_t) + (bp + 2) * sizeof(uint64_t) + 64));
            HULO_CUDA(h->scratch3.reserve(n_rows * 2 * sizeof(uint32_t)));
            uint64_t *d_row_off = h->scratch1.as<uint64_t>();
            double *d_F = reinterpret_cast<double *>(d_row_off + (bp + 1));
            double *d_thr = d_F + 9 * bp;
            HULO_CUDA(cudaMemcpyAsync(h->items.ptr, items.data(), items.size() * sizeof(GuidedItem), cudaMemcpyHostToDevice, h->stream));
            HULO_CUDA(cudaMemcpyAsync(d_row_off, row_off.data(), (bp + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
            HULO_CUDA(cudaMemcpyAsync(d_F, F + 9 * p0, bp * 9 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            HULO_CUDA(cudaMemcpyAsync(d_thr, error_th + p0, bp * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            int32_t *val = h->scratch0.as<int32_t>();
            const unsigned grid = (unsigned)std::min<size_t>(items.size(), (size_t)h->sm_count * 8);
            guided_kernel<<<grid, kGuidedThreads, 0, h->stream>>>(SI.rows, SI.d_xy, SJ.rows, SJ.d_xy, d_F, d_thr, dist_ratio,
                                                                  h->items.as<GuidedItem>(), (uint32_t)items.size(), val);
            HULO_CUDA(cudaGetLastError());
            h->launches++;
            uint8_t *s2 = h->scratch2.as<uint8_t>();
            uint64_t *d_total = reinterpret_cast<uint64_t *>(s2);
            uint64_t *d_seg_out = d_total + 1;
            uint32_t *d_blocks = reinterpret_cast<uint32_t *>(d_seg_out + (bp + 1));
            uint32_t *o_i = h->scratch3.as<uint32_t>(), *o_j = o_i + n_rows;
            { int nl = 0; HULO_CUDA(compact_launch(val, nullptr, (uint32_t)n_rows, d_row_off, (uint32_t)bp, d_blocks, nullptr, o_i, o_j,
                                     nullptr, d_seg_out, d_total, h->stream, &nl)); h->launches += nl; }
            h_seg_out.resize(bp + 2);
            HULO_CUDA(cudaMemcpyAsync(h_seg_out.data(), d_total, (bp + 2) * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
            HULO_CUDA(cudaStreamSynchronize(h->stream));
            const uint64_t total = h_seg_out[0];
            hi.resize(std::max<uint64_t>(total, 1));
            hj.resize(std::max<uint64_t>(total, 1));
            if (total) {
                HULO_CUDA(cudaMemcpyAsync(hi.data(), o_i, total * 4, cudaMemcpyDeviceToHost, h->stream));
                HULO_CUDA(cudaMemcpyAsync(hj.data(), o_j, total * 4, cudaMemcpyDeviceToHost, h->stream));
                HULO_CUDA(cudaStreamSynchronize(h->stream));
            }
            // per pair: drop matches whose position 4-tuple (as floats) repeats
            // (IndMatchDecorator::getDeduplicated), first occurrence kept, order preserved
            for (size_t p = 0; p < bp; ++p) {
                const uint32_t I = pairs[2 * (p0 + p)], J = pairs[2 * (p0 + p) + 1];
                const GuidedGroups::Seg *gI = dedup ? &groups_i.of(SI, I) : nullptr;
                const GuidedGroups::Seg *gJ = dedup ? &groups_j.of(SJ, J) : nullptr;
                seen.clear();
                for (uint64_t m = h_seg_out[1 + p]; m < h_seg_out[2 + p]; ++m) {
                    if (dedup && gI->member[hi[m]]) {
                        const uint64_t k = ((uint64_t)(uint32_t)gI->rep[hi[m]] << 32) | (uint32_t)gJ->rep[hj[m]];
                        if (std::find(seen.begin(), seen.end(), k) != seen.end()) continue;
                        seen.push_back(k);
                    }
                    if (total_out < cap) {
                        if (out_i == nullptr || out_j == nullptr) { set_error("guided matching: null output"); return HULO_ERR_ARG; }
                        out_i[total_out] = hi[m];
                        out_j[total_out] = hj[m];
                    } else {
                        overflow = true;
                    }
                    ++total_out;
                }
                if (pair_offsets) pair_offsets[p0 + p + 1] = total_out;
            }
        } else if (pair_offsets) {
            for (size_t p = 0; p < bp; ++p) pair_offsets[p0 + p + 1] = total_out;
        }
        p0 = p1;
    }
    *n_out = total_out;
    if (overflow) {
        set_error("guided matching: %zu matches, capacity %zu", total_out, cap);
        return HULO_ERR_CAPACITY;
    }
    return HULO_OK;
}

}  // namespace hulo

using namespace hulo;

extern "C" {

int hulo_guided_match(hulo_gpu *h, const hulo_db *db, const double *xy, const uint32_t *pairs, size_t n_pairs,
                      const double *F, const double *error_th, double dist_ratio, int dedup, uint64_t *pair_offsets,
                      uint32_t *out_i, uint32_t *out_j, size_t cap, size_t *n_out) {
    HULO_ARG(h != nullptr && db != nullptr && n_out != nullptr, "null argument");
    *n_out = 0;
    if (pair_offsets) pair_offsets[0] = 0;
    if (n_pairs == 0) return HULO_OK;
    HULO_ARG(pairs != nullptr && F != nullptr && error_th != nullptr, "null argument");
    HULO_ARG(db->n == 0 || xy != nullptr, "feature positions are null");
    HULO_ARG(dist_ratio > 0.0, "the distance ratio must be positive");
    HULO_CUDA(cudaSetDevice(h->device));
    const size_t n_seg = db->seg.size() - 1;
    for (size_t p = 0; p < n_pairs; ++p) {
        HULO_ARG(pairs[2 * p] < n_seg && pairs[2 * p + 1] < n_seg, "pair refers to a segment that does not exist");
        HULO_ARG(db->seg[pairs[2 * p + 1] + 1] - db->seg[pairs[2 * p + 1]] <= kMaxChunkRows, "image with more than 4 Mi descriptors");
    }
    // feature positions of every row of the table
    HULO_CUDA(h->stageA.reserve(std::max<size_t>(db->n, 1) * sizeof(double2)));
    if (db->n) HULO_CUDA(cudaMemcpyAsync(h->stageA.ptr, xy, db->n * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
    GuidedSide side;
    side.rows = db->rows;
    side.seg = db->seg.data();
    side.n_seg = n_seg;
    side.h_xy = xy;
    side.d_xy = h->stageA.as<double2>();
    GuidedGroups groups;                  // one table on both sides: one set of position groups
    side.groups = &groups;
    return guided_match_sides(h, side, side, pairs, n_pairs, F, error_th, dist_ratio, dedup, pair_offsets, out_i, out_j,
                              cap, n_out);
}

}  // extern "C"
