#include "match_utils_gpu.h"

#include <algorithm>
#include <iostream>
#include <stdexcept>

#include "../../../include/hulo_gpu.h"
#include "desc_files.h"

namespace hulo {

namespace {
void must(int status, const char *what) {
    if (status != HULO_OK) throw std::runtime_error(std::string(what) + ": " + hulo_last_error());
}
}  // namespace

struct BowTable {
    hulo_bow *index = nullptr;
    std::map<std::size_t, uint32_t> row_of_view;
    std::size_t dim = 0;
};

struct GpuSession::State {
    hulo_gpu *gpu = nullptr;
    std::map<std::string, std::shared_ptr<Table>> cache;
    std::map<std::string, std::shared_ptr<BowTable>> bow_cache;
    ~State() {
        for (auto &kv : bow_cache)
            if (kv.second && kv.second->index) hulo_bow_destroy(kv.second->index);
        for (auto &kv : cache)
            if (kv.second && kv.second->db) hulo_db_free(kv.second->db);
        if (gpu) hulo_gpu_destroy(gpu);
    }
};

GpuSession::GpuSession(int device) : st_(std::make_shared<State>()) {
    must(hulo_gpu_create(device, &st_->gpu), "hulo_gpu_create");
}
hulo_gpu *GpuSession::gpu() const { return st_->gpu; }
std::map<std::string, std::shared_ptr<BowTable>> &GpuSession::bowCache() { return st_->bow_cache; }

void GpuSession::clearCache() {
    for (auto &kv : st_->cache)
        if (kv.second && kv.second->db) hulo_db_free(kv.second->db);
    st_->cache.clear();
    for (auto &kv : st_->bow_cache)
        if (kv.second && kv.second->index) hulo_bow_destroy(kv.second->index);
    st_->bow_cache.clear();
}

// ------------------------------------------------------------------ selectViewByBoF
void selectViewByBoF(GpuSession &s, const std::vector<float> &bow, const std::string &matchDir,
                     const std::set<std::size_t> &viewList, const Views &views, int knn,
                     std::set<std::size_t> &selectedViewList) {
    if (!(knn >= 0 && (std::size_t)knn < viewList.size()))
        throw std::invalid_argument("hulo::selectViewByBoF: knn must be smaller than the number of views");   // :30
    std::shared_ptr<BowTable> t;
    auto &cache = s.bowCache();
    auto it = cache.find(matchDir);
    if (it != cache.end()) {
        t = it->second;
    } else {
        t = std::make_shared<BowTable>();
        std::vector<float> all;
        std::vector<double> vec;
        for (const auto &kv : views) {
            int r = 0, c = 0;
            if (!readMatBin(bowPath(matchDir, kv.second.s_Img_path), r, c, vec) || vec.empty()) continue;
            if (t->dim == 0) t->dim = vec.size();
            if (vec.size() != t->dim) throw std::runtime_error("hulo::selectViewByBoF: .bow vectors of different length");
            t->row_of_view[kv.first] = (uint32_t)t->row_of_view.size();
            for (double v : vec) all.push_back((float)v);                 // convertTo(CV_32FC1), :44-46
        }
        if (t->row_of_view.empty()) throw std::runtime_error("hulo::selectViewByBoF: no .bow file in " + matchDir);
        must(hulo_bow_create(s.gpu(), all.data(), t->row_of_view.size(), t->dim, &t->index), "hulo_bow_create");
        cache[matchDir] = t;
    }
    if (bow.size() != t->dim) throw std::invalid_argument("hulo::selectViewByBoF: query vector of the wrong length");
    std::vector<uint32_t> subset;
    std::vector<std::size_t> view_of;                                      // position in viewList order -> view id
    for (std::size_t v : viewList) {
        subset.push_back(t->row_of_view.at(v));                            // .at(): every listed view must have a vector
        view_of.push_back(v);
    }
    std::map<uint32_t, std::size_t> view_of_row;
    for (std::size_t k = 0; k < subset.size(); ++k) view_of_row[subset[k]] = view_of[k];
    std::vector<int32_t> idx((std::size_t)std::max(knn, 1));
    must(hulo_bow_knn(t->index, bow.data(), subset.data(), subset.size(), (std::size_t)knn, idx.data(), nullptr),
         "hulo_bow_knn");
    selectedViewList.clear();
    for (int k = 0; k < knn; ++k) selectedViewList.insert(view_of_row.at((uint32_t)idx[k]));
}

std::shared_ptr<GpuSession::Table> GpuSession::table(const Views &views, const std::string &sMatchesDir,
                                                     const std::vector<std::size_t> &view_ids) {
    std::string key = sMatchesDir + "|";
    for (std::size_t v : view_ids) key += std::to_string(v) + ",";
    auto it = st_->cache.find(key);
    if (it != st_->cache.end()) return it->second;
    auto t = std::make_shared<Table>();
    std::vector<uint8_t> all, rows;
    std::vector<uint64_t> off(1, 0);
    for (std::size_t v : view_ids) {
        std::size_t n = 0;
        // an unreadable file behaves as an image without descriptors
        readAKAZEBin(descPath(sMatchesDir, views.at(v).s_Img_path, true), rows, n);
        all.insert(all.end(), rows.begin(), rows.end());
        off.push_back(off.back() + n);
        t->seg_of_view[v] = (uint32_t)t->view_ids.size();
        t->rows_of_view[v] = n;
        t->view_ids.push_back(v);
    }
    static const uint8_t none = 0;
    must(hulo_db_upload(st_->gpu, all.empty() ? &none : all.data(), (std::size_t)off.back(), HULO_ROW_BYTES,
                        off.data(), view_ids.size(), &t->db),
         "hulo_db_upload");
    st_->cache[key] = t;
    return t;
}

GpuSession &defaultSession() {
    static GpuSession s(0);
    return s;
}

// ------------------------------------------------------------------ matchAKAZE
static void match_pair_list(GpuSession &s, const Views &views, const std::string &sMatchesDir,
                            const std::vector<Pair> &pairs, float fDistRatio, PairWiseMatches &matches,
                            std::map<std::size_t, std::size_t> *rows_of_view) {
    std::vector<std::size_t> ids;
    for (const Pair &p : pairs) { ids.push_back(p.first); ids.push_back(p.second); }
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    auto t = s.table(views, sMatchesDir, ids);
    if (rows_of_view) *rows_of_view = t->rows_of_view;
    std::vector<uint32_t> seg_pairs(2 * pairs.size());
    for (std::size_t k = 0; k < pairs.size(); ++k) {
        seg_pairs[2 * k] = t->seg_of_view.at(pairs[k].first);
        seg_pairs[2 * k + 1] = t->seg_of_view.at(pairs[k].second);
    }
    std::vector<uint64_t> off(pairs.size() + 1, 0);
    std::vector<uint32_t> oi(1), oj(1);
    std::size_t n = 0;
    int rc = hulo_match_pairs(s.gpu(), t->db, seg_pairs.data(), pairs.size(), fDistRatio, HULO_PAIR_REFERENCE,
                              off.data(), oi.data(), oj.data(), 0, &n);
    if (rc == HULO_ERR_CAPACITY) {
        oi.resize(n); oj.resize(n);
        rc = hulo_match_pairs(s.gpu(), t->db, seg_pairs.data(), pairs.size(), fDistRatio, HULO_PAIR_REFERENCE,
                              off.data(), oi.data(), oj.data(), n, &n);
    }
    must(rc, "hulo_match_pairs");
    for (std::size_t k = 0; k < pairs.size(); ++k) {
        // a pair without a surviving match never gets a key (insertion happens on push_back,
        // MatchUtils.cpp:148); a pair listed twice appends twice, as the reference does
        for (uint64_t m = off[k]; m < off[k + 1]; ++m) matches[pairs[k]].push_back(IndMatch(oi[m], oj[m]));
    }
}

void matchAKAZE(GpuSession &s, const Views &views, const std::string &sMatchesDir, const std::vector<Pair> &pairs,
                const float fDistRatio, PairWiseMatches &matches) {
    std::cout << "Start putative matching" << std::endl;
    match_pair_list(s, views, sMatchesDir, pairs, fDistRatio, matches, nullptr);
}

// ------------------------------------------------------------------ trackAKAZE
void propagateTracks(std::size_t n_frames, std::size_t maxFrameDist, const std::vector<int> &feat_number,
                     PairWiseMatches &matches) {
    if (n_frames < 2) return;
    std::vector<std::vector<int>> trackPointer(n_frames - 1);
    for (std::size_t f = 0; f + 1 < n_frames; ++f) {
        trackPointer[f].assign((std::size_t)feat_number[f], -1);
        // operator[] on purpose: the reference creates the (f, f+1) key here when it is absent
        for (const IndMatch &m : matches[Pair(f, f + 1)]) trackPointer[f][m.i_] = (int)m.j_;
    }
    for (std::size_t f = 0; f + 1 < n_frames; ++f) {
        const std::size_t lim = std::min(f + maxFrameDist, n_frames);
        for (std::size_t to = f + 2; to < lim; ++to) {
            for (std::size_t i = 0; i < trackPointer[f].size(); ++i) {
                const int t = trackPointer[f][i];
                if (t == -1) continue;
                const int nx = trackPointer[to - 1][(std::size_t)t];
                trackPointer[f][i] = nx;
                if (nx != -1) matches[Pair(f, to)].push_back(IndMatch((uint32_t)i, (uint32_t)nx));
            }
        }
    }
}

void trackAKAZE(GpuSession &s, const Views &views, const std::string &sMatchesDir, const std::size_t maxFrameDist,
                const float fDistRatio, PairWiseMatches &matches) {
    std::cout << "Start putative matching" << std::endl;
    if (views.size() < 2) return;
    // consecutive frames: (id, id + 1) for every view but the last in map order (:164-237);
    // like the reference this assumes consecutive view ids starting at 0
    std::vector<Pair> pairs;
    std::size_t k = 0;
    for (auto it = views.begin(); k + 1 < views.size(); ++it, ++k) pairs.push_back(Pair(it->first, it->first + 1));
    std::map<std::size_t, std::size_t> rows_of_view;
    match_pair_list(s, views, sMatchesDir, pairs, fDistRatio, matches, &rows_of_view);
    std::vector<int> feat_number(views.size() - 1, 0);
    for (const Pair &p : pairs)
        if (p.first < feat_number.size()) feat_number[p.first] = (int)rows_of_view[p.first];   // :183
    propagateTracks(views.size(), maxFrameDist, feat_number, matches);
}

// ------------------------------------------------------------------ matchAKAZEToQuery
void matchAKAZEToQuery(GpuSession &s, const Views &views, const std::string &sMatchesDir,
                       const std::string &sQueryMatchesDir, const std::vector<std::size_t> &pairs,
                       const std::size_t queryInd, const float fDistRatio, PairWiseMatches &matches,
                       FeatDistMap &featDist) {
    std::cout << "Start putative matching" << std::endl;
    std::vector<uint8_t> q;
    std::size_t nq = 0;
    // the query descriptor file is named by s_Img_path itself, not its basename part (:295)
    readAKAZEBin(descPath(sQueryMatchesDir, views.at(queryInd).s_Img_path, false), q, nq);
    if (nq < 1) return;                                                            // :299-301
    for (std::size_t v : pairs) matches.insert(std::make_pair(Pair(v, queryInd), IndMatches()));   // :314-319

    // the resident table holds every map view (all but the query); `pairs` selects segments
    std::vector<std::size_t> ids;
    for (const auto &kv : views)
        if (kv.first != queryInd) ids.push_back(kv.first);
    auto t = s.table(views, sMatchesDir, ids);
    std::vector<uint32_t> segs(pairs.size());
    std::size_t total_rows = 0;
    for (std::size_t k = 0; k < pairs.size(); ++k) {
        segs[k] = t->seg_of_view.at(pairs[k]);
        total_rows += t->rows_of_view.at(pairs[k]);
    }
    const std::size_t cap = std::max<std::size_t>(total_rows, 1);
    std::vector<uint32_t> ov(cap), oi(cap), oj(cap), counts(std::max<std::size_t>(pairs.size(), 1));
    std::vector<int32_t> od(cap);
    std::size_t n = 0;
    must(hulo_match_to_query(s.gpu(), t->db, segs.data(), pairs.size(), q.data(), nq, HULO_ROW_BYTES, fDistRatio,
                             ov.data(), oi.data(), oj.data(), od.data(), cap, &n, counts.data()),
         "hulo_match_to_query");
    std::size_t m = 0;
    for (std::size_t k = 0; k < pairs.size(); ++k) {
        const Pair key(pairs[k], queryInd);
        IndMatches ind;
        std::map<std::size_t, int> fd;
        for (uint32_t c = 0; c < counts[k]; ++c, ++m) {
            ind.push_back(IndMatch(oi[m], oj[m]));
            fd[oj[m]] = od[m];                                                     // last i wins, :351
        }
        matches[key] = ind;                                                        // :358
        featDist[key] = fd;                                                        // :359
    }
}

// ------------------------------------------------------------------ geometricMatch
uint64_t g_geometricSeed = 0x5eedULL;

void geometricMatch(GpuSession &s, const Views &views, const RegionsProvider &regions_provider,
                    const PairWiseMatches &map_putativeMatches, PairWiseMatches &map_geometricMatches,
                    int ransacRound, double geomPrec, bool bGuided_matching) {
    if (bGuided_matching)
        throw std::invalid_argument("hulo::geometricMatch: guided matching needs the descriptor directory (sMatchesDir overload)");
    geometricMatch(s, views, regions_provider, std::string(), map_putativeMatches, map_geometricMatches, ransacRound,
                   geomPrec, false);
}

void geometricMatch(GpuSession &s, const Views &views, const RegionsProvider &regions_provider,
                    const std::string &sMatchesDir, const PairWiseMatches &map_putativeMatches,
                    PairWiseMatches &map_geometricMatches, int ransacRound, double geomPrec, bool bGuided_matching) {
    map_geometricMatches.clear();                                                  // assignment at :415
    std::vector<double> xI, xJ;
    std::vector<uint64_t> off(1, 0), seeds;
    std::vector<int32_t> sizes;
    std::vector<const PairWiseMatches::value_type *> kept;
    for (const auto &kv : map_putativeMatches) {
        const View &vi = views.at(kv.first.first), &vj = views.at(kv.first.second);   // .at() like :385-403
        const FeatureLocations &fi = regions_provider.at(kv.first.first), &fj = regions_provider.at(kv.first.second);
        for (const IndMatch &m : kv.second) {
            xI.push_back(fi.at(m.i_).first); xI.push_back(fi.at(m.i_).second);
            xJ.push_back(fj.at(m.j_).first); xJ.push_back(fj.at(m.j_).second);
        }
        off.push_back(off.back() + kv.second.size());
        sizes.push_back((int32_t)vi.ui_width); sizes.push_back((int32_t)vi.ui_height);
        sizes.push_back((int32_t)vj.ui_width); sizes.push_back((int32_t)vj.ui_height);
        seeds.push_back(g_geometricSeed + 1000003ull * ((uint64_t)kv.first.first * 1000003ull + (uint64_t)kv.first.second));
        kept.push_back(&kv);
    }
    const std::size_t P = kept.size();
    if (P) {
        std::vector<int32_t> valid(P), inl(std::max<std::size_t>((std::size_t)off.back(), 1));
        std::vector<uint32_t> ninl(P);
        std::vector<double> F(9 * P), err(P);
        must(hulo_geometric_filter(s.gpu(), xI.data(), xJ.data(), off.data(), P, sizes.data(), geomPrec,
                                   (std::size_t)std::max(ransacRound, 0), g_geometricSeed, seeds.data(), valid.data(),
                                   ninl.data(), inl.data(), F.data(), err.data(), nullptr),
             "hulo_geometric_filter");
        for (std::size_t p = 0; p < P; ++p) {
            if (!valid[p]) continue;
            IndMatches &out = map_geometricMatches[kept[p]->first];
            for (uint32_t c = 0; c < ninl[p]; ++c) out.push_back(kept[p]->second[(std::size_t)inl[off[p] + c]]);
        }
        if (bGuided_matching && !map_geometricMatches.empty()) {
            // Geometry_guided_matching over all features of the surviving pairs
            std::vector<std::size_t> ids;
            for (const auto &kv : map_geometricMatches) { ids.push_back(kv.first.first); ids.push_back(kv.first.second); }
            std::sort(ids.begin(), ids.end());
            ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
            auto t = s.table(views, sMatchesDir, ids);
            std::vector<double> xy;
            for (std::size_t v : t->view_ids) {
                const FeatureLocations &f = regions_provider.at(v);
                if (f.size() != t->rows_of_view.at(v))
                    throw std::runtime_error("hulo::geometricMatch: view " + std::to_string(v) + " has " + std::to_string(f.size()) +
                                             " features but " + std::to_string(t->rows_of_view.at(v)) + " descriptors");
                for (const auto &pt : f) { xy.push_back(pt.first); xy.push_back(pt.second); }
            }
            std::vector<uint32_t> seg_pairs;
            std::vector<double> gF, gthr;
            std::vector<Pair> gkeys;
            for (std::size_t p = 0; p < P; ++p) {
                if (!valid[p]) continue;
                seg_pairs.push_back(t->seg_of_view.at(kept[p]->first.first));
                seg_pairs.push_back(t->seg_of_view.at(kept[p]->first.second));
                gF.insert(gF.end(), F.begin() + 9 * p, F.begin() + 9 * p + 9);
                gthr.push_back(err[p] * err[p]);                                   // Square(m_dPrecision_robust)
                gkeys.push_back(kept[p]->first);
            }
            const std::size_t G = gkeys.size();
            std::vector<uint64_t> goff(G + 1, 0);
            std::vector<uint32_t> gi(1), gj(1);
            std::size_t n = 0;
            static const double none2[2] = {0, 0};
            int rc = hulo_guided_match(s.gpu(), t->db, xy.empty() ? none2 : xy.data(), seg_pairs.data(), G, gF.data(),
                                       gthr.data(), 0.6 * 0.6, 1, goff.data(), gi.data(), gj.data(), 0, &n);
            if (rc == HULO_ERR_CAPACITY) {
                gi.resize(n); gj.resize(n);
                rc = hulo_guided_match(s.gpu(), t->db, xy.data(), seg_pairs.data(), G, gF.data(), gthr.data(), 0.6 * 0.6, 1,
                                       goff.data(), gi.data(), gj.data(), n, &n);
            }
            must(rc, "hulo_guided_match");
            for (std::size_t g = 0; g < G; ++g) {
                IndMatches &out = map_geometricMatches[gkeys[g]];                  // the key stays even when empty
                out.clear();
                for (uint64_t m = goff[g]; m < goff[g + 1]; ++m) out.push_back(IndMatch(gi[m], gj[m]));
            }
        }
    }
    std::cout << "number of putative matches : " << map_putativeMatches.size() << std::endl;
    std::cout << "number of geometric matches : " << map_geometricMatches.size() << std::endl;
}

// ------------------------------------------------------------------ default-session overloads
void geometricMatch(const Views &views, const RegionsProvider &r, const PairWiseMatches &p, PairWiseMatches &g,
                    int ransacRound, double geomPrec, bool bGuided_matching) {
    geometricMatch(defaultSession(), views, r, p, g, ransacRound, geomPrec, bGuided_matching);
}
void matchAKAZE(const Views &views, const std::string &d, const std::vector<Pair> &pairs, const float r,
                PairWiseMatches &m) {
    matchAKAZE(defaultSession(), views, d, pairs, r, m);
}
void trackAKAZE(const Views &views, const std::string &d, const std::size_t maxFrameDist, const float r,
                PairWiseMatches &m) {
    trackAKAZE(defaultSession(), views, d, maxFrameDist, r, m);
}
void matchAKAZEToQuery(const Views &views, const std::string &d, const std::string &qd,
                       const std::vector<std::size_t> &pairs, const std::size_t queryInd, const float r,
                       PairWiseMatches &m, FeatDistMap &fd) {
    matchAKAZEToQuery(defaultSession(), views, d, qd, pairs, queryInd, r, m, fd);
}

}  // namespace hulo
