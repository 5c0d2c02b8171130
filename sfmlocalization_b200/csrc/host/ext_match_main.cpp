// hulo_ext_match -- the matching stages of the reference's ExtFeatAndMatch CLI
// (ExtFeatAndMatch/src/computeFeaturesAndMatches.cpp:150-246) on the GPU.  It reads the
// .desc files the extraction stage wrote into <matchdir>, matches, and writes
// <matchdir>/matches.putative.txt; then, like the reference (:194-246), drops the pairs with
// fewer than -mm matches, runs the F-matrix geometric filter (hulo::geometricMatch) on the
// rest using the .feat files and the image sizes of sfm_data.json, and writes
// <matchdir>/matches.f.txt -- the files the rest of the reference pipeline consumes
// (openMVG_main_GlobalSfM, cleanSfM.py).
//
//   hulo_ext_match <matchdir> [-f=0.6] [-v=0] [-p=<pairfile>] [-mf=0] [-r=4096] [-mm=60] [-g=4.0]
//                  [--views=<id path per line>] [--out=<file>] [--rank=R --world=W] [--device=D]
//                  [--putative-only]
//
// Flags -f/-v/-p/-mf/-r/-mm/-g have the reference's meaning (computeFeaturesAndMatches.cpp:49-64;
// -g is read as an integer there, :92, and here); the extraction flags (-c -t -o -l -sm) are
// accepted and ignored so the Python drivers' command lines keep working
// (reconstructGraph.py:158-163); -gm switches guided matching on like the reference.  When a .feat
// file or an image size is missing the geometric stage is skipped with a message (the reference
// would abort: "Cannot construct regions providers").
// The view list comes from <matchdir>/sfm_data.json like the reference, or from --views.
// --rank/--world shard the pair list across processes (one per GPU, no collective); each
// rank writes <out>.rank<R> (and matches.f.txt.rank<R>), rank order concatenation equals the
// single-GPU file.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "desc_files.h"
#include "match_utils_gpu.h"
#include "pair_lists.h"

using namespace hulo;

static bool flag(const char *arg, const char *name, std::string &val) {
    const size_t n = strlen(name);
    if (strncmp(arg, name, n) == 0 && arg[n] == '=') { val = arg + n + 1; return true; }
    return false;
}

int main(int argc, char **argv) {
    std::string sMatchesDir, sPairFile, sViews, sOut, v;
    float fDistRatio = 0.6f;
    int videoMatchFrame = 0, rank = 0, world = 1, device = -1;
    int ransacRound = 4096, minMatch = 60, geomError = 4;
    bool bGuided = false, putativeOnly = false;
    size_t maxFrameDist = 0;
    for (int a = 1; a < argc; ++a) {
        if (flag(argv[a], "-f", v)) fDistRatio = (float)atof(v.c_str());
        else if (flag(argv[a], "-v", v)) videoMatchFrame = atoi(v.c_str());
        else if (flag(argv[a], "-p", v)) sPairFile = v;
        else if (flag(argv[a], "-mf", v)) maxFrameDist = (size_t)atoll(v.c_str());
        else if (flag(argv[a], "-r", v)) ransacRound = atoi(v.c_str());
        else if (flag(argv[a], "-mm", v)) minMatch = atoi(v.c_str());
        else if (flag(argv[a], "-g", v)) geomError = atoi(v.c_str());
        else if (strcmp(argv[a], "-gm") == 0 || flag(argv[a], "-gm", v)) bGuided = true;
        else if (strcmp(argv[a], "--putative-only") == 0) putativeOnly = true;
        else if (flag(argv[a], "--views", v)) sViews = v;
        else if (flag(argv[a], "--out", v)) sOut = v;
        else if (flag(argv[a], "--rank", v)) rank = atoi(v.c_str());
        else if (flag(argv[a], "--world", v)) world = atoi(v.c_str());
        else if (flag(argv[a], "--device", v)) device = atoi(v.c_str());
        else if (argv[a][0] == '-') continue;   // other reference flags: not this stage's business
        else if (sMatchesDir.empty()) sMatchesDir = argv[a];
    }
    if (sMatchesDir.empty()) {
        std::cerr << "usage: hulo_ext_match <matchdir> [-f=] [-v=] [-p=] [-mf=] [--views=] [--out=] [--rank= --world=]\n";
        return 1;
    }
    // the reference asserts that at most one of pair file / video window / tracking is given
    const int modes = (!sPairFile.empty()) + (videoMatchFrame > 0) + (maxFrameDist > 0);
    if (modes > 1) { std::cerr << "specify at most one of -p, -v, -mf\n"; return 1; }
    std::cout << "Matches directory : " << sMatchesDir << std::endl;

    Views views;
    const bool ok = sViews.empty() ? readViewsFromSfmData(sMatchesDir + "/sfm_data.json", views)
                                   : readViewsFromList(sViews, views);
    if (!ok) { std::cerr << "Cannot load the view list" << std::endl; return EXIT_FAILURE; }
    if (sOut.empty()) sOut = sMatchesDir + "/matches.putative.txt";

    try {
        GpuSession session(device >= 0 ? device : rank);
        PairWiseMatches matches;
        std::cout << "Start Putative Matching..." << std::endl;
        if (maxFrameDist != 0) {
            if (world > 1) { std::cerr << "tracking mode is a chain: it does not shard\n"; return 1; }
            trackAKAZE(session, views, sMatchesDir, maxFrameDist, fDistRatio, matches);
        } else {
            std::vector<Pair> pairs;
            if (!sPairFile.empty()) {
                readPairFile(sPairFile, pairs);
            } else if (videoMatchFrame > 0) {
                generateVideoMatchPairs(views, pairs, videoMatchFrame);
                removeDupPairs(pairs);
            } else {
                generateAllPairs(views, pairs);
            }
            std::cout << "Total number of pairs : " << pairs.size() << std::endl;
            if (world > 1) {
                std::vector<size_t> ids;
                for (const auto &kv : views) ids.push_back(kv.first);
                auto t = session.table(views, sMatchesDir, ids);
                std::vector<Pair> mine;
                for (size_t k : partitionPairs(pairs, t->rows_of_view, rank, world)) mine.push_back(pairs[k]);
                pairs.swap(mine);
                sOut += ".rank" + std::to_string(rank);
            }
            matchAKAZE(session, views, sMatchesDir, pairs, fDistRatio, matches);
        }
        if (!exportPairWiseMatches(matches, sOut)) {
            std::cerr << "Cannot write " << sOut << std::endl;
            return EXIT_FAILURE;
        }
        std::cout << "wrote " << matches.size() << " pairs to " << sOut << std::endl;

        // ---- geometric matching (computeFeaturesAndMatches.cpp:194-246)
        if (!putativeOnly) {
            std::cout << "Start Geometric Matching..." << std::endl;
            for (auto it = matches.cbegin(); it != matches.cend();) {                 // :222-232
                if ((int)it->second.size() < minMatch) it = matches.erase(it);
                else ++it;
            }
            Views used;
            for (const auto &kv : matches) {
                used[kv.first.first] = views.at(kv.first.first);
                used[kv.first.second] = views.at(kv.first.second);
            }
            RegionsProvider regions;
            bool have = loadRegions(used, sMatchesDir, regions);
            for (const auto &kv : used) have = have && kv.second.ui_width > 0 && kv.second.ui_height > 0;
            if (!have) {
                std::cout << "geometric matching skipped: .feat files or image sizes are missing" << std::endl;
            } else {
                PairWiseMatches geometric;
                geometricMatch(session, views, regions, sMatchesDir, matches, geometric, ransacRound, (double)geomError,
                               bGuided);
                std::string sF = sMatchesDir + "/matches.f.txt";
                if (world > 1) sF += ".rank" + std::to_string(rank);
                if (!exportPairWiseMatches(geometric, sF)) {
                    std::cerr << "Cannot write " << sF << std::endl;
                    return EXIT_FAILURE;
                }
                std::cout << "wrote " << geometric.size() << " pairs to " << sF << std::endl;
            }
        }
    } catch (const std::exception &e) {
        std::cerr << "hulo_ext_match: " << e.what() << std::endl;
        return EXIT_FAILURE;
    }
    return 0;
}
