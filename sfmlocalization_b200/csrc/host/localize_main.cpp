// hulo_localize -- the reference's OpenMVGLocalization_AKAZE CLI
// (OpenMVGLocalization_AKAZE/src/localization.cpp:64-600) from the putative matching on, on the
// GPU: for every query it writes <outDir>/<basename>.json with the keys the reference's Python
// drivers read (filename, sfm_data, matches_dir and, when localised, K, R, t, pair:
// localization.cpp:100-144; consumers sfmMergeGraph.py:260-269, mergeSfM.py:50-66).
//
//   hulo_localize <query .desc file or folder> <sfmDir> <matchDir> <outDir>
//                 [-f=0.6] [-r=200] [-g=4.0] [-gm] [-x= -y= -z= -d=-1] [-i=1] [--width=W --height=H]
//                 [--device=D] [--seed=S] [--rank=R --world=W] [--amat=<A.yml>]
//
// The reference takes image files and extracts AKAZE features itself (localization.cpp:312-330);
// extraction is upstream of the accelerated path, so this tool takes the query's extracted regions:
// <name>.desc (FileUtils.cpp:77-92) with <name>.feat next to it.  The image size comes from
// --width/--height or, by default, from intrinsic 0 of sfm_data.json.  -f -r -g -x -y -z -d -i have
// the reference's meaning; -k=knnbow selects the knn views nearest in bag-of-features space, from the
// views' .bow files and <name>.bow next to the query's .desc (the reference computes the query's
// vector from the image with the -a / -p models, localization.cpp:386-412); -w -a -p are accepted
// and ignored; -gm switches guided matching on (localization.cpp:82).  --rank/--world shard a folder
// of queries over processes (one per GPU, each holding the map; no collective): process R handles
// every W-th query and writes its own result files.  --amat gives the OpenCV YAML file with the 3 x 4
// (or 4 x 4) matrix "A" that takes the model to global coordinates, like the server's aMatFile
// (LocalizeEngine.cc:113-144): landmarks and camera centres are transformed before localising.
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include "desc_files.h"
#include "localize_engine.h"

using namespace hulo;

static bool flag(const char *arg, const char *name, std::string &val) {
    const size_t n = strlen(name);
    if (strncmp(arg, name, n) == 0 && arg[n] == '=') { val = arg + n + 1; return true; }
    return false;
}
static bool is_dir(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
static std::string basename_part(const std::string &path) {
    std::string b = path;
    const size_t slash = b.find_last_of("/\\");
    if (slash != std::string::npos) b = b.substr(slash + 1);
    const size_t dot = b.find_last_of('.');
    if (dot != std::string::npos && dot != 0) b = b.substr(0, dot);
    return b;
}
static void fmt(std::ostream &os, double v) {
    char buf[64];
    snprintf(buf, sizeof buf, "%.6g", v);   // Eigen IOFormat precision 6
    os << buf;
}
static void save_result(const std::string &out_dir, const std::string &query, const std::string &sfm_data,
                        const std::string &matches_dir, const LocalizeEngine::LastResult *r) {
    const std::string path = out_dir + "/" + basename_part(query) + ".json";
    std::ofstream os(path);
    if (!os.is_open()) { std::cerr << "cannot write out result to " << path << std::endl; return; }
    os << "{\n\t\"filename\": \"" << query << "\",\n\t\"sfm_data\": \"" << sfm_data << "\",\n\t\"matches_dir\": \""
       << matches_dir << "\"";
    if (r) {
        auto mat = [&](const double *M) {
            os << "[";
            for (int i = 0; i < 3; ++i) {
                os << "[";
                for (int j = 0; j < 3; ++j) { fmt(os, M[3 * i + j]); if (j < 2) os << ","; }
                os << "]" << (i < 2 ? ",\n" : "");
            }
            os << "]";
        };
        os << ",\n\t\"K\": "; mat(r->K);
        os << ",\n\t\"R\": "; mat(r->R);
        os << ",\n\t\"t\": [";
        for (int i = 0; i < 3; ++i) { fmt(os, r->t_out[i]); if (i < 2) os << ",\n"; }
        os << "],\n\t\"pair\": [";
        for (size_t k = 0; k < r->inlier_pairs.size(); ++k)
            os << "[" << r->inlier_pairs[k].first << "," << r->inlier_pairs[k].second << "]"
               << (k + 1 < r->inlier_pairs.size() ? "," : "");
        os << "]";
    }
    os << "\n}\n";
}

int main(int argc, char **argv) {
    std::vector<std::string> pos;
    std::string v, sAmat;
    float fDistRatio = 0.6f;
    int ransacRound = 200, locEvryNFrame = 1, device = -1, rank = 0, world = 1, knnbow = 0;
    bool bGuided = false;
    double geomPrec = 4.0, cenX = 0, cenY = 0, cenZ = 0, cenRadius = -1.0;
    size_t width = 0, height = 0;
    unsigned long long seed = 1;
    for (int a = 1; a < argc; ++a) {
        if (flag(argv[a], "-f", v)) fDistRatio = (float)atof(v.c_str());
        else if (flag(argv[a], "-r", v)) ransacRound = atoi(v.c_str());
        else if (flag(argv[a], "-g", v)) geomPrec = atof(v.c_str());
        else if (flag(argv[a], "-x", v)) cenX = atof(v.c_str());
        else if (flag(argv[a], "-y", v)) cenY = atof(v.c_str());
        else if (flag(argv[a], "-z", v)) cenZ = atof(v.c_str());
        else if (flag(argv[a], "-d", v)) cenRadius = atof(v.c_str());
        else if (flag(argv[a], "-i", v)) locEvryNFrame = atoi(v.c_str());
        else if (flag(argv[a], "-k", v)) knnbow = atoi(v.c_str());
        else if (flag(argv[a], "--width", v)) width = (size_t)atoll(v.c_str());
        else if (flag(argv[a], "--height", v)) height = (size_t)atoll(v.c_str());
        else if (flag(argv[a], "--device", v)) device = atoi(v.c_str());
        else if (flag(argv[a], "--seed", v)) seed = strtoull(v.c_str(), nullptr, 10);
        else if (flag(argv[a], "--amat", v)) sAmat = v;
        else if (flag(argv[a], "--rank", v)) rank = atoi(v.c_str());
        else if (flag(argv[a], "--world", v)) world = atoi(v.c_str());
        else if (strcmp(argv[a], "-gm") == 0) bGuided = true;
        else if (flag(argv[a], "-gm", v)) bGuided = v != "false" && v != "0";
        else if (argv[a][0] == '-' && !(argv[a][1] >= '0' && argv[a][1] <= '9')) continue;
        else pos.push_back(argv[a]);
    }
    if (pos.size() < 4) {
        std::cerr << "usage: hulo_localize <query .desc file or folder> <sfmDir> <matchDir> <outDir> [-f=] [-r=] [-g=] "
                     "[-x= -y= -z= -d=] [-i=] [--width= --height=]\n";
        return 1;
    }
    const std::string sQuery = pos[0], sSfMDir = pos[1], sMatchesDir = pos[2], sOutputFolder = pos[3];
    std::cout << "Start localizing input image." << std::endl;

    std::vector<std::string> list;
    if (is_dir(sQuery)) {
        if (DIR *d = opendir(sQuery.c_str())) {
            while (dirent *e = readdir(d)) {
                const std::string name = e->d_name;
                if (name.size() > 5 && name.substr(name.size() - 5) == ".desc") list.push_back(sQuery + "/" + name);
            }
            closedir(d);
        }
        std::sort(list.begin(), list.end());
        if (list.empty()) { std::cout << ".desc file is not found in input directory" << std::endl; return EXIT_FAILURE; }
    } else {
        list.push_back(sQuery);
    }
    if (locEvryNFrame <= 0) {
        std::cout << "Number of frame set to skip is invalid. Reset to localize every frame" << std::endl;
        locEvryNFrame = 1;
    }
    const std::string sSfM_data = sSfMDir + (sSfMDir.back() == '/' ? "" : "/") + "sfm_data.json";
    try {
        if (world < 1 || rank < 0 || rank >= world) { std::cerr << "bad --rank/--world\n"; return 1; }
        LocalizeEngine engine(sSfMDir, sMatchesDir, sAmat, fDistRatio, ransacRound, geomPrec, bGuided, 0, knnbow,
                              device >= 0 ? device : rank);
        if (cenRadius > 0) engine.setLocalViews({cenX, cenY, cenZ}, cenRadius);
        const Intrinsic &cam = engine.scene().intrinsics.at(0);
        if (width == 0) width = cam.width;
        if (height == 0) height = cam.height;
        int imageNumber = 0, matchNextNFrame = 0, n_ok = 0;
        for (const std::string &q : list) {
            imageNumber++;
            if ((imageNumber - 1) % world != rank) continue;          // another process's query
            // video mode: localise every i-th frame, and the frames right after a success (:293-301, :583-587)
            if (imageNumber % locEvryNFrame == 0) {
            } else if (matchNextNFrame <= 0) {
                continue;
            } else {
                matchNextNFrame--;
            }
            std::vector<uint8_t> rows;
            size_t n = 0;
            FeatureLocations feats;
            readAKAZEBin(q, rows, n);
            readFeatFile(q.substr(0, q.size() - 4) + "feat", feats);
            if (feats.size() != n) {
                std::cerr << "cannot load region of " << q << " (" << n << " descriptors, " << feats.size() << " features)\n";
                save_result(sOutputFolder, q, sSfM_data, sMatchesDir, nullptr);
                continue;
            }
            std::cout << "image # " << imageNumber << "/" << list.size() << std::endl;
            std::vector<double> p2, p3, times;
            std::vector<int> inl;
            std::vector<float> qbow;
            if (knnbow > 0) {
                int br = 0, bc = 0;
                std::vector<double> bv;
                if (readMatBin(q.substr(0, q.size() - 4) + "bow", br, bc, bv))
                    for (double x : bv) qbow.push_back((float)x);
            }
            const std::vector<double> pose = engine.localize(rows.data(), n, 64, feats, width, height, "", false, p2, p3, inl,
                                                             true, times, std::vector<double>(), -1.0, seed + imageNumber,
                                                             qbow.empty() ? nullptr : &qbow);
            if (times.size() == 6)
                std::cout << "Putative matching: " << times[3] << " s\nGeometric matching: " << times[4] << " s\nPnP: "
                          << times[5] << " s\n";
            if (pose.empty()) {
                save_result(sOutputFolder, q, sSfM_data, sMatchesDir, nullptr);
                continue;
            }
            std::cout << "#inliers = " << engine.last().inlier_pairs.size() << "\ncomplete" << std::endl;
            save_result(sOutputFolder, q, sSfM_data, sMatchesDir, &engine.last());
            matchNextNFrame = locEvryNFrame - 1;
            ++n_ok;
        }
        std::cout << "localized " << n_ok << " of " << list.size() << " queries" << std::endl;
    } catch (const std::exception &e) {
        std::cerr << "hulo_localize: " << e.what() << std::endl;
        return EXIT_FAILURE;
    }
    return 0;
}
