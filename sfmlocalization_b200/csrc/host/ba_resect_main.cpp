// hulo_ba_resect -- the resection stage of the reference's OpenMVG_BA tool
// (OpenMVG_BA/src/adjust_sfm_data.cpp:57-155) on the GPU:
//
//   hulo_ba_resect <in sfm_data.json> <out sfm_data.json> [-c=] [-r=0] [--device=D] [--seed=S] [--iter=4096]
//
// loads the reconstruction, re-resects every view with more than 10 observations against the
// structure (all views in one batched call) and writes the reconstruction with the new poses --
// the file the reference saves as sfm_data_b4bd.json before bundle adjustment (:152-155).  The
// Ceres bundle adjustment the reference runs afterwards for the tokens of -c (:157-244) is outside
// the accelerated path: a non-empty -c is refused, -r is accepted and ignored.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "resect_views.h"

using namespace hulo;

static bool flag(const char *arg, const char *name, std::string &val) {
    const size_t n = strlen(name);
    if (strncmp(arg, name, n) == 0 && arg[n] == '=') { val = arg + n + 1; return true; }
    return false;
}

int main(int argc, char **argv) {
    std::string in, out, v, bd_command;
    int device = 0;
    uint64_t seed = 0x5eed;
    size_t max_iter = 4096;
    for (int i = 1; i < argc; ++i) {
        if (flag(argv[i], "-c", v)) bd_command = v;
        else if (flag(argv[i], "-r", v)) {}
        else if (flag(argv[i], "--device", v)) device = atoi(v.c_str());
        else if (flag(argv[i], "--seed", v)) seed = strtoull(v.c_str(), nullptr, 0);
        else if (flag(argv[i], "--iter", v)) max_iter = (size_t)strtoull(v.c_str(), nullptr, 0);
        else if (in.empty()) in = argv[i];
        else if (out.empty()) out = argv[i];
    }
    if (in.empty() || out.empty()) {
        std::cerr << "usage: hulo_ba_resect <in sfm_data.json> <out sfm_data.json> [-c=] [-r=0] [--device=D] [--seed=S] [--iter=4096]" << std::endl;
        return 1;
    }
    if (!bd_command.empty()) {
        std::cerr << "hulo_ba_resect: bundle adjustment (-c=" << bd_command << ") is not part of this tool; run it on the output" << std::endl;
        return 1;
    }
    std::cout << "Reading sfm_data.json file : " << in << std::endl;
    SfMScene scene;
    if (!loadSfMData(in, scene)) {
        std::cerr << std::endl << "The input sfm_data.json file \"" << in << "\" cannot be read." << std::endl;
        return EXIT_FAILURE;
    }
    hulo_gpu *h = nullptr;
    if (hulo_gpu_create(device, &h) != HULO_OK) {
        std::cerr << "hulo_ba_resect: no CUDA device " << device << std::endl;
        return EXIT_FAILURE;
    }
    ResectViewsReport rep;
    const int rc = resectViews(h, scene, max_iter, seed, &rep);
    if (rc != HULO_OK) {
        std::cerr << "hulo_ba_resect: " << hulo_last_error() << std::endl;
        hulo_gpu_destroy(h);
        return EXIT_FAILURE;
    }
    hulo_gpu_destroy(h);
    if (rep.too_few_warning) std::cout << "Warning: there is/are frames with too few matches." << std::endl;
    std::cout << "Resected " << rep.resected << " of " << rep.attempted << " views (" << rep.views << " in the file); gather "
              << rep.ms_gather << " ms, resection " << rep.ms_resect << " ms" << std::endl;
    if (!saveSfMDataPoses(in, out, scene.poses)) {
        std::cerr << "Cannot write " << out << std::endl;
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}
