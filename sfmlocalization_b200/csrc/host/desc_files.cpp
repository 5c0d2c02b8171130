#include "desc_files.h"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace hulo {

bool readAKAZEBin(const std::string &filename, std::vector<uint8_t> &rows, std::size_t &count) {
    rows.clear();
    count = 0;
    std::ifstream f(filename, std::ios::binary);
    if (!f.is_open()) return false;
    uint64_t n = 0;
    f.read(reinterpret_cast<char *>(&n), sizeof n);
    if (!f) return false;
    rows.resize((std::size_t)n * 64);
    if (n) f.read(reinterpret_cast<char *>(rows.data()), (std::streamsize)rows.size());
    if (!f) { rows.clear(); return false; }
    count = (std::size_t)n;
    return true;
}

bool saveAKAZEBin(const std::string &filename, const uint8_t *rows, std::size_t count, std::size_t width) {
    std::ofstream f(filename, std::ios::binary);
    if (!f.is_open()) return false;
    const uint64_t n = count;
    f.write(reinterpret_cast<const char *>(&n), sizeof n);
    uint8_t row[64];
    const std::size_t w = width < 64 ? width : 64;
    for (std::size_t i = 0; i < count; ++i) {
        memcpy(row, rows + i * width, w);
        if (w < 64) memset(row + w, 0, 64 - w);
        f.write(reinterpret_cast<const char *>(row), 64);
    }
    return (bool)f;
}

bool exportPairWiseMatches(const PairWiseMatches &matches, const std::string &filename) {
    std::ofstream f(filename);
    if (!f.is_open()) return false;
    for (const auto &kv : matches) {
        f << kv.first.first << ' ' << kv.first.second << '\n' << kv.second.size() << '\n';
        for (const IndMatch &m : kv.second) f << m.i_ << ' ' << m.j_ << '\n';
    }
    return (bool)f;
}

bool importPairWiseMatches(const std::string &filename, PairWiseMatches &matches) {
    std::ifstream f(filename);
    if (!f.is_open()) return false;
    std::size_t I, J, n;
    while (f >> I >> J >> n) {
        IndMatches v(n);
        for (std::size_t k = 0; k < n; ++k)
            if (!(f >> v[k].i_ >> v[k].j_)) return false;
        matches[Pair(I, J)] = v;
    }
    return true;
}

bool readPairFile(const std::string &filename, std::vector<Pair> &pairs) {
    std::ifstream f(filename);
    if (!f.is_open()) return false;
    std::size_t a, b;
    while (f >> a >> b) pairs.push_back(Pair(a, b));
    return true;
}

bool readViewsFromList(const std::string &filename, Views &views) {
    std::ifstream f(filename);
    if (!f.is_open()) return false;
    std::size_t id;
    std::string path;
    while (f >> id >> path) views[id] = View{id, path};
    return true;
}

// Minimal scanner: inside the "views" array every element carries "filename": "<path>" and
// "id_view": <n> (cereal writes filename first).  Stops at the "intrinsics" key.
bool readViewsFromSfmData(const std::string &sfm_data_json, Views &views) {
    std::ifstream f(sfm_data_json);
    if (!f.is_open()) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string s = ss.str();
    std::size_t pos = s.find("\"views\"");
    if (pos == std::string::npos) return false;
    const std::size_t end = s.find("\"intrinsics\"", pos);
    for (;;) {
        std::size_t kf = s.find("\"filename\"", pos);
        if (kf == std::string::npos || (end != std::string::npos && kf > end)) break;
        std::size_t q0 = s.find('"', s.find(':', kf) + 1);
        std::size_t q1 = s.find('"', q0 + 1);
        if (q0 == std::string::npos || q1 == std::string::npos) return false;
        const std::string path = s.substr(q0 + 1, q1 - q0 - 1);
        std::size_t ki = s.find("\"id_view\"", q1);
        if (ki == std::string::npos) return false;
        const std::size_t id = (std::size_t)std::strtoull(s.c_str() + s.find(':', ki) + 1, nullptr, 10);
        View view{id, path};
        // "width" / "height" sit between "filename" and "id_view" in the cereal layout
        const std::size_t kw = s.find("\"width\"", q1), kh = s.find("\"height\"", q1);
        if (kw != std::string::npos && kw < ki) view.ui_width = (std::size_t)std::strtoull(s.c_str() + s.find(':', kw) + 1, nullptr, 10);
        if (kh != std::string::npos && kh < ki) view.ui_height = (std::size_t)std::strtoull(s.c_str() + s.find(':', kh) + 1, nullptr, 10);
        views[id] = view;
        pos = ki + 9;
    }
    return !views.empty();
}

bool readFeatFile(const std::string &filename, FeatureLocations &feats) {
    feats.clear();
    std::ifstream f(filename);
    if (!f.is_open()) return false;
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream ls(line);
        double x, y;
        if (ls >> x >> y) feats.push_back(std::make_pair(x, y));
    }
    return true;
}

std::string featPath(const std::string &dir, const std::string &img_path) {
    std::string p = descPath(dir, img_path, true);
    return p.substr(0, p.size() - 4) + "feat";
}

bool loadRegions(const Views &views, const std::string &dir, RegionsProvider &regions) {
    for (const auto &kv : views)
        if (!readFeatFile(featPath(dir, kv.second.s_Img_path), regions[kv.first])) return false;
    return true;
}

bool readMatBin(const std::string &filename, int &rows, int &cols, std::vector<double> &values) {
    values.clear();
    rows = cols = 0;
    std::ifstream f(filename, std::ios::binary);
    if (!f.is_open()) return false;
    int32_t r = 0, c = 0, type = 0;
    f.read(reinterpret_cast<char *>(&r), 4);
    if (!f || r == 0) return (bool)f;                       // an empty matrix stores only its row count (:66-68)
    f.read(reinterpret_cast<char *>(&c), 4);
    f.read(reinterpret_cast<char *>(&type), 4);
    if (!f || r < 0 || c < 0 || (type >> 3) != 0) return false;   // one channel only
    const size_t n = (size_t)r * (size_t)c;
    values.resize(n);
    auto load = [&](auto tag) {
        typedef decltype(tag) T;
        std::vector<T> buf(n);
        f.read(reinterpret_cast<char *>(buf.data()), (std::streamsize)(n * sizeof(T)));
        for (size_t k = 0; k < n; ++k) values[k] = (double)buf[k];
        return (bool)f;
    };
    bool ok = false;
    switch (type & 7) {
        case 0: ok = load((uint8_t)0); break;
        case 1: ok = load((int8_t)0); break;
        case 2: ok = load((uint16_t)0); break;
        case 3: ok = load((int16_t)0); break;
        case 4: ok = load((int32_t)0); break;
        case 5: ok = load((float)0); break;
        case 6: ok = load((double)0); break;
        default: ok = false;
    }
    if (!ok) { values.clear(); return false; }
    rows = r; cols = c;
    return true;
}

bool saveMatBin(const std::string &filename, int rows, int cols, int cv_type, const double *values) {
    std::ofstream f(filename, std::ios::binary);
    if (!f.is_open()) return false;
    const int32_t r = rows, c = cols, t = cv_type;
    f.write(reinterpret_cast<const char *>(&r), 4);
    if (rows == 0) return (bool)f;
    f.write(reinterpret_cast<const char *>(&c), 4);
    f.write(reinterpret_cast<const char *>(&t), 4);
    const size_t n = (size_t)rows * (size_t)cols;
    for (size_t k = 0; k < n; ++k) {
        if (cv_type == 5) { const float v = (float)values[k]; f.write(reinterpret_cast<const char *>(&v), 4); }
        else if (cv_type == 6) { f.write(reinterpret_cast<const char *>(&values[k]), 8); }
        else if (cv_type == 4) { const int32_t v = (int32_t)values[k]; f.write(reinterpret_cast<const char *>(&v), 4); }
        else if (cv_type == 0) { const uint8_t v = (uint8_t)values[k]; f.write(reinterpret_cast<const char *>(&v), 1); }
        else return false;
    }
    return (bool)f;
}

std::string bowPath(const std::string &dir, const std::string &img_path) {
    std::string p = descPath(dir, img_path, true);
    return p.substr(0, p.size() - 4) + "bow";
}

std::string descPath(const std::string &dir, const std::string &img_path, bool strip_extension) {
    std::string base = img_path;
    const std::size_t slash = base.find_last_of("/\\");
    if (slash != std::string::npos) base = base.substr(slash + 1);
    if (strip_extension) {
        const std::size_t dot = base.find_last_of('.');
        if (dot != std::string::npos && dot != 0) base = base.substr(0, dot);
    }
    std::string d = dir;
    if (!d.empty() && d.back() != '/') d += '/';
    return d + base + ".desc";
}

}  // namespace hulo
