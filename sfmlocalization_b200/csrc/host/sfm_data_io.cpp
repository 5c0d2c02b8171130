#include "sfm_data_io.h"

#include <cmath>
#include <fstream>
#include <sstream>

#include "json_min.h"

namespace hulo {

void Intrinsic::K(double out[9]) const {
    out[0] = focal; out[1] = 0; out[2] = ppx;
    out[3] = 0; out[4] = focal; out[5] = ppy;
    out[6] = 0; out[7] = 0; out[8] = 1;
}

namespace {
// r2 (1 + k1 r2 + k2 r2^2 + k3 r2^3)^2: squared radius after distortion
double disto_functor(const std::vector<double> &k, double r2) {
    const double k1 = k.size() > 0 ? k[0] : 0.0, k2 = k.size() > 1 ? k[1] : 0.0, k3 = k.size() > 2 ? k[2] : 0.0;
    const double f = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3));
    return r2 * f * f;
}
// radial_distortion::bisection_Radius_Solve: the undistorted r2 whose distorted value is r2
double bisection_radius_solve(const std::vector<double> &k, double r2, double epsilon = 1e-10) {
    double lower = r2, upper = r2;
    while (disto_functor(k, lower) > r2) lower /= 1.05;
    while (disto_functor(k, upper) < r2) upper *= 1.05;
    while (epsilon < upper - lower) {
        const double mid = 0.5 * (lower + upper);
        if (disto_functor(k, mid) > r2) upper = mid;
        else lower = mid;
    }
    return 0.5 * (lower + upper);
}
}  // namespace

std::pair<double, double> Intrinsic::get_ud_pixel(double x, double y) const {
    bool has = false;
    for (double d : disto) has = has || d != 0.0;
    if (!has) return std::make_pair(x, y);
    const double cx = (x - ppx) / focal, cy = (y - ppy) / focal;      // ima2cam
    const double r2 = cx * cx + cy * cy;
    const double radius = r2 == 0.0 ? 1.0 : std::sqrt(bisection_radius_solve(disto, r2) / r2);
    return std::make_pair(focal * radius * cx + ppx, focal * radius * cy + ppy);   // cam2ima(remove_disto)
}

static const json::Value *unwrap(const json::Value *v) {
    // cereal: {"polymorphic_id":..,"ptr_wrapper":{"id":..,"data":{...}}} or the data object itself
    if (!v) return nullptr;
    if (const json::Value *d = v->path({"ptr_wrapper", "data"})) return d;
    return v;
}

bool loadSfMData(const std::string &sfm_data_json, SfMScene &scene) {
    std::ifstream f(sfm_data_json);
    if (!f.is_open()) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string text = ss.str();
    json::Value root;
    json::Parser parser(text);
    if (!parser.parse(root) || root.type != json::Value::Object) return false;
    if (const json::Value *rp = root.get("root_path")) scene.root_path = rp->str;

    if (const json::Value *views = root.get("views"))
        for (const json::Value &kv : views->arr) {
            const json::Value *d = unwrap(kv.get("value"));
            if (!d) return false;
            View v;
            v.id_view = (std::size_t)(d->get("id_view") ? d->get("id_view")->number() : kv.get("key")->number());
            if (const json::Value *fn = d->get("filename")) v.s_Img_path = fn->str;
            if (const json::Value *w = d->get("width")) v.ui_width = (std::size_t)w->number();
            if (const json::Value *h = d->get("height")) v.ui_height = (std::size_t)h->number();
            if (const json::Value *i = d->get("id_intrinsic")) v.id_intrinsic = (std::size_t)i->number();
            if (const json::Value *p = d->get("id_pose")) v.id_pose = (std::size_t)p->number();
            scene.views[v.id_view] = v;
        }
    if (const json::Value *intr = root.get("intrinsics"))
        for (const json::Value &kv : intr->arr) {
            const json::Value *val = kv.get("value");
            const json::Value *d = unwrap(val);
            if (!d || !kv.get("key")) return false;
            Intrinsic in;
            if (const json::Value *n = val->get("polymorphic_name")) in.type = n->str;
            if (const json::Value *w = d->get("width")) in.width = (std::size_t)w->number();
            if (const json::Value *h = d->get("height")) in.height = (std::size_t)h->number();
            if (const json::Value *fl = d->get("focal_length")) in.focal = fl->number();
            if (const json::Value *pp = d->get("principal_point"))
                if (pp->arr.size() >= 2) { in.ppx = pp->arr[0].number(); in.ppy = pp->arr[1].number(); }
            for (const char *key : {"disto_k1", "disto_k3"})
                if (const json::Value *k = d->get(key))
                    for (const json::Value &c : k->arr) in.disto.push_back(c.number());
            scene.intrinsics[(std::size_t)kv.get("key")->number()] = in;
        }
    if (const json::Value *ext = root.get("extrinsics"))
        for (const json::Value &kv : ext->arr) {
            const json::Value *d = kv.get("value");
            if (!d || !kv.get("key")) return false;
            Pose p;
            if (const json::Value *r = d->get("rotation"))
                for (std::size_t i = 0; i < 3 && i < r->arr.size(); ++i)
                    for (std::size_t j = 0; j < 3 && j < r->arr[i].arr.size(); ++j) p.R[3 * i + j] = r->arr[i].arr[j].number();
            if (const json::Value *c = d->get("center"))
                for (std::size_t i = 0; i < 3 && i < c->arr.size(); ++i) p.center[i] = c->arr[i].number();
            scene.poses[(std::size_t)kv.get("key")->number()] = p;
        }
    if (const json::Value *st = root.get("structure")) {
        std::map<std::size_t, Landmark> sorted;
        for (const json::Value &kv : st->arr) {
            const json::Value *d = kv.get("value");
            if (!d || !kv.get("key")) return false;
            Landmark lm;
            lm.id = (std::size_t)kv.get("key")->number();
            if (const json::Value *x = d->get("X"))
                for (std::size_t i = 0; i < 3 && i < x->arr.size(); ++i) lm.X[i] = x->arr[i].number();
            if (const json::Value *obs = d->get("observations"))
                for (const json::Value &o : obs->arr) {
                    const json::Value *ov = o.get("value");
                    if (!ov || !o.get("key") || !ov->get("id_feat")) return false;
                    Observation ob;
                    ob.id_view = (std::size_t)o.get("key")->number();
                    ob.id_feat = (std::size_t)ov->get("id_feat")->number();
                    if (const json::Value *x = ov->get("x"))
                        for (std::size_t i = 0; i < 2 && i < x->arr.size(); ++i) ob.x[i] = x->arr[i].number();
                    lm.obs.push_back(ob);
                }
            sorted[lm.id] = std::move(lm);
        }
        for (auto &kv : sorted) scene.landmarks.push_back(std::move(kv.second));
    }
    return !scene.views.empty();
}

bool saveSfMDataPoses(const std::string &in_json, const std::string &out_json, const std::map<std::size_t, Pose> &poses) {
    std::ifstream f(in_json);
    if (!f.is_open()) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string text = ss.str();
    json::Value root;
    json::Parser parser(text);
    if (!parser.parse(root) || root.type != json::Value::Object) return false;
    json::Value *ext = root.find("extrinsics");
    if (!ext) {
        root.obj.emplace_back("extrinsics", json::Value());
        ext = &root.obj.back().second;
    }
    ext->type = json::Value::Array;
    ext->arr.clear();
    for (const auto &kv : poses) {
        json::Value rot;
        rot.type = json::Value::Array;
        for (int i = 0; i < 3; ++i) {
            json::Value row;
            row.type = json::Value::Array;
            for (int j = 0; j < 3; ++j) {
                row.arr.emplace_back();
                row.arr.back().set_number(kv.second.R[3 * i + j]);
            }
            rot.arr.push_back(std::move(row));
        }
        json::Value center;
        center.type = json::Value::Array;
        for (int i = 0; i < 3; ++i) {
            center.arr.emplace_back();
            center.arr.back().set_number(kv.second.center[i]);
        }
        json::Value value;
        value.type = json::Value::Object;
        value.obj.emplace_back("rotation", std::move(rot));
        value.obj.emplace_back("center", std::move(center));
        json::Value key;
        key.set_number((double)kv.first);
        json::Value entry;
        entry.type = json::Value::Object;
        entry.obj.emplace_back("key", std::move(key));
        entry.obj.emplace_back("value", std::move(value));
        ext->arr.push_back(std::move(entry));
    }
    std::string out;
    json::dump(root, out);
    out += "\n";
    std::ofstream o(out_json);
    if (!o.is_open()) return false;
    o << out;
    return (bool)o;
}

bool readOpenCVMatrix(const std::string &yaml_file, const std::string &name, int &rows, int &cols,
                      std::vector<double> &data) {
    std::ifstream f(yaml_file);
    if (!f.is_open()) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string s = ss.str();
    std::size_t pos = s.find("\n" + name + ":");
    if (pos == std::string::npos) pos = s.rfind(name + ":", 0) == 0 ? 0 : std::string::npos;
    if (pos == std::string::npos) return false;
    auto field = [&](const char *key, std::size_t from) -> std::size_t {
        const std::size_t k = s.find(key, from);
        return k == std::string::npos ? k : k + strlen(key);
    };
    const std::size_t r = field("rows:", pos), c = field("cols:", pos), d = field("data:", pos);
    if (r == std::string::npos || c == std::string::npos || d == std::string::npos) return false;
    rows = atoi(s.c_str() + r);
    cols = atoi(s.c_str() + c);
    const std::size_t b0 = s.find('[', d), b1 = s.find(']', d);
    if (b0 == std::string::npos || b1 == std::string::npos) return false;
    data.clear();
    const char *p = s.c_str() + b0 + 1, *end = s.c_str() + b1;
    while (p < end) {
        char *e = nullptr;
        const double v = strtod(p, &e);
        if (e == p) { ++p; continue; }
        data.push_back(v);
        p = e;
    }
    return (int)data.size() == rows * cols;
}

}  // namespace hulo
