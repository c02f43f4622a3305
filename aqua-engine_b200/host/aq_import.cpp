/*
 * aq_import.cpp — asset import: Wavefront OBJ/MTL -> scene JSON + BSON .mesh files
 * (the reference did this with the `tobj` crate, Cargo.toml:23).
 *
 * The mapping is the one the shipped data was produced with, reverse-engineered in SURVEY §2.5
 * from CornellBox-Original.{obj,mtl} <-> cbox.json and living_room.mtl <-> room.json:
 *   - one TriangleMesh per OBJ group / material run, named  <obj>_<group>_<index>.mesh,
 *     polygons fan-triangulated (0,1,2),(0,2,3)..., unreferenced vertices dropped, vertex
 *     normals = un-normalised mean of the unit face normals of the faces that use the vertex
 *     (scenes/CornellBox-Original_leftWall_4.mesh: n = (0.99991262, 0.01004899, 0.00492531));
 *   - Principled:  color = Srgb(oetf(Kd))  (Ks when Kd == 0;  Image(map_Kd) when present),
 *     roughness = sqrt(2/(Ns+2)),  metallic = 0 if Ks == 0 else 1/(1+max(Kd)),  Ke/d/Ni dropped,
 *     every other input at the Blender default found in scenes/cbox.json:5-63;
 *   - BSON layout of SURVEY §2.4 (f64 coordinates, i64 indices, keys name/vertices/normals/
 *     texcoords/indices).
 * Camera and light are not in an OBJ: a perspective camera framing the bounds and one point
 * light are written so that the result renders as is.
 */
#include "aqua_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <sstream>
#include <string>
#include <vector>

namespace {

thread_local std::string g_ierr;

struct Mtl {
    std::string name;
    float kd[3] = {0.8f, 0.8f, 0.8f}, ks[3] = {0, 0, 0};
    float ns = 0.f;
    std::string map_kd;
};

struct Group {
    std::string name, mtl;
    std::vector<std::vector<int>> faces_v, faces_t; /* 0-based global indices, -1 = none */
};

double oetf(double c) { return c <= 0.0031308 ? 12.92 * c : 1.055 * std::pow(c, 1.0 / 2.4) - 0.055; }

/* ---- BSON writer for the subset of SURVEY §2.4 */
struct BsonOut {
    std::vector<uint8_t> b;
    void i32(int32_t v) { b.insert(b.end(), (uint8_t*)&v, (uint8_t*)&v + 4); }
    void f64(double v) { b.insert(b.end(), (uint8_t*)&v, (uint8_t*)&v + 8); }
    void i64(int64_t v) { b.insert(b.end(), (uint8_t*)&v, (uint8_t*)&v + 8); }
    void key(uint8_t type, const std::string& k) {
        b.push_back(type);
        b.insert(b.end(), k.begin(), k.end());
        b.push_back(0);
    }
    size_t begin_doc() {
        size_t at = b.size();
        i32(0);
        return at;
    }
    void end_doc(size_t at) {
        b.push_back(0);
        int32_t len = (int32_t)(b.size() - at);
        std::memcpy(&b[at], &len, 4);
    }
    template <class T>
    void tuples(const std::string& k, const std::vector<T>& flat, int arity, bool integer) {
        key(0x04, k);
        size_t a = begin_doc();
        for (size_t i = 0; i < flat.size() / (size_t)arity; ++i) {
            key(0x04, std::to_string(i));
            size_t t = begin_doc();
            for (int c = 0; c < arity; ++c) {
                key(integer ? 0x12 : 0x01, std::to_string(c));
                if (integer)
                    i64((int64_t)flat[i * arity + c]);
                else
                    f64((double)flat[i * arity + c]);
            }
            end_doc(t);
        }
        end_doc(a);
    }
};

bool write_mesh(const std::string& path, const std::string& name, const std::vector<float>& pos,
                const std::vector<float>& nrm, const std::vector<float>& uv, const std::vector<uint32_t>& idx) {
    BsonOut o;
    size_t d = o.begin_doc();
    o.key(0x02, "name");
    o.i32((int32_t)name.size() + 1);
    o.b.insert(o.b.end(), name.begin(), name.end());
    o.b.push_back(0);
    o.tuples("vertices", pos, 3, false);
    o.tuples("normals", nrm, 3, false);
    o.tuples("texcoords", uv, 2, false);
    o.tuples("indices", idx, 3, true);
    o.end_doc(d);
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = std::fwrite(o.b.data(), 1, o.b.size(), f) == o.b.size();
    std::fclose(f);
    return ok;
}

std::string base_name(const std::string& p) {
    size_t s = p.find_last_of('/');
    std::string f = s == std::string::npos ? p : p.substr(s + 1);
    size_t d = f.find_last_of('.');
    return d == std::string::npos ? f : f.substr(0, d);
}
std::string dir_name(const std::string& p) {
    size_t s = p.find_last_of('/');
    return s == std::string::npos ? "." : p.substr(0, s);
}

bool load_mtl(const std::string& path, std::map<std::string, Mtl>* out) {
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return false;
    char line[4096];
    Mtl* cur = nullptr;
    while (std::fgets(line, sizeof line, f)) {
        std::istringstream ss(line);
        std::string t;
        if (!(ss >> t) || t[0] == '#') continue;
        if (t == "newmtl") {
            std::string n;
            ss >> n;
            cur = &(*out)[n];
            cur->name = n;
        } else if (cur) {
            if (t == "Kd") ss >> cur->kd[0] >> cur->kd[1] >> cur->kd[2];
            else if (t == "Ks") ss >> cur->ks[0] >> cur->ks[1] >> cur->ks[2];
            else if (t == "Ns") ss >> cur->ns;
            else if (t == "map_Kd") {
                std::string rest;
                std::getline(ss, rest);
                size_t a = rest.find_first_not_of(" \t"), b = rest.find_last_not_of(" \t\r\n");
                if (a != std::string::npos) cur->map_kd = rest.substr(a, b - a + 1);
            }
        }
    }
    std::fclose(f);
    return true;
}

void json_tex(std::ostringstream& o, const char* key, const std::string& variant, const float* v, int n,
              const std::string* img, bool last = false) {
    o << "        \"" << key << "\": {\"" << variant << "\": ";
    char buf[64];
    if (img) {
        std::string p = *img;
        std::string esc;
        for (char c : p) esc += (c == '\\' || c == '/') ? std::string("\\\\") : std::string(1, c);
        o << "\"" << esc << "\"";
    } else if (n == 1) {
        std::snprintf(buf, sizeof buf, "%.17g", (double)v[0]);
        o << buf;
        if (!std::strchr(buf, '.') && !std::strchr(buf, 'e')) o << ".0";
    } else {
        o << "[";
        for (int i = 0; i < n; ++i) {
            std::snprintf(buf, sizeof buf, "%.17g", (double)v[i]);
            o << buf;
            if (!std::strchr(buf, '.') && !std::strchr(buf, 'e')) o << ".0";
            if (i + 1 < n) o << ", ";
        }
        o << "]";
    }
    o << "}" << (last ? "" : ",") << "\n";
}

}  // namespace

#define AQ_HOST_CATCH                                                      \
    catch (const std::bad_alloc&) { g_ierr = "out of memory"; return AQ_ERR_OOM; } \
    catch (const std::exception& e) { g_ierr = e.what(); return AQ_ERR_IO; }       \
    catch (...) { g_ierr = "internal error"; return AQ_ERR_IO; }

extern "C" {

const char* aq_host_import_last_error(void) { return g_ierr.c_str(); }

/* obj_path -> <out_dir>/<scene_name>.json + <out_dir>/<obj>_<group>_<i>.mesh; returns the number
 * of meshes written, or a negative aq_status */
int aq_host_import_obj(const char* obj_path, const char* out_dir, const char* scene_name) try {
    if (!obj_path || !out_dir || !scene_name) {
        g_ierr = "null argument";
        return AQ_ERR_BAD_ARG;
    }
    FILE* f = std::fopen(obj_path, "r");
    if (!f) {
        g_ierr = std::string("cannot read ") + obj_path;
        return AQ_ERR_IO;
    }
    std::vector<float> V, VT;
    std::vector<Group> groups;
    std::map<std::string, Mtl> mtls;
    std::string cur_group = "default", cur_mtl;
    bool need_new = true;
    char line[8192];
    while (std::fgets(line, sizeof line, f)) {
        std::istringstream ss(line);
        std::string t;
        if (!(ss >> t) || t[0] == '#') continue;
        if (t == "v") {
            float x, y, z;
            ss >> x >> y >> z;
            V.insert(V.end(), {x, y, z});
        } else if (t == "vt") {
            float u = 0, v = 0;
            ss >> u >> v;
            VT.insert(VT.end(), {u, v});
        } else if (t == "mtllib") {
            std::string n;
            ss >> n;
            load_mtl(dir_name(obj_path) + "/" + n, &mtls);
        } else if (t == "g" || t == "o") {
            std::string n;
            ss >> n;
            if (n != cur_group || groups.empty()) need_new = true;
            cur_group = n;
        } else if (t == "usemtl") { /* a new mesh starts when the group or the material CHANGES */
            std::string n;
            ss >> n;
            if (n != cur_mtl) need_new = true;
            cur_mtl = n;
        } else if (t == "f") {
            if (need_new || groups.empty()) {
                groups.push_back(Group{cur_group, cur_mtl, {}, {}});
                need_new = false;
            }
            std::vector<int> fv, ft;
            std::string tok;
            while (ss >> tok) {
                int vi = 0, ti = 0;
                const char* s = tok.c_str();
                vi = std::atoi(s);
                const char* sl = std::strchr(s, '/');
                if (sl && sl[1] != '/' && sl[1] != 0) ti = std::atoi(sl + 1);
                int nv = (int)(V.size() / 3), nt = (int)(VT.size() / 2);
                fv.push_back(vi < 0 ? nv + vi : vi - 1); /* negative = relative to the end */
                ft.push_back(ti == 0 ? -1 : (ti < 0 ? nt + ti : ti - 1));
            }
            if (fv.size() >= 3) {
                groups.back().faces_v.push_back(fv);
                groups.back().faces_t.push_back(ft);
            }
        }
    }
    std::fclose(f);

    const std::string obj_base = base_name(obj_path);
    float bmin[3] = {INFINITY, INFINITY, INFINITY}, bmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    std::ostringstream shapes;
    std::vector<std::string> used_mtls;
    int n_written = 0;
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        const Group& G = groups[gi];
        if (G.faces_v.empty()) continue;
        /* compact vertex list in order of first use; (v, vt) pairs are distinct vertices */
        std::map<std::pair<int, int>, uint32_t> remap;
        std::vector<float> pos, uv, nrm;
        std::vector<uint32_t> idx;
        bool has_uv = false;
        for (auto& ft : G.faces_t)
            for (int t : ft) has_uv |= t >= 0;
        for (size_t fi = 0; fi < G.faces_v.size(); ++fi) {
            const auto& fv = G.faces_v[fi];
            const auto& ft = G.faces_t[fi];
            std::vector<uint32_t> loc;
            for (size_t k = 0; k < fv.size(); ++k) {
                if (fv[k] < 0 || (size_t)fv[k] * 3 + 2 >= V.size()) {
                    g_ierr = "face references a missing vertex";
                    return AQ_ERR_IO;
                }
                auto key = std::make_pair(fv[k], has_uv ? ft[k] : -1);
                auto it = remap.find(key);
                if (it == remap.end()) {
                    uint32_t id = (uint32_t)(pos.size() / 3);
                    pos.insert(pos.end(), {V[3 * (size_t)fv[k]], V[3 * (size_t)fv[k] + 1], V[3 * (size_t)fv[k] + 2]});
                    if (has_uv) {
                        if (ft[k] >= 0 && (size_t)ft[k] * 2 + 1 < VT.size())
                            uv.insert(uv.end(), {VT[2 * (size_t)ft[k]], VT[2 * (size_t)ft[k] + 1]});
                        else
                            uv.insert(uv.end(), {0.f, 0.f});
                    }
                    it = remap.emplace(key, id).first;
                }
                loc.push_back(it->second);
            }
            for (size_t k = 1; k + 1 < loc.size(); ++k) idx.insert(idx.end(), {loc[0], loc[k], loc[k + 1]});
        }
        /* vertex normals: mean of the unit normals of the TRIANGLES (after fan triangulation) that
         * use the vertex — NOT normalised (leftWall v0 is the mean of two triangle
         * normals of a slightly non-planar quad, |n| = 0.999975) */
        std::vector<double> acc(pos.size(), 0.0);
        std::vector<int> cnt(pos.size() / 3, 0);
        for (size_t t = 0; t < idx.size() / 3; ++t) {
            uint32_t a = idx[3 * t], b = idx[3 * t + 1], c = idx[3 * t + 2];
            float e1[3], e2[3];
            for (int k = 0; k < 3; ++k) {
                e1[k] = pos[3 * (size_t)b + k] - pos[3 * (size_t)a + k];
                e2[k] = pos[3 * (size_t)c + k] - pos[3 * (size_t)a + k];
            }
            float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
            float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            if (!(len > 0.f)) continue;
            for (uint32_t v : {a, b, c}) {
                for (int k = 0; k < 3; ++k) acc[3 * (size_t)v + k] += (double)(n[k] / len);
                cnt[v]++;
            }
        }
        nrm.resize(pos.size());
        for (size_t v = 0; v < cnt.size(); ++v)
            for (int k = 0; k < 3; ++k) nrm[3 * v + k] = cnt[v] ? (float)(acc[3 * v + k] / cnt[v]) : 0.f;
        for (size_t i = 0; i < pos.size(); ++i) {
            bmin[i % 3] = std::fmin(bmin[i % 3], pos[i]);
            bmax[i % 3] = std::fmax(bmax[i % 3], pos[i]);
        }
        std::string file = obj_base + "_" + G.name + "_" + std::to_string(n_written) + ".mesh";
        if (!write_mesh(std::string(out_dir) + "/" + file, G.name, pos, nrm, uv, idx)) {
            g_ierr = "cannot write " + file;
            return AQ_ERR_IO;
        }
        std::string mname = G.mtl.empty() ? G.name : G.mtl;
        if (std::find(used_mtls.begin(), used_mtls.end(), mname) == used_mtls.end()) used_mtls.push_back(mname);
        shapes << (n_written ? ",\n" : "") << "    {\"Mesh\": [\"" << file << "\", {\"Named\": \"" << mname << "\"}]}";
        ++n_written;
    }
    if (n_written == 0) {
        g_ierr = "no faces in OBJ";
        return AQ_ERR_IO;
    }

    /* ---- scene JSON */
    std::ostringstream o;
    o << "{\n  \"named_bsdfs\": {\n";
    for (size_t mi = 0; mi < used_mtls.size(); ++mi) {
        Mtl m;
        auto it = mtls.find(used_mtls[mi]);
        if (it != mtls.end()) m = it->second;
        bool ks0 = m.ks[0] == 0.f && m.ks[1] == 0.f && m.ks[2] == 0.f;
        bool kd0 = m.kd[0] == 0.f && m.kd[1] == 0.f && m.kd[2] == 0.f;
        const float* c = (kd0 && !ks0) ? m.ks : m.kd;
        float srgb[3] = {(float)oetf(c[0]), (float)oetf(c[1]), (float)oetf(c[2])};
        float metallic = ks0 ? 0.f : 1.f / (1.f + std::max(m.kd[0], std::max(m.kd[1], m.kd[2])));
        float rough = std::sqrt(2.f / (m.ns + 2.f));
        float zero = 0.f, half = 0.5f, cc = 0.03f, ior = 1.45f, ssr[3] = {1.0f, 0.2f, 0.1f};
        o << "    \"" << used_mtls[mi] << "\": {\"Principled\": {\n";
        if (!m.map_kd.empty())
            json_tex(o, "color", "Image", nullptr, 0, &m.map_kd);
        else
            json_tex(o, "color", "Srgb", srgb, 3, nullptr);
        json_tex(o, "subsurface", "Float", &zero, 1, nullptr);
        json_tex(o, "subsurface_radius", "Float3", ssr, 3, nullptr);
        json_tex(o, "subsurface_color", "Float", &zero, 1, nullptr);
        json_tex(o, "metallic", "Float", &metallic, 1, nullptr);
        json_tex(o, "specular", "Float", &zero, 1, nullptr);
        json_tex(o, "specular_tint", "Float", &zero, 1, nullptr);
        json_tex(o, "roughness", "Float", &rough, 1, nullptr);
        json_tex(o, "anisotropic", "Float", &zero, 1, nullptr);
        json_tex(o, "anisotropic_rotation", "Float", &zero, 1, nullptr);
        json_tex(o, "sheen", "Float", &zero, 1, nullptr);
        json_tex(o, "sheen_tint", "Float", &half, 1, nullptr);
        json_tex(o, "clearcoat", "Float", &zero, 1, nullptr);
        json_tex(o, "clearcoat_roughness", "Float", &cc, 1, nullptr);
        json_tex(o, "ior", "Float", &ior, 1, nullptr);
        json_tex(o, "transmission", "Float", &zero, 1, nullptr);
        json_tex(o, "emission", "Float", &zero, 1, nullptr);
        o << "        \"hint\": \"ltc\"\n    }}" << (mi + 1 < used_mtls.size() ? "," : "") << "\n";
    }
    float cx = 0.5f * (bmin[0] + bmax[0]), cy = 0.5f * (bmin[1] + bmax[1]);
    float ext = std::max(bmax[0] - bmin[0], bmax[1] - bmin[1]);
    float fov = 30.f, dist = 0.5f * ext / std::tan(0.5f * fov * 3.14159265f / 180.f) * 1.1f;
    char buf[512];
    std::snprintf(buf, sizeof buf,
                  "  },\n  \"camera\": {\"Perspective\": {\"res\": [512, 512], \"fov\": %.9g, \"lens_radius\": 0.0, \"focal\": 1.0,\n"
                  "    \"transform\": {\"translate\": [%.9g, %.9g, %.9g], \"rotate\": [0.0, 0.0, 0.0], \"scale\": [1.0, 1.0, 1.0]}}},\n",
                  (double)fov, (double)cx, (double)cy, (double)(bmax[2] + dist));
    o << buf;
    std::snprintf(buf, sizeof buf,
                  "  \"lights\": [{\"Point\": {\"pos\": [%.9g, %.9g, %.9g], \"emission\": {\"Srgb\": [1.0, 1.0, 1.0]}}}],\n",
                  (double)cx, (double)(bmin[1] + 0.85f * (bmax[1] - bmin[1])), (double)(0.5f * (bmin[2] + bmax[2])));
    o << buf;
    o << "  \"shapes\": [\n" << shapes.str() << "\n  ]\n}\n";
    std::string jp = std::string(out_dir) + "/" + scene_name + ".json";
    FILE* jf = std::fopen(jp.c_str(), "w");
    if (!jf) {
        g_ierr = "cannot write " + jp;
        return AQ_ERR_IO;
    }
    std::string js = o.str();
    std::fwrite(js.data(), 1, js.size(), jf);
    std::fclose(jf);
    return n_written;
} AQ_HOST_CATCH

}  // extern "C"
