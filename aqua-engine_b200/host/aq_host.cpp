/*
 * aq_host.cpp — host-side scene ingest (C++ mirror of the Rust host `arukas`).
 *
 * Follows the reference's on-disk data contract only (there is no reference code):
 *   Scene / Camera / Light / Shape / Bsdf / Texture serde enums   scenes/cbox.json:1-627
 *   integrator config                                             scenes/integrator.json:1-8
 *   BSON TriangleMesh {name,vertices,normals,texcoords,indices}   scenes/ *.mesh
 * Output: one flat aq_scene_desc (include/aqua_cuda.h) with all shapes concatenated in
 * shapes[] order, so global primitive ids follow SURVEY §2.4's prefix table.
 */
#include "aqua_host.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <vector>

int aq_jpeg_decode_file(const char* path, uint32_t* w, uint32_t* h, std::vector<uint8_t>* rgba,
                        std::string* err);

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& m) {
    g_err = m;
    return code;
}

/* no C++ exception may cross the C ABI: every extern "C" entry is a function-try-block */
#define AQ_HOST_CATCH                                                                  \
    catch (const std::bad_alloc&) { return fail(AQ_ERR_OOM, "out of memory"); }        \
    catch (const std::exception& e) { return fail(AQ_ERR_IO, std::string("internal error: ") + e.what()); } \
    catch (...) { return fail(AQ_ERR_IO, "internal error"); }

bool read_file(const std::string& path, std::vector<uint8_t>* out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out->resize((size_t)n);
    size_t got = n ? std::fread(out->data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    return got == (size_t)n;
}

/* ------------------------------------------------------------------ tiny JSON DOM */
struct JVal {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    double num = 0;
    bool b = false;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj; /* keeps file order */
    const JVal* get(const char* k) const {
        for (auto& kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

struct JParser {
    const char* p;
    const char* e; /* *e == 0: the buffer is NUL-terminated so strtod / strncmp cannot run past it */
    std::string err;
    int depth = 0;
    static constexpr int kMaxDepth = 64;
    void ws() {
        while (p < e && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p;
    }
    bool parse(JVal* v) {
        if (depth >= kMaxDepth) return bad("nesting too deep");
        ++depth;
        const bool ok = parse_inner(v);
        --depth;
        return ok;
    }
    bool parse_inner(JVal* v) {
        ws();
        if (p >= e) return bad("unexpected end");
        char c = *p;
        if (c == '{') {
            ++p;
            v->kind = JVal::Obj;
            ws();
            if (p < e && *p == '}') {
                ++p;
                return true;
            }
            for (;;) {
                ws();
                std::string k;
                if (!str(&k)) return false;
                ws();
                if (p >= e || *p != ':') return bad("expected ':'");
                ++p;
                JVal c2;
                if (!parse(&c2)) return false;
                v->obj.emplace_back(std::move(k), std::move(c2));
                ws();
                if (p < e && *p == ',') {
                    ++p;
                    continue;
                }
                if (p < e && *p == '}') {
                    ++p;
                    return true;
                }
                return bad("expected ',' or '}'");
            }
        }
        if (c == '[') {
            ++p;
            v->kind = JVal::Arr;
            ws();
            if (p < e && *p == ']') {
                ++p;
                return true;
            }
            for (;;) {
                JVal c2;
                if (!parse(&c2)) return false;
                v->arr.push_back(std::move(c2));
                ws();
                if (p < e && *p == ',') {
                    ++p;
                    continue;
                }
                if (p < e && *p == ']') {
                    ++p;
                    return true;
                }
                return bad("expected ',' or ']'");
            }
        }
        if (c == '"') {
            v->kind = JVal::Str;
            return str(&v->str);
        }
        if (e - p >= 4 && !std::strncmp(p, "true", 4)) {
            v->kind = JVal::Bool;
            v->b = true;
            p += 4;
            return true;
        }
        if (e - p >= 5 && !std::strncmp(p, "false", 5)) {
            v->kind = JVal::Bool;
            v->b = false;
            p += 5;
            return true;
        }
        if (e - p >= 4 && !std::strncmp(p, "null", 4)) {
            p += 4;
            return true;
        }
        char* end = nullptr;
        double d = std::strtod(p, &end);
        if (end == p || end > e) return bad("bad token");
        v->kind = JVal::Num;
        v->num = d;
        p = end;
        return true;
    }
    bool str(std::string* out) {
        if (p >= e || *p != '"') return bad("expected string");
        ++p;
        out->clear();
        while (p < e && *p != '"') {
            if (*p == '\\' && p + 1 < e) {
                ++p;
                switch (*p) {
                    case 'n': out->push_back('\n'); break;
                    case 't': out->push_back('\t'); break;
                    case 'r': out->push_back('\r'); break;
                    case 'b': out->push_back('\b'); break;
                    case 'f': out->push_back('\f'); break;
                    case 'u': {
                        unsigned cp = 0;
                        for (int i = 0; i < 4 && p + 1 < e; ++i) {
                            ++p;
                            cp = cp * 16 + (unsigned)(std::isdigit((unsigned char)*p) ? *p - '0'
                                                      : (std::tolower(*p) - 'a' + 10));
                        }
                        if (cp < 0x80)
                            out->push_back((char)cp);
                        else if (cp < 0x800) {
                            out->push_back((char)(0xC0 | (cp >> 6)));
                            out->push_back((char)(0x80 | (cp & 0x3F)));
                        } else {
                            out->push_back((char)(0xE0 | (cp >> 12)));
                            out->push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
                            out->push_back((char)(0x80 | (cp & 0x3F)));
                        }
                        break;
                    }
                    default: out->push_back(*p); break; /* \" \\ \/ */
                }
                ++p;
            } else {
                out->push_back(*p++);
            }
        }
        if (p >= e) return bad("unterminated string");
        ++p;
        return true;
    }
    bool bad(const char* m) {
        err = m;
        return false;
    }
};

bool parse_json_file(const std::string& path, JVal* root, std::string* err) {
    std::vector<uint8_t> buf;
    if (!read_file(path, &buf)) {
        *err = "cannot read " + path;
        return false;
    }
    buf.push_back(0); /* NUL-terminate for strtod */
    JParser jp{(const char*)buf.data(), (const char*)buf.data() + buf.size() - 1, {}};
    if (!jp.parse(root)) {
        *err = "JSON parse error in " + path + ": " + jp.err;
        return false;
    }
    return true;
}

/* ------------------------------------------------------------------ BSON TriangleMesh */
struct Mesh {
    std::string name;
    std::vector<float> pos, nrm, uv;
    std::vector<uint32_t> idx;
};

struct Bson {
    const uint8_t* d;
    size_t n;
    std::string err;
    static int32_t i32(const uint8_t* p) {
        int32_t v;
        std::memcpy(&v, p, 4);
        return v;
    }
    /* iterate the elements of the document at [off, off+len) */
    template <class F>
    bool each(size_t off, F&& f) {
        if (off > n || n - off < 5) return bad("truncated document");
        int32_t len = i32(d + off);
        if (len < 5 || (size_t)len > n - off) return bad("bad document length");
        size_t p = off + 4, end = off + (size_t)len - 1;
        if (d[end] != 0) return bad("missing terminator");
        while (p < end) {
            uint8_t type = d[p++];
            const char* key = (const char*)(d + p);
            size_t kl = strnlen(key, end - p);
            if (kl >= end - p) return bad("unterminated element name");
            p += kl + 1;
            size_t vsize;
            switch (type) {
                case 0x01: vsize = 8; break;
                case 0x12: vsize = 8; break;
                case 0x10: vsize = 4; break;
                case 0x08: vsize = 1; break;
                case 0x0A: vsize = 0; break;
                case 0x02: {
                    if (end - p < 4) return bad("truncated string");
                    const int32_t sl = i32(d + p); /* bytes incl. the NUL: at least 1 */
                    if (sl < 1) return bad("bad string length");
                    vsize = 4 + (size_t)sl;
                    break;
                }
                case 0x03:
                case 0x04: {
                    if (end - p < 4) return bad("truncated subdocument");
                    const int32_t dl = i32(d + p); /* a document is at least length + terminator */
                    if (dl < 5) return bad("bad subdocument length");
                    vsize = (size_t)dl;
                    break;
                }
                default: return bad("unsupported BSON element type");
            }
            if (vsize > end - p) return bad("element overruns document");
            if (type == 0x02 && d[p + vsize - 1] != 0) return bad("string is not NUL-terminated");
            if (!f(type, key, p, vsize)) return false;
            p += vsize;
        }
        return true;
    }
    bool bad(const char* m) {
        err = m;
        return false;
    }
    bool number(uint8_t type, size_t off, double* out) {
        if (type == 0x01) {
            std::memcpy(out, d + off, 8);
            return true;
        }
        if (type == 0x12) {
            int64_t v;
            std::memcpy(&v, d + off, 8);
            *out = (double)v;
            return true;
        }
        if (type == 0x10) {
            *out = (double)i32(d + off);
            return true;
        }
        return bad("expected a number");
    }
};

template <class T>
bool bson_array_of_tuples(Bson& b, size_t off, int arity, std::vector<T>* out) {
    return b.each(off, [&](uint8_t type, const char*, size_t voff, size_t) {
        if (type != 0x04) return b.bad("expected array of arrays");
        int k = 0;
        bool ok = b.each(voff, [&](uint8_t t2, const char*, size_t v2, size_t) {
            double x;
            if (!b.number(t2, v2, &x)) return false;
            out->push_back((T)x);
            ++k;
            return true;
        });
        if (!ok) return false;
        if (k != arity) return b.bad("wrong tuple arity");
        return true;
    });
}

bool load_mesh(const std::string& path, Mesh* m, std::string* err, bool* missing) {
    std::vector<uint8_t> buf;
    *missing = false;
    if (!read_file(path, &buf)) {
        *missing = true;
        *err = "cannot read " + path;
        return false;
    }
    Bson b{buf.data(), buf.size(), {}};
    if (buf.size() < 5 || (size_t)Bson::i32(buf.data()) != buf.size()) {
        *err = path + ": BSON length does not match file size";
        return false;
    }
    bool ok = b.each(0, [&](uint8_t type, const char* key, size_t off, size_t vsize) {
        if (!std::strcmp(key, "name") && type == 0x02) {
            m->name.assign((const char*)buf.data() + off + 4, vsize - 5);
            return true;
        }
        if (type != 0x04) return true; /* ignore unknown scalars */
        if (!std::strcmp(key, "vertices")) return bson_array_of_tuples(b, off, 3, &m->pos);
        if (!std::strcmp(key, "normals")) return bson_array_of_tuples(b, off, 3, &m->nrm);
        if (!std::strcmp(key, "texcoords")) return bson_array_of_tuples(b, off, 2, &m->uv);
        if (!std::strcmp(key, "indices")) return bson_array_of_tuples(b, off, 3, &m->idx);
        return true;
    });
    if (!ok) {
        *err = path + ": " + b.err;
        return false;
    }
    size_t nv = m->pos.size() / 3;
    if (!m->nrm.empty() && m->nrm.size() != m->pos.size()) {
        *err = path + ": normals/vertices count mismatch";
        return false;
    }
    if (!m->uv.empty() && m->uv.size() != nv * 2) {
        *err = path + ": texcoords/vertices count mismatch";
        return false;
    }
    for (uint32_t i : m->idx)
        if (i >= nv) {
            *err = path + ": index out of range";
            return false;
        }
    return true;
}

/* ------------------------------------------------------------------ Texture enum */
double srgb_to_linear(double c) {
    return c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4);
}

struct TexVal {
    float v[3] = {0, 0, 0};
    std::string image; /* non-empty => Texture::Image */
};

bool parse_texture(const JVal* t, TexVal* out, std::string* err) {
    if (!t || t->kind != JVal::Obj || t->obj.size() != 1) {
        *err = "Texture must be an externally tagged enum";
        return false;
    }
    const std::string& tag = t->obj[0].first;
    const JVal& v = t->obj[0].second;
    if (tag == "Float" && v.kind == JVal::Num) {
        out->v[0] = out->v[1] = out->v[2] = (float)v.num;
        return true;
    }
    if ((tag == "Float3" || tag == "Srgb") && v.kind == JVal::Arr && v.arr.size() == 3) {
        for (int i = 0; i < 3; ++i) {
            double c = v.arr[i].num;
            out->v[i] = (float)(tag == "Srgb" ? srgb_to_linear((double)(float)c) : c);
        }
        return true;
    }
    if (tag == "Image" && v.kind == JVal::Str) {
        out->image = v.str;
        for (char& c : out->image)
            if (c == '\\') c = '/';
        return true;
    }
    *err = "unknown Texture variant '" + tag + "'";
    return false;
}

std::string dir_of(const std::string& p) {
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string(".") : p.substr(0, k);
}

bool vec3(const JVal* v, float* out) {
    if (!v || v->kind != JVal::Arr || v->arr.size() != 3) return false;
    for (int i = 0; i < 3; ++i) out[i] = (float)v->arr[i].num;
    return true;
}

}  // namespace

struct aq_host_scene {
    aq_scene_desc desc{};
    aq_host_scene_info info{};
    std::vector<float> pos, nrm, uv;
    std::vector<uint32_t> idx, tri_mat;
    std::vector<aq_material> mats;
    std::vector<std::string> mat_names;
    std::vector<aq_texture> texs;
    std::vector<std::vector<uint8_t>> tex_data;
    std::vector<aq_point_light> lights;
    struct Range {
        uint32_t first, count, mat;
    };
    std::vector<Range> shapes;
};

extern "C" {

const char* aq_host_last_error(void) { return g_err.c_str(); }
float aq_host_srgb_to_linear(float c) { return (float)srgb_to_linear((double)c); }
void aq_host_free(void* p) { std::free(p); }

int aq_host_jpeg_decode(const char* path, uint32_t* width, uint32_t* height, uint8_t** rgba8) try {
    if (!path || !width || !height || !rgba8) return fail(AQ_ERR_BAD_ARG, "null argument");
    std::vector<uint8_t> px;
    std::string err;
    int rc = aq_jpeg_decode_file(path, width, height, &px, &err);
    if (rc != 0) return fail(rc, err);
    *rgba8 = (uint8_t*)std::malloc(px.size());
    if (!*rgba8) return fail(AQ_ERR_OOM, "out of memory");
    std::memcpy(*rgba8, px.data(), px.size());
    return AQ_OK;
} AQ_HOST_CATCH

int aq_host_mesh_load(const char* path, char* name_out, size_t name_cap, uint32_t* n_verts,
                      uint32_t* n_tris, float** positions, float** normals, float** uvs,
                      uint32_t* n_uvs, uint32_t** indices) try {
    if (!path) return fail(AQ_ERR_BAD_ARG, "null path");
    Mesh m;
    std::string err;
    bool missing;
    if (!load_mesh(path, &m, &err, &missing)) return fail(AQ_ERR_IO, err);
    if (name_out && name_cap) std::snprintf(name_out, name_cap, "%s", m.name.c_str());
    if (n_verts) *n_verts = (uint32_t)(m.pos.size() / 3);
    if (n_tris) *n_tris = (uint32_t)(m.idx.size() / 3);
    if (n_uvs) *n_uvs = (uint32_t)(m.uv.size() / 2);
    auto dup = [](const void* src, size_t bytes) -> void* {
        void* p = std::malloc(bytes ? bytes : 1);
        if (p && bytes) std::memcpy(p, src, bytes);
        return p;
    };
    /* a mesh without normals yields n_verts zero vectors (the shader then uses the geometric
     * normal), exactly as aq_host_scene_load does: the caller may always read n_verts * 3 floats */
    if (m.nrm.empty()) m.nrm.assign(m.pos.size(), 0.0f);
    if (positions) *positions = (float*)dup(m.pos.data(), m.pos.size() * 4);
    if (normals) *normals = (float*)dup(m.nrm.data(), m.nrm.size() * 4);
    if (uvs) *uvs = (float*)dup(m.uv.data(), m.uv.size() * 4);
    if (indices) *indices = (uint32_t*)dup(m.idx.data(), m.idx.size() * 4);
    return AQ_OK;
} AQ_HOST_CATCH

int aq_host_scene_load(const char* json_path, aq_host_scene** out) try {
    if (!json_path || !out) return fail(AQ_ERR_BAD_ARG, "null argument");
    JVal root;
    std::string err;
    if (!parse_json_file(json_path, &root, &err)) return fail(AQ_ERR_IO, err);
    std::string base = dir_of(json_path);
    std::unique_ptr<aq_host_scene> S(new aq_host_scene);

    /* ---- named_bsdfs (scenes/cbox.json:2-515) */
    const JVal* bs = root.get("named_bsdfs");
    if (!bs || bs->kind != JVal::Obj) return fail(AQ_ERR_IO, "scene: missing named_bsdfs");
    std::map<std::string, uint32_t> mat_index, tex_index;
    for (auto& kv : bs->obj) {
        const JVal* pr = kv.second.get("Principled");
        if (!pr) return fail(AQ_ERR_UNSUPPORTED, "bsdf '" + kv.first + "': only Principled is supported");
        aq_material m;
        std::memset(&m, 0, sizeof m);
        m.color_tex = -1;
        struct F {
            const char* key;
            float* dst;
            int n;
        } fields[] = {{"metallic", &m.metallic, 1},
                      {"roughness", &m.roughness, 1},
                      {"specular", &m.specular, 1},
                      {"specular_tint", &m.specular_tint, 1},
                      {"sheen", &m.sheen, 1},
                      {"sheen_tint", &m.sheen_tint, 1},
                      {"clearcoat", &m.clearcoat, 1},
                      {"clearcoat_roughness", &m.clearcoat_roughness, 1},
                      {"ior", &m.ior, 1},
                      {"transmission", &m.transmission, 1},
                      {"subsurface", &m.subsurface, 1},
                      {"anisotropic", &m.anisotropic, 1},
                      {"anisotropic_rotation", &m.anisotropic_rotation, 1},
                      {"emission", m.emission, 3},
                      {"subsurface_color", m.subsurface_color, 3},
                      {"subsurface_radius", m.subsurface_radius, 3}};
        TexVal tv;
        if (!parse_texture(pr->get("color"), &tv, &err))
            return fail(AQ_ERR_IO, "bsdf '" + kv.first + "'.color: " + err);
        /* decode an image once per scene; returns its texture index */
        auto texture_of = [&](const std::string& image, uint32_t* id_out, std::string* why) -> int {
            auto it = tex_index.find(image);
            if (it == tex_index.end()) {
                uint32_t w = 0, h = 0;
                std::vector<uint8_t> px;
                std::string path = base + "/" + image;
                int rc = aq_jpeg_decode_file(path.c_str(), &w, &h, &px, why);
                if (rc != 0) {
                    *why = "texture " + path + ": " + *why;
                    return rc;
                }
                uint32_t id = (uint32_t)S->tex_data.size();
                S->tex_data.push_back(std::move(px));
                aq_texture t;
                t.width = w;
                t.height = h;
                t.rgba8 = nullptr;
                S->texs.push_back(t);
                it = tex_index.emplace(image, id).first;
            }
            *id_out = it->second;
            return 0;
        };
        if (!tv.image.empty()) {
            uint32_t id = 0;
            int rc = texture_of(tv.image, &id, &err);
            if (rc != 0) return fail(rc, err);
            m.color_tex = (int32_t)id;
            m.color[0] = m.color[1] = m.color[2] = 1.0f;
        } else {
            std::memcpy(m.color, tv.v, sizeof m.color);
        }
        for (auto& f : fields) {
            const JVal* v = pr->get(f.key);
            if (!v) continue; /* serde default would be an error; be lenient */
            TexVal t2;
            if (!parse_texture(v, &t2, &err))
                return fail(AQ_ERR_IO, "bsdf '" + kv.first + "'." + f.key + ": " + err);
            if (!t2.image.empty()) { /* Texture::Image on a scalar / colour parameter: constant 1 x texel */
                static const struct { const char* key; int slot; } kSlots[] = {
                    {"metallic", AQ_PTEX_METALLIC}, {"roughness", AQ_PTEX_ROUGHNESS}, {"specular", AQ_PTEX_SPECULAR},
                    {"specular_tint", AQ_PTEX_SPECULAR_TINT}, {"sheen", AQ_PTEX_SHEEN}, {"sheen_tint", AQ_PTEX_SHEEN_TINT},
                    {"transmission", AQ_PTEX_TRANSMISSION}, {"clearcoat", AQ_PTEX_CLEARCOAT},
                    {"clearcoat_roughness", AQ_PTEX_CLEARCOAT_ROUGHNESS}, {"ior", AQ_PTEX_IOR},
                    {"subsurface", AQ_PTEX_SUBSURFACE}, {"subsurface_color", AQ_PTEX_SUBSURFACE_COLOR}};
                int slot = -1;
                for (auto& ks : kSlots)
                    if (!std::strcmp(ks.key, f.key)) slot = ks.slot;
                if (slot < 0)
                    return fail(AQ_ERR_UNSUPPORTED, "bsdf '" + kv.first + "'." + f.key +
                                                        ": Image is not supported on this parameter (emission, anisotropic*, subsurface_radius)");
                uint32_t id = 0;
                int rc = texture_of(t2.image, &id, &err);
                if (rc != 0) return fail(rc, err);
                if (id >= 255u) return fail(AQ_ERR_UNSUPPORTED, "more than 255 textures referenced by non-colour parameters");
                m.param_tex[slot] = (uint8_t)(id + 1u);
                for (int i = 0; i < f.n; ++i) f.dst[i] = 1.0f;
                continue;
            }
            for (int i = 0; i < f.n; ++i) f.dst[i] = t2.v[i];
        }
        mat_index[kv.first] = (uint32_t)S->mats.size();
        S->mats.push_back(m);
        S->mat_names.push_back(kv.first);
    }
    for (size_t i = 0; i < S->texs.size(); ++i) S->texs[i].rgba8 = S->tex_data[i].data();

    /* ---- camera (scenes/cbox.json:516-543) */
    const JVal* cam = root.get("camera");
    const JVal* per = cam ? cam->get("Perspective") : nullptr;
    if (!per) return fail(AQ_ERR_UNSUPPORTED, "camera: only Perspective is supported");
    aq_camera& C = S->desc.camera;
    const JVal* res = per->get("res");
    if (!res || res->kind != JVal::Arr || res->arr.size() != 2) return fail(AQ_ERR_IO, "camera.res");
    C.res[0] = (uint32_t)res->arr[0].num;
    C.res[1] = (uint32_t)res->arr[1].num;
    C.fov = per->get("fov") ? (float)per->get("fov")->num : 45.f;
    C.lens_radius = per->get("lens_radius") ? (float)per->get("lens_radius")->num : 0.f;
    C.focal = per->get("focal") ? (float)per->get("focal")->num : 1.f;
    const JVal* tr = per->get("transform");
    C.scale[0] = C.scale[1] = C.scale[2] = 1.f;
    if (tr) {
        vec3(tr->get("translate"), C.translate);
        vec3(tr->get("rotate"), C.rotate);
        vec3(tr->get("scale"), C.scale);
    }

    /* ---- lights (scenes/cbox.json:544-561) */
    const JVal* ls = root.get("lights");
    if (ls && ls->kind == JVal::Arr)
        for (auto& l : ls->arr) {
            const JVal* pt = l.get("Point");
            if (!pt) return fail(AQ_ERR_UNSUPPORTED, "light: only Point is supported");
            aq_point_light pl;
            if (!vec3(pt->get("pos"), pl.pos)) return fail(AQ_ERR_IO, "light.pos");
            TexVal tv;
            if (!parse_texture(pt->get("emission"), &tv, &err) || !tv.image.empty())
                return fail(AQ_ERR_IO, "light.emission: " + err);
            std::memcpy(pl.intensity, tv.v, sizeof pl.intensity);
            S->lights.push_back(pl);
        }

    /* ---- shapes (scenes/cbox.json:562-627): Mesh[path, Named(bsdf)] */
    const JVal* sh = root.get("shapes");
    if (!sh || sh->kind != JVal::Arr) return fail(AQ_ERR_IO, "scene: missing shapes");
    bool any_uv = false;
    float bmin[3] = {INFINITY, INFINITY, INFINITY}, bmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (auto& s : sh->arr) {
        const JVal* me = s.get("Mesh");
        if (!me || me->kind != JVal::Arr || me->arr.size() != 2 || me->arr[0].kind != JVal::Str)
            return fail(AQ_ERR_UNSUPPORTED, "shape: only Mesh[path, bsdf] is supported");
        const JVal* named = me->arr[1].get("Named");
        if (!named || named->kind != JVal::Str)
            return fail(AQ_ERR_UNSUPPORTED, "shape bsdf: only Named is supported");
        auto mi = mat_index.find(named->str);
        if (mi == mat_index.end()) return fail(AQ_ERR_IO, "unresolved bsdf '" + named->str + "'");
        S->info.n_shapes++;
        std::string mp = me->arr[0].str;
        for (char& c : mp)
            if (c == '\\') c = '/';
        Mesh m;
        bool missing = false;
        if (!load_mesh(base + "/" + mp, &m, &err, &missing)) {
            if (missing) { /* tolerated: .MISSING_LARGE_BLOBS:1 */
                S->info.n_meshes_missing++;
                S->shapes.push_back({(uint32_t)(S->idx.size() / 3), 0u, mi->second});
                std::fprintf(stderr, "[aqua_host] warning: %s (skipped)\n", err.c_str());
                continue;
            }
            return fail(AQ_ERR_IO, err);
        }
        S->info.n_meshes_loaded++;
        uint32_t vbase = (uint32_t)(S->pos.size() / 3), nv = (uint32_t)(m.pos.size() / 3);
        uint32_t tbase = (uint32_t)(S->idx.size() / 3), nt = (uint32_t)(m.idx.size() / 3);
        S->pos.insert(S->pos.end(), m.pos.begin(), m.pos.end());
        if (m.nrm.empty()) m.nrm.assign(m.pos.size(), 0.f); /* (0,0,0) => geometric normal */
        S->nrm.insert(S->nrm.end(), m.nrm.begin(), m.nrm.end());
        if (m.uv.empty()) {
            m.uv.assign((size_t)nv * 2, 0.f);
        } else {
            any_uv = true;
        }
        S->uv.insert(S->uv.end(), m.uv.begin(), m.uv.end());
        for (uint32_t i : m.idx) S->idx.push_back(vbase + i);
        S->tri_mat.insert(S->tri_mat.end(), nt, mi->second);
        S->shapes.push_back({tbase, nt, mi->second});
        for (size_t i = 0; i < m.pos.size(); ++i) {
            bmin[i % 3] = std::fmin(bmin[i % 3], m.pos[i]);
            bmax[i % 3] = std::fmax(bmax[i % 3], m.pos[i]);
        }
    }

    aq_scene_desc& D = S->desc;
    D.n_verts = (uint32_t)(S->pos.size() / 3);
    D.n_tris = (uint32_t)(S->idx.size() / 3);
    D.positions = S->pos.data();
    D.normals = S->nrm.data();
    D.uvs = any_uv ? S->uv.data() : nullptr;
    D.indices = S->idx.data();
    D.tri_material = S->tri_mat.data();
    D.n_materials = (uint32_t)S->mats.size();
    D.materials = S->mats.data();
    D.n_textures = (uint32_t)S->texs.size();
    D.textures = S->texs.data();
    D.n_lights = (uint32_t)S->lights.size();
    D.lights = S->lights.data();
    S->info.n_verts = D.n_verts;
    S->info.n_tris = D.n_tris;
    S->info.n_materials = D.n_materials;
    S->info.n_textures = D.n_textures;
    S->info.n_lights = D.n_lights;
    std::memcpy(S->info.bounds_min, bmin, sizeof bmin);
    std::memcpy(S->info.bounds_max, bmax, sizeof bmax);
    *out = S.release();
    return AQ_OK;
} AQ_HOST_CATCH

void aq_host_scene_free(aq_host_scene* s) { delete s; }
const aq_scene_desc* aq_host_scene_desc(const aq_host_scene* s) { return s ? &s->desc : nullptr; }
int aq_host_scene_get_info(const aq_host_scene* s, aq_host_scene_info* info) try {
    if (!s || !info) return fail(AQ_ERR_BAD_ARG, "null argument");
    *info = s->info;
    return AQ_OK;
} AQ_HOST_CATCH
const char* aq_host_material_name(const aq_host_scene* s, uint32_t i) {
    return (s && i < s->mat_names.size()) ? s->mat_names[i].c_str() : nullptr;
}
int aq_host_shape_range(const aq_host_scene* s, uint32_t shape, uint32_t* first_tri,
                        uint32_t* n_tris, uint32_t* material) try {
    if (!s || shape >= s->shapes.size()) return fail(AQ_ERR_BAD_ARG, "shape index out of range");
    if (first_tri) *first_tri = s->shapes[shape].first;
    if (n_tris) *n_tris = s->shapes[shape].count;
    if (material) *material = s->shapes[shape].mat;
    return AQ_OK;
} AQ_HOST_CATCH

int aq_host_integrator_load(const char* json_path, aq_integrator_cfg* cfg, char* type_out,
                            size_t type_cap) try {
    if (!json_path || !cfg) return fail(AQ_ERR_BAD_ARG, "null argument");
    JVal root;
    std::string err;
    if (!parse_json_file(json_path, &root, &err)) return fail(AQ_ERR_IO, err);
    if (root.kind != JVal::Obj) return fail(AQ_ERR_IO, "integrator: expected an object");
    const JVal* ty = root.get("type");
    std::string t = (ty && ty->kind == JVal::Str) ? ty->str : "pt";
    /* "nrc" (integrator.json:2): its batch_size/training_iters/learning_rate/visualize_cache
     * keys (:4,:6-8) are read by aq_host_integrator_load_nrc */
    if (t != "nrc" && t != "pt" && t != "path")
        return fail(AQ_ERR_UNSUPPORTED, "integrator type '" + t + "' is not supported");
    std::memset(cfg, 0, sizeof *cfg);
    const JVal* spp = root.get("spp");
    const JVal* md = root.get("max_depth");
    cfg->spp_begin = 0;
    cfg->spp_end = spp ? (uint32_t)spp->num : 16u;
    cfg->max_depth = md ? (uint32_t)md->num : 5u;
    const JVal* seed = root.get("seed");
    cfg->seed = seed ? (uint32_t)seed->num : 0u;
    if (type_out && type_cap) std::snprintf(type_out, type_cap, "%s", t.c_str());
    return AQ_OK;
} AQ_HOST_CATCH

/* the NRC-only keys of scenes/integrator.json (:4 batch_size, :6 training_iters,
 * :7 learning_rate, :8 visualize_cache); absent keys keep the reference file's values */
int aq_host_integrator_load_nrc(const char* json_path, aq_nrc_cfg* nrc) try {
    if (!json_path || !nrc) return fail(AQ_ERR_BAD_ARG, "null argument");
    JVal root;
    std::string err;
    if (!parse_json_file(json_path, &root, &err)) return fail(AQ_ERR_IO, err);
    if (root.kind != JVal::Obj) return fail(AQ_ERR_IO, "integrator: expected an object");
    const JVal* b = root.get("batch_size");
    const JVal* it = root.get("training_iters");
    const JVal* lr = root.get("learning_rate");
    const JVal* vc = root.get("visualize_cache");
    nrc->batch_size = b ? (uint32_t)b->num : 512u;
    nrc->training_iters = it ? (uint32_t)it->num : 2048u;
    nrc->learning_rate = lr ? (float)lr->num : 1.0e-3f;
    nrc->visualize_cache = (vc && vc->kind == JVal::Bool && vc->b) ? 1u : 0u;
    return AQ_OK;
} AQ_HOST_CATCH

int aq_host_write_ppm(const char* path, const float* film, uint32_t width, uint32_t height) try {
    if (!path || !film) return fail(AQ_ERR_BAD_ARG, "null argument");
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(AQ_ERR_IO, std::string("cannot write ") + path);
    std::fprintf(f, "P6\n%u %u\n255\n", width, height);
    std::vector<uint8_t> row((size_t)width * 3);
    for (uint32_t y = 0; y < height; ++y) {
        for (uint32_t x = 0; x < width; ++x) {
            const float* p = film + 4 * ((size_t)y * width + x);
            float w = p[3] > 0.f ? 1.f / p[3] : 0.f;
            for (int c = 0; c < 3; ++c) {
                float v = p[c] * w;
                v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
                float s = v <= 0.0031308f ? 12.92f * v : 1.055f * std::pow(v, 1.f / 2.4f) - 0.055f;
                row[3 * (size_t)x + c] = (uint8_t)std::lround(s * 255.f);
            }
        }
        std::fwrite(row.data(), 1, row.size(), f);
    }
    std::fclose(f);
    return AQ_OK;
} AQ_HOST_CATCH

/* ---- PNG: 8-bit RGBA, filter 0, zlib stream of stored (uncompressed) deflate blocks */
static uint32_t crc32_update(uint32_t c, const uint8_t* d, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t v = i;
            for (int k = 0; k < 8; ++k) v = (v & 1) ? 0xEDB88320u ^ (v >> 1) : v >> 1;
            table[i] = v;
        }
        init = true;
    }
    for (size_t i = 0; i < n; ++i) c = table[(c ^ d[i]) & 0xFF] ^ (c >> 8);
    return c;
}
static void png_chunk(FILE* f, const char* type, const std::vector<uint8_t>& data) {
    uint8_t len[4] = {(uint8_t)(data.size() >> 24), (uint8_t)(data.size() >> 16), (uint8_t)(data.size() >> 8), (uint8_t)data.size()};
    std::fwrite(len, 1, 4, f);
    std::fwrite(type, 1, 4, f);
    if (!data.empty()) std::fwrite(data.data(), 1, data.size(), f);
    uint32_t c = crc32_update(0xFFFFFFFFu, (const uint8_t*)type, 4);
    c = crc32_update(c, data.data(), data.size()) ^ 0xFFFFFFFFu;
    uint8_t crc[4] = {(uint8_t)(c >> 24), (uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c};
    std::fwrite(crc, 1, 4, f);
}

int aq_host_write_png(const char* path, const uint8_t* rgba8, uint32_t width, uint32_t height) try {
    if (!path || !rgba8 || !width || !height) return fail(AQ_ERR_BAD_ARG, "aq_host_write_png: bad argument");
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(AQ_ERR_IO, std::string("cannot write ") + path);
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::fwrite(sig, 1, 8, f);
    std::vector<uint8_t> ihdr = {(uint8_t)(width >> 24), (uint8_t)(width >> 16), (uint8_t)(width >> 8), (uint8_t)width,
                                 (uint8_t)(height >> 24), (uint8_t)(height >> 16), (uint8_t)(height >> 8), (uint8_t)height,
                                 8, 6, 0, 0, 0};
    png_chunk(f, "IHDR", ihdr);
    /* raw scanlines: filter byte 0 + RGBA */
    const size_t row = (size_t)width * 4 + 1, raw_n = row * height;
    std::vector<uint8_t> raw(raw_n);
    for (uint32_t y = 0; y < height; ++y) {
        raw[y * row] = 0;
        std::memcpy(&raw[y * row + 1], rgba8 + (size_t)y * width * 4, (size_t)width * 4);
    }
    std::vector<uint8_t> z;
    z.reserve(raw_n + raw_n / 65535 * 5 + 16);
    z.push_back(0x78);
    z.push_back(0x01);
    uint32_t a = 1, b = 0; /* adler32 */
    for (size_t off = 0; off < raw_n;) {
        size_t n = std::min<size_t>(65535, raw_n - off);
        z.push_back(off + n == raw_n ? 1 : 0);
        z.push_back((uint8_t)(n & 0xFF));
        z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)(~n & 0xFF));
        z.push_back((uint8_t)((~n >> 8) & 0xFF));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        for (size_t i = 0; i < n; ++i) {
            a = (a + raw[off + i]) % 65521u;
            b = (b + a) % 65521u;
        }
        off += n;
    }
    uint32_t ad = (b << 16) | a;
    z.push_back((uint8_t)(ad >> 24));
    z.push_back((uint8_t)(ad >> 16));
    z.push_back((uint8_t)(ad >> 8));
    z.push_back((uint8_t)ad);
    png_chunk(f, "IDAT", z);
    png_chunk(f, "IEND", {});
    std::fclose(f);
    return AQ_OK;
} AQ_HOST_CATCH

int aq_host_write_pfm(const char* path, const float* film, uint32_t width, uint32_t height) try {
    if (!path || !film) return fail(AQ_ERR_BAD_ARG, "null argument");
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(AQ_ERR_IO, std::string("cannot write ") + path);
    std::fprintf(f, "PF\n%u %u\n-1.0\n", width, height);
    std::vector<float> row((size_t)width * 3);
    for (uint32_t y = 0; y < height; ++y) { /* PFM stores the bottom row first */
        const float* src = film + 4 * (size_t)(height - 1 - y) * width;
        for (uint32_t x = 0; x < width; ++x) {
            float w = src[4 * x + 3] > 0.f ? 1.f / src[4 * x + 3] : 0.f;
            for (int c = 0; c < 3; ++c) row[3 * (size_t)x + c] = src[4 * x + c] * w;
        }
        std::fwrite(row.data(), 4, row.size(), f);
    }
    std::fclose(f);
    return AQ_OK;
} AQ_HOST_CATCH

void aq_host_srgb_thresholds(float* t255) {
    /* level k+1 is chosen when round(255*oetf(v)) >= k+1, i.e. oetf(v) >= (k+0.5)/255 */
    for (int k = 0; k < 255; ++k) {
        double s = ((double)k + 0.5) / 255.0;
        double lin = s <= 0.04045 ? s / 12.92 : std::pow((s + 0.055) / 1.055, 2.4);
        t255[k] = (float)lin;
    }
}

}  // extern "C"
