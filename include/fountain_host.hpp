/*
 * fountain_host.hpp -- C++17 host side above the C ABI (include/fountain_gpu.h).
 *
 * The reference is a compiled (Rust) crate whose toolchain is not available here, so the
 * host-side mirror of its operator interface for this path is written in C++: the same
 * names, argument meaning and error behaviour as the Rust types a maintainer would keep
 * (Scene, TriangleMesh, Sphere, *Material, InfiniteAreaLight, DiffuseAreaLight,
 * PerspectiveCamera, Film, RandomSampler, PathIntegrator, DirectLightingIntegrator,
 * SamplerIntegrator::render_parallel).  fountain_b200/api.py is the Python twin of this file.
 *
 * Header-only.  The device library is bound at run time (dlopen), so the same host code can
 * be pointed at any library exporting the ABI under a symbol prefix: the product binds
 * `libfountain_gpu.so` / "ftn_"; the test-suite may bind a checker library instead.  Nothing
 * in here computes an intersection or a radiance value: without a device library there is no
 * result (Library::open throws), never a CPU fallback.
 *
 * Error behaviour: every non-zero ABI status becomes a fountain::Error carrying the status
 * code and ftn_last_error(); NaN radiance (the reference's check_radiance panic,
 * integrator/mod.rs:285) is FTN_ERR_NAN_RADIANCE.
 */
#ifndef FOUNTAIN_HOST_HPP
#define FOUNTAIN_HOST_HPP

#include <dlfcn.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "fountain_gpu.h"

namespace fountain {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// ------------------------------------------------------------------------------------------
// Library: the ABI bound at run time
// ------------------------------------------------------------------------------------------
class Library {
public:
    // path: shared object; prefix: symbol prefix of the ABI functions ("ftn_" for the product)
    static std::shared_ptr<Library> open(const std::string& path, const std::string& prefix = "ftn_") {
        return std::shared_ptr<Library>(new Library(path, prefix));
    }
    // $FTN_GPU_LIB, else libfountain_gpu.so through the loader's search path
    static std::shared_ptr<Library> open_gpu() {
        const char* p = std::getenv("FTN_GPU_LIB");
        auto lib = open(p ? p : "libfountain_gpu.so", "ftn_");
        int n = 0;
        lib->check(lib->device_count(&n), "ftn_device_count");
        if (n <= 0) throw Error(FTN_ERR_NO_DEVICE, "fountain: no CUDA device (there is no CPU fallback)");
        return lib;
    }
    ~Library() { if (h_) dlclose(h_); }
    Library(const Library&) = delete;
    Library& operator=(const Library&) = delete;

    void check(int status, const char* what) const {
        if (status == 0) return;
        const char* msg = last_error ? last_error() : "";
        throw Error(status, std::string(what) + " failed (" + std::to_string(status) + "): " + (msg ? msg : ""));
    }

    uint32_t (*abi_version)() = nullptr;
    const char* (*last_error)() = nullptr;
    int (*device_count)(int*) = nullptr;
    int (*set_device)(int) = nullptr;
    int (*scene_create)(const FtnSceneDesc*, FtnScene**) = nullptr;
    int (*scene_destroy)(FtnScene*) = nullptr;
    int (*bvh_build)(FtnScene*) = nullptr;
    int (*bvh_debug_morton)(const FtnScene*, uint32_t*, uint32_t*) = nullptr;
    int (*scene_world_bound)(const FtnScene*, float*) = nullptr;
    int (*scene_stats)(const FtnScene*, FtnStats*) = nullptr;
    int (*intersect)(const FtnScene*, size_t, const FtnRay*, FtnHit*) = nullptr;
    int (*intersect_test)(const FtnScene*, size_t, const FtnRay*, uint8_t*) = nullptr;
    int (*render)(const FtnScene*, const FtnCamera*, const FtnFilm*, const FtnSampler*, const FtnIntegrator*,
                  FtnPixel*, FtnStats*) = nullptr;
    int (*film_pixel_count)(const FtnFilm*, int32_t*, int32_t*) = nullptr;

private:
    void* h_ = nullptr;
    template <class F> void bind(F& f, const std::string& prefix, const char* name, bool required = true) {
        void* s = dlsym(h_, (prefix + name).c_str());
        if (!s && required) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: symbol " + prefix + name + " missing");
        f = reinterpret_cast<F>(s);
    }
    Library(const std::string& path, const std::string& prefix) {
        h_ = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!h_) throw Error(FTN_ERR_NO_DEVICE, std::string("fountain: cannot load ") + path + ": " + dlerror());
        bind(abi_version, prefix, "abi_version");
        bind(last_error, prefix, "last_error");
        bind(device_count, prefix, "device_count", false);
        bind(set_device, prefix, "set_device", false);
        bind(scene_create, prefix, "scene_create");
        bind(scene_destroy, prefix, "scene_destroy");
        bind(bvh_build, prefix, "bvh_build");
        bind(bvh_debug_morton, prefix, "bvh_debug_morton");
        bind(scene_world_bound, prefix, "scene_world_bound");
        bind(scene_stats, prefix, "scene_stats");
        bind(intersect, prefix, "intersect");
        bind(intersect_test, prefix, "intersect_test");
        bind(render, prefix, "render");
        bind(film_pixel_count, prefix, "film_pixel_count");
        if (abi_version() != FTN_ABI_VERSION)
            throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: ABI version mismatch");
    }
};
using LibraryPtr = std::shared_ptr<Library>;

// ------------------------------------------------------------------------------------------
// geometry: Vec3f / Transform (geometry/transform.rs:6-160)
// ------------------------------------------------------------------------------------------
struct Vec3f {
    float x = 0, y = 0, z = 0;
    Vec3f() = default;
    Vec3f(float a, float b, float c) : x(a), y(b), z(c) {}
};
using Point3f = Vec3f;

struct Spectrum {                      // spectrum/mod.rs: RGB triple
    float r = 0, g = 0, b = 0;
    Spectrum() = default;
    Spectrum(float v) : r(v), g(v), b(v) {}                       // Spectrum::uniform
    Spectrum(float r_, float g_, float b_) : r(r_), g(g_), b(b_) {}
    std::array<float, 3> into_array() const { return {r, g, b}; }
};

// `Transform { t, invt }`; built in double, rounded to f32 once when it crosses the ABI
class Transform {
public:
    using M = std::array<double, 16>;   // row-major, m[4*r + c], maths convention (M * column vector)
    M m, minv;

    Transform() : m(eye()), minv(eye()) {}
    explicit Transform(const M& a) : m(a), minv(invert(a)) {}
    Transform(const M& a, const M& ai) : m(a), minv(ai) {}

    static Transform identity() { return Transform(); }                                   // transform.rs:18
    static Transform from_flat(const std::array<double, 16>& col_major) {                 // transform.rs:34-42
        M a;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) a[4 * r + c] = col_major[4 * c + r];
        return Transform(a);
    }
    static Transform translate(double x, double y, double z) {                            // transform.rs:62-66
        M a = eye(), b = eye();
        a[3] = x; a[7] = y; a[11] = z;
        b[3] = -x; b[7] = -y; b[11] = -z;
        return Transform(a, b);
    }
    static Transform scale(double sx, double sy, double sz) {                             // transform.rs:68-72
        M a = eye(), b = eye();
        a[0] = sx; a[5] = sy; a[10] = sz;
        b[0] = 1.0 / sx; b[5] = 1.0 / sy; b[10] = 1.0 / sz;
        return Transform(a, b);
    }
    static Transform rotate(double theta_deg, double ax, double ay, double az) {          // transform.rs:74-78
        double n = std::sqrt(ax * ax + ay * ay + az * az);
        double x = ax / n, y = ay / n, z = az / n;
        double th = theta_deg * M_PI / 180.0, s = std::sin(th), c = std::cos(th);
        M a = eye();
        a[0] = c + x * x * (1 - c);     a[1] = x * y * (1 - c) - z * s; a[2] = x * z * (1 - c) + y * s;
        a[4] = y * x * (1 - c) + z * s; a[5] = c + y * y * (1 - c);     a[6] = y * z * (1 - c) - x * s;
        a[8] = z * x * (1 - c) - y * s; a[9] = z * y * (1 - c) + x * s; a[10] = c + z * z * (1 - c);
        return Transform(a, transpose(a));
    }
    // world-to-camera, as the pbrt LookAt directive (transform.rs:44-56)
    static Transform look_at(const std::array<double, 3>& pos, const std::array<double, 3>& look,
                             const std::array<double, 3>& up) {
        auto sub = [](auto a, auto b) { return std::array<double, 3>{a[0] - b[0], a[1] - b[1], a[2] - b[2]}; };
        auto nrm = [](std::array<double, 3> a) {
            double l = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
            return std::array<double, 3>{a[0] / l, a[1] / l, a[2] / l};
        };
        auto cross = [](auto a, auto b) {
            return std::array<double, 3>{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        };
        auto d = nrm(sub(look, pos));
        auto right = nrm(cross(nrm(up), d));
        auto new_up = cross(d, right);
        M cam = eye();
        for (int r = 0; r < 3; ++r) {
            cam[4 * r + 0] = right[r]; cam[4 * r + 1] = new_up[r]; cam[4 * r + 2] = d[r]; cam[4 * r + 3] = pos[r];
        }
        return Transform(invert(cam), cam);
    }
    static Transform camera_look_at(const std::array<double, 3>& pos, const std::array<double, 3>& look,
                                    const std::array<double, 3>& up) {                    // transform.rs:58-60
        return look_at(pos, look, up).inverse();
    }
    static Transform perspective(double fov_deg, double near_, double far_) {              // transform.rs:105-115
        M p = eye();
        p[10] = far_ / (far_ - near_); p[11] = -far_ * near_ / (far_ - near_);
        p[14] = 1.0; p[15] = 0.0;
        double inv_tan = 1.0 / std::tan(fov_deg * M_PI / 180.0 / 2.0);
        return scale(inv_tan, inv_tan, 1.0) * Transform(p);
    }

    Transform inverse() const { return Transform(minv, m); }                              // transform.rs:121-123
    Transform operator*(const Transform& rhs) const {                                      // transform.rs:164-170
        return Transform(mul(m, rhs.m), mul(rhs.minv, minv));
    }
    bool swaps_handedness() const {                                                        // transform.rs:126-128
        double det = m[0] * (m[5] * m[10] - m[6] * m[9]) - m[1] * (m[4] * m[10] - m[6] * m[8])
                   + m[2] * (m[4] * m[9] - m[5] * m[8]);
        return det < 0.0;
    }
    bool is_identity() const { return m == eye(); }

    void flat(float out[16]) const { to_flat(m, out); }        // column-major f32 for the ABI
    void flat_inv(float out[16]) const { to_flat(minv, out); }

    // f32 evaluation in cgmath's operation order (Matrix4::transform_point multiplies by 1/w):
    // TriangleMesh::new moves the vertices to world space on the host (triangle.rs:42-51)
    Point3f apply_point_f32(const Point3f& p) const {
        float a[16];
        for (int i = 0; i < 16; ++i) a[i] = (float)m[i];
        float c[4];
        for (int r = 0; r < 4; ++r) c[r] = ((a[4 * r] * p.x + a[4 * r + 1] * p.y) + a[4 * r + 2] * p.z) + a[4 * r + 3] * 1.0f;
        float iw = 1.0f / c[3];
        return {c[0] * iw, c[1] * iw, c[2] * iw};
    }
    Vec3f apply_normal_f32(const Vec3f& n) const {                                         // transform.rs:134-140
        float a[16];
        for (int i = 0; i < 16; ++i) a[i] = (float)minv[i];
        return {(a[0] * n.x + a[4] * n.y) + a[8] * n.z, (a[1] * n.x + a[5] * n.y) + a[9] * n.z,
                (a[2] * n.x + a[6] * n.y) + a[10] * n.z};
    }

private:
    static M eye() { return {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }
    static M transpose(const M& a) {
        M t;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) t[4 * r + c] = a[4 * c + r];
        return t;
    }
    static M mul(const M& a, const M& b) {
        M o{};
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                double s = 0;
                for (int k = 0; k < 4; ++k) s += a[4 * r + k] * b[4 * k + c];
                o[4 * r + c] = s;
            }
        return o;
    }
    static void to_flat(const M& a, float out[16]) {
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[4 * c + r] = (float)a[4 * r + c];
    }
    static M invert(const M& a) {      // Gauss-Jordan with partial pivoting
        double w[4][8];
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { w[r][c] = a[4 * r + c]; w[r][4 + c] = (r == c); }
        for (int i = 0; i < 4; ++i) {
            int p = i;
            for (int r = i + 1; r < 4; ++r) if (std::fabs(w[r][i]) > std::fabs(w[p][i])) p = r;
            if (w[p][i] == 0.0) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: singular transform");
            if (p != i) for (int c = 0; c < 8; ++c) std::swap(w[p][c], w[i][c]);
            double inv = 1.0 / w[i][i];
            for (int c = 0; c < 8; ++c) w[i][c] *= inv;
            for (int r = 0; r < 4; ++r) if (r != i) {
                double f = w[r][i];
                if (f != 0.0) for (int c = 0; c < 8; ++c) w[r][c] -= f * w[i][c];
            }
        }
        M o;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) o[4 * r + c] = w[r][4 + c];
        return o;
    }
};

// geometry/mod.rs:88-102 `Ray::new(origin, dir)`: t_max = inf, time = 0
struct Ray : FtnRay {
    Ray(const Point3f& origin, const Vec3f& dir, float t_max_ = INFINITY, float time_ = 0.0f) {
        o[0] = origin.x; o[1] = origin.y; o[2] = origin.z;
        d[0] = dir.x; d[1] = dir.y; d[2] = dir.z;
        t_max = t_max_; time = time_;
    }
};

// ------------------------------------------------------------------------------------------
// PLY (stays on the host side of the boundary, as plydough does in the reference;
// loaders/constructors.rs:94-190 reads x,y,z,[nx,ny,nz],[u,v] + vertex_indices lists)
// ------------------------------------------------------------------------------------------
struct PlyData {
    std::vector<float> vertices, normals, uvs;
    std::vector<uint32_t> indices;
};

inline PlyData load_ply(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: cannot open " + path);
    struct Prop { std::string name, type, count_type; bool list; };
    struct Elem { std::string name; size_t count; std::vector<Prop> props; };
    std::vector<Elem> elems;
    std::string line, format;
    std::getline(f, line);
    if (line.substr(0, 3) != "ply") throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: not a PLY file: " + path);
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ss(line);
        std::string tok;
        ss >> tok;
        if (tok == "format") ss >> format;
        else if (tok == "element") { Elem e; ss >> e.name >> e.count; elems.push_back(e); }
        else if (tok == "property") {
            Prop p; std::string t; ss >> t;
            if (t == "list") { p.list = true; ss >> p.count_type >> p.type >> p.name; }
            else { p.list = false; p.type = t; ss >> p.name; }
            if (elems.empty()) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: PLY property before element");
            elems.back().props.push_back(p);
        } else if (tok == "end_header") break;
    }
    const bool ascii = format == "ascii";
    if (!ascii && format != "binary_little_endian")
        throw Error(FTN_ERR_UNSUPPORTED, "fountain: PLY format " + format + " not supported");
    auto size_of = [](const std::string& t) -> int {
        if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
        if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
        if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
        if (t == "double" || t == "float64") return 8;
        throw Error(FTN_ERR_UNSUPPORTED, "fountain: PLY type " + t);
    };
    auto read_num = [&](const std::string& t) -> double {
        if (ascii) { double v; f >> v; return v; }
        char b[8]; int n = size_of(t); f.read(b, n);
        if (t == "float" || t == "float32") { float v; std::memcpy(&v, b, 4); return v; }
        if (t == "double" || t == "float64") { double v; std::memcpy(&v, b, 8); return v; }
        if (t == "char" || t == "int8") { int8_t v; std::memcpy(&v, b, 1); return v; }
        if (t == "uchar" || t == "uint8") { uint8_t v; std::memcpy(&v, b, 1); return v; }
        if (t == "short" || t == "int16") { int16_t v; std::memcpy(&v, b, 2); return v; }
        if (t == "ushort" || t == "uint16") { uint16_t v; std::memcpy(&v, b, 2); return v; }
        if (t == "int" || t == "int32") { int32_t v; std::memcpy(&v, b, 4); return v; }
        uint32_t v; std::memcpy(&v, b, 4); return v;
    };
    PlyData out;
    bool has_n = false, has_uv = false;
    for (const Elem& e : elems) {
        if (e.name == "vertex") {
            int ix = -1, iy = -1, iz = -1, inx = -1, iny = -1, inz = -1, iu = -1, iv = -1;
            for (size_t i = 0; i < e.props.size(); ++i) {
                const std::string& n = e.props[i].name;
                if (n == "x") ix = i; else if (n == "y") iy = i; else if (n == "z") iz = i;
                else if (n == "nx") inx = i; else if (n == "ny") iny = i; else if (n == "nz") inz = i;
                else if (n == "u") iu = i; else if (n == "v") iv = i;
            }
            if (ix < 0 || iy < 0 || iz < 0) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: PLY without x,y,z");
            has_n = inx >= 0 && iny >= 0 && inz >= 0;
            has_uv = iu >= 0 && iv >= 0;
            std::vector<double> row(e.props.size());
            for (size_t k = 0; k < e.count; ++k) {
                for (size_t i = 0; i < e.props.size(); ++i) row[i] = read_num(e.props[i].type);
                out.vertices.insert(out.vertices.end(), {(float)row[ix], (float)row[iy], (float)row[iz]});
                if (has_n) out.normals.insert(out.normals.end(), {(float)row[inx], (float)row[iny], (float)row[inz]});
                if (has_uv) out.uvs.insert(out.uvs.end(), {(float)row[iu], (float)row[iv]});
            }
        } else {
            for (size_t k = 0; k < e.count; ++k)
                for (const Prop& p : e.props) {
                    if (!p.list) { read_num(p.type); continue; }
                    int n = (int)read_num(p.count_type);
                    std::vector<uint32_t> v(n);
                    for (int i = 0; i < n; ++i) v[i] = (uint32_t)read_num(p.type);
                    if (e.name == "face" && p.name == "vertex_indices") {
                        // constructors.rs:155-157: any face that is not a triangle is fatal
                        if (n != 3) throw Error(FTN_ERR_UNSUPPORTED, "fountain: PLY face with unsupported vertex count " + std::to_string(n));
                        out.indices.insert(out.indices.end(), {v[0], v[1], v[2]});
                    }
                }
        }
    }
    if (!f && !f.eof()) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: truncated PLY " + path);
    return out;
}

// ------------------------------------------------------------------------------------------
// shapes / materials / lights
// ------------------------------------------------------------------------------------------
// shapes/triangle.rs:29-74: vertices (and normals) are moved to world space at construction
class TriangleMesh {
public:
    std::vector<uint32_t> vertex_indices;
    std::vector<float> vertices, normals, tex_coords;
    bool reverse_orientation;
    Transform object_to_world;

    TriangleMesh(const Transform& o2w, std::vector<uint32_t> idx, const std::vector<float>& v,
                 const std::vector<float>& n = {}, std::vector<float> uv = {}, bool reverse = false)
        : vertex_indices(std::move(idx)), tex_coords(std::move(uv)), reverse_orientation(reverse), object_to_world(o2w) {
        if (vertex_indices.size() % 3 != 0) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: indices not a multiple of 3");
        if (v.size() % 3 != 0) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: vertices not a multiple of 3");
        if (!n.empty() && n.size() != v.size()) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: normals/vertices mismatch");
        if (!tex_coords.empty() && tex_coords.size() / 2 != v.size() / 3)
            throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: uvs/vertices mismatch");
        const bool ident = o2w.is_identity();
        vertices = v;
        normals = n;
        if (!ident) {
            for (size_t i = 0; i + 2 < v.size(); i += 3) {
                Point3f p = o2w.apply_point_f32({v[i], v[i + 1], v[i + 2]});
                vertices[i] = p.x; vertices[i + 1] = p.y; vertices[i + 2] = p.z;
            }
            for (size_t i = 0; i + 2 < n.size(); i += 3) {
                Vec3f q = o2w.apply_normal_f32({n[i], n[i + 1], n[i + 2]});
                normals[i] = q.x; normals[i + 1] = q.y; normals[i + 2] = q.z;
            }
        }
    }
    size_t n_triangles() const { return vertex_indices.size() / 3; }
    size_t n_vertices() const { return vertices.size() / 3; }
    bool flip_normals() const { return reverse_orientation ^ object_to_world.swaps_handedness(); }   // shapes/mod.rs:27-29

    // make_triangle_mesh_from_ply (loaders/constructors.rs:94-190)
    static std::shared_ptr<TriangleMesh> from_ply(const std::string& path, const Transform& o2w = Transform::identity(),
                                                  bool reverse = false) {
        PlyData d = load_ply(path);
        return std::make_shared<TriangleMesh>(o2w, std::move(d.indices), d.vertices, d.normals, std::move(d.uvs), reverse);
    }
};

struct Sphere {                         // shapes/sphere.rs:30-58
    Transform object_to_world;
    bool reverse_orientation = false;
    float radius = 1.0f, z_min = -1.0f, z_max = 1.0f, phi_max = 360.0f;
    Sphere(const Transform& o2w, bool reverse, float r) : object_to_world(o2w), reverse_orientation(reverse), radius(r),
                                                           z_min(-r), z_max(r) {}
    Sphere(const Transform& o2w, bool reverse, float r, float zmin, float zmax, float phimax)
        : object_to_world(o2w), reverse_orientation(reverse), radius(r), z_min(zmin), z_max(zmax), phi_max(phimax) {}
};

struct UVMapping {                       // texture/mapping.rs:13-53; defaults (constructors.rs:251-254)
    float scale_u = 1.0f, scale_v = 1.0f, offset_u = 0.0f, offset_v = 0.0f;
};
struct SpectrumTexture;
struct Material {
    virtual ~Material() = default;
    virtual void fill(FtnMaterial& m) const = 0;
    // Textures of the parameters the loader reads with get_texture_or_default (constructors.rs:192-238): Ks, eta, k, Kr, Kt,
    // roughnesses, sigma, the glass index.  They travel through the scene's texture table (FtnSceneDesc::textures); float
    // parameters read component 0.  Kd (matte, plastic) / Kr (mirror) may also use the constructor's inline slot.
    std::vector<std::pair<int, std::shared_ptr<SpectrumTexture>>> param_textures;
    Material& with_texture(FtnMaterialParam p, const SpectrumTexture& t);
};
// mipmap.rs:19-143 `MIPMap<Spectrum>`: the pyramid an ImageTexture filters.  Level l is max(1, w >> l) x max(1, h >> l)
// and there are 1 + floor(log2(max(w, h))) levels (:107-121).  The reference halves each level with the `resize`
// crate's Triangle filter (0.4.3, a third-party dependency that is not in its repository); from_image restates that
// filter for a factor of two -- weights (1, 3, 3, 1) / 8 per axis, renormalised at the borders.  Host-side setup.
struct MIPMap {
    int width = 0, height = 0, wrap = FTN_WRAP_REPEAT;
    int n_levels = 0;
    std::vector<float> packed;            // the ABI's layout: levels concatenated, RGB, row-major with s fastest
    static std::vector<float> halve(const std::vector<float>& src, int w, int h, int axis) {
        const int n = axis == 0 ? w : h, m = std::max(1, n / 2);
        const int ow = axis == 0 ? m : w, oh = axis == 0 ? h : m;
        std::vector<float> out((size_t)3 * ow * oh);
        static const double wt[4] = {1.0, 3.0, 3.0, 1.0};
        for (int y = 0; y < oh; ++y) for (int x = 0; x < ow; ++x) for (int c = 0; c < 3; ++c) {
            double acc = 0.0, ws = 0.0;
            for (int k = 0; k < 4; ++k) {
                const int j = 2 * (axis == 0 ? x : y) + k - 1;
                if (j < 0 || j >= n) continue;
                const int sx = axis == 0 ? j : x, sy = axis == 0 ? y : j;
                acc += wt[k] * src[3 * ((size_t)sy * w + sx) + c]; ws += wt[k];
            }
            out[3 * ((size_t)y * ow + x) + c] = (float)(acc / ws);
        }
        return out;
    }
    // image: RGB f32, row-major (s fastest), width * height * 3
    static std::shared_ptr<MIPMap> from_image(const std::vector<float>& image, int w, int h, int wrap_mode = FTN_WRAP_REPEAT) {
        if (w < 1 || h < 1 || image.size() != (size_t)3 * w * h) throw Error(FTN_ERR_INVALID_ARGUMENT, "MIPMap: image must hold width * height * 3 floats");
        auto mp = std::make_shared<MIPMap>();
        mp->width = w; mp->height = h; mp->wrap = wrap_mode;
        mp->n_levels = 1; for (int m = std::max(w, h); m > 1; m >>= 1) ++mp->n_levels;
        std::vector<float> level = image;
        int cw = w, ch = h;
        mp->packed = level;
        for (int l = 1; l < mp->n_levels; ++l) {
            if (cw > 1) { level = halve(level, cw, ch, 0); cw = std::max(1, cw / 2); }
            if (ch > 1) { level = halve(level, cw, ch, 1); ch = std::max(1, ch / 2); }
            mp->packed.insert(mp->packed.end(), level.begin(), level.end());
        }
        return mp;
    }
};

// A spectrum texture for Kd: ConstantTexture (texture/mod.rs:34-42), Checkerboard2DTexture over two constant
// spectra with AAMethod::None (checkerboard.rs:10-64), UVTexture (uv.rs:6-24) or ImageTexture (image.rs:8-34)
struct SpectrumTexture {
    int type = FTN_TEXTURE_CONSTANT;
    Spectrum value, tex1, tex2;
    UVMapping mapping;
    std::shared_ptr<MIPMap> mipmap;      // IMAGE
    static SpectrumTexture image(std::shared_ptr<MIPMap> mp, UVMapping m = {}) {
        SpectrumTexture t(Spectrum(0.0f)); t.type = FTN_TEXTURE_IMAGE; t.mipmap = std::move(mp); t.mapping = m; return t;
    }
    SpectrumTexture(Spectrum constant) : value(constant) {}
    SpectrumTexture(float constant) : value(constant) {}
    static SpectrumTexture checkerboard(Spectrum tex1, Spectrum tex2, UVMapping m = {}) {
        SpectrumTexture t(Spectrum(0.0f)); t.type = FTN_TEXTURE_CHECKERBOARD; t.tex1 = tex1; t.tex2 = tex2; t.mapping = m; return t;
    }
    static SpectrumTexture uv(UVMapping m = {}) { SpectrumTexture t(Spectrum(0.0f)); t.type = FTN_TEXTURE_UV; t.mapping = m; return t; }
    void fill_kd(FtnMaterial& m) const {
        m.kd_texture = type;
        m.kd[0] = value.r; m.kd[1] = value.g; m.kd[2] = value.b;
        m.tex1[0] = tex1.r; m.tex1[1] = tex1.g; m.tex1[2] = tex1.b;
        m.tex2[0] = tex2.r; m.tex2[1] = tex2.g; m.tex2[2] = tex2.b;
        m.uv_scale[0] = mapping.scale_u; m.uv_scale[1] = mapping.scale_v;
        m.uv_delta[0] = mapping.offset_u; m.uv_delta[1] = mapping.offset_v;
        if (type == FTN_TEXTURE_IMAGE) {
            if (!mipmap) throw Error(FTN_ERR_INVALID_ARGUMENT, "image texture without a MIPMap");
            m.image = mipmap->packed.data(); m.image_width = mipmap->width; m.image_height = mipmap->height;
            m.image_levels = mipmap->n_levels; m.image_wrap = mipmap->wrap;
        }
    }
};
inline Material& Material::with_texture(FtnMaterialParam p, const SpectrumTexture& t) {
    param_textures.emplace_back((int)p, std::make_shared<SpectrumTexture>(t));
    return *this;
}
inline FtnTexture to_abi_texture(const SpectrumTexture& t) {
    FtnTexture o{};
    o.type = t.type;
    o.value[0] = t.value.r; o.value[1] = t.value.g; o.value[2] = t.value.b;
    o.tex1[0] = t.tex1.r; o.tex1[1] = t.tex1.g; o.tex1[2] = t.tex1.b;
    o.tex2[0] = t.tex2.r; o.tex2[1] = t.tex2.g; o.tex2[2] = t.tex2.b;
    o.uv_scale[0] = t.mapping.scale_u; o.uv_scale[1] = t.mapping.scale_v;
    o.uv_delta[0] = t.mapping.offset_u; o.uv_delta[1] = t.mapping.offset_v;
    if (t.type == FTN_TEXTURE_IMAGE) {
        if (!t.mipmap) throw Error(FTN_ERR_INVALID_ARGUMENT, "image texture without a MIPMap");
        o.image = t.mipmap->packed.data(); o.image_width = t.mipmap->width; o.image_height = t.mipmap->height;
        o.image_levels = t.mipmap->n_levels; o.image_wrap = t.mipmap->wrap;
    }
    return o;
}
struct MatteMaterial : Material {        // material/matte.rs; Kd default 0.5, sigma default 0 (constructors.rs:192-196)
    SpectrumTexture kd;
    float sigma;                         // degrees; != 0 selects Oren-Nayar (matte.rs:42-49)
    explicit MatteMaterial(SpectrumTexture kd_ = SpectrumTexture(0.5f), float sigma_ = 0.0f) : kd(kd_), sigma(sigma_) {}
    void fill(FtnMaterial& m) const override {
        m.type = FTN_MATERIAL_MATTE;
        kd.fill_kd(m);
        m.sigma = sigma;
    }
};
struct MetalMaterial : Material {        // material/metal.rs; roughness 0.01, remap true (constructors.rs:213-230)
    Spectrum eta, k;
    float u_roughness, v_roughness;
    bool remap_roughness;
    MetalMaterial(Spectrum eta_, Spectrum k_, float roughness = 0.01f, bool remap = true)
        : eta(eta_), k(k_), u_roughness(roughness), v_roughness(roughness), remap_roughness(remap) {}
    MetalMaterial(Spectrum eta_, Spectrum k_, float u, float v, bool remap)
        : eta(eta_), k(k_), u_roughness(u), v_roughness(v), remap_roughness(remap) {}
    void fill(FtnMaterial& m) const override {
        m.type = FTN_MATERIAL_METAL;
        m.eta[0] = eta.r; m.eta[1] = eta.g; m.eta[2] = eta.b;
        m.k[0] = k.r; m.k[1] = k.g; m.k[2] = k.b;
        m.u_roughness = u_roughness; m.v_roughness = v_roughness; m.remap_roughness = remap_roughness;
    }
};
struct PlasticMaterial : Material {      // material/plastic.rs; Kd = Ks = .25, roughness .1 (constructors.rs:232-238)
    SpectrumTexture kd;
    Spectrum ks;
    float roughness;
    bool remap_roughness;
    explicit PlasticMaterial(SpectrumTexture kd_ = SpectrumTexture(0.25f), Spectrum ks_ = Spectrum(0.25f), float rough = 0.1f, bool remap = true)
        : kd(kd_), ks(ks_), roughness(rough), remap_roughness(remap) {}
    void fill(FtnMaterial& m) const override {
        m.type = FTN_MATERIAL_PLASTIC;
        kd.fill_kd(m);
        m.ks[0] = ks.r; m.ks[1] = ks.g; m.ks[2] = ks.b;
        m.u_roughness = m.v_roughness = roughness; m.remap_roughness = remap_roughness;
    }
};

struct MirrorMaterial : Material {       // material/mirror.rs; Kr default 0.9 (constructors.rs:207-210)
    SpectrumTexture kr;                  // constant or textured (mirror.rs:23)
    explicit MirrorMaterial(SpectrumTexture kr_ = SpectrumTexture(0.9f)) : kr(kr_) {}
    void fill(FtnMaterial& m) const override {
        m.type = FTN_MATERIAL_MIRROR;
        kr.fill_kd(m);                   // the ABI's texture slot serves Kd (matte, plastic) or Kr (mirror)
        m.kr[0] = kr.value.r; m.kr[1] = kr.value.g; m.kr[2] = kr.value.b;
    }
};

struct GlassMaterial : Material {        // material/glass.rs; defaults of constructors.rs:198-205
    Spectrum kr{1.0f}, kt{1.0f};
    float eta = 1.5f, u_roughness = 0.0f, v_roughness = 0.0f;
    bool remap_roughness = true;         // roughness 0 then becomes a small alpha: rough glass (alphas of exactly 0 are todo!() in the reference)
    GlassMaterial() {}
    GlassMaterial(Spectrum kr_, Spectrum kt_, float eta_, float ur, float vr, bool remap = true)
        : kr(kr_), kt(kt_), eta(eta_), u_roughness(ur), v_roughness(vr), remap_roughness(remap) {}
    void fill(FtnMaterial& m) const override {
        m.type = FTN_MATERIAL_GLASS;
        m.kr[0] = kr.r; m.kr[1] = kr.g; m.kr[2] = kr.b;
        m.kt[0] = kt.r; m.kt[1] = kt.g; m.kt[2] = kt.b;
        m.eta[0] = m.eta[1] = m.eta[2] = eta;
        m.u_roughness = u_roughness; m.v_roughness = v_roughness; m.remap_roughness = remap_roughness;
    }
};

struct DiffuseAreaLight {                // light/diffuse.rs:24-41
    Spectrum emit;
    explicit DiffuseAreaLight(Spectrum l = Spectrum(1.0f)) : emit(l) {}
};

struct InfiniteAreaLight {               // light/infinite.rs:23-61
    std::vector<float> texels;           // RGB, row-major, width*height*3
    int width = 1, height = 1;
    Transform light_to_world;
    static InfiniteAreaLight new_uniform(Spectrum l, const Transform& l2w = Transform::identity()) {
        InfiniteAreaLight e;
        e.texels = {l.r, l.g, l.b};
        e.light_to_world = l2w;
        return e;
    }
    static InfiniteAreaLight new_envmap(std::vector<float> rgb, int w, int h, const Transform& l2w = Transform::identity()) {
        if ((size_t)w * h * 3 != rgb.size()) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: env map size mismatch");
        InfiniteAreaLight e;
        e.texels = std::move(rgb); e.width = w; e.height = h; e.light_to_world = l2w;
        return e;
    }
};

struct PointLight {                      // light/point.rs:16-28
    Point3f world_point;
    Spectrum intensity;
    PointLight(const Transform& light_to_world, Spectrum i) : world_point(light_to_world.apply_point_f32({0, 0, 0})), intensity(i) {}
};

struct DistantLight {                    // light/distant.rs:18-31
    Spectrum radiance;
    Vec3f dir_to_light;                  // normalised in f32, cgmath order: v * (1 / |v|)
    DistantLight(Spectrum l, Vec3f d) : radiance(l) {
        float inv = 1.0f / std::sqrt((d.x * d.x + d.y * d.y) + d.z * d.z);
        dir_to_light = Vec3f(d.x * inv, d.y * inv, d.z * inv);
    }
    static DistantLight from_to(Point3f from, Point3f to, Spectrum l) {
        return DistantLight(l, Vec3f(from.x - to.x, from.y - to.y, from.z - to.z));
    }
};

// One explicit `LightSource` (scene/mod.rs:32-49 takes them in file order)
struct Light {
    int type = FTN_LIGHT_INFINITE;
    InfiniteAreaLight infinite;
    Point3f point; Vec3f direction; Spectrum intensity;
    Light(InfiniteAreaLight l) : type(FTN_LIGHT_INFINITE), infinite(std::move(l)) {}
    Light(const PointLight& l) : type(FTN_LIGHT_POINT), point(l.world_point), intensity(l.intensity) {}
    Light(const DistantLight& l) : type(FTN_LIGHT_DISTANT), direction(l.dir_to_light), intensity(l.radiance) {}
};

// primitive.rs:25-29; a TriangleMesh stands for one primitive per triangle (loaders/pbrt.rs:275-317)
struct GeometricPrimitive {
    std::shared_ptr<TriangleMesh> mesh;
    std::shared_ptr<Sphere> sphere;
    std::shared_ptr<Material> material;
    std::shared_ptr<DiffuseAreaLight> light;
    // a light on a mesh = `AreaLightSource "diffuse"` in front of the shape: every triangle carries its own
    // DiffuseAreaLight<Triangle> (loaders/pbrt.rs:275-316)
    GeometricPrimitive(std::shared_ptr<TriangleMesh> m, std::shared_ptr<Material> mat = nullptr,
                       std::shared_ptr<DiffuseAreaLight> l = nullptr)
        : mesh(std::move(m)), material(std::move(mat)), light(std::move(l)) {}
    GeometricPrimitive(std::shared_ptr<Sphere> s, std::shared_ptr<Material> mat = nullptr,
                       std::shared_ptr<DiffuseAreaLight> l = nullptr)
        : sphere(std::move(s)), material(std::move(mat)), light(std::move(l)) {}
};

// ------------------------------------------------------------------------------------------
// Scene (scene/mod.rs:14-68): flattens into FtnSceneDesc, uploads, builds the aggregate
// ------------------------------------------------------------------------------------------
class Scene {
public:
    Scene(LibraryPtr lib, const std::vector<GeometricPrimitive>& prims, const std::vector<Light>& lights = {},
          bool build = true)
        : lib_(std::move(lib)) {
        std::vector<const Material*> mats;
        auto mat_id = [&](const std::shared_ptr<Material>& m) -> int32_t {
            if (!m) return -1;
            for (size_t i = 0; i < mats.size(); ++i) if (mats[i] == m.get()) return (int32_t)i;
            mats.push_back(m.get());
            return (int32_t)mats.size() - 1;
        };
        std::vector<float> pos, nrm, uvs;
        std::vector<uint32_t> idx;
        std::vector<FtnMeshDesc> meshes;
        std::vector<FtnSphere> spheres;
        int with_n = 0, with_uv = 0, n_mesh = 0;
        for (const auto& p : prims) if (p.mesh) { ++n_mesh; with_n += !p.mesh->normals.empty(); with_uv += !p.mesh->tex_coords.empty(); }
        if ((with_n && with_n != n_mesh) || (with_uv && with_uv != n_mesh))
            throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: either every mesh carries normals/uvs or none does");
        uint32_t vbase = 0, first = 0;
        // primitive ids of the ABI: triangles first (mesh order, tri_id), then spheres
        for (const auto& p : prims) {
            if (!p.mesh) continue;
            const TriangleMesh& m = *p.mesh;
            pos.insert(pos.end(), m.vertices.begin(), m.vertices.end());
            nrm.insert(nrm.end(), m.normals.begin(), m.normals.end());
            uvs.insert(uvs.end(), m.tex_coords.begin(), m.tex_coords.end());
            for (uint32_t i : m.vertex_indices) idx.push_back(i + vbase);
            vbase += (uint32_t)m.n_vertices();
            FtnMeshDesc d{};
            d.first_tri = first; d.n_tris = (uint32_t)m.n_triangles(); d.material_id = mat_id(p.material);
            d.flags = m.flip_normals() ? (uint32_t)FTN_MESH_FLIP_NORMALS : 0u;
            d.emissive = p.light != nullptr;
            if (p.light) { d.emit[0] = p.light->emit.r; d.emit[1] = p.light->emit.g; d.emit[2] = p.light->emit.b; }
            meshes.push_back(d);
            first += d.n_tris;
        }
        for (const auto& p : prims) {
            if (!p.sphere) continue;
            const Sphere& s = *p.sphere;
            FtnSphere c{};
            s.object_to_world.flat(c.object_to_world);
            s.object_to_world.flat_inv(c.world_to_object);
            c.radius = s.radius; c.z_min = s.z_min; c.z_max = s.z_max; c.phi_max_deg = s.phi_max;
            c.reverse_orientation = s.reverse_orientation;
            c.material_id = mat_id(p.material);
            c.emissive = p.light != nullptr;
            if (p.light) { c.emit[0] = p.light->emit.r; c.emit[1] = p.light->emit.g; c.emit[2] = p.light->emit.b; }
            spheres.push_back(c);
        }
        std::vector<FtnMaterial> cm(mats.size());
        std::vector<FtnTexture> ct;          // the texture table
        for (size_t i = 0; i < mats.size(); ++i) {
            cm[i] = FtnMaterial{}; mats[i]->fill(cm[i]);
            for (const auto& pt : mats[i]->param_textures) {
                if (pt.first < 0 || pt.first >= FTN_PARAM_COUNT) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: unknown material parameter");
                ct.push_back(to_abi_texture(*pt.second));
                cm[i].param_texture[pt.first] = (uint32_t)ct.size();
            }
        }
        std::vector<FtnLight> cl(lights.size());
        for (size_t i = 0; i < lights.size(); ++i) {
            cl[i] = FtnLight{};
            cl[i].type = lights[i].type;
            Transform l2w = lights[i].type == FTN_LIGHT_INFINITE ? lights[i].infinite.light_to_world : Transform::identity();
            l2w.flat(cl[i].light_to_world);
            l2w.flat_inv(cl[i].world_to_light);
            if (lights[i].type == FTN_LIGHT_INFINITE) {
                cl[i].texels = lights[i].infinite.texels.data();
                cl[i].width = lights[i].infinite.width; cl[i].height = lights[i].infinite.height;
            }
            cl[i].point[0] = lights[i].point.x; cl[i].point[1] = lights[i].point.y; cl[i].point[2] = lights[i].point.z;
            cl[i].direction[0] = lights[i].direction.x; cl[i].direction[1] = lights[i].direction.y; cl[i].direction[2] = lights[i].direction.z;
            cl[i].intensity[0] = lights[i].intensity.r; cl[i].intensity[1] = lights[i].intensity.g; cl[i].intensity[2] = lights[i].intensity.b;
        }
        FtnSceneDesc d{};
        d.abi_version = FTN_ABI_VERSION;
        d.positions = pos.data();
        d.normals = nrm.empty() ? nullptr : nrm.data();
        d.uvs = uvs.empty() ? nullptr : uvs.data();
        d.n_vertices = vbase;
        d.indices = idx.data();
        d.n_triangles = first;
        d.meshes = meshes.data(); d.n_meshes = (uint32_t)meshes.size();
        d.spheres = spheres.data(); d.n_spheres = (uint32_t)spheres.size();
        d.materials = cm.data(); d.n_materials = (uint32_t)cm.size();
        d.lights = cl.data(); d.n_lights = (uint32_t)cl.size();
        d.textures = ct.empty() ? nullptr : ct.data(); d.n_textures = (uint32_t)ct.size();
        n_triangles_ = first;
        lib_->check(lib_->scene_create(&d, &handle_), "ftn_scene_create");
        if (build) this->build();
    }
    ~Scene() { if (handle_) lib_->scene_destroy(handle_); }
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;

    void build() { lib_->check(lib_->bvh_build(handle_), "ftn_bvh_build"); }               // BVH::build, bvh.rs:27

    // Scene::intersect (scene/mod.rs:51) over a batch; prim == FTN_NO_HIT on a miss
    std::vector<FtnHit> intersect(const std::vector<Ray>& rays) const {
        std::vector<FtnHit> hits(rays.size());
        lib_->check(lib_->intersect(handle_, rays.size(), rays.data(), hits.data()), "ftn_intersect");
        return hits;
    }
    // Scene::intersect_test (scene/mod.rs:55) over a batch
    std::vector<uint8_t> intersect_test(const std::vector<Ray>& rays) const {
        std::vector<uint8_t> out(rays.size());
        lib_->check(lib_->intersect_test(handle_, rays.size(), rays.data(), out.data()), "ftn_intersect_test");
        return out;
    }
    std::pair<Point3f, Point3f> world_bound() const {                                      // scene/mod.rs:66
        float b[6];
        lib_->check(lib_->scene_world_bound(handle_, b), "ftn_scene_world_bound");
        return {Point3f(b[0], b[1], b[2]), Point3f(b[3], b[4], b[5])};
    }
    void morton_codes_and_order(std::vector<uint32_t>& codes, std::vector<uint32_t>& order) const {
        codes.assign(n_triangles_, 0); order.assign(n_triangles_, 0);
        lib_->check(lib_->bvh_debug_morton(handle_, codes.data(), order.data()), "ftn_bvh_debug_morton");
    }
    FtnStats stats() const {
        FtnStats s{};
        lib_->check(lib_->scene_stats(handle_, &s), "ftn_scene_stats");
        return s;
    }
    const FtnScene* handle() const { return handle_; }
    const LibraryPtr& library() const { return lib_; }
    uint32_t n_triangles() const { return n_triangles_; }

private:
    LibraryPtr lib_;
    FtnScene* handle_ = nullptr;
    uint32_t n_triangles_ = 0;
};

// ------------------------------------------------------------------------------------------
// sensor
// ------------------------------------------------------------------------------------------
// camera/mod.rs:85-114 (CameraProjection::new :50-72); screen window default from
// PbrtHeader::make_camera (loaders/pbrt.rs:440-451)
class PerspectiveCamera {
public:
    Transform camera_to_world, raster_to_camera;
    float lens_radius, focal_dist, shutter_open, shutter_close;

    PerspectiveCamera(const Transform& c2w, int xres, int yres, float fov, float lens_radius_ = 0.0f, float focal_dist_ = 1e6f,
                      float shutter_open_ = 0.0f, float shutter_close_ = 1.0f)
        : camera_to_world(c2w), lens_radius(lens_radius_), focal_dist(focal_dist_), shutter_open(shutter_open_),
          shutter_close(shutter_close_) {
        float aspect = (float)xres / (float)yres;
        double x0, y0, x1, y1;
        if (aspect > 1.0f) { x0 = -aspect; x1 = aspect; y0 = -1.0; y1 = 1.0; }
        else { x0 = -1.0; x1 = 1.0; y0 = -1.0 / aspect; y1 = 1.0 / aspect; }
        Transform camera_to_screen = Transform::perspective(fov, 1.0e-2, 1000.0);
        Transform screen_to_raster = Transform::scale(xres, yres, 1.0) * Transform::scale(1.0 / (x1 - x0), 1.0 / (y0 - y1), 1.0)
                                   * Transform::translate(-x0, -y1, 0.0);
        raster_to_camera = camera_to_screen.inverse() * screen_to_raster.inverse();
    }
    FtnCamera to_abi() const {
        FtnCamera c{};
        camera_to_world.flat(c.camera_to_world);
        raster_to_camera.flat(c.raster_to_camera);
        c.lens_radius = lens_radius; c.focal_distance = focal_dist;
        c.shutter_open = shutter_open; c.shutter_close = shutter_close;
        return c;
    }
};

struct BoxFilter {                      // filter/mod.rs:10-32
    float radius_x = 0.5f, radius_y = 0.5f;
};

// film.rs:18-81.  After a render `pixels` holds the XYZ sums and filter-weight sums (film.rs:24)
class Film {
public:
    int x_resolution, y_resolution;
    std::array<float, 4> crop_window;   // x0, x1, y0, y1 (pbrt `cropwindow` order)
    BoxFilter filter;
    int width = 0, height = 0;          // cropped pixel bounds extent (film.rs:49-58)
    std::vector<FtnPixel> pixels;

    Film(const LibraryPtr& lib, int xres, int yres, std::array<float, 4> crop = {0.0f, 1.0f, 0.0f, 1.0f}, BoxFilter f = {})
        : x_resolution(xres), y_resolution(yres), crop_window(crop), filter(f) {
        FtnFilm a = to_abi();
        int32_t w = 0, h = 0;
        lib->check(lib->film_pixel_count(&a, &w, &h), "ftn_film_pixel_count");
        width = w; height = h;
        pixels.assign((size_t)w * h, FtnPixel{});
    }
    FtnFilm to_abi() const {
        FtnFilm f{};
        f.x_resolution = x_resolution; f.y_resolution = y_resolution;
        for (int i = 0; i < 4; ++i) f.crop_window[i] = crop_window[i];
        f.filter_radius[0] = filter.radius_x; f.filter_radius[1] = filter.radius_y;
        return f;
    }
    // film.rs:195-210: XYZ -> RGB (spectrum/mod.rs:28-34), times 1/weight, clamped at 0
    std::pair<std::vector<Spectrum>, std::pair<uint32_t, uint32_t>> into_spectrum_buffer() const {
        std::vector<Spectrum> img(pixels.size());
        for (size_t i = 0; i < pixels.size(); ++i) {
            const float* xyz = pixels[i].xyz;
            float r = (3.240479f * xyz[0] + -1.537150f * xyz[1]) + -0.498535f * xyz[2];
            float g = (-0.969256f * xyz[0] + 1.875991f * xyz[1]) + 0.041556f * xyz[2];
            float b = (0.055648f * xyz[0] + -0.204043f * xyz[1]) + 1.057311f * xyz[2];
            float w = pixels[i].filter_weight_sum;
            if (w != 0.0f) {
                float inv = 1.0f / w;
                r = std::max(0.0f, r * inv); g = std::max(0.0f, g * inv); b = std::max(0.0f, b * inv);
            }
            img[i] = Spectrum(r, g, b);
        }
        return {img, {(uint32_t)width, (uint32_t)height}};
    }
};

struct RandomSampler {                  // sampler/random.rs:6-21
    int samples_per_pixel;
    uint64_t seed;
    int mode;
    static RandomSampler new_with_seed(int spp, uint64_t seed, int mode = FTN_SAMPLER_COUNTER) { return {spp, seed, mode}; }
    FtnSampler to_abi(int sample_begin = 0, int sample_stride = 1) const {
        FtnSampler s{};
        s.samples_per_pixel = samples_per_pixel; s.seed = seed; s.mode = mode;
        s.sample_begin = sample_begin; s.sample_stride = sample_stride;
        return s;
    }
};

// ------------------------------------------------------------------------------------------
// integrators
// ------------------------------------------------------------------------------------------
struct IntegratorRadiance {              // integrator/mod.rs:27-37 (the radiance estimator's parameters)
    virtual ~IntegratorRadiance() = default;
    virtual FtnIntegrator to_abi() const = 0;
};
struct PathIntegrator : IntegratorRadiance {        // integrator/path.rs:10-20
    int max_depth;
    float rr_threshold;
    PathIntegrator(int depth, float rr) : max_depth(depth), rr_threshold(rr) {}
    static PathIntegrator new_(int depth, float rr) { return PathIntegrator(depth, rr); }
    FtnIntegrator to_abi() const override { return {FTN_INTEGRATOR_PATH, max_depth, rr_threshold}; }
};
enum class LightStrategy { UniformSampleAll, UniformSampleOne };
struct DirectLightingIntegrator : IntegratorRadiance {   // integrator/direct_lighting.rs:20-25
    LightStrategy strategy;
    int max_depth;
    DirectLightingIntegrator(LightStrategy s, int depth) : strategy(s), max_depth(depth) {
        // uniform_sample_all_lights is unimplemented!() in the reference (direct_lighting.rs:108-116)
        if (s != LightStrategy::UniformSampleOne) throw Error(FTN_ERR_UNSUPPORTED, "fountain: UniformSampleAll is unimplemented");
    }
    FtnIntegrator to_abi() const override { return {FTN_INTEGRATOR_DIRECT_LIGHTING, max_depth, 0.0f}; }
};

// integrator/mod.rs:22-25, 218-227
template <class R>
class SamplerIntegrator {
public:
    PerspectiveCamera camera;
    R radiance;
    FtnStats last_stats{};
    SamplerIntegrator(PerspectiveCamera cam, R rad) : camera(std::move(cam)), radiance(std::move(rad)) {}

    // Renders into film.pixels.  Throws Error(FTN_ERR_NAN_RADIANCE) where the reference panics in
    // check_radiance.  (sample_begin, sample_stride) select this caller's share of the sample
    // indices when several devices render one film.
    void render_parallel(const Scene& scene, Film& film, const RandomSampler& sampler, int sample_begin = 0, int sample_stride = 1) {
        FtnCamera c = camera.to_abi();
        FtnFilm f = film.to_abi();
        FtnSampler s = sampler.to_abi(sample_begin, sample_stride);
        FtnIntegrator it = radiance.to_abi();
        std::vector<FtnPixel> out((size_t)film.width * film.height);
        FtnStats st{};
        const LibraryPtr& lib = scene.library();
        lib->check(lib->render(scene.handle(), &c, &f, &s, &it, out.data(), &st), "ftn_render");
        film.pixels.swap(out);
        last_stats = st;
    }
};

}  // namespace fountain
#endif  // FOUNTAIN_HOST_HPP
