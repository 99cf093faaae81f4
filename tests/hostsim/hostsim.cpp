// TEST HARNESS ONLY -- not part of the product.
//
// Compiles the FTN_HD device functions of fountain_b200/csrc/*.cuh with g++ and drives them
// sequentially on the CPU behind the same C ABI (prefix `sim_`), so that the CPU-only test tier
// (`pytest -m "not gpu"`, no GPU in the build container) can check the kernels' per-thread logic
// -- LBVH topology + BVH2x64 emission + traversal, watertight triangle, EFloat sphere, BSDFs,
// env-map tables / sampling, path logic, film gather -- against the oracle before any GPU
// minute is spent.  What it cannot cover (radix sort, scans, atomics, queue compaction, launch
// glue) is covered by the `-m gpu` tests.  libfountain_gpu.so never links this file and has no
// CPU execution path.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>
#include <list>
#include "../../fountain_b200/csrc/ftn_path.cuh"
#include "../../fountain_b200/csrc/ftn_lbvh.cuh"
#include "../../fountain_b200/csrc/ftn_ploc.cuh"
#include "../../fountain_b200/csrc/ftn_bvh8_build.cuh"

using namespace ftn;

namespace {
thread_local std::string g_err;
int fail(int code, const char* m) { g_err = m; return code; }

struct SimScene {
    std::vector<float> pos, nrm, uv; std::vector<uint32_t> idx;
    std::vector<MeshData> meshes; std::vector<MaterialData> mats;
    std::vector<SphereData> spheres; std::vector<LightData> lights;
    std::vector<std::vector<F4>> env_tex; std::vector<std::vector<float>> env_f;   // storage behind EnvLightData pointers
    std::list<std::vector<uint32_t>> env_g;
    std::list<std::vector<F4>> images;   // storage behind MaterialData::image / TextureData::image
    std::vector<TextureData> texs;
    std::vector<F4> nodes, tris; uint32_t n_nodes = 0; bool wide = false; uint32_t bvh_levels = 0;
    std::vector<uint32_t> codes, order;
    float bounds[6]; bool built = false; uint32_t n_tris = 0;
    SceneView view() const {
        SceneView v;
        v.bvh.nodes = nodes.data(); v.bvh.tris = tris.data(); v.bvh.n_nodes = n_nodes; v.bvh.n_tris = n_tris; v.bvh.wide = wide ? 1u : 0u;
        v.pos = pos.data(); v.nrm = nrm.empty() ? nullptr : nrm.data(); v.uv = uv.empty() ? nullptr : uv.data(); v.idx = idx.data();
        v.meshes = meshes.data(); v.materials = mats.data(); v.textures = texs.data(); v.spheres = spheres.data(); v.n_spheres = (uint32_t)spheres.size();
        v.lights = lights.data(); v.n_lights = (uint32_t)lights.size(); v.n_tris = n_tris; v.refill_threshold = 20; v.vote_bias = 14; v.vote = true;
        return v;
    }
};
M4 to_m4(const float* f) { M4 m; std::memcpy(m.m, f, 64); return m; }
float roughness_to_alpha_host(float roughness) {   // same as scene.cu
    float rough = std::fmax(roughness, 1.0e-3f);
    float x = std::log(rough);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
}  // namespace

#define SIM_API extern "C" __attribute__((visibility("default")))

SIM_API uint32_t sim_abi_version(void) { return FTN_ABI_VERSION; }
SIM_API const char* sim_last_error(void) { return g_err.c_str(); }

SIM_API int sim_scene_create(const FtnSceneDesc* d, SimScene** out) {
    SimScene* s = new SimScene();
    s->n_tris = d->n_triangles;
    s->pos.assign(d->positions, d->positions + 3 * (size_t)d->n_vertices);
    if (d->normals) s->nrm.assign(d->normals, d->normals + 3 * (size_t)d->n_vertices);
    if (d->uvs) s->uv.assign(d->uvs, d->uvs + 2 * (size_t)d->n_vertices);
    s->idx.assign(d->indices, d->indices + 3 * (size_t)d->n_triangles);
    s->texs.resize(d->n_textures);
    for (uint32_t t = 0; t < d->n_textures; ++t) {   // as scene.cu
        s->images.emplace_back();
        if (texture_from_abi(d->textures[t], &s->texs[t], &s->images.back()) != FTN_OK) { delete s; g_err = "texture table: bad texture"; return FTN_ERR_INVALID_ARGUMENT; }
        if (s->texs[t].type == FTN_TEXTURE_IMAGE) s->texs[t].image = s->images.back().data();
    }
    for (uint32_t m = 0; m < d->n_materials; ++m) {
        const FtnMaterial fm = fold_constant_param_textures(d->materials[m], d->textures, d->n_textures); MaterialData md; std::memset(&md, 0, sizeof(md)); md.type = fm.type;
        for (int c = 0; c < 3; ++c) { md.kd[c] = fm.kd[c]; md.ks[c] = fm.ks[c]; md.eta[c] = fm.eta[c]; md.k[c] = fm.k[c]; }
        float ur = fm.u_roughness, vr = fm.v_roughness;
        if (fm.type == FTN_MATERIAL_MIRROR) for (int c = 0; c < 3; ++c) md.kd[c] = fm.kr[c];   // Kr travels in the kd slot
        if (fm.type == FTN_MATERIAL_GLASS) for (int c = 0; c < 3; ++c) { md.kd[c] = fm.kr[c]; md.ks[c] = fm.kt[c]; }
        md.kd_texture = (fm.type == FTN_MATERIAL_MATTE || fm.type == FTN_MATERIAL_PLASTIC || fm.type == FTN_MATERIAL_MIRROR) ? fm.kd_texture : 0;
        if (fm.type == FTN_MATERIAL_MATTE) {   // matte.rs:42-49: sigma clamped to [0, 90] degrees; != 0 -> OrenNayar::new (reflection/mod.rs:259-267)
            const float sigma = std::fmin(std::fmax(fm.sigma, 0.0f), 90.0f);
            if (sigma != 0.0f) {
                const float sr = sigma * (float)(3.14159265358979323846 / 180.0), s2 = sr * sr;
                md.type = FTN_CLASS_OREN_NAYAR;
                md.alpha_x = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));   // a
                md.alpha_y = 0.45f * s2 / (s2 + 0.09f);              // b
            }
        }
        for (int c = 0; c < 3; ++c) { md.tex1[c] = fm.tex1[c]; md.tex2[c] = fm.tex2[c]; }
        for (int c = 0; c < 2; ++c) { md.uv_scale[c] = fm.uv_scale[c]; md.uv_delta[c] = fm.uv_delta[c]; }
        md.image = nullptr; md.img_w = md.img_h = md.img_levels = md.img_wrap = 0;
        if (md.kd_texture == FTN_TEXTURE_IMAGE) {   // as scene.cu
            s->images.emplace_back();
            if (!pack_image_pyramid(fm, &s->images.back())) { delete s; g_err = "image texture: bad pyramid description"; return FTN_ERR_INVALID_ARGUMENT; }
            md.image = s->images.back().data(); md.img_w = fm.image_width; md.img_h = fm.image_height; md.img_levels = fm.image_levels; md.img_wrap = fm.image_wrap;
        }
        if (fm.type == FTN_MATERIAL_PLASTIC) vr = ur;
        if (fm.remap_roughness) { ur = roughness_to_alpha_host(ur); vr = roughness_to_alpha_host(vr); }
        if (md.type != FTN_CLASS_OREN_NAYAR) { md.alpha_x = ur; md.alpha_y = vr; }   // Oren-Nayar keeps (a, b) there
        if (material_params_from_abi(fm, d->n_textures, [&](uint32_t id) { return s->texs[id - 1].type == FTN_TEXTURE_IMAGE; }, &md) != FTN_OK) { delete s; g_err = "material: param_texture id out of range"; return FTN_ERR_INVALID_ARGUMENT; }
        if (fm.type == FTN_MATERIAL_GLASS && !(md.ptex[FTN_PARAM_UROUGHNESS] || md.ptex[FTN_PARAM_VROUGHNESS]) && ur == 0.0f && vr == 0.0f) { delete s; g_err = "smooth glass is todo!() in the reference (glass.rs:66)"; return FTN_ERR_UNSUPPORTED; }
        s->mats.push_back(md);
    }
    s->env_tex.reserve(d->n_lights); s->env_f.reserve(5 * d->n_lights);
    for (uint32_t l = 0; l < d->n_lights; ++l) {
        const FtnLight& fl = d->lights[l];
        LightData ld; std::memset(&ld, 0, sizeof(ld)); ld.type = 0; ld.sphere = -1;
        if (fl.type == FTN_LIGHT_POINT || fl.type == FTN_LIGHT_DISTANT) {   // as scene.cu
            ld.type = fl.type == FTN_LIGHT_POINT ? 2 : 3;
            for (int c = 0; c < 3; ++c) { ld.emit[c] = fl.intensity[c]; ld.vec[c] = fl.type == FTN_LIGHT_POINT ? fl.point[c] : fl.direction[c]; }
            s->lights.push_back(ld);
            continue;
        }
        EnvLightData& e = ld.env;
        e.w = fl.width; e.h = fl.height; e.nu = fl.height; e.nv = fl.width;
        int mx = std::max(fl.width, fl.height), lv = 0; while ((1 << (lv + 1)) <= mx) ++lv;
        e.levels = 1 + lv; e.l2w = to_m4(fl.light_to_world); e.w2l = to_m4(fl.world_to_light);
        const size_t n = (size_t)fl.width * fl.height;
        s->env_tex.emplace_back(n);
        for (size_t i = 0; i < n; ++i) { F4 t; t.x = fl.texels[3 * i]; t.y = fl.texels[3 * i + 1]; t.z = fl.texels[3 * i + 2]; t.w = 0; s->env_tex.back()[i] = t; }
        e.texels = s->env_tex.back().data();
        s->env_f.emplace_back(n); std::vector<float>& func = s->env_f.back();
        for (size_t k = 0; k < n; ++k) func[k] = env_func_value(e, (int)k);                       // k_env_func
        s->env_f.emplace_back((size_t)e.nv * (e.nu + 1)); std::vector<float>& cdf = s->env_f.back();
        s->env_f.emplace_back(e.nv); std::vector<float>& integ = s->env_f.back();
        for (int v = 0; v < e.nv; ++v) dist_row_build(&func[(size_t)v * e.nu], e.nu, &cdf[(size_t)v * (e.nu + 1)], &integ[v]);   // k_env_row_cdf
        s->env_f.emplace_back(e.nv + 1); std::vector<float>& mcdf = s->env_f.back();
        dist_row_build(integ.data(), e.nv, mcdf.data(), &e.marg_integral);
        e.cond_func = func.data(); e.cond_cdf = cdf.data(); e.cond_integral = integ.data(); e.marg_cdf = mcdf.data();
        s->env_g.emplace_back((size_t)e.nv * (e.nu + 1)); std::vector<uint32_t>& cg = s->env_g.back();                    // k_env_guide
        for (int v = 0; v < e.nv; ++v) for (int g = 0; g <= e.nu; ++g) cg[(size_t)v * (e.nu + 1) + g] = env_guide_entry(&cdf[(size_t)v * (e.nu + 1)], e.nu, g);
        s->env_g.emplace_back((size_t)e.nv + 1); std::vector<uint32_t>& mg = s->env_g.back();
        for (int g = 0; g <= e.nv; ++g) mg[g] = env_guide_entry(mcdf.data(), e.nv, g);
        e.cond_guide = cg.data(); e.marg_guide = mg.data();
        s->lights.push_back(ld);
    }
    build_mesh_table(d, &s->meshes, &s->lights);   // as scene.cu: the per-triangle area lights of emissive meshes
    for (uint32_t i = 0; i < d->n_spheres; ++i) {
        const FtnSphere& fs = d->spheres[i]; SphereData sd;
        sd.o2w = to_m4(fs.object_to_world); sd.w2o = to_m4(fs.world_to_object);
        const float r = fs.radius; sd.radius = r;
        sd.z_min = std::fmin(std::fmax(std::fmin(fs.z_min, fs.z_max), -r), r);
        sd.z_max = std::fmin(std::fmax(std::fmax(fs.z_min, fs.z_max), -r), r);
        sd.theta_min = std::acos(std::fmin(std::fmax(fs.z_min / r, -1.0f), 1.0f));
        sd.theta_max = std::acos(std::fmin(std::fmax(fs.z_max / r, -1.0f), 1.0f));
        sd.phi_max = std::fmin(std::fmax(fs.phi_max_deg, 0.0f), 360.0f) * (3.14159265358979323846f / 180.0f);
        sd.reverse_orientation = fs.reverse_orientation; sd.material = fs.material_id; sd.light = -1;
        sd.emit[0] = fs.emit[0]; sd.emit[1] = fs.emit[1]; sd.emit[2] = fs.emit[2];
        sd.area = sd.phi_max * sd.radius * (sd.z_max - sd.z_min);
        if (fs.emissive) {
            LightData ld; std::memset(&ld, 0, sizeof(ld)); ld.type = 1; ld.sphere = (int)i;
            ld.emit[0] = fs.emit[0]; ld.emit[1] = fs.emit[1]; ld.emit[2] = fs.emit[2];
            sd.light = (int)s->lights.size(); s->lights.push_back(ld);
        }
        s->spheres.push_back(sd);
    }
    *out = s;
    return FTN_OK;
}
SIM_API int sim_scene_destroy(SimScene* s) { delete s; return FTN_OK; }

// the device build pipeline of scene.cu as sequential loops over the same per-element bodies
SIM_API int sim_bvh_build(SimScene* s) {
    if (s->built) return FTN_OK;
    const uint32_t n = s->n_tris;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (n > 0) {
        std::vector<F4> tri_lo(n), tri_hi(n);
        float cmin[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmax[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (uint32_t i = 0; i < n; ++i) {   // k_tri_bounds
            float c[3]; tri_bounds_centroid(s->pos.data(), s->idx.data(), i, &tri_lo[i], &tri_hi[i], c);
            const float l[3] = {tri_lo[i].x, tri_lo[i].y, tri_lo[i].z}, h[3] = {tri_hi[i].x, tri_hi[i].y, tri_hi[i].z};
            for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], l[a]); hi[a] = fmaxf(hi[a], h[a]); cmin[a] = fminf(cmin[a], c[a]); cmax[a] = fmaxf(cmax[a], c[a]); }
        }
        s->codes.resize(n); s->order.resize(n);
        for (uint32_t i = 0; i < n; ++i) s->codes[i] = tri_morton(tri_lo[i], tri_hi[i], cmin, cmax);   // k_morton
        std::iota(s->order.begin(), s->order.end(), 0u);
        std::stable_sort(s->order.begin(), s->order.end(), [&](uint32_t a, uint32_t b) { return s->codes[a] < s->codes[b]; });   // radix_sort_pairs
        std::vector<uint32_t> keys(n);
        for (uint32_t i = 0; i < n; ++i) keys[i] = s->codes[s->order[i]];
        std::vector<F4> leaf_lo(n), leaf_hi(n);
        for (uint32_t i = 0; i < n; ++i) { leaf_lo[i] = tri_lo[s->order[i]]; leaf_hi[i] = tri_hi[s->order[i]]; }
        s->tris.resize((size_t)FTN_TRI_F4 * (size_t)n);
        std::vector<uint32_t> final_order = s->order;   // leaf order of the emitted tree (PLOC: depth-first order)
        const char* builder_env = getenv("FTN_BVH_BUILDER");
        const bool use_ploc = builder_env ? std::string(builder_env) == "ploc" : n >= 65536u;   // the policy of scene.cu
        const char* layout_env = getenv("FTN_BVH_LAYOUT");
        const bool wide = layout_env ? std::string(layout_env) != "bvh2" : n >= 65536u;   // the policy of scene.cu
        // k_bvh8_level of scene.cu, level by level, one node at a time
        struct HostAlloc {
            uint32_t next_nodes = 1, next_tris = 0;
            void operator()(uint32_t ni, uint32_t nt, uint32_t* cb, uint32_t* tb) { *cb = next_nodes; next_nodes += ni; *tb = next_tris; next_tris += nt; }
        };
        auto collapse_wide = [&](const LbvhArrays& a8, const F4* llo, const F4* lhi, uint32_t single_count) -> int {
            const size_t max_nodes = bvh8_max_nodes(n);
            std::vector<F4> nodes8(max_nodes * FTN_NODE8_F4); std::vector<uint32_t> brefs(max_nodes, 0u), order8(n, 0xFFFFFFFFu);
            HostAlloc alloc;
            uint32_t begin = 0, end = 1, levels = 0;
            while (begin < end) {
                ++levels;
                for (uint32_t w = begin; w < end; ++w) {
                    F4 blo, bhi;
                    if (single_count) { blo.x = lo[0]; blo.y = lo[1]; blo.z = lo[2]; blo.w = 0; bhi.x = hi[0]; bhi.y = hi[1]; bhi.z = hi[2]; bhi.w = 0; }
                    else bvh8_ref_box(a8, llo, lhi, brefs[w], &blo, &bhi);
                    bvh8_collapse_node(a8, llo, lhi, brefs[w], single_count, blo, bhi, w, alloc, nodes8.data(), brefs.data(), final_order.data(), order8.data());
                    if (alloc.next_nodes > max_nodes) return fail(FTN_ERR_CUDA, "bvh8 collapse: node bound exceeded");
                }
                begin = end; end = alloc.next_nodes;
            }
            if (alloc.next_tris != n) return fail(FTN_ERR_CUDA, "bvh8 collapse: triangle count mismatch");
            if (levels + 2u > (uint32_t)FTN_STACK8_SIZE) return fail(FTN_ERR_CUDA, "bvh8 collapse: too deep");
            for (uint32_t i = 0; i < n; ++i) if (order8[i] == 0xFFFFFFFFu) return fail(FTN_ERR_CUDA, "bvh8 collapse: triangle order has holes");
            nodes8.resize((size_t)alloc.next_nodes * FTN_NODE8_F4);
            s->nodes = nodes8; s->n_nodes = alloc.next_nodes; s->wide = true; s->bvh_levels = levels;
            final_order = order8;
            if (getenv("FTN_DEBUG_BUILD")) fprintf(stderr, "[sim] BVH8q: %u triangles, %u records in %u levels\n", n, s->n_nodes, levels);
            return FTN_OK;
        };
        if (wide && n <= (uint32_t)FTN_LEAF8_MAX) {
            LbvhArrays none; std::memset(&none, 0, sizeof(none));
            if (int rc = collapse_wide(none, leaf_lo.data(), leaf_hi.data(), n)) return rc;
            for (uint32_t i = 0; i < n; ++i) lbvh_gather_tri(s->pos.data(), s->idx.data(), final_order.data(), i, s->meshes.data(), (uint32_t)s->meshes.size(), s->tris.data());
        } else if (!wide && n <= (uint32_t)FTN_LEAF_MAX) {
            for (uint32_t i = 0; i < n; ++i) lbvh_gather_tri(s->pos.data(), s->idx.data(), final_order.data(), i, s->meshes.data(), (uint32_t)s->meshes.size(), s->tris.data());
            s->nodes.resize(FTN_NODE_F4); lbvh_emit_single(n, lo, hi, s->nodes.data()); s->n_nodes = 1;
        } else {
            const size_t ni = n - 1;
            std::vector<uint32_t> left(ni), right(ni), first(ni), last(ni), parent(2 * (size_t)n - 1, PLOC_NONE), arrive(ni, 0), survive(ni), new_index(ni);
            std::vector<F4> node_lo(ni), node_hi(ni);
            LbvhArrays a; a.left = left.data(); a.right = right.data(); a.first = first.data(); a.last = last.data();
            a.parent = parent.data(); a.arrive = arrive.data(); a.node_lo = node_lo.data(); a.node_hi = node_hi.data();
            bool ploc_done = false;
            if (use_ploc) {   // the kernels of scene.cu's ploc_build(), one element at a time
                std::vector<uint32_t> cl0(n), cl1(n), nn(n), mg(n), va(n), ms(n), vs(n);
                for (uint32_t i = 0; i < n; ++i) cl0[i] = LBVH_LEAF_FLAG | i;
                uint32_t c = n, created = 0, rounds = 0;
                uint32_t *cin = cl0.data(), *cout = cl1.data();
                while (c > 1) {
                    ++rounds;
                    for (uint32_t i = 0; i < c; ++i) nn[i] = ploc_nearest(a, leaf_lo.data(), leaf_hi.data(), cin, c, i);
                    uint32_t m = 0, v = 0;
                    for (uint32_t i = 0; i < c; ++i) { ploc_flags(nn.data(), i, mg.data(), va.data()); }
                    for (uint32_t i = 0; i < c; ++i) { ms[i] = m; vs[i] = v; m += mg[i]; v += va[i]; }
                    if (m == 0) return fail(FTN_ERR_CUDA, "PLOC round without a merge");
                    for (uint32_t i = 0; i < c; ++i) ploc_merge(a, leaf_lo.data(), leaf_hi.data(), cin, cout, nn.data(), mg.data(), va.data(), ms.data(), vs.data(), n, created, i);
                    created += m; c -= m; std::swap(cin, cout);
                }
                if (created != ni || cin[0] != 0u) return fail(FTN_ERR_CUDA, "PLOC did not end at root 0");
                const uint32_t ploc_rounds = rounds;
                std::vector<uint32_t> newpos(n); uint32_t max_depth = 0;
                for (uint32_t l = 0; l < n; ++l) { uint32_t d; newpos[l] = ploc_dfs_position(a, n, LBVH_LEAF_FLAG | l, &d); max_depth = std::max(max_depth, d); }
                if (getenv("FTN_DEBUG_BUILD")) fprintf(stderr, "[sim] PLOC: %u triangles, %u rounds, tree depth %u\n", n, ploc_rounds, max_depth);
                const char* depth_env = getenv("FTN_PLOC_MAX_DEPTH");
                const uint32_t depth_limit = depth_env ? (uint32_t)atoi(depth_env) : (uint32_t)FTN_STACK_SIZE - 4u;
                if (max_depth <= depth_limit) {
                    for (size_t i = 0; i < ni; ++i) { uint32_t d; first[i] = ploc_dfs_position(a, n, (uint32_t)i, &d); last[i] = first[i] + arrive[i] - 1u; }
                    std::vector<F4> l2(n), h2(n); std::vector<uint32_t> seen(n, 0);
                    for (uint32_t l = 0; l < n; ++l) { l2[newpos[l]] = leaf_lo[l]; h2[newpos[l]] = leaf_hi[l]; final_order[newpos[l]] = s->order[l]; seen[newpos[l]]++; }
                    for (uint32_t l = 0; l < n; ++l) if (seen[l] != 1u) return fail(FTN_ERR_CUDA, "PLOC leaf positions are not a permutation");
                    for (size_t i = 0; i < ni; ++i) {
                        if (left[i] & LBVH_LEAF_FLAG) left[i] = LBVH_LEAF_FLAG | newpos[left[i] & ~LBVH_LEAF_FLAG];
                        if (right[i] & LBVH_LEAF_FLAG) right[i] = LBVH_LEAF_FLAG | newpos[right[i] & ~LBVH_LEAF_FLAG];
                    }
                    leaf_lo = l2; leaf_hi = h2;
                    ploc_done = true;
                } else {
                    std::fill(parent.begin(), parent.end(), PLOC_NONE); std::fill(arrive.begin(), arrive.end(), 0u);
                }
            }
            if (!ploc_done) {
            for (size_t i = 0; i < ni; ++i) lbvh_topology_node(keys.data(), (int)n, (int)i, a);   // k_lbvh_topology
            for (uint32_t leaf = 0; leaf < n; ++leaf) {   // k_lbvh_refit: second arrival joins
                uint32_t node = parent[n - 1 + leaf];
                while (node != 0xFFFFFFFFu) {
                    if (arrive[node]++ == 0u) break;
                    lbvh_join_children(a, leaf_lo.data(), leaf_hi.data(), node);
                    node = parent[node];
                }
            }
            for (size_t i = 0; i < ni; ++i) { if (arrive[i] != 2u) return fail(FTN_ERR_CUDA, "refit did not reach every node twice"); }
            }
            if (wide) { if (int rc = collapse_wide(a, leaf_lo.data(), leaf_hi.data(), 0u)) return rc; }
            for (uint32_t i = 0; i < n; ++i) lbvh_gather_tri(s->pos.data(), s->idx.data(), final_order.data(), i, s->meshes.data(), (uint32_t)s->meshes.size(), s->tris.data());
            if (!wide) {
            uint32_t run = 0;
            for (size_t i = 0; i < ni; ++i) survive[i] = ploc_done ? ploc_survives(a, (int)i) : lbvh_survives(a, (int)i);   // k_lbvh_survive
            std::vector<uint32_t> is_record(ni);
            for (size_t i = 0; i < ni; ++i) { is_record[i] = lbvh_is_record(a, survive.data(), (int)i); new_index[i] = run; run += is_record[i]; }   // mark + scan
            s->n_nodes = run; s->nodes.resize((size_t)FTN_NODE_F4 * (size_t)run);
            for (size_t i = 0; i < ni; ++i) if (is_record[i]) lbvh_emit_node(a, leaf_lo.data(), leaf_hi.data(), survive.data(), new_index.data(), (int)i, s->nodes.data());
            }
        }
    }
    // sphere bounds + light preprocessing exactly as bvh_build() in scene.cu
    for (const SphereData& sd : s->spheres) {
        const float omin[3] = {-sd.radius, -sd.radius, sd.z_min}, omax[3] = {sd.radius, sd.radius, sd.z_max};
        for (int c = 0; c < 8; ++c) {
            const float p[3] = {(c & 4) ? omax[0] : omin[0], (c & 2) ? omax[1] : omin[1], (c & 1) ? omax[2] : omin[2]};
            const float* m = sd.o2w.m; float q[4];
            for (int r = 0; r < 4; ++r) q[r] = ((m[r] * p[0] + m[4 + r] * p[1]) + m[8 + r] * p[2]) + m[12 + r] * 1.0f;
            const float iw = 1.0f / q[3];
            for (int a = 0; a < 3; ++a) { const float v = q[a] * iw; lo[a] = std::fmin(lo[a], v); hi[a] = std::fmax(hi[a], v); }
        }
    }
    for (int c = 0; c < 3; ++c) { s->bounds[c] = lo[c]; s->bounds[3 + c] = hi[c]; }
    for (LightData& ld : s->lights) {
        if (ld.type != 0 && ld.type != 3) continue;
        float c[3]; for (int a = 0; a < 3; ++a) c[a] = (lo[a] + hi[a]) / 2.0f;
        const float dx = hi[0] - c[0], dy = hi[1] - c[1], dz = hi[2] - c[2];
        ld.env.world_radius = std::sqrt((dx * dx + dy * dy) + dz * dz);
        ld.env.world_center[0] = c[0]; ld.env.world_center[1] = c[1]; ld.env.world_center[2] = c[2];
    }
    s->built = true;
    return FTN_OK;
}

SIM_API int sim_bvh_debug_morton(const SimScene* s, uint32_t* codes, uint32_t* order) {
    if (codes) std::copy(s->codes.begin(), s->codes.end(), codes);
    if (order) std::copy(s->order.begin(), s->order.end(), order);
    return FTN_OK;
}
SIM_API int sim_scene_world_bound(const SimScene* s, float out[6]) { std::memcpy(out, s->bounds, 24); return FTN_OK; }
SIM_API int sim_scene_stats(const SimScene* s, FtnStats* st) {
    std::memset(st, 0, sizeof(*st)); st->bvh_nodes = s->n_nodes; st->bvh_node_bytes = s->wide ? FTN_NODE8_BYTES : FTN_NODE_BYTES; st->bvh_tri_bytes = FTN_TRI_BYTES; return FTN_OK;
}

static RayF to_rayf(const FtnRay& r) { RayF q; q.o = V3(r.o[0], r.o[1], r.o[2]); q.d = V3(r.d[0], r.d[1], r.d[2]); q.t_max = r.t_max; q.time = r.time; return q; }

// k_intersect_batch, one "thread" at a time
SIM_API int sim_intersect(const SimScene* s, size_t n, const FtnRay* rays, FtnHit* hits) {
    const SceneView sc = s->view();
    for (size_t i = 0; i < n; ++i) {
        const RayF ray = to_rayf(rays[i]);
        SceneHit h; TraceCounters tc; tc.nodes = tc.tris = 0;
        scene_intersect<false, false>(sc, ray, &h, &tc);
        FtnHit out;
        if (h.slot == FTN_NO_HIT_SLOT) { out.prim = FTN_NO_HIT; out.t = ray.t_max; out.b1 = 0; out.b2 = 0; }
        else if (h.slot & FTN_SPHERE_SLOT_FLAG) { out.prim = sc.n_tris + (h.slot & ~FTN_SPHERE_SLOT_FLAG); out.t = h.t; out.b1 = 0; out.b2 = 0; }
        else { out.prim = f2u(sc.bvh.tris[(size_t)FTN_TRI_F4 * (size_t)h.slot].w); out.t = h.t; out.b1 = h.tri.b1; out.b2 = h.tri.b2; }
        hits[i] = out;
    }
    return FTN_OK;
}
SIM_API int sim_intersect_test(const SimScene* s, size_t n, const FtnRay* rays, uint8_t* out) {
    const SceneView sc = s->view();
    for (size_t i = 0; i < n; ++i) {
        SceneHit h; TraceCounters tc; tc.nodes = tc.tris = 0;
        scene_intersect<true, false>(sc, to_rayf(rays[i]), &h, &tc);
        out[i] = h.slot != FTN_NO_HIT_SLOT;
    }
    return FTN_OK;
}
SIM_API int sim_intersect_count(const SimScene* s, size_t n, const FtnRay* rays, FtnHit* hits, uint64_t* counters) {
    const SceneView sc = s->view();
    for (size_t i = 0; i < n; ++i) {
        SceneHit h; TraceCounters tc; tc.nodes = tc.tris = 0;
        scene_intersect<false, true>(sc, to_rayf(rays[i]), &h, &tc);
        counters[0] += tc.nodes; counters[1] += tc.tris;
        (void)hits;
    }
    return FTN_OK;
}

SIM_API int sim_film_pixel_count(const FtnFilm* f, int32_t* w, int32_t* h) {
    FilmGeom g; if (film_geometry(f, &g) != FTN_OK) return fail(FTN_ERR_INVALID_ARGUMENT, "bad film");
    if (w) *w = g.crop_max[0] - g.crop_min[0];
    if (h) *h = g.crop_max[1] - g.crop_min[1];
    return FTN_OK;
}

// render_device() of render.cu with every kernel replaced by a loop over its per-path body.
SIM_API int sim_render(const SimScene* s, const FtnCamera* cam, const FtnFilm* film, const FtnSampler* smp,
                       const FtnIntegrator* integ, FtnPixel* out_pixels, FtnStats* stats) {
    if (smp->mode != FTN_SAMPLER_COUNTER) return fail(FTN_ERR_UNSUPPORTED, "counter sampler only");
    FilmGeom fg; if (film_geometry(film, &fg) != FTN_OK) return fail(FTN_ERR_INVALID_ARGUMENT, "bad film");
    const int fw = fg.crop_max[0] - fg.crop_min[0], fh = fg.crop_max[1] - fg.crop_min[1];
    const int sbw = fg.sb_max[0] - fg.sb_min[0], sbh = fg.sb_max[1] - fg.sb_min[1];
    const size_t n_spix = (size_t)sbw * sbh;
    const int n_samples = (smp->sample_begin < smp->samples_per_pixel) ? (smp->samples_per_pixel - smp->sample_begin + smp->sample_stride - 1) / smp->sample_stride : 0;
    int s_per_pass = (int)std::max<size_t>(1, ((size_t)1 << 16) / std::max<size_t>(1, n_spix));   // small passes: exercises multi-pass accumulation
    s_per_pass = std::min(s_per_pass, std::max(1, n_samples));
    const size_t P = n_spix * (size_t)s_per_pass;
    std::vector<float4> L(P); std::vector<float2> pfilm(P); std::vector<uint8_t> spill(n_spix);
    std::vector<float4> accum((size_t)fw * fh, make_float4(0, 0, 0, 0));
    const SceneView sc = s->view();
    uint32_t err = 0;
    uint64_t rays_closest = 0, rays_any = 0, camera_samples = 0;
    bool has_area = false; for (const LightData& l : s->lights) if (l.type == 1 || l.type == FTN_LIGHT_TYPE_TRIANGLE) has_area = true;
    PassParams pp; std::memset(&pp, 0, sizeof(pp));
    pp.film = fg; pp.cam = *cam; pp.seed_key = sampler_seed_key(smp->seed);
    pp.spp = smp->samples_per_pixel; pp.s_stride = smp->sample_stride;
    pp.integrator = integ->type; pp.max_depth = integ->max_depth; pp.rr_threshold = integ->rr_threshold;
    const int reach = (int)std::ceil(std::max(fg.radius[0], fg.radius[1]) + 0.5f);
    for (int done = 0; done < n_samples; done += s_per_pass) {
        const int sc_n = std::min(s_per_pass, n_samples - done);
        pp.s_first = smp->sample_begin + done * smp->sample_stride; pp.s_count = sc_n; pp.n_paths = (uint32_t)(n_spix * (size_t)sc_n);
        std::fill(spill.begin(), spill.end(), 0);
        for (uint32_t path = 0; path < pp.n_paths; ++path) {
            float fx, fy; bool spills;
            RayF ray = raygen_path(pp, path, &fx, &fy, &spills);   // k_raygen
            if (spills) spill[path / (uint32_t)pp.s_count] = 1;
            pfilm[path] = make_float2(fx, fy);
            V3 Lp = v3s(0.0f), beta = v3s(1.0f); uint32_t state = 0;
            RayDiff cd; cd.rx_o = cd.rx_d = cd.ry_o = cd.ry_d = v3s(0.0f);   // PathArrays::diff of the kernels
            ++camera_samples;
            for (int it = 0; it < integ->max_depth + 2 + 4096; ++it) {
                SceneHit h; TraceCounters tc; tc.nodes = tc.tris = 0;
                scene_intersect<false, false>(sc, ray, &h, &tc);   // k_extend
                ++rays_closest;
                if (h.slot == FTN_NO_HIT_SLOT) {   // k_shade_miss
                    const bool add = (pp.integrator == FTN_INTEGRATOR_DIRECT_LIGHTING) || ((state & FTN_STATE_BOUNCES) == 0u) || (state & FTN_STATE_SPECULAR);
                    if (add) Lp = Lp + beta * scene_env_radiance(sc, ray.d);
                    break;
                }
                ShadeOut o;
                const int material = hit_material(sc, h.slot);   // the queue k_extend would bin this path into
                const int mclass = material < 0 ? -1 : sc.materials[material].type;
                const RayDiff* carried = (state & FTN_STATE_HAS_DIFF) ? &cd : nullptr;
                switch (mclass) {   // k_shade<QUEUE>
                    case FTN_MATERIAL_MATTE: shade_surface<FTN_MATERIAL_MATTE>(sc, pp, path, ray, h.slot, state, beta, Lp, &o, &err, carried); break;
                    case FTN_MATERIAL_METAL: shade_surface<FTN_MATERIAL_METAL>(sc, pp, path, ray, h.slot, state, beta, Lp, &o, &err, carried); break;
                    case FTN_MATERIAL_PLASTIC: shade_surface<FTN_MATERIAL_PLASTIC>(sc, pp, path, ray, h.slot, state, beta, Lp, &o, &err, carried); break;
                    case FTN_MATERIAL_MIRROR: shade_surface<FTN_MATERIAL_MIRROR>(sc, pp, path, ray, h.slot, state, beta, Lp, &o, &err, carried); break;
                    case FTN_MATERIAL_GLASS: shade_surface<FTN_MATERIAL_GLASS>(sc, pp, path, ray, h.slot, state, beta, Lp, &o, &err, carried); break;
                    case FTN_CLASS_OREN_NAYAR: shade_surface<FTN_CLASS_OREN_NAYAR>(sc, pp, path, ray, h.slot, state, beta, Lp, &o, &err, carried); break;
                    default: shade_surface<-1>(sc, pp, path, ray, h.slot, state, beta, Lp, &o, &err, carried); break;
                }
                Lp = o.L;
                if (o.direct.has_shadow) {   // k_shadow
                    RayF sr; sr.o = o.direct.sh_o; sr.d = o.direct.sh_d; sr.t_max = rn_sub(1.0f, 0.0001f); sr.time = ray.time;
                    SceneHit sh; scene_intersect<true, false>(sc, sr, &sh, &tc); ++rays_any;
                    if (sh.slot == FTN_NO_HIT_SLOT) Lp = Lp + o.direct.sh_L;
                }
                if (o.direct.has_mis) {   // k_mis
                    RayF mr; mr.o = o.direct.mis_o; mr.d = o.direct.mis_d; mr.t_max = FTN_INF; mr.time = ray.time;
                    SceneHit mh;
                    if (has_area) scene_intersect<false, false>(sc, mr, &mh, &tc); else scene_intersect<true, false>(sc, mr, &mh, &tc);
                    ++rays_closest;
                    const V3 inc = mis_incident(sc, sc.lights[o.direct.mis_light], mr, mh.slot);
                    if (!is_black(inc)) Lp = Lp + o.direct.mis_w * inc;
                }
                if (!o.alive) break;
                if (o.has_diff) cd = o.diff;
                ray.o = o.next_o; ray.d = o.next_d; ray.t_max = FTN_INF; beta = o.beta; state = o.state;
            }
            L[path] = make_float4(Lp.x, Lp.y, Lp.z, 0.0f);
        }
        for (int i = 0; i < fw * fh; ++i) film_gather_pixel(pp, pfilm.data(), L.data(), spill.data(), i, reach, &accum[i], &err);   // k_film_accumulate
    }
    for (int i = 0; i < fw * fh; ++i) {   // k_film_resolve into a zeroed film
        float4 p = make_float4(0, 0, 0, 0);
        film_resolve_pixel(accum[i], &p);
        out_pixels[i].xyz[0] = p.x; out_pixels[i].xyz[1] = p.y; out_pixels[i].xyz[2] = p.z; out_pixels[i].filter_weight_sum = p.w;
    }
    if (stats) { std::memset(stats, 0, sizeof(*stats)); stats->camera_samples = camera_samples; stats->rays_closest = rays_closest; stats->rays_any = rays_any; stats->bvh_nodes = s->n_nodes; stats->bvh_node_bytes = FTN_NODE_BYTES; stats->bvh_tri_bytes = 48; }
    if (err & ERR_NAN) return fail(FTN_ERR_NAN_RADIANCE, "NaN radiance");
    if (err & ERR_UNSUPPORTED) return fail(FTN_ERR_UNSUPPORTED, "unsupported");
    return FTN_OK;
}

// ---- per-function hooks mirroring the oracle's orc_kat_* ------------------------------------------------------
SIM_API int sim_kat_triangle_intersect(const float p0[3], const float p1[3], const float p2[3], const FtnRay* r, float out[4]) {
    TriHit h; const RayF ray = to_rayf(*r);
    if (!triangle_intersect(V3(p0[0], p0[1], p0[2]), V3(p1[0], p1[1], p1[2]), V3(p2[0], p2[1], p2[2]), ray.o, make_ray_shear(ray.d), ray.t_max, &h)) return 0;
    out[0] = h.t; out[1] = h.b0; out[2] = h.b1; out[3] = h.b2; return 1;
}
SIM_API int sim_kat_sphere_intersect(const FtnSphere* fs, const FtnRay* r, float out[13]) {
    FtnSceneDesc d; std::memset(&d, 0, sizeof(d)); d.abi_version = FTN_ABI_VERSION; d.spheres = fs; d.n_spheres = 1;
    SimScene* s; sim_scene_create(&d, &s);
    SphereHit h; const bool hit = sphere_intersect(s->spheres[0], to_rayf(*r), &h);
    if (hit) {
        out[0] = h.t; out[1] = h.p.x; out[2] = h.p.y; out[3] = h.p.z; out[4] = h.p_err.x; out[5] = h.p_err.y; out[6] = h.p_err.z;
        out[7] = h.n.x; out[8] = h.n.y; out[9] = h.n.z; out[10] = h.wo.x; out[11] = h.wo.y; out[12] = h.wo.z;
    }
    delete s;
    return hit ? 1 : 0;
}
SIM_API void sim_kat_camera_ray(const FtnCamera* cam, float fx, float fy, float lx, float ly, float tu, FtnRay* out) {
    const RayF r = camera_ray(*cam, fx, fy, lx, ly, tu);
    out->o[0] = r.o.x; out->o[1] = r.o.y; out->o[2] = r.o.z; out->d[0] = r.d.x; out->d[1] = r.d.y; out->d[2] = r.d.z; out->t_max = r.t_max; out->time = r.time;
}
SIM_API void sim_kat_offset_ray_origin(const float p[3], const float e[3], const float n[3], const float d[3], float out[3]) {
    const V3 r = offset_ray_origin(V3(p[0], p[1], p[2]), V3(e[0], e[1], e[2]), V3(n[0], n[1], n[2]), V3(d[0], d[1], d[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
SIM_API void sim_kat_bsdf(const FtnMaterial* m, const float wo[3], const float wi[3], const float u[2], float out[12]) {
    FtnSceneDesc d; std::memset(&d, 0, sizeof(d)); d.abi_version = FTN_ABI_VERSION; d.materials = m; d.n_materials = 1;
    SimScene* s; if (sim_scene_create(&d, &s) != FTN_OK) { for (int k = 0; k < 12; ++k) out[k] = 0.0f; return; }
    Bsdf b; bsdf_init(&b, V3(0, 0, 1), V3(0, 0, 1), V3(1, 0, 0));
    const V3 o(wo[0], wo[1], wo[2]), i(wi[0], wi[1], wi[2]);
    V3 f; float pdf; ScatterSample sm; bool ok;
    const SceneView scv = s->view(); bool unsupported = false;
#define SIM_BSDF(M) { material_bsdf<M, true>(scv, s->mats[0], 0.0f, 0.0f, TexDiffs{0.0f, 0.0f, 0.0f, 0.0f}, &b, &unsupported); f = bsdf_f<M>(b, o, i, BXDF_ALL); pdf = bsdf_pdf<M>(b, o, i, BXDF_ALL); ok = bsdf_sample_f<M>(b, o, u[0], u[1], BXDF_ALL, &sm); }
    if (s->mats[0].type == FTN_CLASS_OREN_NAYAR) SIM_BSDF(FTN_CLASS_OREN_NAYAR)
    else if (m->type == FTN_MATERIAL_MATTE) SIM_BSDF(FTN_MATERIAL_MATTE)
    else if (m->type == FTN_MATERIAL_METAL) SIM_BSDF(FTN_MATERIAL_METAL)
    else if (m->type == FTN_MATERIAL_MIRROR) SIM_BSDF(FTN_MATERIAL_MIRROR)
    else if (m->type == FTN_MATERIAL_GLASS) SIM_BSDF(FTN_MATERIAL_GLASS)
    else SIM_BSDF(FTN_MATERIAL_PLASTIC)
#undef SIM_BSDF
    out[0] = f.x; out[1] = f.y; out[2] = f.z; out[3] = pdf;
    out[4] = ok ? 1.0f : 0.0f;
    if (ok) { out[5] = sm.f.x; out[6] = sm.f.y; out[7] = sm.f.z; out[8] = sm.wi.x; out[9] = sm.wi.y; out[10] = sm.wi.z; out[11] = sm.pdf; }
    else for (int k = 5; k < 12; ++k) out[k] = 0.0f;
    delete s;
}
SIM_API int sim_kat_env(const SimScene* s, const float u[2], float out[11]) {
    const EnvLightData& e = s->lights[0].env;
    V3 wi, L; float pdf;
    if (!env_sample(e, u[0], u[1], &wi, &pdf, &L)) return FTN_ERR_UNSUPPORTED;
    out[0] = wi.x; out[1] = wi.y; out[2] = wi.z; out[3] = pdf; out[4] = L.x; out[5] = L.y; out[6] = L.z;
    out[7] = env_pdf(e, wi);
    const V3 le = env_emitted(e, wi); out[8] = le.x; out[9] = le.y; out[10] = le.z;
    return FTN_OK;
}
// guided cdf search (search_sorted_le_guided over env_guide_entry) against the full binary search (sampling.rs:66-81):
// builds the Distribution1D of func[0..n) the way the device does and returns the number of u's whose index differs
SIM_API int sim_kat_guided_search(const float* func, int n, const float* u, int n_u, int* first_bad) {
    std::vector<float> cdf((size_t)n + 1); float integral;
    dist_row_build(func, n, cdf.data(), &integral);
    std::vector<uint32_t> guide((size_t)n + 1);
    for (int g = 0; g <= n; ++g) guide[g] = env_guide_entry(cdf.data(), n, g);
    int bad = 0;
    for (int k = 0; k < n_u; ++k) {
        const int a = search_sorted_le(cdf.data(), n + 1, u[k]), b = search_sorted_le_guided(cdf.data(), guide.data(), n, u[k]);
        if (a != b) { if (!bad && first_bad) *first_bad = k; ++bad; }
    }
    return bad;
}
// the device's mip_lookup_trilinear on a pyramid in the ABI's layout; (dsdx, 0, 0, 0) differentials give width = 2 |dsdx|
SIM_API int sim_kat_mipmap_lookup(const float* pyramid, int w, int h, int levels, int wrap, float s, float t, float width, float out[3]) {
    FtnMaterial fm; std::memset(&fm, 0, sizeof(fm));
    fm.image = pyramid; fm.image_width = w; fm.image_height = h; fm.image_levels = levels; fm.image_wrap = wrap;
    std::vector<F4> texels;
    if (!pack_image_pyramid(fm, &texels)) return FTN_ERR_INVALID_ARGUMENT;
    MaterialData md; std::memset(&md, 0, sizeof(md));
    md.image = texels.data(); md.img_w = w; md.img_h = h; md.img_levels = levels; md.img_wrap = wrap;
    const V3 v = mip_lookup_trilinear(md, s, t, 0.5f * width, 0.0f, 0.0f, 0.0f);
    out[0] = v.x; out[1] = v.y; out[2] = v.z;
    return FTN_OK;
}
SIM_API float sim_kat_counter_uniform(uint64_t seed, uint64_t sample_index, uint32_t dim) {
    return sampler_uniform(sampler_sample_key(sampler_seed_key(seed), sample_index), dim);
}
SIM_API float sim_kat_gamma(int n) {
    switch (n) { case 1: return gamma_n(1); case 2: return gamma_n(2); case 3: return gamma_n(3); case 4: return gamma_n(4);
                 case 5: return gamma_n(5); case 6: return gamma_n(6); case 7: return gamma_n(7); default: return gamma_n(8); }
}

// ---- SIMT execution model of the persistent traversal loop (design tool, not a test) -----------------
// Replays ftn_trace_persistent.cuh's loop structure for warps of 32 lanes over a ray batch and
// counts ISSUE SLOTS per phase (one slot = one warp-wide pass through a phase body, whatever the
// number of participating lanes) next to the lane-level work, for different loop policies.
//   ip[0] refill threshold   ip[1] node-phase exit: leave when fewer than this many lanes want a node
//   ip[2] leaf-phase exit: leave when fewer than this many lanes still hold a leaf (1 = drain)
//   ip[3] postponed leaves a lane may hold before it must stop walking (0, 1)
//   ip[4] warps simulated round-robin
// out: [0] node slots [1] node lane-steps [2] triangle slots [3] triangle lane-tests
//      [4] leaf-phase entries (slots) [5] leaf lane-entries [6] refill rounds [7] rays
SIM_API int sim_warp_model(const SimScene* s, size_t n, const FtnRay* rays, const int* ip, double* out) {
    const SceneView sc = s->view();
    const BvhView bvh = sc.bvh;
    const int refill_thresh = ip[0], node_exit = ip[1], leaf_exit = ip[2], postpone = ip[3], n_warps = std::max(1, ip[4]);
    struct Lane {
        bool has_ray = false, finished = false;
        RayF ray; RaySlab slab; RayShear shear; float t_max = 0; uint32_t slot = FTN_NO_HIT_SLOT; TriHit tri;
        int stack[FTN_STACK_SIZE]; int sp = 0, cur = FTN_TRAVERSAL_DONE, leaf = 0;
        uint32_t tri_i = 0;   // progress inside the current leaf
    };
    struct Warp { Lane l[32]; bool done = false; };
    std::vector<Warp> warps(n_warps);
    size_t next = 0;
    double node_slots = 0, node_work = 0, tri_slots = 0, tri_work = 0, leaf_slots = 0, leaf_work = 0, refills = 0;
    TraceCounters tc; tc.nodes = tc.tris = 0;
    int live = n_warps;
    while (live > 0) {
        for (Warp& w : warps) {
            if (w.done) continue;
            // flush + refill
            int idle = 0;
            for (Lane& L : w.l) { if (L.finished) { L.has_ray = false; L.finished = false; } if (!L.has_ray) ++idle; }
            if (idle && next < n) {
                refills += 1;
                for (Lane& L : w.l) {
                    if (L.has_ray || next >= n) continue;
                    L.ray = to_rayf(rays[next++]); L.has_ray = true; L.slot = FTN_NO_HIT_SLOT; L.t_max = L.ray.t_max;
                    L.slab = make_ray_slab(L.ray.o, L.ray.d); L.shear = make_ray_shear(L.ray.d); L.sp = 0; L.cur = 0; L.leaf = 0;
                }
            }
            int with_ray = 0; for (Lane& L : w.l) with_ray += L.has_ray;
            if (!with_ray) { w.done = true; --live; continue; }
            const int thresh = (next >= n) ? 1 : refill_thresh;
            auto promote = [&]() {   // a lane stopped on a leaf with a free postponed slot takes it and pops
                for (Lane& L : w.l) {
                    if (!(L.has_ray && !L.finished)) continue;
                    if (L.leaf >= 0 && L.cur < 0 && L.cur != FTN_TRAVERSAL_DONE) { L.leaf = L.cur; L.cur = (L.sp > 0) ? L.stack[--L.sp] : FTN_TRAVERSAL_DONE; }
                }
            };
            auto count_node = [&]() { int c = 0; for (Lane& L : w.l) c += (L.has_ray && !L.finished && L.cur >= 0); return c; };
            auto count_leaf = [&]() { int c = 0; for (Lane& L : w.l) c += (L.has_ray && !L.finished && L.leaf < 0); return c; };
            auto node_slot = [&]() {
                node_slots += 1;
                for (Lane& L : w.l) {
                    if (!(L.has_ray && !L.finished && L.cur >= 0)) continue;
                    node_work += 1;
                    L.cur = node_step(bvh, L.cur, L.slab, L.t_max, L.stack, L.sp);
                    if (postpone && L.cur < 0 && L.cur != FTN_TRAVERSAL_DONE && L.leaf >= 0) {
                        L.leaf = L.cur; L.cur = (L.sp > 0) ? L.stack[--L.sp] : FTN_TRAVERSAL_DONE;
                    }
                }
            };
            auto leaf_slot = [&]() {
                leaf_slots += 1;
                uint32_t maxc = 0;
                for (Lane& L : w.l) if (L.has_ray && !L.finished && L.leaf < 0) maxc = std::max(maxc, ((~(uint32_t)L.leaf) & 3u) + 1u);
                tri_slots += maxc;
                for (Lane& L : w.l) {
                    if (!(L.has_ray && !L.finished && L.leaf < 0)) continue;
                    leaf_work += 1;
                    tri_work += ((~(uint32_t)L.leaf) & 3u) + 1u;
                    leaf_step<false, false>(bvh, L.leaf, L.ray.o, L.shear, &L.t_max, &L.slot, &L.tri, &tc);
                    L.leaf = 0;
                }
            };
            for (;;) {
                if (ip[5] == 2) {
                    // per-step vote, ONE triangle per leaf slot (lanes keep a cursor into their leaf)
                    for (;;) {
                        promote();
                        const int nn = count_node(), nl = count_leaf();
                        if (nn == 0 && nl == 0) break;
                        if (nl == 0 || nn * 16 >= nl * ip[6]) node_slot();
                        else {
                            leaf_slots += 1; tri_slots += 1;
                            for (Lane& L : w.l) {
                                if (!(L.has_ray && !L.finished && L.leaf < 0)) continue;
                                leaf_work += 1; tri_work += 1;
                                const uint32_t ref = ~(uint32_t)L.leaf, first = ref >> 2, cnt = ref & 3u;
                                const int one = (int)~((first << 2) | 0u);
                                leaf_step<false, false>(bvh, one, L.ray.o, L.shear, &L.t_max, &L.slot, &L.tri, &tc);
                                L.leaf = cnt ? (int)~(((first + 1) << 2) | (cnt - 1)) : 0;
                            }
                        }
                        int active = 0;
                        for (Lane& L : w.l) active += (L.has_ray && !L.finished && !(L.cur == FTN_TRAVERSAL_DONE && L.leaf >= 0));
                        if (active < thresh) break;
                    }
                } else if (ip[5] == 1) {
                    // per-step vote: run the phase the (weighted) majority of lanes wants
                    for (;;) {
                        promote();
                        const int nn = count_node(), nl = count_leaf();
                        if (nn == 0 && nl == 0) break;
                        if (nl == 0 || nn * 16 >= nl * ip[6]) node_slot(); else leaf_slot();
                        int active = 0;
                        for (Lane& L : w.l) active += (L.has_ray && !L.finished && !(L.cur == FTN_TRAVERSAL_DONE && L.leaf >= 0));
                        if (active < thresh) break;
                    }
                } else {
                    // phase 1: nodes
                    for (;;) {
                        const int want = count_node();
                        if (want == 0) break;
                        promote();
                        const int nl = count_leaf();
                        if (want < node_exit && nl > want) break;
                        node_slot();
                    }
                    // phase 2: leaves
                    for (;;) {
                        promote();
                        const int holding = count_leaf();
                        if (holding == 0) break;
                        if (holding < leaf_exit && count_node() > holding) break;
                        leaf_slot();
                    }
                }
                int active = 0;
                for (Lane& L : w.l) {
                    if (L.has_ray && !L.finished && L.cur == FTN_TRAVERSAL_DONE && L.leaf >= 0) L.finished = true;
                    active += (L.has_ray && !L.finished);
                }
                if (active < thresh) break;
            }
        }
    }
    out[0] = node_slots; out[1] = node_work; out[2] = tri_slots; out[3] = tri_work; out[4] = leaf_slots; out[5] = leaf_work; out[6] = refills; out[7] = (double)n;
    return FTN_OK;
}
// ---- the same design tool for the BVH8q loop (trace_persistent8) ----------------------------------------------------
//   ip[0] refill threshold   ip[1] vote bias (node step when 16 * #node lanes >= bias * #leaf lanes)   ip[2] warps
//   ip[3] leaf granularity: 0 = one leaf child (<= 3 triangles) per leaf slot, 1 = one triangle per leaf slot
//   ip[4] postponing: 0 = a lane holding a triangle group waits for a leaf step; 1 = when the warp runs a node step such a
//         lane pushes its triangle group on its stack and walks on (groups popped later are tested then)
//   ip[5] cooperative leaf step: 1 = the pending (lane, triangle) pairs of the warp are spread over all 32 lanes
// out: [0] node slots [1] node lane-steps [2] triangle slots [3] triangle lane-tests [4] loop iterations [5] refill rounds
//      [6] rays [7] max stack depth
SIM_API int sim_warp_model8(const SimScene* s, size_t n, const FtnRay* rays, const int* ip, double* out) {
    const SceneView sc = s->view();
    const BvhView bvh = sc.bvh;
    if (!bvh.wide) return FTN_ERR_INVALID_ARGUMENT;
    const int refill_thresh = ip[0], bias = ip[1], n_warps = std::max(1, ip[2]), one_tri = ip[3], postpone = ip[4], coop = ip[5];
    struct Entry { uint32_t base, bits; bool tri; };
    struct Lane {
        bool has_ray = false, finished = false;
        RayF ray; Ray8 r8; RayShear shear; float t_max = 0; uint32_t slot = FTN_NO_HIT_SLOT; TriHit tri;
        std::vector<Entry> stack;
        uint32_t ng_base = 0, ng_bits = 0, tg_base = 0, tg_bits = 0;
        uint32_t cur_first = 0, cur_count = 0;   // one_tri: the leaf child being tested
    };
    struct Warp { Lane l[32]; bool done = false; };
    std::vector<Warp> warps(n_warps);
    size_t next = 0;
    double node_slots = 0, node_work = 0, tri_slots = 0, tri_work = 0, iters = 0, refills = 0, max_sp = 0;
    TraceCounters tc; tc.nodes = tc.tris = 0;
    int live = n_warps;
    auto pending_tris = [&](const Lane& L) { return (L.tg_bits & 0xFFu) != 0u || L.cur_count != 0u; };
    auto normalise = [&](Lane& L) {   // nothing left in either group: take the next group from the stack
        while (!pending_tris(L) && !(L.ng_bits & 0xFFu) && !L.stack.empty()) {
            const Entry e = L.stack.back(); L.stack.pop_back();
            if (e.tri) { L.tg_base = e.base; L.tg_bits = e.bits; } else { L.ng_base = e.base; L.ng_bits = e.bits; }
        }
    };
    while (live > 0) {
        for (Warp& w : warps) {
            if (w.done) continue;
            int idle = 0;
            for (Lane& L : w.l) { if (L.finished) { L.has_ray = false; L.finished = false; } if (!L.has_ray) ++idle; }
            if (idle && next < n) {
                refills += 1;
                for (Lane& L : w.l) {
                    if (L.has_ray || next >= n) continue;
                    L.ray = to_rayf(rays[next++]); L.has_ray = true; L.slot = FTN_NO_HIT_SLOT; L.t_max = L.ray.t_max;
                    L.r8 = make_ray8(L.ray.o, L.ray.d); L.shear = make_ray_shear(L.ray.d); L.stack.clear();
                    L.ng_base = 0; L.ng_bits = (1u << L.r8.octinv) | (1u << 8); L.tg_bits = 0; L.cur_count = 0;
                }
            }
            int with_ray = 0; for (Lane& L : w.l) with_ray += L.has_ray;
            if (!with_ray) { w.done = true; --live; continue; }
            const int thresh = (next >= n) ? 1 : refill_thresh;
            for (;;) {
                int nn = 0, nl = 0;
                for (Lane& L : w.l) { if (!L.has_ray || L.finished) continue; if (pending_tris(L)) ++nl; else if (L.ng_bits & 0xFFu) ++nn; }
                if (nn + nl < thresh || nn + nl == 0) break;
                iters += 1;
                const bool node_phase = nl == 0 || 16 * nn >= bias * nl;
                if (node_phase) {
                    node_slots += 1;
                    for (Lane& L : w.l) {
                        if (!L.has_ray || L.finished) continue;
                        if (pending_tris(L)) {
                            if (!postpone || L.cur_count != 0u || !(L.ng_bits & 0xFFu)) continue;
                            Entry e; e.base = L.tg_base; e.bits = L.tg_bits; e.tri = true; L.stack.push_back(e); L.tg_bits = 0;
                        }
                        if (!(L.ng_bits & 0xFFu)) continue;
                        node_work += 1;
                        const uint32_t node = node8_pop_child(L.ng_base, L.ng_bits, L.r8.octinv);
                        if (L.ng_bits & 0xFFu) { Entry e; e.base = L.ng_base; e.bits = L.ng_bits; e.tri = false; L.stack.push_back(e); }
                        max_sp = std::max(max_sp, (double)L.stack.size());
                        const Node8Hits h = node8_test(bvh.nodes, node, L.r8, L.t_max);
                        L.ng_base = h.child_base; L.ng_bits = h.ng_bits; L.tg_base = h.tri_base; L.tg_bits = h.tg_bits;
                    }
                } else {
                    uint32_t max_tests = 0, all_tests = 0;
                    for (Lane& L : w.l) {
                        if (!L.has_ray || L.finished || !pending_tris(L)) continue;
                        if (L.cur_count == 0u) node8_pop_leaf(L.tg_base, L.tg_bits, L.r8.octinv, &L.cur_first, &L.cur_count);
                        const uint32_t k = one_tri ? 1u : L.cur_count;
                        tris8_test<false, false>(bvh, L.cur_first, k, L.ray.o, L.shear, &L.t_max, &L.slot, &L.tri, &tc);
                        L.cur_first += k; L.cur_count -= k;
                        tri_work += k; max_tests = std::max(max_tests, k); all_tests += k;
                    }
                    tri_slots += coop ? (all_tests + 31u) / 32u : max_tests;
                }
                for (Lane& L : w.l) {
                    if (!L.has_ray || L.finished) continue;
                    normalise(L);
                    if (!pending_tris(L) && !(L.ng_bits & 0xFFu)) L.finished = true;
                }
            }
        }
    }
    out[0] = node_slots; out[1] = node_work; out[2] = tri_slots; out[3] = tri_work; out[4] = iters; out[5] = refills; out[6] = (double)n; out[7] = max_sp;
    return FTN_OK;
}

// slab test as the kernels run it: mode 0 = per-ray dispatch (nan_free fast form when allowed), 1 = exact form.
// returns bit0 = accepted, bit1 = ray is nan_free; out[0] = entry distance
SIM_API int sim_kat_slab_test(const float bmin[3], const float bmax[3], const FtnRay* ray, float out[1], int mode) {
    const RaySlab s = make_ray_slab(V3(ray->o[0], ray->o[1], ray->o[2]), V3(ray->d[0], ray->d[1], ray->d[2]));
    float e = 0.0f;
    const bool hit = (mode == 0 && s.nan_free) ? slab_test<true>(s, bmin[0], bmax[0], bmin[1], bmax[1], bmin[2], bmax[2], ray->t_max, &e)
                                               : slab_test<false>(s, bmin[0], bmax[0], bmin[1], bmax[1], bmin[2], bmax[2], ray->t_max, &e);
    out[0] = e;
    return (hit ? 1 : 0) | (s.nan_free ? 2 : 0);
}

// ---- design experiment: what would a SAH-quality tree buy?  (binned top-down SAH, same node format) ----
namespace {
struct SahBuilder {
    const SimScene* s; std::vector<F4> lo, hi; std::vector<float> cx[3];
    std::vector<uint32_t> prims;      // permutation being partitioned
    std::vector<F4> nodes; std::vector<uint32_t> order_out;
    int leaf_max; float c_trav, c_isect;
    static float area(const float* l, const float* h) { float dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2]; return 2.0f * (dx * dy + dy * dz + dz * dx); }
    void bounds(uint32_t b, uint32_t e, float* l, float* h) const {
        for (int a = 0; a < 3; ++a) { l[a] = FLT_MAX; h[a] = -FLT_MAX; }
        for (uint32_t i = b; i < e; ++i) { const F4& L = lo[prims[i]]; const F4& H = hi[prims[i]];
            l[0] = fminf(l[0], L.x); l[1] = fminf(l[1], L.y); l[2] = fminf(l[2], L.z); h[0] = fmaxf(h[0], H.x); h[1] = fmaxf(h[1], H.y); h[2] = fmaxf(h[2], H.z); }
    }
    // returns child reference (>= 0 node index, < 0 leaf)
    int build(uint32_t b, uint32_t e) {
        const uint32_t n = e - b;
        float bl[3], bh[3]; bounds(b, e, bl, bh);
        int best_axis = -1; uint32_t best_mid = 0; float best_cost = FLT_MAX;
        if (n > 1) {
            const int NB = 32;
            for (int ax = 0; ax < 3; ++ax) {
                float cmin = FLT_MAX, cmax = -FLT_MAX;
                for (uint32_t i = b; i < e; ++i) { cmin = fminf(cmin, cx[ax][prims[i]]); cmax = fmaxf(cmax, cx[ax][prims[i]]); }
                if (!(cmax > cmin)) continue;
                struct Bin { float l[3], h[3]; uint32_t c; } bins[NB];
                for (auto& bn : bins) { bn.c = 0; for (int a = 0; a < 3; ++a) { bn.l[a] = FLT_MAX; bn.h[a] = -FLT_MAX; } }
                const float scale = NB / (cmax - cmin);
                for (uint32_t i = b; i < e; ++i) {
                    int k = std::min(NB - 1, (int)((cx[ax][prims[i]] - cmin) * scale));
                    Bin& bn = bins[k]; bn.c++; const F4& L = lo[prims[i]]; const F4& H = hi[prims[i]];
                    bn.l[0] = fminf(bn.l[0], L.x); bn.l[1] = fminf(bn.l[1], L.y); bn.l[2] = fminf(bn.l[2], L.z);
                    bn.h[0] = fmaxf(bn.h[0], H.x); bn.h[1] = fmaxf(bn.h[1], H.y); bn.h[2] = fmaxf(bn.h[2], H.z);
                }
                float ra[NB]; uint32_t rc[NB]; float l[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, h[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}; uint32_t c = 0;
                for (int k = NB - 1; k > 0; --k) { for (int a = 0; a < 3; ++a) { l[a] = fminf(l[a], bins[k].l[a]); h[a] = fmaxf(h[a], bins[k].h[a]); } c += bins[k].c; ra[k] = c ? area(l, h) : 0.0f; rc[k] = c; }
                for (int a = 0; a < 3; ++a) { l[a] = FLT_MAX; h[a] = -FLT_MAX; } c = 0;
                for (int k = 0; k < NB - 1; ++k) {
                    for (int a = 0; a < 3; ++a) { l[a] = fminf(l[a], bins[k].l[a]); h[a] = fmaxf(h[a], bins[k].h[a]); } c += bins[k].c;
                    if (c == 0 || rc[k + 1] == 0) continue;
                    const float cost = area(l, h) * c + ra[k + 1] * rc[k + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = ax; best_mid = (uint32_t)k; }
                }
                if (best_axis == ax) {   // remember the split plane for this axis
                    split_plane[ax] = cmin + (best_mid + 1) / scale;
                }
            }
        }
        const float leaf_cost = c_isect * n;
        const float split_cost = best_axis >= 0 ? c_trav + c_isect * best_cost / fmaxf(area(bl, bh), 1e-30f) : FLT_MAX;
        if (n <= (uint32_t)leaf_max && (best_axis < 0 || leaf_cost <= split_cost)) {
            const uint32_t first = (uint32_t)order_out.size();
            for (uint32_t i = b; i < e; ++i) order_out.push_back(prims[i]);
            return (int)~((first << 2) | (n - 1));
        }
        uint32_t mid;
        if (best_axis < 0) mid = b + n / 2;   // coincident centroids: split in the middle
        else {
            const float plane = split_plane[best_axis]; const int ax = best_axis;
            mid = (uint32_t)(std::partition(prims.begin() + b, prims.begin() + e, [&](uint32_t p) { return cx[ax][p] < plane; }) - prims.begin());
            if (mid == b || mid == e) mid = b + n / 2;
        }
        const size_t me = nodes.size() / 4; nodes.resize(nodes.size() + 4);
        float l0[3], h0[3], l1[3], h1[3]; bounds(b, mid, l0, h0); bounds(mid, e, l1, h1);
        const int c0 = build(b, mid), c1 = build(mid, e);
        F4 n0, n1, nz, ci;
        n0.x = l0[0]; n0.y = h0[0]; n0.z = l0[1]; n0.w = h0[1]; n1.x = l1[0]; n1.y = h1[0]; n1.z = l1[1]; n1.w = h1[1];
        nz.x = l0[2]; nz.y = h0[2]; nz.z = l1[2]; nz.w = h1[2]; ci.x = u2f((uint32_t)c0); ci.y = u2f((uint32_t)c1); ci.z = 0; ci.w = 0;
        nodes[4 * me] = n0; nodes[4 * me + 1] = n1; nodes[4 * me + 2] = nz; nodes[4 * me + 3] = ci;
        return (int)me;
    }
    float split_plane[3];
};
}  // namespace
SIM_API int sim_bvh_rebuild_sah(SimScene* s, int leaf_max, float c_trav, float c_isect) {
    const uint32_t n = s->n_tris;
    if (n <= (uint32_t)FTN_LEAF_MAX) return FTN_OK;
    SahBuilder B; B.s = s; B.leaf_max = leaf_max; B.c_trav = c_trav; B.c_isect = c_isect;
    B.lo.resize(n); B.hi.resize(n); for (int a = 0; a < 3; ++a) B.cx[a].resize(n);
    for (uint32_t i = 0; i < n; ++i) { float c[3]; tri_bounds_centroid(s->pos.data(), s->idx.data(), i, &B.lo[i], &B.hi[i], c); for (int a = 0; a < 3; ++a) B.cx[a][i] = c[a]; }
    B.prims.resize(n); std::iota(B.prims.begin(), B.prims.end(), 0u);
    B.nodes.reserve(4 * (size_t)n);
    const int root = B.build(0, n);
    if (root != 0) return fail(FTN_ERR_CUDA, "sah root");
    s->nodes = B.nodes; s->n_nodes = (uint32_t)(B.nodes.size() / 4);
    s->order = B.order_out;
    for (uint32_t i = 0; i < n; ++i) lbvh_gather_tri(s->pos.data(), s->idx.data(), s->order.data(), i, s->meshes.data(), (uint32_t)s->meshes.size(), s->tris.data());
    return FTN_OK;
}

// ---- design experiment: SAH tree rotations (Kensler 2008) on the emitted BVH2x64 records --------------------
// One bottom-up pass (children carry larger indices than their parents in both builders): at node X = (A, B)
// with A = (A0, A1) interior, swapping B with A0 or A1 (or symmetrically a child of B with A) is applied when it
// shrinks the surface area of the rebuilt child.  Returns the number of rotations applied.
namespace {
struct Box6 { float lx, hx, ly, hy, lz, hz; };
inline float area6(const Box6& b) { const float dx = b.hx - b.lx, dy = b.hy - b.ly, dz = b.hz - b.lz; return 2.0f * (dx * dy + dy * dz + dz * dx); }
inline Box6 join6(const Box6& a, const Box6& b) { return {fminf(a.lx, b.lx), fmaxf(a.hx, b.hx), fminf(a.ly, b.ly), fmaxf(a.hy, b.hy), fminf(a.lz, b.lz), fmaxf(a.hz, b.hz)}; }
inline Box6 child_box(const F4* nd, int k) {
    return k == 0 ? Box6{nd[0].x, nd[0].y, nd[0].z, nd[0].w, nd[2].x, nd[2].y} : Box6{nd[1].x, nd[1].y, nd[1].z, nd[1].w, nd[2].z, nd[2].w};
}
inline void set_child(F4* nd, int k, const Box6& b, int ref) {
    if (k == 0) { nd[0].x = b.lx; nd[0].y = b.hx; nd[0].z = b.ly; nd[0].w = b.hy; nd[2].x = b.lz; nd[2].y = b.hz; nd[3].x = u2f((uint32_t)ref); }
    else { nd[1].x = b.lx; nd[1].y = b.hx; nd[1].z = b.ly; nd[1].w = b.hy; nd[2].z = b.lz; nd[2].w = b.hz; nd[3].y = u2f((uint32_t)ref); }
}
inline int child_ref(const F4* nd, int k) { return (int)f2u(k == 0 ? nd[3].x : nd[3].y); }
}  // namespace
SIM_API int sim_bvh_rotate(SimScene* s) {
#if FTN_BVH_WIDTH == 2
    int applied = 0;
    for (int x = (int)s->n_nodes - 1; x >= 0; --x) {
        F4* X = s->nodes.data() + 4 * (size_t)x;
        if (child_ref(X, 1) == FTN_TRAVERSAL_DONE) continue;
        float best = 0.0f; int best_side = -1, best_g = -1;
        for (int side = 0; side < 2; ++side) {          // side: the child of X that is rebuilt (A); the other one (B) moves down
            const int a = child_ref(X, side);
            if (a < 0) continue;
            const F4* A = s->nodes.data() + 4 * (size_t)a;
            const Box6 bB = child_box(X, 1 - side), bA = child_box(X, side);
            for (int g = 0; g < 2; ++g) {               // g: the grandchild that moves up
                const Box6 keep = child_box(A, 1 - g);
                const float delta = area6(join6(bB, keep)) - area6(bA);
                if (delta < best) { best = delta; best_side = side; best_g = g; }
            }
        }
        if (best_side < 0) continue;
        const int a = child_ref(X, best_side);
        F4* A = s->nodes.data() + 4 * (size_t)a;
        const Box6 bB = child_box(X, 1 - best_side); const int rB = child_ref(X, 1 - best_side);
        const Box6 bG = child_box(A, best_g); const int rG = child_ref(A, best_g);
        const Box6 keep = child_box(A, 1 - best_g);
        set_child(A, best_g, bB, rB);                   // A = (B, kept grandchild)
        set_child(X, best_side, join6(bB, keep), a);    // X's box of A
        set_child(X, 1 - best_side, bG, rG);            // the grandchild takes B's place
        ++applied;
    }
    return applied;
#else
    (void)s; return 0;
#endif
}
