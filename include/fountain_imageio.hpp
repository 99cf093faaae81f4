// Host-side image ingestion for the C++ host (include/fountain_host.hpp): what src/imageio/mod.rs and
// src/imageio/exr.rs do between a file name and the texels of a MIPMap / an InfiniteAreaLight.
// Stays on the host side of the C ABI (the device receives texels).  Needs zlib (-lz).
//
//   load_image         imageio/mod.rs:127-150   .exr / .pfm / .png -> RGB f32, width, height
//   load_texels        imageio/mod.rs:82-125    gamma (default: everything but .exr / .pfm), scale, flip_y
//   read_exr           imageio/exr.rs:11-46     first layer, R / G / B as f16 or f32; scan lines, NO / RLE / ZIPS / ZIP
//   read_png                                    8-bit RGB / RGBA, non-interlaced (what the reference's match accepts:
//                                               ImageRgb8 / ImageRgba8, everything else is `unimplemented!()`)
//   make_infinite_area_light   loaders/constructors.rs:339-359
// The reference decodes through the third-party `image` and `exr` crates; these are the formats its own code path accepts.
#pragma once
#include <zlib.h>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include "fountain_host.hpp"

namespace fountain {
namespace imageio {

struct Image { std::vector<float> rgb; int width = 0, height = 0; };

inline std::vector<uint8_t> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: cannot open " + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
inline std::vector<uint8_t> inflate_all(const uint8_t* src, size_t n, size_t expect) {
    std::vector<uint8_t> out(expect);
    uLongf len = (uLongf)expect;
    if (uncompress(out.data(), &len, src, (uLong)n) != Z_OK) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: corrupt zlib stream");
    out.resize(len);
    return out;
}
inline float half_to_float(uint16_t h) {
    const uint32_t s = (h >> 15) & 1u, e = (h >> 10) & 31u, m = h & 1023u;
    uint32_t bits;
    if (e == 0) {
        if (m == 0) bits = s << 31;
        else { int k = 0; uint32_t mm = m; while (!(mm & 1024u)) { mm <<= 1; ++k; } bits = (s << 31) | ((uint32_t)(113 - k) << 23) | ((mm & 1023u) << 13); }
    } else if (e == 31) bits = (s << 31) | 0x7F800000u | (m << 13);
    else bits = (s << 31) | ((e + 112u) << 23) | (m << 13);
    float f; std::memcpy(&f, &bits, 4); return f;
}

// ---- OpenEXR, single-part scan-line ---------------------------------------------------------------------------
inline Image read_exr(const std::string& path) {
    const std::vector<uint8_t> d = read_file(path);
    auto u32 = [&](size_t p) { uint32_t v; std::memcpy(&v, &d.at(p + 3) - 3, 4); return v; };
    auto i32 = [&](size_t p) { return (int32_t)u32(p); };
    if (d.size() < 16 || u32(0) != 20000630u) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: " + path + " is not an OpenEXR file");
    if (u32(4) & 0x1A00u) throw Error(FTN_ERR_UNSUPPORTED, "fountain: tiled / multi-part / deep OpenEXR files are not read");
    size_t pos = 8;
    std::map<std::string, std::vector<uint8_t>> attrs;
    while (d.at(pos) != 0) {
        std::string name((const char*)&d[pos]); pos += name.size() + 1;
        std::string type((const char*)&d[pos]); pos += type.size() + 1;
        const int32_t size = i32(pos); pos += 4;
        attrs[name] = std::vector<uint8_t>(d.begin() + pos, d.begin() + pos + size); pos += size;
    }
    ++pos;
    struct Chan { std::string name; int type; };
    std::vector<Chan> chans;
    const std::vector<uint8_t>& cl = attrs.at("channels");
    for (size_t cp = 0; cl.at(cp) != 0;) {
        Chan c; c.name = (const char*)&cl[cp]; cp += c.name.size() + 1;
        int32_t t, xs, ys; std::memcpy(&t, &cl[cp], 4); std::memcpy(&xs, &cl[cp + 8], 4); std::memcpy(&ys, &cl[cp + 12], 4); cp += 16;
        if (xs != 1 || ys != 1) throw Error(FTN_ERR_UNSUPPORTED, "fountain: subsampled OpenEXR channels");
        c.type = t; chans.push_back(c);
    }
    const int comp = attrs.at("compression").at(0);
    const int lines = comp == 3 ? 16 : 1;
    if (comp < 0 || comp > 3) throw Error(FTN_ERR_UNSUPPORTED, "fountain: OpenEXR compression other than NO / RLE / ZIPS / ZIP");
    int32_t win[4]; std::memcpy(win, attrs.at("dataWindow").data(), 16);
    const int w = win[2] - win[0] + 1, h = win[3] - win[1] + 1;
    auto px_bytes = [](int t) { return t == 1 ? 2 : 4; };
    size_t row_bytes = 0; for (const Chan& c : chans) row_bytes += (size_t)px_bytes(c.type) * w;
    Image img; img.width = w; img.height = h; img.rgb.assign((size_t)3 * w * h, 0.0f);
    int found = 0;
    const int n_blocks = (h + lines - 1) / lines;
    for (int b = 0; b < n_blocks; ++b) {
        uint64_t off; std::memcpy(&off, &d.at(pos + 8 * (size_t)b + 7) - 7, 8);
        const int y = i32(off); const int32_t size = i32(off + 4);
        const int nl = std::min(lines, win[3] - y + 1);
        const size_t expect = row_bytes * nl;
        std::vector<uint8_t> blk(d.begin() + off + 8, d.begin() + off + 8 + size);
        if (comp != 0 && (size_t)size < expect) {
            std::vector<uint8_t> t;
            if (comp == 1) {   // RLE
                for (size_t i = 0; i < blk.size() && t.size() < expect;) {
                    const int c = (int8_t)blk[i++];
                    if (c < 0) { t.insert(t.end(), blk.begin() + i, blk.begin() + i - c); i += -c; }
                    else { t.insert(t.end(), (size_t)c + 1, blk.at(i)); ++i; }
                }
            } else t = inflate_all(blk.data(), blk.size(), expect);
            for (size_t i = 1; i < t.size(); ++i) t[i] = (uint8_t)(t[i - 1] + t[i] - 128);     // predictor
            blk.resize(t.size());
            const size_t half = (t.size() + 1) / 2;
            for (size_t i = 0; i < t.size(); ++i) blk[i] = (i & 1) ? t[half + i / 2] : t[i / 2];   // de-interleave
        }
        size_t p = 0;
        for (int ly = 0; ly < nl; ++ly) for (const Chan& c : chans) {
            const int k = c.name == "R" ? 0 : c.name == "G" ? 1 : c.name == "B" ? 2 : -1;
            if (k >= 0 && b == 0 && ly == 0) ++found;
            for (int x = 0; x < w; ++x) {
                float v;
                if (c.type == 1) { uint16_t hv; std::memcpy(&hv, &blk.at(p + 2 * x + 1) - 1, 2); v = half_to_float(hv); }
                else if (c.type == 2) std::memcpy(&v, &blk.at(p + 4 * x + 3) - 3, 4);
                else { uint32_t uv; std::memcpy(&uv, &blk.at(p + 4 * x + 3) - 3, 4); v = (float)uv; }
                if (k >= 0) img.rgb[3 * ((size_t)(y - win[1] + ly) * w + x) + k] = v;
            }
            p += (size_t)px_bytes(c.type) * w;
        }
    }
    if (found != 3) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: " + path + " lacks an R, G or B channel");
    return img;
}

// ---- PFM --------------------------------------------------------------------------------------------------------
inline Image read_pfm(const std::string& path) {
    const std::vector<uint8_t> d = read_file(path);
    std::istringstream hs(std::string(d.begin(), d.begin() + std::min<size_t>(d.size(), 256)));
    std::string kind; int w, h; double scale;
    hs >> kind >> w >> h >> scale;
    hs.get();
    const size_t data_at = (size_t)hs.tellg();
    const int nch = kind == "PF" ? 3 : kind == "Pf" ? 1 : 0;
    if (!nch || d.size() < data_at + (size_t)4 * w * h * nch) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: bad PFM file " + path);
    Image img; img.width = w; img.height = h; img.rgb.resize((size_t)3 * w * h);
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) for (int c = 0; c < 3; ++c) {
        uint8_t b[4]; std::memcpy(b, &d[data_at + 4 * (((size_t)(h - 1 - y) * w + x) * nch + (nch == 3 ? c : 0))], 4);
        if (scale > 0) { std::swap(b[0], b[3]); std::swap(b[1], b[2]); }      // positive scale = big endian
        float v; std::memcpy(&v, b, 4);
        img.rgb[3 * ((size_t)y * w + x) + c] = v;
    }
    return img;
}

// ---- PNG: 8-bit RGB / RGBA, non-interlaced ------------------------------------------------------------------------
inline Image read_png(const std::string& path) {
    const std::vector<uint8_t> d = read_file(path);
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (d.size() < 8 || std::memcmp(d.data(), sig, 8)) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: " + path + " is not a PNG file");
    auto be32 = [&](size_t p) { return ((uint32_t)d.at(p) << 24) | ((uint32_t)d.at(p + 1) << 16) | ((uint32_t)d.at(p + 2) << 8) | d.at(p + 3); };
    int w = 0, h = 0, depth = 0, color = 0, interlace = 0;
    std::vector<uint8_t> z;
    for (size_t p = 8; p + 8 <= d.size();) {
        const uint32_t len = be32(p); const std::string type((const char*)&d[p + 4], 4);
        if (type == "IHDR") { w = (int)be32(p + 8); h = (int)be32(p + 12); depth = d.at(p + 16); color = d.at(p + 17); interlace = d.at(p + 20); }
        else if (type == "IDAT") z.insert(z.end(), d.begin() + p + 8, d.begin() + p + 8 + len);
        else if (type == "IEND") break;
        p += 12 + (size_t)len;
    }
    if (depth != 8 || (color != 2 && color != 6) || interlace) throw Error(FTN_ERR_UNSUPPORTED, "fountain: only 8-bit RGB / RGBA non-interlaced PNG images are read");
    const int bpp = color == 2 ? 3 : 4;
    const size_t stride = (size_t)w * bpp;
    std::vector<uint8_t> raw = inflate_all(z.data(), z.size(), (stride + 1) * h);
    if (raw.size() != (stride + 1) * h) throw Error(FTN_ERR_INVALID_ARGUMENT, "fountain: truncated PNG data");
    std::vector<uint8_t> cur(stride), prev(stride, 0);
    Image img; img.width = w; img.height = h; img.rgb.resize((size_t)3 * w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* line = &raw[(stride + 1) * y];
        const int ft = line[0];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
            int pred = 0;
            if (ft == 1) pred = a; else if (ft == 2) pred = b; else if (ft == 3) pred = (a + b) / 2;
            else if (ft == 4) { const int pa = std::abs(b - c), pb = std::abs(a - c), pc = std::abs(a + b - 2 * c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }
            cur[i] = (uint8_t)(line[1 + i] + pred);
        }
        for (int x = 0; x < w; ++x) for (int c = 0; c < 3; ++c) img.rgb[3 * ((size_t)y * w + x) + c] = (float)cur[(size_t)x * bpp + c] / 255.0f;   // Spectrum::from_rgb8
        prev = cur;
    }
    return img;
}

inline std::string extension(const std::string& path) {
    const size_t dot = path.rfind('.');
    std::string e = dot == std::string::npos ? "" : path.substr(dot + 1);
    for (char& ch : e) ch = (char)std::tolower((unsigned char)ch);
    return e;
}
inline Image load_image(const std::string& path) {           // imageio/mod.rs:127-150
    const std::string e = extension(path);
    if (e == "exr") return read_exr(path);
    if (e == "pfm") return read_pfm(path);
    if (e == "png") return read_png(path);
    throw Error(FTN_ERR_UNSUPPORTED, "fountain: no decoder for ." + e + " files on this host");
}
inline float inverse_gamma_correct(float v) {                 // imageio/mod.rs:169-175
    return v <= 0.04045f ? v * 1.0f / 12.92f : std::pow((v + 0.055f) * 1.0f / 1.055f, 2.4f);
}
// imageio/mod.rs:19-43, 82-125.  gamma: -1 = by extension (everything but .exr / .pfm is sRGB), 0 / 1 = as given
struct ImageTexInfo { std::string filename; int wrap_mode = FTN_WRAP_REPEAT; float scale = 1.0f; int gamma = -1; bool flip_y = false; };
inline Image load_texels(const ImageTexInfo& info) {
    Image img = load_image(info.filename);
    const std::string e = extension(info.filename);
    const bool gamma = info.gamma < 0 ? !(e == "exr" || e == "pfm") : info.gamma != 0;
    for (float& v : img.rgb) v = (gamma ? inverse_gamma_correct(v) : v) * info.scale;
    if (info.flip_y)
        for (int y = 0; y < img.height / 2; ++y)
            for (int i = 0; i < 3 * img.width; ++i) std::swap(img.rgb[(size_t)3 * img.width * y + i], img.rgb[(size_t)3 * img.width * (img.height - 1 - y) + i]);
    return img;
}
inline std::shared_ptr<MIPMap> load_mipmap(const ImageTexInfo& info) {
    const Image img = load_texels(info);
    return MIPMap::from_image(img.rgb, img.width, img.height, info.wrap_mode);
}
// loaders/constructors.rs:339-359: `LightSource "infinite"`: the map scaled by scale[0], never gamma-corrected
inline InfiniteAreaLight make_infinite_area_light(const std::string& mapname, float scale0 = 1.0f, const Transform& l2w = Transform::identity()) {
    ImageTexInfo info; info.filename = mapname; info.scale = scale0; info.gamma = 0;
    Image img = load_texels(info);
    return InfiniteAreaLight::new_envmap(std::move(img.rgb), img.width, img.height, l2w);
}

}  // namespace imageio
}  // namespace fountain
