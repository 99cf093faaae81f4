"""Host-side image ingestion: what src/imageio/mod.rs does between a file name and the texels a MIPMap / an
InfiniteAreaLight is built from.  Stays on the host side of the ABI -- the device receives texels.

  load_image        imageio/mod.rs:127-150   decode -> (h, w, 3) f32; 8-bit images through Spectrum::from_rgb8 (v / 255)
  load_mipmap       imageio/mod.rs:82-125    gamma (default: everything but .exr), scale, flip_y, MIPMap::new
  ImageTexInfo      imageio/mod.rs:19-43
  read_exr          imageio/exr.rs:11-46     first layer, channels R / G / B as f16 or f32
  write_exr         imageio/exr.rs:48-87     R, G, B as f32 (the reference writes RLE; this writer stores ZIP or raw scan
                                             lines, any reader accepts either)

The reference decodes through third-party crates (`image`, `exr`); here 8-bit formats go through the image library of the
Python host (Pillow) and OpenEXR is read natively: single-part scan-line files, NO / ZIPS / ZIP / RLE compression, HALF /
FLOAT channels -- what `exr` + the reference's writer produce and what environment maps ship as.  PFM is read as well.
"""
import struct
import zlib

import numpy as np

from . import api


class ImageTexInfo:
    """imageio/mod.rs:19-43."""

    def __init__(self, filename, wrap_mode="repeat", scale=1.0, gamma=None, flip_y=False):
        self.filename, self.wrap_mode, self.scale, self.gamma, self.flip_y = str(filename), wrap_mode, float(scale), gamma, bool(flip_y)


def inverse_gamma_correct(v):
    """imageio/mod.rs:169-175, evaluated in f32 like the reference."""
    v = np.asarray(v, dtype=np.float32)
    lo = v * np.float32(1.0) / np.float32(12.92)
    hi = np.power((v + np.float32(0.055)) * np.float32(1.0) / np.float32(1.055), np.float32(2.4), dtype=np.float32)
    return np.where(v <= np.float32(0.04045), lo, hi).astype(np.float32)


def gamma_correct(v):
    """imageio/mod.rs:161-167."""
    v = np.asarray(v, dtype=np.float32)
    return np.where(v <= np.float32(0.0031308), np.float32(12.92) * v,
                    np.float32(1.055) * np.power(np.maximum(v, 0), np.float32(1.0 / 2.4), dtype=np.float32) - np.float32(0.055)).astype(np.float32)


# ---- OpenEXR (scan-line, single part) ------------------------------------------------------------------------------
_EXR_MAGIC = 20000630
_PIXEL_BYTES = {0: 4, 1: 2, 2: 4}          # UINT, HALF, FLOAT
_LINES_PER_BLOCK = {0: 1, 1: 1, 2: 1, 3: 16}   # NO, RLE, ZIPS, ZIP


def _exr_unpredict(buf):
    """Undo the byte-delta predictor and the even/odd byte interleave of EXR's ZIP / RLE blocks."""
    a = np.frombuffer(buf, dtype=np.uint8).astype(np.int64)
    if len(a) > 1:
        a[1:] -= 128
        a = np.cumsum(a) & 0xFF
    a = a.astype(np.uint8)
    half = (len(a) + 1) // 2
    out = np.empty(len(a), dtype=np.uint8)
    out[0::2] = a[:half]
    out[1::2] = a[half:]
    return out.tobytes()


def _exr_predict(raw):
    a = np.frombuffer(raw, dtype=np.uint8)
    t = np.concatenate([a[0::2], a[1::2]]).astype(np.int64)
    d = t.copy()
    d[1:] = (t[1:] - t[:-1] + 128 + 256) & 0xFF
    return d.astype(np.uint8).tobytes()


def _exr_unrle(buf, expect):
    out = bytearray()
    i, n = 0, len(buf)
    while i < n and len(out) < expect:
        c = struct.unpack_from("b", buf, i)[0]
        i += 1
        if c < 0:
            out += buf[i:i - c]
            i += -c
        else:
            out += bytes([buf[i]]) * (c + 1)
            i += 1
    return bytes(out)


def read_exr(path):
    """imageio/exr.rs:11-46: (h, w, 3) f32 from the R, G, B channels of the first layer."""
    data = open(path, "rb").read()
    magic, version = struct.unpack_from("<iI", data, 0)
    if magic != _EXR_MAGIC:
        raise ValueError("%s: not an OpenEXR file" % path)
    if version & 0x200:
        raise NotImplementedError("%s: tiled OpenEXR files are not read" % path)
    if version & 0x1800:
        raise NotImplementedError("%s: multi-part / deep OpenEXR files are not read" % path)
    pos, attrs = 8, {}
    while data[pos] != 0:
        e = data.index(b"\0", pos); name = data[pos:e].decode(); pos = e + 1
        e = data.index(b"\0", pos); typ = data[pos:e].decode(); pos = e + 1
        size = struct.unpack_from("<i", data, pos)[0]; pos += 4
        attrs[name] = (typ, data[pos:pos + size]); pos += size
    pos += 1
    channels, cp, cdata = [], 0, attrs["channels"][1]
    while cdata[cp] != 0:
        e = cdata.index(b"\0", cp); cname = cdata[cp:e].decode(); cp = e + 1
        ptype, _plinear, xs, ys = struct.unpack_from("<iB3xii", cdata, cp); cp += 16
        if xs != 1 or ys != 1:
            raise NotImplementedError("%s: subsampled channels" % path)
        channels.append((cname, ptype))
    compression = attrs["compression"][1][0]
    if compression not in _LINES_PER_BLOCK:
        raise NotImplementedError("%s: OpenEXR compression %d (NO, RLE, ZIPS and ZIP are read)" % (path, compression))
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    lines = _LINES_PER_BLOCK[compression]
    n_blocks = (h + lines - 1) // lines
    offsets = struct.unpack_from("<%dQ" % n_blocks, data, pos)
    row_bytes = sum(_PIXEL_BYTES[t] for _, t in channels) * w
    planes = {c: np.zeros((h, w), dtype=np.float32) for c, _ in channels}
    for off in offsets:
        y, size = struct.unpack_from("<ii", data, off)
        blk = data[off + 8:off + 8 + size]
        nl = min(lines, y1 - y + 1)
        expect = row_bytes * nl
        if compression in (2, 3) and size < expect:
            blk = _exr_unpredict(zlib.decompress(blk))
        elif compression == 1 and size < expect:
            blk = _exr_unpredict(_exr_unrle(blk, expect))
        p = 0
        for ly in range(nl):
            for cname, ptype in channels:       # channels are stored in alphabetical order, line by line
                nb = _PIXEL_BYTES[ptype] * w
                dt = {0: "<u4", 1: "<f2", 2: "<f4"}[ptype]
                planes[cname][y - y0 + ly] = np.frombuffer(blk, dtype=dt, count=w, offset=p).astype(np.float32)
                p += nb
    missing = [c for c in "RGB" if c not in planes]
    if missing:
        raise ValueError("%s: no %s channel" % (path, "/".join(missing)))
    return np.stack([planes["R"], planes["G"], planes["B"]], axis=-1)


def write_exr(path, rgb, compression="zip", half=False):
    """imageio/exr.rs:48-87: one layer, channels B, G, R (alphabetical) as f32 (or f16), scan lines in increasing y."""
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    h, w = rgb.shape[:2]
    comp = {"none": 0, "zips": 2, "zip": 3}[compression]
    ptype, dt = (1, "<f2") if half else (2, "<f4")

    def attr(name, typ, payload):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload
    ch = b"".join(c.encode() + b"\0" + struct.pack("<iB3xii", ptype, 0, 1, 1) for c in "BGR") + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    header = struct.pack("<iI", _EXR_MAGIC, 2) + attr("channels", "chlist", ch) + attr("compression", "compression", bytes([comp])) \
        + attr("dataWindow", "box2i", box) + attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", b"\0") \
        + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) + attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0)) \
        + attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0"
    lines = _LINES_PER_BLOCK[comp]
    blocks = []
    for y in range(0, h, lines):
        raw = b"".join(rgb[yy, :, c].astype(dt).tobytes() for yy in range(y, min(y + lines, h)) for c in (2, 1, 0))
        payload = raw
        if comp:
            z = zlib.compress(_exr_predict(raw))
            if len(z) < len(raw):
                payload = z
        blocks.append(struct.pack("<ii", y, len(payload)) + payload)
    off = len(header) + 8 * len(blocks)
    table = b""
    for b in blocks:
        table += struct.pack("<Q", off)
        off += len(b)
    with open(path, "wb") as f:
        f.write(header + table + b"".join(blocks))


def read_pfm(path):
    """Portable float map (RGB `PF` or grey `Pf`), bottom row first in the file."""
    with open(path, "rb") as f:
        kind = f.readline().strip()
        dims = f.readline().split()
        while len(dims) < 2:
            dims += f.readline().split()
        w, h = int(dims[0]), int(dims[1])
        scale = float(f.readline().strip())
        nch = {b"PF": 3, b"Pf": 1}[kind]
        a = np.frombuffer(f.read(4 * w * h * nch), dtype="<f4" if scale < 0 else ">f4").reshape(h, w, nch)[::-1]
    return np.ascontiguousarray(np.repeat(a, 3, axis=2) if nch == 1 else a, dtype=np.float32)


def load_image(path):
    """imageio/mod.rs:127-150: (h, w, 3) f32 in the file's own encoding (no gamma handling here)."""
    p = str(path)
    ext = p.rsplit(".", 1)[-1].lower() if "." in p else ""
    if ext == "exr":
        return read_exr(p)
    if ext == "pfm":
        return read_pfm(p)
    from PIL import Image           # the Python host's image library, as `image` is the Rust host's
    with Image.open(p) as im:
        if im.mode not in ("RGB", "RGBA"):          # `_ => unimplemented!()` in the reference
            raise NotImplementedError("%s: only 8-bit RGB / RGBA images are read (mode %s)" % (p, im.mode))
        a = np.asarray(im.convert("RGB"), dtype=np.uint8)
    return (a.astype(np.float32) / np.float32(255.0)).astype(np.float32)      # Spectrum::from_rgb8, spectrum/mod.rs:123-130


def load_texels(info):
    """load_mipmap (imageio/mod.rs:82-125) up to MIPMap::new: gamma (by default everything but .exr / .pfm is sRGB),
    scale, flip_y.  Returns (h, w, 3) f32 texels."""
    image = load_image(info.filename)
    gamma = info.gamma
    if gamma is None:
        if "." not in info.filename:
            raise ValueError("No extension on image file %r" % info.filename)
        gamma = info.filename.rsplit(".", 1)[-1].lower() not in ("exr", "pfm")
    if gamma:
        image = inverse_gamma_correct(image)
    image = (image * np.float32(info.scale)).astype(np.float32)
    if info.flip_y:
        image = image[::-1]
    return np.ascontiguousarray(image)


_MIPMAPS = {}


def get_mipmap(info):
    """imageio/mod.rs:60-79: MIPMaps are cached per (file, wrap, scale, gamma, flip)."""
    key = (info.filename, info.wrap_mode, np.float32(info.scale).tobytes(), info.gamma, info.flip_y)
    if key not in _MIPMAPS:
        _MIPMAPS[key] = api.MIPMap(load_texels(info), info.wrap_mode)
    return _MIPMAPS[key]


def make_infinite_area_light(L=1.0, scale=1.0, mapname=None, light_to_world=None):
    """loaders/constructors.rs:339-359: `LightSource "infinite"` -- uniform L when there is no map, otherwise the map
    scaled by scale[0], never gamma-corrected, wrap Repeat."""
    if mapname is None:
        return api.InfiniteAreaLight.new_uniform(L, light_to_world)
    s = float(np.asarray(scale, dtype=np.float32).reshape(-1)[0])
    mip = get_mipmap(ImageTexInfo(mapname, "repeat", s, False, False))
    return api.InfiniteAreaLight.new_envmap(mip.levels[0], light_to_world)


def make_image_texture(filename, wrap="repeat", scale=1.0, gamma=None, mapping=None):
    """make_imagemap_spect (loaders/constructors.rs:295-319): an ImageTexture over the cached MIPMap of the file; flip_y is
    true there (images are stored top row first, texture space has t up)."""
    return api.ImageTexture(get_mipmap(ImageTexInfo(filename, wrap, scale, gamma, True)), mapping)
