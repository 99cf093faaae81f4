"""Host-side image ingestion (fountain_b200/imageio.py = src/imageio/mod.rs + exr.rs): decode, gamma, scale, flip, and
the env light / image texture built from a file.  CPU tier."""
import struct
import zlib

import numpy as np
import pytest

from fountain_b200 import api, imageio


def _img(h, w, seed, hi=4.0):
    rng = np.random.default_rng(seed)
    return (rng.random((h, w, 3)) ** 3 * hi).astype(np.float32)


@pytest.mark.parametrize("compression", ["none", "zips", "zip"])
@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("shape", [(1, 1), (7, 13), (33, 16), (64, 48)])
def test_exr_round_trip(tmp_path, compression, half, shape):
    img = _img(shape[0], shape[1], 3)
    p = str(tmp_path / "a.exr")
    imageio.write_exr(p, img, compression, half)
    back = imageio.read_exr(p)
    want = img.astype(np.float16).astype(np.float32) if half else img
    assert back.shape == img.shape and np.array_equal(back, want)


def _rle(raw):
    out, i = bytearray(), 0
    while i < len(raw):
        j = i
        while j + 1 < len(raw) and raw[j + 1] == raw[i] and j - i < 126:
            j += 1
        if j - i >= 2:
            out += struct.pack("b", j - i) + raw[i:i + 1]; i = j + 1
        else:
            k = i
            while k < len(raw) and k - i < 127 and not (k + 2 < len(raw) and raw[k] == raw[k + 1] == raw[k + 2]):
                k += 1
            out += struct.pack("b", -(k - i)) + raw[i:k]; i = k
    return bytes(out)


def test_exr_rle_blocks_as_the_reference_writes(tmp_path):
    """imageio/exr.rs:80 writes Compression::RLE: re-pack a file's scan lines as RLE blocks and read it back."""
    img = np.zeros((9, 20, 3), np.float32); img[2:5, 3:9] = [1.5, 0.25, 7.0]; img[7, :, 1] = 0.5
    p = str(tmp_path / "a.exr")
    imageio.write_exr(p, img, "none")
    data = bytearray(open(p, "rb").read())
    i = data.index(b"compression\0compression\0") + len(b"compression\0compression\0") + 4
    data[i] = 1
    h, w = img.shape[:2]
    table_at = data.index(b"screenWindowWidth\0float\0") + len(b"screenWindowWidth\0float\0") + 4 + 4 + 1
    row = 12 * w
    blocks, off = [], table_at + 8 * h
    for y in range(h):
        o = struct.unpack_from("<Q", data, table_at + 8 * y)[0]
        raw = bytes(data[o + 8:o + 8 + row])
        z = _rle(imageio._exr_predict(raw))
        blocks.append(struct.pack("<ii", y, len(z)) + z if len(z) < row else struct.pack("<ii", y, row) + raw)
    out = bytearray(data[:table_at])
    for b in blocks:
        out += struct.pack("<Q", off); off += len(b)
    out += b"".join(blocks)
    q = str(tmp_path / "rle.exr")
    open(q, "wb").write(out)
    assert np.array_equal(imageio.read_exr(q), img)


def test_png_gamma_scale_flip(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (6, 9, 3), dtype=np.uint8)
    a[0, 0] = [0, 10, 11]                       # around the 0.04045 knee of the sRGB curve (10/255 < knee < 11/255)
    p = str(tmp_path / "t.png")
    Image.fromarray(a, "RGB").save(p)
    raw = imageio.load_image(p)
    assert np.array_equal(raw, a.astype(np.float32) / np.float32(255.0))          # Spectrum::from_rgb8
    t = imageio.load_texels(imageio.ImageTexInfo(p, "repeat", 2.0, None, True))    # default gamma for non-exr: true
    want = (imageio.inverse_gamma_correct(raw) * np.float32(2.0))[::-1]
    assert np.array_equal(t, want)
    v = a[0, 0].astype(np.float64) / 255.0
    ref = np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4) * 2.0
    assert np.allclose(t[-1, 0], ref, rtol=1e-5, atol=1e-7)
    rgba = str(tmp_path / "t_rgba.png")
    Image.fromarray(np.dstack([a, np.full(a.shape[:2], 128, np.uint8)]), "RGBA").save(rgba)
    assert np.array_equal(imageio.load_image(rgba), raw)                            # to_rgb() drops alpha
    lin = imageio.load_texels(imageio.ImageTexInfo(p, "repeat", 1.0, False, False))
    assert np.array_equal(lin, raw)
    grey = str(tmp_path / "g.png")
    Image.fromarray(a[..., 0], "L").save(grey)
    with pytest.raises(NotImplementedError):                                        # `_ => unimplemented!()`
        imageio.load_image(grey)


def test_pfm(tmp_path):
    img = _img(5, 8, 9)
    p = str(tmp_path / "a.pfm")
    with open(p, "wb") as f:
        f.write(b"PF\n8 5\n-1.0\n" + img[::-1].astype("<f4").tobytes())
    assert np.array_equal(imageio.load_image(p), img)
    assert np.array_equal(imageio.load_texels(imageio.ImageTexInfo(p)), img)        # linear by default, like .exr


def test_env_light_from_exr_file_equals_texels(tmp_path, oracle, orc_backend):
    """make_infinite_area_light (constructors.rs:339-359): mapname -> texels scaled by scale[0], no gamma; the oracle's
    importance sampling of the file-built light equals that of the light built from the same array."""
    import ctypes as C
    from fountain_b200 import _abi as A
    img = _img(16, 32, 11, hi=50.0)
    p = str(tmp_path / "sky.exr")
    imageio.write_exr(p, img, "zip")
    la = imageio.make_infinite_area_light(scale=(0.5, 9, 9), mapname=p)
    lb = api.InfiniteAreaLight.new_envmap(img * np.float32(0.5))
    assert np.array_equal(la.texels, lb.texels)
    sa, sb = api.Scene([], [la], backend=orc_backend), api.Scene([], [lb], backend=orc_backend)
    rng = np.random.default_rng(2)
    for _ in range(50):
        u = (A.f32 * 2)(*rng.random(2)); oa, ob = (A.f32 * 11)(), (A.f32 * 11)()
        assert oracle.library().orc_kat_env(sa.handle, u, oa) == oracle.library().orc_kat_env(sb.handle, u, ob)
        assert list(oa) == list(ob)
    assert imageio.make_infinite_area_light(L=3.0).texels.shape == (1, 1, 3)


def test_image_texture_from_file_is_cached_and_flipped(tmp_path):
    from PIL import Image
    a = np.random.default_rng(1).integers(0, 256, (8, 8, 3), dtype=np.uint8)
    p = str(tmp_path / "tex.png")
    Image.fromarray(a, "RGB").save(p)
    t1 = imageio.make_image_texture(p, "clamp", 1.5)
    t2 = imageio.make_image_texture(p, "clamp", 1.5)
    assert t1.mipmap is t2.mipmap                                                    # get_mipmap's cache (imageio/mod.rs:60-79)
    want = (imageio.inverse_gamma_correct(a.astype(np.float32) / np.float32(255.0)) * np.float32(1.5))[::-1]
    assert np.array_equal(t1.mipmap.levels[0], want)
    assert len(t1.mipmap.levels) == 4 and t1.mipmap.wrap == api.MIPMap.WRAP["clamp"]


# ---- the C++ host's decoders (include/fountain_imageio.hpp) against the Python host's ---------------------------------
def _cpp_dump(tmp_path, path, *texel_args):
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-s", "-C", os.path.join(root, "tests", "cpp"), "imageio_dump"], check=True)
    out = str(tmp_path / "dump.bin")
    p = subprocess.run([os.path.join(root, "tests", "cpp", "imageio_dump"), path, out, *[str(a) for a in texel_args]], capture_output=True, text=True)
    if p.returncode != 0:
        return p.returncode, p.stderr
    raw = open(out, "rb").read()
    w, h = struct.unpack_from("<ii", raw, 0)
    return 0, np.frombuffer(raw, dtype="<f4", offset=8).reshape(h, w, 3)


@pytest.mark.parametrize("compression,half", [("none", False), ("zips", True), ("zip", False), ("zip", True)])
def test_cpp_host_reads_exr(tmp_path, compression, half):
    img = _img(37, 21, 4)
    p = str(tmp_path / "a.exr")
    imageio.write_exr(p, img, compression, half)
    rc, got = _cpp_dump(tmp_path, p)
    assert rc == 0, got
    assert np.array_equal(got, imageio.read_exr(p))


def test_cpp_host_reads_png_pfm_and_applies_gamma_scale_flip(tmp_path):
    from PIL import Image
    a = np.random.default_rng(8).integers(0, 256, (11, 7, 4), dtype=np.uint8)
    png = str(tmp_path / "t.png")
    Image.fromarray(a, "RGBA").save(png)
    rc, got = _cpp_dump(tmp_path, png)
    assert rc == 0, got
    assert np.array_equal(got, imageio.load_image(png))
    rc, got = _cpp_dump(tmp_path, png, 2.0, -1, 1)
    want = imageio.load_texels(imageio.ImageTexInfo(png, "repeat", 2.0, None, True))
    assert rc == 0 and np.allclose(got, want, rtol=2e-6, atol=0)           # powf: libm vs numpy, an ulp apart at most
    img = _img(5, 8, 9)
    pfm = str(tmp_path / "a.pfm")
    with open(pfm, "wb") as f:
        f.write(b"PF\n8 5\n-1.0\n" + img[::-1].astype("<f4").tobytes())
    rc, got = _cpp_dump(tmp_path, pfm)
    assert rc == 0 and np.array_equal(got, img)
    grey = str(tmp_path / "g.png")
    Image.fromarray(a[..., 0], "L").save(grey)
    rc, err = _cpp_dump(tmp_path, grey)
    assert rc == 1 and "8-bit RGB / RGBA" in err
