"""Minimal PLY reader (ASCII and binary little-endian triangles) standing in for the `plydough`
crate the reference's loader uses (src/loaders/constructors.rs:94-190): x,y,z [,nx,ny,nz]
[,u,v] per vertex and `vertex_indices` faces with exactly three indices.  Host-side only.
"""
import os

import numpy as np

_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4",
              "float": "f4", "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2",
              "int32": "i4", "uint32": "u4", "float32": "f4", "float64": "f8"}


def load_ply(path):
    """Returns dict(vertices (V,3) f32, normals (V,3) f32 | None, uvs (V,2) f32 | None, indices (T,3) u32)."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    header = data[:end].decode("ascii").splitlines()
    body = data[end:]
    fmt = None
    elements = []   # (name, count, [(prop name, type | ('list', count type, item type))])
    for line in header:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append((tok[1], int(tok[2]), []))
        elif tok[0] == "property":
            if tok[1] == "list":
                elements[-1][2].append((tok[4], ("list", tok[2], tok[3])))
            else:
                elements[-1][2].append((tok[2], tok[1]))
    out = {}
    if fmt == "ascii":
        lines = body.decode("ascii").split("\n")
        pos = 0
        for name, count, props in elements:
            rows = lines[pos:pos + count]
            pos += count
            if name == "vertex":
                # values are decimal f32 literals; parse straight to float32
                arr = np.array([r.split() for r in rows], dtype=np.float32)
                out["vertex"] = {p[0]: arr[:, i] for i, p in enumerate(props)}
            elif name == "face":
                arr = np.array([r.split() for r in rows], dtype=np.int64)
                if not np.all(arr[:, 0] == 3):
                    raise ValueError("Face with unsupported vertex count found")   # constructors.rs:155-157
                out["face"] = arr[:, 1:4]
    elif fmt == "binary_little_endian":
        off = 0
        for name, count, props in elements:
            if name == "vertex":
                dt = np.dtype([(p[0], "<" + _PLY_TYPES[p[1]]) for p in props])
                arr = np.frombuffer(body, dtype=dt, count=count, offset=off)
                off += dt.itemsize * count
                out["vertex"] = {p[0]: arr[p[0]].astype(np.float32) for p in props}
            elif name == "face":
                (pname, (_, ct, it)) = props[0]
                dt = np.dtype([("n", "<" + _PLY_TYPES[ct]), ("v", "<" + _PLY_TYPES[it], (3,))])
                arr = np.frombuffer(body, dtype=dt, count=count, offset=off)
                off += dt.itemsize * count
                if not np.all(arr["n"] == 3):
                    raise ValueError("Face with unsupported vertex count found")
                out["face"] = arr["v"].astype(np.int64)
    else:
        raise ValueError("unsupported PLY format %r" % fmt)
    v = out["vertex"]
    res = {"vertices": np.stack([v["x"], v["y"], v["z"]], axis=1).astype(np.float32),
           "normals": None, "uvs": None,
           "indices": np.ascontiguousarray(out["face"], dtype=np.uint32)}
    if all(k in v for k in ("nx", "ny", "nz")):
        res["normals"] = np.stack([v["nx"], v["ny"], v["nz"]], axis=1).astype(np.float32)
    if all(k in v for k in ("u", "v")):
        res["uvs"] = np.stack([v["u"], v["v"]], axis=1).astype(np.float32)
    return res


_CACHE = {}


def load_ply_cached(path):
    """Parse once per process (the arrays are treated as read-only by the callers)."""
    key = (os.path.abspath(path), os.path.getmtime(path))
    if key not in _CACHE:
        _CACHE[key] = load_ply(path)
    return _CACHE[key]
