"""Point-cloud file exchange in libpointmatcher's three text formats:
`DataPoints::load / save` for `.csv`, legacy ASCII `.vtk` (POLYDATA) and ASCII `.ply`.

Host-side only (fixture exchange with a libpointmatcher installation); nothing here is on
the registration path.  A cloud is `(features, descriptors)`: features 4 x N float32 with a
last row of ones, descriptors {label: span x N float32}, libpointmatcher's own layout.

Column naming follows upstream's IO.cpp [UPSTREAM-RECALLED]: x y z -> features;
nx ny nz -> `normals`; a scalar column `name` -> descriptor `name`; vector columns
`name_x name_y name_z` (CSV) / `VECTORS|NORMALS name` (VTK) -> one span-3 descriptor.
"""
from __future__ import annotations

import os
import re

import numpy as np

_CSV_NORMALS = ("nx", "ny", "nz")


def _features(xyz: np.ndarray) -> np.ndarray:
    xyz = np.asarray(xyz, dtype=np.float32).reshape(3, -1)
    return np.asfortranarray(np.vstack([xyz, np.ones((1, xyz.shape[1]), np.float32)]))


def _check(features, descriptors):
    f = np.asarray(features, dtype=np.float32)
    if f.ndim != 2 or f.shape[0] != 4:
        raise ValueError("features must be 4 x N (x, y, z, pad)")
    d = {}
    for k, v in (descriptors or {}).items():
        v = np.asarray(v, dtype=np.float32)
        v = v.reshape(1, -1) if v.ndim == 1 else v
        if v.shape[1] != f.shape[1]:
            raise ValueError(f"descriptor {k}: {v.shape[1]} columns for {f.shape[1]} points")
        d[k] = v
    return f, d


# ------------------------------------------------------------------------------- CSV
def _csv_columns(descriptors):
    cols = []
    for k, v in descriptors.items():
        if k == "normals" and v.shape[0] == 3:
            cols += [(n, v[i]) for i, n in enumerate(_CSV_NORMALS)]
        elif v.shape[0] == 1:
            cols.append((k, v[0]))
        elif v.shape[0] == 3:
            cols += [(f"{k}_{a}", v[i]) for i, a in enumerate("xyz")]
        else:
            cols += [(f"{k}_{i}", v[i]) for i in range(v.shape[0])]
    return cols


def save_csv(path, features, descriptors=None):
    f, d = _check(features, descriptors)
    cols = [("x", f[0]), ("y", f[1]), ("z", f[2])] + _csv_columns(d)
    data = np.stack([c for _, c in cols], axis=1) if f.shape[1] else np.zeros((0, len(cols)), np.float32)
    with open(path, "w") as fh:
        fh.write(",".join(n for n, _ in cols) + "\n")
        np.savetxt(fh, data, fmt="%.9g", delimiter=",")


def load_csv(path):
    with open(path) as fh:
        lines = [ln.strip() for ln in fh if ln.strip() and not ln.lstrip().startswith("#")]
    if not lines:
        return _features(np.zeros((3, 0))), {}
    split = lambda ln: [t for t in re.split(r"[,;\s]+", ln) if t]
    first = split(lines[0])
    has_header = any(re.search(r"[A-Za-df-z_]", t) for t in first)  # 'e' alone may be an exponent
    names = first if has_header else ["x", "y", "z"] + [f"c{i}" for i in range(len(first) - 3)]
    rows = lines[1:] if has_header else lines
    data = np.array([[float(t) for t in split(ln)] for ln in rows], dtype=np.float32).reshape(len(rows), len(names))
    return _columns_to_cloud(names, {n: data[:, i] for i, n in enumerate(names)}, path)


def _columns_to_cloud(names, col, path):
    """named scalar columns -> (features, descriptors) by upstream's naming conventions"""
    for a in "xyz":
        if a not in col:
            raise ValueError(f"{path}: no '{a}' column (libpointmatcher needs x, y, z)")
    feats = _features(np.stack([col["x"], col["y"], col["z"]]))
    desc, used = {}, {"x", "y", "z"}
    if all(n in col for n in _CSV_NORMALS):
        desc["normals"] = np.stack([col[n] for n in _CSV_NORMALS])
        used.update(_CSV_NORMALS)
    for n in names:
        if n in used:
            continue
        m = re.match(r"(.+)_x$", n)
        if m and all(f"{m.group(1)}_{a}" in col for a in "xyz"):
            desc[m.group(1)] = np.stack([col[f"{m.group(1)}_{a}"] for a in "xyz"])
            used.update(f"{m.group(1)}_{a}" for a in "xyz")
            continue
        m = re.match(r"(.+)_0$", n)
        if m:
            k = 0
            while f"{m.group(1)}_{k}" in col:
                k += 1
            desc[m.group(1)] = np.stack([col[f"{m.group(1)}_{i}"] for i in range(k)])
            used.update(f"{m.group(1)}_{i}" for i in range(k))
            continue
        desc[n] = col[n][None]
        used.add(n)
    return feats, {k: np.asfortranarray(v.astype(np.float32)) for k, v in desc.items()}


# ------------------------------------------------------------------------------- VTK
def save_vtk(path, features, descriptors=None):
    f, d = _check(features, descriptors)
    n = f.shape[1]
    with open(path, "w") as fh:
        fh.write("# vtk DataFile Version 3.0\nFile created by pgslam_b200\nASCII\nDATASET POLYDATA\n")
        fh.write(f"POINTS {n} float\n")
        np.savetxt(fh, f[:3].T, fmt="%.9g")
        fh.write(f"VERTICES {n} {2 * n}\n")
        np.savetxt(fh, np.stack([np.ones(n, np.int64), np.arange(n)], axis=1), fmt="%d")
        if d:
            fh.write(f"POINT_DATA {n}\n")
        for k, v in d.items():
            if k == "normals" and v.shape[0] == 3:
                fh.write(f"NORMALS {k} float\n")
            elif v.shape[0] == 1:
                fh.write(f"SCALARS {k} float 1\nLOOKUP_TABLE default\n")
            elif v.shape[0] == 3:
                fh.write(f"VECTORS {k} float\n")
            elif v.shape[0] == 9:
                fh.write(f"TENSORS {k} float\n")
            else:
                fh.write(f"SCALARS {k} float {v.shape[0]}\nLOOKUP_TABLE default\n")
            np.savetxt(fh, v.T.reshape(-1, 3) if v.shape[0] == 9 else v.T, fmt="%.9g")


def load_vtk(path):
    with open(path) as fh:
        text = fh.read()
    tok = text.split("\n", 2)
    if len(tok) < 3 or not tok[0].startswith("# vtk DataFile"):
        raise ValueError(f"{path}: not a legacy VTK file")
    words = tok[2].split()
    if words[0].upper() != "ASCII":
        raise ValueError(f"{path}: only ASCII VTK files are supported")
    i = 1
    feats, desc, n = None, {}, 0

    def floats(count):
        nonlocal i
        out = np.array(words[i:i + count], dtype=np.float64).astype(np.float32)
        if out.size != count:
            raise ValueError(f"{path}: truncated data block")
        i += count
        return out

    while i < len(words):
        w = words[i].upper()
        if w == "DATASET":
            if words[i + 1].upper() not in ("POLYDATA", "UNSTRUCTURED_GRID"):
                raise ValueError(f"{path}: unsupported DATASET {words[i + 1]}")
            i += 2
        elif w == "POINTS":
            n = int(words[i + 1])
            i += 3
            feats = _features(floats(3 * n).reshape(n, 3).T)
        elif w in ("VERTICES", "LINES", "POLYGONS", "CELLS"):
            i += 3 + int(words[i + 2])
        elif w == "CELL_TYPES":
            i += 2 + int(words[i + 1])
        elif w == "POINT_DATA":
            i += 2
        elif w in ("NORMALS", "VECTORS"):
            name = words[i + 1]
            i += 3
            desc[name] = floats(3 * n).reshape(n, 3).T
        elif w == "TENSORS":
            name = words[i + 1]
            i += 3
            desc[name] = floats(9 * n).reshape(n, 9).T
        elif w == "SCALARS":
            name = words[i + 1]
            span = 1
            i += 3
            if i < len(words) and re.fullmatch(r"\d+", words[i]):
                span = int(words[i])
                i += 1
            if i < len(words) and words[i].upper() == "LOOKUP_TABLE":
                i += 2
            desc[name] = floats(span * n).reshape(n, span).T
        elif w == "COLOR_SCALARS":
            name, span = words[i + 1], int(words[i + 2])
            i += 3
            desc[name] = floats(span * n).reshape(n, span).T
        else:
            raise ValueError(f"{path}: unsupported VTK keyword {words[i]}")
    if feats is None:
        raise ValueError(f"{path}: no POINTS block")
    return feats, {k: np.asfortranarray(v) for k, v in desc.items()}


# ------------------------------------------------------------------------------- PLY
def save_ply(path, features, descriptors=None):
    f, d = _check(features, descriptors)
    cols = [("x", f[0]), ("y", f[1]), ("z", f[2])] + _csv_columns(d)
    with open(path, "w") as fh:
        fh.write(f"ply\nformat ascii 1.0\ncomment created by pgslam_b200\nelement vertex {f.shape[1]}\n")
        for name, _ in cols:
            fh.write(f"property float {name}\n")
        fh.write("end_header\n")
        if f.shape[1]:
            np.savetxt(fh, np.stack([c for _, c in cols], axis=1), fmt="%.9g")


def load_ply(path):
    with open(path, "rb") as fh:
        raw = fh.read()
    end = raw.find(b"end_header")
    if not raw.startswith(b"ply") or end < 0:
        raise ValueError(f"{path}: not a PLY file")
    header = raw[:end].decode("ascii", "replace").splitlines()
    body = raw[raw.index(b"\n", end) + 1:]
    fmt, n, props, in_vertex = None, 0, [], False
    for ln in header:
        t = ln.split()
        if not t:
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            in_vertex = t[1] == "vertex"
            if in_vertex:
                n = int(t[2])
        elif t[0] == "property" and in_vertex:
            if t[1] == "list":
                raise ValueError(f"{path}: list properties on vertices are not supported")
            props.append((t[2], t[1]))
    np_types = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4",
                "float": "f4", "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2",
                "int32": "i4", "uint32": "u4", "float32": "f4", "float64": "f8"}
    if fmt == "ascii":
        vals = np.array(body.split()[:n * len(props)], dtype=np.float64).reshape(n, len(props))
        col = {name: vals[:, j].astype(np.float32) for j, (name, _) in enumerate(props)}
    elif fmt in ("binary_little_endian", "binary_big_endian"):
        e = "<" if fmt.endswith("little_endian") else ">"
        dt = np.dtype([(name, e + np_types[t]) for name, t in props])
        rec = np.frombuffer(body, dtype=dt, count=n)
        col = {name: rec[name].astype(np.float32) for name, _ in props}
    else:
        raise ValueError(f"{path}: unknown PLY format {fmt}")
    return _columns_to_cloud([name for name, _ in props], col, path)


_LOADERS = {".csv": load_csv, ".vtk": load_vtk, ".ply": load_ply}
_SAVERS = {".csv": save_csv, ".vtk": save_vtk, ".ply": save_ply}


def load(path):
    """DataPoints::load: dispatch on the file extension -> (features 4xN, {label: span x N})."""
    ext = os.path.splitext(path)[1].lower()
    if ext not in _LOADERS:
        raise ValueError(f"unknown point-cloud file extension '{ext}' (csv, vtk, ply)")
    return _LOADERS[ext](path)


def save(path, features, descriptors=None):
    """DataPoints::save: dispatch on the file extension."""
    ext = os.path.splitext(path)[1].lower()
    if ext not in _SAVERS:
        raise ValueError(f"unknown point-cloud file extension '{ext}' (csv, vtk, ply)")
    _SAVERS[ext](path, features, descriptors)
