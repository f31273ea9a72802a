"""Mesh distance fields for the global-SDF build: the reference's baked `.sdf` files, the MeshDistanceField
records of its `.scene` files, and synthetic fields for tests.  Harness-side Python; the build itself is lux_ddgi_build_global_sdf.

File format (SDFBaker.cpp:158-204, reader :207-240), cereal portable binary, little endian:
    uvec3 size | i32 mipCount | u64 n, n bytes fp16 (d/maxDistance + 1)/2 [z][y][x] | per further mip: u64 n, n bytes
"""
import ctypes as C
import json
import os
import struct
from dataclasses import dataclass, field

import numpy as np

from . import abi


@dataclass
class MeshSDF:
    levels: list          # np.float16 [d][h][w] per mip
    aabb_min: np.ndarray  # padded local bounds (MeshDistanceField.aabb)
    aabb_max: np.ndarray
    uvw_mul: np.ndarray   # localToUVWMul = 1/size
    uvw_add: np.ndarray   # localToUVWAdd = -min/size
    max_distance: float
    world: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))  # [row][col]
    name: str = ""


def read_sdf_file(path):
    """-> list of np.float16 volumes [d][h][w], one per mip."""
    raw = open(path, "rb").read()
    sx, sy, sz, mips = struct.unpack_from("<IIIi", raw, 0)
    off, levels = 16, []
    for m in range(mips):
        (n,) = struct.unpack_from("<Q", raw, off)
        off += 8
        w, h, d = max(sx >> m, 1), max(sy >> m, 1), max(sz >> m, 1)
        assert n == w * h * d * 2, (path, m, n, (w, h, d))
        levels.append(np.frombuffer(raw, dtype="<f2", count=w * h * d, offset=off).reshape(d, h, w).copy())
        off += n
    assert off == len(raw), (path, off, len(raw))
    return levels


def box_filter_mips(level0: np.ndarray, count: int = 3):
    """The baker's 2x2x2 box filter (SDFBaker.cpp:166-198): float sum in dz, dy, dx order, * 1/8, packHalf."""
    levels = [np.ascontiguousarray(level0, dtype=np.float16)]
    for _ in range(1, count):
        src = levels[-1].astype(np.float32)
        d, h, w = (max(s // 2, 1) for s in src.shape)
        acc = np.zeros((d, h, w), dtype=np.float32)
        for dz in range(2):
            for dy in range(2):
                for dx in range(2):
                    acc = (acc + src[dz:2 * d:2, dy:2 * h:2, dx:2 * w:2]).astype(np.float32)
        levels.append((acc * np.float32(0.125)).astype(np.float16))
    return levels


def synthetic(kind: str, size=(16, 16, 16), half=(1.0, 1.0, 1.0), pad: float = 0.25, world=None, name="") -> MeshSDF:
    """Analytic box / sphere baked like SDFBaker: padded bounds, voxel-centre samples, value (d/maxDistance + 1)/2."""
    half = np.asarray(half, dtype=np.float32)
    mn, mx = -(half + np.float32(pad)), half + np.float32(pad)
    w, h, d = size
    xs = mn[0] + (np.arange(w, dtype=np.float32) + 0.5) * (mx[0] - mn[0]) / w
    ys = mn[1] + (np.arange(h, dtype=np.float32) + 0.5) * (mx[1] - mn[1]) / h
    zs = mn[2] + (np.arange(d, dtype=np.float32) + 0.5) * (mx[2] - mn[2]) / d
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    if kind == "sphere":
        dist = np.sqrt(X * X + Y * Y + Z * Z) - half[0]
    else:
        q = np.stack([np.abs(X) - half[0], np.abs(Y) - half[1], np.abs(Z) - half[2]], -1)
        dist = np.linalg.norm(np.maximum(q, 0), axis=-1) + np.minimum(q.max(-1), 0)
    max_distance = float(np.linalg.norm(mx - mn))
    enc = ((dist / np.float32(max_distance) + 1) * 0.5).astype(np.float16)
    size_v = (mx - mn).astype(np.float32)
    return MeshSDF(box_filter_mips(enc), mn, mx, (1.0 / size_v).astype(np.float32), (-mn / size_v).astype(np.float32), max_distance,
                   np.eye(4, dtype=np.float32) if world is None else np.asarray(world, dtype=np.float32), name)


def to_ctypes(meshes):
    """-> (array of abi.MeshSDF, keep-alive list).  worldMatrix is stored column-major."""
    arr = (abi.MeshSDF * len(meshes))()
    keep = []
    for a, m in zip(arr, meshes):
        lv = [np.ascontiguousarray(l, dtype=np.float16) for l in m.levels]
        keep.append(lv)
        for i in range(3):
            a.mips[i] = lv[i].ctypes.data if i < len(lv) else None
        d, h, w = lv[0].shape
        a.size[:] = [w, h, d]
        a.mipCount = len(lv)
        a.aabbMin[:] = [float(x) for x in m.aabb_min]
        a.aabbMax[:] = [float(x) for x in m.aabb_max]
        a.localToUVWMul[:] = [float(x) for x in m.uvw_mul]
        a.localToUVWAdd[:] = [float(x) for x in m.uvw_add]
        a.maxDistance = float(m.max_distance)
        a.worldMatrix[:] = [float(x) for x in np.asarray(m.world, dtype=np.float32).T.reshape(-1)]
    return arr, keep


# ----------------------------------------------------------------------------------------------------------------------
# the reference's .scene files (cereal JSON of an entt snapshot): component pools are `count, (entity, record) * count`
# ----------------------------------------------------------------------------------------------------------------------
def _vec(d):
    return np.array([d[f"value{i}"] for i in range(len(d))], dtype=np.float32)


def load_scene_meshes(scene_path: str, asset_root: str):
    """MeshDistanceField records of a reference scene + the entities' world matrices -> list of MeshSDF.
    Transforms: Position / Rotation (Euler degrees, stored) / Scale / Offset per entity, parents from the Hierarchy pool."""
    vals = list(json.load(open(scene_path)).values())
    transforms, parents, names, fields = {}, {}, {}, {}
    i = 0
    while i < len(vals) - 1:
        v, nxt = vals[i], vals[i + 1]
        if isinstance(nxt, dict) and isinstance(v, int):
            if "Position" in nxt:
                transforms[v] = nxt
            elif set(nxt.keys()) == {"value0", "value1", "value2", "value3"} and all(isinstance(x, int) for x in nxt.values()):
                parents.setdefault(v, nxt["value0"])  # Hierarchy: parent, first, next, prev (4294967295 = null)
            elif list(nxt.keys()) == ["value0"] and isinstance(nxt["value0"], str):
                names.setdefault(v, nxt["value0"])
            elif isinstance(nxt.get("value0"), str) and nxt["value0"].endswith(".sdf"):
                fields[v] = nxt
            i += 2
        else:
            i += 1

    def local(e):
        t = transforms.get(e)
        if t is None:
            return np.eye(4, dtype=np.float64)
        p, r, s = _vec(t["Position"]), np.radians(_vec(t["Rotation"]).astype(np.float64)), _vec(t["Scale"])
        cx, cy, cz = np.cos(r)
        sx, sy, sz = np.sin(r)
        rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        m = np.eye(4)
        m[:3, :3] = (rz @ ry @ rx) * s[None, :]
        m[:3, 3] = p
        off = np.array([[t["Offset"][f"value{c}"][f"value{rr}"] for c in range(4)] for rr in range(4)], dtype=np.float64)
        return m @ off

    def world(e, depth=0):
        m = local(e)
        p = parents.get(e, 4294967295)
        return m if p == 4294967295 or p not in transforms or depth > 32 else world(p, depth + 1) @ m

    out = []
    for e, f in fields.items():
        bb = f["value1"]
        mn, mx = _vec(bb["value0"]), _vec(bb["value1"])
        levels = read_sdf_file(os.path.join(asset_root, f["value0"]))
        out.append(MeshSDF(levels, mn, mx, _vec(f["value3"]), _vec(f["value2"]), float(f["value4"]), world(e).astype(np.float32),
                           os.path.basename(f["value0"])))
    return out
