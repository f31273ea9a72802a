"""Synthetic inputs for the DDGI probe update: global SDF + mip, surface cache, probe volume, per-frame rotations.

These are INPUT fixtures: they build the buffers the reference's sibling modules would hand to the
DDGI pass, in the reference's layouts:

* global SDF: R16F [z][y][x], value = clamp(d_world / (2D), -1, 1)          (Shaders/SDF/SDFRasterizeModel.glsl:61)
* mip: restatement of Shaders/SDF/GlobalSDFMipmap.comp:32-68 + the 4 flood passes of
  Engine/DDGI/GlobalDistanceField.cpp:537-573                                 (quirks kept: mixed units, point sample)
* surface cache: ObjectBuffer / TileBuffer records as built by Engine/DDGI/GlobalSurfaceAtlas.cpp:226-280,
  chunk lists as built by Shaders/SDF/SDFCulling.comp:36-101, light-cache (RGBA16F) and depth (D32F) atlases whose
  depth satisfies the SAMPLING formula of Shaders/SDF/AtlasCommon.glsl:62-84 (SURVEY.md A.3 caveat)
* rotation: angle-axis -> mat4 exactly as Engine/DDGI/DDGIRenderer.cpp:267-272, from a SplitMix64 stream

Heavy arrays are torch tensors so the same code runs on the CPU (tests, small scenes) and on the GPU (512^3 / 1024^3).
Nothing here is on the product path: the engine only ever sees the resulting buffers.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import abi

MASK64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & MASK64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def uniform(self, lo=0.0, hi=1.0) -> float:
        return lo + (hi - lo) * ((self.next() >> 11) * (1.0 / (1 << 53)))


def frame_rotation(frame: int, seed: int = 0x4C555847) -> np.ndarray:
    """Column-major mat4 (16 floats) = mat4_cast(angleAxis(U(0,1)*2pi, normalize(U(-1,1)^3))), DDGIRenderer.cpp:267-272."""
    rng = SplitMix64(seed + frame)
    axis = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-1, 1)], dtype=np.float64)
    axis /= np.linalg.norm(axis)
    angle = rng.uniform(0, 1) * math.pi * 2.0
    return rotation_from_axis_angle(axis, angle)


def rotation_from_axis_angle(axis, angle) -> np.ndarray:
    s, c = math.sin(angle * 0.5), math.cos(angle * 0.5)
    x, y, z, w = axis[0] * s, axis[1] * s, axis[2] * s, c
    m = np.zeros((4, 4), dtype=np.float64)  # m[col][row]
    m[0] = [1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y), 0]
    m[1] = [2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x), 0]
    m[2] = [2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y), 0]
    m[3] = [0, 0, 0, 1]
    return m.reshape(16).astype(np.float32)


def identity_rotation() -> np.ndarray:
    return np.eye(4, dtype=np.float32).reshape(16)


@dataclass
class Scene:
    name: str
    uniform: abi.DDGIUniform
    sdf_data: abi.GlobalSDFData
    sdf: torch.Tensor  # float16 [res][res][res*cascades]
    mip: torch.Tensor  # float16 [res/4][res/4][res/4*cascades]
    atlas_data: Optional[abi.GlobalSurfaceAtlasData] = None
    chunks: Optional[np.ndarray] = None  # uint32[64000]
    cull: Optional[np.ndarray] = None  # uint32[]
    objects: Optional[np.ndarray] = None  # abi.OBJECT_DTYPE
    tiles: Optional[np.ndarray] = None  # abi.TILE_DTYPE
    light: Optional[torch.Tensor] = None  # float16 [res][res][4]
    depth: Optional[torch.Tensor] = None  # float32 [res][res]
    sky_face: int = 0
    sky: Optional[np.ndarray] = None  # float16 [6][n][n][4]
    meta: dict = field(default_factory=dict)

    @property
    def probes(self) -> int:
        return abi.probe_count(self.uniform)

    @property
    def rays(self) -> int:
        return self.uniform.raysPerProbe

    def sdf_bytes(self) -> int:
        return self.sdf.numel() * 2 + self.mip.numel() * 2


# ------------------------------------------------------------------------------------------------------------------
# SDF helpers
# ------------------------------------------------------------------------------------------------------------------
def make_sdf_data(center, half_extent: float, res: int) -> abi.GlobalSDFData:
    d = abi.GlobalSDFData()
    d.cascadePosDistance[0][:] = [float(center[0]), float(center[1]), float(center[2]), float(half_extent)]
    d.cascadeVoxelSize[:] = [2.0 * half_extent / res, 0.0, 0.0, 0.0]
    d.cascadesCount = 1
    d.resolution = float(res)
    d.nearPlane = 0.1
    d.farPlane = 1000.0
    return d


def voxel_centers(center, half_extent: float, res: int, device, z0=0, z1=None):
    """World positions of voxel centres (GlobalDistanceField.cpp:758-759) for z-slices [z0, z1)."""
    z1 = res if z1 is None else z1
    vox = 2.0 * half_extent / res
    ax = lambda c, lo, hi: (torch.arange(lo, hi, device=device, dtype=torch.float32) + 0.5) * vox + (c - half_extent)
    return ax(center[0], 0, res), ax(center[1], 0, res), ax(center[2], z0, z1)


def sd_box(px, py, pz, c, h, rot_y: float = 0.0):
    """Exact box SDF; the box is rotated by rot_y radians about +y around its centre."""
    x, y, z = px - c[0], py - c[1], pz - c[2]
    if rot_y != 0.0:
        cs, sn = math.cos(rot_y), math.sin(rot_y)
        x, z = cs * x - sn * z, sn * x + cs * z  # world -> local = R^T
    qx, qy, qz = x.abs() - h[0], y.abs() - h[1], z.abs() - h[2]
    outside = torch.sqrt(qx.clamp(min=0) ** 2 + qy.clamp(min=0) ** 2 + qz.clamp(min=0) ** 2)
    inside = torch.maximum(qx, torch.maximum(qy, qz)).clamp(max=0)
    return outside + inside


def encode_sdf(d_world: torch.Tensor, half_extent: float) -> torch.Tensor:
    return (d_world / (2.0 * half_extent)).clamp(-1.0, 1.0).to(torch.float16)


def _combine(sdf, dist):
    """combineDistanceToSDF, Shaders/SDF/SDFCommon.glsl:18-37 (dist is a python float >= 0)."""
    out = torch.sqrt(sdf.clamp(min=0) ** 2 + dist * dist)
    if dist <= 0:
        out = torch.where(sdf <= 0, sdf, out)
    return out


def _mip_pass(src: torch.Tensor, out_res: int, scale: int, src_res: int, max_distance: float) -> torch.Tensor:
    """One dispatch of GlobalSDFMipmap.comp (single cascade): out[c] = min_o combine(src[clamp(c*scale+o)], |o|*maxD/srcRes)."""
    dev = src.device
    idx = torch.arange(out_res, device=dev) * scale

    def tap(ox, oy, oz):
        ix = (idx + ox).clamp(0, src_res - 1)
        iy = (idx + oy).clamp(0, src_res - 1)
        iz = (idx + oz).clamp(0, src_res - 1)
        v = src[iz][:, iy][:, :, ix].to(torch.float32)
        dist = math.sqrt(ox * ox + oy * oy + oz * oz) * (max_distance / float(src_res))
        return _combine(v, np.float32(dist).item())

    m = tap(0, 0, 0)
    for o in [(1, 0, 0), (0, 1, 0), (0, 0, 1), (-1, 0, 0), (0, -1, 0), (0, 0, -1)]:
        m = torch.minimum(m, tap(*o))
    return m.to(torch.float16)


def build_mip(sdf: torch.Tensor, res: int, half_extent: float, flood: bool = True) -> torch.Tensor:
    """Mip of one cascade: downsample (x4) + 4 flood passes (GlobalDistanceField.cpp:537-573, 825-841)."""
    mres = res // 4
    max_distance = 2.0 * half_extent
    mip = _mip_pass(sdf, mres, 4, res, max_distance)
    if flood:
        for _ in range(4):
            mip = _mip_pass(mip, mres, 1, mres, max_distance)
    return mip.contiguous()


# ------------------------------------------------------------------------------------------------------------------
# Surface cache
# ------------------------------------------------------------------------------------------------------------------
RIGHT = np.array([1.0, 0.0, 0.0])
UP = np.array([0.0, 1.0, 0.0])


def _tile_rotation(face: int) -> np.ndarray:
    """3x3 (row-major) rotation of glm::lookAt for tile `face` (GlobalSurfaceAtlas.cpp:247-262): rows s, u, -f."""
    z_axis = np.zeros(3)
    z_axis[face // 2] = 1.0 if (face & 1) else -1.0
    y_axis = RIGHT if face in (2, 3) else UP
    f = z_axis / np.linalg.norm(z_axis)
    s = np.cross(f, y_axis)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    return np.stack([s, u, -f])


TILE_ROT = [_tile_rotation(i) for i in range(6)]


@dataclass
class BoxObject:
    center: np.ndarray  # (3,)
    half: np.ndarray  # (3,)
    rot_y: float = 0.0
    albedo: tuple = (0.73, 0.73, 0.73)
    emissive: tuple = (0.0, 0.0, 0.0)
    tag: int = 0


def _box_rotation(rot_y: float) -> np.ndarray:
    cs, sn = math.cos(rot_y), math.sin(rot_y)
    return np.array([[cs, 0, sn], [0, 1, 0], [-sn, 0, cs]])  # local -> world


def build_surface_cache(boxes, atlas_res: int, cell: int, chunk_size: float, shade_fn, device="cpu", margin: int = 1, gbuffer: dict = None):
    """Objects/tiles/chunk lists + light and depth atlases for a list of BoxObject.

    Each object gets 6 tiles (one per face) of (cell - 2*margin)^2 texels laid out on a regular grid of `cell`-texel cells.
    Tile index 0 is a dummy because tileOffset == 0 means "no tile" (AtlasCommon.glsl:143).
    shade_fn(world_pos[n,3], world_normal[n,3], obj_index[n]) -> rgb[n,3] float32 torch tensors.
    """
    n = len(boxes)
    per_row = atlas_res // cell
    assert n * 6 <= per_row * per_row, "atlas too small"
    tile_px = cell - 2 * margin

    objects = np.zeros(n, dtype=abi.OBJECT_DTYPE)
    tiles = np.zeros(1 + 6 * n, dtype=abi.TILE_DTYPE)
    rot3 = np.zeros((n, 3, 3))
    for k, b in enumerate(boxes):
        Rm = _box_rotation(b.rot_y)
        rot3[k] = Rm
        m = np.eye(4)
        m[:3, :3] = Rm
        m[:3, 3] = b.center
        objects["objectBounds"][k] = [*b.center, float(np.linalg.norm(b.half))]
        objects["transform"][k] = m.T.reshape(16)  # column-major
        objects["extends"][k] = [*b.half, 1.0]
        for face in range(6):
            t = 1 + k * 6 + face
            objects["tileOffset"][k][face] = t
            cellx, celly = (t - 1) % per_row, (t - 1) // per_row
            x, y = cellx * cell + margin, celly * cell + margin
            Rt = TILE_ROT[face]
            m4 = np.eye(4)
            m4[:3, :3] = Rt
            bounds = np.abs(Rt @ np.maximum(b.half, 0.05)) * 2.0
            tiles["extends"][t] = np.array([x, y, tile_px - 1, tile_px - 1], dtype=np.float64) / atlas_res
            tiles["transform"][t] = m4.T.reshape(16)
            tiles["objectBounds"][t] = [*bounds, 0.0]

    # ---- atlases (vectorised over all tiles) ----
    nt = 6 * n
    dev = torch.device(device)
    light = torch.zeros((atlas_res, atlas_res, 4), dtype=torch.float16, device=dev)
    depth = torch.ones((atlas_res, atlas_res), dtype=torch.float32, device=dev)
    t_idx = torch.arange(nt, device=dev)
    obj_idx = t_idx // 6
    face_idx = t_idx % 6
    centers = torch.tensor(np.stack([b.center for b in boxes]), dtype=torch.float32, device=dev)
    halves = torch.tensor(np.stack([np.maximum(b.half, 0.05) for b in boxes]), dtype=torch.float32, device=dev)
    Rm_t = torch.tensor(rot3, dtype=torch.float32, device=dev)
    Rt_t = torch.tensor(np.stack(TILE_ROT), dtype=torch.float32, device=dev)  # [6,3,3]
    cellx = (t_idx % per_row) * cell + margin
    celly = (t_idx // per_row) * cell + margin
    k = torch.arange(tile_px, device=dev, dtype=torch.float32)
    uv = ((k + 0.5) / float(tile_px - 1)).clamp(0, 1)  # texel centre -> tileUV (inverse of AtlasCommon.glsl:65-67)

    chunk_tiles = max(1, (1 << 22) // (tile_px * tile_px))
    for s in range(0, nt, chunk_tiles):
        e = min(nt, s + chunk_tiles)
        oi, fi = obj_idx[s:e], face_idx[s:e]
        Rt = Rt_t[fi]  # [m,3,3]
        bounds = (Rt @ halves[oi].unsqueeze(-1)).squeeze(-1).abs() * 2.0  # [m,3]
        tpx = (uv[None, None, :] - 0.5) * bounds[:, 0, None, None]  # [m,1,px] along x
        tpy = (uv[None, :, None] - 0.5) * bounds[:, 1, None, None]  # [m,py,1]
        tpx, tpy = torch.broadcast_tensors(tpx, tpy)
        tpz = (0.5 * bounds[:, 2])[:, None, None].expand_as(tpx)
        tp = torch.stack([tpx, tpy, tpz], dim=-1)  # [m,py,px,3] tile space
        lp = torch.einsum("mji,mpqj->mpqi", Rt, tp)  # local = Rt^T * tp
        wp = torch.einsum("mij,mpqj->mpqi", Rm_t[oi], lp) + centers[oi][:, None, None, :]
        nl = Rt[:, 2, :]  # tile +z axis in local space = -f = outward normal of the captured face
        wn = torch.einsum("mij,mj->mi", Rm_t[oi], nl)[:, None, None, :].expand_as(wp)
        rgb = shade_fn(wp.reshape(-1, 3), wn.reshape(-1, 3), oi[:, None, None].expand(wp.shape[:3]).reshape(-1))
        rgb = rgb.reshape(e - s, tile_px, tile_px, 3)
        ys = (celly[s:e, None, None] + torch.arange(tile_px, device=dev)[None, :, None]).expand(e - s, tile_px, tile_px)
        xs = (cellx[s:e, None, None] + torch.arange(tile_px, device=dev)[None, None, :]).expand(e - s, tile_px, tile_px)
        light[ys, xs, :3] = rgb.to(torch.float16)
        light[ys, xs, 3] = 1.0
        depth[ys, xs] = 0.5  # tileDepth of the captured face (tp.z / bounds.z = +1/2), SURVEY A.3
        if gbuffer is not None:  # per-texel surface G-buffer (what SDFDeferredColor.frag captures), for the bounce refresh
            gbuffer.setdefault("texel", []).append((ys * atlas_res + xs).reshape(-1).cpu())
            gbuffer.setdefault("pos", []).append(wp.reshape(-1, 3).cpu())
            gbuffer.setdefault("normal", []).append(wn.reshape(-1, 3).cpu())
            gbuffer.setdefault("object", []).append(oi[:, None, None].expand(wp.shape[:3]).reshape(-1).cpu())

    # ---- chunk lists, SDFCulling.comp:36-101 ----
    NC = abi.CHUNKS_RESOLUTION
    chunks = np.zeros(NC**3, dtype=np.uint32)
    cmin_axis = (np.arange(NC) - NC * 0.5) * chunk_size
    lists = [[] for _ in range(NC**3)]
    for oi_, b in enumerate(boxes):
        c, r = np.asarray(b.center, dtype=np.float64), float(np.float32(np.linalg.norm(b.half)))
        lo = np.clip(np.floor((c - r) / chunk_size + NC * 0.5).astype(int), 0, NC - 1)
        hi = np.clip(np.floor((c + r) / chunk_size + NC * 0.5).astype(int), 0, NC - 1)
        xs_ = np.arange(lo[0], hi[0] + 1)
        ys_ = np.arange(lo[1], hi[1] + 1)
        zs_ = np.arange(lo[2], hi[2] + 1)
        dx = np.clip(c[0], cmin_axis[xs_], cmin_axis[xs_] + chunk_size) - c[0]
        dy = np.clip(c[1], cmin_axis[ys_], cmin_axis[ys_] + chunk_size) - c[1]
        dz = np.clip(c[2], cmin_axis[zs_], cmin_axis[zs_] + chunk_size) - c[2]
        d2 = dz[:, None, None] ** 2 + dy[None, :, None] ** 2 + dx[None, None, :] ** 2
        zi, yi, xi = np.nonzero(d2 <= r * r)  # boxIntersectsSphere, AtlasCommon.glsl:34-38
        addr = (zs_[zi] * NC + ys_[yi]) * NC + xs_[xi]
        for a in addr:
            lists[a].append(oi_)
    cull = [1]  # [0] = allocation counter, starts at 1 (GlobalSurfaceAtlas.cpp:612-613)
    for a, l in enumerate(lists):
        if not l:
            continue
        chunks[a] = len(cull)
        cull.append(len(l))
        cull.extend(l)
    cull[0] = len(cull)
    cull = np.asarray(cull, dtype=np.uint32)

    if gbuffer is not None:
        for k in ("texel", "pos", "normal", "object"):
            gbuffer[k] = torch.cat(gbuffer[k]).numpy()
    data = abi.GlobalSurfaceAtlasData()
    data.cameraPos[:] = [0.0, 0.0, 0.0]
    data.chunkSize = chunk_size
    data.culledObjectsCapacity = len(cull)
    data.resolution = atlas_res
    data.objectsCount = n
    data.padding = 0
    return data, chunks, cull, objects, tiles, light, depth


# ------------------------------------------------------------------------------------------------------------------
# C1: procedural Cornell box
# ------------------------------------------------------------------------------------------------------------------
def cornell_boxes():
    wall = 0.7
    boxes = [
        BoxObject(np.array([0.0, -5.0 - wall, 0.0]), np.array([6.4, wall, 6.4]), 0.0, (0.73, 0.73, 0.73), tag=0),  # floor
        BoxObject(np.array([0.0, 5.0 + wall, 0.0]), np.array([6.4, wall, 6.4]), 0.0, (0.73, 0.73, 0.73), tag=1),  # ceiling
        BoxObject(np.array([-5.0 - wall, 0.0, 0.0]), np.array([wall, 6.4, 6.4]), 0.0, (0.65, 0.05, 0.05), tag=2),  # left, red
        BoxObject(np.array([5.0 + wall, 0.0, 0.0]), np.array([wall, 6.4, 6.4]), 0.0, (0.12, 0.45, 0.15), tag=3),  # right, green
        BoxObject(np.array([0.0, 0.0, -5.0 - wall]), np.array([6.4, 6.4, wall]), 0.0, (0.73, 0.73, 0.73), tag=4),  # back
        BoxObject(np.array([0.0, 0.0, 5.0 + wall]), np.array([6.4, 6.4, wall]), 0.0, (0.73, 0.73, 0.73), tag=5),  # front
        BoxObject(np.array([1.6, -3.5, 1.4]), np.array([1.5, 1.5, 1.5]), math.radians(17.0), (0.73, 0.73, 0.73), tag=6),
        BoxObject(np.array([-1.7, -2.0, -1.5]), np.array([1.5, 3.0, 1.5]), math.radians(-17.0), (0.73, 0.73, 0.73), tag=7),
    ]
    return boxes


CASCADE_DISTANCE_SCALES = (1.0, 2.5, 5.0, 10.0)  # GlobalDistanceField.cpp:600


def make_sdf_data_cascades(centers, half_extents, res: int) -> abi.GlobalSDFData:
    """GlobalSDFData of K nested cascades (GlobalDistanceField.cpp:262-279): cascade k occupies x in [k*res, (k+1)*res) of the volume."""
    d = make_sdf_data(centers[0], half_extents[0], res)
    for k, (c, h) in enumerate(zip(centers, half_extents)):
        d.cascadePosDistance[k][:] = [float(c[0]), float(c[1]), float(c[2]), float(h)]
        d.cascadeVoxelSize[k] = 2.0 * float(h) / res
    d.cascadesCount = len(centers)
    return d


def cornell_scene(res: int = 64, counts=(8, 8, 8), rays: int = 64, atlas_res: int = 512, device="cpu", with_atlas=True,
                  hysteresis=0.98, gamma=5.0, cascades: int = 1) -> Scene:
    """cascades > 1: the reference's nested cascades (half extents in the ratio 1 : 2.5 : 5 : 10, the outermost = the room's 6.4), inner ones
    off-centre; each is the same analytic field sampled on its own grid, side by side along x as the reference lays them out."""
    D = 6.4
    dev = torch.device(device)
    center = (0.0, 0.0, 0.0)
    boxes = cornell_boxes()

    def field(c, half):
        xs, ys, zs = voxel_centers(c, half, res, dev)
        px, py, pz = xs[None, None, :], ys[None, :, None], zs[:, None, None]
        d = -sd_box(px, py, pz, (0.0, 0.0, 0.0), (5.0, 5.0, 5.0))
        for b in boxes[6:]:
            d = torch.minimum(d, sd_box(px, py, pz, b.center, b.half, b.rot_y))
        f = encode_sdf(d.expand(res, res, res), half).contiguous()
        return f, build_mip(f, res, half)

    if cascades == 1:
        sdf, mip = field(center, D)
        sdf_data = make_sdf_data(center, D, res)
    else:
        assert 1 < cascades <= len(CASCADE_DISTANCE_SCALES)
        halves = [D * CASCADE_DISTANCE_SCALES[k] / CASCADE_DISTANCE_SCALES[cascades - 1] for k in range(cascades)]
        centers = [tuple(float(np.float32(o * (D - h))) + 0.0 for o in (0.25, -0.4, 0.15)) for h in halves]  # the outermost stays at the origin
        parts = [field(c, h) for c, h in zip(centers, halves)]
        sdf = torch.cat([p[0] for p in parts], dim=2).contiguous()
        mip = torch.cat([p[1] for p in parts], dim=2).contiguous()
        sdf_data = make_sdf_data_cascades(centers, halves, res)

    span = 8.4
    step = [span / (counts[0] - 1) if counts[0] > 1 else 1.2, span / (counts[1] - 1) if counts[1] > 1 else 1.2,
            span / (counts[2] - 1) if counts[2] > 1 else 1.2]
    if tuple(counts) == (8, 8, 8):
        step = [1.2, 1.2, 1.2]
    uni = abi.make_uniform((-4.2, -4.2, -4.2), step, counts, rays, hysteresis=hysteresis, gamma=gamma)
    sc = Scene("cornell", uni, sdf_data, sdf, mip)

    if with_atlas:
        albedo = torch.tensor([b.albedo for b in boxes], dtype=torch.float32, device=dev)
        light_pos = torch.tensor([0.0, 4.6, 0.0], device=dev)

        def shade(wp, wn, oi):
            l = light_pos - wp
            r2 = (l * l).sum(-1, keepdim=True).clamp(min=0.25)
            ndl = ((l * wn).sum(-1, keepdim=True) / torch.sqrt(r2)).clamp(min=0)
            rgb = albedo[oi] * (40.0 * ndl / r2)
            panel = (oi == 1) & (wp[:, 0].abs() < 1.5) & (wp[:, 2].abs() < 1.5) & (wn[:, 1] < -0.5)
            return torch.where(panel[:, None], torch.full_like(rgb, 15.0), rgb)

        cell = atlas_res // 8
        gb = {}
        (sc.atlas_data, sc.chunks, sc.cull, sc.objects, sc.tiles, sc.light, sc.depth) = build_surface_cache(
            boxes, atlas_res, cell, 2.0 * D / abi.CHUNKS_RESOLUTION, shade, device=device, gbuffer=gb)
        alb = np.asarray([b.albedo for b in boxes], dtype=np.float32)
        gb["albedo"] = alb[gb["object"]]
        gb["metallic"] = np.zeros(len(gb["texel"]), dtype=np.float32)
        sc.meta["gbuffer"] = gb
    sc.meta.update({"D": D, "res": res, "voxel": 2 * D / res})
    return sc


# ------------------------------------------------------------------------------------------------------------------
# C4 / C5: synthetic city
# ------------------------------------------------------------------------------------------------------------------
def city_boxes(lots: int, half_extent: float, seed: int = 7):
    """lots x lots buildings: footprint U(4,9), height U(6,120) (scaled down for cascades smaller than 512^3)."""
    rng = SplitMix64(seed)
    pitch = 2.0 * half_extent / lots
    hscale = min(1.0, half_extent / 256.0)
    boxes = []
    for j in range(lots):
        for i in range(lots):
            fx, fz = rng.uniform(4, 9), rng.uniform(4, 9)
            h = rng.uniform(6, 120) * hscale
            emissive = rng.uniform() < 0.05
            hue = rng.uniform()
            cx = -half_extent + (i + 0.5) * pitch
            cz = -half_extent + (j + 0.5) * pitch
            alb = (0.35 + 0.4 * hue, 0.45 + 0.2 * (1 - hue), 0.4 + 0.3 * abs(0.5 - hue))
            boxes.append(BoxObject(np.array([cx, h * 0.5, cz]), np.array([fx * 0.5, h * 0.5, fz * 0.5]), 0.0, alb,
                                   (10.0, 8.0, 5.0) if emissive else (0.0, 0.0, 0.0), tag=len(boxes)))
    return boxes, pitch


def city_scene(res: int = 512, lots: int = 48, counts=(64, 16, 64), rays: int = 512, atlas_res: int = 4096, device="cpu",
               with_atlas=True, hysteresis=0.98, gamma=5.0, start=None, step=(7.5, 8.0, 7.5), seed: int = 7) -> Scene:
    """Synthetic city: voxel 1.0, D = res/2, cascade centred at (0, D - 8, 0) so the ground slab
    occupies the 8 lowest voxel layers."""
    D = res * 0.5
    dev = torch.device(device)
    boxes, pitch = city_boxes(lots, D, seed)
    center = (0.0, D - 8.0, 0.0)
    vox = 2.0 * D / res
    bc = torch.tensor(np.stack([b.center for b in boxes]), dtype=torch.float32, device=dev).reshape(lots, lots, 3)
    bh = torch.tensor(np.stack([b.half for b in boxes]), dtype=torch.float32, device=dev).reshape(lots, lots, 3)
    xs, ys, zs = voxel_centers(center, D, res, dev)
    reach = 4
    cap = 2.0 * D  # clamp() maps it to 1.0
    lot_x = ((xs + D) / pitch).floor().long().clamp(0, lots - 1)  # [res]
    lot_z = ((zs + D) / pitch).floor().long().clamp(0, lots - 1)
    sdf = torch.empty((res, res, res), dtype=torch.float16, device=dev)
    y_chunk = 8 if dev.type == "cuda" else 4
    # horizontal part per neighbouring lot, shared by all y
    offs = [(dj, di) for dj in range(-reach, reach + 1) for di in range(-reach, reach + 1)]
    z_chunk = 64 if dev.type == "cuda" else 16
    for z0 in range(0, res, z_chunk):
        z1 = min(res, z0 + z_chunk)
        zz = zs[z0:z1]
        lz = lot_z[z0:z1]
        best = torch.full((z1 - z0, res, res), cap, dtype=torch.float32, device=dev)  # [z][y][x]
        for dj, di in offs:
            jj = (lz + dj)
            ii = (lot_x + di)
            valid = ((jj >= 0) & (jj < lots))[:, None] & ((ii >= 0) & (ii < lots))[None, :]  # [z][x]
            jj, ii = jj.clamp(0, lots - 1), ii.clamp(0, lots - 1)
            c = bc[jj][:, ii]  # [z][x][3]
            h = bh[jj][:, ii]
            qx = (xs[None, :] - c[..., 0]).abs() - h[..., 0]  # [z][x]
            qz = (zz[:, None] - c[..., 2]).abs() - h[..., 2]
            qx = torch.where(valid, qx, torch.full_like(qx, cap))
            qy = (ys[None, :, None] - c[..., 1][:, None, :]).abs() - h[..., 1][:, None, :]  # [z][y][x]
            qxb, qzb = qx[:, None, :], qz[:, None, :]
            outside = torch.sqrt(qxb.clamp(min=0) ** 2 + qy.clamp(min=0) ** 2 + qzb.clamp(min=0) ** 2)
            inside = torch.maximum(qxb, torch.maximum(qy, qzb)).clamp(max=0)
            best = torch.minimum(best, outside + inside)
        ground = ys[None, :, None].expand_as(best)  # plane y = 0, solid below
        best = torch.minimum(best, ground)
        sdf[z0:z1] = encode_sdf(best, D)
    mip = build_mip(sdf, res, D)

    if start is None:
        ext = [step[0] * (counts[0] - 1), step[2] * (counts[2] - 1)]
        start = (-ext[0] * 0.5, 2.0, -ext[1] * 0.5)
    uni = abi.make_uniform(start, step, counts, rays, hysteresis=hysteresis, gamma=gamma)
    sc = Scene("city", uni, make_sdf_data(center, D, res), sdf, mip)
    sc.meta = {"D": D, "res": res, "voxel": vox, "lots": lots}

    # sky: constant horizon blue, 1x1 faces
    sky = np.zeros((6, 1, 1, 4), dtype=np.float16)
    sky[..., 0], sky[..., 1], sky[..., 2], sky[..., 3] = 0.5, 0.7, 1.0, 1.0
    sc.sky_face, sc.sky = 1, sky

    if with_atlas:
        ground = BoxObject(np.array([0.0, -8.0, 0.0]), np.array([D, 8.0, D]), 0.0, (0.3, 0.3, 0.32), tag=len(boxes))
        allb = boxes + [ground]
        albedo = torch.tensor([b.albedo for b in allb], dtype=torch.float32, device=dev)
        emis = torch.tensor([b.emissive for b in allb], dtype=torch.float32, device=dev)
        sun = torch.tensor([0.35, 0.8, 0.48], device=dev)
        sun = sun / sun.norm()

        def shade(wp, wn, oi):
            ndl = (wn * sun).sum(-1, keepdim=True).clamp(min=0)
            rgb = albedo[oi] * (3.0 * ndl + 0.3)
            win = ((wp[:, 1] * 0.5).floor().long() % 2 == 0) & (wn[:, 1].abs() < 0.5)  # window bands on side faces
            return rgb + torch.where(win[:, None], emis[oi], torch.zeros_like(rgb))

        (sc.atlas_data, sc.chunks, sc.cull, sc.objects, sc.tiles, sc.light, sc.depth) = build_surface_cache(
            allb, atlas_res, 32, 2.0 * D / abi.CHUNKS_RESOLUTION * 1.0, shade, device=device)
    return sc


# ------------------------------------------------------------------------------------------------------------------
# C2 / C3: the reference's dark-room-emissive scene (BASELINE configs[1], [2])
# ------------------------------------------------------------------------------------------------------------------
DARK_ROOM_FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c2_dark_room.npz")


def dark_room_scene(counts=(16, 8, 16), rays: int = 256, atlas_res: int = 2048, device="cpu", with_atlas=True, hysteresis=0.98, gamma=0.85,
                    fixture: str = DARK_ROOM_FIXTURE) -> Scene:
    """Global SDF of Assets/dark-room-emissive.scene merged from its 100 baked mesh fields (tests/golden/make_c2_dark_room.py wrote the
    fixture in the build container: 128^3, half extent 88, centre 0).  Probe volume over the world AABB of the meshes as
    ddgi::on_game_start lays it out (start = AABB min, counts as configured, step = extent / (counts - 1)); gamma 0.85 is the shipped
    scene's ddgiGamma.  Surface cache: one OBB object per mesh field (its world AABB), six tiles each; albedo from a hash of the mesh name,
    emissive 10 on meshes whose name contains "Light" / "Emissive", one point light under the ceiling."""
    import zlib

    fx = np.load(fixture)
    D, res = float(fx["half_extent"]), int(fx["resolution"])
    dev = torch.device(device)
    sdf = torch.from_numpy(fx["sdf"].view(np.float16).copy()).to(dev)
    mip = torch.from_numpy(fx["mip"].view(np.float16).copy()).to(dev)
    cen, half = fx["box_center"].astype(np.float64), fx["box_half"].astype(np.float64)
    lo, hi = (cen - half).min(0), (cen + half).max(0)
    inset = 2.0
    start = lo + inset
    step = (hi - lo - 2 * inset) / (np.asarray(counts, dtype=np.float64) - 1)
    uni = abi.make_uniform(tuple(float(x) for x in start), tuple(float(x) for x in step), counts, rays, hysteresis=hysteresis, gamma=gamma)
    sc = Scene("dark-room", uni, make_sdf_data((0.0, 0.0, 0.0), D, res), sdf, mip)
    sc.meta = {"D": D, "res": res, "voxel": 2 * D / res, "build_stats": [int(x) for x in fx["stats"]]}
    if with_atlas:
        boxes = []
        for k, (c, h, name, em) in enumerate(zip(cen, half, fx["names"], fx["emissive"])):
            hsh = zlib.crc32(str(name).encode())
            albedo = tuple(0.25 + 0.6 * ((hsh >> (8 * i)) & 255) / 255.0 for i in range(3))
            boxes.append(BoxObject(c, np.maximum(h, 0.05), 0.0, albedo, (10.0, 10.0, 10.0) if em else (0.0, 0.0, 0.0), tag=k))
        albedo_t = torch.tensor([b.albedo for b in boxes], dtype=torch.float32, device=dev)
        emis_t = torch.tensor([b.emissive for b in boxes], dtype=torch.float32, device=dev)
        light_pos = torch.tensor([float((lo[0] + hi[0]) * 0.5), float(hi[1] - 6.0), float((lo[2] + hi[2]) * 0.5)], device=dev)

        def shade(wp, wn, oi):
            l = light_pos - wp
            r2 = (l * l).sum(-1, keepdim=True).clamp(min=4.0)
            ndl = ((l * wn).sum(-1, keepdim=True) / torch.sqrt(r2)).clamp(min=0)
            return albedo_t[oi] * (1500.0 * ndl / r2) + emis_t[oi]

        per_row = int(math.ceil(math.sqrt(len(boxes) * 6)))
        cell = atlas_res // per_row
        gb = {}
        (sc.atlas_data, sc.chunks, sc.cull, sc.objects, sc.tiles, sc.light, sc.depth) = build_surface_cache(
            boxes, atlas_res, cell, 2.0 * D / abi.CHUNKS_RESOLUTION, shade, device=device, gbuffer=gb)
        alb = np.asarray([b.albedo for b in boxes], dtype=np.float32)
        gb["albedo"] = alb[gb["object"]]
        gb["metallic"] = np.zeros(len(gb["texel"]), dtype=np.float32)
        sc.meta["gbuffer"] = gb
    return sc


CONFIGS = {
    # name: (builder, kwargs) — SURVEY.md §8d
    "c1": (cornell_scene, dict(res=64, counts=(8, 8, 8), rays=64, atlas_res=512)),
    "c2": (dark_room_scene, dict(counts=(16, 8, 16), rays=256, atlas_res=2048)),
    "c3": (dark_room_scene, dict(counts=(32, 16, 32), rays=256, atlas_res=2048)),
    "c4": (city_scene, dict(res=512, lots=48, counts=(64, 16, 64), rays=512, atlas_res=4096)),
    "c5": (city_scene, dict(res=1024, lots=96, counts=(128, 32, 128), rays=1024, atlas_res=8192)),
    # reduced cities for tests
    "city64": (city_scene, dict(res=64, lots=6, counts=(8, 4, 8), rays=64, atlas_res=512)),
    "city128": (city_scene, dict(res=128, lots=12, counts=(16, 8, 16), rays=128, atlas_res=1024)),
}


def build(name: str, device="cpu", **over) -> Scene:
    fn, kw = CONFIGS[name]
    kw = dict(kw)
    kw.update(over)
    return fn(device=device, **kw)
