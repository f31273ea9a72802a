"""ctypes binding of oracle/libddgi_oracle.so — TEST INFRASTRUCTURE.

Import this only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (luxgi_b200) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from luxgi_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libddgi_oracle.so")
_lib = None


class SceneDesc(C.Structure):
    _fields_ = [
        ("ddgi", C.POINTER(abi.DDGIUniform)),
        ("sdfData", C.POINTER(abi.GlobalSDFData)),
        ("sdf", C.c_void_p),
        ("mip", C.c_void_p),
        ("atlasData", C.POINTER(abi.GlobalSurfaceAtlasData)),
        ("chunks", C.c_void_p),
        ("cull", C.c_void_p),
        ("objects", C.c_void_p),
        ("tiles", C.c_void_p),
        ("light", C.c_void_p),
        ("depth", C.c_void_p),
        ("skyFace", C.c_int32),
        ("sky", C.c_void_p),
    ]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("mipTaps", "texTaps", "hits", "tileSamples", "steps", "objectsVisited")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "ddgi_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "all"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_threads.restype = C.c_int
        L.oracle_f2h.restype = C.c_uint16
        L.oracle_f2h.argtypes = [C.c_float]
        L.oracle_h2f.restype = C.c_float
        L.oracle_h2f.argtypes = [C.c_uint16]
        L.oracle_sample3d.restype = C.c_float
        L.oracle_sample3d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        L.oracle_trace.restype = C.c_int
        L.oracle_trace.argtypes = [C.POINTER(SceneDesc), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.POINTER(Counters)]
        L.oracle_blend.restype = C.c_int
        L.oracle_blend.argtypes = [C.POINTER(abi.DDGIUniform), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_blend_ids.restype = C.c_int
        L.oracle_blend_ids.argtypes = [C.POINTER(abi.DDGIUniform), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.oracle_border_ids.restype = C.c_int
        L.oracle_border_ids.argtypes = [C.POINTER(abi.DDGIUniform), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_border.restype = C.c_int
        L.oracle_border.argtypes = [C.POINTER(abi.DDGIUniform), C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.oracle_border_offsets.restype = C.c_int
        L.oracle_border_offsets.argtypes = [C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _np(t, dtype=None):
    """torch tensor / numpy array -> contiguous numpy (fp16 viewed as uint16)."""
    if t is None:
        return None
    if hasattr(t, "detach"):
        t = t.detach().cpu().contiguous().numpy()
    a = np.ascontiguousarray(t)
    if a.dtype == np.float16:
        a = a.view(np.uint16)
    if dtype is not None and a.dtype != dtype:
        a = a.astype(dtype)
    return a


class OracleScene:
    """Host copies of a luxgi_b200.scenes.Scene in the layout the oracle reads."""

    def __init__(self, scene):
        self.scene = scene
        self.uniform = scene.uniform
        self.sdf = _np(scene.sdf)
        self.mip = _np(scene.mip)
        self.chunks = _np(scene.chunks)
        self.cull = _np(scene.cull)
        self.objects = _np(scene.objects)
        self.tiles = _np(scene.tiles)
        self.light = _np(scene.light)
        self.depth = _np(scene.depth)
        self.sky = _np(scene.sky)
        d = SceneDesc()
        d.ddgi = C.pointer(scene.uniform)
        d.sdfData = C.pointer(scene.sdf_data)
        d.sdf, d.mip = _ptr(self.sdf), _ptr(self.mip)
        if scene.atlas_data is not None:
            d.atlasData = C.pointer(scene.atlas_data)
            d.chunks, d.cull, d.objects, d.tiles = _ptr(self.chunks), _ptr(self.cull), _ptr(self.objects), _ptr(self.tiles)
            d.light, d.depth = _ptr(self.light), _ptr(self.depth)
        d.skyFace = scene.sky_face
        d.sky = _ptr(self.sky)
        self.desc = d

    def trace(self, rot16, probe_begin=0, count=None, probe_ids=None, want_steps=False):
        """Returns (radiance u16 [n][R][4], dirDist u16 [n][R][4], steps u16 [n][R] | None, counters dict)."""
        R = self.uniform.raysPerProbe
        ids = None
        if probe_ids is not None:
            ids = np.ascontiguousarray(probe_ids, dtype=np.int32)
            count = len(ids)
        elif count is None:
            count = abi.probe_count(self.uniform) - probe_begin
        rad = np.zeros((count, R, 4), dtype=np.uint16)
        dd = np.zeros((count, R, 4), dtype=np.uint16)
        steps = np.zeros((count, R), dtype=np.uint16) if want_steps else None
        rot = np.ascontiguousarray(rot16, dtype=np.float32)
        cn = Counters()
        rc = lib().oracle_trace(C.byref(self.desc), _ptr(rot), probe_begin, count, _ptr(ids), _ptr(rad), _ptr(dd), _ptr(steps), C.byref(cn))
        assert rc == 0, rc
        return rad, dd, steps, cn.as_dict()


def new_atlases(u):
    irr = np.zeros((u.irradianceTextureHeight, u.irradianceTextureWidth, 4), dtype=np.uint16)
    dep = np.zeros((u.depthTextureHeight, u.depthTextureWidth, 2), dtype=np.uint16)
    return irr, dep


def blend(u, rad, dd, prev_irr, prev_dep, out_irr, out_dep, first_frame, probe_begin=0, count=None, ray_row_offset=0, naive=False):
    if count is None:
        count = rad.shape[0]
    rc = lib().oracle_blend(C.byref(u), _ptr(rad), _ptr(dd), ray_row_offset, _ptr(prev_irr), _ptr(prev_dep), _ptr(out_irr),
                            _ptr(out_dep), int(bool(first_frame)), probe_begin, count, int(bool(naive)))
    assert rc == 0, rc


def blend_ids(u, rad, dd, prev_irr, prev_dep, out_irr, out_dep, first_frame, probe_ids, naive=False):
    """Blend + border of a probe list: row k of rad/dd belongs to probe probe_ids[k]."""
    ids = np.ascontiguousarray(probe_ids, dtype=np.int32)
    rc = lib().oracle_blend_ids(C.byref(u), _ptr(rad), _ptr(dd), 0, _ptr(prev_irr), _ptr(prev_dep), _ptr(out_irr), _ptr(out_dep),
                                int(bool(first_frame)), 0, len(ids), _ptr(ids), int(bool(naive)))
    assert rc == 0, rc
    rc = lib().oracle_border_ids(C.byref(u), _ptr(out_irr), _ptr(out_dep), _ptr(ids), len(ids))
    assert rc == 0, rc


def sample_irradiance(u, irr, dep, P, N, Wo):
    P, N, Wo = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in (P, N, Wo))
    out = np.empty_like(P)
    L = lib()
    L.oracle_sample_irradiance.restype = C.c_int
    L.oracle_sample_irradiance.argtypes = [C.POINTER(abi.DDGIUniform), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.oracle_sample_irradiance(C.byref(u), _ptr(irr), _ptr(dep), len(P), _ptr(P), _ptr(N), _ptr(Wo), _ptr(out))
    assert rc == 0, rc
    return out


def sample_probe(u, irr, dep, g_depth, g_normal, camera_position, view_proj_inv):
    g_depth = np.ascontiguousarray(g_depth, dtype=np.float32)
    g_normal = np.ascontiguousarray(g_normal, dtype=np.float32)
    h, w = g_depth.shape
    cam = np.ascontiguousarray(camera_position, dtype=np.float32).reshape(4)
    vpi = np.ascontiguousarray(view_proj_inv, dtype=np.float32).reshape(16)
    out = np.empty((h, w, 4), dtype=np.float32)
    L = lib()
    L.oracle_sample_probe.restype = C.c_int
    L.oracle_sample_probe.argtypes = [C.POINTER(abi.DDGIUniform), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.oracle_sample_probe(C.byref(u), _ptr(irr), _ptr(dep), w, h, _ptr(g_depth), _ptr(g_normal), _ptr(cam), _ptr(vpi), _ptr(out))
    assert rc == 0, rc
    return out


def indirect_light(u, irr, dep, light, base, texel, P, N, albedo, metallic, intensity, camera_pos):
    """In-place indirect-light refresh of `light` (uint16 RGBA16F atlas) for the listed texels."""
    texel = np.ascontiguousarray(texel, dtype=np.uint32)
    P, N, albedo = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in (P, N, albedo))
    metallic = np.ascontiguousarray(metallic, dtype=np.float32)
    cam = np.ascontiguousarray(camera_pos, dtype=np.float32).reshape(3)
    L = lib()
    L.oracle_indirect_light.restype = C.c_int
    L.oracle_indirect_light.argtypes = [C.POINTER(abi.DDGIUniform)] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 5 + [C.c_float, C.c_void_p]
    rc = L.oracle_indirect_light(C.byref(u), _ptr(irr), _ptr(dep), _ptr(light), _ptr(base), len(texel), _ptr(texel), _ptr(P), _ptr(N), _ptr(albedo),
                                 _ptr(metallic), float(intensity), _ptr(cam))
    assert rc == 0, rc


def set_unfused(on):
    """Literal two-rounding blend arithmetic of the shipped SPIR-V instead of the contract's explicit FMAs."""
    lib().oracle_set_unfused(int(bool(on)))


def border(u, irr, dep, probe_begin=0, count=None):
    if count is None:
        count = abi.probe_count(u)
    rc = lib().oracle_border(C.byref(u), _ptr(irr), _ptr(dep), probe_begin, count)
    assert rc == 0, rc


class OraclePipeline:
    """The reference's frame protocol (DDGIRenderer.cpp:333-343,390-395) driven through the oracle."""

    def __init__(self, scene, probe_begin=0, count=None):
        self.os = scene if isinstance(scene, OracleScene) else OracleScene(scene)
        self.u = self.os.uniform
        self.irr = [new_atlases(self.u)[0] for _ in range(2)]
        self.dep = [new_atlases(self.u)[1] for _ in range(2)]
        self.frames = 0
        self.ping = 0
        self.probe_begin = probe_begin
        self.count = abi.probe_count(self.u) - probe_begin if count is None else count
        self.rad = self.dd = None
        self.counters = None

    def update(self, rot16, naive=False):
        self.rad, self.dd, _, self.counters = self.os.trace(rot16, self.probe_begin, self.count)
        w = 1 - self.ping
        blend(self.u, self.rad, self.dd, self.irr[self.ping], self.dep[self.ping], self.irr[w], self.dep[w], self.frames == 0,
              self.probe_begin, self.count, ray_row_offset=self.probe_begin, naive=naive)
        border(self.u, self.irr[w], self.dep[w], self.probe_begin, self.count)
        self.ping = 1 - self.ping
        self.frames += 1

    @property
    def irradiance(self):
        return self.irr[self.ping]

    @property
    def depth(self):
        return self.dep[self.ping]


# ----------------------------------------------------------------------------------------------------------------------
# global SDF build (row f3)
# ----------------------------------------------------------------------------------------------------------------------
def sdf_object_data(mesh, cascade_level):
    """ObjectRasterizeData of a luxgi_b200.meshsdf.MeshSDF for one cascade level (chunkCalculate)."""
    from luxgi_b200 import meshsdf

    arr, keep = meshsdf.to_ctypes([mesh])
    out = abi.ObjectRasterizeData()
    L = lib()
    L.oracle_sdf_object_data.restype = None
    L.oracle_sdf_object_data.argtypes = [C.POINTER(abi.MeshSDF), C.c_int, C.POINTER(abi.ObjectRasterizeData)]
    L.oracle_sdf_object_data(arr, int(cascade_level), C.byref(out))
    return out


def sdf_rasterize_chunk(sdf_bits, objs, meshes, mip_level, centre, D, res, cascade, chunk, ids, read, groups=None):
    """One SDFRasterizeModel dispatch on `sdf_bits` (uint16 [res][res][res*cascades]) in place."""
    L = lib()
    voxel = np.float32(2 * D) / np.float32(res)
    mul = (C.c_float * 3)(*[float(np.float32(2 * D) / np.float32(res))] * 3)
    add = (C.c_float * 3)(*[float(np.float32(c) - np.float32(D) + voxel * np.float32(0.5)) for c in centre])
    oarr = (abi.ObjectRasterizeData * len(objs))(*objs)
    lv = [np.ascontiguousarray(m.levels[mip_level], dtype=np.float16) for m in meshes]
    ptrs = (C.c_void_p * len(lv))(*[a.ctypes.data for a in lv])
    sizes = np.ascontiguousarray([[a.shape[2], a.shape[1], a.shape[0]] for a in lv], dtype=np.int32)
    cc = (C.c_int32 * 3)(*[int(x) for x in chunk])
    idarr = (C.c_uint32 * max(1, len(ids)))(*[int(i) for i in ids])
    L.oracle_sdf_rasterize_chunk.restype = C.c_int
    L.oracle_sdf_rasterize_chunk.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rc = L.oracle_sdf_rasterize_chunk(mul, add, float(np.float32(2 * D)), int(res), int(cascade), cc, len(ids), idarr, int(bool(read)), oarr, ptrs,
                                      _ptr(sizes), _ptr(sdf_bits), int(sdf_bits.shape[2]))
    assert rc == 0, rc


def sdf_mip_pass(src, dst, out_res, res, scale, tex_off, mip_off, max_distance):
    L = lib()
    L.oracle_sdf_mip_pass.restype = C.c_int
    L.oracle_sdf_mip_pass.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
    rc = L.oracle_sdf_mip_pass(_ptr(src), src.shape[2], src.shape[1], _ptr(dst), dst.shape[2], dst.shape[1], int(out_res), int(res), int(scale),
                               int(tex_off), int(mip_off), float(max_distance))
    assert rc == 0, rc


def sdf_build_mip(sdf_data, sdf_bits):
    res, casc = int(sdf_data.resolution), int(sdf_data.cascadesCount)
    mip = np.zeros((res // 4, res // 4, res // 4 * casc), dtype=np.uint16)
    L = lib()
    L.oracle_sdf_build_mip.restype = C.c_int
    L.oracle_sdf_build_mip.argtypes = [C.POINTER(abi.GlobalSDFData), C.c_void_p, C.c_void_p]
    assert L.oracle_sdf_build_mip(C.byref(sdf_data), _ptr(sdf_bits), _ptr(mip)) == 0
    return mip


def sdf_build(sdf_data, meshes, min_object_radius=0.0):
    """One-shot global SDF build -> (sdf uint16 [res][res][res*casc], mip uint16, stats dict)."""
    from luxgi_b200 import meshsdf

    res, casc = int(sdf_data.resolution), int(sdf_data.cascadesCount)
    sdf = np.zeros((res, res, res * casc), dtype=np.uint16)
    mip = np.zeros((res // 4, res // 4, res // 4 * casc), dtype=np.uint16)
    arr, keep = meshsdf.to_ctypes(meshes)
    stats = (C.c_int32 * 4)()
    L = lib()
    L.oracle_sdf_build.restype = C.c_int
    L.oracle_sdf_build.argtypes = [C.POINTER(abi.GlobalSDFData), C.POINTER(abi.MeshSDF), C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.oracle_sdf_build(C.byref(sdf_data), arr, len(meshes), float(min_object_radius), _ptr(sdf), _ptr(mip), stats)
    assert rc == 0, rc
    return sdf, mip, dict(zip(("chunks", "models", "dropped_by_overflow", "chunks_out_of_range"), [int(x) for x in stats]))


def surface_cull(atlas_data, objects, order=None, emulate_slot0=False, capacity_words=None):
    """SDFCulling.comp restated: -> (chunks uint32[64000], cull uint32[capacity_words])."""
    n = abi.CHUNKS_RESOLUTION ** 3
    cap = int(capacity_words if capacity_words is not None else max(int(atlas_data.culledObjectsCapacity), 1))
    chunks = np.zeros(n, dtype=np.uint32)
    cull = np.zeros(cap, dtype=np.uint32)
    objs = np.ascontiguousarray(objects)
    o = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
    L = lib()
    L.oracle_surface_cull.restype = C.c_int
    L.oracle_surface_cull.argtypes = [C.POINTER(abi.GlobalSurfaceAtlasData), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32]
    rc = L.oracle_surface_cull(C.byref(atlas_data), _ptr(objs), _ptr(o), 0 if o is None else len(o), int(bool(emulate_slot0)), _ptr(chunks), _ptr(cull), cap)
    assert rc == 0, rc
    return chunks, cull


def trace_global_sdf(sdf_data, sdf, mip, traces, start_bias=0.0):
    """tracyGlobalSDF for arbitrary rays (abi.SDF_TRACE_DTYPE records) -> abi.SDF_HIT_DTYPE records."""
    traces = np.ascontiguousarray(traces, dtype=abi.SDF_TRACE_DTYPE)
    hits = np.zeros(len(traces), dtype=abi.SDF_HIT_DTYPE)
    s, m = _np(sdf), _np(mip)
    L = lib()
    L.oracle_trace_global_sdf.restype = C.c_int
    L.oracle_trace_global_sdf.argtypes = [C.POINTER(abi.GlobalSDFData), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p]
    rc = L.oracle_trace_global_sdf(C.byref(sdf_data), _ptr(s), _ptr(m), len(traces), _ptr(traces), float(start_bias), _ptr(hits))
    assert rc == 0, rc
    return hits


Light, make_light = abi.Light, abi.make_light


def octohedral_to_direction(e):
    e = np.ascontiguousarray(e, dtype=np.float32).reshape(-1, 2)
    out = np.zeros((len(e), 3), dtype=np.float32)
    L = lib()
    L.oracle_octohedral_to_direction.restype = None
    L.oracle_octohedral_to_direction.argtypes = [C.c_float, C.c_float, C.c_void_p]
    for i in range(len(e)):
        L.oracle_octohedral_to_direction(float(e[i, 0]), float(e[i, 1]), out[i].ctypes.data_as(C.c_void_p))
    return out


def surface_direct_light(sdf_data, sdf, mip, light, camera_pos_bias, light_cache, texel, P, N, albedo, metallic_roughness):
    """SDFDeferredLight.frag for the listed atlas texels, blended additively into `light_cache` (uint16 RGBA16F atlas) in place."""
    texel = np.ascontiguousarray(texel, dtype=np.uint32)
    P, N, albedo = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in (P, N, albedo))
    mr = np.ascontiguousarray(metallic_roughness, dtype=np.float32).reshape(-1, 2)
    cam = np.ascontiguousarray(camera_pos_bias, dtype=np.float32).reshape(4)
    s, m = _np(sdf), _np(mip)
    L = lib()
    L.oracle_surface_direct_light.restype = C.c_int
    L.oracle_surface_direct_light.argtypes = [C.POINTER(abi.GlobalSDFData), C.c_void_p, C.c_void_p, C.POINTER(Light), C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
    rc = L.oracle_surface_direct_light(C.byref(sdf_data), _ptr(s), _ptr(m), C.byref(light), _ptr(cam), _ptr(light_cache), len(texel), _ptr(texel), _ptr(P), _ptr(N),
                                       _ptr(albedo), _ptr(mr))
    assert rc == 0, rc


def _noise(sobol, scrambling):
    sobol = np.ascontiguousarray(sobol, dtype=np.uint8).reshape(256, 4)
    scrambling = np.ascontiguousarray(scrambling, dtype=np.uint8).reshape(128, 128, 4)
    return sobol, scrambling


def sdf_reflection(scene, irr, dep, push, g_depth, g_normal, g_pbr, sobol, scrambling, out):
    """SDFReflection.comp over a G-buffer; `out` (uint16 [h][w][4], RGBA16F) is updated in place.  scene: OracleScene or Scene."""
    osc = scene if isinstance(scene, OracleScene) else OracleScene(scene)
    g_depth = np.ascontiguousarray(g_depth, dtype=np.float32)
    h, w = g_depth.shape
    g_normal = np.ascontiguousarray(g_normal, dtype=np.float32).reshape(h, w, 4)
    g_pbr = np.ascontiguousarray(g_pbr, dtype=np.float32).reshape(h, w, 4)
    sobol, scrambling = _noise(sobol, scrambling)
    assert out.dtype == np.uint16 and out.shape == (h, w, 4) and out.flags.c_contiguous
    L = lib()
    L.oracle_sdf_reflection.restype = C.c_int
    L.oracle_sdf_reflection.argtypes = [C.POINTER(SceneDesc), C.c_void_p, C.c_void_p, C.POINTER(abi.ReflectionPushConstants), C.c_int, C.c_int] + [C.c_void_p] * 6
    rc = L.oracle_sdf_reflection(C.byref(osc.desc), _ptr(irr), _ptr(dep), C.byref(push), w, h, _ptr(g_depth), _ptr(g_normal), _ptr(g_pbr), _ptr(sobol),
                                 _ptr(scrambling), _ptr(out))
    assert rc == 0, rc


def sdf_shadow(sdf_data, sdf, mip, light, view_proj_inv, num_frames, shadow_bias, g_depth, g_normal, sobol, scrambling, out_mask):
    """SDFShadow.comp over a G-buffer; `out_mask` (uint32 [h/4][w/8]) is updated in place."""
    g_depth = np.ascontiguousarray(g_depth, dtype=np.float32)
    h, w = g_depth.shape
    g_normal = np.ascontiguousarray(g_normal, dtype=np.float32).reshape(h, w, 4)
    vpi = np.ascontiguousarray(view_proj_inv, dtype=np.float32).reshape(16)
    sobol, scrambling = _noise(sobol, scrambling)
    assert out_mask.dtype == np.uint32 and out_mask.shape == (h // 4, w // 8) and out_mask.flags.c_contiguous
    s, m = _np(sdf), _np(mip)
    L = lib()
    L.oracle_sdf_shadow.restype = C.c_int
    L.oracle_sdf_shadow.argtypes = [C.POINTER(abi.GlobalSDFData), C.c_void_p, C.c_void_p, C.POINTER(abi.Light), C.c_void_p, C.c_uint32, C.c_float, C.c_int, C.c_int] + [C.c_void_p] * 5
    rc = L.oracle_sdf_shadow(C.byref(sdf_data), _ptr(s), _ptr(m), C.byref(light), _ptr(vpi), int(num_frames), float(shadow_bias), w, h, _ptr(g_depth), _ptr(g_normal),
                             _ptr(sobol), _ptr(scrambling), _ptr(out_mask))
    assert rc == 0, rc


def open_space_stats(sdf_data, sdf, mip, traces, start_bias=0.0, cell=8):
    """Validation aid for the engine's open-space table (OpenTable in ddgi_oracle.cpp): -> (dict of counts, table bits uint32[])."""
    traces = np.ascontiguousarray(traces, dtype=abi.SDF_TRACE_DTYPE)
    s, m = _np(sdf), _np(mip)
    mres = int(sdf_data.resolution) // 4
    cells = (mres // cell) ** 3 * int(sdf_data.cascadesCount)
    bits = np.zeros((cells + 31) // 32, dtype=np.uint32)
    out = np.zeros(9, dtype=np.uint64)
    L = lib()
    L.oracle_open_space_stats.restype = C.c_int
    L.oracle_open_space_stats.argtypes = [C.POINTER(abi.GlobalSDFData), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int]
    rc = L.oracle_open_space_stats(C.byref(sdf_data), _ptr(s), _ptr(m), len(traces), _ptr(traces), float(start_bias), _ptr(out), _ptr(bits), int(cell))
    assert rc == 0, rc
    return dict(zip(("mip_taps", "open_steps", "violations", "open_cells", "cells", "near_steps", "near_violations", "near_tex_used", "tex_used"),
                    (int(x) for x in out))), bits


def trace_global_sdf_open_skip(sdf_data, sdf, mip, traces, start_bias=0.0, cell=8):
    """tracyGlobalSDF with the experimental march's table-driven control flow -> (hits, (mip taps, full-resolution taps) actually taken)."""
    traces = np.ascontiguousarray(traces, dtype=abi.SDF_TRACE_DTYPE)
    hits = np.zeros(len(traces), dtype=abi.SDF_HIT_DTYPE)
    taps = np.zeros(2, dtype=np.uint64)
    s, m = _np(sdf), _np(mip)
    L = lib()
    L.oracle_trace_global_sdf_open_skip.restype = C.c_int
    L.oracle_trace_global_sdf_open_skip.argtypes = [C.POINTER(abi.GlobalSDFData), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int]
    rc = L.oracle_trace_global_sdf_open_skip(C.byref(sdf_data), _ptr(s), _ptr(m), len(traces), _ptr(traces), float(start_bias), _ptr(hits), _ptr(taps), int(cell))
    assert rc == 0, rc
    return hits, (int(taps[0]), int(taps[1]))


def step_classes(sdf_data, sdf, mip, traces, max_steps=96, start_bias=0.0, cell=8):
    """Per-step class of every ray, uint8 [n][max_steps]: 0 ended, 1 open (no tap), 2 near (full-resolution tap only), 3 near (both taps), 4 undecided (both)."""
    traces = np.ascontiguousarray(traces, dtype=abi.SDF_TRACE_DTYPE)
    out = np.zeros((len(traces), max_steps), dtype=np.uint8)
    s, m = _np(sdf), _np(mip)
    L = lib()
    L.oracle_step_classes.restype = C.c_int
    L.oracle_step_classes.argtypes = [C.POINTER(abi.GlobalSDFData), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_int]
    rc = L.oracle_step_classes(C.byref(sdf_data), _ptr(s), _ptr(m), len(traces), _ptr(traces), float(start_bias), int(max_steps), _ptr(out), int(cell))
    assert rc == 0, rc
    return out
