"""Python host binding of libluxddgi.so (ctypes over the C ABI in include/luxddgi.h).

The class mirrors the reference's pass interface (Code/Maple/src/Engine/DDGI/DDGIRenderer.cpp): a `DDGIPipeline` owns
what `DDGIPipelineInternal` owns (ray buffers, 2x2 ping-pong atlases, frames, pingPong) and exposes the systems under
their reference names — trace_rays, probe_update, border_update, end_frame — plus `update` (= the whole pass).

There is NO CPU path here: if the library or a B200 is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

_LIB = None
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libluxddgi.so")

EXPORTS = [
    "lux_ddgi_version", "lux_ddgi_last_error", "lux_ddgi_uniform_from_volume", "lux_ddgi_uniform_finalize", "lux_ddgi_create",
    "lux_ddgi_destroy", "lux_ddgi_set_uniform", "lux_ddgi_set_global_sdf", "lux_ddgi_set_surface_atlas",
    "lux_ddgi_update_surface_light_cache", "lux_ddgi_set_skybox", "lux_ddgi_trace_rays", "lux_ddgi_probe_update",
    "lux_ddgi_border_update", "lux_ddgi_end_frame", "lux_ddgi_update", "lux_ddgi_synchronize", "lux_ddgi_get_buffer",
    "lux_ddgi_download", "lux_ddgi_download_async", "lux_ddgi_download_rows_async", "lux_ddgi_set_ray_buffers", "lux_ddgi_restore", "lux_ddgi_get_state", "lux_ddgi_shard_layout", "lux_ddgi_shard_layout_ex", "lux_ddgi_download_shard_async",
    "lux_ddgi_get_stage_ms", "lux_ddgi_sample_irradiance", "lux_ddgi_sample_probe", "lux_ddgi_indirect_light",
    "lux_ddgi_get_surface_light_cache", "lux_ddgi_build_global_sdf", "lux_ddgi_build_sdf_mip", "lux_ddgi_sdf_file_read", "lux_ddgi_download_fence", "lux_ddgi_wait_fence", "lux_ddgi_set_nccl_comm", "lux_ddgi_update_surface_light_cache_rows", "lux_ddgi_cull_surface_objects",
    "lux_ddgi_get_surface_cull_lists", "lux_ddgi_trace_global_sdf", "lux_ddgi_surface_direct_light",
    "lux_ddgi_sdf_reflection", "lux_ddgi_sdf_shadow", "lux_ddgi_measure_l2_read_bandwidth", "lux_ddgi_update_global_sdf_region",
]


class LuxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libluxddgi error {code}: {msg}")
        self.code = code


def load():
    """dlopen libluxddgi.so.  Raises if it has not been built (python -m luxgi_b200.build)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("LUX_DDGI_LIB", _LIB_PATH)  # tuning hook: an alternative build of the same library
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -m luxgi_b200.build` (there is no CPU fallback)")
    L = C.CDLL(path)
    vp, i32, u32, sz = C.c_void_p, C.c_int32, C.c_uint32, C.c_size_t
    L.lux_ddgi_version.restype = u32
    L.lux_ddgi_last_error.restype = C.c_char_p
    sig = {
        "lux_ddgi_uniform_from_volume": [C.POINTER(abi.IrradianceVolume), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(abi.DDGIUniform)],
        "lux_ddgi_uniform_finalize": [C.POINTER(abi.DDGIUniform)],
        "lux_ddgi_create": [C.POINTER(abi.DDGIUniform), C.POINTER(abi.CreateInfo), C.POINTER(vp)],
        "lux_ddgi_destroy": [vp],
        "lux_ddgi_set_uniform": [vp, C.POINTER(abi.DDGIUniform)],
        "lux_ddgi_set_global_sdf": [vp, C.POINTER(abi.GlobalSDFData), vp, vp, i32],
        "lux_ddgi_set_surface_atlas": [vp, C.POINTER(abi.GlobalSurfaceAtlasData), vp, vp, sz, vp, sz, vp, sz, vp, vp, i32],
        "lux_ddgi_update_surface_light_cache": [vp, vp, i32],
        "lux_ddgi_set_skybox": [vp, i32, vp, i32],
        "lux_ddgi_trace_rays": [vp, C.POINTER(abi.TracePushConstants)],
        "lux_ddgi_probe_update": [vp],
        "lux_ddgi_border_update": [vp],
        "lux_ddgi_end_frame": [vp],
        "lux_ddgi_update": [vp, C.POINTER(C.c_float)],
        "lux_ddgi_synchronize": [vp],
        "lux_ddgi_get_buffer": [vp, i32, C.POINTER(vp), C.POINTER(sz)],
        "lux_ddgi_download": [vp, i32, vp, sz],
        "lux_ddgi_download_async": [vp, i32, vp, sz],
        "lux_ddgi_download_rows_async": [vp, i32, i32, i32, vp],
        "lux_ddgi_set_ray_buffers": [vp, vp, vp, i32],
        "lux_ddgi_restore": [vp, vp, vp, i32, i32],
        "lux_ddgi_get_state": [vp, C.POINTER(abi.State)],
        "lux_ddgi_shard_layout": [C.POINTER(abi.DDGIUniform), i32, i32, C.POINTER(abi.State)],
        "lux_ddgi_shard_layout_ex": [C.POINTER(abi.DDGIUniform), i32, i32, C.c_uint32, C.POINTER(abi.State)],
        "lux_ddgi_download_shard_async": [vp, i32, vp],
        "lux_ddgi_get_stage_ms": [vp, C.POINTER(abi.StageTimes)],
        "lux_ddgi_sample_irradiance": [vp, i32, vp, vp, vp, vp, i32],
        "lux_ddgi_sample_probe": [vp, i32, i32, vp, vp, vp, vp, vp, i32],
        "lux_ddgi_indirect_light": [vp, vp, i32, vp, vp, vp, vp, vp, C.c_float, vp, i32],
        "lux_ddgi_get_surface_light_cache": [vp, C.POINTER(vp), C.POINTER(sz)],
        "lux_ddgi_surface_direct_light": [vp, C.POINTER(abi.Light), vp, i32, vp, vp, vp, vp, vp, i32],
        "lux_ddgi_measure_l2_read_bandwidth": [vp, sz, i32, C.POINTER(C.c_float)],
        "lux_ddgi_sdf_reflection": [vp, C.POINTER(abi.ReflectionPushConstants), i32, i32, vp, vp, vp, vp, vp, vp, i32],
        "lux_ddgi_sdf_shadow": [vp, C.POINTER(abi.Light), vp, C.c_uint32, C.c_float, i32, i32, vp, vp, vp, vp, vp, i32],
        "lux_ddgi_build_global_sdf": [vp, C.POINTER(abi.GlobalSDFData), C.POINTER(abi.MeshSDF), i32, C.c_float],
        "lux_ddgi_build_sdf_mip": [vp],
        "lux_ddgi_update_global_sdf_region": [vp, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), vp, C.c_int, C.c_int32],
        "lux_ddgi_download_fence": [vp, C.POINTER(C.c_uint64)],
        "lux_ddgi_wait_fence": [vp, C.c_uint64],
        "lux_ddgi_set_nccl_comm": [vp, vp],
        "lux_ddgi_cull_surface_objects": [vp, C.c_uint32],
        "lux_ddgi_trace_global_sdf": [vp, i32, vp, C.c_float, vp, i32],
        "lux_ddgi_get_surface_cull_lists": [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(sz)],
        "lux_ddgi_update_surface_light_cache_rows": [vp, vp, i32, i32, i32],
        "lux_ddgi_sdf_file_read": [C.c_char_p, C.POINTER(C.c_uint32 * 3), C.POINTER(i32), C.POINTER(C.c_uint64), vp],
    }
    for name, argtypes in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _LIB = L
    return L


def _check(rc):
    if rc != abi.LUX_OK:
        raise LuxError(rc, load().lux_ddgi_last_error().decode("utf-8", "replace"))


def _host_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _as_host(t, dtype=None):
    """numpy array or CPU torch tensor -> contiguous numpy array (fp16 kept as is)."""
    if hasattr(t, "detach"):
        t = t.detach().contiguous().numpy()
    a = np.ascontiguousarray(t)
    if dtype is not None and a.dtype != dtype:
        a = a.astype(dtype)
    return a


def _is_cuda_tensor(t):
    return hasattr(t, "is_cuda") and t.is_cuda


def _mem(t):
    """-> (pointer, kind, keepalive)"""
    if _is_cuda_tensor(t):
        t = t.contiguous()
        return C.c_void_p(t.data_ptr()), abi.MEM_DEVICE, t
    a = _as_host(t)
    return _host_ptr(a), abi.MEM_HOST, a


def uniform_from_volume(volume: abi.IrradianceVolume, aabb_min, aabb_max) -> abi.DDGIUniform:
    """ddgi::on_game_start grid derivation (DDGIRenderer.cpp:663-682)."""
    u = abi.DDGIUniform()
    mn = (C.c_float * 3)(*[float(v) for v in aabb_min])
    mx = (C.c_float * 3)(*[float(v) for v in aabb_max])
    _check(load().lux_ddgi_uniform_from_volume(C.byref(volume), mn, mx, C.byref(u)))
    return u


def shard_layout(uniform: abi.DDGIUniform, rank: int, world: int, flags: int = 0) -> abi.State:
    """Probe range and atlas row ranges of one shard - a z-slab, or interleaved z-layers with abi.FLAG_SHARD_INTERLEAVED (host arithmetic only;
    works without a GPU)."""
    st = abi.State()
    _check(load().lux_ddgi_shard_layout_ex(C.byref(uniform), int(rank), int(world), int(flags), C.byref(st)))
    return st


class DeviceView:
    """Zero-copy view of an engine buffer for torch (`torch.as_tensor(view, device='cuda')`) via __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}
        self._owner = owner


class DDGIPipeline:
    """One probe volume (or one z-slab shard of it) on one GPU."""

    def __init__(self, uniform: abi.DDGIUniform, device=0, rank=0, world=1, flags=0, stream=None):
        self._lib = load()
        self.uniform = uniform
        self._keep = []
        info = abi.CreateInfo(int(device), int(rank), int(world), int(flags), C.c_void_p(stream) if stream else None)
        h = C.c_void_p()
        _check(self._lib.lux_ddgi_create(C.byref(uniform), C.byref(info), C.byref(h)))
        self._h = h
        self.rays = uniform.raysPerProbe
        st = self.state()
        self.probe_begin, self.probe_count = st.probeBegin, st.probeCount

    # ---- lifetime -------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.lux_ddgi_destroy(self._h)
            self._h = None
            self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- inputs (what trace_rays::system binds, DDGIRenderer.cpp:304-316) -------------------------------------
    def set_global_sdf(self, sdf_data: abi.GlobalSDFData, sdf, mip):
        ps, ks, k1 = _mem(sdf)
        pm, km, k2 = _mem(mip)
        if ks != km:
            raise ValueError("sdf and mip must both be host or both be device arrays")
        _check(self._lib.lux_ddgi_set_global_sdf(self._h, C.byref(sdf_data), ps, pm, ks))
        self._sdf_data = sdf_data
        if ks == abi.MEM_DEVICE:
            self._keep += [k1, k2]

    def set_surface_atlas(self, atlas_data, chunks, cull, objects, tiles, light, depth):
        if atlas_data is None:
            _check(self._lib.lux_ddgi_set_surface_atlas(self._h, None, None, None, 0, None, 0, None, 0, None, None, 0))
            return
        dev = _is_cuda_tensor(light)
        if dev:
            import torch

            to_dev = lambda a: torch.as_tensor(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(light.device)
            tens = [to_dev(chunks), to_dev(cull), to_dev(objects), to_dev(tiles), light.contiguous(), depth.contiguous()]
            ptrs = [C.c_void_p(t.data_ptr()) for t in tens]
            kind = abi.MEM_DEVICE
            self._keep += tens
        else:
            arrs = [_as_host(chunks, np.uint32), _as_host(cull, np.uint32), _as_host(objects), _as_host(tiles), _as_host(light),
                    _as_host(depth, np.float32)]
            ptrs = [_host_ptr(a) for a in arrs]
            kind = abi.MEM_HOST
        _check(self._lib.lux_ddgi_set_surface_atlas(self._h, C.byref(atlas_data), ptrs[0], ptrs[1], len(cull), ptrs[2], len(objects),
                                                     ptrs[3], len(tiles), ptrs[4], ptrs[5], kind))

    def update_surface_light_cache(self, light):
        p, k, keep = _mem(light)
        _check(self._lib.lux_ddgi_update_surface_light_cache(self._h, p, k))
        if k == abi.MEM_DEVICE:
            self._keep.append(keep)

    def update_surface_light_cache_ptr(self, host_ptr):
        _check(self._lib.lux_ddgi_update_surface_light_cache(self._h, C.c_void_p(host_ptr), abi.MEM_HOST))

    def update_surface_light_cache_rows_ptr(self, host_ptr, row_begin, row_count):
        _check(self._lib.lux_ddgi_update_surface_light_cache_rows(self._h, C.c_void_p(host_ptr), int(row_begin), int(row_count), abi.MEM_HOST))

    def set_skybox(self, face_size, faces):
        if not face_size or faces is None:
            _check(self._lib.lux_ddgi_set_skybox(self._h, 0, None, 0))
            return
        p, k, keep = _mem(faces)
        _check(self._lib.lux_ddgi_set_skybox(self._h, int(face_size), p, k))
        if k == abi.MEM_DEVICE:
            self._keep.append(keep)

    def set_scene(self, scene):
        """Bind every input of a luxgi_b200.scenes.Scene."""
        self.set_global_sdf(scene.sdf_data, scene.sdf, scene.mip)
        if scene.atlas_data is not None:
            self.set_surface_atlas(scene.atlas_data, scene.chunks, scene.cull, scene.objects, scene.tiles, scene.light, scene.depth)
        self.set_skybox(scene.sky_face, scene.sky)

    def set_uniform(self, uniform):
        _check(self._lib.lux_ddgi_set_uniform(self._h, C.byref(uniform)))
        self.uniform = uniform

    def set_ray_buffers(self, radiance, direction_distance):
        pr, kr, _ = _mem(radiance)
        pd, kd, _ = _mem(direction_distance)
        assert kr == kd
        _check(self._lib.lux_ddgi_set_ray_buffers(self._h, pr, pd, kr))

    # ---- the systems, in RenderGraph order (RenderGraph.cpp:98-114) ----------------------------------------------
    def trace_rays(self, orientation, num_frames=0):
        pc = abi.TracePushConstants()
        pc.randomOrientation[:] = [float(v) for v in np.asarray(orientation, dtype=np.float32).reshape(16)]
        pc.numFrames = num_frames
        pc.infiniteBounces = 1 if num_frames else 0
        pc.intensity = 1.0
        _check(self._lib.lux_ddgi_trace_rays(self._h, C.byref(pc)))

    def probe_update(self):
        _check(self._lib.lux_ddgi_probe_update(self._h))

    def border_update(self):
        _check(self._lib.lux_ddgi_border_update(self._h))

    def end_frame(self):
        _check(self._lib.lux_ddgi_end_frame(self._h))

    def update(self, orientation):
        rot = np.ascontiguousarray(orientation, dtype=np.float32).reshape(16)
        _check(self._lib.lux_ddgi_update(self._h, rot.ctypes.data_as(C.POINTER(C.c_float))))

    def synchronize(self):
        _check(self._lib.lux_ddgi_synchronize(self._h))

    # ---- outputs -------------------------------------------------------------------------------------------------
    def state(self) -> abi.State:
        st = abi.State()
        _check(self._lib.lux_ddgi_get_state(self._h, C.byref(st)))
        return st

    def stage_ms(self) -> abi.StageTimes:
        t = abi.StageTimes()
        _check(self._lib.lux_ddgi_get_stage_ms(self._h, C.byref(t)))
        return t

    def _shape(self, buf):
        u = self.uniform
        if buf in (abi.BUF_GLOBAL_SDF, abi.BUF_GLOBAL_SDF_MIP):
            res, casc = int(self._sdf_data.resolution), int(self._sdf_data.cascadesCount)
            r = res if buf == abi.BUF_GLOBAL_SDF else res // 4
            return (r, r, r * casc)
        if buf in (abi.BUF_RADIANCE, abi.BUF_DIRECTION_DISTANCE):
            return (self.probe_count, self.rays, 4)
        if buf in (abi.BUF_IRRADIANCE, abi.BUF_IRRADIANCE_PREV):
            return (u.irradianceTextureHeight, u.irradianceTextureWidth, 4)
        return (u.depthTextureHeight, u.depthTextureWidth, 2)

    def buffer_ptr(self, buf):
        p, n = C.c_void_p(), C.c_size_t()
        _check(self._lib.lux_ddgi_get_buffer(self._h, buf, C.byref(p), C.byref(n)))
        return p.value, n.value

    def device_view(self, buf) -> DeviceView:
        p, _ = self.buffer_ptr(buf)
        return DeviceView(p, self._shape(buf), "<f2", self)

    def download(self, buf) -> np.ndarray:
        """uint16 (fp16 bit patterns) array of the buffer's natural shape."""
        out = np.empty(self._shape(buf), dtype=np.uint16)
        _check(self._lib.lux_ddgi_download(self._h, buf, _host_ptr(out), out.nbytes))
        return out

    def download_async_ptr(self, buf, host_ptr, nbytes):
        _check(self._lib.lux_ddgi_download_async(self._h, buf, C.c_void_p(host_ptr), nbytes))

    def download_rows_async_ptr(self, buf, row_begin, row_count, host_ptr):
        _check(self._lib.lux_ddgi_download_rows_async(self._h, buf, int(row_begin), int(row_count), C.c_void_p(host_ptr)))

    def download_shard_async_ptr(self, buf, host_ptr):
        """The shard's own rows of an atlas, packed in shard-local layer order, to pinned host memory (see State.own_rows)."""
        _check(self._lib.lux_ddgi_download_shard_async(self._h, buf, C.c_void_p(host_ptr)))

    def trace_global_sdf(self, traces, start_bias=0.0) -> np.ndarray:
        """tracyGlobalSDF for arbitrary rays: traces = abi.SDF_TRACE_DTYPE records -> abi.SDF_HIT_DTYPE records."""
        traces = np.ascontiguousarray(traces, dtype=abi.SDF_TRACE_DTYPE)
        hits = np.zeros(len(traces), dtype=abi.SDF_HIT_DTYPE)
        _check(self._lib.lux_ddgi_trace_global_sdf(self._h, len(traces), _host_ptr(traces), float(start_bias), _host_ptr(hits), abi.MEM_HOST))
        return hits

    def cull_surface_objects(self, capacity_words=0):
        """SDFCulling.comp on device over the bound object buffer -> (chunks uint32[64000], cull uint32[words])."""
        import torch

        _check(self._lib.lux_ddgi_cull_surface_objects(self._h, int(capacity_words)))
        pc, pl, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        _check(self._lib.lux_ddgi_get_surface_cull_lists(self._h, C.byref(pc), C.byref(pl), C.byref(n)))
        chunks = torch.as_tensor(DeviceView(pc.value, (abi.CHUNKS_RESOLUTION ** 3,), "<i4", self), device="cuda").cpu().numpy().view(np.uint32)
        cull = torch.as_tensor(DeviceView(pl.value, (n.value,), "<i4", self), device="cuda").cpu().numpy().view(np.uint32)
        return chunks, cull

    def set_nccl_comm(self, comm_ptr):
        """Bind an ncclComm_t (integer address, e.g. luxgi_b200.nccl.NcclComm.ptr): lux_ddgi_update then all-gathers the atlas rows itself."""
        _check(self._lib.lux_ddgi_set_nccl_comm(self._h, C.c_void_p(comm_ptr)))

    def download_fence(self) -> int:
        f = C.c_uint64()
        _check(self._lib.lux_ddgi_download_fence(self._h, C.byref(f)))
        return f.value

    def wait_fence(self, fence: int):
        _check(self._lib.lux_ddgi_wait_fence(self._h, fence))

    def restore(self, irradiance, depth, frames, ping_pong):
        a, b = _as_host(irradiance), _as_host(depth)
        _check(self._lib.lux_ddgi_restore(self._h, _host_ptr(a), _host_ptr(b), int(frames), int(ping_pong)))

    # ---- consumer side (SampleProbe.comp / sampleIrradiance) --------------------------------------------------------
    def sample_irradiance(self, P, N, Wo) -> np.ndarray:
        P, N, Wo = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in (P, N, Wo))
        out = np.empty_like(P)
        _check(self._lib.lux_ddgi_sample_irradiance(self._h, len(P), _host_ptr(P), _host_ptr(N), _host_ptr(Wo), _host_ptr(out), abi.MEM_HOST))
        return out

    def sample_probe(self, depth, normals, camera_position, view_proj_inv) -> np.ndarray:
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        normals = np.ascontiguousarray(normals, dtype=np.float32)
        h, w = depth.shape
        cam = np.ascontiguousarray(camera_position, dtype=np.float32).reshape(4)
        vpi = np.ascontiguousarray(view_proj_inv, dtype=np.float32).reshape(16)
        out = np.empty((h, w, 4), dtype=np.float32)
        _check(self._lib.lux_ddgi_sample_probe(self._h, w, h, _host_ptr(depth), _host_ptr(normals), _host_ptr(cam), _host_ptr(vpi), _host_ptr(out),
                                               abi.MEM_HOST))
        return out

    # ---- global SDF build ---------------------------------------------------------------------
    def build_global_sdf(self, sdf_data, meshes, min_object_radius=0.0):
        """Merge mesh distance fields (luxgi_b200.meshsdf.MeshSDF) into the cascades of `sdf_data` on device, build the mip, bind both."""
        from . import meshsdf

        arr, keep = meshsdf.to_ctypes(meshes)
        _check(self._lib.lux_ddgi_build_global_sdf(self._h, C.byref(sdf_data), arr, len(meshes), float(min_object_radius)))
        self._sdf_data = sdf_data

    def build_sdf_mip(self):
        _check(self._lib.lux_ddgi_build_sdf_mip(self._h))

    def update_global_sdf_region(self, cascade, chunk_min, chunk_max, texels, rebuild_mip=True):
        """texels: numpy uint16 / float16 box [dz][dy][dx] of the (clipped) chunk range, or None when the bound device volume was updated in place."""
        lo = (C.c_int32 * 3)(*[int(v) for v in chunk_min])
        hi = (C.c_int32 * 3)(*[int(v) for v in chunk_max])
        ptr = None
        if texels is not None:
            texels = np.ascontiguousarray(texels)
            ptr = texels.ctypes.data_as(C.c_void_p)
        _check(self._lib.lux_ddgi_update_global_sdf_region(self._h, int(cascade), lo, hi, ptr, abi.MEM_HOST, 1 if rebuild_mip else 0))

    @property
    def global_sdf(self):
        return self.download(abi.BUF_GLOBAL_SDF)

    @property
    def global_sdf_mip(self):
        return self.download(abi.BUF_GLOBAL_SDF_MIP)

    def indirect_light(self, base_light, texel, P, N, albedo, metallic, intensity, camera_pos):
        """Infinite-bounce refresh of the surface light cache (SDFAtlasIndirectLight.frag) for a list of atlas texels."""
        texel = np.ascontiguousarray(texel, dtype=np.uint32)
        P, N, albedo = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in (P, N, albedo))
        metallic = np.ascontiguousarray(metallic, dtype=np.float32)
        cam = np.ascontiguousarray(camera_pos, dtype=np.float32).reshape(3)
        base = None if base_light is None else _as_host(base_light)
        _check(self._lib.lux_ddgi_indirect_light(self._h, None if base is None else _host_ptr(base), len(texel), _host_ptr(texel), _host_ptr(P),
                                                 _host_ptr(N), _host_ptr(albedo), _host_ptr(metallic), float(intensity), _host_ptr(cam), abi.MEM_HOST))

    def surface_direct_light(self, light, camera_pos_bias, texel, P, N, albedo, metallic_roughness):
        """Direct lighting of surface-cache texels by one light (SDFDeferredLight.frag), added into the light cache; camera_pos_bias = (xyz, shadow bias)."""
        texel = np.ascontiguousarray(texel, dtype=np.uint32)
        P, N, albedo = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in (P, N, albedo))
        mr = np.ascontiguousarray(metallic_roughness, dtype=np.float32).reshape(-1, 2)
        cam = np.ascontiguousarray(camera_pos_bias, dtype=np.float32).reshape(4)
        assert len(P) == len(N) == len(albedo) == len(mr) == len(texel)
        _check(self._lib.lux_ddgi_surface_direct_light(self._h, C.byref(light), _host_ptr(cam), len(texel), _host_ptr(texel), _host_ptr(P), _host_ptr(N),
                                                       _host_ptr(albedo), _host_ptr(mr), abi.MEM_HOST))

    def sdf_reflection(self, push, depth, normals, pbr, sobol, scrambling, out=None) -> np.ndarray:
        """SDFReflection.comp over a G-buffer (depth [h][w], normals / pbr [h][w][4] float32, RGBA8 blue-noise texels) -> RGBA16F bits [h][w][4];
        `out` carries the previous image (pixels with depth == 1 keep it)."""
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        h, w = depth.shape
        normals = np.ascontiguousarray(normals, dtype=np.float32).reshape(h, w, 4)
        pbr = np.ascontiguousarray(pbr, dtype=np.float32).reshape(h, w, 4)
        sobol = np.ascontiguousarray(sobol, dtype=np.uint8).reshape(256, 4)
        scrambling = np.ascontiguousarray(scrambling, dtype=np.uint8).reshape(128, 128, 4)
        out = np.zeros((h, w, 4), dtype=np.uint16) if out is None else np.ascontiguousarray(out, dtype=np.uint16).reshape(h, w, 4)
        _check(self._lib.lux_ddgi_sdf_reflection(self._h, C.byref(push), w, h, _host_ptr(depth), _host_ptr(normals), _host_ptr(pbr), _host_ptr(sobol),
                                                 _host_ptr(scrambling), _host_ptr(out), abi.MEM_HOST))
        return out

    def sdf_shadow(self, light, view_proj_inv, num_frames, shadow_bias, depth, normals, sobol, scrambling, out=None) -> np.ndarray:
        """SDFShadow.comp over a G-buffer -> uint32 visibility words [h/4][w/8] (bit (y%4)*8 + x%8 = visible); `out` carries the previous words."""
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        h, w = depth.shape
        normals = np.ascontiguousarray(normals, dtype=np.float32).reshape(h, w, 4)
        vpi = np.ascontiguousarray(view_proj_inv, dtype=np.float32).reshape(16)
        sobol = np.ascontiguousarray(sobol, dtype=np.uint8).reshape(256, 4)
        scrambling = np.ascontiguousarray(scrambling, dtype=np.uint8).reshape(128, 128, 4)
        out = np.zeros((h // 4, w // 8), dtype=np.uint32) if out is None else np.ascontiguousarray(out, dtype=np.uint32)
        _check(self._lib.lux_ddgi_sdf_shadow(self._h, C.byref(light), _host_ptr(vpi), int(num_frames), float(shadow_bias), w, h, _host_ptr(depth),
                                             _host_ptr(normals), _host_ptr(sobol), _host_ptr(scrambling), _host_ptr(out), abi.MEM_HOST))
        return out

    def measure_l2_read_bandwidth(self, nbytes=0, repeats=5) -> float:
        """GB/s of read-only sweeps of an L2-resident buffer (default 64 MiB): the denominator of the request-level roofline."""
        out = C.c_float()
        _check(self._lib.lux_ddgi_measure_l2_read_bandwidth(self._h, int(nbytes), int(repeats), C.byref(out)))
        return float(out.value)

    def surface_light_cache(self) -> np.ndarray:
        p, n = C.c_void_p(), C.c_size_t()
        _check(self._lib.lux_ddgi_get_surface_light_cache(self._h, C.byref(p), C.byref(n)))
        import torch

        res = int(round((n.value // 8) ** 0.5))
        view = DeviceView(p.value, (res, res, 4), "<f2", self)
        self.synchronize()
        return torch.as_tensor(view, device="cuda").cpu().numpy().view(np.uint16)

    @property
    def radiance(self):
        return self.download(abi.BUF_RADIANCE)

    @property
    def direction_distance(self):
        return self.download(abi.BUF_DIRECTION_DISTANCE)

    @property
    def irradiance(self):
        return self.download(abi.BUF_IRRADIANCE)

    @property
    def depth(self):
        return self.download(abi.BUF_DEPTH)
