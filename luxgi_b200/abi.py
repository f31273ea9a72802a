"""ctypes mirrors of the POD structs in include/luxddgi.h.

Field-for-field twins of the reference's GLSL blocks (see the header for file:line citations):
DDGIUniform (Shaders/DDGI/DDGICommon.glsl:11-31), GlobalSDFData (Shaders/SDF/GlobalSDFData.glsl:4-12),
GlobalSurfaceAtlasData / ObjectBuffer / TileBuffer (Shaders/SDF/AtlasCommon.glsl:8-32), trace push constants
(Shaders/DDGI/GISDFRays.comp:53-60).  Sizes are asserted at import time.
"""
import ctypes as C

import numpy as np

IRRADIANCE_OCT_SIZE = 8
DEPTH_OCT_SIZE = 16
GLOBAL_SDF_WORLD_SIZE = 60000.0
CHUNKS_RESOLUTION = 40


class DDGIUniform(C.Structure):
    _fields_ = [
        ("startPosition", C.c_float * 4),
        ("step", C.c_float * 4),
        ("probeCounts", C.c_int32 * 4),
        ("maxDistance", C.c_float),
        ("sharpness", C.c_float),
        ("hysteresis", C.c_float),
        ("normalBias", C.c_float),
        ("ddgiGamma", C.c_float),
        ("irradianceProbeSideLength", C.c_int32),
        ("irradianceTextureWidth", C.c_int32),
        ("irradianceTextureHeight", C.c_int32),
        ("depthProbeSideLength", C.c_int32),
        ("depthTextureWidth", C.c_int32),
        ("depthTextureHeight", C.c_int32),
        ("raysPerProbe", C.c_int32),
    ]


class IrradianceVolume(C.Structure):
    _fields_ = [
        ("probeDistance", C.c_float),
        ("infiniteBounce", C.c_int32),
        ("raysPerProbe", C.c_int32),
        ("hysteresis", C.c_float),
        ("intensity", C.c_float),
        ("normalBias", C.c_float),
        ("depthSharpness", C.c_float),
        ("ddgiGamma", C.c_float),
    ]


class TracePushConstants(C.Structure):
    _fields_ = [
        ("randomOrientation", C.c_float * 16),
        ("numFrames", C.c_uint32),
        ("infiniteBounces", C.c_uint32),
        ("numLights", C.c_int32),
        ("intensity", C.c_float),
    ]


class GlobalSDFData(C.Structure):
    _fields_ = [
        ("cascadePosDistance", (C.c_float * 4) * 4),
        ("cascadeVoxelSize", C.c_float * 4),
        ("cascadesCount", C.c_uint32),
        ("resolution", C.c_float),
        ("nearPlane", C.c_float),
        ("farPlane", C.c_float),
    ]


class ObjectRasterizeData(C.Structure):  # ObjectRasterizeData.glsl:7-17, std430, 176 bytes
    _fields_ = [
        ("worldToVolume", C.c_float * 16),
        ("volumeToWorld", C.c_float * 16),
        ("volumeToUVWMul", C.c_float * 3),
        ("mipOffset", C.c_float),
        ("volumeToUVWAdd", C.c_float * 3),
        ("decodeMul", C.c_float),
        ("volumeLocalBoundsExtent", C.c_float * 3),
        ("decodeAdd", C.c_float),
    ]


class MeshSDF(C.Structure):  # LuxMeshSDF: component::MeshDistanceField + world matrix
    _fields_ = [
        ("mips", C.c_void_p * 3),
        ("size", C.c_uint32 * 3),
        ("mipCount", C.c_int32),
        ("aabbMin", C.c_float * 3),
        ("aabbMax", C.c_float * 3),
        ("localToUVWMul", C.c_float * 3),
        ("localToUVWAdd", C.c_float * 3),
        ("maxDistance", C.c_float),
        ("worldMatrix", C.c_float * 16),
    ]


SDF_RASTERIZE_MODEL_MAX_COUNT = 28
SDF_RASTERIZE_CHUNK_SIZE = 32


class GlobalSurfaceAtlasData(C.Structure):
    _fields_ = [
        ("cameraPos", C.c_float * 3),
        ("chunkSize", C.c_float),
        ("culledObjectsCapacity", C.c_uint32),
        ("resolution", C.c_uint32),
        ("objectsCount", C.c_uint32),
        ("padding", C.c_uint32),
    ]


class ObjectBuffer(C.Structure):
    _fields_ = [
        ("objectBounds", C.c_float * 4),
        ("tileOffset", C.c_uint32 * 6),
        ("padding", C.c_int32 * 2),
        ("transform", C.c_float * 16),
        ("extends", C.c_float * 4),
    ]


class TileBuffer(C.Structure):
    _fields_ = [
        ("extends", C.c_float * 4),
        ("transform", C.c_float * 16),
        ("objectBounds", C.c_float * 4),
    ]


class CreateInfo(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world", C.c_int32),
        ("flags", C.c_uint32),
        ("stream", C.c_void_p),
    ]


class State(C.Structure):
    _fields_ = [
        ("frames", C.c_int32),
        ("pingPong", C.c_int32),
        ("probeBegin", C.c_int32),
        ("probeCount", C.c_int32),
        ("irradianceRowBegin", C.c_int32),
        ("irradianceRowCount", C.c_int32),
        ("depthRowBegin", C.c_int32),
        ("depthRowCount", C.c_int32),
        ("kernelLaunches", C.c_uint64),
        ("layerProbes", C.c_int32),
        ("layerStride", C.c_int32),
        ("unitLayers", C.c_int32),
    ]

    def own_rows(self, side):
        """Atlas rows owned by the shard (side = 8 | 16), in shard-local order: what lux_ddgi_download_shard_async packs."""
        S, begin = (side + 2) * self.unitLayers, (self.irradianceRowBegin if side == 8 else self.depthRowBegin)  # rows per interleave unit
        units = self.probeCount // self.layerProbes
        return [begin + k * self.layerStride * S + r for k in range(units) for r in range(S)]

    def own_probes(self):
        """Probe ids of the shard in shard-local order."""
        L = self.layerProbes
        return [self.probeBegin + (l // L) * self.layerStride * L + l % L for l in range(self.probeCount)]


class StageTimes(C.Structure):
    _fields_ = [
        ("setup_ms", C.c_float),
        ("trace_ms", C.c_float),
        ("blend_ms", C.c_float),
        ("border_ms", C.c_float),
        ("total_ms", C.c_float),
        ("march_ms", C.c_float),
        ("shade_ms", C.c_float),
    ]


SIZES = {
    DDGIUniform: 96,
    TracePushConstants: 80,
    GlobalSDFData: 96,
    GlobalSurfaceAtlasData: 32,
    ObjectBuffer: 128,
    TileBuffer: 96,
}
for _t, _n in SIZES.items():
    assert C.sizeof(_t) == _n, (_t.__name__, C.sizeof(_t), _n)

# numpy structured dtypes for bulk construction of the SSBO contents
import numpy as np  # noqa: E402

OBJECT_DTYPE = np.dtype(
    [("objectBounds", "<f4", 4), ("tileOffset", "<u4", 6), ("padding", "<i4", 2), ("transform", "<f4", 16), ("extends", "<f4", 4)]
)
TILE_DTYPE = np.dtype([("extends", "<f4", 4), ("transform", "<f4", 16), ("objectBounds", "<f4", 4)])
assert OBJECT_DTYPE.itemsize == 128 and TILE_DTYPE.itemsize == 96

# LuxStatus / flags / buffer ids (include/luxddgi.h)
LUX_OK = 0
MEM_HOST, MEM_DEVICE = 0, 1
FLAG_STAGE_TIMERS = 1 << 0
FLAG_UNFUSED_BORDER = 1 << 1
FLAG_SDF_TEXTURE = 1 << 2
FLAG_SDF_LOADS = 1 << 3
FLAG_TRACE_SIMPLE = 1 << 4
FLAG_NO_PREFILTER = 1 << 5
FLAG_SHADE_UNSORTED = 1 << 6
FLAG_NO_PIPELINE = 1 << 7
FLAG_MARCH_ROWS = 1 << 9  # force row chunks in the march
FLAG_MARCH_BEAMS = 1 << 10  # force beam chunks
FLAG_BLEND_TC = 1 << 11  # opt-in tensor-core blend (tolerance path)
FLAG_BLEND_LISTS = 1 << 12  # force the list form of the FP32 blend (default from 148 x 64 probes per shard)
FLAG_BLEND_TILES = 1 << 13  # force the tiled form
FLAG_BLEND_TC_MMA_SYNC = 1 << 14  # with FLAG_BLEND_TC: mma.sync kernels instead of tcgen05 / TMA
FLAG_SHARD_INTERLEAVED = 1 << 15  # multi-GPU: rank g owns z-layers g, g + world, ... (balanced) instead of one z-slab


def flag_shard_blocks(log2_layers):
    """LUX_DDGI_FLAG_SHARD_BLOCKS: interleave blocks of 2**log2_layers z-layers (0 = single layers)."""
    return FLAG_SHARD_INTERLEAVED | (int(log2_layers) << 16)


FLAG_MARCH_PROBE_MAJOR = 1 << 8  # A/B: the round-1 march work order
BUF_RADIANCE, BUF_DIRECTION_DISTANCE, BUF_IRRADIANCE, BUF_DEPTH, BUF_IRRADIANCE_PREV, BUF_DEPTH_PREV, BUF_GLOBAL_SDF, BUF_GLOBAL_SDF_MIP = range(8)


def make_uniform(start, step, counts, rays, max_distance=None, sharpness=50.0, hysteresis=0.98, normal_bias=1.0, gamma=5.0):
    """DDGIUniform with the atlas sizing of init::initializeProbeGrid (DDGIRenderer.cpp:181-191).

    max_distance defaults to 1.5 * min(step) (DDGIRenderer.cpp:674 with a per-axis step)."""
    u = DDGIUniform()
    u.startPosition[:] = [float(start[0]), float(start[1]), float(start[2]), 1.0]
    u.step[:] = [float(step[0]), float(step[1]), float(step[2]), 0.0]
    u.probeCounts[:] = [int(counts[0]), int(counts[1]), int(counts[2]), 1]
    u.maxDistance = float(max_distance if max_distance is not None else 1.5 * min(step))
    u.sharpness = sharpness
    u.hysteresis = hysteresis
    u.normalBias = normal_bias
    u.ddgiGamma = gamma
    u.irradianceProbeSideLength = IRRADIANCE_OCT_SIZE
    u.depthProbeSideLength = DEPTH_OCT_SIZE
    xy = int(counts[0]) * int(counts[1])
    u.irradianceTextureWidth = (IRRADIANCE_OCT_SIZE + 2) * xy + 2
    u.irradianceTextureHeight = (IRRADIANCE_OCT_SIZE + 2) * int(counts[2]) + 2
    u.depthTextureWidth = (DEPTH_OCT_SIZE + 2) * xy + 2
    u.depthTextureHeight = (DEPTH_OCT_SIZE + 2) * int(counts[2]) + 2
    u.raysPerProbe = int(rays)
    return u


def probe_count(u):
    return u.probeCounts[0] * u.probeCounts[1] * u.probeCounts[2]


class Light(C.Structure):  # LuxLight = Common/Light.glsl:19-29, 64 bytes (row f4: lux_ddgi_surface_direct_light)
    _fields_ = [("color", C.c_float * 4), ("position", C.c_float * 4), ("direction", C.c_float * 4), ("intensity", C.c_float), ("radius", C.c_float),
                ("type", C.c_float), ("angle", C.c_float)]


LIGHT_DIRECTIONAL, LIGHT_SPOT, LIGHT_POINT = 0.0, 1.0, 2.0


def make_light(values16) -> Light:
    """color[4], position[4], direction[4], intensity, radius, type, angle -> Light."""
    v = [float(x) for x in values16]
    assert len(v) == 16
    l = Light()
    l.color[:], l.position[:], l.direction[:] = v[0:4], v[4:8], v[8:12]
    l.intensity, l.radius, l.type, l.angle = v[12:16]
    return l


class ReflectionPushConstants(C.Structure):  # LuxReflectionPushConstants = SDFReflection.comp:66-78, 112 bytes
    _fields_ = [("bias", C.c_float), ("trim", C.c_float), ("intensity", C.c_float), ("roughDDGIIntensity", C.c_float), ("numLights", C.c_uint32),
                ("numFrames", C.c_uint32), ("sampleGI", C.c_uint32), ("approximateWithDDGI", C.c_uint32), ("cameraPosition", C.c_float * 4),
                ("viewProjInv", C.c_float * 16)]


def make_reflection_push(camera_position, view_proj_inv, num_frames=0, trim=1.0, rough_ddgi_intensity=1.0, approximate_with_ddgi=1) -> ReflectionPushConstants:
    p = ReflectionPushConstants()
    p.bias, p.trim, p.intensity, p.roughDDGIIntensity = 0.0, float(trim), 1.0, float(rough_ddgi_intensity)
    p.numLights, p.numFrames, p.sampleGI, p.approximateWithDDGI = 0, int(num_frames), 0, int(approximate_with_ddgi)
    p.cameraPosition[:] = [float(x) for x in list(camera_position)[:3]] + [1.0]
    p.viewProjInv[:] = [float(x) for x in np.asarray(view_proj_inv, dtype=np.float32).reshape(16)]
    return p


assert C.sizeof(ReflectionPushConstants) == 112 and C.sizeof(Light) == 64


# GlobalSDFTrace / GlobalSDFHit (row f4: lux_ddgi_trace_global_sdf), as numpy record dtypes
SDF_TRACE_DTYPE = np.dtype([("worldPosition", "<f4", 3), ("minDistance", "<f4"), ("worldDirection", "<f4", 3), ("maxDistance", "<f4"),
                            ("stepScale", "<f4"), ("needsHitNormal", "<u4")])
SDF_HIT_DTYPE = np.dtype([("hitNormal", "<f4", 3), ("hitTime", "<f4"), ("hitCascade", "<u4"), ("stepsCount", "<u4"), ("hitSDF", "<f4")])
assert SDF_TRACE_DTYPE.itemsize == 40 and SDF_HIT_DTYPE.itemsize == 28
