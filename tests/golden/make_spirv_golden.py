"""Golden vectors from the reference's SHIPPED SPIR-V binaries (run in the build container, where /root/reference exists):

    python tests/golden/make_spirv_golden.py

Executes Assets/shaders/spv/DDGI/{GISDFRays,IrradianceProbeUpdate,DepthProbeUpdate,IrradianceBorderUpdate,
DepthBorderUpdate}.comp.spv with oracle/spirv/interp.py on a small Cornell scene (SDF 32^3, 2x2x1 probes, 32 rays,
surface cache 256^2, a cube sky with six distinct faces, 2 frames so that the hysteresis branch runs) and writes
tests/golden/spirv_golden.npz: the INPUT scene parameters needed to rebuild the scene (it is procedural, seedless) and the
OUTPUTS of every stage.  tests/test_spirv_golden.py replays the same inputs through the C++ oracle and compares.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import abi, scenes  # noqa: E402
from oracle.spirv import interp as si  # noqa: E402

SPV = "/root/reference/Assets/shaders/spv/DDGI"
F = np.float32


def golden_scene():
    sc = scenes.cornell_scene(res=32, counts=(2, 2, 1), rays=32, atlas_res=256, hysteresis=0.9, gamma=2.2)
    sky = np.zeros((6, 1, 1, 4), dtype=np.float16)
    for f in range(6):
        sky[f, 0, 0] = [0.25 * (f + 1), 0.1 * (6 - f), 0.5 + 0.05 * f, 1.0]
    sc.sky_face, sc.sky = 1, sky
    return sc


def vec(a):
    return [F(x) for x in a]


def ddgi_block(u):
    return [[vec(u.startPosition), vec(u.step), [int(x) & si.M32 for x in u.probeCounts], F(u.maxDistance), F(u.sharpness), F(u.hysteresis),
             F(u.normalBias), F(u.ddgiGamma), u.irradianceProbeSideLength, u.irradianceTextureWidth, u.irradianceTextureHeight,
             u.depthProbeSideLength, u.depthTextureWidth, u.depthTextureHeight, u.raysPerProbe]]


def mat_cols(m16):
    return [vec(m16[c * 4:c * 4 + 4]) for c in range(4)]


def bind_trace(mod, sc, rot, rad_bits, dd_bits):
    a, d = sc.atlas_data, sc.sdf_data
    atlas = [vec(a.cameraPos), F(a.chunkSize), a.culledObjectsCapacity, a.resolution, a.objectsCount, a.padding]
    sdf = [[vec(d.cascadePosDistance[i]) for i in range(4)], vec(d.cascadeVoxelSize), d.cascadesCount, F(d.resolution), F(d.nearPlane), F(d.farPlane)]
    tiles = [[[vec(t["extends"]), mat_cols(t["transform"]), vec(t["objectBounds"])] for t in sc.tiles]]
    objs = [[[vec(o["objectBounds"]), [int(x) for x in o["tileOffset"]], [0, 0], mat_cols(o["transform"]), vec(o["extends"])] for o in sc.objects]]
    res = {
        0: si.StorageImage(np.zeros((1, 1, 4), dtype=np.uint16)),
        1: si.Texture3D(sc.sdf.numpy()), 2: si.Texture3D(sc.mip.numpy()),
        3: si.Texture2D(sc.light.numpy(), repeat=True), 4: si.Texture2D(sc.depth.numpy(), repeat=False),
        5: tiles, 6: objs, 7: [[int(x) for x in sc.chunks]], 8: [[int(x) for x in sc.cull]],
        9: si.TextureCube(sc.sky), 10: [atlas, sdf], 11: ddgi_block(sc.uniform),
        12: si.StorageImage(rad_bits), 13: si.StorageImage(dd_bits),
    }
    for b, v in res.items():
        gid = mod.global_by_binding(0, b)
        if gid is not None:
            mod.storage[gid] = [v]
    (pc,) = mod.global_by_storage(9)  # PushConstant
    mod.storage[pc] = [[mat_cols(rot), 0, 0, 0, F(1.0)]]
    return res


def run_trace(sc, rot):
    mod = si.Module(os.path.join(SPV, "GISDFRays.comp.spv"))
    P, R = sc.probes, sc.rays
    rad = np.zeros((P, R, 4), dtype=np.uint16)
    dd = np.zeros((P, R, 4), dtype=np.uint16)
    res = bind_trace(mod, sc, rot, rad, dd)
    n = si.dispatch(mod, [(gx, gy, 0) for gy in range(P) for gx in range((R + 15) // 16)])
    return rad, dd, n, res[1].taps, res[2].taps


def run_blend(sc, which, rad, dd, prev_irr, prev_dep, out_irr, out_dep, first_frame):
    mod = si.Module(os.path.join(SPV, f"{which}ProbeUpdate.comp.spv"))
    u = sc.uniform
    bind = {(0, 0): si.StorageImage(out_irr), (0, 1): si.StorageImage(out_dep), (1, 0): si.StorageImage(prev_irr), (1, 1): si.StorageImage(prev_dep),
            (1, 2): ddgi_block(u), (2, 0): si.StorageImage(rad), (2, 1): si.StorageImage(dd)}
    for (s, b), v in bind.items():
        gid = mod.global_by_binding(s, b)
        if gid is not None:
            mod.storage[gid] = [v]
    (pc,) = mod.global_by_storage(9)
    mod.storage[pc] = [[1 if first_frame else 0]]
    xy = u.probeCounts[0] * u.probeCounts[1]
    return si.dispatch(mod, [(gx, gy, 0) for gy in range(u.probeCounts[2]) for gx in range(xy)])


def run_border(sc, which, irr, dep):
    mod = si.Module(os.path.join(SPV, f"{which}BorderUpdate.comp.spv"))
    u = sc.uniform
    for b, v in ((0, si.StorageImage(irr)), (1, si.StorageImage(dep))):
        gid = mod.global_by_binding(0, b)
        if gid is not None:
            mod.storage[gid] = [v]
    xy = u.probeCounts[0] * u.probeCounts[1]
    return si.dispatch(mod, [(gx, gy, 0) for gy in range(u.probeCounts[2]) for gx in range(xy)])


def main():
    sc = golden_scene()
    u = sc.uniform
    out = {}
    irr = [np.zeros((u.irradianceTextureHeight, u.irradianceTextureWidth, 4), dtype=np.uint16) for _ in range(2)]
    dep = [np.zeros((u.depthTextureHeight, u.depthTextureWidth, 2), dtype=np.uint16) for _ in range(2)]
    ping = 0
    for frame in range(2):
        rot = scenes.frame_rotation(frame)
        t0 = time.time()
        rad, dd, n, taps, mtaps = run_trace(sc, rot)
        print(f"frame {frame}: trace {n} SPIR-V instructions, {taps} tex taps, {mtaps} mip taps, {time.time() - t0:.1f} s", flush=True)
        w = 1 - ping
        t0 = time.time()
        n1 = run_blend(sc, "Irradiance", rad, dd, irr[ping], dep[ping], irr[w], dep[w], frame == 0)
        n2 = run_blend(sc, "Depth", rad, dd, irr[ping], dep[ping], irr[w], dep[w], frame == 0)
        print(f"frame {frame}: blend {n1} + {n2} instructions, {time.time() - t0:.1f} s", flush=True)
        out[f"f{frame}_irradiance_interior"] = irr[w].copy()
        out[f"f{frame}_depth_interior"] = dep[w].copy()
        run_border(sc, "Irradiance", irr[w], dep[w])
        run_border(sc, "Depth", irr[w], dep[w])
        out[f"f{frame}_radiance"] = rad
        out[f"f{frame}_direction_distance"] = dd
        out[f"f{frame}_irradiance"] = irr[w].copy()
        out[f"f{frame}_depth"] = dep[w].copy()
        out[f"f{frame}_rotation"] = rot
        out[f"f{frame}_tex_taps"] = np.int64(taps)
        out[f"f{frame}_mip_taps"] = np.int64(mtaps)
        ping = w
    # inputs, so the replay does not depend on regenerating the procedural scene bit for bit
    import ctypes as C

    out["in_uniform"] = np.frombuffer(bytes(sc.uniform), dtype=np.uint8)
    out["in_sdf_data"] = np.frombuffer(bytes(sc.sdf_data), dtype=np.uint8)
    out["in_atlas_data"] = np.frombuffer(bytes(sc.atlas_data), dtype=np.uint8)
    out["in_sdf"] = sc.sdf.numpy().view(np.uint16)
    out["in_mip"] = sc.mip.numpy().view(np.uint16)
    out["in_light"] = sc.light.numpy().view(np.uint16)
    out["in_depth"] = sc.depth.numpy()
    out["in_chunks"], out["in_cull"] = sc.chunks, sc.cull
    out["in_objects"] = sc.objects.view(np.uint8)
    out["in_tiles"] = sc.tiles.view(np.uint8)
    out["in_sky"] = sc.sky.view(np.uint16)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spirv_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
