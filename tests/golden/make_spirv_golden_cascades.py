"""Golden vectors for the CASCADED global SDF (the reference always runs 2 cascades, GlobalDistanceField.cpp:182-191) from the SHIPPED SPIR-V:

    python tests/golden/make_spirv_golden_cascades.py        (build container only: needs /root/reference)

Executes Assets/shaders/spv/DDGI/GISDFRays.comp.spv with oracle/spirv/interp.py on the small Cornell scene of make_spirv_golden.py with 2 and
with 4 nested cascades (half extents 1 : 2.5 [: 5 : 10], inner cascades off-centre, side by side along x in one volume as the reference lays
them out; most probes sit outside cascade 0, so rays enter it from outside, leave it, and continue in the next cascade).  One frame each.
A third vector is an OPEN scene (reduced city, 4x4-texel cube sky) for the miss / sky path.  Stores the ray buffers and the tap counts; the scene is procedural (luxgi_b200.scenes.cornell_scene(cascades=K)) and its SDF / mip bytes are
pinned by a CRC.  tests/test_spirv_golden.py::test_cascaded_trace_matches_shipped_spirv replays them through the oracle."""
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import scenes  # noqa: E402
from tests.golden import make_spirv_golden as base  # noqa: E402

CASCADES = (2, 4)


def golden_scene(k):
    sc = scenes.cornell_scene(res=32, counts=(3, 3, 2), rays=32, atlas_res=256, hysteresis=0.9, gamma=2.2, cascades=k)
    sky = np.zeros((6, 1, 1, 4), dtype=np.float16)
    for f in range(6):
        sky[f, 0, 0] = [0.25 * (f + 1), 0.1 * (6 - f), 0.5 + 0.05 * f, 1.0]
    sc.sky_face, sc.sky = 1, sky
    return sc


def open_scene():
    """A reduced city (open sky, emissive window bands, ground slab) with a 4x4-texel cube sky: most rays leave the cascade and take the
    bilinear sky texel, long open-space steps take the `stepDistance = chunkSizeDistance` branch."""
    sc = scenes.city_scene(res=32, lots=3, counts=(3, 2, 3), rays=32, atlas_res=256, hysteresis=0.9, gamma=2.2)
    sc.sky_face, sc.sky = 4, np.random.default_rng(5).uniform(0.0, 3.0, (6, 4, 4, 4)).astype(np.float16)
    return sc


def crc(sc):
    return np.uint32(zlib.crc32(sc.sdf.numpy().tobytes() + sc.mip.numpy().tobytes()))


def main():
    out = {}
    for k in CASCADES:
        sc = golden_scene(k)
        rot = scenes.frame_rotation(k)
        t0 = time.time()
        rad, dd, n, taps, mtaps = base.run_trace(sc, rot)
        print(f"{k} cascades: {n} SPIR-V instructions, {taps} tex taps, {mtaps} mip taps, {time.time() - t0:.1f} s", flush=True)
        out[f"c{k}_rotation"], out[f"c{k}_radiance"], out[f"c{k}_direction_distance"] = rot, rad, dd
        out[f"c{k}_tex_taps"], out[f"c{k}_mip_taps"], out[f"c{k}_crc"] = np.int64(taps), np.int64(mtaps), crc(sc)
    sc = open_scene()
    rot = scenes.frame_rotation(1)
    t0 = time.time()
    rad, dd, n, taps, mtaps = base.run_trace(sc, rot)
    print(f"open city: {n} SPIR-V instructions, {taps} tex taps, {mtaps} mip taps, {time.time() - t0:.1f} s; "
          f"misses {(dd.view(np.float16)[..., 3] >= 60000).mean():.2f}", flush=True)
    out["open_rotation"], out["open_radiance"], out["open_direction_distance"] = rot, rad, dd
    out["open_tex_taps"], out["open_mip_taps"], out["open_crc"] = np.int64(taps), np.int64(mtaps), crc(sc)
    out["open_light_crc"] = np.uint32(zlib.crc32(sc.light.numpy().tobytes() + sc.depth.numpy().tobytes()))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spirv_golden_cascades.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
