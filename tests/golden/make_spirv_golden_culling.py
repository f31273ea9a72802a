"""Golden vectors for the surface-cache culling (first half) from the reference's SHIPPED SPIR-V:

    python tests/golden/make_spirv_golden_culling.py        (build container only: needs /root/reference)

Executes Assets/shaders/spv/SDF/SDFCulling.comp.spv with oracle/spirv/interp.py on the Cornell surface cache (8 objects, chunk size
0.32) for the 4x4x4 workgroups listed in GROUPS (chunks around the room, its walls and empty space), twice: with a generous
culledObjectsCapacity and with one so small that most lists overflow (the `objectsStart + objectsSize > capacity` branch).
Stores the dispatch order (chunk addresses in execution order) and the resulting chunk / cull buffers; the inputs are procedural
(scenes.cornell_scene).  tests/test_spirv_golden.py::test_surface_culling_matches_shipped_spirv replays them through the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import abi, scenes  # noqa: E402
from oracle.spirv import interp as si  # noqa: E402

SPV = "/root/reference/Assets/shaders/spv/SDF/SDFCulling.comp.spv"
F = np.float32
GROUPS = [(gx, gy, gz) for gz in (1, 4, 5, 8) for gy in (2, 5, 8) for gx in (1, 4, 5, 6, 9)]  # 60 of the 1000 workgroups
CAPACITIES = (4096, 150)


def golden_scene():
    return scenes.cornell_scene(res=32, counts=(2, 2, 2), rays=32, atlas_res=256)


def vec(a):
    return [F(x) for x in a]


def mat_cols(m16):
    return [vec(m16[c * 4:c * 4 + 4]) for c in range(4)]


def dispatch_order():
    out = []
    for (gx, gy, gz) in GROUPS:
        for z in range(4):
            for y in range(4):
                for x in range(4):
                    out.append(((gz * 4 + z) * abi.CHUNKS_RESOLUTION + gy * 4 + y) * abi.CHUNKS_RESOLUTION + gx * 4 + x)
    return np.asarray(out, dtype=np.int32)


def run(sc, capacity):
    mod = si.Module(SPV)
    a = sc.atlas_data
    chunks = [0] * abi.CHUNKS_RESOLUTION ** 3
    cull = [1] + [0] * 8191
    objs = [[[vec(o["objectBounds"]), [int(x) for x in o["tileOffset"]], [0, 0], mat_cols(o["transform"]), vec(o["extends"])] for o in sc.objects]]
    for b, v in ((0, objs), (1, [chunks]), (2, [cull])):
        mod.storage[mod.global_by_binding(0, b)] = [v]
    (pc,) = mod.global_by_storage(9)
    mod.storage[pc] = [[[vec(a.cameraPos), F(a.chunkSize), int(capacity), int(a.resolution), int(a.objectsCount), 0]]]
    n = si.dispatch(mod, GROUPS)
    return np.asarray(chunks, dtype=np.uint32), np.asarray(cull, dtype=np.uint32), n


def main():
    sc = golden_scene()
    out = {"order": dispatch_order(), "capacities": np.asarray(CAPACITIES, dtype=np.int32)}
    total = 0
    for cap in CAPACITIES:
        chunks, cull, n = run(sc, cap)
        out[f"chunks_{cap}"], out[f"cull_{cap}"] = chunks, cull
        total += n
        print("capacity", cap, "lists", int((chunks[1:] != 0).sum()), "words used", int(cull[0]), "slot0", int(chunks[0]))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spirv_golden_culling.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "SPIR-V instructions executed:", total)


if __name__ == "__main__":
    main()
