"""Fuzz of the oracle's SDFCulling restatement against the reference's SHIPPED SDFCulling.comp.spv executed live (build container only):

    python tests/golden/fuzz_culling_vs_spirv.py [seed] [seconds]

Random object buffers (1-14 oriented boxes with random bounding spheres), chunk sizes, culledObjectsCapacity (generous to overflowing) and random
subsets of the 1000 workgroups; the oracle, run in the same execution order, must reproduce the chunk and cull buffers word for word."""
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import abi  # noqa: E402
from oracle import binding as o  # noqa: E402
from oracle.spirv import interp as si  # noqa: E402
from tests.golden.make_spirv_golden_culling import SPV, mat_cols, vec  # noqa: E402

F = np.float32
N = abi.CHUNKS_RESOLUTION


def run(seed=0, seconds=300.0, max_configs=None, verbose=True):
    rng = np.random.default_rng(seed)
    k = bad = lists = 0
    t0 = time.time()
    while time.time() - t0 < seconds and (max_configs is None or k < max_configs):
        n = int(rng.integers(1, 15))
        chunk = float(rng.choice([0.2, 0.32, 0.5]))
        objects = np.zeros(n, dtype=abi.OBJECT_DTYPE)
        for j in range(n):
            c = rng.uniform(-0.45, 0.45, 3) * chunk * N
            ext = rng.uniform(0.1, 3.0, 3)
            a = float(rng.uniform(0, 2 * math.pi))
            m = np.eye(4)
            m[:3, :3] = [[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]]
            m[:3, 3] = c
            objects["objectBounds"][j] = [*c, float(np.linalg.norm(ext) * rng.uniform(0.7, 1.3))]
            objects["transform"][j] = m.T.reshape(16)
            objects["extends"][j] = [*ext, 1.0]
            objects["tileOffset"][j] = 1 + 6 * j + np.arange(6)
        groups = [tuple(int(x) for x in rng.integers(0, 10, 3)) for _ in range(5)]
        groups = list(dict.fromkeys(groups))
        cap = int(rng.choice([60, 300, 4096]))
        mod = si.Module(SPV)
        chunks = [0] * N ** 3
        cull = [1] + [0] * 8191
        objs = [[[vec(ob["objectBounds"]), [int(x) for x in ob["tileOffset"]], [0, 0], mat_cols(ob["transform"]), vec(ob["extends"])] for ob in objects]]
        for b, v in ((0, objs), (1, [chunks]), (2, [cull])):
            mod.storage[mod.global_by_binding(0, b)] = [v]
        (pc,) = mod.global_by_storage(9)
        mod.storage[pc] = [[[vec([0, 0, 0, 0]), F(chunk), cap, 256, n, 0]]]
        si.dispatch(mod, groups)
        order = np.asarray([((gz * 4 + z) * N + gy * 4 + y) * N + gx * 4 + x for (gx, gy, gz) in groups for z in range(4) for y in range(4) for x in range(4)], dtype=np.int32)
        data = abi.GlobalSurfaceAtlasData()
        data.cameraPos[:] = [0.0, 0.0, 0.0]
        data.chunkSize, data.culledObjectsCapacity, data.resolution, data.objectsCount, data.padding = chunk, cap, 256, n, 0
        got_chunks, got_cull = o.surface_cull(data, objects, order=order, emulate_slot0=True, capacity_words=8192)
        want_chunks, want_cull = np.asarray(chunks, dtype=np.uint32), np.asarray(cull, dtype=np.uint32)
        if not (np.array_equal(got_chunks, want_chunks) and np.array_equal(got_cull, want_cull)):
            bad += 1
            if verbose:
                print("MISMATCH config", k, int((got_chunks != want_chunks).sum()), int((got_cull != want_cull).sum()), flush=True)
        lists += int((want_chunks[1:] != 0).sum())
        k += 1
    if verbose:
        print("configs", k, "non-empty chunk lists", lists, "mismatches", bad, "in", round(time.time() - t0), "s")
    return k, lists, bad


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, float(sys.argv[2]) if len(sys.argv) > 2 else 300.0)
