"""Fuzz of the oracle's global-SDF build restatement against the reference's SHIPPED SDFRasterizeModel(NoRead).comp.spv and GlobalSDFMipmap.comp.spv
executed live (build container only):

    python tests/golden/fuzz_sdfbuild_vs_spirv.py [seed] [seconds]

Per configuration: 2-4 random synthetic mesh distance fields (spheres / boxes of random resolution, extent, maxDistance margin, rotation and
translation), a random cascade (mesh mip 0 or 1), the NoRead pipeline on a random subset of the models followed by the READ_DISTANCE pipeline on the
rest, on 3 random 8x8x8 workgroups of the chunk; then one GlobalSDFMipmap pass (4x downsample or flood, random offsets) over a volume of random fp16
distances.  Every voxel the binaries wrote must be bit-identical."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import meshsdf  # noqa: E402
from oracle import binding as o  # noqa: E402
from tests.golden import make_spirv_golden_sdfbuild as g  # noqa: E402


def run(seed=0, seconds=300.0, max_configs=None, verbose=True):
    rng = np.random.default_rng(seed)
    RES = g.RES
    one = np.float16(1.0).view(np.uint16)
    k = bad = voxels = touched = 0
    t0 = time.time()
    while time.time() - t0 < seconds and (max_configs is None or k < max_configs):
        nm = int(rng.integers(2, 5))
        meshes = []
        for j in range(nm):
            kind = str(rng.choice(["sphere", "box"]))
            dims = tuple(int(x) for x in rng.choice([8, 12, 16, 20], 3))
            ext = tuple(float(x) for x in rng.uniform(0.4, 1.6, 3))
            meshes.append(meshsdf.synthetic(kind, dims, ext, float(rng.uniform(0.15, 0.4)), g.rot_y(float(rng.uniform(0, 360)), [float(x) for x in rng.uniform(-2.5, 2.5, 3)]), f"m{j}"))
        cascade = int(rng.integers(0, 2))
        (centre, D) = g.cascades()[cascade]
        objs = [o.sdf_object_data(m, cascade) for m in meshes]
        ids = [int(x) for x in rng.permutation(nm)]
        split = int(rng.integers(1, nm + 1))
        g.GROUPS = list(dict.fromkeys(tuple(int(x) for x in rng.integers(0, 4, 3)) for _ in range(3)))
        mask = np.zeros((RES, RES, RES), dtype=bool)
        for gx, gy, gz in g.GROUPS:
            mask[gz * 8:gz * 8 + 8, gy * 8:gy * 8 + 8, gx * 8:gx * 8 + 8] = True
        want = np.full((RES, RES, RES * g.CASC), one, dtype=np.uint16)
        got = want.copy()
        g.run_rasterize(False, want, objs, meshes, cascade, (0, 0, 0), ids[:split])
        o.sdf_rasterize_chunk(got, objs, meshes, cascade, centre, D, RES, cascade, (0, 0, 0), ids[:split], read=False)
        sl = (slice(None), slice(None), slice(cascade * RES, (cascade + 1) * RES))
        ok = np.array_equal(got[sl][mask], want[sl][mask])
        got[sl][~mask] = want[sl][~mask]
        if split < nm:
            g.run_rasterize(True, want, objs, meshes, cascade, (0, 0, 0), ids[split:])
            o.sdf_rasterize_chunk(got, objs, meshes, cascade, centre, D, RES, cascade, (0, 0, 0), ids[split:], read=True)
            ok = ok and np.array_equal(got[sl][mask], want[sl][mask])
        # one mip pass over random distances
        mres = RES // 4
        flood = bool(rng.integers(0, 2))
        if flood:
            src = rng.uniform(-1, 1, (mres, mres, mres * g.CASC)).astype(np.float16).view(np.uint16)
            wm = np.full((mres, mres, mres), one, dtype=np.uint16)
            gm = wm.copy()
            g.run_mip(src, wm, mres, mres, 1, cascade * mres, 0, 2 * D)
            o.sdf_mip_pass(src, gm, mres, mres, 1, cascade * mres, 0, 2 * D)
        else:
            src = rng.uniform(-1, 1, (RES, RES, RES * g.CASC)).astype(np.float16).view(np.uint16)
            wm = np.full((mres, mres, mres * g.CASC), one, dtype=np.uint16)
            gm = wm.copy()
            g.run_mip(src, wm, mres, RES, 4, cascade * RES, cascade * mres, 2 * D)
            o.sdf_mip_pass(src, gm, mres, RES, 4, cascade * RES, cascade * mres, 2 * D)
        ok = ok and np.array_equal(gm, wm)
        if not ok:
            bad += 1
            if verbose:
                print("MISMATCH config", k, "cascade", cascade, "flood", flood, flush=True)
        voxels += int(mask.sum()) + wm.size
        touched += int((want[sl][mask] != one).sum())
        k += 1
    if verbose:
        print("configs", k, "voxels", voxels, "of which rasterised below 1.0:", touched, "mismatches", bad, "in", round(time.time() - t0), "s")
    return k, touched, bad


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, float(sys.argv[2]) if len(sys.argv) > 2 else 300.0)
