"""Fuzz of the oracle's trace against the reference's SHIPPED GISDFRays.comp.spv (build container only: needs /root/reference):

    python tests/golden/fuzz_oracle_vs_spirv.py [seed] [seconds]

Random small scenes — the Cornell room with 1-4 cascades, open cities with random lots / probe placement under a 2x2-texel cube sky, probe grids
pushed into geometry and onto cascade faces — random frame rotations; every ray buffer value and both tap counts of the C++ oracle must equal what the
interpreter gets from the binary.  Last run: 571 configurations, 73 088 rays, 0 mismatches (seed 1, 420 s).
tests/test_spirv_golden.py::test_fuzz_oracle_vs_shipped_spirv runs a bounded slice of it when /root/reference is present."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import scenes  # noqa: E402
from oracle import binding as o  # noqa: E402
from tests.golden import make_spirv_golden as base  # noqa: E402


def random_scene(rng, k):
    kind = k % 3
    if kind == 0:
        sc = scenes.cornell_scene(res=32, counts=(2, 2, 2), rays=16, atlas_res=256, cascades=int(rng.integers(1, 5)))
    elif kind == 1:
        sc = scenes.city_scene(res=32, lots=int(rng.integers(2, 4)), counts=(2, 2, 2), rays=16, atlas_res=256, seed=int(rng.integers(0, 1000)),
                               start=tuple(float(x) for x in rng.uniform([-12, 1, -12], [-2, 20, -2])), step=tuple(float(x) for x in rng.uniform(3, 9, 3)))
        sc.sky_face, sc.sky = 2, rng.uniform(0, 3, (6, 2, 2, 4)).astype(np.float16)
    else:  # probes pushed around: some inside geometry, some on cascade faces
        sc = scenes.cornell_scene(res=32, counts=(2, 1, 2), rays=32, atlas_res=256)
        sc.uniform.startPosition[:3] = [float(x) for x in rng.choice([-6.4, -5.0, -3.3, 0.0, 6.4], 3)]
        sc.uniform.step[:3] = [float(x) for x in rng.uniform(0.5, 6.0, 3)]
    if sc.sky is None:
        sc.sky_face, sc.sky = 1, rng.uniform(0, 3, (6, 1, 1, 4)).astype(np.float16)
    return sc


def run(seed=0, seconds=300.0, max_configs=None, verbose=True):
    rng = np.random.default_rng(seed)
    bad = total = k = 0
    t0 = time.time()
    while time.time() - t0 < seconds and (max_configs is None or k < max_configs):
        sc = random_scene(rng, k)
        rot = scenes.frame_rotation(int(rng.integers(0, 100000)))
        rad, dd, _, taps, mtaps = base.run_trace(sc, rot)
        orad, odd, _, cn = o.OracleScene(sc).trace(rot)
        same = np.array_equal(rad, orad) and np.array_equal(dd, odd) and cn["texTaps"] == taps and cn["mipTaps"] == mtaps
        total += rad.shape[0] * rad.shape[1]
        if not same:
            bad += 1
            if verbose:
                print("MISMATCH config", k, int((rad != orad).sum()), int((dd != odd).sum()), cn["texTaps"], taps, cn["mipTaps"], mtaps, flush=True)
        k += 1
    if verbose:
        print("configs", k, "rays", total, "mismatching configs", bad, "in", round(time.time() - t0), "s")
    return k, total, bad


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, float(sys.argv[2]) if len(sys.argv) > 2 else 300.0)
