"""Fuzz of the oracle's blend + border against the reference's SHIPPED {Irradiance,Depth}ProbeUpdate.comp.spv and {Irradiance,Depth}BorderUpdate.comp.spv
(build container only: needs /root/reference):

    python tests/golden/fuzz_blend_vs_spirv.py [seed] [seconds]

Random probe-volume parameters (hysteresis, gamma, sharpness, maxDistance), random fp16 ray buffers (zero / large radiance, hit distances from 0 to the
60000 of a miss, weights that fall below the 1e-8 gate), random previous atlases, first and later frames.  In the oracle's `unfused` mode (the literal
arithmetic of the binaries) interiors and borders must be bit-identical; in the contract's FMA mode within 1 fp16 ulp.
Last run: 136 configurations, 399 456 atlas values, 0 mismatches, worst FMA-mode difference 1 ulp (seed 2, 600 s)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import abi  # noqa: E402
from oracle import binding as o  # noqa: E402
from tests.golden import make_spirv_golden as base  # noqa: E402
from tests.util import ulp16_diff  # noqa: E402


class _Sc:
    pass


def run(seed=0, seconds=300.0, max_configs=None, verbose=True):
    rng = np.random.default_rng(seed)
    k = bad = texels = 0
    worst = 0
    t0 = time.time()
    while time.time() - t0 < seconds and (max_configs is None or k < max_configs):
        counts = (int(rng.integers(1, 3)), 1, int(rng.integers(1, 3)))
        R = int(rng.choice([8, 16, 24]))
        u = abi.make_uniform((0, 0, 0), (1, 1, 1), counts, R, max_distance=float(rng.uniform(0.5, 5.0)), sharpness=float(rng.choice([1.0, 8.0, 50.0, 80.0])),
                             hysteresis=float(rng.choice([0.0, 0.5, 0.9, 0.98])), gamma=float(rng.choice([0.85, 1.0, 2.2, 5.0])))
        P = abi.probe_count(u)
        d = rng.normal(size=(1, R, 3)).astype(np.float32)  # the frame's directions are the same for every probe (GISDFRays.comp:73); the hoisted blend relies on it
        d = np.repeat(d / np.linalg.norm(d, axis=-1, keepdims=True), P, axis=0)
        dist = rng.choice([0.0, 0.004, 0.3, 2.0, 7.5, 60000.0], (P, R, 1)).astype(np.float32) * rng.uniform(0.5, 1.0, (P, R, 1)).astype(np.float32)
        dd = np.concatenate([d, dist], -1).astype(np.float16).view(np.uint16)
        rad = (rng.choice([0.0, 0.01, 1.0, 40.0], (P, R, 1)) * rng.uniform(0, 1, (P, R, 4))).astype(np.float16).view(np.uint16)
        first = bool(rng.integers(0, 2))
        prev_i, prev_d = o.new_atlases(u)
        prev_i[...] = rng.uniform(0, 2, prev_i.shape).astype(np.float16).view(np.uint16)
        prev_d[...] = rng.uniform(0, 3, prev_d.shape).astype(np.float16).view(np.uint16)
        sc = _Sc()
        sc.uniform = u
        gi, gd = o.new_atlases(u)
        base.run_blend(sc, "Irradiance", rad, dd, prev_i, prev_d, gi, gd, first)
        base.run_blend(sc, "Depth", rad, dd, prev_i, prev_d, gi, gd, first)
        for unfused in (True, False):
            o.set_unfused(unfused)
            try:
                for naive in (True, False):
                    oi, od = o.new_atlases(u)
                    o.blend(u, rad, dd, prev_i, prev_d, oi, od, first_frame=first, naive=naive)
                    ui, ud = int(ulp16_diff(oi, gi).max()), int(ulp16_diff(od, gd).max())
                    worst = max(worst, ui, ud) if not unfused else worst
                    if (unfused and (ui or ud)) or (not unfused and max(ui, ud) > 1):
                        bad += 1
                        if verbose:
                            print("MISMATCH blend config", k, "unfused", unfused, "naive", naive, "ulp", ui, ud, flush=True)
            finally:
                o.set_unfused(False)
        bi, bd = gi.copy(), gd.copy()
        base.run_border(sc, "Irradiance", bi, bd)
        base.run_border(sc, "Depth", bi, bd)
        oi, od = gi.copy(), gd.copy()
        o.border(u, oi, od)
        if not (np.array_equal(oi, bi) and np.array_equal(od, bd)):
            bad += 1
            if verbose:
                print("MISMATCH border config", k, flush=True)
        texels += gi.size + gd.size
        k += 1
    if verbose:
        print("configs", k, "atlas values", texels, "mismatches", bad, "worst FMA-mode ulp", worst, "in", round(time.time() - t0), "s")
    return k, texels, bad


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, float(sys.argv[2]) if len(sys.argv) > 2 else 300.0)
