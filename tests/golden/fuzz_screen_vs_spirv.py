"""Fuzz of the oracle's SDFReflection / SDFShadow restatements against the reference's SHIPPED binaries executed live (build container only):

    python tests/golden/fuzz_screen_vs_spirv.py [seed] [seconds]

Random G-buffers per configuration (depths incl. sky and far pixels, octahedral normals incl. the exact poles that flip importanceSampleGGX's `up`
vector, roughness on the branch thresholds 0.05 / 0.45), random noise textures, frame numbers, trims, light parameters; the RGBA16F reflection image
and the R32UI shadow words must be bit-identical."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import abi  # noqa: E402
from oracle import binding as o  # noqa: E402
from tests.golden import make_spirv_golden as base  # noqa: E402
from tests.golden import make_spirv_golden_screen as gs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run(seed=0, seconds=300.0, max_configs=None, verbose=True):
    rng = np.random.default_rng(seed)
    sc = base.golden_scene()
    g = np.load(os.path.join(HERE, "spirv_golden.npz"))
    u = abi.DDGIUniform.from_buffer_copy(g["in_uniform"].tobytes())
    u.normalBias = 0.1
    sc.uniform = u
    irr, dep = g["f1_irradiance"], g["f1_depth"]
    k = bad = px = 0
    t0 = time.time()
    while time.time() - t0 < seconds and (max_configs is None or k < max_configs):
        sobol = rng.integers(0, 256, (1, 256, 4), dtype=np.uint8)
        scr = rng.integers(0, 256, (128, 128, 4), dtype=np.uint8)
        # reflection: one 16 x 16 workgroup
        W, H = 16, 16
        eye, vpi = gs.camera(W, H)
        eye = eye + rng.uniform(-0.5, 0.5, 3)
        depth, nrm, pbr = gs.gbuffer(W, H, int(rng.integers(0, 1 << 30)))
        nrm[2, :4, :2] = [[0.0, 0.0], [1e-4, 0.0], [0.0, -1e-4], [0.02, 0.02]]  # around the +z pole: |N.z| >= 0.999 takes the other `up`
        pbr[3, :6, 1] = [0.05, 0.45, np.nextafter(np.float32(0.05), 0), np.nextafter(np.float32(0.45), 1), 0.0, 1.0]
        approx, frames, trim, inten = int(rng.integers(0, 2)), int(rng.integers(0, 1000)), float(rng.uniform(0.1, 1.0)), float(rng.uniform(0.2, 2.0))
        want, _ = gs.run_reflection(sc, u, irr, dep, depth, nrm, pbr, sobol, scr, eye, vpi, approx, frames, trim, inten)
        push = abi.make_reflection_push(eye, vpi, frames, trim, inten, approx)
        got = np.full((H, W, 4), 0x3555, dtype=np.uint16)
        o.sdf_reflection(sc, irr, dep, push, depth, nrm, pbr, sobol, scr, got)
        if not np.array_equal(got, want):
            bad += 1
            if verbose:
                print("MISMATCH reflection config", k, int((got != want).any(-1).sum()), "pixels", flush=True)
        # shadow: 2 x 2 workgroups of 8 x 4
        W, H = 16, 8
        _, vpi = gs.camera(W, H)
        depth, nrm, _ = gs.gbuffer(W, H, int(rng.integers(0, 1 << 30)))
        ltype = float(rng.integers(0, 3))
        ldir = rng.normal(size=3); ldir /= np.linalg.norm(ldir)
        light = ([1, 1, 1, 1], [*rng.uniform(-4, 4, 3), 1.0], [*ldir, float(rng.uniform(0.0, 0.5))], float(rng.uniform(0.5, 5)), float(rng.uniform(5, 80)), ltype, float(rng.uniform(0.1, 0.9)))
        frames, bias = int(rng.integers(0, 1000)), float(rng.uniform(0.01, 0.3))
        want, _ = gs.run_shadow(sc, light, depth, nrm, sobol, scr, vpi, frames, bias)
        col, lpos, ld, inten, radius, lt, angle = light
        got = np.full((H // 4, W // 8), 0xDEADBEEF, dtype=np.uint32)
        o.sdf_shadow(sc.sdf_data, sc.sdf, sc.mip, abi.make_light(list(col) + list(lpos) + list(ld) + [inten, radius, lt, angle]), vpi, frames, bias, depth, nrm, sobol, scr, got)
        if not np.array_equal(got, want):
            bad += 1
            if verbose:
                print("MISMATCH shadow config", k, [hex(int(a)) for a in got.reshape(-1)], [hex(int(a)) for a in want.reshape(-1)], flush=True)
        px += 256 + 128
        k += 1
    if verbose:
        print("configs", k, "pixels", px, "mismatches", bad, "in", round(time.time() - t0), "s")
    return k, px, bad


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, float(sys.argv[2]) if len(sys.argv) > 2 else 300.0)
