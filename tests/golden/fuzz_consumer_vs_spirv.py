"""Fuzz of the oracle's SampleProbe / sampleIrradiance restatement against the reference's SHIPPED SampleProbe.comp.spv executed live (build container only):

    python tests/golden/fuzz_consumer_vs_spirv.py [seed] [seconds]

Random cameras (inside and outside the probe volume), depths incl. sky pixels, octahedral normals, normalBias and atlases from random fp16 values
(so every bilinear tap, border and Chebyshev branch sees arbitrary data); the RGBA32F image must be bit-identical."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import abi  # noqa: E402
from oracle import binding as o  # noqa: E402
from oracle.spirv import interp as si  # noqa: E402
from tests.golden import make_spirv_golden_consumer as mc  # noqa: E402
from tests.golden.make_spirv_golden import ddgi_block, mat_cols, vec  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run(seed=0, seconds=300.0, max_configs=None, verbose=True):
    rng = np.random.default_rng(seed)
    g = np.load(os.path.join(HERE, "spirv_golden.npz"))
    k = bad = px = 0
    t0 = time.time()
    while time.time() - t0 < seconds and (max_configs is None or k < max_configs):
        u = abi.DDGIUniform.from_buffer_copy(g["in_uniform"].tobytes())
        u.normalBias = float(rng.choice([0.0, 0.1, 0.5]))
        u.maxDistance = float(rng.uniform(1.0, 12.0))
        if k % 2:
            irr, dep = g["f1_irradiance"], g["f1_depth"]
        else:
            irr = rng.uniform(0, 4, g["f1_irradiance"].shape).astype(np.float16).view(np.uint16)
            dep = rng.uniform(0, 9, g["f1_depth"].shape).astype(np.float16).view(np.uint16)
        W, H = 16, 16
        eye = rng.uniform(-9, 9, 3); target = rng.uniform(-3, 3, 3)
        vp = mc.look_at_perspective(eye.copy(), target.copy(), np.array([0.0, 1.0, 0.0]), np.radians(float(rng.uniform(40, 100))), W / H, 0.1, 50.0)
        vpi = np.linalg.inv(vp).astype(np.float32)
        depth = rng.uniform(0.5, 0.9999, (H, W)).astype(np.float32)
        depth[rng.random((H, W)) < 0.05] = 1.0
        nrm = np.zeros((H, W, 4), dtype=np.float32)
        nrm[..., :2] = rng.uniform(-1, 1, (H, W, 2))
        mod = si.Module(mc.SPV)
        out = mc.OutImage(H, W)
        bind = {0: out, 1: si.Texture2D(irr.view(np.float16), repeat=True), 2: si.Texture2D(dep.view(np.float16), repeat=True), 3: ddgi_block(u),
                4: mc.FloatTexture(depth), 5: mc.FloatTexture(nrm), 6: [vec([*eye, 1.0]), mat_cols(vpi.T.reshape(16))]}
        for b, v in bind.items():
            mod.storage[mod.global_by_binding(0, b)] = [v]
        lx, ly, _ = mod.local_size
        si.dispatch(mod, [(gx, gy, 0) for gy in range((H + ly - 1) // ly) for gx in range((W + lx - 1) // lx)])
        got = o.sample_probe(u, irr, dep, depth, nrm, np.array([*eye, 1.0], dtype=np.float32), vpi.T.reshape(16).copy())
        same = np.array_equal(got.view(np.uint32), out.a.view(np.uint32)) or (np.isnan(got) == np.isnan(out.a)).all() and np.array_equal(np.nan_to_num(got), np.nan_to_num(out.a))
        if not same:
            bad += 1
            if verbose:
                print("MISMATCH config", k, int((got.view(np.uint32) != out.a.view(np.uint32)).any(-1).sum()), "pixels", flush=True)
        px += W * H
        k += 1
    if verbose:
        print("configs", k, "pixels", px, "mismatches", bad, "in", round(time.time() - t0), "s")
    return k, px, bad


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, float(sys.argv[2]) if len(sys.argv) > 2 else 300.0)
