"""Fuzz of the oracle's surface-cache lighting restatements against the reference's SHIPPED SDFDeferredLight.frag.spv and SDFAtlasIndirectLight.frag.spv
executed live, one invocation per fragment (build container only):

    python tests/golden/fuzz_light_vs_spirv.py [seed] [seconds]

Per configuration 64 random surface texels of the Cornell surface cache with perturbed normals and random albedo / metallic / roughness, a random
light (directional / point / spot, random position, direction, intensity, radius, cone), camera and shadow bias; the fp16 values the additive pass
leaves in an empty light cache must be bit-identical."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import abi  # noqa: E402
from oracle import binding as o  # noqa: E402
from oracle.spirv import interp as si  # noqa: E402
from tests.golden import make_spirv_golden_directlight as dl  # noqa: E402
from tests.golden.make_spirv_golden import ddgi_block, vec  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
F = np.float32
G = 8
SPV_DIRECT = "/root/reference/Assets/shaders/spv/SDF/SDFDeferredLight.frag.spv"
SPV_INDIRECT = "/root/reference/Assets/shaders/spv/SDF/SDFAtlasIndirectLight.frag.spv"
TILES = [[[vec([0, 0, 1, 1]), [vec([1, 0, 0, 0]), vec([0, 1, 0, 0]), vec([0, 0, 1, 0]), vec([0, 0, 0, 1])], vec([1, 1, 1, 1])]]]


def run_frag(mod, n):
    gids = {mod.names.get(x): x for x in mod.globals_}
    res = np.zeros((n, 4), dtype=np.float32)
    for k in range(n):
        inv = si.Invocation(mod, {}, {})
        inv.g[gids["inTileUV"]] = si.Ptr([vec([((k % G) + 0.5) / G, ((k // G) + 0.5) / G])])
        inv.g[gids["inTileAddress"]] = si.Ptr([0])
        inv.g[gids["inPosition"]] = si.Ptr([vec([0, 0, 0, 1])])
        for _ in inv.run():
            pass
        res[k] = inv.g[gids["outColor"]].load()
    return res


def bind(mod, table):
    for b, v in table.items():
        gid = mod.global_by_binding(0, b)
        if gid is not None:
            mod.storage[gid] = [v]


def run(seed=0, seconds=300.0, max_configs=None, verbose=True):
    rng = np.random.default_rng(seed)
    sc = dl.golden_scene()
    gb = sc.meta["gbuffer"]
    g = np.load(os.path.join(HERE, "spirv_golden.npz"))
    u = abi.DDGIUniform.from_buffer_copy(g["in_uniform"].tobytes())
    u.normalBias = 0.1
    irr, dep = g["f1_irradiance"], g["f1_depth"]
    d = sc.sdf_data
    sdf_block = [[vec(d.cascadePosDistance[i]) for i in range(4)], vec(d.cascadeVoxelSize), d.cascadesCount, F(d.resolution), F(d.nearPlane), F(d.farPlane)]
    n = G * G
    k = bad = texels = 0
    t0 = time.time()
    while time.time() - t0 < seconds and (max_configs is None or k < max_configs):
        pick = rng.choice(len(gb["texel"]), n, replace=False)
        pos = gb["pos"][pick].astype(np.float32)
        nrm = gb["normal"][pick].astype(np.float32) + rng.normal(scale=0.3, size=(n, 3)).astype(np.float32)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        octn = dl.oct_encode(nrm)
        alb = rng.uniform(0, 1, (n, 3)).astype(np.float32)
        pbr = np.stack([rng.choice([0.0, 0.3, 1.0], n), rng.uniform(0.02, 1.0, n)], -1).astype(np.float32)
        color = np.concatenate([alb, pos[:, :1]], -1).reshape(G, G, 4)
        normal = np.concatenate([octn, pos[:, 1:]], -1).reshape(G, G, 4)
        pbrt = np.concatenate([pbr, np.zeros((n, 2), np.float32)], -1).reshape(G, G, 4)
        cam = [*rng.uniform(-4, 4, 3), float(rng.uniform(0.01, 0.3))]
        ldir = rng.normal(size=3); ldir /= np.linalg.norm(ldir)
        lv = [1.0, float(rng.uniform(0.2, 1)), float(rng.uniform(0.2, 1)), 1.0, *rng.uniform(-4.5, 4.5, 3), 1.0, *ldir, 0.0, float(rng.uniform(0.3, 6)), float(rng.uniform(1, 80)),
              float(rng.integers(0, 3)), float(rng.uniform(0.05, 0.95))]
        light = [vec(lv[0:4]), vec(lv[4:8]), vec(lv[8:12]), F(lv[12]), F(lv[13]), F(lv[14]), F(lv[15])]
        normals = o.octohedral_to_direction(octn)
        texel = np.arange(n, dtype=np.uint32)
        # direct
        mod = si.Module(SPV_DIRECT)
        bind(mod, {0: TILES, 1: si.Texture2D(color, repeat=False), 2: si.Texture2D(normal, repeat=False), 4: si.Texture2D(pbrt, repeat=False),
                   6: [vec(cam), light, sdf_block], 7: si.Texture3D(sc.mip.numpy()), 8: si.Texture3D(sc.sdf.numpy())})
        want = run_frag(mod, n).astype(np.float16).view(np.uint16)
        cache = np.zeros((n, 4), dtype=np.uint16)
        o.surface_direct_light(sc.sdf_data, sc.sdf, sc.mip, abi.make_light(lv), np.float32(cam), cache, texel, pos, normals, alb, pbr)
        if not np.array_equal(cache, want):
            bad += 1
            if verbose:
                print("MISMATCH direct config", k, "type", lv[14], int((cache != want).sum()), "values", flush=True)
        # indirect
        mod = si.Module(SPV_INDIRECT)
        cam_i = [cam[0], cam[1], cam[2], float(rng.uniform(0.2, 2.0))]
        bind(mod, {0: TILES, 1: si.Texture2D(irr.view(np.float16), repeat=True), 2: si.Texture2D(dep.view(np.float16), repeat=True), 3: ddgi_block(u),
                   4: si.Texture2D(color, repeat=False), 5: si.Texture2D(normal, repeat=False), 6: si.Texture2D(pbrt, repeat=False), 7: [vec(cam_i)]})
        want = run_frag(mod, n).astype(np.float16).view(np.uint16)
        cache = np.zeros((n, 4), dtype=np.uint16)
        o.indirect_light(u, irr, dep, cache, None, texel, pos, normals, alb, pbr[:, 0].copy(), cam_i[3], np.float32(cam_i[:3]))
        if not np.array_equal(cache, want):
            bad += 1
            if verbose:
                print("MISMATCH indirect config", k, int((cache != want).sum()), "values", flush=True)
        texels += 2 * n
        k += 1
    if verbose:
        print("configs", k, "texels", texels, "mismatches", bad, "in", round(time.time() - t0), "s")
    return k, texels, bad


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, float(sys.argv[2]) if len(sys.argv) > 2 else 300.0)
