"""Golden vectors for the surface-cache direct light from the reference's SHIPPED SPIR-V:

    python tests/golden/make_spirv_golden_directlight.py        (build container only: needs /root/reference)

Executes Assets/shaders/spv/SDF/SDFDeferredLight.frag.spv with oracle/spirv/interp.py, one fragment invocation per test texel: 256 surface
points of the Cornell surface cache (position / oct-encoded normal / albedo / metallic, roughness in a 16x16 G-buffer, sampled at texel
centres so that the bilinear fetch returns the texel itself), lit by a directional, a point and a spot light, each with a shadow ray through
the 32^3 global SDF (tracyGlobalSDF with start bias 2).  Stores the G-buffer values, the lights and the shader's outColor per texel.
tests/test_spirv_golden.py::test_surface_direct_light_matches_shipped_spirv replays them through the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import scenes  # noqa: E402
from oracle.spirv import interp as si  # noqa: E402

SPV = "/root/reference/Assets/shaders/spv/SDF/SDFDeferredLight.frag.spv"
F = np.float32
G = 16  # G-buffer side
LIGHTS = {  # color, position, direction, intensity, radius, type, angle
    "directional": ([1.0, 0.95, 0.9, 1.0], [0, 0, 0, 1], [-0.35, -0.8, -0.48, 0.0], 2.0, 0.0, 0.0, 0.0),
    "point": ([1.0, 0.8, 0.6, 1.0], [0.5, 3.5, -0.7, 1.0], [0, 0, 0, 0], 3.0, 40.0, 2.0, 0.0),
    "spot": ([0.6, 0.8, 1.0, 1.0], [-1.0, 4.0, 1.0, 1.0], [-0.19611613, 0.98058068, 0.0, 0.0], 4.0, 60.0, 1.0, 0.6),  # direction as the shader uses it: surface -> light
}
CAMERA = [0.3, 0.2, 4.5, 0.08]  # xyz + shadow bias


def vec(a):
    return [F(x) for x in a]


def golden_scene():
    return scenes.cornell_scene(res=32, counts=(2, 2, 2), rays=32, atlas_res=256)


def oct_encode(n):
    n = n / np.abs(n).sum(-1, keepdims=True)
    p = n[:, :2].copy()
    neg = n[:, 2] <= 0
    q = (1.0 - np.abs(p[:, ::-1])) * np.where(p >= 0, 1.0, -1.0)
    p[neg] = q[neg]
    return p.astype(np.float32)


def gbuffer(sc):
    gb = sc.meta["gbuffer"]
    rng = np.random.default_rng(3)
    pick = rng.choice(len(gb["texel"]), G * G, replace=False)
    pos, nrm, alb = gb["pos"][pick].astype(np.float32), gb["normal"][pick].astype(np.float32), gb["albedo"][pick].astype(np.float32)
    nrm = nrm + rng.normal(scale=0.15, size=nrm.shape).astype(np.float32)  # not only axis-aligned normals
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    pbr = np.stack([rng.choice([0.0, 0.3, 1.0], G * G), rng.uniform(0.05, 1.0, G * G)], -1).astype(np.float32)
    return pos, oct_encode(nrm), alb, pbr


def main():
    sc = golden_scene()
    pos, octn, alb, pbr = gbuffer(sc)
    color = np.concatenate([alb, pos[:, :1]], -1).reshape(G, G, 4)
    normal = np.concatenate([octn, pos[:, 1:]], -1).reshape(G, G, 4)
    pbrt = np.concatenate([pbr, np.zeros((G * G, 2), np.float32)], -1).reshape(G, G, 4)
    d = sc.sdf_data
    sdf_block = [[vec(d.cascadePosDistance[i]) for i in range(4)], vec(d.cascadeVoxelSize), d.cascadesCount, F(d.resolution), F(d.nearPlane), F(d.farPlane)]
    tiles = [[[vec([0, 0, 1, 1]), [vec([1, 0, 0, 0]), vec([0, 1, 0, 0]), vec([0, 0, 1, 0]), vec([0, 0, 0, 1])], vec([1, 1, 1, 1])]]]
    out = {"pos": pos, "oct_normal": octn, "albedo": alb, "pbr": pbr, "camera": np.float32(CAMERA)}
    total = 0
    for name, (col, lpos, ldir, inten, radius, ltype, angle) in LIGHTS.items():
        mod = si.Module(SPV)
        light = [vec(col), vec(lpos), vec(ldir), F(inten), F(radius), F(ltype), F(angle)]
        bind = {0: tiles, 1: si.Texture2D(color, repeat=False), 2: si.Texture2D(normal, repeat=False), 4: si.Texture2D(pbrt, repeat=False),
                6: [vec(CAMERA), light, sdf_block], 7: si.Texture3D(sc.mip.numpy()), 8: si.Texture3D(sc.sdf.numpy())}
        for b, v in bind.items():
            gid = mod.global_by_binding(0, b)
            if gid is not None:
                mod.storage[gid] = [v]
        gids = {mod.names.get(g): g for g in mod.globals_}
        res = np.zeros((G * G, 4), dtype=np.float32)
        for k in range(G * G):
            inv = si.Invocation(mod, {}, {})
            inv.g[gids["inTileUV"]] = si.Ptr([vec([((k % G) + 0.5) / G, ((k // G) + 0.5) / G])])
            inv.g[gids["inTileAddress"]] = si.Ptr([0])
            inv.g[gids["inPosition"]] = si.Ptr([vec([0, 0, 0, 1])])
            for _ in inv.run():
                pass
            res[k] = inv.g[gids["outColor"]].load()
            total += inv.count
        out[f"out_{name}"] = res
        out[f"light_{name}"] = np.float32(list(col) + list(lpos) + list(ldir) + [inten, radius, ltype, angle])
        lit = (res[:, :3].sum(1) > 0).mean()
        print(name, "lit fraction", float(lit), "mean", res[:, :3].mean(0))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spirv_golden_directlight.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "SPIR-V instructions executed:", total)


if __name__ == "__main__":
    main()
