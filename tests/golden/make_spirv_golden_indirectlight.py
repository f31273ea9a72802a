"""Golden vectors for the infinite-bounce refresh from the reference's SHIPPED SPIR-V:

    python tests/golden/make_spirv_golden_indirectlight.py        (build container only: needs /root/reference)

Executes Assets/shaders/spv/SDF/SDFAtlasIndirectLight.frag.spv with oracle/spirv/interp.py, one fragment invocation per test texel: 256 surface
points of the Cornell surface cache (position / oct-encoded normal / albedo / metallic in a 16x16 G-buffer sampled at texel centres) against
the frame-1 probe atlases of spirv_golden.npz (themselves written by the shipped blend / border binaries), cameraPos.w = the bounce intensity.
Stores the G-buffer values and the shader's outColor per texel; tests/test_spirv_golden.py replays them through the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import abi  # noqa: E402
from oracle.spirv import interp as si  # noqa: E402
from tests.golden import make_spirv_golden_directlight as dl  # noqa: E402
from tests.golden.make_spirv_golden import ddgi_block, vec  # noqa: E402

SPV = "/root/reference/Assets/shaders/spv/SDF/SDFAtlasIndirectLight.frag.spv"
HERE = os.path.dirname(os.path.abspath(__file__))
F = np.float32
G = dl.G
CAMERA = [0.3, 0.2, 4.5, 1.2]  # xyz + intensity (the shipped scene's 1.2)


def main():
    g = np.load(os.path.join(HERE, "spirv_golden.npz"))
    u = abi.DDGIUniform.from_buffer_copy(g["in_uniform"].tobytes())
    u.normalBias = 0.1
    irr, dep = g["f1_irradiance"], g["f1_depth"]
    sc = dl.golden_scene()
    pos, octn, alb, pbr = dl.gbuffer(sc)
    alb = alb.copy()
    alb[::7] = [0.95, 0.99, 0.2]  # exercises min(albedo, 0.9)
    color = np.concatenate([alb, pos[:, :1]], -1).reshape(G, G, 4)
    normal = np.concatenate([octn, pos[:, 1:]], -1).reshape(G, G, 4)
    pbrt = np.concatenate([pbr, np.zeros((G * G, 2), np.float32)], -1).reshape(G, G, 4)
    tiles = [[[vec([0, 0, 1, 1]), [vec([1, 0, 0, 0]), vec([0, 1, 0, 0]), vec([0, 0, 1, 0]), vec([0, 0, 0, 1])], vec([1, 1, 1, 1])]]]
    mod = si.Module(SPV)
    bind = {0: tiles, 1: si.Texture2D(irr.view(np.float16), repeat=True), 2: si.Texture2D(dep.view(np.float16), repeat=True), 3: ddgi_block(u),
            4: si.Texture2D(color, repeat=False), 5: si.Texture2D(normal, repeat=False), 6: si.Texture2D(pbrt, repeat=False), 7: [vec(CAMERA)]}
    for b, v in bind.items():
        gid = mod.global_by_binding(0, b)
        if gid is not None:
            mod.storage[gid] = [v]
    gids = {mod.names.get(x): x for x in mod.globals_}
    res = np.zeros((G * G, 4), dtype=np.float32)
    total = 0
    for k in range(G * G):
        inv = si.Invocation(mod, {}, {})
        inv.g[gids["inTileUV"]] = si.Ptr([vec([((k % G) + 0.5) / G, ((k // G) + 0.5) / G])])
        inv.g[gids["inTileAddress"]] = si.Ptr([0])
        inv.g[gids["inPosition"]] = si.Ptr([vec([0, 0, 0, 1])])
        for _ in inv.run():
            pass
        res[k] = inv.g[gids["outColor"]].load()
        total += inv.count
    path = os.path.join(HERE, "spirv_golden_indirectlight.npz")
    np.savez_compressed(path, uniform=np.frombuffer(bytes(u), dtype=np.uint8), irradiance=irr, depth_atlas=dep, pos=pos, oct_normal=octn, albedo=alb,
                        metallic=pbr[:, 0].copy(), camera=np.float32(CAMERA), out=res)
    print("wrote", path, "SPIR-V instructions executed:", total, "mean rgb", res[:, :3].mean(0), "nonzero", int((res[:, :3].sum(1) > 0).sum()))


if __name__ == "__main__":
    main()
