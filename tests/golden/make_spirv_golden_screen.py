"""Golden vectors for the screen-space tracyGlobalSDF users from the reference's SHIPPED SPIR-V:

    python tests/golden/make_spirv_golden_screen.py        (build container only: needs /root/reference)

Executes Assets/shaders/spv/SDF/SDFReflection.comp.spv (one 16x16 workgroup, approximateWithDDGI on and off) and SDFShadow.comp.spv (a 16x8
G-buffer = 2x2 workgroups of 8x4, a directional, a point and a spot light) with oracle/spirv/interp.py on the Cornell scene of
make_spirv_golden.py (SDF 32^3, 256^2 surface cache, six-colour cube sky) with the frame-1 atlases of spirv_golden.npz as the probe volume.
G-buffer: random depths (a few sky pixels, some far pixels whose rays start outside the cascade and reach the sky), random octahedral normals,
roughness from every branch of the reflection shader; the blue-noise textures are seeded random RGBA8 texels (the ABI takes raw texels).
Stores the inputs and the images the shaders wrote; tests/test_spirv_golden.py replays them through the oracle."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import abi  # noqa: E402
from oracle.spirv import interp as si  # noqa: E402
from tests.golden import make_spirv_golden as base  # noqa: E402
from tests.golden.make_spirv_golden import ddgi_block, mat_cols, vec  # noqa: E402
from tests.golden.make_spirv_golden_consumer import look_at_perspective  # noqa: E402

SPV = "/root/reference/Assets/shaders/spv/SDF"
HERE = os.path.dirname(os.path.abspath(__file__))
F = np.float32
LIGHTS = {  # color, position, direction (w = light radius for the soft shadow), intensity, radius, type, angle
    "directional": ([1.0, 0.95, 0.9, 1.0], [0, 0, 0, 1], [-0.35, -0.8, -0.48, 0.05], 2.0, 0.0, 0.0, 0.0),
    "point": ([1.0, 0.8, 0.6, 1.0], [0.5, 3.5, -0.7, 1.0], [0, 0, 0, 0.3], 3.0, 40.0, 2.0, 0.0),
    "spot": ([0.6, 0.8, 1.0, 1.0], [-1.0, 4.0, 1.0, 1.0], [-0.19611613, 0.98058068, 0.0, 0.2], 4.0, 60.0, 1.0, 0.6),
}


class UIntImage:  # r32ui storage image
    def __init__(self, h, w, fill):
        self.a = np.full((h, w), fill, dtype=np.uint32)

    def write(self, c, texel):
        y, x = si.s32(c[1]), si.s32(c[0])
        if 0 <= y < self.a.shape[0] and 0 <= x < self.a.shape[1]:
            self.a[y, x] = int(texel[0]) & si.M32


def unorm(tex_u8):
    return (tex_u8.astype(np.float32) / np.float32(255.0)).astype(np.float32)


def noise_textures():
    rng = np.random.default_rng(23)
    return rng.integers(0, 256, (1, 256, 4), dtype=np.uint8), rng.integers(0, 256, (128, 128, 4), dtype=np.uint8)


def camera(W, H):
    eye = np.array([0.3, 0.4, 3.9]); target = np.array([-0.2, -0.5, -1.0])
    vp = look_at_perspective(eye, target, np.array([0.0, 1.0, 0.0]), np.radians(70), W / H, 0.1, 50.0)
    return eye, np.linalg.inv(vp).astype(np.float32).T.reshape(16).copy()  # column-major


def gbuffer(W, H, seed):
    rng = np.random.default_rng(seed)
    depth = rng.uniform(0.90, 0.985, (H, W)).astype(np.float32)
    far = rng.random((H, W)) < 0.12
    depth[far] = rng.uniform(0.9975, 0.9995, int(far.sum())).astype(np.float32)  # beyond the cascade: rays that start outside and reach the sky
    depth[0, 0] = depth[1, 5] = depth[H - 1, W - 2] = 1.0  # sky pixels: early-out (and, for the shadow, a workgroup that stores nothing)
    nrm = np.zeros((H, W, 4), dtype=np.float32)
    nrm[..., :2] = rng.uniform(-1, 1, (H, W, 2))
    pbr = np.zeros((H, W, 4), dtype=np.float32)
    pbr[..., 1] = rng.choice(np.float32([0.01, 0.04, 0.2, 0.44, 0.5, 0.9]), (H, W))
    pbr[..., 0] = rng.uniform(0, 1, (H, W))
    return depth, nrm, pbr


def bind(mod, table):
    for b, v in table.items():
        gid = mod.global_by_binding(0, b)
        if gid is not None:
            mod.storage[gid] = [v]


def run_reflection(sc, u, irr, dep, depth, nrm, pbr, sobol, scr, eye, vpi, approximate, frames, trim, ddgi_intensity):
    mod = si.Module(os.path.join(SPV, "SDFReflection.comp.spv"))
    H, W = depth.shape
    out = np.full((H, W, 4), 0x3555, dtype=np.uint16)  # a recognisable previous content
    a, d = sc.atlas_data, sc.sdf_data
    atlas = [vec(a.cameraPos), F(a.chunkSize), a.culledObjectsCapacity, a.resolution, a.objectsCount, a.padding]
    sdf = [[vec(d.cascadePosDistance[i]) for i in range(4)], vec(d.cascadeVoxelSize), d.cascadesCount, F(d.resolution), F(d.nearPlane), F(d.farPlane)]
    tiles = [[[vec(t["extends"]), mat_cols(t["transform"]), vec(t["objectBounds"])] for t in sc.tiles]]
    objs = [[[vec(o["objectBounds"]), [int(x) for x in o["tileOffset"]], [0, 0], mat_cols(o["transform"]), vec(o["extends"])] for o in sc.objects]]
    bind(mod, {0: si.StorageImage(out), 1: si.Texture3D(sc.sdf.numpy()), 2: si.Texture3D(sc.mip.numpy()), 3: si.Texture2D(sc.light.numpy(), repeat=True),
               4: si.Texture2D(sc.depth.numpy(), repeat=False), 5: tiles, 6: objs, 7: [[int(x) for x in sc.chunks]], 8: [[int(x) for x in sc.cull]],
               9: si.TextureCube(sc.sky), 10: [atlas, sdf], 11: si.Texture2D(np.zeros((H, W, 4), np.float32), repeat=False),
               12: si.Texture2D(nrm, repeat=False), 13: si.Texture2D(depth, repeat=False), 14: si.Texture2D(pbr, repeat=False),
               15: si.Texture2D(unorm(sobol), repeat=False), 16: si.Texture2D(unorm(scr), repeat=False),
               17: si.Texture2D(irr.view(np.float16), repeat=True), 18: si.Texture2D(dep.view(np.float16), repeat=True), 19: ddgi_block(u)})
    (pc,) = mod.global_by_storage(9)
    mod.storage[pc] = [[F(0.0), F(trim), F(1.0), F(ddgi_intensity), 0, frames, 0, 1 if approximate else 0, vec([*eye, 1.0]), mat_cols(vpi)]]
    n = si.dispatch(mod, [(gx, gy, 0) for gy in range((H + 15) // 16) for gx in range((W + 15) // 16)])
    return out, n


def run_shadow(sc, light, depth, nrm, sobol, scr, vpi, frames, shadow_bias):
    mod = si.Module(os.path.join(SPV, "SDFShadow.comp.spv"))
    H, W = depth.shape
    out = UIntImage(H // 4, W // 8, 0xDEADBEEF)
    d = sc.sdf_data
    sdf = [[vec(d.cascadePosDistance[i]) for i in range(4)], vec(d.cascadeVoxelSize), d.cascadesCount, F(d.resolution), F(d.nearPlane), F(d.farPlane)]
    col, lpos, ldir, inten, radius, ltype, angle = light
    lt = [vec(col), vec(lpos), vec(ldir), F(inten), F(radius), F(ltype), F(angle)]
    bind(mod, {0: out, 1: si.Texture2D(nrm, repeat=False), 2: si.Texture2D(depth, repeat=False), 3: si.Texture3D(sc.sdf.numpy()), 4: si.Texture3D(sc.mip.numpy()),
               5: [lt, sdf, mat_cols(vpi), frames, F(shadow_bias)], 6: si.Texture2D(unorm(sobol), repeat=False), 7: si.Texture2D(unorm(scr), repeat=False)})
    n = si.dispatch(mod, [(gx, gy, 0) for gy in range(H // 4) for gx in range(W // 8)])
    return out.a, n


def main():
    sc = base.golden_scene()
    g = np.load(os.path.join(HERE, "spirv_golden.npz"))
    u = abi.DDGIUniform.from_buffer_copy(g["in_uniform"].tobytes())
    u.normalBias = 0.1
    irr, dep = g["f1_irradiance"], g["f1_depth"]
    sobol, scr = noise_textures()
    out = {"uniform": np.frombuffer(bytes(u), dtype=np.uint8), "irradiance": irr, "depth_atlas": dep, "sobol": sobol, "scrambling": scr}
    # reflection
    W, H = 16, 16
    eye, vpi = camera(W, H)
    depth, nrm, pbr = gbuffer(W, H, 31)
    out.update(refl_depth=depth, refl_normal=nrm, refl_pbr=pbr, refl_camera=np.float32([*eye, 1.0]), refl_view_proj_inv=vpi)
    for approx in (1, 0):
        frames, trim, inten = 5 + approx, 0.85, 1.3
        t0 = time.time()
        img, n = run_reflection(sc, u, irr, dep, depth, nrm, pbr, sobol, scr, eye, vpi, approx, frames, trim, inten)
        print(f"reflection approximateWithDDGI={approx}: {n} SPIR-V instructions, {time.time() - t0:.1f} s; written pixels {(img[..., 0] != 0x3555).sum()}", flush=True)
        out[f"refl_out_{approx}"] = img
        out[f"refl_params_{approx}"] = np.float32([frames, trim, inten])
    # shadow
    W, H = 16, 8
    eye, vpi = camera(W, H)
    depth, nrm, _ = gbuffer(W, H, 37)
    out.update(shadow_depth=depth, shadow_normal=nrm, shadow_view_proj_inv=vpi, shadow_frames=np.uint32(9), shadow_bias=np.float32(0.08))
    for name, light in LIGHTS.items():
        t0 = time.time()
        mask, n = run_shadow(sc, light, depth, nrm, sobol, scr, vpi, 9, 0.08)
        print(f"shadow {name}: {n} SPIR-V instructions, {time.time() - t0:.1f} s; masks {[hex(int(m)) for m in mask.reshape(-1)]}", flush=True)
        out[f"shadow_out_{name}"] = mask
        col, lpos, ldir, inten, radius, ltype, angle = light
        out[f"shadow_light_{name}"] = np.float32(list(col) + list(lpos) + list(ldir) + [inten, radius, ltype, angle])
    path = os.path.join(HERE, "spirv_golden_screen.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
