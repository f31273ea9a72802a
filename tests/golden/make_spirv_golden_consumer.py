"""Golden vectors for the consumer row from the reference's shipped SampleProbe.comp.spv:

    python tests/golden/make_spirv_golden_consumer.py

Inputs: the frame-1 atlases of tests/golden/spirv_golden.npz (themselves produced by the shipped blend/border binaries), a
synthetic 16x12 G-buffer (depth + octahedral normals) looking into the Cornell scene, camera position and viewProjInv.
Output: tests/golden/spirv_golden_consumer.npz with the INDIRECT_LIGHTING image written by the shader."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from luxgi_b200 import abi  # noqa: E402
from oracle.spirv import interp as si  # noqa: E402
from tests.golden.make_spirv_golden import ddgi_block, mat_cols, vec  # noqa: E402

SPV = "/root/reference/Assets/shaders/spv/DDGI/SampleProbe.comp.spv"
HERE = os.path.dirname(os.path.abspath(__file__))
F = np.float32


class OutImage:
    def __init__(self, h, w):
        self.a = np.zeros((h, w, 4), dtype=np.float32)

    def write(self, c, texel):
        y, x = si.s32(c[1]), si.s32(c[0])
        if 0 <= y < self.a.shape[0] and 0 <= x < self.a.shape[1]:
            self.a[y, x] = [float(t) for t in texel]


class FloatTexture(si.Texture2D):
    def __init__(self, data):
        super().__init__(np.asarray(data, dtype=np.float32), repeat=False)


def look_at_perspective(eye, target, up, fovy, aspect, near, far):
    f = target - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up); s /= np.linalg.norm(s)
    u = np.cross(s, f)
    view = np.eye(4)
    view[0, :3], view[1, :3], view[2, :3] = s, u, -f
    view[:3, 3] = -view[:3, :3] @ eye
    t = np.tan(fovy / 2)
    proj = np.zeros((4, 4))
    proj[0, 0] = 1 / (aspect * t); proj[1, 1] = 1 / t
    proj[2, 2] = far / (near - far); proj[2, 3] = -(far * near) / (far - near); proj[3, 2] = -1  # depth zero-to-one
    return proj @ view


def main():
    g = np.load(os.path.join(HERE, "spirv_golden.npz"))
    u = abi.DDGIUniform.from_buffer_copy(g["in_uniform"].tobytes())
    u.normalBias = 0.1  # IrradianceVolume default (DDGIRenderer.h:51); keeps the visibility term active
    irr, dep = g["f1_irradiance"], g["f1_depth"]
    W, H = 16, 12
    eye = np.array([0.3, 0.4, 3.9]); target = np.array([-0.2, -0.5, -1.0])
    vp = look_at_perspective(eye, target, np.array([0.0, 1.0, 0.0]), np.radians(70), W / H, 0.1, 50.0)
    vpi = np.linalg.inv(vp).astype(np.float32)
    rng = np.random.default_rng(7)
    depth = rng.uniform(0.90, 0.999, (H, W)).astype(np.float32)
    depth[0, :3] = 1.0  # sky pixels exercise the early-out
    nrm = np.zeros((H, W, 4), dtype=np.float32)
    nrm[..., :2] = rng.uniform(-1, 1, (H, W, 2))
    mod = si.Module(SPV)
    out = OutImage(H, W)
    bind = {0: out, 1: si.Texture2D(irr.view(np.float16), repeat=True), 2: si.Texture2D(dep.view(np.float16), repeat=True), 3: ddgi_block(u),
            4: FloatTexture(depth), 5: FloatTexture(nrm), 6: [vec([*eye, 1.0]), mat_cols(vpi.T.reshape(16))]}
    for b, v in bind.items():
        gid = mod.global_by_binding(0, b)
        assert gid is not None, b
        mod.storage[gid] = [v]
    n = si.dispatch(mod, [(0, 0, 0)])
    print("SampleProbe.comp.spv:", n, "SPIR-V instructions")
    np.savez_compressed(os.path.join(HERE, "spirv_golden_consumer.npz"), uniform=np.frombuffer(bytes(u), dtype=np.uint8), irradiance=irr, depth_atlas=dep,
                        g_depth=depth, g_normal=nrm, camera=np.array([*eye, 1.0], dtype=np.float32), view_proj_inv=vpi.T.reshape(16).copy(), out=out.a)
    print("non-zero pixels", int((out.a[..., 3] > 0).sum()), "mean rgb", out.a[..., :3].mean(axis=(0, 1)))


if __name__ == "__main__":
    main()
