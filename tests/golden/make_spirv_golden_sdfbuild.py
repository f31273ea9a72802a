"""Golden vectors for the global-SDF build from the reference's SHIPPED SPIR-V binaries:

    python tests/golden/make_spirv_golden_sdfbuild.py        (build container only: needs /root/reference)

Executes Assets/shaders/spv/SDF/{SDFRasterizeModelNoRead,SDFRasterizeModel,GlobalSDFMipmap}.comp.spv with oracle/spirv/interp.py:
  * three synthetic mesh distance fields (sphere 16^3, box 12x16x20, box 16^3 rotated + translated), 3 box-filtered mips each,
    merged into a 2-cascade global SDF of 32^3 per cascade: cascade 0 (mesh mip 0) through the NoRead pipeline followed by the
    READ_DISTANCE pipeline on the same chunk, cascade 1 (mesh mip 1, texture x offset 32) through NoRead;
  * the mip: the 4x downsample of each cascade and two flood passes of cascade 1 (Mip -> Tmp -> Mip).
Only a subset of the 8x8x8 workgroups of each chunk is run (the interpreter is slow); untouched voxels keep the cleared value.
tests/test_spirv_golden.py::test_sdf_build_* replays the same dispatches through the C++ oracle and compares bit for bit.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import meshsdf  # noqa: E402
from oracle.spirv import interp as si  # noqa: E402

SPV = "/root/reference/Assets/shaders/spv/SDF"
F = np.float32
RES, CASC = 32, 2
GROUPS = [(gx, gy, gz) for gz in (0, 2) for gy in (1, 3) for gx in (0, 1, 2, 3)]  # 16 of the 64 workgroups of a chunk


def vec(a):
    return [F(x) for x in a]


def mat_cols(m16):
    return [vec(m16[c * 4:c * 4 + 4]) for c in range(4)]


def rot_y(deg, t):
    c, s = np.cos(np.radians(deg)), np.sin(np.radians(deg))
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    m[:3, 3] = t
    return m


def golden_meshes():
    tr = np.eye(4, dtype=np.float32)
    tr[:3, 3] = [-1.5, 0.25, 1.0]
    return [meshsdf.synthetic("sphere", (16, 16, 16), (1.2, 1.2, 1.2), 0.3, tr, "sphere"),
            meshsdf.synthetic("box", (12, 16, 20), (0.6, 1.0, 1.4), 0.25, rot_y(0.0, [1.6, -0.5, -0.8]), "box"),
            meshsdf.synthetic("box", (16, 16, 16), (0.9, 0.7, 0.8), 0.25, rot_y(31.0, [0.4, 1.1, 1.9]), "rotbox")]


def cascades():
    """(centre, half extent) per cascade: cascade 1 = 2.5x cascade 0, as the reference's cascadesDistanceScales."""
    return [((0.0, 0.0, 0.0), 4.0), ((0.0, 0.0, 0.0), 10.0)]


def object_records(objs):
    return [[[mat_cols(o.worldToVolume), mat_cols(o.volumeToWorld), vec(o.volumeToUVWMul), F(o.mipOffset), vec(o.volumeToUVWAdd), F(o.decodeMul),
              vec(o.volumeLocalBoundsExtent), F(o.decodeAdd)] for o in objs]]


def run_rasterize(read, sdf_bits, objs, meshes, cascade, chunk, ids):
    mod = si.Module(os.path.join(SPV, "SDFRasterizeModel.comp.spv" if read else "SDFRasterizeModelNoRead.comp.spv"))
    (centre, D) = cascades()[cascade]
    voxel = 2 * D / RES
    mul = [(2 * D) / RES] * 3
    add = [c - D + voxel * 0.5 for c in centre]
    ubo = [vec(np.float32(mul)), F(2 * D), vec(np.float32(add)), RES, cascade, vec([0, 0, 0])]
    tex = [si.Texture3DMips(m.levels) for m in meshes]
    bind = {0: ubo, 1: object_records(objs), 2: si.StorageImage3D(sdf_bits), 3: tex}
    for b, v in bind.items():
        gid = mod.global_by_binding(0, b)
        if gid is not None:
            mod.storage[gid] = [v]
    (pc,) = mod.global_by_storage(9)
    mod.storage[pc] = [[[int(c) & si.M32 for c in chunk], len(ids), [int(i) for i in ids] + [0] * (28 - len(ids))]]
    return si.dispatch(mod, GROUPS)


def run_mip(src_bits, dst_bits, out_res, res, scale, tex_off, mip_off, max_distance):
    mod = si.Module(os.path.join(SPV, "GlobalSDFMipmap.comp.spv"))
    for b, v in ((0, si.StorageImage3D(src_bits)), (1, si.StorageImage3D(dst_bits))):
        mod.storage[mod.global_by_binding(0, b)] = [v]
    (pc,) = mod.global_by_storage(9)
    mod.storage[pc] = [[res, scale, tex_off, mip_off, F(max_distance)]]
    g = out_res // 4
    return si.dispatch(mod, [(x, y, z) for z in range(g) for y in range(g) for x in range(g)])


def main():
    from oracle import binding as ob

    meshes = golden_meshes()
    one = np.float16(1.0).view(np.uint16)
    sdf = np.full((RES, RES, RES * CASC), one, dtype=np.uint16)
    stages = {}
    objs0 = [ob.sdf_object_data(m, 0) for m in meshes]
    objs1 = [ob.sdf_object_data(m, 1) for m in meshes]
    n = run_rasterize(False, sdf, objs0, meshes, 0, (0, 0, 0), [0, 1])
    stages["c0_noread"] = sdf.copy()
    n += run_rasterize(True, sdf, objs0, meshes, 0, (0, 0, 0), [2])
    stages["c0_read"] = sdf.copy()
    n += run_rasterize(False, sdf, objs1, meshes, 1, (0, 0, 0), [2, 0, 1])
    stages["c1_noread"] = sdf.copy()
    mres = RES // 4
    mip = np.full((mres, mres, mres * CASC), one, dtype=np.uint16)
    tmp = np.full((mres, mres, mres), one, dtype=np.uint16)
    for c, (_, D) in enumerate(cascades()):
        n += run_mip(sdf, mip, mres, RES, 4, c * RES, c * mres, 2 * D)
    stages["mip_down"] = mip.copy()
    D1 = cascades()[1][1]
    n += run_mip(mip, tmp, mres, mres, 1, 1 * mres, 0, 2 * D1)
    stages["flood_tmp"] = tmp.copy()
    n += run_mip(tmp, mip, mres, mres, 1, 0, 1 * mres, 2 * D1)
    stages["flood_mip"] = mip.copy()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spirv_golden_sdfbuild.npz")
    np.savez_compressed(out, groups=np.asarray(GROUPS, dtype=np.int32), **stages)
    print("wrote", out, "SPIR-V instructions executed:", n)
    touched = (stages["c1_noread"] != one).sum()
    print("voxels != 1.0:", int(touched), "min value", float(stages['c1_noread'].view(np.float16).min()))


if __name__ == "__main__":
    main()
