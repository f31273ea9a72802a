"""BASELINE configs[1]/[2] input ("dark-room-emissive scene SDF (Assets/sdf) 128^3"): built in the build container, where
/root/reference exists, and committed because the GPU box has no reference tree:

    python tests/golden/make_c2_dark_room.py

Reads the 100 MeshDistanceField records + transforms of Assets/dark-room-emissive.scene and their baked Assets/sdf/*.sdf volumes
(luxgi_b200/meshsdf.py), merges them with the ORACLE's restatement of the reference's global-SDF build (oracle_sdf_build: chunk
lists incl. the 28-model overflow behaviour, SDFRasterizeModel, GlobalSDFMipmap + flood) into one cascade of 128^3 voxels, half
extent 88 (voxel 1.375, 44-unit chunks), centred at the origin as the reference does (camera-independent viewPosition = 0,
GlobalDistanceField.cpp:640,677), and writes tests/golden/c2_dark_room.npz:
    sdf, mip          uint16 fp16 bits [128][128][128], [32][32][32]
    box_center/half   world AABB of every mesh field (the surface-cache fixture builds one OBB object per mesh from these)
    names, emissive   mesh names; emissive = name contains "Light" or "Emissive"
    stats             chunks dispatched, model references, references dropped by the overflow behaviour, chunks out of range
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import meshsdf, scenes  # noqa: E402
from oracle import binding as ob  # noqa: E402

ASSETS = "/root/reference/Assets"
D, RES = 88.0, 128


def main():
    meshes = meshsdf.load_scene_meshes(os.path.join(ASSETS, "dark-room-emissive.scene"), ASSETS)
    data = scenes.make_sdf_data((0.0, 0.0, 0.0), D, RES)
    sdf, mip, stats = ob.sdf_build(data, meshes, 0.0)
    centers, halves = [], []
    for m in meshes:
        c, e = (m.aabb_min + m.aabb_max) * 0.5, (m.aabb_max - m.aabb_min) * 0.5
        centers.append(m.world[:3, :3] @ c + m.world[:3, 3])
        halves.append(np.abs(m.world[:3, :3]) @ e)
    names = [m.name.rsplit("_", 1)[0] for m in meshes]
    emissive = np.array([("Light" in n) or ("Emissive" in n) for n in names])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c2_dark_room.npz")
    np.savez_compressed(out, sdf=sdf, mip=mip, half_extent=np.float32(D), resolution=np.int32(RES), box_center=np.float32(centers),
                        box_half=np.float32(halves), names=np.array(names), emissive=emissive,
                        stats=np.array([stats[k] for k in ("chunks", "models", "dropped_by_overflow", "chunks_out_of_range")], dtype=np.int32))
    f = sdf.view(np.float16).astype(np.float32)
    print("wrote", out, os.path.getsize(out), "bytes;", stats, "| occupied", float((f < 1).mean()), "inside", float((f <= 0).mean()),
          "| emissive meshes:", int(emissive.sum()), sorted(set(np.array(names)[emissive]))[:6])


if __name__ == "__main__":
    main()
