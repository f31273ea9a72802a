"""The C++ oracle against golden vectors produced by EXECUTING THE REFERENCE'S SHIPPED SPIR-V
(Assets/shaders/spv/DDGI/*.comp.spv) with oracle/spirv/interp.py — see tests/golden/make_spirv_golden.py.

What this pins: operation order, constants, control flow, indexing and storage formats of all five shaders.
* trace (GISDFRays.comp.spv): ray buffers must match BIT FOR BIT, and so must the number of SDF / mip taps;
* border: bit for bit;
* blend: bit for bit in the oracle's `unfused` mode (the literal two-rounding arithmetic of the binaries), and within the
  north-star tolerance (1e-3 rel / 1e-4 abs) in the contract's FMA mode that the CUDA engine implements.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from luxgi_b200 import abi, scenes
from tests.util import compare_atlas

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spirv_golden.npz")


@pytest.fixture(scope="module")
def golden():
    g = np.load(PATH)
    u = abi.DDGIUniform.from_buffer_copy(g["in_uniform"].tobytes())
    sd = abi.GlobalSDFData.from_buffer_copy(g["in_sdf_data"].tobytes())
    ad = abi.GlobalSurfaceAtlasData.from_buffer_copy(g["in_atlas_data"].tobytes())
    sc = scenes.Scene("golden", u, sd, torch.from_numpy(g["in_sdf"].view(np.float16).copy()), torch.from_numpy(g["in_mip"].view(np.float16).copy()),
                      atlas_data=ad, chunks=g["in_chunks"], cull=g["in_cull"], objects=g["in_objects"].view(abi.OBJECT_DTYPE).copy(),
                      tiles=g["in_tiles"].view(abi.TILE_DTYPE).copy(), light=torch.from_numpy(g["in_light"].view(np.float16).copy()),
                      depth=torch.from_numpy(g["in_depth"].copy()), sky_face=1, sky=g["in_sky"].view(np.float16).copy())
    return g, sc


def test_trace_matches_shipped_spirv_bit_for_bit(oracle, golden):
    g, sc = golden
    osc = oracle.OracleScene(sc)
    for f in range(2):
        rad, dd, _, cn = osc.trace(g[f"f{f}_rotation"])
        assert np.array_equal(dd, g[f"f{f}_direction_distance"]), f"frame {f}: direction / hit distance differ"
        assert np.array_equal(rad, g[f"f{f}_radiance"]), f"frame {f}: radiance differs"
        assert cn["texTaps"] == int(g[f"f{f}_tex_taps"]) and cn["mipTaps"] == int(g[f"f{f}_mip_taps"])
        hits = (dd.view(np.float16)[..., 3] < 60000).sum()
        assert hits > 0.5 * dd.shape[0] * dd.shape[1]  # the vectors exercise the surface-cache path, not just misses


@pytest.mark.parametrize("unfused", [True, False])
def test_blend_and_border_match_shipped_spirv(oracle, golden, unfused):
    g, sc = golden
    u = sc.uniform
    oracle.set_unfused(unfused)
    try:
        irr = [oracle.new_atlases(u)[0] for _ in range(2)]
        dep = [oracle.new_atlases(u)[1] for _ in range(2)]
        ping = 0
        for f in range(2):
            w = 1 - ping
            # previous atlases come from the golden run, so frame 1 is checked independently of frame 0
            prev_i = g[f"f{f - 1}_irradiance"] if f else irr[ping]
            prev_d = g[f"f{f - 1}_depth"] if f else dep[ping]
            oracle.blend(u, g[f"f{f}_radiance"], g[f"f{f}_direction_distance"], prev_i, prev_d, irr[w], dep[w], first_frame=(f == 0), naive=True)
            for name, got, want in (("irradiance", irr[w], g[f"f{f}_irradiance_interior"]), ("depth", dep[w], g[f"f{f}_depth_interior"])):
                rep = compare_atlas(name, got, want)
                print(f, unfused, rep)
                assert rep["out_of_tolerance"] == 0, rep
                if unfused:
                    assert rep["mismatched_bits"] == 0, rep
                else:
                    assert rep["max_ulp16"] <= 1, rep
            # border on the golden interiors must reproduce the golden bordered atlases exactly
            bi, bd = g[f"f{f}_irradiance_interior"].copy(), g[f"f{f}_depth_interior"].copy()
            oracle.border(u, bi, bd)
            assert np.array_equal(bi, g[f"f{f}_irradiance"]) and np.array_equal(bd, g[f"f{f}_depth"])
            ping = w
    finally:
        oracle.set_unfused(False)


def test_hoisted_blend_equals_golden_too(oracle, golden):
    g, sc = golden
    u = sc.uniform
    oracle.set_unfused(True)
    try:
        irr, dep = oracle.new_atlases(u)
        oracle.blend(u, g["f1_radiance"], g["f1_direction_distance"], g["f0_irradiance"], g["f0_depth"], irr, dep, first_frame=False, naive=False)
        assert np.array_equal(irr, g["f1_irradiance_interior"]) and np.array_equal(dep, g["f1_depth_interior"])
    finally:
        oracle.set_unfused(False)


def test_sample_probe_matches_shipped_spirv(oracle):
    """Consumer row: the oracle's SampleProbe / sampleIrradiance restatement against the image written by the
    reference's shipped SampleProbe.comp.spv (tests/golden/make_spirv_golden_consumer.py).  Bit for bit: every operation of
    this shader is either pinned by the binary or fixed by the numerics contract (pow through binary64, bilinear atlas taps)."""
    g = np.load(os.path.join(os.path.dirname(PATH), "spirv_golden_consumer.npz"))
    u = abi.DDGIUniform.from_buffer_copy(g["uniform"].tobytes())
    got = oracle.sample_probe(u, g["irradiance"], g["depth_atlas"], g["g_depth"], g["g_normal"], g["camera"], g["view_proj_inv"])
    want = g["out"]
    assert (want[..., 3] > 0).sum() > 150 and np.isfinite(want).all()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"max abs diff {np.abs(got - want).max()}"


def test_sdf_build_matches_shipped_spirv(oracle):
    """Row f3: the oracle's restatement of SDFRasterizeModel(.NoRead) and GlobalSDFMipmap against outputs of the reference's shipped
    SPIR-V (tests/golden/make_spirv_golden_sdfbuild.py): two cascades (mesh mip 0 / 1, x offset), NoRead then READ_DISTANCE on the same
    chunk, the 4x min-downsample and two flood passes.  Bit for bit on every voxel the golden run dispatched."""
    from tests.golden import make_spirv_golden_sdfbuild as g

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "spirv_golden_sdfbuild.npz"))
    meshes = g.golden_meshes()
    RES, CASC = g.RES, g.CASC
    one = np.float16(1.0).view(np.uint16)
    mask = np.zeros((RES, RES, RES), dtype=bool)
    for gx, gy, gz in gold["groups"]:
        mask[gz * 8:gz * 8 + 8, gy * 8:gy * 8 + 8, gx * 8:gx * 8 + 8] = True
    sdf = np.full((RES, RES, RES * CASC), one, dtype=np.uint16)
    objs = [[oracle.sdf_object_data(m, c) for m in meshes] for c in range(CASC)]
    (c0, D0), (c1, D1) = g.cascades()

    def check(stage, cascade):
        want, got = gold[stage][:, :, cascade * RES:(cascade + 1) * RES], sdf[:, :, cascade * RES:(cascade + 1) * RES]
        assert np.array_equal(got[mask], want[mask]), f"{stage}: {(got[mask] != want[mask]).sum()} of {mask.sum()} voxels differ"
        got[~mask] = want[~mask]  # voxels the golden run did not dispatch keep its (cleared) value for the next stage

    oracle.sdf_rasterize_chunk(sdf, objs[0], meshes, 0, c0, D0, RES, 0, (0, 0, 0), [0, 1], read=False)
    check("c0_noread", 0)
    oracle.sdf_rasterize_chunk(sdf, objs[0], meshes, 0, c0, D0, RES, 0, (0, 0, 0), [2], read=True)
    check("c0_read", 0)
    oracle.sdf_rasterize_chunk(sdf, objs[1], meshes, 1, c1, D1, RES, 1, (0, 0, 0), [2, 0, 1], read=False)
    check("c1_noread", 1)
    assert np.array_equal(sdf, gold["c1_noread"])
    assert (sdf != one).sum() > 10000 and sdf.view(np.float16).min() < 0  # the meshes really landed in the volume

    mres = RES // 4
    mip = np.full((mres, mres, mres * CASC), one, dtype=np.uint16)
    tmp = np.full((mres, mres, mres), one, dtype=np.uint16)
    for c, (_, D) in enumerate(g.cascades()):
        oracle.sdf_mip_pass(sdf, mip, mres, RES, 4, c * RES, c * mres, 2 * D)
    assert np.array_equal(mip, gold["mip_down"]), f"{(mip != gold['mip_down']).sum()} mip texels differ"
    oracle.sdf_mip_pass(mip, tmp, mres, mres, 1, mres, 0, 2 * D1)
    assert np.array_equal(tmp, gold["flood_tmp"])
    oracle.sdf_mip_pass(tmp, mip, mres, mres, 1, 0, mres, 2 * D1)
    assert np.array_equal(mip, gold["flood_mip"])


def test_surface_culling_matches_shipped_spirv(oracle):
    """Row f4 (first half): the oracle's restatement of SDFCulling.comp executed in the golden run's dispatch order reproduces the shipped
    binary's chunk and cull buffers word for word, including the list-overflow branch and the element-0 store of empty chunks."""
    from tests.golden import make_spirv_golden_culling as g

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "spirv_golden_culling.npz"))
    sc = g.golden_scene()
    for cap in gold["capacities"]:
        data = abi.GlobalSurfaceAtlasData.from_buffer_copy(bytes(sc.atlas_data))
        data.culledObjectsCapacity = int(cap)
        chunks, cull = oracle.surface_cull(data, sc.objects, order=gold["order"], emulate_slot0=True, capacity_words=8192)
        assert np.array_equal(chunks, gold[f"chunks_{cap}"]), f"capacity {cap}: {(chunks != gold[f'chunks_{cap}']).sum()} chunk words differ"
        assert np.array_equal(cull, gold[f"cull_{cap}"]), f"capacity {cap}: {(cull != gold[f'cull_{cap}']).sum()} cull words differ"
    assert (gold["chunks_4096"][1:] != 0).sum() > 500 > (gold["chunks_150"][1:] != 0).sum() > 0


def test_surface_direct_light_matches_shipped_spirv(oracle):
    """Row f4: the oracle's restatement of SDFDeferredLight.frag (fetchLight for the three light types, shadow ray with start bias 2, BRDF)
    against the shipped binary executed per fragment: the shading added to an empty RGBA16F light cache equals fp16(outColor) bit for bit."""
    from tests.golden import make_spirv_golden_directlight as g

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "spirv_golden_directlight.npz"))
    sc = g.golden_scene()
    n = len(gold["pos"])
    normals = oracle.octohedral_to_direction(gold["oct_normal"])
    texel = np.arange(n, dtype=np.uint32)
    lit = 0
    for name in g.LIGHTS:
        cache = np.zeros((n, 4), dtype=np.uint16)
        oracle.surface_direct_light(sc.sdf_data, sc.sdf, sc.mip, oracle.make_light(gold[f"light_{name}"]), gold["camera"], cache, texel, gold["pos"], normals,
                                    gold["albedo"], gold["pbr"])
        want = gold[f"out_{name}"].astype(np.float16).view(np.uint16)
        assert np.array_equal(cache, want), f"{name}: {(cache != want).sum()} of {cache.size} fp16 values differ"
        lit += int((gold[f"out_{name}"][:, :3].sum(1) > 0).sum())
    assert lit > 60  # the three lights really shade something


@pytest.mark.parametrize("k", [2, 4])
def test_cascaded_trace_matches_shipped_spirv(oracle, k):
    """The cascaded global SDF (the reference always runs 2 cascades): GISDFRays.comp.spv executed on the Cornell scene with 2 and 4 nested,
    off-centre cascades laid side by side in one volume (tests/golden/make_spirv_golden_cascades.py).  Ray buffers and tap counts of the
    oracle equal the shipped binary's bit for bit: cascade entry from outside, the hand-over to the next cascade, per-cascade voxel sizes in
    the thickness test, the hit bias and the surface threshold, and the filter bleed across cascade seams in x."""
    import zlib

    from tests.golden import make_spirv_golden_cascades as g

    gold = np.load(os.path.join(os.path.dirname(PATH), "spirv_golden_cascades.npz"))
    sc = g.golden_scene(k)
    assert np.uint32(zlib.crc32(sc.sdf.numpy().tobytes() + sc.mip.numpy().tobytes())) == gold[f"c{k}_crc"], "the procedural scene changed: regenerate the fixture"
    rad, dd, _, cn = oracle.OracleScene(sc).trace(gold[f"c{k}_rotation"])
    assert np.array_equal(dd, gold[f"c{k}_direction_distance"]), f"{(dd != gold[f'c{k}_direction_distance']).sum()} direction / distance values differ"
    assert np.array_equal(rad, gold[f"c{k}_radiance"]), f"{(rad != gold[f'c{k}_radiance']).sum()} radiance values differ"
    assert cn["texTaps"] == int(gold[f"c{k}_tex_taps"]) and cn["mipTaps"] == int(gold[f"c{k}_mip_taps"])
    # the fixture really crosses cascades: rays from the same origins hit in every cascade
    u = sc.uniform
    tr = np.zeros(abi.probe_count(u) * u.raysPerProbe, dtype=abi.SDF_TRACE_DTYPE)
    d = dd.view(np.float16)[..., :3].astype(np.float32).reshape(-1, 3)
    tr["worldDirection"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    ids = np.repeat(np.arange(abi.probe_count(u)), u.raysPerProbe)
    grid = np.stack([ids % u.probeCounts[0], (ids // u.probeCounts[0]) % u.probeCounts[1], ids // (u.probeCounts[0] * u.probeCounts[1])], -1)
    tr["worldPosition"] = (np.float32(list(u.startPosition)[:3]) + grid.astype(np.float32) * np.float32(list(u.step)[:3])).astype(np.float32)
    tr["maxDistance"], tr["stepScale"] = abi.GLOBAL_SDF_WORLD_SIZE, 1.0
    hits = oracle.trace_global_sdf(sc.sdf_data, sc.sdf, sc.mip, tr, 0.0)
    per_cascade = np.bincount(hits["hitCascade"][hits["hitTime"] >= 0], minlength=k)
    assert (per_cascade > 0).all(), per_cascade


def _screen_golden():
    from tests.golden import make_spirv_golden as base

    g = np.load(os.path.join(os.path.dirname(PATH), "spirv_golden_screen.npz"))
    sc = base.golden_scene()
    sc.uniform = abi.DDGIUniform.from_buffer_copy(g["uniform"].tobytes())
    return g, sc


def test_sdf_reflection_matches_shipped_spirv(oracle):
    """Row f4: the oracle's restatement of SDFReflection.comp (mirror / GGX-sampled / DDGI-approximated branches, reflection ray through the
    global SDF, surface cache sampled with normal = -R, skybox on a miss, blue-noise sampler) against the RGBA16F image written by the shipped
    binary (tests/golden/make_spirv_golden_screen.py), bit for bit, with approximateWithDDGI on and off; depth == 1 pixels stay untouched."""
    g, sc = _screen_golden()
    for approx in (1, 0):
        frames, trim, inten = g[f"refl_params_{approx}"]
        push = abi.make_reflection_push(g["refl_camera"], g["refl_view_proj_inv"], int(frames), trim, inten, approx)
        out = np.full(g[f"refl_out_{approx}"].shape, 0x3555, dtype=np.uint16)
        oracle.sdf_reflection(sc, g["irradiance"], g["depth_atlas"], push, g["refl_depth"], g["refl_normal"], g["refl_pbr"], g["sobol"], g["scrambling"], out)
        want = g[f"refl_out_{approx}"]
        assert np.array_equal(out, want), f"approximateWithDDGI={approx}: {(out != want).any(-1).sum()} of {out.shape[0] * out.shape[1]} pixels differ"
        assert (want[..., 0] == 0x3555).sum() == 3 and (want.view(np.float16)[..., :3].astype(np.float32).sum(-1) > 0).sum() > 150
    assert not np.array_equal(g["refl_out_0"], g["refl_out_1"])  # the DDGI branch really ran


def test_sdf_shadow_matches_shipped_spirv(oracle):
    """Row f4: the oracle's restatement of SDFShadow.comp (soft-shadow disk sample from the blue noise, shadow ray with start bias 0, one
    visibility bit per pixel of an 8x4 workgroup) against the R32UI words written by the shipped binary for a directional, a point and a spot
    light; the workgroup whose first pixel is sky keeps its previous word."""
    from tests.golden import make_spirv_golden_screen as gs

    g, sc = _screen_golden()
    h, w = g["shadow_depth"].shape
    seen = 0
    for name in gs.LIGHTS:
        out = np.full((h // 4, w // 8), 0xDEADBEEF, dtype=np.uint32)
        oracle.sdf_shadow(sc.sdf_data, sc.sdf, sc.mip, oracle.make_light(g[f"shadow_light_{name}"]), g["shadow_view_proj_inv"], int(g["shadow_frames"]),
                          float(g["shadow_bias"]), g["shadow_depth"], g["shadow_normal"], g["sobol"], g["scrambling"], out)
        want = g[f"shadow_out_{name}"]
        assert np.array_equal(out, want), f"{name}: {[hex(int(a)) for a in out.reshape(-1)]} != {[hex(int(a)) for a in want.reshape(-1)]}"
        assert want[0, 0] == 0xDEADBEEF
        seen += sum(bin(int(v)).count("1") for v in want.reshape(-1)[1:])
    assert seen > 40  # visible and shadowed pixels both occur


def test_indirect_light_matches_shipped_spirv(oracle):
    """Row f1: the oracle's restatement of SDFAtlasIndirectLight.frag (min(albedo, 0.9), (albedo - albedo * metallic) / PI, intensity * diffuse *
    sampleIrradiance) against the shipped binary executed per fragment (tests/golden/make_spirv_golden_indirectlight.py): what the additive pass
    leaves in an empty RGBA16F light cache equals fp16(outColor) bit for bit, alpha included; a second pass on top adds again."""
    g = np.load(os.path.join(os.path.dirname(PATH), "spirv_golden_indirectlight.npz"))
    u = abi.DDGIUniform.from_buffer_copy(g["uniform"].tobytes())
    n = len(g["pos"])
    normals = oracle.octohedral_to_direction(g["oct_normal"])
    texel = np.arange(n, dtype=np.uint32)
    cache = np.zeros((n, 4), dtype=np.uint16)
    cam = g["camera"]
    oracle.indirect_light(u, g["irradiance"], g["depth_atlas"], cache, None, texel, g["pos"], normals, g["albedo"], g["metallic"], float(cam[3]), cam[:3])
    want = g["out"].astype(np.float16).view(np.uint16)
    assert np.array_equal(cache, want), f"{(cache != want).sum()} of {cache.size} fp16 values differ"
    assert (g["out"][:, :3].sum(1) > 0).sum() > 150 and (g["albedo"] > 0.9).any()
    # with a base atlas: listed texels = base + indirect (alpha = base alpha + 1)
    base = np.random.default_rng(1).uniform(0, 2, (n, 4)).astype(np.float16)
    cache2 = np.zeros((n, 4), dtype=np.uint16)
    oracle.indirect_light(u, g["irradiance"], g["depth_atlas"], cache2, base.view(np.uint16), texel, g["pos"], normals, g["albedo"], g["metallic"], float(cam[3]), cam[:3])
    want2 = (base.astype(np.float32) + g["out"].astype(np.float32)).astype(np.float16)
    d = np.abs(cache2.view(np.float16).astype(np.float32) - want2.astype(np.float32))
    assert (d <= 2e-3 * np.abs(want2.astype(np.float32)) + 1e-6).all()  # one fp32 add then one fp16 rounding, against fp16(out) added in fp32


def test_open_scene_trace_matches_shipped_spirv(oracle):
    """The miss / sky path of GISDFRays.comp.spv: a reduced open city under a 4x4-texel cube sky (71 % of the rays leave the cascade and take
    a bilinear sky texel; long open-space steps take the `stepDistance = chunkSizeDistance` branch; hits see emissive window bands).  Ray
    buffers and tap counts of the oracle equal the shipped binary's bit for bit."""
    import zlib

    from tests.golden import make_spirv_golden_cascades as g

    gold = np.load(os.path.join(os.path.dirname(PATH), "spirv_golden_cascades.npz"))
    sc = g.open_scene()
    assert np.uint32(zlib.crc32(sc.sdf.numpy().tobytes() + sc.mip.numpy().tobytes())) == gold["open_crc"], "the procedural scene changed: regenerate the fixture"
    assert np.uint32(zlib.crc32(sc.light.numpy().tobytes() + sc.depth.numpy().tobytes())) == gold["open_light_crc"], "the procedural surface cache changed"
    rad, dd, _, cn = oracle.OracleScene(sc).trace(gold["open_rotation"])
    assert np.array_equal(dd, gold["open_direction_distance"]) and np.array_equal(rad, gold["open_radiance"])
    assert cn["texTaps"] == int(gold["open_tex_taps"]) and cn["mipTaps"] == int(gold["open_mip_taps"])
    miss = dd.view(np.float16)[..., 3] >= 60000
    assert 0.5 < miss.mean() < 0.9 and len(np.unique(rad[miss][:, :3], axis=0)) > 20  # distinct bilinear sky values (every probe shares the 32 directions)


@pytest.mark.skipif(not os.path.exists("/root/reference/Assets/shaders/spv/DDGI/GISDFRays.comp.spv"), reason="needs the reference's shipped SPIR-V (build container only)")
def test_fuzz_oracle_vs_shipped_spirv(oracle):
    """A bounded slice of tests/golden/fuzz_oracle_vs_spirv.py: random scenes (1-4 cascades, open cities with sky, probes inside geometry and on
    cascade faces) and rotations, the oracle's ray buffers and tap counts against the shipped GISDFRays.comp.spv executed live.  The full run
    (571 configurations, 73 088 rays) found 0 mismatches."""
    from tests.golden import fuzz_oracle_vs_spirv as fz

    configs, rays, bad = fz.run(seed=7, seconds=60.0, max_configs=12, verbose=False)
    assert configs == 12 and rays > 1000 and bad == 0


@pytest.mark.skipif(not os.path.exists("/root/reference/Assets/shaders/spv/DDGI/DepthProbeUpdate.comp.spv"), reason="needs the reference's shipped SPIR-V (build container only)")
def test_fuzz_blend_vs_shipped_spirv(oracle):
    """A bounded slice of tests/golden/fuzz_blend_vs_spirv.py: random volume parameters, ray buffers (zero / large radiance, misses, weights under the
    gate), previous atlases, first and later frames against the shipped {Irradiance,Depth}{ProbeUpdate,BorderUpdate}.comp.spv executed live: bit for bit
    in `unfused` mode (literal and hoisted formulation), within 1 fp16 ulp in the contract's FMA mode, borders bit for bit."""
    from tests.golden import fuzz_blend_vs_spirv as fz

    configs, values, bad = fz.run(seed=11, seconds=120.0, max_configs=3, verbose=False)
    assert configs == 3 and values > 3000 and bad == 0


@pytest.mark.skipif(not os.path.exists("/root/reference/Assets/shaders/spv/SDF/SDFReflection.comp.spv"), reason="needs the reference's shipped SPIR-V (build container only)")
def test_fuzz_screen_shaders_vs_shipped_spirv(oracle):
    """A bounded slice of tests/golden/fuzz_screen_vs_spirv.py: SDFReflection / SDFShadow executed live on random G-buffers (normals at the +z pole,
    roughness on the 0.05 / 0.45 thresholds), noise textures, frame numbers and lights; the oracle's image / words are bit-identical."""
    from tests.golden import fuzz_screen_vs_spirv as fz

    configs, px, bad = fz.run(seed=5, seconds=60.0, max_configs=3, verbose=False)
    assert configs == 3 and bad == 0


@pytest.mark.skipif(not os.path.exists("/root/reference/Assets/shaders/spv/SDF/SDFDeferredLight.frag.spv"), reason="needs the reference's shipped SPIR-V (build container only)")
def test_fuzz_surface_lighting_vs_shipped_spirv(oracle):
    """A bounded slice of tests/golden/fuzz_light_vs_spirv.py: SDFDeferredLight.frag / SDFAtlasIndirectLight.frag executed live per fragment with random
    lights (all three types), cameras, shadow biases and G-buffer values; the oracle's light-cache values are bit-identical."""
    from tests.golden import fuzz_light_vs_spirv as fz

    configs, texels, bad = fz.run(seed=9, seconds=60.0, max_configs=6, verbose=False)
    assert configs == 6 and bad == 0


@pytest.mark.skipif(not os.path.exists("/root/reference/Assets/shaders/spv/DDGI/SampleProbe.comp.spv"), reason="needs the reference's shipped SPIR-V (build container only)")
def test_fuzz_sample_probe_vs_shipped_spirv(oracle):
    """A bounded slice of tests/golden/fuzz_consumer_vs_spirv.py: SampleProbe.comp.spv executed live with a random camera (possibly outside the probe
    volume), random depths / normals / normalBias and atlases of random fp16 values; the oracle's RGBA32F image is bit-identical."""
    from tests.golden import fuzz_consumer_vs_spirv as fz

    configs, px, bad = fz.run(seed=4, seconds=60.0, max_configs=1, verbose=False)
    assert configs == 1 and px == 256 and bad == 0


@pytest.mark.skipif(not os.path.exists("/root/reference/Assets/shaders/spv/SDF/SDFCulling.comp.spv"), reason="needs the reference's shipped SPIR-V (build container only)")
def test_fuzz_culling_vs_shipped_spirv(oracle):
    """A bounded slice of tests/golden/fuzz_culling_vs_spirv.py: SDFCulling.comp.spv executed live on random object buffers, chunk sizes, capacities and
    workgroup subsets; the oracle in the same execution order reproduces chunk and cull buffers word for word."""
    from tests.golden import fuzz_culling_vs_spirv as fz

    configs, lists, bad = fz.run(seed=3, seconds=60.0, max_configs=25, verbose=False)
    assert configs == 25 and lists > 200 and bad == 0


@pytest.mark.skipif(not os.path.exists("/root/reference/Assets/shaders/spv/SDF/SDFRasterizeModel.comp.spv"), reason="needs the reference's shipped SPIR-V (build container only)")
def test_fuzz_sdf_build_vs_shipped_spirv(oracle):
    """A bounded slice of tests/golden/fuzz_sdfbuild_vs_spirv.py: SDFRasterizeModelNoRead / SDFRasterizeModel / GlobalSDFMipmap executed live on random mesh
    distance fields, cascades, model splits, workgroups and mip inputs; every voxel the binaries wrote is reproduced bit for bit."""
    from tests.golden import fuzz_sdfbuild_vs_spirv as fz

    configs, touched, bad = fz.run(seed=6, seconds=60.0, max_configs=5, verbose=False)
    assert configs == 5 and touched > 1000 and bad == 0
