"""Parity of the CUDA engine (through the C ABI) against the CPU oracle on identical inputs.

north_star tolerances: hit distances within 1e-3 scene units, atlas texels within 1e-3 relative / 1e-4 absolute.
Under the numerics contract the engine is expected to be BIT-IDENTICAL to the oracle; the tests assert
the stated tolerances and additionally require that at most a handful of fp16 values differ at all."""
import os

import numpy as np
import pytest

from luxgi_b200 import abi, ddgi, scenes
from tests.util import compare_atlas, compare_hit_distance, f16

pytestmark = pytest.mark.gpu


def run_engine(sc, rots, flags=0, staged=False, rank=0, world=1):
    pipe = ddgi.DDGIPipeline(sc.uniform, flags=flags, rank=rank, world=world)
    pipe.set_scene(sc)
    for f, rot in enumerate(rots):
        if staged:  # the reference's three systems called one by one
            pipe.trace_rays(rot, num_frames=f)
            pipe.probe_update()
            pipe.border_update()
            pipe.end_frame()
        else:
            pipe.update(rot)
    pipe.synchronize()
    return pipe


def assert_rays_match(pipe, orc, max_flips=0):
    rad, dd = pipe.radiance, pipe.direction_distance
    rep = compare_hit_distance(dd, orc.dd)
    print("hit distance:", rep)
    assert rep["frac_within_tol"] >= 0.999, rep
    assert np.array_equal(dd[..., :3], orc.dd[..., :3]), "ray directions differ"
    r = compare_atlas("radiance", rad, orc.rad)
    print(r)
    assert r["out_of_tolerance"] <= max(max_flips, int(1e-3 * r["texels"])), r
    assert rep["mismatched_bits"] <= max_flips and r["mismatched_bits"] <= max_flips, (rep, r)


def assert_atlases_match(pipe, orc, max_flips=0):
    for name, got, want in (("irradiance", pipe.irradiance, orc.irradiance), ("depth", pipe.depth, orc.depth)):
        rep = compare_atlas(name, got, want)
        print(rep)
        assert rep["out_of_tolerance"] == 0, rep
        assert rep["mismatched_bits"] <= max_flips, rep


def test_c1_single_frame(oracle):
    """BASELINE configs[0]: Cornell 64^3, 8x8x8 probes, 64 rays, 1 frame."""
    sc = scenes.build("c1")
    rot = scenes.frame_rotation(0)
    orc = oracle.OraclePipeline(sc)
    orc.update(rot)
    pipe = run_engine(sc, [rot])
    assert_rays_match(pipe, orc)
    assert_atlases_match(pipe, orc)
    st = pipe.state()
    assert (st.frames, st.pingPong) == (1, 1) and st.kernelLaunches >= 6
    # outer pad rows/columns are never written
    irr = pipe.irradiance
    assert not irr[0].any() and not irr[-1].any() and not irr[:, 0].any() and not irr[:, -1].any()
    pipe.close()


def test_c1_eight_frames_hysteresis(oracle):
    sc = scenes.build("c1")
    rots = [scenes.frame_rotation(f) for f in range(8)]
    orc = oracle.OraclePipeline(sc)
    for r in rots:
        orc.update(r)
    pipe = run_engine(sc, rots)
    assert_rays_match(pipe, orc)
    assert_atlases_match(pipe, orc)
    assert pipe.state().frames == 8
    pipe.close()


def test_staged_systems_and_unfused_border_equal_fused(oracle):
    sc = scenes.cornell_scene(res=32, counts=(4, 4, 4), rays=96, atlas_res=256)  # R not a multiple of 32/64
    rots = [scenes.frame_rotation(f) for f in range(3)]
    orc = oracle.OraclePipeline(sc)
    for r in rots:
        orc.update(r)
    a = run_engine(sc, rots)
    b = run_engine(sc, rots, flags=abi.FLAG_UNFUSED_BORDER, staged=True)
    c = run_engine(sc, rots, staged=True)  # fused borders + the (idempotent) standalone border pass
    ai, ad = a.irradiance, a.depth
    for p in (a, b, c):
        assert_atlases_match(p, orc)
        assert np.array_equal(p.irradiance, ai) and np.array_equal(p.depth, ad)
        p.close()


def test_blend_stage_alone_on_oracle_rays(oracle):
    """Feed oracle-produced ray buffers through lux_ddgi_set_ray_buffers: isolates blend + border parity."""
    sc = scenes.build("c1")
    orc = oracle.OraclePipeline(sc)
    pipe = ddgi.DDGIPipeline(sc.uniform)
    for f in range(3):
        orc.update(scenes.frame_rotation(f))
        pipe.set_ray_buffers(orc.rad, orc.dd)
        pipe.probe_update()
        pipe.border_update()
        pipe.end_frame()
    assert_atlases_match(pipe, orc)
    pipe.close()


def test_blend_skips_gated_rays_even_when_their_radiance_is_infinite(oracle):
    """ProbeUpdate.glsl:93 skips a ray whose weight is below 1e-8; the engine stores such weights as exact zeros and multiplies.  An fp16 Inf
    radiance (an HDR sky or light cache above 65504 rounded by the RGBA16F store) must therefore not reach texels it is gated from: Inf * 0 is
    NaN and would stay in the atlas through the hysteresis.  Texels that DO see the ray become Inf in the oracle and the engine alike."""
    sc = scenes.build("c1")
    u = sc.uniform
    osc = oracle.OracleScene(sc)
    rot = scenes.frame_rotation(0)
    rad, dd, _, _ = osc.trace(rot)
    rad = rad.copy()
    rng = np.random.default_rng(3)
    for p_, r_ in zip(rng.integers(0, rad.shape[0], 40), rng.integers(0, rad.shape[1], 40)):
        rad[p_, r_, int(rng.integers(0, 3))] = 0x7c00  # +Inf
    irr = [oracle.new_atlases(u)[0] for _ in range(2)]
    dep = [oracle.new_atlases(u)[1] for _ in range(2)]
    pipe = ddgi.DDGIPipeline(u)
    for f in range(2):
        oracle.blend(u, rad, dd, irr[f % 2], dep[f % 2], irr[1 - f % 2], dep[1 - f % 2], first_frame=(f == 0))
        oracle.border(u, irr[1 - f % 2], dep[1 - f % 2])
        pipe.set_ray_buffers(rad, dd)
        pipe.probe_update()
        pipe.border_update()
        pipe.end_frame()
    want, got = irr[0], pipe.irradiance
    assert np.array_equal(got, want), f"{(got != want).sum()} irradiance values differ"
    assert np.array_equal(pipe.depth, dep[0])
    wf = f16(want)[..., :3]
    assert np.isinf(wf).any() and not np.isnan(wf).any() and np.isfinite(wf).mean() > 0.5  # some texels see the Inf rays, most do not, none is NaN
    pipe.close()


def test_uniform_field_known_answer_on_gpu():
    """SURVEY §7.3 closed form, no oracle involved: every ray returns L and d."""
    u = abi.make_uniform((0, 0, 0), (1, 1, 1), (4, 4, 2), 128, max_distance=6.0, gamma=5.0)
    sc = scenes.cornell_scene(res=32, counts=(4, 4, 2), rays=128, with_atlas=False)
    pipe = ddgi.DDGIPipeline(u)
    pipe.set_global_sdf(sc.sdf_data, sc.sdf, sc.mip)
    pipe.trace_rays(scenes.frame_rotation(1))  # only to obtain the frame's fp16 directions
    dd = pipe.direction_distance.copy()
    dd[..., 3] = np.float16(3.0).view(np.uint16)
    rad = np.zeros_like(dd)
    rad[..., :3] = np.array([0.5, 0.25, 2.0], dtype=np.float16).view(np.uint16)
    pipe.set_ray_buffers(rad, dd)
    pipe.probe_update()
    pipe.end_frame()
    irr, dep = f16(pipe.irradiance), f16(pipe.depth)
    S = 10
    blk = irr[2:10, 2:10]
    for c, L in enumerate((0.5, 0.25, 2.0)):
        assert np.allclose(blk[..., c], (L / 2) ** 0.2, rtol=2e-3)
    assert np.all(blk[..., 3] == 1.0)
    dq = 3.0 - 0.01
    assert np.allclose(dep[2:18, 2:18, 0], dq / 2, rtol=2e-3) and np.allclose(dep[2:18, 2:18, 1], dq * dq / 2, rtol=2e-3)
    # fused border = mirror rule
    t = irr[1:11, 1:11]
    assert np.array_equal(t[0, 1:-1], t[1, 1:-1][::-1]) and np.array_equal(t[1:-1, 0], t[1:-1, 1][::-1])
    assert np.array_equal(t[0, 0], t[-2, -2]) and np.array_equal(t[-1, -1], t[1, 1])
    pipe.close()


def test_city_with_sky_and_emissive(oracle):
    sc = scenes.build("city64")
    rots = [scenes.frame_rotation(f) for f in range(2)]
    orc = oracle.OraclePipeline(sc)
    for r in rots:
        orc.update(r)
    pipe = run_engine(sc, rots)
    assert_rays_match(pipe, orc)
    assert_atlases_match(pipe, orc)
    assert (f16(pipe.direction_distance)[..., 3] == 60000.0).mean() > 0.05  # sky rays exist
    pipe.close()


def test_axis_aligned_ray_and_no_atlas(oracle):
    import math
    import ctypes as C

    sc = scenes.cornell_scene(res=32, counts=(2, 2, 2), rays=64, with_atlas=False)
    out = (C.c_float * 3)()
    oracle.lib().oracle_spherical_fibonacci(7, 64, None, out)
    v = np.array(out[:], dtype=np.float64)
    tgt = np.array([1.0, 0.0, 0.0])
    axis = np.cross(v, tgt)
    rot = scenes.rotation_from_axis_angle(axis / np.linalg.norm(axis), math.atan2(np.linalg.norm(axis), float(v @ tgt)))
    orc = oracle.OraclePipeline(sc)
    orc.update(rot)
    pipe = run_engine(sc, [rot])
    assert_rays_match(pipe, orc)
    assert_atlases_match(pipe, orc)
    pipe.close()


def test_z_slab_shards_reassemble_the_full_volume(oracle):
    """Two shards (rank 0/1 of world 2) on one GPU write disjoint atlas rows whose union is the single-context result."""
    sc = scenes.build("c1")
    rots = [scenes.frame_rotation(f) for f in range(2)]
    full = run_engine(sc, rots)
    parts = [run_engine(sc, rots, rank=r, world=2) for r in range(2)]
    irr, dep = np.zeros_like(full.irradiance), np.zeros_like(full.depth)
    for p in parts:
        st = p.state()
        assert st.probeCount == full.probe_count // 2
        irr[st.irradianceRowBegin: st.irradianceRowBegin + st.irradianceRowCount] = p.irradiance[st.irradianceRowBegin: st.irradianceRowBegin + st.irradianceRowCount]
        dep[st.depthRowBegin: st.depthRowBegin + st.depthRowCount] = p.depth[st.depthRowBegin: st.depthRowBegin + st.depthRowCount]
        other = np.ones(irr.shape[0], dtype=bool)
        other[st.irradianceRowBegin: st.irradianceRowBegin + st.irradianceRowCount] = False
        assert not p.irradiance[other].any()  # a shard never touches rows it does not own
    assert np.array_equal(irr, full.irradiance) and np.array_equal(dep, full.depth)
    assert np.array_equal(np.concatenate([p.radiance for p in parts]), full.radiance)
    for p in parts + [full]:
        p.close()


def test_interleaved_layer_shards_reassemble_the_full_volume(oracle):
    """LUX_DDGI_FLAG_SHARD_INTERLEAVED: four shards on one GPU, rank g owning the z-layers g, g + 4, ... (the balanced multi-GPU layout), write
    disjoint atlas rows whose union is the single-context result; the packed own-row download returns exactly those rows; both blend forms."""
    sc = scenes.cornell_scene(res=32, counts=(5, 3, 8), rays=96, atlas_res=256)
    rots = [scenes.frame_rotation(f) for f in range(2)]
    full = run_engine(sc, rots)
    for blend, log2b in ((abi.FLAG_BLEND_LISTS, 0), (abi.FLAG_BLEND_TILES, 0), (abi.FLAG_BLEND_LISTS, 1)):  # single layers, blocks of 2 layers
        parts = [run_engine(sc, rots, flags=abi.flag_shard_blocks(log2b) | blend, rank=r, world=4) for r in range(4)]
        irr, dep = np.zeros_like(full.irradiance), np.zeros_like(full.depth)
        rad = np.zeros_like(full.radiance)
        for p in parts:
            st = p.state()
            assert st.layerStride == 4 and st.unitLayers == 1 << log2b and st.probeCount == full.probe_count // 4
            ri, rd = st.own_rows(8), st.own_rows(16)
            irr[ri], dep[rd] = p.irradiance[ri], p.depth[rd]
            other = np.ones(irr.shape[0], dtype=bool)
            other[ri] = False
            assert not p.irradiance[other].any()  # a shard never touches rows it does not own
            rad[st.own_probes()] = p.radiance
            import torch
            pin = torch.empty(len(ri) * irr.shape[1] * irr.shape[2] * 2, dtype=torch.uint8).pin_memory()
            p.download_shard_async_ptr(abi.BUF_IRRADIANCE, pin.data_ptr())
            p.wait_fence(p.download_fence())
            assert np.array_equal(pin.numpy().view(np.uint16).reshape(len(ri), irr.shape[1], irr.shape[2]), p.irradiance[ri])
        assert np.array_equal(irr, full.irradiance) and np.array_equal(dep, full.depth)
        assert np.array_equal(rad, full.radiance)
        for p in parts:
            p.close()
    full.close()


def test_restore_resumes_bit_identically():
    sc = scenes.cornell_scene(res=32, counts=(4, 4, 4), rays=64, atlas_res=256)
    rots = [scenes.frame_rotation(f) for f in range(4)]
    a = run_engine(sc, rots)
    b = run_engine(sc, rots[:2])
    st = b.state()
    c = ddgi.DDGIPipeline(sc.uniform)
    c.set_scene(sc)
    c.restore(b.irradiance, b.depth, st.frames, st.pingPong)
    for r in rots[2:]:
        c.update(r)
    assert np.array_equal(c.irradiance, a.irradiance) and np.array_equal(c.depth, a.depth)
    for p in (a, b, c):
        p.close()


@pytest.mark.parametrize("flags,name", [(0, "wavefront+tld4"), (abi.FLAG_SDF_LOADS, "wavefront+loads"), (abi.FLAG_TRACE_SIMPLE, "simple"),
                                        (abi.FLAG_NO_PREFILTER, "wavefront, full object lists"),
                                        (abi.FLAG_SHADE_UNSORTED, "wavefront, hits shaded in ray order"),
                                        (abi.FLAG_SHADE_UNSORTED | abi.FLAG_SDF_LOADS, "wavefront+loads, ray order"),
                                        (abi.FLAG_MARCH_PROBE_MAJOR, "wavefront, probe-major march order"),
                                        (abi.FLAG_MARCH_PROBE_MAJOR | abi.FLAG_SHADE_UNSORTED, "probe-major march order, ray-order shade"),
                                        (abi.FLAG_MARCH_ROWS, "row chunks (the shape large volumes get)"),
                                        (abi.FLAG_MARCH_ROWS | abi.FLAG_SDF_LOADS, "row chunks + loads"),
                                        (abi.FLAG_MARCH_ROWS | abi.FLAG_MARCH_PROBE_MAJOR | abi.FLAG_SHADE_UNSORTED, "row chunks, probe-major, ray-order shade")])
@pytest.mark.parametrize("cfg", ["c1", "city64"])
def test_trace_variants_match_oracle(oracle, flags, name, cfg):
    """Every trace kernel variant (thread-per-ray, wavefront with explicit loads, wavefront with texture gathers)
    must reproduce the oracle's ray buffers bit for bit."""
    sc = scenes.build(cfg)
    rots = [scenes.frame_rotation(f) for f in range(2)]
    orc = oracle.OraclePipeline(sc)
    for r in rots:
        orc.update(r)
    pipe = run_engine(sc, rots, flags=flags)
    assert_rays_match(pipe, orc)
    assert_atlases_match(pipe, orc)
    pipe.close()


def test_ragged_sizes(oracle):
    """Probe count not a multiple of 32, rays not a multiple of 16/32, non-cubic grid."""
    sc = scenes.cornell_scene(res=32, counts=(3, 5, 2), rays=50, atlas_res=256)
    rots = [scenes.frame_rotation(f) for f in range(2)]
    orc = oracle.OraclePipeline(sc)
    for r in rots:
        orc.update(r)
    for flags in (0, abi.FLAG_MARCH_ROWS, abi.FLAG_SDF_LOADS, abi.FLAG_MARCH_ROWS | abi.FLAG_SDF_LOADS, abi.FLAG_TRACE_SIMPLE):
        pipe = run_engine(sc, rots, flags=flags)
        assert_rays_match(pipe, orc)
        assert_atlases_match(pipe, orc)
        pipe.close()


def test_64_frame_convergence_shipped_scene_parameters(oracle):
    """north_star: converged irradiance over 64 frames within tolerance.  Probe-volume parameters of the shipped
    dark-room-emissive.scene (hysteresis 0.98, gamma 0.85, sharpness 50, 256 rays) on the Cornell SDF, 16x8x16
    probes as in BASELINE configs[1]; every frame re-quantises the feedback to fp16, so any drift would accumulate."""
    sc = scenes.cornell_scene(res=64, counts=(16, 8, 16), rays=256, atlas_res=512, hysteresis=0.98, gamma=0.85)
    orc = oracle.OraclePipeline(sc)
    pipe = ddgi.DDGIPipeline(sc.uniform)
    pipe.set_scene(sc)
    prev = None
    deltas = []
    for f in range(64):
        rot = scenes.frame_rotation(f)
        orc.update(rot)
        pipe.update(rot)
        if f in (0, 7, 31, 63):
            assert_atlases_match(pipe, orc)
        cur = f16(pipe.irradiance)[..., :3]
        if prev is not None:
            deltas.append(float(np.abs(cur - prev).mean()))
        prev = cur
    assert_rays_match(pipe, orc)
    # the temporal filter has settled: frame-to-frame change is the (1-h)-weighted ray noise, small against the signal
    assert np.mean(deltas[-8:]) <= np.mean(deltas[:8]) and np.mean(deltas[-8:]) < 0.01 * float(cur.mean())
    assert pipe.state().frames == 64
    pipe.close()


def test_large_probe_count_big_tile_kernels(oracle):
    """4096 probes: the 32-probe depth tiles (blend_depth_kernel<32>; irradiance<64> needs >= 18 944 probes and is covered by the full-size test) and many march chunks per warp."""
    sc = scenes.cornell_scene(res=32, counts=(16, 16, 16), rays=96, atlas_res=256)
    rots = [scenes.frame_rotation(f) for f in range(2)]
    orc = oracle.OraclePipeline(sc)
    for r in rots:
        orc.update(r)
    pipe = run_engine(sc, rots)
    assert_rays_match(pipe, orc)
    assert_atlases_match(pipe, orc)
    pipe.close()


@pytest.mark.parametrize("counts,rays", [((3, 5, 2), 50), ((8, 8, 8), 64), ((5, 4, 7), 300), ((4, 4, 5), 1024)])
def test_list_blend_equals_oracle_and_tiled_blend(oracle, counts, rays):
    """The list form of the FP32 blend (blend_lists.inc; the default from 148 x 64 probes per shard, forced here): per group of 2 x 2 texels the
    frame's rays with a non-zero weight are walked in ray order, so every texel sees the reference's sum (ProbeUpdate.glsl:66-103) - bit for bit
    the oracle and the tiled kernels.  Ragged cases: probe counts that are no multiple of 64, ray counts that are no multiple of the 64- / 256-ray
    phases (50: one partial phase; 300: 5 irradiance / 2 depth phases; 1024: 16 / 4), three frames so that the hysteresis branch runs."""
    sc = scenes.cornell_scene(res=32, counts=counts, rays=rays, atlas_res=256)
    rots = [scenes.frame_rotation(f) for f in range(3)]
    orc = oracle.OraclePipeline(sc)
    for r in rots:
        orc.update(r)
    a = run_engine(sc, rots, flags=abi.FLAG_BLEND_LISTS)
    b = run_engine(sc, rots, flags=abi.FLAG_BLEND_TILES)
    c = run_engine(sc, rots, flags=abi.FLAG_BLEND_LISTS | abi.FLAG_UNFUSED_BORDER, staged=True)
    for p_ in (a, b, c):
        assert_atlases_match(p_, orc)
    assert np.array_equal(a.irradiance, b.irradiance) and np.array_equal(a.depth, b.depth)
    assert np.array_equal(a.irradiance, c.irradiance) and np.array_equal(a.depth, c.depth)
    for p_ in (a, b, c):
        p_.close()


def test_list_blend_skips_gated_rays_with_infinite_radiance(oracle):
    """Same contract as the tiled kernels (test_blend_skips_gated_rays_...): a gated (zero) weight must not meet an fp16 Inf."""
    sc = scenes.build("c1")
    u = sc.uniform
    osc = oracle.OracleScene(sc)
    rad, dd, _, _ = osc.trace(scenes.frame_rotation(0))
    rad = rad.copy()
    rng = np.random.default_rng(5)
    for p_, r_ in zip(rng.integers(0, rad.shape[0], 40), rng.integers(0, rad.shape[1], 40)):
        rad[p_, r_, int(rng.integers(0, 3))] = 0x7c00  # +Inf
    irr = [oracle.new_atlases(u)[0] for _ in range(2)]
    dep = [oracle.new_atlases(u)[1] for _ in range(2)]
    pipe = ddgi.DDGIPipeline(u, flags=abi.FLAG_BLEND_LISTS)
    for f in range(2):
        oracle.blend(u, rad, dd, irr[f % 2], dep[f % 2], irr[1 - f % 2], dep[1 - f % 2], first_frame=(f == 0))
        oracle.border(u, irr[1 - f % 2], dep[1 - f % 2])
        pipe.set_ray_buffers(rad, dd)
        pipe.probe_update()
        pipe.border_update()
        pipe.end_frame()
    assert np.array_equal(pipe.irradiance, irr[0]), f"{(pipe.irradiance != irr[0]).sum()} irradiance values differ"
    assert np.array_equal(pipe.depth, dep[0])
    wf = f16(irr[0])[..., :3]
    assert np.isinf(wf).any() and not np.isnan(wf).any()
    pipe.close()


def probe_tiles(atlas, u, ids, side):
    S = side + 2
    per_row = u.probeCounts[0] * u.probeCounts[1]
    return np.stack([atlas[1 + (p // per_row) * S: 1 + (p // per_row) * S + S, 1 + (p % per_row) * S: 1 + (p % per_row) * S + S] for p in ids])


@pytest.mark.parametrize("cfg,sample", [("c4", 4096), ("c5", 2048)])
def test_full_size_config_on_a_stratified_probe_subsample(oracle, cfg, sample):
    """BASELINE configs[3] and [4] at FULL size on one GPU (city 512^3, 64x16x64 probes, 512 rays = 33.5 M rays per update; city 1024^3,
    128x32x128 probes, 1024 rays = 537 M rays per update, 2 GiB SDF); the oracle replays a stratified subsample of probes
    from the same device-generated inputs, two frames (the second through the hysteresis branch).  Rays and the subsample's atlas tiles
    (interior + border) must match bit for bit."""
    import torch

    def ray_rows(pipe, buf, ids):  # gather the subsample's rows on the device: the C5 ray buffers are 4.3 GB each
        pipe.synchronize()
        t = torch.as_tensor(pipe.device_view(buf), device="cuda")
        return t[torch.as_tensor(ids, device="cuda", dtype=torch.long)].cpu().numpy().view(np.uint16)

    sc = scenes.build(cfg, device="cuda")
    u = sc.uniform
    P = sc.probes
    ids = (np.arange(sample, dtype=np.int64) * P // sample).astype(np.int32)
    osc = oracle.OracleScene(sc)  # host copies of the device-generated inputs
    pipe = ddgi.DDGIPipeline(u)
    pipe.set_scene(sc)
    irr = [oracle.new_atlases(u)[0] for _ in range(2)]
    dep = [oracle.new_atlases(u)[1] for _ in range(2)]
    for f in range(2):
        rot = scenes.frame_rotation(f)
        pipe.update(rot)
        rad, dd, _, _ = osc.trace(rot, probe_ids=ids)
        grad, gdd = ray_rows(pipe, abi.BUF_RADIANCE, ids), ray_rows(pipe, abi.BUF_DIRECTION_DISTANCE, ids)
        assert np.array_equal(gdd, dd), f"frame {f}: direction/distance differ on {(gdd != dd).sum()} values"
        assert np.array_equal(grad, rad), f"frame {f}: radiance differs on {(grad != rad).sum()} values"
        oracle.blend_ids(u, rad, dd, irr[f % 2], dep[f % 2], irr[1 - f % 2], dep[1 - f % 2], first_frame=(f == 0), probe_ids=ids)
        for name, got, want, side in (("irradiance", pipe.irradiance, irr[1 - f % 2], 8), ("depth", pipe.depth, dep[1 - f % 2], 16)):
            rep = compare_atlas(name, probe_tiles(got, u, ids, side), probe_tiles(want, u, ids, side))
            print(cfg, f, rep)
            assert rep["out_of_tolerance"] == 0 and rep["mismatched_bits"] == 0, rep
    # size-independent properties on the FULL atlases: border idempotence and the mirror rule on every probe
    full_i, full_d = pipe.irradiance, pipe.depth
    pipe.border_update()
    pipe.synchronize()
    assert np.array_equal(pipe.irradiance, full_i) and np.array_equal(pipe.depth, full_d)
    t = full_i[1:-1, 1:-1].reshape(u.probeCounts[2], 10, -1, 10, 4).transpose(0, 2, 1, 3, 4)  # [z][xy][10][10][4]
    assert np.array_equal(t[:, :, 0, 1:-1], t[:, :, 1, 1:-1][:, :, ::-1]) and np.array_equal(t[:, :, 1:-1, 0], t[:, :, 1:-1, 1][:, :, ::-1])
    assert np.array_equal(t[:, :, 0, 0], t[:, :, -2, -2]) and np.array_equal(t[:, :, -1, -1], t[:, :, 1, 1])
    assert np.isfinite(f16(full_i)).all() and np.isfinite(f16(full_d)).all()
    pipe.close()
    del sc
    torch.cuda.empty_cache()


def test_consumer_sample_irradiance_and_sample_probe(oracle):
    """The step after the path: sampleIrradiance / SampleProbe.comp read the just-written atlases with
    bilinear taps through the 1-texel borders — so this also proves the border + outer-pad layout end to end."""
    import os

    sc = scenes.build("c1")
    sc.uniform.normalBias = 0.1
    rots = [scenes.frame_rotation(f) for f in range(3)]
    orc = oracle.OraclePipeline(sc)
    for r in rots:
        orc.update(r)
    pipe = run_engine(sc, rots)
    rng = np.random.default_rng(3)
    n = 4096
    P = rng.uniform(-4.8, 4.8, (n, 3)).astype(np.float32)
    P[:64] = rng.uniform(-9, 9, (64, 3))  # outside the probe grid: base cell clamps
    N = rng.standard_normal((n, 3)).astype(np.float32)
    N /= np.linalg.norm(N, axis=1, keepdims=True)
    Wo = rng.standard_normal((n, 3)).astype(np.float32)
    Wo /= np.linalg.norm(Wo, axis=1, keepdims=True)
    want = oracle.sample_irradiance(sc.uniform, orc.irradiance, orc.depth, P, N, Wo)
    got = pipe.sample_irradiance(P, N, Wo)
    assert np.isfinite(want).all() and want.mean() > 0.01
    assert np.allclose(got, want, rtol=1e-3, atol=1e-4)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{(got != want).sum()} of {got.size} values differ"
    # the shipped-SPIR-V golden G-buffer through the engine's SampleProbe kernel (atlases = the engine's own)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spirv_golden_consumer.npz"))
    want = oracle.sample_probe(sc.uniform, orc.irradiance, orc.depth, g["g_depth"], g["g_normal"], g["camera"], g["view_proj_inv"])
    got = pipe.sample_probe(g["g_depth"], g["g_normal"], g["camera"], g["view_proj_inv"])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"max abs diff {np.abs(got - want).max()}"
    assert (got[0, :3] == 0).all()  # depth == 1 pixels are cleared
    pipe.close()


def test_infinite_bounce_refresh_closes_the_loop(oracle):
    """BASELINE configs[2] in miniature: every 16 frames the surface light cache is rebuilt as
    emissive + direct + intensity * (albedo_0.9 - albedo*metal)/pi * sampleIrradiance(previous atlases) and the trace then
    sees the brighter cache.  Engine (lux_ddgi_indirect_light) and oracle must stay bit-identical through two refreshes."""
    sc = scenes.cornell_scene(res=32, counts=(8, 4, 8), rays=64, atlas_res=256)
    sc.uniform.normalBias = 0.1
    gb = sc.meta["gbuffer"]
    base = sc.light.numpy().view(np.uint16).copy()
    cam = np.array([0.0, 0.0, 4.0], dtype=np.float32)
    orc = oracle.OraclePipeline(sc)
    pipe = ddgi.DDGIPipeline(sc.uniform)
    pipe.set_scene(sc)
    means = []
    for f in range(34):
        if f and f % 16 == 0:  # GI_FRAMES cadence (GlobalSurfaceAtlas.cpp:50, 842-847)
            o_light = base.copy()
            oracle.indirect_light(sc.uniform, orc.irradiance, orc.depth, o_light, base, gb["texel"], gb["pos"], gb["normal"], gb["albedo"],
                                  gb["metallic"], 1.2, cam)
            orc.os.light[...] = o_light  # the oracle scene reads this array in place
            pipe.indirect_light(base, gb["texel"], gb["pos"], gb["normal"], gb["albedo"], gb["metallic"], 1.2, cam)
            got_light = pipe.surface_light_cache()
            assert np.array_equal(got_light, o_light), f"light cache differs on {(got_light != o_light).sum()} values after refresh at frame {f}"
            means.append(float(f16(o_light)[..., :3].mean()))
        rot = scenes.frame_rotation(f)
        orc.update(rot)
        pipe.update(rot)
    assert_rays_match(pipe, orc)
    assert_atlases_match(pipe, orc)
    assert means[0] > float(f16(base)[..., :3].mean()) and means[1] >= means[0]  # each bounce adds energy
    pipe.close()


@pytest.mark.parametrize("cfg", ["city128", "city64"])
def test_march_work_order_does_not_change_results(oracle, cfg):
    """The wavefront march visits direction clusters (outer) x spatially tiled probe units (inner), in beam chunks (2 probes x 32 adjacent
    directions; small volumes) or row chunks (32 probes x 2 directions; volumes beyond the TLB's reach); records are addressed in that order
    by every later stage.  Three frames (hysteresis on) must equal the oracle bit for bit in both chunk shapes, in probe-major order
    (LUX_DDGI_FLAG_MARCH_PROBE_MAJOR) and with the weights computed on the context's stream (LUX_DDGI_FLAG_NO_PIPELINE)."""
    sc = scenes.build(cfg)
    rots = [scenes.frame_rotation(f) for f in range(3)]
    orc = oracle.OraclePipeline(sc)
    for r in rots:
        orc.update(r)
    pipe = run_engine(sc, rots)
    serial = run_engine(sc, rots, flags=abi.FLAG_NO_PIPELINE)
    major = run_engine(sc, rots, flags=abi.FLAG_MARCH_PROBE_MAJOR)
    rows = run_engine(sc, rots, flags=abi.FLAG_MARCH_ROWS)
    beams = run_engine(sc, rots, flags=abi.FLAG_MARCH_BEAMS)
    assert_rays_match(pipe, orc)
    assert_atlases_match(pipe, orc)
    for other in (serial, major, rows, beams):
        assert np.array_equal(pipe.irradiance, other.irradiance) and np.array_equal(pipe.depth, other.depth)
        assert np.array_equal(pipe.radiance, other.radiance) and np.array_equal(pipe.direction_distance, other.direction_distance)
        other.close()
    pipe.close()


def test_pipelined_row_downloads_with_fences(oracle):
    """Frame-pipelined consumer (bench.py's e2e loop): rows of frame f are copied on the download stream while frame f+1 computes;
    waiting on frame f's fence one frame later delivers exactly frame f's atlas rows."""
    import torch

    sc = scenes.build("city64")
    orc = oracle.OraclePipeline(sc)
    pipe = ddgi.DDGIPipeline(sc.uniform)
    pipe.set_scene(sc)
    st = pipe.state()
    u = sc.uniform
    bufs = [torch.empty(st.irradianceRowCount * u.irradianceTextureWidth * 8, dtype=torch.uint8).pin_memory() for _ in range(2)]
    fences, wants = [], []
    for f in range(4):
        rot = scenes.frame_rotation(f)
        orc.update(rot)
        wants.append(orc.irradiance[st.irradianceRowBegin:st.irradianceRowBegin + st.irradianceRowCount].copy())
        pipe.update(rot)
        pipe.download_rows_async_ptr(abi.BUF_IRRADIANCE, st.irradianceRowBegin, st.irradianceRowCount, bufs[f & 1].data_ptr())
        fences.append(pipe.download_fence())
        if f >= 1:
            pipe.wait_fence(fences[f - 1])
            got = bufs[(f - 1) & 1].numpy().view(np.uint16).reshape(wants[f - 1].shape)
            assert np.array_equal(got, wants[f - 1]), f"frame {f - 1}"
    pipe.wait_fence(fences[-1])
    assert np.array_equal(bufs[1].numpy().view(np.uint16).reshape(wants[3].shape), wants[3])
    pipe.close()


@pytest.mark.parametrize("cfg", ["c1", "city64"])
def test_surface_culling_on_device_matches_oracle(oracle, cfg):
    """Row f4 (first half): lux_ddgi_cull_surface_objects rebuilds the chunk / culled-object lists from the bound object buffer.  Word for
    word equal to the oracle's SDFCulling restatement executed in ascending chunk order (the layout the engine defines), with room for
    every list and with the shader's capacity rule dropping most of them; tracing on the rebuilt lists equals the oracle tracing on its own."""
    sc = scenes.build(cfg)
    pipe = ddgi.DDGIPipeline(sc.uniform)
    pipe.set_scene(sc)
    data = abi.GlobalSurfaceAtlasData.from_buffer_copy(bytes(sc.atlas_data))
    data.culledObjectsCapacity = 1 << 30
    want_chunks, want_cull = oracle.surface_cull(data, sc.objects, capacity_words=1 + 64000 * (len(sc.objects) + 1))
    chunks, cull = pipe.cull_surface_objects()
    used = int(want_cull[0])
    assert int(cull[0]) == used and len(cull) >= used
    assert np.array_equal(chunks, want_chunks) and np.array_equal(cull[:used], want_cull[:used])
    # the lists the fixture generator wrote describe the same sets (its float64 box-sphere test may differ on exact ties only)
    same = sum(1 for a in np.nonzero(want_chunks)[0][:2000] if sc.chunks[a] and
               np.array_equal(want_cull[want_chunks[a]:want_chunks[a] + 1 + want_cull[want_chunks[a]]], sc.cull[sc.chunks[a]:sc.chunks[a] + 1 + sc.cull[sc.chunks[a]]]))
    assert same >= 0.99 * min(2000, int((want_chunks != 0).sum()))
    # tracing on the rebuilt lists
    sc2 = scenes.build(cfg)
    sc2.chunks, sc2.cull = want_chunks, want_cull[:used].copy()
    orc = oracle.OraclePipeline(sc2)
    for f in range(2):
        rot = scenes.frame_rotation(f)
        orc.update(rot)
        pipe.update(rot)
    assert_rays_match(pipe, orc)
    assert_atlases_match(pipe, orc)
    # the shader's capacity rule
    small = max(64, used // 5)
    data.culledObjectsCapacity = small
    want_chunks_s, want_cull_s = oracle.surface_cull(data, sc.objects, capacity_words=used + 8)
    chunks_s, cull_s = pipe.cull_surface_objects(capacity_words=small)
    assert 0 < (chunks_s != 0).sum() < (chunks != 0).sum()
    assert np.array_equal(chunks_s, want_chunks_s) and int(cull_s[0]) == int(want_cull_s[0]) and np.array_equal(cull_s[1:small], want_cull_s[1:small])
    pipe.close()


@pytest.mark.parametrize("cfg,flags", [("c1", 0), ("city64", 0), ("city64", abi.FLAG_SDF_LOADS)])
def test_generic_global_sdf_trace_matches_oracle(oracle, cfg, flags):
    """Row f4: lux_ddgi_trace_global_sdf = tracyGlobalSDF for arbitrary rays (shadow / reflection / surface-cache light rays): random origins,
    directions, maxDistance, stepScale in {0.5, 1, 2}, needsHitNormal on and off, start bias 0 (GISDFRays, SDFShadow) and 2 (SDFDeferredLight).
    Every field of every GlobalSDFHit equals the oracle's, bit for bit."""
    sc = scenes.build(cfg, with_atlas=False)
    pipe = ddgi.DDGIPipeline(sc.uniform, flags=flags)
    pipe.set_global_sdf(sc.sdf_data, sc.sdf, sc.mip)
    rng = np.random.default_rng(11)
    n = 20000
    c = np.array([sc.sdf_data.cascadePosDistance[0][i] for i in range(3)], dtype=np.float32)
    D = float(sc.sdf_data.cascadePosDistance[0][3])
    traces = np.zeros(n, dtype=abi.SDF_TRACE_DTYPE)
    traces["worldPosition"] = (c + rng.uniform(-1.1 * D, 1.1 * D, size=(n, 3))).astype(np.float32)  # some origins outside the cascade
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[:64] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 64)] * rng.choice([-1.0, 1.0], size=(64, 1)).astype(np.float32)  # axis aligned
    traces["worldDirection"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    traces["minDistance"] = rng.uniform(0, 1, n).astype(np.float32)
    traces["maxDistance"] = np.where(rng.random(n) < 0.3, abi.GLOBAL_SDF_WORLD_SIZE, rng.uniform(0.5, 3 * D, n)).astype(np.float32)
    traces["stepScale"] = rng.choice(np.float32([0.5, 1.0, 2.0]), n)
    traces["needsHitNormal"] = rng.integers(0, 2, n)
    for bias in (0.0, 2.0):
        want = oracle.trace_global_sdf(sc.sdf_data, sc.sdf, sc.mip, traces, bias)
        got = pipe.trace_global_sdf(traces, bias)
        assert (want["hitTime"] >= 0).mean() > 0.2 and (want["hitTime"] < 0).mean() > 0.05
        for f in abi.SDF_HIT_DTYPE.names:
            a, b = got[f].view(np.uint32), want[f].view(np.uint32)
            assert np.array_equal(a, b), f"bias {bias}: {f} differs on {(a != b).sum()} of {a.size} values"
    pipe.close()


def test_surface_direct_light_matches_shipped_spirv_and_oracle(oracle):
    """Row f4: lux_ddgi_surface_direct_light = SDFDeferredLight.frag (fetchLight, shadow ray with start bias 2, BRDF) added into the RGBA16F light cache.
    (a) against the shipped SPIR-V's outColor on the golden G-buffer (tests/golden/spirv_golden_directlight.npz), bit for bit, for a directional,
    a point and a spot light; (b) against the oracle on every G-buffer texel of the Cornell surface cache, the three lights accumulated one after
    another on top of the scene's own light cache (additive blend, alpha counts the passes)."""
    import os

    from tests.golden import make_spirv_golden_directlight as g

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "spirv_golden_directlight.npz"))
    sc = g.golden_scene()
    res = int(sc.atlas_data.resolution)
    for flags in (0, abi.FLAG_SDF_LOADS):
        pipe = ddgi.DDGIPipeline(sc.uniform, flags=flags)
        pipe.set_scene(sc)
        # (a) golden replay into an empty light cache
        n = len(gold["pos"])
        normals = oracle.octohedral_to_direction(gold["oct_normal"])  # the value the shader decodes (a pure function of the stored input)
        texel = np.arange(n, dtype=np.uint32)
        for name in g.LIGHTS:
            pipe.update_surface_light_cache(np.zeros((res, res, 4), dtype=np.uint16))
            pipe.surface_direct_light(abi.make_light(gold[f"light_{name}"]), gold["camera"], texel, gold["pos"], normals, gold["albedo"], gold["pbr"])
            got = pipe.surface_light_cache().reshape(-1, 4)
            want = gold[f"out_{name}"].astype(np.float16).view(np.uint16)
            assert np.array_equal(got[:n], want), f"{name}: {(got[:n] != want).sum()} of {want.size} fp16 values differ from the shipped shader"
            assert not got[n:].any(), "texels outside the list were written"
        # (b) whole G-buffer, three lights on top of the existing cache
        gb = sc.meta["gbuffer"]
        rng = np.random.default_rng(5)
        k = len(gb["texel"])
        nrm = gb["normal"].astype(np.float32) + rng.normal(scale=0.1, size=(k, 3)).astype(np.float32)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        pbr = np.stack([rng.choice([0.0, 0.5, 1.0], k), rng.uniform(0.05, 1.0, k)], -1).astype(np.float32)
        base = np.ascontiguousarray(sc.light.numpy().view(np.uint16)).reshape(res, res, 4).copy()
        pipe.update_surface_light_cache(base)
        want = base.reshape(-1, 4).copy()
        for name in g.LIGHTS:
            light = abi.make_light(gold[f"light_{name}"])
            oracle.surface_direct_light(sc.sdf_data, sc.sdf, sc.mip, light, gold["camera"], want, gb["texel"], gb["pos"], nrm, gb["albedo"], pbr)
            pipe.surface_direct_light(light, gold["camera"], gb["texel"], gb["pos"], nrm, gb["albedo"], pbr)
        got = pipe.surface_light_cache().reshape(-1, 4)
        assert (want != base.reshape(-1, 4)).any(1).sum() == len(np.unique(gb["texel"]))
        assert np.array_equal(got, want), f"flags {flags}: {(got != want).sum()} of {want.size} fp16 values differ from the oracle"
        # argument errors come back as status codes
        bad = abi.make_light(gold["light_point"])
        bad.type = 7.0
        with pytest.raises(ddgi.LuxError):
            pipe.surface_direct_light(bad, gold["camera"], texel, gold["pos"], normals, gold["albedo"], gold["pbr"])
        pipe.close()


@pytest.mark.parametrize("cascades", [2, 4])
def test_cascaded_global_sdf_matches_oracle(oracle, cascades):
    """The reference always runs a CASCADED global SDF (2 cascades, GlobalDistanceField.cpp:182-191; the structs carry 4): nested, off-centre
    cascades side by side in one volume, most probes outside cascade 0.  Every trace variant reproduces the oracle's ray buffers and atlases bit
    for bit over 3 frames, and the generic ray entry every GlobalSDFHit field incl. hitCascade; the oracle itself is pinned on this layout by
    the shipped GISDFRays.comp.spv (tests/test_spirv_golden.py::test_cascaded_trace_matches_shipped_spirv)."""
    sc = scenes.cornell_scene(res=64, counts=(8, 4, 8), rays=96, atlas_res=256, cascades=cascades)
    rots = [scenes.frame_rotation(f) for f in range(3)]
    orc = oracle.OraclePipeline(sc)
    for r in rots:
        orc.update(r)
    for flags in (0, abi.FLAG_SDF_LOADS, abi.FLAG_TRACE_SIMPLE, abi.FLAG_SHADE_UNSORTED, abi.FLAG_NO_PREFILTER, abi.FLAG_MARCH_ROWS,
                  abi.FLAG_MARCH_ROWS | abi.FLAG_SDF_LOADS):
        pipe = run_engine(sc, rots, flags=flags)
        assert_rays_match(pipe, orc)
        assert_atlases_match(pipe, orc)
        if flags in (0, abi.FLAG_SDF_LOADS):
            rng = np.random.default_rng(17)
            n = 20000
            traces = np.zeros(n, dtype=abi.SDF_TRACE_DTYPE)
            traces["worldPosition"] = rng.uniform(-7.0, 7.0, size=(n, 3)).astype(np.float32)  # some origins outside every cascade
            d = rng.normal(size=(n, 3)).astype(np.float32)
            traces["worldDirection"] = d / np.linalg.norm(d, axis=1, keepdims=True)
            traces["maxDistance"] = np.where(rng.random(n) < 0.5, abi.GLOBAL_SDF_WORLD_SIZE, rng.uniform(0.5, 12.0, n)).astype(np.float32)
            traces["stepScale"] = rng.choice(np.float32([0.5, 1.0, 2.0]), n)
            traces["needsHitNormal"] = rng.integers(0, 2, n)
            for bias in (0.0, 2.0):
                want = oracle.trace_global_sdf(sc.sdf_data, sc.sdf, sc.mip, traces, bias)
                got = pipe.trace_global_sdf(traces, bias)
                assert (np.bincount(want["hitCascade"][want["hitTime"] >= 0], minlength=cascades) > 0).all()
                for f in abi.SDF_HIT_DTYPE.names:
                    a, b = got[f].view(np.uint32), want[f].view(np.uint32)
                    assert np.array_equal(a, b), f"flags {flags} bias {bias}: {f} differs on {(a != b).sum()} of {a.size} values"
        pipe.close()


def _screen_inputs(w, h, seed, eye=(0.3, 0.4, 3.9), target=(-0.2, -0.5, -1.0)):
    from tests.golden.make_spirv_golden_consumer import look_at_perspective

    rng = np.random.default_rng(seed)
    vp = look_at_perspective(np.array(eye, dtype=np.float64), np.array(target, dtype=np.float64), np.array([0.0, 1.0, 0.0]), np.radians(70), w / h, 0.1, 50.0)
    vpi = np.linalg.inv(vp).astype(np.float32).T.reshape(16).copy()
    depth = rng.uniform(0.90, 0.985, (h, w)).astype(np.float32)
    far = rng.random((h, w)) < 0.1
    depth[far] = rng.uniform(0.9975, 0.9995, int(far.sum())).astype(np.float32)
    depth[rng.random((h, w)) < 0.05] = 1.0
    depth[0, 0] = 1.0
    nrm = np.zeros((h, w, 4), dtype=np.float32)
    nrm[..., :2] = rng.uniform(-1, 1, (h, w, 2))
    pbr = np.zeros((h, w, 4), dtype=np.float32)
    pbr[..., 1] = rng.choice(np.float32([0.01, 0.04, 0.2, 0.44, 0.5, 0.9]), (h, w))
    sobol = rng.integers(0, 256, (256, 4), dtype=np.uint8)
    scr = rng.integers(0, 256, (128, 128, 4), dtype=np.uint8)
    return vpi, depth, nrm, pbr, sobol, scr


def test_sdf_reflection_matches_shipped_spirv_and_oracle(oracle):
    """Row f4: lux_ddgi_sdf_reflection = SDFReflection.comp.  (a) the RGBA16F image of the shipped SPIR-V on the golden G-buffer
    (tests/golden/spirv_golden_screen.npz), bit for bit, approximateWithDDGI on and off, texture and load SDF paths; (b) against the oracle on a
    150 x 130 G-buffer (ragged against every block size) in the 2-cascade Cornell scene after two probe updates of the engine itself."""
    from tests.golden import make_spirv_golden as base

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "spirv_golden_screen.npz"))
    sc = base.golden_scene()
    sc.uniform = abi.DDGIUniform.from_buffer_copy(g["uniform"].tobytes())
    for flags in (0, abi.FLAG_SDF_LOADS):
        pipe = ddgi.DDGIPipeline(sc.uniform, flags=flags)
        pipe.set_scene(sc)
        pipe.restore(g["irradiance"], g["depth_atlas"], 2, 0)
        for approx in (1, 0):
            frames, trim, inten = g[f"refl_params_{approx}"]
            push = abi.make_reflection_push(g["refl_camera"], g["refl_view_proj_inv"], int(frames), trim, inten, approx)
            got = pipe.sdf_reflection(push, g["refl_depth"], g["refl_normal"], g["refl_pbr"], g["sobol"], g["scrambling"], out=np.full(g["refl_out_1"].shape, 0x3555, dtype=np.uint16))
            want = g[f"refl_out_{approx}"]
            assert np.array_equal(got, want), f"flags {flags} approximateWithDDGI={approx}: {(got != want).any(-1).sum()} pixels differ from the shipped shader"
        pipe.close()
    # (b)
    sc = scenes.cornell_scene(res=32, counts=(4, 4, 4), rays=64, atlas_res=256, cascades=2)
    sky = np.zeros((6, 2, 2, 4), dtype=np.float16)
    sky[...] = np.random.default_rng(2).uniform(0, 2, sky.shape)
    sc.sky_face, sc.sky = 2, sky
    sc.uniform.normalBias = 0.1
    pipe = ddgi.DDGIPipeline(sc.uniform)
    pipe.set_scene(sc)
    for f in range(2):
        pipe.update(scenes.frame_rotation(f))
    irr, dep = pipe.irradiance, pipe.depth
    w, h = 150, 130
    vpi, depth, nrm, pbr, sobol, scr = _screen_inputs(w, h, 41)
    for approx in (1, 0):
        push = abi.make_reflection_push([0.3, 0.4, 3.9], vpi, 3, 0.9, 1.1, approx)
        want = np.full((h, w, 4), 0x1234, dtype=np.uint16)
        oracle.sdf_reflection(sc, irr, dep, push, depth, nrm, pbr, sobol, scr, want)
        got = pipe.sdf_reflection(push, depth, nrm, pbr, sobol, scr, out=np.full((h, w, 4), 0x1234, dtype=np.uint16))
        assert (want[..., 0] == 0x1234).sum() >= (depth == 1.0).sum() > 100
        assert np.array_equal(got, want), f"approximateWithDDGI={approx}: {(got != want).any(-1).sum()} of {w * h} pixels differ from the oracle"
    pipe.close()


def test_sdf_shadow_matches_shipped_spirv_and_oracle(oracle):
    """Row f4: lux_ddgi_sdf_shadow = SDFShadow.comp.  (a) the R32UI words of the shipped SPIR-V on the golden G-buffer for a directional, a point
    and a spot light, incl. the workgroup that stores nothing; (b) against the oracle on a 160 x 96 G-buffer in the 2-cascade Cornell scene."""
    from tests.golden import make_spirv_golden as base
    from tests.golden import make_spirv_golden_screen as gs

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "spirv_golden_screen.npz"))
    sc = base.golden_scene()
    h, w = g["shadow_depth"].shape
    for flags in (0, abi.FLAG_SDF_LOADS):
        pipe = ddgi.DDGIPipeline(sc.uniform, flags=flags)
        pipe.set_global_sdf(sc.sdf_data, sc.sdf, sc.mip)
        for name in gs.LIGHTS:
            got = pipe.sdf_shadow(abi.make_light(g[f"shadow_light_{name}"]), g["shadow_view_proj_inv"], int(g["shadow_frames"]), float(g["shadow_bias"]),
                                  g["shadow_depth"], g["shadow_normal"], g["sobol"], g["scrambling"], out=np.full((h // 4, w // 8), 0xDEADBEEF, dtype=np.uint32))
            want = g[f"shadow_out_{name}"]
            assert np.array_equal(got, want), f"flags {flags} {name}: {[hex(int(a)) for a in got.reshape(-1)]} != {[hex(int(a)) for a in want.reshape(-1)]}"
        with pytest.raises(ddgi.LuxError):
            pipe.sdf_shadow(abi.make_light(g["shadow_light_point"]), g["shadow_view_proj_inv"], 0, 0.1, g["shadow_depth"][:6], g["shadow_normal"][:6], g["sobol"], g["scrambling"])
        pipe.close()
    # (b)
    sc = scenes.cornell_scene(res=32, counts=(2, 2, 2), rays=32, atlas_res=256, cascades=2, with_atlas=False)
    pipe = ddgi.DDGIPipeline(sc.uniform)
    pipe.set_global_sdf(sc.sdf_data, sc.sdf, sc.mip)
    w, h = 160, 96
    vpi, depth, nrm, _, sobol, scr = _screen_inputs(w, h, 43)
    bits = 0
    for name in gs.LIGHTS:
        light = abi.make_light(g[f"shadow_light_{name}"])
        want = np.full((h // 4, w // 8), 0xDEADBEEF, dtype=np.uint32)
        oracle.sdf_shadow(sc.sdf_data, sc.sdf, sc.mip, light, vpi, 7, 0.05, depth, nrm, sobol, scr, want)
        got = pipe.sdf_shadow(light, vpi, 7, 0.05, depth, nrm, sobol, scr, out=np.full((h // 4, w // 8), 0xDEADBEEF, dtype=np.uint32))
        assert np.array_equal(got, want), f"{name}: {(got != want).sum()} of {want.size} words differ from the oracle"
        assert want[0, 0] == 0xDEADBEEF  # pixel (0, 0) is sky
        bits += sum(bin(int(v)).count("1") for v in want.reshape(-1) if v != 0xDEADBEEF)
    assert 500 < bits < 0.9 * 3 * w * h
    pipe.close()


def test_l2_bandwidth_measurement():
    """lux_ddgi_measure_l2_read_bandwidth (the denominator of bench.py's request-level roofline): an L2-resident sweep must come out well above
    HBM bandwidth and below anything physical (measured: 21.5 TB/s for 64 MiB; 17 TB/s for 1 GiB, which the staggered blocks also serve mostly from L2)."""
    sc = scenes.cornell_scene(res=32, counts=(2, 2, 2), rays=32, atlas_res=256, with_atlas=False)
    pipe = ddgi.DDGIPipeline(sc.uniform)
    l2 = pipe.measure_l2_read_bandwidth()
    big = pipe.measure_l2_read_bandwidth(1 << 30, repeats=2)
    print("L2 read sweep GB/s:", l2, " 1 GiB sweep GB/s:", big)
    assert 8000.0 < l2 < 100000.0 and 3000.0 < big < 100000.0
    with pytest.raises(ddgi.LuxError):
        pipe.measure_l2_read_bandwidth(repeats=0)
    pipe.close()
