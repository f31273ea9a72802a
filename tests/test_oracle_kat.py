"""Known-answer tests of the CPU oracle.  The reference ships no tests for this path, so these
closed-form checks — plus the reference-derived fixtures under tests/golden/ — are what pins the restatement."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from luxgi_b200 import abi, scenes
from tests.util import f16

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---------------------------------------------------------------------------------------------------------------
# fp16 conversion = numpy's IEEE binary16 (RTNE, subnormals, overflow)
# ---------------------------------------------------------------------------------------------------------------
def test_f2h_matches_ieee_rtne(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    vals = np.concatenate([
        rng.standard_normal(20000).astype(np.float32) * 10.0 ** rng.integers(-9, 6, 20000).astype(np.float32),
        np.array([0.0, -0.0, 65504.0, 65519.99, 65520.0, 1e9, -1e9, 6.1e-5, 6.0e-8, 2.98e-8, 2.99e-8, 5.96e-8, 60000.0, 1.0, 0.333], dtype=np.float32),
        np.arange(0, 65536, dtype=np.uint16).view(np.float16).astype(np.float32)[: 0x7C00],
    ])
    # halfway cases between consecutive halfs
    h = np.arange(1, 0x7BFF, 37, dtype=np.uint16)
    mid = (h.view(np.float16).astype(np.float64) + (h + 1).astype(np.uint16).view(np.float16).astype(np.float64)) / 2
    vals = np.concatenate([vals, mid.astype(np.float32), -mid.astype(np.float32)])
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    got = np.array([L.oracle_f2h(float(v)) for v in vals], dtype=np.uint16)
    assert np.array_equal(got, want)


def test_h2f_is_exact(oracle):
    L = oracle.lib()
    bits = np.arange(0, 0x7C01, 13, dtype=np.uint16)
    got = np.array([L.oracle_h2f(int(b)) for b in bits], dtype=np.float32)
    assert np.array_equal(got, bits.view(np.float16).astype(np.float32))
    assert L.oracle_h2f(0xFC00) == -math.inf


# ---------------------------------------------------------------------------------------------------------------
# sphericalFibonacci / octahedral mapping / probe placement
# ---------------------------------------------------------------------------------------------------------------
def test_spherical_fibonacci(oracle):
    L = oracle.lib()
    out = (C.c_float * 3)()
    for R in (64, 256, 1024):
        zs = []
        for i in range(R):
            L.oracle_spherical_fibonacci(i, R, None, out)
            v = np.array(out[:], dtype=np.float64)
            assert abs(np.linalg.norm(v) - 1.0) < 2e-6
            zs.append(v[2])
            assert abs(v[2] - (1.0 - (2 * i + 1) / R)) < 1e-6  # cos(theta) = 1 - (2i+1)/R
            phi = 2 * math.pi * ((i * 0.6180340051651001) % 1.0)
            if 1 - v[2] ** 2 > 1e-4:
                assert abs(math.atan2(v[1], v[0]) - math.atan2(math.sin(phi), math.cos(phi))) < 3e-3 * max(1, i / 50)
        assert abs(np.mean(zs)) < 1e-6  # symmetric in z


def test_rotated_directions_are_unit_and_rotation_is_applied(oracle):
    L = oracle.lib()
    rot = scenes.frame_rotation(3)
    m = rot.reshape(4, 4).T[:3, :3].astype(np.float64)  # column-major -> matrix
    assert np.allclose(m @ m.T, np.eye(3), atol=1e-6)
    a, b = (C.c_float * 3)(), (C.c_float * 3)()
    for i in (0, 5, 63):
        L.oracle_spherical_fibonacci(i, 64, None, a)
        L.oracle_spherical_fibonacci(i, 64, rot.ctypes.data_as(C.c_void_p), b)
        assert np.allclose(m @ np.array(a[:]), np.array(b[:]), atol=1e-6)


def oct_encode(v):
    v = np.asarray(v, dtype=np.float64)
    l1 = np.abs(v).sum()
    r = v[:2] / l1
    if v[2] < 0:
        r = (1 - np.abs(r[::-1])) * np.where(r >= 0, 1.0, -1.0)
    return r


def test_oct_decode_roundtrip(oracle):
    L = oracle.lib()
    out = (C.c_float * 3)()
    for side in (8, 16):
        for j in range(side):
            for i in range(side):
                L.oracle_oct_decode(i, j, side, out)
                v = np.array(out[:], dtype=np.float64)
                assert abs(np.linalg.norm(v) - 1) < 1e-6
                o = oct_encode(v)  # consumer mapping (DDGICommon.glsl:60-72) inverts the texel centre
                want = (np.array([i, j]) + 0.5) * 2 / side - 1
                assert np.allclose(o, want, atol=1e-6)


def test_probe_location(oracle):
    L = oracle.lib()
    u = abi.make_uniform((-1.0, 2.0, 3.0), (0.5, 0.25, 2.0), (3, 4, 5), 32)
    out = (C.c_float * 3)()
    for idx in range(60):
        L.oracle_probe_location(C.byref(u), idx, out)
        x, y, z = idx % 3, (idx % 12) // 3, idx // 12
        assert np.allclose(out[:], [-1 + 0.5 * x, 2 + 0.25 * y, 3 + 2.0 * z])


def test_inverse4(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(0)
    for _ in range(20):
        ang = rng.uniform(0, 6.28)
        m = np.eye(4)
        m[:3, :3] = [[math.cos(ang), 0, math.sin(ang)], [0, 1, 0], [-math.sin(ang), 0, math.cos(ang)]]
        m[:3, 3] = rng.uniform(-50, 50, 3)
        cm = np.ascontiguousarray(m.T.reshape(16), dtype=np.float32)
        o = np.zeros(16, dtype=np.float32)
        L.oracle_inverse4(cm.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
        assert np.allclose(o.reshape(4, 4).T @ m, np.eye(4), atol=2e-5)


# ---------------------------------------------------------------------------------------------------------------
# software texture unit
# ---------------------------------------------------------------------------------------------------------------
def test_trilinear_sampling(oracle):
    L = oracle.lib()
    n = 8
    z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    vol = (0.125 * x + 0.25 * y - 0.0625 * z).astype(np.float16)  # linear field, exactly representable
    bits = np.ascontiguousarray(vol).view(np.uint16)
    p = bits.ctypes.data_as(C.c_void_p)
    # texel centres reproduce texels
    for (i, j, k) in [(0, 0, 0), (3, 4, 5), (7, 7, 7)]:
        got = L.oracle_sample3d(p, n, n, n, (i + 0.5) / n, (j + 0.5) / n, (k + 0.5) / n)
        assert got == float(vol[k, j, i])
    # a linear field is reproduced exactly between texel centres
    got = L.oracle_sample3d(p, n, n, n, 3.0 / n, 4.25 / n, 2.75 / n)
    assert got == pytest.approx(0.125 * 2.5 + 0.25 * 3.75 - 0.0625 * 2.25, abs=1e-6)
    # clamp-to-edge
    assert L.oracle_sample3d(p, n, n, n, -0.3, 0.5 / n, 0.5 / n) == float(vol[0, 0, 0])
    assert L.oracle_sample3d(p, n, n, n, 1.7, 7.5 / n, 7.5 / n) == float(vol[7, 7, 7])


# ---------------------------------------------------------------------------------------------------------------
# Blend: uniform field, hysteresis trajectory, naive == hoisted
# ---------------------------------------------------------------------------------------------------------------
def _uniform_rays(oracle, u, L_rgb, dist, rot):
    P, R = abi.probe_count(u), u.raysPerProbe
    rad = np.zeros((P, R, 4), dtype=np.float16)
    rad[..., :3] = np.asarray(L_rgb, dtype=np.float16)
    dd = np.zeros((P, R, 4), dtype=np.float16)
    out = (C.c_float * 3)()
    for r in range(R):
        oracle.lib().oracle_spherical_fibonacci(r, R, rot.ctypes.data_as(C.c_void_p), out)
        dd[:, r, :3] = np.array(out[:], dtype=np.float32).astype(np.float16)
    dd[..., 3] = np.float16(dist)
    return rad.view(np.uint16), dd.view(np.uint16)


def test_uniform_field_known_answer(oracle):
    """Every ray returns radiance L and distance d => irradiance texel = pow(L/2, 1/gamma), depth texel =
    (min(maxD, d-0.01)/2, min(maxD, d-0.01)^2/2) (ProbeUpdate.glsl:75,96-98,133-141)."""
    u = abi.make_uniform((0, 0, 0), (1, 1, 1), (2, 2, 1), 128, max_distance=6.0, gamma=5.0)
    Lrgb, d = (0.5, 0.25, 2.0), 3.0
    rad, dd = _uniform_rays(oracle, u, Lrgb, d, scenes.frame_rotation(1))
    irr, dep = oracle.new_atlases(u)
    oracle.blend(u, rad, dd, None, None, irr, dep, first_frame=True)
    S = 10
    for p in range(4):
        blk = f16(irr[2:10, 2 + p * S: 10 + p * S, :])
        for c in range(3):
            want = (Lrgb[c] / 2) ** (1 / 5.0)
            assert np.allclose(blk[..., c], want, rtol=2e-3), (p, c)  # sum(rgb*w)/(2 sum w) = L/2 up to fp32 rounding + fp16 store
        assert np.all(blk[..., 3] == 1.0)
    dq = float(np.float16(d)) - np.float32(0.01)
    dblk = f16(dep[2:18, 2:18, :])
    assert np.allclose(dblk[..., 0], dq / 2, rtol=2e-3)
    assert np.allclose(dblk[..., 1], dq * dq / 2, rtol=2e-3)
    # clamp to maxDistance for misses (60000)
    rad, dd = _uniform_rays(oracle, u, Lrgb, 60000.0, scenes.frame_rotation(1))
    oracle.blend(u, rad, dd, None, None, irr, dep, first_frame=True)
    assert np.allclose(f16(dep[2:18, 2:18, 0]), 3.0, rtol=1e-3) and np.allclose(f16(dep[2:18, 2:18, 1]), 18.0, rtol=1e-3)
    # pad rows/columns and (before the border pass) the border ring stay zero
    assert not irr[0].any() and not irr[:, 0].any() and not irr[1].any()


def test_hysteresis_closed_form_trajectory(oracle):
    """Constant input => x_n = fp16(fma(x_{n-1}, h, x*(1-h))) every frame (ProbeUpdate.glsl:144-145)."""
    u = abi.make_uniform((0, 0, 0), (1, 1, 1), (1, 1, 1), 64, hysteresis=0.9, gamma=1.0)
    rot = scenes.identity_rotation()
    rad0, dd = _uniform_rays(oracle, u, (0.0, 0.0, 0.0), 2.0, rot)
    rad1, _ = _uniform_rays(oracle, u, (1.0, 1.0, 1.0), 2.0, rot)
    a = [oracle.new_atlases(u) for _ in range(2)]
    oracle.blend(u, rad0, dd, None, None, a[1][0], a[1][1], first_frame=True)  # frame 0: black
    ping = 1
    x = np.float16(0.0)
    h = np.float32(0.9)
    new = np.float32(0.5)  # (1 * sum w) / (2 sum w), gamma 1
    for n in range(40):
        w = 1 - ping
        oracle.blend(u, rad1, dd, a[ping][0], a[ping][1], a[w][0], a[w][1], first_frame=False)
        ping = w
        x = np.float16(np.float32(math.fma(float(np.float32(x)), float(h), float(new * (np.float32(1.0) - h))))) if hasattr(math, "fma") else x
        got = f16(a[ping][0][2:10, 2:10, 0])
        if hasattr(math, "fma"):
            # all 64 texels follow the same trajectory up to the 1-ulp spread of sum(rgb w)/(2 sum w) around 0.5
            assert np.abs(got - float(x)).max() <= 2 * float(np.spacing(np.float16(x)))
    assert 0.49 < float(got.mean()) <= 0.5 + 1e-3  # converged towards 0.5, stalls within fp16 resolution


def test_naive_equals_hoisted_bitwise(oracle):
    sc = scenes.cornell_scene(res=32, counts=(2, 2, 2), rays=64, atlas_res=256)
    osc = oracle.OracleScene(sc)
    rad, dd, _, _ = osc.trace(scenes.frame_rotation(2))
    u = sc.uniform
    outs = []
    for naive in (True, False):
        irr, dep = oracle.new_atlases(u)
        oracle.blend(u, rad, dd, None, None, irr, dep, first_frame=True, naive=naive)
        irr2, dep2 = oracle.new_atlases(u)
        oracle.blend(u, rad, dd, irr, dep, irr2, dep2, first_frame=False, naive=naive)
        outs.append((irr, dep, irr2, dep2))
    for x, y in zip(*outs):
        assert np.array_equal(x, y)


# ---------------------------------------------------------------------------------------------------------------
# Border: closed-form mirror rule, the reference's own tables, idempotence
# ---------------------------------------------------------------------------------------------------------------
def test_border_offsets_equal_reference_tables(oracle):
    """tests/golden/border_offsets.json is extracted from the reference's BorderUpdate.glsl:25-133 by
    tests/golden/make_border_offsets.py; the oracle's mirror rule must reproduce both tables entry by entry."""
    golden = json.load(open(os.path.join(GOLDEN, "border_offsets.json")))
    for side, key in ((8, "irradiance"), (16, "depth")):
        buf = np.zeros((4 * side + 4, 4), dtype=np.int32)
        n = oracle.lib().oracle_border_offsets(side, buf.ctypes.data_as(C.c_void_p))
        assert n == 4 * side + 4 == len(golden[key])
        assert buf.tolist() == golden[key]


def test_border_mirror_rule_and_idempotence(oracle):
    u = abi.make_uniform((0, 0, 0), (1, 1, 1), (3, 2, 2), 32)
    rng = np.random.default_rng(5)
    irr, dep = oracle.new_atlases(u)
    for img, side in ((irr, 8), (dep, 16)):
        S = side + 2
        for p in range(12):
            px, py = p % 6, p // 6
            img[2 + py * S: 2 + py * S + side, 2 + px * S: 2 + px * S + side] = rng.integers(1, 30000, (side, side, img.shape[2]))
    before = (irr.copy(), dep.copy())
    oracle.border(u, irr, dep)
    for img, b, side in ((irr, before[0], 8), (dep, before[1], 16)):
        S = side + 2
        for p in range(12):
            ox, oy = 1 + (p % 6) * S, 1 + (p // 6) * S
            t = img[oy: oy + S, ox: ox + S]
            assert np.array_equal(t[1:-1, 1:-1], b[oy + 1: oy + S - 1, ox + 1: ox + S - 1])  # interior untouched
            assert np.array_equal(t[0, 1:-1], t[1, 1:-1][::-1])      # top row mirrors x from interior row 1
            assert np.array_equal(t[-1, 1:-1], t[-2, 1:-1][::-1])    # bottom
            assert np.array_equal(t[1:-1, 0], t[1:-1, 1][::-1])      # left column mirrors y from interior column 1
            assert np.array_equal(t[1:-1, -1], t[1:-1, -2][::-1])    # right
            assert np.array_equal(t[0, 0], t[-2, -2]) and np.array_equal(t[0, -1], t[-2, 1])
            assert np.array_equal(t[-1, 0], t[1, -2]) and np.array_equal(t[-1, -1], t[1, 1])
        assert not img[0].any() and not img[-1].any() and not img[:, 0].any() and not img[:, -1].any()  # outer pad
    again = (irr.copy(), dep.copy())
    oracle.border(u, irr, dep)
    assert np.array_equal(irr, again[0]) and np.array_equal(dep, again[1])


# ---------------------------------------------------------------------------------------------------------------
# Trace on analytic SDFs
# ---------------------------------------------------------------------------------------------------------------
def _plane_scene(res=64, D=8.0, h=-2.0, counts=(2, 1, 2), rays=256):
    """Half-space y < h is solid."""
    import torch

    xs, ys, zs = scenes.voxel_centers((0, 0, 0), D, res, "cpu")
    d = (ys[None, :, None] - h).expand(res, res, res)
    sdf = scenes.encode_sdf(d, D).contiguous()
    mip = scenes.build_mip(sdf, res, D)
    u = abi.make_uniform((-1.0, 1.0, -1.0), (2.0, 1.0, 2.0), counts, rays)
    return scenes.Scene("plane", u, scenes.make_sdf_data((0, 0, 0), D, res), sdf, mip)


def test_trace_plane_hit_distance_and_miss(oracle):
    sc = _plane_scene()
    osc = oracle.OracleScene(sc)
    rad, dd, steps, cn = osc.trace(scenes.identity_rotation(), want_steps=True)
    d = f16(dd)
    vox = sc.sdf_data.cascadeVoxelSize[0]
    dirs = d[0, :, :3]
    dist = d[0, :, 3]
    down = dirs[:, 1] < -0.5  # steep enough to reach the plane inside the cascade
    analytic = (1.0 - (-2.0)) / -dirs[down, 1]
    assert np.all(np.abs(dist[down] - analytic) <= 1.0 * vox + 2e-3 * analytic)  # within one voxel (+ fp16 storage)
    up = dirs[:, 1] > 0.05
    assert np.all(dist[up] == 60000.0)  # misses carry GLOBAL_SDF_WORLD_SIZE, exactly representable in fp16
    assert np.all(f16(rad)[0, up, :3] == 0.0)  # 1x1 black fallback sky
    assert steps.max() <= 250 and cn["hits"] > 0 and cn["mipTaps"] >= cn["steps"]


def test_trace_probe_inside_geometry_is_black_and_far(oracle):
    sc = _plane_scene(h=3.0)  # probes at y = 1 are inside the solid half-space
    osc = oracle.OracleScene(sc)
    rad, dd, _, cn = osc.trace(scenes.frame_rotation(0))
    assert np.all(f16(dd)[..., 3] == 60000.0) and not f16(rad)[..., :3].any()  # GISDFRays.comp:93-96
    assert cn["hits"] == rad.shape[0] * rad.shape[1]


def test_trace_axis_aligned_rays_are_finite(oracle):
    """lineHitAABB divides by zero for axis-aligned rays; the contract's select-based min/max keep it finite."""
    sc = _plane_scene(rays=64)
    osc = oracle.OracleScene(sc)
    out = (C.c_float * 3)()
    oracle.lib().oracle_spherical_fibonacci(10, 64, None, out)
    v = np.array(out[:], dtype=np.float64)
    # rotation taking ray 10 onto -y exactly
    tgt = np.array([0.0, -1.0, 0.0])
    axis = np.cross(v, tgt)
    ang = math.atan2(np.linalg.norm(axis), float(v @ tgt))
    rot = scenes.rotation_from_axis_angle(axis / np.linalg.norm(axis), ang)
    rad, dd, _, _ = osc.trace(rot)
    d = f16(dd)
    assert np.isfinite(d).all()
    assert abs(d[0, 10, 1] + 1.0) < 2e-3 and abs(d[0, 10, 3] - 3.0) <= 0.3


def test_cornell_frame_is_sane(oracle):
    sc = scenes.build("c1")
    p = oracle.OraclePipeline(sc)
    p.update(scenes.frame_rotation(0))
    irr, dep = f16(p.irradiance), f16(p.depth)
    assert np.isfinite(irr).all() and np.isfinite(dep).all()
    assert irr[2:-2, 2:-2, :3].mean() > 0.1
    assert 0 < dep[..., 0].max() <= 0.5 * sc.uniform.maxDistance + 1e-3  # r = mean/2 <= maxDistance/2
    c = p.counters
    assert c["mipTaps"] == c["steps"] + c["hits"]  # the hitting step breaks before step++ (SDFCommon.glsl:143-186)
    assert c["texTaps"] >= 6 * c["hits"]  # 6 normal taps per hit


def test_generic_sdf_trace_reproduces_the_probe_trace(oracle):
    """Row f4: oracle_trace_global_sdf is the same tracyGlobalSDF the probe trace calls (pinned by GISDFRays.comp.spv).  With that
    shader's arguments (bias 0, stepScale 1, maxDistance = world size) its hit times give the probe trace's stored distances, its
    normals are unit vectors; needsHitNormal = false returns a zero normal; a short maxDistance turns far hits into misses."""
    sc = scenes.cornell_scene(res=32, counts=(2, 2, 2), rays=64)
    rot = scenes.frame_rotation(3)
    osc = oracle.OracleScene(sc)
    rad, dd, _, _ = osc.trace(rot)
    dirs = dd.view(np.float16).astype(np.float32)[0, :, :3]
    # exact fp32 directions: rebuild them the way the engine / oracle do
    L = oracle.lib()
    d32 = np.zeros((sc.rays, 3), dtype=np.float32)
    import ctypes as C
    L.oracle_spherical_fibonacci.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    r16 = np.ascontiguousarray(rot, dtype=np.float32)
    for i in range(sc.rays):
        L.oracle_spherical_fibonacci(i, sc.rays, r16.ctypes.data_as(C.c_void_p), d32[i].ctypes.data_as(C.c_void_p))
    assert np.allclose(d32, dirs, atol=2e-3)
    u = sc.uniform
    traces = np.zeros(sc.probes * sc.rays, dtype=abi.SDF_TRACE_DTYPE)
    k = 0
    for p in range(sc.probes):
        o = [u.startPosition[a] + u.step[a] * c for a, c in enumerate((p % 2, (p // 2) % 2, p // 4))]
        for r in range(sc.rays):
            traces[k]["worldPosition"], traces[k]["worldDirection"] = o, d32[r]
            traces[k]["maxDistance"], traces[k]["stepScale"], traces[k]["needsHitNormal"] = abi.GLOBAL_SDF_WORLD_SIZE, 1.0, 1
            k += 1
    hits = oracle.trace_global_sdf(sc.sdf_data, sc.sdf, sc.mip, traces, 0.0)
    stored = dd.view(np.float16).astype(np.float32)[..., 3].reshape(-1)
    hit = hits["hitTime"] >= 0
    inside = hit & (hits["hitSDF"] <= 0) & (hits["hitTime"] <= sc.sdf_data.cascadeVoxelSize[0])
    want = np.where(hit & ~inside, np.maximum(hits["hitTime"] + np.float32(0.5) * np.float32(sc.sdf_data.cascadeVoxelSize[0]), 0), np.float32(abi.GLOBAL_SDF_WORLD_SIZE))
    assert np.array_equal(want.astype(np.float16), stored.astype(np.float16)) and hit.mean() > 0.9
    n = np.linalg.norm(hits["hitNormal"][hit], axis=1)
    assert np.allclose(n, 1.0, atol=1e-5)
    t2 = traces.copy()
    t2["needsHitNormal"] = 0
    t2["maxDistance"] = 3.0
    h2 = oracle.trace_global_sdf(sc.sdf_data, sc.sdf, sc.mip, t2, 0.0)
    assert (h2["hitNormal"] == 0).all() and (h2["hitTime"] >= 0).sum() < hit.sum()
    near = hit & (hits["hitTime"] < 2.0)
    assert np.array_equal(h2["hitTime"][near], hits["hitTime"][near])  # maxDistance only shortens the ray


def test_fma_blend_stays_within_one_fp16_ulp_of_the_literal_blend_over_64_frames(oracle):
    """The engine's blend takes the single-rounding reading GLSL allows (acc = fma(value, weight, acc), mix = fma(prev, h, new * (1 - h))); the
    shipped SPIR-V spells OpVectorTimesScalar + OpFAdd and x * (1 - a) + y * a.  The two feed back through fp16 atlases every frame, so the
    question is whether they drift apart: 8 x 4 x 8 probes x 256 rays, hysteresis 0.98, gamma 0.85 (the shipped scene's), 64 frames -
    never more than one fp16 ulp apart and always inside the north-star tolerance (1e-3 relative / 1e-4 absolute)."""
    sc = scenes.cornell_scene(res=32, counts=(8, 4, 8), rays=256, atlas_res=256, hysteresis=0.98, gamma=0.85)
    osc = oracle.OracleScene(sc)
    lit, fma = oracle.OraclePipeline(osc), oracle.OraclePipeline(osc)
    worst_ulp, differing = 0, 0
    try:
        for f in range(64):
            rot = scenes.frame_rotation(f)
            oracle.set_unfused(True)
            lit.update(rot)
            oracle.set_unfused(False)
            fma.update(rot)
            for a, b in ((lit.irradiance, fma.irradiance), (lit.depth, fma.depth)):
                ulp = np.abs(a.astype(np.int32) - b.astype(np.int32))  # same-sign fp16 bit patterns: difference = ulps
                same_sign = (a >> 15) == (b >> 15)
                assert same_sign[ulp > 0].all()
                worst_ulp = max(worst_ulp, int(ulp.max()))
                x, y = f16(a), f16(b)
                assert np.all(np.abs(x - y) <= 1e-4 + 1e-3 * np.abs(x)), f"frame {f}: outside the north-star tolerance"
                if f == 63:
                    differing += int((ulp > 0).sum())
    finally:
        oracle.set_unfused(False)
    assert worst_ulp <= 1, worst_ulp
    assert differing > 0  # the two readings are really different arithmetic (otherwise this test pins nothing)


def test_depth_minus_one_quirk_is_unreachable_from_fp16():
    """ProbeUpdate.glsl:75-77 replaces a ray distance of exactly -1 by maxDistance (`if (d == -1.0f)` after d = min(maxDistance, dist - 0.01)).  The
    distances come from an RGBA16F image, and no fp16 value gives dist - 0.01f == -1.0f in binary32, so the branch is dead for every input the path
    can see (unless maxDistance itself is -1).  The bit-exact kernels keep the statement; the tensor-core producer (blend_umma.inc) relies on this."""
    with np.errstate(invalid="ignore"):
        h = np.arange(65536, dtype=np.uint16).view(np.float16).astype(np.float32)
        d = h - np.float32(0.01)
    assert int((d == np.float32(-1.0)).sum()) == 0
