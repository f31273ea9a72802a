"""The C-ABI library loads, exports every symbol include/luxddgi.h declares, keeps the reference's struct layouts,
and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

from luxgi_b200 import abi, ddgi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "luxddgi.h")).read()
    return sorted(set(re.findall(r"LUX_API\s+[\w\s\*]+?\b(lux_ddgi_\w+)\s*\(", text)))


def test_header_declares_what_python_binds():
    assert declared_symbols() == sorted(ddgi.EXPORTS)


def test_library_exports_every_declared_symbol(engine_lib):
    for name in declared_symbols():
        assert hasattr(engine_lib, name), name
    assert engine_lib.lux_ddgi_version() == 0x00010000


def test_struct_layouts_match_reference_blocks():
    # DDGICommon.glsl:11-31 (scalar layout)
    U = abi.DDGIUniform
    assert (U.startPosition.offset, U.step.offset, U.probeCounts.offset, U.maxDistance.offset) == (0, 16, 32, 48)
    assert (U.sharpness.offset, U.hysteresis.offset, U.normalBias.offset, U.ddgiGamma.offset) == (52, 56, 60, 64)
    assert (U.irradianceProbeSideLength.offset, U.irradianceTextureWidth.offset, U.irradianceTextureHeight.offset) == (68, 72, 76)
    assert (U.depthProbeSideLength.offset, U.depthTextureWidth.offset, U.depthTextureHeight.offset, U.raysPerProbe.offset) == (80, 84, 88, 92)
    # GlobalSDFData.glsl:4-12 (std140)
    S = abi.GlobalSDFData
    assert (S.cascadePosDistance.offset, S.cascadeVoxelSize.offset, S.cascadesCount.offset, S.resolution.offset) == (0, 64, 80, 84)
    # AtlasCommon.glsl:8-32
    O = abi.ObjectBuffer
    assert (O.objectBounds.offset, O.tileOffset.offset, O.padding.offset, O.transform.offset, O.extends.offset) == (0, 16, 40, 48, 112)
    T = abi.TileBuffer
    assert (T.extends.offset, T.transform.offset, T.objectBounds.offset) == (0, 16, 80)
    A = abi.GlobalSurfaceAtlasData
    assert (A.cameraPos.offset, A.chunkSize.offset, A.culledObjectsCapacity.offset, A.resolution.offset, A.objectsCount.offset) == (0, 12, 16, 20, 24)
    assert C.sizeof(abi.TracePushConstants) == 80


def test_uniform_from_volume_matches_on_game_start(engine_lib):
    # the shipped scene's volume: probeDistance 24, AABB [-158.55,44.02,-129.94]..[-4.25,107.69,-24.39] -> 8x4x6
    v = abi.IrradianceVolume(24.0, 1, 256, 0.98, 1.2, 0.1, 50.0, 0.85)
    u = ddgi.uniform_from_volume(v, (-158.55, 44.02, -129.94), (-4.25, 107.69, -24.39))
    assert list(u.probeCounts)[:3] == [8, 4, 6]
    assert u.maxDistance == pytest.approx(36.0)
    assert u.raysPerProbe == 256
    assert (u.irradianceTextureWidth, u.irradianceTextureHeight) == (10 * 32 + 2, 10 * 6 + 2)
    assert (u.depthTextureWidth, u.depthTextureHeight) == (18 * 32 + 2, 18 * 6 + 2)
    assert u.startPosition[0] == pytest.approx(-158.55)


def test_invalid_arguments_return_status_codes(engine_lib):
    h = C.c_void_p()
    assert engine_lib.lux_ddgi_create(None, None, C.byref(h)) == -1
    u = abi.make_uniform((0, 0, 0), (1, 1, 1), (2, 2, 2), 32)
    u.irradianceTextureWidth += 1
    assert engine_lib.lux_ddgi_create(C.byref(u), None, C.byref(h)) == -1
    assert b"atlas sizes" in engine_lib.lux_ddgi_last_error()
    u = abi.make_uniform((0, 0, 0), (1, 1, 1), (2, 2, 3), 32)
    info = abi.CreateInfo(0, 0, 2, 0, None)
    assert engine_lib.lux_ddgi_create(C.byref(u), C.byref(info), C.byref(h)) == -1  # world must divide Z


def test_no_gpu_means_error_not_fallback(engine_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    u = abi.make_uniform((0, 0, 0), (1, 1, 1), (2, 2, 2), 32)
    with pytest.raises(ddgi.LuxError) as e:
        ddgi.DDGIPipeline(u)
    assert e.value.code == -4  # LUX_ERR_NO_DEVICE


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "luxgi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f in ("scenes.py",), f"{f} mentions the oracle"


def test_header_is_plain_c_and_cpp(tmp_path):
    """include/luxddgi.h is the boundary a C or C++ host binds: it must compile on its own as C99 and as C++17, and the struct sizes the
    Python layer assumes must be the compiler's."""
    import shutil
    import subprocess

    inc = os.path.join(ROOT, "include")
    src = tmp_path / "abi_check.c"
    src.write_text('#include "luxddgi.h"\n#include <stdio.h>\nint main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(LuxDDGIUniform), '
                   'sizeof(LuxTracePushConstants), sizeof(LuxGlobalSDFData), sizeof(LuxGlobalSurfaceAtlasData), sizeof(LuxObjectBuffer), sizeof(LuxTileBuffer), '
                   'sizeof(LuxObjectRasterizeData), sizeof(LuxGlobalSDFTrace) + sizeof(LuxGlobalSDFHit)); return 0; }\n')
    cc = shutil.which("gcc") or "/usr/bin/gcc"
    exe = tmp_path / "abi_check"
    subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout.split()]
    assert sizes == [96, 80, 96, 32, 128, 96, 176, 40 + 28]
    cxx = shutil.which("g++") or "/usr/bin/g++"
    cpp = tmp_path / "abi_check.cpp"
    cpp.write_text('#include "luxddgi.h"\nstatic_assert(sizeof(LuxDDGIUniform) == 96, "DDGIUniform");\nint main() { return sizeof(LuxMeshSDF) == 0; }\n')
    subprocess.run([cxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", inc, str(cpp)], check=True)
