"""Randomised engine-vs-oracle parity (collected last on purpose: everything deterministic runs first).

The same scene families the oracle is fuzzed on against the reference's shipped SPIR-V (tests/golden/fuzz_oracle_vs_spirv.py: the Cornell room
with 1-4 cascades, open cities with random probe placement under a 2x2-texel sky, probe grids pushed into geometry and onto cascade faces), with
fixed seeds, two frames each, texture and load SDF paths, tiled and list form of the FP32 blend, row and beam shape of the march.  Ray buffers and
atlases must be bit-identical to the oracle's."""
import numpy as np
import pytest

from luxgi_b200 import abi, ddgi, scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2])
def test_fuzz_engine_vs_oracle(oracle, seed):
    from tests.golden.fuzz_oracle_vs_spirv import random_scene

    rng = np.random.default_rng(seed)
    for k in range(9):
        sc = random_scene(rng, k)
        rots = [scenes.frame_rotation(int(rng.integers(0, 100000))) for _ in range(2)]
        orc = oracle.OraclePipeline(sc)
        for r in rots:
            orc.update(r)
        for flags in (0, abi.FLAG_SDF_LOADS, abi.FLAG_BLEND_LISTS | abi.FLAG_MARCH_ROWS, abi.FLAG_BLEND_LISTS | abi.FLAG_MARCH_BEAMS | abi.FLAG_SDF_LOADS):
            pipe = ddgi.DDGIPipeline(sc.uniform, flags=flags)
            pipe.set_scene(sc)
            for r in rots:
                pipe.update(r)
            pipe.synchronize()
            for name, got, want in (("radiance", pipe.radiance, orc.rad), ("direction_distance", pipe.direction_distance, orc.dd),
                                    ("irradiance", pipe.irradiance, orc.irradiance), ("depth", pipe.depth, orc.depth)):
                assert np.array_equal(got, want), f"seed {seed} config {k} flags {flags}: {name} differs in {(got != want).sum()} of {got.size} fp16 values"
            pipe.close()
