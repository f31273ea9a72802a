"""Round-2 experiment, NOT collected by pytest (run it by hand on a B200):

    python tests/experimental/open_skip_check.py [--bench c4] [--bench c5]

LUX_DDGI_FLAG_OPEN_SKIP: the wavefront march looks each step's position up in a conservative two-bit table over the mip volume (one cell per 8x8x8
mip texels).  OPEN = every texel a trilinear tap in that cell can touch is >= chunkSizeDistance * (1 + 2^-10): the step takes the reference's
`stepDistance = chunkSizeDistance` branch with no tap at all.  NEAR = every such texel is < chunkSizeDistance * (1 - 2^-10): only the
full-resolution tap is taken, the mip tap only when that one is >= 2 * margin.  Results must be bit-identical.
The CPU side of the argument is already checked (tests/test_oracle_kat.py::test_open_space_table_is_conservative: 0 violations; C4: 10.8 % of the
march steps are open and 72 % need no mip tap; C5: 44.6 % open and 36 % no mip tap).

What this script does: builds libluxddgi_experimental.so (-DLUX_EXPERIMENTAL_OPEN_SKIP; the shipped library stays untouched), checks parity of
every small configuration against the oracle with the flag on, and with --bench times the update with and without the flag (same library, same
process, alternating) so that the variant can be adopted or dropped on a measurement."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from luxgi_b200 import build as lux_build  # noqa: E402

os.environ["LUX_DDGI_LIB"] = lux_build.build_experimental()

import numpy as np  # noqa: E402

from luxgi_b200 import abi, ddgi, scenes  # noqa: E402
from oracle import binding as oracle  # noqa: E402


def parity(name, sc, frames=2):
    rots = [scenes.frame_rotation(f) for f in range(frames)]
    orc = oracle.OraclePipeline(sc)
    for r in rots:
        orc.update(r)
    for flags in (abi.FLAG_OPEN_SKIP, abi.FLAG_OPEN_SKIP | abi.FLAG_SDF_LOADS, abi.FLAG_OPEN_SKIP | abi.FLAG_SHADE_UNSORTED):
        pipe = ddgi.DDGIPipeline(sc.uniform, flags=flags)
        pipe.set_scene(sc)
        for r in rots:
            pipe.update(r)
        pipe.synchronize()
        ok = (np.array_equal(pipe.radiance, orc.rad) and np.array_equal(pipe.direction_distance, orc.dd) and np.array_equal(pipe.irradiance, orc.irradiance)
              and np.array_equal(pipe.depth, orc.depth))
        print(f"parity {name} flags {flags:#x}: {'bit-identical' if ok else 'MISMATCH'}", flush=True)
        pipe.close()
        if not ok:
            raise SystemExit(1)


def bench(workload, steps=8, warmup=3):
    import torch

    sc = scenes.build(workload, device="cuda")
    out = {}
    pipes = {name: ddgi.DDGIPipeline(sc.uniform, flags=flags | abi.FLAG_STAGE_TIMERS) for name, flags in (("shipped", 0), ("open_skip", abi.FLAG_OPEN_SKIP))}
    for p in pipes.values():
        p.set_scene(sc)
    for rep in range(2):  # alternate so that clocks / thermals hit both alike
        for name, p in pipes.items():
            ms = {"march": 0.0, "trace": 0.0, "total": 0.0}
            for f in range(warmup + steps):
                p.update(scenes.frame_rotation(f))
                if f >= warmup:
                    p.synchronize()
                    t = p.stage_ms()
                    ms["march"] += t.march_ms / steps; ms["trace"] += t.trace_ms / steps; ms["total"] += t.total_ms / steps
            out[(name, rep)] = ms
            print(workload, name, rep, {k: round(v, 3) for k, v in ms.items()}, flush=True)
    same = all(np.array_equal(getattr(pipes["shipped"], b), getattr(pipes["open_skip"], b)) for b in ("irradiance", "depth"))
    print(workload, "atlases identical between the two variants after", 2 * (warmup + steps), "frames:", same)
    for p in pipes.values():
        p.close()
    torch.cuda.synchronize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bench", action="append", default=[])
    args = ap.parse_args()
    t0 = time.time()
    parity("c1", scenes.build("c1"))
    parity("city64", scenes.build("city64"))
    parity("city128", scenes.build("city128", counts=(16, 16, 16)))
    parity("cornell 2 cascades", scenes.cornell_scene(res=64, counts=(8, 4, 8), rays=96, atlas_res=256, cascades=2), frames=3)
    parity("cornell 32^3 (mip 8^3: one cell per axis)", scenes.cornell_scene(res=32, counts=(3, 5, 2), rays=50, atlas_res=256))
    print(f"parity done in {time.time() - t0:.1f} s")
    for w in args.bench:
        bench(w)


if __name__ == "__main__":
    main()
