"""Quick GPU sanity of LUX_DDGI_FLAG_OPEN_SKIP (see open_skip_check.py): the experimental march against the shipped march of the SAME library on a
city with open sky, bit for bit.  Needs luxgi_b200/libluxddgi_experimental.so (python -m luxgi_b200.build --experimental)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["LUX_DDGI_LIB"] = os.path.join(ROOT, "luxgi_b200", "libluxddgi_experimental.so")

import numpy as np  # noqa: E402

from luxgi_b200 import abi, ddgi, scenes  # noqa: E402

t0 = time.time()
name = sys.argv[1] if len(sys.argv) > 1 else "city128"
sc = scenes.build(name, device="cuda", **({"counts": (16, 16, 16)} if name == "city128" else {}))
print("scene", name, round(time.time() - t0, 1), "s", flush=True)
res = {}
for tag, flags in (("shipped", abi.FLAG_STAGE_TIMERS), ("open_skip", abi.FLAG_STAGE_TIMERS | abi.FLAG_OPEN_SKIP)):
    p = ddgi.DDGIPipeline(sc.uniform, flags=flags)
    p.set_scene(sc)
    ms = []
    for f in range(4):
        p.update(scenes.frame_rotation(f))
        p.synchronize()
        ms.append(round(p.stage_ms().march_ms, 4))
    res[tag] = (p.radiance, p.direction_distance, p.irradiance, p.depth)
    print(tag, "march ms per frame", ms, flush=True)
    p.close()
same = [bool(np.array_equal(a, b)) for a, b in zip(res["shipped"], res["open_skip"])]
print("radiance / direction_distance / irradiance / depth identical:", same, round(time.time() - t0, 1), "s")
sys.exit(0 if all(same) else 1)
