#!/usr/bin/env python
"""bench.py — SDF probe rays/s and ms per DDGI volume update (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c4|c3|c2|c1|city128] [--impl lux|reference]

One "step" = one DDGI volume update (trace + blend + border [+ all-gather of the updated atlases at N > 1]) of the
named workload (default c5: BASELINE.json quotes its metric on no single configuration, so the bench line is the largest one that fits a GPU).  N > 1 is launched by torchrun (one rank per GPU); the probe volume is sharded into z-slabs, the SDF
and the surface cache are replicated, the updated atlas rows are exchanged with one in-place NCCL all-gather per atlas
per step.  Total work is fixed as N grows => "scaling": "strong".

Printed keys beyond the base contract:
  ms_per_update          the second half of the BASELINE metric
  stage_ms               setup / trace / blend of the last timed step (CUDA events on the engine's stream)
  roofline               dominant kernel (trace): algorithmic HBM bytes per launch / measured launch time vs measured HBM peak
  roofline_blend         the same for the blend stage (HBM bytes and FP32 FMA rate)
  cpu_baseline           the CPU oracle (a port of the reference shaders) timed on this box's host cores on a bounded sample
  e2e                    the same metric through the C ABI with HOST buffers: per step the light cache is uploaded from pinned
                         memory and both updated atlases are read back into pinned memory

--impl reference times the reference's algorithm on the host CPU (the oracle port; the reference's own shaders cannot run
here: no Vulkan/glslang in the image) on a bounded probe sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "sdf_probe_rays_per_s"
UNIT = "rays/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def workload_desc(name, sc):
    """The `config` object: the same keys and values in both arms (the driver compares them)."""
    u = sc.uniform
    return {"workload": f"{name}: {sc.name} SDF {int(sc.sdf_data.resolution)}^3 fp16, {u.probeCounts[0]}x{u.probeCounts[1]}x{u.probeCounts[2]} probes, "
                        f"{u.raysPerProbe} rays/probe",
            "probes": sc.probes, "rays_per_probe": u.raysPerProbe, "sdf_res": int(sc.sdf_data.resolution),
            "surface_atlas_res": int(sc.atlas_data.resolution) if sc.atlas_data is not None else 0,
            "l2_policy": "inputs larger than L2 (no flush)"}


def algorithmic_bytes(sc, probes):
    """Compulsory HBM bytes per launch.
    trace (whole stage): two RGBA16F stores per ray + SDF + mip read once + both surface atlases once (upper bound);
    march: SDF + mip read once + one 20-byte record written per ray;   shade: record read + two RGBA16F stores per ray + SDF once
    (normal taps) + both surface atlases once;   blend+border: ray buffers read once + previous interiors + new interiors and borders."""
    R = sc.uniform.raysPerProbe
    atlas = (int(sc.atlas_data.resolution) ** 2) * (8 + 4) if sc.atlas_data is not None else 0
    sdf = sc.sdf.numel() * 2
    return {"trace": 16 * probes * R + sc.sdf_bytes() + atlas,
            "march": 20 * probes * R + sc.sdf_bytes(),
            "shade": (20 + 16) * probes * R + sdf + atlas,
            "blend": 16 * probes * R + probes * (64 * 8 + 256 * 4) + probes * (100 * 8 + 324 * 4)}


def measured_traffic(workload, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/),
    valid only for the configuration they were taken on."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
        return t.get(f"{workload}@{world}", {})
    except Exception:
        return {}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on host cores
# ----------------------------------------------------------------------------------------------------------------------
def oracle_sample_ids(sc, n):
    """n probes stratified over the whole volume: one per stratum of P / n consecutive ids, at a pseudo-random offset inside the stratum
    (a plain stride of P / n is a multiple of the grid's x extent on the bench volumes, i.e. a sample of ONE face of the volume,
    whose rays leave the SDF early: round 1's sample under-counted the march steps per ray by a third)."""
    P = sc.probes
    n = min(n, P)
    i = np.arange(n, dtype=np.int64)
    lo = i * P // n
    width = np.maximum((i + 1) * P // n - lo, 1)
    off = (i * 2654435761 + 40503) % width
    return (lo + off).astype(np.int32)


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def time_oracle(sc, rot, sample_probes, steps, warmup):
    """One step = trace + literal (naive) blend + border of `sample_probes` stratified probes on all host threads."""
    from oracle import binding as ob

    ob.lib().oracle_set_threads(host_threads())  # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every core it may run on regardless
    osc = ob.OracleScene(sc)
    ids = oracle_sample_ids(sc, sample_probes)
    n = len(ids)
    R = sc.uniform.raysPerProbe
    # blend timing runs on a compact n-probe volume with the same rays per probe (its cost is independent of probe position)
    from luxgi_b200 import abi

    side = 1
    while side * side < n:
        side *= 2
    ub = abi.make_uniform((0, 0, 0), (1, 1, 1), (side, max(1, n // side), 1), R, max_distance=sc.uniform.maxDistance,
                          sharpness=sc.uniform.sharpness, hysteresis=sc.uniform.hysteresis, gamma=sc.uniform.ddgiGamma)
    nb = abi.probe_count(ub)
    irr = [ob.new_atlases(ub)[0] for _ in range(2)]
    dep = [ob.new_atlases(ub)[1] for _ in range(2)]
    times, t_trace, t_blend = [], 0.0, 0.0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        rad, dd, _, counters = osc.trace(rot, probe_ids=ids)
        t1 = time.perf_counter()
        ob.blend(ub, rad[:nb], dd[:nb], irr[0], dep[0], irr[1], dep[1], first_frame=(it == 0), naive=True)
        ob.border(ub, irr[1], dep[1])
        t2 = time.perf_counter()
        irr.reverse(); dep.reverse()
        if it >= warmup:
            times.append(t2 - t0); t_trace += t1 - t0; t_blend += t2 - t1
    total = sum(times)
    return {"rays_per_s": n * R * len(times) / total, "sec_per_step": total / len(times), "trace_frac": t_trace / total,
            "sample_probes": int(n), "rays": int(n * R), "threads": ob.lib().oracle_threads(), "counters": counters}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from luxgi_b200 import scenes

    dev = "cuda" if torch.cuda.is_available() else "cpu"  # the fixture generator may use the GPU; the timed code is CPU only
    sc = scenes.build(args.workload, device=dev)
    rot = scenes.frame_rotation(0)
    sample = args.reference_probes
    r = time_oracle(sc, rot, sample, args.steps, args.warmup)
    full_ms = sc.probes * sc.uniform.raysPerProbe / r["rays_per_s"] * 1e3
    line = {"impl": "reference", "metric": METRIC, "value": r["rays_per_s"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_desc(args.workload, sc),
            "ms_per_update_extrapolated": full_ms,
            "cpu_baseline": {"value": r["rays_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                             "sample": f"{r['sample_probes']} stratified probes x {sc.uniform.raysPerProbe} rays per step "
                                       f"(trace + literal blend + border), OpenMP over probes; trace share {r['trace_frac']:.2f}"},
            "e2e": {"value": r["rays_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def multi_gpu_parity(sc, local, rank, world, flags, stream, dev, frames=2, sample=256):
    """N > 1 only, after the timed region: fresh contexts run `frames` updates with the in-library all-gather; every rank checksums its FULL
    atlases on the device (they must be identical on all ranks: own rows + the rows NCCL delivered), and rank 0 replays a stratified probe
    subsample of the WHOLE volume on the CPU oracle and compares those probes' atlas tiles (interior + border) bit for bit."""
    import torch
    import torch.distributed as dist

    from luxgi_b200 import abi, ddgi, nccl, scenes
    from oracle import binding as ob

    u = sc.uniform
    pipe = ddgi.DDGIPipeline(u, device=local, rank=rank, world=world, flags=flags, stream=stream.cuda_stream)
    pipe.set_scene(sc)
    comm = nccl.NcclComm(rank, world, local)
    pipe.set_nccl_comm(comm.ptr)
    ids = oracle_sample_ids(sc, sample)
    if rank == 0:
        ob.lib().oracle_set_threads(host_threads())
        osc = ob.OracleScene(sc)
        irr = [ob.new_atlases(u)[0] for _ in range(2)]
        dep = [ob.new_atlases(u)[1] for _ in range(2)]
    for f in range(frames):
        rot = scenes.frame_rotation(f)
        pipe.update(rot)
        if rank == 0:
            rad, dd, _, _ = osc.trace(rot, probe_ids=ids)
            ob.blend_ids(u, rad, dd, irr[f % 2], dep[f % 2], irr[1 - f % 2], dep[1 - f % 2], first_frame=(f == 0), probe_ids=ids)
    pipe.synchronize()
    sums = []
    for buf in (abi.BUF_IRRADIANCE, abi.BUF_DEPTH):
        t = torch.as_tensor(pipe.device_view(buf), device=dev).reshape(-1).view(torch.int16).to(torch.int64)
        w = (torch.arange(t.numel(), device=dev, dtype=torch.int64) % 65521) + 1
        sums += [int(t.sum().item()), int((t * w).sum().item() & 0x7fffffffffffffff)]
    mine = torch.tensor(sums, device=dev, dtype=torch.int64)
    every = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    same = all(bool(torch.equal(e, every[0])) for e in every)
    ok_oracle, bad = True, 0
    if rank == 0:
        per_row = u.probeCounts[0] * u.probeCounts[1]
        for got, want, side in ((pipe.irradiance, irr[frames % 2], 8), (pipe.depth, dep[frames % 2], 16)):
            S = side + 2
            for p_ in ids:
                r0, c0 = 1 + (int(p_) // per_row) * S, 1 + (int(p_) % per_row) * S
                bad += int((got[r0:r0 + S, c0:c0 + S] != want[r0:r0 + S, c0:c0 + S]).sum())
        ok_oracle = bad == 0
    pipe.close()
    comm.destroy()
    return {"status": "ok" if (same and ok_oracle) else "FAILED", "frames": frames, "atlas_checksums_equal_on_all_ranks": same,
            "probes_checked_against_oracle": int(len(ids)), "differing_fp16_values": bad}


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def run_lux(args):
    import torch
    import torch.distributed as dist

    from luxgi_b200 import abi, ddgi, scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, (world, args.gpus)

    sc = scenes.build(args.workload, device=dev)
    u = sc.uniform
    if u.probeCounts[2] % world:
        raise SystemExit(f"{world} GPUs do not divide Z={u.probeCounts[2]}")
    stream = torch.cuda.Stream(device=dev)
    flags = {"texture": 0, "loads": abi.FLAG_SDF_LOADS, "simple": abi.FLAG_TRACE_SIMPLE}[args.trace]
    if args.no_pipeline:
        flags |= abi.FLAG_NO_PIPELINE
    if args.unsorted:
        flags |= abi.FLAG_SHADE_UNSORTED
    flags |= args.extra_flags
    shards = args.shards
    if shards == "auto":  # blocks of 8 z-layers dealt out round-robin from 4 GPUs on (measured on C5: 33.9 -> 33.2 ms at N = 4, 17.7 -> 17.3 ms at N = 8; at
        shards = "blocks8" if world >= 4 and u.probeCounts[2] % (world * 8) == 0 else "slabs"  # N = 2 the two slabs are evenly loaded already)
    if shards != "slabs" and args.allgather == "lib":  # z-layers dealt out in blocks (no effect at N = 1)
        flags |= abi.flag_shard_blocks({"interleaved": 0, "blocks2": 1, "blocks4": 2, "blocks8": 3}[shards])
    shard_rank, shard_world = (rank, world) if args.emulate_shard is None else tuple(int(x) for x in args.emulate_shard.split('/'))
    # The measured pipe runs lux_ddgi_update as shipped (blend weights on a second stream during the march, no stage events); the per-stage
    # times and the kernel roofline come from a second, serialized pass below (LUX_DDGI_FLAG_STAGE_TIMERS = one batch, one stream).
    pipe = ddgi.DDGIPipeline(u, device=local, rank=shard_rank, world=shard_world, flags=flags, stream=stream.cuda_stream)
    pipe.set_scene(sc)
    comm = None
    if world > 1 and args.allgather == "lib":  # the exchange step behind the C ABI: lux_ddgi_update all-gathers the rows itself
        from luxgi_b200 import nccl

        comm = nccl.NcclComm(rank, world, local)
        pipe.set_nccl_comm(comm.ptr)
    st = pipe.state()
    P, R = sc.probes, u.raysPerProbe
    l2_gbs = pipe.measure_l2_read_bandwidth() if rank == 0 else None  # denominator of the request-level roofline; ~0.1 s, outside every timed region
    rot_cache = {}

    def rot_of(f):
        if f not in rot_cache:
            rot_cache[f] = scenes.frame_rotation(f)
        return rot_cache[f]

    views = {}

    def atlas_rows(buf, row_begin, rows_total):
        ptr, _ = pipe.buffer_ptr(buf)
        if (buf, ptr) not in views:
            t = torch.as_tensor(pipe.device_view(buf), device=dev)
            views[(buf, ptr)] = t
        t = views[(buf, ptr)]
        return t[1:1 + rows_total].reshape(-1), t[row_begin:row_begin + rows_total // world].reshape(-1)

    ag_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    ag_done = []

    def step(f):
        """One volume update.  At N > 1 the all-gather of the atlases written by step f runs on its own stream and overlaps the
        trace of step f+1 (the trace never reads the atlases); step f+2 is the first to overwrite the rows it sends from, so
        that step waits for it."""
        if world > 1 and comm is None and not args.sync_allgather and len(ag_done) >= 2:
            stream.wait_event(ag_done[-2])
        pipe.update(rot_of(f))
        if world > 1 and comm is None:  # --allgather torch: the same exchange issued by the host through torch.distributed
            side = stream if args.sync_allgather else ag_stream
            if side is not stream:
                ev = torch.cuda.Event()
                ev.record(stream)
                side.wait_event(ev)
            with torch.cuda.stream(side):
                full, mine = atlas_rows(abi.BUF_IRRADIANCE, st.irradianceRowBegin, st.irradianceRowCount * world)
                dist.all_gather_into_tensor(full, mine)
                full, mine = atlas_rows(abi.BUF_DEPTH, st.depthRowBegin, st.depthRowCount * world)
                dist.all_gather_into_tensor(full, mine)
                done = torch.cuda.Event()
                done.record(side)
            ag_done.append(done)
            del ag_done[:-2]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------------------------------------------------
    f = 0
    for _ in range(args.warmup):
        step(f); f += 1
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = pipe.state().kernelLaunches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        step(f); f += 1
    with torch.cuda.stream(stream):
        e1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    ms_total = e0.elapsed_time(e1)
    launches = pipe.state().kernelLaunches - launches0
    clocks = sampler.stop() if rank == 0 else None
    P_timed = st.probeCount * world  # == P unless --emulate-shard

    e2e_value, e2e_s, light_bytes, d2h_bytes, region_ms = None, None, 0, 0, None
    if not args.no_e2e:
        # ---- end to end through the C ABI with host buffers --------------------------------------------------------------
        light_bytes = int(sc.atlas_data.resolution) ** 2 * 8
        pin_light = torch.empty(light_bytes, dtype=torch.uint8).pin_memory()
        pin_light.copy_(sc.light.reshape(-1).view(torch.uint8).cpu())
        irr_row_bytes, dep_row_bytes = u.irradianceTextureWidth * 8, u.depthTextureWidth * 4
        atlas_res = int(sc.atlas_data.resolution)
        light_row_bytes, light_rows = atlas_res * 8, atlas_res // world
        light_row0 = rank * light_rows
        if comm is not None and atlas_res % world:
            raise SystemExit("surface atlas rows do not divide by the number of GPUs")
        # two sets of pinned result buffers: frame f's rows land while frame f+1 computes; the host waits for frame f-1's copies
        # before it issues frame f+1, i.e. it receives EVERY frame's atlases, one frame late (how a renderer consumes them)
        pin_irr = [torch.empty(st.irradianceRowCount * irr_row_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
        pin_dep = [torch.empty(st.depthRowCount * dep_row_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
        fences = []

        def e2e_step(f):
            if not args.e2e_skip_h2d:  # H2D of this frame's light cache through the C ABI
                if comm is not None:     # sharded: every rank uploads its 1/N of the rows, the library all-gathers them over NVLink
                    pipe.update_surface_light_cache_rows_ptr(pin_light.data_ptr() + light_row0 * light_row_bytes, light_row0, light_rows)
                else:
                    pipe.update_surface_light_cache_ptr(pin_light.data_ptr())
            step(f)
            k = f & 1
            if not args.e2e_skip_d2h:
                pipe.download_shard_async_ptr(abi.BUF_IRRADIANCE, pin_irr[k].data_ptr())  # the shard's own rows, packed (one strided copy)
                pipe.download_shard_async_ptr(abi.BUF_DEPTH, pin_dep[k].data_ptr())
            fences.append(pipe.download_fence())
            if len(fences) > 1:
                pipe.wait_fence(fences.pop(0))  # frame f-1 is on the host now

        def e2e_drain():
            while fences:
                pipe.wait_fence(fences.pop(0))
            pipe.synchronize()

        for _ in range(max(1, args.warmup // 2)):
            e2e_step(f); f += 1
        e2e_drain()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step(f); f += 1
        e2e_drain()  # the last frame's atlases are on the host inside the timed region too
        barrier()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        e2e_value = P * R * args.steps / e2e_s


        d2h_bytes = int(pin_irr[0].numel() + pin_dep[0].numel())
        # per-frame boundary cost of a scene change: one dirty 32^3 rasterize chunk patched into the bound volume + the cascade's mip rebuilt
        # (lux_ddgi_update_global_sdf_region; the texels are the volume's own, so results do not change)
        res = int(sc.sdf_data.resolution)
        if res >= 64 and rank == 0:
            box = sc.sdf.reshape(res, res, -1)[32:64, 32:64, 32:64].contiguous().cpu().numpy().view(np.uint16)
            pipe.update_global_sdf_region(0, (1, 1, 1), (1, 1, 1), box, rebuild_mip=True)
            pipe.synchronize()
            t0 = time.perf_counter()
            for _ in range(8):
                pipe.update_global_sdf_region(0, (1, 1, 1), (1, 1, 1), box, rebuild_mip=True)
            pipe.synchronize()
            region_ms = (time.perf_counter() - t0) / 8 * 1e3
    comm_used = comm is not None
    # ---- serialized stage pass: same workload, one batch on one stream, CUDA events around every stage -----------------------
    pipe.close()
    if comm is not None:
        comm.destroy()
    views.clear()
    pipe = ddgi.DDGIPipeline(u, device=local, rank=shard_rank, world=shard_world, flags=flags | abi.FLAG_STAGE_TIMERS, stream=stream.cuda_stream)
    pipe.set_scene(sc)
    stage = {"setup": 0.0, "trace": 0.0, "blend": 0.0, "march": 0.0, "shade": 0.0}
    for i in range(2 + args.steps):
        pipe.update(rot_of(f)); f += 1
        if i >= 2:
            pipe.synchronize()
            t = pipe.stage_ms()
            stage["setup"] += t.setup_ms; stage["trace"] += t.trace_ms; stage["blend"] += t.blend_ms
            stage["march"] += t.march_ms; stage["shade"] += t.shade_ms
    torch.cuda.synchronize()
    tm = torch.tensor([ms_total, stage["trace"], stage["blend"], stage["setup"], stage["march"], stage["shade"]], device=dev, dtype=torch.float64)
    per_rank = None
    if world > 1:
        allr = [torch.zeros_like(tm) for _ in range(world)]
        dist.all_gather(allr, tm)
        per_rank = [[round(float(x) / args.steps, 4) for x in t.tolist()[1:3]] for t in allr]
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_total, trace_ms, blend_ms, setup_ms, march_ms, shade_ms = [float(x) for x in tm.tolist()]
    # ---- opt-in tensor-core blend (LUX_DDGI_FLAG_BLEND_TC, tolerance path): its stage time on the same workload, reported beside the line ------
    blend_tc = None
    if world == 1 and not args.no_tc_ab and not (flags & abi.FLAG_BLEND_TC):
        pipe.close()
        pipe = ddgi.DDGIPipeline(u, device=local, rank=shard_rank, world=shard_world, flags=flags | abi.FLAG_BLEND_TC | abi.FLAG_STAGE_TIMERS,
                                 stream=stream.cuda_stream)
        pipe.set_scene(sc)
        n_tc, tc_blend, tc_total = min(args.steps, 4), 0.0, 0.0
        for i in range(2 + n_tc):
            pipe.update(rot_of(f)); f += 1
            if i >= 2:
                pipe.synchronize()
                t = pipe.stage_ms()
                tc_blend += t.blend_ms; tc_total += t.setup_ms + t.trace_ms + t.blend_ms
        blend_tc = {"blend_ms": tc_blend / n_tc, "update_ms_stage_sum": tc_total / n_tc, "steps": n_tc,
                    "kernels": "umma::blend_irradiance_umma_kernel+umma::blend_depth_umma_kernel (tcgen05.mma, TMEM accumulators, TMA operands)",
                    "note": "NOT part of value / e2e: opt-in path held to the north-star tolerance (1e-3 rel / 1e-4 abs), the default blend is the bit-exact FP32 one"}
    parity = None
    if world > 1 and args.emulate_shard is None and not args.no_parity:
        pipe.close()
        pipe = None
        parity = multi_gpu_parity(sc, local, rank, world, flags, stream, dev)
    ms_per_step = ms_total / args.steps
    value = P_timed * R * args.steps / (ms_total * 1e-3)

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        hbm = float(peaks["hbm_gbs"])
        probes_rank = st.probeCount
        stages_b = algorithmic_bytes(sc, probes_rank)
        march_launch_ms, shade_launch_ms = march_ms / args.steps, shade_ms / args.steps
        trace_launch_ms = trace_ms / args.steps
        blend_launch_ms = blend_ms / args.steps
        traffic = measured_traffic(args.workload, world)

        def roof(kernel, nbytes, ms):
            a = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else None
            return {"bound": "hbm", "kernel": kernel, "achieved": a, "peak": hbm, "unit": "GB/s", "frac": (a / hbm) if a else None,
                    "traffic": traffic.get(kernel), "peak_source": peak_kind, "algorithmic_bytes_per_launch": nbytes, "ms_per_launch": ms}

        def tsum(*names):
            v = [traffic.get(n) for n in names]
            return None if any(x is None for x in v) else float(sum(v))

        shade_names = ("shade_kernel",) if args.unsorted else ("classify", "scatter", "shade_sorted")
        traffic = dict(traffic)
        traffic["shade_stage"] = tsum(*shade_names)
        traffic["trace_stage"] = tsum("march_kernel", *shade_names)
        traffic["blend_stage"] = tsum("blend_irradiance", "blend_depth")
        dominant = "march_kernel" if march_launch_ms >= shade_launch_ms else "shade_stage"
        if march_launch_ms == 0.0:  # --trace simple: one kernel
            dominant = "trace_stage"
        roofs = {"march_kernel": roof("march_kernel", stages_b["march"], march_launch_ms),
                 "shade_stage": roof("shade_stage", stages_b["shade"], shade_launch_ms),
                 "trace_stage": roof("trace_stage", stages_b["trace"], trace_launch_ms),
                 "blend": roof("blend_stage", stages_b["blend"], blend_launch_ms)}
        roofs["shade_stage"]["kernels"] = "shade_kernel" if args.unsorted else "classify_rows_kernel+scatter_kernel+shade_sorted_kernel (+3 scan kernels)"
        roofs["blend"]["kernels"] = ("blend_irradiance_tc_kernel+blend_depth_tc_kernel" if flags & abi.FLAG_BLEND_TC else
                                    "blend_irradiance_lists_kernel+blend_depth_lists_kernel" if probes_rank >= 148 * 64 and not flags & abi.FLAG_BLEND_TILES
                                    else "blend_irradiance_kernel+blend_depth_kernel")
        roofs["blend"]["fp32_tfma_per_s_dense_equivalent"] = probes_rank * R * 704 / (blend_launch_ms * 1e-3) / 1e12
        # Issue-slot roofline of the march (what actually binds it): warp instructions per launch from the committed ncu capture of this
        # configuration, divided by the launch time measured here, against 4 schedulers x SMs x the SM clock sampled during the timed region.
        issue_roof = None
        try:
            inst = traffic.get("march_kernel_warp_instructions")
            mhz = (clocks or {}).get("sm_mhz")
            if inst and mhz and march_launch_ms > 0 and probes_rank == sc.probes:
                sms = torch.cuda.get_device_properties(dev).multi_processor_count
                peak = 4.0 * sms * mhz * 1e6
                ach = inst / (march_launch_ms * 1e-3)
                issue_roof = {"bound": "issue", "kernel": "march_kernel", "achieved": ach / 1e9, "peak": peak / 1e9, "unit": "G warp instructions/s",
                              "frac": ach / peak, "warp_instructions_per_launch": inst,
                              "source": "profiles/traffic.json (smsp__inst_executed.sum of the ncu capture of this configuration) / launch time measured live"}
        except Exception:
            issue_roof = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "ms_per_update": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_desc(args.workload, sc),
            "parallelism": {"layout": (f"zlayers-{shards}-{world}" if flags & abi.FLAG_SHARD_INTERLEAVED and world > 1 else f"zslab{world}"),
                            "sharding": (f"probe z-layers dealt out round-robin in blocks ({shards}: rank g owns blocks g, g + N, ...: evenly loaded), " if flags & abi.FLAG_SHARD_INTERLEAVED and world > 1 else "probe z-slabs, ") + "SDF + surface cache replicated, in-place NCCL all-gather of atlas rows per step"
                            + ("" if world == 1 else ((" issued by lux_ddgi_update on the library's gather stream" if args.allgather == "lib" else " issued by the host (torch.distributed)")
                                                      + (" on the compute stream" if args.sync_allgather else ", overlapped with the next step's trace")))},
            "stage_ms": {"setup": setup_ms / args.steps, "trace": trace_launch_ms, "march": march_launch_ms, "shade": shade_launch_ms,
                         "blend_border": blend_launch_ms,
                         "update_minus_stage_sum": ms_per_step - (setup_ms + trace_ms + blend_ms) / args.steps,
                         "note": "stages timed in a second, serialized pass (one batch, one stream, CUDA events on the engine's stream); "
                                 "ms_per_update is the shipped update (blend weights computed on a second stream during the march)"},
            "trace_rays_per_s": probes_rank * world * R / (trace_launch_ms * 1e-3),
            "per_rank_trace_blend_ms": per_rank,
            "roofline": roofs[dominant],
            "roofline_issue": issue_roof,
            "roofline_stages": {k: v for k, v in roofs.items() if k != dominant and v["ms_per_launch"] > 0},
            "clocks": clocks,
            "e2e": None if e2e_value is None else {"value": e2e_value, "unit": UNIT, "ms_per_update": e2e_s / args.steps * 1e3,
                    "h2d_bytes_per_step": (light_bytes // world if comm_used else light_bytes) + 64, "d2h_bytes_per_step": d2h_bytes,
                    "sdf_region_update_ms": region_ms,
                    "sdf_region_update_note": "not part of the timed steps: one dirty 32^3 chunk (64 KiB from host memory) patched into the bound global SDF + that cascade's mip rebuilt on device (lux_ddgi_update_global_sdf_region), the per-frame call of a renderer whose scene changed; a full lux_ddgi_set_global_sdf re-upload is the alternative it replaces",
                    "note": "per rank and per step: light cache H2D from pinned memory (N > 1: own 1/N of its rows, all-gathered over NVLink by the library), own atlas rows D2H into pinned memory; the host waits for frame f-1's rows while frame f computes (every frame delivered, one frame late), all copies complete inside the timed region"},
            "blend_tc": blend_tc,
            "gpu_launches": int(launches),
            "multi_gpu_parity": parity,
            "wall_s_timed_region": t_wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            r = time_oracle(sc, rot_of(0), args.cpu_probes, 1, 0)
            line["cpu_baseline"] = {"value": r["rays_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                                    "sample": f"{r['sample_probes']} stratified probes x {R} rays, 1 update (trace + literal blend + border), "
                                              f"{r['sec_per_step']:.1f} s; trace share {r['trace_frac']:.2f}",
                                    "counters_per_ray": {k: v / r["rays"] for k, v in r["counters"].items()}}
        if "cpu_baseline" in line and l2_gbs:
            # request-level (L2) roofline of the trace, SURVEY §8d B_trace_req: 16 B per trilinear tap (8 fp16 texels; the oracle's tap counters on
            # the stratified sample, normal taps included), 16 B of ray-record stores per ray, 112 B per surface-cache tile sample (one depth gather
            # + three colour gathers); scaled from the sample to this rank's rays
            cpr = line["cpu_baseline"]["counters_per_ray"]
            req = probes_rank * R * (16.0 * (cpr["mipTaps"] + cpr["texTaps"]) + 16.0 + 112.0 * cpr["tileSamples"])
            a = req / (trace_launch_ms * 1e-3) / 1e9
            t_hbm, t_l2 = stages_b["trace"] / (hbm * 1e9), req / (l2_gbs * 1e9)
            line["roofline_l2"] = {"bound": "l2", "kernel": "trace_stage", "achieved": a, "peak": l2_gbs, "unit": "GB/s", "frac": a / l2_gbs,
                                   "request_bytes_per_launch": req, "ms_per_launch": trace_launch_ms,
                                   "peak_source": "measured live before the timed region: lux_ddgi_measure_l2_read_bandwidth (64 MiB read-only sweep by every SM, best of 5)",
                                   "floor_ms": {"hbm": t_hbm * 1e3, "l2": t_l2 * 1e3}, "binding": "l2" if t_l2 > t_hbm else "hbm"}
        if parity is not None:
            sys.stderr.write(f"multi_gpu_parity: {parity['status']}\n")
        print(json.dumps(line), flush=True)
    if pipe is not None:
        pipe.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c5",
                    help="c5 (default) = BASELINE configs[4], the largest configuration that fits one GPU: city 1024^3, 128x32x128 probes, 1024 rays; "
                         "c4 = configs[3] (512^3, 64x16x64, 512 rays; the configuration the ncu profiles under profiles/ were taken on); c2 / c3 / c1")
    ap.add_argument("--impl", default="lux", choices=["lux", "reference"])
    ap.add_argument("--cpu-probes", type=int, default=4096)
    ap.add_argument("--reference-probes", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--trace", default="texture", choices=["texture", "loads", "simple"],
                    help="SDF read path / trace kernel variant: wavefront + tld4 gathers (default), wavefront + fp16 loads, thread-per-ray")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-tc-ab", action="store_true", help="skip the short extra pass that times the opt-in tensor-core blend (N = 1 only)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the post-run parity pass (atlas checksums across ranks + oracle subsample on rank 0)")
    ap.add_argument("--e2e-skip-h2d", action="store_true", help="diagnosis only: e2e leg without the light-cache upload (the printed e2e is then NOT an end-to-end number)")
    ap.add_argument("--e2e-skip-d2h", action="store_true", help="diagnosis only: e2e leg without the atlas downloads")
    ap.add_argument("--emulate-shard", default=None, help="r/w: run shard r of w on one GPU without any collective (profiling aid)")
    ap.add_argument("--allgather", default="lib", choices=["lib", "torch"],
                    help="N > 1: exchange inside lux_ddgi_update (ncclComm bound through the C ABI, default) or issued by this script through torch.distributed")
    ap.add_argument("--shards", default="auto", choices=["auto", "interleaved", "blocks2", "blocks4", "blocks8", "slabs"],
                    help="N > 1: which probe z-layers a rank owns - one contiguous z-slab, or blocks of 1 / 2 / 4 / 8 layers dealt out round-robin "
                         "(LUX_DDGI_FLAG_SHARD_INTERLEAVED: evenly loaded ranks; single layers cost locality, blocks of 8 keep it). auto = blocks8 from 4 GPUs on")
    ap.add_argument("--sync-allgather", action="store_true", help="all-gather on the compute stream (no overlap with the next trace)")
    ap.add_argument("--no-pipeline", action="store_true", help="A/B: one batch on one stream instead of two-stream probe batches")
    ap.add_argument("--extra-flags", type=lambda v: int(v, 0), default=0, help="A/B: LUX_DDGI_FLAG_* bits OR-ed into the context flags")
    ap.add_argument("--unsorted", action="store_true", help="A/B: shade hits in ray order (no counting sort by culling chunk)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_lux(args)


if __name__ == "__main__":
    main()
