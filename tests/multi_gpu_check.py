"""Multi-GPU parity check of the in-library exchange step (not collected by pytest: needs N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py

Twice - once per shard layout (z-slabs; LUX_DDGI_FLAG_SHARD_INTERLEAVED: rank g owns the z-layers g, g + N, ...) - every rank runs its shard of a
small city volume for 4 frames with an ncclComm bound through lux_ddgi_set_nccl_comm and then
(frame 2 relights through the sharded light-cache upload) and then
compares its FULL atlases (own rows + the rows the all-gather delivered) with the CPU oracle's unsharded run, bit for bit; consumers
(lux_ddgi_sample_irradiance at points spread over the whole volume) must see the complete atlas as well.  Prints one line per rank."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from luxgi_b200 import abi, ddgi, nccl, scenes
    from oracle import binding as ob

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    all_ok = 1
    for layout, flags in (("z-slabs", 0), ("interleaved layers", abi.FLAG_SHARD_INTERLEAVED)):
        sc = scenes.build("city128", counts=(16, 8, 16), rays=64)
        sc.uniform.normalBias = 0.3
        orc = ob.OraclePipeline(sc)
        pipe = ddgi.DDGIPipeline(sc.uniform, device=local, rank=rank, world=world, flags=flags)
        pipe.set_scene(sc)
        comm = nccl.NcclComm(rank, world, local)
        pipe.set_nccl_comm(comm.ptr)
        res = int(sc.atlas_data.resolution)
        rows = res // world
        for f in range(4):
            rot = scenes.frame_rotation(f)
            if f == 2:  # sharded relight: every rank uploads its rows of a brighter light cache, the library all-gathers them
                new_light = (sc.light.numpy().astype(np.float32) * 1.5).astype(np.float16)
                orc.os.light[...] = new_light.view(np.uint16)
                mine = torch.from_numpy(np.ascontiguousarray(new_light[rank * rows:(rank + 1) * rows])).pin_memory()
                pipe.update_surface_light_cache_rows_ptr(mine.data_ptr(), rank * rows, rows)
            orc.update(rot)
            pipe.update(rot)
        ok_i = np.array_equal(pipe.irradiance, orc.irradiance)
        ok_d = np.array_equal(pipe.depth, orc.depth)
        rng = np.random.default_rng(5)
        u = sc.uniform
        lo = np.array([u.startPosition[i] for i in range(3)])
        hi = lo + np.array([u.step[i] * (u.probeCounts[i] - 1) for i in range(3)])
        P = rng.uniform(lo, hi, size=(512, 3)).astype(np.float32)
        N = rng.normal(size=(512, 3)).astype(np.float32)
        N /= np.linalg.norm(N, axis=1, keepdims=True)
        want = ob.sample_irradiance(sc.uniform, orc.irradiance, orc.depth, P, N, -N)
        got = pipe.sample_irradiance(P, N, -N)
        ok_s = np.array_equal(got.view(np.uint32), want.view(np.uint32))
        print(f"{layout}: rank {rank}/{world}: irradiance {'OK' if ok_i else 'DIFF'} depth {'OK' if ok_d else 'DIFF'} consumer {'OK' if ok_s else 'DIFF'}", flush=True)
        flag = torch.tensor([int(ok_i and ok_d and ok_s)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        all_ok = min(all_ok, int(flag.item()))
        pipe.close()
        comm.destroy()
    dist.destroy_process_group()
    sys.exit(0 if all_ok == 1 else 1)


if __name__ == "__main__":
    main()
