"""Multi-rank host logic on CPU: z-slab shard layout (C ABI, no GPU needed) + the in-place all-gather of atlas rows
(SURVEY §8e), world_size 2 over gloo.  The CPU oracle stands in for the kernels: each rank traces + blends + borders only
its own probe slab, the ranks exchange their atlas rows exactly as bench.py does over NCCL, and every rank must end up
with the atlases a single process computes for the whole volume."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from luxgi_b200 import abi, ddgi, scenes


def test_shard_layout_partitions_probes_and_rows(engine_lib):
    u = abi.make_uniform((0, 0, 0), (1, 1, 1), (4, 3, 8), 32)
    P = abi.probe_count(u)
    for world in (1, 2, 4, 8):
        probes, irr_rows, dep_rows = [], [], []
        for r in range(world):
            st = ddgi.shard_layout(u, r, world)
            assert st.probeCount == P // world and st.probeBegin == r * st.probeCount
            probes += list(range(st.probeBegin, st.probeBegin + st.probeCount))
            irr_rows += list(range(st.irradianceRowBegin, st.irradianceRowBegin + st.irradianceRowCount))
            dep_rows += list(range(st.depthRowBegin, st.depthRowBegin + st.depthRowCount))
        assert probes == list(range(P))
        assert irr_rows == list(range(1, u.irradianceTextureHeight - 1))  # everything but the two outer pad rows
        assert dep_rows == list(range(1, u.depthTextureHeight - 1))
    with pytest.raises(ddgi.LuxError):
        ddgi.shard_layout(u, 0, 3)  # 3 does not divide Z = 8


def test_interleaved_shard_layout_partitions_probes_and_rows(engine_lib):
    """LUX_DDGI_FLAG_SHARD_INTERLEAVED: the z-layers are dealt out in blocks of B layers, rank g owning the blocks g, g + world, ...; the shards still
    partition probes and rows, and round k of the exchange (blocks k * world .. k * world + world - 1) is a contiguous range of rows to which rank g
    contributes the g-th part."""
    u = abi.make_uniform((0, 0, 0), (1, 1, 1), (4, 3, 16), 32)
    P, xy = abi.probe_count(u), 12
    for log2b in (0, 1, 2):
        B = 1 << log2b
        for world in (1, 2, 4):
            probes, irr_rows, dep_rows = [], [], []
            for r in range(world):
                st = ddgi.shard_layout(u, r, world, abi.flag_shard_blocks(log2b))
                inter = world > 1
                assert st.probeCount == P // world and st.layerStride == (world if inter else 1)
                assert st.unitLayers == (B if inter else 1) and st.layerProbes == st.unitLayers * xy
                own = st.own_probes()
                assert len(own) == st.probeCount and (not inter or all(((p // xy) // B) % world == r for p in own))
                probes += own
                irr_rows += st.own_rows(8)
                dep_rows += st.own_rows(16)
                for k in range(st.probeCount // st.layerProbes):  # the rank's part of round k sits at offset r inside the round's block of rows
                    if inter:
                        assert st.own_rows(8)[k * 10 * B] == 1 + (k * world + r) * 10 * B and st.own_rows(16)[k * 18 * B] == 1 + (k * world + r) * 18 * B
            assert sorted(probes) == list(range(P))
            assert sorted(irr_rows) == list(range(1, u.irradianceTextureHeight - 1))
            assert sorted(dep_rows) == list(range(1, u.depthTextureHeight - 1))
    with pytest.raises(ddgi.LuxError):
        ddgi.shard_layout(u, 0, 4, abi.flag_shard_blocks(3))  # 4 ranks x 8 layers do not divide Z = 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as ob

    sc = scenes.cornell_scene(res=32, counts=(4, 2, 4), rays=32, atlas_res=256)
    u = sc.uniform
    st = ddgi.shard_layout(u, rank, world)
    pipe = ob.OraclePipeline(sc, probe_begin=st.probeBegin, count=st.probeCount)
    for f in range(2):
        pipe.update(scenes.frame_rotation(f))
        for atlas, rb, rc in ((pipe.irradiance, st.irradianceRowBegin, st.irradianceRowCount), (pipe.depth, st.depthRowBegin, st.depthRowCount)):
            t = torch.from_numpy(atlas.view(np.uint8))  # bytes (gloo has no 16-bit types); in place: the gathered rows land in the rank's own atlas
            full = t[1:1 + rc * world].reshape(-1)
            mine = t[rb:rb + rc].reshape(-1).clone()
            dist.all_gather_into_tensor(full, mine)
    np.save(os.path.join(out_dir, f"irr{rank}.npy"), pipe.irradiance)
    np.save(os.path.join(out_dir, f"dep{rank}.npy"), pipe.depth)
    dist.destroy_process_group()


def test_two_rank_allgather_reassembles_full_atlases(oracle, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sc = scenes.cornell_scene(res=32, counts=(4, 2, 4), rays=32, atlas_res=256)
    ref = oracle.OraclePipeline(sc)
    for f in range(2):
        ref.update(scenes.frame_rotation(f))
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"irr{r}.npy"), ref.irradiance)
        assert np.array_equal(np.load(tmp_path / f"dep{r}.npy"), ref.depth)
