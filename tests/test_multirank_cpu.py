"""world_size-2 gloo test of the N>1 path (CPU): frame sharding with halo frames + all-gather of match tables must
reproduce the single-process tables exactly.  The per-rank tables come from the oracle here (the CUDA path produces
them on the GPU box; the sharding / gathering logic under test is identical)."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL, CAP, NW = 6, 640, 32


def _tables_for(frame_ids):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
    import __graft_entry__ as g
    import synth
    O = g.load_oracle()
    voc = synth.vocabulary(NW)
    orb = O.OrbOracle(500, 1.2, 8, 20, 7)
    feats = []
    for f in frame_ids:
        k, d = orb.extract(synth.frame(320, 240, f))
        feats.append((k, d, O.feature_vector_csr(O.bow_assign(d, voc))))
    tab = np.full((len(frame_ids) - 1, CAP), -1, np.int32)
    for i in range(len(frame_ids) - 1):
        (k1, d1, f1), (k2, d2, f2) = feats[i], feats[i + 1]
        _, m = O.search_by_bow(d1, d2, f1, f2, np.ones(len(d1), np.uint8), k1["angle"], k2["angle"], 0.7, True)
        tab[i, :len(m)] = m
    return tab


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_package()
    from sslpl_b200 import batch
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = batch.FrameShard(TOTAL, world, rank)
    local = torch.from_numpy(_tables_for(sh.frame_ids()))
    full = batch.all_gather_tables(local)
    # the packed form used by bench.py: point table + a (synthetic) line table in ONE collective
    lines = (local[:, :40] * 7 + rank).contiguous()
    pg = batch.PackedGather(sh.per_rank, CAP, 40, world, "cpu")
    pg.stage_lines(lines)
    gp, gl = pg.gather(local)
    assert torch.equal(gp, full) and torch.equal(gl[rank * sh.per_rank:(rank + 1) * sh.per_rank], lines)
    q.put((rank, sh.frame_ids(), sh.pair_ids(), full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gather_equals_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2, 3] and res[1][1] == [3, 4, 5, 0]            # block + cyclic halo
    assert res[0][2] == [0, 1, 2] and res[1][2] == [3, 4, 5]
    single = _tables_for(list(range(TOTAL)) + [0])
    for r in range(world):
        assert np.array_equal(res[r][3], single), f"rank {r}: gathered table differs from the single-process table"
    assert (single >= 0).sum() > 50


def test_shard_validation():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.load_package()
    from sslpl_b200 import batch
    import pytest
    with pytest.raises(ValueError):
        batch.FrameShard(7, 2, 0)
    assert batch.FrameShard(512, 8, 7).frame_ids()[-1] == 0 and len(batch.FrameShard(512, 8, 3).frame_ids()) == 65
