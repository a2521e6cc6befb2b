"""Host-side logic of the N > 1 path, exercised with two gloo ranks on CPU: every rank takes its
shard of one instruction set (cut at cluster gaps), the per-rank summaries are gathered and must
cover the set exactly once, in time order, with balanced load."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import load_c0_config
from tests.golden.synth_instructions import c0_like


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from wfsim_b200.sharding import shard_instructions, signal_time
    cfg = load_c0_config()
    inst = c0_like(400, seed=3, event_rate=50.0)
    parts = shard_instructions(inst, world, cfg)
    mine = parts[rank]
    st = signal_time(inst, cfg['drift_velocity_liquid'])
    summary = dict(rank=rank, n=len(mine), amp=int(inst['amp'][mine].sum()),
                   tmin=int(st[mine].min()), tmax=int(st[mine].max()), idx=mine.tolist())
    gathered = [None] * world
    dist.all_gather_object(gathered, summary)
    # a data-path reduction every rank can check: total photo-quanta
    tot = torch.tensor([summary['amp']], dtype=torch.int64)
    dist.all_reduce(tot)
    if rank == 0:
        q.put((gathered, int(tot.item()), int(inst['amp'].sum()), len(inst),
               int(cfg['right_raw_extension'])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_covers_and_orders():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tot, want_tot, n, rext = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tot == want_tot
    idx = np.concatenate([g['idx'] for g in gathered])
    assert len(idx) == n and len(np.unique(idx)) == n            # exact cover
    assert gathered[0]['tmax'] + rext < gathered[1]['tmin']      # cut at a cluster gap, time ordered
    amps = np.array([g['amp'] for g in gathered], float)
    assert amps.min() / amps.max() > 0.7                          # balanced by photon-count proxy


def test_shard_edge_cases():
    from wfsim_b200.sharding import shard_instructions, merge_results
    cfg = load_c0_config()
    inst = c0_like(3, seed=1)
    parts = shard_instructions(inst, 8, cfg)                      # more shards than events
    assert len(parts) == 8 and sum(len(p) for p in parts) == len(inst)
    assert all(len(p) == 0 for p in shard_instructions(inst[:0], 4, cfg))
    # S1 and S2 of one event closer than rext stay together
    near = inst.copy()
    near['z'] = -0.5
    for p in shard_instructions(near, 3, cfg):
        if len(p):
            assert len(p) % 2 == 0
