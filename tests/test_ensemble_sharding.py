"""N > 1 path on CPU: rows of the ensemble sharded over ranks (gloo, world_size 2), oracle arithmetic
per shard, one all-reduce of the observables -- must equal the single-process result."""
import os
import socket

import numpy as np
import pytest

from richmol_b200.ensemble import shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 24, 500, 501):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from oracle import port as oracle
    from richmol_b200 import convert_units as cu, synth
    from richmol_b200.ensemble import allreduce_sum, shard_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = synth.ocs(4)
        h0, pol, cos2 = (oracle.OracleTensor(m[k]) for k in ("h0", "pol", "cos2"))
        pol.mul(-0.5 * cu.AUpol_x_Vm_to_invcm())
        vecs = oracle.init_state(h0, temp=2.0)
        mine = shard_rows(vecs).copy()                       # this rank's ensemble members
        fac = oracle.exp_factor(0.01)
        phase = oracle.h0_phase(h0, fac)
        cos2.field([0, 0, 1])
        cm = cos2.tomat()
        obs = []
        for step in range(3):
            pol.field([1e9 * step, 0.0, 4e9])
            mine = oracle.update_step(pol, mine, fac, phase=phase)
            local = np.array([sum(np.vdot(v, cm.dot(v)) for v in mine), sum(np.vdot(v, v) for v in mine)])
            obs.append(allreduce_sum(local))                 # the path's only collective
        if rank == 0:
            np.save(out, np.array(obs))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    from oracle import port as oracle
    from richmol_b200 import convert_units as cu, synth
    out = str(tmp_path / "obs.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    m = synth.ocs(4)
    h0, pol, cos2 = (oracle.OracleTensor(m[k]) for k in ("h0", "pol", "cos2"))
    pol.mul(-0.5 * cu.AUpol_x_Vm_to_invcm())
    vecs = oracle.init_state(h0, temp=2.0)
    fac = oracle.exp_factor(0.01)
    phase = oracle.h0_phase(h0, fac)
    cos2.field([0, 0, 1])
    cm = cos2.tomat()
    ref = []
    for step in range(3):
        pol.field([1e9 * step, 0.0, 4e9])
        vecs = oracle.update_step(pol, vecs, fac, phase=phase)
        ref.append([sum(np.vdot(v, cm.dot(v)) for v in vecs), sum(np.vdot(v, v) for v in vecs)])
    assert np.allclose(got, np.array(ref), rtol=1e-13, atol=1e-15)
