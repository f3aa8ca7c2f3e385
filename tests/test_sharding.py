"""The N > 1 path on CPU: two gloo ranks shard 7 streams, run the (oracle) chain on their own
streams only, and gather the averaged u8 spectra to rank 0 in global stream order -- the same
helpers bench.py uses with NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_streams, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft
    from oracle import pyoracle as po
    pkg = graft.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = pkg.sharding.streams_for_rank(n_streams, world, rank)
    rows = []
    for sid in mine:
        iq = pkg.synth.s2_tones(1024 * 6, seed=1000 + sid)
        ps = po.Spectrum(1024).rows(iq, K=6)[0]
        rows.append(po.db_payload(ps, 6, 0)[0])
    local = torch.as_tensor(np.stack(rows)) if rows else torch.zeros((0, 1024), dtype=torch.uint8)
    out = pkg.sharding.gather_spectra(local, n_streams, world, rank, dist=dist)
    t = pkg.sharding.max_over_ranks(1.0 + rank, world, dist=dist)
    q.put((rank, mine, None if out is None else out.numpy(), t))
    dist.barrier()
    dist.destroy_process_group()


def test_streams_partition():
    import __graft_entry__ as graft
    sh = graft.load_package().sharding
    for n, w in ((256, 1), (256, 2), (256, 8), (7, 2), (5, 8)):
        parts = [sh.streams_for_rank(n, w, r) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert all(len(p) == sh.local_count(n, w, r) for r, p in enumerate(parts))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


@pytest.mark.timeout(180)
def test_two_rank_gloo_gather(po, synth):
    import torch.multiprocessing as mp
    world, n_streams = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    got.sort(key=lambda g: g[0])
    assert got[0][1] == [0, 2, 4, 6] and got[1][1] == [1, 3, 5]
    assert got[1][2] is None
    gathered = got[0][2]
    assert gathered.shape == (n_streams, 1024)
    for sid in range(n_streams):
        iq = synth.s2_tones(1024 * 6, seed=1000 + sid)
        want = po.db_payload(po.Spectrum(1024).rows(iq, K=6)[0], 6, 0)[0]
        assert np.array_equal(gathered[sid], want)
    assert got[0][3] == got[1][3] == 2.0          # max over ranks
