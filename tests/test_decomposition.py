"""CPU: host-side logic of the multi-GPU brick decomposition (LAMMPS `processors Px Py Pz` semantics) exercised by
two gloo ranks: every particle has exactly one owner, the neighbour-link tables of the two ranks mirror each other
(so the grouped ncclSend/ncclRecv of the halo exchange pair up), and periodic links carry opposite shifts."""
import os
import socket
import sys

import numpy as np
import pytest

import sedifoam_b200 as sb
from sedifoam_b200 import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_auto_grid_minimises_surface():
    sb.build_library()
    assert list(sb.decomp_grid(8, (1.0, 1.0, 1.0))) == [2, 2, 2]
    assert list(sb.decomp_grid(8, (8.0, 1.0, 1.0))) == [8, 1, 1]
    assert list(sb.decomp_grid(4, (1.0, 3.0, 1.0))) in ([1, 4, 1], [1, 2, 2], [2, 2, 1])
    g = sb.decomp_grid(6, (3.0, 2.0, 1.0))
    assert int(np.prod(g)) == 6


def test_links_mirror_each_other_all_ranks():
    grid = np.array([2, 2, 2], np.int32); per = np.array([1, 0, 1], np.int32); prd = np.array([1.0, 2.0, 3.0])
    tables = [sb.decomp_links(r, grid, per, prd) for r in range(8)]
    for r, (peers, offs, sh) in enumerate(tables):
        assert len(peers) > 0
        for p, o, s in zip(peers, offs, sh):
            pp, po, ps = tables[p]
            # the peer has the opposite link back to r with the opposite shift
            hit = [k for k in range(len(pp)) if pp[k] == r and np.array_equal(po[k], -o)]
            assert len(hit) == 1
            assert np.allclose(ps[hit[0]], -s)
        # no link across a non-periodic face
        c = [r % 2, (r // 2) % 2, r // 4]
        for o in offs:
            assert 0 <= c[1] + o[1] <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = cases.sediment_column(dims=(9, 11, 7), phi=0.4, jitter_frac=0.2)
    lo, hi = case["box_lo"], case["box_hi"]
    grid = sb.decomp_grid(world, hi - lo)
    owner = sb.decomp_owner(case["x"], lo, hi, grid)
    mine = np.flatnonzero(owner == rank)
    counts = [None] * world
    dist.all_gather_object(counts, mine.tolist())
    per = np.array([1 if p == "p" else 0 for p in case["periodic"]], np.int32)
    links = sb.decomp_links(rank, grid, per, hi - lo)
    all_links = [None] * world
    dist.all_gather_object(all_links, (links[0].tolist(), links[1].tolist(), links[2].tolist()))
    ok = True
    union = sorted(sum(counts, []))
    ok = ok and union == list(range(len(case["tag"])))           # every particle owned exactly once
    # the engine's own view agrees (host-only queries: no GPU)
    e = sb.Lammps(); cases.apply(case, e)
    ok = ok and e.get_local_n() == len(case["tag"])
    for p, o, s in zip(*links):
        pp, po, ps = all_links[p]
        ok = ok and any(pp[k] == rank and po[k] == [-v for v in o] for k in range(len(pp)))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok, len(mine), grid.tolist()))


def test_two_gloo_ranks_agree_on_ownership_and_links():
    import multiprocessing as mp
    sb.build_library()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=30)
    assert all(r[1] for r in res), res
    assert sum(r[2] for r in res) == 9 * 11 * 7
    assert res[0][3] == res[1][3] and int(np.prod(res[0][3])) == 2
