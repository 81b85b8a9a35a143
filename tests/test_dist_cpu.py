"""CPU-only: host logic of the environment-sharded update, incl. a world_size-2 gloo run."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from a2c_ppo_acktr import _dist


def test_shard_rows_partition_property():
    T, N = 6, 10
    perm = torch.randperm(T * N).numpy()
    shards = [_dist.ShardInfo(3, r, lo, hi, N) for r, (lo, hi) in enumerate([(0, 3), (3, 7), (7, 10)])]
    got = []
    for sh in shards:
        loc = _dist.shard_rows(perm, sh)
        t, n = loc // sh.n_local, loc % sh.n_local
        got.append(t * N + n + sh.lo)
    merged = np.concatenate(got)
    assert sorted(merged.tolist()) == sorted(perm.tolist())
    # order inside a shard follows the global permutation
    for sh, g in zip(shards, got):
        owned = [x for x in perm.tolist() if sh.lo <= x % N < sh.hi]
        assert g.tolist() == owned


def test_shard_epoch_padding_and_capacity():
    sh = _dist.ShardInfo(2, 1, 4, 8, 8)
    perm = torch.randperm(128 * 8).numpy()
    cap = _dist.default_capacity(256, sh)
    idx, counts = _dist.shard_epoch(perm, 256, 4, sh, cap)
    assert idx.shape[0] == 4 and idx.shape[1] >= counts.max() and idx.shape[1] % 4 == 0
    assert counts.sum() == 128 * 4
    for k in range(4):
        assert not idx[k, counts[k]:].any()


def test_shard_env_chunks_partition_property():
    """Recurrent minibatches of the ONE global randperm(N) (reference storage.py:152): the shards' pieces of every
    chunk are disjoint, cover it, and keep the permutation's order."""
    N, nmb = 64, 4
    perm = torch.randperm(N).numpy()
    shards = [_dist.ShardInfo(4, r, 16 * r, 16 * (r + 1), N) for r in range(4)]
    pieces = [_dist.shard_env_chunks(perm, N // nmb, nmb, sh) for sh in shards]
    for k in range(nmb):
        chunk = perm[k * 16:(k + 1) * 16].tolist()
        merged = [int(e) + sh.lo for sh, pc in zip(shards, pieces) for e in pc[k]]
        assert sorted(merged) == sorted(chunk)
        for sh, pc in zip(shards, pieces):
            assert [int(e) + sh.lo for e in pc[k]] == [e for e in chunk if sh.lo <= e < sh.hi]
    sh = _dist.ShardInfo(8, 3, 768, 1024, 2048)
    cap = _dist.env_capacity(512, sh, 4)
    assert 64 < cap <= 112 and cap % 4 == 0
    counts = [len(c) for _ in range(200) for c in _dist.shard_env_chunks(torch.randperm(2048).numpy(), 512, 4, sh)]
    assert max(counts) <= cap and abs(np.mean(counts) - 64) < 2
    assert _dist.round_up_envs(0, 4) == 4 and _dist.round_up_envs(65, 4) == 68 and _dist.round_up_envs(3, 1) == 3
    small = _dist.ShardInfo(2, 0, 0, 2, 4)
    assert _dist.env_capacity(2, small, 1) == 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T, N = 16, 6
    g = torch.Generator().manual_seed(0)
    returns, values = torch.randn(T, N, generator=g), torch.randn(T, N, generator=g)
    lo, hi = (0, 3) if rank == 0 else (3, 6)
    sh = _dist.ShardInfo(world, rank, lo, hi, N)
    x = (returns - values)[:, lo:hi].double()
    moments = torch.tensor([x.sum(), (x * x).sum(), float(x.numel())], dtype=torch.float64)
    _dist.all_reduce_sum(moments)                      # the 3-double exchange of the sharded update
    mean = moments[0] / moments[2]
    std = ((moments[1] - moments[2] * mean * mean) / (moments[2] - 1)).sqrt()
    ref = (returns - values)
    ok = abs(mean.item() - ref.mean().item()) < 1e-6 and abs(std.item() - ref.std().item()) < 1e-6
    # gradient exchange: each rank contributes the rows it owns; the sum equals the global batch
    torch.manual_seed(7)
    perm = torch.randperm(T * N).numpy()
    rows = _dist.shard_rows(perm[:24], sh)
    w = torch.arange(T * N, dtype=torch.float64).view(T, N)[:, lo:hi].reshape(-1)
    part = torch.tensor([w[rows].sum()], dtype=torch.float64)
    _dist.all_reduce_sum(part)
    want = torch.arange(T * N, dtype=torch.float64)[perm[:24]].sum()
    ok = ok and abs(part.item() - want.item()) < 1e-9
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_world2_gloo_moments_and_partition():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0] and out[1]


def _plan_loop(perms, E_global, n_mb, world, n_local):
    """The plan of `_dist.balance_update_plan` written as the plain loop over minibatches and ranks (the definition the
    vectorised implementation must reproduce entry for entry)."""
    import numpy as np
    target = E_global // world
    out = [dict(envs=[[None] * n_mb for _ in perms], send=[[] for _ in range(world)], recv=[[] for _ in range(world)],
                n_guests=0, first_send=[0] * world, first_recv=[0] * world) for _ in range(world)]
    for e, perm in enumerate(perms):
        for k in range(n_mb):
            chunk = np.asarray(perm, dtype=np.int64)[k * E_global:(k + 1) * E_global]
            keep, surplus = [], []
            for r in range(world):
                mine = [int(c) - r * n_local for c in chunk if c // n_local == r]
                keep.append(mine[:target])
                surplus += [(r, c) for c in mine[target:]]
            for r in range(world):
                while len(keep[r]) < target:
                    src, col = surplus.pop(0)
                    slot = out[r]["n_guests"]
                    out[r]["n_guests"] += 1
                    out[src]["send"][r].append(col)
                    out[r]["recv"][src].append(slot)
                    if e == 0:
                        out[src]["first_send"][r] += 1
                        out[r]["first_recv"][src] += 1
                    keep[r].append(n_local + slot)
                out[r]["envs"][e][k] = keep[r]
            assert not surplus
    return out


def test_balance_update_plan_equals_its_definition():
    import numpy as np
    import torch
    from a2c_ppo_acktr import _dist
    for world, n_total, nmb, epochs in ((2, 512, 4, 4), (8, 2048, 4, 4), (4, 64, 2, 3), (2, 8, 2, 3), (4, 16, 4, 2), (16, 4096, 4, 4)):
        n_local = n_total // world
        g = torch.Generator().manual_seed(17 + world)
        perms = [torch.randperm(n_total, generator=g).numpy() for _ in range(epochs)]
        E_g = n_total // nmb
        got = _dist.balance_update_plan(perms, E_g, nmb, world, n_local)
        want = _plan_loop(perms, E_g, nmb, world, n_local)
        for r in range(world):
            assert got[r]["n_guests"] == want[r]["n_guests"]
            assert got[r]["first_send"] == want[r]["first_send"] and got[r]["first_recv"] == want[r]["first_recv"]
            for d in range(world):
                assert list(got[r]["send"][d]) == want[r]["send"][d], (r, d)
                assert list(got[r]["recv"][d]) == want[r]["recv"][d], (r, d)
            for e in range(epochs):
                for k in range(nmb):
                    assert got[r]["envs"][e][k].dtype == np.int64
                    assert list(got[r]["envs"][e][k]) == want[r]["envs"][e][k], (r, e, k)


def test_balance_update_plan_moves_the_minimum_and_keeps_the_reference_composition():
    """`_dist.balance_update_plan`: every rank gets exactly E / world environments of every global minibatch, kept
    environments stay where they are, the union over ranks is the reference's chunk perm[k E : (k + 1) E], the number of
    moves is the sum of the surpluses (the minimum), and both ends of every transfer agree on counts and order."""
    import numpy as np
    import torch
    from a2c_ppo_acktr import _dist
    for world, n_total, nmb, epochs in ((2, 512, 4, 4), (8, 2048, 4, 4), (4, 64, 2, 3)):
        n_local = n_total // world
        g = torch.Generator().manual_seed(world)
        perms = [torch.randperm(n_total, generator=g).numpy() for _ in range(epochs)]
        E_g = n_total // nmb
        plans = _dist.balance_update_plan(perms, E_g, nmb, world, n_local)
        # what every guest slot holds: replay the transfers
        guest = [dict() for _ in range(world)]
        for s in range(world):
            for d in range(world):
                assert len(plans[s]["send"][d]) == len(plans[d]["recv"][s])
                for col, slot in zip(plans[s]["send"][d], plans[d]["recv"][s]):
                    guest[d][int(slot)] = int(col) + s * n_local
        for r in range(world):
            assert sorted(guest[r]) == list(range(plans[r]["n_guests"]))
        moves = 0
        for e in range(epochs):
            for k in range(nmb):
                chunk = perms[e][k * E_g:(k + 1) * E_g]
                got = []
                for r in range(world):
                    cols = plans[r]["envs"][e][k]
                    assert len(cols) == E_g // world
                    got += [int(c) + r * n_local if c < n_local else guest[r][int(c) - n_local] for c in cols]
                assert sorted(got) == sorted(int(x) for x in chunk)
                owner = chunk // n_local
                moves += sum(max(int((owner == r).sum()) - E_g // world, 0) for r in range(world))
        assert moves == sum(p["n_guests"] for p in plans)


def _migration_worker(rank, world, port, out):
    """The environment migration of the balanced scheme over a real process group (gloo): every rank plans from the same
    permutations, packs one record per travelling environment, and the two all-to-alls (first epoch / later epochs) must
    leave every guest slot holding the record of the environment the plan put there."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local, nmb, epochs, width = 16, 2, 3, 24
    n_total = n_local * world
    g = torch.Generator().manual_seed(5)                       # same seed everywhere: same permutations, same plan
    perms = [torch.randperm(n_total, generator=g).numpy() for _ in range(epochs)]
    plans = _dist.balance_update_plan(perms, n_total // nmb, nmb, world, n_local)
    mine = plans[rank]
    # record of global environment e: [e, e + 1, ...] as bytes
    def record(e):
        return (torch.arange(width, dtype=torch.int64) + int(e)).to(torch.uint8)
    guests = torch.zeros(max(p["n_guests"] for p in plans) + 1, width, dtype=torch.uint8)
    for phase in (0, 1):
        cut_s, cut_r = mine["first_send"], mine["first_recv"]
        snd = [x[:c] if phase == 0 else x[c:] for x, c in zip(mine["send"], cut_s)]
        rcv = [x[:c] if phase == 0 else x[c:] for x, c in zip(mine["recv"], cut_r)]
        cols = np.concatenate(snd + [np.zeros(0, np.int64)])
        send = torch.stack([record(rank * n_local + c) for c in cols]) if len(cols) else torch.zeros(0, width, dtype=torch.uint8)
        recv = torch.zeros(sum(len(x) for x in rcv), width, dtype=torch.uint8)
        _dist.Exchange(send, [len(x) for x in snd], recv, [len(x) for x in rcv]).run()
        slots = np.concatenate(rcv + [np.zeros(0, np.int64)])
        if len(slots):
            guests[torch.from_numpy(slots)] = recv
    # what every guest slot must hold, replayed from the senders' side of the plan
    ok = True
    for s in range(world):
        for col, slot in zip(plans[s]["send"][rank], mine["recv"][s]):
            ok = ok and torch.equal(guests[int(slot)], record(s * n_local + int(col)))
    # and every minibatch of this rank is E / world environments of the reference's chunk
    E_g = n_total // nmb
    slot_env = {int(slot): s * n_local + int(col) for s in range(world) for col, slot in zip(plans[s]["send"][rank], mine["recv"][s])}
    for e in range(epochs):
        for k in range(nmb):
            envs = [rank * n_local + int(c) if c < n_local else slot_env[int(c) - n_local] for c in mine["envs"][e][k]]
            ok = ok and len(envs) == E_g // world and set(envs) <= set(int(x) for x in perms[e][k * E_g:(k + 1) * E_g])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_world2_gloo_environment_migration_in_two_phases():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_migration_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0] and out[1]
