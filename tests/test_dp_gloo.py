"""CPU, world_size 2, gloo: the N>1 host logic (weight broadcast, utterance sharding, token gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sopro_b200.config import SoproTTSConfig
    from sopro_b200.dp import broadcast_state_dict, gather_token_lists, shard_range
    from sopro_b200.weights import param_specs, synth_state_dict
    from tests.cases import SMALL_CFG

    cfg = SoproTTSConfig(**SMALL_CFG)
    specs = [(k, v[0]) for k, v in param_specs(cfg, 64).items() if k.startswith("ar.")]
    sd = synth_state_dict(cfg, 64, 0, only_prefix=("ar.",)) if rank == 0 else None
    got = broadcast_state_dict(sd, specs, src=0)
    ref = synth_state_dict(cfg, 64, 0, only_prefix=("ar.",))
    ok = all(torch.equal(got[k], ref[k]) for k, _ in specs)
    lo, hi = shard_range(7, rank, world)
    toks = gather_token_lists([[rank, i] for i in range(lo, hi)])
    out[rank] = (ok, (lo, hi), toks)
    dist.destroy_process_group()


def test_broadcast_shard_gather_world2():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][0] and res[1][0]
    assert res[0][1] == (0, 3) and res[1][1] == (3, 7)
    assert res[0][2] == res[1][2] == [[0, 0], [0, 1], [0, 2], [1, 3], [1, 4], [1, 5], [1, 6]]


def test_shard_range_covers_everything():
    from sopro_b200.dp import shard_range

    for n in (1, 7, 64, 512, 513):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1
