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


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sopro_b200.model as model_mod
    from sopro_b200.config import SoproTTSConfig
    from sopro_b200.dp import DataParallelTTS
    from sopro_b200.weights import synth_state_dict
    from tests.cases import SMALL_CFG

    class _StubTTS:  # stands where SoproTTS.from_state_dict builds the CUDA engines (no GPU in this test)
        def __init__(self, sd):
            self.sd, self.calls = sd, []

        def synthesize_batch(self, texts, *, ref, seeds=None, **kw):
            self.calls.append((list(texts), seeds))
            return [f"wav:{t}" for t in texts]

    orig = model_mod.SoproTTS.from_state_dict
    model_mod.SoproTTS.from_state_dict = classmethod(lambda cls, cfg, sd, tok, msd, **kw: _StubTTS(sd))
    try:
        cfg = SoproTTSConfig(**SMALL_CFG)
        sd0 = synth_state_dict(cfg, 64, 0) if rank == 0 else None
        dp = DataParallelTTS(cfg, sd0, None, None, device="cpu", text_vocab=64)
        ref_sd = synth_state_dict(cfg, 64, 0)
        same = set(dp.tts.sd) == set(ref_sd) and all(torch.equal(dp.tts.sd[k], ref_sd[k].float()) for k in ref_sd)
        texts = [f"t{i}" for i in range(5)]
        wavs, (lo, hi) = dp.synthesize_batch(texts, ref=None, seeds=[10, 11, 12, 13, 14], max_frames=3)
        out[rank] = (same, (lo, hi), wavs, dp.tts.calls[0][1])
    finally:
        model_mod.SoproTTS.from_state_dict = orig
        dist.destroy_process_group()


def test_data_parallel_tts_broadcasts_weights_and_shards_the_batch():
    """sopro_b200.dp.DataParallelTTS at world 2 (gloo): every rank ends up with rank 0's checkpoint and synthesises its
    own contiguous slice of the global batch, with that slice's seeds."""
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_dp_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][0] and res[1][0]
    assert res[0][1] == (0, 2) and res[1][1] == (2, 5)
    assert res[0][2] == ["wav:t0", "wav:t1"] and res[1][2] == ["wav:t2", "wav:t3", "wav:t4"]
    assert res[0][3] == [10, 11] and res[1][3] == [12, 13, 14]
