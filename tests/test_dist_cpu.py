"""N > 1 host logic on CPU: utterance sharding and the one collective of the path (the waveform gather), run
with world_size 2 over gloo.  No CUDA compute is involved (the per-rank synthesis is replaced by a known signal)."""
import os
import random
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minimax_speech_b200.pipeline import PlannedGather, gather_plan, gather_waveforms, shard_utterances, utterance_cost


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _signal(uid, n):
    return torch.arange(n, dtype=torch.float32) * 1e-4 + float(uid)


def _worker(rank, world, port, lengths, hop, ret, planned=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard_utterances(lengths, world)[rank]
        smax = max(lengths[i] for i in mine) * hop
        wav = torch.zeros(len(mine), 1, smax)
        for j, i in enumerate(mine):
            wav[j, 0, :lengths[i] * hop] = _signal(i, lengths[i] * hop)
        plan = gather_plan(shard_utterances(lengths, world), [n * hop for n in lengths]) if planned else None
        res = gather_waveforms(wav, [lengths[i] * hop for i in mine], mine, dst=0, plan=plan)
        if planned:  # the persistent-buffer form used by bench.py: same result, call after call
            g = PlannedGather(plan, wav.device, dst=0)
            for _ in range(2):
                again = g(wav)
                assert (again is None) == (rank != 0)
                if rank == 0:
                    assert sorted(again) == sorted(res) and all(torch.equal(again[i], res[i]) for i in res)
        if rank == 0:
            ok = sorted(res) == list(range(len(lengths)))
            for i, w in res.items():
                ok = ok and w.shape == (lengths[i] * hop,) and torch.equal(w, _signal(i, lengths[i] * hop))
            ret.put(bool(ok))
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("planned", [False, True])
def test_gather_waveforms_world2_gloo(planned):
    """planned: sizes from the host-side gather plan (no metadata collective, no device read-back)."""
    rng = random.Random(0)
    lengths = [rng.randint(2, 30) for _ in range(9)]  # odd count: ranks hold different numbers of utterances
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, 48, ret, planned)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get() is True


def test_shard_utterances_balances_cost_and_covers_everything():
    rng = random.Random(0)
    lengths = [50 * rng.randint(2, 30) for _ in range(256)]  # BASELINE configs[4]: 256 mixed 2-30 s utterances
    for world in (1, 2, 4, 8):
        shards = shard_utterances(lengths, world)
        assert sorted(i for s in shards for i in s) == list(range(256))
        loads = [sum(utterance_cost(lengths[i]) for i in s) for s in shards]
        assert max(loads) / (sum(loads) / world) < 1.02
