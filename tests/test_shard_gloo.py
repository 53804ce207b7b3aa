"""The N > 1 host path on CPU: two ranks over gloo shard the utterances, 'decode' their slices with
the CPU oracle (tiny model, sampled with the counter-based RNG keyed by the global utterance id),
gather the ragged code tensors on rank 0 and reduce timings with MAX.  The gathered result must
equal the single-process run: sharding cannot change any id."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _decode_slice(prompts, seq_ids, n_frames):
    from oracle.dualar_oracle import DualAROracle, OracleSettings
    from oracle.sampler_oracle import OracleSampler
    from smoltts_b200.config import named_config
    from smoltts_b200.synth import make_state_dict

    cfg = named_config("smoltts_byte_tiny")
    orc = DualAROracle(cfg, make_state_dict(cfg, seed=0), dtype=torch.float32, max_seq_len=128)
    st = OracleSettings(default_temp=0.8, default_fast_temp=0.8, top_k=30, top_p=0.95, seed=99)
    outs = []
    with torch.no_grad():
        for p, sid in zip(prompts, seq_ids):
            frames = orc.generate(p, st, fixed_frames=n_frames, sampler=OracleSampler(seq_ids=[sid]))
            outs.append(torch.tensor([f.vq for f in frames], dtype=torch.int32).t().contiguous())
    return outs


def _all_prompts():
    from smoltts_b200.config import named_config
    from smoltts_b200.synth import byte_prompt, prompt_grid

    cfg = named_config("smoltts_byte_tiny")
    return [prompt_grid(byte_prompt(6 + b, seed=40 + b), cfg) for b in range(5)]


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist

    from smoltts_b200.shard import gather_utterances, max_over_ranks, shard

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine, ids = shard(_all_prompts(), rank, world)
        outs = _decode_slice(mine, ids, n_frames)
        dist.barrier()
        gathered = gather_utterances(outs, dist)
        side = dist.new_group(backend="gloo")      # bench.py gathers through a gloo side group next to the NCCL default group
        again = gather_utterances(outs, dist, group=side)
        assert (again is None) == (gathered is None)
        if gathered is not None:
            assert all(torch.equal(a, b) for a, b in zip(gathered, again))
        t = max_over_ranks([float(rank + 1), 10.0 - rank], dist)
        if rank == 0:
            q.put(([g.tolist() for g in gathered], t, ids))
        else:
            assert gathered is None
            q.put((None, t, ids))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_matches_single_process():
    n_frames, world = 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gathered = [r for r in results if r[0] is not None]
    assert len(gathered) == 1
    codes, t, _ = gathered[0]
    assert all(r[1] == [2.0, 10.0] for r in results)                 # MAX over ranks
    assert sorted(i for r in results for i in r[2]) == [0, 1, 2, 3, 4]
    want = _decode_slice(_all_prompts(), list(range(5)), n_frames)
    assert codes == [w.tolist() for w in want]


# ---- the product entry point (ShardedGenerator / generate_sharded) with a CPU stand-in engine ----------------------
def _oracle_factory(spec, rank):
    torch.set_num_threads(1)
    return {"rank": rank, "frames": spec["frames"]}


def _oracle_run(engine, prompts, settings, seq_ids, kwargs):
    return _decode_slice(prompts, seq_ids, engine["frames"])


@pytest.mark.timeout(300)
def test_sharded_generator_orders_and_gathers_like_one_process():
    """Two worker processes behind ShardedGenerator (the CPU oracle standing in for the CUDA engine): utterances come back
    in the caller's order with the ids of a single-process run; a second call reuses the workers; an empty shard is fine."""
    from smoltts_b200.shard import ShardedGenerator

    prompts = _all_prompts()
    want = [w.tolist() for w in _decode_slice(prompts, list(range(5)), 2)]
    with ShardedGenerator({"frames": 2}, gpus=2, factory=_oracle_factory, run=_oracle_run) as g:
        assert [t.tolist() for t in g.generate(prompts, None)] == want
        assert [t.tolist() for t in g.generate(prompts[:1], None)] == want[:1]   # worker 1 gets an empty slice


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a CUDA device")
def test_sharded_generator_fails_loudly_when_a_worker_cannot_start():
    from smoltts_b200.shard import ShardedGenerator

    with pytest.raises(RuntimeError, match="failed to start"):
        ShardedGenerator({"model": "smoltts_byte_tiny"}, gpus=1)   # default factory: needs a CUDA device, there is none here
