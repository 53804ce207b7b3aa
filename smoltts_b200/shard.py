"""Utterance sharding across the GPUs of one box.

The decode path has no cross-utterance term, so multi-GPU is plain data parallelism: every rank owns
a full weight replica, a private KV pool and a contiguous slice of the utterances; there is no
collective on the decode path.  ``torch.distributed`` is used only to gather the emitted codes on the
host and to reduce timings (max over ranks).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def partition(n_items: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [start, end) ranges; the first ``n_items % world`` ranks get one more."""
    base, extra = divmod(n_items, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((start, start + n))
        start += n
    return out


def shard(items: Sequence, rank: int, world: int) -> Tuple[list, List[int]]:
    """This rank's slice of ``items`` and the global ids of its elements (the RNG's ``seq_id``)."""
    lo, hi = partition(len(items), world)[rank]
    return list(items[lo:hi]), list(range(lo, hi))


def gather_utterances(local: Sequence[torch.Tensor], dist=None, dst: int = 0, group=None) -> Optional[List[torch.Tensor]]:
    """Host-side gather of per-utterance code tensors (ragged) onto ``dst`` in global utterance order.
    ``dist`` is the ``torch.distributed`` module; ``group`` an optional process group of it (a gloo group when the default
    one is NCCL: the codes are host tensors)."""
    local = [t.cpu() for t in local]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    bucket = [None] * world if rank == dst else None
    dist.gather_object(local, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    return [t for part in bucket for t in part]


def max_over_ranks(values: Sequence[float], dist=None, device=None) -> List[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


# ------------------------------------------------------------------------------------------------
# Product entry point: one worker process per GPU, utterances sharded, codes gathered on the caller
# ------------------------------------------------------------------------------------------------
def _engine_from_spec(spec: dict, device_index: int):
    """Default worker factory: the CUDA engine on `device_index`, weights from a checkpoint directory
    (``{"path": ...}``, reference layout: config.json + model.safetensors) or seeded synthetic weights
    (``{"model": "smoltts_byte_150m", "seed": 0}``, benchmarks / tests)."""
    import torch

    from .config import named_config
    from .model import RQTransformer
    from .synth import make_state_dict

    torch.cuda.set_device(device_index)
    kw = dict(spec.get("engine", {}))
    if "path" in spec:
        return RQTransformer.from_pretrained(spec["path"], **kw)
    cfg = named_config(spec["model"])
    model = RQTransformer(cfg, **kw)
    model.load_state_dict(make_state_dict(cfg, seed=int(spec.get("seed", 0))))
    return model


def _default_run(engine, prompts, settings, seq_ids, kwargs):
    from .generate import generate_batch

    return generate_batch(engine, prompts, settings, seq_ids=seq_ids, **kwargs)


def _worker_main(rank: int, spec: dict, factory, run, tasks, results):
    try:
        engine = factory(spec, rank)
        results.put((rank, "ready", None))
    except Exception as e:  # the caller re-raises: a GPU without the CUDA library must fail loudly, not idle
        results.put((rank, "error", repr(e)))
        return
    while True:
        job = tasks.get()
        if job is None:
            return
        job_id, prompts, settings, seq_ids, kwargs = job
        try:
            outs = run(engine, prompts, settings, seq_ids, kwargs) if prompts else []
            results.put((rank, job_id, [t.cpu().contiguous() for t in outs]))
        except Exception as e:
            results.put((rank, "error", repr(e)))


class ShardedGenerator:
    """``gpus`` worker processes, each with a full weight replica and a private KV pool on its own GPU (SURVEY 8(e):
    the path shards by utterance, no collective on the decode path).  ``generate`` gives worker r the contiguous
    slice ``partition(len(prompts), gpus)[r]`` (``seq_id`` = global utterance index, so sampling does not depend on
    the sharding) and returns every utterance's codes in the caller's order."""

    def __init__(self, spec: dict, gpus: int, factory=_engine_from_spec, run=_default_run, start_method: str = "spawn"):
        import torch.multiprocessing as mp

        if gpus < 1:
            raise ValueError("gpus must be >= 1")
        self.gpus = gpus
        ctx = mp.get_context(start_method)
        self._results = ctx.Queue()
        self._tasks = [ctx.Queue() for _ in range(gpus)]
        self._procs = [ctx.Process(target=_worker_main, args=(r, spec, factory, run, self._tasks[r], self._results), daemon=True)
                       for r in range(gpus)]
        for p in self._procs:
            p.start()
        self._job = 0
        for _ in range(gpus):
            rank, tag, err = self._results.get()
            if tag == "error":
                self.close()
                raise RuntimeError(f"worker {rank} failed to start: {err}")

    def generate(self, prompts: Sequence[torch.Tensor], settings, **kwargs) -> List[torch.Tensor]:
        self._job += 1
        parts = partition(len(prompts), self.gpus)
        for r, (lo, hi) in enumerate(parts):
            self._tasks[r].put((self._job, [p.cpu() for p in prompts[lo:hi]], settings, list(range(lo, hi)), kwargs))
        got: List[Optional[list]] = [None] * self.gpus
        for _ in range(self.gpus):
            rank, tag, payload = self._results.get()
            if tag == "error":
                raise RuntimeError(f"worker {rank}: {payload}")
            assert tag == self._job
            got[rank] = payload
        return [t for part in got for t in part]

    def close(self) -> None:
        for q in self._tasks:
            try:
                q.put(None)
            except Exception:
                pass
        for p in self._procs:
            p.join(timeout=30)
            if p.is_alive():
                p.terminate()
        self._procs = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def generate_sharded(spec: dict, prompts: Sequence[torch.Tensor], settings, gpus: int, **kwargs) -> List[torch.Tensor]:
    """One-shot form: start the workers, decode, gather, stop.  ``spec`` as for ``ShardedGenerator``."""
    with ShardedGenerator(spec, gpus) as g:
        return g.generate(prompts, settings, **kwargs)
