"""Host-side partitioning used by the map pump, the examples and bench.py: items are independent, so the
path shards by contiguous item ranges with no data-path collective."""
from __future__ import annotations


def shard_bounds(n_items: int, world: int) -> list[tuple[int, int]]:
    """Contiguous [begin, end) per rank, sizes differing by at most one, in rank order."""
    if world < 1 or n_items < 0:
        raise ValueError("world must be >= 1 and n_items >= 0")
    base, extra = divmod(n_items, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


def wave_split(n_items: int, n_gpus: int) -> list[tuple[int, int]]:
    """The C++ dispatcher's split of one wave over the replica pool (engine.cu dispatcher_main):
    ceil(n/G) items per replica in order, trailing replicas possibly empty."""
    per = -(-n_items // n_gpus)
    return [(min(n_items, g * per), min(n_items, (g + 1) * per)) for g in range(n_gpus)]


def batches_of(n_items: int, batch: int, drop_remainder: bool = True):
    """Item ranges of the .map() inputs; the reference drops the final partial batch
    (06_gpu_and_ml/embeddings/text_embeddings_inference.py:156-163)."""
    full = n_items // batch
    for i in range(full):
        yield i * batch, (i + 1) * batch
    if not drop_remainder and n_items % batch:
        yield full * batch, n_items
