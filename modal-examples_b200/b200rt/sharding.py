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


MIN_ITEMS_PER_REPLICA = 8  # engine.cu: a wave uses fewer replicas rather than giving one fewer items than this


def wave_split(n_items: int, n_gpus: int, first: int = 0) -> list[tuple[int, int]]:
    """The C++ dispatcher's split of one wave over the replica pool (engine.cu dispatcher_main): as many replicas as get at
    least MIN_ITEMS_PER_REPLICA items each, contiguous ranges of ceil(n / used) items, starting at replica ``first`` (the
    dispatcher rotates it when a wave does not need the whole pool).  Returns one (begin, end) per replica; unused ones empty."""
    used = min(n_gpus, max(1, n_items // MIN_ITEMS_PER_REPLICA))
    if used == n_gpus:
        first = 0
    per = -(-n_items // used) if n_items else 0
    out = [(0, 0)] * n_gpus
    for k in range(used):
        out[(first + k) % n_gpus] = (min(n_items, k * per), min(n_items, (k + 1) * per))
    return out


def bucket_of(length: int, max_seq: int = 512) -> int:
    """Length bucket of an item (engine.cu bucket_of): items travel padded to the next multiple of 64 tokens."""
    return min(max_seq, (int(length) + 63) // 64 * 64)


def length_runs(lens) -> list[tuple[int, list[int]]]:
    """How b200rt_submit groups one input's items (engine.cu submit_impl): stable sort by bucket, one run per bucket.
    Returns [(bucket, [item indices in travel order])]."""
    order = sorted(range(len(lens)), key=lambda i: bucket_of(lens[i]))
    runs: list[tuple[int, list[int]]] = []
    for i in order:
        b = bucket_of(lens[i])
        if runs and runs[-1][0] == b:
            runs[-1][1].append(i)
        else:
            runs.append((b, [i]))
    return runs


def batches_of(n_items: int, batch: int, drop_remainder: bool = True):
    """Item ranges of the .map() inputs; the reference drops the final partial batch
    (06_gpu_and_ml/embeddings/text_embeddings_inference.py:156-163)."""
    full = n_items // batch
    for i in range(full):
        yield i * batch, (i + 1) * batch
    if not drop_remainder and n_items % batch:
        yield full * batch, n_items
