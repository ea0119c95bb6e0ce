"""GPU (`-m gpu`): the replica pool and the scheduler's contracts -- ONE process driving min(visible, 8) replicas through
the scatter kernel / fused peer gather (reference fan-out: text_embeddings_inference.py:79-86, :167),
ticket ownership between b200rt_wait and b200rt_poll_any, the zero-copy submit variant, length-bucket coalescing, and
device-order between the scheduler's stream and callers' streams on one replica's workspace."""
import threading

import numpy as np
import pytest

from oracle import bge_ref as R

pytestmark = pytest.mark.gpu


def _visible():
    import torch

    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def pool():
    import b200rt

    n = min(_visible(), 8)
    b200rt.init(n)
    g = R.BertGeometry(layers=2)
    flat = R.make_weights(g, 3, "trained")
    model = b200rt.EmbedModel(R.geometry_dict(g), R.pack_blob(flat, g))
    yield b200rt, model, g, flat, n
    b200rt.shutdown()


def test_pool_results_are_bit_identical_to_one_replica_and_match_the_oracle(pool):
    """The same inputs through a pool of G replicas (items scattered into peer HBM, rows gathered by each shard's last
    kernel) and through replica 0 alone (b200rt_embed_device) give the same bits; a sample is checked against the oracle."""
    import torch

    rt, model, g, flat, G = pool
    n = 40 * G + 5
    ids, lens = R.synth_ragged(n, 256, seed=11, min_len=2)
    s0 = rt.stats()
    got = model.embed(ids, lens)  # one ticket, one or more waves over all G replicas
    s1 = rt.stats()
    if G > 1:
        assert s1["peer_bytes"] > s0["peer_bytes"], "a multi-replica wave must move ids/rows over NVLink"
    # replica 0 alone, device-resident, in chunks
    with torch.cuda.device(0):
        d_ids = torch.from_numpy(ids).cuda()
        d_lens = torch.from_numpy(lens).cuda()
        d_out = torch.empty((n, 768), dtype=torch.float32, device="cuda")
        st = torch.cuda.Stream()
        step = 32
        for i in range(0, n, step):
            model.embed_device(0, d_ids[i:].data_ptr(), d_lens[i:].data_ptr(), min(step, n - i), 256, d_out[i:].data_ptr(), st.cuda_stream)
        st.synchronize()
        alone = d_out.cpu().numpy()
    assert np.array_equal(got, alone), "pool and single-replica results differ"
    ref = R.forward_np(flat, ids[:6], lens[:6], g, dtype=np.float64)
    assert R.rel_l2(got[:6], ref).max() <= 1e-3


def test_small_waves_rotate_over_replicas(pool):
    rt, model, g, flat, G = pool
    ids = R.synth_ids(2, 64, 5)
    base = model.embed(ids)
    for _ in range(2 * G + 1):  # consecutive small waves land on different replicas; every replica gives the same bits
        assert np.array_equal(model.embed(ids), base)


def test_wait_owns_its_ticket_against_poll_any(pool):
    """a ticket that a thread is blocked on in b200rt_wait must never be handed to b200rt_poll_any, whichever
    model object it was submitted through."""
    rt, model, g, flat, G = pool
    model2 = rt.EmbedModel(R.geometry_dict(g), R.pack_blob(flat, g))
    ids = R.synth_ids(64, 128, 9)
    results, errors = {}, []

    def waiter(k, m):
        try:
            t = m.submit(ids, tag=("w", k))
            results[("w", k)] = m.wait(t, 120_000)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ths = [threading.Thread(target=waiter, args=(k, model if k % 2 else model2)) for k in range(8)]
    polled = []
    for t in ths:
        t.start()
    own = [model.submit(ids, tag=("p", k)) for k in range(4)] + [model2.submit(ids, tag=("p", 4 + k)) for k in range(4)]
    while len(polled) < len(own):
        t = rt.poll_any(120_000)
        assert t is not None, "poll_any timed out"
        assert t.tag[0] == "p", f"poll_any returned a ticket owned by a blocked waiter: {t.tag}"
        polled.append(t)
    for t in ths:
        t.join()
    assert not errors, errors
    assert len(results) == 8 and all(v is not None for v in results.values())
    assert sorted(t.tag[1] for t in polled) == list(range(8))
    first = polled[0].out
    for t in polled:
        assert np.array_equal(t.out, first)
    for v in results.values():
        assert np.array_equal(v, first)
    assert rt.poll_any(50) is None


def test_zero_copy_submit_matches_the_copying_path(pool):
    rt, model, g, flat, G = pool
    n, S = 96, 128
    ids = R.synth_ids(n, S, 21)
    plain = model.embed(ids.copy())
    pin_ids = rt.PinnedBuffer((n, S), np.int32)
    pin_out = rt.PinnedBuffer((n, 768), np.float32)
    try:
        pin_ids.array[:] = ids
        s0 = rt.stats()
        tks = [model.submit(pin_ids.array[i:i + 32], None, out=pin_out.array[i:i + 32], borrow_ids=True) for i in range(0, n, 32)]
        for t in tks:
            assert model.wait(t, 60_000) is not None
        s1 = rt.stats()
        assert np.array_equal(pin_out.array, plain)
        assert s1["h2d_bytes"] - s0["h2d_bytes"] == n * (S * 4 + 4)
        # lent but pageable ids (no pinned registration): still correct, staged through the scheduler's own buffers
        assert np.array_equal(model.wait(model.submit(ids, borrow_ids=True)), plain)
        with pytest.raises(rt.B200RTError):
            model.submit(ids.astype(np.int64), borrow_ids=True)  # a converted temporary must not be lent
    finally:
        pin_ids.free()
        pin_out.free()


def test_tickets_of_different_max_len_share_a_wave(pool):
    """Length buckets of 64 tokens: inputs padded to 70, 100 and 128 tokens travel in one wave (padded to 128) and each
    item's embedding equals the one it gets alone at its own padded length."""
    rt, model, g, flat, G = pool
    specs = [(5, 70, 31), (7, 100, 32), (3, 128, 33), (4, 127, 34)]
    inputs = [R.synth_ragged(n, S, seed=sd, min_len=3) for n, S, sd in specs]
    solo = [model.embed(i, l) for i, l in inputs]
    n_fill = 6  # more full waves than wave slots: the dispatcher is blocked on a slot while the small tickets queue up
    s0 = rt.stats()
    for _ in range(n_fill):
        model.submit(R.synth_ids(rt.wave_capacity_items() * G, 512, 1))
    tks = [model.submit(i, l) for i, l in inputs]
    outs = [model.wait(t, 120_000) for t in tks]
    s1 = rt.stats()
    for o, s in zip(outs, solo):
        assert np.array_equal(o, s)
    assert s1["waves"] - s0["waves"] < n_fill + len(specs), "same-bucket tickets were not coalesced"
    for _ in range(n_fill):  # reap the filler tickets
        assert rt.poll_any(120_000) is not None


def test_ragged_input_is_split_into_length_buckets(pool):
    """One input with lengths 1..512: its items travel in 64-token length buckets (each padded to its own bucket, not to the
    input's max_len), come back in the input's order, and every item equals the embedding it gets alone."""
    rt, model, g, flat, G = pool
    n = 96
    ids, lens = R.synth_ragged(n, 512, seed=41, min_len=1)
    s0 = rt.stats()
    got = model.embed(ids, lens)
    s1 = rt.stats()
    tokens_padded_to_max = n * 512
    moved = s1["h2d_bytes"] - s0["h2d_bytes"] - 4 * n
    assert moved < 0.75 * 4 * tokens_padded_to_max, "ids were staged at max_len instead of per-bucket lengths"
    assert moved >= 4 * int(lens.sum())
    for i in (0, 1, 17, 50, 95):
        L = int(lens[i])
        alone = model.embed(ids[i:i + 1, :L].copy(), lens[i:i + 1])
        assert np.array_equal(got[i:i + 1], alone), i
    ref = R.forward_np(flat, ids[:4], lens[:4], g, dtype=np.float64)
    assert R.rel_l2(got[:4], ref).max() <= 1e-3
    # pinned, lent buffers take the same route
    pin_ids = rt.PinnedBuffer((n, 512), np.int32)
    pin_out = rt.PinnedBuffer((n, 768), np.float32)
    try:
        pin_ids.array[:] = ids
        assert model.wait(model.submit(pin_ids.array, lens, out=pin_out.array, borrow_ids=True), 60_000) is not None
        assert np.array_equal(pin_out.array, got)
    finally:
        pin_ids.free()
        pin_out.free()


def test_callers_stream_and_scheduler_are_ordered_on_the_workspace(pool):
    """b200rt_embed_device on a caller's stream and submit/wait traffic share one workspace per replica; the
    library orders them on the device, so interleaving them must not corrupt either."""
    import torch

    rt, model, g, flat, G = pool
    ids = R.synth_ids(48, 256, 77)
    ref = model.embed(ids)
    with torch.cuda.device(0):
        d_ids = torch.from_numpy(ids).cuda()
        d_lens = torch.full((48,), 256, dtype=torch.int32, device="cuda")
        outs = [torch.zeros((48, 768), dtype=torch.float32, device="cuda") for _ in range(6)]
        streams = [torch.cuda.Stream() for _ in range(3)]
        tks = []
        for k in range(6):
            model.embed_device(0, d_ids.data_ptr(), d_lens.data_ptr(), 48, 256, outs[k].data_ptr(), streams[k % 3].cuda_stream)
            tks.append(model.submit(ids))
        for t in tks:
            assert np.array_equal(model.wait(t, 60_000), ref)
        for s in streams:
            s.synchronize()
        for o in outs:
            assert np.array_equal(o.cpu().numpy(), ref)


def test_failed_ticket_is_reported_with_its_identity(pool):
    rt, model, g, flat, G = pool
    with pytest.raises(rt.B200RTError) as e:
        model.submit(np.full((2, 8), -1, np.int32), tag="bad")  # rejected at submit: never becomes a ticket
    assert e.value.code == rt.E_INVALID
    t = model.submit(R.synth_ids(1, 8, 1), tag="good")
    assert model.wait(t, 60_000) is not None
    with pytest.raises(rt.TicketError) as e2:
        model.wait(t, 10)  # already reaped: the error names the ticket
    assert e2.value.ticket is t
