"""GPU: Face 2 of the boundary -- the `text-embeddings-router` stand-in driven the way the reference drives
TEI: Popen with the reference's flags, TCP readiness poll (text_embeddings_inference.py:37-51), then
`POST /embed {"inputs": [...]}` -> list of 768-vectors in input order (:97-104)."""
import json
import os
import socket
import subprocess
import time
import urllib.error
import urllib.request

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "modal-examples_b200")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _post(port, payload, path="/embed"):
    req = urllib.request.Request(f"http://127.0.0.1:{port}{path}", data=json.dumps(payload).encode(), headers={"Content-Type": "application/json"})
    try:
        with urllib.request.urlopen(req, timeout=120) as r:
            return r.status, json.loads(r.read())
    except urllib.error.HTTPError as e:
        return e.code, json.loads(e.read())


@pytest.fixture(scope="module")
def router():
    port = _free_port()
    env = dict(os.environ, PATH=os.path.join(PKG, "bin") + os.pathsep + os.environ["PATH"], B200RT_GPUS="1")
    proc = subprocess.Popen(["text-embeddings-router", "--model-id", "BAAI/bge-base-en-v1.5", "--port", str(port)], env=env)
    deadline = time.time() + 300
    while True:  # the reference's readiness loop
        try:
            socket.create_connection(("127.0.0.1", port), timeout=1).close()
            break
        except (socket.timeout, ConnectionRefusedError, OSError):
            assert proc.poll() is None, f"launcher exited unexpectedly with code {proc.returncode}"
            assert time.time() < deadline, "router did not become ready"
            time.sleep(0.2)
    yield port
    proc.terminate()
    proc.wait(timeout=60)


def test_embed_contract_and_parity_with_direct_engine(router):
    import b200rt
    from tei_router.server import GEOMETRY, random_blob
    from tei_router.tokenizer import WordPiece

    texts = [f"Show HN: item {i} embeds {'very ' * (i % 7)}long sentences on eight B200s, naïve café #{i}!" for i in range(32)]
    status, out = _post(router, {"inputs": texts})
    assert status == 200 and len(out) == 32 and all(len(v) == 768 for v in out)
    got = np.array(out, np.float32)
    assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)
    # the same token ids through the C ABI in this process must give the same vectors (same weights, same kernels)
    b200rt.init(devices=[0])
    model = b200rt.EmbedModel(GEOMETRY, random_blob())
    tok = WordPiece()
    rows = [tok.encode(t, 512, True) for t in texts]
    lens = np.array([len(r) for r in rows], np.int32)
    ids = np.zeros((32, int(lens.max())), np.int32)
    for i, r in enumerate(rows):
        ids[i, : len(r)] = r
    direct = model.embed(ids, lens)
    assert np.abs(direct - got).max() < 1e-6
    b200rt.shutdown()
    # a single string is accepted like a list of one
    status, one = _post(router, {"inputs": texts[3]})
    assert status == 200 and np.abs(np.array(one[0], np.float32) - got[3]).max() < 1e-6


def test_error_behaviour(router):
    status, body = _post(router, {"inputs": ["x"] * 33})  # > --max-client-batch-size 32 (TEI default, = BATCH_SIZE)
    assert status == 413 and body["error_type"] == "Validation"
    status, body = _post(router, {"wrong": 1})
    assert status == 422
    status, _ = _post(router, {"inputs": ["x"]}, path="/nope")
    assert status == 404
    with urllib.request.urlopen(f"http://127.0.0.1:{router}/health", timeout=10) as r:
        assert r.status == 200
