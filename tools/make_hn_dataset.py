#!/usr/bin/env python
"""Seed the `tei-hn-data` volume with a synthetic `dataset.jsonl` in the reference's format -- one JSON array of
[id:int, text:str<=512 chars] pairs (06_gpu_and_ml/embeddings/text_embeddings_inference.py:117-127) -- so that
`modal run .../text_embeddings_inference.py::embed_dataset` can run unchanged in-box (the real script downloads
it from BigQuery, which needs the network).  With MODAL_SHIM_LINK_MOUNTS=1 the shim links /data to the volume."""
import argparse
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
import modal  # noqa: E402

WORDS = ("show hn ask launch rust python gpu kernel database startup open source release benchmark faster memory compiler "
         "browser linux cloud model training inference latency throughput embedding search index vector").split()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000)
    a = ap.parse_args()
    vol = modal.Volume.from_name("tei-hn-data", create_if_missing=True)
    rnd = random.Random(0)
    data = [[i, " ".join(rnd.choice(WORDS) for _ in range(rnd.randint(3, 90)))[:512]] for i in range(a.rows)]
    path = os.path.join(vol.local_path, "dataset.jsonl")
    with open(path, "w") as f:
        json.dump(data, f)
    print(f"wrote {a.rows} rows to {path}; /data linked: {vol.mount_at('/data')}")


if __name__ == "__main__":
    main()
