"""TEST INFRASTRUCTURE: records the `huggingface_hub` calls the reference's embedding scripts make instead of reaching the
network (see ../README.md): HfApi.create_repo / upload_folder (wikipedia/main.py:236-256), snapshot_download
(amazon_embeddings.py:128-131, text_embeddings_inference.py:54-56)."""
import json
import os

__version__ = "0.0-stub"


def _log(**kw):
    path = os.environ.get("FAKE_HF_LOG")
    if path:
        with open(path, "a") as f:
            f.write(json.dumps(kw) + "\n")


class HfApi:
    def __init__(self, token=None, **_kw):
        self.token = token

    def create_repo(self, repo_id, **kw):
        _log(call="create_repo", repo_id=repo_id, token=self.token, **{k: v for k, v in kw.items() if isinstance(v, (str, bool, int))})

    def upload_folder(self, folder_path, repo_id, **kw):
        _log(call="upload_folder", folder_path=folder_path, repo_id=repo_id, files=sorted(os.listdir(folder_path)))


def snapshot_download(repo_id, cache_dir=None, **_kw):
    _log(call="snapshot_download", repo_id=repo_id, cache_dir=cache_dir)
    return cache_dir or ""
