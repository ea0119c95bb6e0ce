"""``modal.config``: the two values scripts read must be ``str``
(06_gpu_and_ml/llm-serving/openai_compatible/load_test.py:7-8)."""
import os

_profile = os.environ.get("MODAL_PROFILE", "local")


class _Config(dict):
    def get(self, key, default=None):
        return super().get(key, default)


config = _Config(environment=os.environ.get("MODAL_ENVIRONMENT", "main"), workspace="local", token_id="", token_secret="",
                 server_url="in-box://b200rt")
