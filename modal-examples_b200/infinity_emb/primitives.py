"""The two enums the reference's image example passes to ``EngineArgs`` (image_embeddings_infinity.py:127, 302-303)."""
import enum


class Dtype(str, enum.Enum):
    float32 = "float32"
    float16 = "float16"
    bfloat16 = "bfloat16"
    auto = "auto"


class InferenceEngine(str, enum.Enum):
    torch = "torch"
    optimum = "optimum"
    ctranslate2 = "ctranslate2"
    neuron = "neuron"
    debugengine = "debugengine"


class Device(str, enum.Enum):
    cpu = "cpu"
    cuda = "cuda"
    auto = "auto"
