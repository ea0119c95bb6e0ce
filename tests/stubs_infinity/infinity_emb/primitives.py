import enum


class Dtype(str, enum.Enum):
    float32 = "float32"
    float16 = "float16"


class InferenceEngine(str, enum.Enum):
    torch = "torch"
    optimum = "optimum"
