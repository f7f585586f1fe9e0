from posenet.models.mobilenet_v1 import MobileNetV1, MOBILENET_V1_CHECKPOINTS  # noqa: F401
