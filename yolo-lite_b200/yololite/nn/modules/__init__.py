"""nn modules of the YOLO11 detection path (reference yololite/nn/modules/__init__.py exports many more
model families; only what cfg/yolo11.yaml instantiates is in scope)."""
from .block import C2PSA, C3, DFL, SPPF, Attention, Bottleneck, C2f, C3k, C3k2, PSABlock
from .conv import Concat, Conv, DWConv, autopad
from .head import Detect

__all__ = ("Conv", "DWConv", "Concat", "autopad", "DFL", "SPPF", "C2f", "C3", "C3k", "C3k2", "Bottleneck",
           "Attention", "PSABlock", "C2PSA", "Detect")
