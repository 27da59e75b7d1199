"""yololite — B200-native drop-in for the YOLO11 detection inference path of dongyunjinyu/YOLO-Lite.

Same import name and public surface as the reference package (`from yololite import YOLOLite`), but every
tensor op on the path runs in hand-written sm_100a CUDA behind libyl11.so (include/yl11.h).  Unlike the
reference's __init__ (yololite/__init__.py:4-5) importing this package has no side effects on the process
environment.
"""

__version__ = "0.1.0"
__all__ = ("YOLOLite",)


def __getattr__(name):
    if name == "YOLOLite":
        from .engine.model import YOLOLite

        return YOLOLite
    raise AttributeError(name)
