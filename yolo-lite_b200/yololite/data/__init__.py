"""Inference-side data handling: LetterBox and the in-memory / file sources of the predictor.
`dataset.RectValLoader` is the minimal rect validation loader (harness for `YOLOLite.val(data=...)`); the reference's
training dataset, augmentation and stream loaders serve training and video I/O: out of scope."""
from .augment import LetterBox, letterbox_batch_cuda
from .dataset import RectValLoader, build_val_loader
from .loaders import load_inference_source

__all__ = ("LetterBox", "letterbox_batch_cuda", "load_inference_source", "RectValLoader", "build_val_loader")
