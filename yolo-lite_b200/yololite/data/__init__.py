"""Inference-side data handling: LetterBox and the in-memory / file sources of the predictor.
(The reference's dataset, augmentation and stream loaders serve training and video I/O: out of scope.)"""
from .augment import LetterBox, letterbox_batch_cuda
from .loaders import load_inference_source

__all__ = ("LetterBox", "letterbox_batch_cuda", "load_inference_source")
