"""Inference sources (reference yololite/data/loaders.py LoadTensor :480-548, LoadPilAndNumpy :415-477,
LoadImagesAndVideos :248-412 for still images; data/build.py:143-176).  Each source iterates batches of
(paths, images, info-strings).  Video files, streams and screenshots are I/O features outside the hot path."""
from __future__ import annotations

from pathlib import Path
from types import SimpleNamespace

import cv2
import numpy as np
import torch
from PIL import Image

IMG_FORMATS = {"bmp", "dng", "jpeg", "jpg", "mpo", "png", "tif", "tiff", "webp", "pfm"}


class LoadTensor:
    """A BCHW image tensor (reference data/loaders.py:480-548).

    * float32 / float16: values in [0, 1], used as they are.
    * uint8: image bytes 0..255; the model's ingest kernel divides by 255 on the device (the reference crashes on a
      uint8 tensor — `torch.finfo(uint8)`, loaders.py:525 — so this is an extension: 4x fewer PCIe / HBM bytes).
    * The reference also accepts a FLOAT tensor holding 0..255 values: `im.max() > 1 + eps` -> warning + `/255`
      (loaders.py:525-530).  That check is a full pass over the batch on the host (more expensive than the whole
      GPU pipeline for a 64-image batch), so it is opt-in here: `LoadTensor.check_range = True` restores it."""

    #: run the reference's `im.max() > 1` range check (and rescale by 1/255) on float tensors
    check_range = False

    def __init__(self, im0: torch.Tensor, stride=32):
        if im0.dim() == 3:
            im0 = im0.unsqueeze(0)
        if im0.dim() != 4 or im0.shape[2] % stride or im0.shape[3] % stride:
            raise ValueError(f"torch.Tensor inputs should be BCHW i.e. shape(1, 3, 640, 640) divisible by stride "
                             f"{stride}. Input shape{tuple(im0.shape)} is incompatible.")
        if self.check_range and im0.is_floating_point() and im0.max() > 1.0 + torch.finfo(im0.dtype).eps:
            import warnings

            warnings.warn(f"torch.Tensor inputs should be normalized 0.0-1.0 but max value is {float(im0.max())}. "
                          f"Dividing input by 255.")
            im0 = im0.float() / 255.0
        self.im0 = im0
        self.bs = im0.shape[0]
        self.mode = "image"
        self.paths = [f"image{i}.jpg" for i in range(self.bs)]
        self.source_type = SimpleNamespace(stream=False, screenshot=False, from_img=False, tensor=True)

    def __iter__(self):
        yield self.paths, self.im0, [""] * self.bs

    def __len__(self):
        return self.bs


class LoadPilAndNumpy:
    """A list of HWC BGR numpy arrays and/or PIL images, served as one batch."""

    def __init__(self, im0):
        if not isinstance(im0, (list, tuple)):
            im0 = [im0]
        self.paths = [getattr(im, "filename", "") or f"image{i}.jpg" for i, im in enumerate(im0)]
        self.im0 = [self._single(im) for im in im0]
        self.bs = len(self.im0)
        self.mode = "image"
        self.source_type = SimpleNamespace(stream=False, screenshot=False, from_img=True, tensor=False)

    @staticmethod
    def _single(im):
        if isinstance(im, Image.Image):
            if im.mode != "RGB":
                im = im.convert("RGB")
            im = np.asarray(im)[:, :, ::-1]      # RGB -> BGR
        if not isinstance(im, np.ndarray) or im.ndim != 3:
            raise TypeError(f"expected a PIL image or an HWC numpy array, got {type(im)}")
        return np.ascontiguousarray(im)

    def __iter__(self):
        yield self.paths, self.im0, [""] * self.bs

    def __len__(self):
        return self.bs


class LoadImages:
    """Image files (a path, a directory or a list of paths), `batch` images per iteration."""

    def __init__(self, path, batch=1):
        files = []
        for p in (path if isinstance(path, (list, tuple)) else [path]):
            p = Path(p)
            if p.is_dir():
                files += sorted(str(q) for q in p.iterdir() if q.suffix[1:].lower() in IMG_FORMATS)
            elif p.is_file():
                files.append(str(p))
            else:
                raise FileNotFoundError(f"{p} does not exist")
        files = [f for f in files if Path(f).suffix[1:].lower() in IMG_FORMATS]
        if not files:
            raise FileNotFoundError(f"no images found in {path} (video / stream sources are out of scope)")
        self.files, self.bs, self.mode = files, max(int(batch), 1), "image"
        self.source_type = SimpleNamespace(stream=False, screenshot=False, from_img=False, tensor=False)

    def __iter__(self):
        for i in range(0, len(self.files), self.bs):
            chunk = self.files[i:i + self.bs]
            ims = []
            for f in chunk:
                im = cv2.imread(f)
                if im is None:
                    raise FileNotFoundError(f"Image Not Found {f}")
                ims.append(im)
            yield chunk, ims, [f"image {i + j + 1}/{len(self.files)} {f}: " for j, f in enumerate(chunk)]

    def __len__(self):
        return (len(self.files) + self.bs - 1) // self.bs


def load_inference_source(source=None, batch=1, vid_stride=1, buffer=False):
    if isinstance(source, torch.Tensor):
        return LoadTensor(source)
    if isinstance(source, (np.ndarray, Image.Image)):
        return LoadPilAndNumpy(source)
    if isinstance(source, (list, tuple)) and source and all(isinstance(s, (np.ndarray, Image.Image)) for s in source):
        return LoadPilAndNumpy(source)
    if isinstance(source, (str, Path)) or (isinstance(source, (list, tuple)) and source):
        return LoadImages(source, batch=batch)
    raise TypeError(f"unsupported inference source {type(source)} (tensor, numpy, PIL, image path(s) supported)")
