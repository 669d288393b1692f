"""Visualisation callback of the reference (mvae/callbacks.py:16-138) as plain host code.

Every `print_every_n_batches` batches it writes three collages under `<run_folder>/images/`:

    img_<epoch>_<batch>.png             reconstructions  decode(encode(images))          (callbacks.py:77-84)
    samples_<epoch>_<batch>.png         decode(z), z ~ N(mean(enc), std(enc))            (callbacks.py:86-113)
    interpolations_<epoch>_<batch>.png  decode of linear mixes of consecutive encodings  (callbacks.py:115-134)

It goes through the z-domain entry points `vae.model_encode.predict` / `vae.model_decode.predict`, like the reference.
The reference imports a `collage` helper that does not exist at its HEAD and needs matplotlib and
scikit-image, which this image does not have: the collage here is a square grid, the resize is nearest-neighbour
(`skimage.transform.resize(order=0)`), and PIL writes the PNG (single-channel images inverted like cmap "gray_r").
"""
from __future__ import annotations

import math
import os

import numpy as np

from .custom_logger import logger


def collage(images: np.ndarray) -> np.ndarray:
    """(N, H, W, C) in [0, 1] -> one (rows*H, cols*W, C) grid, cols = ceil(sqrt(N)), unused cells zero."""
    images = np.asarray(images)
    n, h, w, c = images.shape
    cols = int(math.ceil(math.sqrt(n)))
    rows = int(math.ceil(n / cols))
    out = np.zeros((rows * h, cols * w, c), dtype=images.dtype)
    for i in range(n):
        r, q = divmod(i, cols)
        out[r * h:(r + 1) * h, q * w:(q + 1) * w, :] = images[i]
    return out


def resize_nearest(x: np.ndarray, shape) -> np.ndarray:
    """skimage.transform.resize(x, shape, order=0): nearest-neighbour sampling at pixel centres."""
    h, w = x.shape[:2]
    ys = np.minimum((np.arange(shape[0]) + 0.5) * h / shape[0], h - 1).astype(np.int64)
    xs = np.minimum((np.arange(shape[1]) + 0.5) * w / shape[1], w - 1).astype(np.int64)
    return x[ys][:, xs]


class SaveIntermediateResultsCallback:
    def __init__(self, run_folder, print_every_n_batches, initial_epoch, images, vae, resize_shape=(256, 256), seed=0):
        self._vae = vae
        self._images = np.asarray(images, dtype=np.float32)
        self._epoch = initial_epoch
        self._run_folder = run_folder
        self._resize_shape = resize_shape
        self._print_every_n_batches = print_every_n_batches
        self._images_path = os.path.join(self._run_folder, "images")
        self._rng = np.random.default_rng(seed)
        os.makedirs(self._images_path, exist_ok=True)

    def save_collage(self, samples: np.ndarray, batch: int, prefix: str) -> str:
        from PIL import Image
        x = np.clip(self._vae.normalize(np.asarray(samples, dtype=np.float32)), 0.0, 1.0)   # [0, 1] (callbacks.py:50-51)
        x = resize_nearest(collage(x), self._resize_shape)
        path = os.path.join(self._images_path, f"{prefix}_" + str(self._epoch).zfill(3) + "_" + str(batch) + ".png")
        if x.shape[-1] == 1:
            Image.fromarray(np.uint8(np.round((1.0 - x[..., 0]) * 255.0)), mode="L").save(path)      # cmap "gray_r"
        else:
            Image.fromarray(np.uint8(np.round(x[..., :3] * 255.0)), mode="RGB").save(path)
        return path

    def interpolations(self, encodings: np.ndarray) -> np.ndarray:
        """callbacks.py:115-130: row j mixes encoding j into encoding j+1 over sqrt(N) steps."""
        n = encodings.shape[0]
        out = np.zeros_like(encodings)
        s = int(round(math.sqrt(n)))
        for j in range(s):
            start, end = encodings[j, :], encodings[min(j + 1, n - 1), :]
            for i in range(s):
                k = j * s + i
                if k >= n:
                    continue
                mix = float(i) / float(max(s - 1, 1))
                out[k, :] = start * (1.0 - mix) + end * mix
        return out

    def on_batch_end(self, batch, logs=None):
        if batch % self._print_every_n_batches != 0:
            return []
        written = []
        encodings = self._vae.model_encode.predict(self._images)
        written.append(self.save_collage(self._vae.model_decode.predict(encodings), batch, "img"))
        mean, std = float(np.mean(encodings)), float(np.std(encodings))
        logger.info("encodings_mean: {0:.4g}, encodings_std: {1:.4g}".format(mean, std))
        z = self._rng.normal(loc=mean, scale=std, size=encodings.shape).astype(np.float32)
        written.append(self.save_collage(self._vae.model_decode.predict(z), batch, "samples"))
        written.append(self.save_collage(self._vae.model_decode.predict(self.interpolations(encodings)), batch,
                                         "interpolations"))
        return written

    def on_epoch_begin(self, epoch, logs=None):
        self._epoch += 1
