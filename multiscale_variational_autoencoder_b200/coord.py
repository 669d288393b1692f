"""CoordinateChannel layers (mirrors the rank-2, channels_last path of mvae/coord.py:88-133).

In the model the channels are never materialised: `MultiscaleVAE(coord_conv="xy"|"xyr")` generates them inside the
conv_base A-operand load (mvae_conv2d_fwd coord_mode).  This standalone layer exists for API parity."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


class CoordinateChannel2D:
    def __init__(self, use_radius=False, data_format=None, **kwargs):
        if data_format not in [None, "channels_first", "channels_last"]:
            raise ValueError('`data_format` must be either "channels_last", "channels_first" or None.')
        if data_format == "channels_first":
            raise NotImplementedError("only channels_last is supported (the reference model is NHWC)")
        self.rank, self.use_radius, self.data_format = 2, use_radius, "channels_last"

    def __call__(self, inputs):
        was_numpy = not torch.is_tensor(inputs)
        x = torch.as_tensor(np.asarray(inputs, dtype=np.float32) if was_numpy else inputs)
        if x.dim() != 4:
            raise ValueError("CoordinateChannel2D expects (batch, H, W, channels)")
        dev = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        _lib.require_b200(dev.index or 0)
        x = x.to(dev, torch.float32).contiguous()
        B, H, W, C = x.shape
        y = torch.empty((B, H, W, C + (3 if self.use_radius else 2)), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().mvae_coord_channels(x.data_ptr(), y.data_ptr(), B, H, W, C, int(self.use_radius),
                                                       torch.cuda.current_stream(dev).cuda_stream), "coord_channels")
        return y.cpu().numpy() if was_numpy else y

    call = __call__

    def compute_output_shape(self, input_shape):
        out = list(input_shape)
        out[-1] = input_shape[-1] + (3 if self.use_radius else 2)
        return tuple(out)

    def get_config(self):
        return {"rank": 2, "use_radius": self.use_radius, "data_format": self.data_format}
