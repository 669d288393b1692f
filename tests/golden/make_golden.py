"""Generate the golden vectors under tests/golden/ by EXECUTING the reference where that is possible.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

  gaussian_kernel.npz   output of the reference's `layer_blocks.gaussian_kernel` (layer_blocks.py:980-1002)
  coord_channels.npz    output of the reference's `CoordinateChannel2D` (coord.py:88-133) on a numpy
                        Keras-backend shim (oracle/ref_shim.py)
  step_cfg1_small.npz   NOT from the reference (TensorFlow is unavailable): outputs of the fp64 oracle on
                        a seeded tiny model, kept so the oracle itself cannot drift silently.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def main():
    from oracle import ref_shim
    lb, coord = ref_shim.load_reference()

    out = {}
    for size, nsig in [((3, 3), (2, 2)), ((3, 3), (1, 1)), ((5, 5), (2, 2)), ((3, 5), (1, 2))]:
        out[f"k{size[0]}x{size[1]}_n{nsig[0]}x{nsig[1]}"] = lb.gaussian_kernel(size, nsig)
    np.savez(os.path.join(HERE, "gaussian_kernel.npz"), **out)

    rng = np.random.default_rng(5)
    out = {}
    for name, shape, use_radius in [("a", (2, 4, 6, 3), False), ("b", (3, 5, 2, 1), True),
                                    ("c", (1, 8, 8, 3), True)]:
        x = rng.standard_normal(shape).astype(np.float32)
        layer = coord.CoordinateChannel2D(use_radius=use_radius)
        out[name + "_x"] = x
        out[name + "_y"] = np.asarray(layer(x), dtype=np.float32)
        out[name + "_r"] = np.asarray(use_radius)
    np.savez(os.path.join(HERE, "coord_channels.npz"), **out)

    # oracle self-pin (fp64) -----------------------------------------------------------------
    import torch
    from oracle.mvae_oracle import OracleMVAE
    torch.manual_seed(0)
    m = OracleMVAE((8, 8, 3), [4, 2], encoder={"filters": [8, 8], "kernel_size": [(3, 3), (3, 3)],
                                                "strides": [(2, 2), (1, 1)]},
                   sample_std=0.5, dtype=torch.float64, seed=11)
    m.compile(0.01, 1.0, 0.1)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(4, 8, 8, 3, generator=g, dtype=torch.float64) * 255
    eps = [torch.randn(4, z, generator=g, dtype=torch.float64) for z in m.z_dims]
    res, grads = m.loss_and_grads(x, eps)
    np.savez(os.path.join(HERE, "step_cfg1_small.npz"),
             loss=res["loss"].detach().numpy(), out=res["out"].detach().numpy(),
             kl=res["kl_loss"].detach().numpy(), r=res["r_loss"].detach().numpy(),
             gnorm=np.array([float(grads[n].norm()) for n in sorted(grads)]))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
