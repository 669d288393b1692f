"""Generate the golden vectors under tests/golden/ by EXECUTING the reference where that is possible.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

  gaussian_kernel.npz   output of the reference's `layer_blocks.gaussian_kernel` (layer_blocks.py:980-1002)
  coord_channels.npz    output of the reference's `CoordinateChannel2D` (coord.py:88-133) on a numpy
                        Keras-backend shim (oracle/ref_shim.py)
  compile_losses.npz    outputs of the reference's OWN loss closures (multiscale_vae.py:453-495: vae_r_loss,
                        vae_r_experimental_loss, vae_kl_loss, vae_loss), captured by running MultiscaleVAE.compile on a
                        recording stub (oracle/ref_shim.py load_reference_losses) and evaluated on seeded arrays
  schedule.npz          the reference's step_decay_schedule (schedule.py:7-21) over a grid of arguments
  train_fit_call.json   what the reference's train() (multiscale_vae.py:508-557) passes to keras fit()
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

    # compile() losses: the reference's closures on seeded arrays ---------------------------------------------
    out = {}
    cases = [("a", (4, 32, 32, 3), 24, 1.0, 0.1), ("b", (3, 16, 8, 3), 10, 1.0, 1.0), ("c", (2, 10, 6, 1), 5, 0.5, 2.0),
             ("d", (5, 4, 4, 2), 3, 1.0, 0.1), ("e", (2, 64, 64, 3), 8, 1.0, 0.1)]
    for name, shape, zt, rf, kf in cases:
        y = rng.random(shape) * 255.0
        yh = np.clip(y + rng.standard_normal(shape) * 30.0, 0.0, 255.0)
        mu = rng.standard_normal((shape[0], zt))
        lv = rng.standard_normal((shape[0], zt)) * 0.5
        ref = ref_shim.load_reference_losses(shape[1:], mu, lv, rf, kf)
        out.update({name + "_y": y, name + "_yh": yh, name + "_mu": mu, name + "_lv": lv,
                    name + "_factors": np.array([rf, kf]),
                    name + "_vae_r_loss": ref["metrics"][0](y, yh),
                    name + "_vae_kl_loss": ref["metrics"][1](y, yh),
                    name + "_vae_r_experimental_loss": ref["vae_r_experimental_loss"](y, yh),
                    name + "_vae_loss": ref["loss"](y, yh)})
        assert ref["optimizer"].kwargs == {"lr": 0.01, "clipnorm": 1.0}
    np.savez(os.path.join(HERE, "compile_losses.npz"), **out)

    # schedule.py -----------------------------------------------------------------------------------------------
    _, sched = ref_shim.load_reference_model_module()
    rows = []
    for lr0, decay, step in [(0.01, 0.5, 1), (0.01, 1, 1), (0.001, 0.9, 3), (0.05, 0.25, 2)]:
        fn = sched.step_decay_schedule(initial_lr=lr0, decay_factor=decay, step_size=step).schedule
        rows += [[lr0, decay, step, e, fn(e)] for e in range(8)]
    np.savez(os.path.join(HERE, "schedule.npz"), rows=np.array(rows, dtype=np.float64))

    # train(): what reaches fit() ---------------------------------------------------------------------------------
    import json
    import tempfile
    calls = {}
    for tag, kw in [("plain", {}), ("resume_ckpt", dict(initial_epoch=2, save_checkpoint_weights=True, lr_decay=0.5))]:
        d = tempfile.mkdtemp()
        x = np.zeros((20, 8, 8, 3), dtype=np.float32)
        args, kwargs = ref_shim.run_reference_train(x, 8, 5, d, **kw)
        cbs = kwargs.pop("callbacks")
        calls[tag] = dict(x_is_target=args[0] is args[1], kwargs=kwargs,
                          callbacks=[type(c).__name__ for c in cbs],
                          checkpoint_files=[os.path.relpath(c.args[0], d) for c in cbs if type(c).__name__ == "ModelCheckpoint"],
                          viz_first_n=int(cbs[1].args[3].shape[0]), viz_every=cbs[1].args[1])
    with open(os.path.join(HERE, "train_fit_call.json"), "w") as f:
        json.dump(calls, f, indent=1, sort_keys=True)

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
