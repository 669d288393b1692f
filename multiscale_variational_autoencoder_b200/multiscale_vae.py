"""`MultiscaleVAE`: the reference's model-builder API (mvae/multiscale_vae.py:11-587) on B200 kernels.

Same constructor / `compile` / `train` / `encoder` / `decoder` / `model_trainable` / `learning_rate` / `normalize`
surface as the reference class; extra keyword-only switches settle the north-star-vs-reference differences
 and default to the behaviour of mvae/multiscale_vae.py:

    coord_conv   None | "xy" | "xyr"   CoordinateChannel2D (coord.py:88-133) in front of every conv_base
    logvar_scale 1.0 (multiscale_vae.py:378) | 0.5 (multiscale_vae_.py:34)
    diff_mode    "no_upsample" (multiscale_vae.py:314) | "laplacian" (layer_blocks.py:74)
    precision    "fp32" (CUDA-core, tight parity) | "tf32" (tcgen05 tensor cores)

`train()` applies the training-time corruption of the reference's input transform (GaussianNoise(1/(max-min)) in
normalised space + SpatialDropout2D(0.1), multiscale_vae.py:58-59,139-147) with device-side RNG, like `fit` does in the
training phase; `train_on_batch` (parity tests, bench) feeds the clean batch unless `corrupt=True`.  eps of the sampling
layer is drawn on the device per step, or supplied by the caller.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch

from . import _lib, callbacks, schedule
from .custom_logger import logger
from .engine import Engine, ParamStore, Spec
from ._lib import PREC_FP32, PREC_TF32


class _ModelView:
    """Stands in for the keras.Model objects returned by `.encoder`, `.decoder`, `.model_trainable`."""

    def __init__(self, owner, kind):
        self._o, self._kind = owner, kind

    def predict(self, x, batch_size=None, eps=None):
        return getattr(self._o, "_predict_" + self._kind)(np.asarray(x, dtype=np.float32), eps)

    __call__ = predict

    def to_json(self):
        o = self._o
        return json.dumps(dict(class_name="MultiscaleVAE", model=self._kind, input_dims=list(o._inputs_dims),
                               z_dims=list(o._z_latent_dims), encoder=o._encoder_config, decoder=o._decoder_config,
                               min_value=o._min_value, max_value=o._max_value, sample_std=o._sample_std,
                               variables={k: list(v["shape"]) for k, v in o._ps.entries.items()}))

    def summary(self, print_fn=print):
        o = self._o
        tot = tr = 0
        print_fn(f'Model: "{o._name}_{self._kind}"')
        for k, e in o._ps.entries.items():
            n = int(np.prod(e["shape"]))
            tot += n
            tr += n if e["trainable"] else 0
            print_fn(f"  {k:70s} {str(e['shape']):22s} {n}")
        print_fn(f"Total params: {tot}\nTrainable params: {tr}\nNon-trainable params: {tot - tr}")

    def count_params(self):
        return sum(int(np.prod(e["shape"])) for e in self._o._ps.entries.values())


class MultiscaleVAE:
    def __init__(self, input_dims, z_dims, compress_output=False,
                 encoder={"filters": [32], "kernel_size": [(3, 3)], "strides": [(1, 1)]},
                 decoder=None, min_value=0.0, max_value=255.0, sample_std=0.01, channels_index=2, *,
                 coord_conv=None, logvar_scale=1.0, diff_mode="no_upsample", precision="fp32", device=None, seed=7):
        # --- argument checking (multiscale_vae.py:35-38)
        if encoder is None:
            raise ValueError("encoder cannot be None")
        if not all(i > 0 for i in z_dims):
            raise ValueError("z_dims elements should be > 0")
        if channels_index != 2 or len(input_dims) != 3:
            raise ValueError("only HxWxC inputs (channels_index=2) are supported")
        # --- decoder is reverse encoder (multiscale_vae.py:40-45)
        if decoder is None:
            decoder = {"filters": encoder["filters"][::-1], "strides": encoder["strides"][::-1],
                       "kernel_size": encoder["kernel_size"][::-1]}
        self._name = "mvae"
        self._levels = len(z_dims)
        self._z_latent_dims = list(z_dims)
        self._inputs_dims = tuple(input_dims)
        self._encoder_config, self._decoder_config = encoder, decoder
        self._compress_output = compress_output          # stored, unused (as in the reference, :60)
        self._min_value, self._max_value = float(min_value), float(max_value)
        self._sample_std = sample_std
        self._channels_index = channels_index
        self._training_dropout = 0.1                                     # multiscale_vae.py:58
        self._training_noise_std = 1.0 / (self._max_value - self._min_value)   # multiscale_vae.py:59
        self._precision = {"fp32": PREC_FP32, "tf32": PREC_TF32}[precision]
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self._device = torch.device(device)
        _lib.require_b200(self._device.index or 0)
        self._spec = Spec(input_dims, z_dims, encoder, decoder, min_value, max_value, sample_std, coord_conv,
                          logvar_scale, diff_mode)
        self._build(seed)

    # ==========================================================================================================
    def _build(self, seed):
        logger.info("Building multiscale VAE plan")
        self._ps = ParamStore(self._device, seed)
        self._spec.declare_params(self._ps)
        with torch.cuda.device(self._device):
            self._ps.finalize()
        self._engines = {}
        self._graphs = {}
        self._model_encoder = _ModelView(self, "encoder")
        self._model_decoder = _ModelView(self, "decoder")
        self._model_trainable = _ModelView(self, "trainable")
        self._learning_rate = None
        self._lr_dev = torch.zeros(1, dtype=torch.float32, device=self._device)
        self._r_loss_factor = self._kl_loss_factor = 1.0
        self._clip_norm = 1.0
        self._gen = torch.Generator(device=self._device).manual_seed(4321)
        self._world, self._dist = 1, None
        self.use_cuda_graph = True
        self.parallel_levels = True

    def _engine(self, B, training, corrupt=False):
        key = (int(B), bool(training), bool(corrupt))
        if key not in self._engines:
            with torch.cuda.device(self._device):
                e = Engine(self._spec, self._ps, B, training, self._precision, corrupt=corrupt)
            e.r_factor, e.kl_factor = self._r_loss_factor, self._kl_loss_factor
            self._engines[key] = e
        return self._engines[key]

    # ==========================================================================================================
    def compile(self, learning_rate, r_loss_factor=1.0, kl_loss_factor=1.0, clip_norm=1.0):
        """multiscale_vae.py:437-504: loss = r*r_factor + kl*kl_factor (+ regularisers), Adagrad(lr, clipnorm)."""
        self.learning_rate = learning_rate
        self._r_loss_factor, self._kl_loss_factor, self._clip_norm = float(r_loss_factor), float(kl_loss_factor), clip_norm
        self._ps.acc = torch.full_like(self._ps.flat, 0.1)      # Keras Adagrad initial_accumulator_value
        for e in self._engines.values():
            e.r_factor, e.kl_factor = self._r_loss_factor, self._kl_loss_factor
        self._graphs.clear()

    def enable_data_parallel(self, bucket_mb=32.0):
        """Data-parallel over the batch: one process per GPU, gradients all-reduced in buckets over NCCL.  Weights are
        broadcast from rank 0; every rank draws its own eps / noise / dropout masks (device RNG seeded with 4321 + rank)
        and `train()` gives it its own shard of every epoch's permutation, so the replicas see different samples."""
        from .dist import GradAllReduce
        self._dist = GradAllReduce(self._ps, self._device, bucket_mb)
        self._world = self._dist.world
        self._dist.broadcast_params()
        self._gen.manual_seed(4321 + self._dist.rank)
        self._graphs.clear()
        # Default: ONE bucketed all-reduce of the flat gradient buffer between the backward graph and the optimiser graph.
        # MVAE_DP_INGRAPH=1 exchanges per level from inside the step instead (NCCL captured into the graph, launched from
        # each level's stream when its backward ends).  Measured on 2 x B200, cfg2: 1.58 ms/step against 1.39 -- the NCCL
        # kernels of ten small exchanges take SMs from the fused tile kernels, which need all 148 SMs for their 256 tiles --
        # and the process did not exit cleanly with NCCL work captured in live graphs, so it stays an experiment.
        self._dp_ingraph = os.environ.get("MVAE_DP_INGRAPH", "0") == "1"
        self._dp_ranges = self._dist.level_ranges(self._levels)
        self._dp_early = self._dist.early_ranges()

    # ---- one training step ----------------------------------------------------------------------------------
    def _step_body(self, eng):
        d = self._dist
        if d is None or (d.peer is None and not self._dp_ingraph):
            eng.forward_backward(parallel=self.parallel_levels)     # NCCL transport: the caller exchanges after the graph
            return
        if d.peer is not None and not self._dp_ingraph:
            self._peer_step(eng)
            return
        works = []
        eng.on_level_grads = lambda i: works.extend(d.allreduce_level(self._dp_ranges[i]))
        try:
            eng.forward_backward(parallel=self.parallel_levels)
        finally:
            eng.on_level_grads = None
        eng._stream()
        d.wait_all(works)          # the step's stream continues (optimiser) once every level has been exchanged

    def _peer_step(self, eng):
        """Forward + backward with the peer-memory exchange inside (captured with the step).  The big Dense weight gradients
        are exchanged the moment they exist: the kernel goes out on the side stream that carries the Dense weight-gradient
        launch, behind it, and runs under the rest of the level's backward chain (one channel per level, so two levels'
        exchanges may overlap).  What is left -- the many small convolution gradients -- is one multi-range launch at the
        end of the step."""
        d, dev = self._dist, self._device
        per_level = not (eng.batch_levels and eng._batched) and not eng._use_coarse()     # one chain per level: g(i) knows i
        early = self._dp_early if (per_level and os.environ.get("MVAE_DP_EARLY", "1") == "1") else {}
        ctas = int(os.environ.get("MVAE_DP_EARLY_CTAS", "32"))
        prio = os.environ.get("MVAE_DP_EARLY_PRIO", "1") == "1"
        done = []

        def dense_wgrad_issued(op, level):
            r = early.get(op.wname)
            if r is None or level >= d.peer.CHANNELS - 1:
                return
            # lane 0 of the level's side streams is where the Dense weight gradient was just launched.  The exchange follows
            # it on a lane of its own with the HIGHEST stream priority: the weight-gradient lanes have the lowest, and an
            # exchange queued there would only be scheduled once the chains leave SMs idle -- at the end of the backward
            # pass, where nothing hides it any more
            if prio:
                cur = torch.cuda.current_stream(dev)
                lane0 = eng._side_streams.get((cur.cuda_stream, 0))
                def launch():
                    if lane0 is not None:
                        torch.cuda.current_stream(dev).wait_stream(lane0)
                    d.peer.allreduce(ranges=[r], channel=level, ctas=ctas, stream=torch.cuda.current_stream(dev).cuda_stream)
                eng.side(launch, lane="comm", priority=2)
            else:
                eng.side(lambda: d.peer.allreduce(ranges=[r], channel=level, ctas=ctas, stream=eng.s), lane=0)
            done.append(r)

        eng.on_dense_wgrad = dense_wgrad_issued if early else None
        try:
            eng.forward_backward(parallel=self.parallel_levels)
        finally:
            eng.on_dense_wgrad = None
        eng._stream()
        d.peer.allreduce(ranges=d.leftover_ranges(done), channel=d.peer.CHANNELS - 1)

    def _opt_body(self, eng):
        eng.optimizer_step(self._lr_dev, self._clip_norm, 1.0 / self._world)

    def train_step_device(self, eng):
        """Enqueue one step on eng.x / eng.eps (already on the device).  Returns nothing; read eng.scalars later."""
        if self._ps.acc is None:
            raise RuntimeError("call compile() before training")
        # NCCL transport: one exchange between the backward graph and the optimiser graph, issued from the host
        after = self._dist is not None and not self._dp_ingraph and self._dist.peer is None
        if not self.use_cuda_graph:
            self._step_body(eng)
            if after:
                self._dist.allreduce()
            self._opt_body(eng)
            return
        key = id(eng)
        if key not in self._graphs:
            # warm-up outside capture (lazy module load, stream creation), then capture
            side = torch.cuda.Stream(self._device)
            side.wait_stream(torch.cuda.current_stream(self._device))
            with torch.cuda.stream(side):
                snap_p, snap_a = self._ps.flat.clone(), self._ps.acc.clone()
                self._step_body(eng)
                self._opt_body(eng)
                self._ps.flat.copy_(snap_p)
                self._ps.acc.copy_(snap_a)
            torch.cuda.current_stream(self._device).wait_stream(side)
            torch.cuda.synchronize(self._device)
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            lib = _lib.load()
            n0 = lib.mvae_kernel_launch_count()
            with torch.cuda.graph(g1):
                self._step_body(eng)
            with torch.cuda.graph(g2):
                self._opt_body(eng)
            # kernel nodes of one replay (every launch of the library counts; the memsets of the arenas do not)
            eng.kernels_per_step = int(lib.mvae_kernel_launch_count() - n0)
            self._graphs[key] = (g1, g2)
        g1, g2 = self._graphs[key]
        g1.replay()
        if after:
            self._dist.allreduce()
        g2.replay()

    # ---- input pipeline: the next batch travels host -> device on a copy stream while the current step runs -----------
    def stage_batch(self, eng, x_pinned, eps_pinned=None):
        """Enqueue the H2D copies of a FUTURE step's inputs (pinned host tensors: x (B,H,W,C) raw units, optionally the
        per-level eps) on the engine's copy stream.  Batches are consumed in order by train_step_staged()."""
        st = getattr(eng, "_stage", None)
        if st is None:
            st = eng._stage = dict(stream=torch.cuda.Stream(self._device), queue=[], next=0,
                                   x=[torch.empty_like(eng.x) for _ in range(2)],
                                   eps=[[torch.empty_like(e) for e in eng.eps] for _ in range(2)],
                                   ready=[torch.cuda.Event() for _ in range(2)], used=[None, None])
        if len(st["queue"]) >= 2:
            raise RuntimeError("stage_batch: two batches are already staged; run train_step_staged() first")
        slot = st["next"]
        st["next"] ^= 1
        with torch.cuda.stream(st["stream"]):
            if st["used"][slot] is not None:
                st["stream"].wait_event(st["used"][slot])      # the step that read this slot has copied it out
            st["x"][slot].copy_(x_pinned, non_blocking=True)
            if eps_pinned is not None:
                for d, h in zip(st["eps"][slot], eps_pinned):
                    d.copy_(h, non_blocking=True)
            st["ready"][slot].record(st["stream"])
        st["queue"].append((slot, eps_pinned is not None))

    def train_step_staged(self, eng, corrupt=False):
        """One training step on the oldest staged batch (eps drawn on the device when none was staged); corrupt=True
        applies the training-phase input corruption after the batch has landed in the step's input buffer."""
        st = eng._stage
        slot, has_eps = st["queue"].pop(0)
        cur = torch.cuda.current_stream(self._device)
        cur.wait_event(st["ready"][slot])
        eng.x.copy_(st["x"][slot], non_blocking=True)
        if has_eps:
            for d, s_ in zip(eng.eps, st["eps"][slot]):
                d.copy_(s_, non_blocking=True)
        else:
            self._load_eps(eng, None)
        ev = torch.cuda.Event()
        ev.record(cur)
        st["used"][slot] = ev
        if corrupt:
            self._corrupt(eng)
        self.train_step_device(eng)

    def train_on_batch(self, x, eps=None, corrupt=False, noise=None, keep=None):
        """x: (B,H,W,C) numpy (host) or CUDA tensor in raw [min,max] units.  Returns dict of python floats.
        corrupt=True applies the reference's training-phase input corruption (noise ~ N(0,1) like x and keep (B,C) of 0/1
        may be supplied for determinism, else they are drawn on the device)."""
        B = x.shape[0]
        eng = self._engine(B, True, corrupt)
        self._load_input(eng, x)
        self._load_eps(eng, eps)
        if corrupt:
            self._corrupt(eng, noise, keep)
        self.train_step_device(eng)
        return self.read_losses(eng)

    def _corrupt(self, eng, noise=None, keep=None):
        """GaussianNoise + SpatialDropout2D of the "multiscale" model (multiscale_vae.py:139-147), device RNG."""
        if noise is None:
            if not hasattr(eng, "_noise"):
                eng._noise = torch.empty_like(eng.x)
                eng._keep = torch.empty((eng.B, eng.spec.C), dtype=torch.float32, device=self._device)
            noise = eng._noise.normal_(0.0, 1.0, generator=self._gen)
        else:
            noise = torch.as_tensor(noise, dtype=torch.float32).to(self._device).contiguous()
        if keep is None:
            if not hasattr(eng, "_keep"):
                eng._keep = torch.empty((eng.B, eng.spec.C), dtype=torch.float32, device=self._device)
            keep = eng._keep.uniform_(0.0, 1.0, generator=self._gen).ge_(self._training_dropout)
        else:
            keep = torch.as_tensor(keep, dtype=torch.float32).to(self._device).contiguous()
        eng._corrupt_args = (noise, keep)      # keep them alive until the kernel ran
        eng.corrupt_input(noise, keep, self._training_noise_std, self._training_dropout)

    def _load_input(self, eng, x):
        if torch.is_tensor(x):
            eng.x.copy_(x.to(torch.float32), non_blocking=True)
        else:
            eng.x.copy_(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)), non_blocking=True)

    def _load_eps(self, eng, eps):
        for i, e in enumerate(eng.eps):
            if eps is None:
                e.normal_(0.0, 1.0, generator=self._gen)
            else:
                e.copy_(torch.as_tensor(eps[i], dtype=torch.float32), non_blocking=True)

    @staticmethod
    def _loss_dict(s):
        return dict(loss=s[0] + s[4], vae_r_loss=s[2], vae_kl_loss=s[3], r_loss=s[1], reg_loss=s[4])

    @classmethod
    def read_losses(cls, eng):
        """The five loss scalars of the last step: ONE device-to-host read of eng.loss5 (they sit side by side)."""
        return cls._loss_dict(eng.loss5.tolist())

    def per_scale_elbo(self, x, eps=None, training=True):
        """Per-scale ELBO of a batch (north star (3); the loss form of multiscale_vae_.py:340-353 on this model's
        pyramid): for every scale i the reconstruction error mean_b sum_hw mean_c |x_i - m_i| in raw units (x_i: the
        image at scale i, m_i: the partial merge of the decoder outputs i..L-1) and the analytic KL of that scale's
        latents, plus elbo_i = r_loss_factor * recon_i + kl_loss_factor * kl_i.  training=True uses batch statistics
        in the BatchNorm layers like the training step (moving statistics are updated); returns numpy arrays (L,)
        and the per-sample (L, B) tensors."""
        x = np.asarray(x, dtype=np.float32) if not torch.is_tensor(x) else x
        eng = self._engine(x.shape[0], bool(training))
        self._load_input(eng, x)
        self._load_eps(eng, eps)
        if training:
            eng.forward_train(parallel=False)
        else:
            eng.encode(parallel=False)
            eng.decode(parallel=False)
        recon, kl = eng.per_scale_elbo()
        recon, kl = recon.cpu(), kl.cpu()
        elbo = recon * self._r_loss_factor + kl * self._kl_loss_factor
        return dict(recon=recon.mean(1).numpy(), kl=kl.mean(1).numpy(), elbo=elbo.mean(1).numpy(),
                    recon_per_sample=recon.numpy(), kl_per_sample=kl.numpy())

    # ==========================================================================================================
    def train(self, x_train, batch_size, epochs, run_folder, print_every_n_batches=100, initial_epoch=0, step_size=1,
              lr_decay=1, save_checkpoint_weights=False):
        """multiscale_vae.py:508-557 (`fit(x, x, batch_size, shuffle=True, epochs, initial_epoch, callbacks)`): shuffled
        mini-batch epochs in which EVERY sample is used once -- the last batch of an epoch is short when batch_size does
        not divide the data, as in Keras (tests/golden/train_fit_call.json) -- step-decay learning rate per epoch
        (schedule.py:7-21), the visualisation callback every `print_every_n_batches` batches (callbacks.py:66-134), and
        optional weight checkpoints named after the epoch and its mean loss (:524-537; .npz, h5py is not available).
        Returns the per-epoch history (sample-weighted means of the batch losses, like Keras reports them).

        Data-parallel (enable_data_parallel): `batch_size` is the per-process batch; every rank passes the same x_train
        and trains on its own shard perm[rank::world] of the epoch's permutation (truncated to equal length, so all ranks
        take the same number of steps); rank 0 alone writes images and checkpoints, after the BatchNorm moving statistics
        have been averaged over the ranks."""
        x_train = np.asarray(x_train, dtype=np.float32)
        n_all = x_train.shape[0]
        rank, world = (self._dist.rank, self._world) if self._dist is not None else (0, 1)
        n = n_all // world
        if n <= 0 or batch_size <= 0:
            raise ValueError(f"train: {n_all} samples cannot feed {world} rank(s) with batch_size {batch_size}")
        lr_fn = schedule.step_decay_schedule(initial_lr=self._learning_rate, decay_factor=lr_decay, step_size=step_size)
        weights_path = os.path.join(run_folder, "weights")
        os.makedirs(weights_path, exist_ok=True)
        viz = None
        if rank == 0:
            viz = callbacks.SaveIntermediateResultsCallback(run_folder, print_every_n_batches, initial_epoch,
                                                            x_train[0:16, :, :, :], self)
        steps, tail = divmod(n, batch_size)
        eng = self._engine(batch_size, True, corrupt=True) if steps else None
        eng_tail = self._engine(tail, True, corrupt=True) if tail else None
        stage = [torch.empty((batch_size,) + self._inputs_dims, dtype=torch.float32).pin_memory() for _ in range(2)] \
            if steps else []
        lib = _lib.load()
        sums = torch.zeros(5, dtype=torch.float32, device=self._device)
        history = []
        for epoch in range(initial_epoch, epochs):
            if viz is not None:
                viz.on_epoch_begin(epoch)
            self.learning_rate = float(lr_fn(epoch))
            # the same permutation on every rank (seeded by the epoch), each rank keeps its own slice of it
            perm = np.random.default_rng(1234 + epoch).permutation(n_all)[rank::world][:n]
            sums.zero_()
            t0, last = time.time(), None

            def fill(it):
                # gather batch `it` into a pinned buffer and start its H2D copy; it overlaps the step before it
                buf = stage[it & 1]
                if it >= 2:
                    eng._stage["ready"][it & 1].synchronize()      # the copy that last read this pinned buffer is done
                idx = np.sort(perm[it * batch_size:(it + 1) * batch_size])
                buf.copy_(torch.from_numpy(x_train[idx]))
                self.stage_batch(eng, buf)

            def after_batch(it, e, count):
                nonlocal last
                _lib.check(lib.mvae_accumulate(sums.data_ptr(), e.loss5.data_ptr(), 5, float(count),
                                               torch.cuda.current_stream(self._device).cuda_stream), "accumulate")
                if it % print_every_n_batches == 0 or it == steps + (1 if tail else 0) - 1:
                    last = self.read_losses(e)
                    logger.info("epoch %d batch %d/%d loss %.4f vae_r_loss %.4f vae_kl_loss %.4f", epoch + 1, it + 1,
                                steps + (1 if tail else 0), last["loss"], last["vae_r_loss"], last["vae_kl_loss"])
                if viz is not None:
                    viz.on_batch_end(it)
                if self._dist is not None and it % print_every_n_batches == 0:
                    # rank 0 has just written the collages (host work of a few hundred milliseconds): the other ranks wait
                    # here, on the host, rather than inside the next step's exchange kernel
                    torch.cuda.synchronize(self._device)
                    torch.distributed.barrier()

            if steps:
                if hasattr(eng, "_stage"):
                    eng._stage["queue"].clear()
                    eng._stage["next"] = 0
                fill(0)
                for it in range(steps):
                    if it + 1 < steps:
                        fill(it + 1)
                    self.train_step_staged(eng, corrupt=True)
                    after_batch(it, eng, batch_size)
            if tail:
                # the short last batch of the epoch (Keras fit trains on it): its own engine, sized for `tail` samples
                idx = np.sort(perm[steps * batch_size:])
                self._load_input(eng_tail, x_train[idx])
                self._load_eps(eng_tail, None)
                self._corrupt(eng_tail)
                self.train_step_device(eng_tail)
                after_batch(steps, eng_tail, tail)
            mean = self._loss_dict((sums / float(n)).tolist())
            dt = time.time() - t0
            entry = dict(mean, epoch=epoch + 1, images_per_sec=n * world / max(dt, 1e-9), lr=self._learning_rate,
                         batches=steps + (1 if tail else 0), last_batch=last)
            history.append(entry)
            logger.info("epoch %d done: loss %.4f (%.0f images/s)", epoch + 1, mean["loss"], entry["images_per_sec"])
            if save_checkpoint_weights:
                if self._dist is not None:
                    self._dist.average_moving_stats()
                if rank == 0:
                    self.save_weights(os.path.join(weights_path, "weights-%03d-%.2f.npz" % (epoch + 1, mean["loss"])))
                    self.save_weights(os.path.join(weights_path, "weights.npz"))
            if self._dist is not None:
                # rank-local host work (checkpoint, visualisation, logging) is over on every rank before the next epoch's
                # first step: the gradient exchange is a kernel that waits for its peers on the device
                torch.cuda.synchronize(self._device)
                torch.distributed.barrier()
        return history

    # ==========================================================================================================
    def _predict_encoder(self, x, eps=None):
        """`_model_encoder` (multiscale_vae.py:228-243): sampled z of every level, concatenated (B, sum z)."""
        eng = self._engine(x.shape[0], False)
        self._load_input(eng, x)
        self._load_eps(eng, eps)
        eng.encode(parallel=False)
        return torch.cat([t.data.view(x.shape[0], -1) for t in eng.zT], dim=1).cpu().numpy()

    def _predict_decoder(self, z, eps=None):
        """`_model_decoder` (multiscale_vae.py:247-257): (B, sum z) -> denormalised, clipped image."""
        B = z.shape[0]
        if z.shape[1] != sum(self._z_latent_dims):
            raise ValueError(f"expected latent width {sum(self._z_latent_dims)}, got {z.shape[1]}")
        eng = self._engine(B, False)
        zt = torch.from_numpy(np.ascontiguousarray(z, dtype=np.float32)).to(self._device)
        o = 0
        for i, zd in enumerate(self._z_latent_dims):          # split Lambda, multiscale_vae.py:96-106
            eng.zT[i].data.view(B, zd).copy_(zt[:, o:o + zd])
            o += zd
        eng.decode(parallel=False)
        return eng.out.cpu().numpy()

    def _predict_trainable(self, x, eps=None):
        """`_model_trainable` in inference mode: encode -> decode."""
        eng = self._engine(x.shape[0], False)
        self._load_input(eng, x)
        self._load_eps(eng, eps)
        eng.encode(parallel=False)
        eng.decode(parallel=False)
        return eng.out.cpu().numpy()

    # ---- aliases asked for by mvae/vae.py:12-81 and used by mvae/callbacks.py:77-128 ------------------------------
    def encode(self, x, eps=None):
        return self._predict_encoder(np.asarray(x, dtype=np.float32), eps)

    def sample(self, z):
        return self._predict_decoder(np.asarray(z, dtype=np.float32))

    def predict(self, x, eps=None):
        return self._predict_trainable(np.asarray(x, dtype=np.float32), eps)

    @property
    def z_dim(self):
        return int(sum(self._z_latent_dims))

    @property
    def input_dim(self):
        return self._inputs_dims

    @property
    def model_encode(self):
        return self._model_encoder

    @property
    def model_decode(self):
        return self._model_decoder

    # ==========================================================================================================
    def save_weights(self, filename):
        np.savez(filename, **{k: v.numpy() for k, v in self._ps.state_dict().items()})

    def load_weights(self, filename):
        """The reference's load_weights is a stub (multiscale_vae.py:561-562); here a Keras-named .npz is loaded."""
        if filename is None:
            return
        with np.load(filename) as f:
            self._ps.load_state_dict({k: f[k] for k in f.files})

    def state_dict(self):
        return self._ps.state_dict()

    def load_state_dict(self, sd):
        self._ps.load_state_dict(sd)

    # ==========================================================================================================
    @property
    def encoder(self):
        return self._model_encoder

    @property
    def decoder(self):
        return self._model_decoder

    @property
    def model_trainable(self):
        return self._model_trainable

    @property
    def learning_rate(self):
        return self._learning_rate

    @learning_rate.setter
    def learning_rate(self, value):
        self._learning_rate = value
        if value is not None:
            self._lr_dev.fill_(float(value))

    def normalize(self, v):
        return (v - self._min_value) / (self._max_value - self._min_value)
