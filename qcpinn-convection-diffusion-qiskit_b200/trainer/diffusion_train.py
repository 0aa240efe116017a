"""Convection-diffusion PINN training loop (drop-in for reference trainer/diffusion_train.py).

``train(model, nIter=10000, batch_size=128, log_NTK=False, update_lam=False)`` keeps the reference
contract: ``batch_size`` residual points plus ``batch_size // 3`` initial-condition and
``batch_size // 3`` x=0 boundary points per step, loss ``2 MSE_r + 4 MSE_bc + 2 MSE_ic``,
``clip_grad_norm_(1.0)`` (0.1 for the CV solver), Adam, ``ReduceLROnPlateau.step(loss)`` and the
loss appended to ``model.loss_history`` every step (in the captured step both happen on the device
and the host objects are brought up to date when they are read: same decisions and values, no
``loss.item()`` round trip per step, :class:`DevicePlateau`); ``model.epochs + 1`` iterations; a log
line and a checkpoint every ``args["print_every"]`` steps (logged / saved right AFTER that step's
parameter update; the reference does it just before).  ``nIter``, ``log_NTK`` and ``update_lam`` are
accepted and unused, as in the reference.

The step itself is exposed as :class:`TrainStep` so ``bench.py`` times exactly what ``train`` runs.
Under ``torch.distributed`` (world size > 1) every rank draws its own points and the gradients and
the loss ride in one exchange before clipping: one peer-memory kernel per rank on an NVLink node
(``dist.PeerAllReduce``), one flat NCCL / gloo all-reduce otherwise.
"""

import os
import time

import torch

from ..data.diffusion_dataset import Sampler, r, training_boxes, u
from ..nn.pde import diffusion_operator


# diffusion_operator's defaults (sigma = 1, D = 0.01, v = (1, 1)) as residual coefficients
DIFFUSION_COEFFS = (1.0, 1.0, 1.0, -0.01, -0.01)


def fetch_minibatch(sampler, N, out=None, rnd=None):
    if out is None and rnd is None:
        return sampler.sample(N)
    return sampler.sample(N, out=out, rnd=rnd)


class DevicePlateau:
    """Device-resident twin of ``model.scheduler`` (``ReduceLROnPlateau``) for the CUDA-graph step.

    The reference loop ends every step with ``scheduler.step(loss)`` and ``loss.item()``
    (trainer/diffusion_train.py:86-90): a host round trip that a replayed graph would have to wait
    for.  Here one single-thread kernel at the end of the captured step (``qcp_plateau_step``)
    repeats the scheduler's arithmetic on six device doubles, multiplies the device-resident
    learning rate when patience runs out and appends the loss to a device ring; the host objects
    (``scheduler.best / num_bad_epochs / cooldown_counter / last_epoch``, ``model.loss_history``)
    are brought up to date in batches -- whenever somebody reads them (``model.loss_history``,
    ``scheduler.state_dict()``, ``scheduler.step()``), or when the ring is full.  Same decisions,
    same numbers, no synchronisation per step.  One instance per model, shared by its TrainSteps.
    """

    CAPACITY = 4096

    def __init__(self, model):
        from .. import functional as F

        self.model, self.sched = model, model.scheduler
        self.lr = model.optimizer.param_groups[0]["lr"]
        dev = self.lr.device
        self.state = torch.zeros(F.PLATEAU_STATE, dtype=torch.float64, device=dev)
        self.hist = torch.zeros(self.CAPACITY, dtype=torch.float32, device=dev)
        self.pending = 0              # replays recorded on the device, not yet seen by the host
        self.mirror = None            # host-side state the device state was last aligned with
        model._lazy_flushers.append(self.flush)
        self.sched._qcp_flushers = tuple(self.sched._qcp_flushers) + (self.flush,)

    @staticmethod
    def of(model):
        """The model's twin (created on first use), or None when the host scheduler has to stay in
        charge: no / foreign scheduler, several parameter groups, learning rate not on the device."""
        from ..nn.DVPDESolver import _RankConsistentPlateau

        sched, opt = getattr(model, "scheduler", None), getattr(model, "optimizer", None)
        ok = (isinstance(sched, _RankConsistentPlateau) and opt is not None
              and len(opt.param_groups) == 1 and len(sched.min_lrs) == 1
              and torch.is_tensor(opt.param_groups[0]["lr"]) and opt.param_groups[0]["lr"].is_cuda
              and opt.param_groups[0]["lr"].dtype == torch.float32)
        twin = model.__dict__.get("_device_plateau")
        if ok and twin is not None and twin.sched is sched and twin.lr is opt.param_groups[0]["lr"]:
            return twin
        if twin is not None:              # scheduler / optimizer were replaced: retire the old twin
            twin.retire()
            model.__dict__.pop("_device_plateau", None)
        if not ok:
            return None
        twin = DevicePlateau(model)
        model.__dict__["_device_plateau"] = twin
        return twin

    def retire(self):
        self.flush()
        if self.flush in self.model._lazy_flushers:
            self.model._lazy_flushers.remove(self.flush)
        self.sched._qcp_flushers = tuple(f for f in self.sched._qcp_flushers if f != self.flush)

    def config(self):
        s = self.sched
        return {"mode_max": s.mode == "max", "threshold_abs": s.threshold_mode == "abs",
                "threshold": s.threshold, "factor": s.factor, "patience": s.patience,
                "cooldown": s.cooldown, "min_lr": s.min_lrs[0], "eps": s.eps}

    def _host_state(self):
        s = self.sched
        return (float(s.best), int(s.num_bad_epochs), int(s.cooldown_counter), int(s.last_epoch))

    def align(self):
        """Before a replay: if the host scheduler moved since the last exchange (eager steps, a
        restored checkpoint), drain the device side and upload the host state."""
        if self.mirror != self._host_state():
            self.flush()
            host = self._host_state()
            self.state.copy_(torch.tensor(host + (0.0, 0.0), dtype=torch.float64), non_blocking=False)
            self.mirror = host

    def record(self, plan, metric):
        """Inside the captured step, after the optimizer: one scheduler step on the device."""
        from .. import functional as F

        F.plateau_step(plan, metric, self.state, self.lr, self.hist, self.config())

    def replayed(self):
        self.pending += 1
        if self.pending >= self.CAPACITY:
            self.flush()

    def flush(self):
        if self.pending == 0:
            return
        n, self.pending = self.pending, 0
        state = self.state.cpu()                              # synchronises with the last replay
        values = self.hist[:n].cpu().tolist()
        if int(state[4]) != n:
            raise RuntimeError(f"DevicePlateau: {int(state[4])} steps recorded on the device, "
                               f"{n} replays counted on the host")
        self.state[4:5].zero_()
        self.model.__dict__["_loss_history"].extend(values)
        s = self.sched
        s.best = float(state[0])
        s.num_bad_epochs, s.cooldown_counter, s.last_epoch = (int(state[i]) for i in (1, 2, 3))
        s._last_lr = [g["lr"].clone() if torch.is_tensor(g["lr"]) else g["lr"]
                      for g in s.optimizer.param_groups]
        self.mirror = self._host_state()


class TrainStep:
    """One optimisation step of the reference loop.

    Eager mode runs the phases exactly in the reference's order.  On a CUDA device (unless
    ``args["cuda_graph"]`` is False) the step -- sampling, three model calls, loss, backward,
    gradient all-reduce, clipping, Adam -- is captured ONCE as a CUDA graph after a few eager
    steps and then replayed: the ~190 small launches of a step then cost microseconds of host time
    instead of ~3 ms, which is what strong scaling to 8 GPUs (0.5 M points per rank) needs.  On
    the fused route the plateau scheduler runs inside the graph too (:class:`DevicePlateau`), and
    the points of step k + 1 are drawn at the end of step k, next to the single-CTA tail (gradient
    reduction, clip, Adam), into the static batch the next replay reads.

    ``host_sync`` (attribute, may be flipped between calls): True = every call returns the loss as a
    Python float (one device->host read per step, like the reference's ``loss.item()``); False =
    the call returns the device scalar and never waits for the GPU, ``model.loss_history`` and the
    scheduler object catch up when they are read.
    """

    EAGER_STEPS_BEFORE_CAPTURE = 3

    def __init__(self, model, batch_size=128, averager=None, use_graph=None, host_sync=True):
        self.model = model
        self.batch_size = batch_size
        boxes = training_boxes(model.device)
        self.ics_sampler = Sampler(3, boxes["ics"], u, name="Initial Condition", device=model.device)
        self.bcs_sampler = [
            Sampler(3, boxes["bc1"], u, name="Dirichlet BC1", device=model.device),
            Sampler(3, boxes["bc2"], u, name="Dirichlet BC2", device=model.device),
        ]
        self.res_sampler = Sampler(3, boxes["dom"], r, name="Forcing", device=model.device)
        self.averager = averager
        self.max_norm = 0.1 if model.args["solver"] == "CV" else 1
        on_cuda = model.device is not None and torch.device(model.device).type == "cuda"
        if use_graph is None:
            use_graph = bool(model.args.get("cuda_graph", True))
        self.use_graph = bool(use_graph) and on_cuda
        self.fuse_calls = bool(model.args.get("fuse_model_calls", True))
        # whole-step fast path: forward -> MSE seeds -> adjoint kernels -> flat gradient, without an
        # autograd graph (same numbers; args["fuse_step"] = False keeps the autograd route)
        self.fuse_step = (bool(model.args.get("fuse_step", True)) and on_cuda
                          and getattr(model, "supports_fused_step", lambda: False)()
                          and model.optimizer is not None)
        self._eager_calls = 0
        self._graphs = {}          # "device" / "host" -> (graph, outputs, static batch, launches, ahead)
        self.last_terms = None     # (loss, loss_r, loss_bc, loss_ic) tensors of the last step
        self._stages = []          # two device staging slots for prefetched host batches
        self._copy_stream = None
        self.host_sync = bool(host_sync)
        self._plateau = None       # DevicePlateau the captured graphs were built with
        self._plateau_config = None
        self._capture_stream = None
        self._peer = None          # PeerAllReduce | False (unavailable) | None (not tried yet)

    def draw(self):
        """The random numbers of one :meth:`sample` call, in its order (pass back as ``rnd``)."""
        n = self.batch_size
        return (self.ics_sampler.draw(n // 3), self.bcs_sampler[0].draw(n // 3), self.res_sampler.draw(n))

    def sample(self, out=None, rnd=None):
        """One batch (X_ics, u_ics, X_bcs, u_bcs, X_res, r_res); ``out`` = six tensors of a previous
        call to overwrite (the fused CUDA samplers write them in place), ``rnd`` = :meth:`draw`."""
        n = self.batch_size
        o = (None, None, None) if out is None else (out[0:2], out[2:4], out[4:6])
        q = (None, None, None) if rnd is None else rnd
        X_ics, u_ics = fetch_minibatch(self.ics_sampler, n // 3, o[0], q[0])
        X_bcs, u_bcs = fetch_minibatch(self.bcs_sampler[0], n // 3, o[1], q[1])   # only BC face x=0 is used
        X_res, r_res = fetch_minibatch(self.res_sampler, n, o[2], q[2])
        return X_ics, u_ics, X_bcs, u_bcs, X_res, r_res

    def flush(self):
        """Bring ``model.loss_history`` and the scheduler object up to date (no-op when current);
        raises if a rank ever missed a peer-memory gradient exchange."""
        if self._plateau is not None:
            self._plateau.flush()
        if self._peer:
            self._peer.check()

    def close(self):
        """Drain pending host state and release the captured graphs and the peer-memory buffer
        (do this before ``torch.distributed.destroy_process_group()``: graphs of the NCCL route hold
        kernels of the communicator)."""
        self.flush()
        self._graphs.clear()
        if self._peer:
            self._peer.close()
        self._peer = None

    def _peer_exchange(self, n_values):
        """The node-local peer-memory exchange (created collectively on first use), or None when the
        NCCL all-reduce has to do it."""
        if self._peer is None:
            from ..dist import PeerAllReduce

            made = PeerAllReduce.create(n_values, self.model.quantum_layer.params.device,
                                        self.averager.group, self.max_norm)
            self._peer = made if made is not None else False
        return self._peer or None

    def objective(self, batch=None):
        """Returns (loss, seconds, loss_r, loss_bc, loss_ic) like the reference's objective_fn."""
        model = self.model
        start = time.time()
        if model.optimizer is not None:
            model.optimizer.zero_grad()
        X_ics, u_ics, X_bcs, u_bcs, X_res, r_res = self.sample() if batch is None else batch
        X_ics.requires_grad_(True)
        many = getattr(model, "forward_many", None) if self.fuse_calls else None
        if many is not None:
            # same three model calls, issued behind one autograd node (shared gradient reduction)
            u_bc1_pred, u_ics_pred, (_, r_pred) = many(
                [(X_bcs, None), (X_ics, None), (X_res, DIFFUSION_COEFFS)])
        else:
            u_bc1_pred = model.forward(X_bcs)
            u_ics_pred = model.forward(X_ics)
            t_r, x_r, y_r = X_res[:, 0:1], X_res[:, 1:2], X_res[:, 2:3]
            _, r_pred = diffusion_operator(model, t_r, x_r, y_r)
        loss_r = model.loss_fn(r_pred, r_res)
        loss_bc1 = model.loss_fn(u_bc1_pred, u_bcs)
        loss_ics = model.loss_fn(u_ics_pred, u_ics)
        loss = 2.0 * loss_r + 4.0 * loss_bc1 + 2.0 * loss_ics
        return loss, time.time() - start, loss_r, loss_bc1, loss_ics

    def _device_update(self, loss, run_backward=True):
        """backward -> (all-reduce) -> clip -> Adam; returns the (rank-averaged) loss tensor."""
        model = self.model
        if run_backward:
            loss.backward()
        if self.averager is not None:
            # grads and the scheduler metric share one all-reduce
            loss = self.averager.average(extras=[loss])[0].clone()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=self.max_norm)
        if model.optimizer is not None:
            model.optimizer.step()
        return loss.detach()

    def _fused_device_step(self, batch=None, ahead=None):
        """sample -> train_step_grads -> (all-reduce) -> average + clip -> Adam, on the flat
        gradient buffer.  Returns (loss, loss_r, loss_bc, loss_ic) device scalars (rank-averaged
        loss, local terms).  ``ahead`` (graph capture with device-drawn points) = the static batch
        this step reads; the NEXT step's points are drawn into it once the gradients are packed."""
        from .. import functional as F

        model = self.model
        if ahead is not None:
            batch = ahead
        elif batch is None:
            # the samplers (six small launches) run on the side stream while the main stream casts
            # the weights and runs the single-CTA qcp_prepare of this step's angles
            dev = model.quantum_layer.params.device
            main = torch.cuda.current_stream(dev)
            side = model._value_stream(dev)
            if side != main:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    batch = self.sample()
                model.prepare_step()
                main.wait_stream(side)
                for t in batch:
                    t.record_stream(main)
            else:
                batch = self.sample()
        side, refill = None, None
        if ahead is not None:
            # the NEXT step's random numbers are drawn first thing on the side stream (no
            # dependencies); they are mapped into the static batch as soon as the last adjoint kernel
            # has read it, next to this stream's single-CTA tail (partial reduction, all-reduce,
            # clip, Adam, scheduler)
            dev = model.quantum_layer.params.device
            main = torch.cuda.current_stream(dev)
            side = model._value_stream(dev)
            if side != main:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    rnd = self.draw()

                def refill():
                    side.wait_stream(main)            # main has joined every adjoint kernel here
                    with torch.cuda.stream(side):
                        self.sample(out=ahead, rnd=rnd)
            else:
                side = None
                rnd = self.draw()

                def refill():
                    self.sample(out=ahead, rnd=rnd)
        flat, numel = model.train_step_grads(batch, DIFFUSION_COEFFS, after_adjoints=refill)
        world = 1
        if self.averager is not None:
            import torch.distributed as dist

            # gradients and the scheduler metric share one exchange; the local terms are kept
            local_terms = flat[numel + 1:numel + 4].clone()
            peer = self._peer_exchange(numel + 1)
            self.averager.calls += 1
            if peer is not None:
                # one kernel per rank over NVLink peer memory: exchange, ordered sum, mean, clip
                peer.reduce_clip(flat, numel, 1, self.max_norm)
                F._count(1)
                world = 0
            else:
                dist.all_reduce(flat[:numel + 1], op=dist.ReduceOp.SUM, group=self.averager.group)
                world = self.averager.world
        else:
            local_terms = flat[numel + 1:numel + 4]
        if world:
            plan = model._plan(model.quantum_layer.params.device)
            F.clip_grads(plan, flat, numel, 1, 1.0 / world, self.max_norm)
        model.optimizer.step()
        if side is not None:
            torch.cuda.current_stream(model.quantum_layer.params.device).wait_stream(side)
        return flat[numel].clone(), local_terms[0], local_terms[1], local_terms[2]

    # -- host-fed batches ----------------------------------------------------------------------------
    def prefetch(self, batch):
        """Start the host->device copy of a FUTURE step's batch on a side stream, so that it overlaps
        the step running now (data-loader style double buffering; two staging slots).  ``batch`` =
        the six pinned host tensors of :meth:`sample`; pass the same object to ``step(batch)``
        later."""
        dev = self.model.device
        if not self._stages or any(d.shape != s.shape for d, s in zip(self._stages[0]["buf"], batch)):
            self._stages = [{"buf": tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in batch),
                             "src": None, "ready": None, "free": None} for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=dev)
        slot = next((s for s in self._stages if s["src"] is None), None)
        if slot is None:
            raise RuntimeError("TrainStep.prefetch: both staging slots hold batches that were never used")
        if slot["free"] is not None:                     # its previous consumer is done reading
            self._copy_stream.wait_event(slot["free"])
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(slot["buf"], batch):
                dst.copy_(src, non_blocking=True)
            slot["ready"] = torch.cuda.Event()
            slot["ready"].record(self._copy_stream)
        slot["src"] = batch

    def _take_prefetched(self, batch):
        """The staging slot holding ``batch`` if it was prefetched (made stream-ordered), else None."""
        if batch is None:
            return None
        for slot in self._stages:
            if slot["src"] is batch:
                torch.cuda.current_stream(self.model.device).wait_event(slot["ready"])
                return slot
        return None

    def _release_stage(self, slot):
        slot["src"] = None
        slot["free"] = torch.cuda.Event()
        slot["free"].record(torch.cuda.current_stream(self.model.device))

    def _host_update(self, loss):
        """plateau scheduler -> loss.item() -> history (needs the host value)."""
        model = self.model
        value = loss.item()
        if model.scheduler is not None:
            model.scheduler.step(value)
        model.loss_history.append(value)
        return value

    def update(self, loss, run_backward=True):
        """Eager tail of a step: backward -> (all-reduce) -> clip -> Adam -> scheduler -> item()."""
        return self._host_update(self._device_update(loss, run_backward))

    # -- CUDA-graph path -----------------------------------------------------------------------
    def _capture(self, kind, batch):
        model = self.model
        static_batch = None
        if kind == "host":
            static_batch = tuple(torch.empty(t.shape, dtype=t.dtype, device=model.device)
                                 for t in batch)
        from .. import functional as F

        plan = model._plan(model.quantum_layer.params.device)
        plateau = self._plateau if self.fuse_step else None
        ahead = None
        if kind == "device" and self.fuse_step:
            ahead = tuple(t.detach() for t in self.sample())   # the first replay's points
        plan.invalidate()          # the captured step must contain its own qcp_prepare launch
        graph = torch.cuda.CUDAGraph()
        mode = "thread_local" if self.averager is not None else "global"
        launches0 = F.launch_counter
        # captured on a high-priority stream: kernel nodes inherit it, so whenever the residual chain
        # (this stream) and the IC/BC chain (the model's default-priority side stream) both have
        # CTAs waiting, the block scheduler places the residual chain's first and the IC/BC kernels
        # fill what is left (measured +1.1 % at 4 194 304 points, +3.7 % at 524 288; QCP_GRAPH_PRIORITY=0
        # captures on a default-priority stream)
        if self._capture_stream is None:
            high = os.environ.get("QCP_GRAPH_PRIORITY", "1") != "0"
            self._capture_stream = torch.cuda.Stream(device=model.device, priority=-1 if high else 0)
        with torch.cuda.graph(graph, stream=self._capture_stream, capture_error_mode=mode):
            if self.fuse_step:
                reduced, loss_r, loss_bc, loss_ic = self._fused_device_step(static_batch, ahead)
                if plateau is not None:
                    plateau.record(plan, reduced)
            else:
                loss, _, loss_r, loss_bc, loss_ic = self.objective(static_batch)
                reduced = self._device_update(loss)
        plan.invalidate()
        outs = (reduced, loss_r.detach(), loss_bc.detach(), loss_ic.detach())
        # kernels of this library inside one replay (capture only recorded them)
        launches = F.launch_counter - launches0
        F.launch_counter = launches0
        self._graphs[kind] = (graph, outs, static_batch, launches, ahead)

    def _graph_step(self, batch):
        kind = "device" if batch is None else "host"
        plateau = DevicePlateau.of(self.model) if self.fuse_step else None
        if plateau is not self._plateau or \
                (plateau is not None and plateau.config() != self._plateau_config):
            self._graphs.clear()       # scheduler replaced or reconfigured: its constants are baked in
            self._plateau = plateau
            self._plateau_config = plateau.config() if plateau is not None else None
        if kind not in self._graphs:
            self._capture(kind, batch)
        from .. import functional as F

        if plateau is not None:
            plateau.align()
        graph, outs, static_batch, launches, _ = self._graphs[kind]
        if static_batch is not None:
            staged = self._take_prefetched(batch)
            with torch.no_grad():                       # X_ics is a leaf that requires grad
                for dst, src in zip(static_batch, staged["buf"] if staged is not None else batch):
                    dst.copy_(src, non_blocking=True)   # device (staged) or pinned-host sources
            if staged is not None:
                self._release_stage(staged)
        graph.replay()
        # the replayed fused Adam changed every parameter behind Python's back (no version bump,
        # no optimizer hook): move the cache keys on, so that an eager forward / evaluation between
        # replays re-casts the weights and re-runs qcp_prepare instead of reusing stale copies
        self.model.quantum_layer.mark_updated()
        F.launch_counter += launches
        self.last_terms = outs
        if plateau is None:
            return self._host_update(outs[0])
        plateau.replayed()         # scheduler step and loss record happened inside the graph
        return outs[0].item() if self.host_sync else outs[0]

    def steady(self, host_batches=False):
        """True once a call does only steady-state work: no eager start-up steps and no CUDA-graph
        capture left for this kind of batch (device-sampled or host-fed)."""
        return not self.use_graph or ("host" if host_batches else "device") in self._graphs

    def __call__(self, batch=None):
        if self.use_graph and self._eager_calls >= self.EAGER_STEPS_BEFORE_CAPTURE:
            return self._graph_step(batch)
        self._eager_calls += 1
        if batch is not None:
            staged = self._take_prefetched(batch)
            if staged is not None:
                batch = tuple(t.clone() for t in staged["buf"])
                self._release_stage(staged)
            else:
                batch = tuple(t.to(self.model.device, non_blocking=True) for t in batch)
        if self.fuse_step:
            reduced, loss_r, loss_bc, loss_ic = self._fused_device_step(batch)
        else:
            loss, _, loss_r, loss_bc, loss_ic = self.objective(batch)
            reduced = self._device_update(loss)
        self.last_terms = (reduced, loss_r.detach(), loss_bc.detach(), loss_ic.detach())
        return self._host_update(reduced)


def _make_averager(model):
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return None
    from ..dist import GradientAverager

    avg = GradientAverager(model, extra=1)
    avg.enabled = False            # explicit call in TrainStep.update, no engine callback needed
    sched = model.scheduler
    if hasattr(sched, "_qcp_enabled"):
        sched._qcp_enabled = False  # the metric is already averaged
    return avg


def train(model, nIter=10000, batch_size=128, log_NTK=False, update_lam=False):
    step = TrainStep(model, batch_size, _make_averager(model), host_sync=False)
    rank0 = True
    if step.averager is not None:
        import torch.distributed as dist

        rank0 = dist.get_rank() == 0
    t0 = time.time()
    model.logger.print(f"Starting training for {model.epochs} epochs...")
    model.logger.print(f"Batch size: {batch_size}")

    every = model.args["print_every"]
    t_log, it_log = t0, -1
    for it in range(model.epochs + 1):
        step()                                   # one full optimisation step (eager or graph replay)
        loss, loss_r, loss_bc1, loss_ics = step.last_terms
        if it % every == 0 or it == 0 or model.args.get("use_ibm_hardware", False):
            # graph replays do not wait for the GPU: the .item() reads below do, so step times are
            # averages over the steps since the previous log line
            loss_value = loss.item()
            now = time.time()
            elapsed = now - t0
            dt = (now - t_log) / (it - it_log)
            t_log, it_log = now, it
            eta = elapsed / (it + 1) * (model.epochs - it)
            lr = float(model.optimizer.param_groups[0]["lr"]) if model.optimizer else 0.0
            model.logger.print(
                "Epoch: %d/%d [%.1f%%] | Loss: %.2e | Loss_res: %.2e | Loss_bcs: %.2e | "
                "loss_ics: %.2e | lr: %.2e | Epoch_time: %.2fs | Total: %.1fs | ETA: %.1fs"
                % (it, model.epochs, 100.0 * it / model.epochs if model.epochs > 0 else 0,
                   loss_value, loss_r.item(), loss_bc1.item(), loss_ics.item(), lr, dt, elapsed, eta))
            if it > 0 and it % every == 0 and rank0:
                model.save_state()

    step.close()                                 # loss history / scheduler object complete (and synced)
    if model.device is not None and torch.device(model.device).type == "cuda":
        torch.cuda.synchronize(model.device)
    total = time.time() - t0
    model.logger.print(
        f"Training completed in {total:.2f} seconds ({total / 60:.2f} minutes)")
