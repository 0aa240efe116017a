"""Convection-diffusion PINN training loop (drop-in for reference trainer/diffusion_train.py).

``train(model, nIter=10000, batch_size=128, log_NTK=False, update_lam=False)`` keeps the reference
contract: ``batch_size`` residual points plus ``batch_size // 3`` initial-condition and
``batch_size // 3`` x=0 boundary points per step, loss ``2 MSE_r + 4 MSE_bc + 2 MSE_ic``,
``clip_grad_norm_(1.0)`` (0.1 for the CV solver), Adam, ``ReduceLROnPlateau.step(loss)`` and a
``loss.item()`` per step; ``model.epochs + 1`` iterations; a log line and a checkpoint every
``args["print_every"]`` steps (logged / saved right AFTER that step's parameter update; the
reference does it just before).  ``nIter``, ``log_NTK`` and ``update_lam`` are accepted and unused,
as in the reference.

The step itself is exposed as :class:`TrainStep` so ``bench.py`` times exactly what ``train`` runs.
Under ``torch.distributed`` (world size > 1) every rank draws its own points and the gradients and
loss terms ride in one flat all-reduce before clipping.
"""

import time

import torch

from ..data.diffusion_dataset import Sampler, r, training_boxes, u
from ..nn.pde import diffusion_operator


# diffusion_operator's defaults (sigma = 1, D = 0.01, v = (1, 1)) as residual coefficients
DIFFUSION_COEFFS = (1.0, 1.0, 1.0, -0.01, -0.01)


def fetch_minibatch(sampler, N):
    return sampler.sample(N)


class TrainStep:
    """One optimisation step of the reference loop.

    Eager mode runs the phases exactly in the reference's order.  On a CUDA device (unless
    ``args["cuda_graph"]`` is False) the step -- sampling, three model calls, loss, backward,
    gradient all-reduce, clipping, Adam -- is captured ONCE as a CUDA graph after a few eager
    steps and then replayed: the ~190 small launches of a step then cost microseconds of host time
    instead of ~3 ms, which is what strong scaling to 8 GPUs (0.5 M points per rank) needs.  The
    plateau scheduler and ``loss.item()`` stay outside the graph (they need the host value).
    """

    EAGER_STEPS_BEFORE_CAPTURE = 3

    def __init__(self, model, batch_size=128, averager=None, use_graph=None):
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
        self._graphs = {}          # "device" / "host" -> (graph, static outputs, static batch)
        self.last_terms = None     # (loss, loss_r, loss_bc, loss_ic) tensors of the last step
        self._stages = []          # two device staging slots for prefetched host batches
        self._copy_stream = None

    def sample(self):
        n = self.batch_size
        X_ics, u_ics = fetch_minibatch(self.ics_sampler, n // 3)
        X_bcs, u_bcs = fetch_minibatch(self.bcs_sampler[0], n // 3)   # only BC face x=0 is used
        X_res, r_res = fetch_minibatch(self.res_sampler, n)
        return X_ics, u_ics, X_bcs, u_bcs, X_res, r_res

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

    def _fused_device_step(self, batch=None):
        """sample -> train_step_grads -> (all-reduce) -> average + clip -> Adam, on the flat
        gradient buffer.  Returns (loss, loss_r, loss_bc, loss_ic) device scalars (rank-averaged
        loss, local terms)."""
        from .. import functional as F

        model = self.model
        if batch is None:
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
        flat, numel = model.train_step_grads(batch, DIFFUSION_COEFFS)
        world = 1
        if self.averager is not None:
            import torch.distributed as dist

            # gradients and the scheduler metric share one all-reduce; the local terms are kept
            local_terms = flat[numel + 1:numel + 4].clone()
            dist.all_reduce(flat[:numel + 1], op=dist.ReduceOp.SUM, group=self.averager.group)
            world = self.averager.world
            self.averager.calls += 1
        else:
            local_terms = flat[numel + 1:numel + 4]
        plan = model._plan(model.quantum_layer.params.device)
        F.clip_grads(plan, flat, numel, 1, 1.0 / world, self.max_norm)
        model.optimizer.step()
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
        plan.invalidate()          # the captured step must contain its own qcp_prepare launch
        graph = torch.cuda.CUDAGraph()
        mode = "thread_local" if self.averager is not None else "global"
        launches0 = F.launch_counter
        with torch.cuda.graph(graph, capture_error_mode=mode):
            if self.fuse_step:
                reduced, loss_r, loss_bc, loss_ic = self._fused_device_step(static_batch)
            else:
                loss, _, loss_r, loss_bc, loss_ic = self.objective(static_batch)
                reduced = self._device_update(loss)
        plan.invalidate()
        outs = (reduced, loss_r.detach(), loss_bc.detach(), loss_ic.detach())
        # kernels of this library inside one replay (capture only recorded them)
        launches = F.launch_counter - launches0
        F.launch_counter = launches0
        self._graphs[kind] = (graph, outs, static_batch, launches)

    def _graph_step(self, batch):
        kind = "device" if batch is None else "host"
        if kind not in self._graphs:
            self._capture(kind, batch)
        from .. import functional as F

        graph, outs, static_batch, launches = self._graphs[kind]
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
        return self._host_update(outs[0])

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
    step = TrainStep(model, batch_size, _make_averager(model))
    rank0 = True
    if step.averager is not None:
        import torch.distributed as dist

        rank0 = dist.get_rank() == 0
    t0 = time.time()
    model.logger.print(f"Starting training for {model.epochs} epochs...")
    model.logger.print(f"Batch size: {batch_size}")

    every = model.args["print_every"]
    step_times = []
    for it in range(model.epochs + 1):
        t_step = time.time()
        step()                                   # one full optimisation step (eager or graph replay)
        loss, loss_r, loss_bc1, loss_ics = step.last_terms
        dt = time.time() - t_step
        step_times.append(dt)
        if it % every == 0 or it == 0 or model.args.get("use_ibm_hardware", False):
            elapsed = time.time() - t0
            mean_dt = sum(step_times) / len(step_times)
            eta = mean_dt * (model.epochs - it)
            lr = float(model.optimizer.param_groups[0]["lr"]) if model.optimizer else 0.0
            model.logger.print(
                "Epoch: %d/%d [%.1f%%] | Loss: %.2e | Loss_res: %.2e | Loss_bcs: %.2e | "
                "loss_ics: %.2e | lr: %.2e | Epoch_time: %.2fs | Total: %.1fs | ETA: %.1fs"
                % (it, model.epochs, 100.0 * it / model.epochs if model.epochs > 0 else 0,
                   loss.item(), loss_r.item(), loss_bc1.item(), loss_ics.item(), lr, dt, elapsed, eta))
            if it > 0 and it % every == 0 and rank0:
                model.save_state()

    total = time.time() - t0
    model.logger.print(
        f"Training completed in {total:.2f} seconds ({total / 60:.2f} minutes)")
