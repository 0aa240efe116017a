"""Convection-diffusion PINN training loop (drop-in for reference trainer/diffusion_train.py).

``train(model, nIter=10000, batch_size=128, log_NTK=False, update_lam=False)`` keeps the reference
contract: ``batch_size`` residual points plus ``batch_size // 3`` initial-condition and
``batch_size // 3`` x=0 boundary points per step, loss ``2 MSE_r + 4 MSE_bc + 2 MSE_ic``,
``clip_grad_norm_(1.0)`` (0.1 for the CV solver), Adam, ``ReduceLROnPlateau.step(loss)`` and a
``loss.item()`` per step; ``model.epochs + 1`` iterations; a log line and a checkpoint every
``args["print_every"]`` steps.  ``nIter``, ``log_NTK`` and ``update_lam`` are accepted and unused,
as in the reference.

The step itself is exposed as :class:`TrainStep` so ``bench.py`` times exactly what ``train`` runs.
Under ``torch.distributed`` (world size > 1) every rank draws its own points and the gradients and
loss terms ride in one flat all-reduce before clipping.
"""

import time

import torch

from ..data.diffusion_dataset import Sampler, r, training_boxes, u
from ..nn.pde import diffusion_operator


def fetch_minibatch(sampler, N):
    return sampler.sample(N)


class TrainStep:
    """One optimisation step of the reference loop, split into sample / loss / update phases."""

    def __init__(self, model, batch_size=128, averager=None):
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
        u_bc1_pred = model.forward(X_bcs)
        u_ics_pred = model.forward(X_ics)
        t_r, x_r, y_r = X_res[:, 0:1], X_res[:, 1:2], X_res[:, 2:3]
        _, r_pred = diffusion_operator(model, t_r, x_r, y_r)
        loss_r = model.loss_fn(r_pred, r_res)
        loss_bc1 = model.loss_fn(u_bc1_pred, u_bcs)
        loss_ics = model.loss_fn(u_ics_pred, u_ics)
        loss = 2.0 * loss_r + 4.0 * loss_bc1 + 2.0 * loss_ics
        return loss, time.time() - start, loss_r, loss_bc1, loss_ics

    def update(self, loss, run_backward=True):
        """backward -> (all-reduce) -> clip -> Adam -> plateau scheduler -> loss.item()."""
        model = self.model
        if run_backward:
            loss.backward()
        if self.averager is not None:
            # grads and the scheduler metric share one all-reduce
            loss = self.averager.average(extras=[loss])[0].clone()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=self.max_norm)
        if model.optimizer is not None:
            model.optimizer.step()
        loss = loss.detach()
        if model.scheduler is not None:
            model.scheduler.step(loss)
        value = loss.item()
        model.loss_history.append(value)
        return value

    def __call__(self, batch=None):
        loss, *_ = self.objective(batch)
        return self.update(loss)


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
        loss, dt, loss_r, loss_bc1, loss_ics = step.objective()
        step_times.append(dt)
        if it % every == 0 or it == 0 or model.args.get("use_ibm_hardware", False):
            elapsed = time.time() - t0
            mean_dt = sum(step_times) / len(step_times)
            eta = mean_dt * (model.epochs - it)
            lr = model.optimizer.param_groups[0]["lr"] if model.optimizer else 0.0
            model.logger.print(
                "Epoch: %d/%d [%.1f%%] | Loss: %.2e | Loss_res: %.2e | Loss_bcs: %.2e | "
                "loss_ics: %.2e | lr: %.2e | Epoch_time: %.2fs | Total: %.1fs | ETA: %.1fs"
                % (it, model.epochs, 100.0 * it / model.epochs if model.epochs > 0 else 0,
                   loss.item(), loss_r.item(), loss_bc1.item(), loss_ics.item(), lr, dt, elapsed, eta))
            if it > 0 and it % every == 0 and rank0:
                model.save_state()
        step.update(loss)

    total = time.time() - t0
    model.logger.print(
        f"Training completed in {total:.2f} seconds ({total / 60:.2f} minutes)")
