"""The sync-free CUDA-graph train step: device-resident plateau scheduler (qcp_plateau_step), the
packed gradient / objective (qcp_pack_step), points of the next step drawn at the end of the current
one, and the lazily drained loss history -- all against the host route they replace (reference
trainer/diffusion_train.py:81-90: backward, clip, Adam, scheduler.step(loss), loss.item())."""

import math

import pytest
import torch

import qcpinn_b200 as qb
from helpers import F

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)

ARGS = {
    "batch_size": 64, "epochs": 4, "lr": 0.005, "seed": 1, "print_every": 2,
    "num_qubits": 4, "num_quantum_layers": 1, "classic_network": [3, 50, 1],
    "q_ansatz": "cascade", "problem": "diffusion", "solver": "DV", "encoding": "None",
}


def _model(tmp_path, **over):
    torch.manual_seed(0)
    return qb.DVPDESolver(dict(ARGS, **over), qb.Logging(str(tmp_path)), device=DEV)


def _plan():
    prog = qb.program.compile_program("cascade", 4, 1, None)
    return F.Plan(prog, 0, torch.float64, 50, DEV)


@pytest.mark.parametrize("mode,threshold_mode", [("min", "rel"), ("min", "abs"), ("max", "rel"), ("max", "abs")])
def test_plateau_kernel_repeats_the_torch_scheduler(mode, threshold_mode):
    """qcp_plateau_step against torch.optim.lr_scheduler.ReduceLROnPlateau on the host, step by
    step: best, num_bad_epochs, cooldown_counter, last_epoch and the float32 learning rate (bit
    for bit), through improvements, plateaus, cooldown, the min_lr clamp and the eps rule."""
    plan = _plan()
    g = torch.Generator().manual_seed(5)
    base = torch.cat([torch.linspace(3.0, 1.0, 12), torch.full((14,), 1.0), torch.linspace(1.0, 0.2, 6),
                      torch.full((40,), 0.2)])
    seq = (base + 1e-3 * torch.randn(base.numel(), generator=g)).float()
    if mode == "max":
        seq = -seq
    seq[20] = float("nan")                      # a NaN metric is "not better" on both sides
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=torch.tensor(0.005, dtype=torch.float32))
    kw = dict(mode=mode, factor=0.5, patience=2, threshold=1e-2, threshold_mode=threshold_mode,
              cooldown=1, min_lr=4e-4, eps=1e-8)
    host = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, **kw)
    cfg = {"mode_max": mode == "max", "threshold_abs": threshold_mode == "abs", "threshold": 1e-2,
           "factor": 0.5, "patience": 2, "cooldown": 1, "min_lr": 4e-4, "eps": 1e-8}
    state = torch.tensor([host.best, 0.0, 0.0, float(host.last_epoch), 0.0, 0.0], dtype=torch.float64, device=DEV)
    lr = torch.tensor(0.005, dtype=torch.float32, device=DEV)
    hist = torch.zeros(16, dtype=torch.float32, device=DEV)       # shorter than the sequence on purpose
    metric = torch.zeros(1, dtype=torch.float32, device=DEV)
    reductions = 0
    for i, v in enumerate(seq.tolist()):
        before = float(opt.param_groups[0]["lr"])
        host.step(v)
        reductions += float(opt.param_groups[0]["lr"]) != before
        metric.fill_(v)
        F.plateau_step(plan, metric, state, lr, hist, cfg)
        st = state.cpu().tolist()
        hb = host.best
        assert (st[0] == hb) or (math.isinf(st[0]) and math.isinf(hb)), (i, st, hb)
        assert (int(st[1]), int(st[2]), int(st[3])) == (host.num_bad_epochs, host.cooldown_counter,
                                                        host.last_epoch), (i, st)
        assert lr.item() == opt.param_groups[0]["lr"].item(), i
    assert reductions >= 3 and int(state[5]) == reductions
    assert abs(lr.item() - max(0.005 * 0.5 ** reductions, 4e-4)) < 1e-9   # halved, clamped at min_lr
    assert int(state[4]) == seq.numel()                            # all steps counted ...
    got = hist.cpu()
    want = seq[:16]
    assert torch.equal(got[~want.isnan()], want[~want.isnan()])    # ... the ring keeps what fits


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_pack_step_is_the_cast_plus_weighted_objective(dtype):
    tdt = {"float64": torch.float64, "float32": torch.float32}[dtype]
    prog = qb.program.compile_program("cascade", 4, 1, None)
    plan = F.Plan(prog, 0, tdt, 50, DEV)
    for n in (0, 1, 717, 5000):
        torch.manual_seed(n)
        grads = torch.randn(n, dtype=tdt, device=DEV)
        terms = torch.rand(3, dtype=torch.float64, device=DEV)
        flat = torch.full((n + 4,), -7.0, dtype=torch.float32, device=DEV)
        F.pack_step(plan, grads, terms, (2.0, 4.0, 2.0), flat, n)
        assert torch.equal(flat[:n], grads.float())
        t = terms.cpu()
        assert abs(flat[n].item() - (2 * t[0] + 4 * t[1] + 2 * t[2]).item()) < 1e-6
        assert torch.equal(flat[n + 1:], terms.float())


def _run(tmp_path, tag, steps, host_sync, use_graph=None, patience=None):
    from qcpinn_b200.trainer.diffusion_train import TrainStep

    model = _model(tmp_path / tag)
    if patience is not None:
        model.scheduler.patience = patience
        model.scheduler.factor = 0.5
        model.scheduler.threshold = 0.5        # only a halved loss counts as progress: the lr drops
    step = TrainStep(model, 96, use_graph=use_graph, host_sync=host_sync)
    torch.manual_seed(4242)
    out = [step() for _ in range(steps)]
    return model, step, out


def test_async_replays_match_synchronous_steps(tmp_path):
    """host_sync=False (device scalar returned, nothing waits) vs host_sync=True vs the eager route
    without a graph, same seed: same batches (the captured step draws the NEXT step's points at its
    end), same loss history, same scheduler state and learning rate, same parameters."""
    n = 14
    m_sync, s_sync, out_sync = _run(tmp_path, "sync", n, True, patience=1)
    m_lazy, s_lazy, out_lazy = _run(tmp_path, "lazy", n, False, patience=1)
    m_eager, s_eager, out_eager = _run(tmp_path, "eager", n, True, use_graph=False, patience=1)
    assert s_sync.steady() and s_lazy.steady() and s_lazy._plateau is not None
    assert all(isinstance(v, float) for v in out_sync)
    assert torch.is_tensor(out_lazy[-1]) and out_lazy[-1].is_cuda      # replays return the device scalar
    assert m_lazy._device_plateau.pending > 0                            # nothing drained yet
    h_sync, h_lazy, h_eager = m_sync.loss_history, m_lazy.loss_history, m_eager.loss_history
    assert m_lazy._device_plateau.pending == 0
    assert len(h_sync) == len(h_lazy) == len(h_eager) == n
    assert h_sync == out_sync
    assert h_lazy == h_sync                                              # bit-identical
    assert all(abs(a - b) <= 1e-6 * abs(b) for a, b in zip(h_eager, h_sync)), (h_eager, h_sync)
    sd_sync, sd_lazy, sd_eager = (m.scheduler.state_dict() for m in (m_sync, m_lazy, m_eager))
    for k in ("best", "num_bad_epochs", "cooldown_counter", "last_epoch"):
        assert sd_sync[k] == sd_lazy[k], k
    assert sd_eager["last_epoch"] == sd_sync["last_epoch"] == n
    assert not any(k.startswith("_qcp_") for k in sd_lazy)
    lr = [float(m.optimizer.param_groups[0]["lr"]) for m in (m_sync, m_lazy, m_eager)]
    assert lr[0] == lr[1] == lr[2] and lr[0] < 0.005                    # patience 1: the lr did drop
    for a, b in zip(m_sync.parameters(), m_lazy.parameters()):
        assert torch.equal(a, b)


def test_host_scheduler_calls_between_replays_are_respected(tmp_path):
    """A host-side scheduler.step() / load_state_dict() between replays drains the device side
    first and is uploaded before the next replay; the step counter stays continuous."""
    model, step, _ = _run(tmp_path, "mix", 6, False)
    twin = model._device_plateau
    pending = twin.pending
    assert pending == 6 - step.EAGER_STEPS_BEFORE_CAPTURE
    model.scheduler.step(123.0)                     # host call: flush, then the host step itself
    assert twin.pending == 0 and model.scheduler.last_epoch == 7
    assert len(model.__dict__["_loss_history"]) == 6
    step()
    step()
    assert model.scheduler.state_dict()["last_epoch"] == 9
    saved = model.scheduler.state_dict()
    for _ in range(3):
        step()
    model.scheduler.load_state_dict(saved)          # back to epoch 9 (history keeps all steps)
    step()
    assert model.scheduler.state_dict()["last_epoch"] == 10
    assert len(model.loss_history) == 6 + 2 + 3 + 1
    # reconfiguring the scheduler re-captures the step with the new constants
    model.scheduler.patience = 0
    model.scheduler.factor = 0.25
    model.scheduler.threshold = 10.0                # nothing counts as an improvement any more
    lr0 = float(model.optimizer.param_groups[0]["lr"])
    for _ in range(3):
        step()
    assert float(model.optimizer.param_groups[0]["lr"]) < lr0


def test_ring_overflow_drains_itself(tmp_path, monkeypatch):
    from qcpinn_b200.trainer import diffusion_train as dt

    monkeypatch.setattr(dt.DevicePlateau, "CAPACITY", 4)
    model, step, _ = _run(tmp_path, "ring", 14, False)
    assert len(model.__dict__["_loss_history"]) >= 8        # drained in batches of 4, unprompted
    assert len(model.loss_history) == 14
    assert all(math.isfinite(v) for v in model.loss_history)
