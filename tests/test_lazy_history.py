"""Host-side contract of the lazily drained training state (no GPU): ``model.loss_history`` stays the
reference's list attribute, readers and writers run the registered flushers first, and the plateau
scheduler's ``state_dict`` / ``load_state_dict`` / ``step`` do the same and never leak hook state
into checkpoints (reference nn/DVPDESolver.py:20-27, :120-130)."""

import torch

import qcpinn_b200 as qb

ARGS = {
    "batch_size": 64, "epochs": 4, "lr": 0.005, "seed": 1, "print_every": 2,
    "num_qubits": 4, "num_quantum_layers": 1, "classic_network": [3, 50, 1],
    "q_ansatz": "cascade", "problem": "diffusion", "solver": "DV", "encoding": "None",
}


def test_loss_history_is_a_list_that_flushes_before_use(tmp_path):
    model = qb.DVPDESolver(dict(ARGS), qb.Logging(str(tmp_path)), device="cpu")
    assert model.loss_history == [] and isinstance(model.loss_history, list)
    model.loss_history.append(1.5)
    calls = []

    def flusher():
        calls.append(len(model.__dict__["_loss_history"]))
        if len(calls) == 1:
            model.__dict__["_loss_history"].append(2.5)      # a pending device-side entry arrives

    model._lazy_flushers.append(flusher)
    assert model.loss_history == [1.5, 2.5] and calls == [1]
    model.loss_history = [9.0]                               # assignment (restore) flushes first too
    assert calls == [1, 2] and model.loss_history == [9.0]
    assert "loss_history" not in dict(model.named_parameters()) and "_loss_history" in model.__dict__


def test_scheduler_hooks_flush_and_stay_out_of_checkpoints(tmp_path):
    model = qb.DVPDESolver(dict(ARGS), qb.Logging(str(tmp_path)), device="cpu")
    sched = model.scheduler
    calls = []
    sched._qcp_flushers = (lambda: calls.append("f"),)
    sched._qcp_enabled = False
    sd = sched.state_dict()
    assert calls == ["f"] and not any(k.startswith("_qcp_") for k in sd)
    assert {"best", "num_bad_epochs", "cooldown_counter", "last_epoch", "patience", "factor"} <= set(sd)
    sched.step(torch.tensor(0.25))
    assert calls == ["f", "f"] and sched.last_epoch == 1 and sched.best == 0.25
    sched.load_state_dict(sd)
    assert calls == ["f", "f", "f"] and sched.last_epoch == 0
    path = tmp_path / "ckpt.pth"
    model.save_state(str(path))                              # picklable: no callables inside
    state = qb.DVPDESolver.load_state(str(path))
    assert state["scheduler"]["last_epoch"] == 0 and state["loss_history"] == []


def test_sampler_overwrites_given_buffers():
    from qcpinn_b200.data.diffusion_dataset import Sampler, training_boxes, u

    s = Sampler(3, training_boxes("cpu")["dom"], u, device="cpu")
    torch.manual_seed(3)
    X0, y0 = s.sample(7)
    X, y = torch.zeros(7, 3), torch.zeros(7, 1)
    torch.manual_seed(3)
    out = s.sample(7, out=(X, y))
    assert out[0] is X and out[1] is y and torch.equal(X, X0) and torch.equal(y, y0)
