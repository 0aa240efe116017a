"""Host-side data-parallel logic on CPU: world_size-2 gloo runs of the gradient averager (engine
callback path used with an unmodified trainer, explicit path used by TrainStep) and of the
rank-consistent plateau scheduler.  SURVEY.md section 4, item 6 (data-parallel equivalence)."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import qcpinn_b200 as qb


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(3, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.rand(64, 3, generator=g), torch.rand(64, 1, generator=g)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from qcpinn_b200.dist import GradientAverager, init_from_env
    from qcpinn_b200.nn.DVPDESolver import _RankConsistentPlateau

    r, w, _ = init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    X, Y = _data()
    shard = slice(rank * 32, (rank + 1) * 32)

    # (1) engine-callback path: plain loss.backward() ends with averaged grads
    model = _make_model()
    avg = GradientAverager(model)
    loss = ((model(X[shard]) - Y[shard]) ** 2).mean()
    loss.backward()
    assert avg.calls == 1
    g1 = [p.grad.clone() for p in model.parameters()]
    # second backward -> exactly one more all-reduce
    for p in model.parameters():
        p.grad = None
    ((model(X[shard]) - Y[shard]) ** 2).mean().backward()
    assert avg.calls == 2
    avg.remove()

    # (2) explicit path with a loss scalar riding in the same buffer
    model2 = _make_model()
    avg2 = GradientAverager(model2, extra=1)
    avg2.enabled = False
    loss2 = ((model2(X[shard]) - Y[shard]) ** 2).mean()
    loss2.backward()
    mean_loss = avg2.average(extras=[loss2])[0].item()
    g2 = [p.grad.clone() for p in model2.parameters()]

    # (3) scheduler sees the same metric on every rank
    opt = torch.optim.SGD(model2.parameters(), lr=1.0)
    sched = _RankConsistentPlateau(opt, mode="min", factor=0.5, patience=0)
    sched._qcp_enabled = True
    sched.step(torch.tensor(1.0))
    sched.step(torch.tensor(0.0 if rank == 0 else 4.0))   # mean 2.0 > best 1.0 on BOTH ranks
    lr = opt.param_groups[0]["lr"]

    torch.save({"g1": g1, "g2": g2, "mean_loss": mean_loss, "lr": lr},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gradient_average_equals_full_batch(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = [torch.load(tmp_path / f"rank{r}.pt") for r in range(2)]

    model = _make_model()
    X, Y = _data()
    full = ((model(X) - Y) ** 2).mean()
    full.backward()
    want = [p.grad for p in model.parameters()]
    for key in ("g1", "g2"):
        for r in range(2):
            for got, ref in zip(res[r][key], want):
                assert torch.allclose(got, ref, rtol=1e-5, atol=1e-7)
    assert abs(res[0]["mean_loss"] - full.item()) < 1e-6
    assert res[0]["mean_loss"] == res[1]["mean_loss"]
    assert res[0]["lr"] == res[1]["lr"] == 0.5


def test_averager_requires_initialised_process_group():
    from qcpinn_b200.dist import GradientAverager

    if dist.is_initialized():
        pytest.skip("process group already up")
    with pytest.raises(RuntimeError):
        GradientAverager(_make_model())
