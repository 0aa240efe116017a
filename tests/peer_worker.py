"""Worker of tests/test_gpu_peer_allreduce.py (one process per GPU, launched by torchrun): the
peer-memory gradient exchange (qcp_peer_allreduce_clip) against the NCCL all-reduce + clip route it
replaces, inside the data-parallel train step."""
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qcpinn_b200 as qb  # noqa: E402
from qcpinn_b200.dist import PeerAllReduce, init_from_env  # noqa: E402
from qcpinn_b200.trainer.diffusion_train import TrainStep, _make_averager  # noqa: E402

ARGS = {
    "batch_size": 64, "epochs": 4, "lr": 0.005, "seed": 1, "print_every": 2,
    "num_qubits": 4, "num_quantum_layers": 1, "classic_network": [3, 50, 1],
    "q_ansatz": "cascade", "problem": "diffusion", "solver": "DV", "encoding": "None",
}


def run(rank, dev, peer_on, steps=9):
    os.environ["QCP_PEER_ALLREDUCE"] = "1" if peer_on else "0"
    torch.manual_seed(0)
    model = qb.DVPDESolver(dict(ARGS), qb.Logging(os.path.join(tempfile.gettempdir(), f"peer_r{rank}")), device=dev)
    step = TrainStep(model, 96, _make_averager(model), host_sync=False)
    torch.manual_seed(100 + rank)                      # every rank draws its own points
    for _ in range(steps):
        step()
    step.flush()
    assert step.steady()
    assert bool(step._peer) == peer_on, (step._peer, PeerAllReduce.last_failure)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    calls = step._peer.calls if step._peer else 0
    if step._peer:
        step._peer.close()
    # the captured graphs (NCCL kernels inside on the NCCL route) must be gone before the process
    # group is destroyed: nothing of the step leaves this function
    return list(model.loss_history), flat, calls


def main():
    import faulthandler

    faulthandler.dump_traceback_later(100, exit=True)      # a hang prints its stack and ends the process
    rank, world, local = init_from_env()
    assert world >= 2
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    # 1. the exchange by itself (both parities, clip active and inactive) against NCCL
    peer = PeerAllReduce.create(721, dev, None, 1.0)
    assert peer is not None, PeerAllReduce.last_failure
    for _ in range(3):
        assert peer.self_test(1.0)
    peer.check()
    # 2. inside the train step: same trajectory as the NCCL route, parameters bit-identical on all ranks
    h_peer, p_peer, calls = run(rank, dev, True)
    h_nccl, p_nccl, _ = run(rank, dev, False)
    assert len(h_peer) == len(h_nccl) == 9
    assert all(abs(a - b) <= 2e-6 * abs(b) for a, b in zip(h_peer, h_nccl)), (h_peer, h_nccl)
    assert float((p_peer - p_nccl).abs().max()) <= 1e-5 * float(p_nccl.abs().max())
    gathered = [torch.empty_like(p_peer) for _ in range(world)]
    dist.all_gather(gathered, p_peer)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "parameters differ between ranks"
    hist = [torch.empty(9, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(hist, torch.tensor(h_peer, dtype=torch.float64, device=dev))
    assert all(torch.equal(h, hist[0]) for h in hist), "rank-averaged loss differs between ranks"
    dist.barrier()
    if rank == 0:
        print(f"PEER_OK world={world} exchanges={calls}", flush=True)
    peer.close()
    del peer
    import gc

    gc.collect()
    torch.cuda.synchronize(dev)
    dist.barrier()
    print(f"rank {rank}: teardown", flush=True)
    dist.destroy_process_group()
    print(f"rank {rank}: done", flush=True)


if __name__ == "__main__":
    main()
