"""``DVPDESolver`` -- drop-in for reference nn/DVPDESolver.py on the fused sm_100a path.

Kept from the reference: constructor signature ``(args, logger, data=None, device=None)``; the
``args`` keys read (``batch_size, num_qubits, epochs, classic_network, lr`` required;
``encoding`` optional); submodule names ``preprocessor`` / ``postprocessor`` / ``quantum_layer``
(so state dicts and ``model.pth`` checkpoints interchange); Adam + ReduceLROnPlateau(0.9, 1000) +
MSELoss; xavier-normal / zero-bias on the pre MLP only; ``forward(x: (B,3)) -> (B,1)`` float32
with the "log then re-raise" error convention; ``save_state`` / ``load_state`` dict layout.

Different by design: ``forward`` is ONE fused kernel (pre MLP -> circuit -> post MLP), and
``taylor_residual`` (used by ``nn.pde.diffusion_operator``) is ONE fused Taylor-mode kernel; both
have hand-written adjoint kernels.  The whole module, including ``quantum_layer``, must live on a
CUDA device -- there is no CPU path.
"""

from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import functional as F
from ..utils.logger import Logging
from .DVQuantumLayer import DVQuantumLayer


class _RankConsistentPlateau(torch.optim.lr_scheduler.ReduceLROnPlateau):
    """ReduceLROnPlateau whose metric is averaged over data-parallel ranks, so every rank takes
    the same LR decisions when the unmodified trainer calls ``scheduler.step(loss)``."""

    _qcp_group = None
    _qcp_enabled = False
    # callables that bring this object up to date with a device-resident twin (the CUDA-graph
    # train step advances best / num_bad_epochs / cooldown_counter / last_epoch on the device,
    # ``trainer.diffusion_train.DevicePlateau``); run before anything reads or changes the state
    _qcp_flushers = ()

    def _qcp_flush(self):
        for f in tuple(self._qcp_flushers):
            f()

    def state_dict(self):
        self._qcp_flush()
        return {k: v for k, v in super().state_dict().items() if not k.startswith("_qcp_")}

    def load_state_dict(self, state_dict):
        self._qcp_flush()
        return super().load_state_dict(state_dict)

    def step(self, metrics, *a, **kw):
        self._qcp_flush()
        if self._qcp_enabled and torch.is_tensor(metrics):
            import torch.distributed as dist

            m = metrics.detach().clone()
            dist.all_reduce(m, op=dist.ReduceOp.SUM, group=self._qcp_group)
            metrics = m / dist.get_world_size(self._qcp_group)
        return super().step(metrics, *a, **kw)


class DVPDESolver(nn.Module):
    def __init__(self, args, logger: Logging, data=None, device=None):
        super().__init__()
        self.logger = logger
        self.device = device
        self.args = args
        self.data = data
        self.batch_size = self.args["batch_size"]
        self.num_qubits = self.args["num_qubits"]
        self.epochs = self.args["epochs"]
        self.optimizer = None
        self.scheduler = None
        self._lazy_flushers = []      # see loss_history
        self.loss_history = []
        self.encoding = self.args.get("encoding", "angle")
        self.draw_quantum_circuit_flag = True
        self.classic_network = self.args["classic_network"]
        self.total_training_time = 0
        self.total_memory_peak = 0

        hidden = self.classic_network[-2]
        self.preprocessor = nn.Sequential(
            nn.Linear(self.classic_network[0], hidden),
            nn.Tanh(),
            nn.Linear(hidden, self.num_qubits),
        ).to(self.device)
        self.postprocessor = nn.Sequential(
            nn.Linear(self.num_qubits, hidden),
            nn.Tanh(),
            nn.Linear(hidden, self.classic_network[-1]),
        ).to(self.device)
        self.activation = nn.Tanh()

        # The reference leaves quantum_layer on the CPU (PennyLane moves data itself); the fused
        # kernels need every parameter on the model's device.
        self.quantum_layer = self._build_quantum_layer(self.args).to(self.device)
        # opt-in slow path (args["diff_mode"] = "autograd"): the three sub-modules are chained with
        # ordinary differentiable torch ops like the reference's forward (nn/DVPDESolver.py:81-110),
        # so nested autograd.grad(create_graph=True) calls work; the fused entry points are hidden
        # so that nn.pde's operators take the reference's generic nested-autograd formulation
        self.diff_mode = self.args.get("diff_mode", "kernels")
        if self.diff_mode == "autograd":
            self.taylor_residual = None
            self.taylor_streams_grad = None
            self.forward_many = None

        # On a CUDA device Adam is built capturable with a device-resident learning rate, so a
        # whole train step can be replayed as one CUDA graph (trainer.diffusion_train.TrainStep)
        # and ReduceLROnPlateau's updates still reach the captured kernels.
        on_cuda = torch.device(self.device).type == "cuda" if self.device is not None else False
        trainable = [p for p in self.parameters() if p.requires_grad]
        if on_cuda:
            lr0 = torch.tensor(float(self.args["lr"]), dtype=torch.float32, device=self.device)
            self.optimizer = torch.optim.Adam(trainable, lr=lr0, capturable=True, fused=True)
        else:
            self.optimizer = torch.optim.Adam(trainable, lr=self.args["lr"])
        # fused / capturable optimizers update parameters without touching their version counters
        self.optimizer.register_step_post_hook(lambda *_: self.quantum_layer.mark_updated())
        self.scheduler = self._make_scheduler()
        self.loss_fn = torch.nn.MSELoss()
        self.log_path = self.logger.get_output_dir()
        self._initialize_weights()
        self._dp = None

    # ``loss_history`` is the reference's plain list attribute.  The CUDA-graph train step records
    # the per-step loss on the device and drains it in batches (no host round trip per step), so
    # every read first runs the registered flushers.
    @property
    def loss_history(self):
        for f in tuple(self.__dict__.get("_lazy_flushers", ())):
            f()
        return self.__dict__["_loss_history"]

    @loss_history.setter
    def loss_history(self, value):
        for f in tuple(self.__dict__.get("_lazy_flushers", ())):
            f()
        self.__dict__["_loss_history"] = value

    # construction hooks (the single-file trainer's HybridQPINN overrides them)
    def _build_quantum_layer(self, args):
        return DVQuantumLayer(args)

    def _make_scheduler(self):
        return _RankConsistentPlateau(self.optimizer, mode="min", factor=0.9, patience=1000)

    def _initialize_weights(self):
        for layer in self.preprocessor:
            if isinstance(layer, nn.Linear):
                nn.init.xavier_normal_(layer.weight)
                if layer.bias is not None:
                    nn.init.zeros_(layer.bias)

    # -- fused path ----------------------------------------------------------------------------
    def _check_fused(self):
        net = self.classic_network
        if net[0] not in (2, 3) or net[-1] < 1:
            raise NotImplementedError(
                f"fused kernels cover classic_network=[3, H, k] and [2, H, k]; got {net}")

    @property
    def n_outputs(self):
        return int(self.classic_network[-1])

    def _mlp_tensors(self, output=None):
        """The eight MLP tensors the kernels read.  Multi-output solvers (Navier-Stokes: u, v, p,
        reference nn/pde.py:2-25) run the single-output kernels once per output row of the last
        Linear layer (``output`` = row index); autograd sums the shared parameters' gradients."""
        pre, post = self.preprocessor, self.postprocessor
        w1 = pre[0].weight
        if w1.shape[1] == 2:
            # two-input solvers (wave / Klein-Gordon / Helmholtz, reference nn/pde.py:26-52,73-95):
            # the coordinates ride in the kernel's x / y slots behind a zero t column; autograd
            # slices the gradient of the padded weight back
            w1 = torch.nn.functional.pad(w1, (1, 0))
        w4, b4 = post[2].weight, post[2].bias
        if self.n_outputs != 1:
            if output is None:
                raise NotImplementedError(
                    f"this entry point covers single-output solvers; the model has {self.n_outputs} "
                    "outputs (use forward() / taylor_streams_grad())")
            w4, b4 = w4[output:output + 1], b4[output:output + 1]
        return (w1, pre[0].bias, pre[2].weight, pre[2].bias,
                post[0].weight, post[0].bias, w4, b4)

    def _kernel_input(self, x):
        """(B, 3) kernel input: two-input solvers get a zero t column in front."""
        if x.shape[1] == 2 and self.classic_network[0] == 2:
            return torch.nn.functional.pad(x, (1, 0))
        return x

    def _plan(self, device) -> F.Plan:
        self._check_fused()
        # the module's tensors are float32 (like the reference's): let a float64 plan read / write
        # them directly instead of casting around every call
        return self.quantum_layer.plan(device, hidden=self.classic_network[-2],
                                       io_dtype=torch.float32)

    def _device_of(self, x):
        dev = self.quantum_layer.params.device
        if x.device != dev:
            raise RuntimeError(f"input on {x.device} but model parameters on {dev}")
        return dev

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        try:
            if x.dim() != 2:
                raise ValueError(f"Expected 2D input tensor, got shape {x.shape}")
            if self.draw_quantum_circuit_flag:
                self.draw_quantum_circuit(x)
                self.draw_quantum_circuit_flag = False
            if self.diff_mode == "autograd":
                q = self.quantum_layer(self.preprocessor(x)).to(torch.float32)    # (n, B)
                return self.postprocessor(q.T.reshape(-1, self.num_qubits))
            plan = self._plan(self._device_of(x))
            if self.n_outputs != 1:
                cols = [F.solver_value(plan, self._kernel_input(x), self.quantum_layer.params,
                                       self._mlp_tensors(o), self.quantum_layer.theta_key())
                        for o in range(self.n_outputs)]
                return torch.cat(cols, dim=1).to(torch.float32)
            u = F.solver_value(plan, self._kernel_input(x), self.quantum_layer.params,
                               self._mlp_tensors(), self.quantum_layer.theta_key())
            return u.to(torch.float32)
        except Exception as e:
            self.logger.print(f"Forward pass failed: {str(e)}")
            raise

    def taylor_residual(self, X: torch.Tensor, coeffs):
        """(u, r) with r = c0 u_t + c1 u_x + c2 u_y + c3 u_xx + c4 u_yy, each (B,1) float32."""
        try:
            if X.dim() != 2:
                raise ValueError(f"Expected 2D input tensor, got shape {X.shape}")
            if self.draw_quantum_circuit_flag:
                self.draw_quantum_circuit(X)
                self.draw_quantum_circuit_flag = False
            plan = self._plan(self._device_of(X))
            u, r = F.solver_residual(plan, self._kernel_input(X), self.quantum_layer.params,
                                     self._mlp_tensors(), coeffs, self.quantum_layer.theta_key())
            return u.to(torch.float32), r.to(torch.float32)
        except Exception as e:
            self.logger.print(f"Forward pass failed: {str(e)}")
            raise

    def forward_many(self, batches):
        """All model calls of one train step behind one autograd node (n <= 4; falls back to
        separate calls otherwise).  ``batches`` = [(X, coeffs or None), ...]: ``None`` means
        ``forward(X)`` -> u, a 5-tuple of PDE coefficients means ``taylor_residual`` -> (u, r).
        Returns a list with one entry per batch (u, or the pair (u, r)), float32 like forward()."""
        try:
            for X, _ in batches:
                if X.dim() != 2:
                    raise ValueError(f"Expected 2D input tensor, got shape {X.shape}")
            if self.draw_quantum_circuit_flag:
                self.draw_quantum_circuit(batches[0][0])
                self.draw_quantum_circuit_flag = False
            plan = self._plan(self._device_of(batches[0][0]))
            if not plan.fused_engine:
                return [self.forward(X) if co is None else self.taylor_residual(X, co)
                        for X, co in batches]
            batches = [(self._kernel_input(X), co) for X, co in batches]
            flat = F.solver_many(plan, batches, self.quantum_layer.params, self._mlp_tensors(),
                                 self.quantum_layer.theta_key())
            out, i = [], 0
            for _, co in batches:
                if co is None:
                    out.append(flat[i].to(torch.float32))
                    i += 1
                else:
                    out.append((flat[i].to(torch.float32), flat[i + 1].to(torch.float32)))
                    i += 2
            return out
        except Exception as e:
            self.logger.print(f"Forward pass failed: {str(e)}")
            raise

    # -- fused train step (no autograd graph) ------------------------------------------------------
    N_LOSS_SLOTS = 4            # loss, loss_r, loss_bc, loss_ic behind the flat gradient

    def flat_grad_buffer(self):
        """One persistent float32 buffer [all parameter gradients | 4 loss scalars]; every
        ``p.grad`` is a view into it, in ``parameters()`` order (= the kernels' gradient order)."""
        buf = getattr(self, "_flat_grad", None)
        params = [p for p in self.parameters() if p.requires_grad]
        numel = sum(p.numel() for p in params)
        if buf is None or buf.numel() != numel + self.N_LOSS_SLOTS or buf.device != params[0].device:
            buf = torch.zeros(numel + self.N_LOSS_SLOTS, dtype=torch.float32, device=params[0].device)
            self._flat_grad = buf
        off = 0
        for p in params:
            n = p.numel()
            view = buf[off:off + n].view_as(p)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                p.grad = view
            off += n
        return buf, numel

    def supports_fused_step(self):
        try:
            dev = self.quantum_layer.params.device
            return dev.type == "cuda" and self.classic_network[0] == 3 and \
                self.diff_mode == "kernels" and \
                self.n_outputs == 1 and self._plan(dev).fused_engine and \
                all(p.dtype == torch.float32 for p in self.parameters())
        except Exception:
            return False

    def prepare_step(self):
        """Cast the current weights to the plan dtype and run ``qcp_prepare`` for the current angles
        (both cached on ``theta_key``): lets a trainer issue them early, next to unrelated work."""
        plan = self._plan(self.quantum_layer.params.device)
        key = self.quantum_layer.theta_key()
        tt, _ = plan.typed_weights(self.quantum_layer.params, self._mlp_tensors(), key)
        plan.prepare(tt, key)

    def train_step_grads(self, batch, coeffs, weights=(2.0, 4.0, 2.0), after_adjoints=None):
        """The reference objective ``w_r MSE_r + w_bc MSE_bc + w_ic MSE_ic`` (reference
        trainer/diffusion_train.py:30-49) and ALL its parameter gradients without an autograd
        graph: forward kernels -> MSE seeds -> adjoint kernels -> one cast into the flat gradient
        buffer (``qcp_pack_step``).  ``batch`` = (X_ics, u_ics, X_bcs, u_bcs, X_res, r_res), float32 on the model's
        device.  Returns the flat buffer and the number of gradient elements; the buffer's last
        four slots hold (loss, loss_r, loss_bc, loss_ic).  ``after_adjoints()`` is called once every
        kernel that reads the batch has been issued and joined on the current stream (a trainer
        refills its static batch there, next to the reductions)."""
        X_ics, u_ics, X_bcs, u_bcs, X_res, r_res = batch
        dev = self._device_of(X_res)
        plan = self._plan(dev)
        if not plan.fused_engine:
            raise NotImplementedError("train_step_grads covers the n <= 4 engine")
        key = self.quantum_layer.theta_key()
        tt, mt = plan.typed_weights(self.quantum_layer.params, self._mlp_tensors(), key)
        plan.prepare(tt, key)
        w_r, w_bc, w_ic = weights
        nb, ni, nr = X_bcs.shape[0], X_ics.shape[0], X_res.shape[0]
        X_val = torch.cat([X_bcs.detach(), X_ics.detach()]).to(plan.io_dtype)
        X_r = X_res.detach().to(plan.io_dtype).contiguous()
        ws_v = plan.workspace(nb + ni, F.MODE_VALUE)
        ws_r = plan.workspace(nr, F.MODE_RESIDUAL)
        terms = torch.zeros(3, dtype=torch.float64, device=dev)            # r, bc, ic
        t_bc, t_ic = u_bcs.detach().reshape(-1).contiguous(), u_ics.detach().reshape(-1).contiguous()
        t_r = r_res.detach().reshape(-1).contiguous()
        # the IC/BC chain and the residual chain are independent until the partial reduction: run
        # the (small) IC/BC kernels on a side stream so they fill the tails of the residual kernels
        main = torch.cuda.current_stream(dev)
        side = self._value_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            u_v, _, _ = plan.solver_forward(X_val, mt, F.MODE_VALUE, save=ws_v)
            gu_v = torch.empty_like(u_v)
            F.mse_seed(plan, u_v[:nb], t_bc, w_bc, gu_v[:nb], terms[1:2])
            F.mse_seed(plan, u_v[nb:], t_ic, w_ic, gu_v[nb:], terms[2:3])
        _, r_r, _ = plan.solver_forward(X_r, mt, F.MODE_RESIDUAL, coeffs, save=ws_r)
        gr = torch.empty_like(r_r)
        F.mse_seed(plan, r_r, t_r, w_r, gr, terms[0:1])
        # residual chain first; the IC/BC adjoints (side stream) are held back until its post-MLP
        # adjoint is done, i.e. until its contraction adjoint is ready too: the block scheduler then
        # places that one first (higher stream priority in the captured step) and the IC/BC kernels
        # fill its tail.  Without the gate a 2 us race at the end of post_backward decides whether
        # they slip in front instead (+90 us per step at 524 288 points, profiles/r02_timeline_8gpu*)
        gate = os.environ.get("QCP_GATE_VALUE", "1") != "0"
        views, _ = F.solver_backward_many(
            plan, [(X_r, None, gr, F.MODE_RESIDUAL, coeffs, ws_r, False, None),
                   (X_val, gu_v, None, F.MODE_VALUE, None, ws_v, False, side, gate)], mt, tt,
            after_adjoints=after_adjoints)
        flat, numel = self.flat_grad_buffer()
        # cast into the optimizer's float32 buffer + weighted objective and its terms: one launch
        F.pack_step(plan, views[0]._base, terms, (w_r, w_bc, w_ic), flat, numel)
        return flat, numel

    def _value_stream(self, dev):
        """Side stream of the IC/BC chain in :meth:`train_step_grads` (the current stream itself when
        ``QCP_OVERLAP=0``)."""
        if os.environ.get("QCP_OVERLAP", "1") == "0":
            return torch.cuda.current_stream(dev)
        streams = self.__dict__.setdefault("_value_streams", {})
        if dev not in streams:
            streams[dev] = torch.cuda.Stream(device=dev)
        return streams[dev]

    def taylor_streams_grad(self, X: torch.Tensor):
        """(B, n_outputs, 6) = (u, u_t, u_x, u_y, u_xx, u_yy) of every output, connected to every
        parameter (first-order adjoint kernels): what residual operators that multiply streams
        need (``nn.pde.navier_stokes_2D_operator``).  The coordinates are detached."""
        try:
            if X.dim() != 2:
                raise ValueError(f"Expected 2D input tensor, got shape {X.shape}")
            plan = self._plan(self._device_of(X))
            Xk = self._kernel_input(X).detach()
            key = self.quantum_layer.theta_key()
            outs = [F.solver_streams_grad(plan, Xk, self.quantum_layer.params,
                                          self._mlp_tensors(o if self.n_outputs != 1 else None), key)
                    for o in range(self.n_outputs)]
            return torch.stack(outs, dim=1).to(torch.float32)
        except Exception as e:
            self.logger.print(f"Forward pass failed: {str(e)}")
            raise

    def taylor_streams(self, X: torch.Tensor):
        """No-grad evaluation helper: (B,6) = u, u_t, u_x, u_y, u_xx, u_yy."""
        plan = self._plan(self._device_of(X))
        with torch.no_grad():
            return F.solver_streams(plan, self._kernel_input(X), self.quantum_layer.params,
                                    self._mlp_tensors())[2]

    # -- data parallel -------------------------------------------------------------------------
    def enable_data_parallel(self, process_group=None):
        """Average gradients over ranks at the end of every backward pass (one flat all-reduce,
        issued from an autograd-engine callback so an unmodified trainer needs no changes)."""
        from ..dist import GradientAverager

        self._dp = GradientAverager(self, process_group)
        self.scheduler._qcp_group = process_group
        self.scheduler._qcp_enabled = True
        return self._dp

    # -- checkpoint ----------------------------------------------------------------------------
    def save_state(self, path=None):
        state = {
            "args": self.args,
            "classic_network": self.classic_network,
            "quantum_params": self.quantum_layer.state_dict(),
            "preprocessor": self.preprocessor.state_dict(),
            "quantum_layer": self.quantum_layer.state_dict(),
            "postprocessor": self.postprocessor.state_dict(),
            "optimizer": self.optimizer.state_dict(),
            "scheduler": self.scheduler.state_dict(),
            "loss_history": self.loss_history,
            "log_path": self.log_path,
        }
        model_path = os.path.join(self.log_path, "model.pth") if path is None else path
        with open(model_path, "wb") as f:
            torch.save(state, f)
        self.logger.print(f"Model state saved to {model_path}")

    @classmethod
    def load_state(cls, file_path, map_location=None):
        if map_location is None:
            map_location = torch.device("cpu")
        with open(file_path, "rb") as f:
            return torch.load(f, map_location=map_location, weights_only=False)

    def restore(self, state):
        """Load a ``save_state`` dict (ours or the reference's) back into this model."""
        self.preprocessor.load_state_dict(state["preprocessor"])
        self.quantum_layer.load_state_dict(state["quantum_layer"])
        self.quantum_layer.mark_updated()
        self.postprocessor.load_state_dict(state["postprocessor"])
        if "optimizer" in state:
            self._load_optimizer_state(state["optimizer"])
        if "scheduler" in state:
            self.scheduler.load_state_dict(state["scheduler"])
        self.loss_history = list(state.get("loss_history", []))
        return self

    def _load_optimizer_state(self, opt_state):
        """``optimizer.load_state_dict`` replaces ``param_groups`` wholesale: a reference checkpoint
        brings a float lr without capturable / fused, one of ours read with the default
        ``map_location="cpu"`` a CPU lr tensor.  Either would break the CUDA-graph step (capture
        fails, or the lr is frozen into the graph and plateau reductions stop applying), so the
        invariants of ``__init__`` are re-installed: device-resident lr tensor (the SAME tensor
        object a captured graph already reads), capturable + fused, ``step`` counters on the device."""
        lr_tensors = [g["lr"] for g in self.optimizer.param_groups]
        self.optimizer.load_state_dict(opt_state)
        dev = self.quantum_layer.params.device
        if dev.type != "cuda":
            return
        for g, lr0 in zip(self.optimizer.param_groups, lr_tensors):
            loaded = float(g["lr"])
            if isinstance(lr0, torch.Tensor) and lr0.device.type == "cuda":
                lr0.fill_(loaded)
                g["lr"] = lr0
            else:
                g["lr"] = torch.tensor(loaded, dtype=torch.float32, device=dev)
            g["capturable"] = True
            g["fused"] = True
            g["foreach"] = False
        for st in self.optimizer.state.values():
            if "step" in st:
                st["step"] = torch.as_tensor(st["step"], dtype=torch.float32).to(dev)
            for k in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
                if k in st:
                    st[k] = st[k].to(dev)

    def draw_quantum_circuit(self, x):
        # matplotlib / qml.draw_mpl are not part of this stack: log the gate program instead.
        if self.draw_quantum_circuit_flag:
            try:
                self.logger.print("The circuit used in the study:")
                self.logger.print(self.quantum_layer.describe())
            except Exception as e:
                self.logger.print(f"Failed to draw quantum circuit: {str(e)}")
