"""Plans and ``torch.autograd.Function`` wrappers over the C-ABI (PyTorch is plumbing only).

``Plan`` owns one ``qcp_plan_t`` (compiled gate program + device workspaces).  The three autograd
functions mirror the three ways the reference enters its hot path:

* ``layer_apply``      -- ``DVQuantumLayer.forward``            (reference nn/DVQuantumLayer.py:151-154)
* ``solver_value``     -- ``DVPDESolver.forward``               (reference nn/DVPDESolver.py:81-110)
* ``solver_residual``  -- ``diffusion_operator(model, t, x, y)`` (reference nn/pde.py:53-72)

Backward passes are hand-written adjoint kernels (first order only: the residual already carries
the second derivatives forward in Taylor mode, so nothing needs to be differentiated twice).
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch
from torch.autograd.function import once_differentiable

from . import _lib
from .program import ENC_AMPLITUDE, ENC_ANGLE, ENC_NONE, CircuitProgram

MODE_VALUE, MODE_RESIDUAL = 1, 6
_DTYPE_CODE = {torch.float32: 0, torch.float64: 1}
MLP_NAMES = ("w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4")

launch_counter = 0   # CUDA kernels launched by this library (bench.py reports it as gpu_launches)
# True: the forward saves its Taylor jets (2*n*S values per point) and the backward runs as three
# lean kernels over them; False: one fused backward kernel recomputes the forward from X.
SAVE_JETS = True


def _count(n: int) -> None:
    global launch_counter
    launch_counter += n


def encoding_code(name) -> int:
    # reference nn/DVQuantumLayer.py:177-182: anything but "amplitude" (incl. "None") means angle
    return ENC_AMPLITUDE if name == "amplitude" else ENC_ANGLE


class Plan:
    """One compiled circuit on one device for one dtype."""

    def __init__(self, program: CircuitProgram, encoding: int, dtype: torch.dtype, hidden: int,
                 device: torch.device, io_dtype: torch.dtype | None = None):
        if dtype not in _DTYPE_CODE:
            raise ValueError(f"qcpinn_b200 supports float32/float64 plans, got {dtype}")
        self.lib = _lib.require_cuda()
        self.program = program
        if program.reupload:
            encoding = ENC_NONE       # the program holds its own per-sample gates
        self.encoding = encoding
        self.dtype = dtype
        self.hidden = int(hidden)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(
                f"qcpinn_b200 runs on CUDA devices only (got {self.device}); there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n = program.n_qubits
        self.n_theta = program.n_theta
        self._handle = ctypes.c_void_p()
        self._key = None
        self._wkey = None
        self._wcache = None
        ops = np.ascontiguousarray(program.ops, dtype=np.int32)
        consts = np.ascontiguousarray(program.consts, dtype=np.complex128).view(np.float64)
        with torch.cuda.device(self.device):
            rc = self.lib.qcp_plan_create(
                ctypes.byref(self._handle), self.n, encoding, _DTYPE_CODE[dtype], self.hidden,
                ops.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ops.shape[0],
                consts.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), program.consts.shape[0],
                self.n_theta)
        _lib.check(rc, "qcp_plan_create")
        self.num_features = self.lib.qcp_plan_num_features(self._handle)
        # "feature" (n <= 4) | "register" (5..10 qubits in registers) | "tiled" (up to 16 qubits,
        # HBM slab swept through registers) | "global" (gate by gate: QCP_ENGINE=L, and every
        # program with per-sample gates inside the layers)
        self.engine = ("feature", "global", "register", "tiled")[self.lib.qcp_plan_engine(self._handle)]
        self.fused_engine = self.engine == "feature"   # the others are per-sample statevector engines
        # element type of X / u / r / grad_u / grad_r / grad_X (weights, jets, grads stay `dtype`)
        self.io_dtype = dtype
        if io_dtype is not None and io_dtype != dtype and self.fused_engine and dtype == torch.float64:
            _lib.check(self.lib.qcp_plan_set_io_dtype(self._handle, _DTYPE_CODE[io_dtype]),
                       "qcp_plan_set_io_dtype")
            self.io_dtype = io_dtype

    def describe(self) -> str:
        buf = ctypes.create_string_buffer(512)
        _lib.check(self.lib.qcp_plan_describe(self._handle, buf, 512), "qcp_plan_describe")
        return buf.value.decode()

    def __del__(self):
        h = getattr(self, "_handle", None)
        try:
            if h is not None and h.value:
                self.lib.qcp_plan_destroy(h)
                h.value = None
        except Exception:   # interpreter shutdown: ctypes may already be torn down
            pass

    # -- helpers -------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _t(self, x: torch.Tensor) -> torch.Tensor:
        if x.device != self.device:
            raise RuntimeError(f"tensor on {x.device}, plan on {self.device}")
        return x.detach().to(self.dtype).contiguous()

    def _io(self, x: torch.Tensor) -> torch.Tensor:
        """Caller-facing array (X, grad_u, grad_r) in the plan's I/O dtype."""
        if x.device != self.device:
            raise RuntimeError(f"tensor on {x.device}, plan on {self.device}")
        return x.detach().to(self.io_dtype).contiguous()

    def invalidate(self):
        self._key = None
        self._wkey = None

    def typed_weights(self, theta, mlp, epoch=None):
        """(theta, 8 MLP tensors) in the plan dtype, contiguous.  When a cast is needed (float32
        parameters, float64 plan) all nine tensors are packed and converted with two kernels and
        the result is cached until any of them changes (3 model calls per step share it)."""
        tensors = [theta] + list(mlp)
        if all(t.dtype == self.dtype and t.is_contiguous() for t in tensors):
            return theta.detach().reshape(-1), [t.detach() for t in mlp]
        # `epoch` (bumped by the solver after every optimizer step) guards against in-place
        # updates that do not touch the tensors' version counters (fused optimizers)
        key = (epoch,) + tuple((t.data_ptr(), t._version) for t in tensors)
        if epoch is None:
            self._wkey = None        # no epoch information: never trust the cache
        if getattr(self, "_wkey", None) != key:
            for t in tensors:
                if t.device != self.device:
                    raise RuntimeError(f"tensor on {t.device}, plan on {self.device}")
            flat = torch.cat([t.detach().reshape(-1) for t in tensors]).to(self.dtype)
            views, off = [], 0
            for t in tensors:
                views.append(flat[off:off + t.numel()].view(t.shape))
                off += t.numel()
            self._wcache = (views[0].reshape(-1), views[1:])
            self._wkey = key
        return self._wcache

    def prepare(self, theta: torch.Tensor, key=None) -> None:
        """theta (dtype T, n_theta) -> feature matrix C.  ``key`` lets callers skip repeats."""
        if key is not None and key == self._key:
            return
        if theta.numel() != self.n_theta:
            raise ValueError(f"expected {self.n_theta} circuit angles, got {theta.numel()}")
        with torch.cuda.device(self.device):
            rc = self.lib.qcp_prepare(self._handle, ctypes.c_void_p(theta.data_ptr()), self._stream())
        _lib.check(rc, "qcp_prepare")
        _count(1)
        self._key = key

    def feature_matrix(self) -> torch.Tensor:
        out = np.empty((self.n, self.num_features), dtype=np.float64)
        with torch.cuda.device(self.device):
            rc = self.lib.qcp_feature_matrix(
                self._handle, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), self._stream())
        _lib.check(rc, "qcp_feature_matrix")
        return torch.from_numpy(out)

    def _mlp(self, tensors) -> _lib.QcpMlp:
        m = _lib.QcpMlp()
        for name, t in zip(MLP_NAMES, tensors):
            setattr(m, name, t.data_ptr())
        return m

    # -- raw calls (tensors already dtype T, contiguous, on device) ----------------------------
    def layer_forward(self, z):
        b = z.shape[0]
        q = torch.empty((self.n, b), dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.qcp_layer_forward(self._handle, ctypes.c_void_p(z.data_ptr()), b,
                                            ctypes.c_void_p(q.data_ptr()), self._stream())
        _lib.check(rc, "qcp_layer_forward")
        _count(1 if b else 0)
        return q

    def layer_backward(self, theta, z, grad_q, need_gz=True):
        b = z.shape[0]
        gz = torch.empty_like(z) if need_gz else None
        gtheta = torch.empty(self.n_theta, dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.qcp_layer_backward(
                self._handle, ctypes.c_void_p(theta.data_ptr()), ctypes.c_void_p(z.data_ptr()),
                ctypes.c_void_p(grad_q.data_ptr()), b,
                ctypes.c_void_p(gz.data_ptr() if gz is not None else None),
                ctypes.c_void_p(gtheta.data_ptr()), self._stream())
        _lib.check(rc, "qcp_layer_backward")
        _count(3 if b else 1)
        return gz, gtheta

    # fraction of the device memory the saved statevectors of one call may take (engines R / T);
    # beyond it the backward recomputes the forward instead
    STATE_SAVE_BUDGET = 0.35

    def _state_save_wanted(self, batch, mode) -> bool:
        if self.fused_engine:
            return False
        es = 8 if self.dtype == torch.float64 else 4
        state_bytes = 2 * es * mode * batch * (1 << self.n)
        total = torch.cuda.get_device_properties(self.device).total_memory
        return state_bytes <= self.STATE_SAVE_BUDGET * total

    def _set_state_save(self, want: bool) -> None:
        if want != getattr(self, "_state_save", True):
            _lib.check(self.lib.qcp_plan_set_state_save(self._handle, int(want)),
                       "qcp_plan_set_state_save")
            self._state_save = want

    def _sync_state_flag(self, save) -> None:
        """The library sizes / interprets a workspace by the plan's state-save flag: put the flag in
        the state the workspace was allocated with before every call that receives it."""
        if save is not None and not self.fused_engine:
            self._set_state_save(bool(getattr(save, "_qcp_has_state", True)))

    def workspace(self, batch, mode):
        """Uninitialised saved-jet workspace for one (batch, mode) forward/backward pair (engines
        R / T: plus the final psi streams when they fit the memory budget)."""
        want = self._state_save_wanted(batch, mode)
        if not self.fused_engine:
            self._set_state_save(want)
        n = self.lib.qcp_solver_workspace_elems(self._handle, batch, mode)
        ws = torch.empty(max(int(n), 1), dtype=self.dtype, device=self.device)
        ws._qcp_has_state = want
        return ws

    def solver_forward(self, X, mlp, mode, coeffs=None, want_streams=False, save=None):
        b = X.shape[0]
        u = torch.empty(b, dtype=self.io_dtype, device=self.device)
        r = torch.empty(b, dtype=self.io_dtype, device=self.device) if mode == MODE_RESIDUAL else None
        streams = (torch.empty((b, 6), dtype=self.io_dtype, device=self.device)
                   if (want_streams and mode == MODE_RESIDUAL) else None)
        c = (ctypes.c_double * 5)(*coeffs) if coeffs is not None else None
        m = self._mlp(mlp)
        self._sync_state_flag(save)
        with torch.cuda.device(self.device):
            rc = self.lib.qcp_solver_forward(
                self._handle, ctypes.byref(m), ctypes.c_void_p(X.data_ptr()), b, mode, c,
                ctypes.c_void_p(u.data_ptr()),
                ctypes.c_void_p(r.data_ptr() if r is not None else None),
                ctypes.c_void_p(streams.data_ptr() if streams is not None else None),
                ctypes.c_void_p(save.data_ptr() if save is not None else None),
                self._stream())
        _lib.check(rc, "qcp_solver_forward")
        _count(1 if b else 0)
        return u, r, streams

    def solver_backward(self, X, mlp, theta, grad_u, grad_r, mode, coeffs=None, need_gx=False,
                        save=None):
        b = X.shape[0]
        sizes = [t.numel() for t in mlp] + [self.n_theta]
        flat = torch.empty(sum(sizes), dtype=self.dtype, device=self.device)
        views, off = [], 0
        for t, sz in zip(list(mlp) + [theta], sizes):
            views.append(flat[off:off + sz].view(t.shape))
            off += sz
        gx = torch.empty_like(X) if need_gx else None
        c = (ctypes.c_double * 5)(*coeffs) if coeffs is not None else None
        m = self._mlp(mlp)
        g = self._mlp(views[:8])
        self._sync_state_flag(save)
        with torch.cuda.device(self.device):
            rc = self.lib.qcp_solver_backward(
                self._handle, ctypes.byref(m), ctypes.c_void_p(theta.data_ptr()),
                ctypes.c_void_p(X.data_ptr()),
                ctypes.c_void_p(grad_u.data_ptr() if grad_u is not None else None),
                ctypes.c_void_p(grad_r.data_ptr() if grad_r is not None else None),
                b, mode, c, ctypes.c_void_p(save.data_ptr() if save is not None else None),
                ctypes.byref(g), ctypes.c_void_p(views[8].data_ptr()),
                ctypes.c_void_p(gx.data_ptr() if gx is not None else None), self._stream())
        _lib.check(rc, "qcp_solver_backward")
        _count(5 if (save is not None and b > 0) else 3)
        return views, gx


def _solver_backward_streams(plan: Plan, X, mlp, theta, grad_streams, save=None):
    """Adjoint of the six-stream forward for a per-point cotangent of ALL streams [B,6]."""
    flat, views = _flat_grad_views(plan, mlp, theta)
    m = plan._mlp(mlp)
    g = plan._mlp(views[:8])
    plan._sync_state_flag(save)
    with torch.cuda.device(plan.device):
        rc = plan.lib.qcp_solver_backward_streams(
            plan._handle, ctypes.byref(m), ctypes.c_void_p(theta.data_ptr()),
            ctypes.c_void_p(X.data_ptr()), ctypes.c_void_p(grad_streams.data_ptr()), X.shape[0],
            ctypes.c_void_p(save.data_ptr() if save is not None else None), ctypes.byref(g),
            ctypes.c_void_p(views[8].data_ptr()), plan._stream())
    _lib.check(rc, "qcp_solver_backward_streams")
    _count(6 if (save is not None and X.shape[0] > 0) else 3)
    return views


def _flat_grad_views(plan: Plan, mlp, theta):
    sizes = [t.numel() for t in mlp] + [plan.n_theta]
    flat = torch.empty(sum(sizes), dtype=plan.dtype, device=plan.device)
    views, off = [], 0
    for t, sz in zip(list(mlp) + [theta], sizes):
        views.append(flat[off:off + sz].view(t.shape))
        off += sz
    return flat, views


def solver_backward_many(plan: Plan, items, mlp, theta, after_adjoints=None):
    """Adjoint kernels of several model calls, ONE partial reduction and ONE theta-gradient kernel.
    ``items`` = [(X, grad_u, grad_r, mode, coeffs, save, need_gx[, stream[, gate]]), ...]; an item's
    adjoint kernels run on its ``stream`` (a torch stream already ordered after the producers of its
    inputs) and the reduction waits for it; with ``gate`` they also wait for the post-MLP adjoint of
    the PREVIOUS item (``qcp_solver_backward_after_post``).  ``after_adjoints()`` runs between the joined adjoint
    launches and the reduction.  Returns (views, [gx])."""
    lib = plan.lib
    flat, views = _flat_grad_views(plan, mlp, theta)
    m = plan._mlp(mlp)
    g = plan._mlp(views[:8])
    gxs = []
    with torch.cuda.device(plan.device):
        stream = plan._stream()
        _lib.check(lib.qcp_solver_backward_begin(plan._handle), "qcp_solver_backward_begin")
        main = torch.cuda.current_stream(plan.device)
        joins = []
        for X, gu, gr, mode, coeffs, save, need_gx, *rest in items:
            side = rest[0] if rest and rest[0] is not None and rest[0] != main else None
            item_stream = ctypes.c_void_p(side.cuda_stream) if side is not None else stream
            gx = torch.empty_like(X) if need_gx else None
            gxs.append(gx)
            c = (ctypes.c_double * 5)(*coeffs) if coeffs is not None else None
            if len(rest) > 1 and rest[1] and side is not None:
                _lib.check(lib.qcp_solver_backward_after_post(plan._handle, item_stream),
                           "qcp_solver_backward_after_post")
            rc = lib.qcp_solver_backward_add(
                plan._handle, ctypes.byref(m), ctypes.c_void_p(X.data_ptr()),
                ctypes.c_void_p(gu.data_ptr() if gu is not None else None),
                ctypes.c_void_p(gr.data_ptr() if gr is not None else None),
                X.shape[0], mode, c, ctypes.c_void_p(save.data_ptr() if save is not None else None),
                ctypes.c_void_p(gx.data_ptr() if gx is not None else None), item_stream)
            _lib.check(rc, "qcp_solver_backward_add")
            if side is not None:
                joins.append(side)
            if X.shape[0]:
                _count(3 if save is not None else 1)
        for side in joins:
            main.wait_stream(side)
        if after_adjoints is not None:
            after_adjoints()
        rc = lib.qcp_solver_backward_finish(
            plan._handle, ctypes.c_void_p(theta.data_ptr()), ctypes.byref(g),
            ctypes.c_void_p(views[8].data_ptr()), stream)
        _lib.check(rc, "qcp_solver_backward_finish")
        # reduce + theta_grad; with saved jets everywhere the d C window is reduced separately on
        # the library's auxiliary stream (theta_grad runs next to the pre-MLP adjoints)
        _count(3 if items and all(it[5] is not None and it[0].shape[0] for it in items) else 2)
    return views, gxs


def mse_seed(plan: Plan, pred, target, weight, grad_out, loss_slot):
    """loss_slot (device float64 scalar view) += mean((pred - target)^2);
    grad_out = 2 * weight * (pred - target) / n.  All arrays float32, contiguous."""
    n = pred.numel()
    if n == 0:
        return
    with torch.cuda.device(plan.device):
        rc = plan.lib.qcp_mse_seed(
            ctypes.c_void_p(pred.data_ptr()), ctypes.c_void_p(target.data_ptr()), n, float(weight),
            ctypes.c_void_p(grad_out.data_ptr()), ctypes.c_void_p(loss_slot.data_ptr()),
            plan._stream())
    _lib.check(rc, "qcp_mse_seed")
    _count(1)


def pack_step(plan: Plan, grads, terms, weights, flat, n_grad):
    """flat[:n_grad] = float32(grads); flat[n_grad] = weights . terms; flat[n_grad+1 : n_grad+4] = terms
    (``grads`` in the plan dtype, ``terms`` three device doubles).  One launch."""
    if grads.numel() != n_grad or not grads.is_contiguous() or grads.dtype != plan.dtype:
        raise ValueError("pack_step: gradient vector does not match the flat buffer")
    w_r, w_bc, w_ic = (float(w) for w in weights)
    with torch.cuda.device(plan.device):
        rc = plan.lib.qcp_pack_step(
            ctypes.c_void_p(grads.data_ptr()), _DTYPE_CODE[plan.dtype],
            int(n_grad), ctypes.c_void_p(terms.data_ptr()), w_r, w_bc, w_ic,
            ctypes.c_void_p(flat.data_ptr()), plan._stream())
    _lib.check(rc, "qcp_pack_step")
    _count(1)


PLATEAU_STATE = 6     # best, num_bad_epochs, cooldown_counter, last_epoch, recorded, reductions


def plateau_step(plan: Plan, metric, state, lr, history, cfg):
    """Device-side ``ReduceLROnPlateau.step(metric)``; ``cfg`` = dict(mode_max, threshold_abs,
    threshold, factor, patience, cooldown, min_lr, eps); ``history`` float32 ring or None."""
    if state.dtype != torch.float64 or state.numel() != PLATEAU_STATE or lr.dtype != torch.float32 \
            or metric.dtype != torch.float32:
        raise ValueError("plateau_step: state float64[6], lr / metric float32 expected")
    with torch.cuda.device(plan.device):
        rc = plan.lib.qcp_plateau_step(
            ctypes.c_void_p(metric.data_ptr()), ctypes.c_void_p(state.data_ptr()),
            ctypes.c_void_p(lr.data_ptr()),
            ctypes.c_void_p(history.data_ptr() if history is not None else 0),
            int(history.numel()) if history is not None else 0,
            int(cfg["mode_max"]), int(cfg["threshold_abs"]), float(cfg["threshold"]),
            float(cfg["factor"]), int(cfg["patience"]), int(cfg["cooldown"]), float(cfg["min_lr"]),
            float(cfg["eps"]), plan._stream())
    _lib.check(rc, "qcp_plateau_step")
    _count(1)


def clip_grads(plan: Plan, flat, n_grad, n_extra, pre_scale, max_norm):
    """In place on the flat float32 buffer [grads | extras]: x pre_scale, then clip_grad_norm_."""
    with torch.cuda.device(plan.device):
        rc = plan.lib.qcp_clip_grads(ctypes.c_void_p(flat.data_ptr()), int(n_grad), int(n_extra),
                                     float(pre_scale), float(max_norm), plan._stream())
    _lib.check(rc, "qcp_clip_grads")
    _count(1)


def _grad_in(plan: Plan, g, like, io=False):
    if g is None:
        return None
    return g.detach().to(plan.io_dtype if io else plan.dtype).reshape(like).contiguous()


class _LayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: Plan, key, z, theta):
        zt, tt = plan._t(z), plan._t(theta).reshape(-1)
        plan.prepare(tt, key)
        ctx.plan, ctx.key = plan, key
        ctx.save_for_backward(zt, tt)
        ctx.in_dtypes = (z.dtype, theta.dtype)
        ctx.theta_shape = theta.shape
        return plan.layer_forward(zt)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_q):
        plan = ctx.plan
        zt, tt = ctx.saved_tensors
        plan.prepare(tt, ctx.key)
        gq = _grad_in(plan, grad_q, (plan.n, zt.shape[0]))
        gz, gtheta = plan.layer_backward(tt, zt, gq, need_gz=ctx.needs_input_grad[2])
        gz = gz.to(ctx.in_dtypes[0]) if gz is not None else None
        gtheta = gtheta.to(ctx.in_dtypes[1]).view(ctx.theta_shape) if ctx.needs_input_grad[3] else None
        return None, None, gz, gtheta


class _SolverFn(torch.autograd.Function):
    """mode VALUE -> u (B,1);  mode RESIDUAL -> (u, r), both (B,1)."""

    @staticmethod
    def forward(ctx, plan: Plan, key, mode, coeffs, X, theta, *mlp):
        Xt = plan._io(X)
        tt, mt = plan.typed_weights(theta, mlp, key)
        plan.prepare(tt, key)
        needs_grad = any(ctx.needs_input_grad[4:])
        save = plan.workspace(Xt.shape[0], mode) if (needs_grad and SAVE_JETS and Xt.shape[0]) else None
        u, r, _ = plan.solver_forward(Xt, mt, mode, coeffs, save=save)
        ctx.save = save
        ctx.plan, ctx.key, ctx.mode, ctx.coeffs = plan, key, mode, coeffs
        ctx.save_for_backward(Xt, tt, *mt)
        ctx.in_dtypes = [X.dtype, theta.dtype] + [w.dtype for w in mlp]
        ctx.theta_shape = theta.shape
        if mode == MODE_RESIDUAL:
            return u.view(-1, 1), r.view(-1, 1)
        return u.view(-1, 1)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_u, grad_r=None):
        plan = ctx.plan
        Xt, tt, *mt = ctx.saved_tensors
        plan.prepare(tt, ctx.key)
        b = Xt.shape[0]
        gu = _grad_in(plan, grad_u, (b,), io=True)
        gr = _grad_in(plan, grad_r, (b,), io=True) if ctx.mode == MODE_RESIDUAL else None
        need_gx = ctx.needs_input_grad[4] and ctx.mode == MODE_VALUE
        views, gx = plan.solver_backward(Xt, mt, tt, gu, gr, ctx.mode, ctx.coeffs, need_gx,
                                         save=ctx.save)
        ctx.save = None   # consumed (overwritten with cotangents)
        dt = ctx.in_dtypes
        gX = gx.to(dt[0]) if gx is not None else None
        if len(set(dt[1:])) == 1 and dt[1] != plan.dtype:
            # one cast kernel for the flat gradient buffer instead of nine
            flat = views[0]._base.to(dt[1])
            off, cast = 0, []
            for v in views:
                cast.append(flat[off:off + v.numel()].view(v.shape))
                off += v.numel()
            views = cast
        gtheta = views[8].to(dt[1]).view(ctx.theta_shape) if ctx.needs_input_grad[5] else None
        gm = [v.to(d) if need else None
              for v, d, need in zip(views[:8], dt[2:], ctx.needs_input_grad[6:])]
        return (None, None, None, None, gX, gtheta, *gm)


class _SolverStreamsFn(torch.autograd.Function):
    """X -> streams [B,6] = (u, u_t, u_x, u_y, u_xx, u_yy), differentiable with respect to every
    parameter: the building block of residual operators that are nonlinear in the streams
    (Navier-Stokes, reference nn/pde.py:2-25).  The coordinates are not differentiated."""

    @staticmethod
    def forward(ctx, plan: Plan, key, X, theta, *mlp):
        Xt = plan._io(X)
        tt, mt = plan.typed_weights(theta, mlp, key)
        plan.prepare(tt, key)
        needs_grad = any(ctx.needs_input_grad[3:])
        save = (plan.workspace(Xt.shape[0], MODE_RESIDUAL)
                if (needs_grad and SAVE_JETS and Xt.shape[0]) else None)
        _, _, streams = plan.solver_forward(Xt, mt, MODE_RESIDUAL, (0.0,) * 5, want_streams=True,
                                            save=save)
        ctx.save = save
        ctx.plan, ctx.key = plan, key
        ctx.save_for_backward(Xt, tt, *mt)
        ctx.in_dtypes = [theta.dtype] + [w.dtype for w in mlp]
        ctx.theta_shape = theta.shape
        return streams

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_streams):
        plan = ctx.plan
        Xt, tt, *mt = ctx.saved_tensors
        plan.prepare(tt, ctx.key)
        gs = _grad_in(plan, grad_streams, (Xt.shape[0], 6), io=True)
        views = _solver_backward_streams(plan, Xt, mt, tt, gs, save=ctx.save)
        ctx.save = None
        gtheta, gm = _cast_grads(plan, views, ctx.in_dtypes[0], ctx.in_dtypes[1:], ctx.theta_shape,
                                 ctx.needs_input_grad[3], ctx.needs_input_grad[4:12])
        return (None, None, None, gtheta, *gm)


def solver_streams_grad(plan: Plan, X, theta, mlp, key=None):
    """Differentiable six-stream forward (n <= 4): [B,6]."""
    if not plan.fused_engine:
        raise NotImplementedError("differentiable Taylor streams cover the n <= 4 engine")
    return _SolverStreamsFn.apply(plan, key, X, theta, *mlp)


def _cast_grads(plan, views, dt_theta, dt_mlp, theta_shape, needs_theta, needs_mlp):
    dts = [dt_theta] + list(dt_mlp)
    if len(set(dts)) == 1 and dts[0] != plan.dtype:
        flat = views[0]._base.to(dts[0])      # one cast kernel for the flat buffer instead of nine
        off, cast = 0, []
        for v in views:
            cast.append(flat[off:off + v.numel()].view(v.shape))
            off += v.numel()
        views = cast
    gtheta = views[8].to(dt_theta).view(theta_shape) if needs_theta else None
    gm = [v.to(d) if need else None for v, d, need in zip(views[:8], dt_mlp, needs_mlp)]
    return gtheta, gm


class _SolverManyFn(torch.autograd.Function):
    """Several model calls of one train step behind ONE autograd node (n <= 4): the forward
    kernels run back to back, the backward launches every call's adjoint kernels and shares one
    partial-sum reduction and one theta-gradient kernel, and autograd sees a single gradient per
    parameter (no 3-way AccumulateGrad adds).  ``specs`` = ((mode, coeffs), ...) per input X."""

    @staticmethod
    def forward(ctx, plan: Plan, key, specs, theta, *rest):
        mlp, Xs = rest[:8], rest[8:]
        tt, mt = plan.typed_weights(theta, mlp, key)
        plan.prepare(tt, key)
        needs_grad = any(ctx.needs_input_grad[3:])
        outs, saved = [], []
        for (mode, coeffs), X in zip(specs, Xs):
            Xt = plan._io(X)
            save = plan.workspace(Xt.shape[0], mode) if (needs_grad and SAVE_JETS and Xt.shape[0]) else None
            u, r, _ = plan.solver_forward(Xt, mt, mode, coeffs, save=save)
            saved.append((Xt, save))
            outs.append(u.view(-1, 1))
            if mode == MODE_RESIDUAL:
                outs.append(r.view(-1, 1))
        ctx.plan, ctx.key, ctx.specs = plan, key, specs
        ctx.saved = saved
        ctx.save_for_backward(tt, *mt)
        ctx.in_dtypes = [theta.dtype] + [w.dtype for w in mlp]
        ctx.x_dtypes = [X.dtype for X in Xs]
        ctx.theta_shape = theta.shape
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        plan = ctx.plan
        tt, *mt = ctx.saved_tensors
        plan.prepare(tt, ctx.key)
        items, gi = [], 0
        n_x = len(ctx.specs)
        for i, ((mode, coeffs), (Xt, save)) in enumerate(zip(ctx.specs, ctx.saved)):
            b = Xt.shape[0]
            gu = _grad_in(plan, grads[gi], (b,), io=True)
            gi += 1
            gr = None
            if mode == MODE_RESIDUAL:
                gr = _grad_in(plan, grads[gi], (b,), io=True)
                gi += 1
            need_gx = ctx.needs_input_grad[3 + 1 + 8 + i] and mode == MODE_VALUE
            items.append((Xt, gu, gr, mode, coeffs, save, need_gx))
        views, gxs = solver_backward_many(plan, items, mt, tt)
        ctx.saved = None
        gtheta, gm = _cast_grads(plan, views, ctx.in_dtypes[0], ctx.in_dtypes[1:], ctx.theta_shape,
                                 ctx.needs_input_grad[3], ctx.needs_input_grad[4:12])
        gX = [g.to(d) if g is not None else None for g, d in zip(gxs, ctx.x_dtypes)]
        return (None, None, None, gtheta, *gm, *gX)


def solver_many(plan: Plan, batches, theta, mlp, key=None):
    """``batches`` = [(X, coeffs or None), ...]; returns a flat tuple: u for value entries,
    (u, r) for residual entries, in order."""
    specs = tuple((MODE_RESIDUAL, tuple(float(c) for c in co)) if co is not None else (MODE_VALUE, None)
                  for _, co in batches)
    return _SolverManyFn.apply(plan, key, specs, theta, *mlp, *[X for X, _ in batches])


def layer_apply(plan: Plan, z, theta, key=None):
    return _LayerFn.apply(plan, key, z, theta)


def solver_value(plan: Plan, X, theta, mlp, key=None):
    return _SolverFn.apply(plan, key, MODE_VALUE, None, X, theta, *mlp)


def solver_residual(plan: Plan, X, theta, mlp, coeffs, key=None):
    return _SolverFn.apply(plan, key, MODE_RESIDUAL, tuple(float(c) for c in coeffs), X, theta, *mlp)


def solver_streams(plan: Plan, X, theta, mlp, coeffs=(1.0, 1.0, 1.0, -0.01, -0.01)):
    """No-grad helper for tests / evaluation: (u, r, streams[B,6])."""
    Xt, tt = plan._io(X), plan._t(theta).reshape(-1)
    mt = [plan._t(w) for w in mlp]
    plan.prepare(tt, None)
    return plan.solver_forward(Xt, mt, MODE_RESIDUAL, tuple(coeffs), want_streams=True)


def fma_peak(dtype: torch.dtype, device=None, iters: int = 20000) -> float:
    """Measured FMA-pipe FLOP/s of the current device (roofline denominator)."""
    lib = _lib.require_cuda()
    device = torch.device(device if device is not None else "cuda")
    out = ctypes.c_double(0.0)
    with torch.cuda.device(device):
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        rc = lib.qcp_bench_fma(_DTYPE_CODE[dtype], iters, ctypes.byref(out), stream)
    _lib.check(rc, "qcp_bench_fma")
    return out.value
