"""Collocation samplers and the analytic convection-diffusion target (reference
data/diffusion_dataset.py:5-56).  Cheap element-wise torch work that stays on the model's device.

Gaussian pulse ``u = exp(-100((x-.5)^2 + (y-.5)^2)) exp(-t)``; forcing
``r = u_t + v_x u_x + v_y u_y - D (u_xx + u_yy)`` in closed form.
"""

import torch

default_D = 0.01
default_v_x = 1.0
default_v_y = 1.0


class Sampler:
    """Uniform points in the axis-aligned box ``coords`` (2, dim) with targets ``func(points)``."""

    def __init__(self, dim, coords, func, name=None, device="cpu"):
        self.dim = dim
        self.coords = coords
        self.func = func
        self.name = name
        self.device = device

    def draw(self, N):
        """The uniform random numbers of one ``sample(N)`` call (to be passed back as ``rnd``)."""
        return torch.rand(N, self.dim, device=self.device)

    def sample(self, N, out=None, rnd=None):
        """``out`` = (points, targets) of an earlier call to overwrite in place (static buffers of a
        captured train step), ``rnd`` = numbers drawn earlier by :meth:`draw`; the reference
        signature is ``sample(N)``."""
        pts = self.draw(N) if rnd is None else rnd
        if pts.shape != (N, self.dim):
            raise ValueError(f"Sampler.sample: rnd has shape {tuple(pts.shape)}, expected {(N, self.dim)}")
        fused = self._fused(pts, out)
        if fused is not None:
            return fused
        lo, hi = self.coords[0:1, :], self.coords[1:2, :]
        pts = lo + (hi - lo) * pts
        vals = self.func(pts.to(self.device))
        if out is not None:
            out[0].copy_(pts)
            out[1].copy_(vals)
            return out[0], out[1]
        return pts, vals

    def _fused(self, rnd, out=None):
        """CUDA fast path for the two analytic targets of this module: one kernel maps the random
        numbers into the box and evaluates ``u`` / ``r`` (instead of ~15 / ~30 torch launches)."""
        kind = _FUSED_KIND.get(self.func)
        if kind is None or not rnd.is_cuda or self.dim != 3 or rnd.dtype != torch.float32:
            return None
        import ctypes

        from .. import _lib

        if getattr(self, "_lo_hi", None) is None:
            box = self.coords.detach().to("cpu", torch.float32).reshape(-1).tolist()
            self._lo_hi = (ctypes.c_float * 6)(*box)
        lib = _lib.require_cuda()
        if out is not None:
            X, y = out
            if X.shape != rnd.shape or y.numel() != rnd.shape[0] or X.dtype != torch.float32 \
                    or y.dtype != torch.float32 or not (X.is_contiguous() and y.is_contiguous()) \
                    or X.device != rnd.device or y.device != rnd.device:
                raise ValueError("Sampler.sample(out=...): buffers do not match this draw")
        else:
            X = torch.empty_like(rnd)
            y = torch.empty((rnd.shape[0], 1), dtype=torch.float32, device=rnd.device)
        with torch.cuda.device(rnd.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(rnd.device).cuda_stream)
            rc = lib.qcp_sample_targets(
                ctypes.c_void_p(rnd.data_ptr()), rnd.shape[0], self._lo_hi, kind,
                default_D, default_v_x, default_v_y,
                ctypes.c_void_p(X.data_ptr()), ctypes.c_void_p(y.data_ptr()), stream)
        _lib.check(rc, "qcp_sample_targets")
        return X, y


def _pulse(txy):
    dx = txy[:, 1:2] - 0.5
    dy = txy[:, 2:3] - 0.5
    return dx, dy, torch.exp(-100 * (dx ** 2 + dy ** 2)) * torch.exp(-txy[:, 0:1])


def u(txy):
    return _pulse(txy)[2]


def u_t(txy):
    return -u(txy)


def u_x(txy):
    dx, _, val = _pulse(txy)
    return -200 * dx * val


def u_y(txy):
    _, dy, val = _pulse(txy)
    return -200 * dy * val


def u_xx(txy):
    dx, _, val = _pulse(txy)
    return (40000 * dx ** 2 - 400) * val


def u_yy(txy):
    _, dy, val = _pulse(txy)
    return (40000 * dy ** 2 - 400) * val


def r(txy, Diffusion=default_D, v_x=default_v_x, v_y=default_v_y):
    dx, dy, val = _pulse(txy)
    lap = (40000 * dx ** 2 - 400) * val + (40000 * dy ** 2 - 400) * val
    return -val + v_x * (-200 * dx * val) + v_y * (-200 * dy * val) - Diffusion * lap


_FUSED_KIND = {u: 0, r: 1}


def _box(lo, hi, device):
    return torch.tensor([lo, hi], dtype=torch.float32, device=device)


def training_boxes(device):
    """The four boxes of reference trainer/diffusion_train.py:9-20 (t, x, y order)."""
    return {
        "ics": _box([0.0, 0.0, 0.0], [0.0, 1.0, 1.0], device),
        "bc1": _box([0.0, 0.0, 0.0], [1.0, 0.0, 1.0], device),
        "bc2": _box([0.0, 1.0, 0.0], [1.0, 1.0, 1.0], device),
        "dom": _box([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], device),
    }


def generate_training_dataset(device):
    b = training_boxes(device)
    ics_sampler = Sampler(3, b["ics"], u, name="Initial Condition", device=device)
    bc1 = Sampler(3, b["bc1"], u, name="Dirichlet BC1", device=device)
    bc2 = Sampler(3, b["bc2"], u, name="Dirichlet BC2", device=device)
    res_sampler = Sampler(3, b["dom"], r, name="Forcing", device=device)
    return [ics_sampler, [bc1, bc2], res_sampler]
