"""DDPM schedule and the reverse sampling loop (reference: pointnet2/util.py:118-123,154-255).

``sampling`` keeps the reference's signature.  What changed underneath:
  * the Gaussian noise z_t is generated ON THE DEVICE by a counter-based Philox kernel fused with the
    posterior-mean update (one launch per step) -- the reference draws it with the CPU generator and
    copies 24 KB * B to the GPU every step (util.py:118-123,248-249);
  * the step index lives in a device tensor updated in place -- no per-step H2D (reference :229);
  * schedule scalars are read from a host copy -- no device->host sync inside the loop;
  * ``noise`` lets tests inject the exact z sequence of another run (parity mode).
"""
import ctypes

import numpy as np
import torch

from ._lib import call, dptr, stream_ptr


class AverageMeter(object):
    """Running average (reference util.py:7-25)."""

    def __init__(self, name, fmt=":f"):
        self.name, self.fmt = name, fmt
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def __str__(self):
        return ("{name} {val" + self.fmt + "} ({avg" + self.fmt + "})").format(**self.__dict__)


def calc_diffusion_hyperparams(T, beta_0, beta_T):
    """Linear-beta DDPM schedule in fp32, computed by the same recurrences as the reference
    (util.py:154-181): Alpha_bar[t] = Alpha_bar[t-1]*Alpha[t]; Sigma[t]^2 = Beta[t](1-Abar[t-1])/(1-Abar[t])."""
    Beta = torch.linspace(beta_0, beta_T, T)
    Alpha = 1 - Beta
    Alpha_bar = Alpha + 0
    Beta_tilde = Beta + 0
    for t in range(1, T):
        Alpha_bar[t] *= Alpha_bar[t - 1]
        Beta_tilde[t] *= (1 - Alpha_bar[t - 1]) / (1 - Alpha_bar[t])
    Sigma = torch.sqrt(Beta_tilde)
    return {"T": T, "Beta": Beta, "Alpha": Alpha, "Alpha_bar": Alpha_bar, "Sigma": Sigma}


class DeviceNoise:
    """Counter-based N(0,1) stream on the device (Philox4x32-10 + Box-Muller in libpdr_b200).

    ``stream`` selects an independent sub-stream of the same seed (it is folded into the Philox key through a
    splitmix64 round), so that calls, batches and ranks that share a seed do not share noise."""

    def __init__(self, seed=0, stream=0):
        z = (int(seed) + 0x9E3779B97F4A7C15 * int(stream)) & ((1 << 64) - 1)
        if stream:
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & ((1 << 64) - 1)
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & ((1 << 64) - 1)
            z ^= z >> 31
        self.seed = z
        self.offset = 0

    def _advance(self, count):
        off = self.offset
        self.offset += (count + 3) // 4
        return off

    def normal(self, size, device):
        x = torch.empty(size, dtype=torch.float32, device=device)
        with torch.cuda.device(x.device):
            call("pdr_normal_fill", x.numel(), dptr(x), ctypes.c_uint64(self.seed),
                 ctypes.c_uint64(self._advance(x.numel())), stream_ptr(x))
        return x

    def affine_update(self, x, eps, scale_x, scale_eps, sigma, noise=None):
        """x <- x*scale_x + eps*scale_eps + sigma*z in place (z injected if ``noise`` is given)."""
        assert x.is_contiguous() and eps.is_contiguous() and x.dtype == torch.float32 and eps.dtype == torch.float32
        if noise is not None:
            noise = noise.to(device=x.device, dtype=torch.float32).contiguous()
        off = self._advance(x.numel()) if (noise is None and sigma != 0.0) else 0
        with torch.cuda.device(x.device):
            call("pdr_affine_noise_update", x.numel(), dptr(x), dptr(eps), ctypes.c_float(scale_x),
                 ctypes.c_float(scale_eps), ctypes.c_float(sigma), dptr(noise), ctypes.c_uint64(self.seed),
                 ctypes.c_uint64(off), stream_ptr(x))
        return x


def std_normal(size, device="cuda", rng=None):
    """Standard Gaussian tensor on the device (reference util.py:118-123 draws on the CPU)."""
    return (rng or _default_rng()).normal(size, torch.device(device))


_DEFAULT_RNG = None


def _rank():
    d = torch.distributed
    return d.get_rank() if (d.is_available() and d.is_initialized()) else 0


def _default_rng():
    """The process-wide stream every unseeded draw comes from.  Like the reference's global CPU generator
    (util.py:118-123) it ADVANCES between calls, so two unseeded sampling() calls -- two batches of an evaluation,
    the trials of generate_samples.py --num_trials -- never see the same x_T / z_t; unlike it, ranks that were
    seeded alike still draw from different sub-streams.  Re-created when torch.manual_seed changes the seed."""
    global _DEFAULT_RNG
    key = (torch.initial_seed(), _rank())
    if _DEFAULT_RNG is None or _DEFAULT_RNG[0] != key:
        _DEFAULT_RNG = (key, DeviceNoise(seed=key[0], stream=key[1]))
    return _DEFAULT_RNG[1]


def chain_rng(seed, stream=0):
    """Noise source of one sampling chain: the advancing process-wide stream when ``seed`` is None, otherwise the
    reproducible stream (seed, stream)."""
    return _default_rng() if seed is None else DeviceNoise(seed, stream=stream)


def _host_schedule(_dh):
    """fp32 host copies of the schedule (one transfer before the loop instead of syncs inside it)."""
    get = lambda k: _dh[k].detach().float().cpu().numpy().astype(np.float32)
    return get("Alpha"), get("Alpha_bar"), get("Sigma")


def sampling(net, size, diffusion_hyperparams, print_every_n_steps=100, label=0, verbose=True, condition=None,
             return_multiple_t_slices=False, t_slices=[5, 10, 20, 50, 100, 200, 400, 600, 800],
             use_a_precomputed_XT=False, step=100, XT=None, noise=None, seed=None, device=None, noise_stream=0):
    """Ancestral sampling p(x_0|x_T) = prod_t p_theta(x_{t-1}|x_t).  reference util.py:184-255.

    Extra keyword arguments (not in the reference): ``noise`` -- callable ``(t, size) -> tensor`` (t = T for
    the initial x_T) replaying a given noise sequence; ``seed`` -- Philox seed of a reproducible chain (None: the
    process-wide stream, which keeps advancing from call to call like the reference's global generator);
    ``noise_stream`` -- sub-stream of ``seed`` (callers fold the batch index and the rank into it); ``device``.
    """
    _dh = diffusion_hyperparams
    T = _dh["T"]
    Alpha, Alpha_bar, Sigma = _host_schedule(_dh)
    assert len(Alpha) == T and len(Alpha_bar) == T and len(Sigma) == T and len(size) == 3
    if device is None:
        device = condition.device if condition is not None else torch.device("cuda", torch.cuda.current_device())
    rng = chain_rng(seed, noise_stream)
    print("begin sampling, total number of reverse steps = %s" % T)
    result_slices = {}

    def draw(t):
        return noise(t, size).to(device=device, dtype=torch.float32) if noise is not None else rng.normal(size, device)

    if label is not None and isinstance(label, int):
        label = torch.full((size[0],), label, dtype=torch.long, device=device)
    if use_a_precomputed_XT:
        x = (XT.to(device) + float(Sigma[step]) * draw(T)).contiguous()
        start_iter = step - 1
    else:
        x = draw(T).contiguous()
        start_iter = T - 1
    ts = torch.empty((size[0],), dtype=torch.float32, device=device)
    one = np.float32(1.0)
    with torch.no_grad():
        for t in range(start_iter, -1, -1):
            if verbose:
                print("t%d x max %.2f min %.2f" % (t, x.max(), x.min()))
            if print_every_n_steps > 0 and t % print_every_n_steps == 0:
                print("reverse step: %d" % t, flush=True)
            ts.fill_(float(t))
            if condition is None:
                eps = net(x, ts=ts, label=label)
            else:
                eps = net(x, condition, ts=ts, label=label, use_retained_condition_feature=True)
            eps = eps.contiguous()
            inv = one / np.sqrt(Alpha[t])
            c_eps = (one - Alpha[t]) / np.sqrt(one - Alpha_bar[t])
            sigma = float(Sigma[t]) if t > 0 else 0.0
            if return_multiple_t_slices and t in t_slices:
                rng.affine_update(x, eps, float(inv), float(-c_eps * inv), 0.0)
                result_slices[t] = x.clone()  # slices are the posterior means, without noise (util.py:246-247)
                if t > 0:
                    z = noise(t, size) if noise is not None else None
                    rng.affine_update(x, eps, 1.0, 0.0, sigma, noise=z)
            else:
                z = noise(t, size) if (noise is not None and t > 0) else None
                rng.affine_update(x, eps, float(inv), float(-c_eps * inv), sigma, noise=z)
    if condition is not None and hasattr(net, "reset_cond_features"):
        net.reset_cond_features()
    return (x, result_slices) if return_multiple_t_slices else x
