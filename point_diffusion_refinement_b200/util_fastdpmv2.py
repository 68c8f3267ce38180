"""FastDPM sampling (VAR / STEP, Kong & Ping 2021) behind the reference's
``fast_sampling_function_v2`` surface (reference: pointnet2/util_fastdpmv2.py:186-476).

Schedule construction (bisection for the variance schedule, continuous-time mapping through the
log-Gamma approximation) is host-side numpy exactly as in the reference; the per-step tensor update
``x <- x*a + c*eps + sigma*z`` is one fused device kernel with in-kernel Philox noise.
"""
import numpy as np
import torch

from .util import DeviceNoise, calc_diffusion_hyperparams, chain_rng  # noqa: F401


def bisearch(f, domain, target, eps=1e-8):
    """Bisection used by the reference for monotonically DEcreasing f (util_fastdpmv2.py:186-209):
    moves right when f(x) < target, left when f(x) > (1 +- eps) target."""
    sign = -1 if target < 0 else 1
    left, right = domain
    x = (left + right) / 2
    for _ in range(1000):
        x = (left + right) / 2
        fx = f(x)
        if fx < target:
            right = x
        elif fx > (1 + sign * eps) * target:
            left = x
        else:
            break
    return x


def get_VAR_noise(S, diffusion_config, schedule="linear"):
    """S noise levels eta whose prod(1-eta) matches the full schedule's alpha_bar_T (:212-236)."""
    b0, bT, T = diffusion_config["beta_0"], diffusion_config["beta_T"], diffusion_config["T"]
    target = np.prod(1 - np.linspace(b0, bT, T))
    if schedule == "linear":
        g = lambda x: np.linspace(b0, x, S)
        domain = (b0, 0.99)
    elif schedule == "quadratic":
        g = lambda x: np.array([b0 * (1 + i * x) ** 2 for i in range(S)])
        domain = (0.0, 0.95 / np.sqrt(b0) / S)
    else:
        raise NotImplementedError
    f = lambda x: np.prod(1 - g(x))
    return g(bisearch(f, domain, target, eps=1e-4))


def get_STEP_step(S, diffusion_config, schedule="linear"):
    """S discrete steps out of T (:239-259)."""
    T = diffusion_config["T"]
    if schedule == "linear":
        c = (T - 1.0) / (S - 1.0)
        taus = [np.floor(i * c) for i in range(S)]
    elif schedule == "quadratic":
        taus = np.linspace(0, np.sqrt(T * 0.8), S) ** 2
    else:
        raise NotImplementedError
    return [int(s) for s in taus]


def _log_gamma(x):
    y = x - 1  # Stirling with the 1/(12y) correction (:261-264)
    return np.log(2 * np.pi * y) / 2 + y * (np.log(y) - 1) + np.log(1 + 1 / (12 * y))


def _log_cont_noise(t, beta_0, beta_T, T):
    """log alpha_bar extended to continuous t (:267-272)."""
    delta_beta = (beta_T - beta_0) / (T - 1)
    _c = (1.0 - beta_0) / delta_beta
    t_1 = t + 1
    return t_1 * np.log(delta_beta) + _log_gamma(_c + 1) - _log_gamma(_c - t_1 + 1)


def _gamma_bar(user_defined_eta):
    g = (1 - torch.from_numpy(np.asarray(user_defined_eta)).to(torch.float32)).clone()
    for t in range(1, len(g)):
        g[t] *= g[t - 1]
    return g


def _precompute_VAR_steps(diffusion_hyperparams, user_defined_eta):
    """Continuous step tau for each of the user noise levels, from the last to the first (:275-304)."""
    _dh = diffusion_hyperparams
    T = _dh["T"]
    Alpha_bar = _dh["Alpha_bar"].detach().float().cpu()
    Beta = _dh["Beta"].detach().float().cpu()
    assert len(Alpha_bar) == T
    Gamma_bar = _gamma_bar(user_defined_eta)
    T_user = len(Gamma_bar)
    assert Gamma_bar[0] <= Alpha_bar[0] and Gamma_bar[-1] >= Alpha_bar[-1]
    # The reference passes 0-d fp32 arrays here (Beta[0].cpu().numpy(), :295-297), so its Stirling formula
    # runs in whatever precision the installed numpy promotes fp32-with-python-float to: a fp32/fp64
    # mixture under the numpy 1.x it was written for, pure fp32 under numpy >= 2 -- where the result is so
    # noisy (+-0.9 step, last tau 0.497) that the reference's own `assert abs(tau) < 0.1` (:353) fails.
    # We evaluate it in float64 on the fp32 schedule endpoints, which reproduces the reference run with
    # a float64 `Beta` (tests/golden/schedules.pt: "taus_f64") and always satisfies that assert.
    b0, bT = float(Beta[0]), float(Beta[-1])
    ab = Alpha_bar.numpy()
    steps = []
    for t in range(T_user - 1, -1, -1):
        gb = Gamma_bar[t].numpy()
        t_adapted = None
        hits = np.nonzero((ab[:-1] >= gb) & (gb > ab[1:]))[0]
        if len(hits):
            i = int(hits[0])
            t_adapted = bisearch(f=lambda _t: _log_cont_noise(_t, b0, bT, T), domain=(i - 0.01, i + 1.01),
                                 target=float(np.log(gb)))
        if t_adapted is None:
            t_adapted = T - 1
        steps.append(t_adapted)
    return steps


def _run_chain(net, size, taus, coef_fn, label, verbose, condition, noise, seed, device, noise_stream=0):
    if device is None:
        device = condition.device if condition is not None else torch.device("cuda", torch.cuda.current_device())
    rng = chain_rng(seed, noise_stream)
    n_steps = len(taus)
    draw = (lambda i: noise(i, size).to(device=device, dtype=torch.float32)) if noise is not None else None
    x = (draw(-1) if draw else rng.normal(size, device)).contiguous()
    if label is not None and isinstance(label, int):
        label = torch.full((size[0],), label, dtype=torch.long, device=device)
    ts = torch.empty((size[0],), dtype=torch.float32, device=device)
    with torch.no_grad():
        for i, tau in enumerate(taus):
            if verbose:
                print("t %.2f x max %.2f min %.2f" % (tau, x.max(), x.min()))
            ts.fill_(float(tau))
            if condition is None:
                eps = net(x, ts=ts, label=label)
            else:
                eps = net(x, condition, ts=ts, label=label, use_retained_condition_feature=True)
            scale_x, c, sigma = coef_fn(i, tau, i == n_steps - 1)
            z = draw(i) if (draw and sigma != 0.0) else None
            rng.affine_update(x, eps.contiguous(), scale_x, c, sigma, noise=z)
    if condition is not None and hasattr(net, "reset_cond_features"):
        net.reset_cond_features()
    return x


def _ddim_coefficients(a_cur, a_next, kappa, last):
    """x *= sqrt(a_next/a_cur); x += c*eps + sigma*z  (:353-373), in fp32 like the reference."""
    f = np.float32
    a_cur = f(a_cur)
    if last:
        a_next, sigma = f(1.0), f(0.0)
    else:
        a_next = f(a_next)
        sigma = f(kappa) * np.sqrt((f(1) - a_next) / (f(1) - a_cur) * (f(1) - a_cur / a_next))
    scale_x = np.sqrt(a_next / a_cur)
    c = np.sqrt(f(1) - a_next - sigma ** 2) - np.sqrt(f(1) - a_cur) * scale_x
    return float(scale_x), float(c), float(sigma)


def VAR_sampling(net, size, diffusion_hyperparams, user_defined_eta, kappa, continuous_steps,
                 print_every_n_steps=100, label=0, verbose=True, condition=None, noise=None, seed=None,
                 device=None, noise_stream=0):
    """:307-381."""
    _dh = diffusion_hyperparams
    T = _dh["T"]
    Alpha_bar = _dh["Alpha_bar"].detach().float().cpu()
    assert len(_dh["Alpha"]) == T and len(Alpha_bar) == T and len(_dh["Sigma"]) == T and len(size) == 3
    assert 0.0 <= kappa <= 1.0
    Gamma_bar = _gamma_bar(user_defined_eta).numpy()
    T_user = len(Gamma_bar)
    assert Gamma_bar[0] <= Alpha_bar[0] and Gamma_bar[-1] >= Alpha_bar[-1]
    print("begin sampling, total number of reverse steps = %s" % T_user)

    def coef(i, tau, last):
        if last:
            assert abs(tau) < 0.1
        cur = Gamma_bar[T_user - 1 - i]
        nxt = None if last else Gamma_bar[T_user - 1 - i - 1]
        return _ddim_coefficients(cur, nxt, kappa, last)

    return _run_chain(net, size, list(continuous_steps), coef, label, verbose, condition, noise, seed, device,
                      noise_stream)


def STEP_sampling(net, size, diffusion_hyperparams, user_defined_steps, kappa, print_every_n_steps=100, label=0,
                  verbose=True, condition=None, noise=None, seed=None, device=None, noise_stream=0):
    """:384-452."""
    _dh = diffusion_hyperparams
    T = _dh["T"]
    Alpha_bar = _dh["Alpha_bar"].detach().float().cpu().numpy()
    assert len(_dh["Alpha"]) == T and len(Alpha_bar) == T and len(_dh["Sigma"]) == T and len(size) == 3
    assert 0.0 <= kappa <= 1.0
    steps = sorted(list(user_defined_steps), reverse=True)
    print("begin sampling, total number of reverse steps = %s" % len(steps))

    def coef(i, tau, last):
        if last:
            assert tau == 0
        nxt = None if last else Alpha_bar[steps[i + 1]]
        return _ddim_coefficients(Alpha_bar[tau], nxt, kappa, last)

    return _run_chain(net, size, steps, coef, label, verbose, condition, noise, seed, device, noise_stream)


def fast_sampling_function_v2(net, size, diffusion_hyperparams, diffusion_config, length=100, sampling_method="var",
                              schedule="quadratic", kappa=0.0, print_every_n_steps=100, label=0, verbose=True,
                              condition=None, noise=None, seed=None, device=None, noise_stream=0):
    """:455-476."""
    assert sampling_method in ["var", "step"]
    assert schedule in ["quadratic", "linear"]
    extra = dict(print_every_n_steps=print_every_n_steps, label=label, verbose=verbose, condition=condition,
                 noise=noise, seed=seed, device=device, noise_stream=noise_stream)
    if sampling_method == "var":
        eta = get_VAR_noise(length, diffusion_config, schedule)
        taus = _precompute_VAR_steps(diffusion_hyperparams, eta)
        return VAR_sampling(net, size, diffusion_hyperparams, eta, kappa, taus, **extra)
    steps = get_STEP_step(length, diffusion_config, schedule)
    return STEP_sampling(net, size, diffusion_hyperparams, steps, kappa, **extra)
