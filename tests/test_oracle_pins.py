"""Pins for the CPU oracle (SURVEY.md §8c "pins the new repo must create").

The reference holds no golden vectors for this path and torchdiffeq/torchsde are absent, so the
oracle is anchored on analytic answers, order conditions, an independent DP5 (scipy), finite
differences, nn.GRUCell and Random123's Philox known-answer vectors.
"""
from fractions import Fraction as Fr

import numpy as np
import pytest
import scipy.integrate
import scipy.linalg
import torch

from oracle import torchdiffeq_restatement as tdq
from oracle import torchsde_restatement as tsde
from oracle.latent_motion import ODEFunc, SDEFunc, sample_z_m_odernn
from oracle.philox import philox4x32_10, normals


class Lin(torch.nn.Module):
    def __init__(self, A):
        super().__init__()
        self.A = A

    def forward(self, t, y):
        return y @ self.A.T


def _expm_traj(A, y0, t):
    return torch.stack([y0 @ torch.tensor(scipy.linalg.expm(A.numpy() * float(tt))).T for tt in t])


# ---- (1) analytic -------------------------------------------------------------------------------

@pytest.mark.parametrize("method,tol", [("rk4", 1e-6), ("dopri5", 1e-8)])
def test_linear_ode_matches_expm(method, tol):
    torch.manual_seed(0)
    A = torch.randn(4, 4, dtype=torch.float64) * 0.5
    y0 = torch.randn(3, 4, dtype=torch.float64)
    t = torch.linspace(0, 1, 16, dtype=torch.float64)
    sol = tdq.odeint(Lin(A), y0, t, method=method, rtol=1e-9, atol=1e-11)
    assert torch.equal(sol[0], y0)
    assert (sol - _expm_traj(A, y0, t)).abs().max() < tol


def test_decreasing_time_is_time_reversal():
    torch.manual_seed(1)
    A = torch.randn(3, 3, dtype=torch.float64) * 0.3
    y0 = torch.randn(2, 3, dtype=torch.float64)
    t = torch.linspace(1, 0, 9, dtype=torch.float64)
    for method in ("rk4", "dopri5"):
        sol = tdq.odeint(Lin(A), y0, t, method=method, rtol=1e-10, atol=1e-12)
        ex = torch.stack([y0 @ torch.tensor(scipy.linalg.expm(A.numpy() * (float(tt) - 1.0))).T for tt in t])
        assert (sol - ex).abs().max() < 1e-6


# ---- (2) order of convergence --------------------------------------------------------------------

def test_rk4_38_rule_is_order_4():
    torch.manual_seed(2)
    A = torch.randn(4, 4, dtype=torch.float64)
    y0 = torch.randn(1, 4, dtype=torch.float64)
    errs = []
    for n in (8, 16, 32, 64):
        t = torch.linspace(0, 1, n + 1, dtype=torch.float64)
        sol = tdq.odeint(Lin(A), y0, t, method="rk4")
        errs.append(float((sol[-1] - _expm_traj(A, y0, t[-1:])[0]).abs().max()))
    slopes = [np.log2(errs[i] / errs[i + 1]) for i in range(3)]
    assert all(3.7 < s < 4.4 for s in slopes), slopes


@pytest.mark.parametrize("method,order", [("euler", 1), ("midpoint", 2)])
def test_euler_and_midpoint_orders(method, order):
    torch.manual_seed(3)
    A = torch.randn(4, 4, dtype=torch.float64)
    y0 = torch.randn(1, 4, dtype=torch.float64)
    errs = []
    for n in (32, 64, 128, 256):
        t = torch.linspace(0, 1, n + 1, dtype=torch.float64)
        sol = tdq.odeint(Lin(A), y0, t, method=method)
        errs.append(float((sol[-1] - _expm_traj(A, y0, t[-1:])[0]).abs().max()))
    slopes = [np.log2(errs[i] / errs[i + 1]) for i in range(3)]
    assert all(order - 0.3 < s < order + 0.4 for s in slopes), slopes


def test_midpoint_evaluates_the_field_at_the_half_step():
    seen = []

    class F(torch.nn.Module):
        def forward(self, t, y):
            seen.append(float(t))
            return -y

    tdq.odeint(F(), torch.ones(1, 1, dtype=torch.float64), torch.tensor([0.0, 0.4], dtype=torch.float64), method="midpoint")
    assert np.allclose(seen, [0.0, 0.2])


def test_rk4_is_the_38_rule_not_classic():
    """One step of y' = y: 3/8 rule and classic RK4 agree to O(h^5) but differ in rounding-free rational
    arithmetic for a nonlinear field; check the stage times 1/3, 2/3 are what the field sees."""
    seen = []

    class F(torch.nn.Module):
        def forward(self, t, y):
            seen.append(float(t))
            return -y * y

    y0 = torch.ones(1, 1, dtype=torch.float64)
    tdq.odeint(F(), y0, torch.tensor([0.0, 0.3], dtype=torch.float64), method="rk4")
    assert np.allclose(seen, [0.0, 0.1, 0.2, 0.3])


def test_dopri5_local_order_5():
    torch.manual_seed(3)
    A = torch.randn(3, 3, dtype=torch.float64)
    y0 = torch.randn(1, 3, dtype=torch.float64)
    errs = []
    for h in (0.2, 0.1, 0.05):
        # a single forced step of size h: first_step=h, tolerances so loose it is accepted, output at t=h
        sol = tdq.odeint(Lin(A), y0, torch.tensor([0.0, h], dtype=torch.float64), method="dopri5",
                         rtol=1e3, atol=1e3, options={"first_step": h})
        assert tdq.last_step_log().accepted == [True]
        errs.append(float((sol[-1] - _expm_traj(A, y0, torch.tensor([h], dtype=torch.float64))[0]).abs().max()))
    slopes = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert all(5.5 < s < 6.5 for s in slopes), slopes  # local error O(h^6)


# ---- (3) Butcher / interpolation identities -------------------------------------------

def test_dopri5_tableau_order_conditions():
    c = [Fr(0), Fr(1, 5), Fr(3, 10), Fr(4, 5), Fr(8, 9), Fr(1), Fr(1)]
    beta = [
        [Fr(1, 5)],
        [Fr(3, 40), Fr(9, 40)],
        [Fr(44, 45), Fr(-56, 15), Fr(32, 9)],
        [Fr(19372, 6561), Fr(-25360, 2187), Fr(64448, 6561), Fr(-212, 729)],
        [Fr(9017, 3168), Fr(-355, 33), Fr(46732, 5247), Fr(49, 176), Fr(-5103, 18656)],
        [Fr(35, 384), Fr(0), Fr(500, 1113), Fr(125, 192), Fr(-2187, 6784), Fr(11, 84)],
    ]
    for i, row in enumerate(beta):
        assert sum(row) == c[i + 1]
        assert np.allclose([float(x) for x in row], tdq.DOPRI5.beta[i].numpy())
    b = beta[-1] + [Fr(0)]
    assert sum(b) == 1
    assert sum(bi * ci for bi, ci in zip(b, c)) == Fr(1, 2)
    assert sum(bi * ci ** 2 for bi, ci in zip(b, c)) == Fr(1, 3)
    assert sum(bi * ci ** 3 for bi, ci in zip(b, c)) == Fr(1, 4)
    assert sum(bi * ci ** 4 for bi, ci in zip(b, c)) == Fr(1, 5)
    bstar = [Fr(1951, 21600), Fr(0), Fr(22642, 50085), Fr(451, 720), Fr(-12231, 42400), Fr(649, 6300), Fr(1, 60)]
    assert sum(bstar) == 1
    for p in (1, 2, 3):
        assert sum(bi * ci ** p for bi, ci in zip(bstar, c)) == Fr(1, p + 1)
    assert sum(bi * ci ** 4 for bi, ci in zip(bstar, c)) != Fr(1, 5)
    cerr = [float(x - y) for x, y in zip(b, bstar)]
    assert np.allclose(cerr, tdq.DOPRI5.c_error.numpy(), atol=1e-16)
    cmid = [Fr(6025192743, 30085553152), Fr(0), Fr(51252292925, 65400821598), Fr(-2691868925, 45128329728),
            Fr(187940372067, 1594534317056), Fr(-1776094331, 19743644256), Fr(11237099, 235043384)]
    cmid = [x / 2 for x in cmid]
    assert np.allclose([float(x) for x in cmid], tdq.DOPRI5.c_mid.numpy())
    assert abs(float(sum(cmid)) - 0.5) < 1e-12
    assert abs(float(sum(m * ci for m, ci in zip(cmid, c))) - 1 / 8) < 1e-12
    assert abs(float(sum(m * ci ** 2 for m, ci in zip(cmid, c))) - 1 / 24) < 1e-12
    assert abs(float(sum(m * ci ** 3 for m, ci in zip(cmid, c))) - 1 / 64) < 1e-12


def test_38_rule_order_conditions():
    c = [Fr(0), Fr(1, 3), Fr(2, 3), Fr(1)]
    b = [Fr(1, 8), Fr(3, 8), Fr(3, 8), Fr(1, 8)]
    a = [[], [Fr(1, 3)], [Fr(-1, 3), Fr(1)], [Fr(1), Fr(-1), Fr(1)]]
    for i in range(1, 4):
        assert sum(a[i]) == c[i]
    for p in range(4):
        assert sum(bi * ci ** p for bi, ci in zip(b, c)) == Fr(1, p + 1)
    assert sum(b[i] * sum(a[i][j] * c[j] for j in range(i)) for i in range(4)) == Fr(1, 6)
    assert sum(b[i] * c[i] * sum(a[i][j] * c[j] for j in range(i)) for i in range(4)) == Fr(1, 8)
    assert sum(b[i] * sum(a[i][j] * c[j] ** 2 for j in range(i)) for i in range(4)) == Fr(1, 12)
    assert sum(b[i] * sum(a[i][j] * sum(a[j][k] * c[k] for k in range(j)) for j in range(i)) for i in range(4)) == Fr(1, 24)


def test_quartic_interpolant_endpoints_and_slopes():
    torch.manual_seed(4)
    y0, y1, ym, f0, f1 = (torch.randn(5, dtype=torch.float64) for _ in range(5))
    dt = torch.tensor(0.37, dtype=torch.float64)
    co = tdq._interp_fit(y0, y1, ym, f0, f1, dt)
    t0, t1 = torch.tensor(1.0, dtype=torch.float64), torch.tensor(1.37, dtype=torch.float64)
    ev = lambda t: tdq._interp_evaluate(co, t0, t1, torch.tensor(t, dtype=torch.float64))
    assert torch.allclose(ev(1.0), y0) and torch.allclose(ev(1.37), y1) and torch.allclose(ev(1.185), ym)
    eps = 1e-7  # one-sided difference: error ~ eps/2 * |p''|, p'' = O(1e2) for random data
    assert torch.allclose((ev(1.0 + eps) - ev(1.0)) / eps, f0, atol=1e-3)
    assert torch.allclose((ev(1.37) - ev(1.37 - eps)) / eps, f1, atol=1e-3)


# ---- (4) independent DP5 ----------------------------------------------------------------------------

def test_dopri5_vs_scipy_rk45_on_odefunc():
    torch.manual_seed(5)
    f = ODEFunc(16, 16).double()
    with torch.no_grad():
        for p in f.parameters():
            p.mul_(3.0)
    y0 = torch.randn(8, 16, dtype=torch.float64)
    t = torch.linspace(0, 1, 16, dtype=torch.float64)
    sol = tdq.odeint(f, y0, t, method="dopri5", rtol=1e-9, atol=1e-11)

    def rhs(tt, y):
        with torch.no_grad():
            return f(None, torch.from_numpy(y).view(8, 16)).reshape(-1).numpy()

    ref = scipy.integrate.solve_ivp(rhs, (0, 1), y0.reshape(-1).numpy(), method="RK45", t_eval=t.numpy(),
                                    rtol=1e-11, atol=1e-13)
    assert np.abs(sol.detach().numpy().reshape(16, -1) - ref.y.T).max() < 1e-6


def _scipy_first_attempt(RK, f, y0, rtol, atol, t_bound=1.0):
    """SciPy's own pieces of an explicit adaptive Runge-Kutta step (scipy.integrate._ivp: select_initial_step, rk_step, the
    method's tableau and error estimator, RMS norm) on the WHOLE batch as one state vector — the batch-global norm torchdiffeq
    applies to a (B, D) state.  Returns (h0, y1, error_norm of the first attempted step)."""
    from scipy.integrate._ivp import common, rk

    B, D = y0.shape

    def fun(tt, y):
        with torch.no_grad():
            return f(None, torch.from_numpy(y).view(B, D)).reshape(-1).numpy()

    y = y0.reshape(-1).numpy().copy()
    f0 = fun(0.0, y)
    h0 = common.select_initial_step(fun, 0.0, y, t_bound, np.inf, f0, 1.0, RK.error_estimator_order, rtol, atol)
    K = np.empty((RK.n_stages + 1, y.size))
    y1, f1 = rk.rk_step(fun, 0.0, y, f0, h0, RK.A, RK.B, RK.C, K)
    scale = atol + np.maximum(np.abs(y), np.abs(y1)) * rtol
    err = common.norm(np.dot(K.T, RK.E) * h0 / scale)
    return h0, y1, err


@pytest.mark.parametrize("method,rk_name,scale", [("dopri5", "RK45", 1.0), ("dopri5", "RK45", 4.0), ("bosh3", "RK23", 2.0)])
def test_adaptive_step_pieces_equal_scipys_independent_implementation(method, rk_name, scale):
    """Algorithmic pin against an independent implementation that IS installed: SciPy's RK45 is the Dormand-Prince 5(4) pair
    and RK23 the Bogacki-Shampine 3(2) pair torchdiffeq calls dopri5 / bosh3, and torchdiffeq's initial-step heuristic is
    SciPy's (Hairer's).  In fp64 the restatement's first step size, the state after the first attempted step and its error
    ratio equal what SciPy's own select_initial_step / rk_step / error estimator give on the same field — tableau, FSAL
    bookkeeping, error weights, tolerance scale and the RMS norm over the whole batch all have to agree for that.
    One documented difference: torchdiffeq's dopri5 is the Dormand-Prince-SHAMPINE tableau, whose error weights
    (35/384 - 1951/21600, ...) are exactly 2/3 of the classical pair's b - b^ (35/384 - 5179/57600 = 71/57600 against 71/86400),
    so its error ratio is 2/3 of SciPy's; bosh3's weights are SciPy's.  (The two CONTROLLERS differ by design: torchdiffeq never
    shrinks an accepted step; the dense outputs differ too.)"""
    from scipy.integrate._ivp import rk

    RK = getattr(rk, rk_name)
    k_err = {"dopri5": 1.5, "bosh3": 1.0}[method]            # SciPy's error estimate / torchdiffeq's
    tab = {"dopri5": tdq.DOPRI5, "bosh3": tdq.BOSH3}[method]
    assert np.allclose(np.abs(tab.c_error.numpy()) * k_err, np.abs(RK.E), rtol=1e-14, atol=0)
    assert np.allclose(tab.c_sol.numpy()[:RK.n_stages], RK.B, rtol=1e-14, atol=1e-18)
    torch.manual_seed(11)
    f = ODEFunc(16, 16).double()
    with torch.no_grad():
        for p in f.parameters():
            p.mul_(scale)
    y0 = torch.randn(24, 16, dtype=torch.float64)
    rtol, atol = 1e-5, 1e-6
    h0, y1, err = _scipy_first_attempt(RK, f, y0, rtol, atol)
    with torch.no_grad():
        # the output time is the end of the first step: the dense output at x = 1 is y1 itself
        sol = tdq.odeint(f, y0, torch.tensor([0.0, h0], dtype=torch.float64), method=method, rtol=rtol, atol=atol)
    log = tdq.last_step_log()
    assert abs(log.dt0 - h0) <= 1e-12 * h0, (log.dt0, h0)
    assert abs(log.dt[0] - h0) <= 1e-12 * h0
    # (after the heuristic's tiny first step the error estimate is ~1e-13 absolute: its own fp64 round-off is ~1e-7 of it)
    assert abs(k_err * log.error_ratio[0] - err) <= 1e-6 * err, (log.error_ratio[0], err)
    assert log.accepted[0] and np.abs(sol[1].numpy().reshape(-1) - y1).max() <= 1e-13 * np.abs(y1).max()
    # ... and a forced over-long first step: same rejected-attempt error ratio
    from scipy.integrate._ivp import common
    B, D = y0.shape

    def fun(tt, y):
        with torch.no_grad():
            return f(None, torch.from_numpy(y).view(B, D)).reshape(-1).numpy()

    y = y0.reshape(-1).numpy().copy()
    K = np.empty((RK.n_stages + 1, y.size))
    big = 2.0
    yb, _ = rk.rk_step(fun, 0.0, y, fun(0.0, y), big, RK.A, RK.B, RK.C, K)
    errb = common.norm(np.dot(K.T, RK.E) * big / (atol + np.maximum(np.abs(y), np.abs(yb)) * rtol))
    with torch.no_grad():
        tdq.odeint(f, y0, torch.tensor([0.0, 1.0], dtype=torch.float64), method=method, rtol=rtol, atol=atol,
                   options={"first_step": big})
    log = tdq.last_step_log()
    assert abs(k_err * log.error_ratio[0] - errb) <= 1e-9 * errb and log.accepted[0] == (errb <= k_err)


def test_dopri5_step_sequence_is_independent_of_output_times():
    torch.manual_seed(6)
    f = ODEFunc(16, 16)
    y0 = torch.randn(32, 16)
    # The controller (safety 0.9) keeps a smooth tanh field at error_ratio ~ 0.5, so rejections are forced
    # with an over-long first step on 8x weights.
    opts = {"first_step": 1.0}
    with torch.no_grad():
        for p in f.parameters():
            p.mul_(8.0)
        tdq.odeint(f, y0, torch.linspace(0, 1, 16), method="dopri5", rtol=1e-5, atol=1e-5, options=opts)
        a = tdq.last_step_log()
        tdq.odeint(f, y0, torch.tensor([0.0, 1.0]), method="dopri5", rtol=1e-5, atol=1e-5, options=opts)
        b = tdq.last_step_log()
    assert a.accepted == b.accepted and a.dt == b.dt
    assert a.n_rejected > 0 and not a.accepted[0]
    assert a.nfe == 1 + 6 * len(a.accepted)


# ---- (5) adjoint vs autograd-through-solver vs finite differences -----------------------------------

def _loss(solve, f, y0, t, g, **kw):
    return (solve(f, y0, t, **kw) * g).sum()


@pytest.mark.parametrize("method,kw", [("rk4", {}), ("dopri5", dict(rtol=1e-9, atol=1e-11))])
def test_gradients_adjoint_autograd_fd(method, kw):
    torch.manual_seed(7)
    f = ODEFunc(6, 5).double()
    y0 = torch.randn(4, 6, dtype=torch.float64, requires_grad=True)
    t = torch.linspace(0, 1, 6 if method == "rk4" else 4, dtype=torch.float64)
    if method == "rk4":
        t = torch.linspace(0, 1, 41, dtype=torch.float64)  # fine grid so the continuous adjoint ~ discrete gradient
    g = torch.randn(len(t), 4, 6, dtype=torch.float64)
    params = list(f.parameters())

    ga = torch.autograd.grad(_loss(tdq.odeint_adjoint, f, y0, t, g, method=method, **kw), [y0] + params)
    gb = torch.autograd.grad(_loss(tdq.odeint, f, y0, t, g, method=method, **kw), [y0] + params)
    for a, b in zip(ga, gb):
        assert (a - b).abs().max() < 2e-5 * max(1.0, float(b.abs().max()))

    # fp64 central finite differences on a few coordinates against autograd-through-solver
    eps = 1e-6
    with torch.no_grad():
        for tensor, grad in zip([y0] + params, gb):
            flat = tensor.view(-1)
            for idx in (0, flat.numel() // 2, flat.numel() - 1):
                old = flat[idx].item()
                flat[idx] = old + eps
                lp = _loss(tdq.odeint, f, y0, t, g, method=method, **kw)
                flat[idx] = old - eps
                lm = _loss(tdq.odeint, f, y0, t, g, method=method, **kw)
                flat[idx] = old
                fd = float(lp - lm) / (2 * eps)
                assert abs(fd - float(grad.view(-1)[idx])) < 1e-5 * max(1.0, abs(fd))


def test_rk4_adjoint_is_one_38_step_per_interval_on_augmented_state():
    """A.4: with method='rk4' every backward interval is ONE 3/8-rule step of the augmented system, so the
    adjoint's parameter gradient equals this closed-form restatement (A.4 last bullet)."""
    torch.manual_seed(8)
    f = ODEFunc(16, 16)
    y0 = torch.randn(32, 16, requires_grad=True)
    t = torch.linspace(0, 1, 16)
    g = torch.randn(16, 32, 16)
    sol = tdq.odeint_adjoint(f, y0, t, method="rk4")
    grads = torch.autograd.grad((sol * g).sum(), [y0] + list(f.parameters()))

    W1, b1, W2, b2 = [p.detach() for p in f.parameters()]

    def F(y, a):
        h = torch.tanh(y @ W1.T + b1)
        fe = h @ W2.T + b2
        gh = a @ W2
        delta = gh * (1 - h * h)
        return -fe, delta @ W1, (delta.T @ y, delta.sum(0), a.T @ h, a.sum(0))

    sol = sol.detach()
    a = g[-1].clone()
    acc = [torch.zeros_like(p) for p in (W1, b1, W2, b2)]
    for i in range(15, 0, -1):
        dt = (-t[i - 1]) - (-t[i])
        y = sol[i]
        k1y, k1a, k1p = F(y, a)
        k2y, k2a, k2p = F(y + dt * k1y / 3, a + dt * k1a / 3)
        k3y, k3a, k3p = F(y + dt * (k2y - k1y / 3), a + dt * (k2a - k1a / 3))
        k4y, k4a, k4p = F(y + dt * (k1y - k2y + k3y), a + dt * (k1a - k2a + k3a))
        a = a + (k1a + 3 * (k2a + k3a) + k4a) * dt * 0.125 + g[i - 1]
        acc = [A + (p1 + 3 * (p2 + p3) + p4) * dt * 0.125 for A, p1, p2, p3, p4 in zip(acc, k1p, k2p, k3p, k4p)]
    for mine, ref in zip([a] + acc, grads):
        assert (mine - ref).abs().max() <= 2e-5 * float(ref.abs().max())


# ---- (6) GRU jump ------------------------------------------------------------------------------------

def test_odernn_loop_shapes_and_gru_formula():
    torch.manual_seed(9)
    f = ODEFunc(16, 16)
    gru = torch.nn.GRUCell(16, 16)
    h0 = torch.randn(3, 16)
    eps = torch.randn(4, 3, 16)
    with torch.no_grad():
        z = sample_z_m_odernn(f, gru, h0, eps, adjoint=False, rtol=1e-5, atol=1e-5)
        assert z.shape == (12, 16)
        hp = tdq.odeint(f, h0, torch.tensor([0.0, 1.0]), rtol=1e-5, atol=1e-5)[-1]
        Wi, Wh, bi, bh = gru.weight_ih, gru.weight_hh, gru.bias_ih, gru.bias_hh
        gi = eps[0] @ Wi.T + bi
        gh = hp @ Wh.T + bh
        r = torch.sigmoid(gi[:, :16] + gh[:, :16])
        zz = torch.sigmoid(gi[:, 16:32] + gh[:, 16:32])
        n = torch.tanh(gi[:, 32:] + r * gh[:, 32:])
        h1 = (1 - zz) * n + zz * hp
    assert torch.allclose(z.view(3, 4, 16)[:, 0], h1, atol=1e-6)


# ---- (7) Philox + SDE grid -----------------------------------------------------------------------------

def test_philox4x32_10_random123_known_answers():
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, out in kat:
        r = philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert [int(x) for x in r] == out


def test_philox_normals_moments():
    z = normals(1234, np.arange(200000), step=3, d_block=1)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    assert abs((z ** 4).mean() - 3) < 0.05


def test_sde_euler_grid_is_41_steps_with_fp32_time():
    ts = torch.linspace(0, 1, 16).float()
    pairs = tsde.step_grid(ts, 2.5e-2)
    assert len(pairs) == 41  # SURVEY Appendix B (measured upstream behaviour: fp32 accumulation)
    assert float(pairs[-1][1]) == 1.0 and float(pairs[-1][1] - pairs[-1][0]) < 1e-6


def test_sde_euler_given_increments_matches_hand_loop():
    torch.manual_seed(10)
    sde = SDEFunc(16, 16)
    y0 = torch.randn(5, 16)
    ts = torch.linspace(0, 1, 16).float()
    pairs = tsde.step_grid(ts, 2.5e-2)
    dW = torch.stack([torch.randn(5, 16) * float(b - a) ** 0.5 for a, b in pairs])
    with torch.no_grad():
        sol = tsde.sdeint(sde, y0, ts, bm=tsde.TableBrownian(dW), method="euler", dt=2.5e-2)
        y = y0
        for k, (a, b) in enumerate(pairs[:2]):
            y = y + sde.f(a, y) * (b - a) + sde.g(a, y) * dW[k]
    assert sol.shape == (16, 5, 16) and torch.equal(sol[0], y0)
    # frame 1 (t=1/15) lies between step 2 (t=.05) and step 3 (t=.075)
    with torch.no_grad():
        y2 = y
        a, b = pairs[2]
        y3 = y2 + sde.f(a, y2) * (b - a) + sde.g(a, y2) * dW[2]
        w = (ts[1] - a) / (b - a)
    assert torch.allclose(sol[1], (1 - w) * y2 + w * y3, atol=1e-6)


# ---- torchsde's stochastic adjoint (f3) ---------------------------------------------------------------------------------
def _adjoint_step_closed_form(params, y, a, v):
    """One evaluation of the augmented adjoint field in closed form for the two tanh MLPs — the arithmetic
    csrc/sde_small.cu::sde_adjoint_bwd_kernel performs (its header derives it)."""
    fW1, fb1, fW2, fb2, gW1, gb1, gW2, gb2 = params
    hf = torch.tanh(y @ fW1.T + fb1)
    f = hf @ fW2.T + fb2
    df = (a @ fW2) * (1 - hf * hf)
    z = y @ gW1.T + gb1
    h = torch.tanh(z)
    s = 1 - h * h
    g = h @ gW2.T + gb2
    q, p = g @ gW2, a @ gW1.T
    gdg = (s * q) @ gW1
    a_dg = (s * (a @ gW2)) @ gW1
    r = p * s
    gbar = r @ gW2.T
    hbar = gbar @ gW2 - 2 * h * p * q
    zbar = s * hbar

    def vjp_g(c):
        d = (c @ gW2) * s
        return d @ gW1, d.T @ y, d.sum(0), c.T @ h, c.sum(0)

    ey, eW1, eb1, eW2, eb2 = vjp_g(a_dg)
    wy, wW1, wb1, wW2, wb2 = vjp_g(a * v)
    f_o = [-(f - gdg), df @ fW1 - zbar @ gW1 + ey, df.T @ y, df.sum(0), a.T @ hf, a.sum(0),
           -(zbar.T @ y + (s * q).T @ a) + eW1, -zbar.sum(0) + eb1, -(g.T @ r + gbar.T @ h) + eW2, -gbar.sum(0) + eb2]
    g_o = [-(g * v), wy, None, None, None, None, wW1, wb1, wW2, wb2]
    return f_o, g_o


def test_sde_adjoint_field_closed_form_equals_autograd_restatement():
    from oracle import torchsde_restatement as tsde
    from oracle.latent_motion import SDEFunc
    torch.manual_seed(0)
    sde = SDEFunc(16, 16).double()
    params = list(sde.parameters())
    y, a = torch.randn(5, 16, dtype=torch.float64), torch.randn(5, 16, dtype=torch.float64)
    v = 0.15 * torch.randn(5, 16, dtype=torch.float64)
    f_out, g_out = tsde.adjoint_f_and_g_prod(sde, params, torch.tensor(-0.5), y, a, v)
    f_c, g_c = _adjoint_step_closed_form([p.detach() for p in params], y, a, v)
    for x, ref in zip(f_c, f_out):
        assert float((x - ref).abs().max()) <= 1e-13
    for x, ref in zip(g_c, g_out):
        assert float(((0 * ref if x is None else x) - ref).abs().max()) <= 1e-13


def test_sde_adjoint_grids_of_the_reference_call():
    """models/mocogan_sde.py:57-59: 41 forward steps, 45 reverse steps (3 per output interval: 0.025, 0.025, 0.01667), and
    the union grid both sides are measured on."""
    from oracle import torchsde_restatement as tsde
    ts = torch.linspace(0, 1, 16).float()
    rev = tsde.reverse_step_grid(ts, 2.5e-2)
    assert [len(p) for _, p in rev] == [3] * 15
    hs = [float(s1 - s0) for s0, s1 in rev[0][1]]
    assert abs(hs[0] - 0.025) < 1e-6 and abs(hs[1] - 0.025) < 1e-6 and abs(hs[2] - (1 / 15 - 0.05)) < 1e-6
    grid = tsde.adjoint_time_grid(ts, 2.5e-2)
    assert len(grid) == 80 and float(grid[0]) == 0.0 and float(grid[-1]) == 1.0


def test_sde_stochastic_adjoint_converges_to_discrete_gradient_for_diagonal_jacobian():
    """Pin of the restatement's signs and correction terms: for a diffusion whose Jacobian is diagonal (what torchsde's
    'diagonal' formulas assume) the stochastic adjoint converges to the pathwise (discrete) gradient as dt -> 0."""
    from oracle import torchsde_restatement as tsde

    class DiagSDE(torch.nn.Module):
        noise_type, sde_type = "diagonal", "ito"

        def __init__(self):
            super().__init__()
            self.drift = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 4))
            self.w, self.b = torch.nn.Parameter(torch.randn(4)), torch.nn.Parameter(torch.randn(4))

        def f(self, t, y):
            return self.drift(y)

        def g(self, t, y):
            return 0.7 * torch.tanh(self.w * y + self.b) + 0.3

    torch.manual_seed(0)
    ts = torch.linspace(0, 1, 3).float()
    B = 1500
    sde = DiagSDE().double()
    y0, up = torch.randn(B, 4, dtype=torch.float64), torch.randn(3, B, 4, dtype=torch.float64)
    errs = []
    for dt in (0.1, 0.00625):
        grid = tsde.adjoint_time_grid(ts, dt)
        gen = torch.Generator().manual_seed(1)
        inc = torch.randn(len(grid) - 1, B, 4, dtype=torch.float64, generator=gen) * torch.diff(grid).sqrt().view(-1, 1, 1)
        bm = tsde.GridBrownian(grid, inc)

        def run(fn):
            y = y0.clone().requires_grad_(True)
            sol = fn(sde, y, ts, bm=bm, method="euler", dt=dt)
            return torch.autograd.grad((sol * up).sum(), [y] + list(sde.parameters()))

        g_disc, g_adj = run(tsde.sdeint), run(tsde.sdeint_adjoint)
        errs.append(max(float((a - b).norm() / b.norm()) for a, b in zip(g_adj, g_disc)))
    assert errs[1] < 0.06 and errs[1] < 0.5 * errs[0], errs   # O(sqrt(dt)) strong error of the Euler adjoint: 16x smaller dt
