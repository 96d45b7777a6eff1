"""Property tests (hypothesis; SURVEY §4 "Property tests"): random batch sizes, field shapes, grid lengths, non-uniform and
decreasing output grids.  CPU part: invariants of the oracle and of the host-side logic.  GPU part: the fixed-grid kernels
against the oracle on whatever shape hypothesis draws.  Examples are derandomised: the same cases run every time."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import gan_ode_b200 as gode
from gan_ode_b200 import dist as gdist
from oracle import torchdiffeq_restatement as tdq
from tests.helpers import clone_to, make_field, rel_err

CPU = settings(derandomize=True, max_examples=25, deadline=None)
DEV = "cuda"


def _grid(data, T, decreasing):
    """A strictly monotone, non-uniform grid of T points on [0, 1]."""
    gaps = data.draw(st.lists(st.floats(0.05, 1.0), min_size=T - 1, max_size=T - 1))
    t = torch.tensor([0.0] + list(np.cumsum(gaps) / sum(gaps)), dtype=torch.float32)
    t[-1] = 1.0
    return t.flip(0).contiguous() if decreasing else t


# ---- CPU: oracle invariants -----------------------------------------------------------------------------------------------
@CPU
@given(B=st.integers(1, 9), D=st.sampled_from([2, 5, 16]), H=st.sampled_from([3, 16]), T=st.integers(2, 9), data=st.data())
def test_oracle_decreasing_grid_is_the_reversed_field_on_the_negated_grid(B, D, H, T, data):
    """misc.py::_check_inputs: a decreasing t is solved as increasing -t with the field negated — for every fixed-grid
    method and for dopri5, on any non-uniform grid."""
    f = make_field(D, H, seed=B + T).double()
    t = _grid(data, T, decreasing=True).double()
    y0 = torch.randn(B, D, dtype=torch.float64)

    class Neg(torch.nn.Module):
        def forward(self, s, x):
            return -f(-s, x)

    for method, kw in (("rk4", {}), ("midpoint", {}), ("dopri5", dict(rtol=1e-8, atol=1e-10))):
        with torch.no_grad():
            a = tdq.odeint(f, y0, t, method=method, **kw)
            b = tdq.odeint(Neg(), y0, -t, method=method, **kw)
        assert torch.allclose(a, b, rtol=1e-12, atol=1e-14), method


@CPU
@given(B=st.integers(1, 6), D=st.sampled_from([3, 16]), T=st.integers(2, 7), data=st.data())
def test_oracle_rk4_adjoint_equals_autograd_through_the_solver_as_the_grid_refines(B, D, T, data):
    """adjoint.py vs plain autograd: both are gradients of the same map up to the discretisation error of one extra 3/8
    step per interval, i.e. O(h^4) — halving every interval must shrink their difference by ~16."""
    f = make_field(D, 7, seed=D + T, scale=1.5).double()
    t1 = _grid(data, T, decreasing=False).double()
    t2 = torch.sort(torch.cat([t1, (t1[1:] + t1[:-1]) / 2])).values
    y0 = torch.randn(B, D, dtype=torch.float64)
    g = torch.randn(B, D, dtype=torch.float64)

    def gap(t):
        out = []
        for solve in (tdq.odeint_adjoint, tdq.odeint):
            y = y0.clone().requires_grad_(True)
            sol = solve(f, y, t, method="rk4")
            out.append(torch.autograd.grad((sol[-1] * g).sum(), [y] + list(f.parameters())))
        return max(float((a - b).abs().max()) for a, b in zip(*out))

    d1, d2 = gap(t1), gap(t2)
    assert d2 <= d1 / 6 + 1e-13, (d1, d2)


@CPU
@given(B=st.integers(1, 40), T=st.integers(2, 12), data=st.data())
def test_oracle_batch_rows_are_independent_under_fixed_grid_methods(B, T, data):
    """Fixed-grid solves have no coupling between trajectories: solving a batch equals solving its rows (this is what makes
    the batch shard across GPUs without a data-path collective)."""
    f = make_field(16, 16, seed=T)
    t = _grid(data, T, decreasing=data.draw(st.booleans()))
    y0 = torch.randn(B, 16)
    cut = data.draw(st.integers(0, B))
    with torch.no_grad():
        whole = tdq.odeint(f, y0, t, method="rk4")
        parts = [tdq.odeint(f, y0[a:b], t, method="rk4") for a, b in ((0, cut), (cut, B)) if b > a]
    assert torch.allclose(whole, torch.cat(parts, 1), rtol=1e-6, atol=1e-7)


# ---- CPU: host-side logic ---------------------------------------------------------------------------------------------------
@CPU
@given(B=st.integers(0, 5000), world=st.integers(1, 64))
def test_shard_bounds_partition_any_batch(B, world):
    cuts = [gdist.shard_bounds(B, r, world) for r in range(world)]
    assert cuts[0][0] == 0 and cuts[-1][1] == B
    assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
    sizes = [hi - lo for lo, hi in cuts]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@CPU
@given(T=st.integers(2, 40), decreasing=st.booleans(), dtype=st.sampled_from([torch.float32, torch.float64]), data=st.data())
def test_host_step_table_matches_torch_arithmetic_on_any_monotone_grid(T, decreasing, dtype, data):
    """odeint._host_steps: the dt table that rides in the launch parameters is t[j+1]-t[j] in t's dtype rounded to fp32
    (what multiplies the fp32 state in torchdiffeq's fixed-grid driver), and the fp64 grid handed to dopri5 is increasing."""
    from gan_ode_b200 import odeint as api_mod  # noqa: F401  (the function object; the module is reached through it)
    import importlib
    api = importlib.import_module("gan_ode_b200.odeint")
    t = _grid(data, T, decreasing).to(dtype)
    t64, dt32, fsign = api._host_steps(t)
    assert fsign == (-1.0 if decreasing else 1.0)
    assert np.all(np.diff(t64) > 0) and t64.dtype == np.float64
    expect = (t[1:] - t[:-1]).to(torch.float32).numpy()
    assert dt32.dtype == np.float32 and np.array_equal(dt32, expect)


# ---- GPU: the fixed-grid kernels on drawn shapes ------------------------------------------------------------------------------
def _need_gpu():
    assert torch.cuda.is_available(), "GPU tests need a B200 (no CPU fallback on the product path)"


@pytest.mark.gpu
@settings(derandomize=True, max_examples=16, deadline=None)
@given(B=st.integers(1, 200), shape=st.sampled_from([(16, 16), (16, 16), (32, 32), (32, 64), (64, 256)]), T=st.integers(2, 32),
       decreasing=st.booleans(), layout=st.sampled_from(["tbd", "btd"]), adjoint=st.booleans(), data=st.data())
def test_rk4_kernels_match_the_oracle_on_drawn_shapes(B, shape, T, decreasing, layout, adjoint, data):
    _need_gpu()
    D, H = shape
    if (D, H) != (16, 16):
        adjoint = True    # backprop-through-solver exists for the reference shape only; the reference always uses the adjoint
    f = make_field(D, H, seed=B + T)
    t = _grid(data, T, decreasing)
    torch.manual_seed(B * 31 + T)
    y0, g = torch.randn(B, D), torch.randn(T, B, D)

    def run(mod, field, y, gg, **kw):
        y = y.clone().requires_grad_(True)
        sol = (mod.odeint_adjoint if adjoint else mod.odeint)(field, y, t, method="rk4", **kw)
        if sol.shape[0] != T:          # (B, T, D) view handed back for layout='btd'
            sol = sol.transpose(0, 1)
        return sol.detach(), torch.autograd.grad((sol * gg).sum(), [y] + list(field.parameters()))

    ref_sol, ref_g = run(tdq, f, y0, g)
    out_sol, out_g = run(gode, clone_to(f, DEV), y0.to(DEV), g.to(DEV), options={"layout": layout})
    errs = [rel_err(out_sol, ref_sol)] + [rel_err(a, b) for a, b in zip(out_g, ref_g)]
    import os
    if os.environ.get("GODE_TEST_VERBOSE"):
        print("B=%d D=%d H=%d T=%d dec=%s %s adj=%s" % (B, D, H, T, decreasing, layout, adjoint), ["%.1e" % e for e in errs])
    assert errs[0] <= 1e-5, errs
    # gradients: the ORACLE is fp32 here too (its own rounding over up to 31 steps x 200 trajectories is in the difference;
    # the fixed-size cases in test_gpu_parity.py separate the two with an fp64 oracle)
    assert max(errs[1:]) <= 5e-5, errs
