"""the data-parallel training harness (scripts/train_dp_harness.py).

CPU (needs /root/reference): the harness builds the reference's OWN nets (models/mocogan_ode.py::VideoGenerator with
dim_hidden passed, models/mocogan.py::VideoDiscriminator / PatchImageDiscriminator) and runs one iteration of the
reference's loop (ucf_moco_ode.py:113-163) with them — the solver behind `import torchdiffeq` is the CPU oracle here; and
the stand-in nets used where the reference tree is absent have exactly the reference's state_dict keys and shapes.
GPU: two iterations of the same loop through the gan_ode_b200 shim (tcgen05 forward + adjoint, D=64 / H=256)."""
import importlib.util
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
spec = importlib.util.spec_from_file_location("train_dp_harness", os.path.join(ROOT, "scripts", "train_dp_harness.py"))
harness = importlib.util.module_from_spec(spec)
spec.loader.exec_module(harness)

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not mounted")


def _oracle_shim():
    from oracle import torchdiffeq_restatement as tdq
    m = types.ModuleType("torchdiffeq")
    m.odeint, m.odeint_adjoint = tdq.odeint, tdq.odeint_adjoint
    return m


@needs_ref
def test_stand_in_nets_have_the_reference_nets_keys_and_shapes(monkeypatch):
    monkeypatch.setitem(sys.modules, "torchdiffeq", _oracle_shim())
    ref = harness.build_nets(use_reference=True)
    mine = harness.build_nets(use_reference=False)
    assert "reference" in ref[3] and "stand-ins" in mine[3]
    for r, m in zip(ref[:3], mine[:3]):
        rs, ms = r.state_dict(), m.state_dict()
        assert list(rs.keys()) == list(ms.keys()), (type(r).__name__, set(rs) ^ set(ms))
        for k in rs:
            assert rs[k].shape == ms[k].shape, (type(r).__name__, k)
    # same outputs for the same weights and noise: the samplers are restated call for call
    mine[0].load_state_dict(ref[0].state_dict())
    import numpy as np
    outs = []
    for g in (ref[0], mine[0]):
        torch.manual_seed(3)
        np.random.seed(3)
        with torch.no_grad():
            v, _ = g.sample_videos(1)
            i, _ = g.sample_images(1)
        outs.append((v, i))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@needs_ref
def test_one_iteration_with_the_reference_nets_on_cpu():
    res = harness.run_harness(iters=1, batch=1, use_reference=True, device="cpu", solver_module=_oracle_shim(), warmup=0)
    assert "reference classes" in res["nets"] and res["ode_trajectories_per_iter"] == 3 * (1 + 32)
    assert all(l == l and abs(l) < 1e3 for l in res["losses"].values()), res["losses"]
    assert res["ode_param_max_update"] > 0          # the generator step reached the ODE parameters through the adjoint


@pytest.mark.gpu
def test_harness_runs_on_the_gpu_through_the_shim():
    assert torch.cuda.is_available()
    import gan_ode_b200 as gode
    try:
        res = harness.run_harness(iters=2, batch=2, precision="bf16", warmup=1)
    finally:
        gode.config.layout, gode.config.precision = "tbd", "fp32"
        sys.modules.pop("torchdiffeq", None)
        sys.modules.pop("torchsde", None)
    assert res["n_gpus"] == 1 and res["ode_forward_ms_per_iter"] > 0
    assert all(l == l and abs(l) < 1e3 for l in res["losses"].values()), res["losses"]
    assert res["ode_param_max_update"] > 0
