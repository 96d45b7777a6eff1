"""Drop-in check against the UNMODIFIED reference files, in this container only (they are absent on the GPU box):
the reference's generator runs with this repo's solver signatures plugged in as `torchdiffeq` (here: the CPU oracle,
which has the same signatures as the CUDA boundary), and the tests' stand-in caller equals the real one."""
import importlib
import os
import sys
import types

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not mounted")


@pytest.fixture()
def ref_models(monkeypatch):
    from oracle import torchdiffeq_restatement as tdq
    shim = types.ModuleType("torchdiffeq")
    shim.odeint, shim.odeint_adjoint = tdq.odeint, tdq.odeint_adjoint
    monkeypatch.setitem(sys.modules, "torchdiffeq", shim)
    monkeypatch.syspath_prepend(REF)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    for name in [m for m in sys.modules if m == "models" or m.startswith("models.")]:
        monkeypatch.delitem(sys.modules, name)
    mod = importlib.import_module("models.mocogan_ode")
    yield mod
    for name in [m for m in sys.modules if m == "models" or m.startswith("models.")]:
        sys.modules.pop(name, None)


def test_reference_generator_runs_on_our_signatures(ref_models):
    torch.manual_seed(0)
    gen = ref_models.VideoGeneratorMNISTODE(1, 50, 0, 16, 16)
    assert sum(p.numel() for p in gen.ode_fn.parameters()) == 544
    assert [n for n, _ in gen.ode_fn.named_parameters()] == ["fn.0.weight", "fn.0.bias", "fn.2.weight", "fn.2.bias"]
    vids, _ = gen.sample_videos(4)
    assert tuple(vids.shape) == (4, 1, 16, 28, 28)
    vids.mean().backward()
    assert gen.ode_fn.fn[0].weight.grad is not None and gen.ode_fn.fn[0].weight.grad.abs().sum() > 0


def test_boundary_recognises_the_reference_field(ref_models):
    from gan_ode_b200 import recognise_field
    gen = ref_models.VideoGeneratorMNISTODE(1, 50, 0, 16, 16)
    W1, b1, W2, b2 = recognise_field(gen.ode_fn)
    assert W1.shape == (16, 16) and b1.shape == (16,) and W2.shape == (16, 16) and b2.shape == (16,)


def test_stand_in_caller_equals_reference_caller(ref_models):
    from tests.caller_model import LatentMotionODE
    torch.manual_seed(1)
    gen = ref_models.VideoGeneratorMNISTODE(1, 50, 0, 16, 16)
    mine = LatentMotionODE(16, 16)
    mine.load_state_dict({k: v for k, v in gen.state_dict().items() if k.startswith(("ode_fn.", "linear."))})
    torch.manual_seed(2)
    ref_codes = gen.sample_z_m(8)
    torch.manual_seed(2)
    my_codes = mine.sample_z_m(8)
    assert torch.equal(ref_codes, my_codes)


def test_odernn_stand_in_equals_reference_caller(ref_models, monkeypatch):
    """models/mocogan_ode_rnn.py imports a non-existent `on_dev` package: alias it to `models`."""
    import models
    from tests.caller_model import LatentMotionODERNN
    monkeypatch.setitem(sys.modules, "on_dev", models)
    monkeypatch.setitem(sys.modules, "on_dev.mocogan_ode", ref_models)
    rnn_mod = importlib.import_module("models.mocogan_ode_rnn")
    torch.manual_seed(3)
    gen = rnn_mod.VideoGeneratorMNISTODERNN(1, 50, 0, 16, 4)
    mine = LatentMotionODERNN(16, 4)
    mine.load_state_dict({k: v for k, v in gen.state_dict().items() if k.startswith(("ode_fn.", "recurrent."))})
    torch.manual_seed(4)
    ref_codes = gen.sample_z_m(3)
    torch.manual_seed(4)
    my_codes = mine.sample_z_m(3)
    assert ref_codes.shape == (12, 16)
    assert torch.equal(ref_codes, my_codes)
