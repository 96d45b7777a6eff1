"""gan_ode_b200 — B200-native latent-motion ODE hot path of chechaohp/gan-ode.

Public surface = the solver signatures the reference imports:
    from gan_ode_b200 import odeint, odeint_adjoint          # torchdiffeq's (models/mocogan_ode.py:4)
or, without touching the reference files, `gan_ode_b200.install_shims()` registers this package as
`torchdiffeq` / `torchsde` in sys.modules so `from torchdiffeq import odeint_adjoint as odeint` resolves here.
"""
from .odeint import check_status, config, last_adjoint_log, last_step_log, odeint, odeint_adjoint, recognise_field  # noqa: F401
from ._lib import GodeError  # noqa: F401
from .graphed import GraphedSolvePipeline, GraphedSolveStep  # noqa: F401
from .sdeint import GridBrownian, PhiloxBrownian, TableBrownian, adjoint_grid, sdeint, sdeint_adjoint  # noqa: F401
from .odernn import gru_jump, odernn_codes  # noqa: F401
from .fused import FusedLatentSampler, fused_sample_z  # noqa: F401

__version__ = "0.1.0"


def install_shims():
    """Make `import torchdiffeq` / `import torchsde` resolve to this package (reference files stay unmodified)."""
    import sys
    import types

    td = types.ModuleType("torchdiffeq")
    td.odeint = odeint
    td.odeint_adjoint = odeint_adjoint
    td.__version__ = "0.2.2+gan_ode_b200"
    sys.modules["torchdiffeq"] = td
    ts = types.ModuleType("torchsde")
    ts.sdeint = sdeint
    ts.sdeint_adjoint = sdeint_adjoint
    sys.modules["torchsde"] = ts
    return td
