"""deqsci_b200 — B200-native (sm_100a) DE-GAP reconstruction hot path of IndigoPurple/DEQSCI.

The sub-packages mirror the reference's module paths (`utils.cg_utils`, `operators.operator`,
`networks.ffdnet.models`, `networks.provable.model.SimpleCNN_models`,
`solvers.equilibrium_solvers_yaping`, `solvers.new_equilibrium_utils_yaping`) with the same
names, signatures and checkpoint keys; underneath they call hand-written CUDA through the C-ABI of
include/deqsci.h (deqsci_b200/libdeqsci.so).  `install_reference_aliases()` exposes them under the
reference's top-level names so its scripts run unchanged."""
import importlib
import sys

from ._lib import DeqsciError, LIB_PATH  # noqa: F401

_ALIASES = [
    "utils", "utils.cg_utils", "operators", "operators.operator", "networks", "networks.ffdnet",
    "networks.ffdnet.models", "networks.ffdnet.functions", "networks.provable", "networks.provable.model",
    "networks.provable.model.SimpleCNN_models", "networks.provable.model.conv_sn_chen",
    "networks.provable.model.models", "networks.provable.model.realSN_models",
    "networks.provable.model.Spectral_Normalize_chen", "solvers",
    "solvers.equilibrium_solvers_yaping", "solvers.new_equilibrium_utils_yaping",
    "training", "training.sci_equilibrium_training", "utils.sci_dataloader", "utils.metrics",
]


def install_reference_aliases():
    """Registers deqsci_b200.<path> as <path> in sys.modules for every mirrored reference module."""
    for name in _ALIASES:
        sys.modules[name] = importlib.import_module("deqsci_b200." + name)
