"""busca_b200: BUSCA's per-frame association hot path as hand-written sm_100a CUDA kernels behind the
reference's own plug-in API (busca.network.BUSCA, busca.tracking.center_distance, busca/option.py).

Importing this package does not touch the GPU; constructing ``BUSCA`` (or calling a tracking function) loads
``libbusca_b200.so`` and fails loudly if it - or a B200 - is missing.  See DESIGN.md and INTEGRATION.md.
"""
__all__ = ["install_as_busca"]


def install_as_busca():
    """Make ``import busca`` resolve to this package so the UNMODIFIED adapters/* files run on it:
    ``from busca.network import BUSCA``, ``from busca.tracking import center_distance``,
    ``from busca.option import load_args_from_config``, ``from busca.visualization import plot_box``."""
    import importlib
    import sys
    pkg = importlib.import_module(__name__)
    sys.modules["busca"] = pkg
    for sub in ("network", "tracking", "option", "custom_layers", "visualization"):
        mod = importlib.import_module(f"{__name__}.{sub}")
        sys.modules[f"busca.{sub}"] = mod
        setattr(pkg, sub, mod)
    return pkg
