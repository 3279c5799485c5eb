"""Place-holder for the reference's busca/custom_layers.py (TransformerEncoderLayer / TransformerEncoder).

The four post-LN encoder layers execute inside libbusca_b200.so (busca_transformer / busca_associate:
QKV GEMM, warp-shuffle attention, out-proj + residual, LayerNorm, FFN, LayerNorm); there is no Python
layer object to build.  What this module records is the one behavioural fact a user of those classes
must know:

EFFECTIVE_ACTIVATION - the YAML says ``activation: gelu`` but the reference executes ReLU in every
layer: ``TransformerEncoder`` clones its layer with ``copy.deepcopy`` (custom_layers.py:44-45), deepcopy
calls ``TransformerEncoderLayer.__setstate__`` (custom_layers.py:24-27), which sees no 'activation' key in
the instance ``__dict__`` (sub-modules live in ``_modules``) and injects ``F.relu`` as an instance
attribute that shadows the nn.GELU sub-module.  Pinned by tests/golden/assoc_*.npz.
"""
EFFECTIVE_ACTIVATION = "relu"


def effective_activation(configured: str, follow_reference: bool = True) -> str:
    """Activation the FFN must run to reproduce the reference for a YAML value ``configured``."""
    if configured not in ("relu", "gelu"):
        raise RuntimeError("activation should be relu/gelu, not {}".format(configured))
    return EFFECTIVE_ACTIVATION if follow_reference else configured
