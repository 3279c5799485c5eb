"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/busca_b200.h declares (no compute
call is made - there is no GPU here), the product path fails loudly without an sm_100 device, and the host-side logic
that mirrors the reference (memory sampling, option loading, the deliberate activation trap, the filler box)."""
import argparse
import os
import re
from types import SimpleNamespace

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(REPO, "include", "busca_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(busca_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from busca_b200 import _lib
    L = _lib.load()
    names = header_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == names, sorted(set(_lib.EXPORTS) ^ set(names))
    assert b"sm_100a" in L.busca_version()


def _no_cuda_device():
    import torch
    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_cuda_device(), reason="needs a box without a GPU")
def test_no_cpu_fallback_without_device():
    """busca_create must refuse to start without a device: there is no CPU path to fall back to."""
    from busca_b200 import _lib
    from busca_b200.engine import Engine
    with pytest.raises(_lib.BuscaError):
        Engine(device=0, bank_slots=8)


def test_cpu_device_is_rejected():
    from busca_b200.network import _device_index
    assert _device_index("cuda:3") == 3 and _device_index("cuda") == 0 and _device_index(2) == 2 and _device_index(None) == 0
    assert _device_index(SimpleNamespace(index=1)) == 1
    with pytest.raises(RuntimeError):
        _device_index("cpu")


def test_memory_indices_match_the_oracle():
    """BUSCA._memory_indices (host mirror of network.py:247-279) against the oracle's restatement, every history length."""
    from busca_b200.network import BUSCA
    from oracle import network as onet
    for L in (1, 2, 5, 11, 30):
        for n in range(0, 3 * L + 7):
            for broader in (True, False):
                assert BUSCA._memory_indices(n, L, broader) == onet.sample_memory(n, L, broader), (n, L, broader)


def test_option_loading_and_merge(capsys):
    from busca_b200 import option
    for name, thresh, select in (("bytetrack_mot20.yml", 0.3, False), ("bytetrack_mot17.yml", None, None)):
        targs, trargs = option.load_args_from_config(os.path.join(REPO, "busca_b200", "configs", name))
        assert targs.transformer is trargs.transformer and trargs.dataset is not None
        t = targs.transformer
        assert (t.dim_embedding, t.trans_dim, t.nhead, t.ff_size, t.num_layer) == (512, 512, 4, 1024, 4)
        assert t.input_flavour == "MEM-SEP-CAN-BAD" and t.output_flavour == "CAN"
        assert targs.seq_len >= 1 and targs.num_candidates >= 1
        if thresh is not None:
            assert targs.busca_thresh == thresh and targs.select_highest_candidate is select
    cli = argparse.Namespace(busca_thresh=0.77, seq_len=None, brand_new=7)
    merged = option.merge_args(targs, cli, verbose=True)
    out = capsys.readouterr().out
    assert merged.busca_thresh == 0.77 and merged.seq_len == targs.seq_len and merged.brand_new == 7
    assert "Overriding busca_thresh" in out and "Setting brand_new" in out
    assert targs.busca_thresh != 0.77                                     # the base namespace is not mutated


def test_activation_trap_and_filler_box():
    from busca_b200 import custom_layers, tracking
    assert custom_layers.effective_activation("gelu") == "relu"           # the reference's deepcopy/__setstate__ trap
    assert custom_layers.effective_activation("gelu", follow_reference=False) == "gelu"
    with pytest.raises(RuntimeError):
        custom_layers.effective_activation("swish")
    fmin = np.finfo("float32").min
    b64 = tracking.missing_candidate_bbox(legacy_float64=True)
    assert b64.dtype == np.float64 and b64.tolist() == [float(fmin), float(fmin), float(fmin) / 100.0, float(fmin) / 100.0]
    b32 = tracking.missing_candidate_bbox(seq_len=3, flavour="ltwh", legacy_float64=False)
    assert b32.dtype == np.float32 and b32.shape == (3, 4) and b32[0, 2] == -(np.float32(fmin) / np.float32(100.0))
    with pytest.raises(ValueError):
        tracking.missing_candidate_bbox(flavour="xyxy")
    assert tracking.center_distance([], []).shape == (0, 0)               # empty inputs never reach the device
    crop = np.full((384, 128, 3), 255, np.uint8)
    n = tracking.normalize_crop(crop)
    assert n.dtype == np.float32 and np.allclose(n[0, 0], (1.0 - np.array([0.406, 0.456, 0.485])) / np.array([0.225, 0.224, 0.299]), rtol=1e-6)


def test_install_as_busca_gives_the_adapters_their_imports():
    """INTEGRATION.md route A: after install_as_busca() every `from busca.X import Y` the reference adapters execute
    (adapters/*/…byte_tracker.py, strong_sort.py, tracker.py: network.BUSCA, tracking.center_distance, option.load_args_from_config /
    merge_args, visualization.plot_box) resolves to this package - no device needed for the imports themselves."""
    import importlib
    import sys
    import busca_b200
    saved = {k: v for k, v in sys.modules.items() if k == "busca" or k.startswith("busca.")}
    try:
        busca_b200.install_as_busca()
        for mod, names in (("busca.network", ["BUSCA"]), ("busca.tracking", ["center_distance", "get_bbox_crop", "missing_candidate_bbox"]),
                           ("busca.option", ["load_args_from_config", "merge_args"]), ("busca.visualization", ["plot_box"])):
            m = importlib.import_module(mod)
            assert m.__name__.startswith("busca_b200"), (mod, m.__name__)
            for n in names:
                assert hasattr(m, n), (mod, n)
    finally:
        for k in [k for k in sys.modules if k == "busca" or k.startswith("busca.")]:
            del sys.modules[k]
        sys.modules.update(saved)
