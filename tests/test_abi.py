"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/routeformer_b200.h declares."""
import ctypes
import os
import re

import pytest

from routeformer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "routeformer_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 28
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    # and the binding covers every declared entry point
    bound = set(_lib.SIGNATURES) | {"rf_abi_version", "rf_last_error"}
    assert set(names) == bound, set(names) ^ bound


def test_abi_version_and_struct_mirrors():
    lib = _lib.load()
    assert lib.rf_abi_version() == 1
    for which, struct in _lib.STRUCTS.items():
        assert lib.rf_struct_size(which) == ctypes.sizeof(struct), struct.__name__
    assert lib.rf_struct_size(99) == -1


def test_argument_errors_do_not_need_a_gpu():
    """Validation happens before any CUDA call: a bad argument returns RF_ERR_INVALID_ARGUMENT with a message."""
    lib = _lib.load()
    p = _lib.RfGemmParams()
    assert lib.rf_gemm_tf32(ctypes.byref(p), None) == -1
    assert b"rf_gemm_tf32" in lib.rf_last_error()
    assert lib.rf_median_downsample(None, None, 1, 10, 2, 10, None) == -1


def test_only_sm100a_code_is_shipped():
    import subprocess, shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_build_stamp_is_location_independent(tmp_path):
    """The in-tree library travels with the repository snapshot to the GPU box, where the tree sits under another path: the
    freshness stamp must depend on file contents only, or every rank of a torchrun launch would rebuild (and race on) the .so."""
    import glob
    import os
    import shutil

    from routeformer_b200 import build as B

    src = sorted(glob.glob(os.path.join(B.CSRC, "*.cu")) + glob.glob(os.path.join(B.CSRC, "*.cuh")))
    copy_dir = tmp_path / "elsewhere" / "csrc"
    copy_dir.mkdir(parents=True)
    copies = [shutil.copy(p, copy_dir / os.path.basename(p)) for p in src]
    assert B._digest(src) == B._digest([str(c) for c in copies])
    (copy_dir / os.path.basename(src[0])).write_text("// changed")
    assert B._digest(src) != B._digest([str(c) for c in copies])


def test_patch_reference_rebinds_the_path_symbols():
    """routeformer_b200.compat.patch_reference(): the experiment's own import lines then resolve to the CUDA implementations."""
    import sys
    import types

    import routeformer_b200 as R
    from routeformer_b200.compat import patch_reference

    saved = {k: v for k, v in sys.modules.items() if k == "routeformer" or k.startswith("routeformer.")}
    try:
        for name in ("routeformer", "routeformer.models", "routeformer.models.gps_backbone", "routeformer.score",
                     "routeformer.losses", "routeformer.losses.future_discounted_mse"):
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
        sys.modules["routeformer.models.gps_backbone"].Transformer = "reference class, outside the path"
        done = patch_reference()
        from routeformer import Routeformer  # noqa: I001
        from routeformer.losses.future_discounted_mse import FutureDiscountedLoss
        from routeformer.models import RouteformerConfig
        from routeformer.models.gps_backbone import GPSBackboneConfig, Informer, Transformer
        from routeformer.score import ade, fde
        assert Routeformer is R.Routeformer and RouteformerConfig is R.RouteformerConfig and Informer is R.Informer
        assert GPSBackboneConfig is R.GPSBackboneConfig and FutureDiscountedLoss is R.FutureDiscountedLoss
        assert ade is R.ade and fde is R.fde and Transformer == "reference class, outside the path"
        assert "routeformer.models.cross_modal_transformer" not in done  # not importable here: skipped, not an error
    finally:
        for k in [k for k in sys.modules if k == "routeformer" or k.startswith("routeformer.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No fallback: without the CUDA library the loader raises (it never degrades to a CPU / eager path)."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "librouteformer_b200.so"))
    monkeypatch.setenv("RF_LIB_PATH", str(tmp_path / "librouteformer_b200.so"))  # (also keeps load() from rebuilding in-tree)
    with pytest.raises(_lib.LibraryMissing, match="no fallback"):
        _lib.load()


def test_cpu_tensors_are_refused():
    """The product path has no CPU implementation: ops and the model raise on host tensors instead of computing something."""
    import torch

    import routeformer_b200 as R
    from routeformer_b200 import ops

    a = torch.zeros(8, 8)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        ops.gemm(a, a, torch.zeros(8, 8))
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        ops.layernorm_fwd(a, torch.ones(8), torch.zeros(8), torch.empty(8, 8), torch.empty(8), torch.empty(8))
    g = R.GPSBackboneConfig(seq_len=40, label_len=40, pred_len=30, factor=4, distil=True, dropout=0.0, activation="relu", d_model=32,
                            n_heads=2, e_layers=1, d_layers=1, d_ff=64)
    model = R.Routeformer(R.RouteformerConfig(gps_backbone_config=g, decoder_mode="smart"), gps_backbone=R.Informer).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model({"gps": torch.cumsum(torch.randn(2, 40, 2), dim=1)})
