"""CPU tests of the host mirror of the reference interface (simhand_b200/ops.py): `install()` rebinding, the lazy weight
handles, and the loud failure without a CUDA device (no CPU fallback)."""
import types

import pytest
import torch

import simhand_b200
from simhand_b200 import ops


def _fake_reference_modules():
    """Stand-ins for src.models.utils and a model module that did `from src.models.utils import ...`
    (simhand_w_model.py:14-30): the names exist before install() and point somewhere else."""
    utils = types.ModuleType("src.models.utils")
    model = types.ModuleType("src.models.unsupervised.simhand_w_model")
    for name in ops._DROP_INS:
        setattr(utils, name, lambda *a, **k: "reference")
    for name in ("get_weights_linear", "vanila_weights_contrastive_loss", "vanila_pos_weights_contrastive_loss"):
        setattr(model, name, getattr(utils, name))
    model.unrelated = object()
    return utils, model


def test_install_rebinds_only_names_the_module_has():
    utils, model = _fake_reference_modules()
    unrelated = model.unrelated
    simhand_b200.install(utils, model)
    for name in ops._DROP_INS:
        assert getattr(utils, name) is getattr(ops, name)
    assert model.get_weights_linear is ops.get_weights_linear
    assert model.vanila_weights_contrastive_loss is ops.vanila_weights_contrastive_loss
    assert not hasattr(model, "vanila_neg_weights_contrastive_loss")     # never imported there: not invented
    assert model.unrelated is unrelated


def test_lazy_handles_describe_but_never_materialise_silently():
    j = torch.zeros(6, 21, 3)
    pw, nw = ops.get_weights_linear(j[:, :, :2], j[:, :, :2], "mpjpe")
    assert tuple(pw.shape) == (6,) and tuple(nw.shape) == (12, 12)
    assert nw.dim() == 2 and nw.size(0) == 12 and nw.dtype == torch.float32 and nw.device.type == "cpu"
    assert "LazyWeights" in repr(nw)
    with pytest.raises(AttributeError, match="materialize"):
        nw.mean()                              # a 1 GiB matrix at 2N = 16384 is never built behind the caller's back
    with pytest.raises(ValueError):
        ops.get_weights_linear(j[:, :, :2], j[:, :, :2], "nope")


def test_no_cpu_fallback():
    z = torch.nn.functional.normalize(torch.randn(6, 16), dim=1)
    j = torch.rand(6, 21, 3)
    pw, nw = ops.get_weights_linear(j[:, :, :2], j[:, :, :2], "mpjpe")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.vanila_weights_contrastive_loss(z, z, pw, nw)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pw.materialize()
    with pytest.raises(ValueError, match="same get_weights_linear call"):
        pw2, nw2 = ops.get_weights_linear(j[:, :, :2], j[:, :, :2], "mpjpe")
        ops.vanila_weights_contrastive_loss(z, z, pw, nw2)


def test_exact_weights_switch(monkeypatch):
    from simhand_b200 import _lib
    monkeypatch.delenv("SMH_Q16", raising=False)
    monkeypatch.delenv("SMH_EXACT_WEIGHTS", raising=False)
    assert ops.step_flags("fp16") == _lib.DIMS_Q16_TILES
    assert ops.step_flags("fp16", exact_weights=True) == 0
    assert ops.step_flags("fp32") == 0
    assert ops.step_flags("fp16", ops.make_weighting("non_linear", "mpjpe", 1.0, 1.0)) == 0
    monkeypatch.setenv("SMH_EXACT_WEIGHTS", "1")
    assert ops.step_flags("fp16") == 0
    assert ops.step_flags("fp16", exact_weights=False) == _lib.DIMS_Q16_TILES
