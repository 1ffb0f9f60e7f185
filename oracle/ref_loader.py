"""TEST INFRASTRUCTURE ONLY: executes the reference's own loss functions without importing
its module (the module imports kornia/comet_ml/easydict, which are not installed) and without
copying its text: the FunctionDef nodes are compiled straight from the reference checkout.

Only usable where the checkout exists (the build container); the GPU box uses tests/golden/.
"""
from __future__ import annotations

import ast
import os

import numpy as np
import torch
from torch import Tensor

REF_ROOT = os.environ.get("SIMHAND_REF", "/root/reference")
REF_UTILS = os.path.join(REF_ROOT, "src", "models", "utils.py")

_NAMES = (
    "get_weights_linear",                     # src/models/utils.py:218-261
    "vanila_weights_contrastive_loss",        # src/models/utils.py:391-427
    "vanila_contrastive_loss",                # src/models/utils.py:157-189
    "vanila_pos_weights_contrastive_loss",    # src/models/utils.py:430-465
    "vanila_neg_weights_contrastive_loss",    # src/models/utils.py:468-501
)


def reference_available() -> bool:
    return os.path.isfile(REF_UTILS)


def load_reference_functions(names=_NAMES) -> dict:
    if not reference_available():
        raise FileNotFoundError(f"reference checkout not found at {REF_UTILS}")
    ns = {"torch": torch, "Tensor": Tensor, "np": np}
    with open(REF_UTILS) as fh:
        tree = ast.parse(fh.read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), REF_UTILS, "exec"), ns)
    missing = [n for n in names if n not in ns]
    if missing:
        raise RuntimeError(f"reference functions not found: {missing}")
    return ns
