"""simhand_b200 -- B200-native similarity-weighted NT-Xent loss (SiMHand handclr_w / peclr_w / simclr_w).

Public API (mirrors `src/models/utils.py` of the reference):
    get_weights_linear, vanila_weights_contrastive_loss, weighted_ntxent, l2_normalize, install,
    get_transformed_projections (fused normalise/translate/rotate/normalise), HostPipeline (host-buffer front end)
"""
from .ops import (LazyWeights, apply_pca, get_weights_linear_with_pca, get_weights_nonlinear_with_pca, get_transformed_projections, get_weights_linear, get_weights_nonlinear, install, l2_normalize, mpjpe_weights, run_step,  # noqa: F401
                  vanila_contrastive_loss, vanila_neg_weights_contrastive_loss,
                  vanila_pos_weights_contrastive_loss, vanila_weights_contrastive_loss, weighted_ntxent)

from .pipeline import HostPipeline  # noqa: F401,E402
from .head import FusedProjectionHead  # noqa: F401,E402

__version__ = "0.1.0"
