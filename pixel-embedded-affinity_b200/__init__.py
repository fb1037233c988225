"""pea_b200 -- B200-native (sm_100a) implementation of the Pixel-Embedded-Affinity hot path:
embedding -> multi-offset affinities -> weighted-MSE affinity loss (+ hand-derived backward), and
instance labels -> ground-truth affinities / masks / class-balance weights.

The directory is named ``pixel-embedded-affinity_b200``; import it as ``pea_b200`` (alias package at the repo
root).  Importing loads ``csrc/libpea_b200.so`` and fails loudly if it has not been built: there is no CPU path.
"""
from . import _lib  # noqa: F401  (loads the shared library or raises)
from .loss import WeightedMSE, MSELoss, BCELoss, WeightedBCE
from .loss_embedding_mse_2d import (embedding_loss, ema_embedding_loss, embedding2affs, embedding_loss_labels,
                                    embedding_loss_multi)
from .loss_embedding_mse_3d import (embedding_loss_norm1, embedding_loss_norm5, ema_embedding_loss_norm1,
                                    ema_embedding_loss_norm5, inf_embedding_loss_norm1, inf_embedding_loss_norm5,
                                    embedding_loss_norm5_labels, embedding_loss_norm1_multi,
                                    embedding_loss_norm5_self_and_cross)
from .affinity_ours import gen_offsets, multi_offset, gen_affs_ours
from .data_affinity import mknhood3d, seg_to_aff
from .data_segmentation import weight_binary_ratio
from .data_consistency import convert_consistency_flip
from .provider_valid import VolumeBlender
from .gt import make_targets_2d, make_targets_3d, make_targets_pyramid_2d, make_targets_pyramid_3d

__version__ = "0.1.0"
