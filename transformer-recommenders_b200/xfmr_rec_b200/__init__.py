"""xfmr_rec_b200 — B200-native scoring-and-loss / full-catalog top-k path of
yxtay/transformer-recommenders behind the reference's own Python interfaces.

    losses   mirror of xfmr_rec/losses.py      (LossConfig, 7 losses, LogitsStatistics, registry)
    models   compute_embeds of xfmr_rec/models.py:366-419 without the O(M^2 D) candidate copy
    index    exact full-catalog search with the LanceIndex.search surface (index.py:214-255)
    metrics  compute_retrieval_metrics of xfmr_rec/metrics.py:17-79 (+ batched device version)
    evaluate batched validation loop: exact search with history excluded + the 7 metrics for U users
    dist     catalog sharding + NCCL all-gather merge, data-parallel loss reduction
    step     the whole scoring-and-loss train step as one sync-free, CUDA-graph-replayed call
    data     SeqBatch construction on the device (SeqDataset sampling + collate, data.py:669-805)
    encoder  the sequence encoder (BertModel(is_decoder=True) on item embeddings), forward + backward
    service  request / response types of xfmr_rec/service.py and the ItemIndex service surface
    ops      tensor-level wrappers over the C ABI (include/xfmr_b200.h)
"""

from . import _native, ops  # noqa: F401
from . import data, dist, encoder, evaluate, index, losses, metrics, models, params, service, step  # noqa: F401
from .losses import (  # noqa: F401
    LOSS_CLASSES,
    AlignmentContrastiveLoss,
    AlignmentLoss,
    ContrastiveLoss,
    EmbedLoss,
    InfoNCELoss,
    LogitsStatistics,
    LossConfig,
    LossType,
    NCELoss,
    PairwiseHingeLoss,
    PairwiseLogisticLoss,
    PoolCandidates,
    SampledCandidates,
)

from .step import PoolLossStep  # noqa: F401,E402

__version__ = "0.1.0"
