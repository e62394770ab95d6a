"""linkless_link_prediction_b200 — B200-native (sm_100a) implementation of the LLP training / scoring hot path.

Drop-in surface of snap-research/linkless-link-prediction for that path (same class, function and flag names):
``models`` (MLP / SAGE / LinkPredictor), ``sageconv`` (SAGEConv / SAGEConv_updated), ``train_teacher_gnn``
(train / test_transductive / test_production / main), ``main`` (student KD: train / train_minibatch /
neighbor_samplers / kl_loss / main), ``shims`` (negative_sampling / random_walk / Evaluator / seed_everything).
All arithmetic runs in hand-written CUDA kernels behind the C-ABI of ``libllp_b200.so`` (``include/llp_b200.h``).
"""
from . import _native  # noqa: F401
from .ops import compute_dtype, set_compute_dtype  # noqa: F401
from .models import MLP, SAGE, LinkPredictor  # noqa: F401
from .sageconv import SAGEConv, SAGEConv_updated  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .shims import Data, Evaluator, negative_sampling, random_walk, seed_everything  # noqa: F401

__version__ = "0.1.0"
