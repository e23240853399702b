"""edl_b200 -- a Blackwell (B200, sm_100a)-native elastic deep-learning engine.

Capabilities mirror elasticdeeplearning/edl (``paddle_edl``): elastic collective launcher, distill
service (DistillReader + teacher discovery/balancing), service registry, elastic data reader and
train-state checkpointing -- rebuilt on PyTorch + hand-written sm_100a CUDA kernels + NVLink 5 /
NVSwitch peer-memory collectives.  ``import paddle_edl`` / ``import edl`` are aliases of this package.
"""
__version__ = "0.1.0"

from .api import (size, rank, local_rank, is_leader_trainer, init_distributed, notify_end_one_batch,  # noqa: E402,F401
                  notify_end_one_epoch, trainer_env)
from .utils.state import State, TorchState, PaddleState, DataCheckpoint, TrainStatus, EpochAttr  # noqa: E402,F401
from .checkpoint import save_check_point, load_check_point  # noqa: E402,F401
