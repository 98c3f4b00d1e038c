"""Checkpoint descriptions — API mirror of holocron/models/checkpoints.py (TrainingRecipe :31-42, Metric :45-49, Dataset
:52-57, Evaluation :60-65, LoadingMeta :68-77, PreProcessing :80-87, Checkpoint :90-97, _handle_legacy_pretrained :100-109).

Plain metadata records: what a checkpoint was trained / evaluated on, where its ``state_dict`` lives (``meta.url``; ``file://``
URLs work without a network) and how inputs must be pre-processed. The tables of released checkpoints of the reference
(GitHub release URLs) are not shipped: ``pretrained=True`` needs an explicit ``checkpoint=``."""
import logging
from dataclasses import dataclass
from enum import Enum
from typing import Dict, List, Tuple, Union

from torchvision.transforms.functional import InterpolationMode

__all__ = ["Checkpoint", "Dataset", "Evaluation", "LoadingMeta", "Metric", "PreProcessing", "TrainingRecipe"]

logger = logging.getLogger(__name__)


@dataclass
class TrainingRecipe:
    """Commit, script and arguments that produced the checkpoint."""

    commit: Union[str, None]
    script: Union[str, None]
    args: Union[str, None]


class Metric(str, Enum):
    TOP1_ACC = "top1-accuracy"
    TOP5_ACC = "top5-accuracy"


class Dataset(str, Enum):
    IMAGENET1K = "imagenet-1k"
    IMAGENETTE = "imagenette"
    CIFAR10 = "cifar10"


@dataclass
class Evaluation:
    dataset: Dataset
    results: Dict[Metric, float]


@dataclass
class LoadingMeta:
    url: str
    sha256: str
    size: int
    arch: str
    num_params: int
    categories: List[str]


@dataclass
class PreProcessing:
    input_shape: Tuple[int, ...]
    mean: Tuple[float, ...]
    std: Tuple[float, ...]
    interpolation: InterpolationMode = InterpolationMode.BILINEAR


@dataclass
class Checkpoint:
    """Everything needed to run a model in the conditions of its checkpoint."""

    evaluation: Evaluation
    meta: LoadingMeta
    pre_processing: PreProcessing
    recipe: TrainingRecipe


def _handle_legacy_pretrained(pretrained: bool = False, checkpoint: Union[Checkpoint, None] = None,
                              default_checkpoint: Union[Checkpoint, None] = None) -> Union[Checkpoint, None]:
    checkpoint = checkpoint or (default_checkpoint if pretrained else None)
    if pretrained and checkpoint is None:
        logger.warning("Invalid model URL, using default initialization.")
    return checkpoint
