"""Checkpoint descriptions — the record types of holocron/models/checkpoints.py (TrainingRecipe :31-42, Metric :45-49, Dataset
:52-57, Evaluation :60-65, LoadingMeta :68-77, PreProcessing :80-87, Checkpoint :90-97, _handle_legacy_pretrained :100-109).

Plain metadata: what a checkpoint was trained / evaluated on, where its ``state_dict`` lives (``meta.url``; ``file://`` URLs work
without a network) and how inputs must be pre-processed. The records are generated from one field table below (same class
names, field names, order, defaults and keyword construction as the reference's dataclasses, so ``Checkpoint(evaluation=...,
meta=..., pre_processing=..., recipe=...)`` objects are interchangeable in user code). The reference's tables of released
checkpoints (GitHub release URLs) are not shipped: ``pretrained=True`` needs an explicit ``checkpoint=``."""
import logging
from dataclasses import field, is_dataclass, make_dataclass
from enum import Enum
from typing import Dict, List, Tuple, Union

from torchvision.transforms.functional import InterpolationMode

__all__ = ["Checkpoint", "Dataset", "Evaluation", "LoadingMeta", "Metric", "PreProcessing", "TrainingRecipe"]

logger = logging.getLogger(__name__)

# string-valued enumerations (members compare equal to their values, like the reference's `class X(str, Enum)`)
Metric = Enum("Metric", {"TOP1_ACC": "top1-accuracy", "TOP5_ACC": "top5-accuracy"}, type=str, module=__name__)
Metric.__doc__ = "Evaluation metric"
Dataset = Enum("Dataset", {"IMAGENET1K": "imagenet-1k", "IMAGENETTE": "imagenette", "CIFAR10": "cifar10"}, type=str,
               module=__name__)
Dataset.__doc__ = "Training / evaluation dataset"

_OPT_STR = Union[str, None]
_RECORDS = {
    # name: (docstring, [(field, type[, default]) ...])
    "TrainingRecipe": ("Commit, script and command-line arguments that produced the checkpoint.",
                       [("commit", _OPT_STR), ("script", _OPT_STR), ("args", _OPT_STR)]),
    "Evaluation": ("Results of the model's evaluation on a dataset.",
                   [("dataset", Dataset), ("results", Dict[Metric, float])]),
    "LoadingMeta": ("What is needed to fetch and load the parameters.",
                    [("url", str), ("sha256", str), ("size", int), ("arch", str), ("num_params", int), ("categories", List[str])]),
    "PreProcessing": ("Input pre-processing the checkpoint expects.",
                      [("input_shape", Tuple[int, ...]), ("mean", Tuple[float, ...]), ("std", Tuple[float, ...]),
                       ("interpolation", InterpolationMode, field(default=InterpolationMode.BILINEAR))]),
}
for _name, (_doc, _fields) in _RECORDS.items():
    globals()[_name] = make_dataclass(_name, _fields, namespace={"__doc__": _doc, "__module__": __name__})
TrainingRecipe, Evaluation, LoadingMeta, PreProcessing = (globals()[n] for n in _RECORDS)
Checkpoint = make_dataclass(
    "Checkpoint", [("evaluation", Evaluation), ("meta", LoadingMeta), ("pre_processing", PreProcessing), ("recipe", TrainingRecipe)],
    namespace={"__doc__": "Everything needed to run a model in the conditions of its checkpoint.", "__module__": __name__})
assert all(is_dataclass(c) for c in (TrainingRecipe, Evaluation, LoadingMeta, PreProcessing, Checkpoint))


def _handle_legacy_pretrained(pretrained: bool = False, checkpoint=None, default_checkpoint=None):
    """``pretrained=True`` selects the default checkpoint unless one was given explicitly; warns when there is none."""
    chosen = checkpoint if checkpoint else (default_checkpoint if pretrained else None)
    if pretrained and chosen is None:
        logger.warning("Invalid model URL, using default initialization.")
    return chosen
