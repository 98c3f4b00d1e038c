"""The trainer CLASSES of SURVEY §8 f1 (reference holocron/trainer: Trainer.fit_n_epochs / find_lr / check_setup / save / load,
ClassificationTrainer, BinaryClassificationTrainer, SegmentationTrainer, DetectionTrainer, assign_iou) against golden runs of
the UNMODIFIED reference classes on the same scenarios (tests/_trainer_cases.py, tests/golden/make_golden.py --trainers):
stock-torch models and optimizers on the CPU, i.e. the host-driven path of holocron_b200.trainer.Trainer - every epoch's
evaluation metrics, the final parameters, the step / epoch counters, the checkpoint layout, the printed summaries, parameter
groups, frozen layers, the learning-rate finder's recordings. The device-driven path (TrainStep) is covered on the GPU by
tests/test_gpu_trainer.py."""
import math

import pytest
import torch

import holocron_b200 as hb

import _trainer_cases as cases
from conftest import load_golden


def _close(a, b, tol=1e-5):
    if a is None or b is None:
        assert a is b
        return
    assert math.isclose(a, b, rel_tol=tol, abs_tol=tol), (a, b)


def _same_state(got, want, tol=1e-5):
    assert list(got) == list(want)
    for k in want:
        g, w = got[k].double(), want[k].double()
        assert g.shape == w.shape and float((g - w).norm()) <= tol * (1.0 + float(w.norm())), k


@pytest.fixture(scope="module")
def runs():
    ours = {}
    cases.run_scenarios(hb.trainer, lambda tag, rec: ours.__setitem__(tag, rec))
    return ours, load_golden("trainers")


def test_classification_fit_n_epochs_matches_reference(runs):
    ours, ref = runs
    a, b = ours["cls_fit"], ref["cls_fit"]
    assert len(a["metrics"]) == len(b["metrics"]) == 2
    for ma, mb in zip(a["metrics"], b["metrics"]):
        assert set(ma) == set(mb) == {"val_loss", "acc1", "acc5"}
        for k in mb:
            _close(ma[k], mb[k])
    _same_state(a["state"], b["state"])
    assert (a["step"], a["epoch"]) == (b["step"], b["epoch"]) == (12, 2)
    _close(a["min_loss"], b["min_loss"])
    assert a["ckpt_keys"] == b["ckpt_keys"] == ["epoch", "min_loss", "model", "step"]
    assert (a["ckpt_epoch"], a["ckpt_step"]) == (b["ckpt_epoch"], b["ckpt_step"])
    assert a["msg"] == b["msg"]


def test_frozen_layers_param_groups_cosine_and_nan_skip_match_reference(runs):
    ours, ref = runs
    a, b = ours["cls_frozen_cosine"], ref["cls_frozen_cosine"]
    assert a["groups"] == b["groups"] and a["frozen"] == b["frozen"] == ["0.weight"] and a["step"] == b["step"] == 5
    _same_state(a["state"], b["state"])
    for k in b["metrics"]:
        _close(a["metrics"][k], b["metrics"][k])


def test_binary_and_segmentation_trainers_match_reference(runs):
    ours, ref = runs
    for tag in ("binary", "segmentation"):
        a, b = ours[tag], ref[tag]
        _same_state(a["state"], b["state"])
        assert set(a["metrics"]) == set(b["metrics"])
        for k in b["metrics"]:
            _close(a["metrics"][k], b["metrics"][k])
        assert a["msg"] == b["msg"]


def test_detection_trainer_and_assign_iou_match_reference(runs):
    ours, ref = runs
    a, b = ours["detection"], ref["detection"]
    _close(a["w"], b["w"])
    assert set(a["metrics"]) == set(b["metrics"]) == {"loc_err", "clf_err", "det_err", "val_loss"}
    for k in b["metrics"]:
        _close(a["metrics"][k], b["metrics"][k])
    assert a["assign"] == b["assign"] and a["msg"] == b["msg"] and a["msg_none"] == b["msg_none"]


def test_lr_finder_and_check_setup_match_reference(runs):
    ours, ref = runs
    a, b = ours["find_lr"], ref["find_lr"]
    assert len(a["lrs"]) == len(b["lrs"]) == len(a["losses"]) == 6
    for x, y in zip(a["lrs"] + a["losses"], b["lrs"] + b["losses"]):
        _close(x, y)
    _same_state(a["state_after_check"], b["state_after_check"])
    assert a["too_many"] == b["too_many"] is not None


def test_trainer_errors_and_checkpoint_round_trip(tmp_path):
    model = cases.cls_model()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    tr = hb.trainer.ClassificationTrainer(model, cases.cls_batches(2, 1), cases.cls_batches(1, 2), torch.nn.CrossEntropyLoss(), opt,
                                          gpu=None, output_file=str(tmp_path / "ckpt.pth"))
    with pytest.raises(ValueError):
        tr.fit_n_epochs(1, 1e-3, sched_type="linear")
    with pytest.raises(AssertionError):
        tr.plot_recorder()
    for p in model.parameters():
        p.requires_grad_(False)
    with pytest.raises(AssertionError):
        tr._reset_opt(1e-3)
    for p in model.parameters():
        p.requires_grad_(True)
    tr.epoch, tr.step, tr.min_loss = 3, 17, 0.25
    tr.save(str(tmp_path / "ckpt.pth"))
    other = hb.trainer.ClassificationTrainer(cases.cls_model(), [], [], torch.nn.CrossEntropyLoss(),
                                             torch.optim.Adam(model.parameters(), lr=1e-3), gpu=None)
    other.load(torch.load(tmp_path / "ckpt.pth", map_location="cpu"))
    assert (other.start_epoch, other.epoch, other.step, other.min_loss) == (3, 3, 17, 0.25)
    if not torch.cuda.is_available():
        with pytest.raises(AssertionError):
            hb.trainer.ClassificationTrainer(model, [], [], torch.nn.CrossEntropyLoss(), opt, gpu=0)
    # NaN tolerance of the host-driven path (reference core.py:153-159)
    flaky = cases.FlakyCrossEntropy(bad=1)
    flaky.forward = lambda out, target: torch.nn.functional.cross_entropy(out, target) * float("nan")
    tr = hb.trainer.ClassificationTrainer(cases.cls_model(), cases.cls_batches(4, 1), cases.cls_batches(1, 2), flaky,
                                          torch.optim.SGD(model.parameters(), lr=1e-3), gpu=None, skip_nan_loss=True, nan_tolerance=2)
    tr._reset_scheduler(1e-3, 1, "cosine")
    with pytest.raises(ValueError, match="NaN or inf for more than 2 steps"):
        tr._fit_epoch()
