"""Host-side trainer semantics against golden runs of the UNMODIFIED reference Trainer (tests/golden/make_golden.py --trainer:
holocron/trainer/core.py _reset_opt / _reset_scheduler / _fit_epoch on a tiny RepVGG): the per-iteration (lr, beta1) the
reference's scheduler sets, and the freezing / parameter-splitting helpers of holocron/trainer/utils.py. No GPU needed."""
import torch

import holocron_b200 as hb
from holocron_b200.models.classification.repvgg import RepVGG
from holocron_b200.trainer import freeze_bn, freeze_model, lr_schedule_table, split_normalization_params

from conftest import load_golden


def tiny():
    torch.manual_seed(0)
    return RepVGG([1, 1, 1], [16, 32, 64], 1, 1, num_classes=10)


class _Opt:
    """Stand-in with the attributes lr_schedule_table reads (the fused optimizers themselves need CUDA tensors only in step())."""
    def __init__(self, lr, betas):
        self.param_groups = [{"lr": lr, "betas": betas}]


def test_schedule_table_reproduces_the_reference_scheduler():
    g = load_golden("trainer")
    for tag in ("acc2_clip_onecycle", "nan_skip_cosine"):
        d = g[tag]
        cfg = d["cfg"]
        table = lr_schedule_table(_Opt(cfg["lr"], (0.95, 0.99)), cfg["lr"], 8, cfg["sched"])
        assert table.shape == (8, 2)
        assert torch.allclose(table[:, 0].double(), d["lrs"], rtol=1e-6, atol=0), (tag, table[:, 0], d["lrs"])
        if cfg["sched"] == "onecycle":       # OneCycleLR also cycles beta1 (cycle_momentum defaults to True)
            assert torch.allclose(table[:, 1].double(), d["beta1s"], rtol=1e-6, atol=0)
        else:
            assert bool((table[:, 1] == -1).all()) and bool((d["beta1s"] == 0.95).all())


def test_freeze_helpers_match_reference():
    g = load_golden("trainer")
    m = tiny()
    freeze_model(m.train(), "features.1")
    assert [n for n, p in m.named_parameters() if not p.requires_grad] == g["freeze"]["frozen"]
    assert [n for n, mod in m.named_modules() if isinstance(mod, torch.nn.BatchNorm2d) and not mod.training] == g["freeze"]["bn_eval"]
    norm, other = split_normalization_params(tiny())
    assert len(norm) == g["split"]["norm"] and len(other) == g["split"]["other"]
    assert sum(p.numel() for p in norm) == g["split"]["norm_numel"] and sum(p.numel() for p in other) == g["split"]["other_numel"]
    m2 = tiny()
    for p in m2.features[0].parameters():
        p.requires_grad_(False)
    freeze_bn(m2.train())
    frozen = [mod for mod in m2.features[0].modules() if isinstance(mod, torch.nn.BatchNorm2d)]
    assert frozen and all((not b.training) and (not b.track_running_stats) for b in frozen)
    assert all(b.training for b in m2.features[1].modules() if isinstance(b, torch.nn.BatchNorm2d))


def test_tiny_model_state_matches_reference_init():
    """The golden runs start from the reference's seeded init; this package's tree must produce the same one."""
    g = load_golden("trainer")
    sd = tiny().state_dict()
    assert list(sd) == list(g["acc2_clip_onecycle"]["state"])
    assert hb.optim.AdaBelief is not None


def test_checkpoint_layout_is_the_reference_one(tmp_path):
    """TrainStep.save writes the reference Trainer's checkpoint dict (core.py:106-121: epoch / step / min_loss / model,
    legacy serialisation); TrainStep.load resumes from a dict of that layout, e.g. one written by the reference."""
    from holocron_b200.trainer import TrainStep
    g = load_golden("trainer")
    m = tiny()
    ts = TrainStep(m, torch.nn.CrossEntropyLoss(), torch.optim.SGD(m.parameters(), lr=0.1), graph=False)
    ts.epoch, ts.iterations, ts.min_loss = 3, 17, 0.25
    path = tmp_path / "ckpt.pth"
    ts.save(str(path))
    state = torch.load(path, weights_only=False)
    assert list(state) == ["epoch", "step", "min_loss", "model"] and state["epoch"] == 3 and state["step"] == 17
    assert list(state["model"]) == list(g["acc2_clip_onecycle"]["state"])       # the reference model's state_dict keys
    # a checkpoint as the reference writes it (its final golden state) loads into this package's model
    ref_state = {"epoch": 1, "step": 8, "min_loss": 1.5, "model": g["acc2_clip_onecycle"]["state"]}
    m2 = tiny()
    ts2 = TrainStep(m2, torch.nn.CrossEntropyLoss(), torch.optim.SGD(m2.parameters(), lr=0.1), graph=False)
    ts2.load(ref_state)
    assert ts2.epoch == 1 and ts2.iterations == 8 and ts2.min_loss == 1.5
    for k, v in m2.state_dict().items():
        assert torch.equal(v, g["acc2_clip_onecycle"]["state"][k])
