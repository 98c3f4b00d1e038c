"""GPU tests of the captured training step (holocron_b200.trainer.TrainStep) against golden runs of the UNMODIFIED reference
Trainer (tests/golden/trainer.pt, generated on the CPU in fp32 by tests/golden/make_golden.py --trainer from
holocron/trainer/core.py): gradient accumulation + clip_grad_norm_ + OneCycleLR (lr and beta1), and NaN-loss skipping +
CosineAnnealingLR - per-iteration losses, final parameters, optimizer step count. Also the fused AdamP update against the
reference formulas (oracle/optim.py)."""
import pytest
import torch

import holocron_b200 as hb
from holocron_b200.models.classification.repvgg import RepVGG
from holocron_b200.trainer import TrainStep, lr_schedule_table

from conftest import load_golden

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def tiny():
    torch.manual_seed(0)
    return RepVGG([1, 1, 1], [16, 32, 64], 1, 1, num_classes=10)


def batches(n, nan_at=None):
    g = torch.Generator().manual_seed(31)
    out = []
    for i in range(n):
        x = (torch.rand(8, 3, 32, 32, generator=g) - 0.45) / 0.225
        if nan_at is not None and i == nan_at:
            x = x.clone()
            x[0, 0, 0, 0] = float("nan")
        out.append((x, torch.randint(0, 10, (8,), generator=g)))
    return out


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("tag", ["acc2_clip_onecycle", "nan_skip_cosine"])
def test_train_step_reproduces_reference_trainer(tag, graph):
    d = load_golden("trainer")[tag]
    cfg = d["cfg"]
    model = tiny().cuda().to(memory_format=torch.channels_last).train()
    opt = hb.optim.AdaBelief(model.parameters(), lr=cfg["lr"], betas=(0.95, 0.99), eps=1e-6, capturable=True)
    table = lr_schedule_table(opt, cfg["lr"], 8, cfg["sched"])
    step = TrainStep(model, torch.nn.CrossEntropyLoss(), opt, gradient_acc=cfg["gradient_acc"], grad_clip=cfg["gradient_clip"],
                     skip_nan_loss=cfg["skip_nan_loss"], nan_tolerance=5, schedule=table, graph=graph)
    losses = []
    for x, t in batches(8, cfg["nan_at"]):
        losses.append(step(x.cuda(), t.cuda()).float().item())
    st = step.state()
    ref = d["losses"]
    print(f"\n[trainer {tag} graph={graph}] losses", [round(v, 4) for v in losses], "reference", [round(float(v), 4) for v in ref], st)
    assert st["iter"] == 8 and st["opt_steps"] == d["opt_steps"]
    for i, (a, b) in enumerate(zip(losses, ref.tolist())):
        if b != b:
            assert a != a, "the NaN batch must produce a NaN loss here too"
        else:
            assert abs(a - b) / abs(b) < 2e-2, (i, a, b)
    # the last row of the schedule was the one in effect during the last iteration
    assert abs(st["lr"] - float(d["lrs"][-1 if cfg["gradient_acc"] == 1 else -1])) / float(d["lrs"][-1]) < 1e-5
    # parameters after the 4 (or 7) updates: all of them as one vector, and the direction of the total update. (Per-tensor
    # relative errors are meaningless for the BatchNorm biases: they start at exactly 0 and have moved by a few lr only.)
    sd = model.state_dict()
    init = tiny().state_dict()
    keys = [k for k, v in d["state"].items() if v.dtype.is_floating_point and "running" not in k]
    ours = torch.cat([sd[k].detach().float().cpu().flatten() for k in keys])
    ref_p = torch.cat([d["state"][k].flatten() for k in keys])
    p0 = torch.cat([init[k].flatten() for k in keys])
    e_all = rel_l2(ours, ref_p)
    cos = torch.nn.functional.cosine_similarity(ours - p0, ref_p - p0, dim=0).item()
    # running statistics: the NaN batch poisons them in the reference as well (its forward pass is not skipped) - the NaN
    # pattern must be identical, the finite entries close
    stats = 0.0
    for k, v in d["state"].items():
        if "running" in k:
            mine = sd[k].detach().float().cpu()
            assert torch.equal(torch.isnan(mine), torch.isnan(v)), k
            ok = ~torch.isnan(v)
            if ok.any():
                stats = max(stats, rel_l2(mine[ok], v[ok]))
    print(f"[trainer {tag}] parameters rel-L2 {e_all:.5f}, update cosine {cos:.4f}, running statistics {stats:.4f}")
    assert e_all < 2e-2 and cos > 0.9 and stats < 5e-2
    step.check()   # nan_run never exceeded the tolerance


def test_nan_tolerance_raises_like_the_reference():
    model = tiny().cuda().to(memory_format=torch.channels_last).train()
    opt = hb.optim.AdaBelief(model.parameters(), lr=1e-3, capturable=True)
    step = TrainStep(model, torch.nn.CrossEntropyLoss(), opt, skip_nan_loss=True, nan_tolerance=2, graph=False)
    x, t = batches(1, nan_at=0)[0]
    before = [p.detach().clone() for p in model.parameters()]
    for _ in range(3):
        step(x.cuda(), t.cuda())
    assert all(torch.equal(a, b) for a, b in zip(before, model.parameters()))   # every update skipped
    assert step.state()["opt_steps"] == 0 and step.state()["nan_run"] == 3
    with pytest.raises(ValueError):
        step.check()


def test_grad_clip_matches_torch_clip_grad_norm():
    from holocron_b200._lib import lib, ptr, stream_ptr
    import ctypes
    torch.manual_seed(0)
    g = torch.randn(1_000_003, device="cuda")
    ref = g.clone()
    p = torch.nn.Parameter(torch.zeros_like(ref))
    p.grad = ref
    norm = torch.nn.utils.clip_grad_norm_([p], 0.7)
    scratch = torch.empty(lib().hb_grad_clip_partials_max(), device="cuda", dtype=torch.float64)
    ctl = torch.zeros(8, device="cuda")
    pad = torch.zeros(g.numel() + 4, device="cuda")
    buf = pad[:g.numel()]
    buf.copy_(g)
    assert lib().hb_grad_clip_norm(ptr(buf), buf.numel(), ctypes.c_float(0.7), ptr(scratch), ptr(ctl), stream_ptr()) == 0
    assert rel_l2(buf, p.grad) < 1e-6 and abs(ctl[7].item() - norm.item()) / norm.item() < 1e-6
    # below the threshold nothing changes
    small = g * 1e-6
    keep = small.clone()
    assert lib().hb_grad_clip_norm(ptr(small), small.numel(), ctypes.c_float(0.7), ptr(scratch), None, stream_ptr()) == 0
    assert torch.equal(small, keep)


@pytest.mark.parametrize("amsgrad,wd", [(False, 0.0), (True, 1e-2)])
def test_adamp_matches_reference_formulas(amsgrad, wd):
    g = load_golden("optim")["adamp"]      # reference AdamP trajectory (tests/golden/make_golden.py gen_optim)
    ps, gs = g["params"], g["grads"]
    params = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
    opt = hb.optim.AdamP(params, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd, amsgrad=amsgrad, delta=0.1)
    for it in range(1, 4):
        for p, gr in zip(params, gs):
            p.grad = (gr * it).cuda()
        opt.step()
    for p, r in zip(params, g["after"][f"adamp_{int(amsgrad)}"]):
        assert rel_l2(p, r) < 1e-5
    sd = opt.state_dict()
    assert set(sd["state"][0]) >= {"step", "exp_avg", "exp_avg_sq"} and sd["state"][0]["step"] == 3
    assert isinstance(opt, torch.optim.Adam)


@pytest.mark.parametrize("tag", ["acc2_clip_onecycle", "nan_skip_cosine"])
def test_trainer_class_device_path_reproduces_reference_trainer(tag, tmp_path):
    """The reference's call sequence (Trainer(...); freeze_model; _reset_opt; _reset_scheduler; _fit_epoch) through
    holocron_b200.trainer.ClassificationTrainer on cuda:0 with the fused AdaBelief: the device path is selected (TrainStep:
    accumulation, clipping, schedule table and NaN skipping on the device), counters and final parameters match the golden
    run of the reference's own Trainer; then evaluate() / save() / load() on the device."""
    d = load_golden("trainer")[tag]
    cfg = d["cfg"]
    model = tiny()
    data = batches(8, cfg["nan_at"])
    opt = hb.optim.AdaBelief(model.parameters(), lr=1e-3, betas=(0.95, 0.99), eps=1e-6, capturable=True)
    tr = hb.trainer.ClassificationTrainer(model, data, data[:2], torch.nn.CrossEntropyLoss(), opt, gpu=0,
                                          output_file=str(tmp_path / "ckpt.pth"), skip_nan_loss=cfg["skip_nan_loss"], nan_tolerance=5,
                                          gradient_acc=cfg["gradient_acc"], gradient_clip=cfg["gradient_clip"], log_every=4)
    assert next(tr.model.parameters()).is_cuda
    hb.trainer.freeze_model(tr.model.train(), None)
    tr._reset_opt(cfg["lr"], None)
    tr._reset_scheduler(cfg["lr"], 1, cfg["sched"])
    assert tr._train_step is not None                      # fused optimizer with capturable=True -> device path
    tr._fit_epoch()
    st = tr._train_step.state()
    assert (tr.step, tr.epoch, st["iter"], st["opt_steps"]) == (8, 1, 8, d["opt_steps"])
    sd = tr.model.state_dict()
    init = tiny().state_dict()
    keys = [k for k, v in d["state"].items() if v.dtype.is_floating_point and "running" not in k]
    ours = torch.cat([sd[k].detach().float().cpu().flatten() for k in keys])
    ref_p = torch.cat([d["state"][k].flatten() for k in keys])
    p0 = torch.cat([init[k].flatten() for k in keys])
    e_all = rel_l2(ours, ref_p)
    cos = torch.nn.functional.cosine_similarity(ours - p0, ref_p - p0, dim=0).item()
    print(f"\n[trainer class {tag}] parameters rel-L2 {e_all:.5f}, update cosine {cos:.4f}, recorded losses {tr.loss_recorder}")
    assert e_all < 2e-2 and cos > 0.9
    assert len(tr.loss_recorder) == 2                      # loss read back every `log_every` = 4 iterations
    if cfg["nan_at"] is None:                              # (the NaN batch poisons the running statistics, as in the reference)
        m = tr.evaluate()
        assert set(m) == {"val_loss", "acc1", "acc5"} and m["val_loss"] == m["val_loss"] and 0.0 <= m["acc1"] <= m["acc5"] <= 1.0
        assert "Acc@1" in tr._eval_metrics_str(m)
    tr.save(str(tmp_path / "ckpt.pth"))
    state = torch.load(tmp_path / "ckpt.pth", map_location="cpu")
    assert sorted(state) == ["epoch", "min_loss", "model", "step"] and state["step"] == 8
    tr.load(state)
    assert tr.epoch == 1 and tr.step == 8


def test_trainer_class_fit_n_epochs_and_lr_finder_on_device(tmp_path):
    """fit_n_epochs (two epochs, one-cycle) + find_lr + check_setup end to end on the device path: the loss goes down, the
    checkpoint of the best epoch is written, the finder records one loss per rate."""
    model = tiny()
    data = batches(8)
    opt = hb.optim.AdaBelief(model.parameters(), lr=1e-3, betas=(0.95, 0.99), eps=1e-6, capturable=True)
    seen = []
    tr = hb.trainer.ClassificationTrainer(model, data, data[:3], torch.nn.CrossEntropyLoss(), opt, gpu=0,
                                          output_file=str(tmp_path / "best.pth"), on_epoch_end=lambda m: seen.append(dict(m)))
    tr.fit_n_epochs(2, 2e-3, sched_type="onecycle")
    assert len(seen) == 2 and tr.epoch == 2 and tr.step == 16 and (tmp_path / "best.pth").exists()
    assert all(m["val_loss"] == m["val_loss"] for m in seen) and tr.min_loss == min(m["val_loss"] for m in seen)
    tr.find_lr(start_lr=1e-5, end_lr=1e-2, num_it=6)
    assert len(tr.lr_recorder) == len(tr.loss_recorder) == 6 and abs(tr.lr_recorder[-1] - 1e-2) < 1e-9
    tr.check_setup(lr=1e-3, num_it=6)
    assert len(tr.loss_recorder) == 6 and tr.loss_recorder[-1] < tr.loss_recorder[0]
    with pytest.raises(ValueError):
        tr.find_lr(num_it=100)
