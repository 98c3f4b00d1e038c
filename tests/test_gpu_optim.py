"""GPU parity tests for the fused optimizer steps (AdaBelief / LAMB / TAdam) against the reference's trajectories
(golden fixtures) and the CPU oracle. fp32 state: rtol 1e-4 (north_star allows 1e-3)."""
import pytest
import torch

import holocron_b200 as hb
from oracle import optim as OO

from conftest import load_golden

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-4, atol=1e-6):
    torch.testing.assert_close(a.detach().cpu().float(), b.detach().cpu().float(), rtol=rtol, atol=atol)


@pytest.mark.parametrize("name,cls", [("adabelief", "AdaBelief"), ("adabelief_wd_ams", "AdaBelief"), ("lamb", "LAMB"),
                                      ("lamb_wd", "LAMB"), ("tadam", "TAdam"), ("tadam_wd_ams_dof", "TAdam")])
def test_trajectories_vs_reference(name, cls):
    g = load_golden("optim")
    params = [torch.nn.Parameter(p.clone().cuda()) for p in g["p0"]]
    opt = getattr(hb.optim, cls)(params, **g[name + "_kw"])
    for step in range(3):
        for p, gr in zip(params, g["grads"][step]):
            p.grad = gr.clone().cuda()
        opt.step()
        for p, ref in zip(params, g[name][step]):
            close(p, ref)
    st = opt.state[params[0]]
    assert st["step"] == 3 and isinstance(st["step"], int)
    assert set(st) >= {"step", "exp_avg", "exp_avg_sq"}
    if cls == "TAdam":
        assert st["W_t"].shape == (1,)
    if cls == "LAMB":
        assert "local_lr" in st
    # state_dict round trip keeps working (Optimizer protocol)
    sd = opt.state_dict()
    opt2 = getattr(hb.optim, cls)(params, **g[name + "_kw"])
    opt2.load_state_dict(sd)
    for p, gr in zip(params, g["grads"][0]):
        p.grad = gr.clone().cuda()
    opt2.step()


def test_adabelief_capturable_matches_default():
    """capturable=True (device-side step counter for CUDA-graph replay) follows the same trajectory."""
    g = load_golden("optim")
    params = [torch.nn.Parameter(p.clone().cuda()) for p in g["p0"]]
    opt = hb.optim.AdaBelief(params, capturable=True, **g["adabelief_kw"])
    for step in range(3):
        for p, gr in zip(params, g["grads"][step]):
            p.grad = gr.clone().cuda()
        opt.step()
        for p, ref in zip(params, g["adabelief"][step]):
            close(p, ref)


def test_optimizer_changes_params_like_reference_test():
    # reference tests/test_optim.py:10-39: one step on a 1024 -> 10 classifier layer must change its weight
    # (for TAdam the first update is ~1e-10 - the first w_t is (dof+d)/(sum(g^2)/eps) - so only tiny weights move,
    # exactly as in the reference)
    for cls, kw in ((hb.optim.AdaBelief, {}), (hb.optim.LAMB, {"weight_decay": 2e-5}), (hb.optim.TAdam, {})):
        torch.manual_seed(0)
        lin = torch.nn.Linear(1024, 10).cuda()
        opt = cls(lin.parameters(), lr=1e-4, **kw)
        before = lin.weight.data.clone()
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(lin(torch.rand(4, 1024, device="cuda")), torch.zeros(4, dtype=torch.long, device="cuda"))
        loss.backward()
        opt.step()
        assert lin.weight.grad is not None
        assert not torch.equal(lin.weight.data, before), cls.__name__
        with pytest.raises(ValueError):
            cls(lin.parameters(), lr=-1.0)
        with pytest.raises(ValueError):
            cls(lin.parameters(), betas=(1.1, 0.9))


def test_functional_apis():
    torch.manual_seed(0)
    ps = [torch.randn(100, device="cuda"), torch.randn(7, 9, device="cuda")]
    gs = [torch.randn_like(p) for p in ps]
    ref_p = [p.cpu().clone() for p in ps]
    m = [torch.zeros_like(p) for p in ps]; v = [torch.zeros_like(p) for p in ps]
    hb.optim.adabelief(ps, gs, m, v, [], [1, 1], False, 0.9, 0.999, 1e-2, 0.0, 1e-8)
    for rp, g_ in zip(ref_p, gs):
        OO.adabelief_step(rp, g_.cpu(), torch.zeros_like(rp), torch.zeros_like(rp), 1, 1e-2, 0.9, 0.999, 1e-8)
    for a, b in zip(ps, ref_p):
        close(a, b)
    W = [0.9 / 0.1 * torch.ones(1, device="cuda") for _ in ps]
    ref2 = [p.cpu().clone() for p in ps]
    m = [torch.zeros_like(p) for p in ps]; v = [torch.zeros_like(p) for p in ps]
    hb.optim.tadam(ps, gs, m, v, [], W, [1, 1], False, 0.9, 0.999, 1e-2, 0.0, 1e-8, None)
    for rp, g_ in zip(ref2, gs):
        OO.tadam_step(rp, g_.cpu(), torch.zeros_like(rp), torch.zeros_like(rp), 0.9 / 0.1 * torch.ones(1), 1, 1e-2, 0.9, 0.999, 1e-8)
    for a, b in zip(ps, ref2):
        close(a, b)


def test_full_size_repvgg_a1_parameter_set_vs_oracle():
    """BASELINE config 3 optimizer: AdaBelief(lr=1e-3, betas=(0.95, 0.99), eps=1e-6) over RepVGG-A1's 208 tensors /
    31.4 M parameters; the CUDA step is compared with the oracle on every tensor (3 steps), plus channels_last
    parameter layouts and the step/linearity property p(lr=2a) - p0 == 2 (p(lr=a) - p0)."""
    torch.manual_seed(0)
    model = hb.models.repvgg_a1(num_classes=1000)
    shapes = [tuple(p.shape) for p in model.parameters()]
    assert len(shapes) == 208 and sum(torch.Size(s).numel() for s in shapes) > 31_000_000
    params = []
    for s in shapes:
        t = torch.randn(*s) * 0.05
        if len(s) == 4:
            t = t.contiguous(memory_format=torch.channels_last)
        params.append(t)
    dev = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    opt = hb.optim.AdaBelief(dev, lr=1e-3, betas=(0.95, 0.99), eps=1e-6, weight_decay=1e-2)
    cpu_m = [torch.zeros_like(p) for p in params]; cpu_s = [torch.zeros_like(p) for p in params]
    for step in range(1, 4):
        grads = [torch.randn_like(p) * 0.01 for p in params]
        for d, g_ in zip(dev, grads):
            d.grad = g_.cuda()
        opt.step()
        for p, g_, m, s in zip(params, grads, cpu_m, cpu_s):
            OO.adabelief_step(p, g_, m, s, step, 1e-3, 0.95, 0.99, 1e-6, 1e-2)
    worst = max(((d.detach().cpu() - p).abs().max() / (p.abs().max() + 1e-12)).item() for d, p in zip(dev, params))
    assert worst < 1e-5, worst
    # linearity in lr of a single step from identical state
    p0 = torch.randn(100_003, device="cuda")
    g0 = torch.randn_like(p0)
    outs = []
    for lr in (1e-3, 2e-3):
        p = torch.nn.Parameter(p0.clone()); p.grad = g0.clone()
        hb.optim.AdaBelief([p], lr=lr).step()
        outs.append(p.detach() - p0)
    close(outs[1], 2 * outs[0], 1e-3, 1e-6)   # differences of fp32 parameters: quantised at ulp(p) ~ 1e-7
