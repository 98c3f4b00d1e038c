cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nn_layers.py tests/test_gpu_conv_bn.py -q -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_zoo.py -q -k "rexnet" 2>&1 | tail -3
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from holocron_b200.nn._dwconv import dwconv2d
import torch.nn.functional as TF
torch.manual_seed(0)
for (n, c, h, w, s, pad, bias) in [(2, 96, 37, 29, 1, 1, False), (2, 40, 16, 16, 2, 1, True), (3, 176, 14, 14, 1, 1, False), (2, 24, 9, 7, 2, 1, False), (1, 8, 5, 4, 1, 1, True), (2, 32, 8, 8, 1, 0, False)]:
    x = torch.randn(n, c, h, w, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wt = torch.randn(c, 1, 3, 3, device="cuda", requires_grad=True)
    b = torch.randn(c, device="cuda", requires_grad=True) if bias else None
    y = dwconv2d(x, wt, b, s, pad)
    xr = x.detach().float().requires_grad_(True)
    ref = TF.conv2d(xr, wt, b, s, pad, 1, c)
    g = torch.randn_like(ref).bfloat16()
    y.backward(g)
    ref.backward(g.float())
    rel = lambda a, b_: ((a.float() - b_.float()).norm() / b_.float().norm()).item()
    print((n, c, h, w, s, pad, bias), "fwd", round(rel(y, ref), 5), "dx", round(rel(x.grad, xr.grad), 5))
    assert rel(y, ref) < 4e-3 and rel(x.grad, xr.grad) < 4e-3
print("dw quad ok")
PY
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-eager-baseline"
timeout 400 $B --model rexnet1_0x > gpurun_out/h_rex.json 2> gpurun_out/h_rex.err
python -c "
import json
d=json.load(open('gpurun_out/h_rex.json'))
print('rexnet', round(d['ms_per_step'],3), round(d['value'],1), {k[:10]:(v['ms'],v['frac']) for k,v in d['roofline']['per_family'].items()})
"
