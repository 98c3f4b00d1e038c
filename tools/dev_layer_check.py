"""Dev tool (GPU): run a zoo model and check EVERY convolution launch of the forward/backward pass against
torch.nn.functional.conv2d on the same bf16 operands; prints the launches whose rel-L2 error exceeds 1e-2."""
import sys

import torch

sys.path.insert(0, ".")
import torch.nn.functional as TF

import holocron_b200 as hb
from holocron_b200.nn import _fused

name = sys.argv[1] if len(sys.argv) > 1 else "cspdarknet53"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
orig = _fused.conv2d_forward_raw
bad = []
count = [0]


def checked(x, wf, cout, r, s, stride, pad, dil, bias=None, residual=None, act=0):
    y = orig(x, wf, cout, r, s, stride, pad, dil, bias=bias, residual=residual, act=act)
    w = wf[:cout].permute(0, 3, 1, 2).float()
    ref = TF.conv2d(x.float(), w, None if bias is None else bias[:cout].float(), stride, pad, dil)
    if residual is not None:
        ref = ref + residual.float()
    if act == 1:
        ref = ref.relu()
    err = ((y.float() - ref).norm() / (ref.norm() + 1e-20)).item()
    count[0] += 1
    if not (err < 1e-2) or act > 1:
        bad.append((tuple(x.shape), tuple(wf.shape), cout, stride, pad, dil, act, err))
    return y


_fused.conv2d_forward_raw = checked
orig_bn = _fused.bn_act
bn_bad = []
bn_count = [0]


def act_ref(z, act, slope):
    if act == 1: return z.relu()
    if act == 2: return z.clamp(0, 6)
    if act == 3: return TF.silu(z)
    if act == 4: return TF.leaky_relu(z, slope)
    if act == 5: return z * torch.tanh(TF.softplus(z))
    if act == 6: return 0.5 * z * (z + 2).clamp(0, 2)
    return z


def checked_bn(us, bns, act=0, slope=0.0, residual=None, training=None, res_after_act=False):
    out = orig_bn(us, bns, act, slope, residual, training, res_after_act)
    tr = bns[0].training if training is None else training
    z = 0
    for u, bn in zip(us, bns):
        uf = u.detach().float()[:, :bn.num_features]
        if tr:
            mean = uf.mean((0, 2, 3), keepdim=True); var = uf.var((0, 2, 3), unbiased=False, keepdim=True)
        else:
            mean = bn.running_mean.view(1, -1, 1, 1); var = bn.running_var.view(1, -1, 1, 1)
        z = z + (uf - mean) / torch.sqrt(var + bn.eps) * bn.weight.detach().view(1, -1, 1, 1) + bn.bias.detach().view(1, -1, 1, 1)
    c = bns[0].num_features
    if residual is not None and not res_after_act:
        z = torch.maximum(z, residual.detach().float()[:, :c]) if act == 7 else z + residual.detach().float()[:, :c]
    ref = act_ref(z, act, slope)
    if residual is not None and res_after_act:
        ref = ref + residual.detach().float()[:, :c]
    err = ((out.detach().float()[:, :c] - ref).norm() / (ref.norm() + 1e-20)).item()
    bn_count[0] += 1
    if not (err < 2e-2):
        bn_bad.append((tuple(us[0].shape), len(us), act, residual is not None, res_after_act, tr, err))
    return out


_fused.bn_act = checked_bn

from holocron_b200.models import _blocks
orig_unit = _blocks.conv_bn_act
unit_bad = []
unit_count = [0]


def checked_unit(x, conv, bn, act, residual=None, res_after_act=False, keep_padded=False):
    out = orig_unit(x, conv, bn, act, residual, res_after_act, keep_padded)
    with torch.no_grad():
        cin = conv.in_channels
        xf = x.detach().float()[:, :cin]
        # same operand rounding as the kernels: bf16 inputs and filters, fp32 accumulation
        z = TF.conv2d(xf.to(torch.bfloat16).float(), conv.weight.detach().to(torch.bfloat16).float(),
                      None if conv.bias is None else conv.bias.detach().float(), conv.stride, conv.padding,
                      conv.dilation, conv.groups)
        if bn is not None:
            z = z.to(torch.bfloat16).float()
            if bn.training:
                mean = z.mean((0, 2, 3), keepdim=True); var = z.var((0, 2, 3), unbiased=False, keepdim=True)
            else:
                mean = bn.running_mean.view(1, -1, 1, 1); var = bn.running_var.view(1, -1, 1, 1)
            z = (z - mean) / torch.sqrt(var + bn.eps) * bn.weight.detach().view(1, -1, 1, 1) + bn.bias.detach().view(1, -1, 1, 1)
        code, slope = _fused.act_code(act)
        c = conv.out_channels
        if residual is not None and not res_after_act:
            z = z + residual.detach().float()[:, :c]
        ref = act_ref(z, code, slope)
        if residual is not None and res_after_act:
            ref = ref + residual.detach().float()[:, :c]
        err = ((out.detach().float()[:, :c] - ref).norm() / (ref.norm() + 1e-20)).item()
    unit_count[0] += 1
    if not (err < 3e-2):
        unit_bad.append((tuple(x.shape), tuple(conv.weight.shape), conv.stride, conv.padding, conv.groups, bn is not None,
                         code, residual is not None, res_after_act, keep_padded, tuple(out.shape), err))
    return out


for _name, _mod in list(sys.modules.items()):
    if _name.startswith("holocron_b200.models") and getattr(_mod, "conv_bn_act", None) is orig_unit:
        _mod.conv_bn_act = checked_unit
torch.manual_seed(0)
m = getattr(hb.models, name)(num_classes=10).cuda().train()
if len(sys.argv) > 3:
    import os
    g = torch.load(os.path.join("tests", "golden", "zoo.pt"))[name]
    print("golden logits", g["logits"][0, :5].tolist())
torch.manual_seed(1)
x = torch.rand(2, 3, size, size).cuda()
out = m(x)
out.sum().backward()
print(name, "conv launches checked:", count[0], "suspicious:", len(bad))
for b in bad[:40]:
    print("  ", b)
print("bn_act calls checked:", bn_count[0], "suspicious:", len(bn_bad))
for b in bn_bad[:40]:
    print("  ", b)
print("conv_bn_act units checked:", unit_count[0], "suspicious:", len(unit_bad))
for b in unit_bad[:40]:
    print("  ", b)
print("logits", out[0, :5].tolist())
