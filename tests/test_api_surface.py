"""Drop-in boundary: names, parameter order, kinds and defaults of the public hot-path surface must equal the
reference's, recorded from the unmodified reference by `tests/golden/make_golden.py --api`. Annotations are not compared."""
import functools
import inspect
import json

import holocron_b200 as hb

from conftest import GOLDEN


def _describe(obj):
    target = obj.__init__ if inspect.isclass(obj) else obj
    out = []
    for name, p in inspect.signature(target).parameters.items():
        if name == "self":
            continue
        out.append([name, p.kind.name, None if p.default is inspect.Parameter.empty else repr(p.default)])
    return out


# deliberate, documented deviations: (path, parameter) -> our default
# no network: nothing to download, so the detectors default to an un-pretrained backbone
DEVIATIONS = {(f"models.detection.{n}", "pretrained_backbone"): "False" for n in ("yolov4", "yolov1", "yolov2")}
DEVIATIONS.update({(f"models.segmentation.{n}", "pretrained_backbone"): "False" for n in ("unet_rexnet13", "unet_tvvgg11", "unet_tvresnet34")})


def test_public_signatures_match_reference():
    ref = json.loads((GOLDEN / "api_signatures.json").read_text())
    assert len(ref) >= 45
    problems = []
    for path, sig in ref.items():
        *mods, name = path.split(".")
        try:
            obj = getattr(functools.reduce(getattr, mods, hb), name)
        except AttributeError:
            problems.append(f"{path}: missing")
            continue
        ours = _describe(obj)
        for p in sig:
            if DEVIATIONS.get((path, p[0])) is not None:
                p[2] = DEVIATIONS[(path, p[0])]
        # an optimizer may accept the reference's (Adam's) extra keyword switches through **kwargs
        ref_core = [p for p in sig if p[1] != "VAR_KEYWORD"]
        ours_core = [p for p in ours if p[1] != "VAR_KEYWORD"]
        if ours_core[:len(ref_core)] != ref_core and ref_core[:len(ours_core)] != ours_core:
            problems.append(f"{path}: reference {sig} != ours {ours}")
        elif len(ours_core) < len(ref_core) and not any(p[1] == "VAR_KEYWORD" for p in ours):
            problems.append(f"{path}: parameters {ref_core[len(ours_core):]} not accepted")
    assert not problems, "\n".join(problems)


def test_public_classes_keep_reference_bases_and_properties():
    """isinstance contract and public properties of the reference's classes (recorded by make_golden.py --api), e.g.
    ``isinstance(AdaBelief(...), torch.optim.Adam)`` (reference optim/adabelief.py:16) and ``DropBlock2d.drop_prob``
    (nn/modules/dropblock.py:33-35)."""
    ref = json.loads((GOLDEN / "api_classes.json").read_text())
    assert "optim.AdaBelief" in ref and "nn.DropBlock2d" in ref
    problems = []
    for path, want in ref.items():
        *mods, name = path.split(".")
        cls = getattr(functools.reduce(getattr, mods, hb), name)
        ours = {f"{b.__module__}.{b.__qualname__}" for b in cls.__mro__[1:]}
        missing = [b for b in want["torch_bases"] if b not in ours]
        if missing:
            problems.append(f"{path}: not a subclass of {missing}")
        for prop in want["properties"]:
            if not isinstance(inspect.getattr_static(cls, prop, None), property):
                problems.append(f"{path}: property {prop} missing")
    assert not problems, "\n".join(problems)


def _describe_state_dict(model):
    import hashlib
    sd = model.state_dict()
    lines = [f"{k}:{tuple(v.shape)}:{str(v.dtype).replace('torch.', '')}" for k, v in sd.items()]
    h_vals = hashlib.sha1()
    for v in sd.values():
        h_vals.update(v.detach().contiguous().cpu().numpy().tobytes())
    return {"entries": len(lines), "numel": int(sum(v.numel() for v in sd.values())),
            "layout_sha1": hashlib.sha1("\n".join(lines).encode()).hexdigest(), "values_sha1": h_vals.hexdigest(),
            "first": lines[:3], "last": lines[-3:]}


def test_state_dict_layout_and_seeded_init_match_reference():
    """Checkpoint compatibility and init RNG order for EVERY factory: key order, shapes, dtypes, and - bit for bit - the
    parameter / buffer values produced under torch.manual_seed(0), against hashes recorded from the reference."""
    import torch
    ref = json.loads((GOLDEN / "state_dicts.json").read_text())
    # 17 + 8 (ResNet family) + 4 (MobileOne) + 1 (Res2Net) + 3 (SKNet) + 9 (ConvNeXt) classification factories, yolov4, unet3p, yolov1, yolov2, 7 U-Nets
    assert len(ref) >= 56
    for name, want in ref.items():
        torch.manual_seed(0)
        if name == "yolov4":
            m = hb.models.yolov4(pretrained_backbone=False, num_classes=80)
        elif name == "unet3p":
            m = hb.models.unet3p(num_classes=21)
        elif name in ("yolov1", "yolov2"):
            m = getattr(hb.models, name)(num_classes=20)
        elif name.startswith("unet"):
            m = getattr(hb.models, name)(num_classes=5)
        else:
            m = getattr(hb.models, name)(num_classes=10)
        got = _describe_state_dict(m)
        assert got == want, (name, got, want)
