"""ONNX export of a zoo model — same command line as the reference's scripts/export_to_onnx.py.

    python scripts/export_to_onnx.py repvgg_a0 --checkpoint repvgg_a0.pth --path model.onnx

The model is built with the reference's parameter names, optionally loaded from a ``state_dict`` file, put in inference
form (RepVGG / MobileOne are re-parametrised) and serialised by :func:`holocron_b200.onnx.export_onnx` (opset 14, static
input shape). Runs on the host: the exporter traces over fake tensors and needs no GPU."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from holocron_b200 import models  # noqa: E402
from holocron_b200.onnx import export_onnx  # noqa: E402


@torch.inference_mode()
def main(args):
    if args.pretrained and not isinstance(args.checkpoint, str):
        raise SystemExit("--pretrained needs network access; pass --checkpoint <state_dict file> instead")
    model = models.__dict__[args.arch](pretrained=False).eval()
    if isinstance(args.checkpoint, str):
        model.load_state_dict(torch.load(args.checkpoint, map_location="cpu"), strict=True)
    if args.arch.startswith("repvgg") or args.arch.startswith("mobileone"):
        model.reparametrize()
    export_onnx(model, (args.batch_size, args.in_channels, args.height, args.width), args.path, opset_version=14)
    print(f"{args.arch} -> {args.path} ({os.path.getsize(args.path)} bytes)")


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Holocron model ONNX export",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("arch", type=str, help="Architecture to use")
    parser.add_argument("--height", type=int, default=224, help="The height of the input image")
    parser.add_argument("--width", type=int, default=224, help="The width of the input image")
    parser.add_argument("--in-channels", type=int, default=3, help="The number of channels of the input image")
    parser.add_argument("--batch-size", type=int, default=1, help="The batch size used for the model")
    parser.add_argument("--path", type=str, default="./model.onnx", help="The path of the output file")
    parser.add_argument("--checkpoint", type=str, default=None, help="The checkpoint to restore")
    parser.add_argument("--pretrained", dest="pretrained", help="Use pre-trained models from the modelzoo", action="store_true")
    main(parser.parse_args())
