#!/usr/bin/env python
"""Evaluation driver with the contract of the reference ``CODON_X{4,8,16}/test.py``.

Same five flags (``--batchSize --cuda --threads --pretrained --gpus``, test.py:21-26), same public
functions (``main``, ``test(model)``, ``EvaluationResults(depth_high, output)``), same per-image
pipeline (test.py:116-145): grayscale read, /255, forward, clip to [0,1], x255, truncate to uint8,
write the PNG, masked RMSE against the label, Gaussian SSIM, ``print(name, rmse, ssim)`` per image
and the two means at the end.  The reference hard-codes its data paths (test.py:70-73) and the
checkpoint name (:56) and is broken as shipped; here they are flags:

  --scale {4,8,16}   which CODONNet (default 4)          --mode fp16|bf16|tf32|fp32 (default fp16 = .half())
  --input-depth DIR  pre-upsampled depth PNGs             --input-color DIR   guide images
  --label DIR        ground truth                         --fix DIR           hole-filled GT for SSIM (default: label)
  --out DIR          result PNGs (default CODON_result_save/)   --logfile F   tee of stdout (default ./test_sintel.txt)
  --pretrained FILE  X4.pth-style checkpoint; without it seeded synthetic weights are used (and said so)
  --lr-depth DIR     TRUE low-resolution depth PNGs (H/s x W/s): uploaded as uint8 (s^2 fewer bytes) and upsampled
                     to the guide's size on the GPU (cv2.INTER_CUBIC semantics) instead of reading the offline
                     "Bicubic/X4" images of --input-depth (test.py:77)

Everything between the PNG decode and the PNG encode runs on the GPU: the images are uploaded as uint8, the colour
guide is converted to gray with the arithmetic of ``cv2.imread(path, 0)`` (test.py:118), ``/255`` (test.py:122-123),
the optional bicubic up-sampling, the forward, the quantisation, RMSE and SSIM.  Images are sharded over the GPUs named by ``--gpus`` (one host thread per GPU), or over
ranks when launched under torchrun; the sums are combined with one all-reduce.
"""
from __future__ import annotations

import argparse
import os
import random
import sys

import numpy as np
import torch

from . import engine as _eng
from . import scheduler as _sched
from .Loger import Logger
from .ssim_2 import ssim_exact  # noqa: F401  (re-exported, as the reference's `from ssim_2 import *`)

parser = argparse.ArgumentParser(description="CODON guided depth super-resolution: evaluation (B200 engine)")
parser.add_argument("--batchSize", type=int, default=64, help="kept for compatibility (unused, as in the reference)")
parser.add_argument("--cuda", default=True, action="store_true", help="Use cuda? (always; there is no CPU path)")
parser.add_argument("--threads", type=int, default=1, help="kept for compatibility (unused)")
parser.add_argument("--pretrained", default="", type=str, help="path to X4.pth / X8.pth / X16.pth")
parser.add_argument("--gpus", default="0", type=str, help="comma-separated gpu ids (default: 0)")
parser.add_argument("--scale", type=int, default=4, choices=[4, 8, 16])
parser.add_argument("--mode", default="fp16", choices=sorted(_eng.MODES))
parser.add_argument("--input-depth", default=None)
parser.add_argument("--input-color", default=None)
parser.add_argument("--label", default=None)
parser.add_argument("--fix", default=None)
parser.add_argument("--lr-depth", default=None, help="directory of true low-resolution depth PNGs (upsampled on the GPU)")
parser.add_argument("--out", default="CODON_result_save/")
parser.add_argument("--logfile", default="./test_sintel.txt", help="tee of stdout (reference: ./test_sintel.txt); empty = none")
parser.add_argument("--seed", type=int, default=None)
parser.add_argument("--weights-seed", type=int, default=0, help="seed of the synthetic weights used without --pretrained")

opt = None
model = None


def _imread_gray(path):
    import cv2
    img = cv2.imread(path, 0)                       # test.py:116-118
    if img is None:
        raise FileNotFoundError(path)
    return img


def _imread_any(path):
    """uint8 [H,W] (single-channel file) or [H,W,3] BGR (colour file): the bytes cv2 decodes, no conversion."""
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.dtype != np.uint8:
        return _imread_gray(path)
    if img.ndim == 3 and img.shape[2] == 4:
        img = np.ascontiguousarray(img[:, :, :3])
    return img


def _imwrite(path, img):
    import cv2
    cv2.imwrite(path, img)


def EvaluationResults(depth_high, output):
    """Masked RMSE in grey levels (reference test.py:148-164), computed by codon_masked_rmse."""
    lab = torch.as_tensor(np.ascontiguousarray(depth_high))
    out = torch.as_tensor(np.ascontiguousarray(output))
    if lab.dtype != torch.uint8 or out.dtype != torch.uint8:
        # the reference accepts any numeric arrays; the GPU kernel is integer-exact on uint8 images
        lab, out = lab.to(torch.float64), out.to(torch.float64)
        if not (torch.equal(lab, lab.round()) and torch.equal(out, out.round()) and 0 <= float(lab.min())
                and float(lab.max()) <= 255 and 0 <= float(out.min()) and float(out.max()) <= 255):
            raise _eng.CodonError("EvaluationResults expects 8-bit images")
        lab, out = lab.to(torch.uint8), out.to(torch.uint8)
    return float(_eng.masked_rmse(lab.cuda()[None], out.cuda()[None])[0])


def _build_model(scale):
    from importlib import import_module
    return import_module(f"codon_b200.CODON_x{scale}").CODONNet()


def _process(net, gpu, job, o):
    """One image on one GPU: returns (name, rmse, ssim)."""
    name, depth_path, gray_path, label_path, fix_path = job
    dev = torch.device("cuda", gpu)
    label_u8 = _imread_gray(label_path)
    fix_u8 = _imread_gray(fix_path) if fix_path != label_path else label_u8
    io_dtype = {"fp16": torch.float16, "bf16": torch.bfloat16}.get(o.mode, torch.float32)
    # guide: the decoded bytes go up as uint8; a colour file is converted to gray on the GPU with the arithmetic of
    # cv2.imread(path, 0) (test.py:118), then /255 (test.py:123)
    guide = torch.from_numpy(_imread_any(gray_path)).to(dev)
    gray_u8 = _eng.bgr_to_gray_u8(guide, "imread") if guide.dim() == 3 else guide
    y = _eng.u8_to_unit_f32(gray_u8)[None, None].to(io_dtype)
    H, W = gray_u8.shape
    if getattr(o, "lr_depth", None):
        # true LR depth (H/s x W/s uint8): /255 and bicubic up-sampling to the guide's size on the GPU
        lr = torch.from_numpy(_imread_gray(os.path.join(o.lr_depth, name))).to(dev)
        x = _eng.bicubic_upsample(_eng.u8_to_unit_f32(lr)[None], H, W).clamp_(0.0, 1.0)[None].to(io_dtype)
    else:
        depth_u8 = torch.from_numpy(_imread_gray(depth_path)).to(dev)                    # test.py:116
        x = _eng.u8_to_unit_f32(depth_u8)[None, None].to(io_dtype)                       # test.py:122
    with torch.no_grad():
        out = net.engine(dev).forward(x, y)                                              # test.py:125
    # np.clip(out,0,1); (out*255).astype(uint8), evaluated in the output dtype (test.py:127-132)
    out_u8 = _eng.quantise_u8(out[0, 0].float(), via_half=(io_dtype == torch.float16))
    h, w = out_u8.shape
    lab = torch.from_numpy(label_u8).to(dev)[:h, :w].contiguous()
    fix = torch.from_numpy(fix_u8).to(dev)[:h, :w].contiguous()
    rmse = float(_eng.masked_rmse(lab[None], out_u8[None])[0])                           # test.py:138
    ssim = float(_eng.ssim_gauss(fix[None], out_u8[None])[0])                            # test.py:139
    if o.out:
        os.makedirs(o.out, exist_ok=True)
        _imwrite(os.path.join(o.out, name), out_u8.cpu().numpy())                        # test.py:137
    return name, rmse, ssim


def test(model, options=None):
    """Runs the evaluation loop (reference test.py:64-145); returns (mean_rmse, mean_ssim, n)."""
    o = options or opt
    if o is None:
        o = parser.parse_args([])
    model.eval()                                                                         # test.py:67
    if not (o.input_depth and o.input_color and o.label):
        raise _eng.CodonError("test(): --input-depth, --input-color and --label are required "
                              "(the reference hard-codes them, test.py:70-73)")
    names = sorted(os.listdir(o.input_color))                                            # test.py:100-109
    fix_dir = o.fix or o.label
    jobs = [(n, os.path.join(o.input_depth, n), os.path.join(o.input_color, n), os.path.join(o.label, n),
             os.path.join(fix_dir, n)) for n in names]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        gpus = [int(os.environ.get("LOCAL_RANK", "0"))]
        jobs_here = [jobs[i] for i in _sched.shard_indices(len(jobs), world, rank)]
    else:
        gpus = [int(g) for g in str(o.gpus).split(",") if g.strip() != ""] or [0]
        n_vis = torch.cuda.device_count()
        gpus = [g for g in gpus if g < n_vis] or [0]
        jobs_here = jobs
    results = _sched.MultiGpuExecutor(gpus).map(lambda gpu, job: _process(model, gpu, job, o), jobs_here)
    rmse_sum = ssim_sum = 0.0
    for name, rmse, ssim in results:
        rmse_sum += rmse
        ssim_sum += ssim
        print(name, rmse, ssim)                                                          # test.py:142
    rmse_sum, ssim_sum, n = _sched.reduce_metric_sums(rmse_sum, ssim_sum, len(results))
    if rank == 0:
        print(n)                                                                         # test.py:144
        print(rmse_sum / max(n, 1), ssim_sum / max(n, 1))                                # test.py:145
    return rmse_sum / max(n, 1), ssim_sum / max(n, 1), n


def main(argv=None):
    global opt, model
    opt = parser.parse_args(argv)
    print(opt)
    if not torch.cuda.is_available():
        raise Exception("No GPU found or Wrong gpu id: codon_b200 has no CPU path")      # test.py:37-38
    print("=> use gpu id: '{}'".format(opt.gpus))
    opt.seed = opt.seed if opt.seed is not None else random.randint(1, 10000)            # test.py:39-43
    print("Random Seed: ", opt.seed)
    torch.manual_seed(opt.seed)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    print("===> Building model")
    model = _build_model(opt.scale)                                                      # test.py:48
    model.set_mode(opt.mode)
    if opt.logfile and int(os.environ.get("RANK", "0")) == 0:
        sys.stdout = Logger(opt.logfile)                                                     # test.py:53
    if opt.pretrained:
        from .checkpoint import load_checkpoint
        sd, meta = load_checkpoint(opt.pretrained)                                       # test.py:56-59
        opt.start_epoch = int(meta.get("epoch", 0)) + 1
        from .checkpoint import infer_scale
        # x4/x8 checkpoints carry the never-called attention_c5 / attention_s5 (CODON_x4.py:64-65), x16 ones do not
        # (SURVEY.md 5.4): the checkpoint must be of the family --scale names, and then every key must match
        family = infer_scale(sd)
        want = "x16" if opt.scale == 16 else "x4/x8"
        if family != want:
            raise _eng.CodonError(f"'{opt.pretrained}' is an {family} checkpoint but --scale {opt.scale} was given "
                                  f"({'no ' if want == 'x16' else ''}attention_c5 / attention_s5 keys expected)")
        model.load_state_dict(sd, strict=True)
        print(f"=> loaded '{opt.pretrained}' (epoch {meta.get('epoch', '?')})")
    else:
        from .synthetic import synthetic_state_dict
        model.load_state_dict(synthetic_state_dict(opt.scale, opt.weights_seed))
        print(f"=> no --pretrained given: SYNTHETIC weights seed {opt.weights_seed} (image quality is meaningless)")
    res = test(model, opt)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return res


if __name__ == "__main__":
    main()
