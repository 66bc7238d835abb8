#!/usr/bin/env python
"""Video training entry point on the B200 engine — same CLI as the reference's train_clip2.py (flags :404-489,
method dispatch :258-321, train loop :26-126, optimizer :215-236, poly LR :239-252, checkpoint/resume :179-189,
:347-357) for the two methods on the TCB hot path: ``--method clip_psp`` and ``--method clip_ocr``.

Differences that are the point of this repo:
  * multi-GPU is one process per GPU under torchrun (`python -m torch.distributed.run --nproc-per-node N
    train_clip2.py ...`), clips sharded over ranks, one NCCL gradient all-reduce per step (+ SyncBN statistics with
    ``--syncbn True``) instead of nn.DataParallel; ``--gpu_num`` is checked against WORLD_SIZE;
  * ``--dataroot`` reads VSPW clips with ``vspw_data.VSPWClipTrain`` (bit-identical to the reference's
    ``dataset2.BaseDataset_longclip`` under the same RNG seeds); ``--synthetic True`` feeds seeded synthetic clips instead;
  * ``--precision {bf16x3,bf16,fp32}`` selects the convolution arithmetic (default bf16x3 = parity mode).
"""
import argparse
import gc
import os
import time
from collections import OrderedDict

import torch
import torch.nn as nn

from cvpr2021_vspw_implement_b200 import engine as E
from cvpr2021_vspw_implement_b200 import parallel as P
from cvpr2021_vspw_implement_b200.config import cfg
from cvpr2021_vspw_implement_b200.data import DevicePrefetcher, SyntheticClipTrain
from cvpr2021_vspw_implement_b200.models import Clip_PSP, ClipOCRNet, ModelBuilder, Non_local3d
from cvpr2021_vspw_implement_b200.utils import AverageMeter, parse_devices, setup_logger

OTHER_METHODS = ["netwarp", "ETC", "tdnet", "our_warp", "propnet", "our_warp_merge", "netwarp_ocr", "etc_ocr"]


def build_batch(clip_imgs, clip_gts, it_, method="clip_psp"):
    """train_clip2.py:54-83: clip_psp / clip_ocr take frame 0 of the sampled clip as the current frame; nonlocal3d gets the
    whole clip and supervises every frame."""
    if method == "nonlocal3d":
        return {"clipimgs_data": list(clip_imgs), "cliplabels_data": list(clip_gts), "step": it_}
    return {"img_data": clip_imgs[0], "seg_label": clip_gts[0], "clipimgs_data": list(clip_imgs[1:]),
            "cliplabels_data": list(clip_gts[1:]), "step": it_}


def train(segmentation_module, data_loader, optimizer, bucket, history, epoch, cfg, args, device, rank=0):
    batch_time, data_time = AverageMeter(), AverageMeter()
    ave_total_loss, ave_acc = AverageMeter(), AverageMeter()
    segmentation_module.train(not cfg.TRAIN.fix_bn)
    epoch_iters = len(data_loader)
    max_iters = epoch_iters * cfg.TRAIN.num_epoch
    tic = time.time()
    # The cyclic GC is paused inside the step loop (a gen-2 sweep over a step's tape closures stalls the launching thread
    # for tens of ms and idles the GPU); everything a step allocates is freed by reference counting after backward.
    gc.collect()
    gc.disable()
    # pinned batches are copied to the device one step ahead on a side stream (the reference does a blocking .cuda())
    for i, (clip_imgs, clip_gts) in enumerate(DevicePrefetcher(data_loader, device)):
        batch_data = build_batch(clip_imgs, clip_gts, i + 1, args.method)
        data_time.update(time.time() - tic)
        bucket.zero_grad()  # one memset; every p.grad is a slice of the flat gradient bucket the backward kernels write into
        adjust_learning_rate(optimizer, i + (epoch - 1) * epoch_iters, cfg, max_iters, args)
        loss, acc = segmentation_module(batch_data)
        loss, acc = loss.mean(), acc.mean()
        loss.backward()
        bucket.all_reduce_mean()
        optimizer.step()
        loss_v, acc_v = P.mean_scalar(loss).item(), P.mean_scalar(acc).item()  # one host sync per step, as the reference
        batch_time.update(time.time() - tic)
        tic = time.time()
        ave_total_loss.update(loss_v)
        ave_acc.update(acc_v * 100)
        if rank == 0:
            print('Epoch: [{}][{}/{}], Time: {:.2f}, Data: {:.2f}, lr_encoder: {:.6f}, lr_decoder: {:.6f}, '
                  'Accuracy: {:4.2f}, Loss: {:.6f}'.format(epoch, i, epoch_iters, batch_time.average(), data_time.average(),
                                                           cfg.TRAIN.running_lr_encoder, cfg.TRAIN.running_lr_decoder,
                                                           ave_acc.average(), ave_total_loss.average()))
        history['train']['epoch'].append(epoch - 1 + 1. * i / epoch_iters)
        history['train']['loss'].append(loss_v)
        history['train']['acc'].append(acc_v)
        if (i + 1) % 500 == 0:
            gc.collect()
        if args.max_iters_per_epoch and i + 1 >= args.max_iters_per_epoch:
            break
    gc.enable()


def checkpoint(opt, nets, history, args, epoch):
    """model_epoch_E.pth / opt_epoch_E.pth (train_clip2.py:179-189).  Keys carry the 'module.' prefix the reference's
    DataParallel checkpoints have, because both resume (:350-353) and test_clip2.py (:267-269) strip 7 characters."""
    print('Saving checkpoints...')
    os.makedirs(args.saveroot, exist_ok=True)
    sd = OrderedDict(("module." + k, v) for k, v in nets.state_dict().items())
    torch.save(sd, '{}/model_epoch_{}.pth'.format(args.saveroot, epoch))
    torch.save(opt.state_dict(), '{}/opt_epoch_{}.pth'.format(args.saveroot, epoch))


def _params(gen, seen, keep_duplicates):
    out = []
    for p in gen:
        if keep_duplicates or id(p) not in seen:
            seen.add(id(p))
            out.append(p)
    return out


def create_optimizers(model, cfg, args):
    """SGD with the reference's four parameter groups (train_clip2.py:215-236).  The reference's generators yield
    every parameter 2-5 times (quirk Q10).  By default the duplicates are dropped: ONE update per parameter per step.  That is
    a deliberate difference in the training trajectory — under torch 1.3 (no duplicate check) the reference applies the SGD
    update k times per step to a parameter yielded k times, i.e. an effective learning rate of ~k x lr with compounding
    momentum; torch >= 2 refuses such groups in its fused/foreach paths.  ``--keep_duplicate_params True`` passes the repeats
    through unchanged (torch.optim.SGD, the reference's trajectory and optimizer-checkpoint layout); without it a reference
    ``opt_epoch_E.pth`` is remapped onto the de-duplicated groups on resume (`remap_reference_optimizer_state`)."""
    seen, kd = set(), args.keep_duplicate_params
    wd = cfg.TRAIN.weight_decay
    if args.fix:
        groups = [{'params': _params(model.get_10x_lr_params(), seen, kd), 'lr': args.lr, 'weight_decay': wd},
                  {'params': _params(model.get_10x_lr_params_bias(), seen, kd), 'lr': args.lr, 'weight_decay': 0}]
    else:
        groups = [{'params': _params(model.get_1x_lr_params(), seen, kd), 'lr': args.lr * 0.1, 'weight_decay': wd},
                  {'params': _params(model.get_10x_lr_params(), seen, kd), 'lr': args.lr, 'weight_decay': wd},
                  {'params': _params(model.get_1x_lr_params_bias(), seen, kd), 'lr': args.lr * 0.1, 'weight_decay': 0},
                  {'params': _params(model.get_10x_lr_params_bias(), seen, kd), 'lr': args.lr, 'weight_decay': 0}]
    if getattr(args, "fused_sgd", False) and not kd:
        from cvpr2021_vspw_implement_b200.optim import FusedSGD  # same update rule and state_dict, one launch per step
        return FusedSGD(groups, lr=args.lr, momentum=cfg.TRAIN.beta1, weight_decay=wd)
    return torch.optim.SGD(groups, lr=args.lr, momentum=cfg.TRAIN.beta1, weight_decay=wd)


def remap_reference_optimizer_state(state, optimizer):
    """Make an ``opt_epoch_E.pth`` written by the REFERENCE loadable into this entry point's de-duplicated optimizer.

    The reference's parameter groups repeat every parameter 2-5 times (quirk Q10), and torch's `Optimizer.state_dict()` keeps
    those repeats in ``param_groups[i]['params']`` (positions are counted WITH the repeats: [0, 1, 0, 1] then the next group
    starts at 4; torch 1.3 stored `id(p)` values instead of positions — any hashable key works here).  The momentum buffers
    are keyed by the first key of each parameter.  Mapping: the k-th DISTINCT key of reference group g <-> the k-th parameter
    of our group g.  A state dict without repeats (one written by this entry point) is returned unchanged."""
    groups = state["param_groups"]
    ours = optimizer.state_dict()["param_groups"]
    if len(groups) != len(ours):
        raise ValueError(f"optimizer checkpoint has {len(groups)} parameter groups, this run has {len(ours)} "
                         f"(was it written with a different --fix setting?)")
    if all(len(set(g["params"])) == len(g["params"]) for g in groups) and [len(g["params"]) for g in groups] == [len(g["params"]) for g in ours]:
        return state
    key_map, new_groups = {}, []
    for g_ref, g_our in zip(groups, ours):
        distinct = list(dict.fromkeys(g_ref["params"]))
        if len(distinct) != len(g_our["params"]):
            raise ValueError(f"optimizer checkpoint group holds {len(distinct)} distinct parameters, this run's group {len(g_our['params'])}: "
                             f"not the same model / --method; with --keep_duplicate_params True the file is loaded as it is")
        key_map.update(dict(zip(distinct, g_our["params"])))
        ng = dict(g_ref)
        ng["params"] = list(g_our["params"])
        new_groups.append(ng)
    new_state = {key_map[k]: v for k, v in state["state"].items() if k in key_map}
    return {"state": new_state, "param_groups": new_groups}


def adjust_learning_rate(optimizer, cur_iter, cfg, max_iters, args):
    scale_running_lr = ((1. - float(cur_iter) / max_iters) ** cfg.TRAIN.lr_pow)
    cfg.TRAIN.running_lr_encoder = args.lr * scale_running_lr
    cfg.TRAIN.running_lr_decoder = args.lr * scale_running_lr
    lr = cfg.TRAIN.running_lr_encoder
    scales = [1.0, 1.0] if args.fix else [0.1, 1.0, 0.1, 1.0]
    for g, s in zip(optimizer.param_groups, scales):
        g['lr'] = lr * s


def build_module(cfg, args):
    net_encoder = ModelBuilder.build_encoder(arch=cfg.MODEL.arch_encoder.lower(), fc_dim=cfg.MODEL.fc_dim,
                                             weights=cfg.MODEL.weights_encoder, args=args)
    crit = nn.NLLLoss(ignore_index=255)
    if args.method == 'clip_psp':
        return Clip_PSP(net_encoder, crit, args, deep_sup_scale=0.4)
    if args.method == 'clip_ocr':
        return ClipOCRNet(net_encoder, crit, args, deep_sup_scale=0.4)
    if args.method == 'nonlocal3d':
        return Non_local3d(args, net_encoder, crit)
    # the other methods of the reference are outside the TCB hot path this engine implements
    raise NotImplementedError(f"--method {args.method!r}: only clip_psp / clip_ocr / nonlocal3d run on the B200 engine")


def make_loader(args, world, rank):
    per_rank = args.batchsize // world
    if args.batchsize % world:
        raise ValueError(f"--batchsize {args.batchsize} must divide over {world} ranks")
    if args.synthetic:
        h, w = (int(x) for x in args.synthetic_size.lower().split("x"))
        ds = SyntheticClipTrain(args, length=args.synthetic_clips, height=h, width=w, seed=cfg.TRAIN.seed)
    else:
        from cvpr2021_vspw_implement_b200.vspw_data import VSPWClipTrain  # = dataset2.BaseDataset_longclip (:852-1048)
        ds = VSPWClipTrain(args, 'train')
    sampler = None
    if world > 1:
        sampler = torch.utils.data.distributed.DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=True, drop_last=True)
    return torch.utils.data.DataLoader(ds, batch_size=per_rank, shuffle=sampler is None, sampler=sampler, num_workers=args.workers,
                                       drop_last=True, pin_memory=True)


def main(cfg, args):
    world, rank, local = P.init_from_env(device_offset=args.start_gpu)  # the NCCL communicator is bound to the compute device
    if args.gpu_num != world:
        raise ValueError(f"--gpu_num {args.gpu_num} but WORLD_SIZE={world}: launch one process per GPU with torchrun")
    device = torch.device("cuda", args.start_gpu + local)
    torch.cuda.set_device(device)
    E.set_precision(args.precision)
    syncbn = args.syncbn and world > 1
    E.set_syncbn(syncbn, clamp=args.syncbn_clamp, group=P.make_syncbn_group(args.syncbn_exchange) if syncbn else None)
    torch.manual_seed(cfg.TRAIN.seed)
    segmentation_module = build_module(cfg, args)
    loader_train = make_loader(args, world, rank)
    if rank == 0:
        print('1 Epoch = {} iters'.format(len(loader_train)))
    segmentation_module.cuda(device)
    optimizer = create_optimizers(segmentation_module, cfg, args)
    if args.resume_epoch != 0:
        to_load = torch.load(os.path.join('./resume', 'model_epoch_{}.pth'.format(args.resume_epoch)), map_location=device)
        segmentation_module.load_state_dict(OrderedDict((k[7:], v) for k, v in to_load.items()))  # strip 'module.' (:350-353)
        cfg.TRAIN.start_epoch = args.resume_epoch
        opt_state = torch.load(os.path.join('./resume', 'opt_epoch_{}.pth'.format(args.resume_epoch)), map_location=device)
        if not args.keep_duplicate_params:  # a reference checkpoint repeats every parameter 2-5 times per group (quirk Q10)
            opt_state = remap_reference_optimizer_state(opt_state, optimizer)
        optimizer.load_state_dict(opt_state)
        print('resume from epoch {}'.format(args.resume_epoch))
    P.broadcast_parameters(segmentation_module)
    bucket = P.GradBucket(segmentation_module.parameters())
    history = {'train': {'epoch': [], 'loss': [], 'acc': []}}
    for epoch in range(cfg.TRAIN.start_epoch, cfg.TRAIN.num_epoch):
        if rank == 0:
            print('Epoch {}'.format(epoch))
        if hasattr(loader_train.sampler, "set_epoch"):
            loader_train.sampler.set_epoch(epoch)
        train(segmentation_module, loader_train, optimizer, bucket, history, epoch + 1, cfg, args, device, rank)
        if (epoch + 1) % args.checkpoint_every == 0 and rank == 0 and args.saveroot:
            checkpoint(optimizer, segmentation_module, history, args, epoch + 1)
    if rank == 0:
        print('Training Done!')
    return history


def str2bool(v):
    if isinstance(v, bool):
        return v
    if v.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    raise argparse.ArgumentTypeError('Boolean value expected.')


def make_parser():
    parser = argparse.ArgumentParser(description="PyTorch Semantic Segmentation Training (B200 engine)")
    parser.add_argument("--cfg", default="config/vsp-resnet101dilated-ppm_deepsup_clip.yaml", metavar="FILE", type=str)
    parser.add_argument("--gpus", default="0-3", help="gpus to use, e.g. 0-3 or 0,1,2,3")
    parser.add_argument("--predir", default='../../ade20k-hrnetv2-c1')
    for name, typ, default in (("num_class", int, 124), ("batchsize", int, 16), ("workers", int, 0), ("start_gpu", int, 0),
                               ("gpu_num", int, 1), ("dataroot", str, ''), ("trainfps", int, 1), ("lr", float, 0.02),
                               ("multi_scale", str2bool, False), ("saveroot", str, ''), ("totalepoch", int, 30),
                               ("dataroot2", str, ''), ("usetwodata", str2bool, False), ("cropsize", int, 531),
                               ("validation", str2bool, True), ("lesslabel", str2bool, False), ("clip_num", int, 5),
                               ("dilation_num", int, 3), ("clip_up", str2bool, False), ("clip_middle", str2bool, False),
                               ("fix", str2bool, False), ("othergt", str2bool, False), ("propclip2", str2bool, False),
                               ("early_usecat", str2bool, False), ("earlyfuse", str2bool, False), ("weight_decay", float, 1e-4),
                               ("allsup", str2bool, False), ("allsup_scale", float, 0.3), ("deepsup_scale", float, 0.4),
                               ("linear_combine", str2bool, False), ("distsoftmax", str2bool, False),
                               ("distnearest", str2bool, False), ("temp", float, 3), ("max_distances", str, '10'),
                               ("pre_enc", str, ''), ("pre_dec", str, ''), ("dilation2", str, "2,5,9"), ("resume_epoch", int, 0),
                               ("clipocr_all", str2bool, False), ("use_memory", str2bool, False), ("memory_num", int, 8),
                               ("st_weight", float, 0.1), ("psp_weight", str2bool, False)):
        parser.add_argument("--" + name, type=typ, default=default)
    parser.add_argument("--method", type=str, default='', choices=['clip_psp', 'clip_ocr', 'nonlocal3d'] + OTHER_METHODS)
    # ---- flags of this engine (not in the reference) ----
    parser.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16", "fp32"])
    parser.add_argument("--syncbn", type=str2bool, default=True, help="all-reduce BN statistics over ranks (reference multi-GPU semantics)")
    parser.add_argument("--syncbn_exchange", default="peer", choices=["peer", "nccl"], help="peer = one-shot exchange over NVLink peer memory per BN layer (csrc/peer.cu); nccl = one library all-reduce per layer")
    parser.add_argument("--syncbn_clamp", type=str2bool, default=False, help="clamp(var,eps)^-1/2 as the reference's DataParallel SyncBN (quirk Q4)")
    parser.add_argument("--synthetic", type=str2bool, default=False)
    parser.add_argument("--synthetic_size", type=str, default="480x854")
    parser.add_argument("--synthetic_clips", type=int, default=64)
    parser.add_argument("--max_iters_per_epoch", type=int, default=0)
    parser.add_argument("--checkpoint_every", type=int, default=20)
    parser.add_argument("--keep_duplicate_params", type=str2bool, default=False)
    parser.add_argument("--fused_sgd", type=str2bool, default=True, help="vspw_sgd_momentum_step instead of torch.optim.SGD (same update rule)")
    parser.add_argument("opts", help="Modify config options using the command-line", default=None, nargs=argparse.REMAINDER)
    return parser


def configure(args):
    args.max_distances = [int(dd) for dd in args.max_distances.split(',')]
    cfg.merge_from_file(args.cfg)
    cfg.merge_from_list(args.opts)
    cfg.MODEL.weights_encoder = args.pre_enc
    cfg.MODEL.weights_decoder = args.pre_dec
    cfg.TRAIN.num_epoch = args.totalepoch
    cfg.TRAIN.max_iters = cfg.TRAIN.epoch_iters * cfg.TRAIN.num_epoch
    cfg.TRAIN.weight_decay = args.weight_decay
    cfg.TRAIN.lr_encoder = cfg.TRAIN.lr_decoder = args.lr
    cfg.TRAIN.running_lr_encoder = cfg.TRAIN.running_lr_decoder = args.lr
    return cfg


if __name__ == '__main__':
    args = make_parser().parse_args()
    configure(args)
    rank = int(os.environ.get("RANK", "0"))
    logger = setup_logger(distributed_rank=rank)
    logger.info("Loaded configuration file {}".format(args.cfg))
    logger.info("Running with config:\n{}".format(cfg))
    if rank == 0:
        os.makedirs(cfg.DIR, exist_ok=True)
        with open(os.path.join(cfg.DIR, 'config.yaml'), 'w') as f:
            f.write("{}".format(cfg))
        parse_devices(args.gpus)  # validated like the reference; placement comes from torchrun's LOCAL_RANK
        print(args)
    main(cfg, args)
