#!/usr/bin/env python
"""Video inference / evaluation entry point on the B200 engine — same CLI as the reference's test_clip2.py (flags
:349-405; per-video loop :280-321; metrics :323-331) for ``--method clip_psp`` and ``--method clip_ocr`` (per-frame loop
``test``, :28-89) and ``--method nonlocal3d`` (sliding-window loop ``test_all``, :90-195).

Per batch: ``scores = module(batch_data, segSize=(H, W))`` -> argmax -> Evaluator (global + per video) -> VC metric,
exactly the reference's sequence (test_clip2.py:28-89).  ``--dataroot`` reads the videos of ``--split`` with
``vspw_data.VSPWClipTest`` (= the reference's ``TestDataset_longclip``); ``--synthetic True`` evaluates seeded synthetic
videos instead, and then ``--load`` may be empty (random weights).
Replicas only: videos are independent and the OCR memory bank is per-video state, so multi-GPU inference is one
process per GPU over disjoint video lists.
"""
import argparse
import os
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from cvpr2021_vspw_implement_b200 import engine as E
from cvpr2021_vspw_implement_b200.config import cfg
from cvpr2021_vspw_implement_b200.data import SyntheticClipTest, SyntheticWindowTest
from cvpr2021_vspw_implement_b200.models import Clip_PSP, ClipOCRNet, ModelBuilder, Non_local3d
from cvpr2021_vspw_implement_b200.utils import Evaluator, get_common_device, setup_logger
from train_clip2 import OTHER_METHODS, str2bool


def vspw_palette():
    """The reference's palette (test_clip2.py:24): 22 VOC-style colours, then grey levels i,i,i for i >= 22."""
    voc = [0, 0, 0, 128, 0, 0, 0, 128, 0, 128, 128, 0, 0, 0, 128, 128, 0, 128, 0, 128, 128, 128, 128, 128, 64, 0, 0, 191, 0, 0,
           64, 128, 0, 191, 128, 0, 64, 0, 128, 191, 0, 128, 64, 128, 128, 191, 128, 128, 0, 64, 0, 128, 64, 0, 0, 191, 0,
           128, 191, 0, 0, 64, 128, 128, 64, 128]
    return voc + [c for i in range(22, 256) for c in (i, i, i)]


def test(segmentation_module, loader, gpu, args, evaluator, eval_video, video):
    """Per-frame loop (test_clip2.py:28-89).  Labels and argmax predictions stay on the device: the confusion matrices are
    accumulated there (vspw_confusion_add) and the returned lists hold CUDA tensors for the on-device VC metric
    (utils.get_common_device); a prediction only crosses to the host when --is_save asks for its PNG."""
    segmentation_module.eval()
    gtlist_, predlist_ = [], []
    h = w = 0
    for i, data in enumerate(loader):
        imgs, gts, clip_imgs, _, gtnames = data
        _, _, h, w = imgs.size()
        dev = torch.device("cuda", args.start_gpu)
        imgs, gts = imgs.to(dev), gts.to(dev)
        batch_data = {'img_data': imgs, 'seg_label': gts, 'clipimgs_data': [c.to(dev) for c in clip_imgs]}
        if args.use_memory:
            batch_data['is_clean_memory'] = (i == 0)
        with torch.no_grad():
            scores = segmentation_module(batch_data, segSize=(imgs.size(2), imgs.size(3)))
            pred_d = torch.argmax(scores, dim=1).to(torch.int32)
        evaluator.add_batch_device(gts, pred_d)
        eval_video.add_batch_device(gts, pred_d)
        predlist_.append(pred_d)
        gtlist_.append(gts.squeeze(1))
        if args.is_save:
            pred = pred_d.cpu().numpy()
            for j in range(pred.shape[0]):
                _save_pred(args, video, gtnames[j], pred[j])
    return gtlist_, predlist_, h, w


def _save_pred(args, video, name, pred_hw):
    from PIL import Image
    os.makedirs(os.path.join(args.saveroot, video), exist_ok=True)
    im = Image.fromarray(pred_hw.astype('uint8')).convert('P')
    im.putpalette(vspw_palette())
    im.save(os.path.join(args.saveroot, video, name.split('.')[0] + '.png'))


def test_all(segmentation_module, loader, gpu, args, evaluator, eval_video, video):
    """The reference's whole-clip loop (test_clip2.py:90-195) for models that score EVERY frame of a clip (Non_local3d): the
    loader slides a clip_num window over the video, each frame therefore receives up to clip_num probability maps, and a
    frame is decided -- mean of its maps, argmax -- as soon as clip_num maps have arrived; frames near the ends that never
    collect clip_num maps are decided after the last clip from what they have.  Frames enter the metric lists in the order
    they are decided, as in the reference."""
    segmentation_module.eval()
    dev = torch.device("cuda", args.start_gpu)
    gtlist_, predlist_ = [], []
    target_dic, pred_dic, done = {}, {}, set()
    h = w = 0

    def decide(name, maps):
        prob = torch.cat(maps, dim=0).mean(dim=0, keepdim=True)
        pred_d = torch.argmax(prob, dim=1)
        gts = target_dic.pop(name).to(dev).unsqueeze(0)  # (1, 1, H, W)
        evaluator.add_batch_device(gts, pred_d)
        eval_video.add_batch_device(gts, pred_d)
        predlist_.append(pred_d.to(torch.int32))
        gtlist_.append(gts.squeeze(1))
        if args.is_save:
            _save_pred(args, video, name, pred_d[0].cpu().numpy())

    for data in loader:
        imgs, _, clip_imgs, clip_targets, gtnames = data
        _, _, h, w = imgs.size()
        batch_data = {'clipimgs_data': [c.to(dev) for c in clip_imgs], 'cliplabels_data': clip_targets}
        with torch.no_grad():
            scores = segmentation_module(batch_data, segSize=(imgs.size(2), imgs.size(3)))
        for score, clip_target, gtname in zip(scores, clip_targets, gtnames):
            for ii in range(score.size(0)):
                nn_ = gtname[ii]
                if nn_ in done:
                    continue
                target_dic.setdefault(nn_, clip_target[ii])
                pred_dic.setdefault(nn_, []).append(score[ii].unsqueeze(0))
                if len(pred_dic[nn_]) > args.clip_num - 1:
                    decide(nn_, pred_dic.pop(nn_))
                    done.add(nn_)
    for name, maps in pred_dic.items():
        decide(name, maps)
    return gtlist_, predlist_, h, w


def build_module(cfg, args, num_class):
    net_encoder = ModelBuilder.build_encoder(arch=cfg.MODEL.arch_encoder, fc_dim=cfg.MODEL.fc_dim, weights='')
    crit = nn.NLLLoss(ignore_index=-1)
    if args.method == 'clip_psp':
        return Clip_PSP(net_encoder, crit, args)
    if args.method == 'clip_ocr':
        return ClipOCRNet(net_encoder, crit, args)
    if args.method == 'nonlocal3d':
        return Non_local3d(args, net_encoder, crit)
    raise NotImplementedError(f"--method {args.method!r}: only clip_psp / clip_ocr / nonlocal3d run on the B200 engine")


def main(cfg, gpu, args):
    num_class = 42 if args.lesslabel else args.num_class
    torch.cuda.set_device(gpu)
    E.set_precision(args.precision)
    torch.manual_seed(cfg.TRAIN.seed)
    segmentation_module = build_module(cfg, args, num_class)
    segmentation_module.cuda(args.start_gpu)
    if args.load:
        to_load = torch.load(args.load, map_location=torch.device("cuda:" + str(args.start_gpu)))
        segmentation_module.load_state_dict(OrderedDict((k[7:], v) for k, v in to_load.items()))  # strip 'module.' (:267-271)
    elif not args.synthetic:
        raise ValueError("--load is required (a model_epoch_E.pth written by train_clip2.py)")
    if args.gpu_num > 1:
        raise NotImplementedError("multi-GPU inference = one test_clip2.py process per GPU over disjoint --split lists")
    if args.synthetic:
        videolists = [f"synthetic_{i:03d}" for i in range(args.synthetic_videos)]
    else:
        with open(os.path.join(args.dataroot, args.split + '.txt')) as f:
            videolists = [line[:-1] for line in f.readlines()]
    evaluator, eval_video = Evaluator(num_class), Evaluator(num_class)
    total_vmIOU = total_vfwIOU = 0.0
    total_VC_acc = []
    for video in videolists:
        eval_video.reset()
        whole_clip = args.method == 'nonlocal3d'  # test_clip2.py:300-308: TestDataset_clip + test_all
        if args.synthetic:
            h, w = (int(x) for x in args.synthetic_size.lower().split("x"))
            cls = SyntheticWindowTest if whole_clip else SyntheticClipTest
            test_dataset = cls(args, video, frames=args.synthetic_frames, height=h, width=w, seed=cfg.TRAIN.seed)
        else:
            # = dataset2.TestDataset_clip (:154-337) / TestDataset_longclip (:344-490)
            from cvpr2021_vspw_implement_b200.vspw_data import VSPWClipTest, VSPWWindowTest
            test_dataset = (VSPWWindowTest if whole_clip else VSPWClipTest)(args.dataroot, video, args, is_train=False)
        loader_test = torch.utils.data.DataLoader(test_dataset, batch_size=args.batchsize, shuffle=False, num_workers=0, drop_last=False)
        run = test_all if whole_clip else test
        gtlist_, predlist_, h, w = run(segmentation_module, loader_test, gpu, args, evaluator, eval_video, video)
        # VC_n on the device: one launch per video, 2 integers per window come back
        accs = get_common_device(torch.cat(gtlist_, dim=0), torch.cat(predlist_, dim=0), args.vc_clip_num) if gtlist_ else []
        if accs:
            print(sum(accs) / len(accs))
        total_VC_acc.extend(accs)
        eval_video.sync_device()
        v_mIOU = eval_video.Mean_Intersection_over_Union()
        total_vmIOU += v_mIOU
        total_vfwIOU += eval_video.Frequency_Weighted_Intersection_over_Union()
        print(video, v_mIOU)
    total_vmIOU /= len(videolists)
    total_vfwIOU /= len(videolists)
    evaluator.sync_device()
    Acc, Acc_class = evaluator.Pixel_Accuracy(), evaluator.Pixel_Accuracy_Class()
    mIoU, FWIoU = evaluator.Mean_Intersection_over_Union(), evaluator.Frequency_Weighted_Intersection_over_Union()
    print("Acc:{}, Acc_class:{}, mIoU:{}, fwIoU: {}, video mIOU: {}, video fwIOU: {}".format(Acc, Acc_class, mIoU, FWIoU, total_vmIOU, total_vfwIOU))
    VC_Acc = np.nanmean(np.array(total_VC_acc)) if total_VC_acc else float("nan")
    print("Video Consistency num :{} acc:{}".format(args.vc_clip_num, VC_Acc))
    print('Inference done!')
    return {"Acc": Acc, "Acc_class": Acc_class, "mIoU": mIoU, "fwIoU": FWIoU, "video_mIoU": total_vmIOU, "VC": VC_Acc}


def make_parser():
    parser = argparse.ArgumentParser(description="PyTorch Semantic Segmentation Testing (B200 engine)")
    parser.add_argument("--cfg", default="config/vsp-resnet101dilated-ppm_deepsup_clip.yaml", metavar="FILE", type=str)
    for name, typ, default in (("num_class", int, 124), ("start_gpu", int, 0), ("dataroot", str, ''), ("saveroot", str, ''),
                               ("load_en", str, ''), ("load_de", str, ''), ("load", str, ''), ("batchsize", int, 4), ("split", str, 'val'),
                               ("is_save", str2bool, False), ("lesslabel", str2bool, False), ("use_720p", str2bool, False),
                               ("clip_num", int, 5), ("dilation_num", int, 0), ("gpu_num", int, 1), ("propclip2", str2bool, False),
                               ("early_usecat", str2bool, False), ("earlyfuse", str2bool, False), ("allsup", str2bool, False),
                               ("allsup_scale", float, 0.3), ("deepsup_scale", float, 0.0), ("linear_combine", str2bool, False),
                               ("distsoftmax", str2bool, False), ("distnearest", str2bool, False), ("temp", float, 3),
                               ("max_distances", str, '10'), ("clipocr_all", str2bool, False), ("dilation2", str, "2,5,9"),
                               ("use_memory", str2bool, False), ("memory_num", int, 8), ("vc_clip_num", int, 8),
                               ("psp_weight", str2bool, False)):
        parser.add_argument("--" + name, type=typ, default=default)
    parser.add_argument("--method", type=str, default='', choices=['clip_psp', 'clip_ocr', 'nonlocal3d'] + OTHER_METHODS)
    parser.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16", "fp32"])
    parser.add_argument("--synthetic", type=str2bool, default=False)
    parser.add_argument("--synthetic_size", type=str, default="480x854")
    parser.add_argument("--synthetic_videos", type=int, default=2)
    parser.add_argument("--synthetic_frames", type=int, default=12)
    parser.add_argument("opts", help="Modify config options using the command-line", default=None, nargs=argparse.REMAINDER)
    return parser


if __name__ == '__main__':
    args = make_parser().parse_args()
    args.max_distances = [int(dd) for dd in args.max_distances.split(',')]
    cfg.merge_from_file(args.cfg)
    cfg.merge_from_list(args.opts)
    logger = setup_logger(distributed_rank=0)
    logger.info("Loaded configuration file {}".format(args.cfg))
    logger.info("Running with config:\n{}".format(cfg))
    cfg.MODEL.arch_encoder = cfg.MODEL.arch_encoder.lower()
    cfg.MODEL.arch_decoder = cfg.MODEL.arch_decoder.lower()
    cfg.MODEL.weights_encoder = args.load_en
    cfg.MODEL.weights_decoder = args.load_de
    if args.saveroot:
        os.makedirs(args.saveroot, exist_ok=True)
    main(cfg, args.start_gpu, args)
    print(args)
