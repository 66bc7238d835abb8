"""Non_local3d / NLBlockND on the vspw_b200 tape engine (the reference's genuine dense
cross-frame affinity, `--method nonlocal3d`).

Reference: models/non_local.py:7-152 (NLBlockND) and models/non_local_models.py:9-115 (Non_local3d).  Same constructor
signatures, parameter names/shapes (the 1x1x1 `nn.Conv3d` containers are kept so checkpoints load unchanged), init
(BN of W_z starts at weight = bias = 0: the block is the identity at initialisation), LR-group generators and
forward contract: `feed_dict['clipimgs_data']` / `['cliplabels_data']` are lists of T frames / labels, EVERY frame is
supervised; train -> (mean of the per-frame losses, mean of the per-frame accuracies); eval -> list of T probability maps.

Only what `Non_local3d` uses is implemented: mode='dot', dimension=3, bn_layer=True.  In that mode
y = (theta phi^T / P) g has no softmax, so the engine evaluates it as theta (phi^T g / P) — a 128x128 matrix per clip
instead of the (T h w) x (T h w) affinity (engine.nl_dot_affinity).
"""
import torch
import torch.nn as nn

from .. import engine as E
from .models import _ignore_index
from .resnet import conv_op


class SynchronizedBatchNorm3d(nn.BatchNorm3d):
    def forward(self, input):  # pragma: no cover - containers are never called directly
        raise RuntimeError("SynchronizedBatchNorm3d is executed by the vspw_b200 tape engine, not called directly")


def _conv1(tape, conv, x, bn=None):
    """A kernel-size-1 nn.Conv2d / nn.Conv3d container as a 1x1 conv over NHWC positions."""
    if any(k != 1 for k in conv.kernel_size) or any(s != 1 for s in conv.stride) or any(p != 0 for p in conv.padding):
        raise NotImplementedError("only kernel_size=1, stride=1, padding=0 convolutions appear in NLBlockND")
    return E.conv2d(tape, x, conv.weight, conv.bias, 1, 0, 1, want_stats=bn is not None and bn.training)


class NLBlockND(nn.Module):
    def __init__(self, in_channels, inter_channels=None, mode="embedded", dimension=3, bn_layer=True):
        super().__init__()
        assert dimension in [1, 2, 3]
        if mode not in ["gaussian", "embedded", "dot", "concatenate"]:
            raise ValueError("`mode` must be one of `gaussian`, `embedded`, `dot` or `concatenate`")
        if mode != "dot" or dimension != 3 or not bn_layer:
            raise NotImplementedError("the VSPW code base instantiates NLBlockND(mode='dot', dimension=3, bn_layer=True) only")
        self.mode, self.dimension = mode, dimension
        self.in_channels = in_channels
        self.inter_channels = inter_channels if inter_channels is not None else max(in_channels // 2, 1)
        self.g = nn.Conv3d(self.in_channels, self.inter_channels, kernel_size=1)
        self.W_z = nn.Sequential(nn.Conv3d(self.inter_channels, self.in_channels, kernel_size=1),
                                 SynchronizedBatchNorm3d(self.in_channels))
        nn.init.constant_(self.W_z[1].weight, 0)
        nn.init.constant_(self.W_z[1].bias, 0)
        self.theta = nn.Conv3d(self.in_channels, self.inter_channels, kernel_size=1)
        self.phi = nn.Conv3d(self.in_channels, self.inter_channels, kernel_size=1)

    def graph(self, tape, x, t_frames, n_clips, training):
        g_x = _conv1(tape, self.g, x)
        theta_x = _conv1(tape, self.theta, x)
        phi_x = _conv1(tape, self.phi, x)
        y = E.nl_dot_affinity(tape, theta_x, phi_x, g_x, t_frames, n_clips)
        w_y = E.batchnorm_act(tape, _conv1(tape, self.W_z[0], y, self.W_z[1]), self.W_z[1], relu=False, training=training)
        return E.add_vars(tape, w_y, x)


class Non_local3d(nn.Module):
    def __init__(self, args, net_enc, crit, downsample=False):
        super().__init__()
        if downsample:
            raise NotImplementedError("Non_local3d(downsample=True) is never used by train_clip2.py / test_clip2.py")
        self.encoder = net_enc
        self.downsample = downsample
        self.crit = crit
        self.emb = nn.Conv2d(2048, 256, 1, 1)
        self.nonlocalblock = NLBlockND(in_channels=256, mode="dot", dimension=3, bn_layer=True)
        self.last_layer = nn.Conv2d(512, args.num_class, kernel_size=1, stride=1)

    @staticmethod
    def _walk(modules, want_bias):
        for mod in modules:
            for _, sub in mod.named_modules():
                for key, p in sub.named_parameters():
                    if p.requires_grad and (("bias" in key) == want_bias):
                        yield p

    def get_1x_lr_params(self):
        return self._walk([self.encoder], False)

    def get_10x_lr_params(self):
        return self._walk([self.emb, self.nonlocalblock, self.last_layer], False)

    def get_1x_lr_params_bias(self):
        return self._walk([self.encoder], True)

    def get_10x_lr_params_bias(self):
        return self._walk([self.emb, self.nonlocalblock, self.last_layer], True)

    def pixel_acc(self, pred, label):
        _, preds = torch.max(pred, dim=1)
        valid = (label >= 0).long()
        return torch.sum(valid * (preds == label).long()).float() / (torch.sum(valid).float() + 1e-10)

    def _logits(self, tape, frames, training):
        t_frames, n = len(frames), frames[0].shape[0]
        x = E.Var(E.input_from_frames(frames))
        feat = self.encoder.graph(tape, x)[-1]
        emb = conv_op(tape, self.emb, feat)
        z = self.nonlocalblock.graph(tape, emb, t_frames, n, training)
        cat = E.concat_channels(tape, [emb, z])
        return conv_op(tape, self.last_layer, cat)

    def forward(self, feed_dict, segSize=None):
        frames = list(feed_dict["clipimgs_data"])
        t_frames, n = len(frames), frames[0].shape[0]
        training = self.training
        if segSize is not None:
            def runner(tape):
                logits = self._logits(tape, frames, training)
                outs = []
                for t in range(t_frames):
                    lt = E.slice_images(tape, logits, t * n, (t + 1) * n)
                    outs.append(E.up_softmax(lt, int(segSize[0]), int(segSize[1])))
                return tuple(outs), None

            return list(E.run_graph(self, runner))

        labels = [l.contiguous().float() for l in feed_dict["cliplabels_data"]]
        if len(labels) != t_frames:
            raise ValueError("Non_local3d supervises every frame: len(cliplabels_data) must equal len(clipimgs_data)")
        ignore = _ignore_index(self.crit)

        def runner(tape):
            logits = self._logits(tape, frames, training)
            E.publish("logits", logits)
            terms = [E.nll_term(tape, E.slice_images(tape, logits, t * n, (t + 1) * n), labels[t], ignore, want_acc=True)
                     for t in range(t_frames)]
            loss, acc, gslot = E.loss_mean_of_terms(tape, terms)
            return (loss, acc), lambda g: gslot.__setitem__("g", g.contiguous())

        loss, acc = E.run_graph(self, runner)
        return loss, acc
