"""Parameter/buffer container for the reference's SynchronizedBatchNorm2d.

Reference: models/sync_batchnorm/batchnorm.py:30-98.  On one device (and in eval mode) the reference
class is exactly ``F.batch_norm``; the arithmetic here is done by ``engine.batchnorm_act`` (CUDA).
The class only owns ``weight``, ``bias``, ``running_mean``, ``running_var``, ``num_batches_tracked``
under the same state_dict keys, so reference checkpoints load unchanged.
"""
import torch.nn as nn


class SynchronizedBatchNorm2d(nn.BatchNorm2d):
    def forward(self, input):  # pragma: no cover - containers are never called directly
        raise RuntimeError("SynchronizedBatchNorm2d is executed by the vspw_b200 tape engine, not called directly")


BatchNorm2d = SynchronizedBatchNorm2d
