"""B200-native large-margin cosine-softmax head (drop-in for main_code/utils/criterion.py heads).

Public API:
  ArcFace, CosFace, SphereFace, MV_Softmax, CurricularFace, AdaFace, ElasticCosFace, ElasticArcFace,
  MagFace, VPLArcFace, QAFace - nn.Modules with the reference constructor signatures
  FusedOutput         - result of ``head.fused_loss(feats, labels)``
  ShardedMarginHead   - class-sharded (Partial-FC style) head over torch.distributed / NCCL
  verification        - pair_cosine (CUDA) + the reference's LFW 10-fold protocol on embedding pairs
  HeadSGD             - the reference's SGD(momentum, weight_decay) for the head parameter, fused with the next W prologue
All compute runs in libmargin_head.so (hand-written sm_100a CUDA behind a C ABI, include/margin_head.h).
"""
from .heads import (AdaFace, ArcFace, CosFace, CurricularFace, ElasticArcFace, ElasticCosFace, FusedOutput,
                    HEAD_CLASSES, MagFace, MV_Softmax, QAFace, SphereFace, VPLArcFace)
from .functional import HeadEngine, ShardInfo
from .sharded import ShardedMarginHead, ShardComm, shard_range
from . import _lib
from ._lib import MarginHeadError
from . import verification
from .optim import HeadSGD

__all__ = ["AdaFace", "ArcFace", "CosFace", "CurricularFace", "ElasticArcFace", "ElasticCosFace", "FusedOutput",
           "HEAD_CLASSES", "MagFace", "MarginHeadError", "MV_Softmax", "QAFace", "SphereFace", "VPLArcFace", "HeadEngine", "ShardInfo", "ShardedMarginHead", "ShardComm", "shard_range", "_lib", "verification",
           "HeadSGD"]
