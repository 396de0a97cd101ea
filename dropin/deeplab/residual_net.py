"""`from deeplab.residual_net import Bottleneck` (train.py:39, test.py) -> B200 package."""
from cosnet_b200.backbone import BasicBlock, Bottleneck, ResNet  # noqa: F401
