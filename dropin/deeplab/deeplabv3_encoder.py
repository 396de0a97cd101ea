from cosnet_b200.backbone import ASPP, DepthEncoder_ResNetASPP, Encoder  # noqa: F401
