from cosnet_b200.backbone import LEARNABLE_AFFINE as k_learnable_affine_parameters  # noqa: F401
