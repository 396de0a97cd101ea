"""B200-native co-attention hot path of the COSNet `raa` model (rgbd_segmentation_RAA.py:150-187, :204-238)."""
from .coattention import check_overflow, coattention, coattention_forward_raw, coattention_pair, workspace_bytes  # noqa: F401

__all__ = ["coattention", "coattention_forward_raw", "coattention_pair", "workspace_bytes", "check_overflow"]
