"""Drop-in replacement for the reference's rgbd_segmentation_RAA.py: put this directory first on PYTHONPATH
(or copy it over the reference checkout); `from rgbd_segmentation_RAA import RGBDSegmentation_RAA`
(train.py:38, test.py:41) then resolves to the B200 implementation."""
from cosnet_b200.rgbd_segmentation_raa import RGBDSegmentation_RAA  # noqa: F401
