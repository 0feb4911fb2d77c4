from crowdsam_b200.amg import *  # noqa: F401,F403
from crowdsam_b200.amg import MaskData, rle_to_mask, area_from_rle  # noqa: F401
