"""Drop-in `segment_anything_cs` package (B200 implementation behind the reference's import names)."""
from .build_sam import build_sam, build_sam_vit_b, build_sam_vit_h, build_sam_vit_l, sam_model_registry
from .predictor import SamPredictor
from .automatic_mask_generator import SamAutomaticMaskGenerator
