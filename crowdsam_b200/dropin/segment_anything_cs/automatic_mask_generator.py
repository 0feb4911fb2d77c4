from crowdsam_b200.automask import SamAutomaticMaskGenerator  # noqa: F401
