from crowdsam_b200.transforms import ResizeLongestSide  # noqa: F401
