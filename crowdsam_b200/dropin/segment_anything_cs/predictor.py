from crowdsam_b200.predictor import SamPredictor  # noqa: F401
