from crowdsam_b200.pipeline import CrowdSAM  # noqa: F401
