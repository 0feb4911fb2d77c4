from crowdsam_b200.modules import ImageEncoderViT, MaskDecoder, PromptEncoder, Sam  # noqa: F401
