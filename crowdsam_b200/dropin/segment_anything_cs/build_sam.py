from crowdsam_b200.build import (_build_sam, build_sam, build_sam_vit_b, build_sam_vit_h, build_sam_vit_l,  # noqa: F401
                                 build_sam_vit_t, sam_model_registry)
