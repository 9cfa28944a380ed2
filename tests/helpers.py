"""Shared helpers for the test modules."""


def unet_kwargs(cfg):
    return dict(image_size=32, in_channels=cfg["in_channels"], out_channels=cfg["out_channels"],
                model_channels=cfg["model_channels"], attention_resolutions=list(cfg["attention_resolutions"]),
                num_res_blocks=cfg["num_res_blocks"], channel_mult=list(cfg["channel_mult"]),
                num_heads=cfg["num_heads"], use_spatial_transformer=True, transformer_depth=1,
                context_dim=cfg["context_dim"], use_checkpoint=True, legacy=False,
                latent_size=(cfg["latent_h"], cfg["latent_w"]), max_context_len=40)
