"""128^3 DMTet-grid DDPM (same keys and values as the reference's configs/res128.py:6-62)."""
from configs.default_configs import get_default_configs

_ITER_SIZE = 4
_OVERRIDES = {
    "training": dict(sde="vpsde", continuous=False, reduce_mean=True, batch_size=8, iter_size=_ITER_SIZE,
                     lip_scale=None, snapshot_freq_for_preemption=1000),
    "sampling": dict(method="pc", predictor="ancestral_sampling", corrector="none"),
    "data": dict(dataset="ShapeNet", centered=True, image_size=128, num_channels=4, meta_path="PLACEHOLDER",
                 filter_meta_path="PLACEHOLDER", num_workers=8, aug=True),
    # the reference names 'ddpm_res128_v2' here while registering only 'ddpm_res128'; both are registered in this repo
    "model": dict(name="ddpm_res128_v2", scale_by_sigma=False, num_scales=1000, ema_rate=0.9999,
                  normalization="GroupNorm", nonlinearity="swish", nf=128, ch_mult=(1, 1, 2, 4, 4, 4),
                  num_res_blocks_first=2, num_res_blocks=2, attn_resolutions=(16,), resamp_with_conv=True,
                  conditional=True, dropout=0.1),
    "optim": dict(lr=7e-5 / _ITER_SIZE * 2.0),
    "eval": dict(batch_size=7),
}


def get_config():
    config = get_default_configs()
    for section, values in _OVERRIDES.items():
        for k, v in values.items():
            config[section][k] = v
    config.seed = 42
    return config
