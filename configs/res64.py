"""64^3 DMTet-grid DDPM (same keys and values as the reference's configs/res64.py:6-63)."""
from configs.default_configs import get_default_configs

_OVERRIDES = {
    "training": dict(sde="vpsde", continuous=False, reduce_mean=True, batch_size=48, lip_scale=None,
                     snapshot_freq_for_preemption=1000),
    "sampling": dict(method="pc", predictor="ancestral_sampling", corrector="none"),
    "data": dict(dataset="ShapeNet", centered=True, image_size=64, num_channels=4, meta_path="PLACEHOLDER",
                 filter_meta_path="PLACEHOLDER", num_workers=4, aug=True),
    "model": dict(name="ddpm_res64", scale_by_sigma=False, num_scales=1000, ema_rate=0.9999,
                  normalization="GroupNorm", nonlinearity="swish", nf=128, ch_mult=(1, 1, 2, 4, 4),
                  num_res_blocks_first=2, num_res_blocks=3, attn_resolutions=(16,), resamp_with_conv=True,
                  conditional=True, dropout=0.1),
    "optim": dict(lr=2e-5),
    "eval": dict(batch_size=4, eval_dir="PLACEHOLDER"),
}


def get_config():
    config = get_default_configs()
    for section, values in _OVERRIDES.items():
        for k, v in values.items():
            config[section][k] = v
    config.seed = 42
    return config
