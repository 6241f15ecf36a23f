"""Host-side checks of the training (forward + backward) plan: built GPU-less through mdb_unet_create_dry."""
import ctypes

import pytest

from helpers import full_config, tiny_config


def _dry(cfg, batch, training):
    from meshdiffusion_b200 import _native
    from meshdiffusion_b200.diffusion.models import ddpm
    L = _native.lib()
    c = ddpm._config_c(ddpm.arch_from_config(cfg), batch, "bf16", training=training)
    h = ctypes.c_void_p()
    _native.check(L.mdb_unet_create_dry(ctypes.byref(c), ctypes.byref(h)))
    arena, n = ctypes.c_longlong(), L.mdb_unet_num_params(h)
    _native.check(L.mdb_unet_info(h, None, ctypes.byref(arena), None, None))
    L.mdb_unet_destroy(h)
    return arena.value, n


@pytest.mark.parametrize("name,batch", [("tiny", 3), ("res64", 1), ("res128", 1)])
def test_training_plan_builds_and_frees_everything(name, batch):
    """The backward emitters must hand every activation / gradient / scratch block back to the arena (the builder
    throws on a leak), and the plan must not change the parameter table."""
    cfg = tiny_config("res64", "bf16") if name == "tiny" else full_config(name, "bf16")
    a_inf, n_inf = _dry(cfg, batch, False)
    a_trn, n_trn = _dry(cfg, batch, True)
    assert n_inf == n_trn
    assert a_trn > a_inf
    if name == "res64":
        assert a_trn < 3.5 * 2 ** 30, "res64 training arena grew beyond 3.5 GB per sample"


def test_training_engine_refuses_tf32():
    from meshdiffusion_b200 import _native
    from meshdiffusion_b200.diffusion.models import ddpm
    L = _native.lib()
    cfg = tiny_config("res64", "tf32")
    c = ddpm._config_c(ddpm.arch_from_config(cfg), 1, "tf32", training=True)
    h = ctypes.c_void_p()
    assert L.mdb_unet_create_dry(ctypes.byref(c), ctypes.byref(h)) != 0
    assert b"bf16" in L.mdb_last_error()
