"""Hybrid Demucs v3 (BASELINE configs[1]): oracle against the reference golden vectors, the engine's
host logic through the ABI emulator (CPU), the CUDA path against the oracle and the golden vectors (GPU)."""
import numpy as np
import pytest
import torch

from _fixtures import golden, rel_l2, strided, synth_mix
from abi_emulator import emulated_abi
from demucs_b200 import hdemucs as HD
from demucs_b200.hdemucs_engine import HDemucsEngine
from oracle.hdemucs_oracle import hdemucs_forward
from oracle.make_golden import hdemucs_small_config


def _inputs(g, cfg):
    ls = float(g["layer_scale"])
    W = HD.init_weights(cfg, int(g["seed"]), None if ls < 0 else ls)
    return W, synth_mix(int(g["batch"]), int(g["length"]), 4321 + int(g["seed"]))


def _check(g, out, taps, tol):
    errs = {k[4:]: rel_l2(strided(taps[k[4:]].contiguous(), int(g["tap_stride"])), g[k]) for k in g.files if k.startswith("tap.")}
    e_out = rel_l2(strided(out, int(g["stride"])), g["out"])
    worst = max(errs.items(), key=lambda kv: kv[1])
    print("out", e_out, "worst tap", worst)
    assert e_out < tol and worst[1] < tol, (e_out, worst)


def test_parameter_inventory_matches_reference():
    g = golden("hdemucs_mmi.npz")
    specs = HD.param_specs(HD.hdemucs_mmi_config())
    assert list(specs) == list(g["names"]) and [str(tuple(v[0])) for v in specs.values()] == list(g["shapes"])
    assert HD.count_params(HD.hdemucs_mmi_config()) == 83637832
    with pytest.raises(HD.UnsupportedConfig):
        HD.HDemucsConfig.from_reference_kwargs(sources=["a"], cac=False)


@pytest.mark.parametrize("name", ["hdemucs_small.npz", "hdemucs_small_odd.npz"])
def test_oracle_matches_reference_golden(name):
    g = golden(name)
    cfg = hdemucs_small_config()
    W, mix = _inputs(g, cfg)
    taps = {}
    with torch.no_grad():
        out = hdemucs_forward(W, cfg, mix, taps)
    _check(g, out, {k: v for k, v in taps.items() if f"tap.{k}" in g.files}, 2e-6)


def test_engine_host_logic_matches_golden():
    """Every descriptor, weight packing and buffer of HDemucsEngine through the numpy ABI (fp32 mode)."""
    g = golden("hdemucs_small.npz")
    cfg = hdemucs_small_config()
    W, mix = _inputs(g, cfg)
    with emulated_abi():
        eng = HDemucsEngine(cfg, W, "cpu", mode="fp32")
        taps = {}
        out = eng.forward(mix[:1].contiguous(), taps)
    one = {k: v for k, v in g.items()}
    # the golden holds batch 2; compare item 0 through the oracle (itself pinned to the golden above)
    with torch.no_grad():
        otaps = {}
        want = hdemucs_forward(W, cfg, mix[:1], otaps)
    for k, v in taps.items():
        assert rel_l2(v, otaps[k]) < 2e-5, k
    assert rel_l2(out, want) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("strict", 1e-4)])
@pytest.mark.parametrize("name", ["hdemucs_small.npz", "hdemucs_small_odd.npz"])
def test_gpu_forward_small(name, mode, tol):
    g = golden(name)
    cfg = hdemucs_small_config()
    W, mix = _inputs(g, cfg)
    eng = HDemucsEngine(cfg, W, "cuda:0", mode=mode)
    taps = {}
    out = eng.forward(mix.to("cuda:0"), taps)
    torch.cuda.synchronize()
    _check(g, out, taps, tol)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tol", [("strict", 1e-4)])
def test_gpu_forward_hdemucs_mmi(mode, tol):
    """The hdemucs_mmi geometry (83.6 M parameters, one 7.8 s item) against the reference golden vectors, and a batch of
    16 x 7.8 s (BASELINE configs[1]) whose items must equal the single-item result."""
    g = golden("hdemucs_mmi.npz")
    cfg = HD.hdemucs_mmi_config()
    W, mix = _inputs(g, cfg)
    eng = HDemucsEngine(cfg, W, "cuda:0", mode=mode)
    taps = {}
    out = eng.forward(mix.to("cuda:0"), taps)
    torch.cuda.synchronize()
    _check(g, out, taps, tol)
    batch = synth_mix(16, cfg.frames(343980) * 1024 - 1024 + 980, 11)[..., :343980].contiguous().to("cuda:0")
    full = eng.forward(batch).clone()
    assert torch.isfinite(full).all()
    for b in (0, 9, 15):
        one = eng.forward(batch[b:b + 1].contiguous())
        assert rel_l2(one.cpu(), full[b:b + 1].cpu()) < 1e-4, b


def _both(name, make_args, outs):
    """Run one C-ABI entry point on CUDA tensors and through the numpy restatement on copies; return both outputs."""
    import abi_emulator as E
    from demucs_b200 import _lib
    cpu_args = make_args("cpu")
    gpu_args = make_args("cuda:0")
    E.TABLE[name](*[a.data_ptr() if torch.is_tensor(a) else a for a in cpu_args])
    _lib.call(name, *[a.data_ptr() if torch.is_tensor(a) else a for a in gpu_args])
    torch.cuda.synchronize()
    return [(gpu_args[i].cpu(), cpu_args[i]) for i in outs]


@pytest.mark.gpu
def test_gpu_hdemucs_kernels_match_their_specification():
    """gn_stats / gn_act / lstm_frame / lstm_unframe_add / lstm_bidir / local_state against the numpy statement of the
    header's semantics (which the oracle comparison above validates end to end)."""
    g = torch.Generator().manual_seed(3)
    B, rows, C, G = 3, 37, 64, 4
    x = torch.randn(B, rows, C, generator=g) * 2 + 0.3
    (got, want), = _both("bd_gn_stats", lambda d: (x.to(d), torch.zeros(B * G * 2, dtype=torch.float64, device=d), B, rows, C, G, 0), [1])
    assert torch.allclose(got, want, rtol=1e-6)
    mr = torch.stack([torch.randn(B * G, generator=g) * 0.1, 1 + 0.2 * torch.rand(B * G, generator=g)], 1).contiguous()
    gam, bet = torch.randn(C, generator=g), torch.randn(C, generator=g)
    for act, Co in ((0, C), (1, C), (2, C // 2)):
        add = torch.randn(B, 44, Co, generator=g)        # item pitch 44 rows, 30 kept from row 5
        (got, want), = _both("bd_gn_act", lambda d: (x.to(d), torch.zeros(B, 44, Co, device=d), mr.to(d), gam.to(d), bet.to(d),
                                                     add.to(d), B, rows, 5, 30, C, G, act, 44 * Co, 0), [1])
        assert torch.allclose(got, want, atol=2e-5), act
    T, Cc, nf = 336, 24, 4
    h = torch.randn(B, T, Cc, generator=g)
    (fr, fr_w), = _both("bd_lstm_frame", lambda d: (h.to(d), torch.zeros(B * nf, 200, Cc, device=d), B, T, Cc, nf, 200, 100, 0), [1])
    assert torch.equal(fr, fr_w)
    (got, want), = _both("bd_lstm_unframe_add", lambda d: (fr_w.to(d), h.to(d), torch.zeros(B, T, Cc, device=d), B, T, Cc, nf, 200,
                                                           100, 0), [2])
    assert torch.equal(got, want) and torch.allclose(want, 2 * h)          # frames of h + skip h = 2h
    N, Tt, H = 5, 23, 48
    pre = torch.randn(N, Tt, 2, 4 * H, generator=g)
    whh = torch.randn(2, H, 4 * H, generator=g) / H ** 0.5
    (got, want), = _both("bd_lstm_bidir", lambda d: (pre.to(d), whh.to(d), torch.zeros(N, Tt, 2 * H, device=d),
                                                     torch.zeros(6 * N * H, device=d), N, Tt, H, 0), [2])
    assert rel_l2(got, want) < 2e-6
    N, T, D = 2, 77, 64
    qkc = torch.randn(N, T, 3 * D, generator=g)
    dq = torch.randn(N, T, 16, generator=g) - 2
    (got, want), = _both("bd_local_state", lambda d: (qkc.to(d), dq.to(d), torch.zeros(N, T, D, device=d), N, T, D, 4, 0), [2])
    assert rel_l2(got, want) < 2e-6
    # the shapes of hdemucs_mmi: persistent LSTM with 64 / 32 units per block, several batch chunks; long-T LocalState
    for N, Tt, H in ((64, 40, 192), (16, 30, 384), (200, 12, 192)):
        pre = torch.randn(N, Tt, 2, 4 * H, generator=g)
        whh = torch.randn(2, H, 4 * H, generator=g) / H ** 0.5
        (got, want), = _both("bd_lstm_bidir", lambda d: (pre.to(d), whh.to(d), torch.zeros(N, Tt, 2 * H, device=d),
                                                         torch.zeros(6 * N * H, device=d), N, Tt, H, 0), [2])
        assert rel_l2(got, want) < 3e-6, (N, Tt, H)
    N, T, D = 1, 1500, 64       # beyond the tiled kernel's shared memory: the streaming form
    qkc = torch.randn(N, T, 3 * D, generator=g)
    dq = torch.randn(N, T, 16, generator=g) - 2
    (got, want), = _both("bd_local_state", lambda d: (qkc.to(d), dq.to(d), torch.zeros(N, T, D, device=d), N, T, D, 4, 0), [2])
    assert rel_l2(got, want) < 2e-6


def test_apply_model_on_hdemucs_host_logic():
    """apply_model with a v3 model (no ``valid_length``: chunks run at their own length, apply.py:302-312), split with a
    short last chunk and the shift trick, through the emulated ABI against the oracle's apply."""
    import random
    import demucs_b200 as D
    from oracle.apply_oracle import apply_model_oracle
    cfg = hdemucs_small_config()
    cfg.segment = 2.0
    model = HD.HDemucs.from_config(cfg, init_seed=2, layer_scale=0.5, mode="fp32")
    W = HD.init_weights(cfg, 2, 0.5)
    mix = synth_mix(1, 210001, 6)
    for kw in (dict(shifts=0, overlap=0.25), dict(shifts=1, overlap=0.25), dict(shifts=0, split=False)):
        m = mix[..., :60000] if kw.get("split") is False else mix
        with emulated_abi():
            random.seed(4)
            got = D.apply_model(model, m.clone(), **kw)
        random.seed(4)
        with torch.no_grad():
            want = apply_model_oracle((W, cfg), m, **kw)
        assert got.shape == want.shape and rel_l2(got, want) < 2e-5, kw
