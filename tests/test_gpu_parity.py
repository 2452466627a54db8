"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs and against the golden fixtures generated from the reference.

Tolerances (north_star): per-stem relative L2 <= 1e-4 in fp32/TF32 mode, STFT bins <= 1e-5.
"""
import random

import numpy as np
import pytest
import torch

from oracle.htdemucs_oracle import htdemucs_forward, stft_cac, istft_cac
from oracle.apply_oracle import apply_model_oracle
from _fixtures import (golden, rel_l2, strided, small_config, synth_mix, forward_fixture_inputs,
                       init_weights, htdemucs_config, APPLY_CASES, BAG_WEIGHTS)
import demucs_b200 as D
from demucs_b200 import _lib
from demucs_b200.engine import Engine

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
STEM_TOL = 1e-4
BLOCK_TOL = 2e-5   # fp32 kernels vs fp32 oracle, per block


def stem_errors(got, want):
    S = want.shape[1]
    return [rel_l2(got[:, s], want[:, s]) for s in range(S)]


def test_library_is_the_cuda_one():
    assert _lib.TEST_HOOK is None
    assert _lib.lib().bd_version() >= 1


@pytest.mark.parametrize("L", [343980, 50001, 4096])
def test_stft_istft_kernels(L):
    cfg = small_config()
    eng = Engine(cfg, init_weights(cfg, 0), DEV)
    B, S = 2, 3
    x = synth_mix(B, L, 7)
    xd = x.to(DEV)
    T = (L + 1023) // 1024
    spec = torch.empty(B, T, 2048, 4, device=DEV)
    stats = torch.zeros(4 * B, dtype=torch.float64, device=DEV)
    _lib.call("bd_stft_cac", xd.data_ptr(), eng.window.data_ptr(), eng.twiddle.data_ptr(), spec.data_ptr(),
              stats.data_ptr(), B, 2, L, 0)
    torch.cuda.synchronize()
    want = stft_cac(x.double())
    got = spec.permute(0, 3, 2, 1).cpu()
    if rel_l2(got, want) >= 1e-5:      # diagnostics: where does it differ?
        err = (got.double() - want).abs()          # [B, 4, F, T]
        bad = (err > 1e-4 * want.abs().max()).nonzero()
        print("STFT mismatch:", rel_l2(got, want), "bad elements:", bad.shape[0], "first:", bad[:12].tolist(),
              "frames:", sorted(set(bad[:, 3].tolist()))[:20], "max err", float(err.max()), "max ref", float(want.abs().max()))
        spec2 = torch.empty_like(spec)
        st2 = torch.zeros_like(stats)
        _lib.call("bd_stft_cac", xd.data_ptr(), eng.window.data_ptr(), eng.twiddle.data_ptr(), spec2.data_ptr(),
                  st2.data_ptr(), B, 2, L, 0)
        torch.cuda.synchronize()
        print("second run rel_l2:", rel_l2(spec2.permute(0, 3, 2, 1).cpu(), want), "window/twiddle checksums",
              float(eng.window.double().sum()), float(eng.twiddle.double().abs().sum()))
    assert rel_l2(got, want) < 1e-5
    assert float((got.double() - want).abs().max() / want.abs().max()) < 1e-5   # bins, max-abs flavour
    st = stats.view(B, 4).cpu()
    assert torch.allclose(st[:, 0], want.sum(dim=(1, 2, 3)), rtol=1e-4, atol=1e-3)
    assert torch.allclose(st[:, 1], (want ** 2).sum(dim=(1, 2, 3)), rtol=1e-5)
    assert torch.allclose(st[:, 2], x.double().sum(dim=(1, 2)), rtol=1e-6, atol=1e-6)
    assert torch.allclose(st[:, 3], (x.double() ** 2).sum(dim=(1, 2)), rtol=1e-6)
    # inverse: random spectrogram, identity normalisation
    g = torch.Generator().manual_seed(11)
    z = torch.randn(B, S, 4, 2048, T, generator=g)
    zd = z.permute(0, 4, 3, 1, 2).reshape(B, T, 2048, 4 * S).contiguous().to(DEV)
    norm = torch.zeros(B, 8, device=DEV)
    norm[:, 1] = 1.0
    frames = torch.empty(B * S * 2 * T * 4096, device=DEV)
    out = torch.empty(B, S, 2, L, device=DEV)
    _lib.call("bd_istft_frames", zd.data_ptr(), norm.data_ptr(), eng.window.data_ptr(), eng.twiddle.data_ptr(),
              frames.data_ptr(), B, S, T, 0)
    _lib.call("bd_ola_combine", frames.data_ptr(), None, norm.data_ptr(), out.data_ptr(), B, S, T, L, L, 0)
    torch.cuda.synchronize()
    want = istft_cac(z.double(), L)
    assert rel_l2(out.cpu(), want) < 1e-5
    # fused variant: source-major spectrogram, overlap-add in shared memory, cropped output, + time branch
    zs = z.permute(0, 4, 1, 3, 2).reshape(B, T, S, 2048, 4).contiguous().to(DEV)    # [B,T,S,F,(2c+reim)]
    xt = torch.randn(B, L, 2 * S, generator=g).to(DEV)
    norm[:, 4], norm[:, 5] = 0.25, 1.5
    Lout = L - 37
    out2 = torch.full((B, S, 2, Lout), float("nan"), device=DEV)
    _lib.call("bd_istft_ola", zs.data_ptr(), norm.data_ptr(), eng.window.data_ptr(), eng.twiddle.data_ptr(),
              xt.data_ptr(), out2.data_ptr(), B, S, T, L, Lout, 0)
    torch.cuda.synchronize()
    want2 = want[..., :Lout] + (xt.cpu().double()[:, :Lout] * 1.5 + 0.25).view(B, Lout, S, 2).permute(0, 2, 3, 1)
    assert not torch.isnan(out2).any()
    assert rel_l2(out2.cpu(), want2) < 1e-5


def test_spectral_kernels_match_reference_golden():
    g = golden("spectral.npz")
    cfg = small_config()
    eng = Engine(cfg, init_weights(cfg, 0), DEV)
    for name, L in (("full", 343980), ("odd", 50001)):
        x = synth_mix(2, L, 7).to(DEV)
        T = (L + 1023) // 1024
        spec = torch.empty(2, T, 2048, 4, device=DEV)
        stats = torch.zeros(8, dtype=torch.float64, device=DEV)
        _lib.call("bd_stft_cac", x.data_ptr(), eng.window.data_ptr(), eng.twiddle.data_ptr(), spec.data_ptr(),
                  stats.data_ptr(), 2, 2, L, 0)
        got = strided(spec.permute(0, 3, 2, 1).contiguous(), 97)
        want = g[f"{name}.stft"]
        assert rel_l2(got, want) < 1e-5
        assert np.abs(got - want).max() / np.abs(want).max() < 1e-5


@pytest.mark.parametrize("name", ["small_ls05.npz", "small_short.npz"])
def test_forward_blocks_small(name):
    """Every block tap + the output against the oracle (full tensors) and the reference golden."""
    g = golden(name)
    cfg = small_config()
    W, mix = forward_fixture_inputs(g, cfg)
    taps_o, taps = {}, {}
    with torch.no_grad():
        want = htdemucs_forward(W, cfg, mix, taps_o)
    eng = Engine(cfg, W, DEV)
    got = eng.forward(mix.to(DEV), taps)
    torch.cuda.synchronize()
    for k, v in taps_o.items():
        if k in ("istft", "time_out"):          # the engine trims these to the input length
            v = v[..., :taps[k].shape[-1]]
        assert rel_l2(taps[k].cpu(), v) < BLOCK_TOL, k
    assert max(stem_errors(got.cpu(), want)) < STEM_TOL
    assert rel_l2(strided(got, int(g["stride"])), g["out"]) < STEM_TOL
    for key in g.files:
        if key.startswith("tap."):
            assert rel_l2(strided(taps[key[4:]].contiguous(), int(g["tap_stride"])), g[key]) < 5e-5, key


@pytest.mark.parametrize("name", ["htdemucs_default.npz", "htdemucs_ls05.npz"])
def test_forward_htdemucs_full_size(name):
    """The real htdemucs geometry (41.98 M parameters, 7.8 s segment) against the reference golden
    and the CPU oracle."""
    g = golden(name)
    cfg = htdemucs_config()
    W, mix = forward_fixture_inputs(g, cfg)
    eng = Engine(cfg, W, DEV)
    taps = {}
    got = eng.forward(mix.to(DEV), taps)
    torch.cuda.synchronize()
    assert rel_l2(strided(got, int(g["stride"])), g["out"]) < STEM_TOL
    for key in g.files:
        if key.startswith("tap."):
            assert rel_l2(strided(taps[key[4:]].contiguous(), int(g["tap_stride"])), g[key]) < 1e-4, key
    with torch.no_grad():
        want = htdemucs_forward(W, cfg, mix)
    assert max(stem_errors(got.cpu(), want)) < STEM_TOL


def test_forward_batch_items_are_independent():
    cfg = small_config()
    W = init_weights(cfg, 2, layer_scale=0.5)
    eng = Engine(cfg, W, DEV)
    mix = synth_mix(5, cfg.segment_length, 3).to(DEV)
    full = eng.forward(mix).clone()
    for b in (0, 4):
        one = eng.forward(mix[b:b + 1].contiguous())
        assert rel_l2(one.cpu(), full[b:b + 1].cpu()) < 1e-6


def test_apply_model_matches_reference_golden():
    g = golden("apply_small.npz")
    cfg = small_config()
    models = [D.HTDemucs.from_config(cfg, init_seed=s, layer_scale=0.5).to(DEV) for s in range(2)]
    mix = synth_mix(1, int(g["length"]), 99)
    stride = int(g["stride"])
    for name, kw in APPLY_CASES.items():
        m = mix[..., :50000] if name == "nosplit" else mix
        random.seed(0)
        out = D.apply_model(models[0], m.clone(), device=DEV, **kw)
        assert out.device.type == "cpu" and list(out.shape) == [1, 3, 2, m.shape[-1]]
        assert rel_l2(strided(out, stride), g[name]) < STEM_TOL, name
    random.seed(0)
    out = D.apply_model(D.BagOfModels(models, BAG_WEIGHTS), mix.clone(), shifts=1, device=DEV)
    assert rel_l2(strided(out, stride), g["bag"]) < STEM_TOL


def test_apply_model_full_size_round_trip_properties():
    """BASELINE config 1 geometry (10 s clip, 2 segments).  Size-independent checks:
    (i) equality with the oracle's apply on the same weights, (ii) a clip of one segment length
    with split=False goes through overlap-add unchanged relative to a plain forward."""
    cfg = htdemucs_config()
    model = D.HTDemucs.from_config(cfg, init_seed=0).to(DEV)
    mix = synth_mix(1, 441000, 1234)
    out = D.apply_model(model, mix, shifts=0, split=True, overlap=0.25, device=DEV)
    W = init_weights(cfg, 0)
    with torch.no_grad():
        want = apply_model_oracle((W, cfg), mix, shifts=0, split=True, overlap=0.25)
    assert max(stem_errors(out, want)) < STEM_TOL
    short = mix[..., :cfg.segment_length].contiguous()
    a = D.apply_model(model, short, shifts=0, split=False, device=DEV)
    b = model(short.to(DEV)).cpu()
    # two forwards of the same input: the GroupNorm partial sums meet through order-dependent atomics (last-bit noise),
    # which this random-weight network amplifies ~40x (tools/dev_determinism.py: 1.4e-5 run to run in this mode)
    assert rel_l2(a, b) < 5e-5


def test_separator_front_door():
    cfg = small_config()
    model = D.HTDemucs.from_config(cfg, init_seed=0, layer_scale=0.5)
    sep = D.Separator(model, device=DEV, shifts=0)
    wav = synth_mix(1, 70000, 4)[0]
    keep = wav.clone()
    got_wav, stems = sep.separate_tensor(wav)
    assert set(stems) == set(cfg.sources) and stems["a"].shape == (2, 70000)
    assert torch.allclose(got_wav, keep, atol=1e-6)
    # oracle: api.py:267-290 normalisation around apply_model
    ref = keep.mean(0)
    x = (keep - ref.mean()) / (ref.std() + 1e-8)
    with torch.no_grad():
        want = apply_model_oracle((init_weights(cfg, 0, layer_scale=0.5), cfg), x[None], shifts=0)
    want = want * (ref.std() + 1e-8) + ref.mean()
    for i, s in enumerate(cfg.sources):
        assert rel_l2(stems[s], want[0, i]) < STEM_TOL


def test_errors_are_loud():
    cfg = small_config()
    model = D.HTDemucs.from_config(cfg, init_seed=0).to(DEV)
    with pytest.raises(ValueError):
        model(torch.zeros(1, 2, cfg.segment_length + 1, device=DEV))
    with pytest.raises(D.KernelError):
        D.HTDemucs.from_config(cfg, init_seed=0).forward(torch.zeros(1, 2, 4096))   # CPU tensors: no fallback


BF16_TOL = 1e-2   # the north_star's bound for the bf16 mode (per-stem relative L2)
TF32_TOL = 8e-3   # "tf32" (single-pass kind::tf32) is an auxiliary mode with NO north_star tolerance class: it misses the
#                   1e-4 bound by construction (10-bit operands); this is its own measured envelope (worst block tap
#                   6.2e-3, stems 1.3e-3 .. 3.9e-3), kept as a regression guard


@pytest.mark.parametrize("mode,tol", [("bf16", BF16_TOL), ("tf32", TF32_TOL)])
@pytest.mark.parametrize("name", ["htdemucs_default.npz", "htdemucs_ls05.npz"])
def test_forward_htdemucs_reduced_precision_modes(name, mode, tol):
    """"bf16": bf16 tensors and kind::f16 MMAs through the transformer, single-pass tf32 convolutions -- the
    north_star's bf16 tolerance, per-stem relative L2 <= 1e-2, on both weight fixtures, every block tap included.
    "tf32": single-pass kind::tf32 everywhere (auxiliary; its own envelope)."""
    g = golden(name)
    cfg = htdemucs_config()
    W, mix = forward_fixture_inputs(g, cfg)
    eng = Engine(cfg, W, DEV, mode=mode)
    taps = {}
    got = eng.forward(mix.to(DEV), taps)
    torch.cuda.synchronize()
    errs = {}
    for key in g.files:
        if key.startswith("tap."):
            errs[key[4:]] = rel_l2(strided(taps[key[4:]].contiguous(), int(g["tap_stride"])), g[key])
    e_out = rel_l2(strided(got, int(g["stride"])), g["out"])
    print(name, mode, "out rel-L2", e_out, "worst tap", max(errs.items(), key=lambda kv: kv[1]))
    assert max(errs.values()) < tol
    with torch.no_grad():
        want = htdemucs_forward(W, cfg, mix)
    stems = stem_errors(got.cpu(), want)
    print("per-stem", stems)
    assert max(stems) < tol


@pytest.mark.parametrize("mode,tol", [("bf16", BF16_TOL), ("tf32", TF32_TOL), ("strict", 1e-4)])
def test_forward_blocks_small_tc_modes(mode, tol):
    """Small geometry through the tensor-core arm (16/32-float k-blocks, R0 < 128 tiles, 3-tap transposed
    convs, narrow tiles): every block within TF32 rounding of the fp32 oracle."""
    g = golden("small_ls05.npz")
    cfg = small_config()
    W, mix = forward_fixture_inputs(g, cfg)
    taps_o, taps = {}, {}
    with torch.no_grad():
        want = htdemucs_forward(W, cfg, mix, taps_o)
    eng = Engine(cfg, W, DEV, mode=mode)
    got = eng.forward(mix.to(DEV), taps)
    torch.cuda.synchronize()
    errs = {k: rel_l2(taps[k].cpu(), v) for k, v in taps_o.items()}
    print("small", mode, "taps", {k: f"{e:.1e}" for k, e in errs.items()})
    assert max(errs.values()) < tol
    assert max(stem_errors(got.cpu(), want)) < tol


@pytest.mark.parametrize("mode", ["strict", "tf32x3"])
@pytest.mark.parametrize("name", ["htdemucs_default.npz", "htdemucs_ls05.npz"])
def test_forward_htdemucs_strict_modes(name, mode):
    """Error-compensated tensor-core modes -- "strict" (the default: bf16 hi/lo operand split, three kind::f16
    products) and "tf32x3" (the same with tf32 parts) -- against the reference golden vectors: the north_star's
    fp32/TF32 bound, per-stem relative L2 <= 1e-4, on both weight fixtures, every block tap included."""
    g = golden(name)
    cfg = htdemucs_config()
    W, mix = forward_fixture_inputs(g, cfg)
    eng = Engine(cfg, W, DEV, mode=mode)
    taps = {}
    got = eng.forward(mix.to(DEV), taps)
    torch.cuda.synchronize()
    errs = {}
    for key in g.files:
        if key.startswith("tap."):
            errs[key[4:]] = rel_l2(strided(taps[key[4:]].contiguous(), int(g["tap_stride"])), g[key])
    e_out = rel_l2(strided(got, int(g["stride"])), g["out"])
    print(name, mode, "out rel-L2", e_out, "worst tap", max(errs.items(), key=lambda kv: kv[1]))
    assert max(errs.values()) < 1e-4
    with torch.no_grad():
        want = htdemucs_forward(W, cfg, mix)
    stems = stem_errors(got.cpu(), want)
    print("per-stem", stems)
    assert max(stems) < STEM_TOL


@pytest.mark.parametrize("mode,tol", [("strict", STEM_TOL), ("bf16", BF16_TOL)])
def test_htdemucs_6s_shifts2(mode, tol):
    """BASELINE config 3 in miniature: the 6-stem geometry (htdemucs_6s), overlap 0.25, shifts=2, two segments,
    in the default (fp32-accurate) and the bf16 mode, against the oracle drawing the same shift offsets."""
    from demucs_b200.config import htdemucs_6s_config
    cfg = htdemucs_6s_config()
    model = D.HTDemucs.from_config(cfg, init_seed=3, mode=mode).to(DEV)
    mix = synth_mix(1, 400000, 77)
    random.seed(11)
    out = D.apply_model(model, mix, shifts=2, split=True, overlap=0.25, device=DEV)
    assert list(out.shape) == [1, 6, 2, 400000]
    W = init_weights(cfg, 3)
    random.seed(11)
    with torch.no_grad():
        want = apply_model_oracle((W, cfg), mix, shifts=2, split=True, overlap=0.25)
    errs = stem_errors(out, want)
    print("6s shifts=2", mode, "per-stem", errs)
    assert max(errs) < tol


@pytest.mark.parametrize("mode,tol", [("bf16", BF16_TOL), ("strict", STEM_TOL)])
def test_bag_of_four_single_source_members(mode, tol):
    """BASELINE config 4 in miniature (htdemucs_ft: four fine-tuned members, member i contributes source i only,
    demucs/remote/htdemucs_ft.yaml) in the bf16 mode the config names (and the default one), against the oracle's
    weighted sum."""
    cfg = htdemucs_config()
    weights = [[1.0 if s == m else 0.0 for s in range(4)] for m in range(4)]
    models = [D.HTDemucs.from_config(cfg, init_seed=10 + m, mode=mode).to(DEV) for m in range(4)]
    mix = synth_mix(1, 300000, 5)
    out = D.apply_model(D.BagOfModels(models, weights), mix, shifts=0, split=True, overlap=0.25, device=DEV)
    want = torch.zeros_like(out)
    with torch.no_grad():
        for m in range(4):
            part = apply_model_oracle((init_weights(cfg, 10 + m), cfg), mix, shifts=0, split=True, overlap=0.25)
            want[:, m] = part[:, m]
    errs = stem_errors(out, want)
    print("bag of 4", mode, "per-stem", errs)
    assert max(errs) < tol


def test_full_size_batch_64_items_are_independent():
    """bench.py runs one 64-segment forward per step (11 M rows in the first layers): every item must come out as it
    does alone (guards the 32-bit row arithmetic and the per-item statistics slabs at that size)."""
    cfg = htdemucs_config()
    eng = Engine(cfg, init_weights(cfg, 0, layer_scale=0.5), DEV, mode="strict")
    mix = synth_mix(64, cfg.segment_length, 21).to(DEV)
    full = eng.forward(mix).clone()
    assert torch.isfinite(full).all()
    for b in (0, 37, 63):
        one = eng.forward(mix[b:b + 1].contiguous())
        # an indexing / slab mix-up would be O(1); what remains is the mode's run-to-run noise (order-dependent
        # atomics amplified by the random-weight network: ~1.4e-5 in "strict", profiles/r02_run_to_run.log)
        assert rel_l2(one.cpu(), full[b:b + 1].cpu()) < 1e-4, b


@pytest.mark.parametrize("mode,tol", [("strict", STEM_TOL), ("bf16", BF16_TOL)])
def test_forward_htdemucs_short_input_batch3(mode, tol):
    """The htdemucs geometry with an input shorter than the training segment (the model pads it, htdemucs.py:536-542)
    and an odd batch, on the single-pass tensor-core arm: exercises the ragged tails of the mma.sync kernels."""
    cfg = htdemucs_config()
    W = init_weights(cfg, 4, layer_scale=0.5)
    mix = synth_mix(3, 123457, 8)
    eng = Engine(cfg, W, DEV, mode=mode)
    got = eng.forward(mix.to(DEV))
    torch.cuda.synchronize()
    assert list(got.shape) == [3, 4, 2, 123457]
    with torch.no_grad():
        want = htdemucs_forward(W, cfg, mix)
    errs = stem_errors(got.cpu(), want)
    print("short input", mode, "per-stem", errs)
    assert max(errs) < tol
