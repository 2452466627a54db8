"""CPU: model packages and repositories (demucs_b200/repo.py) -- reference states.py / repo.py / pretrained.py semantics."""
import sys
import types

import pytest
import torch
import yaml

import demucs_b200 as D
from demucs_b200 import repo as R
from _fixtures import small_config


def _model(seed=0):
    return D.HTDemucs.from_config(small_config(), init_seed=seed, layer_scale=0.5)


def test_package_round_trip_fp16_and_checksum(tmp_path):
    m = _model()
    path = R.save_with_checksum(R.serialize_model(m, half=True), tmp_path / "abcd1234.th")
    assert path.name.startswith("abcd1234-") and len(path.stem.split("-")[1]) == 8
    repo = R.LocalRepo(tmp_path)
    assert repo.has_model("abcd1234")
    got = repo.get_model("abcd1234")
    assert isinstance(got, D.HTDemucs) and got.sources == m.sources and float(got.segment) == float(m.segment)
    for (k, a), (_, b) in zip(m.state_dict().items(), got.state_dict().items()):
        assert b.dtype == torch.float32 and torch.equal(b, a.half().float()), k     # fp16 package, fp32 parameters
    # a corrupted file is refused (repo.py:26-39)
    raw = bytearray(path.read_bytes())
    raw[len(raw) // 2] ^= 0xFF
    path.write_bytes(bytes(raw))
    with pytest.raises(R.ModelLoadingError, match="Invalid checksum"):
        R.LocalRepo(tmp_path).get_model("abcd1234")


def test_reference_class_reference_is_remapped(tmp_path):
    """A package written by the REFERENCE pickles ``demucs.htdemucs.HTDemucs``; it must load without that package."""
    m = _model()
    pkg = R.serialize_model(m, half=False)
    fake_root, fake = types.ModuleType("demucs"), types.ModuleType("demucs.htdemucs")

    class HTDemucs:        # what the pickle will name
        pass
    HTDemucs.__module__, HTDemucs.__qualname__ = "demucs.htdemucs", "HTDemucs"
    fake.HTDemucs = HTDemucs
    sys.modules["demucs"], sys.modules["demucs.htdemucs"] = fake_root, fake
    try:
        pkg["klass"] = HTDemucs
        pkg["kwargs"] = dict(pkg["kwargs"], some_future_option=3)       # dropped with a warning when strict=False
        torch.save(pkg, tmp_path / "ref.th")
    finally:
        del sys.modules["demucs"], sys.modules["demucs.htdemucs"]
    with pytest.warns(UserWarning, match="Dropping inexistant parameter some_future_option"):
        got = R.load_model(tmp_path / "ref.th")
    assert isinstance(got, D.HTDemucs)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), got.state_dict().values()))


def test_unsupported_and_malicious_packages_are_refused(tmp_path):
    fake_root, fake = types.ModuleType("demucs"), types.ModuleType("demucs.demucs")

    class Demucs:          # the v1 / v2 time-domain model (members of the mdx bags)
        pass
    Demucs.__module__, Demucs.__qualname__ = "demucs.demucs", "Demucs"
    fake.Demucs = Demucs
    sys.modules["demucs"], sys.modules["demucs.demucs"] = fake_root, fake
    try:
        torch.save({"klass": Demucs, "args": (), "kwargs": {}, "state": {}}, tmp_path / "v2.th")
    finally:
        del sys.modules["demucs"], sys.modules["demucs.demucs"]
    with pytest.raises(R.ModelLoadingError, match="outside the accelerated path"):
        R.load_model(tmp_path / "v2.th")
    # a Hybrid Demucs v3 package loads into demucs_b200.HDemucs
    from demucs_b200 import hdemucs as HD
    from oracle.make_golden import hdemucs_small_config
    v3 = HD.HDemucs.from_config(hdemucs_small_config(), init_seed=1)
    path = R.save_with_checksum(R.serialize_model(v3), tmp_path / "cafe0001.th")
    got = R.LocalRepo(tmp_path).get_model("cafe0001")
    assert isinstance(got, HD.HDemucs) and got.cfg.dconv_comp == 4
    assert torch.equal(got.state_dict()["decoder.0.norm2.weight"], v3.state_dict()["decoder.0.norm2.weight"].half().float())
    import os
    torch.save({"klass": os.system, "args": (), "kwargs": {}, "state": {}}, tmp_path / "evil.th")
    with pytest.raises(Exception, match="no use for"):
        R.load_model(tmp_path / "evil.th")
    pkg = R.serialize_model(_model())
    pkg["state"]["__quantized"] = True
    with pytest.raises(R.ModelLoadingError, match="DiffQ"):
        R.load_model(pkg)


def test_bag_yaml_and_get_model(tmp_path):
    sigs = []
    for seed in range(2):
        sig = f"{seed:08x}"
        R.save_with_checksum(R.serialize_model(_model(seed)), tmp_path / f"{sig}.th")
        sigs.append(sig)
    (tmp_path / "mybag.yaml").write_text(yaml.safe_dump({"models": sigs, "weights": [[1., 0., 1.], [0., 1., 1.]]}))
    bag = R.get_model("mybag", repo=tmp_path)
    assert isinstance(bag, D.BagOfModels) and len(bag.models) == 2 and bag.weights == [[1., 0., 1.], [0., 1., 1.]]
    single = R.get_model(sigs[1], repo=tmp_path)
    assert torch.equal(single.state_dict()["freq_emb.embedding.weight"],
                       _model(1).state_dict()["freq_emb.embedding.weight"].half().float())
    with pytest.raises(R.ModelLoadingError):
        R.get_model("nope", repo=tmp_path)
    # the remote zoo without the network: a clear error, not a download attempt and not random weights
    with pytest.raises(R.ModelLoadingError, match="does not download"):
        R.get_model("htdemucs")
    with pytest.raises(D.api.LoadModelError):
        D.Separator("htdemucs", device="cpu")
    listing = D.api.list_models(tmp_path)
    assert set(listing["single"]) == set(sigs) and "mybag" in listing["bag"]


def test_released_htdemucs_kwargs_are_accepted():
    """The released htdemucs packages record t_dropout=0.02 (grids/mmi.py:22), dconv_mode=3, bottom_channels=512."""
    cfg = D.config.HTDemucsConfig.from_reference_kwargs(
        sources=["drums", "bass", "other", "vocals"], dconv_mode=3, t_dropout=0.02, bottom_channels=512, t_layers=5,
        segment=7.8, t_weight_decay=0.05, rescale=0.1)
    assert cfg.transformer_dim == 512 and cfg.dconv_mode == 3
