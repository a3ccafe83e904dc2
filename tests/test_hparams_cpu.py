"""CPU: the PyYAML-only HyperPyYAML subset loader and the recipe plug-in contract."""
import os

import pytest
import torch

from conftest import ROOT
from ml_vae_b200.hparams import load_hyperpyyaml, recursive_update

CFG = os.path.join(ROOT, "ml_vae_b200", "config", "run_b200.yaml")


def _load(extra=None):
    overrides = {"dataset": "synthetic", "model_class": "b200_vanilla_vae", "model_name": "unit"}
    ov = [extra or {}, overrides, "model: !include:../models/b200_vanilla_vae/model.yaml"]
    with open(CFG) as f:
        return load_hyperpyyaml(f, ov)


def test_run_yaml_loads_and_injects_children():
    hp = _load()
    assert hp["seed"] == 123456 and hp["output_dir"] == "results/unit"
    m = hp["model"]
    assert m["input_size"] == 120 and m["batch_size"] == 8 and m["model_name"] == "unit"   # <n_mels> * 3, injected
    assert m["latent_size"] == 32 and m["kld_weight"] == 0.001
    enc, dec = m["encoder"], m["decoder"]
    assert m["modules"]["encoder"] is enc and m["modules"]["decoder"] is dec                # !ref -> same object
    assert enc.fc_sizes == [120, 64, 64] and dec.fc_sizes == [1024, 64, 64, 120]
    assert hp["compute_features"].hop == 320 and hp["compute_features"].feature_dim == 120
    opt = m["optimizer"]([torch.nn.Parameter(torch.zeros(1))])
    assert isinstance(opt, torch.optim.Adam) and opt.defaults["lr"] == 0.001
    assert m["epoch_counter"].limit == 50


def test_seed_is_applied_before_modules_are_built():
    a = _load()["model"]["encoder"].state_dict()["mean_fc.weight"]
    b = _load()["model"]["encoder"].state_dict()["mean_fc.weight"]
    assert torch.equal(a, b)
    c = _load({"seed": 7})["model"]["encoder"].state_dict()["mean_fc.weight"]
    assert not torch.equal(a, c)


def test_overrides_list_and_extra_overrides():
    hp = _load({"n_mels": 80, "hop_length": 10, "model": {"latent_size": 64, "n_epochs": 1}})
    assert hp["model"]["input_size"] == 240 and hp["model"]["latent_size"] == 64
    assert hp["model"]["encoder"].latent_size == 64 and hp["model"]["epoch_counter"].limit == 1
    recursive_update(hp, {"model": {"n_epochs": 3}})            # prepare_experiment.py:25
    assert hp["model"]["n_epochs"] == 3


def test_placeholder_must_be_overridden():
    with open(CFG) as f, pytest.raises(ValueError, match="PLACEHOLDER"):
        load_hyperpyyaml(f, {"dataset": "x"})


def test_reference_yaml_files_load_when_present():
    """In the build container: the reference's own run.yaml + test_vanilla_vae/model.yaml load through
    this loader once the SpeechBrain classes are swapped for the drop-ins (pure yaml overrides)."""
    ref = "/root/reference/src/config/run.yaml"
    if not os.path.exists(ref):
        pytest.skip("reference checkout not present (GPU box)")
    swap = """
compute_features: !new:ml_vae_b200.features.Fbank
    deltas: True
    sample_rate: !ref <sample_rate>
    hop_length: !ref <hop_length>
    n_fft: !ref <n_fft>
    n_mels: !ref <n_mels>
prepare: null
"""
    inner = {"model": {"epoch_counter": None, "checkpointer": None, "normalizer": None,
                       "encoder": None, "decoder": None, "modules": None}}
    with open(ref) as f:
        hp = load_hyperpyyaml(f, [swap, {"dataset": "d", "model_class": "test_vanilla_vae", "model_name": "n"},
                                  "model: !include:../models/test_vanilla_vae/model.yaml", inner])
    assert hp["model"]["input_size"] == 120 and hp["model"]["kld_weight"] == 0.001
    assert hp["model"]["dec_rnn_hidden_size"] == 512 and hp["kaldi_feature_params"]["hop_length"] == 20
    assert hp["compute_features"].n_mels == 40


def test_recipe_plugin_contract():
    """prepare_experiment.py:47-57: models.<class>.model.SBModel(label_encoder=, modules=, hparams=,
    run_opts=, checkpointer=) with compute_forward / compute_objectives / fit_batch."""
    import importlib
    mod = importlib.import_module("ml_vae_b200.models.b200_vanilla_vae.model")
    assert hasattr(mod, "SBModel")
    for name in ("compute_forward", "compute_objectives", "fit_batch", "compute_and_save_losses", "init_optimizers"):
        assert callable(getattr(mod.SBModel, name))


def test_recipe_declares_the_checkpointer_the_reference_reads(tmp_path):
    """prepare_experiment.py:56 reads hparams['model']['checkpointer']; the stand-in saves / restores every recoverable
    (modules by their reference state_dict keys, the epoch counter) and keeps the best checkpoint by min_key."""
    hp = _load({"output_dir": str(tmp_path)})
    m = hp["model"]
    ck = m["checkpointer"]
    assert set(ck.recoverables) >= {"encoder", "decoder", "epoch_counter"}
    assert ck.recoverables["encoder"] is m["encoder"] and ck.recoverables["epoch_counter"] is m["epoch_counter"]
    opt = m["optimizer"](list(m["encoder"].parameters()))
    ck.add_recoverable("optimizer", opt)                                      # md_model.py:50-52
    next(iter(m["epoch_counter"]))
    want = {k: v.clone() for k, v in m["encoder"].state_dict().items()}
    ck.save_and_keep_only(meta={"loss": 2.0}, min_keys=["loss"])
    with torch.no_grad():
        for p in m["encoder"].parameters():
            p.add_(1.0)
    m["epoch_counter"].current = 7
    ck.save_and_keep_only(meta={"loss": 3.0}, min_keys=["loss"])             # worse: the best one must survive
    path, meta = ck.recover_if_possible(min_key="loss")
    assert meta["loss"] == 2.0 and m["epoch_counter"].current == 1
    for k, v in m["encoder"].state_dict().items():
        assert torch.equal(v, want[k]), k
