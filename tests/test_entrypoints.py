"""options/* + test.py / train.py entry points (reference API surface: flags, defaults, synonyms)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_options_defaults_and_synonyms():
    from shineon_virtual_tryon_b200.options import TestOptions, TrainOptions

    o = TestOptions().parse(["--model", "tom", "--name", "x"])
    assert o.model == "unet_mask" and o.person_inputs == ["agnostic", "densepose"] and o.cloth_inputs == ["cloth"]
    assert (o.ngf, o.num_attn, o.self_attn, o.flow_warp, o.batch_size, o.workers) == (64, 2, False, False, 8, 4)
    assert (o.datamode, o.result_dir, o.is_train, o.gpu_ids, o.n_frames_now) == ("test", "test_results", False, [0], 1)
    o = TrainOptions().parse(["--model", "gmm", "--name", "x", "--gpu_ids", "0,1", "-b", "4", "--accumulated_batches", "16"])
    assert o.model == "warp" and o.person_inputs == ["agnostic", "cocopose"] and o.grid_size == 5
    assert (o.lr, o.keep_epochs, o.decay_epochs, o.accumulated_batches, o.gpu_ids, o.is_train) == (1e-4, 5, 5, 16, [0, 1], True)
    o = TrainOptions().parse(["--model", "unet", "--name", "x", "--self_attn", "--activation", "gelu", "--flow_warp",
                              "--n_frames_total", "5"])
    assert (o.self_attn, o.activation, o.flow_warp, o.n_frames_total, o.n_frames_now, o.pen_flow_mask) == (True, "gelu", True, 5, 5, 1.0)


def test_train_entry_point_rejects_models_without_native_training():
    sys.path.insert(0, ROOT)
    import train

    with pytest.raises(SystemExit) as e:
        train.main(["--model", "gmm", "--name", "x"])
    assert "U-Net try-on stage" in str(e.value)


def test_train_entry_point_refuses_datasets_it_does_not_have():
    """--dataset vvt must not silently train on synthetic noise."""
    sys.path.insert(0, ROOT)
    import train

    with pytest.raises(SystemExit) as e:
        train.main(["--model", "unet", "--name", "x", "--dataset", "vvt"])
    assert "--dataset vvt" in str(e.value) and "synthetic" in str(e.value)


@pytest.mark.gpu
def test_train_entry_point_runs_on_synthetic_data(cuda, tmp_path):
    sys.path.insert(0, ROOT)
    import warnings

    import torch
    import train

    args = ["--model", "unet", "--name", "smoke", "--self_attn", "--activation", "gelu", "-b", "2", "--synthetic_samples", "4",
            "--accumulated_batches", "2", "--experiments_dir", str(tmp_path), "--save_count", "1"]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        rc = train.main(args + ["--max_steps", "2"])
    assert rc == 0
    assert any("RANDOM weights" in str(x.message) for x in w), "training on a random VGG19 must warn"
    ck = torch.load(os.path.join(str(tmp_path), "smoke", "final.ckpt"), map_location="cpu")
    # Lightning-readable (state_dict + hyper_parameters + global_step) and resumable (Adam moments, step, epoch)
    assert {"state_dict", "hyper_parameters", "global_step", "epoch", "b200_optimizer"} <= set(ck)
    assert ck["global_step"] == 2 and ck["hyper_parameters"]["activation"] == "gelu"
    assert ck["b200_optimizer"]["exp_avg_sq"].abs().sum() > 0
    assert os.path.exists(os.path.join(str(tmp_path), "smoke", "step_0000001.ckpt"))  # --save_count
    # resume: the optimiser continues at step 2 (bias correction / schedule), not from scratch
    rc = train.main(args + ["--max_steps", "3", "--checkpoint", os.path.join(str(tmp_path), "smoke", "final.ckpt")])
    assert rc == 0
    ck2 = torch.load(os.path.join(str(tmp_path), "smoke", "final.ckpt"), map_location="cpu")
    assert ck2["global_step"] == 3


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["gmm", "tom"])
def test_test_entry_point_runs_on_synthetic_data(cuda, model, tmp_path):
    sys.path.insert(0, ROOT)
    import importlib

    entry = importlib.import_module("test")
    rc = entry.main(["--model", model, "--name", "smoke", "-b", "2", "--synthetic_samples", "4", "--self_attn",
                     "--activation", "gelu", "--result_dir", str(tmp_path)])
    assert rc == 0
    files = os.listdir(os.path.join(str(tmp_path), "smoke", "test"))
    assert len(files) == 4 and all(f.endswith(".png") for f in files)
