"""Runs the REFERENCE's own training script (train_confignet.py:14-78, the body of its test tests/training_test.py:13-24)
on the reference's own test dataset (tests/test_assets/test_dataset_res_256.pck + _imgs.dat, loaded by the reference's own
NeuralRendererDataset.load / process_metadata) with the one-import switch INTEGRATION.md describes: the `confignet` package
exports the reference's dataset / utility modules and the PRODUCT's ConfigNetFirstStage / ConfigNet.

This container has no GPU and the product has no CPU fallback, so the run is done twice:

  1. as is, on device "cpu": the script must get through argument parsing, dataset loading, config merging with the
     reference's DEFAULT_CONFIG, model construction, setup_training and the host half of the first discriminator step
     (NumPy batch assembly from the memory-mapped image store) and then fail LOUDLY at the first kernel call (CnError);
  2. with the four step methods replaced by recorders: the rest of the script's calls must fit the product's class surface -
     train() signatures of both stages, get_weights(), the unbound ConfigNetFirstStage.set_weights(second_stage_model, ...),
     the image-loss-weight bump - and leave the reference's output layout behind (first_stage/checkpoints/000000.{npz,json,
     _facemodel_distr.pck} and the loss tables; the step-0 checkpoint of the second stage under <output_dir>/checkpoints).

Build container only (/root/reference is read at run time):  python scripts/run_reference_training_script_on_product.py
"""
import importlib
import json
import os
import pickle
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return sys.modules.get(self.__name__ + "." + name, _Anything)


for m in ["tensorflow", "tensorflow.keras", "tensorflow.keras.applications", "tensorflow.keras.layers", "tensorflow.keras.models",
          "tensorflow.keras.backend", "tensorflow.keras.utils", "tensorflow.keras.initializers", "tensorflow.keras.regularizers",
          "tensorflow.keras.constraints", "tensorflow.compat", "tensorflow.compat.v1", "cv2",
          "matplotlib", "matplotlib.pyplot", "transformations", "azureml", "azureml.core", "azureml.core.run"]:
    parts = m.split(".")
    for i in range(1, len(parts) + 1):
        sys.modules.setdefault(".".join(parts[:i]), _Stub(".".join(parts[:i])))


class _OfflineRun:
    pass


sys.modules["azureml.core.run"].Run = types.SimpleNamespace(get_context=lambda: _OfflineRun())     # "running locally"
sys.modules["azureml.core.run"]._OfflineRun = _OfflineRun
sys.modules["tensorflow"].compat = types.SimpleNamespace(v1=types.SimpleNamespace(set_random_seed=lambda seed: None))

import confignet_b200                                               # noqa: E402
from confignet_b200 import _lib as L                                # noqa: E402

# ---- the `confignet` package a maintainer gets after the one-import switch (confignet/__init__.py:3-5 -> confignet_b200)
pkg = types.ModuleType("confignet")
pkg.__path__ = [os.path.join(REF, "confignet")]
sys.modules["confignet"] = pkg
mpkg = types.ModuleType("confignet.metrics")
mpkg.__path__ = [os.path.join(REF, "confignet", "metrics")]
sys.modules["confignet.metrics"] = mpkg
pkg.azure_ml_utils = importlib.import_module("confignet.azure_ml_utils")
pkg.confignet_utils = importlib.import_module("confignet.confignet_utils")
sys.modules["confignet.perceptual_loss"] = _Stub("confignet.perceptual_loss")
pkg.NeuralRendererDataset = importlib.import_module("confignet.neural_renderer_dataset").NeuralRendererDataset
# train_confignet.py:12 reads DEFAULT_CONFIG from the reference's first-stage module: give it the module with the
# reference's own table (imported for that constant only - its classes are not used)
ref_first = importlib.import_module("confignet.confignet_first_stage")
pkg.ConfigNetFirstStage = confignet_b200.ConfigNetFirstStage
pkg.ConfigNet = confignet_b200.ConfigNet
pkg.LatentGAN = confignet_b200.LatentGAN

sys.path.insert(0, REF)
train_confignet = importlib.import_module("train_confignet")
ASSETS = os.path.join(REF, "tests", "test_assets")


def script_args(out_dir):
    ds = os.path.join(ASSETS, "test_dataset_res_256.pck")
    a = "--real_training_set_path %s --synth_training_set_path %s --output_dir %s" % (ds, ds, out_dir)
    a += " --validation_set_path %s --attribute_classifier_path %s" % (ds, os.path.join(out_dir, "no_classifier.json"))
    a += " --stage_1_training_steps 1 --stage_2_training_steps 1 --batch_size 4 --n_samples_for_metrics 10"
    return a.split(" ")


def main():
    # the product classes default to cuda:0; this container has no GPU, so the switch also pins the device
    for cls in (confignet_b200.ConfigNetFirstStage, confignet_b200.ConfigNet):
        init = cls.__init__
        cls.__init__ = (lambda init: lambda self, config, initialize=True, **kw: init(self, config, initialize,
                                                                                      **dict(kw, device="cpu")))(init)

    # ---- 1. unmodified: everything up to the first kernel, then a loud failure
    out1 = tempfile.mkdtemp(prefix="cn_train1_")
    trace = []
    first = confignet_b200.ConfigNetFirstStage
    orig_setup, orig_sample = first.setup_training, first.sample_latent_vector
    first.setup_training = lambda self, *a, **k: (trace.append("setup_training"), orig_setup(self, *a, **k))[1]
    try:
        train_confignet.parse_args(script_args(out1))
        raise SystemExit("the product ran a training step without CUDA - there must be no CPU fallback")
    except L.CnError as e:
        import traceback
        frames = [f.name for f in traceback.extract_tb(e.__traceback__)]
        assert "parse_args" in frames and "train" in frames and "discriminator_training_step" in frames, frames
        assert trace == ["setup_training"]
        print("1. script reached the first kernel of discriminator_training_step and failed loudly: %s" % str(e).splitlines()[0][:100])
    finally:
        first.setup_training = orig_setup

    # ---- 2. step methods replaced by recorders: the rest of the script against the product's class surface
    calls = []

    def recorder(name):
        def f(self, *args):
            sets = [type(a).__name__ for a in args[:-1]]
            calls.append([type(self).__name__, name] + sets)
            return {"loss_sum": 1.0, "term": 0.5}
        return f
    for cls in (confignet_b200.ConfigNetFirstStage, confignet_b200.ConfigNet):
        for name in ("discriminator_training_step", "synth_discriminator_training_step", "latent_discriminator_training_step",
                     "generator_training_step"):
            setattr(cls, name, recorder(name))
        cls.update_smoothed_weights = lambda self, smoother_alpha=0.999: calls.append([type(self).__name__, "update_smoothed_weights"])
    # calculate_metrics (KID / FID, controllability, perceptual loss) needs the GPU networks: recorded here, executed by
    # tests/test_metrics_gpu.py.  setup_training builds the metric objects for real (host half only).
    metric_calls = []
    for cls in (confignet_b200.ConfigNetFirstStage, confignet_b200.ConfigNet):
        cls.calculate_metrics = (lambda cls: lambda self, output_dir, aml_run=None: metric_calls.append(
            [cls.__name__, self.get_training_step_number(), type(self._inception_metric_object).__name__,
             type(getattr(self, "controllability_metrics", None)).__name__]))(cls)
    seen_cfg = []
    orig_train2 = confignet_b200.ConfigNet.train
    confignet_b200.ConfigNet.train = lambda self, *a, **k: (seen_cfg.append(dict(self.config)), orig_train2(self, *a, **k))[1]
    out2 = tempfile.mkdtemp(prefix="cn_train2_")
    # the attribute classifier the script hands to stage 2 (--attribute_classifier_path): a file pair in the reference's format
    from confignet_b200.metrics import CelebaAttributeClassifier
    CelebaAttributeClassifier({"input_shape": [128, 128, 3], "predicted_attributes": ["Smiling", "Young"]}, device="cpu").save(out2, "no_classifier")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")             # stand-in InceptionV3 weights (no download offline)
        train_confignet.parse_args(script_args(out2))
    assert metric_calls == [["ConfigNetFirstStage", 0, "InceptionMetrics", "NoneType"],
                            ["ConfigNet", 0, "InceptionMetrics", "ControllabilityMetrics"]], metric_calls
    names = [c[:2] for c in calls]
    assert names == [["ConfigNetFirstStage", "discriminator_training_step"], ["ConfigNetFirstStage", "synth_discriminator_training_step"],
                     ["ConfigNetFirstStage", "latent_discriminator_training_step"], ["ConfigNetFirstStage", "generator_training_step"],
                     ["ConfigNetFirstStage", "update_smoothed_weights"],
                     ["ConfigNet", "discriminator_training_step"], ["ConfigNet", "synth_discriminator_training_step"],
                     ["ConfigNet", "latent_discriminator_training_step"], ["ConfigNet", "generator_training_step"],
                     ["ConfigNet", "update_smoothed_weights"]], names
    assert all(s == "NeuralRendererDataset" for c in calls for s in c[2:])
    assert calls[7][2:] == ["NeuralRendererDataset"] * 2                      # stage-2 latent-D step: real and synthetic set
    cfg2 = seen_cfg[0]
    assert cfg2["model_type"] == "ConfigNet" and cfg2["batch_size"] == 4 and tuple(cfg2["output_shape"]) == (256, 256, 3)
    assert abs(cfg2["image_loss_weight"] - 10 * ref_first.DEFAULT_CONFIG["image_loss_weight"]) < 1e-12     # train_confignet.py:68
    dims = {k: tuple(v) for k, v in cfg2["facemodel_inputs"].items()}
    assert dims["blendshape_values"][0] == 62 and dims["texture_embedding"][0] == 50, dims   # process_metadata(config, True)
    files = sorted(os.listdir(os.path.join(out2, "first_stage", "checkpoints")))
    assert files == ["000000.json", "000000.npz", "000000_facemodel_distr.pck"], files
    assert sorted(os.listdir(os.path.join(out2, "checkpoints"))) == files
    for d in (os.path.join(out2, "first_stage"), out2):
        for prefix in ("generator_", "discriminator_", "synth_discriminator_", "latent_discriminator_"):
            t = np.loadtxt(os.path.join(d, prefix + "losses.txt"), ndmin=2)
            assert t.shape == (1, 2) and t[0].tolist() == [1.0, 0.5]
    with open(os.path.join(out2, "checkpoints", "000000_facemodel_distr.pck"), "rb") as fp:
        distr = pickle.load(fp)                                               # the reference's fitted distributions travel with the model
    assert sorted(distr.keys()) == sorted(dims.keys()) and distr["blendshape_values"].sample(3)[0].shape == (3, 62)
    print("2. both stages ran through the reference script: %d recorded calls, checkpoints and loss tables in the reference's layout" % len(calls))
    # ---- 3. train_latent_gan.py:11-50 on the second-stage checkpoint phase 2 left behind: the reference's own
    #         confignet_utils.load_confignet resolves config["model_type"] in the `confignet` package -> the product's ConfigNet.load
    pkg.load_confignet = pkg.confignet_utils.load_confignet
    train_latent_gan = importlib.import_module("train_latent_gan")
    gan_init = confignet_b200.LatentGAN.__init__
    confignet_b200.LatentGAN.__init__ = lambda self, config, **kw: gan_init(self, config, **dict(kw, device="cpu"))
    out3 = tempfile.mkdtemp(prefix="cn_train3_")
    gan_args = ("--confignet_path %s --training_set_path %s --output_dir %s --n_training_steps 1 --batch_size 4 --n_samples_for_metrics 10"
                % (os.path.join(out2, "checkpoints", "000000.json"), os.path.join(ASSETS, "test_dataset_res_256.pck"), out3)).split(" ")
    try:
        train_latent_gan.parse_args(gan_args)
        raise SystemExit("the product encoded images without CUDA - there must be no CPU fallback")
    except L.CnError as e:
        import traceback
        frames = [f.name for f in traceback.extract_tb(e.__traceback__)]
        assert "extract_embeddings" in frames and "encode_images" in frames, frames
        print("3. train_latent_gan.py loaded the product checkpoint through the reference's load_confignet and reached the "
              "first kernel of encode_images: %s" % str(e).splitlines()[0][:80])
    gcalls = []
    real_encode_images = confignet_b200.ConfigNet.encode_images
    confignet_b200.ConfigNet.encode_images = lambda self, imgs: (np.zeros((imgs.shape[0], 145), np.float32), np.zeros((imgs.shape[0], 3), np.float32))
    confignet_b200.LatentGAN.discriminator_training_step = lambda self, emb, opt: gcalls.append(["d", tuple(emb.shape)]) or {"loss_sum": 1.0}
    confignet_b200.LatentGAN.generator_training_step = lambda self, opt: gcalls.append(["g"]) or {"loss_sum": 2.0}
    confignet_b200.LatentGAN.update_smoothed_weights = lambda self, smoother_alpha=0.999: gcalls.append(["ema"])
    train_latent_gan.parse_args(gan_args)
    assert gcalls == [["d", (2, 145)], ["g"], ["ema"]], gcalls                  # the test dataset holds two images
    assert sorted(os.listdir(os.path.join(out3, "checkpoints"))) == ["000000.json", "000000.npz"]
    print("   ... and, with the steps recorded, ran its loop and left checkpoints/000000.{json,npz}")
    # ---- 4. evaluation/evaluate_confignet_controllability.py (the body of tests/evaluation_test.py::test_confignet_evaluation)
    #         on the stage-2 checkpoint of phase 2 and the classifier file pair: confignet.ControllabilityMetrics is the
    #         product's class.  Unmodified it must fail loudly at the first kernel; with the four device entry points it calls
    #         replaced by deterministic stand-ins its host logic runs through and leaves the reference's output files.
    pkg.ControllabilityMetrics = confignet_b200.ControllabilityMetrics
    sys.path.insert(0, os.path.join(REF, "evaluation"))
    evaluate = importlib.import_module("evaluate_confignet_controllability")
    out4 = tempfile.mkdtemp(prefix="cn_eval4_")
    confignet_b200.ConfigNet.encode_images = real_encode_images                  # phase 3's stand-in: back to the real method
    for n_iters in (0, 1):
        eval_args = ("--model_path %s --test_set_path %s --output_dir %s --attribute_classifier_path %s --n_samples 2 "
                     "--n_fine_tuning_iters %d --write_images"
                     % (os.path.join(out2, "checkpoints", "000000.json"), os.path.join(ASSETS, "test_dataset_res_256.pck"), out4,
                        os.path.join(out2, "no_classifier.json"), n_iters)).split(" ")
        if n_iters == 0:
            try:
                evaluate.parse_args(eval_args)
                raise SystemExit("the product computed metrics without CUDA - there must be no CPU fallback")
            except L.CnError as e:
                import traceback
                frames = [f.name for f in traceback.extract_tb(e.__traceback__)]
                assert "get_metrics" in frames and "generate_images_for_metric" in frames and "encode_images" in frames, frames
                print("4. evaluate_confignet_controllability.py built the product's ControllabilityMetrics and reached the first "
                      "kernel of encode_images: %s" % str(e).splitlines()[0][:70])
            ecalls = []
            rs = np.random.RandomState(0)
            confignet_b200.ConfigNet.encode_images = lambda self, imgs: (ecalls.append(["encode", len(imgs)]), (rs.randn(len(imgs), 145).astype(np.float32), np.zeros((len(imgs), 3), np.float32)))[1]
            confignet_b200.ConfigNet.fine_tune_on_img = lambda self, imgs, n_iters=50, **kw: (ecalls.append(["fine_tune", len(imgs), n_iters]), (rs.randn(len(imgs), 145).astype(np.float32), np.zeros((len(imgs), 3), np.float32)))[1]
            confignet_b200.ConfigNet.generate_images = lambda self, lat, rot: (ecalls.append(["generate", len(lat)]), rs.randint(0, 256, (len(lat), 256, 256, 3)).astype(np.uint8))[1]
            from confignet_b200.confignet_first_stage import SyntheticEncoderNet
            SyntheticEncoderNet.predict = lambda self, params: (ecalls.append(["synthetic_encoder", len(params)]), rs.randn(1, 145).astype(np.float32))[1]
            from confignet_b200.metrics import CelebaAttributeClassifier as Clf
            Clf.predict_attributes = lambda self, imgs: (ecalls.append(["predict_attributes", len(imgs)]), rs.rand(len(imgs), len(self.config["predicted_attributes"])).astype(np.float32))[1]
            # the stand-in classifier of phase 2 predicts two attributes; the metric configurations name CelebA's
            with open(os.path.join(out2, "no_classifier.json")) as fp:
                meta = json.load(fp)
            meta["config"]["predicted_attributes"] = ["Black_Hair", "Blond_Hair", "Brown_Hair", "Gray_Hair", "Mouth_Slightly_Open", "Smiling",
                                                      "Narrow_Eyes", "Mustache", "No_Beard", "Goatee", "Sideburns", "Young"]
            Clf({"input_shape": [128, 128, 3], "predicted_attributes": meta["config"]["predicted_attributes"]}, device="cpu").save(out2, "no_classifier")
        ecalls.clear()
        evaluate.parse_args(eval_args)
        name = "contr_metrics_tuning_iters_%d_000000" % n_iters
        produced = sorted(os.listdir(out4))
        assert name + ".json" in produced and name + ".csv" in produced and name in produced, produced
        with open(os.path.join(out4, name + ".json")) as fp:
            m = json.load(fp)
        assert len(m) == 10 and "controllability" in m and len(m["smile_config"]) == 4, sorted(m)
        assert np.loadtxt(os.path.join(out4, name + ".csv"), delimiter=",").shape == (4, 9)          # 8 configurations + their mean
        assert os.path.isdir(os.path.join(out4, name))        # --write_images (cv2 is a stub in this harness; the PNG dump itself: tests/test_metrics_cpu.py)
        kinds = [c[0] for c in ecalls]
        if n_iters == 0:
            assert kinds.count("encode") == 1 and kinds.count("generate") == 1 + 16 and kinds.count("predict_attributes") == 16, kinds
        else:
            assert kinds.count("fine_tune") == 2 and kinds.count("generate") == 2 * (1 + 16) and ["fine_tune", 1, 1] in ecalls, kinds
    print("   ... and, with the device entry points replaced, wrote contr_metrics_tuning_iters_{0,1}_000000.{json,csv}")
    import shutil
    shutil.rmtree(out4, ignore_errors=True)
    shutil.rmtree(out3, ignore_errors=True)
    shutil.rmtree(out1, ignore_errors=True)
    shutil.rmtree(out2, ignore_errors=True)
    print("reference training script OK")


if __name__ == "__main__":
    main()
