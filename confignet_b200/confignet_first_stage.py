"""ConfigNetFirstStage on B200: the class surface of the reference
(confignet/confignet_first_stage.py:86-679) over our CUDA networks.

Same constructor / config merging / latent layout / method names / loss-dictionary keys, NumPy in and
NumPy out at the public boundary, so train_confignet.py-style callers work unchanged.  What differs is
underneath: every network call is a sequence of launches of libconfignet_b200.so, the optimizer is one
fused kernel per network over a flat buffer, the EMA stays on the device, and with torch.distributed
initialised each step all-reduces its flat gradient buffer over NCCL.

Out of scope here: TensorBoard/AzureML logging, image and metric
checkpoints.  ``train`` keeps the reference loop structure and loss history only.
"""
import os
import json
import pickle
import time
from collections import OrderedDict
import numpy as np
import torch

from . import netspec, networks, ops, pretrained
from .runtime import ParamGroup, Network, KerasAdam, StepGraphs, InferenceGraphs, shard_rows, world, coalesce_grads

DEFAULT_CONFIG = {
    "model_type": None,
    "latent_dim": 128,
    "output_shape": (128, 128, 3),
    "const_input_shape": (4, 4, 4, 512),
    "n_adain_mlp_layers": 2,
    "n_adain_mlp_units": 128,
    "gen_output_activation": "tanh",
    "n_discr_features_at_layer_0": 48,
    "max_discr_filters": 512,
    "n_discr_layers": 5,
    "discr_conv_kernel_size": 3,
    "latent_regression_weight": 10.0,
    "use_style_discriminator": True,
    "rotation_ranges": ((-30, 30), (-10, 10), (0, 0)),
    "relu_before_in": True,
    "initial_from_rgb_layer_in_discr": True,
    "adain_on_learned_input": False,
    "latent_regressor_rot_weight": 5.0,
    "optimizer": {"lr": 0.0004, "beta_1": 0.0, "beta_2": 0.9, "amsgrad": False},
    "batch_size": 24,
    "n_discriminator_updates": 1,
    "n_generator_updates": 1,
    "latent_distribution": "normal",
    "metrics_checkpoint_period": 1000,
    "image_checkpoint_period": 500,
    "facemodel_inputs": {
        "texture_embedding": (None, 30),
        "geometry_identity_params": (None, 30),
        "blendshape_values": (None, 30),
        "beard_style_embedding": (None, 7),
        "eyebrow_style_embedding": (None, 7),
        "lower_eyelash_style": (None, 2),
        "upper_eyelash_style": (None, 2),
        "head_hair_style_embedding": (None, 9),
        "eye_color": (None, 3),
        "head_hair_color": (None, 3),
        "hdri_embedding": (None, 20),
        "bone_rotations:left_eye": (None, 2),
    },
    "num_synth_encoder_layers": 2,
    "n_latent_discr_layers": 4,
    "image_loss_weight": 0.00005,
    "eye_loss_weight": 5,
    "domain_adverserial_loss_weight": 5.0,
}


def merge_configs(default_config, input_config):
    """confignet_utils.py:39-61 (recursive dictionary merge, input wins)."""
    result = {}
    for name in default_config:
        lhs = default_config[name]
        if name in input_config:
            rhs = input_config[name]
            if isinstance(lhs, dict):
                assert isinstance(rhs, dict)
                result[name] = merge_configs(lhs, rhs)
            else:
                result[name] = rhs
        else:
            result[name] = lhs
    for name in input_config:
        rhs = input_config[name]
        if isinstance(rhs, dict) and name in default_config.keys():
            continue
        result[name] = rhs
    return result


def update_loss_dict(main_loss_dict, new_loss_dict):
    """confignet_utils.py:206-212."""
    for key, val in new_loss_dict.items():
        val = float(val)
        main_loss_dict.setdefault(key, []).append(val)


def _log_loss_vals(loss_dict, output_dir, prefix):
    """the text half of confignet_utils.py:214-241: one row per step, one column per loss term, names in the header"""
    if not loss_dict:
        return
    os.makedirs(output_dir, exist_ok=True)
    names = list(loss_dict.keys())
    np.savetxt(os.path.join(output_dir, prefix + "losses.txt"), np.stack(list(loss_dict.values()), axis=1),
               header="\t".join(names))


def flip_random_subset_of_images(images):
    """confignet_utils.py:198-204 (same NumPy RNG consumption; flips a copy view per image)."""
    flip_or_not = np.random.randint(0, 2, size=images.shape[0])
    for i, flip in enumerate(flip_or_not):
        if flip == 1:
            images[i] = np.fliplr(images[i])
    return images


class _PerParamMLP:
    """synthetic_encoder.per_facemodel_input_mlps[name] (synthetic_encoder.py:19-30): ``predict`` + ``num_in``."""

    def __init__(self, owner, name, num_in, num_layers):
        self.owner, self.name, self.num_in, self.num_layers = owner, name, num_in, num_layers

    def __call__(self, x):
        x = networks._as_dev(x, self.owner.group.device)
        return networks.mlp_fused(x, self.owner.group.params, "mlp_" + self.name, self.num_layers, alpha=0.3)

    def predict(self, x):
        with torch.no_grad():
            return self(x).cpu().numpy()


class SyntheticEncoderNet(Network):
    def __init__(self, group, facemodel_inputs, num_layers):
        super().__init__(group, networks.synthetic_encoder_forward, facemodel_inputs=facemodel_inputs, num_layers=num_layers)
        self.facemodel_param_names = list(facemodel_inputs.keys())
        self.per_facemodel_input_mlps = {n: _PerParamMLP(self, n, facemodel_inputs[n][0], num_layers)
                                         for n in self.facemodel_param_names}

    def __call__(self, inputs):
        dev = self.group.device
        if isinstance(inputs, dict):
            inputs = [inputs[n] for n in self.facemodel_param_names]
        if isinstance(inputs, (list, tuple)):
            inputs = [networks._as_dev(x, dev) for x in inputs]
        else:
            inputs = networks._as_dev(inputs, dev)
        return super().__call__(inputs)


class GeneratorNet(Network):
    def build_input_dict(self, latent_vector, rotation):
        """hologan_generator.py:109-127."""
        d = {}
        zs = latent_vector if isinstance(latent_vector, list) else [latent_vector] * 5
        for k, z in zip(("z_3d_0", "z_3d_1", "z_2d_0", "z_2d_1", "z_2d_2"), zs):
            d[k] = z
        d["rotation"] = rotation
        return d

    def __call__(self, inputs):
        if not isinstance(inputs, dict):
            inputs = self.build_input_dict(inputs[0], inputs[1])
        dev = self.group.device
        zs = [networks._as_dev(inputs[k], dev) for k in ("z_3d_0", "z_3d_1", "z_2d_0", "z_2d_1", "z_2d_2")]
        return super().__call__(None, inputs["rotation"], zs=zs)


GENERATE_CHUNK = 64     # images per generator pass of a large generate_images call (a per-sample network: the batch size only moves last bits)


class ConfigNetFirstStage(StepGraphs):
    def __init__(self, config, initialize=True, device=None, seed=1234):
        self.config = merge_configs(DEFAULT_CONFIG, config)
        self.config["model_type"] = "ConfigNetFirstStage"
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self._seed = seed

        self._graphs = {}                 # step name -> (optimizer, runtime.GraphedFn)
        self._infer = InferenceGraphs()   # captured generate_images per (network, batch)
        self.generator = None
        self.generator_smoothed = None
        self.discriminator = None
        self.latent_regressor = None
        self.latent_discriminator = None
        self.synth_discriminator = None
        self.synthetic_encoder = None
        self.perceptual_loss = None

        self.g_losses, self.d_losses, self.metrics = {}, {}, {}
        self.synth_d_losses, self.latent_d_losses = {}, {}
        self.n_checkpoint_rotations = 6
        self.n_checkpoint_samples = 10

        # confignet_first_stage.py:114-120
        self.config["facemodel_inputs"] = {k: tuple(v) for k, v in self.config["facemodel_inputs"].items() if v[0] is not None}
        self.config["facemodel_inputs"] = OrderedDict(sorted(self.config["facemodel_inputs"].items(), key=lambda t: t[0]))
        self.config["latent_dim"] = 0
        for spec in self.config["facemodel_inputs"]:
            self.config["latent_dim"] += self.config["facemodel_inputs"][spec][1]
        self.facemodel_param_distributions = None
        if initialize:
            self.initialize_network()

    # ---------------------------------------------------------------- construction
    def _discr_args(self):
        c = self.config
        return dict(output_res=c["output_shape"][0], n_layers=c["n_discr_layers"], base=c["n_discr_features_at_layer_0"],
                    max_maps=c["max_discr_filters"], ksize=c["discr_conv_kernel_size"],
                    from_rgb=c["initial_from_rgb_layer_in_discr"])

    def _make_group(self, spec, seed, vgg_like=False):
        return ParamGroup(netspec.init_params(spec, seed, vgg_like=vgg_like), self.device)

    def _get_generator_kwargs(self):
        """confignet_first_stage.py:240-248."""
        c = self.config
        return {"latent_dim": c["latent_dim"], "output_shape": tuple(c["output_shape"][:2]),
                "n_adain_mlp_units": c["n_adain_mlp_units"], "n_adain_mlp_layers": c["n_adain_mlp_layers"],
                "gen_output_activation": c["gen_output_activation"]}

    def initialize_network(self):
        """confignet_first_stage.py:251-287."""
        c = self.config
        s = self._seed
        res = c["output_shape"][0]
        nl = c["n_discr_layers"]
        self.synthetic_encoder = SyntheticEncoderNet(
            self._make_group(netspec.synthetic_encoder_spec(c["facemodel_inputs"], c["num_synth_encoder_layers"]), s + 1),
            c["facemodel_inputs"], c["num_synth_encoder_layers"])
        self.discriminator = Network(self._make_group(netspec.discriminator_spec(**self._discr_args()), s + 2),
                                     networks.discriminator_forward, n_layers=nl)
        self.synth_discriminator = Network(self._make_group(netspec.discriminator_spec(**self._discr_args()), s + 3),
                                           networks.discriminator_forward, n_layers=nl)
        self.latent_discriminator = Network(
            self._make_group(netspec.latent_discriminator_spec(c["latent_dim"], c["n_latent_discr_layers"]), s + 4),
            networks.latent_discriminator_forward, n_layers=c["n_latent_discr_layers"])
        self.latent_regressor = Network(
            self._make_group(netspec.latent_regressor_spec(c["latent_dim"], **self._discr_args()), s + 5),
            networks.latent_regressor_forward, n_layers=nl)
        gspec = netspec.generator_spec(c["latent_dim"], res, c["n_adain_mlp_units"], c["n_adain_mlp_layers"])
        gkw = dict(output_res=res, n_mlp_layers=c["n_adain_mlp_layers"], out_act=networks.output_activation(c["gen_output_activation"]))
        self.generator = GeneratorNet(self._make_group(gspec, s + 6), networks.generator_forward, **gkw)
        self.generator_smoothed = GeneratorNet(self._make_group(gspec, s + 6), networks.generator_forward, **gkw)
        self.generator_smoothed.group.copy_from(self.generator.group)
        # VGG19 (perceptual_loss.py:19-24).  Pretrained ImageNet weights cannot be downloaded offline:
        # seeded He-normal weights stand in; load real ones with perceptual_loss.set_weights().
        self.perceptual_loss = Network(self._make_group(netspec.vgg19_spec(), s + 7, vgg_like=True),
                                       networks.vgg19_activations)
        self.perceptual_loss.group.set_frozen(self.drop_graphs)
        if type(self) is ConfigNetFirstStage:
            # the three networks of the generator step share one gradient allocation: one all-reduce per step (runtime.coalesce_grads)
            coalesce_grads([self.generator.group, self.latent_regressor.group, self.synthetic_encoder.group])
            pretrained.from_config_or_env(self)

    PRETRAINED = (("perceptual_loss", "the VGG19 perceptual-loss network (perceptual_loss.py:19-24)"),)

    def load_pretrained_weights(self, vgg19=None, **other):
        """Loads the ImageNet VGG19 of the perceptual loss from an .npz export (confignet_b200/pretrained.py)."""
        if other:
            raise TypeError("unknown pretrained networks for the first stage: %s" % sorted(other))
        if vgg19 is not None:
            pretrained.load_vgg(self.perceptual_loss.group, vgg19, "VGG19")

    # ---------------------------------------------------------------- weights / io
    def get_weights(self, return_tensors=False):
        """confignet_first_stage.py:129-140."""
        w = {}
        w["generator_weights"] = self.generator.get_weights()
        w["generator_smoothed_weights"] = self.generator_smoothed.get_weights()
        w["discriminator_weights"] = self.discriminator.get_weights()
        w["latent_regressor_weights"] = self.latent_regressor.get_weights()
        w["synthetic_encoder_weights"] = self.synthetic_encoder.get_weights()
        w["latent_discriminator_weights"] = self.latent_discriminator.get_weights()
        w["synth_discriminator_weights"] = self.synth_discriminator.get_weights()
        return w

    def set_weights(self, weights):
        self.generator.set_weights(weights["generator_weights"])
        self.generator_smoothed.set_weights(weights["generator_smoothed_weights"])
        self.discriminator.set_weights(weights["discriminator_weights"])
        self.latent_regressor.set_weights(weights["latent_regressor_weights"])
        self.synthetic_encoder.set_weights(weights["synthetic_encoder_weights"])
        self.latent_discriminator.set_weights(weights["latent_discriminator_weights"])
        self.synth_discriminator.set_weights(weights["synth_discriminator_weights"])

    def get_training_step_number(self):
        return 0 if "loss_sum" not in self.g_losses else len(self.g_losses["loss_sum"]) - 1

    def get_batch_size(self):
        return self.config["batch_size"]

    def get_log_dict(self):
        return {"g_losses": self.g_losses, "d_losses": self.d_losses, "metrics": self.metrics}

    def set_logs(self, log_dict):
        self.g_losses, self.d_losses, self.metrics = log_dict["g_losses"], log_dict["d_losses"], log_dict["metrics"]

    def save(self, output_dir, output_filename):
        """confignet_first_stage.py:173-180: <name>.npz (lists of arrays), <name>.json, _facemodel_distr.pck."""
        weights = {}
        for k, v in self.get_weights().items():
            weights[k] = np.empty(len(v), dtype=object)
            weights[k][:] = v
        np.savez(os.path.join(output_dir, output_filename + ".npz"), **weights)
        with open(os.path.join(output_dir, output_filename + ".json"), "w") as fp:
            json.dump(self.config, fp, indent=4)
        with open(os.path.join(output_dir, output_filename + "_facemodel_distr.pck"), "wb") as fp:
            pickle.dump(self.facemodel_param_distributions, fp)

    @classmethod
    def load(cls, file_path, **kw):
        with open(file_path, "r") as fp:
            config = json.load(fp)
        model = cls(config, **kw)
        weights = np.load(os.path.splitext(file_path)[0] + ".npz", allow_pickle=True)
        model.set_weights(weights)
        log_file = os.path.splitext(file_path)[0] + "_log.json"
        if os.path.exists(log_file):
            with open(log_file, "r") as fp:
                model.set_logs(json.load(fp))
        distr = os.path.splitext(file_path)[0] + "_facemodel_distr.pck"
        if os.path.exists(distr):
            with open(distr, "rb") as fp:
                model.facemodel_param_distributions = pickle.load(fp)
        else:
            print("WARNING: facemodel param distributions not loaded")
        return model

    # ---------------------------------------------------------------- latent layout (integer, bit-exact)
    @property
    def facemodel_input_dim(self):
        return sum(d for d, _ in self.config["facemodel_inputs"].values())

    def get_facemodel_param_idxs_in_latent(self, param_name):
        """confignet_first_stage.py:217-227."""
        dims = list(self.config["facemodel_inputs"].values())
        names = list(self.config["facemodel_inputs"].keys())
        idx = names.index(param_name)
        start = int(np.sum([x[1] for x in dims[:idx]]))
        return range(start, start + dims[idx][1])

    def set_facemodel_param_in_latents(self, latents, param_name, param_value):
        """confignet_first_stage.py:229-239."""
        param_value = np.array(param_value)
        if len(param_value.shape) == 1:
            param_value = param_value[np.newaxis]
        latents_for_param = self.synthetic_encoder.per_facemodel_input_mlps[param_name].predict(param_value)
        idxs = self.get_facemodel_param_idxs_in_latent(param_name)
        new_latents = np.copy(latents)
        new_latents[:, idxs] = latents_for_param
        return new_latents

    # ---------------------------------------------------------------- samplers (host NumPy, reference RNG order)
    def update_smoothed_weights(self, smoother_alpha=0.999):
        """confignet_first_stage.py:393-400, on the device (one kernel over the flat buffer)."""
        ops.ema_update(self.generator_smoothed.group.flat, self.generator.group.flat, smoother_alpha)

    def sample_rotations(self, n_samples, axes=[0, 1, 2]):
        r = np.zeros((n_samples, 3))
        for axis in axes:
            lo, hi = self.config["rotation_ranges"][axis]
            r[:, axis] = np.pi * np.random.uniform(lo, hi, n_samples) / 180
        return r.astype(np.float32)

    def sample_latent_vector(self, n_samples):
        if self.config["latent_distribution"] == "normal":
            return np.random.normal(0, 1, (n_samples, self.config["latent_dim"]))
        elif self.config["latent_distribution"] == "uniform":
            return np.random.uniform(-1, 1, (n_samples, self.config["latent_dim"]))

    def sample_facemodel_params(self, n_samples):
        return [self.facemodel_param_distributions[n].sample(n_samples)[0] for n in self.config["facemodel_inputs"].keys()]

    def sample_synthetic_dataset(self, dataset, n_samples):
        """confignet_first_stage.py:425-435.  gt_imgs / eye_masks are returned as uint8 row gathers (host
        arrays, or device tensors when the dataset lives in HBM); /127.5-1 happens on the device."""
        idxs = np.random.randint(0, dataset.imgs.shape[0], n_samples)
        facemodel_params = [dataset.metadata_inputs[n][idxs] for n in self.config["facemodel_inputs"].keys()]
        render_rotations = dataset.metadata_inputs["rotations"][idxs].astype(np.float32)
        gt_imgs = self._take_rows(dataset.imgs, idxs)
        eye_masks = self._take_rows(dataset.eye_masks, idxs)
        return facemodel_params, render_rotations, gt_imgs, eye_masks

    # ---------------------------------------------------------------- host -> device staging (runtime.StepGraphs._to_device)
    def _take_rows(self, store, idxs):
        if isinstance(store, torch.Tensor):
            # device-resident store: the row indices go up like every other input (pinned, side stream, no host sync)
            return store.index_select(0, self._to_device(np.asarray(idxs, np.int64), torch.int64, count=False))
        return np.copy(store[idxs])

    def _upload_images(self, imgs_u8, flip=None):
        """uint8 (B,H,W,3) batch -> float32 [-1,1] device tensor; optional per-image left-right flip."""
        t = self._to_device(imgs_u8, torch.uint8)
        if t.dtype != torch.uint8:
            raise ValueError("image batches are staged as uint8")
        if flip is not None:
            f = torch.as_tensor(np.asarray(flip).astype(bool), device=self.device)
            t = torch.where(f[:, None, None, None], t.flip(2), t)
        return ops.from_uint8(t)

    def _rank_rows(self, *arrays):
        """Data parallelism: every rank draws the same global batch from the same NumPy stream and keeps
        its own row slice."""
        lo, hi = shard_rows(arrays[0].shape[0]) if world()[1] > 1 else (0, arrays[0].shape[0])
        return [a[lo:hi] for a in arrays]

    def get_discriminator_batch(self, training_set):
        """confignet_first_stage.py:438-450."""
        B = self.get_batch_size()
        img_idxs = np.random.randint(0, training_set.imgs.shape[0], B)
        flips = np.random.randint(0, 2, size=B)          # flip_random_subset_of_images' draw (confignet_utils.py:199)
        latent = self.sample_latent_vector(B).astype(np.float32)
        rotation = self.sample_rotations(B)
        img_idxs, flips, latent, rotation = self._rank_rows(img_idxs, flips, latent, rotation)
        real_imgs = self._upload_images(self._take_rows(training_set.imgs, img_idxs), flips)
        fake_imgs = self.generator.predict_device([self._to_device(latent, torch.float32), rotation])
        return real_imgs, fake_imgs

    def get_synth_discriminator_batch(self, training_set):
        """confignet_first_stage.py:452-464."""
        B = self.get_batch_size()
        img_idxs = np.random.randint(0, training_set.imgs.shape[0], B)
        flips = np.random.randint(0, 2, size=B)
        facemodel_params, rotations = self._sample_synth_metadata(training_set, B)
        sliced = self._rank_rows(img_idxs, flips, rotations, *facemodel_params)
        img_idxs, flips, rotations, facemodel_params = sliced[0], sliced[1], sliced[2], sliced[3:]
        real_imgs = self._upload_images(self._take_rows(training_set.imgs, img_idxs), flips)
        with torch.no_grad():
            latent = self.synthetic_encoder([self._to_device(a, torch.float32) for a in facemodel_params])
            fake_imgs = self.generator([latent, rotations])
        return real_imgs, fake_imgs

    def _sample_synth_metadata(self, dataset, n_samples):
        """sample_synthetic_dataset when only parameters and rotations are used (same single RNG draw);
        avoids gathering images that the caller discards (confignet_first_stage.py:460,494)."""
        idxs = np.random.randint(0, dataset.imgs.shape[0], n_samples)
        facemodel_params = [dataset.metadata_inputs[n][idxs] for n in self.config["facemodel_inputs"].keys()]
        return facemodel_params, dataset.metadata_inputs["rotations"][idxs].astype(np.float32)

    # ---------------------------------------------------------------- training steps
    def _real_from_u8(self, imgs_u8, flips):
        """device half of _upload_images: optional per-image left-right flip, uint8 -> float32 [-1, 1]"""
        if flips is not None:
            imgs_u8 = torch.where(flips[:, None, None, None], imgs_u8.flip(2), imgs_u8)
        return ops.from_uint8(imgs_u8)

    @staticmethod
    def _detached(losses):
        """The loss dictionary handed back to the caller holds plain values (the tape is released)."""
        return OrderedDict((k, v.detach()) for k, v in losses.items())

    # The three big steps are split into a host half (NumPy sampling in the reference's draw order, uploads, the
    # optimizer's iteration count / bias-corrected learning rate) and a device half on device tensors only (forward,
    # losses, backward, gradient packing, all-reduce, Adam) that runs as a CUDA-graph replay after two eager steps
    # (runtime.GraphedFn): ~2000 launches per iteration otherwise leave the GPU idle ~9 % of the time.
    def discriminator_training_step(self, training_set, optimizer):
        """confignet_first_stage.py:466-476 (batch assembly :438-450)."""
        B = self.get_batch_size()
        img_idxs = np.random.randint(0, training_set.imgs.shape[0], B)
        flips = np.random.randint(0, 2, size=B)          # flip_random_subset_of_images' draw (confignet_utils.py:199)
        latent = self.sample_latent_vector(B).astype(np.float32)
        rotation = self.sample_rotations(B)
        img_idxs, flips, latent, rotation = self._rank_rows(img_idxs, flips, latent, rotation)
        real_u8 = self._to_device(self._take_rows(training_set.imgs, img_idxs), torch.uint8)
        flips_d = self._to_device(np.asarray(flips).astype(np.bool_), torch.bool)
        latent_d = self._to_device(latent, torch.float32)
        rot_d = self._to_device(np.asarray(rotation, np.float32), torch.float32)
        optimizer.begin_step(self.device)

        def device_half(real_u8, flips_d, latent_d, rot_d):
            real_imgs = self._real_from_u8(real_u8, flips_d)
            with torch.no_grad():
                fake_imgs = self.generator((latent_d, rot_d))
            losses = networks.compute_discriminator_loss(self.discriminator.params, real_imgs, fake_imgs,
                                                         self.config["n_discr_layers"])
            self._backward(losses["loss_sum"], [self.discriminator])
            return self._detached(losses)
        fn = self._graphed("d", optimizer, device_half, [self.discriminator])
        return self._global_losses(fn(real_u8, flips_d, latent_d, rot_d))

    def synth_discriminator_training_step(self, synth_training_set, optimizer):
        """confignet_first_stage.py:478-488 (batch assembly :452-464)."""
        B = self.get_batch_size()
        img_idxs = np.random.randint(0, synth_training_set.imgs.shape[0], B)
        flips = np.random.randint(0, 2, size=B)
        facemodel_params, rotations = self._sample_synth_metadata(synth_training_set, B)
        sliced = self._rank_rows(img_idxs, flips, rotations, *facemodel_params)
        img_idxs, flips, rotations, facemodel_params = sliced[0], sliced[1], sliced[2], sliced[3:]
        real_u8 = self._to_device(self._take_rows(synth_training_set.imgs, img_idxs), torch.uint8)
        flips_d = self._to_device(np.asarray(flips).astype(np.bool_), torch.bool)
        rot_d = self._to_device(np.asarray(rotations, np.float32), torch.float32)
        fm_d = [self._to_device(a, torch.float32) for a in facemodel_params]
        optimizer.begin_step(self.device)

        def device_half(real_u8, flips_d, rot_d, *fm_d):
            real_imgs = self._real_from_u8(real_u8, flips_d)
            with torch.no_grad():
                fake_imgs = self.generator((self.synthetic_encoder(list(fm_d)), rot_d))
            losses = networks.compute_discriminator_loss(self.synth_discriminator.params, real_imgs, fake_imgs,
                                                         self.config["n_discr_layers"])
            self._backward(losses["loss_sum"], [self.synth_discriminator])
            return self._detached(losses)
        fn = self._graphed("synth_d", optimizer, device_half, [self.synth_discriminator])
        return self._global_losses(fn(real_u8, flips_d, rot_d, *fm_d))

    def latent_discriminator_training_step(self, synth_training_set, optimizer):
        """confignet_first_stage.py:490-504."""
        B = self.get_batch_size()
        real_latents = self.sample_latent_vector(B).astype(np.float32)
        facemodel_params, _ = self._sample_synth_metadata(synth_training_set, B)
        sliced = self._rank_rows(real_latents, *facemodel_params)
        real_latents, facemodel_params = sliced[0], sliced[1:]
        real_d = self._to_device(real_latents, torch.float32)
        fm_d = [self._to_device(a, torch.float32) for a in facemodel_params]
        optimizer.begin_step(self.device)

        def device_half(real_d, *fm_d):
            with torch.no_grad():
                fake_latents = self.synthetic_encoder(list(fm_d))
            losses = networks.compute_latent_discriminator_loss(self.latent_discriminator.params, real_d, fake_latents,
                                                                self.config["n_latent_discr_layers"])
            self._backward(losses["loss_sum"], [self.latent_discriminator])
            return self._detached(losses)
        fn = self._graphed("latent_d", optimizer, device_half, [self.latent_discriminator])
        return self._global_losses(fn(real_d, *fm_d))

    def generator_training_step(self, real_training_set, synth_training_set, optimizer):
        """confignet_first_stage.py:506-560."""
        pretrained.warn_if_standin(self, self.PRETRAINED[:1], "generator_training_step")
        c = self.config
        n_synth = self.get_batch_size() // 2
        n_real = self.get_batch_size() - n_synth
        idxs = np.random.randint(0, synth_training_set.imgs.shape[0], n_synth)     # sample_synthetic_dataset's draw
        facemodel_params = [synth_training_set.metadata_inputs[n][idxs] for n in c["facemodel_inputs"].keys()]
        synth_rot = synth_training_set.metadata_inputs["rotations"][idxs].astype(np.float32)
        real_latents = self.sample_latent_vector(n_real).astype(np.float32)
        real_rot = self.sample_rotations(n_real)
        sliced = self._rank_rows(idxs, synth_rot, *facemodel_params)
        idxs, synth_rot, facemodel_params = sliced[0], sliced[1], sliced[2:]
        real_latents, real_rot = self._rank_rows(real_latents, real_rot)
        gt_u8 = self._to_device(self._take_rows(synth_training_set.imgs, idxs), torch.uint8)
        masks_f = self._to_device(self._take_rows(synth_training_set.eye_masks, idxs), torch.float32)
        real_latents_d = self._to_device(real_latents, torch.float32)
        synth_rot_d = self._to_device(np.asarray(synth_rot, np.float32), torch.float32)
        real_rot_d = self._to_device(np.asarray(real_rot, np.float32), torch.float32)
        fm_d = [self._to_device(a, torch.float32) for a in facemodel_params]
        optimizer.begin_step(self.device)
        fn = self._graphed("g", optimizer, self._generator_step_device,
                           [self.generator, self.latent_regressor, self.synthetic_encoder])
        return self._global_losses(fn(gt_u8, masks_f, real_latents_d, synth_rot_d, real_rot_d, *fm_d))

    def _generator_step_device(self, gt_u8, eye_masks, real_latents_d, synth_rot, real_rot, *fm_d):
        """device half of generator_training_step (confignet_first_stage.py:514-557)"""
        c = self.config
        gt_imgs = self._real_from_u8(gt_u8, None)
        losses = OrderedDict()
        synth_latents = self.synthetic_encoder(list(fm_d))
        out_synth = self.generator((synth_latents, synth_rot))
        out_real = self.generator((real_latents_d, real_rot))
        losses["image_loss"] = c["image_loss_weight"] * networks.perceptual_loss(self.perceptual_loss.params, gt_imgs, out_synth)
        losses["eye_loss"] = c["eye_loss_weight"] * networks.eye_loss(gt_imgs, out_synth, eye_masks)
        for i, o in enumerate(self.synth_discriminator(out_synth).values()):
            losses["GAN_loss_synth_" + str(i)] = networks.gan_g_loss(o)
        for i, o in enumerate(self.discriminator(out_real).values()):
            losses["GAN_loss_real_" + str(i)] = networks.gan_g_loss(o)
        losses["latent_GAN_loss"] = c["domain_adverserial_loss_weight"] * networks.gan_g_loss(self.latent_discriminator(synth_latents))
        stacked_latents = torch.cat((synth_latents, real_latents_d), dim=0)
        stacked_imgs = torch.cat((out_synth, out_real), dim=0)
        stacked_rot = torch.cat((synth_rot, real_rot), dim=0)
        labels = torch.cat((stacked_latents, c["latent_regressor_rot_weight"] * stacked_rot), dim=-1)
        losses["latent_regression_loss"] = c["latent_regression_weight"] * networks.latent_regression_loss(
            self.latent_regressor.params, stacked_imgs, labels, c["n_discr_layers"])
        losses["loss_sum"] = networks._sum(losses.values())
        self._backward(losses["loss_sum"], [self.generator, self.latent_regressor, self.synthetic_encoder])
        return self._detached(losses)

    def setup_training(self, log_dir, synth_training_set, n_samples_for_metrics, real_training_set=None):
        """confignet_first_stage.py:562-595.  Draws from NumPy's global stream what the reference's set-up draws, in its
        order - the InceptionMetrics sample rows (metrics/metrics.py:206; always 1000), the metric latents / rotations,
        the checkpoint latents and the checkpoint rows of the synthetic set - so a seeded run enters its first training
        step at the reference's stream position, and keeps the same checkpoint / metric inputs.  A real training set that
        carries precomputed ``inception_features`` (NeuralRendererDataset does) gets the reference's InceptionMetrics object
        (KID / FID on the B200 InceptionV3, confignet_b200/metrics); one without them only consumes the same draw.
        TensorBoard is out of scope."""
        if real_training_set is None:
            real_training_set = synth_training_set
        if log_dir:
            os.makedirs(log_dir, exist_ok=True)
        if getattr(real_training_set, "inception_features", None) is not None:
            from .metrics.metrics import InceptionMetrics
            self._inception_metric_object = InceptionMetrics(self.config, real_training_set, weights=self.config.get("inception_weights"),
                                                             device=self.device)
        else:
            self._inception_metric_object = None
            self._metric_sample_idxs = np.random.randint(0, real_training_set.imgs.shape[0], 1000)
        self._generator_input_for_metrics = {"latent": self.sample_latent_vector(n_samples_for_metrics),
                                             "rotation": self.sample_rotations(n_samples_for_metrics)}
        n_rot, n_smp = self.n_checkpoint_rotations, self.n_checkpoint_samples
        latent = np.vstack([self.sample_latent_vector(n_smp)] * n_rot)          # samples 0..n-1, repeated per rotation
        lo, hi = self.config["rotation_ranges"][0]
        rotation = np.zeros((n_rot, 3))
        rotation[:, 0] = np.pi * np.linspace(lo, hi, n_rot) / 180
        rotation = np.reshape(np.hstack([rotation] * n_smp), (-1, 3))           # each yaw n_samples times in a row
        self._checkpoint_visualization_input = {"latent": latent, "rotation": rotation}
        self.facemodel_param_distributions = getattr(synth_training_set, "metadata_input_distributions", None)
        facemodel_params, _, gt_imgs, _ = self.sample_synthetic_dataset(synth_training_set, n_smp)
        self._checkpoint_visualization_input["facemodel_params"] = [np.tile(p, (n_rot, 1)) for p in facemodel_params]
        self._checkpoint_visualization_input["gt_imgs"] = gt_imgs               # uint8 rows (the reference keeps float32 copies)

    def run_checkpoints(self, output_dir, iteration_time, aml_run=None, checkpoint_start=None):
        """confignet_first_stage.py:332-375: the cadence and the files a resumed run needs - the loss histories as
        <prefix>losses.txt (confignet_utils.py:239-241) every image_checkpoint_period steps, a checkpoint under
        <output_dir>/checkpoints/<step, 6 digits> every metrics_checkpoint_period steps (step 0 included), preceded by
        calculate_metrics() when setup_training() built the metric objects.  Image grids (OpenCV), loss plots (matplotlib)
        and TensorBoard / AzureML scalars are out of scope."""
        if not output_dir or world()[0] != 0:
            return
        step_number = self.get_training_step_number()
        image_period = step_number % self.config["image_checkpoint_period"] == 0
        if image_period:
            _log_loss_vals(self.synth_d_losses, output_dir, "synth_discriminator_")
            _log_loss_vals(self.latent_d_losses, output_dir, "latent_discriminator_")
        if step_number % self.config["metrics_checkpoint_period"] == 0:
            if getattr(self, "_inception_metric_object", None) is not None:
                print("Running metrics")
                self.calculate_metrics(output_dir, aml_run=aml_run)
            checkpoint_output_dir = os.path.join(output_dir, "checkpoints")
            os.makedirs(checkpoint_output_dir, exist_ok=True)
            self.save(checkpoint_output_dir, str(step_number).zfill(6))
        if image_period:
            _log_loss_vals(self.g_losses, output_dir, "generator_")
            _log_loss_vals(self.d_losses, output_dir, "discriminator_")
            print("Training iteration time: %f" % iteration_time)

    def generate_output_for_metrics(self):
        """confignet_first_stage.py:331-332"""
        return self.generate_images(self._generator_input_for_metrics["latent"], self._generator_input_for_metrics["rotation"])

    def calculate_metrics(self, output_dir, aml_run=None):
        """confignet_first_stage.py:378-385: KID / FID of the smoothed generator's images against the real set's features,
        appended to self.metrics (saved with the checkpoint) and written to <output_dir>/inception_metrics.txt"""
        generated_images = self.generate_output_for_metrics()
        self.metrics.setdefault("training_step_number", []).append(self.get_training_step_number())
        self._inception_metric_object.update_and_log_metrics(generated_images, self.metrics, output_dir, aml_run, None)

    def train(self, real_training_set, synth_training_set, output_dir, log_dir, n_steps=100000,
              n_samples_for_metrics=1000, aml_run=None):
        """confignet_first_stage.py:597-626: loop structure, optimizer sharing, loss history, checkpoint cadence."""
        self.setup_training(log_dir, synth_training_set, n_samples_for_metrics, real_training_set=real_training_set)
        start_step = self.get_training_step_number()
        discriminator_optimizer = KerasAdam(**self.config["optimizer"])
        generator_optimizer = KerasAdam(**self.config["optimizer"])
        for _ in range(start_step, n_steps):
            t0 = time.perf_counter()
            for _ in range(self.config["n_discriminator_updates"]):
                d_loss = self.discriminator_training_step(real_training_set, discriminator_optimizer)
                synth_d_loss = self.synth_discriminator_training_step(synth_training_set, discriminator_optimizer)
                latent_d_loss = self.latent_discriminator_training_step(synth_training_set, discriminator_optimizer)
            for _ in range(self.config["n_generator_updates"]):
                g_loss = self.generator_training_step(real_training_set, synth_training_set, generator_optimizer)
            self.update_smoothed_weights()
            print("[D loss: %f] [synth_D loss: %f] [latent_D_loss: %f] [G loss: %f]" %
                  (d_loss["loss_sum"], synth_d_loss["loss_sum"], latent_d_loss["loss_sum"], g_loss["loss_sum"]))
            update_loss_dict(self.g_losses, g_loss)
            update_loss_dict(self.d_losses, d_loss)
            update_loss_dict(self.synth_d_losses, synth_d_loss)
            update_loss_dict(self.latent_d_losses, latent_d_loss)
            self.last_iteration_time = time.perf_counter() - t0
            self.run_checkpoints(output_dir, self.last_iteration_time, aml_run=aml_run)

    # ---------------------------------------------------------------- evaluation
    def _to_host(self, t):
        """device tensor -> NumPy array the caller owns, through a pinned staging buffer kept per shape (a pageable
        .cpu() costs a staging allocation and a blocking copy per call: 10 % of a batch-1 generate_images)"""
        if not t.is_cuda:
            return t.numpy()
        cache = self.__dict__.setdefault("_host_staging", {})
        key = (tuple(t.shape), t.dtype)
        buf = cache.get(key)
        if buf is None:
            if len(cache) >= 8:
                cache.clear()
            buf = cache[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        buf.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return buf.numpy().copy()

    def _generate_u8(self, net, latent_vector, rotations, borrow=False):
        """generator forward + clip + truncating uint8 cast on the device (confignet_first_stage.py:633-639).
        latent_vector: (B, latent) or a list of 5 of them (hologan_generator.py:109-127); -> uint8 device tensor.
        Batches up to 8 replay a captured graph (runtime.InferenceGraphs)."""
        dev = self.device
        if isinstance(latent_vector, (list, tuple)):
            zs = [networks._as_dev(z, dev) for z in latent_vector]
        else:
            zs = [networks._as_dev(latent_vector, dev)] * 5
        rot = networks._as_dev(rotations, dev).reshape(-1, 3)
        n = rot.shape[0]
        if n > GENERATE_CHUNK:          # the metric passes generate 1000 images per call (keras predict walks them in batches too)
            parts = [InferenceGraphs._eager(net, [z[i:i + GENERATE_CHUNK] for z in zs], rot[i:i + GENERATE_CHUNK])
                     for i in range(0, n, GENERATE_CHUNK)]
            return torch.cat(parts, dim=0)
        if self.config.get("cuda_graphs", True):
            return self._infer.run(net, zs, rot, borrow=borrow)
        return InferenceGraphs._eager(net, zs, rot)

    def generate_images(self, latent_vector, rotations):
        """confignet_first_stage.py:633-639 -> uint8 (B,H,W,3)."""
        return self._to_host(self._generate_u8(self.generator_smoothed, latent_vector, rotations, borrow=True))

    def generate_images_device(self, latent_vector, rotations):
        """generate_images with device tensors in and a uint8 device tensor out (no host round trip)"""
        return self._generate_u8(self.generator_smoothed, latent_vector, rotations)

    def generate_images_from_facemodel(self, facemodel_params, rotations):
        with torch.no_grad():
            latents = self.synthetic_encoder(facemodel_params)
        return self.generate_images(latents, rotations)
