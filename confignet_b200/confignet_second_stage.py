"""ConfigNet (second stage) on B200: the class surface of the reference
(confignet/confignet_second_stage.py:19-403) over our CUDA networks.

Adds to ConfigNetFirstStage: the RealEncoder (keras-applications ResNet50 + two Dense heads,
dnn_models/real_encoder.py:9-34), the second-stage latent-discriminator / generator steps (:132-218) with the
batch-normalised latent regression loss (:93-107), ``encode_images`` (:301-308), ``generate_images`` with the
fine-tuned generator (:310-319) and ``fine_tune_on_img`` (:321-403) with its VGGFace (VGG16) perceptual term.

Out of scope: image checkpoints (OpenCV grids), TensorBoard.
"""
import os
import time
from collections import OrderedDict
import numpy as np
import torch
import torch.distributed as dist

from . import netspec, networks, ops, pretrained
from .confignet_first_stage import (ConfigNetFirstStage, DEFAULT_CONFIG, merge_configs, update_loss_dict, GeneratorNet)
from .runtime import ParamGroup, Network, KerasAdam, allreduce_grads, shard_rows, world, gather_rows, coalesce_grads

PREDICT_BATCH = 32       # keras Model.predict default batch size [TF-2.1]


class RealEncoderNet(Network):
    """RealEncoder (real_encoder.py:9-34): ``enc(imgs) -> (embedding, rotation)``; imgs float32 in [-1, 1]."""

    def __init__(self, group, rotation_ranges):
        self.rotation_range_multiplier = np.pi * np.array(
            [rotation_ranges[0][1], rotation_ranges[1][1], rotation_ranges[2][1]], np.float64) / 180.0
        mult = torch.tensor(self.rotation_range_multiplier, dtype=torch.float32, device=group.device)
        latent_dim = group.params["feature_to_latent_mlp/bias"].shape[0]
        super().__init__(group, networks.real_encoder_forward, weights_order=netspec.real_encoder_keras_order(latent_dim),
                         rotation_range_multiplier=mult)

    def __call__(self, imgs):
        return super().__call__(networks._as_dev(imgs, self.group.device))

    def predict(self, imgs):
        """keras Model.predict -> (embeddings, rotations) NumPy arrays"""
        e, r = self.predict_device(imgs)
        return e.cpu().numpy(), r.cpu().numpy()

    def predict_device(self, imgs):
        """keras Model.predict's batching: forward in batches of 32 without a tape -> (embeddings, rotations) device tensors."""
        embs, rots = [], []
        with torch.no_grad():
            for i in range(0, imgs.shape[0], PREDICT_BATCH):
                e, r = self(imgs[i:i + PREDICT_BATCH])
                embs.append(e); rots.append(r)
        if not embs:
            dev = self.group.device
            return torch.zeros((0, self.group.params["feature_to_latent_mlp/bias"].shape[0]), device=dev), torch.zeros((0, 3), device=dev)
        return torch.cat(embs, dim=0), torch.cat(rots, dim=0)


class ConfigNet(ConfigNetFirstStage):
    def __init__(self, config, initialize=True, device=None, seed=1234):
        self.config = merge_configs(DEFAULT_CONFIG, config)
        super().__init__(self.config, initialize=False, device=device, seed=seed)
        self.config["model_type"] = "ConfigNet"
        self.encoder = None
        self.generator_fine_tuned = None
        self.controllability_metrics = None
        self.perceptual_loss_face_reco = None
        if initialize:
            self.initialize_network()

    # ---------------------------------------------------------------- construction / weights
    def initialize_network(self):
        """confignet_second_stage.py:45-49 (+ :29, the VGGFace perceptual model)."""
        super().initialize_network()
        c, s = self.config, self._seed
        enc = ParamGroup(netspec.init_real_encoder_params(c["latent_dim"], s + 8), self.device, trainable=netspec.is_trainable)
        self.encoder = RealEncoderNet(enc, c["rotation_ranges"])
        # VGG16 with the VGGFace weights (perceptual_loss.py:26-41): not downloadable offline, seeded stand-in
        self.perceptual_loss_face_reco = Network(self._make_group(netspec.vgg16_spec(), s + 9, vgg_like=True),
                                                 networks.vggface_activations)
        self.perceptual_loss_face_reco.group.set_frozen(self.drop_graphs)
        # the four networks of the stage-2 generator step, in _backward's order: one all-reduce per step
        coalesce_grads([self.generator.group, self.latent_regressor.group, self.synthetic_encoder.group, self.encoder.group])
        pretrained.from_config_or_env(self)

    PRETRAINED = (("perceptual_loss", "the VGG19 perceptual-loss network (perceptual_loss.py:19-24)"),
                  ("encoder", "the ResNet50 trunk of the real encoder (real_encoder.py:13)"),
                  ("perceptual_loss_face_reco", "the VGGFace VGG16 network (perceptual_loss.py:26-41)"))

    def load_pretrained_weights(self, vgg19=None, vggface=None, resnet50=None):
        """Loads the third-party networks from .npz exports (confignet_b200/pretrained.py): ImageNet VGG19 (perceptual
        loss), VGGFace VGG16 (fine-tuning face-recognition loss), ImageNet ResNet50 (the encoder's trunk; a checkpoint
        loaded with load() / set_weights() already carries a TRAINED encoder and marks it as such)."""
        if vgg19 is not None:
            pretrained.load_vgg(self.perceptual_loss.group, vgg19, "VGG19")
        if vggface is not None:
            pretrained.load_vgg(self.perceptual_loss_face_reco.group, vggface, "VGGFace VGG16")
        if resnet50 is not None:
            pretrained.load_resnet50(self.encoder.group, resnet50)

    def face_reco_loss(self, gt_imgs, gen_imgs):
        """confignet_second_stage.py:88-91: the VGGFace perceptual loss with (generated, ground truth) in the reference's
        argument order; device tensors (or arrays) in [-1, 1]."""
        dev = self.device
        return networks.perceptual_loss(self.perceptual_loss_face_reco.params, networks._as_dev(gen_imgs, dev),
                                        networks._as_dev(gt_imgs, dev), model_type="VGGFace")

    def get_weights(self, return_tensors=False):
        w = super().get_weights()
        w["real_encoder_weights"] = self.encoder.get_weights()
        return w

    def set_weights(self, weights):
        super().set_weights(weights)
        self.encoder.set_weights(weights["real_encoder_weights"])
        self.encoder.group.pretrained = True          # a checkpoint's encoder: trained, not the seeded stand-in

    # ---------------------------------------------------------------- batch assembly (reference RNG draw order)
    def _draw_image_rows(self, dataset, batch_size):
        """sample_random_batch_of_images (confignet_second_stage.py:109-117): index draw, then the flip draw."""
        idxs = np.random.randint(0, dataset.imgs.shape[0], batch_size)
        flips = np.random.randint(0, 2, size=batch_size)
        return idxs, flips

    def sample_random_batch_of_images(self, dataset, batch_size=None):
        if batch_size is None:
            batch_size = self.get_batch_size()
        idxs, flips = self._draw_image_rows(dataset, batch_size)
        idxs, flips = self._rank_rows(idxs, flips)
        return self._upload_images(self._take_rows(dataset.imgs, idxs), flips)

    def get_discriminator_batch(self, training_set):
        """confignet_second_stage.py:119-130: fakes are reconstructions of encoded training images."""
        B = self.get_batch_size()
        idxs, flips = self._draw_image_rows(training_set, B)
        input_img_idxs = np.random.randint(0, training_set.imgs.shape[0], B)
        idxs, flips, input_img_idxs = self._rank_rows(idxs, flips, input_img_idxs)
        real_imgs = self._upload_images(self._take_rows(training_set.imgs, idxs), flips)
        input_imgs = self._upload_images(self._take_rows(training_set.imgs, input_img_idxs))
        with torch.no_grad():
            latent_vector, rotation = self.encoder(input_imgs)
            fake_imgs = self.generator([latent_vector, rotation])
        return real_imgs, fake_imgs

    # ---------------------------------------------------------------- training steps
    # Same split as in stage 1: a host half (NumPy draws in the reference's order, pinned uploads, Keras-Adam's
    # iteration count) and a device half on device tensors only, replayed as a CUDA graph after two eager calls.
    def discriminator_training_step(self, training_set, optimizer):
        """confignet_first_stage.py:466-476 over ConfigNet.get_discriminator_batch (confignet_second_stage.py:119-130):
        the fakes are reconstructions generator(encoder(training images)) - not prior samples as in stage 1 - and the
        NumPy stream sees (image rows, flips, input image rows)."""
        B = self.get_batch_size()
        idxs, flips = self._draw_image_rows(training_set, B)
        input_img_idxs = np.random.randint(0, training_set.imgs.shape[0], B)
        idxs, flips, input_img_idxs = self._rank_rows(idxs, flips, input_img_idxs)
        real_u8 = self._to_device(self._take_rows(training_set.imgs, idxs), torch.uint8)
        flips_d = self._to_device(np.asarray(flips).astype(np.bool_), torch.bool)
        input_u8 = self._to_device(self._take_rows(training_set.imgs, input_img_idxs), torch.uint8)
        optimizer.begin_step(self.device)

        def device_half(real_u8, flips_d, input_u8):
            real_imgs = self._real_from_u8(real_u8, flips_d)
            with torch.no_grad():
                latent_vector, rotation = self.encoder(self._real_from_u8(input_u8, None))
                fake_imgs = self.generator((latent_vector, rotation))
            losses = networks.compute_discriminator_loss(self.discriminator.params, real_imgs, fake_imgs,
                                                         self.config["n_discr_layers"])
            self._backward(losses["loss_sum"], [self.discriminator])
            return self._detached(losses)
        fn = self._graphed("d2", optimizer, device_half, [self.discriminator])
        return self._global_losses(fn(real_u8, flips_d, input_u8))

    def latent_discriminator_training_step(self, real_training_set, synth_training_set, optimizer):
        """confignet_second_stage.py:132-147: real latents come from the encoder."""
        B = self.get_batch_size()
        idxs, flips = self._draw_image_rows(real_training_set, B)
        facemodel_params, _ = self._sample_synth_metadata(synth_training_set, B)
        sliced = self._rank_rows(idxs, flips, *facemodel_params)
        idxs, flips, facemodel_params = sliced[0], sliced[1], sliced[2:]
        real_u8 = self._to_device(self._take_rows(real_training_set.imgs, idxs), torch.uint8)
        flips_d = self._to_device(np.asarray(flips).astype(np.bool_), torch.bool)
        fm_d = [self._to_device(a, torch.float32) for a in facemodel_params]
        optimizer.begin_step(self.device)

        def device_half(real_u8, flips_d, *fm_d):
            with torch.no_grad():
                real_latents, _ = self.encoder(self._real_from_u8(real_u8, flips_d))
                fake_latents = self.synthetic_encoder(list(fm_d))
            losses = networks.compute_latent_discriminator_loss(self.latent_discriminator.params, real_latents, fake_latents,
                                                                self.config["n_latent_discr_layers"])
            self._backward(losses["loss_sum"], [self.latent_discriminator])
            return self._detached(losses)
        fn = self._graphed("latent_d2", optimizer, device_half, [self.latent_discriminator])
        return self._global_losses(fn(real_u8, flips_d, *fm_d))

    def compute_normalized_latent_regression_loss(self, generator_outputs, labels):
        """confignet_second_stage.py:93-107.  The statistics are over the GLOBAL batch: with data parallelism the
        regressor outputs and labels are all-gathered (B x 148 floats) before the loss kernel."""
        c = self.config
        out = networks.latent_regressor_forward(self.latent_regressor.params, generator_outputs, c["n_discr_layers"])
        return ops.norm_latent_loss(gather_rows(out), gather_rows(labels), float(c["latent_regression_weight"]), 3)

    def generator_training_step(self, real_training_set, synth_training_set, optimizer):
        """confignet_second_stage.py:149-218."""
        pretrained.warn_if_standin(self, self.PRETRAINED[:2], "generator_training_step")
        c = self.config
        n_synth = self.get_batch_size() // 2
        n_real = self.get_batch_size() - n_synth
        idxs = np.random.randint(0, synth_training_set.imgs.shape[0], n_synth)     # sample_synthetic_dataset's draw
        facemodel_params = [synth_training_set.metadata_inputs[n][idxs] for n in c["facemodel_inputs"].keys()]
        synth_rot = synth_training_set.metadata_inputs["rotations"][idxs].astype(np.float32)
        ridxs, rflips = self._draw_image_rows(real_training_set, n_real)
        sliced = self._rank_rows(idxs, synth_rot, *facemodel_params)
        idxs, synth_rot, facemodel_params = sliced[0], sliced[1], sliced[2:]
        ridxs, rflips = self._rank_rows(ridxs, rflips)
        synth_u8 = self._to_device(self._take_rows(synth_training_set.imgs, idxs), torch.uint8)
        masks_f = self._to_device(self._take_rows(synth_training_set.eye_masks, idxs), torch.float32)
        real_u8 = self._to_device(self._take_rows(real_training_set.imgs, ridxs), torch.uint8)
        rflips_d = self._to_device(np.asarray(rflips).astype(np.bool_), torch.bool)
        synth_rot_d = self._to_device(np.asarray(synth_rot, np.float32), torch.float32)
        fm_d = [self._to_device(a, torch.float32) for a in facemodel_params]
        optimizer.begin_step(self.device)
        # the batch-statistics loss all-gathers under data parallelism: a collective inside the step, so no capture there
        fn = self._graphed("g2", optimizer, self._generator_step_device,
                           [self.generator, self.latent_regressor, self.synthetic_encoder, self.encoder], dp_ok=False)
        return self._global_losses(fn(synth_u8, masks_f, real_u8, rflips_d, synth_rot_d, *fm_d))

    def _generator_step_device(self, synth_u8, eye_masks, real_u8, rflips_d, synth_rot, *fm_d):
        """device half of the stage-2 generator_training_step (confignet_second_stage.py:166-216)"""
        c = self.config
        synth_imgs = self._real_from_u8(synth_u8, None)
        real_imgs = self._real_from_u8(real_u8, rflips_d)
        losses = OrderedDict()
        synth_latents = self.synthetic_encoder(list(fm_d))
        out_synth = self.generator((synth_latents, synth_rot))
        real_latents, real_rot = self.encoder(real_imgs)
        out_real = self.generator((real_latents, real_rot))
        p_vgg = self.perceptual_loss.params
        losses["image_loss_synth"] = c["image_loss_weight"] * networks.perceptual_loss(p_vgg, synth_imgs, out_synth)
        losses["image_loss_real"] = c["image_loss_weight"] * networks.perceptual_loss(p_vgg, real_imgs, out_real)
        losses["eye_loss"] = c["eye_loss_weight"] * networks.eye_loss(synth_imgs, out_synth, eye_masks)
        for i, o in enumerate(self.synth_discriminator(out_synth).values()):
            losses["GAN_loss_synth_" + str(i)] = networks.gan_g_loss(o)
        for i, o in enumerate(self.discriminator(out_real).values()):
            losses["GAN_loss_real_" + str(i)] = networks.gan_g_loss(o)
        # domain-adversarial loss: labels 0 for the real latents, 1 for the synthetic ones (:157-161,192-199)
        ld_synth = self.latent_discriminator(synth_latents)
        ld_real = self.latent_discriminator(real_latents)
        losses["latent_GAN_loss"] = c["domain_adverserial_loss_weight"] * networks.gan_d_loss_mixed(ld_real, ld_synth)
        if c["latent_regression_weight"] > 0.0:
            stacked_latents = torch.cat((synth_latents, real_latents), dim=0)
            stacked_imgs = torch.cat((out_synth, out_real), dim=0)
            stacked_rot = torch.cat((synth_rot, real_rot), dim=0)
            labels = torch.cat((stacked_latents, c["latent_regressor_rot_weight"] * stacked_rot), dim=-1)
            losses["latent_regression_loss"] = self.compute_normalized_latent_regression_loss(stacked_imgs, labels)
        losses["loss_sum"] = networks._sum(losses.values())
        self._backward(losses["loss_sum"], [self.generator, self.latent_regressor, self.synthetic_encoder, self.encoder])
        return self._detached(losses)

    def setup_training(self, log_dir, synth_training_set, n_samples_for_metrics, attribute_classifier=None,
                       real_training_set=None, validation_set=None):
        """confignet_second_stage.py:255-266: + the checkpoint / metric rows of the validation set (two more draws from
        the NumPy stream; skipped, like the reference would fail, when there is no validation set) and the
        ControllabilityMetrics object around ``attribute_classifier`` (a CelebaAttributeClassifier or the path of its
        .json; None: no controllability metrics)."""
        super().setup_training(log_dir, synth_training_set, n_samples_for_metrics, real_training_set)
        if validation_set is not None:
            idxs = np.random.randint(0, validation_set.imgs.shape[0], self.n_checkpoint_samples)
            self._checkpoint_visualization_input["input_images"] = self._take_rows(validation_set.imgs, idxs)
            idxs = np.random.randint(0, validation_set.imgs.shape[0], n_samples_for_metrics)
            self._generator_input_for_metrics["input_images"] = self._take_rows(validation_set.imgs, idxs)
        if attribute_classifier is not None:
            from .metrics.metrics import ControllabilityMetrics
            self.controllability_metrics = ControllabilityMetrics(self, attribute_classifier)

    def generate_output_for_metrics(self):
        """confignet_second_stage.py:81-83: reconstructions of the validation rows"""
        latent, rotation = self.encode_images(self._generator_input_for_metrics["input_images"])
        return self.generate_images(latent, rotation)

    def calculate_metrics(self, output_dir, aml_run=None):
        """confignet_second_stage.py:220-253: KID / FID of the reconstructions, the controllability metrics, and the mean
        perceptual (VGG19) loss between the validation images and their reconstructions in batches of 16 -> self.metrics,
        controllability_metrics.json, image_metrics.txt"""
        super().calculate_metrics(output_dir, aml_run)
        input_images = self._generator_input_for_metrics["input_images"]
        if self.controllability_metrics is not None:
            self.controllability_metrics.update_and_log_metrics(input_images, self.metrics, output_dir, aml_run, None)
        latents, rotations = self.encode_images(input_images)
        metric_batch_size = 16
        n_valid_samples = len(input_images)
        perceptual_loss = []
        with torch.no_grad():
            for i in range(1 + n_valid_samples // metric_batch_size):
                start_idx, end_idx = i * metric_batch_size, min(n_valid_samples, (i + 1) * metric_batch_size)
                if end_idx <= start_idx:
                    break           # (the reference evaluates an empty trailing batch - NaN - when n is a multiple of 16)
                gt_imgs = self._images_to_device(input_images[start_idx:end_idx])
                gen_imgs = self.generator_smoothed.predict_device(
                    self.generator_smoothed.build_input_dict(self._to_device(latents[start_idx:end_idx], torch.float32),
                                                             self._to_device(rotations[start_idx:end_idx], torch.float32)))
                perceptual_loss.append(float(networks.perceptual_loss(self.perceptual_loss.params, gt_imgs, gen_imgs)))
        perceptual_loss = float(np.mean(perceptual_loss))
        self.metrics.setdefault("perceptual_loss", []).append(perceptual_loss)
        if aml_run is not None:
            aml_run.log("perceptual_loss", perceptual_loss)
        np.savetxt(os.path.join(output_dir, "image_metrics.txt"), self.metrics["perceptual_loss"])

    def train(self, real_training_set, synth_training_set, validation_set=None, attribute_classifier=None,
              output_dir=None, log_dir=None, n_steps=100000, n_samples_for_metrics=1000, aml_run=None):
        """confignet_second_stage.py:268-299: loop structure, optimizer sharing, loss history, checkpoint cadence."""
        self.setup_training(log_dir, synth_training_set, n_samples_for_metrics, attribute_classifier,
                            real_training_set=real_training_set, validation_set=validation_set)
        start_step = self.get_training_step_number()
        discriminator_optimizer = KerasAdam(**self.config["optimizer"])
        generator_optimizer = KerasAdam(**self.config["optimizer"])
        for _ in range(start_step, n_steps):
            t0 = time.perf_counter()
            for _ in range(self.config["n_discriminator_updates"]):
                d_loss = self.discriminator_training_step(real_training_set, discriminator_optimizer)
                synth_d_loss = self.synth_discriminator_training_step(synth_training_set, discriminator_optimizer)
                latent_d_loss = self.latent_discriminator_training_step(real_training_set, synth_training_set, discriminator_optimizer)
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
    def _images_to_device(self, input_images):
        """uint8 -> /127.5 - 1 on the device; float images are taken as already in [-1, 1]."""
        if isinstance(input_images, np.ndarray) and input_images.dtype == np.uint8:
            return self._upload_images(input_images)
        if isinstance(input_images, torch.Tensor) and input_images.dtype == torch.uint8:
            return ops.from_uint8(input_images.to(self.device))
        return self._to_device(np.asarray(input_images, np.float32) if not isinstance(input_images, torch.Tensor) else input_images,
                               torch.float32)

    def encode_images(self, input_images):
        """confignet_second_stage.py:301-308 -> (embeddings (B, latent) f32, rotations (B, 3) f32) NumPy."""
        embs, rots = [], []
        for i in range(0, input_images.shape[0], PREDICT_BATCH):
            e, r = self.encoder.predict_device(self._images_to_device(input_images[i:i + PREDICT_BATCH]))
            embs.append(e); rots.append(r)
        if not embs:
            return np.zeros((0, self.config["latent_dim"]), np.float32), np.zeros((0, 3), np.float32)
        return torch.cat(embs, dim=0).cpu().numpy(), torch.cat(rots, dim=0).cpu().numpy()

    def generate_images(self, latent_vectors, rotations):
        """confignet_second_stage.py:310-319 -> uint8 (B,H,W,3)."""
        net = self.generator_fine_tuned if self.generator_fine_tuned is not None else self.generator_smoothed
        return self._to_host(self._generate_u8(net, latent_vectors, rotations, borrow=True))

    def generate_images_device(self, latent_vectors, rotations):
        net = self.generator_fine_tuned if self.generator_fine_tuned is not None else self.generator_smoothed
        return self._generate_u8(net, latent_vectors, rotations)

    def _new_fine_tune_state(self, key, shared_arrays, local_arrays, imgs, force_neutral_expression):
        """Variables, optimizer and the device half of one fine-tuning iteration for one case of fine_tune_on_img."""
        c = self.config
        n_imgs = key[0]
        shared = ParamGroup(shared_arrays, self.device)
        local = ParamGroup(local_arrays, self.device,
                           trainable=(lambda k: k != "expr_embeddings") if force_neutral_expression else None)
        pre, post = shared.params["pre_expr_embeddings"], shared.params["post_expr_embeddings"]
        expr, rotations = local.params["expr_embeddings"], local.params["rotations"]
        optimizer = KerasAdam(lr=0.0001, beta_1=0.9, beta_2=0.999)
        gen = self.generator_fine_tuned
        # the tiled pre/post parts the reference returns are the ones built INSIDE the last tape, i.e. their values
        # before the last optimizer update, next to the updated expression part (confignet_second_stage.py:361-364,402):
        # found by executing the reference's fine_tune_on_img (tests/golden/reference_steps.npz) and kept
        pre_tiled, post_tiled = pre.detach().clone(), post.detach().clone()
        out_first = torch.empty((1,) + tuple(imgs.shape[1:]), device=self.device, dtype=torch.float32)

        def iteration(imgs):
            """one fine-tuning iteration on device tensors only (confignet_second_stage.py:359-392, without the update)"""
            losses = OrderedDict()
            with torch.no_grad():
                pre_tiled.copy_(pre); post_tiled.copy_(post)
            embeddings = torch.cat((pre.expand(n_imgs, -1), expr, post.expand(n_imgs, -1)), dim=1)
            out = gen((embeddings, rotations))
            losses["image_loss_real"] = 0.5 * c["image_loss_weight"] * networks.perceptual_loss(self.perceptual_loss.params, imgs, out)
            losses["face_reco_loss"] = 0.5 * c["image_loss_weight"] * networks.perceptual_loss(
                self.perceptual_loss_face_reco.params, out, imgs, model_type="VGGFace")
            for i, o in enumerate(self.discriminator(out).values()):
                losses["GAN_loss_real_" + str(i)] = networks.gan_g_loss(o)
            losses["latent_GAN_loss"] = c["domain_adverserial_loss_weight"] * networks.gan_d_loss(1, self.latent_discriminator(embeddings))
            labels = torch.cat((embeddings, c["latent_regressor_rot_weight"] * rotations), dim=-1)
            losses["latent_regression_loss"] = self.compute_normalized_latent_regression_loss(out, labels)
            losses["loss_sum"] = networks._sum(losses.values())
            self._backward(losses["loss_sum"], [gen.group, shared, local])
            with torch.no_grad():
                out_first.copy_(out[:1])
            return self._detached(losses)

        return dict(shared=shared, local=local, optimizer=optimizer, pre_tiled=pre_tiled, post_tiled=post_tiled,
                    out_first=out_first, iteration=iteration)

    def fine_tune_on_img(self, input_images, n_iters=50, img_output_dir=None, force_neutral_expression=False):
        """confignet_second_stage.py:321-403.  One shared fine-tuned generator and shared pre/post-expression
        embeddings, per-image expression embeddings and rotations; Keras Adam(lr=1e-4) defaults.  With data
        parallelism the images are sharded by rows: generator and pre/post gradients are all-reduced, the per-image
        variables stay rank-local."""
        pretrained.warn_if_standin(self, self.PRETRAINED, "fine_tune_on_img")
        c = self.config
        if isinstance(input_images, np.ndarray) and input_images.ndim == 3:
            input_images = input_images[np.newaxis]
        rank, ws = world()
        n_global = input_images.shape[0]
        if ws > 1:
            lo, hi = shard_rows(n_global)
            input_images = input_images[lo:hi]
        imgs = self._images_to_device(input_images)
        n_imgs = imgs.shape[0]
        pred_emb, pred_rot = self.encoder.predict(imgs)
        if force_neutral_expression:
            n_exp = c["facemodel_inputs"]["blendshape_values"][0]
            pred_emb = self.set_facemodel_param_in_latents(pred_emb, "blendshape_values", np.zeros((1, n_exp), np.float32))

        if self.generator_fine_tuned is None:
            gspec = netspec.generator_spec(c["latent_dim"], c["output_shape"][0], c["n_adain_mlp_units"], c["n_adain_mlp_layers"])
            self.generator_fine_tuned = GeneratorNet(self._make_group(gspec, self._seed + 6), networks.generator_forward,
                                                     output_res=c["output_shape"][0], n_mlp_layers=c["n_adain_mlp_layers"],
                                                     out_act=networks.output_activation(c["gen_output_activation"]))
        self.generator_fine_tuned.group.copy_from(self.generator_smoothed.group)

        expr_idxs = self.get_facemodel_param_idxs_in_latent("blendshape_values")
        mean_emb = np.sum(pred_emb, axis=0, keepdims=True, dtype=np.float32)
        if ws > 1:
            t = torch.from_numpy(mean_emb).to(self.device)
            dist.all_reduce(t)
            mean_emb = t.cpu().numpy()
        mean_emb = mean_emb / np.float32(n_global)
        # The variables, the optimizer and the captured iteration of one (image count, resolution, neutral-expression) case
        # are kept between calls: the reference builds new tf.Variables and a new Adam per call
        # (confignet_second_stage.py:341-358), here the same buffers are re-initialised in place - same values, and the
        # CUDA graph captured by an earlier call is replayed from the first iteration on (bench config 5: a capture per
        # call cost as much as the ten iterations it served).
        shared_arrays = OrderedDict([("pre_expr_embeddings", mean_emb[:, :expr_idxs[0]]),
                                     ("post_expr_embeddings", mean_emb[:, expr_idxs[-1] + 1:])])
        local_arrays = OrderedDict([("rotations", pred_rot), ("expr_embeddings", pred_emb[:, list(expr_idxs)])])
        key = (n_imgs, tuple(imgs.shape[1:]), bool(force_neutral_expression), ws)
        cache = self.__dict__.setdefault("_fine_tune_state", OrderedDict())
        ft = cache.get(key)
        if ft is None:
            while len(cache) >= 2:                                   # two cases stay resident (graph memory pools)
                old_key, _ = cache.popitem(last=False)
                entry = self._graphs.pop("fine_tune:%r" % (old_key,), None)
                if entry is not None:
                    entry[1].release()
            ft = cache[key] = self._new_fine_tune_state(key, shared_arrays, local_arrays, imgs, force_neutral_expression)
        else:
            cache.move_to_end(key)
            ft["shared"].set_weights(list(shared_arrays.values()))
            ft["local"].set_weights(list(local_arrays.values()))
            ft["optimizer"].reset()
            with torch.no_grad():
                ft["pre_tiled"].copy_(ft["shared"].params["pre_expr_embeddings"])
                ft["post_tiled"].copy_(ft["shared"].params["post_expr_embeddings"])
        shared, local, optimizer = ft["shared"], ft["local"], ft["optimizer"]
        expr, rotations = local.params["expr_embeddings"], local.params["rotations"]
        pre_tiled, post_tiled, out_first = ft["pre_tiled"], ft["post_tiled"], ft["out_first"]
        gen = self.generator_fine_tuned
        if img_output_dir is not None and rank == 0:
            os.makedirs(img_output_dir, exist_ok=True)
        self.fine_tune_losses = []
        step = self._graphed("fine_tune:%r" % (key,), optimizer, ft["iteration"], [gen.group, shared, local], dp_ok=False,
                             reduce_groups=[gen.group, shared])
        for step_number in range(n_iters):
            optimizer.begin_step(self.device)
            self.fine_tune_losses.append(step(imgs))
            if img_output_dir is not None and rank == 0:
                np.save(os.path.join(img_output_dir, "output_%02d.npy" % step_number), ops.to_uint8(out_first).cpu().numpy()[0])

        with torch.no_grad():
            embeddings = torch.cat((pre_tiled.expand(n_imgs, -1), expr, post_tiled.expand(n_imgs, -1)), dim=1)
            emb_out, rot_out = embeddings.contiguous(), rotations.detach().contiguous()
            if ws > 1:
                eparts = [torch.empty_like(emb_out) for _ in range(ws)]
                rparts = [torch.empty_like(rot_out) for _ in range(ws)]
                dist.all_gather(eparts, emb_out); dist.all_gather(rparts, rot_out)
                emb_out, rot_out = torch.cat(eparts, dim=0), torch.cat(rparts, dim=0)
        return emb_out.cpu().numpy(), rot_out.cpu().numpy()
