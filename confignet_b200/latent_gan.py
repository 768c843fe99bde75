"""LatentGAN on B200: the class surface of the reference (confignet/latent_gan.py:32-253).

An MLP GAN on (B, latent_dim) embedding vectors: generator / smoothed generator / discriminator are
MLPSimple(num_layers, latent_dim, int(1.5 * latent_dim), latent_dim | 1) with Keras LeakyReLU (alpha 0.3).  Every
Dense layer is a launch of the implicit-GEMM family, the R1 penalty on real embeddings runs through the
second-order-closed Dense / LeakyReLU operators, Adam is one fused kernel per network and the EMA stays on the
device.  Out of scope: TensorBoard logs and Inception metrics.
"""
import os
import json
from collections import OrderedDict
import numpy as np
import torch

from . import netspec, networks, ops
from .confignet_first_stage import merge_configs
from .runtime import ParamGroup, Network, KerasAdam, StepGraphs, shard_rows, world

DEFAULT_CONFIG = {
    "latent_dim": None,
    "optimizer": {"lr": 0.00005, "beta_1": 0.0, "beta_2": 0.9, "amsgrad": False},
    "batch_size": 32,
    "num_mlp_layers": 3,
    "latent_distribution_type": "normal",
    "hidden_layer_size_multiplier": 1.5,
    "n_samples_for_metrics": 1000,
    "verbose_log_period": 500,
    "logging_img_square_size": 6,
}


def _mlp_forward(p, x, num_layers, second_order=False):
    x = networks._as_dev(x, next(iter(p.values())).device)
    f = networks.mlp_diff if second_order else networks.mlp_fused
    return f(x, p, "mlp", num_layers, 0.3)


class LatentGAN(StepGraphs):
    def __init__(self, config, device=None, seed=4321):
        self.config = merge_configs(DEFAULT_CONFIG, config)
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self._seed = seed
        self._graphs = {}                 # step name -> (optimizer, runtime.GraphedFn)
        self.generator = None
        self.generator_smoothed = None
        self.discriminator = None
        self.d_losses, self.g_losses = {}, {}
        self.initialize_network()

    @classmethod
    def load(cls, file_path, **kw):
        with open(file_path, "r") as fp:
            config = json.load(fp)
        gan = cls(config, **kw)
        weights = np.load(os.path.splitext(file_path)[0] + ".npz", allow_pickle=True)
        gan.set_weights(weights)
        return gan

    def save(self, output_dir, output_filename):
        np.savez(os.path.join(output_dir, output_filename + ".npz"), **self.get_weights())
        with open(os.path.join(output_dir, output_filename + ".json"), "w") as fp:
            json.dump(self.config, fp, indent=4)

    def get_weights(self):
        """latent_gan.py:65-80: object arrays of per-variable arrays."""
        weights = {}
        for key, net in (("generator_weights", self.generator), ("smoothed_generator_weights", self.generator_smoothed),
                         ("discriminator_weights", self.discriminator)):
            w = net.get_weights()
            weights[key] = np.empty(len(w), dtype=object)
            weights[key][:] = w
        return weights

    def set_weights(self, weights):
        self.generator.set_weights(weights["generator_weights"])
        self.generator_smoothed.set_weights(weights["smoothed_generator_weights"])
        self.discriminator.set_weights(weights["discriminator_weights"])

    def initialize_network(self):
        """latent_gan.py:88-109."""
        c = self.config
        nl = c["num_mlp_layers"]
        gspec = netspec.latent_gan_mlp_spec(c["latent_dim"], nl, c["hidden_layer_size_multiplier"])
        dspec = netspec.latent_gan_mlp_spec(c["latent_dim"], nl, c["hidden_layer_size_multiplier"], num_out=1)
        mk = lambda spec, seed: ParamGroup(netspec.init_params(spec, seed), self.device)
        self.generator = Network(mk(gspec, self._seed), _mlp_forward, num_layers=nl)
        self.generator_smoothed = Network(mk(gspec, self._seed), _mlp_forward, num_layers=nl)
        self.generator_smoothed.group.copy_from(self.generator.group)
        self.discriminator = Network(mk(dspec, self._seed + 1), _mlp_forward, num_layers=nl)

    def sample_input_latent_vector(self, n_samples):
        if self.config["latent_distribution_type"] == "uniform":
            return np.random.uniform(-1, 1, (n_samples, self.config["latent_dim"]))
        elif self.config["latent_distribution_type"] == "normal":
            return np.random.normal(0, 1, (n_samples, self.config["latent_dim"]))

    # ---------------------------------------------------------------- training steps
    def _rank_rows(self, *arrays):
        lo, hi = shard_rows(arrays[0].shape[0]) if world()[1] > 1 else (0, arrays[0].shape[0])
        return [a[lo:hi] for a in arrays]

    # these steps are launch-latency bound (three-layer MLPs on (B, latent) matrices, ~120 tiny
    # launches each) - host half (NumPy draws, uploads, Adam's iteration count) + a CUDA-graph replay of the device half.
    def discriminator_training_step(self, gt_embeddings, optimizer):
        """latent_gan.py:117-149 (RNG draw order: input latents, then the real-embedding indices)."""
        B = self.config["batch_size"]
        latent_vectors = self.sample_input_latent_vector(B).astype(np.float32)
        real_idxs = np.random.randint(0, gt_embeddings.shape[0], B)
        latent_vectors, real_idxs = self._rank_rows(latent_vectors, real_idxs)
        if isinstance(gt_embeddings, torch.Tensor):         # embedding store in HBM: only the row indices go up
            idx_d = self._to_device(np.asarray(real_idxs, np.int64), torch.int64, count=False)
            real = gt_embeddings.index_select(0, idx_d).to(self.device, torch.float32)
        else:
            real = self._to_device(np.asarray(gt_embeddings[real_idxs], np.float32), torch.float32)
        latent_d = self._to_device(latent_vectors, torch.float32)
        optimizer.begin_step(self.device)
        nl = self.config["num_mlp_layers"]

        def device_half(latent_d, real):
            with torch.no_grad():
                fake = self.generator(latent_d)
            real = real.detach().requires_grad_(True)
            p_d = self.discriminator.params
            o_real = _mlp_forward(p_d, real, nl, second_order=True)
            o_fake = _mlp_forward(p_d, fake.detach(), nl, second_order=True)
            losses = OrderedDict()
            losses["GAN_loss_real"] = networks.gan_d_loss(1, o_real)
            losses["GAN_loss_fake"] = networks.gan_d_loss(0, o_fake)
            losses["gp_loss"] = networks.gradient_regularization(o_real, real)
            losses["loss_sum"] = networks._sum(losses.values())
            self._backward(losses["loss_sum"], [self.discriminator])
            return OrderedDict((k, v.detach()) for k, v in losses.items())
        fn = self._graphed("d", optimizer, device_half, [self.discriminator])
        return self._global_losses(fn(latent_d, real))

    def generator_training_step(self, optimizer):
        """latent_gan.py:151-165."""
        latents = self.sample_input_latent_vector(self.config["batch_size"]).astype(np.float32)
        latents, = self._rank_rows(latents)
        latent_d = self._to_device(latents, torch.float32)
        optimizer.begin_step(self.device)
        nl = self.config["num_mlp_layers"]

        def device_half(latent_d):
            losses = OrderedDict()
            generated = self.generator(latent_d)
            losses["gan_loss"] = networks.gan_g_loss(_mlp_forward(self.discriminator.params, generated, nl))
            losses["loss_sum"] = networks._sum(losses.values())
            self._backward(losses["loss_sum"], [self.generator])
            return OrderedDict((k, v.detach()) for k, v in losses.items())
        fn = self._graphed("g", optimizer, device_half, [self.generator])
        return self._global_losses(fn(latent_d))

    def update_smoothed_weights(self, smoother_alpha=0.999):
        """latent_gan.py:167-174, one kernel over the flat buffer."""
        ops.ema_update(self.generator_smoothed.group.flat, self.generator.group.flat, smoother_alpha)

    def extract_embeddings(self, confignet_model, training_set, max_chunk_size=1000):
        """latent_gan.py:214-230."""
        n_imgs = training_set.imgs.shape[0]
        embeddings = np.zeros((n_imgs, self.config["latent_dim"]), np.float32)
        for begin in range(0, n_imgs, max_chunk_size):
            end = min(begin + max_chunk_size, n_imgs)
            embeddings[begin:end], _ = confignet_model.encode_images(training_set.imgs[begin:end])
        return embeddings

    def setup_logs(self, log_dir, training_set, confignet_model):
        """latent_gan.py:200-212.  Draws from NumPy's global stream what the reference's set-up draws, in its order (the
        logging latents, the InceptionMetrics sample rows - metrics/metrics.py:206 -, the metric latents and rotations), so a
        seeded run enters its first step at the reference's stream position.  TensorBoard and KID / FID are out of scope."""
        if log_dir:
            os.makedirs(log_dir, exist_ok=True)
        n_logged_images = self.config["logging_img_square_size"] ** 2
        self.inputs_for_logs = {"latents": self.sample_input_latent_vector(n_logged_images),
                                "rotations": np.zeros((n_logged_images, 3), np.float32)}
        n = self.config["n_samples_for_metrics"]
        self._metric_sample_idxs = np.random.randint(0, training_set.imgs.shape[0], n)
        self.inputs_for_metrics = {"latents": self.sample_input_latent_vector(n),
                                   "rotations": confignet_model.sample_rotations(n)}

    def write_logs(self, output_dir, step_number, d_loss, g_loss, confignet_model):
        """latent_gan.py:176-198: the checkpoint cadence (<output_dir>/checkpoints/<step, 6 digits> every
        verbose_log_period steps, step 0 included); image grids, TensorBoard scalars and KID / FID are out of scope."""
        if not output_dir or world()[0] != 0:
            return
        if step_number % self.config["verbose_log_period"] == 0:
            checkpoint_output_dir = os.path.join(output_dir, "checkpoints")
            os.makedirs(checkpoint_output_dir, exist_ok=True)
            self.save(checkpoint_output_dir, str(step_number).zfill(6))

    def train(self, training_set, confignet_model, output_dir, log_dir, n_iters):
        """latent_gan.py:232-247: set-up draws, loop structure, one optimizer for both networks, checkpoint cadence."""
        self.setup_logs(log_dir, training_set, confignet_model)
        gt_embeddings = self.extract_embeddings(confignet_model, training_set)
        optimizer = KerasAdam(**self.config["optimizer"])
        for step_number in range(n_iters):
            d_loss = self.discriminator_training_step(gt_embeddings, optimizer)
            g_loss = self.generator_training_step(optimizer)
            self.update_smoothed_weights()
            print("[step: %d] [D loss: %f] [G loss: %f]" % (step_number, d_loss["loss_sum"], g_loss["loss_sum"]))
            for hist, l in ((self.d_losses, d_loss), (self.g_losses, g_loss)):
                for k, v in l.items():
                    hist.setdefault(k, []).append(float(v))
            self.write_logs(output_dir, step_number, d_loss, g_loss, confignet_model)

    def generate_latents(self, n_samples, truncation=1.0):
        """latent_gan.py:249-253 -> (n, latent_dim) float32 NumPy."""
        input_latents = (self.sample_input_latent_vector(n_samples) * truncation).astype(np.float32)
        return self.generator_smoothed.predict(input_latents)
