"""Worker of tests/test_dp_gpu.py: a few first-stage training iterations through the class surface, either in one
process on the global batch or under torch.distributed (NCCL) with the batch sharded by rows.  Rank 0 writes the
per-step losses, the flat gradient the optimizer consumed in iteration 1 and the weights after iterations 1 and 3."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist


def main(out_path, global_batch, n_iters):
    from confignet_b200 import netspec
    from confignet_b200.confignet_first_stage import ConfigNetFirstStage
    from confignet_b200.runtime import KerasAdam
    from confignet_b200.synthetic_data import SyntheticDataset
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda:%d" % local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = {"output_shape": (256, 256, 3), "batch_size": global_batch, "facemodel_inputs": netspec.default_facemodel_inputs(),
           "cuda_graph_warmup": 1}
    model = ConfigNetFirstStage(cfg, device=dev)
    nets = ["synthetic_encoder", "discriminator", "synth_discriminator", "latent_discriminator", "latent_regressor", "generator"]
    for i, name in enumerate(nets):                       # same perturbed weights on every rank
        g = getattr(model, name).group
        arrays = netspec.perturb_params(dict(zip(g.names, g.get_weights())), 500 + i, 0.05)
        g.set_weights([arrays[k] for k in g.names])
    real, synth = SyntheticDataset(24, 256, seed=1), SyntheticDataset(24, 256, seed=2)
    d_opt, g_opt = KerasAdam(**model.config["optimizer"]), KerasAdam(**model.config["optimizer"])
    np.random.seed(11)
    out = {}
    losses = []
    for it in range(n_iters):
        steps = [("d", lambda: model.discriminator_training_step(real, d_opt), ["discriminator"]),
                 ("sd", lambda: model.synth_discriminator_training_step(synth, d_opt), ["synth_discriminator"]),
                 ("ld", lambda: model.latent_discriminator_training_step(synth, d_opt), ["latent_discriminator"]),
                 ("g", lambda: model.generator_training_step(real, synth, g_opt), ["generator", "latent_regressor", "synthetic_encoder"])]
        for tag, fn, touched in steps:
            l = fn()
            losses.append([float(v) for v in l.values()])
            if it == 0:
                for n in touched:      # the gradient Adam consumed: summed over ranks by the all-reduce, scaled 1/world inside the kernel
                    out["grad_%s_%s" % (tag, n)] = (getattr(model, n).group.grad / world).cpu().numpy()
        model.update_smoothed_weights()
        if it in (0, n_iters - 1):
            for n in nets + ["generator_smoothed"]:
                out["w%d_%s" % (it, n)] = getattr(model, n).group.flat.detach().cpu().numpy()
    replayed = sorted(k for k, (_, v) in model._graphs.items() if v.graph is not None)
    out["losses"] = np.array([x for row in losses for x in row])
    out["replayed"] = np.array(replayed)
    if int(os.environ.get("RANK", "0")) == 0:
        np.savez(out_path, **out)
    # clean teardown: graphs first (they hold no NCCL work - the all-reduce runs between two graphs), then the group
    model.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    print("dp_worker done (world %d)" % world, flush=True)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
