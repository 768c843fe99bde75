"""-m gpu, needs 2 GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`): data parallelism on hardware.

Two ranks on 4 images each must reproduce the one-process run on the same 8 images: the global loss values of the first
iteration, the flat gradients the optimizers consumed (one NCCL all-reduce per network between the two captured graphs
of a step, 1/world folded into Adam) and - as far as a sign-like first Adam step (beta_1 = 0) allows - the updated
weights.  The worker processes tear their process group down normally: a hang there
(the round-1 problem with captured NCCL work) fails the test by timeout."""
import os
import subprocess
import sys
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_reproduce_the_single_process_global_batch(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    worker = os.path.join(ROOT, "tests", "dp_worker.py")
    one, two = str(tmp_path / "one.npz"), str(tmp_path / "two.npz")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="0,1")
    r = subprocess.run([sys.executable, worker, one, "8", "3"], env=dict(env, CUDA_VISIBLE_DEVICES="0"), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", worker, two, "8", "3"], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("dp_worker done (world 2)") == 2          # both ranks left through destroy_process_group
    a, b = np.load(one), np.load(two)
    assert list(a["replayed"]) == list(b["replayed"]) == ["d", "g", "latent_d", "synth_d"]
    la, lb = a["losses"], b["losses"]
    n1 = len(la) // 3                                               # loss terms of one iteration
    rel = np.abs(la - lb) / np.maximum(1.0, np.abs(la))
    print("loss differences per iteration:", [float(rel[i * n1:(i + 1) * n1].max()) for i in range(3)])
    assert rel[:19].max() <= 1e-5                                   # D step of iteration 1: identical weights going in
    # the other steps of iteration 1 start from weights that already carry one sign-like Adam update each (lr * g / |g|:
    # an element whose tiny gradient differs in the last bits between 1 x 8 and 2 x 4 images moves by +lr on one side
    # and -lr on the other): 1e-3.  From iteration 2 on the two runs are two DIFFERENT trajectories of a chaotic system
    # (B = 8, untrained GAN; measured 1e-2 at iteration 2, 0.4 at iteration 3 - the same growth two single-GPU runs with a
    # 1e-7 perturbation show): only finiteness is asserted there.
    assert rel[:n1].max() <= 1e-3 and np.isfinite(la).all() and np.isfinite(lb).all()
    for k in a.files:
        if k.startswith("grad_d_") or k.startswith("grad_ld_"):     # gradients at identical weights (first steps of their networks)
            e = np.linalg.norm(a[k] - b[k]) / np.linalg.norm(a[k])
            # two fp32 evaluations of the same gradient with different summation partitions (1 x 8 vs 2 x 4 images, R1
            # double backward included): measured 0.9e-4 .. 1.0e-4 on B200 across kernel revisions
            assert e <= 3e-4, (k, e)
    # weights after iteration 1: the discriminator's first update is sign-like (lr * g / |g|): count disagreeing elements
    for n in ("discriminator", "latent_discriminator"):
        wa, wb = a["w0_" + n], b["w0_" + n]
        frac = float((np.abs(wa - wb) > 1e-6).mean())
        print(n, "elements whose first update differs:", frac)
        assert frac <= 2e-2, (n, frac)
    for k in a.files:
        if k.startswith("w2_"):      # three updates of at most lr per element: the weights themselves stay close
            e = np.linalg.norm(a[k] - b[k]) / np.linalg.norm(a[k])
            assert e <= 5e-2, (k, e)
