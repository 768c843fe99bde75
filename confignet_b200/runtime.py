"""Host-side runtime pieces: flat parameter groups, Keras-protocol network shims, Keras-Adam on flat
buffers, and the data-parallel gradient exchange.

  ParamGroup        one flat HBM buffer per network (+ flat gradient / Adam moments); the per-variable
                    tensors are views, so get_weights()/set_weights() interchange is a memcpy and the
                    optimizer / all-reduce each run as ONE launch per step.
  Network           the slice of the Keras Model protocol the reference's classes use
                    (``m(x)``, ``predict``, ``get_weights``, ``set_weights``, ``trainable_weights``, ``build``).
  KerasAdam         keras.optimizers.Adam semantics [TF-2.1] (confignet_first_stage.py:601-602), incl. the
                    per-optimizer ``iterations`` counter shared by the three discriminators (:608-610).
"""
import ctypes
import math
import os
import weakref
from collections import OrderedDict
import numpy as np
import torch
import torch.distributed as dist
from . import _lib as L
from . import ops


class ParamGroup:
    def __init__(self, spec_or_arrays, device, init=None, trainable=None):
        """spec_or_arrays: OrderedDict name -> ndarray (initial values, Keras layout).
        trainable: optional predicate name -> bool (keras non-trainable weights, e.g. BatchNorm moving statistics,
        stay in the flat buffer for get_weights()/set_weights() but receive no gradient and no update)."""
        arrays = spec_or_arrays
        self.names = list(arrays.keys())
        self.shapes = [tuple(np.shape(arrays[k])) for k in self.names]
        self.sizes = [int(np.prod(s)) if len(s) else 1 for s in self.shapes]
        # 16-byte aligned slots so every view can be read with float4
        self.offsets, off = [], 0
        for n in self.sizes:
            self.offsets.append(off)
            off += (n + 3) // 4 * 4
        self.total = off
        self.frozen, self.on_frozen_change = False, None
        self.device = torch.device(device)
        self.flat = torch.zeros(self.total, device=self.device, dtype=torch.float32)
        self.grad = torch.zeros(self.total, device=self.device, dtype=torch.float32)
        self.params = OrderedDict()
        for name, shape, n, o in zip(self.names, self.shapes, self.sizes, self.offsets):
            v = self.flat[o:o + n].view(shape)
            self.params[name] = v
        if self.flat.is_cuda:      # the conv kernels keep packed images of registered kernels (include/confignet_b200.h)
            L.call("cn_register_params", ops._p(self.flat), self.total * 4)
        self.set_weights([arrays[k] for k in self.names])
        self._train_idx = [i for i, k in enumerate(self.names) if trainable is None or trainable(k)]
        self._trainable = [self.params[self.names[i]] for i in self._train_idx]
        for v in self._trainable:
            v.requires_grad_(True)
        self._off_arr = (ctypes.c_int64 * len(self._train_idx))(*[self.offsets[i] for i in self._train_idx])
        self._n_arr = (ctypes.c_int64 * len(self._train_idx))(*[self.sizes[i] for i in self._train_idx])

    # ---- Keras weight interchange
    def get_weights(self):
        host = self.flat.detach().cpu().numpy()
        return [host[o:o + n].reshape(s).copy() for s, n, o in zip(self.shapes, self.sizes, self.offsets)]

    def set_weights(self, weights):
        if len(weights) != len(self.names):
            raise ValueError("expected %d weight arrays, got %d" % (len(self.names), len(weights)))
        host = np.zeros(self.total, np.float32)
        for w, s, n, o, name in zip(weights, self.shapes, self.sizes, self.offsets, self.names):
            w = np.asarray(w, np.float32)
            if tuple(w.shape) != s:
                raise ValueError("weight %s: expected shape %s, got %s" % (name, s, tuple(w.shape)))
            host[o:o + n] = w.reshape(-1)
        with torch.no_grad():
            self.flat.copy_(torch.from_numpy(host))
        self._changed()

    def copy_from(self, other):
        with torch.no_grad():
            self.flat.copy_(other.flat)
        self._changed()

    def _changed(self):
        """the flat buffer was written by something other than the library's optimizer / EMA entry points"""
        if self.flat.is_cuda:
            L.call("cn_params_changed", ops._p(self.flat))
            if self.frozen and self.on_frozen_change is not None:
                self.on_frozen_change()

    def set_frozen(self, on_change=None):
        """Declares the buffer constant between set_weights() calls (VGG19 / VGGFace, perceptual_loss.py:19-41): captured
        step graphs then keep using its packed kernels instead of re-packing them on every replay.  ``on_change`` is
        called when set_weights() does change it (the owner drops its captured graphs)."""
        self.frozen, self.on_frozen_change = True, on_change
        for v in self.params.values():
            v.requires_grad_(False)
        if self.flat.is_cuda:
            L.call("cn_set_params_frozen", ops._p(self.flat), 1)

    def __del__(self):
        try:
            if self.flat.is_cuda:
                L.call("cn_unregister_params", ops._p(self.flat))
        except Exception:
            pass

    @property
    def trainable_weights(self):
        return list(self._trainable)

    # ---- gradients
    def pack_grads(self, grads):
        """grads: list aligned with ``trainable_weights`` (None = unused variable) -> self.grad (flat)."""
        if len(grads) != len(self._train_idx):
            raise ValueError("expected %d gradients, got %d" % (len(self._train_idx), len(grads)))
        keep = [None if g is None else ops._chk(g) for g in grads]
        ptrs = (ctypes.c_void_p * len(keep))(*[None if g is None else g.data_ptr() for g in keep])
        L.call("cn_multi_copy", len(keep), ptrs, self._off_arr, self._n_arr, ops._p(self.grad), ops._stream())
        return keep       # keeps the sources alive until the copy is enqueued


def _to_numpy(out):
    if isinstance(out, torch.Tensor):
        return out.detach().cpu().numpy()
    if isinstance(out, dict):
        return type(out)((k, _to_numpy(v)) for k, v in out.items())
    if isinstance(out, (tuple, list)):
        return type(out)(_to_numpy(v) for v in out)
    return out


class Network:
    """Keras-Model-like shim around (ParamGroup, forward function)."""

    def __init__(self, group, forward, weights_order=None, **forward_kwargs):
        """weights_order: names in the order keras' get_weights() lists them when that differs from the group's flat
        layout (a nested model lists trainable before non-trainable variables, netspec.real_encoder_keras_order)."""
        self.group = group
        self._forward = forward
        self._kw = forward_kwargs
        self._order = None
        if weights_order is not None:
            if sorted(weights_order) != sorted(group.names):
                raise ValueError("weights_order must be a permutation of the group's variable names")
            self._order = [group.names.index(k) for k in weights_order]

    @property
    def params(self):
        return self.group.params

    def __call__(self, *inputs, **kw):
        k = dict(self._kw); k.update(kw)
        return self._forward(self.group.params, *inputs, **k)

    def predict_device(self, *inputs, **kw):
        """forward without a tape -> device tensors (what the step functions use internally)"""
        with torch.no_grad():
            return self(*inputs, **kw)

    def predict(self, *inputs, **kw):
        """keras Model.predict: forward without a tape -> NumPy arrays (tuples / dicts of them for multi-output models),
        as the reference's callers expect (metrics/metrics.py:54-61, latent_gan.py:249-253)."""
        return _to_numpy(self.predict_device(*inputs, **kw))

    def build(self, input_shape=None):
        return None

    def get_weights(self):
        w = self.group.get_weights()
        return w if self._order is None else [w[i] for i in self._order]

    def set_weights(self, weights):
        weights = list(weights)
        if self._order is not None:
            if len(weights) != len(self._order):
                raise ValueError("expected %d weight arrays, got %d" % (len(self._order), len(weights)))
            by_slot = [None] * len(weights)
            for w, i in zip(weights, self._order):
                by_slot[i] = w
            weights = by_slot
        self.group.set_weights(weights)

    @property
    def trainable_weights(self):
        return self.group.trainable_weights


class KerasAdam:
    """[TF-2.1] Adam (non-amsgrad): t = iterations+1; lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; theta -= lr_t*m/(sqrt(v)+eps), eps = 1e-7."""

    def __init__(self, lr=0.0004, beta_1=0.0, beta_2=0.9, epsilon=1e-7, amsgrad=False, **_):
        if amsgrad:
            raise NotImplementedError("amsgrad is not used by ConfigNet (confignet_first_stage.py:50)")
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, epsilon
        self.iterations = 0
        self.state = {}

    def _lr_t(self, t):
        return self.lr * math.sqrt(1 - self.b2 ** t) / (1 - self.b1 ** t)

    def reset(self):
        """back to a freshly constructed optimizer, keeping the moment buffers' addresses (captured graphs hold them)"""
        self.iterations = 0
        for m, v in self.state.values():
            m.zero_(); v.zero_()

    def _state(self, g):
        st = self.state.get(id(g))
        if st is None:
            st = (torch.zeros_like(g.flat), torch.zeros_like(g.flat))
            self.state[id(g)] = st
        return st

    def apply_flat(self, groups, gscale=1.0):
        """One optimizer step over the flat gradient buffers of ``groups`` (already packed / reduced)."""
        lr_t = self._lr_t(self.iterations + 1)
        for g in groups:
            st = self._state(g)
            ops.adam_ema_step(g.flat, g.grad, st[0], st[1], None, lr_t, self.b1, self.b2, self.eps, 0.0, gscale)
        self.iterations += 1

    # ---- split form for captured steps: the host half advances `iterations` and refreshes the bias-corrected
    #      learning rate in a device scalar; the device half (which may be a CUDA-graph replay) reads it
    def begin_step(self, device):
        if getattr(self, "lr_dev", None) is None:
            self.lr_dev = torch.zeros(1, device=device, dtype=torch.float32)
        # the value travels as a kernel argument (no staging buffer the host could overwrite while it runs ahead of the GPU)
        self.lr_dev.fill_(self._lr_t(self.iterations + 1))
        self.iterations += 1

    def apply_flat_device_lr(self, groups, gscale=1.0):
        for g in groups:
            st = self._state(g)
            ops.adam_ema_step_dev(g.flat, g.grad, st[0], st[1], None, self.lr_dev, self.b1, self.b2, self.eps, 0.0, gscale)


class GraphedFn:
    """Runs the device half of a training step as CUDA-graph replays.

    ``fn(*tensors) -> OrderedDict of scalar tensors`` is everything up to and including the packing of the flat
    gradient buffers of ``groups`` (forward, losses, backward incl. the R1 double backward); ``finish()`` is the
    optimizer launch on those buffers.  On one GPU both are captured into ONE graph.  Under data parallelism the
    gradient all-reduce sits between them and stays OUTSIDE the graphs: graph 1 (``fn``), NCCL all-reduce on the flat
    buffers, graph 2 (``finish``).  No NCCL work is ever captured, so the process group can be torn down normally.

    The first ``warm`` calls run eagerly (they are real steps and let the library create its plans and buffers), the
    next call is captured and replayed, later calls copy their inputs into the captured input tensors and replay.  A
    replay runs the optimizer kernels without the host code that tells the library its packed-weight cache is out of
    date, so every replay marks the flat buffers of ``groups`` as changed (cn_params_changed).  Any capture error
    switches the wrapper back to eager execution for good.  Shapes must not change between calls (new shape -> eager)."""
    ENABLED = os.environ.get("CN_GRAPHS", "1") != "0"
    REPLAYED_LAUNCHES = 0          # library kernels launched through graph replays (cn_launch_count only sees eager ones)
    _live = weakref.WeakSet()      # every wrapper that holds a captured graph (release_graphs)

    def __init__(self, fn, finish=None, groups=(), warm=2, reduce_groups=None):
        self.fn, self.finish, self.groups, self.warm = fn, finish, list(groups), warm
        self.rgroups = self.groups if reduce_groups is None else list(reduce_groups)     # gradients exchanged between ranks
        self.calls, self.graph, self.graph2, self.failed = 0, None, None, False
        self.static_in, self.static_out, self.sig = None, None, None
        self.launches = 0

    @staticmethod
    def _sig(tensors):
        return tuple((tuple(t.shape), t.dtype) for t in tensors)

    def _eager(self, tensors):
        out = self.fn(*tensors)
        if self.finish is not None:
            self.finish(allreduce_grads(self.rgroups))
        return out

    def release(self):
        """drops the captured graphs (and their memory pool); the next call captures again"""
        self.graph = self.graph2 = self.static_in = self.static_out = None
        self.calls = min(self.calls, self.warm)

    def _capture(self, tensors):
        self.sig = self._sig(tensors)
        self.static_in = [t.clone() for t in tensors]
        torch.cuda.synchronize()
        ws = world()[1]
        lib = L.load()
        n0 = int(lib.cn_launch_count(0))
        g = torch.cuda.CUDAGraph()
        g2 = None
        with torch.cuda.graph(g):
            out = self.fn(*self.static_in)
            if self.finish is not None and ws == 1:
                self.finish(1.0)
        if self.finish is not None and ws > 1:
            g.replay()                                 # the captured step's gradients are real before they are reduced
            scale = allreduce_grads(self.rgroups)
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, pool=g.pool()):
                self.finish(scale)
            g2.replay()
            first_done = True
        else:
            first_done = False
        self.launches = int(lib.cn_launch_count(0)) - n0       # recorded, not executed
        lib.cn_launch_count_add(-self.launches)
        self.graph, self.graph2, self.static_out = g, g2, out
        L.call("cn_graphs_captured")
        GraphedFn._live.add(self)
        return first_done

    def __call__(self, *tensors):
        if not GraphedFn.ENABLED or self.failed or ops.PROFILE[0] is not None or not tensors[0].is_cuda:
            return self._eager(tensors)
        self.calls += 1
        if self.calls <= self.warm:
            return self._eager(tensors)
        done = False
        if self.graph is None:
            try:
                done = self._capture(tensors)
            except Exception as e:                                   # pragma: no cover - depends on the driver
                self.failed = True
                self.release()
                torch.cuda.synchronize()
                import warnings
                warnings.warn("CUDA-graph capture of a training step failed (%s); running eagerly" % (str(e)[:200],))
                return self._eager(tensors)
        elif self._sig(tensors) != self.sig:
            return self._eager(tensors)
        else:
            for s, t in zip(self.static_in, tensors):
                s.copy_(t, non_blocking=True)
        if not done:
            self.graph.replay()
            if self.graph2 is not None:
                allreduce_grads(self.rgroups)
                self.graph2.replay()
        for g in self.groups:                   # the replayed optimizer kernels changed these buffers
            L.call("cn_params_changed", ops._p(g.flat))
        GraphedFn.REPLAYED_LAUNCHES += self.launches
        return OrderedDict((k, v.clone()) for k, v in self.static_out.items())


class InferenceGraphs:
    """The small-batch evaluation path (generate_images at batch 1-6 in the demo loop,
    evaluation/confignet_demo.py:154-160; the controllability sweep, metrics/metrics.py:98-102): one captured CUDA graph
    per (generator network, batch size) - ~45 launches replayed as one - whose packed kernels are the ones of the eager
    warm-up call (the buffer is declared frozen for the capture, so the graph holds no pack launches).  The graph is
    re-captured when the network's change counter has moved (training, EMA, set_weights)."""
    MAX_BATCH = 8

    def __init__(self):
        self.cache = {}

    def clear(self):
        self.cache.clear()

    @staticmethod
    def _eager(net, zs, rot):
        keys = ("z_3d_0", "z_3d_1", "z_2d_0", "z_2d_1", "z_2d_2")
        d = dict(zip(keys, zs))
        d["rotation"] = rot
        return ops.to_uint8(net.predict_device(d))

    def run(self, net, zs, rot, borrow=False):
        """zs: 5 device tensors (B, latent); rot: (B, 3) device tensor of Euler angles -> uint8 (B, H, W, 3) device tensor.
        ``borrow``: return the graph's own output buffer (valid until the next replay) instead of a copy - for a caller that
        moves it to the host right away."""
        B = zs[0].shape[0]
        if not GraphedFn.ENABLED or B > self.MAX_BATCH or ops.PROFILE[0] is not None or not zs[0].is_cuda:
            return self._eager(net, zs, rot)
        flat = net.group.flat
        epoch = int(L.load().cn_params_epoch(ops._p(flat)))
        key = (id(net.group), B, zs[0].shape[1])
        e = self.cache.get(key)
        if e is None or e["epoch"] != epoch:
            if e is not None and e.get("failed"):
                return self._eager(net, zs, rot)
            try:
                s_zs = [z.clone() for z in zs]
                s_rot = rot.clone()
                self._eager(net, s_zs, s_rot)                      # plans, buffers and the packed kernels of the CURRENT weights
                torch.cuda.synchronize()
                frozen = net.group.frozen
                L.call("cn_set_params_frozen", ops._p(flat), 1)
                lib = L.load()
                n0 = int(lib.cn_launch_count(0))
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        out = self._eager(net, s_zs, s_rot)
                finally:
                    L.call("cn_set_params_frozen", ops._p(flat), 1 if frozen else 0)
                launches = int(lib.cn_launch_count(0)) - n0              # recorded, not executed
                lib.cn_launch_count_add(-launches)
                L.call("cn_graphs_captured")
                e = self.cache[key] = dict(graph=g, zs=s_zs, rot=s_rot, out=out, epoch=epoch, launches=launches)
            except Exception as ex:                                  # pragma: no cover - depends on the driver
                torch.cuda.synchronize()
                self.cache[key] = dict(failed=True, epoch=epoch)
                import warnings
                warnings.warn("CUDA-graph capture of generate_images failed (%s); running eagerly" % (str(ex)[:200],))
                return self._eager(net, zs, rot)
        if e.get("failed"):
            return self._eager(net, zs, rot)
        for s, z in zip(e["zs"], zs):
            if s.data_ptr() != z.data_ptr():
                s.copy_(z, non_blocking=True)
        e["rot"].copy_(rot, non_blocking=True)
        e["graph"].replay()
        GraphedFn.REPLAYED_LAUNCHES += e["launches"]
        return e["out"] if borrow else e["out"].clone()


def release_graphs():
    """Drops every captured step graph of this process (their private memory pools go back to the allocator)."""
    for g in list(GraphedFn._live):
        g.release()
    GraphedFn._live.clear()
    import gc
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


H2D_BYTES = [0]        # bytes staged host -> device by StepGraphs._to_device (bench.py reads and resets it)


class StepGraphs:
    """What the model classes share around an optimizer step: tape.gradient + packing, the eager whole step, the
    CUDA-graph wrapper per step, global loss values under data parallelism.  Expects ``self._graphs`` (dict),
    ``self.config`` and ``self.device``."""

    def _to_device(self, arr, dtype, count=True):
        """host array -> device tensor of `dtype`: pinned staging + copy on a side stream, so the next step's batch goes
        up while the current step computes and the host never blocks on a pageable copy."""
        if isinstance(arr, torch.Tensor):
            return arr.to(self.device, dtype)
        arr = np.ascontiguousarray(arr)
        if not arr.flags.writeable:            # a slice of the dataset's read-only np.memmap (neural_renderer_dataset.py:346)
            arr = np.array(arr)
        t = torch.from_numpy(arr)
        if count:
            H2D_BYTES[0] += t.numel() * t.element_size()
        if self.device.type != "cuda":
            return t.to(self.device).to(dtype)
        if getattr(self, "_upload_stream", None) is None:
            self._upload_stream = torch.cuda.Stream(self.device)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._upload_stream):
            d = t.pin_memory().to(self.device, non_blocking=True)
        cur.wait_stream(self._upload_stream)
        d.record_stream(cur)
        return d.to(dtype)

    def _backward(self, loss, nets):
        """tape.gradient(loss, trainable_weights) for the networks of one optimizer step, packed into their flat
        gradient buffers -> the groups."""
        groups = [n.group if hasattr(n, "group") else n for n in nets]
        params = [p for g in groups for p in g.trainable_weights]
        grads = torch.autograd.grad(loss, params, allow_unused=True)
        i = 0
        for g in groups:
            k = len(g.trainable_weights)
            g.pack_grads(grads[i:i + k])
            i += k
        return groups

    def _apply(self, optimizer, loss, nets, device_lr=False):
        """optimizer.apply_gradients(zip(tape.gradient(...), trainable_weights)): backward, gradient exchange, Adam.
        device_lr: the host half of the optimizer step (KerasAdam.begin_step) already ran."""
        groups = self._backward(loss, nets)
        gscale = allreduce_grads(groups)
        if device_lr:
            optimizer.apply_flat_device_lr(groups, gscale)
        else:
            optimizer.apply_flat(groups, gscale)

    def _graphed(self, name, optimizer, fn, nets, dp_ok=True, reduce_groups=None):
        """One CUDA-graph wrapper per step, bound to the optimizer whose moment buffers the captured region holds.
        ``fn`` ends with _backward(loss, nets); the wrapper adds the gradient exchange and the Adam launch.
        dp_ok = False: ``fn`` itself contains a collective (the all-gathered batch-statistics loss of stage 2), which
        must not be captured - such a step runs eagerly under data parallelism.  reduce_groups: the groups whose
        gradients are exchanged when that is not all of them (fine-tuning keeps per-image variables rank-local)."""
        groups = [n.group if hasattr(n, "group") else n for n in nets]
        rgroups = groups if reduce_groups is None else list(reduce_groups)

        def finish(gscale, o=optimizer, gr=groups):
            o.apply_flat_device_lr(gr, gscale)
        if not self.config.get("cuda_graphs", True) or (world()[1] > 1 and not dp_ok):
            def eager(*tensors):
                out = fn(*tensors)
                finish(allreduce_grads(rgroups))
                return out
            return eager
        entry = self._graphs.get(name)
        if entry is None or entry[0] is not optimizer:
            # a new optimizer object (a second train() call): its moment buffers and learning-rate scalar are not the
            # ones the old graph captured - drop that graph (and its memory pool) and start over
            entry = self._graphs[name] = (optimizer, GraphedFn(fn, finish, groups, warm=int(self.config.get("cuda_graph_warmup", 2)),
                                                               reduce_groups=rgroups))
        return entry[1]

    def drop_graphs(self):
        """forget the captured step graphs (the next step runs eagerly again and is re-captured)"""
        for _, g in self._graphs.values():
            g.release()
        self._graphs.clear()
        if getattr(self, "_infer", None) is not None:
            self._infer.clear()

    def close(self):
        """release the captured step graphs and their memory pools (call before tearing the process group down)"""
        self.drop_graphs()
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def _global_losses(self, losses):
        """Data parallelism: every loss term is a batch mean over this rank's equal shard; the value of the global batch
        (what a single-GPU run reports and logs) is the mean over ranks - one small all-reduce per step."""
        if world()[1] == 1:
            return losses
        keys = list(losses.keys())
        t = torch.stack([losses[k].reshape(()) for k in keys])
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t = t / world()[1]
        return OrderedDict((k, t[i]) for i, k in enumerate(keys))


# ------------------------------------------------------------------------------------------------ data parallel
def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def coalesce_grads(groups):
    """One allocation behind the flat gradient buffers of the networks ONE optimizer step updates (generator + latent
    regressor + synthetic encoder (+ real encoder)): their gradient exchange is then a single all-reduce over the span
    instead of one call per network - the exchange sits between two graph replays, so every call is exposed latency
    ('one flat gradient buffer per optimizer step').  Must run before any graph is captured over
    the groups (the buffers' addresses are baked into captured launches)."""
    groups = [g for g in groups if g is not None]
    if len(groups) < 2:
        return
    arena = torch.zeros(sum(g.total for g in groups), device=groups[0].device, dtype=torch.float32)
    off = 0
    for g in groups:
        g.grad = arena[off:off + g.total]            # totals are multiples of 4 floats: every slot stays 16-byte aligned
        g._grad_arena = (arena, off)
        off += g.total


def _grad_spans(groups):
    """-> tensors to all-reduce: maximal runs of groups whose buffers are adjacent in one arena collapse into one span"""
    spans, i = [], 0
    while i < len(groups):
        arena_off = getattr(groups[i], "_grad_arena", None)
        j = i
        if arena_off is not None:
            arena, end = arena_off[0], arena_off[1] + groups[i].grad.numel()
            while j + 1 < len(groups):
                nxt = getattr(groups[j + 1], "_grad_arena", None)
                if nxt is None or nxt[0] is not arena or nxt[1] != end:
                    break
                end += groups[j + 1].grad.numel()
                j += 1
            spans.append(arena[arena_off[1]:end] if j > i else groups[i].grad)
        else:
            spans.append(groups[i].grad)
        i = j + 1
    return spans


def allreduce_grads(groups):
    """Sum the flat gradient buffers over ranks (fp32 NCCL all-reduce: one call per run of coalesced buffers, else one
    per network) and return the 1/world scale to fold into the optimizer kernel.  Losses are batch means over equal
    shards, so the averaged gradient equals the single-process gradient of the global batch."""
    rank, ws = world()
    if ws == 1:
        return 1.0
    for span in _grad_spans(list(groups)):
        dist.all_reduce(span, op=dist.ReduceOp.SUM)
    return 1.0 / ws


def shard_rows(n_global):
    """Row slice [lo, hi) of a globally sampled batch owned by this rank."""
    rank, ws = world()
    if n_global % ws != 0:
        raise ValueError("global batch %d is not divisible by world size %d" % (n_global, ws))
    per = n_global // ws
    return rank * per, (rank + 1) * per


class _GatherRows(torch.autograd.Function):
    """All-gather of row shards for the one loss on the path that uses batch statistics
    (compute_normalized_latent_regression_loss, confignet_second_stage.py:96-101).
    Every rank then evaluates the same global loss; the backward keeps this rank's rows and multiplies by the
    world size, which the 1/world gradient averaging of allreduce_grads() undoes."""

    @staticmethod
    def forward(ctx, x):
        rank, ws = world()
        x = x.contiguous()
        parts = [torch.empty_like(x) for _ in range(ws)]
        dist.all_gather(parts, x)
        ctx.rows = x.shape[0]
        return torch.cat(parts, dim=0)

    @staticmethod
    def backward(ctx, g):
        rank, ws = world()
        return g[rank * ctx.rows:(rank + 1) * ctx.rows] * float(ws)


def gather_rows(x):
    return x if world()[1] == 1 else _GatherRows.apply(x)
