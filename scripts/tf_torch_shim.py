"""A torch-backed (float64, CPU) stand-in for the slice of TensorFlow 2.1 / tf.keras that the reference's model code
uses (confignet/dnn_models/*.py, confignet_utils.py, losses.py): enough to EXECUTE the reference's own class bodies -
HologanGenerator, Conv{2,3}dAdaIn, AdaIn, MLPSimple, HologanDiscriminator, DiscrBlock, InstanceNormalization,
HologanLatentRegressor, SyntheticDataEncoder, compute_discriminator_loss with its nested GradientTape - without
TensorFlow (not installable here).  Every layer below is a few lines stating the [TF-2.1] semantics SURVEY.md section 8c
lists (SAME padding rule, LeakyReLU default alpha 0.3, tf.nn.leaky_relu 0.2, LayerNormalization epsilon 1e-3 with
population variance, nearest x2 UpSampling, Keras kernel layouts, cross-correlation); the ARCHITECTURE - which layer
follows which, channel counts, kernel sizes, which latent feeds which AdaIN, reshape orders, loss formulas - comes from
the reference code that runs on top of it.  Used only by scripts/make_golden_models_from_reference.py (build container).
"""
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

DT = torch.float64


def T(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x), dtype=DT if np.asarray(x).dtype.kind == "f" else None)


class NT(torch.Tensor):
    """torch.Tensor that accepts a NumPy array on the left of * + - (tf.Tensor does: losses.py:11 multiplies NumPy
    label arrays with tf.math.softplus(...))."""

    def __rmul__(self, o):
        return torch.Tensor.__mul__(self, T(o))

    def __radd__(self, o):
        return torch.Tensor.__add__(self, T(o))

    def __rsub__(self, o):
        return torch.Tensor.__neg__(torch.Tensor.__sub__(self, T(o)))

    def __mul__(self, o):
        return torch.Tensor.__mul__(self, T(o) if isinstance(o, (np.ndarray, tuple, list)) else o)

    def __add__(self, o):
        return torch.Tensor.__add__(self, T(o) if isinstance(o, (np.ndarray, tuple, list)) else o)

    def __sub__(self, o):
        return torch.Tensor.__sub__(self, T(o) if isinstance(o, (np.ndarray, tuple, list)) else o)

    def __truediv__(self, o):
        return torch.Tensor.__truediv__(self, T(o).to(DT) if isinstance(o, np.ndarray) else o)

    def numpy(self):                       # tf.Tensor.numpy() / tf.Variable.numpy(): the value, whatever the tape state
        return self.detach().as_subclass(torch.Tensor).numpy()


def _ints(shape):
    return tuple(int(s) for s in shape)


# ------------------------------------------------------------------------------------------------ tf.*
tf = types.ModuleType("tensorflow")
tf.float32, tf.int32 = "float32", "int32"
tf.cast = lambda x, dtype: T(x).to(torch.int64 if dtype == "int32" else DT)
tf.convert_to_tensor = lambda x, dtype=None: tf.cast(x, dtype) if dtype else T(x)
tf.shape = lambda x: tuple(T(x).shape)
tf.constant = lambda v, shape=None, name=None, dtype=None: torch.full(_ints(shape), float(v), dtype=DT) if shape else torch.tensor(v, dtype=DT)
tf.tile = lambda x, reps: T(x).repeat(*_ints(reps))
tf.expand_dims = lambda x, axis: T(x).unsqueeze(axis)
tf.matmul = lambda a, b: T(a) @ T(b)
tf.transpose = lambda x, perm: T(x).permute(*perm)
tf.clip_by_value = lambda x, lo, hi: T(x).clamp(lo, hi)
tf.reshape = lambda x, shape: T(x).reshape(_ints(shape))
tf.floor = lambda x: T(x).floor()
tf.range = lambda a, b=None: torch.arange(int(a)) if b is None else torch.arange(int(a), int(b))
tf.stack = lambda xs, axis=0: torch.stack([T(x) for x in xs], dim=axis)
tf.concat = lambda xs, axis: torch.cat([T(x) for x in xs], dim=axis)
tf.gather = lambda x, idx, axis=0: T(x).index_select(axis, torch.as_tensor(idx).to(torch.int64))
tf.gather_nd = lambda params, idx: T(params)[tuple(T(idx).to(torch.int64).unbind(dim=1))]
tf.sin, tf.cos, tf.sqrt, tf.square = (lambda x: T(x).sin()), (lambda x: T(x).cos()), (lambda x: T(x).sqrt()), (lambda x: T(x) ** 2)
tf.squeeze = lambda x, axis=None: T(x).squeeze(axis) if axis is not None else T(x).squeeze()
tf.ones = lambda shape, dtype=None: torch.ones(_ints(shape), dtype=DT)
tf.zeros = lambda shape, dtype=None: torch.zeros(_ints(shape), dtype=DT)


def _axes(axis):
    return tuple(axis) if isinstance(axis, (list, tuple, range)) else axis


def _reduce(fn):
    def f(x, axis=None, keepdims=False):
        if isinstance(x, (list, tuple)) and not isinstance(x, torch.Tensor):
            x = torch.stack([T(v) for v in x])
        x = T(x)
        return fn(x) if axis is None else fn(x, dim=_axes(axis), keepdim=keepdims)
    return f


tf.reduce_mean = _reduce(torch.mean)
tf.reduce_sum = _reduce(torch.sum)
tf.reduce_prod = lambda x, axis=None: int(np.prod([int(v) for v in x]))          # only ever applied to shapes
tf.nn = types.SimpleNamespace(leaky_relu=lambda x, alpha=0.2: F.leaky_relu(T(x), alpha),    # [TF-2.1] default alpha 0.2
                              tanh=lambda x: torch.tanh(T(x)))
tf.math = types.SimpleNamespace(softplus=lambda x: F.softplus(T(x)).as_subclass(NT),
                                reduce_variance=lambda x, axis=None, keepdims=False: T(x).var(dim=axis, unbiased=False, keepdim=keepdims))
tf.losses = types.SimpleNamespace(mean_squared_error=lambda a, b: ((T(a) - T(b)) ** 2).mean(dim=-1))
tf.compat = types.SimpleNamespace(v1=types.SimpleNamespace(initializers=types.SimpleNamespace(ones=lambda: "ones")))


class GradientTape:
    """tf.GradientTape over torch autograd: gradient(target, source) = d sum(target) / d source, differentiable again
    (the reference calls it inside an outer tape: losses.py:26-43 within confignet_first_stage.py:469-472)."""

    def __init__(self, persistent=False):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def watch(self, x):
        x.requires_grad_(True)

    def gradient(self, target, sources):
        if isinstance(sources, (list, tuple)):
            return list(torch.autograd.grad(T(target).sum(), list(sources), create_graph=True, allow_unused=True))
        return torch.autograd.grad(T(target).sum(), sources, create_graph=True)[0]


tf.GradientTape = GradientTape
tf.Variable = lambda x, **kw: T(np.asarray(x, np.float64)).clone().requires_grad_(True).as_subclass(NT)


# ------------------------------------------------------------------------------------------------ keras
def _conv(a):
    """NumPy (and nested list / tuple / dict of NumPy) -> torch, as Keras does for layer inputs"""
    if isinstance(a, np.ndarray):
        return T(a.astype(np.float64) if a.dtype.kind == "f" else a)
    if isinstance(a, list):
        return [_conv(v) for v in a]
    if isinstance(a, tuple):
        return tuple(_conv(v) for v in a)
    if isinstance(a, dict):
        return type(a)((k, _conv(v)) for k, v in a.items())
    return a


def _nt(o):
    """layer outputs behave like tf.Tensor next to NumPy operands"""
    if isinstance(o, torch.Tensor):
        return o.as_subclass(NT)
    if isinstance(o, dict):
        return type(o)((k, _nt(v)) for k, v in o.items())
    if isinstance(o, (list, tuple)):
        return type(o)(_nt(v) for v in o)
    return o


class Layer:
    def __init__(self, name=None, **kw):
        self.name = name
        self.built = False
        self.weights = []          # creation order

    def __setattr__(self, name, value):
        """[TF-2.1] base_layer.py Layer.__setattr__: a Layer, or a list / dict (wrapped into a trackable container, which
        'has weights' even while empty), is appended to self._layers once, by identity, in ASSIGNMENT order."""
        object.__setattr__(self, name, value)
        if name != "weights" and isinstance(value, (Layer, list, dict)):
            tracked = self.__dict__.setdefault("_tracked", [])
            if not any(t is value for t in tracked):
                tracked.append(value)

    def add_weight(self, shape=None, name=None, initializer=None, regularizer=None, constraint=None, **kw):
        w = torch.ones(_ints(shape), dtype=DT) if initializer == "ones" else torch.zeros(_ints(shape), dtype=DT)
        w.requires_grad_(True)
        self.weights.append((name, w))
        return w

    def build(self, input_shape):
        self.built = True

    def __call__(self, *args, **kwargs):
        args = tuple(_conv(a) for a in args)               # Keras accepts NumPy inputs
        if not self.built and args and isinstance(args[0], torch.Tensor):
            self.build(tuple(args[0].shape))
            self.built = True
        return _nt(self.call(*args, **kwargs))


def tracked_layers(layer):
    """[TF-2.1] trackable_layer_utils.filter_empty_layer_containers(layer._layers): containers are flattened in place
    (lists in list order, dict wrappers in SORTED KEY order - data_structures._DictWrapper._values), layers kept once."""
    out, seen, stack = [], set(), list(layer.__dict__.get("_tracked", []))[::-1]
    while stack:
        o = stack.pop()
        if id(o) in seen:
            continue
        seen.add(id(o))
        if isinstance(o, Layer):
            out.append(o)
        elif isinstance(o, dict):
            stack.extend([o[k] for k in sorted(o)][::-1])
        elif isinstance(o, (list, tuple)):
            stack.extend(list(o)[::-1])
    return out


def _dedup(ws):
    out, seen = [], set()
    for w in ws:
        if id(w) not in seen:
            seen.add(id(w))
            out.append(w)
    return out


def keras_trainable_weights(layer):
    """[TF-2.1] Layer.trainable_weights / Network.trainable_weights: own variables in creation order, then the tracked
    sub-layers' (gather_trainable_weights), de-duplicated.  A stand-in for a library model may supply the two lists."""
    if hasattr(layer, "keras_trainable"):
        return list(layer.keras_trainable)
    own = [w for _, w in layer.weights]                    # every variable the shim's layers create is trainable
    return _dedup(own + [w for l in tracked_layers(layer) for w in keras_trainable_weights(l)])


def keras_non_trainable_weights(layer):
    if hasattr(layer, "keras_non_trainable"):
        return list(layer.keras_non_trainable)
    return _dedup([w for l in tracked_layers(layer) for w in keras_non_trainable_weights(l)])


def keras_layer_weights(layer):
    """[TF-2.1] Layer.weights / Network.weights = trainable_weights + non_trainable_weights (so a NESTED model lists all
    its trainable variables before its non-trainable ones)."""
    return _dedup(keras_trainable_weights(layer) + keras_non_trainable_weights(layer))


def keras_get_weights_order(model):
    """[TF-2.1] network.py Network.get_weights: `for layer in self.layers: weights += layer.weights`."""
    return [w for l in tracked_layers(model) for w in keras_layer_weights(l)]


class InputSpec:
    def __init__(self, **kw):
        pass


class Model(Layer):
    def __init__(self, *a, **kw):
        Layer.__init__(self)

    def __call__(self, *args, **kwargs):
        return _nt(self.call(*tuple(_conv(a) for a in args), **kwargs))

    def build(self, input_shape):
        """[TF-2.1] training.py Model.build on a subclassed model: calls the model on placeholders of the given shape(s),
        which builds every layer the call reaches."""
        def zeros(shape):
            return torch.zeros(tuple(1 if d is None else int(d) for d in shape), dtype=DT)
        nested = isinstance(input_shape, list) or (isinstance(input_shape, tuple) and input_shape and isinstance(input_shape[0], (tuple, list)))
        with torch.no_grad():
            self([zeros(sh) for sh in input_shape] if nested else zeros(input_shape))

    def get_weights(self):
        """[TF-2.1] network.py Network.get_weights, as NumPy arrays"""
        return [w.detach().numpy().copy() for w in keras_get_weights_order(self)]

    def set_weights(self, weights):
        """[TF-2.1] network.py Network.set_weights: consumed layer by layer in the same order; shapes must agree"""
        mine = keras_get_weights_order(self)
        if len(mine) != len(weights):
            raise ValueError("You called `set_weights(weights)` on a model with %d variables with a list of %d arrays" % (len(mine), len(weights)))
        with torch.no_grad():
            for w, v in zip(mine, weights):
                v = torch.as_tensor(np.asarray(v)).to(DT)
                if tuple(v.shape) != tuple(w.shape):
                    raise ValueError("Layer weight shape %s not compatible with provided weight shape %s" % (tuple(w.shape), tuple(v.shape)))
                w.copy_(v)


class Sequential(Model):
    def __init__(self, layers=None, name=None):
        Model.__init__(self)
        self.layers = []
        self._out_shape = None
        for l in (layers or []):
            self.add(l)

    def add(self, layer):
        """[TF-2.1] sequential.py Sequential.add: a first layer that was given input_shape is called on an Input of that shape,
        every later layer on the running output - so such a stack owns its variables before it is ever called."""
        shape = getattr(layer, "_input_shape", None) if not self.layers else self._out_shape
        self.layers.append(layer)
        if shape is not None and not layer.built:
            layer.build((None,) + tuple(shape))
            layer.built = True
        self._out_shape = (layer.units,) if isinstance(layer, Dense) and shape is not None else \
            (shape if isinstance(layer, LeakyReLU) else None)

    def call(self, x):
        for l in self.layers:
            x = l(x)
        return x


def _act(a):
    if a is None:
        return lambda x: x
    if callable(a):
        return a
    return {"tanh": torch.tanh, "relu": torch.relu, "linear": (lambda x: x)}[a]


def _same_pad(n, k, s):
    out = -(-n // s)
    tot = max((out - 1) * s + k - n, 0)
    return tot // 2, tot - tot // 2            # [TF-2.1] SAME: the smaller half goes in front


class _ConvND(Layer):
    nd = 2

    def __init__(self, filters, kernel_size, strides=1, padding="valid", activation=None, use_bias=True, input_shape=None,
                 name=None, **kw):
        Layer.__init__(self, name)
        self.filters, self.padding, self.use_bias = int(filters), padding, use_bias
        self.k = _ints(kernel_size) if isinstance(kernel_size, (tuple, list)) else (int(kernel_size),) * self.nd
        self.s = _ints(strides) if isinstance(strides, (tuple, list)) else (int(strides),) * self.nd
        self.activation = _act(activation)

    def build(self, input_shape):
        cin = int(input_shape[-1])
        self.kernel = self.add_weight(shape=self.k + (cin, self.filters), name="kernel")     # Keras layout (k..., Cin, Cout)
        self.bias = self.add_weight(shape=(self.filters,), name="bias") if self.use_bias else None

    def call(self, x):
        nd = self.nd
        xc = x.permute(0, nd + 1, *range(1, nd + 1))                                   # channels-last -> channels-first
        w = self.kernel.permute(nd + 1, nd, *range(nd))                                # -> (Cout, Cin, k...)
        if self.padding == "same":
            pads = []
            for d in reversed(range(nd)):
                pads += list(_same_pad(xc.shape[2 + d], self.k[d], self.s[d]))
            xc = F.pad(xc, pads)
        y = (F.conv2d if nd == 2 else F.conv3d)(xc, w, self.bias, stride=self.s)        # cross-correlation, as TF
        return self.activation(y.permute(0, *range(2, nd + 2), 1))


class Conv2D(_ConvND):
    nd = 2


class Conv3D(_ConvND):
    nd = 3


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None,
                 input_shape=None, name=None, **kw):
        Layer.__init__(self, name)
        self.units, self.use_bias, self.activation = int(units), use_bias, _act(activation)
        self._input_shape = None if input_shape is None else _ints(input_shape)

    def build(self, input_shape):
        self.kernel = self.add_weight(shape=(int(input_shape[-1]), self.units), name="kernel")
        self.bias = self.add_weight(shape=(self.units,), name="bias") if self.use_bias else None

    def call(self, x):
        y = x @ self.kernel
        return self.activation(y + self.bias if self.bias is not None else y)


class LeakyReLU(Layer):
    def __init__(self, alpha=0.3, **kw):                  # [TF-2.1] keras.layers.LeakyReLU default alpha = 0.3
        Layer.__init__(self)
        self.alpha = alpha

    def call(self, x):
        return F.leaky_relu(x, self.alpha)


class _UpSampling(Layer):
    def __init__(self, size=2, **kw):
        Layer.__init__(self)

    def call(self, x):                                    # nearest, factor 2 on every spatial axis
        for d in range(1, x.dim() - 1):
            x = x.repeat_interleave(2, dim=d)
        return x


class Reshape(Layer):
    def __init__(self, target_shape, name=None, **kw):
        Layer.__init__(self, name)
        self.target_shape = _ints(target_shape)

    def call(self, x):
        return x.reshape((x.shape[0],) + self.target_shape)


class Lambda(Layer):
    def __init__(self, fn, name=None, **kw):
        Layer.__init__(self, name)
        self.fn = fn

    def call(self, x):
        return self.fn(x)


class LayerNormalization(Layer):
    def __init__(self, axis=-1, epsilon=1e-3, center=True, scale=True, **kw):      # [TF-2.1] default epsilon 1e-3
        Layer.__init__(self)
        assert not center and not scale, "the reference only uses center=False, scale=False"
        self.axis, self.epsilon = _axes(axis), epsilon

    def call(self, x):
        mu = x.mean(dim=self.axis, keepdim=True)
        var = ((x - mu) ** 2).mean(dim=self.axis, keepdim=True)                    # nn.moments: population variance
        return (x - mu) * torch.rsqrt(var + self.epsilon)


K = types.ModuleType("tensorflow.keras.backend")
K.mean = lambda x, axis=None, keepdims=False: T(x).mean(dim=_axes(axis), keepdim=keepdims) if axis is not None else T(x).mean()
K.std = lambda x, axis=None, keepdims=False: T(x).var(dim=_axes(axis), unbiased=False, keepdim=keepdims).sqrt()
K.int_shape = lambda x: tuple(int(s) for s in x.shape)
K.reshape = lambda x, shape: T(x).reshape(_ints(shape))

_get = types.SimpleNamespace(get=lambda v: v)


def install():
    """Registers the shim under the module names the reference imports."""
    keras = types.ModuleType("tensorflow.keras")
    layers = types.ModuleType("tensorflow.keras.layers")
    for cls in (Layer, InputSpec, Conv2D, Conv3D, Dense, LeakyReLU, Reshape, Lambda, LayerNormalization):
        setattr(layers, cls.__name__, cls)
    layers.UpSampling2D = layers.UpSampling3D = _UpSampling
    models = types.ModuleType("tensorflow.keras.models")
    models.Model, models.Sequential = Model, Sequential
    keras.layers, keras.models, keras.backend = layers, models, K
    keras.initializers = keras.regularizers = keras.constraints = _get
    tf.keras = keras
    mods = {"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.layers": layers, "tensorflow.keras.models": models,
            "tensorflow.keras.backend": K}
    for n in ("initializers", "regularizers", "constraints"):
        m = types.ModuleType("tensorflow.keras." + n)
        m.get = lambda v: v
        setattr(keras, n, m)
        mods["tensorflow.keras." + n] = m
    sys.modules.update(mods)
    return tf
