"""Optimizers with the tf.keras call surface polus uses (polus/training.py:88-96,191,208-211):
`optimizer.learning_rate.read_value()/.assign()`, `optimizer.apply_gradients(zip(grads, weights))`,
`optimizer.variables()`, `optimizer.iterations`.  One fused polus_adam launch covers each contiguous
run of trainable variables in the parameter arena."""
import ctypes as C

import numpy as np

from . import _lib, device, ops
from .tensor import F32, Param, Tensor, arena


class Variable:
    """Host scalar with the tf.Variable methods polus touches (training.py:90-94)."""

    def __init__(self, value, on_change=None):
        self._v = float(value)
        self._on_change = on_change

    def read_value(self):
        return self._v

    def numpy(self):
        return self._v

    def assign(self, v):
        self._v = float(v)
        if self._on_change:
            self._on_change()

    def __float__(self):
        return self._v

    def __mul__(self, o):
        return self._v * o

    __rmul__ = __mul__

    def __repr__(self):
        return f"Variable({self._v})"


class WarmUp:
    """transformers.optimization_tf.WarmUp over PolynomialDecay(power=1) (polus/schedulers.py:5-23)."""

    def __init__(self, initial_learning_rate, warmup_steps, decay_steps, end_learning_rate=1e-7):
        self.initial_learning_rate = float(initial_learning_rate)
        self.warmup_steps = int(warmup_steps)
        self.decay_steps = int(decay_steps)
        self.end_learning_rate = float(end_learning_rate)

    def __call__(self, step):
        if step < self.warmup_steps:
            return self.initial_learning_rate * (step / self.warmup_steps)
        d = min(step - self.warmup_steps, self.decay_steps)
        return (self.initial_learning_rate - self.end_learning_rate) * (1.0 - d / self.decay_steps) + self.end_learning_rate

    # lets BaseTrainer's hvd.size() scaling treat a schedule like a variable
    def read_value(self):
        return self.initial_learning_rate

    def assign(self, v):
        self.initial_learning_rate = float(v)
        if getattr(self, "_on_change", None):
            self._on_change()


class Adam:
    """tf.keras.optimizers.Adam (epsilon 1e-7, outside the bias-corrected sqrt)."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, weight_decay_rate=0.0, name="Adam",
                 **kwargs):
        self._hyper = None          # device float[4] {lr, grad_scale, weight_decay, end_lr}: read by the kernel, so a
        self._hyper_host = None     # captured graph follows learning_rate.assign() / grad_scale changes (no re-capture)
        self._hyper_dirty = True
        self.learning_rate = learning_rate if isinstance(learning_rate, WarmUp) else Variable(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.weight_decay_rate = float(weight_decay_rate)
        self.name = name
        self.grad_scale = 1.0       # set to 1/size by the distributed tape (Horovod op=Average)
        self._state = {}            # chunk id -> (m Buffer, v Buffer, decay Buffer|None)
        self._ranges_key = None
        self._ranges = None
        self._early = []            # (chunk, off, n) spans already updated during this step's backward

    @property
    def lr(self):
        return self.learning_rate

    # ---- hyper-parameters a user may change between steps (polus/training.py:90-94 assigns learning_rate; LR-changing
    # callbacks; reuse of one optimizer across HPO trials): attribute writes mark the device copy stale
    def _touch(self):
        self._hyper_dirty = True

    @property
    def learning_rate(self):
        return self._learning_rate

    @learning_rate.setter
    def learning_rate(self, value):
        if not isinstance(value, (WarmUp, Variable)):
            value = Variable(value)
        value._on_change = self._touch
        self._learning_rate = value
        self._hyper_dirty = True

    @property
    def grad_scale(self):
        return self._grad_scale

    @grad_scale.setter
    def grad_scale(self, value):
        self._grad_scale = float(value)
        self._hyper_dirty = True

    @property
    def weight_decay_rate(self):
        return self._weight_decay_rate

    @weight_decay_rate.setter
    def weight_decay_rate(self, value):
        self._weight_decay_rate = float(value)
        self._hyper_dirty = True

    def sync_hyper(self):
        """Make the device copy of {lr, grad_scale, weight_decay, end_lr} current.  Called by the trainer before every
        step (eager or graph replay) and by _launch; a no-op unless something was assigned since the last call."""
        if self._hyper is None:
            self._hyper = device.Buffer(16, zero=True)
            self._hyper_dirty = True
        if self._hyper_dirty:
            lr = self._learning_rate
            vals = np.array([lr.initial_learning_rate if isinstance(lr, WarmUp) else float(lr), self._grad_scale,
                             self._weight_decay_rate, lr.end_learning_rate if isinstance(lr, WarmUp) else 0.0], np.float32)
            device.upload(self._hyper.ptr, vals)   # stream-ordered on the compute stream, then synchronised
            self._hyper_dirty = False
        return self._hyper.ptr

    @property
    def iterations(self):
        return int(device.download(ops.step_counter(), (1,), np.uint32)[0])

    def _chunk_state(self, ch):
        st = self._state.get(id(ch))
        if st is None:
            m = device.Buffer(ch.capacity * 4, zero=True)
            v = device.Buffer(ch.capacity * 4, zero=True)
            dm = None
            if self.weight_decay_rate > 0:
                dm = device.Buffer(ch.capacity)
                device.upload(dm.ptr, ch.host_decay)
            device.synchronize()
            st = self._state[id(ch)] = (m, v, dm, ch)
        return st

    def variables(self):
        """Optimizer slots as (ptr, nbytes) spans -- what broadcast_init_vars ships (training.py:211)."""
        out = []
        # chunk creation order: the same on every rank (the dict's insertion order follows host object ids)
        for m, v, _, ch in sorted(self._state.values(), key=lambda st: st[3].index):
            out.append(Tensor((ch.used,), F32, ptr=m.ptr, block=m))
            out.append(Tensor((ch.used,), F32, ptr=v.ptr, block=v))
        return out

    def _cfg(self):
        c = _lib.AdamCfg()
        lr = self.learning_rate
        if isinstance(lr, WarmUp):
            c.lr, c.schedule = lr.initial_learning_rate, 1
            c.warmup_steps, c.decay_steps, c.end_lr = lr.warmup_steps, lr.decay_steps, lr.end_learning_rate
        else:
            c.lr, c.schedule, c.warmup_steps, c.decay_steps, c.end_lr = float(lr), 0, 0, 1, 0.0
        c.beta1, c.beta2, c.eps = self.beta_1, self.beta_2, self.epsilon
        c.weight_decay, c.grad_scale = self.weight_decay_rate, self.grad_scale
        return c

    @staticmethod
    def _contiguous_ranges(weights):
        spans = sorted(((w.chunk.index, w.offset, (w.size + 63) & ~63, w.chunk) for w in weights if isinstance(w, Param)),
                       key=lambda s: (s[0], s[1]))
        out = []
        for cid, off, n, ch in spans:
            if out and out[-1][0] is ch and out[-1][1] + out[-1][2] == off:
                out[-1][2] += n
            elif out and out[-1][0] is ch and off < out[-1][1] + out[-1][2]:
                continue  # duplicate variable
            else:
                out.append([ch, off, n])
        return out

    def _launch(self, ch, off, n, increment, stream):
        cfg = self._cfg()
        m, v, dm, _ = self._chunk_state(ch)
        from .tensor import _pool
        if self._hyper is None or (self._hyper_dirty and _pool.trace is None):
            self.sync_hyper()   # never inside a capture: the trainer refreshes it before every step
        _lib.call("polus_adam", ch.p.ptr + off * 4, ch.g.ptr + off * 4, m.ptr + off * 4, v.ptr + off * 4,
                  ch.pb.ptr + off * 2, (dm.ptr + off) if dm is not None else None, n, C.byref(cfg), self._hyper.ptr,
                  ops.step_counter(), 1 if increment else 0, stream)

    def apply_span_early(self, ch, off, n, after=None):
        """Update one contiguous arena span while backward is still running (its gradients are final and, with `after`
        = the collective stream, reduced).  Runs on the optimizer side stream: the HBM-bound update hides under the
        tensor-bound GEMMs of the earlier layers instead of standing alone at the end of the step.  The step counter
        (bias correction, LR schedule, dropout streams) is advanced by apply_gradients, after every span."""
        if id(ch) not in self._state:
            return False  # slots are created on the first (op-by-op) step, never inside a capture
        stream = ops.opt_stream_after(after) if after is not None else ops.opt_stream_after(device.stream(), ops.side_stream_if_dirty())
        self._launch(ch, off, n, False, stream)
        self._early.append((id(ch), off, n))
        return True

    def apply_gradients(self, grads_and_vars):
        pairs = [(g, w) for g, w in grads_and_vars if g is not None]
        weights = [w for _, w in pairs]
        # The kernel reads the gradient arena.  A post_process_grads hook (polus/training.py:187-189) may hand back NEW
        # tensors (clipping, scaling through ops.mul ...): those values are what must be applied, so they are copied
        # over the arena slot first.  Tensors that are the arena views themselves (the default) cost nothing.
        for g, w in pairs:
            if isinstance(w, Param) and isinstance(g, Tensor) and g.ptr != w.grad.ptr:
                if g.size != w.size:
                    raise ValueError(f"gradient of {w.name} has shape {g.shape}, variable has {w.shape}")
                g32 = ops.cast(g, F32)
                _lib.call("polus_memcpy_d2d", w.grad.ptr, g32.ptr, w.grad.nbytes, device.stream())
        key = tuple(id(w) for w in weights)
        if key != self._ranges_key:
            self._ranges_key, self._ranges = key, self._contiguous_ranges(weights)
        ops.opt_stream_join()
        # spans the backward sweep already updated are cut out of the contiguous runs
        todo = []
        for ch, off, n in self._ranges:
            cuts = sorted((o, m) for c, o, m in self._early if c == id(ch) and o >= off and o + m <= off + n)
            pos = off
            for o, m in cuts:
                if o > pos:
                    todo.append((ch, pos, o - pos))
                pos = max(pos, o + m)
            if pos < off + n:
                todo.append((ch, pos, off + n - pos))
        self._early = []
        st = device.stream()
        if not todo:
            ch = self._ranges[0][0]
            self._launch(ch, self._ranges[0][1], 0, True, st)  # nothing left: only advance the step counter
        for i, (ch, off, n) in enumerate(todo):
            self._launch(ch, off, n, i == len(todo) - 1, st)


class AdamWeightDecay(Adam):
    """transformers.AdamWeightDecay: decoupled decay, skipped for LayerNorm / bias variables."""

    def __init__(self, learning_rate=0.001, weight_decay_rate=0.01, **kwargs):
        super().__init__(learning_rate=learning_rate, weight_decay_rate=weight_decay_rate, **kwargs)
