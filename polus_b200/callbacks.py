"""polus.callbacks (reference polus/callbacks.py:24-632): the hooks the training loop calls.

Pure host Python, as in the reference; kept because BaseTrainer.train drives them every step
(polus/training.py:297-338).  The loss they receive is a lazy device scalar, so only callbacks that
actually format or average it pay for a device sync.
"""
import os
import sys
from collections import OrderedDict, defaultdict
from functools import wraps
from timeit import default_timer as timer

import numpy as np

from . import _lib, hvd as _hvd, logger
from .core import get_jit_compile


def runs_if_root(method):
    """Run the hook only on local rank 0 (reference callbacks.py:24-29)."""
    @wraps(method)
    def _impl(self, *method_args, **method_kwargs):
        if _hvd().local_rank() == 0:
            return method(self, *method_args, **method_kwargs)
    return _impl


class IOutput:
    def __init__(self):
        super().__init__()
        if self.__class__.__name__ == "IOutputStream":
            raise Exception("This is an interface that cannot be instantiated")
        self.data = OrderedDict()

    def write(self, key, value):
        self.data[key] = value

    def flush(self):
        out, self.data = self.data, OrderedDict()
        return out


class ICallback:
    def __init__(self):
        super().__init__()
        if self.__class__.__name__ == "ICallback":
            raise Exception("This is an interface that cannot be instantiated")

    def on_train_begin(self):
        pass

    def on_epoch_begin(self, epoch):
        pass

    def on_train_batch_begin(self, epoch, step):
        pass

    def on_train_batch_end(self, epoch, step, loss):
        pass

    def on_epoch_end(self, epoch):
        pass

    def on_train_end(self):
        pass


class CallbackCoordinator(ICallback):
    """Fans every hook out to the registered callbacks, in order; owns the dict they share."""

    def __init__(self, callbacks, trainer, epochs, steps):
        super().__init__()
        self.callbacks = callbacks
        self.trainer = trainer
        self.epochs = epochs
        self.steps = steps
        self.shared_dict = {}
        logger.info(f"{len(self.callbacks)} callbacks were registered to be used")
        self.output_streamers = []
        for c in self.callbacks:
            c.add_coordinator(self)
            if isinstance(c, IOutput):
                self.output_streamers.append(c)

    def has_callback(self, callback_class):
        return any(isinstance(c, callback_class) for c in self.callbacks)

    def on_train_begin(self):
        for c in self.callbacks:
            c.on_train_begin()

    def on_epoch_begin(self, epoch):
        for c in self.callbacks:
            c.on_epoch_begin(epoch)

    def on_train_batch_begin(self, epoch, step):
        for c in self.callbacks:
            c.on_train_batch_begin(epoch, step)

    def on_train_batch_end(self, epoch, step, loss):
        for c in self.callbacks:
            c.on_train_batch_end(epoch, step, loss)

    def on_epoch_end(self, epoch):
        for c in self.callbacks:
            c.on_epoch_end(epoch)

    def on_train_end(self):
        for c in self.callbacks:
            c.on_train_end()


class Callback(ICallback):
    def __init__(self):
        super().__init__()
        self.coordinator = None

    def add_coordinator(self, coordinator):
        self.coordinator = coordinator


class TimerCallback(Callback):
    """Host wall-clock per batch, written to every output streamer (callbacks.py:142-156)."""

    def __init__(self):
        super().__init__()
        self.start = None

    def on_train_batch_begin(self, epoch, step):
        self.start = timer()

    def on_train_batch_end(self, epoch, step, loss):
        for output in self.coordinator.output_streamers:
            output.write("time", timer() - self.start)


class LossSmoothCallback(Callback):
    """Bias-corrected exponential moving average of the loss, beta 0.97 (callbacks.py:158-188)."""

    def __init__(self, beta=0.97, output=False):
        super().__init__()
        self.beta = beta
        self.mov_avg = 0
        self.n = 0
        self.smooth_loss = 0
        self.output = output

    def _maybe_output(self):
        if self.output:
            for output in self.coordinator.output_streamers:
                output.write("smooth loss", self.smooth_loss)

    @runs_if_root
    def on_train_batch_end(self, epoch, step, loss):
        self.n += 1
        self.mov_avg = self.beta * self.mov_avg + (1 - self.beta) * float(loss)
        self.smooth_loss = self.mov_avg / (1 - self.beta ** self.n)
        self.coordinator.shared_dict["smooth_loss"] = self.smooth_loss
        self._maybe_output()

    @runs_if_root
    def on_epoch_end(self, epoch):
        self._maybe_output()


class ValidationDataCallback(Callback):
    """Runs inference over a validation set at epoch end, gathers predictions from every rank and
    feeds the trainer's metrics on rank 0 (callbacks.py:190-261)."""

    def __init__(self, tf_validation, custom_inference_f=None, name=None, show_progress=False, validation_interval=1):
        super().__init__()
        self.tf_validation = tf_validation
        self.name = name
        self.custom_inference_f = custom_inference_f
        self.show_progress = show_progress
        self.validation_interval = validation_interval

    @runs_if_root
    def on_train_begin(self):
        shared = self.coordinator.shared_dict
        shared.setdefault("validation", {})
        if self.name is None:
            self.name = len(shared["validation"])
        shared["validation"][self.name] = {metric.name: [] for metric in self.coordinator.trainer.metrics}

    def get_metrics(self):
        return self.coordinator.shared_dict["validation"][self.name]

    def on_epoch_end(self, epoch):
        if epoch % self.validation_interval:
            return
        from .models import PolusClassifier
        hvd = _hvd()
        trainer = self.coordinator.trainer
        logger.info(f"Running validation for {self.name} set")
        for step, sample in enumerate(self.tf_validation):
            if self.show_progress:
                print(f"{step}", end="\r")
            if self.custom_inference_f is not None:
                y = self.custom_inference_f(trainer.model, sample)
            elif isinstance(sample, (list, tuple)) and len(sample) == 2:
                if isinstance(trainer.model, PolusClassifier):
                    y = trainer.model.inference(sample[0]), sample[1]
                else:
                    logger.warning("We default to just run the model over the validation data, since the models does "
                                   "not extend PolusClassifier neither a custom_inference_f was provided.")
                    y = trainer.model(sample[0]), sample[1]
            else:
                raise ValueError("Sample format outputed by the validator dataset is not supported, change to a dict "
                                 "or a two length tuple")
            all_predictions = hvd.allgather_object(y)
            if hvd.local_rank() == 0:
                for pred in all_predictions:
                    for metric in trainer.metrics:
                        metric.samples_from_batch(pred)
        if hvd.local_rank() == 0:
            results = self.coordinator.shared_dict["validation"][self.name]
            for metric in trainer.metrics:
                results[metric.name].append(metric.evaluate())
            for output in self.coordinator.output_streamers:
                output.write(f"Validation {self.name}", results)


class SaveModelCallback(Callback):
    """strategy in {"every", "best", "end"} (callbacks.py:264-313)."""

    def __init__(self, strategy, validation_name=None, metric_name=None, cache_folder=None, selection_dict_key=None):
        super().__init__()
        self.strategy = strategy
        self.validation_name = validation_name
        self.metric_name = metric_name
        self.cache_folder = cache_folder
        self.selection_dict_key = selection_dict_key
        if self.strategy not in ["every", "best", "end"]:
            logger.warning(f"The selected strategy ({strategy}) is not supported, so this callback will be ignored")
        if self.strategy == "best":
            self.best = 0

    def _save(self, **kw):
        if self.cache_folder is not None:
            kw["base_path"] = self.cache_folder
        self.coordinator.trainer.model.save(**kw)

    @runs_if_root
    def on_epoch_end(self, epoch):
        if self.strategy == "best":
            last = self.coordinator.shared_dict["validation"][self.validation_name][self.metric_name][-1]
            metric = self.selection_dict_key(last) if isinstance(last, dict) else last
            if metric > self.best:
                self.best = metric
                self._save(extension=f"_{self.validation_name}_{self.metric_name}_best")
        elif self.strategy == "every":
            self._save(extension=f"_epoch_{epoch}")

    @runs_if_root
    def on_train_end(self):
        if self.strategy == "end":
            self._save()


class EarlyStop(Callback):
    """NaN guard + patience on the (smoothed) epoch loss (callbacks.py:315-363).  As in the reference,
    `last_loss` is never updated from its initial 1000."""

    def __init__(self, patience=3, use_smooth_loss=True):
        super().__init__()
        self.current_patience = 0
        self.patience = patience
        self.last_loss = 1000
        self.use_smooth_loss = use_smooth_loss

    @runs_if_root
    def on_train_begin(self):
        if self.use_smooth_loss and not self.coordinator.has_callback(LossSmoothCallback):
            logger.warning("LossSmoothCallback was not found on the coordinator, which is a requirement to use smooth "
                           "loss. Therefore this call back will use the normal loss")
            self.use_smooth_loss = False
        self.loss = []

    @runs_if_root
    def on_train_batch_end(self, epoch, step, loss):
        if not self.use_smooth_loss:
            self.loss.append(loss)

    @runs_if_root
    def on_epoch_end(self, epoch):
        if self.use_smooth_loss:
            loss = self.coordinator.shared_dict["smooth_loss"]
        else:
            loss = sum(float(l) for l in self.loss) / max(len(self.loss), 1)
            self.loss = []
        if np.isnan(loss):
            logger.info("The training will stop early since the loss became nan")
            self.coordinator.trainer.early_stop = True
        if self.last_loss < loss:
            self.current_patience += 1
        if self.current_patience > self.patience:
            self.coordinator.trainer.early_stop = True
            logger.info(f"The training will stop early since the loss did not improve in {self.patience} consecutive epochs")


class HPOPruneCallback(Callback):
    """Reports the validation metric of every epoch to the hyper-parameter search backend and stops a pruned trial
    (reference callbacks.py:366-405).  The search driver (polus/hpo.py, optuna) is outside the training step and is
    not part of this package: without an HPO context the reference's callback does nothing but warn, and
    that is the behaviour kept here -- a script that lists it among its callbacks runs unchanged.  `hpo_backend` may
    be handed in directly (any object with report(score, step=) and should_prune(), e.g. an optuna Trial)."""

    def __init__(self, validator_name, metric_name, hpo_backend=None):
        super().__init__()
        self.hpo_backend = hpo_backend
        if self.hpo_backend is None:
            logger.warning("HPOPruneCallback was initialized however, there is no hpo context at the moment")
        self.validator_name = validator_name
        self.metric_name = metric_name

    @runs_if_root
    def on_epoch_end(self, epoch):
        if self.hpo_backend is None:
            return
        score = self.coordinator.shared_dict["validation"][self.validator_name][self.metric_name][-1]
        if not (hasattr(self.hpo_backend, "report") and hasattr(self.hpo_backend, "should_prune")):
            raise ValueError(f"The current {self.hpo_backend} backend is not supported so we do not know how to prune")
        self.hpo_backend.report(score, step=epoch)
        if self.hpo_backend.should_prune():
            message = f"Trial was pruned at epoch {epoch} with a score of {score}."
            try:
                from optuna.exceptions import TrialPruned
            except ImportError:
                self.coordinator.trainer.early_stop = True
                logger.info(message)
                return
            raise TrialPruned(message)


class Profiler(Callback):
    """Profiling window over [steps_interval[0], steps_interval[1]) global steps; stops training when
    the window closes, like the reference (callbacks.py:408-470).  Instead of tf.profiler it brackets
    the window with cudaProfilerStart/Stop, so `ncu/nsys --capture-range=cudaProfilerApi` record exactly
    those steps, and every step of the window is one NVTX range "step <n>" -- the reference's
    tf.profiler.experimental.Trace('step', step_num=step) (callbacks.py:452-461) -- bracketed by CUDA events whose
    device times are written to <logs_dir>/step_times.json when the window closes.
    Enabled by POLUS_PROFILER / POLUS_PROFILER_RANGE (training.py:279-285)."""

    def __init__(self, write_graph=True, steps_interval=[10, 20], logs_dir="logs/tensorboard_logs"):
        super().__init__()
        self.write_graph = write_graph
        self.steps_interval = steps_interval
        self.logs_dir = logs_dir
        self.trace_started = False
        self._range_open = False
        self._events = []   # (global step, start event, stop event)
        self.step_times_ms = {}

    def _event(self):
        import ctypes as C
        from . import device
        e = C.c_void_p()
        _lib.call("polus_event_create", C.byref(e))
        _lib.call("polus_event_record", e, device.stream())
        return e

    @runs_if_root
    def on_train_batch_begin(self, epoch, step):
        n = self.coordinator.trainer.step_counter
        if n >= self.steps_interval[0] and not self.trace_started:
            logger.info("Profiler - trace start!")
            _lib.call("polus_profiler_start")
            self.trace_started = True
        if self.steps_interval[0] <= n < self.steps_interval[1] and self.trace_started:
            logger.info(f"Step - {step}")
            _lib.call("polus_profiler_range_push", f"step {n}".encode())
            self._range_open = True
            self._events.append([n, self._event(), None])

    @runs_if_root
    def on_train_batch_end(self, epoch, step, loss):
        n = self.coordinator.trainer.step_counter
        if self._range_open:
            self._events[-1][2] = self._event()
            _lib.call("polus_profiler_range_pop")
            self._range_open = False
        if n >= self.steps_interval[1] - 1 and self.trace_started:
            _lib.call("polus_device_sync")
            _lib.call("polus_profiler_stop")
            self.trace_started = False
            self.coordinator.trainer.early_stop = True
            self._write_step_times()

    def _write_step_times(self):
        import ctypes as C
        import json
        for n, e0, e1 in self._events:
            if e1 is None:
                continue
            ms = C.c_float()
            _lib.call("polus_event_elapsed_ms", e0, e1, C.byref(ms))
            self.step_times_ms[int(n)] = float(ms.value)
        try:
            os.makedirs(self.logs_dir, exist_ok=True)
            with open(os.path.join(self.logs_dir, "step_times.json"), "w") as f:
                json.dump({"unit": "ms", "steps": self.step_times_ms}, f)
        except OSError as e:
            logger.warning(f"Profiler could not write step_times.json: {e}")


class WandBLogCallback(Callback, IOutput):
    """Weights & Biases logging (callbacks.py:473-558); wandb is imported lazily."""

    def __init__(self, project, init_args, entity=None, additional_info=None, model_config=None, model_name_prefix=""):
        Callback.__init__(self)
        IOutput.__init__(self)
        self.project, self.init_args, self.entity = project, init_args, entity
        self.additional_info = additional_info or {}
        self.model_config = model_config
        self.model_name_prefix = model_name_prefix
        self._run = None

    @runs_if_root
    def on_train_begin(self):
        import wandb
        cfg = dict(self.additional_info)
        if self.model_config is not None:
            cfg.update(self.model_config)
        self._run = wandb.init(project=self.project, entity=self.entity, config=cfg, **self.init_args)

    @runs_if_root
    def on_train_batch_end(self, epoch, step, loss):
        if self._run is not None:
            data = {"loss": float(loss)}
            data.update({k: v for k, v in self.flush().items() if np.isscalar(v)})
            self._run.log(data)

    @runs_if_root
    def on_epoch_end(self, epoch):
        if self._run is not None:
            self._run.log({f"epoch_{k}": v for k, v in self.flush().items() if np.isscalar(v)})

    @runs_if_root
    def on_train_end(self):
        if self._run is not None:
            self._run.finish()


class ConsoleLogCallback(Callback, IOutput):
    """Per-step progress line + per-epoch average loss (callbacks.py:560-632).  `log_interval` (new,
    default 1 = reference behaviour) prints every n-th step so the loop is not forced to sync on the
    device loss at every step."""

    def __init__(self, log_on_train_step=False, log_interval=1):
        Callback.__init__(self)
        IOutput.__init__(self)
        self.log_on_train_step = log_on_train_step
        self.log_interval = max(int(log_interval), 1)
        self.loss_per_epoch = defaultdict(list)

    def _fmt(self, d, sep=" - "):
        if isinstance(d, dict):
            parts = []
            for key, e in d.items():
                out = self._fmt(e, ", ")
                if isinstance(e, dict) or (isinstance(e, list) and e and isinstance(e[0], dict)):
                    out = f"[{out}]"
                parts.append(f"{key}: {out}")
            return sep.join(parts)
        if isinstance(d, list):
            return self._fmt(d[-1], ", ")
        return f"{d:.3f}"

    @runs_if_root
    def on_train_begin(self):
        logger.info(f"Begin training of the model \"{self.coordinator.trainer.model.name}\" for {self.coordinator.epochs} epochs")
        logger.debug(f"The training step will be build with jit_compiler={get_jit_compile()}")

    @runs_if_root
    def on_epoch_begin(self, epoch):
        logger.info(f"Begin epoch {epoch}")

    @runs_if_root
    def on_train_batch_end(self, epoch, step, loss):
        self.loss_per_epoch[epoch].append(loss)
        if step % self.log_interval:
            self.flush()
            return
        line = f"{step}/{self.coordinator.steps} - loss: {loss:.3f} - " + self._fmt(self.flush())
        if self.log_on_train_step:
            logger.info(line)
        else:
            print(line, end="\r")

    @runs_if_root
    def on_epoch_end(self, epoch):
        losses = self.loss_per_epoch[epoch]
        avg_loss = sum(float(l) for l in losses) / len(losses) if losses else 0
        logger.info(f"Average loss: {avg_loss:.3f} - " + self._fmt(self.flush()))

    @runs_if_root
    def on_train_end(self):
        logger.info("End of the training")
