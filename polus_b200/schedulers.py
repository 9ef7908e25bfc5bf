"""polus.schedulers (reference polus/schedulers.py:5-23)."""
from .optimizers import WarmUp


def warmup_scheduler(num_train_steps, max_lr, warmup_percentage=0.1, end_lr=1e-7):
    """Linear warm-up over int(N*pct) steps, then linear decay to 1e-7.  As in the reference the
    `end_lr` argument is accepted but NOT used (schedulers.py:15 hard-codes 1e-7)."""
    num_warmup_steps = int(num_train_steps * warmup_percentage)
    return WarmUp(initial_learning_rate=max_lr, warmup_steps=num_warmup_steps,
                  decay_steps=num_train_steps - num_warmup_steps, end_learning_rate=1e-7)
