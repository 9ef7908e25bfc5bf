"""`polus` import alias of the B200-native package: `from polus.training import ClassifierTrainer`,
`from polus.ner.models import ...`, `import polus.callbacks` ... resolve to the same-named modules of `polus_b200`, so a
script written against bioinformatics-ua/polus runs unchanged (the module map follows the reference package:
training, models, layers, losses, callbacks, data, metrics, core, utils, schedulers, ner.*, ir.*, mock.horovod).

A meta-path finder redirects every `polus.X` import to `polus_b200.X` and registers the SAME module object under both
names (no second copy of any module state: one parameter arena, one PolusContext)."""
import importlib
import importlib.abc
import importlib.util
import sys

import polus_b200 as _impl

_PREFIX, _TARGET = __name__ + ".", _impl.__name__ + "."


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real_name):
        self.real_name = real_name

    def create_module(self, spec):
        mod = importlib.import_module(self.real_name)
        return mod

    def exec_module(self, module):   # already executed under its real name
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_PREFIX):
            return None
        real = _TARGET + fullname[len(_PREFIX):]
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except (ImportError, ValueError):
            return None
        return importlib.util.spec_from_loader(fullname, _AliasLoader(real))


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())

# everything the implementation package exports at top level (PolusContext, logger, hvd, __version__ ...)
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__") or k == "__version__"})
__path__ = []   # a package: submodule imports go through the finder above
