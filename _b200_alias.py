"""Import aliasing so the reference's module paths (`lib.network.rtpose_vgg`, `evaluate.coco_eval`, ...) resolve to
the modules of the hyphen-named package directory `pytorch_realtime_multi-person_pose_estimation_b200/` - one module
object per file, whichever name it is imported under."""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys

PKG = "pytorch_realtime_multi-person_pose_estimation_b200"
_ROOT = os.path.dirname(os.path.abspath(__file__))


def load_package():
    if PKG not in sys.modules:
        pkg_dir = os.path.join(_ROOT, PKG)
        spec = importlib.util.spec_from_file_location(PKG, os.path.join(pkg_dir, "__init__.py"),
                                                      submodule_search_locations=[pkg_dir])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[PKG] = mod
        spec.loader.exec_module(mod)
    return sys.modules[PKG]


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real):
        self.real = real

    def create_module(self, spec):
        return importlib.import_module(self.real)

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def __init__(self):
        self.aliases = {}

    def find_spec(self, fullname, path=None, target=None):
        head = fullname.split(".", 1)[0]
        if head not in self.aliases or "." not in fullname:
            return None
        real = self.aliases[head] + fullname[len(head):]
        load_package()
        if importlib.util.find_spec(real) is None:
            return None
        return importlib.machinery.ModuleSpec(fullname, _AliasLoader(real), is_package=True)


_finder = _AliasFinder()


def install(alias, subpackage):
    """Make `alias` / `alias.*` resolve to PKG.subpackage / PKG.subpackage.*; returns the real top module."""
    load_package()
    if _finder not in sys.meta_path:
        sys.meta_path.insert(0, _finder)
    _finder.aliases[alias] = PKG + "." + subpackage
    real = importlib.import_module(PKG + "." + subpackage)
    sys.modules[alias] = real
    return real
