"""Import the reference's own modules from oracle/_ref/ (built by oracle/build_ref.py)  --  TEST / BASELINE
INFRASTRUCTURE ONLY; the product package never imports this.

``load(cpu=True)`` applies the ``.cuda()`` shim of SURVEY.md 8c (the reference's CPU fall-back only catches
AssertionError) and returns the three modules; ``load(cpu=False)`` leaves ``.cuda()`` alone for the reference's own
CUDA-eager path (call ``torch.cuda.set_device`` first: masks and order matrices go to the current device).
The modules are imported under private names so that they never shadow the drop-in package's modules of the same
names.
"""
import importlib.machinery
import importlib.util
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_ORIG_CUDA = None


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, m + ".bin")) for m in ("flow_realnvp", "modules_realnvp", "utils"))


def load(cpu: bool = True):
    """Returns (flow_realnvp, modules_realnvp, utils) of the reference, or raises ImportError."""
    if not available():
        raise ImportError("oracle/_ref is not built (python oracle/build_ref.py needs /root/reference)")
    import torch
    global _ORIG_CUDA
    warnings.filterwarnings("ignore")
    if _ORIG_CUDA is None:
        _ORIG_CUDA = torch.Tensor.cuda
    if cpu:
        def _no_cuda(self, *a, **k):
            raise AssertionError("reference forced onto the CPU")
        torch.Tensor.cuda = _no_cuda
    else:
        torch.Tensor.cuda = _ORIG_CUDA
    saved = {k: sys.modules.get(k) for k in ("utils", "modules_realnvp", "flow_realnvp")}
    mods = {}
    try:
        for name in ("utils", "modules_realnvp", "flow_realnvp"):       # import order = dependency order
            loader = importlib.machinery.SourcelessFileLoader(name, os.path.join(REF_DIR, name + ".bin"))
            spec = importlib.util.spec_from_loader(name, loader)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod          # the reference modules import each other by these names
            loader.exec_module(mod)
            mods[name] = mod
    finally:
        for k, v in saved.items():           # do not leave the reference's modules under the public names
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mods["flow_realnvp"], mods["modules_realnvp"], mods["utils"]
