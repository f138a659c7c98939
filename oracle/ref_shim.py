"""Import shim for the UNMODIFIED reference (SunJiamei/LRP-imagecaptioning-pytorch).

TEST INFRASTRUCTURE ONLY.  This file is used in the build container (where
/root/reference is mounted) to (a) validate the oracle restatement in
``oracle/lrp_oracle.py`` and (b) generate the golden vectors under
``tests/golden/`` (see ``oracle/make_golden.py``).  It is never imported by the
product path and never runs on the GPU box (the reference does not travel).

What has to be stubbed to import the reference on a CPU-only, offline box
(SURVEY.md Appendix A):
  * matplotlib / skimage / nltk are not installed        -> empty stub modules
  * torchvision.models.utils no longer exists            -> stub with load_state_dict_from_url
  * vgg16(pretrained=True) wants the network             -> force pretrained=False
  * hard-coded .cuda() calls                             -> identity on CPU
  * gridTDmodel.py:444  ``beam_idx = top_words / vocab_size`` (true division on
    torch>=1.6 -> IndexError)                            -> text patch to ``//`` (Q8)
  * lrp_modules.py:14 in-place masked_fill_ on a saved input (Q9)
                                                         -> ``.clone()`` first (only for Linear hooks)
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

def _find_reference():
    """/root/reference in the build container; on the GPU box the copy oracle/stage_reference.py left under
    baseline/_ref (git-ignored, travels with the gpurun snapshot)."""
    env = os.environ.get("LRPX_REFERENCE_ROOT")
    if env:
        return env
    staged = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    for cand in ("/root/reference", staged):
        if os.path.isdir(os.path.join(cand, "LRPtools")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_reference()

# A fixed English stop-word subset (nltk is not installed); the
# synthetic vocabulary used by the tests draws stop words from here.
STOP_WORDS = ["a", "an", "the", "of", "on", "in", "with", "and", "is", "are", "to", "at"]


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "LRPtools"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        plt = _stub("matplotlib.pyplot", imshow=lambda *a, **k: None, show=lambda *a, **k: None)   # display only
        mpl.pyplot = plt
    if "skimage" not in sys.modules:
        sk = _stub("skimage")
        skt = _stub("skimage.transform")
        sk.transform = skt
    if "nltk" not in sys.modules:
        nl = _stub("nltk")
        corpus = _stub("nltk.corpus")

        class _SW:
            @staticmethod
            def words(lang):
                return list(STOP_WORDS)

        corpus.stopwords = _SW
        nl.corpus = corpus
    import torchvision.models as tvm

    if "torchvision.models.utils" not in sys.modules:
        u = _stub("torchvision.models.utils", load_state_dict_from_url=lambda *a, **k: {})
        tvm.utils = u


_loaded = {}


def load_reference_evaluation():
    """evaluation.py of the reference (EvaluationExperiments: block_image, ablation loops), imported with the reference's
    own models / LRPtools visible under their public names for the duration of the import."""
    if "evaluation" in _loaded:
        return _loaded["evaluation"]
    load_reference()
    saved = {k: sys.modules[k] for k in list(sys.modules) if k.split(".")[0] in ("LRPtools", "models", "config")}
    for k in saved:
        del sys.modules[k]
    for k in [k for k in sys.modules if k.startswith("_lrpx_ref_.")]:
        sys.modules[k[len("_lrpx_ref_."):]] = sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        path = os.path.join(REFERENCE_ROOT, "evaluation.py")
        mod = types.ModuleType("_lrpx_ref_evaluation")
        mod.__file__ = path
        exec(compile(open(path).read(), path, "exec"), mod.__dict__)
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k.split(".")[0] in ("LRPtools", "models", "config")]:
            del sys.modules[k]
        sys.modules.update(saved)
    _loaded["evaluation"] = mod
    return mod


def load_reference(cpu: bool = True):
    """Returns a namespace with the reference modules (LRPtools.*, models.*)."""
    if "ns" in _loaded:
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    install_stubs()
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
    # Make sure the reference's own top-level packages win over the repo's mirrors.
    # (the product's mirrors are regular packages and would shadow the reference's namespace packages whatever the
    # path order: take every other tree that carries an LRPtools package off sys.path while importing; the mirrors
    # already imported are put back under their names afterwards)
    mirrors = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("LRPtools", "models")}
    saved_path = list(sys.path)
    sys.path[:] = [REFERENCE_ROOT] + [q for q in saved_path
                                      if not os.path.isdir(os.path.join(q or ".", "LRPtools")) or
                                      os.path.abspath(q or ".") == os.path.abspath(REFERENCE_ROOT)]
    try:
        import LRPtools.utils as ref_utils  # noqa
        import models.vgg as ref_vgg
        import models.resnet as ref_resnet

        # no network: never try to fetch pretrained weights
        _orig_vgg = ref_vgg._vgg
        ref_vgg._vgg = lambda arch, cfg, bn, pretrained, progress, **kw: _orig_vgg(arch, cfg, bn, False, progress, **kw)
        _orig_r101 = ref_resnet.resnet101
        ref_resnet.resnet101 = lambda pretrained=False, **kw: _orig_r101(pretrained=False, **kw)

        # lrp_modules with the Q9 patch (clone before in-place fill)
        src = open(os.path.join(REFERENCE_ROOT, "LRPtools/lrp_modules.py")).read()
        src = src.replace("input_ = module.input[0]\n        input_.masked_fill_",
                          "input_ = module.input[0].clone()\n        input_.masked_fill_")
        spec = importlib.util.spec_from_loader("LRPtools.lrp_modules", loader=None)
        ref_lrp_modules = importlib.util.module_from_spec(spec)
        ref_lrp_modules.__file__ = os.path.join(REFERENCE_ROOT, "LRPtools/lrp_modules.py")
        sys.modules["LRPtools.lrp_modules"] = ref_lrp_modules
        exec(compile(src, ref_lrp_modules.__file__, "exec"), ref_lrp_modules.__dict__)
        import LRPtools
        LRPtools.lrp_modules = ref_lrp_modules
        import LRPtools.lrp_wrapper as ref_lrp_wrapper

        def _load_patched(modname, relpath):
            src = open(os.path.join(REFERENCE_ROOT, relpath)).read()
            src = src.replace("beam_idx = top_words / vocab_size", "beam_idx = top_words // vocab_size")
            spec = importlib.util.spec_from_loader(modname, loader=None)
            mod = importlib.util.module_from_spec(spec)
            mod.__file__ = os.path.join(REFERENCE_ROOT, relpath)
            sys.modules[modname] = mod
            exec(compile(src, mod.__file__, "exec"), mod.__dict__)
            return mod

        ref_grid = _load_patched("models.gridTDmodel", "models/gridTDmodel.py")
        ref_aoa = _load_patched("models.aoamodel", "models/aoamodel.py")
        ref_ada = _load_patched("models.adaptiveattention", "models/adaptiveattention.py")
    finally:
        sys.path[:] = saved_path
    ns = types.SimpleNamespace(utils=ref_utils, lrp_modules=ref_lrp_modules, lrp_wrapper=ref_lrp_wrapper,
                               vgg=ref_vgg, resnet=ref_resnet, gridTDmodel=ref_grid, aoamodel=ref_aoa,
                               adaptiveattention=ref_ada)
    _loaded["ns"] = ns
    # leave the reference's modules registered under private names only
    for k in [k for k in sys.modules if k.split(".")[0] in ("LRPtools", "models")]:
        sys.modules["_lrpx_ref_." + k] = sys.modules.pop(k)
    sys.modules.update(mirrors)
    return ns
