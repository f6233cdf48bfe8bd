"""Loader for the UNMODIFIED reference installed under baseline/_ref -- BENCHMARK / TEST INFRASTRUCTURE ONLY.

baseline/_ref is made once by (recorded in DESIGN.md section 2):
    cp -r /root/reference /tmp/refcopy
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
           --target baseline/_ref /tmp/refcopy
It is git-ignored and travels to the GPU box with the snapshot.  Nothing of it is edited.  The package's
__init__.py imports the matplotlib plotting stack, which this image lacks, so the package object is created here
without running __init__ (its submodules are the reference's own files) and three stand-ins are registered: seaborn
(three no-op setters, used at CRISPRessoCORE.py:144-148), CRISPResso2.plots.CRISPRessoPlot (setMatplotlibDefaults,
:2700) and CRISPResso2.plots.upsetplot (empty) -- the recipe of SURVEY.md Appendix C.  process_fastq and everything it
calls (get_new_variant_object, the two Cython modules, CRISPRessoMultiProcessing) are the reference's stock code.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "CRISPResso2", "CRISPRessoCORE.py"))


def load_core():
    """-> the reference's CRISPRessoCORE module (raises ImportError if baseline/_ref is absent)."""
    if not available():
        raise ImportError("baseline/_ref is not installed")
    if "CRISPResso2.CRISPRessoCORE" in sys.modules and getattr(sys.modules["CRISPResso2"], "_c2b_shim", False):
        return sys.modules["CRISPResso2.CRISPRessoCORE"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)                        # also makes the dist-info (version lookup) visible
    sb = types.ModuleType("seaborn")
    sb.set_context = sb.set = sb.set_style = lambda *a, **k: None
    sys.modules.setdefault("seaborn", sb)
    pkg = types.ModuleType("CRISPResso2")
    pkg.__path__ = [os.path.join(REF_ROOT, "CRISPResso2")]
    pkg.__file__ = os.path.join(REF_ROOT, "CRISPResso2", "__init__.py")
    pkg._c2b_shim = True
    sys.modules["CRISPResso2"] = pkg
    fake = types.ModuleType("CRISPResso2.plots.CRISPRessoPlot")
    fake.setMatplotlibDefaults = lambda *a, **k: None
    sys.modules["CRISPResso2.plots.CRISPRessoPlot"] = fake
    sys.modules["CRISPResso2.plots.upsetplot"] = types.ModuleType("CRISPResso2.plots.upsetplot")
    return importlib.import_module("CRISPResso2.CRISPRessoCORE")


def native_modules():
    """-> (CRISPResso2Align, CRISPRessoCOREResources) of the installed reference."""
    load_core()
    return (importlib.import_module("CRISPResso2.CRISPResso2Align"),
            importlib.import_module("CRISPResso2.CRISPRessoCOREResources"))
