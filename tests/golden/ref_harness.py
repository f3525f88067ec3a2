"""Import the live reference (Dont4rootMe/beast_tokenizer) in THIS container.

Only used by tests/golden/make_golden.py (fixture generation) and by CPU-only
cross-checks that skip when /root/reference is absent.  Never imported by the
product package, by `-m gpu` tests, smoke() or bench.py.

The reference imports `addict` and `matplotlib`, neither of which is installed;
two stub modules stand in.  Nothing of the reference is
copied: it is imported from where it lies.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BEAST_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "beast"))


def import_reference():
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        plt.Figure = object
        plt.Axes = object
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "addict" not in sys.modules:
        class Dict(dict):
            def __getattr__(self, k):
                if k.startswith("__"):
                    raise AttributeError(k)
                return self.setdefault(k, Dict())

            def __setattr__(self, k, v):
                self[k] = v

        m = types.ModuleType("addict")
        m.Dict = Dict
        sys.modules["addict"] = m
    for p in (os.path.join(REFERENCE_ROOT, "MP_lite_PyTorch"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    from beast.beast_bspline_tokenizer import BEASTBsplineTokenizer
    from beast.beast_bspline_bpe_tokenizer import BEASTBsplineBPETokenizer
    from beast.beast_bpe_trainer import FIGBPE, FIGBPEState
    return types.SimpleNamespace(
        BEASTBsplineTokenizer=BEASTBsplineTokenizer,
        BEASTBsplineBPETokenizer=BEASTBsplineBPETokenizer,
        FIGBPE=FIGBPE,
        FIGBPEState=FIGBPEState,
    )
