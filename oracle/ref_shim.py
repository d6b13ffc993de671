"""Import the UNMODIFIED reference (/root/reference) in this container.  TEST INFRASTRUCTURE ONLY.

Used by oracle/make_golden.py (fixture generation) and by tests that are skipped when
/root/reference is absent (it does not exist on the GPU box).  Stubs the four modules the reference
imports but never uses on this path: easydict, h5py, matplotlib, seaborn.
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _default_root():
    """DKD_REFERENCE_ROOT, else the source tree (this container), else the staged byte-identical copy that travels to
    the GPU box (baseline/_ref, oracle/install_ref.py)."""
    env = os.environ.get("DKD_REFERENCE_ROOT")
    if env:
        return env
    return "/root/reference" if os.path.isdir("/root/reference/method") else _STAGED


REF_ROOT = _default_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "method"))


class EasyDict(dict):
    """Minimal stand-in for the absent `easydict` package (attribute access, AttributeError on miss)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {})
        d.update(kw)
        for k, v in d.items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def load():
    """Returns (method.model, method.eval, method.data_provider) of the reference."""
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = EasyDict
        sys.modules["easydict"] = m
    for name in ["matplotlib", "matplotlib.pyplot", "seaborn", "h5py"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import method.model as ref_model
    import method.eval as ref_eval
    import method.data_provider as ref_data
    return ref_model, ref_eval, ref_data


def model_config(visual_input_size, query_input_size, hidden=384, n_heads=4, max_ctx_l=128, max_desc_l=30):
    """model_config of method/train.py:300-314 (+ label_style, which forward() reads at model.py:138)."""
    return EasyDict(
        visual_input_size=visual_input_size, query_input_size=query_input_size,
        inheritance_hidden=hidden, exploration_hidden=hidden, max_ctx_l=max_ctx_l, max_desc_l=max_desc_l,
        input_drop=0.2, drop=0.2, n_heads=n_heads, initializer_range=0.02, margin=0.1,
        use_hard_negative=False, hard_pool_size=20, label_style="soft")


def options(device="cpu", eval_query_bsz=50, eval_context_bsz=200):
    return EasyDict(double_branch=True, kl_intra_weight=0.1, inher_nce_weight=0.04, explore_nce_weight=0.04,
                    collection="tvr", alpha=0.8, belta=0.8, eval_context_bsz=eval_context_bsz,
                    eval_query_bsz=eval_query_bsz, num_workers=0, pin_memory=False, device=device)
