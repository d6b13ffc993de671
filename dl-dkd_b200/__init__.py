"""dkd_b200 — B200-native corpus retrieval scoring path of DL-DKD++ (HuiGuanLab/DL-DKD).

The sources live in `dl-dkd_b200/`; the import name is `dkd_b200` (the `dkd_b200/` shim at the repository root
points Python at this directory: a hyphen is not importable directly).

Public surface (drop-in names of the reference):
    model.DLDKD                      encode_context / encode_query / get_sim_scores /
                                     get_pred_from_raw_query / key_clip_guided_attention
    eval.compute_context_info, eval.compute_query2ctx_info, eval.eval_epoch, eval.eval_q2m
    ops.*                            tensor-level wrappers over the C ABI (include/dkd_b200.h)
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "ops", "build"]
