"""Training-step similarity and losses (BASELINE.json configs[4]): DLDKD.forward of the
reference (method/model.py:100-162) from the encoded vectors onward.

What runs where
  * in-batch similarity, forward AND backward — hand-written kernels (csrc/dkd_train.cu) behind one autograd
    Function: one pass of fp32 dots yields get_sim_scores (:307-329), get_unnormalized_sim_scores (:331-350) and
    the positive video's frame curve that compute_kl_loss gathers (:184-188); the (M, L, N) per-frame tensor the
    reference materialises twice per branch never exists;
  * masked-softmax KL over that curve (:190-195) — one fused forward+backward kernel instead of a Python loop of
    M log_softmax / softmax / kl_div calls;
  * triplet and (soft) NCE losses (:352-388, method/model_components.py:106-233) — one fused call per branch
    (row kernel + column kernel + reduction) returning both loss terms and d loss / d scores; the reference's
    ~150 small tensor ops and per-video Python loops are gone, the order of its torch.randint draws is kept.
The encoders stay PyTorch modules (model.py) so autograd carries the gradients from here into their parameters.
"""
import numpy as np
import torch

from . import ops


class _InBatchSim(torch.autograd.Function):
    """(q (M, D), x (N, L, D), mask (N, L) uint8, labels (M) int32 | None) -> (max_n (M, N), max_u (M, N), curve (M, L))."""

    @staticmethod
    def forward(ctx, q, x, mask_u8, labels):
        q = q.contiguous().float()
        x = x.contiguous().float()
        rq = ops.row_inv_norms(q)
        rx = ops.row_inv_norms(x)
        max_n, arg_n, max_u, arg_u, curve = ops.train_sim_fwd(q, x, rq, rx, mask_u8, labels)
        ctx.save_for_backward(q, x, rq, rx, mask_u8, labels, max_n, arg_n, arg_u, curve)
        if curve is None:
            curve = q.new_zeros((q.shape[0], x.shape[1]))
            ctx.has_curve = False
        else:
            ctx.has_curve = True
        return max_n, max_u, curve

    @staticmethod
    def backward(ctx, g_n, g_u, g_c):
        q, x, rq, rx, mask_u8, labels, max_n, arg_n, arg_u, curve = ctx.saved_tensors
        want_q, want_x = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (want_q or want_x):
            return None, None, None, None

        def prep(g):
            return None if g is None else g.contiguous().float()

        gq, gx = ops.train_sim_bwd(q, x, rq, rx, mask_u8, labels, max_n, arg_n, arg_u, curve, prep(g_n), prep(g_u),
                                   prep(g_c) if ctx.has_curve else None, want_q=want_q, want_x=want_x)
        return gq, gx, None, None


def in_batch_similarity(q, x, mask, labels=None):
    """Differentiable (get_sim_scores max, get_unnormalized_sim_scores max, rows[m, :, labels[m]]).
    q (M, D), x (N, L, D) CUDA fp32; mask (N, L) {0,1}; labels: positive video index per query."""
    mask_u8 = None if mask is None else (mask > 0).to(torch.uint8).contiguous()
    lab = None
    if labels is not None:
        lab = torch.as_tensor(np.asarray(labels), dtype=torch.int32).to(q.device)
    return _InBatchSim.apply(q, x, mask_u8, lab)


class _KLCurve(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, lens, temp):
        loss, dpred = ops.kl_curve_loss(pred.contiguous().float(), target.contiguous().float(), lens, temp)
        ctx.save_for_backward(dpred)
        return loss.sum()

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return g * dpred, None, None, None


def kl_frame_loss(pred_curve, target_curve, mask, labels, temp=0.2):
    """compute_kl_loss(mode='frame_score') (method/model.py:184-197): sum over queries of the KL between the
    teacher's and the student's softmax over the valid frames of the positive video."""
    lab = torch.as_tensor(np.asarray(labels), dtype=torch.long, device=pred_curve.device)
    lens = (mask[lab] > 0).sum(dim=1).to(torch.int32).contiguous()
    return _KLCurve.apply(pred_curve, target_curve.detach(), lens, temp)


def negative_draws(labels, M, N, use_hard_negative, hard_pool_size):
    """The negative-sampling draws of get_clip_triplet_loss (method/model.py:352-388) in the reference's order and
    from the same (default CPU) generator: one torch.randint(0, n_neg) per video for the video -> text side unless
    hard negatives are on (:363-368), then one (M,) torch.randint(1, max_idx) for the text -> video side (:376-380).
    Returns (t2v_draw (M,), v2t_pick (N,)) int32 CPU tensors."""
    if use_hard_negative:
        pick = torch.zeros((N,), dtype=torch.int64)
    else:
        cnt = np.bincount(np.asarray(labels), minlength=N)
        pick = torch.cat([torch.randint(0, int(M - c), size=(1,)) for c in cnt])
    max_idx = min(1 + hard_pool_size, N) if use_hard_negative else N
    draw = torch.randint(1, max_idx, size=(M,))
    return draw.to(torch.int32), pick.to(torch.int32)


class _BranchLosses(torch.autograd.Function):
    """(s_n, s_u, sims | None, labels, draws, config) -> (triplet, nce): value and gradient in one fused pass."""

    @staticmethod
    def forward(ctx, s_n, s_u, sims, labels, draw, pick, margin, soft, alpha, belta):
        s_n, s_u = s_n.contiguous().float(), s_u.contiguous().float()
        self_distil = sims is s_u or (sims is not None and sims.data_ptr() == s_u.data_ptr())
        z = None
        if soft:
            z = s_u if self_distil else sims.contiguous().float()
        terms, g_n, g_u = ops.train_losses(s_n, s_u, z, labels, draw, pick, margin, soft, alpha, belta)
        ctx.save_for_backward(g_n, g_u)
        return terms[0], terms[1]

    @staticmethod
    def backward(ctx, g_trip, g_nce):
        g_n, g_u = ctx.saved_tensors
        return g_trip * g_n, g_nce * g_u, None, None, None, None, None, None, None, None


def branch_losses(s_n, s_u, sims, labels, margin, use_hard_negative, hard_pool_size, soft, alpha, belta):
    """Triplet loss on the cosine maxima + (soft) NCE loss on the raw maxima of one branch.  sims: the soft-target
    source of clip_nce_soft — the teacher's raw maxima (no gradient), or s_u itself for the self-distilled
    exploration branch (method/model.py:146-150); ignored when soft is False."""
    M, N = s_n.shape
    if soft and sims is not s_u and sims.requires_grad:
        raise ValueError("branch_losses: soft targets from a separate tensor must not require grad")
    draw, pick = negative_draws(labels, M, N, use_hard_negative, hard_pool_size)
    lab = torch.as_tensor(np.asarray(labels), dtype=torch.int32)
    dev = s_n.device
    return _BranchLosses.apply(s_n, s_u, sims if soft else None, lab.to(dev), draw.to(dev), pick.to(dev), margin, soft,
                               alpha, belta)


def losses_from_encoded(model, enc, labels, mask):
    """The loss side of DLDKD.forward (method/model.py:113-157) on ENCODED vectors: enc = dict(teacher_q (M, Dt),
    teacher_ctx (N, L, Dt), inher_q / explore_q (M, D), inher_ctx / explore_ctx (N, L, D)); `model` supplies the
    loss weights and config (margin, hard negatives, label_style).  Returns (loss, dict of terms)."""
    cfg = model.config
    # teacher / inheritance / exploration scores: one fused pass each
    _, t_max_u, t_curve = in_batch_similarity(enc["teacher_q"], enc["teacher_ctx"], mask, labels)
    i_max_n, i_max_u, i_curve = in_batch_similarity(enc["inher_q"], enc["inher_ctx"], mask, labels)
    soft = cfg.label_style == "soft"
    hard, pool = cfg.use_hard_negative, cfg.hard_pool_size
    inher_trip, nce = branch_losses(i_max_n, i_max_u, t_max_u, labels, cfg.margin, hard, pool, soft, model.alpha,
                                    model.belta)
    inher_nce = model.inher_nce_weight * nce
    explore_trip = explore_nce = 0
    if model.double_branch:
        e_max_n, e_max_u, _ = in_batch_similarity(enc["explore_q"], enc["explore_ctx"], mask, None)
        explore_trip, nce = branch_losses(e_max_n, e_max_u, e_max_u, labels, cfg.margin, hard, pool, soft,
                                          model.alpha, model.belta)
        explore_nce = model.explore_nce_weight * nce
    kl_intra = model.kl_intra_weight * model.weight * kl_frame_loss(i_curve, t_curve, mask, labels, 0.2)
    kl = kl_intra
    loss = inher_trip + inher_nce + kl + explore_trip + explore_nce
    return loss, {"loss_overall": float(loss), "inher_trip": inher_trip, "inher_nce": inher_nce,
                  "explore_trip": explore_trip, "explore_nce": explore_nce, "kl": kl, "kl_intra": kl_intra}


def forward_losses(model, batch):
    """DLDKD.forward (method/model.py:100-162): same batch keys, same return value (loss, dict of terms)."""
    mask = batch["student_videos_mask"]
    inher_ctx, explore_ctx = model.encode_context(batch["student_videos"], mask)
    inher_q, explore_q = model.encode_query(batch["student_text"], batch["student_text_mask"])
    enc = dict(teacher_q=batch["teacher_text"].squeeze(), teacher_ctx=batch["teacher_videos"], inher_q=inher_q,
               inher_ctx=inher_ctx, explore_q=explore_q, explore_ctx=explore_ctx)
    return losses_from_encoded(model, enc, batch["text_labels"], mask)
