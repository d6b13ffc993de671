"""Training-step similarity and losses (BASELINE.json configs[4]): DLDKD.forward of the
reference (method/model.py:100-162) from the encoded vectors onward.

What runs where
  * in-batch similarity, forward AND backward — hand-written kernels (csrc/dkd_train.cu) behind one autograd
    Function: one pass of fp32 dots yields get_sim_scores (:307-329), get_unnormalized_sim_scores (:331-350) and
    the positive video's frame curve that compute_kl_loss gathers (:184-188); the (M, L, N) per-frame tensor the
    reference materialises twice per branch never exists;
  * masked-softmax KL over that curve (:190-195) — one fused forward+backward kernel instead of a Python loop of
    M log_softmax / softmax / kl_div calls;
  * triplet and (soft) NCE losses (:352-388, method/model_components.py:106-233) — vectorised PyTorch on the
    (M, N) score matrices (82 k elements at batch 128 x 5 captions); the reference's per-video Python loops are
    gone, the arithmetic and the order of torch.randint draws are kept.
The encoders stay PyTorch modules (model.py) so autograd carries the gradients from here into their parameters.
"""
import math

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


def clip_triplet_loss(scores, labels, margin, use_hard_negative, hard_pool_size):
    """get_clip_triplet_loss (method/model.py:352-388) without the per-video Python loop.  The torch.randint
    draws happen in the reference's order (one per video for v2t when not hard-negative, then one (M,) draw for
    t2v), from the default CPU generator, so a seeded run samples the same negatives."""
    M, N = scores.shape
    dev = scores.device
    lab = torch.as_tensor(np.asarray(labels), dtype=torch.long, device=dev)
    onehot = torch.zeros((M, N), dtype=torch.bool, device=dev)
    onehot[torch.arange(M, device=dev), lab] = True
    # v2t: per video, mean of its positive queries vs one negative query
    cnt = onehot.sum(dim=0)
    pos_v = torch.where(onehot, scores, torch.zeros_like(scores)).sum(dim=0) / cnt      # nan for a video with no caption, like torch.mean of an empty slice
    neg_sorted, _ = torch.sort(torch.where(onehot, torch.full_like(scores, -float("inf")), scores), dim=0,
                               descending=True)                                          # (M, N): negatives first
    if use_hard_negative:
        pick = torch.zeros((N,), dtype=torch.long, device=dev)
    else:
        n_neg = (M - cnt).tolist()
        pick = torch.cat([torch.randint(0, int(k), size=(1,)) for k in n_neg]).to(dev)
    neg_v = neg_sorted.gather(0, pick[None, :])[0]
    v2t_loss = (margin + neg_v - pos_v).clamp(min=0).sum()
    # t2v: the positive video vs one of the top-ranked other videos
    rows = torch.arange(M, device=dev)
    pos_t = scores[rows, lab]
    masked = scores.detach().clone()
    masked[rows, lab] = 999
    _, order = torch.sort(masked, descending=True, dim=1)
    max_idx = min(1 + hard_pool_size, N) if use_hard_negative else N
    draw = torch.randint(1, max_idx, size=(M,)).to(dev)
    neg_t = scores[rows, order[rows, draw]]
    t2v_loss = (margin + neg_t - pos_t).clamp(min=0)
    return t2v_loss.sum() / M + v2t_loss / N


def _label_matrix(labels, M, N, dev):
    lab = torch.as_tensor(np.asarray(labels), dtype=torch.long, device=dev)
    I = torch.zeros((M, N), device=dev)
    I[torch.arange(M, device=dev), lab] = 1
    return I, lab


def clip_nce_loss(labels, scores):
    """clip_nce (method/model_components.py:210-233), reduction 'mean'."""
    M, N = scores.shape
    I, lab = _label_matrix(labels, M, N, scores.device)
    t2v_nom = scores[torch.arange(M, device=scores.device), lab]
    t2v_den = torch.logsumexp(scores, dim=1)
    present = I.sum(dim=0) > 0
    v2t_nom = torch.logsumexp(scores.masked_fill((I == 0) & present[None, :], -float("inf")), dim=0)
    v2t_den = torch.logsumexp(scores, dim=0)
    zero = torch.zeros_like(v2t_den)
    v2t = torch.where(present, v2t_den - torch.where(present, v2t_nom, zero), zero)   # absent videos stay 0 - 0
    return torch.mean(t2v_den - t2v_nom) + torch.mean(v2t)


def clip_nce_soft_loss(labels, scores, sims, alpha, belta):
    """clip_nce_soft (method/model_components.py:106-208), reduction 'mean': hard part = first floor(alpha * bsz)
    queries / videos with one-hot targets, soft part = the rest with targets blended with softmax(sims)."""
    M, N = scores.shape
    dev = scores.device
    hardQ, hardV = math.floor(alpha * M), math.floor(alpha * N)
    softQ, softV = M - hardQ, N - hardV
    I, _ = _label_matrix(labels, M, N, dev)
    rowsel = (torch.arange(M, device=dev) >= hardQ)[:, None]
    I_Q = torch.where(rowsel, torch.clamp((1 - belta) * torch.softmax(sims, dim=-1) + belta * I, min=0), I)
    vsel = (torch.arange(N, device=dev) >= hardV)[:, None]
    I_V = torch.where(vsel, torch.clamp((1 - belta) * torch.softmax(sims.T, dim=-1) + belta * I.T, min=0), I.T)  # (N, M)
    lse_rows = torch.logsumexp(scores, dim=1, keepdim=True)
    t2v_nom_h = (I_Q[:hardQ] * scores[:hardQ]).sum()
    t2v_den_h = (I_Q[:hardQ] * lse_rows[:hardQ]).sum()
    t2v_nom_s = (I_Q[hardQ:] * scores[hardQ:]).sum()
    t2v_den_s = (I_Q[hardQ:] * lse_rows[hardQ:]).sum()
    present = I.sum(dim=0) > 0                                                   # videos in label_dict
    v_nom = torch.logsumexp(torch.log(I_V + 1e-12) + scores.T, dim=1)            # (N,)
    v_den = torch.logsumexp(scores, dim=0)
    hard_v = present & (torch.arange(N, device=dev) < hardV)
    soft_v = present & (torch.arange(N, device=dev) >= hardV)
    zero = torch.zeros_like(v_nom)
    v2t_nom_h, v2t_den_h = torch.where(hard_v, v_nom, zero).sum(), torch.where(hard_v, v_den, zero).sum()
    v2t_nom_s, v2t_den_s = torch.where(soft_v, v_nom, zero).sum(), torch.where(soft_v, v_den, zero).sum()
    hard_loss = soft_loss = 0.0
    if hardQ != 0 and hardV != 0:
        hard_loss = (t2v_den_h - t2v_nom_h) / hardQ + (v2t_den_h - v2t_nom_h) / hardV
    if softQ != 0 and softV != 0:
        soft_loss = (t2v_den_s - t2v_nom_s) / softQ + (v2t_den_s - v2t_nom_s) / softV
    return alpha * hard_loss + (1 - alpha) * soft_loss


def losses_from_encoded(model, enc, labels, mask):
    """The loss side of DLDKD.forward (method/model.py:113-157) on ENCODED vectors: enc = dict(teacher_q (M, Dt),
    teacher_ctx (N, L, Dt), inher_q / explore_q (M, D), inher_ctx / explore_ctx (N, L, D)); `model` supplies the
    loss weights and config (margin, hard negatives, label_style).  Returns (loss, dict of terms)."""
    cfg = model.config
    # teacher / inheritance / exploration scores: one fused pass each
    _, t_max_u, t_curve = in_batch_similarity(enc["teacher_q"], enc["teacher_ctx"], mask, labels)
    i_max_n, i_max_u, i_curve = in_batch_similarity(enc["inher_q"], enc["inher_ctx"], mask, labels)
    soft = cfg.label_style == "soft"
    inher_trip = clip_triplet_loss(i_max_n, labels, cfg.margin, cfg.use_hard_negative, cfg.hard_pool_size)
    if soft:
        inher_nce = model.inher_nce_weight * clip_nce_soft_loss(labels, i_max_u, t_max_u, model.alpha, model.belta)
    else:
        inher_nce = model.inher_nce_weight * clip_nce_loss(labels, i_max_u)
    explore_trip = explore_nce = 0
    if model.double_branch:
        e_max_n, e_max_u, _ = in_batch_similarity(enc["explore_q"], enc["explore_ctx"], mask, None)
        explore_trip = clip_triplet_loss(e_max_n, labels, cfg.margin, cfg.use_hard_negative, cfg.hard_pool_size)
        if soft:
            explore_nce = model.explore_nce_weight * clip_nce_soft_loss(labels, e_max_u, e_max_u, model.alpha, model.belta)
        else:
            explore_nce = model.explore_nce_weight * clip_nce_loss(labels, e_max_u)
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
