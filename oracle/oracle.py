"""CPU oracle for the DL-DKD++ corpus retrieval scoring path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker or the timed CPU baseline — never as a product path.

Two groups:

(a-REF)  restatements of code that EXISTS in the reference; every function cites the reference
         file:line it follows.  These are PINNED: tests/golden/ref_*.npz were produced by running
         the unmodified reference (/root/reference, imported through oracle/ref_shim.py) and
         tests/test_oracle_golden.py checks this file against them.

(a-NS)   the two-scale head north_star names (clip proposals + key-clip-guided frame attention).
         It does not exist under /root/reference (SURVEY §8 a-NS) — PARITY UNPINNED by the
         reference: this restatement of the MS-SL formulation is the normative spec; its only
         external anchors are the reference primitives it composes (average_to_fixed_length,
         F.normalize / einsum / torch.max conventions of get_sim_scores, the 0.7/0.3 fusion).

Plain PyTorch fp32 on CPU, written for clarity (direct formulas, no algebraic shortcuts).
"""
import numpy as np
import torch
import torch.nn.functional as F

MASK_FILL = -1e10  # method/model.py:444-445


# ----------------------------------------------------------------------------------- (a-REF)
def mask_logits(target, mask):
    """method/model.py:444-445."""
    return target * mask + (1 - mask) * MASK_FILL


def get_sim_scores(query, context, mask=None):
    """DLDKD.get_sim_scores, method/model.py:307-329.

    query (M, D), context (N, L, D), mask (N, L) float {0,1}.
    Returns (scores (M, N), per_row (M, L, N), argmax (M, N) int64); the reference computes the
    argmax at :327 and drops it.
    """
    q = F.normalize(query, dim=-1)
    c = F.normalize(context, dim=-1)
    per_row = torch.einsum("md,nld->mln", q, c)
    if mask is not None:
        per_row = mask_logits(per_row, mask.transpose(0, 1).unsqueeze(0))
    scores, idx = torch.max(per_row, dim=1)
    return scores, per_row, idx


def get_unnormalized_sim_scores(query, context, mask=None):
    """DLDKD.get_unnormalized_sim_scores, method/model.py:331-350."""
    per_row = torch.einsum("md,nld->mln", query, context)
    if mask is not None:
        per_row = mask_logits(per_row, mask.transpose(0, 1).unsqueeze(0))
    return torch.max(per_row, dim=1)[0]


def fuse_branches(inher_scores, explore_scores, w_inher=0.7, w_explore=0.3):
    """eval_epoch fusion, method/eval.py:254 (numpy fp32)."""
    return w_inher * np.asarray(inher_scores, dtype=np.float32) + w_explore * np.asarray(explore_scores, dtype=np.float32)


def get_gt(video_metas, query_metas):
    """method/eval.py:43-57 (dictionary instead of the O(Nv*Nq) double loop; same output)."""
    pos = {v: i for i, v in enumerate(video_metas)}
    v2t = [[] for _ in video_metas]
    for qi, qid in enumerate(query_metas):
        vi = pos.get(qid.split('#', 1)[0])
        if vi is not None:
            v2t[vi].append(qi)
    t2v = {}
    for vi, qs in enumerate(v2t):
        for qi in qs:
            t2v.setdefault(qi, []).append(vi)
    return v2t, t2v


def gt_ranks(neg_scores, q2m_gts, stable=True):
    """Rank (1-based) of the best ground-truth item per query, eval_q2m method/eval.py:69-82.

    The reference sorts with np.argsort (quicksort: order of exactly tied scores unspecified);
    stable=True fixes the documented tie rule "lower video index first".
    """
    n_q, n_m = neg_scores.shape
    ranks = np.zeros((n_q,), np.int32)
    for i in range(n_q):
        order = np.argsort(neg_scores[i], kind="stable" if stable else "quicksort")
        inv = np.empty(n_m, np.int64)
        inv[order] = np.arange(n_m)
        ranks[i] = min((inv[k] + 1 for k in q2m_gts[i]), default=n_m + 1)
    return ranks


def eval_q2m(neg_scores, q2m_gts):
    """method/eval.py:59-94: (r1, r5, r10, r100, medr, meanr) from NEGATED scores."""
    r = gt_ranks(neg_scores, q2m_gts)
    n_q = neg_scores.shape[0]
    rk = [100.0 * np.count_nonzero(r <= k) / n_q for k in (1, 5, 10, 100)]
    return (rk[0], rk[1], rk[2], rk[3], float(np.median(r)), float(r.mean()))


def t2v_map(neg_scores, t2v_gts):
    """method/eval.py:97-111 with ap_score :26-41: only the first GT counts => mean(1 / rank)."""
    first = {i: [g[0]] for i, g in t2v_gts.items()}
    r = gt_ranks(neg_scores, first)
    return float(np.mean(1.0 / r))


def recall_from_topk(topk_ids, q2m_gts, ks=(1, 5, 10, 100)):
    """R@K from per-query ranked id lists (what the GPU path returns)."""
    n_q = topk_ids.shape[0]
    out = []
    for k in ks:
        hit = 0
        for i in range(n_q):
            gts = set(q2m_gts[i])
            hit += any(int(v) in gts for v in topk_ids[i, :k])
        out.append(100.0 * hit / n_q)
    return tuple(out)


def average_to_fixed_length(x, map_size):
    """method/data_provider.py:30-50.  x: (n, D) torch tensor -> (map_size, D)."""
    n = x.shape[0]
    idxs = torch.arange(0, map_size + 1, 1.0) / map_size * n
    idxs = torch.min(torch.round(idxs).long(), torch.tensor(n - 1))
    out = []
    for i in range(map_size):
        s, e = idxs[i].item(), idxs[i + 1].item()
        out.append(torch.mean(x[s:e], dim=0) if s < e else x[s])
    return torch.stack(out, dim=0)


def l2_normalize_np_array(a, eps=1e-5):
    """method/data_provider.py:71-73."""
    return a / (np.linalg.norm(a, axis=-1, keepdims=True) + eps)


# ------------------------------------------------------------------------------------ (a-NS)
def num_proposals(T):
    return T * (T + 1) // 2


def proposal_index(w, s, T):
    return (w - 1) * T - ((w - 1) * (w - 2)) // 2 + s


def downsample_clips(frames, lengths, T=32):
    """N1: average_to_fixed_length on each video's valid encoded frames. (Nv, L, D) -> (Nv, T, D)."""
    return torch.stack([average_to_fixed_length(frames[n, : int(lengths[n])], T) for n in range(frames.shape[0])])


def build_proposals(clips):
    """N2: (Nv, T, D) -> (Nv, P, D): window length w = 1..T (Identity, then AvgPool1d(w, stride 1)), start s."""
    Nv, T, D = clips.shape
    x = clips.transpose(1, 2)  # (Nv, D, T)
    outs = [clips]
    for w in range(2, T + 1):
        outs.append(F.avg_pool1d(x, kernel_size=w, stride=1).transpose(1, 2))
    return torch.cat(outs, dim=1)


def clip_scale_scores(query, proposals):
    """N3: cosine of every query against every proposal, max / first argmax over proposals.

    query (M, D), proposals (Nv, P, D) -> (s_clip (M, Nv), key_clip (M, Nv) int64, all (M, P, Nv)).
    Same conventions as get_sim_scores (F.normalize eps, einsum layout, torch.max over dim 1).
    """
    return get_sim_scores(query, proposals, None)


def key_clip_guided_attention(key, val, mask, proposals, key_clip):
    """N4 (inference, all pairs): key/val (Nv, L, D) = W_k F, W_v F; mask (Nv, L); proposals
    (Nv, P, D) un-normalised means; key_clip (M, Nv).  Returns g (M, Nv, D).

    logits[m, n, l] = key[n, l] . proposals[n, key_clip[m, n]]   (no 1/sqrt(d) scaling)
    a = softmax over frames with masked frames filled with -1e10; g = sum_l a_l val[n, l].
    """
    M, Nv = key_clip.shape
    g = torch.empty(M, Nv, key.shape[-1])
    for n in range(Nv):
        pk = proposals[n, key_clip[:, n]]                      # (M, D)
        logits = pk @ key[n].T                                  # (M, L)
        logits = mask_logits(logits, mask[n].unsqueeze(0))
        a = torch.softmax(logits, dim=-1)
        g[:, n] = a @ val[n]
    return g


def frame_scale_scores(query, g):
    """N5: cos(q_m, g[m, n])."""
    return torch.einsum("md,mnd->mn", F.normalize(query, dim=-1), F.normalize(g, dim=-1))


def attention_table(key, val, mask, proposals):
    """Query-independent form of N4: attention output for EVERY proposal of every video, (Nv, P, D).
    key_clip_guided_attention(...)[m, n] == attention_table(...)[n, key_clip[m, n]]."""
    Nv, P, D = proposals.shape
    out = torch.empty(Nv, P, D)
    for n in range(Nv):
        logits = proposals[n] @ key[n].T
        logits = mask_logits(logits, mask[n].unsqueeze(0))
        out[n] = torch.softmax(logits, dim=-1) @ val[n]
    return out


def two_scale_branch(query, frames, mask, key_w, key_b, val_w, val_b, T=32, w_clip=0.7, w_frame=0.3):
    """N6 for one branch: encoded query vectors (M, D) and encoded frames (Nv, L, D) ->
    dict(clip, key_clip, frame, branch) with branch = w_clip*clip + w_frame*frame."""
    lengths = mask.sum(dim=1).long()
    clips = downsample_clips(frames, lengths, T)
    props = build_proposals(clips)
    s_clip, _, key_clip = clip_scale_scores(query, props)
    key = F.linear(frames, key_w, key_b)
    val = F.linear(frames, val_w, val_b)
    g = key_clip_guided_attention(key, val, mask, props, key_clip)
    s_frame = frame_scale_scores(query, g)
    return dict(clip=s_clip, key_clip=key_clip, frame=s_frame, branch=w_clip * s_clip + w_frame * s_frame,
                clips=clips, proposals=props)


def topk_ids(scores, K):
    """Ranked ids per query: score descending, lower id first on exact ties."""
    order = np.argsort(-np.asarray(scores), axis=1, kind="stable")
    return order[:, :K]


# ------------------------------------------------------------------- full-size parity checker (tests/, bench.py parity leg)
def two_scale_corpus(frames_by_branch, mask, params_by_branch, T=32):
    """Query-independent side of N1/N2/N4 for every branch: (proposals, key, val) lists (plain formulas above)."""
    lengths = mask.sum(dim=1).long()
    props, keys, vals = [], [], []
    for f, (kw, kb, vw, vb) in zip(frames_by_branch, params_by_branch):
        props.append(build_proposals(downsample_clips(f, lengths, T)))
        keys.append(F.linear(f, kw, kb))
        vals.append(F.linear(f, vw, vb))
    return props, keys, vals


def two_scale_eval_detail(q_by_branch, props, keys, vals, mask, bsz=50, w_clip=0.7, w_frame=0.3, tie_gap=2e-6):
    """N3-N6 + the 0.7/0.3 fusion (method/eval.py:254) for EVERY (query, video) pair, in query batches so that the
    (bsz, P, Nv) proposal-score tensor stays small.  Besides the fused matrix it returns what a parity check at the
    named shapes needs: per branch the clip-scale score, the key clip and `tie` — True where the oracle's own best and
    second-best proposal scores are closer than `tie_gap` (fp32 summation noise: either one is "the" key clip, and the
    frame-scale term follows that choice).  Returns dict(fused (M, Nv) float32, clip [b], key_clip [b], tie (M, Nv))."""
    M = q_by_branch[0].shape[0]
    nb = len(q_by_branch)
    outs = [[] for _ in range(nb)]
    clips = [[] for _ in range(nb)]
    kcs = [[] for _ in range(nb)]
    ties = []
    for lo in range(0, M, bsz):
        tie = None
        for bi, q in enumerate(q_by_branch):
            qb = q[lo: lo + bsz]
            s_clip, allp, kc = get_sim_scores(qb, props[bi], None)
            top2 = torch.topk(allp, 2, dim=1).values
            t = (top2[:, 0] - top2[:, 1]) <= tie_gap
            tie = t if tie is None else (tie | t)
            g = key_clip_guided_attention_batched(keys[bi], vals[bi], mask, props[bi], kc)
            s_frame = frame_scale_scores(qb, g)
            outs[bi].append(w_clip * s_clip + w_frame * s_frame)
            clips[bi].append(s_clip)
            kcs[bi].append(kc.to(torch.int32))
        ties.append(tie)
    sc = [torch.cat(o, dim=0).numpy() for o in outs]
    fused = sc[0] if nb == 1 else fuse_branches(sc[0], sc[1])
    return dict(fused=fused, branch=sc, clip=[torch.cat(c).numpy() for c in clips],
                key_clip=[torch.cat(k).numpy() for k in kcs], tie=torch.cat(ties).numpy())


def compare_ranking(fused_ref, tie, got_ids, K, swap_tol=1e-5):
    """Ranked top-K lists of the device (got_ids (M, K) video ids) against the oracle's dense fused scores.
    Tie pairs (see two_scale_eval_detail) are dropped from both lists (their score follows an equally valid other key
    clip); what remains must be the same sequence up to swaps of oracle scores closer than `swap_tol` (the oracle list
    is taken a few entries longer than K so that dropped tie pairs do not shorten the comparison).  Returns dict(queries_identical, swaps, tie_pairs_in_lists, mismatches) — mismatches must be 0."""
    M = fused_ref.shape[0]
    ref_ids = topk_ids(fused_ref, min(K + 8, fused_ref.shape[1]))
    identical = swaps = ties_in = mismatches = 0
    for m in range(M):
        g, r = list(got_ids[m]), list(ref_ids[m][:K])
        if g == r:
            identical += 1
            continue
        a = [v for v in g if v >= 0 and not tie[m, v]]
        b = [v for v in ref_ids[m] if not tie[m, v]]
        ties_in += len(g) - len(a)
        n = min(len(a), len(b))
        for x, y in zip(a[:n], b[:n]):
            if x != y:
                if abs(float(fused_ref[m, x]) - float(fused_ref[m, y])) <= swap_tol:
                    swaps += 1
                else:
                    mismatches += 1
    return dict(queries=M, queries_identical=identical, swaps=swaps, tie_pairs_in_lists=ties_in, mismatches=mismatches)


def recall_counts(ranks, ks=(1, 5, 10, 100)):
    """Integer R@K numerators (#queries whose best GT ranks within K): what 'R@K identical' compares."""
    r = np.asarray(ranks)
    return [int(np.count_nonzero(r <= k)) for k in ks]


# ------------------------------------------------------------------- CPU baseline (bench.py only)
def key_clip_guided_attention_batched(key, val, mask, proposals, key_clip):
    """Vectorised (bmm) form of key_clip_guided_attention for the timed CPU baseline: same math,
    no Python loop over videos (how MS-SL runs it in inference)."""
    Nv = proposals.shape[0]
    idx = key_clip.T.unsqueeze(-1).expand(-1, -1, proposals.shape[-1])          # (Nv, M, D)
    pk = torch.gather(proposals, 1, idx)                                          # (Nv, M, D)
    logits = torch.bmm(pk, key.transpose(1, 2))                                   # (Nv, M, L)
    logits = mask_logits(logits, mask.unsqueeze(1))
    a = torch.softmax(logits, dim=-1)
    return torch.bmm(a, val).transpose(0, 1)                                      # (M, Nv, D)


def cpu_eval_two_scale(q_by_branch, proposals_by_branch, key_by_branch, val_by_branch, mask, bsz=50,
                       w_clip=0.7, w_frame=0.3, K=100):
    """The two-scale eval loop the way the reference structures its own (method/eval.py:188-216): query
    batches of `bsz`, per branch the corpus-side normalisation repeated in every call
    (method/model.py:319), dense einsum + max, attention, cosine; then numpy fusion (:254) and a full
    argsort ranking per query (:75).  Returns (fused (M, Nv) float32, top-K ids)."""
    M = q_by_branch[0].shape[0]
    outs = [[] for _ in q_by_branch]
    for lo in range(0, M, bsz):
        for bi, q in enumerate(q_by_branch):
            qb = q[lo: lo + bsz]
            s_clip, _, kc = get_sim_scores(qb, proposals_by_branch[bi], None)
            g = key_clip_guided_attention_batched(key_by_branch[bi], val_by_branch[bi], mask,
                                                  proposals_by_branch[bi], kc)
            s_frame = frame_scale_scores(qb, g)
            outs[bi].append(w_clip * s_clip + w_frame * s_frame)
    sc = [torch.cat(o, dim=0).numpy() for o in outs]
    fused = sc[0] if len(sc) == 1 else fuse_branches(sc[0], sc[1])
    order = np.stack([np.argsort(-fused[i])[:K] for i in range(fused.shape[0])])
    return fused, order


def cpu_eval_frame_head(q_by_branch, frames_by_branch, mask, bsz=50, K=100):
    """The reference's shipped eval loop (frame head): method/eval.py:188-216, :254, :75."""
    M = q_by_branch[0].shape[0]
    outs = [[] for _ in q_by_branch]
    for lo in range(0, M, bsz):
        for bi, q in enumerate(q_by_branch):
            s, _, _ = get_sim_scores(q[lo: lo + bsz], frames_by_branch[bi], mask)
            outs[bi].append(s)
    sc = [torch.cat(o, dim=0).numpy() for o in outs]
    fused = sc[0] if len(sc) == 1 else fuse_branches(sc[0], sc[1])
    order = np.stack([np.argsort(-fused[i])[:K] for i in range(fused.shape[0])])
    return fused, order


# ----------------------------------------------------------------------------------- (a-REF) training step
# Restatement of the loss side of DLDKD.forward (method/model.py:100-197, :352-388) and of clip_nce /
# clip_nce_soft (method/model_components.py:106-233), from the ENCODED vectors onward, differentiable (plain
# torch autograd provides the expected gradients).  PINNED by tests/golden/ref_train_step.npz (loss terms and
# parameter gradients of the unmodified reference; oracle/make_golden.py:golden_train_step).
def kl_frame_score(predict_rows, target_rows, mask, labels, temp=0.2):
    """compute_kl_loss(mode='frame_score'), method/model.py:184-197.  *_rows (M, L, N)."""
    loss = 0
    for i, x in enumerate(labels):
        feat_len = int((mask[x] > 0).sum())
        p = F.log_softmax(predict_rows[i, :feat_len, x] / temp, dim=-1)
        t = F.softmax(target_rows[i, :feat_len, x] / temp, dim=-1)
        loss = loss + F.kl_div(p, t, reduction="sum")
    return loss


def clip_triplet_loss(scores, labels, margin, use_hard_negative, hard_pool_size):
    """get_clip_triplet_loss, method/model.py:352-388 (same order of torch.randint draws)."""
    labels = np.array(labels)
    v2t, t2v = scores.t(), scores
    v2t_loss = 0
    for i in range(v2t.shape[0]):
        pos = torch.mean(v2t[i][np.where(labels == i)[0]])
        neg, _ = torch.sort(v2t[i][np.where(labels != i)[0]], descending=True)
        pick = neg[0] if use_hard_negative else neg[torch.randint(0, neg.shape[0], size=(1,))]
        v2t_loss = v2t_loss + (margin + pick - pos).clamp(min=0).sum()
    idx = torch.arange(t2v.shape[0])
    pos = t2v[idx, labels]
    masked = t2v.detach().clone()
    masked[idx, labels] = 999
    _, order = torch.sort(masked, descending=True, dim=1)
    hi = min(1 + hard_pool_size, t2v.shape[1]) if use_hard_negative else t2v.shape[1]
    neg = t2v[idx, order[idx, torch.randint(1, hi, size=(t2v.shape[0],))]]
    return (margin + neg - pos).clamp(min=0).sum() / len(t2v) + v2t_loss / len(v2t)


def _label_dict(labels):
    d = {}
    for index, label in enumerate(labels):
        d.setdefault(label, []).append(index)
    return d


def clip_nce(labels, scores):
    """clip_nce.forward, method/model_components.py:215-233 (reduction 'mean')."""
    M, N = scores.shape
    nom = torch.logsumexp(scores[torch.arange(M), torch.as_tensor(labels)].unsqueeze(1), dim=1)
    den = torch.logsumexp(scores, dim=1)
    vn, vd = [torch.zeros(()) for _ in range(N)], [torch.zeros(()) for _ in range(N)]
    for i, qs in _label_dict(labels).items():
        vn[i] = torch.logsumexp(scores[qs, i], dim=0)
        vd[i] = torch.logsumexp(scores[:, i], dim=0)
    return torch.mean(den - nom) + torch.mean(torch.stack(vd) - torch.stack(vn))


def clip_nce_soft(labels, scores, sims, alpha, belta):
    """clip_nce_soft.forward, method/model_components.py:111-208 (reduction 'mean')."""
    import math
    M, N = scores.shape
    hardQ, hardV = math.floor(alpha * M), math.floor(alpha * N)
    softQ, softV = M - hardQ, N - hardV
    ld = _label_dict(labels)
    I = torch.zeros(M, N)
    for i, qs in ld.items():
        I[qs, i] = 1
    IQ = I.clone()
    IQ[hardQ:] = torch.clamp((1 - belta) * torch.softmax(sims, dim=-1)[hardQ:] + belta * IQ[hardQ:], min=0)
    IV = I.T.clone()
    IV[hardV:] = torch.clamp((1 - belta) * torch.softmax(sims.T, dim=-1)[hardV:] + belta * IV[hardV:], min=0)
    lse = torch.logsumexp(scores, dim=1, keepdim=True)
    t_nom_h, t_den_h = (IQ[:hardQ] * scores[:hardQ]).sum(), (IQ[:hardQ] * lse[:hardQ]).sum()
    t_nom_s, t_den_s = (IQ[hardQ:] * scores[hardQ:]).sum(), (IQ[hardQ:] * lse[hardQ:]).sum()
    v_nom_h = v_den_h = v_nom_s = v_den_s = torch.zeros(())
    for i in ld:
        nom = torch.logsumexp(torch.log(IV[i] + 1e-12) + scores[:, i], dim=0)
        den = torch.logsumexp(scores[:, i], dim=0)
        if i < hardV:
            v_nom_h, v_den_h = v_nom_h + nom, v_den_h + den
        else:
            v_nom_s, v_den_s = v_nom_s + nom, v_den_s + den
    hard = soft = 0.0
    if hardQ != 0 and hardV != 0:
        hard = (t_den_h - t_nom_h) / hardQ + (v_den_h - v_nom_h) / hardV
    if softQ != 0 and softV != 0:
        soft = (t_den_s - t_nom_s) / softQ + (v_den_s - v_nom_s) / softV
    return alpha * hard + (1 - alpha) * soft


def train_losses(enc, labels, mask, margin=0.1, use_hard_negative=True, hard_pool_size=1, label_style="soft",
                 alpha=0.8, belta=0.8, w_kl=0.1, w_inher_nce=0.04, w_explore_nce=0.04):
    """Loss side of DLDKD.forward, method/model.py:113-157.  enc: dict of ENCODED tensors — teacher_q (M, Dt),
    teacher_ctx (N, L, Dt), inher_q / explore_q (M, D), inher_ctx / explore_ctx (N, L, D).  Returns (loss, terms)."""
    _, t_rows, _ = get_sim_scores(enc["teacher_q"], enc["teacher_ctx"], mask)
    t_u = get_unnormalized_sim_scores(enc["teacher_q"], enc["teacher_ctx"], mask)
    i_max, i_rows, _ = get_sim_scores(enc["inher_q"], enc["inher_ctx"], mask)
    i_u = get_unnormalized_sim_scores(enc["inher_q"], enc["inher_ctx"], mask)
    e_max, _, _ = get_sim_scores(enc["explore_q"], enc["explore_ctx"], mask)
    e_u = get_unnormalized_sim_scores(enc["explore_q"], enc["explore_ctx"], mask)
    t = {}
    t["inher_trip"] = clip_triplet_loss(i_max, labels, margin, use_hard_negative, hard_pool_size)
    t["inher_nce"] = w_inher_nce * (clip_nce_soft(labels, i_u, t_u, alpha, belta) if label_style == "soft"
                                    else clip_nce(labels, i_u))
    t["explore_trip"] = clip_triplet_loss(e_max, labels, margin, use_hard_negative, hard_pool_size)
    t["explore_nce"] = w_explore_nce * (clip_nce_soft(labels, e_u, e_u, alpha, belta) if label_style == "soft"
                                        else clip_nce(labels, e_u))
    t["kl"] = w_kl * kl_frame_score(i_rows, t_rows, mask, labels, 0.2)
    loss = t["inher_trip"] + t["inher_nce"] + t["kl"] + t["explore_trip"] + t["explore_nce"]
    return loss, t
