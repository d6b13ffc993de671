"""Drop-in parity on the GPU: the reference-facing entry points (model + eval mirror) against fixtures
produced by the unmodified reference, and the engine's heads/precisions against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import ref_shim
from oracle.datasets import QuerySet, TeacherQuerySet, TeacherVideoSet, VideoSet
from tests import synth
from tests.test_oracle_golden import _load, _tiny_model

pytestmark = pytest.mark.gpu


def _tiny_setup(dkd):
    g = _load("ref_tiny_eval.npz")
    m = _tiny_model(dkd, g).cuda()
    Nv, Nq = int(g["dims"][5]), int(g["dims"][6])
    vids = [torch.from_numpy(g["videos"][i, : g["video_len"][i]]) for i in range(Nv)]
    qs = [torch.from_numpy(g["queries"][i, : g["query_len"][i]]) for i in range(Nq)]
    opt = ref_shim.options(device=torch.device("cuda"), eval_query_bsz=5, eval_context_bsz=4)
    return g, m, VideoSet(vids), QuerySet(qs, Nv), opt


def test_eval_entry_points_match_reference_fixture(ops, dkd):
    """compute_context_info + compute_query2ctx_info + eval_q2m/t2v_map/eval_epoch on the reference's own
    weights and data reproduce the reference's outputs (tests/golden/ref_tiny_eval.npz)."""
    from dkd_b200 import eval as E
    g, m, vset, qset, opt = _tiny_setup(dkd)
    ctx = E.compute_context_info(m, vset, opt)
    assert ctx["video_metas"] == vset.ids and ctx["teacher_frame_feat"] is None
    assert np.array_equal(ctx["video_mask"].cpu().numpy(), g["video_mask"])
    valid = g["video_mask"] > 0
    assert np.allclose(ctx["inher_frame_feat"].cpu().numpy()[valid], g["inher_frame_feat"][valid], atol=5e-6)
    inher, explore, teacher, metas = E.compute_query2ctx_info(m, qset, opt, ctx)
    assert teacher is None
    assert [qset.ids.index(x) for x in metas] == list(g["order"])      # per-batch length sort reproduced
    assert inher.dtype == np.float32 and inher.shape == g["inher_scores"].shape
    assert np.abs(inher - g["inher_scores"]).max() <= 5e-6
    assert np.abs(explore - g["explore_scores"]).max() <= 5e-6
    _, t2v = E.get_gt(ctx["video_metas"], metas)
    fused = O.fuse_branches(inher, explore)
    got = E.eval_q2m(-1 * fused, t2v)
    assert np.allclose(np.array(got, dtype=np.float64), g["metrics_fused"])
    assert np.isclose(E.t2v_map(-1 * fused, t2v), float(g["map_fused"]))
    rsum = E.eval_epoch(m, vset, qset, opt)
    assert np.isclose(rsum, g["metrics_fused"][:4].sum())


def test_get_sim_scores_dropin_matches_fixture(ops, dkd):
    from dkd_b200.model import DLDKD
    g = _load("ref_sim_scores.npz")
    q, ctx, mask = (torch.from_numpy(g[k]).cuda() for k in ("q", "ctx", "mask"))
    s, rows = DLDKD.get_sim_scores(q, ctx, mask)
    assert np.abs(s.cpu().numpy() - g["scores"]).max() <= 2e-6
    valid = (g["mask"].T[None] > 0) & np.ones_like(g["rows"], dtype=bool)
    assert np.abs(rows.cpu().numpy()[valid] - g["rows"][valid]).max() <= 2e-6
    assert (rows.cpu().numpy()[~valid] == np.float32(-1e10)).all()
    s2, _ = DLDKD.get_sim_scores(q, ctx, None)
    assert np.abs(s2.cpu().numpy() - g["scores_nomask"]).max() <= 2e-6


def test_two_scale_entry_points(ops, dkd):
    """get_pred_from_raw_query / key_clip_guided_attention against the oracle's two-scale head
    (parity unpinned by the reference: the oracle restatement is the spec)."""
    from dkd_b200 import eval as E
    g, m, vset, qset, opt = _tiny_setup(dkd)
    opt.scoring = "two_scale"
    ctx = E.compute_context_info(m, vset, opt)
    pc = ctx["prepared"]
    qf = torch.from_numpy(g["queries"]).cuda()
    qmask = (torch.arange(qf.shape[1])[None] < torch.from_numpy(g["query_len"])[:, None]).float().cuda()
    clip, frame, fused = m.get_pred_from_raw_query(qf, qmask, video_proposal_feat=pc, return_fused=True)
    with torch.no_grad():
        qi, qe = m.encode_query(qf, qmask)
    mask = ctx["video_mask"].cpu()
    refs = []
    for bi, (qq, ff) in enumerate(((qi, ctx["inher_frame_feat"]), (qe, ctx["explore_frame_feat"]))):
        kw, kb, vw, vb = (t.detach().cpu() for t in m.attention_params()[bi])
        refs.append(O.two_scale_branch(qq.cpu(), ff.cpu(), mask, kw, kb, vw, vb, T=m.map_size))
    for bi in range(2):
        assert (clip[bi].cpu() - refs[bi]["clip"]).abs().max() <= 2e-6
        assert (frame[bi].cpu() - refs[bi]["frame"]).abs().max() <= 1e-5
    fused_ref = O.fuse_branches(refs[0]["branch"].numpy(), refs[1]["branch"].numpy())
    assert np.abs(fused.cpu().numpy() - fused_ref).max() <= 5e-6
    # attention entry points
    kc = refs[0]["key_clip"].cuda()
    gk = m.key_clip_guided_attention_in_inference(ctx["inher_frame_feat"], None, ctx["video_mask"], kc, branch=0)
    kw, kb, vw, vb = (t.detach().cpu() for t in m.attention_params()[0])
    fr = ctx["inher_frame_feat"].cpu()
    g_ref = O.key_clip_guided_attention(torch.nn.functional.linear(fr, kw, kb), torch.nn.functional.linear(fr, vw, vb),
                                        mask, refs[0]["proposals"], refs[0]["key_clip"])
    g_ref = torch.nn.functional.normalize(g_ref, dim=-1)
    assert (gk.cpu() - g_ref).abs().max() <= 1e-5
    labels = [i % len(vset) for i in range(len(qset))]
    mi = torch.tensor([int(refs[0]["key_clip"][i, labels[i]]) for i in range(len(qset))])
    gt = m.key_clip_guided_attention(ctx["inher_frame_feat"], None, ctx["video_mask"], mi.cuda(), labels, branch=0)
    assert (gt.cpu() - g_ref[torch.arange(len(qset)), torch.tensor(labels)]).abs().max() <= 1e-5
    # per-branch dense two-scale scores through the reference-named eval entry point
    inher, explore, _, metas = E.compute_query2ctx_info(m, qset, opt, ctx)
    order = [qset.ids.index(x) for x in metas]
    assert np.abs(inher - refs[0]["branch"].numpy()[order]).max() <= 5e-6
    assert np.abs(explore - refs[1]["branch"].numpy()[order]).max() <= 5e-6


@pytest.mark.parametrize("head", ["frame", "two_scale"])
def test_bf16_rank_equals_exact_rank(ops, head):
    """Top-100 from the tcgen05 path + fp32 rescoring == top-100 of the exact fp32 path (ids AND scores),
    and both equal the oracle's ranking up to fp32-noise ties."""
    from dkd_b200 import engine
    Nv, L, D, M, K = 700, 128, 384, 300, 100
    frames, mask, lengths = synth.encoded_corpus(Nv, L, D, seed=7, shared=1.5)
    frames2, _, _ = synth.encoded_corpus(Nv, L, D, seed=8, shared=1.5)
    frames2 = frames2 * mask[:, :, None]
    gen = torch.Generator().manual_seed(9)
    params = [(0.05 * torch.randn(D, D, generator=gen), torch.zeros(D), 0.05 * torch.randn(D, D, generator=gen),
               torch.zeros(D)) for _ in range(2)]
    qs = [synth.encoded_queries(M, D, seed=10), synth.encoded_queries(M, D, seed=11)]
    dev = torch.device("cuda")
    pc = engine.prepare_corpus([frames.to(dev), frames2.to(dev)], mask.to(dev),
                               [tuple(t.to(dev) for t in p) for p in params], heads=(head,))
    pq = engine.prepare_queries([q.to(dev) for q in qs])
    s_ex, i_ex = engine.rank(pc, pq, K=K, head=head, precision="exact")
    s_bf, i_bf = engine.rank(pc, pq, K=K, head=head, precision="bf16", Kc=128)
    # certified candidates (engine.rank certify=True): identical to the exact path for EVERY query, both heads
    assert torch.equal(i_bf, i_ex)
    assert torch.equal(s_bf, s_ex)
    # even with no candidate margin at all (Kc = K) the certificate + exact fallback restores equality
    engine.STATS["certify_fallback_queries"] = 0
    s_t, i_t = engine.rank(pc, pq, K=K, head=head, precision="bf16", Kc=K)
    assert torch.equal(i_t, i_ex) and torch.equal(s_t, s_ex)
    assert engine.STATS["certify_fallback_queries"] > 0
    if head == "frame":
        sc = [O.get_sim_scores(q, f, mask)[0].numpy() for q, f in zip(qs, (frames, frames2))]
        fused_ref = O.fuse_branches(sc[0], sc[1])
        top_ref = O.topk_ids(fused_ref, K)
        got = i_ex.cpu().numpy()
        ref_sorted = np.take_along_axis(fused_ref, top_ref, 1)
        assert np.abs(s_ex.cpu().numpy() - ref_sorted).max() <= 2e-6
        diff = got != top_ref
        # any disagreement must be a swap of scores closer than fp32 summation noise
        if diff.any():
            r, c = np.nonzero(diff)
            assert np.abs(fused_ref[r, got[r, c]] - fused_ref[r, top_ref[r, c]]).max() <= 2e-6


def test_sharded_rank_equals_unsharded(ops):
    """Video-sharded prepare + local top-K + merge == single-corpus top-K (the multi-GPU algebra on one GPU)."""
    from dkd_b200 import engine
    Nv, L, D, M, K, G = 301, 64, 128, 90, 100, 4
    frames, mask, _ = synth.encoded_corpus(Nv, L, D, seed=21)
    qs = [synth.encoded_queries(M, D, seed=22)]
    dev = torch.device("cuda")
    pq = engine.prepare_queries([q.to(dev) for q in qs])
    pc = engine.prepare_corpus([frames.to(dev)], mask.to(dev), heads=("frame",))
    s_all, i_all = engine.rank(pc, pq, K=K, head="frame", precision="bf16")
    ls, li = [], []
    for r in range(G):
        lo, hi = engine.shard_range(Nv, r, G)
        pcs = engine.prepare_corpus([frames[lo:hi].to(dev)], mask[lo:hi].to(dev), heads=("frame",), id_base=lo)
        s, i = engine.rank(pcs, pq, K=K, head="frame", precision="bf16")
        ls.append(s)
        li.append(i)
    ms, mi = ops.merge_topk(torch.stack(ls), torch.stack(li))
    assert torch.equal(mi, i_all) and torch.equal(ms, s_all)


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_streamed_rank_equals_resident_rank(ops):
    """engine.rank_streamed over ragged chunks (last chunk short, one chunk smaller than K) and ragged query
    batches == engine.rank over the whole corpus, exact precision, both heads' worth of branches."""
    from dkd_b200 import engine
    Nv, L, D, M, K = 333, 64, 128, 150, 100
    frames, mask, _ = synth.encoded_corpus(Nv, L, D, seed=41)
    frames2, _, _ = synth.encoded_corpus(Nv, L, D, seed=42)
    frames2 = frames2 * mask[:, :, None]
    gen = torch.Generator().manual_seed(43)
    params = [(0.05 * torch.randn(D, D, generator=gen), torch.zeros(D), 0.05 * torch.randn(D, D, generator=gen),
               torch.zeros(D)) for _ in range(2)]
    dev = torch.device("cuda")
    fr = [frames.to(dev), frames2.to(dev)]
    params = [tuple(t.to(dev) for t in p) for p in params]
    qs = [synth.encoded_queries(M, D, seed=44).to(dev), synth.encoded_queries(M, D, seed=45).to(dev)]
    pc = engine.prepare_corpus(fr, mask.to(dev), params, heads=("two_scale",), precisions=("exact",))
    s_all, i_all = engine.rank(pc, engine.prepare_queries(qs), K=K, head="two_scale", precision="exact")
    pqs = engine.split_queries(qs, 64)
    assert [p.M for p in pqs] == [64, 64, 22]
    s_st, i_st = engine.rank_streamed(engine.iter_chunks(fr, mask.to(dev), 128, id_base=0), pqs, params, K=K,
                                      precision="exact")      # chunks of 128, 128, 77 (< K) videos
    assert torch.equal(i_st, i_all) and torch.equal(s_st, s_all)
    with pytest.raises(ValueError):
        engine.rank_streamed(iter(()), pqs, params)


@pytest.mark.parametrize("scoring", ["frame", "two_scale"])
def test_eval_epoch_hot_path_equals_reference_flow(ops, dkd, scoring):
    """eval_epoch(opt.precision='bf16') — engine.rank on the device, no dense device->host copy — returns the same
    R@1+R@5+R@10+R@100 as the reference flow (opt.precision='exact': dense matrices, eval_q2m), and rank_queries'
    ranked lists agree with the ranking of the dense exact scores."""
    from dkd_b200 import eval as E
    g, m, vset, qset, opt = _tiny_setup(dkd)
    opt.scoring = scoring
    opt.precision = "exact"
    rsum_ref = E.eval_epoch(m, vset, qset, opt)
    ctx = E.compute_context_info(m, vset, opt)
    inher, explore, _, metas = E.compute_query2ctx_info(m, qset, opt, ctx)
    fused = O.fuse_branches(inher, explore)
    _, t2v = E.get_gt(ctx["video_metas"], metas)
    ref_metrics = E.eval_q2m(-1 * fused, t2v)
    ref_map = E.t2v_map(-1 * fused, t2v)
    opt.precision = "bf16"
    rsum = E.eval_epoch(m, vset, qset, opt)
    assert np.isclose(rsum, rsum_ref)
    ctx = E.compute_context_info(m, vset, opt)
    s, ids, metas2, dense = E.rank_queries(m, qset, opt, ctx, K=100, return_dense=True)
    assert metas2 == metas and ids.shape[0] == len(qset) and dense.shape == fused.shape
    Nv = len(vset)
    assert np.array_equal(ids.cpu().numpy()[:, :Nv], O.topk_ids(fused, Nv))            # corpus smaller than K
    assert np.abs(s.cpu().numpy()[:, :Nv] - np.take_along_axis(fused, O.topk_ids(fused, Nv), 1)).max() <= 5e-6
    got = E.metrics_from_ranking(ids, dense, t2v)
    assert np.allclose(got[:6], ref_metrics) and np.isclose(got[6], ref_map)
    if scoring == "frame":      # the reference's own numbers for this fixture
        assert np.isclose(rsum, g["metrics_fused"][:4].sum())


def test_get_sim_scores_fast_route_and_grad_guard(ops, dkd):
    from dkd_b200.model import DLDKD
    g = _load("ref_sim_scores.npz")
    q, ctx, mask = (torch.from_numpy(g[k]).cuda() for k in ("q", "ctx", "mask"))
    s, rows = DLDKD.get_sim_scores(q, ctx, mask, want_rows=False)          # tcgen05 kind::tf32 x 3 route
    assert rows is None and np.abs(s.cpu().numpy() - g["scores"]).max() <= 2e-6
    with pytest.raises(RuntimeError, match="inference-only"):
        DLDKD.get_sim_scores(q.clone().requires_grad_(True), ctx, mask)
    with torch.no_grad():
        DLDKD.get_sim_scores(q.clone().requires_grad_(True), ctx, mask)


def test_teacher_items_are_scored(ops, dkd):
    """Datasets whose items carry teacher (CLIP) features: the 4-field item layout of the reference's collate
    functions; teacher scores = get_sim_scores on the raw teacher features (method/eval.py:198-202)."""
    from dkd_b200 import eval as E
    g, m, vset, qset, opt = _tiny_setup(dkd)
    gen = torch.Generator().manual_seed(3)
    Dt = 32
    tv = [torch.randn(len(f), Dt, generator=gen) for f in vset.feats]
    tq = [torch.randn(1, Dt, generator=gen) for _ in qset.feats]
    ctx = E.compute_context_info(m, TeacherVideoSet(vset.feats, tv), opt)
    assert ctx["teacher_frame_feat"].shape[:2] == ctx["video_mask"].shape and ctx["teacher_frame_feat"].shape[2] == Dt
    inher, explore, teacher, metas = E.compute_query2ctx_info(m, TeacherQuerySet(qset.feats, tq, len(vset)), opt, ctx)
    assert np.abs(inher - g["inher_scores"]).max() <= 5e-6
    order = [qset.ids.index(x) for x in metas]
    ref, _, _ = O.get_sim_scores(torch.cat(tq)[order], ctx["teacher_frame_feat"].cpu(), ctx["video_mask"].cpu())
    assert teacher.shape == inher.shape and np.abs(teacher - ref.numpy()).max() <= 2e-6
