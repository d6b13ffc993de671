"""Full-size runs of BASELINE.json's eval configs (synthetic features of the named shapes, random-init DL-DKD++
encoders) through size-independent properties:

  * the tcgen05 path + exact rescoring returns the same top-100 (ids AND scores) as the all-exact path,
    for EVERY query of the config;
  * R@K computed from the ranked top-100 equals R@K computed from the rank of the ground-truth video in the
    dense exact fused scores (eval_q2m route);
  * video-sharded ranking + merge equals the unsharded ranking.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = {
    # BASELINE.json configs[0..2]
    "charades": dict(Nv=1334, L=128, Dv=1024, Nq=3720, Lq=30, Dq=768, H=384, T=32),
    "tvr": dict(Nv=2179, L=128, Dv=3072, Nq=10895, Lq=30, Dq=768, H=384, T=32),
    "activitynet": dict(Nv=4885, L=128, Dv=1024, Nq=17031, Lq=30, Dq=1024, H=384, T=32),
}


def _setup(name, head):
    import bench
    from dkd_b200 import engine
    from dkd_b200.model import DLDKD
    dev = torch.device("cuda")
    shape = SHAPES[name]
    model, frames, mask, qs = bench.synth_encoded(shape, dev, 0, DLDKD)
    if name == "charades":  # ragged corpus for one of the configs: lengths 64..128, padding zeroed like cat_tensor
        g = torch.Generator(device=dev).manual_seed(77)
        lengths = torch.randint(64, shape["L"] + 1, (shape["Nv"],), device=dev, generator=g)
        mask = (torch.arange(shape["L"], device=dev)[None] < lengths[:, None]).float()
        frames = [f * mask[:, :, None] for f in frames]
    pc = engine.prepare_corpus(frames, mask, [tuple(t.detach() for t in p) for p in model.attention_params()],
                               T=shape["T"], heads=(head,))
    pq = engine.prepare_queries([q.contiguous() for q in qs])
    return shape, engine, pc, pq, frames, mask, model


@pytest.mark.parametrize("name,head", [("charades", "two_scale"), ("charades", "frame"), ("tvr", "two_scale"),
                                       ("tvr", "frame"), ("activitynet", "two_scale")])
def test_config_bf16_rank_equals_exact_rank(ops, name, head):
    shape, engine, pc, pq, *_ = _setup(name, head)
    K = 100
    s_ex, i_ex = engine.rank(pc, pq, K=K, head=head, precision="exact")
    s_bf, i_bf = engine.rank(pc, pq, K=K, head=head, precision="bf16", rescore=True, Kc=128)
    torch.cuda.synchronize()
    same_ids = (i_bf == i_ex).all(dim=1)
    assert bool(same_ids.all()), f"{name}/{head}: {int((~same_ids).sum())} of {pq.M} queries differ in their top-{K}"
    assert torch.equal(s_bf, s_ex)
    # ranked lists are sorted (score desc, id asc on ties) and ids are valid and unique
    assert bool((s_ex[:, :-1] >= s_ex[:, 1:]).all())
    assert int(i_ex.min()) >= 0 and int(i_ex.max()) < shape["Nv"]
    assert bool((torch.sort(i_ex, dim=1).values[:, 1:] != torch.sort(i_ex, dim=1).values[:, :-1]).all())


def test_config_recall_from_topk_equals_recall_from_gt_rank(ops):
    """R@1/5/10/100 two ways on the TVR-shaped config: ranked top-100 lists vs rank of the GT video (eval_q2m)."""
    from dkd_b200 import eval as E
    shape, engine, pc, pq, *_ = _setup("tvr", "two_scale")
    Nq, Nv = shape["Nq"], shape["Nv"]
    t2v = {q: [q % Nv] for q in range(Nq)}                      # caption ids vid{q mod Nv}#enc#{q div Nv}
    _, top = engine.rank(pc, pq, K=100, head="two_scale", precision="bf16")
    fused, _ = engine.score_two_scale_head(pc, pq, "exact")
    ptr = torch.arange(Nq + 1, dtype=torch.int32, device="cuda")
    gts = (torch.arange(Nq, device="cuda") % Nv).to(torch.int32)
    ranks = ops.rank_of_gt(fused, ptr, gts).cpu().numpy()
    from_rank = [100.0 * np.count_nonzero(ranks <= k) / Nq for k in (1, 5, 10, 100)]
    from_topk = E.recall_from_topk(top, t2v)
    assert np.allclose(from_rank, from_topk), (from_rank, from_topk)


def test_config_sharded_rank_equals_unsharded(ops):
    """Charades-shaped corpus split over 4 video shards (the 4-GPU algebra on one device): local bf16 rank with
    global ids + dkd_merge_topk == the unsharded rank."""
    shape, engine, pc, pq, frames, mask, model = _setup("charades", "two_scale")
    params = [tuple(t.detach() for t in p) for p in model.attention_params()]
    s_all, i_all = engine.rank(pc, pq, K=100, head="two_scale", precision="bf16")
    ls, li = [], []
    for r in range(4):
        lo, hi = engine.shard_range(shape["Nv"], r, 4)
        pcs = engine.prepare_corpus([f[lo:hi].contiguous() for f in frames], mask[lo:hi].contiguous(), params,
                                    T=shape["T"], heads=("two_scale",), id_base=lo)
        s, i = engine.rank(pcs, pq, K=100, head="two_scale", precision="bf16")
        ls.append(s)
        li.append(i)
    ms, mi = ops.merge_topk(torch.stack(ls), torch.stack(li))
    assert torch.equal(mi, i_all) and torch.equal(ms, s_all)


def test_config_streamed_corpus_equals_resident_corpus(ops):
    """BASELINE.json configs[3] scaled to one test box: 20,000 videos x 32 clips (N(0,1) at D = 384), 3,000
    queries, streamed in chunks of 4,096 videos with query batches of 1,024 (engine.rank_streamed):
      * exact streamed top-100 == exact top-100 over the resident corpus (ids AND scores);
      * tcgen05 + rescoring streamed top-100 == exact streamed top-100 for every query;
      * two ranks' halves streamed separately + dkd_merge_topk == the whole corpus."""
    import bench
    from dkd_b200 import engine
    dev = torch.device("cuda")
    shape = dict(Nv=20_000, L=32, Nq=3_000, H=384, T=32)
    frames, mask, qs, attn = bench.synth_c4(shape, dev, 0)
    pqs = engine.split_queries(qs, 1024)
    assert [p.M for p in pqs] == [1024, 1024, 952]
    s_ex, i_ex = engine.rank_streamed(engine.iter_chunks(frames, mask, 4096), pqs, attn, K=100, precision="exact")
    s_bf, i_bf = engine.rank_streamed(engine.iter_chunks(frames, mask, 4096), pqs, attn, K=100, precision="bf16")
    same = (i_bf == i_ex).all(dim=1)
    assert bool(same.all()), f"{int((~same).sum())} of {shape['Nq']} queries differ"
    assert torch.equal(s_bf, s_ex)
    s_sc, i_sc = engine.rank_streamed(engine.iter_chunks(frames, mask, 4096), pqs, attn, K=100, precision="shortcut")
    assert torch.equal(i_sc, i_ex) and torch.equal(s_sc, s_ex)
    pc = engine.prepare_corpus(frames, mask, attn, T=32, heads=("two_scale",), precisions=("exact",))
    s_all, i_all = engine.rank(pc, engine.prepare_queries(qs), K=100, head="two_scale", precision="exact")
    assert torch.equal(i_all, i_ex) and torch.equal(s_all, s_ex)
    del pc
    halves = []
    for r in range(2):
        lo, hi = engine.shard_range(shape["Nv"], r, 2)
        halves.append(engine.rank_streamed(engine.iter_chunks([f[lo:hi] for f in frames], mask[lo:hi], 4096, id_base=lo),
                                           pqs, attn, K=100, precision="bf16"))
    ms, mi = ops.merge_topk(torch.stack([h[0] for h in halves]), torch.stack([h[1] for h in halves]))
    assert torch.equal(mi, i_ex) and torch.equal(ms, s_ex)


@pytest.mark.parametrize("name", ["tvr", "charades"])
def test_config_dense_error_within_certificate_eps(ops, name):
    """The assumption behind engine.rank's candidate certificate, checked on EVERY (query, video) pair of a full
    config: |approximate fused score - exact fused score| <= CERT_EPS (bf16 operands, 1e-3: the north_star
    tolerance) resp. CERT_EPS_F16 (IEEE-half operands), after the ambiguous-key-clip pass."""
    from dkd_b200 import engine
    import bench
    from dkd_b200.model import DLDKD
    dev = torch.device("cuda")
    shape = SHAPES[name]
    model, frames, mask, qs = bench.synth_encoded(shape, dev, 0, DLDKD)
    pc = engine.prepare_corpus(frames, mask, [tuple(t.detach() for t in p) for p in model.attention_params()],
                               T=shape["T"], heads=("two_scale",), precisions=("exact", "bf16", "fp16"))
    pq = engine.prepare_queries([q.contiguous() for q in qs])
    exact, _ = engine.score_two_scale_head(pc, pq, "exact")
    for precision, eps in (("bf16", engine.CERT_EPS), ("fp16", engine.CERT_EPS_F16)):
        approx, _ = engine.score_two_scale_head(pc, pq, precision)
        err = float((approx - exact).abs().max())
        assert err <= eps, f"{name}/{precision}: dense error {err:.2e} exceeds the certificate bound {eps:.1e}"


@pytest.mark.parametrize("name,head", [("tvr", "two_scale"), ("tvr", "frame"), ("activitynet", "two_scale")])
def test_config_fp16_rank_equals_exact_rank(ops, name, head):
    """IEEE-half GEMM operands (precision="fp16"): same top-100 (ids and scores) as the exact path for every query."""
    import bench
    from dkd_b200 import engine
    from dkd_b200.model import DLDKD
    dev = torch.device("cuda")
    shape = SHAPES[name]
    model, frames, mask, qs = bench.synth_encoded(shape, dev, 0, DLDKD)
    pc = engine.prepare_corpus(frames, mask, [tuple(t.detach() for t in p) for p in model.attention_params()],
                               T=shape["T"], heads=(head,), precisions=("exact", "fp16"))
    pq = engine.prepare_queries([q.contiguous() for q in qs])
    s_ex, i_ex = engine.rank(pc, pq, K=100, head=head, precision="exact")
    engine.STATS["certify_fallback_queries"] = 0
    s_h, i_h = engine.rank(pc, pq, K=100, head=head, precision="fp16", Kc=128)
    assert torch.equal(i_h, i_ex) and torch.equal(s_h, s_ex)
    assert engine.STATS["certify_fallback_queries"] <= pq.M // 100       # the fallback is the exception, not the path


@pytest.mark.parametrize("name", ["tvr", "charades"])
def test_config_shortcut_rank_equals_exact_rank(ops, name):
    """precision="shortcut" (exact clip scores for every pair via the linearity shortcut, approximate fp16 frame
    gather, exact frame rescoring of the candidates): same top-100 as the all-exact path for every query; its dense
    scores differ from the exact ones only by the fp16 frame term."""
    import bench
    from dkd_b200 import engine
    from dkd_b200.model import DLDKD
    dev = torch.device("cuda")
    shape = SHAPES[name]
    model, frames, mask, qs = bench.synth_encoded(shape, dev, 0, DLDKD)
    if name == "charades":
        g = torch.Generator(device=dev).manual_seed(77)
        lengths = torch.randint(64, shape["L"] + 1, (shape["Nv"],), device=dev, generator=g)
        mask = (torch.arange(shape["L"], device=dev)[None] < lengths[:, None]).float()
        frames = [f * mask[:, :, None] for f in frames]
    pc = engine.prepare_corpus(frames, mask, [tuple(t.detach() for t in p) for p in model.attention_params()],
                               T=shape["T"], heads=("two_scale",), precisions=("exact", "shortcut"))
    assert pc.branches[0].prop_b is None and pc.branches[0].prop_h is None      # no GEMM operand at all
    pq = engine.prepare_queries([q.contiguous() for q in qs])
    s_ex, i_ex = engine.rank(pc, pq, K=100, head="two_scale", precision="exact")
    engine.STATS["certify_fallback_queries"] = 0
    s_sc, i_sc = engine.rank(pc, pq, K=100, head="two_scale", precision="shortcut", Kc=128)
    assert torch.equal(i_sc, i_ex) and torch.equal(s_sc, s_ex)
    assert engine.STATS["certify_fallback_queries"] <= pq.M // 100
    exact, _ = engine.score_two_scale_head(pc, pq, "exact")
    approx, per = engine.score_two_scale_head(pc, pq, "shortcut")
    assert float((approx - exact).abs().max()) <= engine.CERT_EPS_F16
