"""Edge cases of the scoring path on the GPU: empty query sets, empty corpus shards, a corpus smaller than K,
single-frame and fully padded videos, argument errors reported as exceptions (never a silent fallback)."""
import pytest
import torch

from tests import synth

pytestmark = pytest.mark.gpu


def _params(D, seed):
    g = torch.Generator().manual_seed(seed)
    return [tuple(t.cuda() for t in (0.05 * torch.randn(D, D, generator=g), torch.zeros(D),
                                     0.05 * torch.randn(D, D, generator=g), torch.zeros(D))) for _ in range(2)]


@pytest.mark.parametrize("head", ["frame", "two_scale"])
def test_empty_queries_and_empty_shard(ops, head):
    from dkd_b200 import engine
    D, L, K = 128, 64, 100
    frames, mask, _ = synth.encoded_corpus(40, L, D, seed=1)
    fr = [frames.cuda(), frames.flip(0).contiguous().cuda() * mask.cuda()[:, :, None]]
    pc = engine.prepare_corpus(fr, mask.cuda(), _params(D, 2), heads=(head,))
    q0 = [torch.empty(0, D, device="cuda"), torch.empty(0, D, device="cuda")]
    s, i = engine.rank(pc, engine.prepare_queries(q0), K=K, head=head, precision="bf16")
    assert s.shape == (0, K) and i.shape == (0, K)
    # empty shard: 0 videos -> every list is padding; merged with a real shard it changes nothing
    pe = engine.prepare_corpus([f[:0] for f in fr], mask.cuda()[:0], _params(D, 2), heads=(head,), id_base=40)
    qs = [synth.encoded_queries(9, D, seed=3).cuda(), synth.encoded_queries(9, D, seed=4).cuda()]
    pq = engine.prepare_queries(qs)
    se, ie = engine.rank(pe, pq, K=K, head=head, precision="bf16")
    assert bool((ie == -1).all()) and bool(torch.isinf(se).all())
    s1, i1 = engine.rank(pc, pq, K=K, head=head, precision="exact")
    assert int((i1 >= 0).sum(dim=1).min()) == 40 and bool((i1[:, 40:] == -1).all())     # corpus smaller than K
    ms, mi = ops.merge_topk(torch.stack([s1, se]), torch.stack([i1, ie]))
    assert torch.equal(mi, i1) and torch.equal(ms, s1)
    sb, ib = engine.rank(pc, pq, K=K, head=head, precision="bf16", Kc=128)
    assert torch.equal(ib[:, :40].sort(dim=1).values, i1[:, :40].sort(dim=1).values)


def test_single_frame_and_fully_padded_videos(ops):
    """A video of one valid frame scores through that frame only; a video with no valid frame scores exactly
    -1e10 on the reference head (mask_logits, method/model.py:444) and ranks last."""
    from dkd_b200 import engine
    from oracle import oracle as O
    D, L = 64, 32
    frames, mask, _ = synth.encoded_corpus(12, L, D, seed=5, min_len=1)
    mask[3] = 0
    mask[3, 0] = 1                      # single frame
    mask[7] = 0                         # fully padded
    frames = frames * mask[:, :, None]
    q = synth.encoded_queries(20, D, seed=6)
    ref, _, _ = O.get_sim_scores(q, frames, mask)
    pc = engine.prepare_corpus([frames.cuda()], mask.cuda(), heads=("frame",))
    pq = engine.prepare_queries([q.cuda()])
    (s, a), = engine.score_frame_head(pc, pq, "exact")
    assert (s.cpu() - ref).abs().max() <= 2e-6
    assert bool((s[:, 7] == -1e10).all()) and bool((a[:, 3] == 0).all())
    # the tcgen05 GEMM (pair tiles: video 7 is the second half of pair 3): same fill value, first argmax
    (sb, ab), = engine.score_frame_head(pc, pq, "bf16")
    assert bool((sb[:, 7] == -1e10).all()) and bool((ab[:, 7] == 0).all()) and bool((ab[:, 3] == 0).all())
    live = [n for n in range(12) if n != 7]
    assert (sb.cpu() - ref)[:, live].abs().max() <= 1e-3
    _, ids = engine.rank(pc, pq, K=12, head="frame", precision="bf16")
    assert bool((ids[:, -1] == 7).all())


def test_argument_errors_raise(ops, dkd):
    from dkd_b200 import _lib
    x = torch.randn(4, 48, device="cuda")                      # D = 48: not a multiple of 64
    with pytest.raises(_lib.DkdError):
        ops.build_proposals(torch.randn(2, 32, 48, device="cuda"))
    with pytest.raises(_lib.DkdError):
        ops.normalize_rows(x.cpu())                            # CPU tensor: no CPU path
    with pytest.raises(_lib.DkdError):
        ops.topk(torch.randn(3, 10, device="cuda"), 1000)      # K > 256
    with pytest.raises(_lib.DkdError):
        ops.score_max_f32(x, torch.randn(2, 200, 48, device="cuda"))   # R > 128


@pytest.mark.parametrize("head,L,T", [("frame", 100, 32), ("frame", 37, 32), ("two_scale", 100, 16), ("two_scale", 50, 12)])
def test_gemm_rows_not_a_multiple_of_16(ops, head, L, T):
    """Corpora whose rows per video are not a multiple of 16 (longest video of 100 frames; map_size 16 -> 136
    proposals): the tcgen05 path pads every video with masked rows, and still returns the exact path's top-K."""
    from dkd_b200 import engine
    D, Nv, M, K = 128, 150, 70, 100
    frames, mask, _ = synth.encoded_corpus(Nv, L, D, seed=11, min_len=max(T, 20) if head == "two_scale" else 3)
    fr = [frames.cuda(), (frames.flip(0) * mask[:, :, None]).contiguous().cuda()]
    pc = engine.prepare_corpus(fr, mask.cuda(), _params(D, 12), T=T, heads=(head,))
    assert pc.Lg % 16 == 0 and pc.Pg % 16 == 0
    pq = engine.prepare_queries([synth.encoded_queries(M, D, seed=13).cuda(), synth.encoded_queries(M, D, seed=14).cuda()])
    s_ex, i_ex = engine.rank(pc, pq, K=K, head=head, precision="exact")
    s_bf, i_bf = engine.rank(pc, pq, K=K, head=head, precision="bf16", Kc=128)
    assert torch.equal(i_bf, i_ex) and torch.equal(s_bf, s_ex)
    if head == "frame":
        (sa, aa), _ = engine.score_frame_head(pc, pq, "bf16")
        (se, ae), _ = engine.score_frame_head(pc, pq, "exact")
        assert float((sa - se).abs().max()) <= 1e-3 and int(aa.max()) < L
