"""Kernel-level parity: every C-ABI entry point against the CPU oracle on the same seeded inputs.

Tolerances (north_star): fp32 paths 2e-6 abs on cosine scores (summation order only); bf16 GEMM path
1e-3 abs vs fp32, and 2e-5 vs an fp32 evaluation of the SAME bf16-rounded operands; indices exact
except where the oracle's own top-2 gap is below the stated noise floor.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O
from tests import synth

pytestmark = pytest.mark.gpu

FP32_TOL = 2e-6
BF16_TOL = 1e-3


def _cuda(*ts):
    return [t.cuda().contiguous() for t in ts]


def _argmax_ok(got, ref_scores_all, ref_arg, tol):
    """got/ref_arg (M, N); ref_scores_all (M, R, N). Mismatches allowed only if the oracle's score at the
    kernel's index is within tol of the oracle max (a near tie)."""
    got = got.long()
    bad = got != ref_arg
    if not bad.any():
        return True, 0
    m, n = torch.nonzero(bad, as_tuple=True)
    s_got = ref_scores_all[m, got[m, n], n]
    s_ref = ref_scores_all[m, ref_arg[m, n], n]
    return bool(((s_ref - s_got).abs() <= tol).all()), int(bad.sum())


def test_normalize_rows(ops):
    x = torch.randn(300, 384, generator=torch.Generator().manual_seed(1))
    x[7] = 0.0  # zero row: F.normalize gives zeros (eps clamp)
    (xc,) = _cuda(x)
    f32, b16 = ops.normalize_rows(xc, want_f32=True, want_bf16=True, rows_pad=384)
    ref = F.normalize(x, dim=-1)
    assert torch.allclose(f32[:300].cpu(), ref, atol=1e-7, rtol=1e-6)
    assert torch.equal(f32[300:].cpu(), torch.zeros(84, 384))
    assert torch.equal(b16[:300].cpu(), f32[:300].cpu().to(torch.bfloat16))
    assert torch.equal(b16[300:].float().cpu(), torch.zeros(84, 384))


@pytest.mark.parametrize("T", [32, 8])
def test_downsample_clips(ops, T):
    frames, mask, lengths = synth.encoded_corpus(37, 128, 64, seed=2, min_len=1)
    fc, lc = _cuda(frames, lengths)
    got = ops.downsample_clips(fc, lc, T=T).cpu()
    ref = O.downsample_clips(frames, lengths, T)
    assert torch.allclose(got, ref, atol=1e-6, rtol=1e-6)


def test_build_proposals(ops):
    frames, mask, lengths = synth.encoded_corpus(19, 128, 384, seed=3)
    clips = O.downsample_clips(frames, lengths, 32)
    (cc,) = _cuda(clips)
    pb, ps, pf = ops.build_proposals(cc, want_bf16=True, want_scale=True, want_f32=True)
    ref = O.build_proposals(clips)  # (Nv, 528, D)
    assert pf.shape == ref.shape == (19, 528, 384)
    assert torch.allclose(pf.cpu(), ref, atol=1e-6, rtol=1e-5)
    refn = F.normalize(ref, dim=-1)
    assert (pb.float().cpu() - refn).abs().max() <= 2 ** -8 * refn.abs().max() + 1e-6
    # prop_scale = 1 / (w * ||mean||)
    w = torch.cat([torch.full((32 - k + 1,), float(k)) for k in range(1, 33)])
    ref_scale = 1.0 / (w[None, :] * ref.norm(dim=-1))
    assert torch.allclose(ps.cpu(), ref_scale, rtol=2e-6)
    # index formula
    assert O.proposal_index(1, 0, 32) == 0 and O.proposal_index(32, 0, 32) == 527
    assert ops.proposal_index(2, 5, 32) == 32 + 5


@pytest.mark.parametrize("M,Nv,L,D", [(50, 23, 128, 384), (130, 9, 77, 64), (1, 3, 16, 512), (64, 5, 128, 32)])
def test_score_max_f32_matches_get_sim_scores(ops, M, Nv, L, D):
    frames, mask, lengths = synth.encoded_corpus(Nv, L, D, seed=10 + M)
    q = synth.encoded_queries(M, D, seed=20 + M)
    s_ref, rows_ref, a_ref = O.get_sim_scores(q, frames, mask)
    qc, fc, mc = _cuda(q, frames, mask.to(torch.uint8))
    qn, _ = ops.normalize_rows(qc)
    xn, _ = ops.normalize_rows(fc)
    om, oa, rows = ops.score_max_f32(qn, xn.view(Nv, L, D), mc, want_rows=True)
    assert (om.cpu() - s_ref).abs().max() <= FP32_TOL
    rows = rows.cpu()
    valid = (mask.T[None] > 0).expand_as(rows_ref)
    assert (rows[valid] - rows_ref[valid]).abs().max() <= FP32_TOL
    assert torch.equal(rows[~valid], torch.full_like(rows[~valid], -1e10))  # mask_logits fill, exact
    ok, nbad = _argmax_ok(oa.cpu(), rows_ref, a_ref, FP32_TOL)
    assert ok, f"{nbad} argmax mismatches beyond fp32 ties"


def test_score_max_f32_no_mask_and_fully_masked_video(ops):
    frames, mask, lengths = synth.encoded_corpus(6, 32, 64, seed=5)
    q = synth.encoded_queries(10, 64, seed=6)
    mask2 = mask.clone()
    mask2[2] = 0  # never happens in the reference data (len >= 1) but must not crash: all -1e10, idx 0
    s_ref, _, a_ref = O.get_sim_scores(q, frames, mask2)
    qc, fc, mc = _cuda(q, frames, mask2.to(torch.uint8))
    qn, _ = ops.normalize_rows(qc)
    xn, _ = ops.normalize_rows(fc)
    om, oa, _ = ops.score_max_f32(qn, xn.view(6, 32, 64), mc)
    assert (om.cpu() - s_ref).abs().max() <= FP32_TOL
    assert torch.equal(om[:, 2].cpu(), torch.full((10,), -1e10))
    assert torch.equal(oa[:, 2].cpu().long(), a_ref[:, 2])
    s_ref2, _, _ = O.get_sim_scores(q, frames, None)
    om2, _, _ = ops.score_max_f32(qn, xn.view(6, 32, 64), None)
    assert (om2.cpu() - s_ref2).abs().max() <= FP32_TOL


@pytest.mark.parametrize("M,Nv,L,D", [(50, 23, 128, 384), (300, 9, 77, 64), (1, 3, 16, 512), (260, 40, 128, 384),
                                      (129, 5, 100, 128)])
def test_score_max_exact_matches_get_sim_scores(ops, M, Nv, L, D):
    """The tcgen05 kind::tf32 exact path (split operands, packed row planes) vs the oracle's get_sim_scores:
    ragged masks, R not a multiple of 16, a fully masked video, several tiles per video; CSR == dense."""
    frames, mask, lengths = synth.encoded_corpus(Nv, L, D, seed=40 + M)
    mask = mask.clone()
    if Nv > 2:
        mask[2] = 0
    q = synth.encoded_queries(M, D, seed=50 + M)
    s_ref, rows_ref, a_ref = O.get_sim_scores(q, frames, mask)
    qc, fc, mc = _cuda(q, frames, mask.to(torch.uint8))
    qn, _ = ops.normalize_rows(qc)
    xn, _ = ops.normalize_rows(fc)
    planes = ops.pack_rows(xn.view(Nv, L, D))
    om, oa = ops.score_max_exact(qn, planes, L, mc)
    torch.cuda.synchronize()
    assert (om.cpu() - s_ref).abs().max() <= FP32_TOL
    if Nv > 2:
        assert torch.equal(om[:, 2].cpu(), torch.full((M,), -1e10)) and (oa[:, 2] == 0).all()
    ok, nbad = _argmax_ok(oa.cpu(), rows_ref, a_ref, FP32_TOL)
    assert ok, f"{nbad} argmax mismatches beyond fp32 ties"
    om2, _ = ops.score_max_exact(qn, planes, L, None)
    assert (om2.cpu() - O.get_sim_scores(q, frames, None)[0]).abs().max() <= FP32_TOL
    # CSR restriction reproduces the dense bits
    g = torch.Generator().manual_seed(60)
    gap = torch.rand(M, Nv, generator=g)
    vid_ptr, q_list, slot = ops.select_pairs_csr(gap.cuda(), 0.4)
    n_e = int(vid_ptr[-1].item())
    cs, ck = ops.score_max_exact(qn, planes, L, mc, csr=(vid_ptr, q_list))
    sl = slot[:n_e].long()
    assert torch.equal(cs[:n_e], om.flatten()[sl]) and torch.equal(ck[:n_e], oa.flatten()[sl])


@pytest.mark.parametrize("M,Nv,D,T", [(70, 11, 384, 32), (5, 3, 64, 32), (300, 160, 384, 32), (133, 7, 128, 8),
                                      (260, 5, 512, 32)])
def test_clip_score_f32(ops, M, Nv, D, T):
    """Exact clip-scale scores (tcgen05 kind::tf32, split operands) vs the oracle's direct formulation."""
    frames, mask, lengths = synth.encoded_corpus(Nv, 128, D, seed=31)
    q = synth.encoded_queries(M, D, seed=32)
    clips = O.downsample_clips(frames, lengths, T)
    props = O.build_proposals(clips)
    s_ref, all_ref, k_ref = O.clip_scale_scores(q, props)
    qc, cc = _cuda(q, clips)
    qn, _ = ops.normalize_rows(qc)
    _, ps, _ = ops.build_proposals(cc, want_bf16=False)
    om, oa = ops.clip_score_f32(qn, cc, ps)
    torch.cuda.synchronize()
    assert (om.cpu() - s_ref).abs().max() <= FP32_TOL
    ok, nbad = _argmax_ok(oa.cpu(), all_ref, k_ref, FP32_TOL)
    assert ok, f"{nbad} key-clip mismatches beyond fp32 ties"


def test_clip_score_f32_csr_and_scatter(ops):
    """CSR restriction (some videos with empty lists, ragged tiles) == the dense result at those pairs; the
    scatter form writes them into dense matrices in place."""
    M, Nv, D = 333, 23, 384
    frames, mask, lengths = synth.encoded_corpus(Nv, 128, D, seed=33)
    q = synth.encoded_queries(M, D, seed=34)
    fc, lc, qc = _cuda(frames, lengths, q)
    clips = ops.downsample_clips(fc, lc)
    _, ps, _ = ops.build_proposals(clips, want_bf16=False)
    qn, _ = ops.normalize_rows(qc)
    dm, da = ops.clip_score_f32(qn, clips, ps)
    g = torch.Generator().manual_seed(35)
    gap = torch.rand(M, Nv, generator=g)
    gap[:, 3] = 1.0                      # video 3: empty list
    gap[:, 5] = 0.0                      # video 5: every query (3 tiles, last one ragged)
    csr = ops.select_pairs_csr(gap.cuda(), 0.3)
    vid_ptr, q_list, slot = csr
    n_e = int(vid_ptr[-1].item())
    assert n_e == int((gap < 0.3).sum())
    cs, ck = ops.clip_score_f32(qn, clips, ps, csr=(vid_ptr, q_list))
    sl = slot[:n_e].long()
    assert torch.equal(cs[:n_e], dm.flatten()[sl]) and torch.equal(ck[:n_e], da.flatten()[sl])
    out_m = torch.full((M, Nv), -7.0, device="cuda")
    out_a = torch.full((M, Nv), -7, dtype=torch.int32, device="cuda")
    ops.clip_score_f32(qn, clips, ps, csr=(vid_ptr, q_list), scatter=(slot, out_m, out_a))
    sel = (gap < 0.3).cuda()
    assert torch.equal(out_m[sel], dm[sel]) and torch.equal(out_a[sel], da[sel])
    assert (out_m[~sel] == -7.0).all() and (out_a[~sel] == -7).all()
    # (begin, count) runs instead of a CSR: same entries, explicit counts
    counts = (vid_ptr[1:] - vid_ptr[:-1]).contiguous()
    out_m2 = torch.full((M, Nv), -7.0, device="cuda")
    out_a2 = torch.full((M, Nv), -7, dtype=torch.int32, device="cuda")
    ops.clip_score_f32(qn, clips, ps, csr=(vid_ptr[:-1].contiguous(), q_list, counts), scatter=(slot, out_m2, out_a2))
    assert torch.equal(out_m2, out_m) and torch.equal(out_a2, out_a)


@pytest.mark.parametrize("M,pad,Nv,R,D,masked", [
    (200, 256, 37, 528, 384, False),   # clip-proposal shape, CTA pair (cta_group::2), ragged M
    (50, 128, 23, 128, 384, True),     # reference frame path with mask, single CTA (cta_group::1)
    (128, 256, 300, 528, 384, False),  # CTA pair, many work items, second query tile all padding
    (300, 128, 150, 528, 384, False),  # three query tiles -> single-CTA kernel
    (700, 256, 61, 128, 384, True),    # CTA pair with mask, three tile pairs
    (10, 128, 4, 64, 128, True),
    # pair tiles (R <= 128 on CTA pairs: one N = 2 R tile spans two videos, dkd_score_bf16.cu kPair)
    (300, 256, 45, 112, 384, True),    # 7 chunks per video (16-column tail load), odd video count
    (513, 256, 128, 128, 384, False),  # unmasked, three tile pairs, even video count
    (100, 256, 7, 48, 128, True),      # 3 chunks per video, D = 128
    (256, 256, 2, 16, 192, True),      # smallest: one pair, one chunk per video; 3 K blocks
    (256, 256, 3, 32, 128, False),     # odd count without mask: the last pair's second half is a repeat
])
def test_score_max_bf16(ops, M, pad, Nv, R, D, masked):
    g = torch.Generator().manual_seed(77)
    x = torch.randn(Nv, R, D, generator=g) + 0.5 * torch.randn(Nv, 1, D, generator=g)
    q = torch.randn(M, D, generator=g)
    mask = None
    if masked:
        lengths = torch.randint(1, R + 1, (Nv,), generator=g)
        mask = (torch.arange(R)[None] < lengths[:, None]).float()
    s_ref, rows_ref, a_ref = O.get_sim_scores(q, x, mask)
    qc, xc = _cuda(q, x)
    Mpad = ops.round_up(M, pad)
    _, qb = ops.normalize_rows(qc, want_f32=False, want_bf16=True, rows_pad=Mpad)
    _, xb = ops.normalize_rows(xc, want_f32=False, want_bf16=True)
    mc = None if mask is None else mask.to(torch.uint8).cuda()
    om, oa, og, fl = ops.score_max_bf16(qb, M, xb, Nv, R, mc, want_gap=True, flag_tau=3e-3)
    torch.cuda.synchronize()
    # (0) flag bits == (gap < tau), and the flagged-pair runs list exactly those pairs
    want_flag = og < 3e-3
    bits = ((fl.cpu().numpy().view(np.uint32)[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(M, -1)[:, :Nv]
    assert np.array_equal(bits.astype(bool), want_flag.cpu().numpy())
    vb, ql, vc, slot = ops.select_flagged(fl, Nv)
    vb_c, vc_c, ql_c, sl_c = vb.cpu().numpy(), vc.cpu().numpy(), ql.cpu().numpy(), slot.cpu().numpy()
    assert int(vc_c.sum()) == int(want_flag.sum())
    for n in range(Nv):
        run = slice(int(vb_c[n]), int(vb_c[n]) + int(vc_c[n]))
        assert sorted(ql_c[run].tolist()) == np.nonzero(want_flag.cpu().numpy()[:, n])[0].tolist()
        assert np.array_equal(sl_c[run], ql_c[run].astype(np.int64) * Nv + n)
    # (1) vs fp32 oracle: north_star tolerance
    assert (om.cpu() - s_ref).abs().max() <= BF16_TOL
    # (2) vs fp32 evaluation of the same bf16 operands: accumulate-order noise only
    qf, xf = qb[:M].float().cpu(), xb.float().cpu().view(Nv, R, D)
    rows_b = torch.einsum("md,nrd->mrn", qf, xf)
    if mask is not None:
        rows_b = O.mask_logits(rows_b, mask.T[None])
    s_b, a_b = rows_b.max(dim=1)
    assert (om.cpu() - s_b).abs().max() <= 2e-5
    ok, nbad = _argmax_ok(oa.cpu(), rows_b, a_b, 2e-5)
    assert ok, f"{nbad} argmax mismatches vs bf16-operand reference"
    # (3) vs oracle argmax: only near ties at bf16 resolution may differ
    ok, nbad = _argmax_ok(oa.cpu(), rows_ref, a_ref, 2 * BF16_TOL)
    assert ok
    # (4) gap = best - runner-up of the same bf16 operands
    if R > 1:
        top2 = torch.topk(rows_b, 2, dim=1).values
        live = top2[:, 1] > -1e9                      # runner-up masked: gap ~ 1e10, not comparable in fp32
        assert ((top2[:, 0] - top2[:, 1]) - og.cpu())[live].abs().max() <= 4e-5
        assert (og.cpu()[~live] > 1e9).all()


@pytest.mark.parametrize("M,pad,Nv,R,D", [(200, 256, 37, 528, 384), (50, 128, 23, 128, 384)])
def test_score_max_f16_operands(ops, M, pad, Nv, R, D):
    """The same GEMM on IEEE-half operands: 8 x tighter than bf16 against the fp32 oracle, accumulate-order noise
    against an fp32 evaluation of the same half operands, proposals built directly as half rows."""
    g = torch.Generator().manual_seed(78)
    x = torch.randn(Nv, R, D, generator=g) + 0.5 * torch.randn(Nv, 1, D, generator=g)
    q = torch.randn(M, D, generator=g)
    s_ref, rows_ref, a_ref = O.get_sim_scores(q, x, None)
    qc, xc = _cuda(q, x)
    Mpad = ops.round_up(M, pad)
    _, _, qh = ops.normalize_rows(qc, want_f32=False, rows_pad=Mpad, want_f16=True)
    _, _, xh = ops.normalize_rows(xc, want_f32=False, want_f16=True)
    om, oa = ops.score_max_bf16(qh, M, xh, Nv, R)
    assert (om.cpu() - s_ref).abs().max() <= BF16_TOL / 8
    rows_h = torch.einsum("md,nrd->mrn", qh[:M].float().cpu(), xh.float().cpu().view(Nv, R, D))
    s_h, a_h = rows_h.max(dim=1)
    assert (om.cpu() - s_h).abs().max() <= 2e-5
    ok, nbad = _argmax_ok(oa.cpu(), rows_h, a_h, 2e-5)
    assert ok, f"{nbad} argmax mismatches vs half-operand reference"
    if R == 528:   # half proposal rows == half rounding of the normalised window means
        frames, mask, lengths = synth.encoded_corpus(9, 128, D, seed=5)
        clips = ops.downsample_clips(frames.cuda(), lengths.cuda())
        ph, ps = ops.build_proposals_f16(clips, want_scale=True)
        pb, ps2, _ = ops.build_proposals(clips)
        refn = F.normalize(O.build_proposals(O.downsample_clips(frames, lengths, 32)), dim=-1)
        assert (ph.float().cpu() - refn).abs().max() <= 2 ** -11 * refn.abs().max() + 1e-6
        assert torch.equal(ps, ps2)


def _branch_params(D, seed):
    g = torch.Generator().manual_seed(seed)
    kw = 0.05 * torch.randn(D, D, generator=g)
    vw = 0.05 * torch.randn(D, D, generator=g)
    kb = 0.01 * torch.randn(D, generator=g)
    vb = 0.01 * torch.randn(D, generator=g)
    return kw, kb, vw, vb


@pytest.mark.parametrize("Nv,L,D", [(7, 128, 384), (5, 50, 64)])
def test_frame_attn_table(ops, Nv, L, D):
    frames, mask, lengths = synth.encoded_corpus(Nv, L, D, seed=41)
    kw, kb, vw, vb = _branch_params(D, 42)
    clips = O.downsample_clips(frames, lengths, 32)
    props = O.build_proposals(clips)
    key, val = F.linear(frames, kw, kb), F.linear(frames, vw, vb)
    ref = F.normalize(O.attention_table(key, val, mask, props), dim=-1)
    kc, vc, cc, lc = _cuda(key, val, clips, lengths)
    tf, tb = ops.frame_attn_table(kc, vc, cc, lc)
    assert (tf.cpu() - ref).abs().max() <= 5e-6
    assert tb.dtype == torch.float16 and (tb.float().cpu() - ref).abs().max() <= 2 ** -11


def test_two_scale_branch_exact_path(ops):
    """N6 end to end (one branch), exact fp32 kernels vs the oracle."""
    Nv, L, D, M = 13, 128, 384, 45
    frames, mask, lengths = synth.encoded_corpus(Nv, L, D, seed=51)
    q = synth.encoded_queries(M, D, seed=52)
    kw, kb, vw, vb = _branch_params(D, 53)
    ref = O.two_scale_branch(q, frames, mask, kw, kb, vw, vb)
    fc, lc, qc = _cuda(frames, lengths, q)
    clips = ops.downsample_clips(fc, lc)
    pb, ps, _ = ops.build_proposals(clips)
    qn, _ = ops.normalize_rows(qc)
    s_clip, k_clip = ops.clip_score_f32(qn, clips, ps)
    key, val = F.linear(frames, kw, kb).cuda(), F.linear(frames, vw, vb).cuda()
    tf, tb = ops.frame_attn_table(key, val, clips, lc)
    fused, fr = ops.frame_fuse(qn, tf, s_clip, k_clip, 0.7, 0.3, 1.0, want_frame=True)
    assert (s_clip.cpu() - ref["clip"]).abs().max() <= FP32_TOL
    # the key clip is bit-exact except where the oracle's own best and second-best proposal scores are closer than
    # fp32 summation noise (documented ties, listed): there either proposal is "the" key clip
    allp = O.clip_scale_scores(q, ref["proposals"])[1]                       # (M, P, Nv)
    top2 = torch.topk(allp, 2, dim=1).values
    tie = (top2[:, 0] - top2[:, 1]) <= 2e-6
    same = k_clip.cpu().long() == ref["key_clip"]
    assert bool((same | tie).all()), f"key clip differs outside ties at {(~(same | tie)).nonzero().tolist()[:5]}"
    assert int(tie.sum()) <= 3, f"{int(tie.sum())} fp32 key-clip ties among {M * Nv} pairs: {tie.nonzero().tolist()[:5]}"
    assert (fr.cpu()[same] - ref["frame"][same]).abs().max() <= 5e-6
    assert (fused.cpu()[same] - ref["branch"][same]).abs().max() <= 5e-6


@pytest.mark.parametrize("M,Nv,D", [(77, 45, 384), (33, 7, 64), (260, 100, 512)])
def test_frame_fuse_fp16_gather(ops, M, Nv, D):
    """The fp16 dense gather (half2 products, fp32 accumulation across 4-product chunks) against the fp32 gather
    on the same table: operand rounding 2^-12 relative + chunk rounding -> well inside 2e-4; ragged M / Nv tails."""
    P = 528
    g = torch.Generator().manual_seed(123)
    table = F.normalize(torch.randn(Nv, P, D, generator=g), dim=-1)
    q = F.normalize(torch.randn(M, D, generator=g), dim=-1)
    clip = torch.rand(M, Nv, generator=g)
    key = torch.randint(0, P, (M, Nv), generator=g, dtype=torch.int32)
    tc, qc, cc, kc = _cuda(table, q, clip, key)
    ref_fused, ref_fr = ops.frame_fuse(qc, tc, cc, kc, 0.7, 0.3, 0.7, want_frame=True)
    want = (q[:, None, :] * table[torch.arange(Nv)[None, :], key.long()]).sum(-1)
    assert (ref_fr.cpu() - want).abs().max() <= 2e-6
    fused = torch.full((M, Nv), 0.25, device="cuda")
    got_fused, got_fr = ops.frame_fuse(qc.half(), tc.half(), cc, kc, 0.7, 0.3, 0.7, fused=fused, accumulate=True,
                                       want_frame=True)
    torch.cuda.synchronize()
    assert (got_fr - ref_fr).abs().max().item() <= 2e-4
    assert (got_fused - (ref_fused + 0.25)).abs().max().item() <= 2e-4


def test_fuse_scores_bit_exact(ops):
    g = torch.Generator().manual_seed(61)
    a, b = torch.randn(1000, 37, generator=g), torch.randn(1000, 37, generator=g)
    ac, bc = _cuda(a, b)
    got = ops.fuse_scores(ac, bc, 0.7, 0.3).cpu().numpy()
    ref = O.fuse_branches(a.numpy(), b.numpy())
    assert np.array_equal(got, ref)  # numpy rounding order of method/eval.py:254, bit exact


@pytest.mark.parametrize("M,Nv,K", [(33, 2179, 100), (5, 40, 100), (17, 1000, 128), (3, 5000, 256)])
def test_topk(ops, M, Nv, K):
    g = torch.Generator().manual_seed(71)
    s = torch.randn(M, Nv, generator=g)
    s[:, ::7] = s[:, 3:4]  # many exact ties: lower id must win
    (sc,) = _cuda(s)
    ts, ti = ops.topk(sc, K, id_base=1000)
    ref = O.topk_ids(s.numpy(), K)
    kk = min(K, Nv)
    assert np.array_equal(ti.cpu().numpy()[:, :kk] - 1000, ref[:, :kk])
    assert np.array_equal(ts.cpu().numpy()[:, :kk], np.take_along_axis(s.numpy(), ref[:, :kk], 1))
    if K > Nv:
        assert (ti.cpu()[:, Nv:] == -1).all()


def test_merge_topk_equals_global_topk(ops):
    g = torch.Generator().manual_seed(81)
    M, Nv, K, G = 21, 1200, 100, 4
    s = torch.randn(M, Nv, generator=g)
    s[:, 5::11] = 0.25
    (sc,) = _cuda(s)
    shard = Nv // G
    ls, li = [], []
    for r in range(G):
        part = sc[:, r * shard:(r + 1) * shard].contiguous()
        a, b = ops.topk(part, K, id_base=r * shard)
        ls.append(a)
        li.append(b)
    ms, mi = ops.merge_topk(torch.stack(ls), torch.stack(li))
    gs, gi = ops.topk(sc, K)
    assert torch.equal(mi, gi) and torch.equal(ms, gs)
    assert np.array_equal(gi.cpu().numpy(), O.topk_ids(s.numpy(), K))


def test_rank_of_gt_and_recall(ops):
    g = torch.Generator().manual_seed(91)
    M, Nv = 64, 300
    s = torch.randn(M, Nv, generator=g)
    s[:, 10] = s[:, 20]
    gts = {i: ([i % Nv] if i % 3 else [i % Nv, (i * 7) % Nv]) for i in range(M)}
    ptr = np.zeros(M + 1, np.int32)
    ids = []
    for i in range(M):
        ids += gts[i]
        ptr[i + 1] = len(ids)
    sc = s.cuda()
    r = ops.rank_of_gt(sc, torch.from_numpy(ptr).cuda(), torch.tensor(ids, dtype=torch.int32).cuda())
    ref = O.gt_ranks(-s.numpy(), gts, stable=True)
    assert np.array_equal(r.cpu().numpy(), ref)


def test_candidate_rescoring_plumbing(ops):
    """candidates -> CSR -> exact clip score per entry -> frame fuse per entry -> sort == dense exact path."""
    Nv, L, D, M, K = 40, 128, 64, 30, 16
    frames, mask, lengths = synth.encoded_corpus(Nv, L, D, seed=101)
    q = synth.encoded_queries(M, D, seed=102)
    kw, kb, vw, vb = _branch_params(D, 103)
    fc, lc, qc = _cuda(frames, lengths, q)
    clips = ops.downsample_clips(fc, lc)
    _, ps, _ = ops.build_proposals(clips, want_bf16=False)
    qn, _ = ops.normalize_rows(qc)
    s_clip, k_clip = ops.clip_score_f32(qn, clips, ps)
    tf, _ = ops.frame_attn_table(F.linear(frames, kw, kb).cuda(), F.linear(frames, vw, vb).cuda(), clips, lc,
                                 want_f16=False)
    fused, _ = ops.frame_fuse(qn, tf, s_clip, k_clip, 0.7, 0.3, 1.0)
    ts, ti = ops.topk(fused, K)
    # perturb candidate order, then rescore
    perm = torch.randperm(K)
    cand = ti[:, perm].contiguous()
    vid_ptr, q_list, slot = ops.candidates_to_csr(cand, Nv)
    cs, ck = ops.clip_score_f32(qn, clips, ps, csr=(vid_ptr, q_list))
    cand_scores = torch.zeros(M, K, device="cuda")
    ops.frame_fuse_csr(qn, tf, cs, ck, (vid_ptr, q_list, slot), 0.7, 0.3, 1.0, cand_scores, False)
    rs, ri = ops.sort_candidates(cand_scores, cand, K)
    assert torch.equal(ri, ti)
    assert torch.equal(rs, ts)


def test_gemm_epilogue_lists_equal_flag_matrix_selection(ops):
    """dkd_score_max_bf16_lists (ambiguous pairs appended to per-video lists by the GEMM's epilogue) + dkd_clip_score_list
    == the bit-matrix route (dkd_score_max_bf16 flags -> dkd_select_flagged -> dkd_clip_score_f32 scatter): same flagged
    pairs per video, same dense matrices afterwards, bit for bit."""
    Nv, L, D, M, tau = 300, 128, 384, 700, 1e-3
    frames, mask, lengths = synth.encoded_corpus(Nv, L, D, seed=61, shared=1.5)
    q = synth.encoded_queries(M, D, seed=62)
    fc, lc, qc = _cuda(frames, lengths, q)
    clips = ops.downsample_clips(fc, lc)
    pb, ps, _ = ops.build_proposals(clips)
    planes = ops.pack_clips(clips)
    Mpad = ops.round_up(M, 256)
    qn, qb = ops.normalize_rows(qc, want_f32=True, want_bf16=True, rows_pad=Mpad)
    qn = qn[:M]
    s1, k1, flags = ops.score_max_bf16(qb, M, pb.view(-1, D), Nv, 528, flag_tau=tau)
    s2, k2, cnt, lst = ops.score_max_bf16_lists(qb, M, pb.view(-1, D), Nv, 528, tau)
    assert torch.equal(s1, s2) and torch.equal(k1, k2)
    vb, ql, vc, slot = ops.select_flagged(flags, Nv)
    assert torch.equal(vc, cnt) and int(cnt.sum()) > 0
    vbc, qlc, vcc, lstc = vb.cpu(), ql.cpu(), vc.cpu(), lst.cpu()
    for n in (0, 1, 17, Nv - 1):
        a = sorted(qlc[int(vbc[n]): int(vbc[n]) + int(vcc[n])].tolist())
        assert a == sorted(lstc[n, : int(vcc[n])].tolist())
    ops.clip_score_f32(qn, planes, ps, csr=(vb, ql, vc), scatter=(slot, s1, k1))
    ops.clip_score_list(qn, planes, ps, cnt, lst, s2, k2)
    assert torch.equal(s1, s2) and torch.equal(k1, k2)
    # and the re-resolved pairs carry the exact kernel's values
    se, ke = ops.clip_score_f32(qn, planes, ps)
    fl = torch.zeros((M, Nv), dtype=torch.bool)
    for n in range(Nv):
        fl[lstc[n, : int(vcc[n])].long(), n] = True
    assert torch.equal(s2.cpu()[fl], se.cpu()[fl]) and torch.equal(k2.cpu()[fl], ke.cpu()[fl])


@pytest.mark.parametrize("M,Nv,K", [(100, 2179, 128), (37, 500, 100), (9, 129, 128), (5, 60, 100), (300, 17432, 128)])
def test_select_topk_equals_topk_set(ops, M, Nv, K):
    """dkd_select_topk (radix select, unsorted) returns exactly the set dkd_topk ranks, including exact-score ties at
    the K-th place (lower id wins), and the K-th best score."""
    g = torch.Generator().manual_seed(M + Nv)
    s = torch.randn(M, Nv, generator=g)
    s[:, ::5] = 0.25                                     # many exact ties, also around the K-th score for some rows
    s[1] = 1.0                                           # a row of all-equal scores: the K lowest ids
    s[2, 7] = float("-inf")
    sc = s.cuda()
    ts, ti = ops.topk(sc, K, 1000)
    si, kth = ops.select_topk(sc, K, 1000)
    kk = min(K, Nv)
    assert torch.equal(torch.sort(si, dim=1).values[:, K - kk:], torch.sort(ti[:, :kk], dim=1).values)
    if Nv > K:
        assert torch.equal(kth, ts[:, K - 1])
        assert bool((si >= 1000).all())
    else:
        assert bool((si[:, kk:] == -1).all()) and bool(torch.isinf(kth).all())
