"""The fused corpus-side encoder (encoder.FusedContextEncoder: tcgen05 kind::tf32 x 3 linears + LayerNorm /
attention kernels) against (a) the reference's own encoded features for the reference's weights
(tests/golden/ref_tiny_eval.npz, produced by the unmodified reference) and (b) the PyTorch mirror of the encoder at the
TVR dimensions.

Stated tolerance: 1e-4 absolute on encoded features of magnitude O(1).  The split-tf32 products are fp32 grade (2^-21
relative), but the tensor core's fp32 accumulator TRUNCATES: every tcgen05.mma costs up to one ulp of the running sum,
so a K = 3072 contraction (1,152 accumulation steps, |sum| ~ 1) is good to ~1e-4 where the K = 384 contractions of the
scoring path are good to 2e-6.  The fused encoder is therefore opt-in (model.enable_fused_encoder());
the default encode_context stays the PyTorch mirror, which reproduces the reference to 5e-6."""
import numpy as np
import pytest
import torch

from oracle import ref_shim
from tests.test_oracle_golden import _load, _tiny_model

pytestmark = pytest.mark.gpu
TOL = 1e-4


def test_linear_exact_and_layernorm_kernels(ops):
    g = torch.Generator().manual_seed(0)
    M, K, N = 1000, 3072, 768
    x = torch.randn(M, K, generator=g).cuda()
    w = (0.02 * torch.randn(N, K, generator=g)).cuda()
    b = torch.randn(N, generator=g).cuda()
    ref = torch.relu((x.double() @ w.double().T + b.double())).float()
    out = ops.linear_exact(x, ops.pack_weight(w), N, bias=b, relu=True)
    assert float((out - ref).abs().max()) <= 2e-4        # K = 3072, |out| ~ 1: accumulator truncation (see above)
    # LayerNorm prologue: (x - mean) * rstd applied inside the GEMM's stager
    ss = ops.row_stats(x, 1e-5)
    xh = torch.nn.functional.layer_norm(x.double(), (K,), eps=1e-5)
    ref2 = (xh @ w.double().T).float()
    out2 = ops.linear_exact(x, ops.pack_weight(w), N, row_scale_shift=ss)
    assert float((out2 - ref2).abs().max()) <= 2e-4
    # ragged sizes: M not a multiple of 128, N not a multiple of 128
    x3, w3 = x[:77, :96].contiguous(), w[:36, :96].contiguous()
    out3 = ops.linear_exact(x3, ops.pack_weight(w3), 36)
    assert float((out3 - (x3.double() @ w3.double().T).float()).abs().max()) <= 2e-6
    # layernorm with position rows and residual, input a column slice of a wider matrix
    D, L = 384, 128
    wide = torch.randn(512, 2 * D, generator=g).cuda()
    res = torch.randn(512, D, generator=g).cuda()
    pos = torch.randn(L, D, generator=g).cuda()
    gam, bet = torch.randn(D, generator=g).cuda(), torch.randn(D, generator=g).cuda()
    got = ops.layernorm_rows(wide[:, D:], gam, bet, 1e-5, residual=res, pos=pos, L=L)
    want = torch.nn.functional.layer_norm(wide[:, D:] + pos.repeat(4, 1) + res, (D,), gam, bet, 1e-5)
    assert float((got - want).abs().max()) <= 5e-6


def test_fused_encoder_reproduces_reference_fixture(ops, dkd):
    """The reference's weights and videos -> the reference's encoded frame features (valid frames)."""
    g = _load("ref_tiny_eval.npz")
    m = _tiny_model(dkd, g).cuda().enable_fused_encoder()
    videos = torch.from_numpy(g["videos"]).cuda()
    vlen = torch.from_numpy(g["video_len"])
    mask = (torch.arange(videos.shape[1])[None] < vlen[:, None]).float().cuda()
    with torch.no_grad():
        fi, fe = m.encode_context(videos, mask)
    valid = mask.bool().cpu().numpy()
    assert np.abs(fi.cpu().numpy()[valid] - g["inher_frame_feat"][valid]).max() <= TOL
    assert np.abs(fe.cpu().numpy()[valid] - g["explore_frame_feat"][valid]).max() <= TOL


def test_fused_encoder_matches_pytorch_mirror_at_tvr_dims(ops, dkd):
    from dkd_b200.model import DLDKD
    cfg = ref_shim.model_config(3072, 768, hidden=384, n_heads=4, max_ctx_l=128, max_desc_l=30)
    torch.manual_seed(0)
    m = DLDKD(cfg, ref_shim.options()).cuda().eval()
    with torch.no_grad():                       # non-trivial LayerNorm parameters and biases (reset_parameters: 1 / 0)
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(37, 128, 3072, device="cuda", generator=gen)
    x = x / (x.norm(dim=-1, keepdim=True) + 1e-5)
    lens = torch.randint(40, 129, (37,), device="cuda", generator=gen)
    mask = (torch.arange(128, device="cuda")[None] < lens[:, None]).float()
    with torch.no_grad():
        ri, re = m.encode_context(x, mask)
        m.enable_fused_encoder()
        fi, fe, kv = m._fused_encoder.encode_context(x, mask, want_key_val=True)
        assert float((fi - ri).abs().max()) <= TOL and float((fe - re).abs().max()) <= TOL
        fi2, _ = m.encode_context(x, None)                         # no mask
        m.enable_fused_encoder(False)
        ri2, _ = m.encode_context(x, None)
        assert float((fi2 - ri2).abs().max()) <= TOL
        kw, kb, vw, vb = m.attention_params()[1]
        assert float((kv[1][0] - torch.nn.functional.linear(re, kw, kb)).abs().max()) <= 2e-4
        assert float((kv[1][1] - torch.nn.functional.linear(re, vw, vb)).abs().max()) <= 2e-4
    with pytest.raises(RuntimeError):
        m.train()
        m.enable_fused_encoder()._fused_encoder.encode_context(x, mask)
