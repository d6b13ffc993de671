"""Packed corpus format (dl-dkd_b200/corpus_io.py): round trips, the reference's resampling / normalisation
(pinned by tests/golden/ref_avg_fixed.npz), BigFile conversion against the reference's own reader, the Dataset
items compute_context_info consumes, and the chunk generator of engine.rank_streamed (CPU here, CUDA in -m gpu)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_shim
from tests import synth
from tests.test_oracle_golden import _load


@pytest.fixture(scope="module")
def cio(dkd):
    from dkd_b200 import corpus_io
    return corpus_io


def _planes(Nv=37, L=16, D=24, seed=1):
    a, mask, lengths = synth.encoded_corpus(Nv, L, D, seed=seed, min_len=1)
    b, _, _ = synth.encoded_corpus(Nv, L, D, seed=seed + 1, min_len=1)
    return [a, b * mask[:, :, None]], mask, lengths


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
def test_round_trip(cio, tmp_path, dtype):
    planes, mask, lengths = _planes()
    ids = [f"v_{n:03d}" for n in range(37)]
    # garbage beyond the valid length must not reach the file
    dirty = [p.clone() for p in planes]
    dirty[0][1, lengths[1]:] = 7.0
    path = cio.write_packed(str(tmp_path / "c.dkd"), dirty, lengths, ids, dtype=dtype)
    pc = cio.PackedCorpus(path)
    assert (pc.Nv, pc.L, pc.D, pc.planes, pc.dtype) == (37, 16, 24, 2, dtype)
    assert pc.ids == ids and np.array_equal(pc.lengths, lengths.numpy())
    assert torch.equal(pc.mask(), mask)
    cast = {"f32": lambda t: t, "bf16": lambda t: t.bfloat16().float(), "f16": lambda t: t.half().float()}[dtype]
    for n in (0, 1, 36):
        for p in range(2):
            assert torch.equal(pc.video(n, p), cast(planes[p][n, : lengths[n]]))
    got = [torch.cat([c[p].clone() for c, _, _ in cio.device_chunks(pc, 10, "cpu")]) for p in range(2)]
    for p in range(2):
        assert torch.equal(got[p], cast(planes[p]))
    bases = [b for _, _, b in cio.device_chunks(pc, 10, "cpu", lo=5, hi=30, id_base=1005)]
    assert bases == [1005, 1015, 1025]
    masks = torch.cat([m for _, m, _ in cio.device_chunks(pc, 10, "cpu", lo=5, hi=30)])
    assert torch.equal(masks, mask[5:30])
    assert list(cio.device_chunks(pc, 10, "cpu", lo=7, hi=7)) == []


def test_rejects_bad_files(cio, tmp_path):
    planes, _, lengths = _planes(Nv=3)
    path = cio.write_packed(str(tmp_path / "ok.dkd"), planes, lengths)
    raw = open(path, "rb").read()
    open(tmp_path / "magic.dkd", "wb").write(b"NOTACORP" + raw[8:])
    open(tmp_path / "short.dkd", "wb").write(raw[:-100])
    open(tmp_path / "tiny.dkd", "wb").write(raw[:20])
    for name in ("magic.dkd", "short.dkd", "tiny.dkd"):
        with pytest.raises(ValueError):
            cio.PackedCorpus(str(tmp_path / name))
    with pytest.raises(ValueError):
        cio.write_packed(str(tmp_path / "x.dkd"), planes, lengths + 100)
    with pytest.raises(ValueError):
        cio.write_packed(str(tmp_path / "x.dkd"), planes, lengths, dtype="int8")
    empty = cio.PackedCorpus(cio.write_packed(str(tmp_path / "e.dkd"), [torch.zeros(0, 4, 8)], np.zeros(0, np.int32)))
    assert empty.Nv == 0 and empty.ids == [] and list(cio.device_chunks(empty, 4, "cpu")) == []


def test_resampling_matches_reference_fixture(cio):
    g = _load("ref_avg_fixed.npz")
    assert np.array_equal(cio.uniform_feature_sampling(g["ufs_x"], 128).astype(np.float32), g["ufs_y"])
    assert np.array_equal(cio.l2_normalize_rows(g["ufs_x"]).astype(np.float32), g["l2_y"])
    short = g["ufs_x"][:50]
    assert cio.uniform_feature_sampling(short, 128) is short


def _write_bigfile(d, names, feats):
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, "shape.txt"), "w").write(f"{feats.shape[0]} {feats.shape[1]}")
    open(os.path.join(d, "id.txt"), "w").write(" ".join(names))
    feats.astype(np.float32).tofile(os.path.join(d, "feature.bin"))


def test_pack_bigfile_and_dataset_items(cio, dkd, tmp_path):
    """A synthetic BigFile directory (frame-level rows in shuffled order) -> packed file; every video equals
    what the reference's per-frame reader + preprocessing produces (live reference when present, else the same
    formulas); the Dataset items collate like the reference's."""
    rng = np.random.default_rng(5)
    lens = {"vidA": 7, "vidB": 200, "vidC": 128, "vidD": 1}
    v2f = {v: [f"{v}_{i}" for i in range(n)] for v, n in lens.items()}
    names = [f for fs in v2f.values() for f in fs]
    rng.shuffle(names)
    feats = rng.standard_normal((len(names), 12)).astype(np.float32)
    _write_bigfile(str(tmp_path / "bf"), names, feats)
    path = cio.pack_bigfile(str(tmp_path / "bf"), v2f, str(tmp_path / "c.dkd"), max_ctx_len=128, dtype="f32")
    pc = cio.PackedCorpus(path)
    assert pc.ids == sorted(v2f) and list(pc.lengths) == [7, 128, 128, 1]
    if ref_shim.available():
        ref_shim.load()
        from utils.basic_utils import BigFile
        from method.data_provider import uniform_feature_sampling, l2_normalize_np_array
        bf = BigFile(str(tmp_path / "bf"))
        for n, vid in enumerate(pc.ids):
            vecs = np.array([bf.read_one(f) for f in v2f[vid]])              # data_provider.py:286-290
            want = l2_normalize_np_array(uniform_feature_sampling(vecs, 128))
            assert want.dtype == np.float64                                  # the reference's loader works in float64
            assert np.array_equal(pc.video(n).numpy(), want.astype(np.float32)), vid   # bit for bit
    row = {f: i for i, f in enumerate(names)}
    want = cio.l2_normalize_rows(cio.uniform_feature_sampling(feats[[row[f] for f in v2f["vidB"]]].astype(np.float64), 128))
    assert np.array_equal(pc.video(1).numpy(), want.astype(np.float32))
    ds = cio.PackedVideoDataset(pc)
    feat, idx, vid = ds[2]
    assert (idx, vid) == (2, "vidC") and feat.shape == (128, 12) and feat.dtype == torch.float32
    from dkd_b200 import eval as E
    clip, m, idxs, vids = E.collate_frame_val([ds[i] for i in range(len(ds))])
    assert clip.shape == (4, 128, 12) and list(vids) == pc.ids and torch.equal(m, pc.mask())


@pytest.mark.gpu
def test_device_chunks_feed_rank_streamed(ops, cio, tmp_path):
    """Encoded corpus written as bf16, streamed from the file through the pinned double buffer == ranking the
    same bf16-rounded corpus resident on the device."""
    from dkd_b200 import engine
    Nv, L, D, M = 300, 32, 128, 64
    planes, mask, lengths = _planes(Nv, L, D, seed=9)
    pc = cio.PackedCorpus(cio.write_packed(str(tmp_path / "enc.dkd"), planes, lengths, dtype="bf16"))
    g = torch.Generator().manual_seed(3)
    params = [tuple(t.cuda() for t in (0.05 * torch.randn(D, D, generator=g), torch.zeros(D),
                                       0.05 * torch.randn(D, D, generator=g), torch.zeros(D))) for _ in range(2)]
    qs = [synth.encoded_queries(M, D, seed=4).cuda(), synth.encoded_queries(M, D, seed=5).cuda()]
    pqs = engine.split_queries(qs, 40)
    s1, i1 = engine.rank_streamed(cio.device_chunks(pc, 64, "cuda"), pqs, params, K=100, precision="exact")
    res = [p.bfloat16().float().cuda() for p in planes]
    s2, i2 = engine.rank_streamed(engine.iter_chunks(res, mask.cuda(), 64), pqs, params, K=100, precision="exact")
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
