"""Property tests (hypothesis) of the host-side integer logic: proposal indexing, shard ranges, the packed corpus
format.  CPU only."""
import os

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import oracle as O


@given(T=st.integers(1, 32))
@settings(max_examples=32, deadline=None)
def test_proposal_index_is_a_bijection_in_window_major_order(T):
    import __graft_entry__ as g
    g.load_package()
    from dkd_b200 import ops
    P = ops.num_proposals(T)
    seen = []
    for w in range(1, T + 1):
        for s in range(0, T - w + 1):
            seen.append(ops.proposal_index(w, s, T))
            assert seen[-1] == O.proposal_index(w, s, T)
    assert seen == list(range(P))                 # w-major, then start: exactly the order build_proposals writes


@given(Nv=st.integers(0, 5000), world=st.integers(1, 16))
@settings(max_examples=200, deadline=None)
def test_shard_ranges_partition_and_balance(Nv, world):
    import __graft_entry__ as g
    g.load_package()
    from dkd_b200 import engine
    edges = [engine.shard_range(Nv, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == Nv
    assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
    sizes = [hi - lo for lo, hi in edges]
    assert all(s >= 0 for s in sizes) and max(sizes) <= (Nv + world - 1) // world


@given(Nv=st.integers(0, 9), L=st.integers(1, 7), D=st.sampled_from([4, 8]), planes=st.integers(1, 3),
       dtype=st.sampled_from(["f32", "bf16", "f16"]), chunk=st.integers(1, 5), seed=st.integers(0, 2 ** 16))
@settings(max_examples=60, deadline=None)
def test_packed_corpus_round_trip(tmp_path_factory, Nv, L, D, planes, dtype, chunk, seed):
    import __graft_entry__ as g
    g.load_package()
    from dkd_b200 import corpus_io as cio
    gen = torch.Generator().manual_seed(seed)
    lengths = torch.randint(0, L + 1, (Nv,), generator=gen).numpy().astype(np.int32)
    data = [torch.randn(Nv, L, D, generator=gen) for _ in range(planes)]
    path = os.path.join(tmp_path_factory.mktemp("c"), "x.dkd")
    pc = cio.PackedCorpus(cio.write_packed(path, data, lengths, dtype=dtype))
    cast = {"f32": lambda t: t, "bf16": lambda t: t.bfloat16().float(), "f16": lambda t: t.half().float()}[dtype]
    valid = (torch.arange(L)[None] < torch.from_numpy(lengths)[:, None]).float()
    got = [[] for _ in range(planes)]
    bases, masks = [], []
    for fr, m, b in cio.device_chunks(pc, chunk, "cpu", id_base=100):
        for p in range(planes):
            got[p].append(fr[p].clone())
        bases.append(b)
        masks.append(m)
    assert bases == list(range(100, 100 + Nv, chunk))
    for p in range(planes):
        want = cast(data[p]) * valid[:, :, None]
        have = torch.cat(got[p]) if got[p] else torch.zeros(0, L, D)
        assert torch.equal(have, want)
    if Nv:
        assert torch.equal(torch.cat(masks), valid)
        n = int(seed % Nv)
        assert torch.equal(pc.video(n, planes - 1), cast(data[planes - 1][n, : lengths[n]]))
