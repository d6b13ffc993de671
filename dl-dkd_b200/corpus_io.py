"""Packed corpus format + loaders: the data format on the input side of the scoring path.

The reference reads video features one FRAME at a time — `BigFile.read_one` re-opens `feature.bin` and seeks for
every frame id (utils/basic_utils.py:27-65, called per frame at method/data_provider.py:288-290) — then
resamples to max_ctx_l frames and L2-normalises per video (data_provider.py:52-73, :299-301).  At TVR size that is
279 k file opens before compute_context_info can start.  Here the same features live in ONE file that is mapped
once:

    offset 0      header  64 bytes: magic "DKDCORP1", version u32, dtype u32 (0 f32, 1 bf16, 2 f16), Nv u64, L u32,
                          D u32, planes u32, ids_bytes u64, data_offset u64
    64            lengths int32[Nv]
    64 + 4 Nv     ids     utf-8, '\\n' separated
    data_offset   data    [planes][Nv][L][D], zero padded beyond each video's length (4096-byte aligned)

A plane is one feature stream: raw visual features (1 plane, D = 3072 at TVR) feeding encode_context, or encoded
frames (2 planes: inheritance, exploration; D = 384) feeding engine.rank_streamed directly.

  write_packed / PackedCorpus           writer and zero-copy (np.memmap) reader
  pack_bigfile                          one sequential pass over a reference BigFile directory -> packed file,
                                        with the reference's own resampling + normalisation
  PackedVideoDataset                    torch Dataset yielding the reference's (feat, index, video_id) items
                                        (data_provider.py:309) — drops into compute_context_info
  device_chunks                         double-buffered pinned-memory H2D pipeline on a side stream yielding
                                        (frames_by_plane, mask, id_base) chunks for engine.rank_streamed
"""
import os
import struct

import numpy as np
import torch

MAGIC = b"DKDCORP1"
HEADER = struct.Struct("<8sIIQIIIQQ")   # magic, version, dtype, Nv, L, D, planes, ids_bytes, data_offset
HEADER_BYTES = 64
DTYPES = {"f32": (0, np.float32, 4), "bf16": (1, np.uint16, 2), "f16": (2, np.float16, 2)}
_BY_CODE = {v[0]: (k, v[1], v[2]) for k, v in DTYPES.items()}


def uniform_feature_sampling(features, max_len):
    """method/data_provider.py:52-68: mean-pool `features` (n, D) down to max_len rows when n > max_len."""
    n = features.shape[0]
    if max_len is None or n <= max_len:
        return features
    idxs = np.arange(0, max_len + 1, 1.0) / max_len * n
    idxs = np.round(idxs).astype(np.int32)
    idxs[idxs > n - 1] = n - 1
    out = np.empty((max_len, features.shape[1]), dtype=features.dtype)
    for i in range(max_len):
        s, e = idxs[i], idxs[i + 1]
        out[i] = np.mean(features[s:e], axis=0) if s < e else features[s]
    return out


def l2_normalize_rows(a, eps=1e-5):
    """method/data_provider.py:71-73."""
    return a / (np.linalg.norm(a, axis=-1, keepdims=True) + eps)


def _to_storage(x, dtype):
    """(…, D) float32 array/tensor -> numpy array in the file's storage type (bf16 as raw uint16, round-to-nearest-even)."""
    t = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x).detach().cpu().float()
    if dtype == "f32":
        return t.numpy()
    if dtype == "f16":
        return t.half().numpy()
    return t.bfloat16().view(torch.int16).numpy().view(np.uint16)


def write_packed(path, planes, lengths, ids=None, dtype="bf16"):
    """planes: list of (Nv, L, D) arrays/tensors (rows beyond lengths[n] are written as zeros); lengths (Nv,)."""
    if dtype not in DTYPES:
        raise ValueError(f"dtype must be one of {sorted(DTYPES)}")
    code, _, _ = DTYPES[dtype]
    Nv, L, D = planes[0].shape
    lengths = np.asarray(lengths, dtype=np.int32)
    if lengths.shape != (Nv,) or (Nv and (lengths.min() < 0 or lengths.max() > L)):
        raise ValueError("lengths must be (Nv,) with 0 <= length <= L")
    ids = [f"vid{n}" for n in range(Nv)] if ids is None else list(ids)
    if len(ids) != Nv or any("\n" in s for s in ids):
        raise ValueError("ids must be Nv strings without newlines")
    blob = "\n".join(ids).encode("utf-8")
    data_offset = (HEADER_BYTES + 4 * Nv + len(blob) + 4095) // 4096 * 4096
    valid = (np.arange(L)[None, :] < lengths[:, None])[:, :, None]
    with open(path, "wb") as f:
        f.write(HEADER.pack(MAGIC, 1, code, Nv, L, D, len(planes), len(blob), data_offset).ljust(HEADER_BYTES, b"\0"))
        f.write(lengths.tobytes())
        f.write(blob)
        f.write(b"\0" * (data_offset - f.tell()))
        for p in planes:
            if tuple(p.shape) != (Nv, L, D):
                raise ValueError("all planes must share the shape (Nv, L, D)")
            for lo in range(0, Nv, 1024):          # bounded staging: a plane can be larger than host memory
                blk = np.asarray(p[lo: lo + 1024].cpu() if isinstance(p, torch.Tensor) else p[lo: lo + 1024], dtype=np.float32)
                f.write(_to_storage(blk * valid[lo: lo + 1024], dtype).tobytes())
    return path


class PackedCorpus:
    """Zero-copy reader: the data section is one np.memmap; chunk() slices it without touching the rest."""

    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            head = f.read(HEADER_BYTES)
            if len(head) < HEADER_BYTES:
                raise ValueError(f"{path}: truncated header")
            magic, version, code, Nv, L, D, planes, ids_bytes, data_offset = HEADER.unpack(head[:HEADER.size])
            if magic != MAGIC or version != 1 or code not in _BY_CODE:
                raise ValueError(f"{path}: not a packed corpus (magic/version/dtype)")
            self.Nv, self.L, self.D, self.planes = int(Nv), int(L), int(D), int(planes)
            self.dtype, self._np_dtype, self._esize = _BY_CODE[code]
            self.lengths = np.frombuffer(f.read(4 * self.Nv), dtype=np.int32).copy()
            blob = f.read(ids_bytes).decode("utf-8")
            self.ids = blob.split("\n") if self.Nv else []
        need = data_offset + self.planes * self.Nv * self.L * self.D * self._esize
        if os.path.getsize(path) < need:
            raise ValueError(f"{path}: truncated data section ({os.path.getsize(path)} < {need} bytes)")
        shape = (self.planes, self.Nv, self.L, self.D)
        self.data = (np.memmap(path, dtype=self._np_dtype, mode="r", offset=data_offset, shape=shape)
                     if self.Nv else np.zeros(shape, self._np_dtype))

    def torch_dtype(self):
        return {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[self.dtype]

    def chunk(self, lo, hi):
        """Storage-typed views (planes, hi - lo, L, D) of videos [lo, hi)."""
        return self.data[:, lo:hi]

    def video(self, n, plane=0):
        """fp32 (length, D) tensor of one video (a copy)."""
        raw = np.ascontiguousarray(self.data[plane, n, : self.lengths[n]])
        raw = raw.copy() if not raw.flags.writeable else raw
        t = torch.from_numpy(raw.view(np.int16) if self.dtype == "bf16" else raw)
        return (t.view(torch.bfloat16) if self.dtype == "bf16" else t).float()

    def mask(self, lo=0, hi=None):
        hi = self.Nv if hi is None else hi
        return torch.from_numpy((np.arange(self.L)[None, :] < self.lengths[lo:hi, None]).astype(np.float32))


def pack_bigfile(feat_dir, video2frames, out_path, max_ctx_len=128, dtype="f32", order=None):
    """Reference BigFile directory (shape.txt 'rows dims', id.txt, feature.bin float32; utils/basic_utils.py:11-26)
    -> packed corpus, in one mapped pass.  video2frames: video id -> list of frame ids (the reference's
    video2frames dict, data_provider.py:271).  Per video: frames in list order, uniform_feature_sampling to
    max_ctx_len, l2_normalize (data_provider.py:299-301), computed in float64 and cast to float32 at the end — bit for bit
    what the reference's loader hands to cat_videos.  order: video ids to write (default: sorted keys)."""
    rows, dims = (int(x) for x in open(os.path.join(feat_dir, "shape.txt")).read().split())
    names = open(os.path.join(feat_dir, "id.txt"), encoding="ISO-8859-1").read().strip().split()
    if len(names) != rows:
        raise ValueError("id.txt does not match shape.txt")
    name2index = dict(zip(names, range(rows)))
    feats = np.memmap(os.path.join(feat_dir, "feature.bin"), dtype=np.float32, mode="r", shape=(rows, dims))
    order = sorted(video2frames) if order is None else list(order)
    Nv = len(order)
    plane = np.zeros((Nv, max_ctx_len, dims), dtype=np.float32)
    lengths = np.zeros((Nv,), dtype=np.int32)
    for n, vid in enumerate(order):
        idx = np.fromiter((name2index[f] for f in video2frames[vid]), dtype=np.int64)
        # float64 like the reference: BigFile.read_one returns Python floats, so np.array(...) of a video's frames is a
        # float64 array and uniform_feature_sampling + l2_normalize_np_array run in float64 (data_provider.py:286-301);
        # the cast to float32 happens last (cat_videos writes into a float32 tensor, :75-86)
        v = l2_normalize_rows(uniform_feature_sampling(np.asarray(feats[idx], dtype=np.float64), max_ctx_len))
        lengths[n] = v.shape[0]
        plane[n, : v.shape[0]] = v.astype(np.float32)
    return write_packed(out_path, [plane], lengths, order, dtype=dtype)


class PackedVideoDataset(torch.utils.data.Dataset):
    """Items (feat (length, D) fp32, index, video_id): what VisDataSet4MS.__getitem__ returns without a teacher
    (method/data_provider.py:309); collate with dkd_b200.eval.collate_frame_val."""

    def __init__(self, corpus, plane=0):
        self.corpus = corpus if isinstance(corpus, PackedCorpus) else PackedCorpus(corpus)
        self.plane = plane
        self.video_ids = self.corpus.ids

    def __len__(self):
        return self.corpus.Nv

    def __getitem__(self, index):
        return self.corpus.video(index, self.plane), index, self.corpus.ids[index]


def device_chunks(corpus, chunk_videos, device, lo=0, hi=None, id_base=None):
    """Generator of (frames_by_plane [fp32 (n, L, D) on `device`], mask (n, L), id_base) over videos [lo, hi):
    the chunk source of engine.rank_streamed for a corpus that lives on disk / in host memory.  Chunk i+1 is copied
    (mapped file -> pinned staging -> device, on a side stream) while the caller scores chunk i; two staging
    buffers and two device buffers are recycled, guarded by events.  A yielded chunk may alias a recycled buffer:
    it is valid until the next chunk is drawn (engine.rank_streamed consumes chunks that way)."""
    hi = corpus.Nv if hi is None else hi
    id_base = lo if id_base is None else id_base
    if hi <= lo:
        return
    device = torch.device(device)
    n_max = min(chunk_videos, hi - lo)
    st_dtype = torch.int16 if corpus.dtype == "bf16" else corpus.torch_dtype()
    shape = (corpus.planes, n_max, corpus.L, corpus.D)
    use_cuda = device.type == "cuda"
    staging = [torch.empty(shape, dtype=st_dtype, pin_memory=use_cuda) for _ in range(2)]
    dbuf = [torch.empty(shape, dtype=st_dtype, device=device) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device) if use_cuda else None
    staged = [None, None]        # event: H2D of buffer b finished
    released = [None, None]      # event: the consumer is done with device buffer b
    starts = list(range(lo, hi, chunk_videos))

    def issue(i):
        b = i % 2
        c_lo, c_hi = starts[i], min(starts[i] + chunk_videos, hi)
        n = c_hi - c_lo
        if use_cuda and staged[b] is not None:
            staged[b].synchronize()              # the previous H2D out of this staging buffer has drained
        src = corpus.chunk(c_lo, c_hi)
        src = src.view(np.int16) if corpus.dtype == "bf16" else src
        np.copyto(staging[b][:, :n].numpy(), src)              # mapped file -> pinned staging, one pass
        if use_cuda:
            with torch.cuda.stream(copy_stream):
                if released[b] is not None:
                    copy_stream.wait_event(released[b])
                dbuf[b][:, :n].copy_(staging[b][:, :n], non_blocking=True)
                staged[b] = torch.cuda.Event()
                staged[b].record(copy_stream)
        else:
            dbuf[b][:, :n].copy_(staging[b][:, :n])
        return n, c_lo

    pending = issue(0)
    for i in range(len(starts)):
        n, c_lo = pending
        b = i % 2
        if i + 1 < len(starts):
            pending = issue(i + 1)               # overlaps with the consumer's work on chunk i
        if use_cuda:
            torch.cuda.current_stream(device).wait_event(staged[b])
        raw = dbuf[b][:, :n]
        x = (raw.view(torch.bfloat16) if corpus.dtype == "bf16" else raw).float()     # f32 files: a view of dbuf[b]
        yield [x[p] for p in range(corpus.planes)], corpus.mask(c_lo, c_lo + n).to(device), id_base + (c_lo - lo)
        # resumed: everything the consumer enqueued on the current stream for this chunk precedes this event,
        # and only then may the copy stream overwrite device buffer b (a chunk is valid until the next one is drawn)
        if use_cuda:
            released[b] = torch.cuda.Event()
            released[b].record(torch.cuda.current_stream(device))
