"""Multi-GPU sharding of one haystack (SURVEY 8e / DESIGN 6).

Start windows are independent (src/search.rs:533: one BFS per start position), so a haystack is cut
into one contiguous byte range per rank; rank g searches the starts in its own range and reads a right
halo of `max_match_graphemes() + 3` grapheme clusters (the reference's own streaming overlap,
src/stream.rs:213-258, plus the look-ahead of the dead-end filter).  Cuts and halo ends lie on
extended-grapheme-cluster boundaries of the WHOLE haystack (`fac_plan_shards`, a pure host function of
libfacgpu.so), and a shard of a non-ASCII haystack is searched with the Unicode grapheme storage even
when its own bytes are ASCII (FAC_TEXT_IS_UNICODE; `\\r\\n` is one grapheme there, src/grapheme.rs:92-98).

There is no data-path collective.  The only exchange is the gather of the per-shard match lists to rank 0:
an all_gather of the counts, then point-to-point sends of exactly count * 32 bytes into rank 0's buffer
(device memory over NCCL, host memory over gloo in the CPU tests).  Each shard list is ranked locally
(`order`), so for Order::Unsorted -- ascending (start, end, pattern) -- the concatenation in shard order
already is the global list; the other orders and the overlap modes (`FuzzyMatches::apply`,
src/matches.rs:7-149) are finished on rank 0 by `fac_matches_apply[_device]`.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import fac_match
from .api import SearchError


def _as_u8(text):
    if isinstance(text, np.ndarray):
        return np.ascontiguousarray(text, dtype=np.uint8)
    return np.frombuffer(bytes(text), dtype=np.uint8)


def plan_shards(max_match_graphemes, text, world, n_bytes=None):
    """[(own_begin, own_end, read_end)] per rank.  `text`: whole haystack (bytes / numpy uint8), or None for an
    ASCII haystack of `n_bytes` bytes that the caller does not hold in one piece."""
    lib = _abi.load_library()
    out = (_abi.fac_shard * world)()
    if text is None:
        ptr, n = None, int(n_bytes)
    else:
        arr = _as_u8(text)
        ptr, n = C.c_void_p(arr.ctypes.data), len(arr)
    st = lib.fac_plan_shards(int(max_match_graphemes), ptr, n, world, out)
    if st != 0:
        raise RuntimeError("fac_plan_shards failed: %s" % lib.fac_last_error_string().decode())
    return [(int(s.own_begin), int(s.own_end), int(s.read_end)) for s in out]


def is_ascii(text):
    arr = _as_u8(text)
    return not bool((arr & 0x80).any())


def search_shard(engine, backend, text, shard, threshold, order=0, whole_is_ascii=True, device_ptr=None, result_on_device=False):
    """Matches that start in shard = (own_begin, own_end, read_end) of `text`, absolute offsets, ranked by `order`.
    device_ptr: address of text[own_begin:read_end] already resident on the engine's device."""
    a, b, end = shard
    flags = 0 if whole_is_ascii else _abi.FAC_TEXT_IS_UNICODE
    if hasattr(backend, "search_ex"):
        if result_on_device:
            flags |= _abi.FAC_RESULT_ON_DEVICE
        if device_ptr is not None:
            return backend.search_ex(engine._h, device_ptr, end - a, 0, b - a, a, threshold, order, 0, flags | _abi.FAC_HAYSTACK_ON_DEVICE)
        sl = np.ascontiguousarray(_as_u8(text)[a:end])
        return backend.search_ex(engine._h, sl.ctypes.data, end - a, 0, b - a, a, threshold, order, 0, flags)
    # CPU checker backends: plain search of the slice, ownership by start (src/stream.rs:262-297).  A trailing
    # non-ASCII scalar behind the halo keeps an ASCII slice of a non-ASCII haystack on the Unicode storage.
    sl = bytes(_as_u8(text)[a:end])
    if not whole_is_ascii and is_ascii(sl):
        sl += " é".encode("utf-8")
    arr, stats = backend.search(engine._h, sl, threshold, order, 0, False)
    keep = [m for m in arr if m.start < b - a]
    out = (fac_match * len(keep))()
    for i, m in enumerate(keep):
        C.memmove(C.byref(out[i]), C.byref(m), C.sizeof(fac_match))
        out[i].start += a
        out[i].end += a
    return out, stats


def records_tensor(arr, device="cpu"):
    """fac_match ctypes array (or DeviceMatches) -> uint8 tensor [n * 32] on `device`."""
    import torch
    if hasattr(arr, "as_tensor"):
        return arr.as_tensor(device)
    n = len(arr)
    if n == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    host = torch.from_numpy(np.frombuffer(arr, dtype=np.uint8, count=n * 32).copy())
    return host.to(device)


def gather_records(local, dist, device="cpu", out=None):
    """Gather every rank's uint8 record tensor to rank 0 with exact sizes.  Returns (buffer, counts) on rank 0
    -- buffer[: sum(counts) * 32] is the concatenation in rank order -- and (None, counts) elsewhere.
    `out`: optional preallocated uint8 tensor on rank 0 that is reused when large enough."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    cnt = torch.tensor([local.numel() // 32], dtype=torch.int64, device=device)
    allc = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, cnt)
    counts = [int(c) for c in allc.cpu().tolist()]
    total = sum(counts)
    if rank == 0:
        buf = out if out is not None and out.numel() >= total * 32 else torch.empty(max(total, 1) * 32, dtype=torch.uint8, device=device)
        ops, pos = [], 0
        for r in range(world):
            nb = counts[r] * 32
            if r == 0:
                buf[:nb].copy_(local)
            elif nb:
                ops.append(dist.P2POp(dist.irecv, buf[pos:pos + nb], r))
            pos += nb
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return buf, counts
    if local.numel():
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, local, 0)]):
            w.wait()
    return None, counts


def search_sharded(engine, backend, text, threshold, order, overlap, dist, device="cpu"):
    """engine.search(text, opts) computed by all ranks of `dist`; the result (fac_match array) lands on rank 0."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if world > 1 and getattr(engine, "_has_auto_beam", False):
        # the auto_beam budget accumulates over all start windows of the call (src/search.rs:1096-1103)
        raise SearchError(_abi.FAC_UNSUPPORTED, "auto_beam engines cannot be sharded: the state budget is cumulative over the whole haystack")
    arr = _as_u8(text)
    ascii_ = is_ascii(arr)
    shard = plan_shards(engine.max_match_graphemes(), arr, world)[rank]
    on_gpu = hasattr(backend, "search_ex") and str(device) != "cpu"
    local, _ = search_shard(engine, backend, arr, shard, threshold, order, ascii_, result_on_device=on_gpu)
    buf, counts = gather_records(records_tensor(local, device), dist, device)
    if rank != 0:
        return None
    total = sum(counts)
    presorted = _abi.FAC_APPLY_PRESORTED if (order == 0 or world == 1) else 0
    if on_gpu:
        final, _ = backend.apply_device(engine._h, buf.data_ptr(), total, order, overlap, presorted)
        return final
    host = (fac_match * total)()
    if total:
        C.memmove(host, buf[: total * 32].cpu().numpy().ctypes.data, total * 32)
    final, _ = backend.apply(engine._h, host, total, order, overlap)
    return final
