"""Multi-GPU sharding of one haystack (SURVEY 8e / DESIGN 6).

Start windows are independent (src/search.rs:533: one BFS per start position), so a haystack is cut
into one contiguous byte range per rank; rank g searches the starts in its own range and reads a right
halo of `max_match_graphemes() + 1` graphemes (the reference's own streaming overlap,
src/stream.rs:213-258).  There is no data-path collective.  The only exchange is the gather of the
per-shard match lists (counts, then padded records) so that `FuzzyMatches::apply` (src/matches.rs:7)
can rank / de-overlap globally -- overlap components may straddle a cut.

Works over any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import ctypes as C

import numpy as np

from ._abi import fac_match

MAX_GRAPHEME_BYTES = 4  # halo in bytes for ASCII / BMP text; callers with longer clusters pass halo_bytes


def plan_shards(n_bytes, world, text=None):
    """[(own_begin, own_end)] per rank; cuts are snapped forward to a UTF-8 scalar boundary when `text` is given
    (ASCII text is never moved)."""
    cuts = [0]
    for r in range(1, world):
        c = (n_bytes * r) // world
        if text is not None:
            while c < n_bytes and (text[c] & 0xC0) == 0x80:
                c += 1
        cuts.append(max(c, cuts[-1]))
    cuts.append(n_bytes)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_slice(n_bytes, own, halo_graphemes, bytes_per_grapheme=1):
    """(read_begin, read_end): the owned range plus the right halo."""
    a, b = own
    return a, min(n_bytes, b + (halo_graphemes + 1) * bytes_per_grapheme)


def search_shard(engine, backend, text, own, threshold, on_device_ptr=None):
    """Matches whose start lies in `own`, offsets absolute.  `text` is the whole haystack (bytes / numpy uint8)."""
    n = len(text)
    a, end = shard_slice(n, own, engine.max_match_graphemes())
    if hasattr(backend, "search_shard"):
        buf = bytes(text[a:end]) if on_device_ptr is None else on_device_ptr
        arr, stats = backend.search_shard(engine._h, buf, end - a, 0, own[1] - a, a, threshold, on_device_ptr is not None)
        return arr, stats
    # CPU checker backends: plain search of the slice, ownership by start (src/stream.rs:262-297)
    arr, stats = backend.search(engine._h, bytes(text[a:end]), threshold, 0, 0, False)
    keep = [m for m in arr if m.start < own[1] - a]
    out = (fac_match * len(keep))()
    for i, m in enumerate(keep):
        C.memmove(C.byref(out[i]), C.byref(m), C.sizeof(fac_match))
        out[i].start += a
        out[i].end += a
    return out, stats


def gather_matches(arr, dist, device="cpu"):
    """All ranks contribute a fac_match array; rank 0 receives the concatenation (others get None).
    Two collectives: all_gather of the counts, all_gather of the records padded to the maximum."""
    import torch
    world = dist.get_world_size()
    n = len(arr)
    cnt = torch.tensor([n], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    rec = np.zeros(mx * 32, dtype=np.uint8)
    if n:
        rec[: n * 32] = np.frombuffer(arr, dtype=np.uint8, count=n * 32)
    t = torch.from_numpy(rec).to(device)
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    if dist.get_rank() != 0:
        return None
    total = sum(counts)
    out = (fac_match * total)()
    pos = 0
    for r in range(world):
        if counts[r]:
            b = parts[r][: counts[r] * 32].cpu().numpy().tobytes()
            C.memmove(C.byref(out, pos * 32), b, len(b))
            pos += counts[r]
    return out


def search_sharded(engine, backend, text, threshold, order, overlap, dist, device="cpu"):
    """engine.search(text, opts) computed by all ranks of `dist`; the result lands on rank 0."""
    rank, world = dist.get_rank(), dist.get_world_size()
    own = plan_shards(len(text), world, text)[rank]
    local, _ = search_shard(engine, backend, text, own, threshold)
    allm = gather_matches(local, dist, device)
    if rank != 0:
        return None
    final, _ = backend.apply(engine._h, allm, len(allm), order, overlap)
    return final
