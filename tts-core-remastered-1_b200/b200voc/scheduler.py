"""Batch scheduler for waveform synthesis (north_star item 3).

Utterances are independent units (no op in vocoder7/generator.py:50-98 mixes batch elements), so
the work shards across GPUs with NO collective in the math; the only communication is one optional
gather of the finished waveforms to a single rank.  Long utterances are additionally cut along
time into (chunk + halo) units that are also independent: the receptive field of the convolutional
stack is < 6 mel frames per side, so chunk + halo >= 6 + discard reproduces
the full-utterance output exactly (with the attention layer enabled the chunked result is the
block-local evaluation of that layer).

Everything here is host logic around a `synth` callable with the Generator.forward signature, so it
is unit-tested on CPU (gloo, world_size 2) with a stand-in callable and runs unchanged on NCCL.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

Synth = Callable[[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor], torch.Tensor]


# ------------------------------------------------------------------ sharding plan (pure host logic)
def plan_shards(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-first assignment of utterance indices to ranks, balancing the sum of frames
    (work is linear in T).  Deterministic; every index appears exactly once."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(lengths[i])
    for s in shards:
        s.sort()
    return shards


def group_by_length(lengths: Sequence[int], max_batch: int) -> List[List[int]]:
    """Batches of utterances with IDENTICAL frame counts (padding an utterance with zero mel frames
    is not equivalent to the convolutions' zero padding, so exact results need exact lengths)."""
    by_len: Dict[int, List[int]] = {}
    for i, t in enumerate(lengths):
        by_len.setdefault(int(t), []).append(i)
    batches = []
    for t in sorted(by_len, reverse=True):
        idx = by_len[t]
        for s in range(0, len(idx), max_batch):
            batches.append(idx[s:s + max_batch])
    return batches


def chunk_plan(T: int, chunk_frames: int, halo: int) -> List[Tuple[int, int, int, int]]:
    """[(in_start, in_end, keep_start, keep_end)] in frames: synthesize [in_start, in_end), keep the
    samples of frames [keep_start, keep_end).  Consecutive keep ranges tile [0, T)."""
    if chunk_frames <= 0 or halo < 0:
        raise ValueError("chunk_frames must be > 0 and halo >= 0")
    plan = []
    s = 0
    while s < T:
        e = min(T, s + chunk_frames)
        plan.append((max(0, s - halo), min(T, e + halo), s, e))
        s = e
    return plan


# ------------------------------------------------------------------ single-device drivers
def synthesize_batch(synth: Synth, mels: Sequence[torch.Tensor], prosodies: Sequence[torch.Tensor],
                     styles: Sequence[torch.Tensor], emotions: Sequence[torch.Tensor], max_batch: int = 16,
                     **kw) -> List[torch.Tensor]:
    """Ragged list of utterances (mel[i]: [80, T_i], prosody[i]: [T_i, 18], style[i]: [128],
    emotion[i]: [6]) -> list of waveforms [1, hop*T_i], batched by identical length."""
    n = len(mels)
    if not (len(prosodies) == len(styles) == len(emotions) == n):
        raise ValueError("mels / prosodies / styles / emotions must have the same length")
    out: List[Optional[torch.Tensor]] = [None] * n
    for idx in group_by_length([m.shape[-1] for m in mels], max_batch):
        wav = synth(torch.stack([mels[i] for i in idx]), torch.stack([prosodies[i] for i in idx]),
                    torch.stack([styles[i] for i in idx]), torch.stack([emotions[i] for i in idx]), **kw)
        for k, i in enumerate(idx):
            out[i] = wav[k].clone()
    return out  # type: ignore[return-value]


def synthesize_long(synth: Synth, mel: torch.Tensor, prosody: torch.Tensor, style: torch.Tensor,
                    emotion: torch.Tensor, chunk_frames: int = 512, halo: int = 8, hop: int = 256,
                    max_batch: int = 16, **kw) -> torch.Tensor:
    """Long-form synthesis: mel[B, 80, T] is cut into chunk_frames-frame chunks with a `halo`-frame
    context on each side, chunks of equal shape are batched together, the halo samples are
    discarded and the kept pieces are concatenated (overlap-discard stitching; any cross-fade over
    the halo would be a no-op because both sides are exact there).  Returns [B, 1, hop*T]."""
    B, _, T = mel.shape
    plan = chunk_plan(T, chunk_frames, halo)
    out = torch.empty(B, 1, hop * T, device=mel.device, dtype=torch.float32)
    # group chunks by (input length, left context) so that stacked chunks share one kernel launch set
    groups: Dict[Tuple[int, int], List[Tuple[int, int, int, int]]] = {}
    for c in plan:
        groups.setdefault((c[1] - c[0], c[2] - c[0]), []).append(c)
    for (_, _), chunks in groups.items():
        per_call = max(1, max_batch // B)
        for s in range(0, len(chunks), per_call):
            part = chunks[s:s + per_call]
            m = torch.cat([mel[:, :, a:b] for (a, b, _, _) in part], 0)
            p = torch.cat([prosody[:, a:b] for (a, b, _, _) in part], 0)
            st = style.repeat(len(part), 1)
            em = emotion.repeat(len(part), 1)
            wav = synth(m, p, st, em, **kw)
            for k, (a, b, ks, ke) in enumerate(part):
                piece = wav[k * B:(k + 1) * B, :, (ks - a) * hop:(ke - a) * hop]
                out[:, :, ks * hop:ke * hop] = piece
    return out


# ------------------------------------------------------------------ host-to-host streaming loop
class StreamingSynthesizer:
    """Host-in / host-out synthesis of a stream of equally shaped batches, software pipelined over
    three CUDA streams: the upload of batch i+1 (pinned host -> device) and the download of batch i-1
    (device -> pinned host) overlap the kernels of batch i.  Device input / output tensors are
    double buffered (`depth` slots, allocated once), so a serving loop's steady-state step time is
    the kernel time, not kernel + PCIe time.  `synth` must accept ``out=`` (Generator.forward does)
    and launch on the current stream.  On a CPU device (unit tests with a stand-in callable) the
    loop degenerates to the sequential one -- same results, no streams."""

    def __init__(self, synth: Synth, device: torch.device, depth: int = 2, **kw):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.synth, self.device, self.depth, self.kw = synth, torch.device(device), depth, kw
        self._slots = None
        self._shape_key = None

    def _alloc(self, batch, out_like: torch.Tensor):
        key = tuple((tuple(t.shape), t.dtype) for t in batch) + ((tuple(out_like.shape), out_like.dtype),)
        if self._slots is not None and key == self._shape_key:
            return
        self._shape_key = key
        self._slots = [dict(inp=[torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in batch],
                            out=torch.empty(out_like.shape, dtype=out_like.dtype, device=self.device))
                       for _ in range(self.depth)]

    def run(self, batches: Sequence[Sequence[torch.Tensor]], host_outs: Sequence[torch.Tensor]) -> None:
        """batches[i] = (mel, prosody, style, emotion) host tensors (pinned for true overlap);
        host_outs[i] = preallocated host tensor that receives batch i's waveforms.  Returns when every
        waveform has landed in host memory."""
        if len(batches) != len(host_outs):
            raise ValueError("one output buffer per batch")
        if not batches:
            return
        if self.device.type != "cuda":
            for b, o in zip(batches, host_outs):
                o.copy_(self.synth(*[t.to(self.device) for t in b], **self.kw))
            return
        self._alloc(batches[0], host_outs[0])
        main = torch.cuda.current_stream(self.device)
        s_in, s_out = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        ev_in = [torch.cuda.Event() for _ in batches]
        ev_comp = [torch.cuda.Event() for _ in batches]
        ev_out = [torch.cuda.Event() for _ in batches]
        start = torch.cuda.Event()
        start.record(main)
        s_in.wait_event(start)
        s_out.wait_event(start)
        for i, (b, o) in enumerate(zip(batches, host_outs)):
            slot = self._slots[i % self.depth]
            with torch.cuda.stream(s_in):
                if i >= self.depth:
                    s_in.wait_event(ev_comp[i - self.depth])      # the slot's inputs have been consumed
                for d, h in zip(slot["inp"], b):
                    d.copy_(h, non_blocking=True)
                ev_in[i].record(s_in)
            main.wait_event(ev_in[i])
            if i >= self.depth:
                main.wait_event(ev_out[i - self.depth])          # the slot's output has been downloaded
            self.synth(*slot["inp"], out=slot["out"], **self.kw)
            ev_comp[i].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_comp[i])
                o.copy_(slot["out"], non_blocking=True)
                ev_out[i].record(s_out)
        main.wait_event(ev_out[-1])          # later work on the caller's stream is ordered after the downloads
        ev_out[-1].synchronize()             # ... and the host may read host_outs as soon as run() returns


class GraphedSynthesizer:
    """Small requests are launch bound: a B = 1, T = 172 forward (BASELINE configs[0]) is ~0.1 ms of tensor work
    behind ~16 kernel launches and the host-side argument checks.  This wrapper captures ONE CUDA graph per exact
    request shape (B, T, out dtype, mel layout) -- the forward never allocates or synchronises after its first call
    -- and replays it on static device buffers: a call is 4 small device copies + one graph launch.  Shapes are
    never padded to a bucket (zero mel frames are not the convolution's zero padding, see ``group_by_length``): a
    serving loop sees few distinct shapes (fixed chunk + halo units of ``synthesize_long``, fixed batch sizes), and
    the least recently used graph is dropped beyond ``max_graphs``.  The result tensor is the graph's static output
    buffer: consume or copy it before the next call with the same shape.  Flag arguments (style_drop, ...) are part
    of the key because they are baked into the captured launches."""

    def __init__(self, gen, max_graphs: int = 16, warmup: int = 2):
        self.gen, self.max_graphs, self.warmup = gen, int(max_graphs), int(warmup)
        self._graphs: Dict[tuple, dict] = {}
        self.captures = 0

    def _capture(self, key, mel, prosody, style, emotion, kw) -> dict:
        dev = mel.device
        static = [torch.empty(t.shape, dtype=torch.float32, device=dev) for t in (mel, prosody, style, emotion)]
        for d, t in zip(static, (mel, prosody, style, emotion)):
            d.copy_(t)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        out = None
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(1, self.warmup)):          # packing, function attributes, workspace: outside the capture
                out = self.gen(*static, **kw)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(graph):
            self.gen(*static, out=out, **kw)
        self.captures += 1
        # the captured launches hold raw pointers into the generator's workspace: keep that tensor alive with the graph
        # (a later, larger request makes the generator allocate a new one; this graph keeps using its own)
        return dict(graph=graph, static=static, out=out, ws=getattr(self.gen, "_workspace", None))

    def __call__(self, mel, prosody, style, emotion, **kw) -> torch.Tensor:
        if not mel.is_cuda:
            raise ValueError("GraphedSynthesizer needs CUDA tensors (there is no CPU path)")
        if "out" in kw or "frame_lengths" in kw or "_tap" in kw:
            raise ValueError("GraphedSynthesizer owns the output buffer; out=/frame_lengths=/_tap= are not supported")
        key = (tuple(mel.shape), tuple(prosody.shape), mel.device.index, tuple(sorted(kw.items())))
        g = self._graphs.pop(key, None)
        if g is None:
            if len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))          # least recently used
            g = self._capture(key, mel, prosody, style, emotion, kw)
        self._graphs[key] = g                                        # most recently used last
        for d, t in zip(g["static"], (mel, prosody, style, emotion)):
            d.copy_(t, non_blocking=True)
        g["graph"].replay()
        return g["out"]


# ------------------------------------------------------------------ multi-GPU (one process per GPU)
def _dist_info(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _comm_device(group=None) -> torch.device:
    """Device the collective's buffers must live on: the current CUDA device under NCCL (also for a rank whose
    shard is empty and therefore has no tensor to infer it from), the host under gloo."""
    import torch.distributed as dist
    if dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_flat(local_flat: torch.Tensor, sizes: Sequence[int], dst: int, group=None) -> Optional[List[torch.Tensor]]:
    """The single final collective of a sharded job, as a true gather: rank r contributes a flat fp32 buffer of
    sizes[r] elements (sizes are known to every rank from the plan: nothing is padded, nothing is exchanged to
    agree on them) and only `dst` receives -- point-to-point sends batched into one NCCL group (gloo: the same
    calls).  Returns the list of per-rank buffers on dst, None elsewhere."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = _comm_device(group)
    if int(local_flat.numel()) != int(sizes[rank]):
        raise ValueError(f"rank {rank} contributes {local_flat.numel()} elements, the plan says {sizes[rank]}")
    if rank == dst:
        parts = [local_flat.to(dev) if r == dst else torch.empty(int(sizes[r]), device=dev, dtype=torch.float32)
                 for r in range(world)]
        ops = [dist.P2POp(dist.irecv, parts[r], r, group) for r in range(world) if r != dst and sizes[r] > 0]
    else:
        parts = None
        send = local_flat.to(dev).contiguous()
        ops = [dist.P2POp(dist.isend, send, dst, group)] if sizes[rank] > 0 else []
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return parts


def sharded_synthesize(synth: Synth, mels: Sequence[torch.Tensor], prosodies: Sequence[torch.Tensor],
                       styles: Sequence[torch.Tensor], emotions: Sequence[torch.Tensor], max_batch: int = 16,
                       gather_to: Optional[int] = None, group=None, **kw):
    """Every rank holds the full (host) work list, synthesizes only its shard, no data-path
    collective.  Returns {index: wav} for the local shard; if `gather_to` is a rank, that rank
    additionally receives every waveform (one gather at the end: the only communication)."""
    rank, world = _dist_info(group)
    lengths = [int(m.shape[-1]) for m in mels]
    mine = plan_shards(lengths, world)[rank]
    wavs = synthesize_batch(synth, [mels[i] for i in mine], [prosodies[i] for i in mine],
                            [styles[i] for i in mine], [emotions[i] for i in mine], max_batch=max_batch, **kw)
    local = {i: w for i, w in zip(mine, wavs)}
    if gather_to is None or world == 1:
        return local
    return gather_waveforms(local, lengths, gather_to, group=group)


def gather_waveforms(local: Dict[int, torch.Tensor], lengths: Sequence[int], dst: int, hop: int = 256, group=None):
    """The single final collective: `dst` gets {index: wav} for every utterance, other ranks get their local dict
    back.  A gather (only dst receives), ragged (every waveform travels at its own length)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    shards = plan_shards(lengths, world)
    sizes = [sum(hop * int(lengths[i]) for i in sh) for sh in shards]
    dev = _comm_device(group)
    flat = (torch.cat([local[i].reshape(-1).to(dev) for i in shards[rank]]) if shards[rank]
            else torch.empty(0, device=dev, dtype=torch.float32))
    parts = gather_flat(flat, sizes, dst, group=group)
    if rank != dst:
        return local
    full = {}
    for r in range(world):
        off = 0
        for i in shards[r]:
            n = hop * int(lengths[i])
            full[i] = parts[r][off:off + n].reshape(1, -1).clone()
            off += n
    return full


def sharded_synthesize_streaming(synth: Synth, device, mels: torch.Tensor, prosodies: torch.Tensor, styles: torch.Tensor,
                                 emotions: torch.Tensor, max_batch: int = 16, hop: int = 256, group=None,
                                 out: Optional[torch.Tensor] = None, **kw):
    """BASELINE configs[3]: n equally long utterances given as HOST tensors (mels[n, 80, T], ..., ideally pinned) are
    sharded over the ranks (contiguous blocks: equal lengths need no balancing) and each rank pushes its shard through
    the host-in / host-out StreamingSynthesizer in batches of `max_batch`.  No collective.  Returns (first index,
    host waveforms [n_local, 1, hop * T]) of this rank's shard; `out` (pinned, that shape) is reused when given -- a
    serving loop allocates it once, pinning 100s of MB costs as much as synthesizing them."""
    rank, world = _dist_info(group)
    n, T = int(mels.shape[0]), int(mels.shape[-1])
    per = (n + world - 1) // world
    lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
    if out is None:
        out = torch.empty(hi - lo, 1, hop * T)
        if torch.device(device).type == "cuda":
            out = out.pin_memory()
    elif tuple(out.shape) != (hi - lo, 1, hop * T):
        raise ValueError(f"out must have shape {(hi - lo, 1, hop * T)}, got {tuple(out.shape)}")
    starts = list(range(lo, hi, max_batch))
    # a ragged last batch has its own shape: the streamer re-allocates its slots for it (one extra allocation per job)
    full = [s for s in starts if min(hi, s + max_batch) - s == max_batch]
    tail = [s for s in starts if s not in full]
    streamer = StreamingSynthesizer(synth, device, depth=2, **kw)
    for group_starts in (full, tail):
        if group_starts:
            streamer.run([(mels[s:min(hi, s + max_batch)], prosodies[s:min(hi, s + max_batch)],
                           styles[s:min(hi, s + max_batch)], emotions[s:min(hi, s + max_batch)]) for s in group_starts],
                         [out[s - lo:min(hi, s + max_batch) - lo] for s in group_starts])
    return lo, out


# ------------------------------------------------------------------ long-form synthesis over several GPUs
def long_units(B: int, T: int, chunk_frames: int, halo: int) -> List[Tuple[int, int, int, int, int]]:
    """Work units of a long-form job: (utterance, in_start, in_end, keep_start, keep_end) in frames, utterance-major.
    Units are independent (chunk + halo), so one utterance may be spread over several GPUs."""
    plan = chunk_plan(T, chunk_frames, halo)
    return [(b, a, e, ks, ke) for b in range(B) for (a, e, ks, ke) in plan]


def synthesize_units(synth: Synth, units: Sequence[Tuple[int, int, int, int, int]], mel: torch.Tensor,
                     prosody: torch.Tensor, style: torch.Tensor, emotion: torch.Tensor, hop: int = 256,
                     max_batch: int = 16, **kw) -> List[torch.Tensor]:
    """Synthesizes the given (chunk + halo) units -- batched by identical input shape and left context, like
    synthesize_long -- and returns the KEPT piece of each, [hop * (keep_end - keep_start)] samples, in `units` order."""
    out: List[Optional[torch.Tensor]] = [None] * len(units)
    groups: Dict[Tuple[int, int], List[int]] = {}
    for k, (b, a, e, ks, ke) in enumerate(units):
        groups.setdefault((e - a, ks - a), []).append(k)
    for (_, _), idx in groups.items():
        for s in range(0, len(idx), max_batch):
            part = idx[s:s + max_batch]
            m = torch.stack([mel[units[k][0], :, units[k][1]:units[k][2]] for k in part])
            p = torch.stack([prosody[units[k][0], units[k][1]:units[k][2]] for k in part])
            st = torch.stack([style[units[k][0]] for k in part])
            em = torch.stack([emotion[units[k][0]] for k in part])
            wav = synth(m, p, st, em, **kw)
            for j, k in enumerate(part):
                _, a, _, ks, ke = units[k]
                out[k] = wav[j, 0, (ks - a) * hop:(ke - a) * hop].clone()
    return out  # type: ignore[return-value]


def sharded_synthesize_long(synth: Synth, mel: torch.Tensor, prosody: torch.Tensor, style: torch.Tensor,
                            emotion: torch.Tensor, chunk_frames: int = 512, halo: int = 8, hop: int = 256,
                            max_batch: int = 16, gather_to: Optional[int] = None, group=None, **kw):
    """BASELINE configs[4]: long-form batch mel[B, 80, T] cut into (chunk + halo) units that are spread over the ranks
    (greedy by input frames; a 60 s utterance spans GPUs), no collective in the math.  Every rank returns
    (units, pieces) of its shard; with `gather_to`, that rank instead returns the stitched [B, 1, hop * T] waveforms
    (one ragged gather, then overlap-discard stitching: the kept pieces tile each utterance exactly).  The result is
    the same samples synthesize_long produces on one device (units are batch-independent)."""
    rank, world = _dist_info(group)
    B, _, T = mel.shape
    units = long_units(B, T, chunk_frames, halo)
    shards = plan_shards([u[2] - u[1] for u in units], world)
    mine = [units[k] for k in shards[rank]]
    pieces = synthesize_units(synth, mine, mel, prosody, style, emotion, hop=hop, max_batch=max_batch, **kw)
    if gather_to is None:
        return mine, pieces
    sizes = [sum(hop * (units[k][4] - units[k][3]) for k in sh) for sh in shards]
    if world == 1:
        parts = [torch.cat([x.reshape(-1) for x in pieces])] if pieces else [torch.empty(0)]
    else:
        dev = _comm_device(group)
        flat = torch.cat([x.reshape(-1).to(dev) for x in pieces]) if pieces else torch.empty(0, device=dev)
        parts = gather_flat(flat, sizes, gather_to, group=group)
    if rank != gather_to:
        return mine, pieces
    out = torch.empty(B, 1, hop * T, device=parts[0].device, dtype=torch.float32)
    for r in range(world):
        off = 0
        for k in shards[r]:
            b, _, _, ks, ke = units[k]
            n = hop * (ke - ks)
            out[b, 0, ks * hop:ke * hop] = parts[r][off:off + n]
            off += n
    return out
