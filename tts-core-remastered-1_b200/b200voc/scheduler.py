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


# ------------------------------------------------------------------ multi-GPU (one process per GPU)
def sharded_synthesize(synth: Synth, mels: Sequence[torch.Tensor], prosodies: Sequence[torch.Tensor],
                       styles: Sequence[torch.Tensor], emotions: Sequence[torch.Tensor], max_batch: int = 16,
                       gather_to: Optional[int] = None, group=None, **kw):
    """Every rank holds the full (host) work list, synthesizes only its shard, no data-path
    collective.  Returns {index: wav} for the local shard; if `gather_to` is a rank, that rank
    additionally receives every waveform (one gather at the end: the only communication)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lengths = [int(m.shape[-1]) for m in mels]
    mine = plan_shards(lengths, world)[rank]
    wavs = synthesize_batch(synth, [mels[i] for i in mine], [prosodies[i] for i in mine],
                            [styles[i] for i in mine], [emotions[i] for i in mine], max_batch=max_batch, **kw)
    local = {i: w for i, w in zip(mine, wavs)}
    if gather_to is None or world == 1:
        return local
    return gather_waveforms(local, lengths, gather_to, group=group)


def gather_waveforms(local: Dict[int, torch.Tensor], lengths: Sequence[int], dst: int, hop: int = 256, group=None):
    """The single final collective: all ranks contribute their shard, `dst` gets {index: wav} for
    every utterance (other ranks get their local dict back).  Implemented as one all_gather of a
    padded [n_max, L_max] buffer per rank (NCCL and gloo both support it)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    shards = plan_shards(lengths, world)
    n_max = max(len(s) for s in shards)
    l_max = hop * max(lengths) if lengths else 0
    dev = next(iter(local.values())).device if local else torch.device("cpu")
    buf = torch.zeros(max(n_max, 1), l_max, device=dev, dtype=torch.float32)
    for k, i in enumerate(shards[rank]):
        buf[k, :hop * lengths[i]] = local[i].reshape(-1)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    if rank != dst:
        return local
    full = {}
    for r in range(world):
        for k, i in enumerate(shards[r]):
            full[i] = parts[r][k, :hop * lengths[i]].reshape(1, -1).clone()
    return full
