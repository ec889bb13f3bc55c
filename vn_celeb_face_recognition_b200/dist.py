"""Multi-GPU plumbing: one process per GPU, frames / crops sharded by index, every rank runs the whole pipeline on its
slice.  The only collective on the path is the ragged all-gather of the per-face results
(embeddings + labels + probabilities) at the end of a step -- NCCL over NVLink on GPUs, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous block partition: item i belongs to rank i // ceil(n/world)."""
    per = (n_items + world - 1) // world
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def all_gather_faces(emb, label, prob, group=None):
    """emb (n_i, D) float, label (n_i,) int64, prob (n_i,) float on this rank -> concatenation over ranks in rank
    order (= single-GPU order for a contiguous shard), plus the per-rank counts.  Two collectives: counts, then ONE
    padded payload (label and prob ride in two extra float columns... kept exact: labels < 2^24)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return emb, label, prob, torch.tensor([emb.shape[0]], device=emb.device)
    dev = emb.device
    n = torch.tensor([emb.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = torch.cat(counts)
    nmax = int(counts.max().item())
    D = emb.shape[1]
    payload = torch.zeros(nmax, D + 2, dtype=torch.float32, device=dev)
    payload[:emb.shape[0], :D] = emb.float()
    payload[:emb.shape[0], D] = label.float()
    payload[:emb.shape[0], D + 1] = prob.float()
    gathered = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    parts = [g[:int(c)] for g, c in zip(gathered, counts.tolist())]
    allp = torch.cat(parts, 0)
    return allp[:, :D].contiguous(), allp[:, D].long(), allp[:, D + 1].contiguous(), counts


def all_gather_faces_padded(emb, label, prob, cap, group=None, stream=None):
    """The same exchange without any host synchronisation: ONE collective over a fixed-capacity payload, so the host
    keeps enqueueing the next step while this one drains (the ragged variant above reads the counts back twice).
    ``cap`` = rows reserved per rank, identical on every rank (e.g. frames_per_rank * max_faces_per_frame).
    Returns (payload (world, cap + 1, D + 2) fp32 on device, D): rank r's faces are payload[r, :n_r] with columns
    [:D] embedding, [D] label, [D + 1] probability; n_r rides in payload[r, cap, 0] (exact: counts < 2^24).
    ``compact_faces`` turns it into the ragged concatenation when the consumer needs it on the host side.
    ``stream`` (CUDA only): run the exchange on that side stream, ordered after what is enqueued on the current stream so
    far; the current stream does NOT wait for it, so the next batch's kernels are not held up by the collective (which also
    waits for the slowest rank).  The result is then valid on ``stream``: make the consumer wait for it."""
    n, D = emb.shape
    if n > cap:
        raise ValueError("all_gather_faces_padded: %d faces on this rank exceed the per-rank capacity %d" % (n, cap))
    if stream is not None:
        stream.wait_stream(torch.cuda.current_stream(emb.device))
        for t in (emb, label, prob):
            t.record_stream(stream)                     # allocated on the current stream, read on the side stream
        with torch.cuda.stream(stream):
            return all_gather_faces_padded(emb, label, prob, cap, group=group)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    mine = torch.zeros(cap + 1, D + 2, dtype=torch.float32, device=emb.device)
    mine[:n, :D] = emb
    mine[:n, D] = label
    mine[:n, D + 1] = prob
    mine[cap, 0] = float(n)
    if world == 1:
        return mine.unsqueeze(0), D
    out = torch.empty(world * (cap + 1), D + 2, dtype=torch.float32, device=emb.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.view(world, cap + 1, D + 2), D


def all_gather_payload(payload, group=None, stream=None, out=None):
    """The step's exchange on the send buffer the fused tail kernel filled (FacePipeline: ``out["payload"]``, (cap + 1, D + 2)
    fp32 -- rows [0, n) = [embedding | label | prob], n in [cap, 0]): ONE ``all_gather_into_tensor`` straight from that
    buffer, no packing ops, no host synchronisation.  Returns ((world, cap + 1, D + 2) fp32, D, event) -- the same layout
    ``all_gather_faces_padded`` produces, so ``compact_faces`` applies.
    ``stream`` (CUDA): run the exchange on that side stream, ordered after what is enqueued on the current stream so far;
    the current stream does not wait for it (the collective waits for the slowest rank, the next batch's kernels need
    not).  ``event`` fires when the exchange has read ``payload``: the producer of the batch that reuses this send buffer
    (FacePipeline alternates two) must wait for it.  ``out``: preallocated result buffer."""
    D = payload.shape[1] - 2
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return payload.unsqueeze(0), D, None
    if out is None:
        out = torch.empty(world * payload.shape[0], payload.shape[1], dtype=payload.dtype, device=payload.device)
    if stream is not None and payload.is_cuda:
        stream.wait_stream(torch.cuda.current_stream(payload.device))
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(out.view(-1, payload.shape[1]), payload, group=group)
            ev = torch.cuda.Event()
            ev.record(stream)
        return out.view(world, payload.shape[0], payload.shape[1]), D, ev
    dist.all_gather_into_tensor(out.view(-1, payload.shape[1]), payload, group=group)
    return out.view(world, payload.shape[0], payload.shape[1]), D, None


class PeerGather:
    """The step's exchange WITHOUT a collective kernel: the send buffers live in symmetric memory (every rank maps every
    peer's buffer over NVLink: torch.distributed._symmetric_memory), and after a signal-pad barrier each rank PULLS the
    peers' payloads with plain device-to-device copies, which the copy engines execute -- no SM is taken from the next
    step's kernels.  An ALTERNATIVE to all_gather_payload, not the default: measured on 2 B200s (64 x 1080p frames per rank)
    a step takes 8.92 ms without any exchange, 9.11 ms with the NCCL all_gather on a side stream and 9.30 ms with this class
    (two signal-pad barriers + world copies per step), so the ~2 % the exchange costs is not the SM footprint of NCCL's
    kernel (NCCL_MAX_NCHANNELS=1 changes nothing either).  Bit-identical result (tests/test_gpu_multi.py).

        pg = PeerGather(group); pg.attach(face_pipeline)            # payload buffers are now peer-readable
        gathered, D, event = pg.gather(out["payload"], stream=side, out=buf)     # same contract as all_gather_payload

    ``event`` fires when every peer has read ``payload`` (second barrier): the producer that reuses the buffer waits for it."""

    def __init__(self, group=None):
        import torch.distributed._symmetric_memory as symm
        self.symm = symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self._handles = {}

    def alloc(self, shape, device):
        t = self.symm.empty(*shape, dtype=torch.float32, device=device)
        t.zero_()
        self._handles[t.data_ptr()] = self.symm.rendezvous(t, self.group)      # collective: every rank allocates in the same order
        return t

    def attach(self, face_pipeline):
        face_pipeline.payload_alloc = self.alloc
        return self

    def gather(self, payload, stream=None, out=None):
        h = self._handles.get(payload.data_ptr())
        if h is None:
            raise RuntimeError("payload was not allocated by this PeerGather (call attach() before the first batch)")
        D = payload.shape[1] - 2
        if out is None:
            out = torch.empty(self.world * payload.shape[0], payload.shape[1], dtype=payload.dtype, device=payload.device)
        out3 = out.view(self.world, payload.shape[0], payload.shape[1])
        cur = torch.cuda.current_stream(payload.device)
        st = stream if stream is not None else cur
        if stream is not None:
            stream.wait_stream(cur)
        with torch.cuda.stream(st):
            h.barrier(channel=0)                                   # every rank's payload is complete
            for step in range(self.world):
                r = (self.rank - step) % self.world
                src = payload if r == self.rank else h.get_buffer(r, payload.shape, payload.dtype)
                out3[r].copy_(src, non_blocking=True)              # device-to-device: copy engine, peer memory over NVLink
            h.barrier(channel=1)                                   # every rank has read every payload
            ev = torch.cuda.Event()
            ev.record(st)
        return out3, D, ev


def compact_faces(payload, D):
    """(payload, D) of all_gather_faces_padded -> (emb, label, prob, counts) exactly as all_gather_faces returns them
    (rank-order concatenation).  Reads the per-rank counts back: this is the one synchronising step."""
    cap = payload.shape[1] - 1
    counts = payload[:, cap, 0].long()
    parts = [payload[r, :c] for r, c in enumerate(counts.tolist())]
    allp = torch.cat(parts, 0)
    return allp[:, :D].contiguous(), allp[:, D].long(), allp[:, D + 1].contiguous(), counts
