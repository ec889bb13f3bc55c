"""Host-side mirror of the reference's pipeline glue for the hot path, plus the fused device pipeline.

Reference-facing functions (same names / arguments / returns as the reference):
  * ``parallel_detect_and_align``  demo_image.py:273-306
  * ``recognize_celeb``            demo_image.py:50-76
  * ``find_embedding``             demo_image.py:30-34
  * ``identify_person``            demo_image.py:113-147
  * ``transforms_default``         data_loader/__init__.py:27-34, 52-56
  * ``cal_embedding`` and helpers  find_embedding.py:11-59
  * ``center_point_dict``          align_face.py:12-48
``FacePipeline`` is the same path without the host round trips between stages: frames (u8, device) -> MTCNN cascade ->
aligned crops -> InceptionResnetV1 -> MLP -> labels, one read-back at the end.
"""
import os
from pathlib import Path

import numpy as np
import torch

from . import _lib, encoder_plan

center_point_dict = {
    "(96, 112)": np.array([[30.2946, 51.6963], [65.5318, 51.5014], [48.0252, 71.7366], [33.5493, 92.3655],
                           [62.7299, 92.2041]], dtype=np.float32),
    "(112, 112)": np.array([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366], [41.5493, 92.3655],
                            [70.7299, 92.2041]], dtype=np.float32),
    "(150, 150)": np.array([[51.287415, 69.23612], [98.48009, 68.97509], [75.03375, 96.075806], [55.646385, 123.7038],
                            [94.72754, 123.48763]], dtype=np.float32),
    "(160, 160)": np.array([[54.706573, 73.85186], [105.045425, 73.573425], [80.036, 102.48086], [59.356144, 131.95071],
                            [101.04271, 131.72014]], dtype=np.float32),
    "(224, 224)": np.array([[76.589195, 103.3926], [147.0636, 103.0028], [112.0504, 143.4732], [83.098595, 184.731],
                            [141.4598, 184.4082]], dtype=np.float32),
}


def transforms_default(face):
    """np.float32 -> (x - 127.5) / 128 -> HWC -> CHW tensor (data_loader/__init__.py:27-34, 52-56)."""
    arr = (np.float32(face) - 127.5) / 128
    return torch.from_numpy(np.ascontiguousarray(np.transpose(arr, (2, 0, 1))))


# ----------------------------------------------------------------------------------------------------------------
# fused device pipeline
# ----------------------------------------------------------------------------------------------------------------
class FacePipeline:
    """detect -> align/crop -> embed -> classify on the device.

    align="similarity": demo_video semantics (get_face_from_boxes + 5-point similarity + cv2.warpAffine);
    align="extract":    MTCNN.extract semantics (box crop + area resize + fixed_image_standardization)."""

    def __init__(self, detector, encoder, classifier=None, target_fs=(160, 160), align="similarity", center_point=None,
                 threshold=0.0, max_faces_per_frame=32, return_faces_u8=False, bgr=False, input_format="rgb"):
        assert target_fs[0] == target_fs[1], "square targets only"
        self.det, self.enc, self.cls = detector, encoder, classifier
        self.S = int(target_fs[0])
        self.mode = 1 if align == "similarity" else 0
        self.template = center_point if center_point is not None else center_point_dict.get(str(tuple(target_fs)))
        if self.mode == 1 and self.template is None:
            raise ValueError("no landmark template for target size %s" % (target_fs,))
        self.threshold = threshold          # float, or the reference's per-class dict {str(label): threshold} (demo_image.py:118-124)
        self.max_faces_per_frame = max_faces_per_frame
        self.return_faces_u8 = return_faces_u8
        #: frames arrive in OpenCV's BGR order (cv2.VideoCapture.read): the channel swap of demo_video.py:107-110 runs on the
        #: device (vnfr_swap_rb_u8) instead of cv2.cvtColor on the host
        self.bgr = bgr
        #: "rgb": frames (B,H,W,3) u8;  "nv12": frames (B, H*3/2, W) u8 as a video decoder delivers them (luma plane + interleaved
        #: half-resolution U,V) -- converted on the device with cv2's BT.601 fixed point (vnfr_nv12_to_rgb_u8): half the
        #: host -> device bytes of RGB frames
        assert input_format in ("rgb", "nv12") and not (bgr and input_format == "nv12")
        self.input_format = input_format
        self._payloads, self._payload_slot, self._thr_class = {}, 0, None
        self._rgb_bufs = {}

    def _nv12_to_rgb(self, raw, out):
        """raw (n, H*3/2, W) u8 device NV12 -> out (n, H, W, 3) u8 RGB on the current stream."""
        n, h32, W = raw.shape
        H = h32 * 2 // 3
        with torch.cuda.device(raw.device):
            _lib.call("vnfr_nv12_to_rgb_u8", _lib.ptr(raw), _lib.ptr(out), n, H, W, _lib.stream_ptr())
        return out

    def _to_rgb(self, frames_dev, slot_key, in_place):
        """BGR device frames -> RGB (in place when the buffer is ours, else into a cached buffer per ``slot_key``)."""
        if not self.bgr:
            return frames_dev
        assert frames_dev.is_contiguous()
        out = frames_dev
        if not in_place:
            key = (slot_key, tuple(frames_dev.shape), frames_dev.device)
            out = self._rgb_bufs.get(key)
            if out is None:
                if len(self._rgb_bufs) > 4:
                    self._rgb_bufs.clear()
                out = self._rgb_bufs[key] = torch.empty_like(frames_dev)
        n_px = frames_dev.numel() // 3
        if n_px % 4:
            raise _lib.VnfrError("bgr=True needs a pixel count that is a multiple of 4")
        with torch.cuda.device(frames_dev.device):
            _lib.call("vnfr_swap_rb_u8", _lib.ptr(frames_dev), _lib.ptr(out), n_px, _lib.stream_ptr())
        return out

    def _payload(self, B, dev):
        """The all-gather send buffer of the next batch: (cap + 1, 514) fp32 -- row f = [embedding (512) | label | prob] of
        face f, the face count in [cap, 0] (dist.all_gather_payload exchanges it as is).  Two buffers alternate so that the
        read-back / exchange of batch i can still be in flight while the tail kernel of batch i+1 writes."""
        cap = max(B * self.max_faces_per_frame, 1)
        key = (cap, dev)
        if key not in self._payloads:
            # ``payload_alloc`` (optional attribute, set by dist.PeerGather.attach): allocate the send buffers in memory that
            # peer GPUs can read (the copy-engine exchange); default: ordinary device memory
            alloc = getattr(self, "payload_alloc", None) or (lambda shape, d: torch.zeros(*shape, dtype=torch.float32, device=d))
            self._payloads = {key: [alloc((cap + 1, 514), dev) for _ in range(2)]}
        self._payload_slot = 1 - self._payload_slot
        return self._payloads[key][self._payload_slot]

    def _thresholds(self, dev):
        """(scalar threshold, per-class device tensor or None) for the fused tail kernel."""
        if isinstance(self.threshold, dict):
            if self._thr_class is None or self._thr_class.device != dev:
                nc = self.cls.num_classes
                self._thr_class = torch.tensor([float(self.threshold[str(i)]) for i in range(nc)], dtype=torch.float32, device=dev)
            return 0.0, self._thr_class
        return float(self.threshold or 0.0), None

    def run_device(self, frames_u8, mark=None, pipelined=False):
        """frames_u8: CUDA uint8 (B,H,W,3).  Returns a dict of DEVICE tensors + the face count (one tiny sync to size
        the encoder batch): count (B,), boxes (B,capf,5), points (B,capf,10), faces_u8 (F,S,S,3), emb (F,512),
        label (F,), prob (F,), face_img (F,).

        ``pipelined=True`` (a stream of batches): the detection cascade of this call does not wait for the work the
        previous call left on the current stream (its encoder / classifier), only for that call's face crops -- cascade
        i+1 and encoder i then share the GPU.  The caller asserts that ``frames_u8`` was complete before the previous
        call returned (e.g. frames resident on the device); results alternate between two workspace slots, so the
        ``count`` / ``boxes`` / ``points`` tensors of a call stay valid until the call after the next one."""
        from .models.mtcnn import CropWorkspaceOverflow
        marked = mark is not None                      # stage markers need the single-stream cascade
        mark = mark or (lambda name: None)
        with torch.no_grad():
            B = frames_u8.shape[0]
            if self.input_format == "nv12":
                self._rgb_slot = 1 - getattr(self, "_rgb_slot", 0)
                H, W = frames_u8.shape[1] * 2 // 3, frames_u8.shape[2]
                key = ("nv12", self._rgb_slot, B, H, W, frames_u8.device)
                if key not in self._rgb_bufs:
                    if len(self._rgb_bufs) > 4:
                        self._rgb_bufs.clear()
                    self._rgb_bufs[key] = torch.empty(B, H, W, 3, dtype=torch.uint8, device=frames_u8.device)
                frames_u8 = self._nv12_to_rgb(frames_u8.contiguous(), self._rgb_bufs[key])
            if self.bgr:
                self._rgb_slot = 1 - getattr(self, "_rgb_slot", 0)
                frames_u8 = self._to_rgb(frames_u8.contiguous(), self._rgb_slot, in_place=False)
            for attempt in range(6):
                if marked or B < 32:
                    ws = self.det.detect_device(frames_u8, mark=mark)
                else:
                    # two halves on two streams: the low-occupancy stage kernels of one half overlap the other half's
                    # P-Net / R-Net / O-Net (measured 8.14 -> 7.74 ms for 64 x 1080p)
                    n = self.device_chunks
                    bounds = [(i * B // n, (i + 1) * B // n) for i in range(n)]
                    prev = getattr(self, "_dev_crops_done", None)
                    if pipelined and attempt == 0 and prev is not None:
                        self._dev_slot = 1 - getattr(self, "_dev_slot", 0)
                        ws = self.det.detect_device_chunked(frames_u8, [prev] * n, bounds, slot=self._dev_slot, wait_current=False)
                    else:
                        ws = self.det.detect_device_chunked(frames_u8, None, bounds, slot=getattr(self, "_dev_slot", 0))
                try:
                    return self._embed_classify(ws, mark, crops_event="_dev_crops_done" if (pipelined and not marked) else None)
                except CropWorkspaceOverflow:          # more candidates than the crop workspaces hold: grow and repeat
                    if attempt == 5:
                        raise
                    self.det.grow_crop_workspace()
                    self._dev_crops_done = None

    def _embed_classify(self, ws, mark, crops_event=None):
        """detections (DetectWorkspace / ResultWorkspace) -> aligned crops -> encoder -> classifier, on the device.
        ``crops_event``: attribute name under which an event recorded right after the face-crop kernel is stored (the
        last reader of the frames and of the detection workspace: what a pipelined next batch has to wait for)."""
        with torch.no_grad():
            dt = self.enc.half_dtype or encoder_plan.HALF
            u8, half, fimg, cap = self.det.face_crops_device(ws, self.mode, self.S, self.det.margin, self.template, dt,
                                                            ws.B * self.max_faces_per_frame, want_u8=self.return_faces_u8)
            mark("face_crops")
            if crops_event is not None:
                ev = torch.cuda.Event()
                ev.record()
                setattr(self, crops_event, ev)
            host = ws.counters[-(ws.B + 1):].cpu().numpy()       # out_count (B) + status: the only mid-pipeline read-back
            self.det.check_status(int(host[-1]))
            F = int(host[:-1].sum())
            out = {"count": ws.out_count, "boxes": ws.out_box, "points": ws.out_pts, "n_faces": F, "faces_u8": None if u8 is None else u8[:F],
                   "face_img": fimg[:F], "count_host": host[:-1].copy(), "ws": ws}
            dev = ws.out_box.device
            payload = self._payload(ws.B, dev)
            out["payload"] = payload
            if F > payload.shape[0] - 1:
                raise _lib.VnfrError("more faces than max_faces_per_frame allows")
            if F == 0:
                payload[-1, 0] = 0.0
                out.update(emb=payload[:0, :512], label=torch.zeros(0, dtype=torch.int64, device=dev),
                           prob=torch.zeros(0, device=dev))
                return out
            # convolutions (one graph replay) + ONE fused tail kernel: pool -> bottleneck -> L2 norm -> MLP -> log-softmax ->
            # argmax -> identify_person's threshold (demo_image.py:131-137), rows written straight into the send buffer
            thr, thr_class = self._thresholds(dev) if self.cls is not None else (0.0, None)
            res = self.enc.embed_s2d(half[:F], self.S, classifier=self.cls, threshold=thr, thr_class=thr_class, payload=payload,
                                     mark=mark)
            mark("tail")
            out["emb"] = res["emb"]
            if self.cls is not None:
                out["label"], out["prob"] = res["label"], res["prob"]
        return out

    def embed_faces(self, faces_u8, payload=None, mark=None):
        """Aligned faces (n, S, S, 3) uint8 -- what parallel_detect_and_align returns, on the host (pinned for an asynchronous
        copy) or on the device -- -> transforms_default + InceptionResnetV1 + MLP + identify_person's threshold on the device
        (recognize_celeb, demo_image.py:50-76, without its host transforms).  Returns the embed_s2d dict (emb, label, prob)
        of device tensors; with ``payload`` the rows go straight into that send buffer."""
        t = torch.as_tensor(faces_u8)
        dev = t.device if t.is_cuda else (self.det._cuda_device() if self.det is not None else torch.device(self.enc.device))
        if not t.is_cuda:
            t = t.to(dev, non_blocking=True)
        n, S = t.shape[0], t.shape[1]
        dt = self.enc.half_dtype or encoder_plan.HALF
        key = (n, S, dt, t.device)
        if getattr(self, "_s2d_key", None) != key:
            self._s2d = torch.zeros(n, (S + 1) // 2, (S + 1) // 2, 16, dtype=dt, device=t.device)
            self._s2d_key = key
        with torch.no_grad(), torch.cuda.device(t.device):
            _lib.call("vnfr_u8hwc_to_s2d16", _lib.ptr(t.contiguous()), n, S, S, _lib.ptr(self._s2d), encoder_plan.dtype_code(dt), _lib.stream_ptr())
            thr, thr_class = self._thresholds(t.device) if self.cls is not None else (0.0, None)
            return self.enc.embed_s2d(self._s2d, S, classifier=self.cls, threshold=thr, thr_class=thr_class, payload=payload, mark=mark)

    #: device-resident frames: the cascade runs as this many sub-batches alternating between two streams
    device_chunks = int(os.environ.get("VNFR_DEVICE_CHUNKS", "2"))
    #: frames per sub-batch of the host-frame path (H2D of sub-batch i+1 overlaps the cascade of sub-batch i).  Measured for
    #: 64 x 1080p with two batches in flight (tools/e2e_pipe_probe.py): 8/4 10.94 ms, 16/8 10.45, 16/16 10.43, 32/32 11.15
    sub_batch = 16
    #: the first sub-batch is smaller: with a single batch in flight nothing can overlap its copy, so it should land quickly
    first_sub_batch = 8

    #: NV12 frames are half the bytes: the copy is no longer what the cascade waits for, and two large sub-batches keep the
    #: cascade kernels as efficient as on device-resident frames (measured 9.50 -> 9.16 ms per 64 x 1080p, tools/e2e_timeline.py)
    sub_batch_nv12 = 32

    def _sub_batches(self, B):
        bounds, b0 = [], 0
        sub = self.sub_batch_nv12 if self.input_format == "nv12" else self.sub_batch
        first = sub if self.input_format == "nv12" else min(self.first_sub_batch, self.sub_batch)
        while b0 < B:
            n = first if b0 == 0 else sub
            bounds.append((b0, min(B, b0 + n)))
            b0 += n
        return bounds

    def _run_host_frames(self, t, dev):
        """Pinned host frames -> device in sub-batches on a copy stream, the detection cascade of each sub-batch
        starting as soon as its frames have landed; crops / encoder / classifier then run once over all faces.

        Two frame buffers / workspace slots alternate between calls and nothing here waits for the work an earlier call
        left on the current stream: the H2D copy and the cascade of batch i+1 overlap the encoder of batch i when the
        caller keeps two batches in flight (``submit``)."""
        nv12 = self.input_format == "nv12"
        if nv12:
            B, H, W = t.shape[0], t.shape[1] * 2 // 3, t.shape[2]
        else:
            B, H, W, _ = t.shape
        key = (B, H, W, dev, nv12)
        if getattr(self, "_fbuf_key", None) != key:
            self._fbufs = [torch.empty(B, H, W, 3, dtype=torch.uint8, device=dev) for _ in range(2)]
            self._rawbufs = [torch.empty_like(t, device=dev) for _ in range(2)] if nv12 else None
            self._fbuf_key = key
            self._copy_stream = torch.cuda.Stream(dev)
            self._host_crops_done = [None, None]
            self._host_slot = 0
        slot = self._host_slot
        self._host_slot = 1 - slot
        buf, cs = self._fbufs[slot], self._copy_stream
        cur = torch.cuda.current_stream(dev)
        if self._host_crops_done[slot] is not None:
            cs.wait_event(self._host_crops_done[slot])       # the batch that used this slot two calls ago has read its frames
        else:
            cs.wait_stream(cur)
        events = []
        bounds = self._sub_batches(B)
        with torch.cuda.stream(cs):
            for b0, b1 in bounds:
                if nv12:                                               # half the bytes over PCIe, colour conversion behind the copy
                    raw = self._rawbufs[slot]
                    raw[b0:b1].copy_(t[b0:b1], non_blocking=True)
                    self._nv12_to_rgb(raw[b0:b1], buf[b0:b1])
                else:
                    buf[b0:b1].copy_(t[b0:b1], non_blocking=True)
                if self.bgr:
                    self._to_rgb(buf[b0:b1], None, in_place=True)      # on the copy stream, right behind the sub-batch's H2D
                ev = torch.cuda.Event()
                ev.record(cs)
                events.append(ev)
        from .models.mtcnn import CropWorkspaceOverflow
        for attempt in range(6):
            with torch.no_grad():
                ws = self.det.detect_device_chunked(buf, events if attempt == 0 else None, bounds, slot=slot,
                                                    wait_current=attempt > 0)
            try:
                out = self._embed_classify(ws, lambda name: None, crops_event="_host_crops_tmp")
                self._host_crops_done[slot] = self._host_crops_tmp
                return out
            except CropWorkspaceOverflow:
                if attempt == 5:
                    raise
                self.det.grow_crop_workspace()

    def submit(self, frames):
        """Asynchronous form of ``__call__`` for a stream of batches: enqueues the whole path for ``frames`` ((B,H,W,3)
        uint8, pinned host memory for overlap) plus the device->host copy of the results and returns a PendingResult;
        ``.result()`` waits for that batch only.  With two batches in flight (submit batch i+1, then collect batch i) the
        H2D copy and the detection cascade of batch i+1 run under the encoder of batch i.  (The one host wait inside
        ``submit`` is the face count of the batch, which sizes its encoder launch.)"""
        dev = self.det._cuda_device()
        t = torch.as_tensor(frames)
        if t.is_cuda:
            out = self.run_device(t)
        elif t.is_pinned() and t.dim() == (3 if self.input_format == "nv12" else 4) and t.shape[0] > self.sub_batch and t.dtype == torch.uint8:
            out = self._run_host_frames(t, dev)
        else:
            out = self.run_device(t.to(dev, non_blocking=True))
        return PendingResult(self, out)

    def __call__(self, frames):
        """frames: (B,H,W,3) uint8 numpy / torch (host or device).  Returns per-frame lists (boxes (n,4) numpy, labels,
        probs) plus the (F,512) embeddings -- the H2D of the frames (overlapped with compute when the host tensor is
        pinned), one D2H of the results."""
        return self.submit(frames).result()


class PendingResult:
    """Results of one FacePipeline.submit(): the device->host copies are enqueued at construction (pinned staging buffers,
    two sets alternating per pipeline), ``result()`` waits for them and builds the per-frame lists."""

    def __init__(self, fp, out):
        self.cnt = out["count_host"]
        self.F = F = out["n_faces"]
        B = len(self.cnt)
        nmax = max(int(self.cnt.max()) if B else 0, 1)
        dev = out["boxes"].device
        key = (B, out["boxes"].shape[1], out["emb"].shape[1], fp.max_faces_per_frame)
        if getattr(fp, "_stage_key", None) != key:
            for o in fp.__dict__.get("_stage_owner", []):          # geometry changed: collect what still uses the old buffers
                if o is not None and o._res is None:
                    o.result()
            fp._stage_owner = [None, None]
            cap = max(B * fp.max_faces_per_frame, 1)
            pin = lambda *shape, dtype: torch.empty(*shape, dtype=dtype).pin_memory()
            fp._stage = [dict(boxes=pin(B, out["boxes"].shape[1], 5, dtype=torch.float32),
                              rows=pin(cap, out["payload"].shape[1], dtype=torch.float32)) for _ in range(2)]
            fp._stage_key, fp._stage_slot = key, 0
        slot = fp._stage_slot
        st = fp._stage[slot]
        fp._stage_slot = 1 - slot
        # the staging buffers alternate: a third submit() before the first result() has been collected would overwrite them,
        # so the previous owner of this slot is materialised first (that wait only costs when more than two are in flight)
        owners = fp.__dict__.setdefault("_stage_owner", [None, None])
        if owners[slot] is not None and owners[slot]._res is None:
            owners[slot].result()
        owners[slot] = self
        self._res = None
        # contiguous -> contiguous pinned copies only: a strided device->host copy goes through a pageable temporary and
        # blocks the host until everything enqueued before it (the encoder of this batch) has finished
        st["boxes"].copy_(out["boxes"], non_blocking=True)
        self.boxes = st["boxes"][:, :nmax]
        if F > st["rows"].shape[0]:
            raise _lib.VnfrError("more faces than max_faces_per_frame allows")
        self.rows = st["rows"][:F]
        self.has_cls = "label" in out
        if F:
            self.rows.copy_(out["payload"][:F], non_blocking=True)        # [embedding | label | prob] rows: one contiguous copy
        self.done = torch.cuda.Event()
        self.done.record(torch.cuda.current_stream(dev))
        self.out = out                                   # keeps the device tensors alive until the copies have run

    def result(self):
        if self._res is not None:
            return self._res
        self.done.synchronize()
        cnt, F = self.cnt, self.F
        boxes = self.boxes.numpy()
        rows = self.rows.numpy()
        D = rows.shape[1] - 2
        lab = rows[:, D].astype(np.int64) if self.has_cls else np.zeros(F, np.int64)
        prob = rows[:, D + 1].copy() if self.has_cls else np.zeros(F, np.float32)
        emb = rows[:, :D].copy()
        res, o = [], 0
        for b in range(len(cnt)):
            n = int(cnt[b])
            res.append({"boxes": boxes[b, :n, :4].copy(), "det_prob": boxes[b, :n, 4].copy(), "labels": lab[o:o + n],
                        "probs": prob[o:o + n], "emb": emb[o:o + n]})
            o += n
        self.out = None
        self._res = res
        return res


# ----------------------------------------------------------------------------------------------------------------
# reference-facing glue
# ----------------------------------------------------------------------------------------------------------------
def parallel_detect_and_align(rgb_images, detection_md, center_point, target_fs, log=False):
    """demo_image.py:273-306: returns (list[list[u8 (S,S,3) RGB]], list[list[box (4,)]]) on the host."""
    dev = detection_md._cuda_device()
    frames = torch.as_tensor(np.stack([np.asarray(im) for im in rgb_images])).to(dev)
    with torch.no_grad():
        from .models.mtcnn import CropWorkspaceOverflow
        for attempt in range(6):
            ws = detection_md.detect_device(frames)
            u8, _, _, _ = detection_md.face_crops_device(ws, 1, int(target_fs[0]), 0, center_point)
            cnt = ws.out_count.cpu().numpy()
            try:
                detection_md.check_status(int(ws.status.item()))
                break
            except CropWorkspaceOverflow:
                if attempt == 5:
                    raise
                detection_md.grow_crop_workspace()
        F = int(cnt.sum())
        faces = u8[:F].cpu().numpy()
        nmax = int(cnt.max()) if len(cnt) else 0
        boxes = ws.out_box[:, :max(nmax, 1), :4].cpu().numpy()
    bth_faces, bth_boxes, o = [], [], 0
    for b in range(len(cnt)):
        n = int(cnt[b])
        bth_faces.append([faces[o + i] for i in range(n)])
        bth_boxes.append([boxes[b, i].copy() for i in range(n)])
        if n == 0 and log:
            print("Face not found in this image !")
        o += n
    return bth_faces, bth_boxes


def find_embedding(image_tensor, embedding_model):
    """demo_image.py:30-34."""
    embedding_model.eval()
    with torch.no_grad():
        embeddings = embedding_model(image_tensor)
    return embeddings.detach()


def identify_person(embeddings, classify_model, name_df, threshold):
    """demo_image.py:113-147: argmax, exp(log-prob), (per-class) threshold -> label or num_classes, name lookup."""
    classify_model.eval()
    with torch.no_grad():
        output = classify_model(embeddings)
    n_classes = output.size(1)
    predictions = torch.argmax(output, dim=1).detach().cpu().numpy()
    probs = torch.exp(output).detach().cpu().numpy()
    chosen_prob = probs[np.arange(len(predictions)), predictions]
    if type(threshold) is float:
        thr = np.full(len(predictions), threshold)
    else:
        thr = np.array([threshold[str(p)] for p in predictions])
    filtered = np.where(chosen_prob >= thr, predictions, n_classes)
    # vectorised replacement of the per-face pandas scan (demo_image.py:139-145): first name per label
    lookup = {}
    for lab, name in zip(name_df["label"].tolist(), name_df["name"].tolist()):
        lookup.setdefault(lab, name)
    return [lookup.get(int(p), "Unknown") for p in filtered]


def recognize_celeb(bth_alg_face_list, device, emb_model, classify_model, transforms, label2name_df, threshold):
    """demo_image.py:50-76."""
    flat = [f for x in bth_alg_face_list for f in x]
    if len(flat) == 0:
        return [[] for _ in bth_alg_face_list]
    tf = torch.stack([transforms(f) for f in flat], dim=0)
    embeddings = find_embedding(tf.to(device), emb_model)
    names = identify_person(embeddings, classify_model, label2name_df, threshold)
    out, c = [], 0
    for x in bth_alg_face_list:
        out.append(names[c:c + len(x)])
        c += len(x)
    return out


# find_embedding.py -------------------------------------------------------------------------------------------------
def create_batch_images(list_files, batch_size):
    """find_embedding.py:11-20 -- including the reference's trailing (possibly empty) batch."""
    n_batchs = len(list_files) // batch_size
    batches = [list_files[i * batch_size:(i + 1) * batch_size] for i in range(n_batchs)]
    batches.append(list_files[n_batchs * batch_size:])
    return batches, n_batchs


def create_image_tensors(data_dir_path, list_files, transforms, pool=None):
    """find_embedding.py:23-32; ``pool`` (a concurrent.futures executor) decodes + transforms the files in parallel
    (PIL / zlib release the GIL), results stay in file order."""
    from PIL import Image
    load = lambda f: transforms(Image.open(str(Path(data_dir_path) / f)))
    items = list(pool.map(load, list_files)) if pool is not None else [load(f) for f in list_files]
    return torch.stack(items, 0)


def save_embeddings(embeddings, list_files, output_dir, pool=None):
    """find_embedding.py:34-42: one ``<stem>.npz`` per image, key ``arr_0``, (512,) fp32 -- the wire format
    VNCelebEmbDataset reads (data_loader/vn_celeb_emb_dataset.py:14).  ``pool``: write the files from worker threads."""
    def save(i):
        np.savez_compressed(str(Path(output_dir) / "{}.npz".format(list_files[i].split(".")[0])), embeddings[i])
    if pool is not None:
        return [pool.submit(save, i) for i in range(embeddings.shape[0])]
    for i in range(embeddings.shape[0]):
        save(i)
    return []


def cal_embedding(data_dir, batch_size, model, transforms, output_dir, device, workers=0, rank=0, world=1):
    """find_embedding.py:45-59.  Unlike the reference an EMPTY trailing batch is skipped instead of crashing in
    torch.stack (the reference raises when len(files) % batch_size == 0).

    Extensions for the offline-embedding configs, same files / same ``.npz`` bytes-on-disk format:
    ``workers`` > 0 decodes batch i+1 and writes batch i-1 on a thread pool while the GPU embeds batch i; ``rank`` /
    ``world`` shard the sorted file list into contiguous blocks (dist.shard_range), each rank writing its own files."""
    from concurrent.futures import ThreadPoolExecutor
    from . import dist as vdist
    os.makedirs(output_dir, exist_ok=True)
    model.eval()
    list_files = sorted(os.listdir(data_dir))
    if world > 1:
        lo, hi = vdist.shard_range(len(list_files), rank, world)
        list_files = list_files[lo:hi]
    batches, n_batchs = create_batch_images(list_files, batch_size)
    batches = [b for b in batches if b]
    pool = ThreadPoolExecutor(max_workers=workers) if workers and workers > 0 else None
    try:
        nxt = pool.submit(create_image_tensors, Path(data_dir), batches[0], transforms, pool) if (pool and batches) else None
        writes = []
        for k, batch_file in enumerate(batches):
            if pool is not None:
                tensors = nxt.result()
                nxt = pool.submit(create_image_tensors, Path(data_dir), batches[k + 1], transforms, None) if k + 1 < len(batches) else None
            else:
                tensors = create_image_tensors(Path(data_dir), batch_file, transforms)
            with torch.no_grad():
                embeddings = model(tensors.to(device)).detach().cpu().numpy()
            writes += save_embeddings(embeddings, batch_file, output_dir, pool)
        for w in writes:
            w.result()
    finally:
        if pool is not None:
            pool.shutdown(wait=True)
