"""Host ingest and reporting around the device pipeline: the reference's demo_video.py main loop without its per-frame host
work.

  * ``FrameBatcher``   demo_video.py:78-101: cv2.VideoCapture.read -> queue of ``n_frames`` BGR frames + (time_in_video, count)
                       per frame, flushed at the end of the video.  Frames are written straight into PINNED batch buffers
                       (two alternate), so the H2D copy of FacePipeline.submit overlaps compute; the BGR -> RGB conversion
                       (cv2.cvtColor, :107-110) is NOT done on the host: FacePipeline(bgr=True) swaps the channels on the
                       device (csrc/encoder_ops.cu vnfr_swap_rb_u8).
  * ``tracker_rows``   demo_video.py:155-181: the tracker CSV rows ``Time,Names,Frame_idx,Bboxes`` with boxes scaled by
                       [w, h, w, h] -- the same Python formatting calls as the reference, so the text is byte-identical
                       (tests/golden/tracker_rows.npz is produced by executing the reference's own lines).
  * ``run_video``      the loop: batches -> FacePipeline.submit (two batches in flight) -> names -> tracker file.
Drawing / frame dumps (draw_boxes_on_image, cv2.imwrite) and the emotion model are out of scope.
"""
import numpy as np
import torch

TRACKER_COLUMNS = ["Time", "Names", "Frame_idx", "Bboxes"]          # demo_video.py:67


def append_log_to_file(file_path, list_items):
    """utils/utils.py:60-64."""
    with open(file_path, "a") as f:
        f.write(",".join(list_items) + "\n")


def tracker_rows(frames_info, bth_names, bth_chosen_boxes, frame_shape):
    """demo_video.py:155-181 for one batch: frames_info = [[time_in_video, count], ...], bth_names = per-frame name lists,
    bth_chosen_boxes = per-frame lists of (4,) boxes (pixels), frame_shape = (h, w, 3).  Returns the text appended to the
    tracker file."""
    rows = []
    for idx, names in enumerate(bth_names):
        bboxes = bth_chosen_boxes[idx]
        row = [str(frames_info[idx][0]), '"' + str(names) + '"', str(frames_info[idx][1])]
        if len(bboxes) == 0:
            scaled_bboxes = []
        else:
            h, w, _ = frame_shape
            scale = np.array([w, h, w, h])
            scaled_bboxes = [list(x / scale) for x in bboxes]
        row.append('"' + str(scaled_bboxes) + '"')
        rows.append(",".join(row) + "\n")
    return "".join(rows)


class FrameBatcher:
    """Iterates over (frames (n, H, W, 3) uint8 pinned BGR, frames_info) batches of a cv2.VideoCapture-like object
    (``read() -> (ret, frame)``, ``get(CAP_PROP_FPS)``), demo_video.py:78-101: ``n_frames`` per batch, a short last batch at
    the end of the video, time_in_video = count / fps with count starting at 1."""

    def __init__(self, cap, n_frames, fps=None, pinned=True, n_buffers=3):
        self.cap, self.n_frames = cap, int(n_frames)
        self.fps = fps if fps is not None else cap.get(5)            # cv2.CAP_PROP_FPS == 5
        self.pinned, self.n_buffers = pinned, n_buffers
        self._bufs, self._slot = None, 0

    def _buffer(self, shape):
        if self._bufs is None or tuple(self._bufs[0].shape[1:]) != tuple(shape):
            mk = lambda: torch.empty((self.n_frames,) + tuple(shape), dtype=torch.uint8)
            self._bufs = [mk().pin_memory() if (self.pinned and torch.cuda.is_available()) else mk() for _ in range(self.n_buffers)]
        self._slot = (self._slot + 1) % self.n_buffers
        return self._bufs[self._slot]

    def __iter__(self):
        count, n, buf, info = 0, 0, None, []
        while True:
            ret, frame = self.cap.read()
            count += 1
            if ret:
                if buf is None:
                    buf = self._buffer(frame.shape)
                buf[n].copy_(torch.from_numpy(frame))
                info.append([count / self.fps, count])
                n += 1
            if n == self.n_frames or (not ret and n > 0):
                yield buf[:n], info
                buf, n, info = None, 0, []
            if not ret:
                return


def run_video(cap, pipeline, label2name, n_frames=16, output_tracker=None, fps=None):
    """The demo_video.py loop on the fused device pipeline.  ``pipeline``: FacePipeline(..., bgr=True) with a classifier;
    ``label2name``: dict label -> name (labels without a name, and the "unknown" label num_classes, give 'Unknown' like
    identify_person, demo_image.py:139-145).  Returns (tracker text, frames processed, faces found)."""
    text, n_done, n_faces = [",".join(TRACKER_COLUMNS) + "\n"], 0, 0
    if output_tracker is not None:
        with open(output_tracker, "w") as f:                         # demo_video.py:71-75
            f.write("")
        append_log_to_file(output_tracker, TRACKER_COLUMNS)
    pending = []

    def collect():
        nonlocal n_done, n_faces
        handle, info, shape = pending.pop(0)
        res = handle.result()
        names = [[label2name.get(int(l), "Unknown") for l in r["labels"]] for r in res]
        boxes = [[b for b in r["boxes"]] for r in res]
        rows = tracker_rows(info, names, boxes, shape)
        text.append(rows)
        if output_tracker is not None:
            with open(output_tracker, "a") as f:
                f.write(rows)
        n_done += len(res)
        n_faces += sum(len(r["labels"]) for r in res)

    for frames, info in FrameBatcher(cap, n_frames, fps=fps):
        pending.append((pipeline.submit(frames), info, tuple(frames.shape[1:])))
        if len(pending) == 2:                                        # two batches in flight
            collect()
    while pending:
        collect()
    return "".join(text), n_done, n_faces
