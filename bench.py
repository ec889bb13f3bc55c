#!/usr/bin/env python
"""Benchmark of the detect -> align -> embed -> classify hot path (BASELINE.json metric: faces/sec on 1080p frames).

    python bench.py --gpus N --steps K --warmup W            # this framework (sm_100a kernels through the C-ABI)
    python bench.py --impl reference ...                     # the reference's CPU path (oracle port) on the host cores

One step = one pass of the whole path over one batch of synthetic 1080p frames (BASELINE config 3: 64 frames per rank,
12 bundled faces pasted per frame, MTCNN(**cfg/detection/mtcnn.json) i.e. min_face_size 50 + keep_all, demo_video
alignment to 160x160, InceptionResnetV1 + MLPModel(512, 1001) with random-init weights).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "faces/sec detect+embed+classify (1080p)"
UNIT = "faces/s"

#: BASELINE.json configs measured by this file.  Per-frame algorithmic work of the detection kernels: SURVEY.md 8(d) / Appendix C.
WORKLOADS = {
    3: dict(workload="pipeline_1080p (BASELINE config 3)", kind="1080p", frames=64, min_face=50, max_faces=32, cpu_frames=8, ref_frames=64,
            frame="1920x1080x3 u8", pnet_gflop=0.818, pyramid_mb=9.11, metric=METRIC),
    4: dict(workload="pipeline_4k_crowded (BASELINE config 4)", kind="4k", frames=8, min_face=20, max_faces=96, cpu_frames=1, ref_frames=2,
            frame="3840x2160x3 u8", pnet_gflop=21.67, pyramid_mb=97.0,
            metric="faces/sec detect+embed+classify (crowded 3840x2160, min_face_size 20)"),
}
ENC_FLOP_PER_FACE = 2.8353e9          # 1 417.66 MMAC per 160x160 crop (conv + last_linear)
MLP_FLOP_PER_FACE = 6.2e6


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def make_frames(n, first_seed, kind="1080p"):
    """Synthetic frames that contain faces (random noise yields zero detections): seeded, frame seed = global index."""
    from vn_celeb_face_recognition_b200 import synthetic
    return synthetic.frames(kind, n, first_seed=first_seed)


def build_models(dev, seed=0, min_face_size=50):
    import torch
    from vn_celeb_face_recognition_b200.models import MTCNN, InceptionResnetV1, MLPModel
    from vn_celeb_face_recognition_b200 import synthetic
    torch.manual_seed(seed)
    det = MTCNN(image_size=160, keep_all=True, min_face_size=min_face_size, device=dev)       # cfg/detection/mtcnn.json
    # random init (BASELINE config) -- the SEEDED, BN-calibrated one: torch's default init + eval-mode BN is degenerate
    # (every input -> the same embedding), which would make label agreement and cosine parity trivially perfect
    enc = InceptionResnetV1(pretrained=None, device=dev).eval()
    enc.load_state_dict(synthetic.encoder_state_dict_seed0())
    cls = MLPModel(512, 1001).to(dev).eval()
    cls.load_state_dict(synthetic.mlp_state_dict(1001, seed=seed))
    return det, enc, cls


class ClockSampler:
    """SM clock / throttle-reason samples DURING the timed region (B200_PROFILING.md recipe): NVML polled every 10 ms from
    a thread (nvidia-smi -lms cannot deliver a first sample inside a ~100 ms region); nvidia-smi is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    #: NVML poll period; VNFR_CLOCK_POLL_MS overrides (probing whether the sampler itself perturbs the timed region)
    POLL_S = float(os.environ.get("VNFR_CLOCK_POLL_MS", "10")) * 1e-3

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc, self.nvml, self._stop = gpu_index, [], None, None, False
        self.sm, self.reasons, self.max_sm = [], set(), None

    def _phys_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.idx])
            except (ValueError, IndexError):
                pass
        return self.idx

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._phys_index())
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._phys_index()), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        names = [("hw_slowdown", getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                 ("hw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                 ("sw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                 ("sw_power_cap", getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4))]
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                r = int(get_reasons(self.h))
                for name, bit in names:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.POLL_S)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self._stop = True
            self.thread.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvidia-smi"}


# --------------------------------------------------------------------------------------------------------------------
_REF_MODELS = {}


def reference_kind():
    """"reference" when the unmodified reference tree is importable here (VNFR_REFERENCE_ROOT, /root/reference or
    baseline/_ref), else "port" (the oracle restatement, pinned to the reference by tests/golden)."""
    if os.environ.get("VNFR_BENCH_PORT"):
        return "port"
    try:
        from oracle import ref_shims
        return "reference" if ref_shims.reference_available() else "port"
    except Exception:
        return "port"


def cpu_reference_leg(frames, enc_sd, mlp_sd, repeats=1, threads=None, min_face_size=50, want_outputs=False):
    """The reference's CPU path on the host cores: parallel_detect_and_align + recognize_celeb (demo_image.py:273-306,
    :50-76) -- the UNMODIFIED reference when its tree is present, else the oracle port.  Returns (faces/s, n_faces, seconds,
    threads, kind[, outputs]) with outputs = (per-frame label lists, (F,512) embeddings, (F,C) log-probs) of the port."""
    import torch
    from oracle import pipeline as opipe, align
    from vn_celeb_face_recognition_b200 import synthetic as synth
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    kind = reference_kind()
    best, nf, outputs = None, 0, None
    if kind == "reference":
        import pandas as pd
        from oracle import ref_shims
        ref = ref_shims.load_reference()
        if "det" not in _REF_MODELS or _REF_MODELS.get("mfs") != min_face_size:
            _REF_MODELS["det"] = ref.models.MTCNN(image_size=160, keep_all=True, device="cpu", min_face_size=min_face_size)
            _REF_MODELS["mfs"] = min_face_size
            enc = ref.models.InceptionResnetV1(pretrained=None, device="cpu").eval()
            enc.load_state_dict(enc_sd)
            mlp = ref.models.MLPModel(512, mlp_sd["dense_2.weight"].shape[0]).eval()
            mlp.load_state_dict(mlp_sd)
            nc = mlp_sd["dense_2.weight"].shape[0]
            _REF_MODELS.update(enc=enc, mlp=mlp, names=pd.DataFrame({"label": np.arange(nc), "name": ["id%d" % i for i in range(nc)]}))
        cp = ref.align_face.center_point_dict["(160, 160)"]
        for _ in range(repeats):
            t0 = time.perf_counter()
            faces, _ = ref.demo_image.parallel_detect_and_align(list(frames), _REF_MODELS["det"], cp, (160, 160))
            ref.demo_image.recognize_celeb(faces, "cpu", _REF_MODELS["enc"], _REF_MODELS["mlp"], ref.data_loader.transforms_default,
                                           _REF_MODELS["names"], 0.0)
            dt = time.perf_counter() - t0
            nf = sum(len(x) for x in faces)
            best = dt if best is None else min(best, dt)
        if want_outputs:
            outputs = opipe.recognize(faces, enc_sd, mlp_sd, 0.0, return_logp=True)
    else:
        sds = synth.mtcnn_state_dicts()
        cp = align.CENTER_POINTS[(160, 160)]
        for _ in range(repeats):
            t0 = time.perf_counter()
            faces, _ = opipe.parallel_detect_and_align(list(frames), sds, cp, (160, 160), min_face_size=min_face_size)
            outputs = opipe.recognize(faces, enc_sd, mlp_sd, 0.0, return_logp=True)
            dt = time.perf_counter() - t0
            nf = sum(len(x) for x in faces)
            best = dt if best is None else min(best, dt)
    if want_outputs:
        return nf / best, nf, best, threads, kind, outputs
    return nf / best, nf, best, threads, kind


def label_parity(gpu_res, outputs):
    """`parity` block of the JSON line: the GPU step's first frames (per-frame dicts of FacePipeline results) against the
    CPU reference path on the same frames: face counts, label agreement, min embedding cosine, and every differing face
    with the reference's own top-1 minus (our label) log-prob margin."""
    ref_labels, ref_emb, ref_logp = outputs
    n_fr = len(ref_labels)
    faces, agree, min_cos, flips, counts_equal, o = 0, 0, 1.0, [], True, 0
    for i in range(n_fr):
        r, rl = gpu_res[i], ref_labels[i]
        if len(r["labels"]) != len(rl):
            counts_equal = False
            o += len(rl)
            continue
        for k, lab in enumerate(rl):
            cos = float((r["emb"][k] * ref_emb[o + k]).sum())
            min_cos = min(min_cos, cos)
            faces += 1
            got = int(r["labels"][k])
            if got == lab:
                agree += 1
            else:
                flips.append({"frame": i, "face": k, "got": got, "ref": int(lab), "cosine": round(cos, 6),
                              "ref_margin": float(ref_logp[o + k].max() - ref_logp[o + k][got])})
        o += len(rl)
    return {"frames": n_fr, "faces": faces, "face_counts_equal": counts_equal, "label_agree": agree, "min_cos": min_cos,
            "flips": flips, "note": "end to end vs the CPU reference path on the same frames; tests/test_gpu_pipeline.py::"
            "test_config3_labels_identical_to_reference separates arithmetic (identical crops: labels identical) from input "
            "perturbation (landmarks agree to ~1e-4 px; the random-init encoder amplifies the resulting crop differences)"}


def cpu_embed_leg(enc_sd, mlp_sd, batch=64, repeats=2, threads=None):
    """Second figure of the metric on the host cores: InceptionResnetV1 + MLP (oracle port) on a batch of 160x160 crops
    (batch 64).  Returns (embeds/s, seconds)."""
    import torch
    from oracle import nets
    from vn_celeb_face_recognition_b200 import synthetic as synth
    torch.set_num_threads(threads or os.cpu_count() or 1)
    x = synth.crops_160(batch, seed=1)
    best = None
    with torch.no_grad():
        for i in range(repeats + 1):                       # first pass = warm-up
            t0 = time.perf_counter()
            nets.mlp_forward(mlp_sd, nets.encoder_forward(enc_sd, x))
            dt = time.perf_counter() - t0
            if i > 0:
                best = dt if best is None else min(best, dt)
    return batch / best, best


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path -- the unmodified reference when its tree is
    present (cpu_baseline.kind "reference"), else the oracle port ("port") --, all host threads, on the same frames/step
    as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from vn_celeb_face_recognition_b200.models import InceptionResnetV1, MLPModel
    from vn_celeb_face_recognition_b200 import synthetic
    enc_sd = synthetic.encoder_state_dict_seed0()            # the same seeded weights as our arm (build_models)
    mlp_sd = synthetic.mlp_state_dict(1001, seed=0)
    wl = WORKLOADS[args.config]
    n = args.ref_frames or wl["ref_frames"]              # config 4: a bounded sample (a 4K / min_face 20 frame costs seconds on the CPU)
    frames = make_frames(n, 0, wl["kind"])
    times, faces = [], 0
    for i in range(args.warmup + args.steps):
        fps, nf, dt, threads, kind = cpu_reference_leg(frames, enc_sd, mlp_sd, min_face_size=wl["min_face"])
        if i >= args.warmup:
            times.append(dt)
            faces += nf
    total = sum(times)
    val = faces / total
    sample = "%d synthetic %s frames (seeds 0..%d) per step (our arm: %d per rank), %d timed steps" % (n, wl["kind"], n - 1, wl["frames"], args.steps)
    eps, edt = cpu_embed_leg(enc_sd, mlp_sd, threads=threads)
    line = {"metric": wl["metric"], "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["workload"], "frames_per_rank": n, "frames_per_step": n,
                       "faces_per_step": faces // max(1, args.steps), "min_face_size": wl["min_face"], "align": "similarity 160x160",
                       "num_classes": 1001},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                             "embeds_per_s": eps, "embeds_sample": "InceptionResnetV1 + MLP on 64 crops, best of 2, %.2f s" % edt},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from vn_celeb_face_recognition_b200 import _lib, pipeline, dist as vdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this framework has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    wl = WORKLOADS[args.config]
    B = args.frames or wl["frames"]
    det, enc, cls = build_models(dev, min_face_size=wl["min_face"])
    enc.chunk = args.chunk
    fp = pipeline.FacePipeline(det, enc, cls, (160, 160), "similarity", max_faces_per_frame=wl["max_faces"])
    frames_np = make_frames(B, rank * B, wl["kind"])
    fp_host = fp
    if args.ingest == "nv12":
        # host frames as a video decoder delivers them (NV12, 1.5 B/px): every leg -- device-resident, e2e, CPU baseline -- sees the
        # SAME pixels, cv2's decode of those NV12 frames; only the e2e path uploads NV12 and converts on the device
        import cv2
        Hh, Ww = frames_np.shape[1:3]
        nv12 = np.empty((B, Hh * 3 // 2, Ww), np.uint8)
        for i in range(B):
            i420 = cv2.cvtColor(frames_np[i], cv2.COLOR_RGB2YUV_I420)
            nv12[i, :Hh] = i420[:Hh]
            nv12[i, Hh:] = np.stack([i420[Hh:Hh + Hh // 4].reshape(Hh // 2, Ww // 2), i420[Hh + Hh // 4:].reshape(Hh // 2, Ww // 2)],
                                    axis=-1).reshape(Hh // 2, Ww)
            frames_np[i] = cv2.cvtColor(nv12[i], cv2.COLOR_YUV2RGB_NV12)
        fp_host = pipeline.FacePipeline(det, enc, cls, (160, 160), "similarity", max_faces_per_frame=wl["max_faces"], input_format="nv12")
        frames_pinned = torch.from_numpy(nv12).pin_memory()
        frames_dev = torch.from_numpy(frames_np).to(dev)
    else:
        frames_pinned = torch.from_numpy(frames_np).pin_memory()
        frames_dev = frames_pinned.to(dev)

    stage_ms = {}
    ev_log = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev_log.append((name, e))

    gather_stream = torch.cuda.Stream(dev) if world > 1 else None
    # the exchange: one NCCL all_gather_into_tensor of the send buffers.  VNFR_PEER_GATHER=1: copy-engine pull from the peers'
    # symmetric send buffers instead (dist.PeerGather: no collective kernel; measured 9.30 ms per step on 2 GPUs against 9.11
    # with NCCL and 8.92 without any exchange, so the SM footprint of the NCCL kernel is not what the exchange costs)
    peer_gather = None
    if world > 1 and os.environ.get("VNFR_PEER_GATHER") and not os.environ.get("VNFR_RAGGED_GATHER"):
        peer_gather = vdist.PeerGather().attach(fp)
    gather_events = []          # exchange i has read its send buffer: the step that reuses the buffer (i + 2) waits for it
    gather_out = [None, None]

    def step_device(timed):
        # production path: a stream of batches (the cascade of step i+1 may start under the encoder of step i; every step
        # still runs all of its own work); the instrumented passes run one batch at a time on one stream
        if len(gather_events) >= 2 and gather_events[-2] is not None:
            torch.cuda.current_stream().wait_event(gather_events[-2])
        out = fp.run_device(frames_dev, mark=mark if timed else None, pipelined=not timed and not args.no_pipeline)
        if world > 1 and not os.environ.get("VNFR_BENCH_NO_EXCHANGE"):      # (diagnostic switch: what the exchange costs)
            if os.environ.get("VNFR_RAGGED_GATHER"):
                vdist.all_gather_faces(out["emb"], out["label"], out["prob"])
            else:
                # the step's exchange: ONE all_gather_into_tensor straight from the send buffer the fused tail kernel filled
                # (no packing ops, no host sync), on a side stream: the collective waits for the slowest rank, the next
                # step's kernels need not
                side = None if (args.no_pipeline or os.environ.get("VNFR_GATHER_MAIN")) else gather_stream
                slot = len(gather_events) & 1
                if gather_out[slot] is None:
                    gather_out[slot] = torch.empty(world * out["payload"].shape[0], out["payload"].shape[1], device=dev)
                if peer_gather is not None:
                    out["gathered"], _, ev = peer_gather.gather(out["payload"], stream=side, out=gather_out[slot])
                else:
                    out["gathered"], _, ev = vdist.all_gather_payload(out["payload"], stream=side, out=gather_out[slot])
                gather_events.append(ev)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`)
    for _ in range(args.warmup):
        out = step_device(False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    faces = 0
    prof_range = bool(os.environ.get("VNFR_PROFILE_RANGE"))     # ncu --profile-from-start off: capture the timed steps only
    if prof_range:
        torch.cuda.profiler.start()
    t0.record()
    for _ in range(args.steps):
        out = step_device(False)
        faces += out["n_faces"]
    if gather_stream is not None:
        torch.cuda.current_stream().wait_stream(gather_stream)       # the last step's exchange is inside the timed region
    t1.record()
    barrier()
    if prof_range:
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.launch_count() - l0
    ms = t0.elapsed_time(t1)
    # per-stage device time: separate instrumented passes (stage markers force the single-stream cascade; the timed loop
    # above runs the production path, whose two detection half-batches overlap on two streams)
    n_inst = 3
    for _ in range(2):                     # untimed: the single-stream path allocates its own workspaces on first use
        step_device(True)
    barrier()
    ev_log.clear()
    for _ in range(n_inst):
        out = step_device(True)
    barrier()
    for (n0, e0), (n1, e1) in zip(ev_log[:-1], ev_log[1:]):
        if n1 != "start":
            stage_ms[n1] = stage_ms.get(n1, 0.0) + e0.elapsed_time(e1) / n_inst

    # ---- end to end through the public API: pinned host frames in, host results out (`e2e`)
    e2e_steps = 1 if args.skip_e2e else args.steps

    def measure_e2e(fph, pinned):
        for _ in range(0 if args.skip_e2e else 3):
            fph(pinned)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nf, res = 0, None
        e0.record()
        if args.no_pipeline:
            for _ in range(e2e_steps):
                res = fph(pinned)
                nf += sum(len(r["labels"]) for r in res)
        else:
            # the streaming form of the public call: two batches in flight -- batch i+1 is submitted (its H2D copy and cascade
            # start under the encoder of batch i), then the results of batch i are collected on the host.  Every step's
            # frames are copied from pinned host memory and every step's results are read back inside the timed region.
            pending = None
            for _ in range(e2e_steps):
                nxt = fph.submit(pinned)
                if pending is not None:
                    res = pending.result()
                    nf += sum(len(r["labels"]) for r in res)
                pending = nxt
            res = pending.result()
            nf += sum(len(r["labels"]) for r in res)
        e1.record()
        barrier()
        return e0.elapsed_time(e1), nf, res

    ms_e2e, faces_e2e, res = measure_e2e(fp_host, frames_pinned)
    nf_step = faces_e2e // max(1, e2e_steps)
    # face counts + status, the (B, capf, 5) box tensor (copied whole: contiguous async copy), label / prob / embedding per face
    d2h = (B + 1) * 4 + B * det.caps[3] * 5 * 4 + nf_step * (8 + 4 + 512 * 4)
    # the same measurement with the host frames in the decoder's native NV12 (half the bytes over PCIe; extra key `e2e_nv12`):
    # at 8 GPUs the RGB frames of all ranks (8 x 398 MB per 10 ms step) exceed what one host can push (~170 GB/s measured)
    ms_nv12 = faces_nv12 = None
    if args.ingest == "rgb" and not args.skip_e2e and not args.no_nv12:
        import cv2
        Hh, Ww = frames_np.shape[1:3]
        nv12 = np.empty((B, Hh * 3 // 2, Ww), np.uint8)
        for i in range(B):
            i420 = cv2.cvtColor(frames_np[i], cv2.COLOR_RGB2YUV_I420)
            nv12[i, :Hh] = i420[:Hh]
            nv12[i, Hh:] = np.stack([i420[Hh:Hh + Hh // 4].reshape(Hh // 2, Ww // 2), i420[Hh + Hh // 4:].reshape(Hh // 2, Ww // 2)],
                                    axis=-1).reshape(Hh // 2, Ww)
        fp_nv = pipeline.FacePipeline(det, enc, cls, (160, 160), "similarity", max_faces_per_frame=wl["max_faces"], input_format="nv12")
        nv_pinned = torch.from_numpy(nv12).pin_memory()
        ms_nv12, faces_nv12, _ = measure_e2e(fp_nv, nv_pinned)

    # ---- second headline figure of BASELINE.json ("embeds/sec"): InceptionResnetV1 + L2-norm + MLP classify on
    # batch-1024 synthetic 160x160 crops (config 2) through the public forward() API, crops resident on the device
    embed = None
    if not args.skip_e2e and args.config == 3:
        torch.manual_seed(1)
        crops = torch.randn(args.embed_batch, 3, 160, 160, device=dev).clamp_(-1, 1)
        for _ in range(3):
            lp = cls(enc(crops))
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(5):
            lp = cls(enc(crops))
        g1.record()
        torch.cuda.synchronize()
        ms_emb = g0.elapsed_time(g1) / 5
        embed = {"value": args.embed_batch / (ms_emb * 1e-3) * world, "unit": "embeds/s", "batch_per_rank": args.embed_batch,
                 "ms_per_batch": ms_emb, "tflops": args.embed_batch * (ENC_FLOP_PER_FACE + MLP_FLOP_PER_FACE) / (ms_emb * 1e-3) / 1e12,
                 "note": "InceptionResnetV1.forward + MLPModel.forward on (B,3,160,160) fp32 device tensors"}

    # ---- config 5 figure: cosine top-5 of a batch of embeddings against this rank's gallery shard (+ NCCL merge)
    topk = None
    if args.config == 3 and args.gallery_rows == 0:
        args.gallery_rows = 131072
    if not args.skip_e2e and args.gallery_rows > 0 and args.config == 3:
        from vn_celeb_face_recognition_b200 import gallery
        gen = torch.Generator(device=dev).manual_seed(2 + rank)
        gshard = gallery.GalleryShard(torch.nn.functional.normalize(torch.randn(args.gallery_rows, 512, device=dev, generator=gen), dim=1),
                                      index_offset=rank * args.gallery_rows)
        qs = torch.nn.functional.normalize(torch.randn(args.gallery_queries, 512, device=dev, generator=gen), dim=1)
        for _ in range(2):
            tv, ti = gallery.merge_topk(*gshard.topk(qs, 5))
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(3):
            tv, ti = gallery.merge_topk(*gshard.topk(qs, 5))
        k1.record()
        torch.cuda.synchronize()
        ms_k = k0.elapsed_time(k1) / 3
        topk = {"queries_per_s": args.gallery_queries / (ms_k * 1e-3), "queries": args.gallery_queries,
                "gallery_rows_total": args.gallery_rows * world, "k": 5, "ms": ms_k,
                "tflops": 2.0 * args.gallery_queries * args.gallery_rows * 512 / (ms_k * 1e-3) / 1e12,
                "note": "every rank scores the same queries against its gallery shard, then all_gather + merge"}
        del gshard

    # ---- reductions over ranks
    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_nv12 or 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_nv12_max = t.tolist()
        c = torch.tensor([faces, faces_e2e, launches, faces_nv12 or 0], device=dev, dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        faces, faces_e2e, launches, faces_nv12_sum = [int(v) for v in c.tolist()]
        if ms_nv12 is not None:
            ms_nv12, faces_nv12 = ms_nv12_max, faces_nv12_sum

    if rank == 0:
        ws = out["ws"]
        cnts = ws.counters.cpu().numpy()
        nseg = ws.B * ws.L
        work = {"pnet_candidates_per_frame": float(np.minimum(cnts[:nseg], ws.caps[0]).sum()) / B,
                "rnet_crops_per_frame": float(cnts[2 * nseg:2 * nseg + B].sum()) / B,
                "onet_crops_per_frame": float(cnts[2 * nseg + B:2 * nseg + 2 * B].sum()) / B,
                "faces_per_frame": float(cnts[2 * nseg + 2 * B:2 * nseg + 3 * B].sum()) / B}
        faces_step_rank = out["n_faces"]
        # roofline of the dominant kernel: the tcgen05 implicit-GEMM convolution (encoder stage)
        enc_ms = stage_ms.get("encoder", 0.0)
        conv_launches_per_step = None
        roof = None
        if enc_ms > 0:
            flops = faces_step_rank * ENC_FLOP_PER_FACE
            achieved = flops / (enc_ms * 1e-3) / 1e12
            # DRAM bytes of those launches from the committed ncu pass (profiles/): same command, same workload
            traffic, traffic_note = None, None
            tp = os.path.join(ROOT, "profiles", "encoder_traffic.json")
            if os.path.exists(tp) and faces_step_rank > 0:
                tj = json.load(open(tp))
                # activations dominate (weights: 47 MB, L2-resident), so the stage's DRAM bytes scale with the faces of the step
                traffic = tj["dram_bytes_per_step"] * faces_step_rank / tj["faces"]
                traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum summed over the %d encoder launches of one step of %d "
                                "faces (%.1f MB per launch)%s, %s" % (tj["conv_launches_per_step"], tj["faces"],
                                                                      tj["dram_bytes_per_launch"] / 1e6,
                                                                      "" if tj["faces"] == faces_step_rank else
                                                                      ", scaled to this step's %d faces" % faces_step_rank, tj["source"]))
            roof = {"kernel": "tcgen05 convolutions (sv_conv_kernel + igemm_conv_kernel + block17_fused_kernel): all conv launches of "
                              "the InceptionResnetV1 stage of one step", "bound": "tensor",
                    "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                    "traffic": traffic, "traffic_note": traffic_note,
                    "peak_source": peaks["source"] + ", sustained (kernel timed inside a long step)",
                    "stage_ms": enc_ms, "share_of_step": enc_ms / sum(stage_ms.values()),
                    "note": "stage times come from instrumented single-stream passes; the timed loop overlaps the two "
                            "detection half-batches on two streams, so ms_per_step < sum(stage_ms)"}
        # the detection kernels against their own roofs (config 4's step is dominated by them)
        det_roofs = []
        if stage_ms.get("pnet", 0) > 0:
            a = B * wl["pnet_gflop"] / stage_ms["pnet"]               # GFLOP / ms = TFLOP/s
            det_roofs.append({"kernel": "pnet_kernel (fp32 FMA conv1/conv2 + split-precision tcgen05 conv3, all pyramid levels)",
                              "bound": "tensor", "achieved": a, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                              "frac": a / peaks["bf16_sustained"], "stage_ms": stage_ms["pnet"],
                              "note": "%.2f GFLOP of fp32-accurate work per frame; conv3 (57 %%) costs 3 fp16 products per fp32 product, "
                                      "conv1/conv2 run on the FMA pipe" % wl["pnet_gflop"]})
        if stage_ms.get("pyramid", 0) > 0:
            a = B * wl["pyramid_mb"] / stage_ms["pyramid"]            # MB / ms = GB/s
            det_roofs.append({"kernel": "pyramid_strip_kernel (all levels, one read of the frame)", "bound": "hbm", "achieved": a,
                              "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": a / peaks["hbm_gbs"], "stage_ms": stage_ms["pyramid"],
                              "note": "%.2f MB per frame (frame read once + levels written)" % wl["pyramid_mb"]})
        if roof is not None and det_roofs:
            dom = max([roof] + det_roofs, key=lambda r: r["stage_ms"])
            if dom is not roof:                                       # the JSON's `roofline` is the step's dominant kernel
                det_roofs = [r for r in det_roofs if r is not dom] + [roof]
                roof = dict(dom, traffic=None, peak_source=peaks["source"], share_of_step=dom["stage_ms"] / sum(stage_ms.values()))
        line = {"metric": wl["metric"], "value": faces / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": _half_name(enc.half_dtype), "data": "synthetic",
                "config": {"workload": wl["workload"], "frames_per_rank": B, "frame": wl["frame"],
                           "faces_per_step": faces // max(1, args.steps), "min_face_size": wl["min_face"], "align": "similarity 160x160",
                           "encoder": "InceptionResnetV1 random-init", "classifier": "MLPModel(512,1001) random-init",
                           "detector_weights": "bundled MTCNN", "detector_dtype": "f32", "encoder_chunk": enc.chunk,
                           "l2_policy": "inputs larger than L2 (%.0f MB of frames per step)" % (frames_np.nbytes / 1e6),
                           "batches_in_flight": 1 if args.no_pipeline else 2, "work_per_frame": work, "collective": ("none" if world == 1 else ("all_gather(emb,label,prob): copy-engine pull from symmetric peer buffers"
                                                                      if os.environ.get("VNFR_PEER_GATHER") and not os.environ.get("VNFR_RAGGED_GATHER")
                                                                      else "all_gather(emb,label,prob)"))},
                "e2e": {"value": faces_e2e / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(frames_pinned.numel()), "ingest": args.ingest,
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / e2e_steps,
                        "api": "FacePipeline.__call__ (one batch at a time)" if args.no_pipeline else
                               "FacePipeline.submit / PendingResult.result, two batches in flight"},
                "e2e_nv12": None if ms_nv12 is None else {
                    "value": faces_nv12 / (ms_nv12 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(frames_np.nbytes // 2),
                    "ms_per_step": ms_nv12 / e2e_steps,
                    "note": "the same e2e measurement with NV12 host frames (a video decoder's native output, 1.5 B/px) converted on the "
                            "device bit-identically to cv2.cvtColor(COLOR_YUV2RGB_NV12): FacePipeline(input_format='nv12')"},
                "gpu_launches": int(launches), "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
                "roofline": roof, "roofline_other": det_roofs, "clocks": clocks, "embed": embed, "gallery_topk": topk}
        if world == 1 and not args.no_cpu_baseline:
            enc_sd = {k: v.detach().float().cpu() for k, v in enc.state_dict().items()}
            mlp_sd = {k: v.detach().float().cpu() for k, v in cls.state_dict().items()}
            n = args.cpu_frames or wl["cpu_frames"]
            fps, nf, dt, threads, kind, outputs = cpu_reference_leg(frames_np[:n], enc_sd, mlp_sd, repeats=2 if args.config == 3 else 1,
                                                                    min_face_size=wl["min_face"], want_outputs=True)
            eps, edt = cpu_embed_leg(enc_sd, mlp_sd, threads=threads)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": "first %d of the step's frames (%d faces), %.2f s" % (n, nf, dt),
                                    "embeds_per_s": eps, "embeds_sample": "InceptionResnetV1 + MLP on 64 crops, best of 2, %.2f s" % edt}
            # parity of the measured step itself: the e2e results of the same frames against the CPU reference path
            line["parity"] = label_parity(res, outputs)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _half_name(dt=None):
    """Arithmetic type of the dominant (tensor-core) part: 16-bit operands, fp32 accumulation.  The detector runs fp32."""
    from vn_celeb_face_recognition_b200 import encoder_plan
    import torch
    return "f16" if (dt or encoder_plan.HALF) == torch.float16 else "bf16"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 3 = 1080p pipeline (default, the headline), 4 = crowded 4K / min_face_size 20, "
                         "2 = embed + classify on 160x160 crops (embeds/s), 5 = offline embedding + sharded cosine top-5")
    ap.add_argument("--frames", type=int, default=0, help="frames per rank per step (default: the config's: 64 x 1080p / 8 x 4K)")
    ap.add_argument("--chunk", type=int, default=1024, help="encoder crops per internal chunk")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the bounded CPU sample of our arm's cpu_baseline / parity block (default 8 / 1)")
    ap.add_argument("--ref-frames", type=int, default=0, help="--impl reference: frames per step (default: our arm's frames per rank)")
    ap.add_argument("--gallery-rows", type=int, default=0, help="gallery rows per rank (config 5: default 125 000; config 3's extra "
                    "gallery_topk figure: default 131 072, -1 = skip)")
    ap.add_argument("--gallery-queries", type=int, default=8192)
    ap.add_argument("--crops-per-rank", type=int, default=122880, help="config 5: crops embedded per rank per step (1 M / 8 GPUs = 125 000; "
                    "rounded down to whole chunks of 4096)")
    ap.add_argument("--embed-batch", type=int, default=1024, help="crops per rank of the embeds/s measurement (config 2)")
    ap.add_argument("--ingest", default="rgb", choices=["rgb", "nv12"], help="host frame format of the e2e leg: packed RGB (what the "
                    "reference's cap.read + cvtColor hands over) or NV12 as a video decoder delivers it (half the H2D bytes, converted on the device)")
    ap.add_argument("--no-nv12", action="store_true", help="skip the extra e2e_nv12 measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: one un-warmed e2e step")
    ap.add_argument("--no-pipeline", action="store_true", help="one batch at a time: no overlap between consecutive steps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.config in (2, 5):
        import bench_embed
        (bench_embed.run_reference if args.impl == "reference" else bench_embed.run_ours)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
