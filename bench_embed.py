"""bench.py --config 2 / 5: the embedding-side configs of BASELINE.json (imported by bench.py; same JSON contract).

config 2  InceptionResnetV1 embedding + L2-norm + MLP classify on synthetic 160x160 aligned crops, batch 1024 per rank
          (metric: embeds/sec).  A step = one batch of u8 crops -> transforms_default -> encoder -> fused tail (labels).
config 5  offline embedding of 125 000 crops per rank + cosine top-5 against a 125 000-row gallery shard per rank, queries
          all-gathered over NCCL, per-shard top-5 lists exchanged and merged (metric: embeds/sec incl. the search).
"""
import json
import os
import time

import numpy as np

import bench as B

ENC_FLOP = B.ENC_FLOP_PER_FACE
METRIC2 = "embeds/sec InceptionResnetV1 + L2-norm + MLP classify (160x160 crops, batch 1024)"
METRIC5 = "embeds/sec offline embedding + cosine top-5 vs sharded gallery"


def crops_u8(n, seed):
    """(n,160,160,3) uint8 aligned crops: the fp32 config-2 crops of synthetic.crops_160 mapped back to bytes."""
    import torch
    from vn_celeb_face_recognition_b200 import synthetic
    x = synthetic.crops_160(n, seed=seed)
    return (x * 128.0 + 127.5).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def cpu_leg(enc_sd, mlp_sd, crops, threads=None):
    """The reference's recognize_celeb model calls on the host cores (transforms_default -> encoder -> MLP -> argmax): the
    unmodified reference modules when the tree is present, else the oracle port.  Returns (embeds/s, seconds, kind, labels)."""
    import torch
    from oracle import nets, pipeline as opipe
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    kind = B.reference_kind()
    x = torch.stack([opipe.transforms_default(f) for f in crops.numpy()])
    with torch.no_grad():
        if kind == "reference":
            from oracle import ref_shims
            ref = ref_shims.load_reference()
            enc = ref.models.InceptionResnetV1(pretrained=None, device="cpu").eval(); enc.load_state_dict(enc_sd)
            mlp = ref.models.MLPModel(512, mlp_sd["dense_2.weight"].shape[0]).eval(); mlp.load_state_dict(mlp_sd)
            f = lambda: mlp(ref.demo_image.find_embedding(x, enc))
        else:
            f = lambda: nets.mlp_forward(mlp_sd, nets.encoder_forward(enc_sd, x))
        f() if len(x) <= 64 else None                       # warm-up on small samples only
        t0 = time.perf_counter()
        lp = f()
        dt = time.perf_counter() - t0
    return len(x) / dt, dt, kind, lp.argmax(1).numpy(), threads


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from vn_celeb_face_recognition_b200 import synthetic
    if args.config == 5:
        print(json.dumps({"impl": "reference", "metric": METRIC5, "unavailable": "the reference has no gallery search (SURVEY.md 8d); "
                          "its embedding side is timed by --config 2 --impl reference"}))
        return
    enc_sd, mlp_sd = synthetic.encoder_state_dict_seed0(), synthetic.mlp_state_dict(1001, seed=0)
    n = args.embed_batch
    crops = crops_u8(n, seed=1)
    times = []
    for i in range(args.warmup + args.steps):
        eps, dt, kind, _, threads = cpu_leg(enc_sd, mlp_sd, crops)
        if i >= args.warmup:
            times.append(dt)
    val = n * len(times) / sum(times)
    print(json.dumps({"metric": METRIC2, "value": val, "unit": "embeds/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "embed_classify_160 (BASELINE config 2)", "batch_per_rank": n, "crop": "160x160x3 u8"},
                      "cpu_baseline": {"value": val, "unit": "embeds/s", "cores": threads, "kind": kind,
                                       "sample": "%d crops per step = our arm's batch, %d timed steps" % (n, args.steps)},
                      "e2e": {"value": val, "unit": "embeds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from vn_celeb_face_recognition_b200 import _lib, pipeline, synthetic
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this framework has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.config == 5:
        return run_config5(args, dev, world, rank)
    peaks = B.load_peaks()
    _, enc, cls = B.build_models(dev)
    enc.chunk = max(args.chunk, args.embed_batch)
    fp = pipeline.FacePipeline(None, enc, cls, (160, 160), "extract")
    n = args.embed_batch
    host = [crops_u8(n, seed=1 + rank * 2 + i).pin_memory() for i in range(2)]       # two alternating batches
    devb = [h.to(dev) for h in host]
    cap = n
    payloads = [torch.zeros(cap + 1, 514, device=dev) for _ in range(2)]
    ev_log = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev_log.append((name, e))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident
    for i in range(args.warmup):
        fp.embed_faces(devb[i & 1], payload=payloads[i & 1])
    barrier()
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        fp.embed_faces(devb[i & 1], payload=payloads[i & 1])
    t1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.launch_count() - l0
    ms = t0.elapsed_time(t1)
    # encoder stage time (events around the convolution graph)
    enc_ms = 0.0
    for i in range(3):
        ev_log.clear()
        mark("start")
        fp.embed_faces(devb[i & 1], payload=payloads[i & 1], mark=mark)
        torch.cuda.synchronize()
        enc_ms += ev_log[0][1].elapsed_time(ev_log[1][1]) / 3
    # ---- end to end: pinned u8 crops in, (emb | label | prob) rows out; the H2D of batch i+1 overlaps the compute of batch i
    copy_stream = torch.cuda.Stream(dev)
    stage_in = [torch.empty_like(devb[0]) for _ in range(2)]
    rows_host = [torch.empty(n, 514).pin_memory() for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i & 1])
            stage_in[i & 1].copy_(host[i & 1], non_blocking=True)
            copied[i & 1].record(copy_stream)

    def e2e_loop(k):
        for e in consumed:
            e.record()
        upload(0)
        for i in range(k):
            if i + 1 < k:
                upload(i + 1)
            torch.cuda.current_stream().wait_event(copied[i & 1])
            fp.embed_faces(stage_in[i & 1], payload=payloads[i & 1])
            consumed[i & 1].record()
            rows_host[i & 1].copy_(payloads[i & 1][:n], non_blocking=True)
        torch.cuda.synchronize()

    e2e_loop(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank == 0:
        achieved = n * ENC_FLOP / (enc_ms * 1e-3) / 1e12
        line = {"metric": METRIC2, "value": world * n * args.steps / (ms * 1e-3), "unit": "embeds/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": B._half_name(enc.half_dtype), "data": "synthetic",
                "config": {"workload": "embed_classify_160 (BASELINE config 2)", "batch_per_rank": n, "crop": "160x160x3 u8",
                           "encoder": "InceptionResnetV1 random-init (seeded, BN-calibrated)", "classifier": "MLPModel(512,1001) random-init",
                           "l2_policy": "two alternating input batches; a step's 2.7 GB of activations evict the inputs",
                           "collective": "none (crops are independent)"},
                "e2e": {"value": world * n * args.steps / (ms_e2e * 1e-3), "unit": "embeds/s", "h2d_bytes_per_step": int(host[0].numel()),
                        "d2h_bytes_per_step": n * 514 * 4, "ms_per_step": ms_e2e / args.steps,
                        "api": "FacePipeline.embed_faces on pinned u8 crops, upload of batch i+1 under the compute of batch i"},
                "gpu_launches": int(launches) * world,
                "roofline": {"kernel": "tcgen05 convolutions of the InceptionResnetV1 stage", "bound": "tensor", "achieved": achieved,
                             "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"], "traffic": None,
                             "peak_source": peaks["source"] + ", sustained", "stage_ms": enc_ms},
                "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            enc_sd = {k: v.detach().float().cpu() for k, v in enc.state_dict().items()}
            mlp_sd = {k: v.detach().float().cpu() for k, v in cls.state_dict().items()}
            m = 64
            eps, dt, kind, lab_ref, threads = cpu_leg(enc_sd, mlp_sd, host[0][:m])
            line["cpu_baseline"] = {"value": eps, "unit": "embeds/s", "cores": threads, "kind": kind,
                                    "sample": "first %d crops of the step's batch, %.2f s" % (m, dt)}
            got = rows_host[(args.steps - 1) & 1] if (args.steps - 1) & 1 == 0 else None
            res = fp.embed_faces(devb[0], payload=payloads[0])
            torch.cuda.synchronize()
            line["parity"] = {"crops": m, "label_agree": int((res["label"][:m].cpu().numpy() == lab_ref).sum())}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_config5(args, dev, world, rank):
    """Offline embedding of ``--crops-per-rank`` crops per rank + cosine top-5 of every crop's embedding against the whole
    gallery (``--gallery-rows`` rows per rank, sharded): all_gather of the query embeddings, ONE fused score-GEMM + top-k launch
    against the local shard, all_to_all of the per-shard lists, merge.  (The crops of a step cycle through a pool of 4096
    distinct pinned crops: 1 M distinct crops per box would need 77 GB of pinned host memory.)"""
    import torch
    import torch.distributed as dist
    from vn_celeb_face_recognition_b200 import _lib, pipeline, gallery, encoder_plan
    peaks = B.load_peaks()
    _, enc, cls = B.build_models(dev)
    chunk = 4096
    enc.chunk = chunk
    fp = pipeline.FacePipeline(None, enc, None, (160, 160), "extract")
    n = (args.crops_per_rank // chunk) * chunk
    n_chunks = n // chunk
    g_rows = args.gallery_rows or 125000
    pool_host = crops_u8(chunk, seed=11 + rank).pin_memory()
    pool_dev = pool_host.to(dev)
    gen = torch.Generator(device=dev).manual_seed(2 + rank)
    shard = gallery.GalleryShard(torch.nn.functional.normalize(torch.randn(g_rows, 512, device=dev, generator=gen), dim=1),
                                 index_offset=rank * g_rows)
    dt = enc.half_dtype or encoder_plan.HALF
    q_local = torch.empty(n, 512, dtype=dt, device=dev)
    q_all = torch.empty(world * n, 512, dtype=dt, device=dev) if world > 1 else q_local
    stage_in = [torch.empty_like(pool_dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(dev)
    copied, consumed = [torch.cuda.Event() for _ in range(2)], [torch.cuda.Event() for _ in range(2)]
    top_host = (torch.empty(n, 5).pin_memory(), torch.empty(n, 5, dtype=torch.int64).pin_memory())
    times = {"embed": 0.0, "search": 0.0}

    def search():
        if world > 1:
            dist.all_gather_into_tensor(q_all, q_local)
        v, i = shard.topk(q_all, 5)                                   # (world*n, 5) against the local shard
        if world > 1:
            # rank s needs the lists of ITS queries from every shard: all_to_all of (n, 5) blocks, then merge
            rv, ri = torch.empty_like(v), torch.empty_like(i)
            dist.all_to_all_single(rv, v)
            dist.all_to_all_single(ri, i)
            v = rv.view(world, n, 5).permute(1, 0, 2).reshape(n, world * 5)
            i = ri.view(world, n, 5).permute(1, 0, 2).reshape(n, world * 5)
            v, i = _merge(v, i)
        return v, i

    def _merge(v, i):
        order = torch.argsort(i, dim=1, stable=True)
        v, i = torch.gather(v, 1, order), torch.gather(i, 1, order)
        order = torch.argsort(v, dim=1, descending=True, stable=True)
        return torch.gather(v, 1, order)[:, :5].contiguous(), torch.gather(i, 1, order)[:, :5].contiguous()

    def step(from_host, timed=None):
        if from_host:
            for e in consumed:
                e.record()

            def upload(c):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[c & 1])
                    stage_in[c & 1].copy_(pool_host, non_blocking=True)
                    copied[c & 1].record(copy_stream)
            upload(0)
        t_a = torch.cuda.Event(enable_timing=True); t_b = torch.cuda.Event(enable_timing=True); t_c = torch.cuda.Event(enable_timing=True)
        t_a.record()
        for c in range(n_chunks):
            if from_host:
                if c + 1 < n_chunks:
                    upload(c + 1)
                torch.cuda.current_stream().wait_event(copied[c & 1])
                src = stage_in[c & 1]
            else:
                src = pool_dev
            res = fp.enc.embed_s2d(_s2d(src), 160, want_half=True)
            if from_host:
                consumed[c & 1].record()
            q_local[c * chunk:(c + 1) * chunk].copy_(res["emb16"])
        t_b.record()
        v, i = search()
        if from_host:
            top_host[0].copy_(v, non_blocking=True)
            top_host[1].copy_(i, non_blocking=True)
        t_c.record()
        if timed is not None:
            torch.cuda.synchronize()
            timed["embed"] += t_a.elapsed_time(t_b)
            timed["search"] += t_b.elapsed_time(t_c)
        return v, i

    s2d_buf = torch.zeros(chunk, 80, 80, 16, dtype=dt, device=dev)

    def _s2d(src):
        _lib.call("vnfr_u8hwc_to_s2d16", _lib.ptr(src), chunk, 160, 160, _lib.ptr(s2d_buf), encoder_plan.dtype_code(dt), _lib.stream_ptr())
        return s2d_buf

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            step(False)
        barrier()
        sampler = B.ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
        if rank == 0:
            sampler.start()
        l0 = _lib.launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            v, i = step(False)
        t1.record()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        launches = _lib.launch_count() - l0
        ms = t0.elapsed_time(t1)
        step(False, times)
        for _ in range(2):
            step(True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step(True)
        torch.cuda.synchronize()
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank == 0:
        enc_tf = n * ENC_FLOP / (times["embed"] * 1e-3) / 1e12
        gal_tf = 2.0 * world * n * g_rows * 512 / (times["search"] * 1e-3) / 1e12
        line = {"metric": METRIC5, "value": world * n * args.steps / (ms * 1e-3), "unit": "embeds/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": B._half_name(enc.half_dtype), "data": "synthetic",
                "config": {"workload": "offline_embed_gallery_top5 (BASELINE config 5)", "crops_per_rank": n, "gallery_rows_per_rank": g_rows,
                           "gallery_rows_total": g_rows * world, "k": 5, "chunk": chunk, "crop": "160x160x3 u8",
                           "l2_policy": "a chunk's activations (11 GB) evict its inputs",
                           "collective": "all_gather(query embeddings) + all_to_all(top-5 lists)" if world > 1 else "none"},
                "e2e": {"value": world * n * args.steps / (ms_e2e * 1e-3), "unit": "embeds/s", "h2d_bytes_per_step": int(n * 76800),
                        "d2h_bytes_per_step": int(n * 5 * 12), "ms_per_step": ms_e2e / args.steps,
                        "api": "pinned u8 crop chunks -> embed_s2d -> GalleryShard.topk (+ NCCL exchange) -> pinned top-5"},
                "gpu_launches": int(launches) * world, "stage_ms": {k: round(x, 3) for k, x in times.items()},
                "roofline": {"kernel": "tcgen05 convolutions of the InceptionResnetV1 stage", "bound": "tensor", "achieved": enc_tf,
                             "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": enc_tf / peaks["bf16_sustained"], "traffic": None,
                             "peak_source": peaks["source"] + ", sustained", "stage_ms": times["embed"]},
                "roofline_other": [{"kernel": "gallery_topk_kernel (score GEMM with fused top-8, scores never leave TMEM) + exchange + merge",
                                    "bound": "tensor", "achieved": gal_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                                    "frac": gal_tf / peaks["bf16_sustained"], "stage_ms": times["search"]}],
                "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            enc_sd = {k: x.detach().float().cpu() for k, x in enc.state_dict().items()}
            mlp_sd = {k: x.detach().float().cpu() for k, x in cls.state_dict().items()}
            eps, dtm, kind, _, threads = cpu_leg(enc_sd, mlp_sd, pool_host[:64])
            # parity of the search on the first queries: exact fp32 top-5 of the same 16-bit operands on the host
            qh = q_local[:256].float().cpu()
            gh = shard.w[:g_rows].float().cpu()
            rv, ri = torch.topk(qh @ gh.t(), 5, dim=1)
            agree = float((ri == i[:256].cpu()).float().mean())
            line["cpu_baseline"] = {"value": eps, "unit": "embeds/s", "cores": threads, "kind": kind,
                                    "sample": "embedding side only (the reference has no gallery search): 64 crops, %.2f s" % dtm}
            line["parity"] = {"queries": 256, "top5_index_agreement_vs_fp32_host": agree}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
