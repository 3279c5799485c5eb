#!/usr/bin/env python
"""bench.py - decisions/sec of BUSCA's per-frame association hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port) on host cores

A "step" is one frame of the hot path: motion proposals + centre-distance/IoU + candidate selection, crop-and-resize
gather of the D detections and T motion proposals, ReID embedding of T*(L+C) patches (two BatchNorm batches), the
Decision Transformer and the decision, for T unmatched tracks.  One decision = one unmatched track evaluated in one
frame, so a step yields T decisions.

`value`  : inputs already resident in HBM (frame, patch bank, boxes): busca_frame_step_dev.
`e2e`    : the same step through the reference-facing plug-in API (BUSCA.get_image_crops / center_distance /
           associate_embeddings) with HOST buffers; H2D of the frame/boxes and D2H of crops/probabilities are inside
           the timed region.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from busca_b200 import sharding, synth  # noqa: E402
from busca_b200.scene import Scene  # noqa: E402

WEIGHTS_PROFILE = "conditioned"      # synth.make_weights: decisions vary from track to track (tests/golden/scene_mot20_cond.npz pins this workload)

WORKLOADS = {
    # BASELINE.json configs[2]: MOT20-scale dense crowd, ~200 unmatched tracks/frame (the scale the metric is quoted on)
    "mot20": dict(T=200, D=300, L=11, C=5, desc="MOT20-scale synthetic 1920x1080 frames, 200 unmatched tracks x 5 proposals, 300 detections, L=11"),
    # configs[1]
    "mot17": dict(T=50, D=60, L=11, C=5, desc="MOT17-scale synthetic 1920x1080 frames, 50 unmatched tracks x 5 proposals, 60 detections, L=11"),
    # configs[4]: long-history stress (40,000 stacked ReID patches per frame, S = 54 tokens per track; 130 GB of ReID workspace)
    "stress": dict(T=1000, D=300, L=30, C=10, desc="long-history stress: 1000 unmatched tracks x 10 proposals, 300 detections, L=30 history frames"),
    # configs[0]
    "cfg1": dict(T=16, D=40, L=11, C=5, desc="1 synthetic 1920x1080 frame, 16 unmatched tracks x 5 proposals, 40 detections"),
}


def conv_flops(n_patches: int):
    """Algorithmic FLOPs (2*MACs) of the ReID convolutions for n_patches, by kernel class:
    8.0075 GFLOP per 384x128 patch in total."""
    out = {"stem_conv7x7": 2.0 * 192 * 64 * 64 * 3 * 49, "conv1x1": 0.0, "conv3x3": 0.0}
    H, W, inpl = 96, 32, 64
    for planes, blocks, stride in synth.RESNET_LAYERS:
        for b in range(blocks):
            s = stride if b == 0 else 1
            Ho, Wo = H // s, W // s
            out["conv1x1"] += 2.0 * H * W * inpl * planes               # conv1
            out["conv3x3"] += 2.0 * Ho * Wo * planes * planes * 9       # conv2 (carries the stride)
            out["conv1x1"] += 2.0 * Ho * Wo * planes * planes * 4       # conv3
            if b == 0:
                out["conv1x1"] += 2.0 * Ho * Wo * inpl * planes * 4     # downsample
            inpl, H, W = planes * 4, Ho, Wo
    return {k: v * n_patches for k, v in out.items()}


class ClockSampler:
    """nvidia-smi clocks line of the profiling recipe, sampled DURING the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._th = None
        self._proc = None

    def start(self):
        try:
            self._proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                           "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self._proc = None
            return
        self._th = threading.Thread(target=self._read, daemon=True)
        self._th.start()

    def _read(self):
        for line in self._proc.stdout:
            if self._stop.is_set():
                break
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        self._stop.set()
        if self._proc:
            self._proc.terminate()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(self.rows)}
        sm = []
        reasons = set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            busy = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
            out["sm_mhz"] = float(np.median(busy))
        out["reasons"] = sorted(reasons)
        return out


# --------------------------------------------------------------------------------------------------------
def build_model(precision: str, device: int, bank_slots: int = 8192):
    from busca_b200.network import BUSCA
    from busca_b200.option import load_args_from_config
    targs, _ = load_args_from_config(os.path.join(REPO, "busca_b200", "configs", "bytetrack_mot20.yml"))
    a = targs.transformer
    a.device = f"cuda:{device}"
    a.precision = precision
    a.bank_slots = bank_slots
    m = BUSCA(a).eval()
    m.load_state_dict(synth.make_weights(0, profile=WEIGHTS_PROFILE))
    return m, targs


def adapter_leg(model, targs, args, device_rounds=False, config="mot20"):
    """The adapter's real call pattern (busca_b200/hosts/bytetrack.py after byte_tracker.py:226-456) on a MOT20-scale synthetic
    sequence: per frame 3 detection-crop calls, one single-box crop call per unmatched track, center_distance without a handle,
    associate_embeddings; tracker state evolves (histories appended, patch-bank slots recycled).  Times only what runs inside
    BUSCA's API (the host tracker's own Python is the caller's cost, reported beside it)."""
    import copy
    from busca_b200 import tracking
    from busca_b200.hosts.bytetrack import ByteTrackHost, DeviceRounds
    warm = 13
    seq = synth.make_sequence(4242, warm + args.adapter_frames, args.adapter_objects, miss=0.33, low_score=0.05, clutter=2.0, frame_ring=6)
    a = copy.copy(targs)
    if config == "mot17":                                   # BASELINE.json configs[1]: the MOT17 YAML (coverage gate, camera-motion compensation, score fusion)
        from busca_b200.option import load_args_from_config
        a, _ = load_args_from_config(os.path.join(REPO, "busca_b200", "configs", "bytetrack_mot17.yml"))
    a.use_busca, a.track_thresh, a.track_buffer, a.match_thresh, a.mot20 = True, 0.6, 30, 0.9, config != "mot17"
    clock = {"t": 0.0}

    def timed(fn):
        def w(*x, **k):
            t0 = time.perf_counter()
            try:
                return fn(*x, **k)
            finally:
                clock["t"] += time.perf_counter() - t0
        return w

    class Timed:                                             # BUSCA as the adapter sees it, every entry point on the clock
        get_image_crops = staticmethod(timed(model.get_image_crops))
        associate_embeddings = staticmethod(timed(model.associate_embeddings))

    host = ByteTrackHost(Timed, a, iou_fn=lambda x, y: model.engine.iou(x, y), center_distance_fn=timed(lambda t, d: tracking.center_distance(t, d)),
                         rounds=DeviceRounds(model.engine) if device_rounds else None,   # the rounds themselves on the device
                         reliable_fn=timed(lambda shape, tracks, p: tracking.is_reliable(shape, tracks, p, engine=model.engine)),
                         camera_motion_fn=timed(lambda prev, cur: model.engine.camera_motion(prev, cur)[0]))
    busca_ms, total_ms, n_unmatched, kept = [], [], [], 0
    pr = None
    if args.profile_e2e:
        import cProfile
        pr = cProfile.Profile()
    for f in range(len(seq.dets)):
        if pr is not None and f == warm:
            pr.enable()
        clock["t"] = 0.0
        t0 = time.perf_counter()
        host.update(seq.dets[f].copy(), [seq.H, seq.W], [seq.H, seq.W], current_frame=seq.frames[f])
        model.engine.sync()
        if f >= warm and host.last_busca is not None:
            total_ms.append((time.perf_counter() - t0) * 1e3)
            busca_ms.append(clock["t"] * 1e3)
            n_unmatched.append(len(host.last_busca[2]))
            kept += len(host.last_busca[0])
    if pr is not None:
        import pstats
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(45)
    if not busca_ms:
        return None
    dec = float(np.sum(n_unmatched))
    return {"frames": len(busca_ms), "objects": args.adapter_objects, "config": config, "bank_slots_in_use": int(model.engine.slots_in_use()) if hasattr(model.engine, "slots_in_use") else None,
            "total_ms_per_frame_p50": round(float(np.percentile(total_ms, 50)), 2), "total_ms_per_frame_p99": round(float(np.percentile(total_ms, 99)), 2),
            "unmatched_tracks_per_frame": round(float(np.mean(n_unmatched)), 1),
            "value": round(dec / (np.sum(busca_ms) * 1e-3), 1), "unit": "decisions/s (time inside BUSCA's API only)",
            "busca_ms_per_frame_p50": round(float(np.percentile(busca_ms, 50)), 2), "busca_ms_per_frame_p99": round(float(np.percentile(busca_ms, 99)), 2),
            "host_tracker_ms_per_frame_p50": round(float(np.percentile(np.array(total_ms) - np.array(busca_ms), 50)), 2),
            "kept_alive": int(kept), "decisions": int(dec),
            "pattern": "3 detection-crop calls + one single-box crop call per unmatched track + center_distance + associate_embeddings per frame"}


def run_ours(args):
    import gc
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = WORKLOADS[args.workload]
    T, D, L, C = wl["T"], wl["D"], wl["L"], wl["C"]
    # BASELINE.json configs[3]: a FIXED job of `--sequences` independent sequences (default 64), sharded BY SEQUENCE over the ranks
    # (no hot-path collective).  One step = frame k of every sequence, so the work per step is fixed and ms_per_step shrinks with
    # N (strong scaling); each rank runs the frames of the sequences it owns back to back on its GPU.
    S = args.sequences
    mine = sharding.partition_sequences([args.steps + args.warmup] * S, world, policy="round_robin")[rank]
    model, targs = build_model(args.precision, local, bank_slots=max(8192, len(mine) * (T * L + D + T) + 4096))
    eng = model.engine
    if args.adapter_only:
        if rank == 0:
            if os.environ.get("BUSCA_DEFER_CROP_COPIES") == "1":
                model.engine.set_option("defer_crop_copies", 1)
            out = adapter_leg(model, targs, args, device_rounds=True, config=args.adapter_config)
            print(json.dumps({"metric": "decisions/sec (adapter pattern, whole sequence)", "impl": "ours", "dtype": args.precision, "e2e_adapter": out}), flush=True)
        return
    frame_sets = {}

    def frames_for(sid):                                     # 4 distinct frame sets (synthesis costs ~1 s each); set 0 = the golden scene's
        k = sid % 4
        if k not in frame_sets:
            from busca_b200.scene import scene_frames
            frame_sets[k] = scene_frames(k, 3)
        return frame_sets[k]

    scenes = [Scene(T, D, L, C, seed=sid, frames=frames_for(sid)) for sid in mine]
    for sc in scenes:
        sc.setup_resident(model, busca_thresh=targs.busca_thresh, select_highest=targs.select_highest_candidate, own_frame=True)
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize(local)
        eng.sync()
        if dist is not None:
            dist.barrier()

    # ---- resident (`value`)
    for _ in range(args.warmup):
        for sc in scenes:
            sc.step_resident()
    eng.sync()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.launches
    run0, tot0 = eng.counter("reid_images_run"), eng.counter("reid_images_total")
    lat = []
    # timed region: exactly K steps, nothing but the product path on the stream (no per-kernel events: an event pair around each of
    # the ~840 launches of a frame costs ~2 ms/frame of device idle time and serialises launches that otherwise overlap their
    # prologue with the predecessor's tail)
    e0.record(stream)
    for _ in range(args.steps):
        for sc in scenes:
            t0 = time.perf_counter()
            sc.step_resident()
            eng.sync()                                        # a tracker needs frame k's decisions before it can submit frame k+1
            lat.append((time.perf_counter() - t0) * 1e3)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = eng.launches - launches0
    # per-kernel durations for the roofline / share tables: the SAME steps again with a CUDA-event pair around every launch (on the
    # engine's stream), directly after the timed region; ms_per_step_profiled is reported next to ms_per_step
    prof_acc = {}
    prof_steps = args.steps if len(scenes) <= 8 else 1
    eng.set_profiling(True)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(prof_steps):
        for sc in scenes:
            sc.step_resident()
            eng.sync()
            for k, v in eng.last_profile().items():
                a = prof_acc.setdefault(k, {"ms": 0.0, "launches": 0, "flops": 0.0, "xflops": 0.0, "kernel": v.get("kernel", "")})
                a["ms"] += v["ms"]
                a["launches"] += v["launches"]
                a["flops"] += v.get("flops", 0.0)
                a["xflops"] += v.get("xflops", 0.0)
    p1.record(stream)
    barrier()
    ms_prof = p0.elapsed_time(p1)
    n_frames_prof = max(1, prof_steps * len(scenes))
    n_frames_timed = max(1, args.steps * len(scenes))          # frames this rank ran inside the timed region
    eng.set_profiling(False)
    results = [sc.read_resident() for sc in scenes]
    scene = scenes[0] if scenes else None
    parity = None
    if scene is not None:
        keep, probs = results[0]["keep"], results[0]["probs"]
        assert np.isfinite(probs).all() and abs(float(probs.sum()) - T) < 1e-2 * T
        # rank 0's first sequence (seed 0) is the scene tests/golden/scene_<workload>_cond.npz holds the UNMODIFIED reference's answer
        # for: the timed step's probabilities / decisions are checked against it (same bounds as tests/test_gpu_scene.py)
        gpath = os.path.join(REPO, "tests", "golden", f"scene_{args.workload}_cond.npz")
        if rank == 0 and os.path.exists(gpath):
            g = np.load(gpath)
            if [int(v) for v in g["meta"]] == [scene.seed, T, D, L, C]:
                tolp, tie = (1e-3, 2e-3) if args.precision == "fp32" else (3.5e-2, 3.5e-2)
                kslot = min(D, C - 1)
                ref_keep = g["reliable"] & (g["probs"][:, kslot] > targs.busca_thresh)
                clear = np.abs(g["probs"][:, kslot] - targs.busca_thresh) > tie
                parity = {"golden": os.path.basename(gpath), "max_abs_dprob": round(float(np.abs(probs - g["probs"]).max()), 6), "tolerance": tolp,
                          "decisions_compared": int(clear.sum()), "decisions_equal": bool(np.array_equal(keep.astype(bool)[clear], ref_keep[clear])),
                          "reference_kept": int(ref_keep.sum())}
                assert parity["max_abs_dprob"] < tolp and parity["decisions_equal"], parity

    # ---- plug-in API (`e2e`): every owned sequence in turn (its host-side crops live only while it runs), K timed frames each
    e2e_s = float("nan")
    h2d = d2h = 0
    if not args.no_e2e:
        e2e_s = 0.0
        barrier()
        for sc in scenes:
            sc.setup_e2e(model)
            for i in range(max(1, min(args.warmup, 2))):
                sc.step_e2e(i, busca_thresh=targs.busca_thresh)
            eng.sync()
            t0 = time.perf_counter()
            for i in range(args.steps):
                sc.step_e2e(i, busca_thresh=targs.busca_thresh)
            eng.sync()
            e2e_s += time.perf_counter() - t0
            h2d, d2h = sc.h2d, sc.d2h
            if args.profile_e2e and rank == 0 and sc is scenes[0]:   # where the host side of the plug-in path spends its time (stderr)
                import cProfile
                import pstats
                pr = cProfile.Profile()
                pr.enable()
                for i in range(args.steps):
                    sc.step_e2e(i)
                eng.sync()
                pr.disable()
                pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(30)
            sc.tracks = sc._hist_crops = None
            gc.collect()
        barrier()
    adapter = adapter_leg(model, targs, args) if (rank == 0 and args.adapter_frames > 0) else None
    if adapter is not None:
        dev = adapter_leg(model, targs, args, device_rounds=True)
        if dev is not None:
            adapter["with_device_rounds"] = {k: dev[k] for k in ("value", "busca_ms_per_frame_p50", "host_tracker_ms_per_frame_p50", "kept_alive", "decisions")}
            adapter["with_device_rounds"]["what"] = "same sequence with the host tracker's rounds on libbusca_b200 (batched Kalman predict / update, IoU cost + assignment per round, duplicate removal: csrc/rounds.cu)"
        model.engine.set_option("defer_crop_copies", 1)
        try:
            dev = adapter_leg(model, targs, args, device_rounds=True)
        finally:
            model.engine.set_option("defer_crop_copies", 0)
        if dev is not None:
            adapter["with_device_rounds_and_deferred_crop_copies"] = {k: dev[k] for k in ("value", "busca_ms_per_frame_p50", "host_tracker_ms_per_frame_p50", "total_ms_per_frame_p50", "kept_alive", "decisions")}
            adapter["with_device_rounds_and_deferred_crop_copies"]["what"] = "plus the opt-in defer_crop_copies: get_image_crops returns before its device->host copy has landed (valid after the next waiting call)"

    # NCCL only here: max over ranks of the device-timed regions, and the gather of the per-rank result tables
    ms_max, e2e_ms_max = sharding.reduce_max([ms, e2e_s * 1e3], dist, device=f"cuda:{local}")
    # result rows of the last frame of every owned sequence: (sequence, frame, track, kept-alive box = the motion proposal, p_kalman)
    rows = []
    for sid, sc, r in zip(mine, scenes, results):
        b = sc.pred_tlwh
        rows += [(sid, args.steps, t, b[t, 0], b[t, 1], b[t, 2], b[t, 3], float(r["probs"][t, min(D, C - 1)])) for t in range(T) if r["keep"][t]]
    table = sharding.gather_results(sharding.pack_results(rows), dist, device=f"cuda:{local}")
    frames_max, = sharding.reduce_max([float(n_frames_timed)], dist, device=f"cuda:{local}")
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md 1.4 PF sustained)"
    # per-layer table -> (a) classes for the share table, (b) the __global__ instantiations ncu names, for the roofline
    by_class, by_kernel = {}, {}
    for k, v in prof_acc.items():                          # "conv1x1_tc[64>256 s1 96x32]" -> class "conv1x1"
        cls = k.split("_tc[")[0]
        a = by_class.setdefault(cls, {"ms": 0.0, "launches": 0})
        a["ms"] += v["ms"]
        a["launches"] += v["launches"]
        if v.get("kernel") and v.get("xflops", 0.0) > 0:
            b = by_kernel.setdefault(v["kernel"], {"ms": 0.0, "launches": 0, "flops": 0.0, "xflops": 0.0})
            b["ms"] += v["ms"]
            b["launches"] += v["launches"]
            b["flops"] += v["flops"]                        # algorithmic: statistics-only passes carry 0
            b["xflops"] += v["xflops"]                      # executed
    detail = {k: round(v["ms"] / n_frames_prof, 4) for k, v in prof_acc.items() if "_tc[" in k}
    total_prof = sum(v["ms"] for v in by_class.values())
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    step_flops = 8.0096e9 * T * (L + C)                     # 8.0096 GFLOP per patch x the stacked patches of one step
    roofline = None
    if by_kernel:
        # Dominant kernel = the __global__ with the most device time.  conv_tc_kernel's template arguments after the first four (BN, KB,
        # DUAL, RESB) only select tuning variants of the SAME kernel for a launch (CTA-pair MMA, ring depth, statistics-only ring), so
        # those instantiations are one family here - otherwise every new variant would split the dominant kernel's time and hand the
        # title to an HBM-bound one.  Statistics-only launches belong to their family and credit 0 FLOPs.
        import re

        def family(k):
            m = re.match(r"(conv_tc_kernel<\d+, \d+, \d+, \d+)(, .*)?>", k)
            return m.group(1) + ", ...>" if m else k

        fam = {}
        for k, v in by_kernel.items():
            f = fam.setdefault(family(k), {"ms": 0.0, "launches": 0, "flops": 0.0, "xflops": 0.0, "members": []})
            for q in ("ms", "launches", "flops", "xflops"):
                f[q] += v[q]
            f["members"].append(k)
        dom = max(fam, key=lambda k: fam[k]["ms"])
        d = fam[dom]
        avg_ms = d["ms"] / d["launches"]
        # ALGORITHMIC FLOPs of the family's launches (every convolution counted once; its statistics-only recomputation
        # passes add time but no FLOPs) / the CUDA-event time of ALL its launches
        achieved = d["flops"] / (d["ms"] * 1e-3) / 1e12
        traffic = step_dram = None
        try:                                                     # DRAM bytes per launch of the same kernel, from the committed ncu capture
            tr = json.load(open(os.path.join(REPO, "profiles", "ncu_traffic.json")))
            mem = [tr["kernels"][k] for k in d["members"] if k in tr["kernels"]]
            traffic = sum(v["dram_bytes_per_launch"] * v["launches"] for v in mem) / max(1, sum(v["launches"] for v in mem)) if mem else None
            step_dram = sum(v["dram_bytes_per_launch"] * v["launches"] for v in tr["kernels"].values() if v.get("dram_bytes_per_launch"))
        except Exception:
            pass
        ms_step = ms / n_frames_timed                            # per FRAME (one sequence): what the per-frame FLOP / byte counts refer to
        # The chain is layer-serialised (a grid-wide BatchNorm dependency after every convolution), so the bound of a frame is the SUM over
        # its launches of max(tensor time, HBM time), not one roofline: per kernel instantiation, algorithmic FLOPs / tensor peak against ncu
        # DRAM bytes / copy peak (an instantiation mixes tensor- and HBM-bound layers, which only lowers this bound: conservative).
        composite = None
        try:
            cb = 0.0
            for k, v in by_kernel.items():
                t_tensor = v["flops"] / n_frames_prof / (peak_tf * 1e12)
                t_hbm = tr["kernels"][k]["dram_bytes_per_launch"] * (v["launches"] / n_frames_prof) / (peak_hbm * 1e9) if k in tr["kernels"] else 0.0
                cb += max(t_tensor, t_hbm)
            hbm_only = sum(v["dram_bytes_per_launch"] * v["launches"] for kk, v in tr["kernels"].items()
                           if kk not in by_kernel and v.get("dram_bytes_per_launch")) / (peak_hbm * 1e9)   # pooling, BN folds, gather ...
            composite = {"bound_ms_per_frame": round((cb + hbm_only) * 1e3, 3), "frac": round((cb + hbm_only) / (ms_step * 1e-3), 5),
                         "what": "sum over kernel instantiations of max(algorithmic FLOPs / sustained bf16 peak, ncu DRAM bytes / copy peak), "
                                 "plus DRAM time of the non-GEMM kernels, over ms_per_frame"}
        except Exception:
            pass
        roofline = {"bound": "tensor", "kernel": dom, "achieved": round(achieved, 3), "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": round(achieved / peak_tf, 5), "traffic": traffic, "peak_source": peak_src,
                    "avg_launch_ms": round(avg_ms, 4), "launches": d["launches"], "instantiations": sorted(d["members"]),
                    "flops_per_launch": round(d["flops"] / d["launches"], 1),
                    "executed_frac": round(d["xflops"] / (d["ms"] * 1e-3) / 1e12 / peak_tf, 5),
                    "share_of_step": round(d["ms"] / max(1e-9, total_prof), 4),
                    "step_frac": round(step_flops / (ms_step * 1e-3) / 1e12 / peak_tf, 5),
                    "step_tflops": round(step_flops / (ms_step * 1e-3) / 1e12, 1),
                    "step_hbm_frac": round(step_dram / (ms_step * 1e-3) / 1e9 / peak_hbm, 5) if step_dram else None,
                    "step_dram_bytes_ncu": step_dram, "composite": composite,
                    "all_conv_kernels": {k: {"tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1), "executed_tflops": round(v["xflops"] / (v["ms"] * 1e-3) / 1e12, 1),
                                             "ms_per_frame": round(v["ms"] / n_frames_prof, 3), "launches_per_frame": v["launches"] / n_frames_prof}
                                         for k, v in by_kernel.items()},
                    "timing": f"CUDA-event pair around every launch on the engine's stream, {prof_steps} step(s) run again directly after the timed region "
                              f"({round(ms_prof / n_frames_prof, 3)} ms/frame with the events in the stream)",
                    "note": "achieved/frac: algorithmic FLOPs (statistics-only recomputation passes credit 0) over the time of all launches of the dominant "
                            "instantiation; step_frac: 8.0096 GFLOP x stacked patches / ms_per_step / measured sustained bf16 peak"}
    prof_acc = by_class
    kernels = {k: {"ms_per_frame": round(v["ms"] / n_frames_prof, 4), "launches_per_frame": v["launches"] / n_frames_prof,
                   "share": round(v["ms"] / max(total_prof, 1e-9), 4)} for k, v in sorted(prof_acc.items(), key=lambda kv: -kv[1]["ms"])}
    cpu = cpu_baseline(args) if not args.no_cpu_baseline else None
    host_rows = host_rows_leg(eng) if (rank == 0 and not args.no_cpu_baseline) else None
    out = {
        "metric": "decisions/sec", "value": round(S * T * args.steps / (ms_max * 1e-3), 3), "unit": "decisions/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_max / args.steps, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": wl["desc"] + f"; fixed job of {S} independent sequences sharded by sequence (BASELINE.json configs[3]), one step = one "
                               f"frame of every sequence", "name": args.workload, "T": T, "D": D, "L": L, "C": C, "patches_per_frame": T * (L + C),
                   "sequences": S, "frames_per_step": S, "frames_per_step_on_slowest_rank": int(frames_max / args.steps),
                   "weights": "random-init, conditioned profile (numpy PCG64 seed 0; synth.make_weights), model_busca.pth layout",
                   "l2": "inputs larger than L2: every frame streams GBs of ReID activations through HBM (L2 is 126 MB)",
                   "parallelism": f"{S} sequences over {world} GPU(s) by sequence (round robin), no hot-path collective; NCCL gathers the result rows"},
        "ms_per_frame": round(ms_max / frames_max, 4),
        "ms_per_frame_profiled": round(ms_prof / n_frames_prof, 4),
        "p50_frame_latency_ms": round(float(np.percentile(lat, 50)), 3), "p99_frame_latency_ms": round(float(np.percentile(lat, 99)), 3),
        "e2e": {"value": round(S * T * args.steps / (e2e_ms_max * 1e-3), 3) if e2e_ms_max == e2e_ms_max else None, "unit": "decisions/s",
                "h2d_bytes_per_step": int(h2d) * S, "d2h_bytes_per_step": int(d2h) * S, "h2d_bytes_per_frame": int(h2d), "d2h_bytes_per_frame": int(d2h),
                "ms_per_frame": round(e2e_ms_max / frames_max, 4) if e2e_ms_max == e2e_ms_max else None,
                "api": "BUSCA.get_image_crops + center_distance + associate_embeddings (host numpy in/out), every sequence for K frames"},
        "e2e_adapter": adapter, "host_rows": host_rows,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kernels,
        "conv_detail_ms_per_frame": detail,
        "cpu_baseline": cpu,
        "kept_tracks": int(table.shape[0]),
        "result_rows_gathered": int(table.shape[0]),
        "parity_vs_reference_golden": parity,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------
def cpu_sample_case(T, D, L, seed=5):
    from oracle import crop as ocrop
    return synth.make_assoc_case(seed, T, D, L, crop_fn=ocrop.get_image_crops)


def time_oracle(T, D, L, C, reps, warm):
    """The reference's algorithm (CPU oracle port: crops, geometry, ReID, Transformer) on the host cores."""
    import torch
    from oracle import geometry as ogeo
    from oracle import network as onet
    torch.set_num_threads(os.cpu_count())
    weights = synth.make_weights(0, profile=WEIGHTS_PROFILE)
    case = cpu_sample_case(T, D, L)
    times = []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        dists = ogeo.center_distance([t.tlbr for t in case.tracks], [d.tlbr for d in case.dets])
        pm, rel = onet.associate(weights, case.tracks, case.dets, dists, L, C, True, kalman=case.kalman)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return times


def host_rows_leg(eng):
    """SURVEY.md 8(f) rows at MOT20 scale, device call (host buffers in and out, wall clock) beside the CPU path of the same row (part of
    the cpu_baseline leg: oracle/ restatements, and cv2 itself for the ECC row when the box has it).  Milliseconds, best of 3."""
    from oracle import coverage as ocov
    from oracle import rounds as ornd
    rng = np.random.default_rng(0)

    def best(fn, reps=3):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return round(min(ts) * 1e3, 3)

    def boxes(n):
        b = synth.random_boxes(rng, n)
        b[:, 2:] += b[:, :2]
        return b

    a = boxes(500)
    b = np.concatenate([a[:350] + rng.normal(0, 4, (350, 4)), boxes(50)])
    sc = rng.uniform(0.1, 1.0, len(b))
    out = {"match_round_500x400": {"device_ms": best(lambda: eng.match_round(a, b, sc, 0.9)),
                                   "cpu_ms": best(lambda: ornd.linear_assignment(ornd.fuse_score(ornd.iou_distance(a, b), sc), 0.9), 1),
                                   "cpu": "numpy IoU + scipy linear_sum_assignment on lapjv's extended matrix"}}
    n = 500
    mean = np.concatenate([rng.uniform(0, 1900, (n, 2)), rng.uniform(0.2, 0.8, (n, 1)), rng.uniform(60, 300, (n, 1)), rng.normal(0, 3, (n, 4))], axis=1)
    cov = np.stack([np.diag(rng.uniform(0.5, 4.0, 8)) for _ in range(n)])
    z = mean[:, :4] + 1.0
    out["kalman_predict_update_500"] = {"device_ms": best(lambda: eng.kalman_update(*eng.kalman_predict(mean, cov, None), z)),
                                        "cpu_ms": best(lambda: [ornd.kf_update(m, c, zz) for m, c, zz in zip(*ornd.kf_multi_predict(mean, cov, None), z)], 1),
                                        "cpu": "numpy multi_predict + one scipy cho_factor / cho_solve update per track, as the reference"}
    cb = boxes(300)
    out["detection_coverage_1080p_300"] = {"device_ms": best(lambda: eng.detection_coverage(cb, 1080, 1920)),
                                           "cpu_ms": best(lambda: ocov.detection_coverage((1080, 1920), cb), 1), "cpu": "numpy canvas fill + count_nonzero"}
    f1 = synth.make_frame(21)
    f2 = synth.make_moved_frame(f1, 0.003, 2.2, -1.3, 21)
    row = {"device_ms": best(lambda: eng.camera_motion(f1, f2))}
    try:
        import cv2
        g1, g2 = cv2.cvtColor(f1, cv2.COLOR_BGR2GRAY), cv2.cvtColor(f2, cv2.COLOR_BGR2GRAY)
        crit = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 100, 1e-5)
        row["cpu_ms"] = best(lambda: cv2.findTransformECC(g1, g2, np.eye(2, 3, dtype=np.float32), cv2.MOTION_EUCLIDEAN, crit), 1)
        row["cpu"] = f"cv2 {cv2.__version__} cvtColor + findTransformECC (the reference's own call), {cv2.getNumThreads()} threads"
    except Exception:
        from oracle import ecc as oecc
        row["cpu_ms"] = best(lambda: oecc.camera_motion(f1, f2), 1)
        row["cpu"] = "numpy restatement of cv2.findTransformECC (cv2 not importable on this box)"
    out["camera_motion_1080p"] = row
    chw = synth.make_detector_tensor(1, 1080, 1920)
    from oracle import ingest as oing
    out["frame_ingest_1080p"] = {"device_ms": best(lambda: eng.ingest_frame(chw, synth.YOLOX_MEANS, synth.YOLOX_STD)),
                                 "cpu_ms": best(lambda: oing.denormalize_frame(chw, synth.YOLOX_MEANS, synth.YOLOX_STD), 1),
                                 "cpu": "numpy de-normalisation (the evaluator's statements)", "note": "device_ms includes the 25 MB host->device and 6 MB device->host copies"}
    return out


def cpu_baseline(args):
    wl = WORKLOADS[args.workload]
    Ts = args.cpu_tracks
    times = time_oracle(Ts, min(wl["D"], 4 * Ts), wl["L"], wl["C"], reps=1, warm=0)
    return {"value": round(Ts / float(np.median(times)), 4), "unit": "decisions/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{Ts} unmatched tracks x {wl['C']} proposals of the same workload ({Ts * (wl['L'] + wl['C'])} ReID patches), "
                      f"1 repetition, torch CPU fp32 with {os.cpu_count()} threads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    Ts = args.cpu_tracks
    times = time_oracle(Ts, min(wl["D"], 4 * Ts), wl["L"], wl["C"], reps=args.steps, warm=min(args.warmup, 1))
    tot = float(np.sum(times))
    val = round(Ts * len(times) / tot, 4)
    sample = (f"each step = {Ts} unmatched tracks x {wl['C']} proposals of the same workload ({Ts * (wl['L'] + wl['C'])} ReID patches); "
              f"oracle port of the reference algorithm, torch CPU fp32, {os.cpu_count()} threads")
    out = {"impl": "reference", "metric": "decisions/sec", "value": val, "unit": "decisions/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(tot / len(times) * 1e3, 3), "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": wl["desc"], "name": args.workload, "T": wl["T"], "D": wl["D"], "L": wl["L"], "C": wl["C"], "sequences": args.sequences},
           "cpu_baseline": {"value": val, "unit": "decisions/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "decisions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mot20", choices=list(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("BUSCA_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--cpu-tracks", type=int, default=16, help="size of the bounded CPU sample (unmatched tracks)")
    ap.add_argument("--sequences", type=int, default=64, help="independent sequences of the fixed job (BASELINE.json configs[3]: 64)")
    ap.add_argument("--adapter-frames", type=int, default=30, help="frames of the adapter-pattern leg after 13 warm-up frames (0 = skip)")
    ap.add_argument("--adapter-objects", type=int, default=560, help="objects of the adapter-pattern sequence (~200 unmatched tracks per frame at 560)")
    ap.add_argument("--adapter-config", default="mot20", choices=["mot20", "mot17"], help="YAML of the adapter-pattern leg (mot17: coverage gate + camera-motion compensation + score fusion)")
    ap.add_argument("--adapter-only", action="store_true", help="run only the adapter-pattern leg (whole sequences: BASELINE.json configs[1] / [2]) with the rounds on the device and print its JSON")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the plug-in API leg (profiler runs)")
    ap.add_argument("--profile-e2e", action="store_true", help="after the timed plug-in leg, cProfile the same loop to stderr")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
