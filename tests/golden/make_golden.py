#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported read-only with the shims next to this file) on CPU.

Run in the build container only:   python tests/golden/make_golden.py [crops geometry pe assoc]      (default set)
                                   python tests/golden/make_golden.py cond scene scene_mot20 adapter adapter_mot17   (conditioned weights)
                                   python tests/golden/make_golden.py coverage rounds                  
                                   python tests/golden/make_golden.py ingest                           (8f row 4)
                                   python tests/golden/make_golden.py ecc                              (8f row 3: cv2 itself)
The GPU box never runs this (it has no /root/reference); it only reads the .npz files.

Inputs are not stored: they are regenerated from seeds by busca_b200.synth (numpy PCG64), so
the fixtures stay small.  What is stored are the reference's outputs at the probe points listed
in SURVEY.md Appendix B.
"""
import ast
import hashlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [os.path.join(HERE, "shims"), REF, REPO]

import numpy as np

np.float = float  # alias removed in numpy>=1.24; the reference still uses it (tracking.py:43)
import torch

torch.set_num_threads(os.cpu_count())

import busca.reid.resnet as _R

_R.load_state_dict_from_url = lambda *a, **k: {}  # no network (resnet.py:346-360)
import busca.encodings as ref_enc
import busca.network as ref_net
import busca.tracking as ref_trk
from busca.option import load_args_from_config

from busca_b200 import synth


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build_reference(weights_seed=0, profile="chaotic"):
    targs, _ = load_args_from_config(os.path.join(REF, "config/ByteTrack/MOT20/config_bytetrack_mot20.yml"))
    a = targs.transformer
    a.device = torch.device("cpu")
    a.reid_weights_file = "no"
    model = ref_net.BUSCA(a).eval()
    sd = synth.make_weights(weights_seed, profile=profile)
    ref_keys = {k for k in model.state_dict().keys() if ".fc." not in k}
    assert ref_keys == set(sd.keys()), (sorted(ref_keys - set(sd))[:5], sorted(set(sd) - ref_keys)[:5])
    for k, v in model.state_dict().items():
        if k in sd:
            assert tuple(v.shape) == tuple(sd[k].shape), (k, v.shape, sd[k].shape)
    path = "/tmp/busca_golden_weights_%s.pth" % profile
    torch.save({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, path)
    model.load_pretrained(path, ignore_reid_fc=True)
    return model, targs


# --------------------------------------------------------------------------------------
def golden_crops(model):
    """a6: get_image_crops / get_bbox_crop (network.py:492-507, tracking.py:62-113)."""
    frame = synth.make_frame(4242)
    H, W = frame.shape[:2]
    rng = np.random.default_rng(7)
    boxes = []
    # inside
    for _ in range(20):
        w, h = rng.uniform(8, 400), rng.uniform(8, 900)
        x, y = rng.uniform(0, W - w), rng.uniform(0, H - h)
        boxes.append([x, y, x + w, y + h])
    # straddling each border
    for _ in range(16):
        w, h = rng.uniform(20, 300), rng.uniform(40, 600)
        cx = rng.choice([0.0, W, rng.uniform(0, W)])
        cy = rng.choice([0.0, H, rng.uniform(0, H)])
        boxes.append([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2])
    # fully outside, degenerate, integer-aligned, exact 2x (INTER_AREA switch), tiny, huge
    boxes += [[-300.0, -300.0, -100.0, -50.0], [W + 5.0, 10.0, W + 80.0, 200.0], [100.0, H + 1.0, 180.0, H + 300.0],
              [500.0, 500.0, 500.0, 700.0], [500.0, 500.0, 600.0, 500.0],
              [100.0, 100.0, 356.0, 868.0], [101.0, 57.0, 357.0, 825.0],
              [64.0, 64.0, 192.0, 448.0], [10.0, 10.0, 11.0, 11.0], [10.2, 10.7, 12.1, 13.9],
              [-50.0, -50.0, W + 50.0, H + 50.0], [0.0, 0.0, float(W), float(H)],
              [1900.5, 1000.25, 1930.75, 1090.5], [-10.5, 500.0, 30.25, 620.0]]
    boxes = np.asarray(boxes, dtype=np.float64)
    crops64 = model.get_image_crops(frame, boxes, normalize=False)
    crops32 = model.get_image_crops(frame, boxes.astype(np.float32), normalize=False)
    assert crops64.dtype == np.uint8 and crops64.shape[1:] == (384, 128, 3)
    empty = model.get_image_crops(frame, [], normalize=False)
    np.savez_compressed(os.path.join(HERE, "crops.npz"), frame_seed=4242, boxes=boxes,
                        sha64=np.array([sha(c) for c in crops64]), sha32=np.array([sha(c) for c in crops32]),
                        full_idx=np.array([0, 20, 24, 36, 39, 41, 42, 46, 49]),
                        full=crops64[[0, 20, 24, 36, 39, 41, 42, 46, 49]],
                        empty_shape=np.array(empty.shape), empty_dtype=str(empty.dtype))
    print("crops:", len(boxes), "boxes; fp32/fp64 digests differ on", int(sum(sha(a) != sha(b) for a, b in zip(crops64, crops32))))


def _extract_function(path, name):
    src = open(path).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"np": np, "torch": torch}
            exec(compile(mod, path, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def golden_geometry():
    """a1-a3: Kalman multi_predict mean, centre distance, IoU (+1 convention)."""
    rng = np.random.default_rng(11)
    a = synth.random_boxes(rng, 37)
    b = synth.random_boxes(rng, 53)
    b[:10] = a[:10] + rng.normal(0, 5, (10, 4))
    a[:, 2:] += a[:, :2]
    b[:, 2:] += b[:, :2]
    b[11] = a[11]                                     # identical boxes -> IoU 1, distance 0
    cd = ref_trk.center_distance(a, b)          # 2-D ndarrays (a list of arrays crashes the reference, tracking.py:45)
    cdw = ref_trk.center_distance(a, b, weight_size=True)
    # IoU: cython_bbox is not in the mount; the reference's own in-tree restatement
    # (adapters/GHOST/src/tracking_utils.py:176-205) is executed instead.
    bbox_overlaps = _extract_function(os.path.join(REF, "adapters/GHOST/src/tracking_utils.py"), "bbox_overlaps")
    iou_ghost = bbox_overlaps(a.copy(), b.copy())
    # Kalman (adapters/CenterTrack/src/lib/utils/mot_online/kalman_filter.py:154-191)
    sys.path.insert(0, os.path.join(REF, "adapters/CenterTrack/src/lib/utils/mot_online"))
    import kalman_filter as KF
    kf = KF.KalmanFilter()
    mean = np.concatenate([rng.uniform(0, 1900, (29, 2)), rng.uniform(0.2, 0.8, (29, 1)), rng.uniform(60, 300, (29, 1)),
                           rng.normal(0, 3, (29, 4))], axis=1)
    cov = np.stack([np.eye(8) * rng.uniform(0.5, 4.0) for _ in range(29)])
    tracked = rng.uniform(size=29) < 0.7
    m_in = mean.copy()
    m_in[~tracked, 7] = 0                             # STrack.multi_predict, byte_tracker.py:54-56
    m_out, _ = kf.multi_predict(m_in, cov)
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), a=a, b=b, center_distance=cd, center_distance_weighted=cdw,
                        iou_ghost=iou_ghost, kf_mean_in=mean, kf_tracked=tracked, kf_mean_out=m_out)
    print("geometry: cd", cd.shape, "iou", iou_ghost.shape)


class Probe:
    """Wrap bound methods of the reference model to capture intermediates without editing it."""

    def __init__(self, model):
        self.m = model
        self.rec = {}
        self._orig = []

    def wrap(self, obj, name, key, multi=False):
        orig = getattr(obj, name)
        rec = self.rec

        def wrapped(*a, **k):
            out = orig(*a, **k)
            if multi:
                rec.setdefault(key, []).append(out)
            else:
                rec[key] = out
            return out

        setattr(obj, name, wrapped)
        self._orig.append((obj, name, orig))

    def restore(self):
        for obj, name, orig in self._orig:
            try:
                delattr(obj, name)
            except AttributeError:
                setattr(obj, name, orig)


def run_assoc(model, case, L, C, flavour64, contiguous=True, **kw):
    """Run the reference's associate_embeddings; with ``flavour64`` the sentinel box is float64 as
    under the pinned numpy 1.23.5.

    ``contiguous``: the reference hands the ReID encoder a permuted (channels-last strided) view
    (network.py:397-398).  On CPU, torch 2.11 then picks its channels-last batch-norm training kernel,
    whose statistics are ~80x less accurate than the NCHW kernel (5e-5 vs 6e-7 per layer against
    fp64; compounding to ~1e-2 on the embeddings over 53 layers).  With contiguous=True the very same
    call is fed ``x.contiguous()`` so the accurate kernels run; the algorithm is unchanged.  The
    primary fixtures use this; the as-is numbers are stored beside them (asis_*)."""
    orig_mcb = ref_trk.missing_candidate_bbox
    saved_fake = model.pos_encoder.distant_fake_bbox
    if flavour64:
        f64 = lambda seq_len=None, flavour="ltrb": orig_mcb(seq_len, flavour).astype(np.float64)
        ref_enc.missing_candidate_bbox = f64
        ref_net.missing_candidate_bbox = f64
        model.pos_encoder.distant_fake_bbox = torch.from_numpy(f64(flavour="ltwh"))
    else:
        assert orig_mcb(flavour="ltwh").dtype == np.float32, "expected numpy>=2 behaviour in this container"
    pr = Probe(model)
    if contiguous:
        enc_fwd = model.reid_encoder.forward
        model.reid_encoder.forward = lambda x: enc_fwd(x.contiguous())
        pr._orig.append((model.reid_encoder, "forward", enc_fwd))
    pr.wrap(model.reid_encoder, "forward", "reid", multi=True)
    pr.wrap(model.pos_encoder, "_get_temporal_ids", "tids")
    pr.wrap(model.pos_encoder, "_get_spatial_ids", "sids")
    pr.wrap(model.pos_encoder, "forward", "input_seq")
    pr.wrap(model.transformer_encoder, "forward", "trans_out")
    pr.wrap(model, "forward", "logits")
    try:
        dists = ref_trk.center_distance(case.tracks, case.dets) if len(case.dets) else np.zeros((len(case.tracks), 0))
        with torch.no_grad():
            probs_matrix, reliable = model.associate_embeddings(
                tracks_embeddings=case.tracks, dets_embeddings=case.dets, dists_matrix=dists, seq_len=L, num_candidates=C,
                use_broader_memory=True, extra_kalman_candidates=case.kalman, normalize_ims=True, **kw)
    finally:
        pr.restore()
        ref_enc.missing_candidate_bbox = orig_mcb
        ref_net.missing_candidate_bbox = orig_mcb
        model.pos_encoder.distant_fake_bbox = saved_fake
    r = pr.rec
    T = len(case.tracks)
    out = dict(dists=dists, probs_matrix=probs_matrix, reliable=reliable,
               mem_emb=r["reid"][0][1].view(T, L, -1).numpy(), can_emb=r["reid"][1][1].view(T, C, -1).numpy(),
               mem_t=r["tids"][0].numpy(), can_t=r["tids"][1].numpy(),
               mem_xy=r["sids"][0][0].numpy(), mem_size=r["sids"][0][1].numpy(),
               can_xy=r["sids"][1][0].numpy(), can_size=r["sids"][1][1].numpy(),
               input_seq=r["input_seq"].numpy(), trans_out=r["trans_out"].numpy(),
               cand_rows=model.logits.numpy(), mem_logits=model.mem_logits.numpy(), logits=r["logits"].numpy())
    out["probs"] = torch.softmax(r["logits"], dim=-1).numpy()
    return out


def fp64_deviation(model, case, L, C):
    """max|emb - emb_fp64| / max|emb_fp64| of the memory batch for the NCHW and the as-is
    (channels-last) CPU paths: the evidence behind run_assoc's ``contiguous`` switch."""
    grab = []
    enc_fwd = model.reid_encoder.forward

    def hook(x):
        grab.append(x.clone())
        return enc_fwd(x)

    model.reid_encoder.forward = hook
    try:
        dists = ref_trk.center_distance(case.tracks, case.dets)
        with torch.no_grad():
            model.associate_embeddings(case.tracks, case.dets, dists, L, C, True, False,
                                       extra_kalman_candidates=case.kalman, normalize_ims=True)
    finally:
        del model.reid_encoder.forward
    x = grab[0]                                   # the memory batch, strided exactly as the reference passes it
    net = model.reid_encoder.model
    with torch.no_grad():
        _, e_cl = net(x, output_option="plain")
        _, e_nchw = net(x.contiguous(), output_option="plain")
        net.double()
        _, e64 = net(x.contiguous().double(), output_option="plain")
        net.float()
    d = lambda a: float((a.double() - e64).abs().max() / e64.abs().max())
    print("fp64 deviation: nchw", d(e_nchw), "channels-last as-is", d(e_cl))
    return d(e_nchw), d(e_cl)


def golden_assoc(model):
    ref_crop = lambda frame, boxes: model.get_image_crops(frame, boxes, normalize=False)
    cases = {
        # name: (seed, T, D, L, C, short_history)
        "assoc_cfg1": (101, 16, 40, 11, 5, 1),       # BASELINE config 1: 16 tracks x 5 proposals
        "assoc_fewdets": (102, 5, 3, 11, 5, 2),      # D < C: sentinel candidates, short histories
        "assoc_nodets": (103, 3, 0, 11, 5, 0),       # Kalman proposal only
    }
    for name, (seed, T, D, L, C, short) in cases.items():
        case = synth.make_assoc_case(seed, T, D, L, crop_fn=ref_crop, short_history=short)
        store = {"meta": np.array([seed, T, D, L, C, short])}
        for fl, tag in ((False, "f32"), (True, "f64")):
            out = run_assoc(model, case, L, C, flavour64=fl, select_highest_candidate=False)
            keep = ["probs_matrix", "reliable", "mem_t", "can_t", "mem_xy", "mem_size", "can_xy", "can_size",
                    "cand_rows", "logits", "probs", "mem_logits"]
            if tag == "f32":
                keep += ["dists", "mem_emb", "can_emb"]
                if T <= 5:
                    keep += ["input_seq", "trans_out"]
            for k in keep:
                store[f"{tag}_{k}"] = out[k]
            print(name, tag, "probs[0]", np.round(out["probs"][0], 4), "reliable", out["reliable"].astype(int))
        # the reference exactly as it runs on CPU here (channels-last BN kernels), for the record
        o = run_assoc(model, case, L, C, flavour64=True, contiguous=False, select_highest_candidate=False)
        for k in ("mem_emb", "can_emb", "probs", "logits"):
            store[f"asis_{k}"] = o[k]
        if name == "assoc_fewdets":
            store["fp64_dev"] = np.array(fp64_deviation(model, case, L, C))
        # host post-processing variants (network.py:415-424); device work is identical
        o = run_assoc(model, case, L, C, flavour64=True, select_highest_candidate=True)
        store["f64_probs_matrix_highest"] = o["probs_matrix"]
        o = run_assoc(model, case, L, C, flavour64=True, select_highest_candidate=True,
                      highest_candidate_minimum_thresh=0.3, keep_highest_value=True)
        store["f64_probs_matrix_highest_keep_thr"] = o["probs_matrix"]
        # a sample of the crops the case is built from, to pin the synthetic generator itself
        store["crop_sha_track0"] = np.array([sha(c) for c in case.tracks[0].images_mem])
        store["crop_sha_kalman"] = np.array([sha(k.images_mem[-1]) for k in case.kalman])
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **store)


def golden_pe(model):
    """a10: closed-form samples of the 211x211x61x512 fp16 table (encodings.py:23-32) plus the three
    separable 1-D tables it is made of, as the reference built them in this container."""
    pe = model.pos_encoder.pe
    assert pe.shape == (211, 211, 61, 512) and pe.dtype == torch.float16
    rng = np.random.default_rng(3)
    idx = np.stack([rng.integers(0, 211, 64), rng.integers(0, 211, 64), rng.integers(0, 61, 64)], axis=1)
    vals = torch.stack([pe[i, j, k] for i, j, k in idx]).numpy()
    np.savez_compressed(os.path.join(HERE, "pe.npz"), idx=idx, vals=vals,
                        tab_x=pe[:, 0, 0, :172].numpy(), tab_y=pe[0, :, 0, 172:344].numpy(), tab_z=pe[0, 0, :, 344:].numpy())
    print("pe samples", vals.shape, vals[0, :4])


def golden_assoc_cond(model):
    """The plug-in call on the CONDITIONED weight set (synth.make_weights(profile="conditioned")): winners vary from track to
    track and the Kalman-slot probabilities straddle busca_thresh, so decisions - and the bf16 path - are really tested."""
    ref_crop = lambda frame, boxes: model.get_image_crops(frame, boxes, normalize=False)
    cases = {"assoc_cfg1_cond": (101, 16, 40, 11, 5, 1), "assoc_fewdets_cond": (102, 5, 3, 11, 5, 2)}
    for name, (seed, T, D, L, C, short) in cases.items():
        case = synth.make_assoc_case(seed, T, D, L, crop_fn=ref_crop, short_history=short)
        out = run_assoc(model, case, L, C, flavour64=True, select_highest_candidate=False)
        store = {"meta": np.array([seed, T, D, L, C, short])}
        for k in ("probs_matrix", "reliable", "mem_emb", "can_emb", "cand_rows", "logits", "probs", "mem_logits", "dists",
                  "mem_xy", "mem_size", "can_xy", "can_size", "mem_t", "can_t"):
            store["f64_" + k] = out[k]
        o = run_assoc(model, case, L, C, flavour64=True, select_highest_candidate=True)
        store["f64_probs_matrix_highest"] = o["probs_matrix"]
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **store)
        p = out["probs"]
        print(name, "winners", np.bincount(p.argmax(1), minlength=C + 2), "p_kalman", np.round(p[:, min(D, C - 1)], 3))


def golden_scene(model, name, T, D, seed, emb_stride):
    """One frame of busca_b200.scene.Scene (what bench.py times and busca_frame_step_dev consumes) through the reference's
    associate_embeddings: the parity pin of the BENCHMARKED entry point at the benchmarked scale."""
    from busca_b200.scene import Scene
    L, C = 11, 5
    sc = Scene(T, D, L, C, seed=seed)
    ref_crop = lambda frame, boxes: model.get_image_crops(frame, boxes, normalize=False)
    tracks, dets, kal = sc.objects(ref_crop)
    case = SimpleCase(tracks, dets, kal)
    import time
    t0 = time.time()
    out = run_assoc(model, case, L, C, flavour64=True, select_highest_candidate=False)
    print(name, "reference run: %.1f s" % (time.time() - t0))
    p = out["probs"]
    kslot = min(D, C - 1)
    store = {"meta": np.array([seed, T, D, L, C]), "emb_stride": emb_stride}
    for k in ("probs_matrix", "reliable", "logits", "probs", "dists", "can_xy", "can_size", "mem_xy", "mem_size"):
        store[k] = out[k]
    store["mem_emb_sub"] = out["mem_emb"][::emb_stride]
    store["can_emb_sub"] = out["can_emb"][::emb_stride]
    store["cand_rows_sub"] = out["cand_rows"][::emb_stride]
    store["crop_sha_kalman"] = np.array([sha(k.images_mem[-1]) for k in kal[:16]])
    store["crop_sha_det"] = np.array([sha(d.images_mem[-1]) for d in dets[:16]])
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **store)
    print(name, "winners", np.bincount(p.argmax(1), minlength=C + 2), "p_kalman>0.3:", int((p[:, kslot] > 0.3).sum()), "of", T,
          "kept:", int((out["reliable"] & (p[:, kslot] > 0.3)).sum()))


class SimpleCase:
    def __init__(self, tracks, dets, kalman):
        self.tracks, self.dets, self.kalman = tracks, dets, kalman


def golden_adapter(weights_path, config="MOT20", out_name="adapter_seq", n_frames=36, n_obj=10, seed=5):
    """Drive the UNMODIFIED adapter (adapters/CenterTrack/src/lib/utils/byte_tracker.py, md5-identical to the ByteTrack copy)
    over a synthetic sequence and record, per frame, the ids / boxes it outputs and what Step 3b decided.  Shims: lap, cython_bbox
    (tests/golden/shims), np.float."""
    import importlib
    sys.path.insert(0, os.path.join(REF, "adapters/CenterTrack/src/lib"))
    bt = importlib.import_module("utils.byte_tracker")
    from utils.mot_online.basetrack import BaseTrack
    seq = synth.make_sequence(seed, n_frames, n_obj, miss=0.25)
    args, _ = load_args_from_config(os.path.join(REF, f"config/ByteTrack/{config}/config_bytetrack_{config.lower()}.yml"))
    args.use_busca, args.device, args.busca_ckpt = True, "cpu", weights_path
    args.track_thresh, args.track_buffer, args.match_thresh, args.mot20 = 0.6, 30, 0.9, config == "MOT20"
    args.transformer.reid_weights_file = "no"
    BaseTrack._count = 0
    tracker = bt.BYTETracker(args)
    # accurate (NCHW) CPU batch-norm kernels, as for the other fixtures (see run_assoc)
    enc_fwd = tracker.busca_tracker.reid_encoder.forward
    tracker.busca_tracker.reid_encoder.forward = lambda x: enc_fwd(x.contiguous())
    f64 = lambda seq_len=None, flavour="ltrb": ref_trk.missing_candidate_bbox(seq_len, flavour).astype(np.float64)
    ref_enc.missing_candidate_bbox = f64
    ref_net.missing_candidate_bbox = f64
    tracker.busca_tracker.pos_encoder.distant_fake_bbox = torch.from_numpy(f64(flavour="ltwh"))
    rec3 = {}
    orig3 = tracker.third_round_association

    def third(strack_pool, considered_dets, extra_kalman_candidates, asoc_thresh):
        m, u = orig3(strack_pool=strack_pool, considered_dets=considered_dets, extra_kalman_candidates=extra_kalman_candidates, asoc_thresh=asoc_thresh)
        rec3["pool_ids"] = [t.track_id for t in strack_pool]
        rec3["matches"] = [(int(i), float(p)) for i, p in m]
        rec3["n"] = len(strack_pool)
        return m, u

    tracker.third_round_association = third
    gate, warps = [], []
    if hasattr(args, "reliable_thresh"):
        orig_rel = tracker.is_reliable

        def rel(current_frame, active_stracks, p):
            r = orig_rel(current_frame=current_frame, active_stracks=active_stracks, p=p)
            rec3["gate"] = bool(r)
            return r
        tracker.is_reliable = rel
    if args.use_camera_motion_compensation:
        import cv2
        orig_ecc = cv2.findTransformECC

        def ecc(**kw):
            cc, w = orig_ecc(**kw)
            rec3["warp"] = np.array(w)
            return cc, w
        bt.cv2.findTransformECC = ecc
    ids, boxes, off, b3_ids, b3_keep, b3_prob, b3_off = [], [], [0], [], [], [], [0]
    import time
    t0 = time.time()
    for f in range(n_frames):
        rec3.clear()
        out = tracker.update(seq.dets[f].copy(), [seq.H, seq.W], [seq.H, seq.W], current_frame=seq.frames[f])
        ids += [t.track_id for t in out]
        boxes += [t.tlwh for t in out]
        off.append(len(ids))
        if rec3.get("n"):
            kept = {i: p for i, p in rec3["matches"]}
            b3_ids += rec3["pool_ids"]
            b3_keep += [i in kept for i in range(rec3["n"])]
            b3_prob += [kept.get(i, -1.0) for i in range(rec3["n"])]
        b3_off.append(len(b3_ids))
        gate.append(int(rec3.get("gate", -1)))
        warps.append(rec3.get("warp", np.full((2, 3), np.nan, np.float32)))
        print("frame", f + 1, "ids", [t.track_id for t in out], "3b pool", rec3.get("pool_ids"), "kept", [rec3["pool_ids"][i] for i, _ in rec3.get("matches", [])],
              "%.0f s" % (time.time() - t0), flush=True)
    ref_enc.missing_candidate_bbox = ref_trk.missing_candidate_bbox
    ref_net.missing_candidate_bbox = ref_trk.missing_candidate_bbox
    if args.use_camera_motion_compensation:
        bt.cv2.findTransformECC = orig_ecc
    np.savez_compressed(os.path.join(HERE, f"{out_name}.npz"), meta=np.array([seed, n_frames, n_obj]), miss=0.25, gate=np.array(gate), warps=np.array(warps),
                        ids=np.array(ids), boxes=np.array(boxes).reshape(-1, 4), off=np.array(off),
                        b3_ids=np.array(b3_ids), b3_keep=np.array(b3_keep, bool), b3_prob=np.array(b3_prob), b3_off=np.array(b3_off))
    print("adapter: kept-alive decisions", int(np.sum(b3_keep)), "of", len(b3_keep))


def golden_coverage():
    """8f row 2: BYTETracker.get_detection_coverage / is_reliable of the UNMODIFIED adapter (called unbound on a stub self - the two
    methods read nothing from the tracker), on seeded boxes that include negative / out-of-frame / reversed / degenerate corners."""
    import importlib
    sys.path.insert(0, os.path.join(REF, "adapters/CenterTrack/src/lib"))
    bt = importlib.import_module("utils.byte_tracker")

    class Stub:
        get_detection_coverage = bt.BYTETracker.get_detection_coverage
        is_reliable = bt.BYTETracker.is_reliable

    class Trk:
        def __init__(self, tlbr, scale):
            self.tlbr, self.scale = tlbr, scale

    rng = np.random.default_rng(23)
    store, cases = {}, []
    for k, (H, W, n, scale) in enumerate([(1080, 1920, 0, 1.0), (1080, 1920, 1, 1.0), (1080, 1920, 60, 1.0), (1080, 1920, 300, 0.75),
                                          (480, 640, 25, 1.6), (97, 1031, 40, 1.0), (720, 2500, 500, 1.0)]):
        b = synth.random_boxes(rng, n) if n else np.zeros((0, 4))
        b = b * np.array([W / 1920.0, H / 1080.0, W / 1920.0, H / 1080.0]) / scale
        b[:, 2:] += b[:, :2]
        if n >= 25:
            b[0] = [-40.7, -12.2, 31.9, 55.5] / np.float64(scale)          # negative corners: int() truncates toward zero
            b[1] = [W - 20.5, H - 30.5, W + 80.0, H + 15.0] / np.float64(scale)
            b[2] = [300.9, 200.9, 250.1, 120.1] / np.float64(scale)       # reversed corners
            b[3] = [-300.0, -300.0, -10.0, -10.0] / np.float64(scale)     # wholly outside
            b[4] = [77.3, 88.8, 77.9, 88.9] / np.float64(scale)           # a single pixel
            b[5] = [-0.9, 10.0, 0.9, 12.0] / np.float64(scale)            # -0.9 -> 0
        frame = np.zeros((H, W, 3), np.uint8)
        trks = [Trk(r.copy(), scale) for r in b]
        out = Stub().get_detection_coverage(frame, trks, [])
        store[f"c{k}_boxes"] = b
        store[f"c{k}_meta"] = np.array([H, W, scale], np.float64)
        store[f"c{k}_scalars"] = np.array([out["area_covered"], out["area_covered_per_obj"], out["max_bbox_area"], out["average_bbox_area"]], np.float64)
        store[f"c{k}_areas"] = np.array(out["bbox_areas"], np.float64)
        ps = [(0.0, 0.0), (5.0, 0.1), (1.0, 0.05), (n * 0.5, 0.0), (float(n), -1e-9)]
        store[f"c{k}_p"] = np.array(ps)
        store[f"c{k}_reliable"] = np.array([Stub().is_reliable(frame, trks, p) for p in ps], bool)
        cases.append(k)
        print("coverage case", k, H, W, n, "covered %.4f" % out["area_covered"], store[f"c{k}_reliable"])
    store["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "coverage.npz"), **store)


def golden_rounds():
    """8f row 1: the UNMODIFIED reference KalmanFilter (initiate / multi_predict / update), matching.iou_distance / fuse_score and
    remove_duplicate_stracks on seeded inputs.  cython_bbox is the shim that executes the reference's in-tree restatement; lap is not
    involved (the assignment has no reference-side golden: third-party, absent)."""
    import importlib
    sys.path.insert(0, os.path.join(REF, "adapters/CenterTrack/src/lib"))
    bt = importlib.import_module("utils.byte_tracker")
    from utils.mot_online import kalman_filter as KF
    from utils.mot_online import matching
    kf = KF.KalmanFilter()
    rng = np.random.default_rng(41)
    n, steps = 48, 6
    boxes = synth.random_boxes(rng, n)
    xyah = np.stack([boxes[:, 0] + boxes[:, 2] / 2, boxes[:, 1] + boxes[:, 3] / 2, boxes[:, 2] / boxes[:, 3], boxes[:, 3]], axis=1)
    mean = np.zeros((n, 8))
    cov = np.zeros((n, 8, 8))
    for i in range(n):
        mean[i], cov[i] = kf.initiate(xyah[i])
    store = {"kf_mean0": mean.copy(), "kf_cov0": cov.copy()}
    for s in range(steps):
        tracked = rng.uniform(size=n) < 0.8
        m_in = mean.copy()
        m_in[~tracked, 7] = 0                                   # STrack.multi_predict (byte_tracker.py:55-56)
        mean_p, cov_p = kf.multi_predict(m_in, cov)
        z = mean_p[:, :4] + rng.normal(0, 1, (n, 4)) * np.array([4.0, 4.0, 0.01, 5.0])
        upd = rng.uniform(size=n) < 0.7
        mean_u, cov_u = mean_p.copy(), cov_p.copy()
        for i in np.where(upd)[0]:
            mean_u[i], cov_u[i] = kf.update(mean_p[i], cov_p[i], z[i])
        store.update({f"kf{s}_tracked": tracked, f"kf{s}_mean_pred": mean_p, f"kf{s}_cov_pred": cov_p, f"kf{s}_z": z, f"kf{s}_upd": upd,
                      f"kf{s}_mean_upd": mean_u, f"kf{s}_cov_upd": cov_u})
        mean, cov = mean_u, cov_u
    store["kf_steps"] = np.array(steps)
    # cost matrices
    for k, (na, nb) in enumerate([(37, 53), (200, 300), (1, 7), (60, 3)]):
        a = synth.random_boxes(rng, na)
        b = synth.random_boxes(rng, nb)
        m = min(na, nb) // 2 + 1
        b[:m] = a[:m] + rng.normal(0, 6, (m, 4))
        a[:, 2:] += a[:, :2]
        b[:, 2:] += b[:, :2]
        sc = rng.uniform(0.1, 0.99, nb).astype(np.float32)        # detector scores reach the tracker as float32
        cost = matching.iou_distance(list(a), list(b))
        fused = matching.fuse_score(cost.copy(), [types.SimpleNamespace(score=v) for v in sc])
        store.update({f"m{k}_a": a, f"m{k}_b": b, f"m{k}_score": sc, f"m{k}_cost": cost, f"m{k}_fused": fused})
    store["m_cases"] = np.array(4)

    class Trk:
        def __init__(self, tlbr, frame_id, start):
            self.tlbr, self.frame_id, self.start_frame = tlbr, frame_id, start

    for k, (na, nb) in enumerate([(40, 25), (5, 1)]):
        a = synth.random_boxes(rng, na)
        b = synth.random_boxes(rng, nb)
        m = min(na, nb)
        b[:m] = a[:m] + rng.normal(0, 3, (m, 4))
        a[:, 2:] += a[:, :2]
        b[:, 2:] += b[:, :2]
        fa, sa = rng.integers(20, 60, na), rng.integers(0, 20, na)
        fb, sb = rng.integers(20, 60, nb), rng.integers(0, 20, nb)
        e = min(3, m)
        fb[:e], sb[:e] = fa[:e], sa[:e]                            # equal ages: the first list's track goes
        ta = [Trk(a[i], int(fa[i]), int(sa[i])) for i in range(na)]
        tb = [Trk(b[i], int(fb[i]), int(sb[i])) for i in range(nb)]
        ra, rb = bt.remove_duplicate_stracks(ta, tb)
        store.update({f"d{k}_a": a, f"d{k}_b": b, f"d{k}_age_a": fa - sa, f"d{k}_age_b": fb - sb,
                      f"d{k}_keep_a": np.array([t in ra for t in ta]), f"d{k}_keep_b": np.array([t in rb for t in tb])})
        print("duplicates case", k, "dropped", na - len(ra), nb - len(rb))
    store["d_cases"] = np.array(2)
    np.savez_compressed(os.path.join(HERE, "rounds.npz"), **store)
    print("rounds golden written")


def golden_ingest():
    """8f row 4: the evaluator's de-normalisation statements (adapters/ByteTrack/yolox/evaluators/mot_evaluator.py:198-204) executed
    as they stand - the source lines are read from the reference file and exec'd with `.cuda()` stripped (no GPU in this container;
    the fp32 multiply and add are IEEE operations on either device) - and write_results (:30-40) through a temp file."""
    path = os.path.join(REF, "adapters/ByteTrack/yolox/evaluators/mot_evaluator.py")
    lines = open(path).read().split("\n")
    first = next(i for i, l in enumerate(lines) if "rgb_means = torch.tensor" in l)
    block = [l.strip().replace(".cuda()", "") for l in lines[first:first + 7]]
    assert block[-1].startswith("vot_img = (vot_img * 255.0).astype(np.uint8)"), block
    rng = np.random.default_rng(61)
    store = {}
    for k, (H, W) in enumerate([(800, 1440), (37, 53), (608, 1088)]):
        means, std = synth.YOLOX_MEANS, synth.YOLOX_STD
        chw = synth.make_detector_tensor(61 + k, H, W)
        pre = types.SimpleNamespace(means=means, std=std)
        ns = {"torch": torch, "np": np, "imgs": [torch.from_numpy(chw)],
              "self": types.SimpleNamespace(dataloader=types.SimpleNamespace(dataset=types.SimpleNamespace(preproc=pre)))}
        exec("\n".join(block), ns)
        store[f"i{k}_seed"] = np.array(61 + k)
        store[f"i{k}_shape"] = np.array([H, W])
        store[f"i{k}_sha"] = sha(ns["vot_img"])
        if H * W < 4000:
            store[f"i{k}_bgr"] = ns["vot_img"]
        print("ingest case", k, H, W, store[f"i{k}_sha"][:16])
    store["i_cases"] = np.array(3)
    # write_results
    import importlib.util, tempfile
    src = open(path).read()
    a = src.index("def write_results(")
    b = src.index("def write_results_no_score(")
    ns = {"logger": types.SimpleNamespace(info=lambda *a, **k: None)}
    exec(src[a:b], ns)
    res = []
    for f in range(1, 6):
        n = int(rng.integers(0, 6))
        res.append((f, [tuple(rng.uniform(-5, 1900, 4)) for _ in range(n)], [int(v) for v in rng.integers(-1, 40, n)],
                    [float(np.float32(v)) for v in rng.uniform(0.1, 1, n)]))
    with tempfile.NamedTemporaryFile("r", suffix=".txt") as tf:
        ns["write_results"](tf.name, res)
        store["mot_txt"] = np.array(open(tf.name).read())
    store["mot_rows"] = np.array([[f, tid, *tlwh, s] for f, tl, ids, sc in res for tlwh, tid, s in zip(tl, ids, sc)], np.float64)
    np.savez_compressed(os.path.join(HERE, "ingest.npz"), **store)


ECC_CASES = [(3, 1080, 1920, 0.004, 3.3, -2.1), (4, 1080, 1920, 0.0, 1.0, 0.0), (5, 1080, 1920, -0.01, -6.5, 4.25), (6, 480, 640, 0.002, 0.4, 0.7),
             (7, 97, 131, 0.0, -1.5, 2.0)]


def golden_ecc():
    """8f row 3: cv2.cvtColor + cv2.findTransformECC exactly as BYTETracker.camera_motion_compensation calls them
    (byte_tracker.py:640-645) on seeded frame pairs."""
    import cv2
    store = {"cases": np.array(ECC_CASES, np.float64), "cv2_version": np.array(cv2.__version__)}
    for k, (seed, H, W, th, tx, ty) in enumerate(ECC_CASES):
        f1 = synth.make_frame(seed, H, W)
        f2 = synth.make_moved_frame(f1, th, tx, ty, seed)
        im1_gray = cv2.cvtColor(f1, cv2.COLOR_BGR2GRAY)
        im2_gray = cv2.cvtColor(f2, cv2.COLOR_BGR2GRAY)
        warp_matrix = np.eye(2, 3, dtype=np.float32)
        criteria = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 100, 0.00001)
        cc, warp_matrix = cv2.findTransformECC(templateImage=im1_gray, inputImage=im2_gray, warpMatrix=warp_matrix, motionType=cv2.MOTION_EUCLIDEAN,
                                               criteria=criteria)
        store[f"e{k}_warp"], store[f"e{k}_cc"] = warp_matrix, np.array(cc)
        store[f"e{k}_gray_sha"] = np.array(sha(im2_gray))
        print("ecc case", k, cc, warp_matrix.ravel())
    np.savez_compressed(os.path.join(HERE, "ecc.npz"), **store)


if __name__ == "__main__":
    which = sys.argv[1:] or ["crops", "geometry", "pe", "assoc"]
    if "ecc" in which:
        golden_ecc()
        sys.exit(0)
    if "ingest" in which:
        golden_ingest()
        sys.exit(0)
    if "coverage" in which or "rounds" in which:
        if "coverage" in which:
            golden_coverage()
        if "rounds" in which:
            golden_rounds()
        sys.exit(0)
    if any(w in ("cond", "scene", "scene_mot20", "adapter", "adapter_mot17") for w in which):
        model, targs = build_reference(profile="conditioned")
        if "cond" in which:
            golden_assoc_cond(model)
        if "scene" in which:
            golden_scene(model, "scene_cfg1_cond", 16, 40, seed=11, emb_stride=1)
        if "scene_mot20" in which:
            golden_scene(model, "scene_mot20_cond", 200, 300, seed=0, emb_stride=8)
        if "adapter" in which:
            golden_adapter("/tmp/busca_golden_weights_conditioned.pth")
        if "adapter_mot17" in which:
            golden_adapter("/tmp/busca_golden_weights_conditioned.pth", config="MOT17", out_name="adapter_seq_mot17", n_frames=24, n_obj=30, seed=8)
        sys.exit(0)
    model, targs = build_reference()
    if "crops" in which:
        golden_crops(model)
    if "geometry" in which:
        golden_geometry()
    if "pe" in which:
        golden_pe(model)
    if "assoc" in which:
        golden_assoc(model)
