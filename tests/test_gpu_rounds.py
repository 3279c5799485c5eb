"""GPU parity tests (-m gpu) of the SURVEY.md 8(f) rows: the host tracker's association rounds (rounds.cu) and the detection-coverage
gate (geometry.cu), through the C ABI, against fixtures of the unmodified reference (tests/golden/rounds.npz, coverage.npz) and the
oracle.  Bit-exact wherever the reference's arithmetic is order-independent; the Kalman update (LAPACK in the reference) within 1e-10."""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from busca_b200 import synth

KF_UPDATE_TOL = 1e-10          # relative to the largest entry of the track's state / covariance; measured ~1e-13


@pytest.fixture(scope="module")
def engine():
    from busca_b200.engine import Engine
    return Engine(device=0, bank_slots=8)


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "rounds.npz"))


# ------------------------------------------------------------------------------------------------ 8f row 2
def test_detection_coverage_equals_reference(engine, golden_dir):
    from busca_b200 import tracking
    c = np.load(os.path.join(golden_dir, "coverage.npz"))
    for k in c["cases"]:
        H, W, scale = c[f"c{k}_meta"]
        H, W = int(H), int(W)
        boxes = c[f"c{k}_boxes"] * scale
        out = tracking.get_detection_coverage((H, W, 3), boxes, engine=engine)
        got = np.array([out["area_covered"], out["area_covered_per_obj"], out["max_bbox_area"], out["average_bbox_area"]])
        assert np.array_equal(got, c[f"c{k}_scalars"]), (k, got, c[f"c{k}_scalars"])
        assert np.array_equal(np.array(out["bbox_areas"], np.float64), c[f"c{k}_areas"])
        rel = [tracking.is_reliable((H, W, 3), boxes, p, engine=engine) for p in c[f"c{k}_p"]]
        assert np.array_equal(np.array(rel, bool), c[f"c{k}_reliable"])


def test_detection_coverage_random_vs_oracle(engine):
    from oracle import coverage as ocov
    rng = np.random.default_rng(77)
    for H, W, n in [(1080, 1920, 700), (33, 47, 20), (2160, 3840, 150), (5, 1025, 3)]:
        b = np.concatenate([rng.uniform(-50, W + 50, (n, 1)), rng.uniform(-50, H + 50, (n, 1)),
                            rng.uniform(-50, W + 50, (n, 1)), rng.uniform(-50, H + 50, (n, 1))], axis=1)
        if n >= 100:
            b[:, 2:] = b[:, :2] + rng.uniform(0, 120, (n, 2))
        cnt, areas = engine.detection_coverage(b, H, W)
        want = ocov.detection_coverage((H, W), b)
        assert cnt == want["nonzero"], (H, W, n)
        assert np.array_equal(areas, np.array(want["bbox_areas"]))


# ------------------------------------------------------------------------------------------------ 8f row 1
def test_kalman_predict_bit_exact(engine, g):
    mean, cov = g["kf_mean0"], g["kf_cov0"]
    for s in range(int(g["kf_steps"])):
        mp, cp = engine.kalman_predict(mean, cov, g[f"kf{s}_tracked"])
        assert np.array_equal(mp, g[f"kf{s}_mean_pred"]), s
        assert np.array_equal(cp, g[f"kf{s}_cov_pred"]), s
        mean, cov = g[f"kf{s}_mean_upd"], g[f"kf{s}_cov_upd"]
    m2, c2 = engine.kalman_predict(g["kf_mean0"], g["kf_cov0"], None)                 # no flags: nothing is zeroed
    from oracle import rounds as ornd
    mo, co = ornd.kf_multi_predict(g["kf_mean0"], g["kf_cov0"], None)
    assert np.array_equal(m2, mo) and np.array_equal(c2, co)


def test_kalman_update_vs_reference(engine, g):
    worst = 0.0
    for s in range(int(g["kf_steps"])):
        mp, cp, z, upd = g[f"kf{s}_mean_pred"], g[f"kf{s}_cov_pred"], g[f"kf{s}_z"], g[f"kf{s}_upd"]
        mu, cu = engine.kalman_update(mp[upd], cp[upd], z[upd])
        wm, wc = g[f"kf{s}_mean_upd"][upd], g[f"kf{s}_cov_upd"][upd]
        em = np.abs(mu - wm).max(axis=1) / np.abs(wm).max(axis=1)
        ec = np.abs(cu - wc).reshape(len(wc), -1).max(axis=1) / np.abs(wc).reshape(len(wc), -1).max(axis=1)
        worst = max(worst, em.max(), ec.max())
    print(f"kalman update: worst relative deviation from scipy/LAPACK {worst:.2e}")
    assert worst < KF_UPDATE_TOL


def test_match_cost_bit_exact(engine, g):
    for k in range(int(g["m_cases"])):
        a, b, sc = g[f"m{k}_a"], g[f"m{k}_b"], g[f"m{k}_score"]
        _, _, cost = engine.match_round(a, b, None, 0.9, want_cost=True)
        assert np.array_equal(cost, g[f"m{k}_cost"]), k
        _, _, fused = engine.match_round(a, b, sc.astype(np.float64), 0.9, want_cost=True)
        assert np.array_equal(fused, g[f"m{k}_fused"]), k


def _check_assignment(cost, thresh, x, y):
    from oracle import rounds as ornd
    xo, yo = ornd.linear_assignment(cost, thresh)
    n, m = cost.shape
    # a valid partial matching
    assert all(0 <= j < m and y[j] == i for i, j in enumerate(x) if j >= 0)
    assert all(0 <= i < n and x[i] == j for j, i in enumerate(y) if i >= 0)
    if not np.array_equal(x, xo):                                   # only legitimate when the optimum is not unique
        assert abs(ornd.assignment_objective(cost, x, thresh) - ornd.assignment_objective(cost, xo, thresh)) < 1e-9
        return False
    return True


def test_assignment_equals_oracle(engine, g):
    rng = np.random.default_rng(123)
    same = total = 0
    for n, m in [(1, 1), (1, 9), (9, 1), (7, 7), (37, 53), (64, 20), (200, 300), (500, 300), (300, 520)]:
        for thresh in (0.5, 0.7, 0.9):
            cost = rng.uniform(0, 1.3, (n, m))
            x, y = engine.linear_assignment(cost, thresh)
            same += _check_assignment(cost, thresh, x, y)
            total += 1
    assert same == total                                            # continuous random costs: the optimum is unique
    # IoU-distance matrices (many exact 1.0 entries), as the rounds see them, straight from boxes
    for k in range(int(g["m_cases"])):
        for sc, thresh in ((None, 0.5), (g[f"m{k}_score"].astype(np.float64), 0.9)):
            x, y, cost = engine.match_round(g[f"m{k}_a"], g[f"m{k}_b"], sc, thresh, want_cost=True)
            assert _check_assignment(cost, thresh, x, y)
    x, y = engine.linear_assignment(np.zeros((0, 5)), 0.9)
    assert len(x) == 0 and (y == -1).all()
    x, y, _ = engine.match_round(np.zeros((3, 4)), np.zeros((0, 4)), None, 0.9)
    assert (x == -1).all() and len(y) == 0


def test_assignment_mot20_scale_latency(engine):
    """A crowded frame: 500 pooled tracks against 400 detections placed on them (IoU costs), one call = cost matrix + assignment."""
    rng = np.random.default_rng(5)
    a = synth.random_boxes(rng, 500)
    b = np.concatenate([a[:350] + rng.normal(0, 4, (350, 4)), synth.random_boxes(rng, 50)])
    a[:, 2:] += a[:, :2]
    b[:, 2:] += b[:, :2]
    x, y, cost = engine.match_round(a, b, None, 0.9, want_cost=True)
    assert _check_assignment(cost, 0.9, x, y)
    t0 = time.perf_counter()
    for _ in range(5):
        engine.match_round(a, b, None, 0.9)
    ms = (time.perf_counter() - t0) / 5 * 1e3
    from oracle import rounds as ornd
    t0 = time.perf_counter()
    ornd.linear_assignment(ornd.iou_distance(a, b), 0.9)
    print(f"match_round 500 x 400: {ms:.2f} ms on the device (host buffers in and out); scipy on the extended matrix: {(time.perf_counter() - t0) * 1e3:.2f} ms")
    assert (x >= 0).sum() >= 300


def test_duplicates_equal_reference(engine, g):
    for k in range(int(g["d_cases"])):
        da, db = engine.duplicate_tracks(g[f"d{k}_a"], g[f"d{k}_age_a"], g[f"d{k}_b"], g[f"d{k}_age_b"], 0.15)
        assert np.array_equal(~da, g[f"d{k}_keep_a"]) and np.array_equal(~db, g[f"d{k}_keep_b"])
    da, db = engine.duplicate_tracks(np.zeros((0, 4)), np.zeros(0), g["d0_b"], g["d0_age_b"], 0.15)
    assert len(da) == 0 and not db.any()


# ------------------------------------------------------------------------------------------------ 8f row 4
def test_frame_ingest_bit_exact(engine, golden_dir):
    """mot_evaluator.py:198-204 on the device: sha256 of the uint8 BGR frame equal to the reference's, from a host tensor and from a
    tensor that already lives in HBM; the ingested frame is the one the crops read."""
    import hashlib
    from oracle import crop as ocrop
    gi = np.load(os.path.join(golden_dir, "ingest.npz"))
    for k in range(int(gi["i_cases"])):
        H, W = (int(v) for v in gi[f"i{k}_shape"])
        chw = synth.make_detector_tensor(int(gi[f"i{k}_seed"]), H, W)
        out = engine.ingest_frame(chw, synth.YOLOX_MEANS, synth.YOLOX_STD)
        assert hashlib.sha256(out.tobytes()).hexdigest() == str(gi[f"i{k}_sha"]), k
        p = engine.dev_alloc(chw.nbytes)
        engine.h2d(p, chw)
        out2 = engine.ingest_frame(p, synth.YOLOX_MEANS, synth.YOLOX_STD, H, W)
        engine.dev_free(p)
        assert np.array_equal(out, out2)
        # crops come from the ingested frame without an upload
        boxes = np.array([[3.2, 4.1, 30.7, 33.9], [-5.0, -3.0, 20.0, 25.0]])
        slots = engine.alloc_slots(2)
        got = engine.crop(boxes, slots)
        engine.free_slots(slots)
        assert np.array_equal(got, ocrop.get_image_crops(out, boxes))
        assert engine.sync_frame(out) is False                      # the returned host copy is recognised as the resident frame


def test_frame_ingest_bandwidth(engine):
    H, W = 1080, 1920
    chw = synth.make_detector_tensor(1, H, W)
    p = engine.dev_alloc(chw.nbytes)
    engine.h2d(p, chw)
    engine.ingest_frame(p, synth.YOLOX_MEANS, synth.YOLOX_STD, H, W, to_host=False)
    engine.set_profiling(True)
    engine.ingest_frame(p, synth.YOLOX_MEANS, synth.YOLOX_STD, H, W, to_host=False)
    prof = engine.last_profile()
    engine.set_profiling(False)
    engine.dev_free(p)
    assert "frame_ingest" in prof, prof
    ms = prof["frame_ingest"]["ms"]
    print(f"frame_ingest 1080p: {ms * 1e3:.1f} us, {H * W * 15 / ms / 1e6:.0f} GB/s of algorithmic bytes (12 B read + 3 B written per pixel)")


# ------------------------------------------------------------------------------------------------ 8f row 3
def test_camera_motion_equals_cv2(engine, golden_dir):
    """BYTETracker.camera_motion_compensation's warp matrix (cv2.cvtColor + cv2.findTransformECC, Euclidean, 100 iterations, eps 1e-5)
    against tests/golden/ecc.npz (cv2 4.13 outputs): translation within 2e-3 px, rotation entries within 2e-6, rho within 1e-5."""
    from test_host_rounds import check_warp, ecc_cases
    worst_t = worst_r = 0.0
    for k, f1, f2, want, want_rho, _sha in ecc_cases(golden_dir):
        warp, rho, it = engine.camera_motion(f1, f2)
        check_warp(k, warp, rho, want, want_rho)
        worst_t, worst_r = max(worst_t, np.abs(warp[:, 2] - want[:, 2]).max()), max(worst_r, np.abs(warp[:, :2] - want[:, :2]).max())
        # the cached path: the same pair again, the current frame first as a 'previous' call, then from the frame in HBM
        engine.camera_motion(f2, f1)                                   # now f1 is the cached previous frame
        engine.upload_frame(f2)
        warp2, rho2, _ = engine.camera_motion(None, None, shape=f2.shape)
        assert np.array_equal(warp, warp2) and rho == rho2, k           # deterministic (fixed-order reduction)
    print(f"camera motion vs cv2: worst translation deviation {worst_t:.2e} px, worst rotation-entry deviation {worst_r:.2e}")


def test_camera_motion_latency_and_errors(engine):
    from busca_b200._lib import BuscaError
    f1 = synth.make_frame(21)
    f2 = synth.make_moved_frame(f1, 0.003, 2.2, -1.3, 21)
    engine.camera_motion(f1, f2)
    engine.set_profiling(True)
    t0 = time.perf_counter()
    warp, rho, it = engine.camera_motion(f1, f2)
    ms = (time.perf_counter() - t0) * 1e3
    prof = engine.last_profile()
    engine.set_profiling(False)
    print(f"camera motion 1080p: {ms:.2f} ms host to host ({it} iterations; kernels: "
          f"{prof['ecc_iteration']['ms']:.3f} ms in {prof['ecc_iteration']['launches']} launches, prepare {prof['ecc_prepare']['ms']:.3f} ms)")
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    b = 255 - a                                                         # anti-correlated: cv2 raises StsNoConv
    with pytest.raises(BuscaError):
        engine.camera_motion(a, b)
    with pytest.raises(BuscaError):
        engine.camera_motion(None, f2[:100, :100])                      # no cached frame of that size


def test_deferred_crop_copies_are_the_same_bytes(golden_dir):
    """Option defer_crop_copies: get_image_crops returns before the device->host copy has landed; after the next waiting call the host
    bytes (and the bank slots the association reads) are identical to the strict path."""
    from types import SimpleNamespace
    from busca_b200.network import BUSCA
    from busca_b200.option import load_args_from_config
    here = os.path.dirname(os.path.abspath(__file__))
    targs, _ = load_args_from_config(os.path.join(os.path.dirname(here), "busca_b200", "configs", "bytetrack_mot20.yml"))
    a = targs.transformer
    a.device = "cuda:0"
    strict = BUSCA(a)
    a.defer_crop_copies = True
    lazy = BUSCA(a)
    frame = synth.make_frame(2)
    rng = np.random.default_rng(4)
    b = synth.random_boxes(rng, 40)
    b[:, 2:] += b[:, :2]
    want_big = strict.get_image_crops(frame, b, normalize=False)
    got_big = lazy.get_image_crops(frame, b, normalize=False)
    singles_w = [strict.get_image_crops(frame, [r], normalize=False)[0] for r in b[:12]]
    singles_g = [lazy.get_image_crops(frame, [r], normalize=False)[0] for r in b[:12]]
    lazy.sync()
    assert np.array_equal(want_big, got_big)
    for w, g2 in zip(singles_w, singles_g):
        assert np.array_equal(w, g2)
    slots = [lazy._registry.lookup(g2) for g2 in singles_g]
    assert all(s is not None for s in slots)
    assert np.array_equal(lazy.engine.bank_download(np.array(slots, np.int32)), np.stack(singles_w))


def test_patch_bank_eviction_and_verification():
    """8f row 4 (avoid_memory_leak, GHOST tracker.py:249-258): bank slots return to the pool when the adapter drops a crop; and the debug
    option verify_patches catches an adapter that rewrites a stored crop in place (the bank would serve the old pixels)."""
    import gc
    from busca_b200.network import BUSCA
    from busca_b200.option import load_args_from_config
    from oracle import crop as ocrop
    here = os.path.dirname(os.path.abspath(__file__))
    targs, _ = load_args_from_config(os.path.join(os.path.dirname(here), "busca_b200", "configs", "bytetrack_mot20.yml"))
    a = targs.transformer
    a.device, a.verify_patches = "cuda:0", True
    m = BUSCA(a).eval()
    m.load_state_dict(synth.make_weights(0, profile="conditioned"))
    base = m.engine.slots_in_use()
    case = synth.make_assoc_case(7, 3, 6, targs.seq_len, crop_fn=lambda f, b: m.get_image_crops(f, b, normalize=False))
    used = m.engine.slots_in_use()
    assert used > base
    from busca_b200 import tracking
    dists = tracking.center_distance(case.tracks, case.dets, engine=m.engine)
    kw = dict(seq_len=targs.seq_len, num_candidates=targs.num_candidates, use_broader_memory=targs.use_broader_memory,
              select_highest_candidate=False, normalize_ims=True)
    pm, rel = m.associate_embeddings(case.tracks, case.dets, dists, extra_kalman_candidates=case.kalman, **kw)   # verification passes on untouched crops
    assert pm is not None and m.engine.slots_in_use() == used                        # temporary slots of the call are returned
    del case, pm, rel, dists
    gc.collect()
    assert m.engine.slots_in_use() == base, (m.engine.slots_in_use(), base)           # the adapter dropped its crops: every slot is back

    def tampered():
        c2 = synth.make_assoc_case(8, 3, 6, targs.seq_len, crop_fn=lambda f, b: m.get_image_crops(f, b, normalize=False))
        d2 = tracking.center_distance(c2.tracks, c2.dets, engine=m.engine)
        kw2 = dict(kw, extra_kalman_candidates=c2.kalman)
        c2.tracks[0].images_mem[-1][10:20, 10:20] ^= 0xFF                            # an adapter writing into a stored crop
        try:
            m.associate_embeddings(c2.tracks, c2.dets, d2, **kw2)
        except RuntimeError as ex:
            return "modified in place" in str(ex)
        return False
    assert tampered()


def test_frame_geometry_batch_equals_per_frame_calls(engine):
    """busca_frame_geometry_batch (B frames of B sequences in one launch) = B calls of busca_frame_geometry, bit for bit."""
    rng = np.random.default_rng(31)
    B, T, D, Cn = 7, 23, 41, 5
    mean = np.concatenate([rng.uniform(0, 1900, (B, T, 2)), rng.uniform(0.2, 0.8, (B, T, 1)), rng.uniform(60, 300, (B, T, 1)), rng.normal(0, 3, (B, T, 4))], axis=2)
    tracked = rng.uniform(size=(B, T)) < 0.7
    det = np.stack([synth.random_boxes(rng, D) for _ in range(B)])
    det[:, :, 2:] += det[:, :, :2]
    got = engine.frame_geometry_batch(mean, tracked, det, Cn)
    for b in range(B):
        want = engine.frame_geometry(mean[b], tracked[b], det[b], Cn)
        for k in ("tlwh", "tlbr", "dist", "iou", "cand"):
            assert np.array_equal(got[k][b], want[k]), (b, k)


def test_round_argument_errors(engine):
    from busca_b200._lib import BuscaError
    with pytest.raises(BuscaError):
        engine.linear_assignment(np.zeros((3, 3)), float("inf"))            # no finite 'unassigned' option: rejected, not a hang
    with pytest.raises(BuscaError):
        engine.linear_assignment(np.zeros((1, 9000)), 0.5)                  # rows + columns beyond the shared-memory solver
    x, y = engine.linear_assignment(np.full((4, 3), np.nan), 0.5)           # NaN costs: nothing is assignable
    assert (x == -1).all() and (y == -1).all()
    x, y = engine.linear_assignment(np.full((2, 2), 0.5), 0.5)              # cost == limit: a tie between matching and not; either is optimal
    assert set(x.tolist()) <= {-1, 0, 1}
