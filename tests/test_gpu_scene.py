"""GPU parity of the BENCHMARKED entry points at the benchmarked scales (-m gpu), on the conditioned weight set
(synth.make_weights(profile="conditioned"): winners vary from track to track, Kalman-slot probabilities straddle busca_thresh):

  * busca_frame_step_dev  (bench.py `value`)  and  the plug-in path get_image_crops + center_distance + associate_embeddings
    (bench.py `e2e`) on busca_b200.scene.Scene - the very object bench.py times - against tests/golden/scene_*.npz, the
    UNMODIFIED reference's associate_embeddings on the same scene (tests/golden/make_golden.py scene / scene_mot20);
  * the bf16 tensor-core path held to the tolerance SURVEY.md Appendix C.6 proposes.

Tolerances (stated once, used everywhere below):
  fp32: probabilities / embeddings within 1e-3 (north_star), decisions identical outside 2e-3 near-ties.
  bf16: embedding cosine >= 0.999 and rel-L2 <= 2.5e-2; probabilities: max |dp| <= 3.5e-2, 99 % of them within 2e-2, 95 % within
        1.5e-2, half within 4e-3; decisions identical outside 3.5e-2 near-ties of the threshold (arg-max: outside 7e-2 top-1 / top-2
        margins, two probabilities being involved).  Measured on a B200 at 200 tracks / 1400 probabilities, 13 runs (the fp32
        reductions are atomics, so runs differ; tests/probe_bf16_sources.py): max 2.1e-2 .. 3.0e-2 (mean 2.5e-2, s.d. 0.2e-2), p99
        1.5-1.7e-2, p95 0.9-1.0e-2, median 1.5e-3 - all of it from the bf16 ReID encoder: with the Decision Transformer in fp32 the
        figures are the same.  CPU emulation of the bf16 roundings (tests/analysis_weights.py, 24 tracks): cosine 0.9998, rel-L2 <=
        2.0e-2, max |dp| 1.3e-2, median 1.6e-3."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from busca_b200 import synth
from busca_b200.scene import Scene

TOL = {"fp32": dict(prob=1e-3, p99=1e-3, p95=1e-3, p50=1e-3, tie=2e-3, rows=1e-3), "bf16": dict(prob=3.5e-2, p99=2e-2, p95=1.5e-2, p50=4e-3, tie=3.5e-2, rows=6e-2)}
BF16_COS, BF16_L2 = 0.999, 2.5e-2
THRESH = 0.3                      # busca_thresh of config_bytetrack_mot20.yml


def build(precision, bank_slots=8192):
    from busca_b200.network import BUSCA
    from busca_b200.option import load_args_from_config
    here = os.path.dirname(os.path.abspath(__file__))
    targs, _ = load_args_from_config(os.path.join(os.path.dirname(here), "busca_b200", "configs", "bytetrack_mot20.yml"))
    a = targs.transformer
    a.device, a.precision, a.bank_slots = "cuda:0", precision, bank_slots
    m = BUSCA(a).eval()
    m.load_state_dict(synth.make_weights(0, profile="conditioned"))
    return m


@pytest.fixture(scope="module")
def models():
    cache = {}

    def get(precision):
        if precision not in cache:
            cache[precision] = build(precision)
        return cache[precision]

    return get


def check_decisions(probs, reliable, ref_probs, ref_reliable, kslot, tol, min_clear=0.8):
    """byte_tracker.py:504-526 on both sides; rows whose reference score is within `tie` of the threshold (or whose top-1 /
    top-2 margin is below it, for the arg-max) are the documented near-ties."""
    assert np.array_equal(reliable, ref_reliable)
    dp = np.abs(probs - ref_probs)
    assert dp.max() < tol["prob"], dp.max()
    assert np.percentile(dp, 99) < tol["p99"], np.percentile(dp, 99)
    assert np.percentile(dp, 95) < tol["p95"], np.percentile(dp, 95)
    assert np.percentile(dp, 50) < tol["p50"], np.percentile(dp, 50)
    clear = np.abs(ref_probs[:, kslot] - THRESH) > tol["tie"]
    keep, ref_keep = reliable & (probs[:, kslot] > THRESH), ref_reliable & (ref_probs[:, kslot] > THRESH)
    assert clear.mean() > min_clear, clear.mean()
    assert np.array_equal(keep[clear], ref_keep[clear])
    assert ref_keep.any() and not ref_keep.all()          # the workload is not degenerate: some tracks are kept alive, some are not
    srt = np.sort(ref_probs, axis=1)
    clear_top = (srt[:, -1] - srt[:, -2]) > 2 * tol["tie"]      # two probabilities, each within the bound: the arg-max can flip inside twice the bound
    assert clear_top.mean() > min_clear - 0.15, clear_top.mean()
    assert np.array_equal(probs.argmax(1)[clear_top], ref_probs.argmax(1)[clear_top])
    assert len(np.unique(ref_probs.argmax(1))) >= 3        # ... and different tracks pick different winners
    return keep, ref_keep, clear


@pytest.mark.parametrize("name,precision", [("scene_cfg1_cond", "fp32"), ("scene_cfg1_cond", "bf16"),
                                            ("scene_mot20_cond", "bf16"), ("scene_mot20_cond", "fp32")])
def test_frame_step_dev_vs_reference(models, golden_dir, name, precision):
    """The device-resident step bench.py reports as `value`: probabilities, proposal table and keep decisions."""
    from oracle import geometry as ogeo
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    seed, T, D, L, C = (int(v) for v in g["meta"])
    m = models(precision)
    sc = Scene(T, D, L, C, seed=seed)
    sc.setup_resident(m, busca_thresh=THRESH)
    n0 = m.engine.launches
    sc.step_resident()
    m.engine.sync()
    assert m.engine.launches - n0 > 100
    out = sc.read_resident()
    tol = TOL[precision]
    # integer work: the proposal table is the reference's (nearest detections by the reference's own distances, Kalman slot last)
    idx, n_avail = ogeo.select_candidates(g["dists"], C, True)
    assert np.array_equal(out["cand"], idx)
    pm = g["probs_matrix"]
    for t in range(0, T, max(1, T // 16)):
        assert np.array_equal(np.nonzero(pm[t])[0], np.sort(idx[t, :n_avail]))
    kslot = min(D, C - 1)
    keep, ref_keep, clear = check_decisions(out["probs"], sc.reliable, g["probs"], g["reliable"], kslot, tol, min_clear=0.8 if T >= 100 else 0.6)
    assert np.array_equal(out["keep"], keep)                  # decide_kernel == the host rule on the same probabilities
    # select_highest_candidate flavour of the decision (MOT17 YAMLs: thresh 0.5 on the one-hot of the arg-max)
    sc.step_args.select_highest, sc.step_args.busca_thresh = 1, 0.5
    sc.step_resident()
    m.engine.sync()
    out2 = sc.read_resident()
    assert np.array_equal(out2["keep"], sc.reliable & (out2["probs"].argmax(1) == kslot))


@pytest.mark.parametrize("name,precision", [("scene_cfg1_cond", "fp32"), ("scene_mot20_cond", "bf16")])
def test_plugin_path_on_the_benchmarked_scene(models, golden_dir, name, precision):
    """bench.py's `e2e` leg (host buffers through the reference-facing API) against the same golden."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    seed, T, D, L, C = (int(v) for v in g["meta"])
    m = models(precision)
    sc = Scene(T, D, L, C, seed=seed)
    sc.setup_e2e(m)
    r = sc.step_e2e(0, busca_thresh=THRESH, details=True)
    tol = TOL[precision]
    assert np.array_equal(r["dists"], g["dists"])              # fp64 centre distances: bit-exact
    pm, ref = r["probs_matrix"], g["probs_matrix"]
    assert np.array_equal(pm > 0, ref > 0)
    assert np.abs(pm - ref).max() < tol["prob"]
    kal, ref_kal = pm[np.arange(T), D + np.arange(T)], ref[np.arange(T), D + np.arange(T)]
    clear = np.abs(ref_kal - THRESH) > tol["tie"]
    assert clear.mean() > 0.8
    assert np.array_equal((r["reliable"] & (kal > THRESH))[clear], (g["reliable"] & (ref_kal > THRESH))[clear])
    stride = int(g["emb_stride"])
    rows, ref_rows = m.logits[::stride], g["cand_rows_sub"]
    assert np.abs(rows - ref_rows).max() / np.abs(ref_rows).max() < tol["rows"]


@pytest.mark.parametrize("name", ["assoc_cfg1_cond", "assoc_fewdets_cond"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_association_conditioned_weights(models, golden_dir, name, precision):
    """Whole path through the C ABI with embeddings exposed: fp32 at 1e-3, bf16 at the stated bound."""
    from busca_b200 import tracking
    m = models(precision)
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    seed, T, D, L, C, short = (int(v) for v in g["meta"])
    case = synth.make_assoc_case(seed, T, D, L, crop_fn=lambda f, b: m.get_image_crops(f, b, normalize=False), short_history=short)
    dists = tracking.center_distance(case.tracks, case.dets, engine=m.engine)
    assert np.array_equal(dists, g["f64_dists"])
    mem_slots = np.full((T, L), -1, np.int32)
    mem_ltwh = np.tile(np.array([250.0, 250.0, 500.0, 500.0]), (T, L, 1))
    for t, tr in enumerate(case.tracks):
        sel = m._memory_indices(len(tr.images_mem), L, True)
        if len(sel) == L:
            for i, j in enumerate(sel):
                mem_slots[t, i] = m._registry.lookup(tr.images_mem[j])
                mem_ltwh[t, i] = tr.tlwh_mem[j] * tr.scale
    det_slots = np.array([m._registry.lookup(d.images_mem[-1]) for d in case.dets], np.int32)
    det_ltwh = np.array([d.tlwh_mem[-1] * d.scale for d in case.dets])
    kal_slots = np.array([m._registry.lookup(k.images_mem[-1]) for k in case.kalman], np.int32)
    kal_ltwh = np.array([k.tlwh * k.scale for k in case.kalman])
    out = m.engine.associate(mem_slots, mem_ltwh, det_slots, det_ltwh, dists, kal_slots, kal_ltwh, L, C,
                             want=("probs", "cand", "mem_emb", "can_emb", "pe_index", "logits"))
    tol = TOL[precision]
    for mine, ref in ((out["mem_emb"], g["f64_mem_emb"]), (out["can_emb"], g["f64_can_emb"])):
        a, b = mine.reshape(-1, 512).astype(np.float64), ref.reshape(-1, 512).astype(np.float64)
        if precision == "fp32":
            assert np.abs(a - b).max() / np.abs(b).max() < 1e-3
        else:
            cos = (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))
            l2 = np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)
            assert cos.min() >= BF16_COS, cos.min()
            assert l2.max() <= BF16_L2, l2.max()
    assert np.array_equal(out["pe_index"][:, :L, 0], g["f64_mem_xy"]) and np.array_equal(out["pe_index"][:, L:, 0], g["f64_can_xy"])
    assert np.array_equal(out["pe_index"][:, :L, 1], g["f64_mem_size"]) and np.array_equal(out["pe_index"][:, L:, 1], g["f64_can_size"])
    ref_p = g["f64_probs"]
    assert np.abs(out["probs"] - ref_p).max() < tol["prob"], np.abs(out["probs"] - ref_p).max()
    kslot = min(D, C - 1)
    clear = np.abs(ref_p[:, kslot] - THRESH) > tol["tie"]
    assert np.array_equal((out["probs"][:, kslot] > THRESH)[clear], (ref_p[:, kslot] > THRESH)[clear])
    srt = np.sort(ref_p, axis=1)
    clear_top = (srt[:, -1] - srt[:, -2]) > 2 * tol["tie"]
    assert np.array_equal(out["probs"].argmax(1)[clear_top], ref_p.argmax(1)[clear_top])
    if T >= 16:
        assert clear.mean() > 0.6 and clear_top.mean() >= 0.4    # 16 rows: the 80 % requirement is made at MOT20 scale (test_frame_step_dev_vs_reference)
        assert len(np.unique(ref_p.argmax(1))) >= 3 and (ref_p[:, kslot] > THRESH).any() and not (ref_p[:, kslot] > THRESH).all()
    # the reference-facing call gives the reference's matrix
    pm, reliable = m.associate_embeddings(case.tracks, case.dets, dists, L, C, use_broader_memory=True, select_highest_candidate=False,
                                          extra_kalman_candidates=case.kalman, normalize_ims=True)
    assert np.array_equal(reliable, g["f64_reliable"]) and np.array_equal(pm > 0, g["f64_probs_matrix"] > 0)
    assert np.abs(pm - g["f64_probs_matrix"]).max() < tol["prob"]
    if precision == "fp32":
        pm2, _ = m.associate_embeddings(case.tracks, case.dets, dists, L, C, True, True, extra_kalman_candidates=case.kalman, normalize_ims=True)
        assert np.array_equal(pm2, g["f64_probs_matrix_highest"])


def test_dedup_equals_stacked_batch_conditioned(models):
    """End of the network, conditioned weights: embeddings of a batch with repeated images, duplicate elimination on vs off.  Same
    maths, different summation order of the statistics (and Gram-matrix vs direct statistics for the repeated images): the bf16 rounding
    of activations turns 1e-7 differences of a BatchNorm scale into an occasional different bf16 value, so the embeddings agree to the
    level of the bf16 path's own error, not bit for bit.  The statistics themselves are compared at 1e-5 in
    tests/test_gpu_conv_tc.py::test_statistics_of_a_deduplicated_batch."""
    m = models("bf16")
    eng = m.engine
    rng = np.random.default_rng(11)
    frame = synth.make_frame(3)
    boxes = synth.random_boxes(rng, 9, *frame.shape[:2])
    boxes[:, 2:] += boxes[:, :2]
    crops = m.get_image_crops(frame, boxes, normalize=False)
    base = np.array([m._registry.lookup(c) for c in crops], np.int32)
    slots = np.concatenate([np.repeat(base, [1, 2, 3, 4, 5, 6, 7, 7, 1]), np.full(4, -1, np.int32)])
    rng.shuffle(slots)
    on = eng.reid_embed(slots)
    eng.set_option("dedup", 0)
    try:
        off = eng.reid_embed(slots)
    finally:
        eng.set_option("dedup", 1)
    cos = (on * off).sum(1) / (np.linalg.norm(on, axis=1) * np.linalg.norm(off, axis=1))
    assert cos.min() > 0.9998, cos.min()                       # measured 0.99990: the size of the bf16 path's own distance to fp32
    assert np.abs(on - off).max() / np.abs(off).max() < 5e-2
