"""CPU: the oracle (oracle/) against the fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle."""
import hashlib
import os

import numpy as np
import pytest

from busca_b200 import synth
from oracle import crop as ocrop
from oracle import encoding as oenc
from oracle import geometry as ogeo
from oracle import network as onet


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_crops_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "crops.npz"))
    frame = synth.make_frame(int(g["frame_seed"]))
    boxes = g["boxes"]
    for i, b in enumerate(boxes):
        c = ocrop.get_bbox_crop(frame, b)
        assert sha(c) == g["sha64"][i], f"box {i} {b}"
        assert sha(ocrop.crop_direct(frame, b)) == g["sha64"][i], f"direct gather, box {i} {b}"
        assert sha(ocrop.get_bbox_crop(frame, b.astype(np.float32))) == g["sha32"][i]
    for j, i in enumerate(g["full_idx"]):
        assert np.array_equal(ocrop.get_bbox_crop(frame, boxes[i]), g["full"][j])
    e = ocrop.get_image_crops(frame, [])
    assert tuple(e.shape) == tuple(g["empty_shape"]) and str(e.dtype) == str(g["empty_dtype"])


def test_resize_matches_cv2_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    sizes = [(1, 1), (2, 3), (256, 768), (128, 384), (127, 383), (129, 385), (64, 192), (499, 1099)]
    sizes += [(int(rng.integers(1, 500)), int(rng.integers(1, 1100))) for _ in range(40)]
    for sw, sh in sizes:
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        ref = cv2.resize(src, (128, 384), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(ref, ocrop.resize_linear_u8(src)), (sw, sh)


def test_geometry_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "geometry.npz"))
    cd = ogeo.center_distance(g["a"], g["b"])
    assert cd.dtype == np.float64 and np.array_equal(cd, g["center_distance"])
    iou = ogeo.bbox_overlaps(g["a"], g["b"])
    assert np.array_equal(iou, g["iou_ghost"])
    assert iou[11, 11] == 1.0 and (iou > 0).sum() > 10
    m = ogeo.kalman_predict_mean(g["kf_mean_in"], g["kf_tracked"])
    assert np.array_equal(m, g["kf_mean_out"])


def test_iou_known_answers():
    # +1 pixel convention (trades/tracker.py:266-285): a 10x10 box [0,0,9,9] vs itself -> 1;
    # touching boxes share a 1-pixel column.
    a = np.array([[0.0, 0.0, 9.0, 9.0]])
    b = np.array([[0.0, 0.0, 9.0, 9.0], [9.0, 0.0, 18.0, 9.0], [10.0, 0.0, 19.0, 9.0], [5.0, 5.0, 14.0, 14.0]])
    iou = ogeo.bbox_overlaps(a, b)[0]
    assert iou[0] == 1.0
    assert iou[1] == 10.0 / 190.0
    assert iou[2] == 0.0
    assert iou[3] == 25.0 / 175.0
    assert ogeo.bbox_overlaps(np.zeros((0, 4)), b).shape == (0, 4)


def test_pe_tables(golden_dir):
    g = np.load(os.path.join(golden_dir, "pe.npz"))
    tx, ty, tz = oenc.pe_tables(512)
    assert np.array_equal(tx, g["tab_x"]) and np.array_equal(ty, g["tab_y"]) and np.array_equal(tz, g["tab_z"])
    for (i, j, k), v in zip(g["idx"], g["vals"]):
        assert np.array_equal(np.concatenate([tx[i], ty[j], tz[k]]), v)
    # closed form: pe[3,7,11,0:4] = sin 3, cos 3, sin(3 f1), cos(3 f1)
    f1 = 10000 ** (-2 / 172)
    want = np.array([np.sin(3), np.cos(3), np.sin(3 * f1), np.cos(3 * f1)]).astype(np.float16)
    assert np.array_equal(tx[3, :4], want)


CASES = ["assoc_cfg1", "assoc_fewdets", "assoc_nodets"]


@pytest.fixture(scope="module")
def assoc_runs(golden_dir, weights):
    """Oracle run of every golden association case, both sentinel flavours (ReID once per case)."""
    out = {}
    for name in CASES:
        g = np.load(os.path.join(golden_dir, name + ".npz"))
        seed, T, D, L, C, short = (int(v) for v in g["meta"])
        case = synth.make_assoc_case(seed, T, D, L, crop_fn=ocrop.get_image_crops, short_history=short)
        assert np.array_equal(np.array([sha(c) for c in case.tracks[0].images_mem]), g["crop_sha_track0"])
        assert np.array_equal(np.array([sha(k.images_mem[-1]) for k in case.kalman]), g["crop_sha_kalman"])
        dists = ogeo.center_distance([t.tlbr for t in case.tracks], [d.tlbr for d in case.dets])
        taps = {}
        pm, rel = onet.associate(weights, case.tracks, case.dets, dists, L, C, True, kalman=case.kalman,
                                 sentinel_fp64=True, taps=taps)
        # second flavour: reuse the embeddings, redo the Transformer only
        import torch
        mem_img, mem_box, can_img, can_box, idx, n_avail, _ = onet.gather_inputs(
            case.tracks, case.dets, dists, L, C, True, case.kalman)
        taps32 = {}
        lg32 = onet.transformer_forward(weights, torch.from_numpy(taps["mem_emb"]), torch.from_numpy(taps["can_emb"]),
                                        mem_box, can_box, False, taps=taps32)
        taps32["logits"] = lg32.numpy()
        taps32["probs"] = torch.softmax(lg32, -1).numpy()
        pm32 = onet.scatter_probs(taps32["probs"], idx, n_avail, D + T)
        out[name] = dict(g=g, case=case, dists=dists, f64=(pm, rel, taps), f32=(pm32, rel, taps32), idx=idx, n_avail=n_avail)
    return out


@pytest.mark.parametrize("name", CASES)
def test_assoc_indices_bit_exact(assoc_runs, name):
    r = assoc_runs[name]
    g = r["g"]
    if r["dists"].size:
        assert np.array_equal(r["dists"], g["f32_dists"])
    for fl in ("f32", "f64"):
        taps = r[fl][2]
        T = len(r["case"].tracks)
        assert np.array_equal(np.broadcast_to(taps["mem_t"], g[f"{fl}_mem_t"].shape), g[f"{fl}_mem_t"])
        assert np.array_equal(np.broadcast_to(taps["can_t"], g[f"{fl}_can_t"].shape), g[f"{fl}_can_t"])
        for k in ("mem_xy", "mem_size", "can_xy", "can_size"):
            assert np.array_equal(taps[k], g[f"{fl}_{k}"]), (fl, k)
        assert np.array_equal(r[fl][1], g[f"{fl}_reliable"])
        # candidate table: non-zero pattern of the reference's probs_matrix
        ref_cols = [set(np.nonzero(row)[0]) for row in g[f"{fl}_probs_matrix"]]
        mine = [set(int(j) for j in row[: r["n_avail"]]) for row in r["idx"]]
        assert ref_cols == mine


@pytest.mark.parametrize("name", CASES)
def test_assoc_floats_within_tolerance(assoc_runs, name):
    """fp32 tolerance from BASELINE.json north_star: 1e-3 relative on embeddings and scores."""
    r = assoc_runs[name]
    g = r["g"]
    t64 = r["f64"][2]

    def rel(a, b):
        return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

    assert rel(t64["mem_emb"], g["f32_mem_emb"]) < 1e-3
    assert rel(t64["can_emb"], g["f32_can_emb"]) < 1e-3
    for fl in ("f32", "f64"):
        pm, _, taps = r[fl]
        assert rel(taps["cand_rows"], g[f"{fl}_cand_rows"]) < 1e-3
        assert np.abs(taps["logits"] - g[f"{fl}_logits"]).max() < 1e-3 * max(1.0, np.abs(g[f"{fl}_logits"]).max())
        assert np.abs(taps["probs"] - g[f"{fl}_probs"]).max() < 1e-3
        assert np.abs(pm - g[f"{fl}_probs_matrix"]).max() < 1e-3
        assert np.array_equal(taps["probs"].argmax(1), g[f"{fl}_probs"].argmax(1))
        if f"{fl}_input_seq" in g:
            assert rel(taps["input_seq"], g[f"{fl}_input_seq"]) < 1e-3
            assert rel(taps["trans_out"], g[f"{fl}_trans_out"]) < 1e-3


@pytest.mark.parametrize("name", CASES)
def test_assoc_select_highest(assoc_runs, name):
    r = assoc_runs[name]
    g = r["g"]
    T, D = len(r["case"].tracks), len(r["case"].dets)
    probs = r["f64"][2]["probs"]
    a = onet.scatter_probs(probs, r["idx"], r["n_avail"], D + T, select_highest_candidate=True)
    assert np.array_equal(a, g["f64_probs_matrix_highest"])
    b = onet.scatter_probs(probs, r["idx"], r["n_avail"], D + T, select_highest_candidate=True,
                           highest_candidate_minimum_thresh=0.3, keep_highest_value=True)
    assert np.abs(b - g["f64_probs_matrix_highest_keep_thr"]).max() < 1e-3
    assert np.array_equal(b > 0, g["f64_probs_matrix_highest_keep_thr"] > 0)


def test_memory_sampling():
    # network.py:261-269: 11 evenly spaced entries, first and last always included
    assert onet.sample_memory(11, 11, True) == list(range(11))
    s = onet.sample_memory(40, 11, True)
    assert s[0] == 0 and s[-1] == 39 and len(s) == 11 and s == [int(i * 3.9) for i in range(11)]
    assert onet.sample_memory(40, 11, False) == list(range(29, 40))
    assert onet.sample_memory(5, 11, True) == list(range(5))
    assert onet.sample_memory(0, 1, True) == []
