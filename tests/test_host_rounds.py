"""CPU: the oracle restatements of the host-tracker rows of SURVEY.md 8(f) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py coverage / rounds)."""
import os

import numpy as np
import pytest

from oracle import coverage as ocov


def coverage_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "coverage.npz"))
    for k in g["cases"]:
        H, W, scale = g[f"c{k}_meta"]
        yield int(k), int(H), int(W), float(scale), g


def test_coverage_oracle_equals_reference(golden_dir):
    for k, H, W, scale, g in coverage_cases(golden_dir):
        boxes = g[f"c{k}_boxes"] * scale                      # the reference multiplies tlbr by track.scale before int()
        out = ocov.detection_coverage((H, W, 3), boxes)
        ref = g[f"c{k}_scalars"]
        got = np.array([out["area_covered"], out["area_covered_per_obj"], out["max_bbox_area"], out["average_bbox_area"]])
        assert np.array_equal(got, ref), (k, got, ref)      # bit-exact: integer raster + the same fp64 operations
        assert np.array_equal(np.array(out["bbox_areas"], np.float64), g[f"c{k}_areas"])
        rel = [ocov.is_reliable((H, W, 3), boxes, p) for p in g[f"c{k}_p"]]
        assert np.array_equal(np.array(rel, bool), g[f"c{k}_reliable"])


def test_coverage_oracle_equals_cv2_rectangle():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for H, W in [(64, 80), (97, 1031), (300, 1100)]:
        b = np.concatenate([rng.uniform(-30, W + 30, (40, 1)), rng.uniform(-30, H + 30, (40, 1)),
                            rng.uniform(-30, W + 30, (40, 1)), rng.uniform(-30, H + 30, (40, 1))], axis=1)
        canvas = np.zeros((H, W, 3), np.uint8)
        for r in b:
            cv2.rectangle(canvas, (int(r[0]), int(r[1])), (int(r[2]), int(r[3])), (255, 255, 255), thickness=-1)
        assert int(np.count_nonzero(canvas[:, :, 0])) == ocov.detection_coverage((H, W), b)["nonzero"]


# ------------------------------------------------------------------------------------------------ 8f row 1: rounds
from oracle import rounds as ornd


@pytest.fixture(scope="module")
def rounds_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "rounds.npz"))


def test_kalman_oracle_equals_reference(rounds_golden):
    g = rounds_golden
    mean, cov = g["kf_mean0"], g["kf_cov0"]
    for i in range(len(mean)):
        m, c = ornd.kf_initiate(mean[i, :4])
        assert np.array_equal(m, mean[i]) and np.array_equal(c, cov[i])
    for s in range(int(g["kf_steps"])):
        mp, cp = ornd.kf_multi_predict(mean, cov, g[f"kf{s}_tracked"])
        assert np.array_equal(mp, g[f"kf{s}_mean_pred"]) and np.array_equal(cp, g[f"kf{s}_cov_pred"])
        mean, cov = mp.copy(), cp.copy()
        for i in np.where(g[f"kf{s}_upd"])[0]:
            mean[i], cov[i] = ornd.kf_update(mp[i], cp[i], g[f"kf{s}_z"][i])
        assert np.array_equal(mean, g[f"kf{s}_mean_upd"]) and np.array_equal(cov, g[f"kf{s}_cov_upd"])   # same LAPACK, same bits


def test_cost_matrices_oracle_equals_reference(rounds_golden):
    g = rounds_golden
    for k in range(int(g["m_cases"])):
        cost = ornd.iou_distance(g[f"m{k}_a"], g[f"m{k}_b"])
        assert np.array_equal(cost, g[f"m{k}_cost"])
        assert np.array_equal(ornd.fuse_score(cost, g[f"m{k}_score"]), g[f"m{k}_fused"])
    assert ornd.iou_distance(np.zeros((0, 4)), g["m0_b"]).shape == (0, 53)


def test_duplicates_oracle_equals_reference(rounds_golden):
    g = rounds_golden
    for k in range(int(g["d_cases"])):
        da, db = ornd.remove_duplicates(g[f"d{k}_a"], g[f"d{k}_age_a"], g[f"d{k}_b"], g[f"d{k}_age_b"])
        assert np.array_equal(~da, g[f"d{k}_keep_a"]) and np.array_equal(~db, g[f"d{k}_keep_b"])


def test_assignment_oracle_is_the_exhaustive_optimum():
    rng = np.random.default_rng(9)
    for n, m in [(1, 1), (2, 3), (4, 4), (5, 3), (3, 6)]:
        for _ in range(6):
            cost = rng.uniform(0, 1.2, (n, m))
            for thresh in (0.5, 0.9):
                x, y = ornd.linear_assignment(cost, thresh)
                assert abs(ornd.assignment_objective(cost, x, thresh) - ornd.brute_force_assignment(cost, thresh)) < 1e-12
                assert all(cost[i, j] <= thresh for i, j in enumerate(x) if j >= 0)
                assert all((y[j] == i) for i, j in enumerate(x) if j >= 0)


def private_column_assignment(cost, limit):
    """Line-by-line numpy mirror of assignment_kernel (busca_b200/csrc/rounds.cu): N rows, columns 1..M real, M+i private to row i at
    `limit`; shortest augmenting paths with potentials; ties -> lowest column."""
    N, M = cost.shape
    C, INF = M + N, 1e300
    u, v, p, way = np.zeros(N + 1), np.zeros(C + 1), np.zeros(C + 1, int), np.zeros(C + 1, int)
    for r in range(1, N + 1):
        minv, used = np.full(C + 1, INF), np.zeros(C + 1, bool)
        p[0], j0 = r, 0
        while True:
            i0 = p[j0]
            best, bestj = INF, None
            for j in range(1, C + 1):
                if used[j] or j == j0:
                    continue
                c = cost[i0 - 1, j - 1] if j <= M else (limit if j - M == i0 else INF)
                if c < INF:
                    cur = c - u[i0] - v[j]
                    if cur < minv[j]:
                        minv[j], way[j] = cur, j0
                if minv[j] < best:
                    best, bestj = minv[j], j
            used[j0] = True
            for j in range(C + 1):
                if used[j]:
                    u[p[j]] += best
                    v[j] -= best
                else:
                    minv[j] -= best
            j0 = bestj
            if p[j0] == 0:
                break
        while j0:
            j1 = way[j0]
            p[j0] = p[j1]
            j0 = j1
    x = np.full(N, -1, int)
    for j in range(1, M + 1):
        if p[j]:
            x[p[j] - 1] = j - 1
    return x


def test_device_assignment_algorithm_equals_the_extended_problem():
    """The formulation the CUDA kernel solves (rows with private 'unassigned' columns) has the optimum of lapjv's extended matrix."""
    rng = np.random.default_rng(10)
    for n, m in [(1, 1), (3, 5), (12, 7), (25, 25), (40, 18)]:
        for thresh in (0.5, 0.9):
            cost = rng.uniform(0, 1.3, (n, m))
            cost[rng.uniform(size=(n, m)) < 0.5] = 1.0            # IoU distance of non-overlapping boxes
            x = private_column_assignment(cost, thresh)
            xo, _ = ornd.linear_assignment(cost, thresh)
            assert np.array_equal(x, xo), (n, m, thresh)


# ------------------------------------------------------------------------------------------------ 8f row 4: ingest + result format
def test_ingest_oracle_equals_reference(golden_dir):
    import hashlib
    from busca_b200 import synth
    from oracle import ingest as oing
    g = np.load(os.path.join(golden_dir, "ingest.npz"))
    for k in range(int(g["i_cases"])):
        H, W = (int(v) for v in g[f"i{k}_shape"])
        out = oing.denormalize_frame(synth.make_detector_tensor(int(g[f"i{k}_seed"]), H, W), synth.YOLOX_MEANS, synth.YOLOX_STD)
        assert out.shape == (H, W, 3) and out.dtype == np.uint8
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == str(g[f"i{k}_sha"])
        if f"i{k}_bgr" in g:
            assert np.array_equal(out, g[f"i{k}_bgr"])


def test_mot_txt_equals_reference(golden_dir):
    from busca_b200.sharding import write_mot_txt
    from oracle import ingest as oing
    g = np.load(os.path.join(golden_dir, "ingest.npz"))
    rows = g["mot_rows"]
    res = {}
    for r in rows:
        res.setdefault(int(r[0]), ([], [], []))
        res[int(r[0])][0].append(tuple(r[2:6]))
        res[int(r[0])][1].append(int(r[1]))
        res[int(r[0])][2].append(float(r[6]))
    assert oing.mot_lines([(f, *v) for f, v in sorted(res.items())]) == str(g["mot_txt"])
    table = np.concatenate([np.zeros((len(rows), 1)), rows], axis=1)          # sharding's result rows: seq, frame, id, x, y, w, h, score
    assert write_mot_txt(table, 0) == str(g["mot_txt"])


# ------------------------------------------------------------------------------------------------ 8f row 3: camera motion (ECC)
ECC_TOL_T, ECC_TOL_R = 2e-3, 2e-6          # pixels of translation / rotation-matrix entries, against cv2.findTransformECC


def ecc_cases(golden_dir):
    from busca_b200 import synth
    g = np.load(os.path.join(golden_dir, "ecc.npz"))
    for k, (seed, H, W, th, tx, ty) in enumerate(g["cases"]):
        f1 = synth.make_frame(int(seed), int(H), int(W))
        yield k, f1, synth.make_moved_frame(f1, th, tx, ty, int(seed)), g[f"e{k}_warp"], float(g[f"e{k}_cc"]), str(g[f"e{k}_gray_sha"])


def check_warp(k, warp, rho, want, want_rho):
    assert np.abs(warp[:, 2] - want[:, 2]).max() < ECC_TOL_T, (k, warp, want)
    assert np.abs(warp[:, :2] - want[:, :2]).max() < ECC_TOL_R, (k, warp, want)
    assert abs(rho - want_rho) < 1e-5, (k, rho, want_rho)


def test_ecc_oracle_equals_cv2_golden(golden_dir):
    import hashlib
    from oracle import ecc as oecc
    for k, f1, f2, want, want_rho, gray_sha in ecc_cases(golden_dir):
        assert hashlib.sha256(oecc.bgr2gray(f2).tobytes()).hexdigest() == gray_sha          # BGR2GRAY is integer-exact
        if f1.shape[0] * f1.shape[1] > 640 * 480 and k != 0:
            continue                                                                        # one full-HD case keeps the CPU suite short
        rho, warp = oecc.camera_motion(f1, f2)
        check_warp(k, warp, rho, want, want_rho)


def test_ecc_building_blocks_equal_cv2_live():
    cv2 = pytest.importorskip("cv2")
    from busca_b200 import synth
    from oracle import ecc as oecc
    f = synth.make_frame(12, 240, 320)
    g = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(g, oecc.bgr2gray(f))
    b = cv2.GaussianBlur(g.astype(np.float32), (5, 5), 0)
    assert np.array_equal(b, oecc.gaussian5(g.astype(np.float32)))
    gx, gy = oecc.gradients(b)
    assert np.array_equal(gx, cv2.filter2D(b, -1, np.array([[-0.5, 0, 0.5]], np.float32)))
    assert np.array_equal(gy, cv2.filter2D(b, -1, np.array([[-0.5], [0], [0.5]], np.float32)))
    M = np.array([[0.9999, -0.0141, 2.79], [0.0141, 0.9999, -1.44]], np.float32)
    w = cv2.warpAffine(b, M, (320, 240), flags=cv2.INTER_LINEAR + cv2.WARP_INVERSE_MAP)
    assert np.abs(w - oecc.warp_linear([b], M)[0]).max() < 1e-4
    mk = cv2.warpAffine(np.ones((240, 320), np.uint8), M, (320, 240), flags=cv2.INTER_NEAREST + cv2.WARP_INVERSE_MAP)
    assert np.array_equal(mk.astype(bool), oecc.warp_mask(M, 240, 320))
