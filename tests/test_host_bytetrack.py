"""The from-scratch ByteTrack-with-BUSCA host driver (busca_b200/hosts/bytetrack.py) against tests/golden/adapter_seq.npz, which
tests/golden/make_golden.py recorded from the UNMODIFIED reference adapter (adapters/CenterTrack/src/lib/utils/byte_tracker.py)
on the same seeded synthetic sequence with the conditioned weights: per frame the ids and boxes of the output tracks, and for
every Step-3b call which pooled tracks were kept alive.

CPU (not gpu): the driver on the oracle - pins the driver's control flow to the adapter's.
GPU: the driver on libbusca_b200 (fp32 and bf16) - BASELINE.json north_star: "resulting track IDs must be bit-exact"."""
import os

import numpy as np
import pytest

from busca_b200 import synth
from busca_b200.hosts.bytetrack import ByteTrackHost
from busca_b200.option import load_args_from_config

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = os.path.join(os.path.dirname(HERE), "busca_b200", "configs", "bytetrack_mot20.yml")


CFG17 = os.path.join(os.path.dirname(HERE), "busca_b200", "configs", "bytetrack_mot17.yml")


def tracker_args(config="MOT20"):
    args, _ = load_args_from_config(CFG if config == "MOT20" else CFG17)
    args.use_busca = True
    args.track_thresh, args.track_buffer, args.match_thresh, args.mot20 = 0.6, 30, 0.9, config == "MOT20"
    return args


def replay(busca, iou_fn, cdist_fn, golden, n_frames=None, near_tie=0.0, config="MOT20", box_atol=1e-6, on_frame=None, **host_kw):
    """Run the driver over the golden's sequence; returns the list of frames whose Step-3b pool contained a documented
    near-tie (|p - busca_thresh| < near_tie in the reference) - from the first such frame on, ids may legitimately differ."""
    g = golden
    seed, total, n_obj = (int(v) for v in g["meta"])
    n_frames = n_frames or total
    seq = synth.make_sequence(seed, total, n_obj, miss=float(g["miss"]))
    args = tracker_args(config)
    host = ByteTrackHost(busca, args, iou_fn=iou_fn, center_distance_fn=cdist_fn, **host_kw)
    for f in range(n_frames):
        if on_frame is not None:
            on_frame(f)
        out = host.update(seq.dets[f].copy(), [seq.H, seq.W], [seq.H, seq.W], current_frame=seq.frames[f])
        a, b = int(g["off"][f]), int(g["off"][f + 1])
        want_ids = g["ids"][a:b].tolist()
        b3a, b3b = int(g["b3_off"][f]), int(g["b3_off"][f + 1])
        if b3b > b3a:
            assert host.last_busca is not None, f"frame {f + 1}: the reference ran Step 3b, the driver did not"
            matches, _u, pk, rel = host.last_busca
            want_keep = g["b3_keep"][b3a:b3b]
            want_prob = g["b3_prob"][b3a:b3b]
            mine_keep = np.zeros(b3b - b3a, bool)
            mine_keep[[m[0] for m in matches]] = True
            if getattr(args, "select_highest_candidate", False):
                tie = np.zeros(b3b - b3a, bool)                # one-hot probabilities: the near-tie is between arg-max candidates, checked by the caller's bound
            else:
                tie = np.abs(np.where(want_prob >= 0, want_prob, pk) - args.busca_thresh) < near_tie if near_tie else np.zeros(b3b - b3a, bool)
            if tie.any() and not np.array_equal(mine_keep, want_keep):
                return f + 1                                   # a documented near-tie flipped: stop comparing here
            assert np.array_equal(mine_keep, want_keep), (f + 1, mine_keep, want_keep, pk)
            kept = want_prob >= 0
            if kept.any():
                tol = max(near_tie, 1e-3)
                assert np.abs(pk[kept] - want_prob[kept]).max() < tol, (f + 1, pk[kept], want_prob[kept])
        assert [t.track_id for t in out] == want_ids, (f + 1, [t.track_id for t in out], want_ids)
        if want_ids:
            assert np.allclose(np.array([t.tlwh for t in out]), g["boxes"][a:b], rtol=0, atol=box_atol)
    return None


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "adapter_seq.npz"))


def test_driver_reproduces_the_reference_adapter_on_cpu(golden):
    """22 frames on the oracle (the first Step-3b calls happen from frame 14 on): same ids, same boxes, same kept-alive sets."""
    from oracle_busca import OracleBUSCA, center_distance, iou
    busca = OracleBUSCA(synth.make_weights(0, profile="conditioned"))
    assert replay(busca, iou, center_distance, golden, n_frames=22) is None
    assert busca.calls >= 5


def test_driver_with_batched_rounds_on_cpu(golden):
    """The driver's batched-round path (what DeviceRounds plugs into) on the oracle: same ids as the reference adapter."""
    from oracle_busca import OracleBUSCA, OracleRounds, center_distance, iou
    busca = OracleBUSCA(synth.make_weights(0, profile="conditioned"))
    assert replay(busca, iou, center_distance, golden, n_frames=16, rounds=OracleRounds()) is None


def test_assignment_matches_lapjv_semantics():
    from busca_b200.hosts.bytetrack import assign
    cost = np.array([[0.1, 0.9, 0.8], [0.95, 0.2, 0.97]])
    m, ua, ub = assign(cost, 0.9)
    assert m.tolist() == [[0, 0], [1, 1]] and ua == [] and ub == [2]
    m, ua, ub = assign(np.array([[0.95]]), 0.9)                # above the limit: stays unassigned
    assert len(m) == 0 and ua == [0] and ub == [0]
    m, ua, ub = assign(np.zeros((0, 3)), 0.9)
    assert len(m) == 0 and ua == [] and ub == [0, 1, 2]


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_track_ids_equal_reference_adapter_gpu(golden, precision):
    """All 36 frames through libbusca_b200 with the adapter's call pattern (3 detection-crop calls per frame, one single-box
    crop call per unmatched track, center_distance without an engine handle).  fp32: decisions outside 2e-3 near-ties;
    bf16: outside 3e-2 (the stated bf16 bound, tests/test_gpu_scene.py)."""
    from busca_b200 import tracking
    from busca_b200.network import BUSCA
    args = tracker_args()
    a = args.transformer
    a.device, a.precision = "cuda:0", precision
    m = BUSCA(a).eval()
    m.load_state_dict(synth.make_weights(0, profile="conditioned"))
    launches0 = m.engine.launches
    stop = replay(m, lambda x, y: m.engine.iou(x, y), lambda t, d: tracking.center_distance(t, d), golden,
                  near_tie=2e-3 if precision == "fp32" else 3e-2)
    assert m.engine.launches - launches0 > 1000
    assert stop is None or stop > 20, f"a near-tie flipped already at frame {stop}"


@pytest.mark.gpu
def test_track_ids_with_the_rounds_on_the_device(golden):
    """the same 36 frames with the association rounds themselves on the device (batched Kalman predict / update,
    IoU cost + assignment per round, duplicate removal: csrc/rounds.cu) - ids, boxes (1e-6) and Step-3b decisions as the reference's."""
    from busca_b200 import tracking
    from busca_b200.hosts.bytetrack import DeviceRounds
    from busca_b200.network import BUSCA
    args = tracker_args()
    a = args.transformer
    a.device, a.precision = "cuda:0", "fp32"
    m = BUSCA(a).eval()
    m.load_state_dict(synth.make_weights(0, profile="conditioned"))
    stop = replay(m, lambda x, y: m.engine.iou(x, y), lambda t, d: tracking.center_distance(t, d), golden, near_tie=2e-3,
                  rounds=DeviceRounds(m.engine))
    prof = m.engine.counter("launches") if hasattr(m.engine, "counter") else None
    assert stop is None or stop > 20, f"a near-tie flipped already at frame {stop}"


# ---- the MOT17 configuration: detection-coverage gate + camera-motion compensation + score fusion + select_highest_candidate ----------
@pytest.fixture(scope="module")
def golden17(golden_dir):
    return np.load(os.path.join(golden_dir, "adapter_seq_mot17.npz"))


def test_driver_mot17_config_on_cpu(golden17):
    """config_bytetrack_mot17.yml (reliable_thresh [15, 0.037], use_camera_motion_compensation, select_highest_candidate) on the oracle:
    ids as the unmodified adapter's over the frames where the gate opens and closes.  The warp matrices are the ones cv2 returned in the
    reference run (the oracle's own ECC is pinned separately: tests/test_host_rounds.py)."""
    from oracle import coverage as ocov
    from oracle_busca import OracleBUSCA, OracleRounds, center_distance, iou
    busca = OracleBUSCA(synth.make_weights(0, profile="conditioned"))
    g = golden17
    frame = [0]

    def reliable(shape, tracks, p):
        return ocov.is_reliable(shape, [np.asarray(t.tlbr) * t.scale for t in tracks], p)

    assert replay(busca, iou, center_distance, g, n_frames=16, config="MOT17", reliable_fn=reliable, camera_motion_fn=lambda prev, cur: g["warps"][frame[0]],
                  rounds=OracleRounds(), on_frame=lambda f: frame.__setitem__(0, f)) is None
    assert (g["gate"][:16] == 1).any() and busca.calls >= 2


@pytest.mark.gpu
def test_mot17_config_all_on_the_device(golden17):
    """The same sequence with EVERYTHING on libbusca_b200: BUSCA (fp32), the rounds (DeviceRounds), the coverage gate
    (busca_detection_coverage) and the camera-motion warp (busca_camera_motion) - ids, boxes and Step-3b decisions as the unmodified
    adapter's (cv2 ECC, scipy/LAPACK Kalman, lapjv-semantics assignment)."""
    from busca_b200 import tracking
    from busca_b200.hosts.bytetrack import DeviceRounds
    from busca_b200.network import BUSCA
    args = tracker_args("MOT17")
    a = args.transformer
    a.device, a.precision = "cuda:0", "fp32"
    m = BUSCA(a).eval()
    m.load_state_dict(synth.make_weights(0, profile="conditioned"))
    g = golden17
    worst = [0.0]

    def camera(prev, cur):
        warp, _rho, _it = m.engine.camera_motion(prev, cur)
        return warp

    stop = replay(m, lambda x, y: m.engine.iou(x, y), lambda t, d: tracking.center_distance(t, d), g, near_tie=2e-3, config="MOT17",
                  reliable_fn=lambda shape, tracks, p: tracking.is_reliable(shape, tracks, p, engine=m.engine), camera_motion_fn=camera,
                  rounds=DeviceRounds(m.engine), box_atol=1e-3)     # the device warp is within ~1e-5 px of cv2's, the centres are float32
    assert stop is None, f"decisions diverged at frame {stop}"
