"""Test infrastructure: the CPU oracle dressed as the reference's BUSCA object (get_image_crops / associate_embeddings),
so host-side drivers can be exercised without a GPU.  Never imported by the product."""
import numpy as np

from oracle import crop as ocrop
from oracle import geometry as ogeo
from oracle import network as onet


class OracleBUSCA:
    def __init__(self, weights):
        self.sd = weights
        self.calls = 0

    def get_image_crops(self, image, bboxes, output_size=None, normalize=True):
        assert not normalize
        boxes = [np.asarray(b, dtype=np.float64).reshape(4) for b in bboxes]
        if not boxes:
            return np.zeros([0, 128, 384, 3])
        return ocrop.get_image_crops(image, np.stack(boxes))

    def associate_embeddings(self, tracks_embeddings, dets_embeddings, dists_matrix, seq_len, num_candidates, use_broader_memory,
                             select_highest_candidate, highest_candidate_minimum_thresh=None, keep_highest_value=False,
                             extra_kalman_candidates=(), plot_results=False, normalize_ims=False):
        self.calls += 1
        return onet.associate(self.sd, tracks_embeddings, dets_embeddings, dists_matrix, seq_len, num_candidates, use_broader_memory,
                              select_highest_candidate, highest_candidate_minimum_thresh, keep_highest_value, kalman=extra_kalman_candidates)


def center_distance(tracks, dets):
    return ogeo.center_distance([t.tlbr for t in tracks], [d.tlbr for d in dets])


def iou(a, b):
    return ogeo.bbox_overlaps(a, b)


class OracleRounds:
    """The DeviceRounds protocol of busca_b200/hosts/bytetrack.py on the oracle (numpy / scipy): exercises the driver's batched
    round path on a box without a GPU."""

    def predict(self, mean, cov, tracked):
        from oracle import rounds as ornd
        return ornd.kf_multi_predict(mean, cov, tracked)

    def update(self, mean, cov, xyah):
        from oracle import rounds as ornd
        out = [ornd.kf_update(m, c, z) for m, c, z in zip(mean, cov, xyah)]
        return np.asarray([o[0] for o in out]), np.asarray([o[1] for o in out])

    def match(self, a_tlbr, b_tlbr, scores, thresh):
        from oracle import rounds as ornd
        cost = ornd.iou_distance(a_tlbr, b_tlbr)
        return ornd.linear_assignment(cost if scores is None else ornd.fuse_score(cost, scores), thresh)[0]

    def duplicates(self, a_tlbr, a_age, b_tlbr, b_age):
        from oracle import rounds as ornd
        return ornd.remove_duplicates(a_tlbr, a_age, b_tlbr, b_age)
