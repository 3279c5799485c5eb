// fp64 box arithmetic shared by geometry.cu and rounds.cu: IEEE operations through the *_rn intrinsics (never contracted into FMAs),
// so the results are bit-identical to numpy / scipy.cdist / cython_bbox.
#pragma once
#include <cuda_runtime.h>

namespace {

struct Box { double x1, y1, x2, y2; };

static __device__ __forceinline__ double center_dist(const Box &a, const Box &b) {
    // (tlbr[:2] + tlbr[2:]) / 2.0 ; cdist 'euclidean': s = dx*dx; s += dy*dy; sqrt(s)
    // x / 2.0 == x * 0.5 bit for bit (scaling by a power of two is the same single rounding either way): no software division
    double acx = __dmul_rn(__dadd_rn(a.x1, a.x2), 0.5), acy = __dmul_rn(__dadd_rn(a.y1, a.y2), 0.5);
    double bcx = __dmul_rn(__dadd_rn(b.x1, b.x2), 0.5), bcy = __dmul_rn(__dadd_rn(b.y1, b.y2), 0.5);
    double dx = __dsub_rn(acx, bcx), dy = __dsub_rn(acy, bcy);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

static __device__ __forceinline__ double box_iou(const Box &a, const Box &q) {
    // cython_bbox: +1 pixel convention; 0 unless iw > 0 and ih > 0
    double iw = __dadd_rn(__dsub_rn(fmin(a.x2, q.x2), fmax(a.x1, q.x1)), 1.0);
    if (!(iw > 0.0)) return 0.0;
    double ih = __dadd_rn(__dsub_rn(fmin(a.y2, q.y2), fmax(a.y1, q.y1)), 1.0);
    if (!(ih > 0.0)) return 0.0;
    double qa = __dmul_rn(__dadd_rn(__dsub_rn(q.x2, q.x1), 1.0), __dadd_rn(__dsub_rn(q.y2, q.y1), 1.0));
    double aa = __dmul_rn(__dadd_rn(__dsub_rn(a.x2, a.x1), 1.0), __dadd_rn(__dsub_rn(a.y2, a.y1), 1.0));
    double inter = __dmul_rn(iw, ih);
    double ua = __dsub_rn(__dadd_rn(aa, qa), inter);
    return __ddiv_rn(inter, ua);
}

}  // namespace
