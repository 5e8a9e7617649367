"""CPU checker for the pose-initialisation path (include/mcba_pnp.h). TEST INFRASTRUCTURE: only tests/, smoke()
and the cpu_baseline leg of tools/bench_pnp.py may import it; the product never does.

It restates ransacP3PDistortion / ransacP3P of the reference (/root/reference/McCalib/src/geometrytools.cpp:709-751,
:233-326) line by line ON TOP OF OpenCV ITSELF (cv2 4.13 is in this image): the reference's own third-party calls —
cv::solvePnP(SOLVEPNP_P3P), cv::projectPoints, cv::fisheye::undistortPoints, cv::solvePnP(SOLVEPNP_ITERATIVE) — are
made through cv2 with the reference's argument conventions (float32 points, fp32 re-projections). Parity of this path
is therefore pinned by the library the reference calls, not by a re-derivation.

Two things are ours, because the reference leaves them undefined (see include/mcba_pnp.h):
  - the draws: the reference re-seeds std::mt19937 from std::random_device at every draw; here draw h of observation b
    is `sample4(seed, b, h, n)` (splitmix64), the same function the CUDA kernel evaluates;
  - a draw whose first three points are collinear, or for which cv::solvePnP reports failure, scores 0 inliers.
"""
import numpy as np

M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def sample4(seed, b, h, n):
    st = splitmix64(seed ^ splitmix64((b * 0x100000001B3 + h) & M64))
    srt, out = [], []
    for k in range(4):
        st = splitmix64(st)
        r = st % (n - k)
        pos = 0
        while pos < k and r >= srt[pos]:
            r += 1
            pos += 1
        srt.insert(pos, r)
        out.append(r)
    return out


def degenerate(P3):
    d12, d13, d23 = P3[0] - P3[1], P3[0] - P3[2], P3[1] - P3[2]
    a12, a13, a23 = d12 @ d12, d13 @ d13, d23 @ d23
    nx = np.cross(d12, d13)
    return not (nx @ nx > 1e-12 * a12 * a13) or not (a23 > 0.0)


def ransac_p3p(scene, image, K, dist, thresh, it, p, refine, seed, b):
    """geometrytools.cpp:233-326. scene [n,3] float32, image [n,2] float32. Returns dict."""
    import cv2
    n = len(scene)
    N, trial, countit, best = it, 0, 0, 0
    inliers = np.zeros(0, int)
    best_r = best_t = None
    best_h = -1
    eps = 0.00001
    if n >= 4:
        while N > trial and countit < it:
            idx = sample4(seed, b, countit, n)
            P4, I4 = scene[idx], image[idx]
            index = np.zeros(0, int)
            ok = False
            if not degenerate(P4[:3].astype(np.float64)):
                ok, rv, tv = cv2.solvePnP(P4, I4, K, dist, flags=cv2.SOLVEPNP_P3P)        # :266-267
            if ok:
                proj, _ = cv2.projectPoints(scene, rv, tv, K, dist)                       # :270-271
                proj = proj.reshape(-1, 2).astype(np.float32)                             # std::vector<cv::Point2f>
                d = np.sqrt((image[:, 0].astype(np.float64) - proj[:, 0]) ** 2 + (image[:, 1].astype(np.float64) - proj[:, 1]) ** 2)
                index = np.nonzero(d < thresh)[0]                                         # :276-283
            trial += 1
            if len(index) > best:                                                         # :287-303
                inliers, best, best_h = index, len(index), countit
                best_r, best_t = rv.copy(), tv.copy()
                frac = best / float(n)
                pno = 1 - frac ** 3
                if pno == 0:
                    pno = eps
                if pno > 1 - eps:
                    pno = 1 - eps
                N = int(np.floor(np.log(1 - p) / np.log(pno) + 0.5))
                trial = 0
            countit += 1
    out = dict(best_h=best_h, n_hyp=countit, inliers=inliers, pose0=np.zeros(6), pose=np.zeros(6))
    if best_h >= 0:
        out["pose0"] = np.concatenate([best_r.ravel(), best_t.ravel()])
        out["pose"] = out["pose0"].copy()
        if refine and best >= 4:                                                          # :314-324
            ok, r2, t2 = cv2.solvePnP(scene[inliers], image[inliers], K, dist, best_r.copy(), best_t.copy(), True,
                                      cv2.SOLVEPNP_ITERATIVE)
            out["pose"] = np.concatenate([r2.ravel(), t2.ravel()])
    return out


def ransac_p3p_distortion(scene, image, intr, model, thresh, it, p=0.99, refine=True, seed=0, b=0):
    """geometrytools.cpp:709-751. intr: the 9-vector of Camera::intrinsics_."""
    import cv2
    K = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1.0]])
    scene = np.ascontiguousarray(scene, np.float32)
    image = np.ascontiguousarray(image, np.float32)
    if model == 0:
        return ransac_p3p(scene, image, K, np.asarray(intr[4:9], np.float64), thresh, it, p, refine, seed, b)
    und = cv2.fisheye.undistortPoints(image.reshape(-1, 1, 2), K, np.asarray(intr[4:8], np.float64)).reshape(-1, 2)
    und = und.astype(np.float32)
    und[:, 0] = und[:, 0] * np.float32(intr[0]) + np.float32(intr[2])                     # :733-740, float arithmetic
    und[:, 1] = und[:, 1] * np.float32(intr[1]) + np.float32(intr[3])
    return ransac_p3p(scene, und, K, np.zeros(5), thresh, it, p, refine, seed, b)


def ransac_all(problem, opt):
    """Every observation run of a mcba.pnp.PnpProblem; same outputs as mcba_pnp_ransac."""
    nb = problem.n_bobs
    res = dict(pose=np.zeros((nb, 6)), pose0=np.zeros((nb, 6)), n_inliers=np.zeros(nb, np.int32),
               inlier_mask=np.zeros(problem.n_obs, np.uint8), best_h=np.zeros(nb, np.int32), n_hyp=np.zeros(nb, np.int32))
    for b in range(nb):
        o0, o1 = int(problem.bobs_ptr[b]), int(problem.bobs_ptr[b + 1])
        cam = int(problem.bobs_cam[b])
        scene = problem.pts_xyz[problem.pt_idx[o0:o1]]
        image = np.stack([problem.u[o0:o1], problem.v[o0:o1]], 1)
        r = ransac_p3p_distortion(scene, image, problem.intrinsics[cam], int(problem.cam_model[cam]), opt.threshold_px,
                                  opt.max_iterations, opt.confidence, bool(opt.refine), int(opt.seed), b)
        res["pose"][b], res["pose0"][b] = r["pose"], r["pose0"]
        res["n_inliers"][b], res["best_h"][b], res["n_hyp"][b] = len(r["inliers"]), r["best_h"], r["n_hyp"]
        res["inlier_mask"][o0 + r["inliers"]] = 1
    return res
