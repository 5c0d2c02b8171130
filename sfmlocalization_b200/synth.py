"""Seeded synthetic inputs for the hot path: random AKAZE/MLDB-like
descriptor rows with planted true matches, and 3D maps with projected, noise-perturbed
observations.  numpy only; shared by tests/ and bench.py so both sides see identical data."""
import numpy as np

ROW = 64
DESC_BITS = 486
# iPhone-6 intrinsics shipped with the reference:
# VisionLocalizeServer/config/camera/iphone6-1920x1080/K.txt:1-3
K_IPHONE6 = np.array([[1861.73, 0.0, 1043.21], [0.0, 1870.67, 644.65], [0.0, 0.0, 1.0]])
IMAGE_WH = (1920, 1080)


def random_rows(n, seed):
    """n x 64 uint8: bits 0..485 i.i.d. Bernoulli(0.5), bits 486..511 zero (61-byte AKAZE
    descriptor zero padded to 64, FileUtils.cpp:77-92)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = rng.integers(0, 256, size=(n, ROW), dtype=np.uint8)
    rows[:, 60] &= 0x3F
    rows[:, 61:] = 0
    return rows


def plant_matches(A, B, seed, frac=0.3, flip_p=0.10):
    """Overwrite a fraction of the rows of A with noisy copies of random rows of B
    (each of the 486 bits flipped with probability flip_p).  Returns (rows, target) with
    target[i] = planted row of B or -1."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nA = A.shape[0]
    target = np.full(nA, -1, np.int64)
    if nA == 0 or B.shape[0] == 0:
        return A, target
    sel = rng.random(nA) < frac
    idx = np.nonzero(sel)[0]
    tgt = rng.integers(0, B.shape[0], size=len(idx))
    target[idx] = tgt
    rows = B[tgt].copy()
    flips = rng.random((len(idx), ROW * 8)) < flip_p
    flips[:, DESC_BITS:] = False
    rows ^= np.packbits(flips, axis=1, bitorder="little")
    A = A.copy()
    A[idx] = rows
    return A, target


def tie_heavy_rows(n, seed, varying_bits=12):
    """Rows in which only `varying_bits` bits differ, so distances tie constantly."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = np.zeros((n, ROW), np.uint8)
    bits = rng.integers(0, 2, size=(n, varying_bits), dtype=np.uint8)
    rows[:, :2] = np.packbits(np.pad(bits, ((0, 0), (0, 16 - varying_bits))), axis=1, bitorder="little")
    return rows


def descriptor_sets(nA, nB, seed, frac=0.3):
    """The standard benchmark pair: database B, searchers A with planted matches."""
    B = random_rows(nB, seed)
    A = random_rows(nA, seed + 7919)
    A, target = plant_matches(A, B, seed + 104729, frac=frac)
    return A, B, target


def image_collection(n_images, rows_per_image, seed, overlap=0.3, jitter=0):
    """Reconstruction-style collection: image k+1 shares `overlap` of its rows (noisy) with
    image k.  Returns (rows, seg_offsets)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    counts = [rows_per_image + (int(rng.integers(-jitter, jitter + 1)) if jitter else 0)
              for _ in range(n_images)]
    segs = []
    prev = None
    for k, c in enumerate(counts):
        rows = random_rows(c, seed * 1000 + k)
        if prev is not None and c > 0 and prev.shape[0] > 0:
            rows, _ = plant_matches(rows, prev, seed * 1000 + 500 + k, frac=overlap)
        segs.append(rows)
        prev = rows
    off = np.zeros(n_images + 1, np.uint64)
    off[1:] = np.cumsum(counts)
    return (np.concatenate(segs, axis=0) if segs else np.zeros((0, ROW), np.uint8)), off


def rodrigues(rv):
    th = np.linalg.norm(rv)
    if th < 1e-12:
        return np.eye(3)
    k = rv / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)


def resection_scene(N, seed, outlier_frac=0.5, noise_px=0.7, K=K_IPHONE6):
    """N 3D points in a 10 x 6 x 10 m box 4..14 m in front of a random camera (rotation <= 30
    deg), projected with K, N(0, noise_px^2) pixel noise, a fraction replaced by uniform
    outliers.  Returns dict(x2d N x 2, X3d N x 3, R, t, inlier_mask, K)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rv = rng.normal(size=3)
    rv *= np.deg2rad(rng.uniform(0, 30)) / np.linalg.norm(rv)
    R = rodrigues(rv)
    t = rng.normal(size=3) * 0.5
    Xc = np.stack([rng.uniform(-5, 5, N), rng.uniform(-3, 3, N), rng.uniform(4, 14, N)], axis=1)
    # keep the points inside the image
    for _ in range(20):
        uvw = Xc @ K.T
        uv = uvw[:, :2] / uvw[:, 2:]
        bad = (uv[:, 0] < 0) | (uv[:, 0] >= IMAGE_WH[0]) | (uv[:, 1] < 0) | (uv[:, 1] >= IMAGE_WH[1])
        if not bad.any():
            break
        nb = int(bad.sum())
        Xc[bad] = np.stack([rng.uniform(-5, 5, nb), rng.uniform(-3, 3, nb), rng.uniform(4, 14, nb)], axis=1)
    X = (Xc - t) @ R          # X_world = R^T (X_cam - t)
    uvw = Xc @ K.T
    x = uvw[:, :2] / uvw[:, 2:] + rng.normal(scale=noise_px, size=(N, 2))
    out = rng.random(N) < outlier_frac
    n_out = int(out.sum())
    x[out] = np.stack([rng.uniform(0, IMAGE_WH[0], n_out), rng.uniform(0, IMAGE_WH[1], n_out)], axis=1)
    return dict(x2d=x, X3d=X, R=R, t=t, inlier_mask=~out, K=K.copy())


def sample_triplets(N, T, seed):
    """T triplets of distinct correspondence indices."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tri = np.empty((T, 3), np.uint32)
    for k in range(T):
        tri[k] = rng.choice(N, size=3, replace=False)
    return tri


def _noisy_copies(base, seed, flip_p):
    rng = np.random.Generator(np.random.PCG64(seed))
    flips = rng.random((base.shape[0], ROW * 8)) < flip_p
    flips[:, DESC_BITS:] = False
    return base ^ np.packbits(flips, axis=1, bitorder="little")


def localization_scene(n_views, feats_per_view, n_landmarks, nq, seed, track_frac=0.6, query_inlier_frac=0.35,
                       noise_px=0.7, flip_p=0.08, K=K_IPHONE6, window=None):
    """A synthetic SfM map and one query image.
    Map: n_landmarks 3D points, each with a base descriptor; every view holds feats_per_view
    features of which track_frac observe a random landmark (descriptor = noisy copy of the
    landmark's) and the rest are clutter.  Query: nq features, query_inlier_frac of them observe
    landmarks (projected with the true pose + pixel noise), the rest are clutter at random
    positions.  window: when given, every view observes landmarks from a window of that many
    consecutive landmark ids (views march through the id range) and the query from the window in
    the middle -- the co-visibility structure of a real map, where a query overlaps a few dozen
    views strongly instead of all views weakly.
    Returns a dict with everything hulo_engine_create / localize need plus truth."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sc = resection_scene(n_landmarks, seed + 1, outlier_frac=0.0, noise_px=0.0, K=K)
    X, R, t = sc["X3d"], sc["R"], sc["t"]
    lm_desc = random_rows(n_landmarks, seed + 2)
    rows, off = [], [0]
    obs_view, obs_feat, obs_lm = [], [], []
    # keypoint positions of the map features (input of the F-matrix geometric filter): every view
    # has its own pose near the query's; drawn from a separate stream so the fields above and below
    # are unchanged by their presence
    rng_xy = np.random.Generator(np.random.PCG64(seed + 5))
    map_xy, view_R, view_t = [], [], []
    for v in range(n_views):
        n_obs = int(feats_per_view * track_frac)
        if window:
            w0 = int((n_landmarks - window) * v / max(1, n_views - 1))
            lms = w0 + rng.choice(window, size=min(n_obs, window), replace=False)
        else:
            lms = rng.choice(n_landmarks, size=min(n_obs, n_landmarks), replace=False)
        d_obs = _noisy_copies(lm_desc[lms], seed * 7919 + v, flip_p)
        d_clutter = random_rows(feats_per_view - len(lms), seed * 104729 + v)
        d = np.concatenate([d_obs, d_clutter], axis=0)
        perm = rng.permutation(d.shape[0])
        d = d[perm]
        inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
        rows.append(d); off.append(off[-1] + d.shape[0])
        obs_view += [v] * len(lms); obs_feat += inv[:len(lms)].tolist(); obs_lm += lms.tolist()
        Rv = rodrigues(rng_xy.normal(size=3) * np.deg2rad(4.0)) @ R
        tv = t + rng_xy.normal(size=3) * 0.4
        Xv = X[lms] @ Rv.T + tv
        uvv = Xv @ K.T
        xy_obs = uvv[:, :2] / uvv[:, 2:] + rng_xy.normal(scale=noise_px, size=(len(lms), 2))
        n_cl = feats_per_view - len(lms)
        xy = np.concatenate([xy_obs, np.stack([rng_xy.uniform(0, IMAGE_WH[0], n_cl),
                                               rng_xy.uniform(0, IMAGE_WH[1], n_cl)], axis=1)])
        map_xy.append(xy[perm]); view_R.append(Rv); view_t.append(tv)
    n_in = int(nq * query_inlier_frac)
    if window:
        q_lms = (n_landmarks - window) // 2 + rng.choice(window, size=min(n_in, window), replace=False)
    else:
        q_lms = rng.choice(n_landmarks, size=min(n_in, n_landmarks), replace=False)
    q_desc = np.concatenate([_noisy_copies(lm_desc[q_lms], seed + 3, flip_p), random_rows(nq - len(q_lms), seed + 4)])
    Xc = X[q_lms] @ R.T + t
    uv = (Xc @ K.T); uv = uv[:, :2] / uv[:, 2:]
    uv += rng.normal(scale=noise_px, size=uv.shape)
    q_xy = np.concatenate([uv, np.stack([rng.uniform(0, IMAGE_WH[0], nq - len(q_lms)),
                                         rng.uniform(0, IMAGE_WH[1], nq - len(q_lms))], axis=1)])
    perm = rng.permutation(nq)
    q_desc, q_xy = q_desc[perm], q_xy[perm]
    q_truth = np.full(nq, -1, np.int64)
    q_truth[np.argsort(perm)[:len(q_lms)]] = q_lms
    return dict(rows=np.concatenate(rows), seg_offsets=np.array(off, np.uint64),
                obs_view=np.array(obs_view, np.uint32), obs_feat=np.array(obs_feat, np.uint32),
                obs_landmark=np.array(obs_lm, np.uint32), landmark_X=X, K=K.copy(), R=R, t=t,
                center=-R.T @ t, q_desc=q_desc, q_xy=q_xy, q_truth=q_truth, lm_desc=lm_desc, window=window,
                map_xy=np.concatenate(map_xy), view_wh=np.tile(np.array(IMAGE_WH, np.int32), (n_views, 1)),
                view_R=np.array(view_R), view_t=np.array(view_t))


def extra_query(scene, nq, seed, query_inlier_frac=0.35, noise_px=0.7, flip_p=0.08):
    """Another query image of the same map from a nearby pose (a few degrees / decimetres away):
    the concurrent requests of the batched server mode.  Returns dict(q_desc, q_xy, R, t, center)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    K, X = scene["K"], scene["landmark_X"]
    dR = rodrigues(rng.normal(size=3) * np.deg2rad(3.0))
    R = dR @ scene["R"]
    t = dR @ scene["t"] + rng.normal(size=3) * 0.15
    Xc = X @ R.T + t
    uv = Xc @ K.T
    uv = uv[:, :2] / uv[:, 2:]
    ok = (Xc[:, 2] > 1.0) & (uv[:, 0] >= 0) & (uv[:, 0] < IMAGE_WH[0]) & (uv[:, 1] >= 0) & (uv[:, 1] < IMAGE_WH[1])
    if scene.get("window"):
        w0 = (len(X) - scene["window"]) // 2 + int(rng.integers(-scene["window"] // 4, scene["window"] // 4 + 1))
        inwin = np.zeros(len(X), bool)
        inwin[max(0, w0):w0 + scene["window"]] = True
        ok &= inwin
    vis = np.nonzero(ok)[0]
    n_in = min(int(nq * query_inlier_frac), len(vis))
    q_lms = rng.choice(vis, size=n_in, replace=False)
    q_desc = np.concatenate([_noisy_copies(scene["lm_desc"][q_lms], seed + 3, flip_p), random_rows(nq - n_in, seed + 4)])
    xy = uv[q_lms] + rng.normal(scale=noise_px, size=(n_in, 2))
    q_xy = np.concatenate([xy, np.stack([rng.uniform(0, IMAGE_WH[0], nq - n_in),
                                         rng.uniform(0, IMAGE_WH[1], nq - n_in)], axis=1)])
    perm = rng.permutation(nq)
    return dict(q_desc=q_desc[perm], q_xy=q_xy[perm], R=R, t=t, center=-R.T @ t)



def two_view_matches(N, seed, outlier_frac=0.4, noise_px=0.5, K=K_IPHONE6):
    """Putative matches between two views of a 3D scene (the input of the F-matrix geometric
    filter): N point pairs in pixels, a fraction replaced by uniform outliers in image J.
    Returns dict(xI, xJ N x 2, inlier_mask, F_true with xJ^T F xI = 0, size (w, h))."""
    rng = np.random.Generator(np.random.PCG64(seed))
    X = np.stack([rng.uniform(-4, 4, N), rng.uniform(-2.5, 2.5, N), rng.uniform(5, 14, N)], axis=1)
    R = rodrigues(rng.normal(size=3) * np.deg2rad(6.0))
    t = np.array([rng.uniform(0.3, 0.8), rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2)])
    uI = X @ K.T; xI = uI[:, :2] / uI[:, 2:]
    XJ = X @ R.T + t
    uJ = XJ @ K.T; xJ = uJ[:, :2] / uJ[:, 2:]
    xI = xI + rng.normal(scale=noise_px, size=xI.shape)
    xJ = xJ + rng.normal(scale=noise_px, size=xJ.shape)
    out = rng.random(N) < outlier_frac
    n_out = int(out.sum())
    xJ[out] = np.stack([rng.uniform(0, IMAGE_WH[0], n_out), rng.uniform(0, IMAGE_WH[1], n_out)], axis=1)
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Ki = np.linalg.inv(K)
    F = Ki.T @ tx @ R @ Ki
    return dict(xI=xI, xJ=xJ, inlier_mask=~out, F_true=F / np.linalg.norm(F), size=IMAGE_WH)
