"""CPU restatement of the marker detector on the step before the hot path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this file; the product path (ar_slam_b200/csrc/detect.cu behind arslam_detect_markers) never does.

Reference call sites: cv::aruco::detectMarkers at ar_slam/src/aruco_detector.cpp:106 and
ar_slam/src/ar_slam_util.cpp:268 (DICT_4X4_50, DetectorParameters defaults, minCornerDistanceRate = 0.1
in loadImages :249-250).  The arithmetic lives in OpenCV (the reference builds against Ubuntu 22.04's
libopencv-dev 4.5.4; this image has the cv2 4.13.0 wheel), which is not under /root/reference.  This file
restates the published algorithm stage by stage:

  to_gray             cv::cvtColor BGR2GRAY, 8-bit fixed point (R 4899, G 9617, B 1868, >> 14)
  adaptive_threshold  cv::adaptiveThreshold MEAN_C / BINARY_INV: box mean with replicated borders, rounded
  find_contours       cv::findContours RETR_LIST / CHAIN_APPROX_NONE: Suzuki-Abe border following
  approx_poly_dp      cv::approxPolyDP, closed curves (Douglas-Peucker with OpenCV's start search and clean-up)
  is_contour_convex   cv::isContourConvex
  find_marker_contours, reorder_corners, filter_too_close, extract_bits, identify, detect_markers
                      aruco_detector.cpp of OpenCV's objdetect module (_findMarkerContours ... detectMarkers)

PARITY IS PINNED: cv2 is importable in the build container and on the GPU box, and tests/test_oracle_detect.py
checks every stage against the cv2 function it restates and the whole pipeline against
cv2.aruco.ArucoDetector.detectMarkers on the reference's own demo images (golden corners committed under
tests/golden/demo_marker_corners.json) and on rendered scenes.
"""
import numpy as np

# cv::aruco::DetectorParameters defaults that the path uses (OpenCV 4.x), with the reference's one override
DEFAULTS = dict(
    adaptiveThreshWinSizeMin=3, adaptiveThreshWinSizeMax=23, adaptiveThreshWinSizeStep=10,
    adaptiveThreshConstant=7.0, minMarkerPerimeterRate=0.03, maxMarkerPerimeterRate=4.0,
    polygonalApproxAccuracyRate=0.03, minCornerDistanceRate=0.05, minDistanceToBorder=3,
    minMarkerDistanceRate=0.125, minGroupDistance=float(np.float32(0.21)), markerBorderBits=1,
    perspectiveRemovePixelPerCell=4, perspectiveRemoveIgnoredMarginPerCell=0.13,
    maxErroneousBitsInBorderRate=0.35, minOtsuStdDev=5.0, errorCorrectionRate=0.6)
REFERENCE_PARAMS = dict(DEFAULTS, minCornerDistanceRate=0.1)      # ar_slam_util.cpp:250

# Freeman codes of cv::findContours: 0 = east, counter-clockwise on the screen (y grows downwards)
DX = (1, 1, 0, -1, -1, -1, 0, 1)
DY = (0, -1, -1, -1, 0, 1, 1, 1)


def to_gray(bgr):
    """cv::cvtColor(BGR2GRAY) for 8-bit images."""
    if bgr.ndim == 2:
        return bgr.copy()
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    return ((b * 1868 + g * 9617 + r * 4899 + (1 << 13)) >> 14).astype(np.uint8)


def box_mean(gray, win):
    """cv::boxFilter(normalize = true, BORDER_REPLICATE) on 8-bit data; win*win is odd, so rounding has no ties."""
    r = win // 2
    p = np.pad(gray.astype(np.int64), r, mode="edge")
    ii = np.zeros((p.shape[0] + 1, p.shape[1] + 1), np.int64)
    ii[1:, 1:] = p.cumsum(0).cumsum(1)
    h, w = gray.shape
    s = ii[win:win + h, win:win + w] - ii[:h, win:win + w] - ii[win:win + h, :w] + ii[:h, :w]
    return (2 * s + win * win) // (2 * win * win)


def adaptive_threshold(gray, win, c):
    """cv::adaptiveThreshold(gray, 255, MEAN_C, THRESH_BINARY_INV, win, c): 255 where gray <= mean - floor(c)."""
    idelta = int(np.floor(c))
    return np.where(gray.astype(np.int64) - box_mean(gray, win) <= -idelta, 255, 0).astype(np.uint8)


def find_contours(binary):
    """cv::findContours(RETR_LIST, CHAIN_APPROX_NONE): list of (n, 2) int32 arrays of (x, y), in cv2's order
    (the border found last in the raster scan comes first).  The image counts as surrounded by zeros."""
    h, w = binary.shape
    img = np.zeros((h + 2, w + 2), np.int8)
    img[1:-1, 1:-1] = binary != 0
    nz = img != 0
    # raster-ordered places where a border can start: a 0 -> 1 step (outer border at x) or a 1 -> 0 step (hole at x - 1)
    step = nz[:, 1:] != nz[:, :-1]
    ys, xs = np.nonzero(step)
    xs = xs + 1
    out = []
    flat = img            # marks: 1 = untouched, 2 = on a traced border, -2 = traced and its east neighbour was examined
    for y, x in zip(ys.tolist(), xs.tolist()):
        p = int(flat[y, x])
        if p != 0:
            if p != 1:
                continue                    # outer border only starts at an untouched pixel
            is_hole, x0 = False, x
        else:
            if flat[y, x - 1] < 1:
                continue                    # the pixel left of the gap already has its east side accounted for
            is_hole, x0 = True, x - 1
        out.append(_trace(flat, x0, y, is_hole))
    out.reverse()
    return out


def _trace(img, x0, y0, is_hole):
    """Suzuki-Abe border following as cv::findContours does it: start search clockwise from west (outer) or east
    (hole), then counter-clockwise from the pixel after the one we came from.  Returns image coordinates."""
    pts = []
    s_end = s = 0 if is_hole else 4
    while True:
        s = (s - 1) & 7
        if img[y0 + DY[s], x0 + DX[s]] != 0:
            break
        if s == s_end:
            break
    if img[y0 + DY[s], x0 + DX[s]] == 0:          # isolated pixel
        img[y0, x0] = -2
        return np.array([[x0 - 1, y0 - 1]], np.int32)
    x1, y1 = x0 + DX[s], y0 + DY[s]
    x3, y3 = x0, y0
    while True:
        s_end = s
        while True:
            s += 1
            x4, y4 = x3 + DX[s & 7], y3 + DY[s & 7]
            if img[y4, x4] != 0:
                break
        s &= 7
        if s != 0 and s - 1 < s_end:                # the search passed east: that neighbour is a zero of this border
            img[y3, x3] = -2
        elif img[y3, x3] == 1:
            img[y3, x3] = 2
        pts.append((x3 - 1, y3 - 1))
        if x4 == x0 and y4 == y0 and x3 == x1 and y3 == y1:
            break
        x3, y3 = x4, y4
        s = (s + 4) & 7
    return np.array(pts, np.int32)


def approx_poly_dp(contour, eps, closed=True):
    """cv::approxPolyDP for a closed integer contour ((n, 2) array); returns the kept points in order.
    The split criterion is the squared distance to the chord as a SEGMENT (what cv2 4.13 does; OpenCV 4.5.4 measured
    to the infinite line, which differs only on contours with one-pixel spikes) -- pinned against cv2.approxPolyDP."""
    assert closed
    src = [(int(a), int(b)) for a, b in np.asarray(contour).reshape(-1, 2)]
    count = len(src)
    if count == 0:
        return np.zeros((0, 2), np.int32)
    eps2 = float(eps) * float(eps)
    dst = []
    stack = []
    pos = 0
    right_start = 0
    le_eps = False
    start_pt = (-1000000, -1000000)
    # 1. approximately the two farthest points: three sweeps, each from the farthest point of the one before
    for _ in range(3):
        max_dist = 0.0
        pos = (pos + right_start) % count
        start_pt = src[pos]
        pos = (pos + 1) % count
        for j in range(1, count):
            pt = src[pos]
            pos = (pos + 1) % count
            dx = float(pt[0] - start_pt[0])
            dy = float(pt[1] - start_pt[1])
            dist = dx * dx + dy * dy
            if dist > max_dist:
                max_dist = dist
                right_start = j
        le_eps = max_dist <= eps2
    # 2. the two halves
    if not le_eps:
        slice_start = pos % count
        right_end = slice_start
        slice_end = right_start = (right_start + slice_start) % count
        stack.append((right_start, right_end))
        stack.append((slice_start, slice_end))
    else:
        dst.append(start_pt)
    # 3. recursive splitting
    while stack:
        s_start, s_end = stack.pop()
        end_pt = src[s_end]
        pos = s_start
        start_pt = src[pos]
        pos = (pos + 1) % count
        if pos != s_end:
            dx = float(end_pt[0] - start_pt[0])
            dy = float(end_pt[1] - start_pt[1])
            len2 = dx * dx + dy * dy
            max_dist = 0.0
            r_start = 0
            while pos != s_end:
                pt = src[pos]
                pos = (pos + 1) % count
                px = float(pt[0] - start_pt[0])
                py = float(pt[1] - start_pt[1])
                proj = px * dx + py * dy
                if proj < 0 or len2 == 0:         # before the chord's start (or a chord of no length): distance to the start
                    dist = px * px + py * py
                elif proj > len2:                 # beyond its end
                    ex = float(pt[0] - end_pt[0])
                    ey = float(pt[1] - end_pt[1])
                    dist = ex * ex + ey * ey
                else:
                    cross = py * dx - px * dy
                    dist = cross * cross / len2
                if dist > max_dist:
                    max_dist = dist
                    r_start = (pos + count - 1) % count
            le_eps = max_dist <= eps2
        else:
            le_eps = True
            start_pt = src[s_start]
        if le_eps:
            dst.append(start_pt)
        else:
            stack.append((r_start, s_end))
            stack.append((s_start, r_start))
    # 4. clean-up: drop points on [almost] straight lines
    count = new_count = len(dst)
    pos = count - 1
    start_pt = dst[pos]
    pos = (pos + 1) % count
    wpos = pos
    pt = dst[pos]
    pos = (pos + 1) % count
    i = 0
    while i < count and new_count > 2:
        end_pt = dst[pos]
        pos = (pos + 1) % count
        dx = float(end_pt[0] - start_pt[0])
        dy = float(end_pt[1] - start_pt[1])
        dist = abs((pt[0] - start_pt[0]) * dy - (pt[1] - start_pt[1]) * dx)
        inner = (pt[0] - start_pt[0]) * (end_pt[0] - pt[0]) + (pt[1] - start_pt[1]) * (end_pt[1] - pt[1])
        if dist * dist <= 0.5 * eps2 * (dx * dx + dy * dy) and dx != 0 and dy != 0 and inner >= 0:
            new_count -= 1
            dst[wpos] = start_pt = end_pt
            wpos = (wpos + 1) % count
            pt = dst[pos]
            pos = (pos + 1) % count
            i += 2
            continue
        dst[wpos] = start_pt = pt
        wpos = (wpos + 1) % count
        pt = end_pt
        i += 1
    return np.array(dst[:new_count], np.int32).reshape(-1, 2)


def is_contour_convex(pts):
    """cv::isContourConvex on integer points."""
    p = [(int(a), int(b)) for a, b in np.asarray(pts).reshape(-1, 2)]
    n = len(p)
    prev_pt = p[(n - 2) % n]
    cur_pt = p[n - 1]
    dx0 = cur_pt[0] - prev_pt[0]
    dy0 = cur_pt[1] - prev_pt[1]
    orientation = 0
    for i in range(n):
        prev_pt = cur_pt
        cur_pt = p[i]
        dx = cur_pt[0] - prev_pt[0]
        dy = cur_pt[1] - prev_pt[1]
        dxdy0 = dx * dy0
        dydx0 = dy * dx0
        orientation |= 1 if dydx0 > dxdy0 else (2 if dydx0 < dxdy0 else 3)
        if orientation == 3:
            return False
        dx0, dy0 = dx, dy
    return True


def find_marker_contours(thresh, params, contours=None):
    """_findMarkerContours: quads (4 x 2 float32, in approxPolyDP's order), their contours and the 'too near the image
    border' flags, in contour order.  cv2 4.13 keeps a quad that is too near the border until the grouping of
    filter_too_close is done (observed: such a quad still absorbs the quads it encloses), so it is flagged, not dropped."""
    rows, cols = thresh.shape
    big = max(rows, cols)
    min_perimeter = int(params["minMarkerPerimeterRate"] * big)         # unsigned int cast in OpenCV
    max_perimeter = int(params["maxMarkerPerimeterRate"] * big)
    if contours is None:
        contours = find_contours(thresh)
    quads, kept, flags = [], [], []
    for c in contours:
        n = len(c)
        if n < min_perimeter or n > max_perimeter:
            continue
        approx = approx_poly_dp(c, float(n) * params["polygonalApproxAccuracyRate"])
        if len(approx) != 4 or not is_contour_convex(approx):
            continue
        min_dist_sq = float(big) * float(big)
        for j in range(4):
            d = approx[j].astype(np.int64) - approx[(j + 1) % 4].astype(np.int64)
            min_dist_sq = min(min_dist_sq, float(d[0] * d[0] + d[1] * d[1]))
        min_corner = float(n) * params["minCornerDistanceRate"]
        if min_dist_sq < min_corner * min_corner:
            continue
        b = params["minDistanceToBorder"]
        near = bool((approx[:, 0] < b).any() or (approx[:, 1] < b).any() or
                    (approx[:, 0] > cols - 1 - b).any() or (approx[:, 1] > rows - 1 - b).any())
        quads.append(approx.astype(np.float32))
        kept.append(c)
        flags.append(near)
    return quads, kept, flags


def reorder_corners(q):
    """_reorderCandidatesCorners: clockwise on the screen; swaps corners 1 and 3 when the cross product is negative."""
    q = q.copy()
    dx1, dy1 = q[1] - q[0]
    dx2, dy2 = q[2] - q[0]
    if float(dx1) * float(dy2) - float(dy1) * float(dx2) < 0.0:
        q[[1, 3]] = q[[3, 1]]
    return q


def _perimeter(q):
    d = q - np.roll(q, -1, axis=0)
    return float(np.sqrt((d.astype(np.float64) ** 2).sum(1)).sum())


def _average_distance(a, b):
    """getAverageDistance: smallest mean corner-to-corner distance over the four cyclic alignments (float32)."""
    best = np.float32(np.finfo(np.float32).max)
    for first in range(4):
        dist = np.float32(0)
        for c in range(4):
            d = a[c] - b[(first + c) % 4]
            dist = np.float32(dist + np.sqrt(np.float32(d[0] * d[0] + d[1] * d[1])))
        dist = np.float32(dist / np.float32(4))
        best = min(best, dist)
    return best


def _average_module_size(q, marker_size, border_bits):
    s = np.float32(0)
    for i in range(4):
        d = q[i] - q[(i + 1) % 4]
        s = np.float32(s + np.sqrt(np.float32(d[0] * d[0] + d[1] * d[1])))
    n = marker_size + 2 * border_bits
    return np.float32(s / np.float32(4 * n))


def filter_too_close(quads, params, marker_size=4, near_border=None):
    """_filterTooCloseCandidates with detectInvertedMarker = false: candidates sorted big to small (stable), near
    duplicates grouped, the largest of a group kept with its 'close contours' as fall-backs for identification.
    A group whose largest member is too near the image border is dropped as a whole; flagged fall-backs are skipped.
    Returns a list of (quad, [fall-back quads])."""
    if near_border is None:
        near_border = [False] * len(quads)
    order = sorted(range(len(quads)), key=lambda i: -np.float32(_perimeter32(quads[i])))
    cand = [quads[i] for i in order]
    flag = [near_border[i] for i in order]
    per = [_perimeter32(q) for q in cand]
    n = len(cand)
    group = [-1] * n
    groups = []
    selected = [True] * n
    rate = np.float32(params["minMarkerDistanceRate"])
    for i in range(n):
        for j in range(i + 1, n):
            if _average_distance(cand[i], cand[j]) < np.float32(per[j] * rate):
                selected[i] = selected[j] = False
                if group[i] < 0 and group[j] < 0:
                    group[i] = group[j] = len(groups)
                    groups.append([i, j])
                elif group[i] > -1 and group[j] == -1:
                    group[j] = group[i]
                    groups[group[i]].append(j)
                elif group[j] > -1 and group[i] == -1:
                    group[i] = group[j]
                    groups[group[j]].append(i)
    close = {i: [] for i in range(n)}
    for g in groups:
        g.sort()
        cur = g[0]
        selected[cur] = True
        for k in g[1:]:
            dist = _average_distance(cand[k], cand[cur])
            module = _average_module_size(cand[k], marker_size, params["markerBorderBits"])
            if dist > np.float32(np.float32(params["minGroupDistance"]) * module):
                cur = k
                if not flag[k]:
                    close[g[0]].append(cand[k])
    return [(cand[i], close[i]) for i in range(n) if selected[i] and not flag[i]]


def _perimeter32(q):
    """MarkerCandidateTree's perimeter: float32 sum of the four side lengths."""
    s = np.float32(0)
    for i in range(4):
        d = q[i] - q[(i + 1) % 4]
        s = np.float32(s + np.sqrt(np.float32(d[0] * d[0] + d[1] * d[1])))
    return s


def perspective_transform(src, dst):
    """cv::getPerspectiveTransform (8 x 8 system, LU with partial pivoting in double)."""
    a = np.zeros((8, 8))
    b = np.zeros(8)
    for i in range(4):
        sx, sy = float(src[i][0]), float(src[i][1])
        dx, dy = float(dst[i][0]), float(dst[i][1])
        a[i] = [sx, sy, 1, 0, 0, 0, -sx * dx, -sy * dx]
        a[i + 4] = [0, 0, 0, sx, sy, 1, -sx * dy, -sy * dy]
        b[i] = dx
        b[i + 4] = dy
    x = np.linalg.solve(a, b)
    return np.append(x, 1.0).reshape(3, 3)


def warp_nearest(gray, m, size):
    """cv::warpPerspective(INTER_NEAREST, BORDER_CONSTANT 0): source pixel = round-half-even of M^-1 (x, y, 1)."""
    inv = np.linalg.inv(m)
    ys, xs = np.mgrid[0:size, 0:size].astype(np.float64)
    xw = inv[0, 0] * xs + inv[0, 1] * ys + inv[0, 2]
    yw = inv[1, 0] * xs + inv[1, 1] * ys + inv[1, 2]
    w = inv[2, 0] * xs + inv[2, 1] * ys + inv[2, 2]
    w = np.where(w != 0, 1.0 / np.where(w != 0, w, 1.0), 0.0)
    sx = np.rint(np.clip(xw * w, -2.0 ** 31, 2.0 ** 31 - 1)).astype(np.int64)
    sy = np.rint(np.clip(yw * w, -2.0 ** 31, 2.0 ** 31 - 1)).astype(np.int64)
    h, wd = gray.shape
    ok = (sx >= 0) & (sx < wd) & (sy >= 0) & (sy < h)
    out = np.zeros((size, size), np.uint8)
    out[ok] = gray[sy[ok], sx[ok]]
    return out


def otsu_threshold(img):
    """cv::threshold(THRESH_OTSU)'s threshold value (getThreshVal_Otsu_8u)."""
    hist = np.bincount(img.ravel(), minlength=256).astype(np.float64)
    n = img.size
    scale = 1.0 / n
    mu = float((np.arange(256) * hist).sum()) * scale
    mu1 = q1 = 0.0
    max_sigma = 0.0
    max_val = 0
    for i in range(256):
        p_i = hist[i] * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < np.finfo(np.float32).eps or max(q1, q2) > 1.0 - np.finfo(np.float32).eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma = sigma
            max_val = i
    return max_val


def extract_bits(gray, quad, params, marker_size=4):
    """_extractBits: perspective removal to (marker_size + 2 border) * cell pixels, Otsu, majority vote per cell."""
    border = params["markerBorderBits"]
    cell = params["perspectiveRemovePixelPerCell"]
    n = marker_size + 2 * border
    margin = int(params["perspectiveRemoveIgnoredMarginPerCell"] * cell)
    size = n * cell
    dst = np.array([[0, 0], [size - 1, 0], [size - 1, size - 1], [0, size - 1]], np.float32)
    m = perspective_transform(quad, dst)
    res = warp_nearest(gray, m, size)
    inner = res[cell // 2:size - cell // 2, cell // 2:size - cell // 2].astype(np.float64)
    mean = inner.mean()
    std = np.sqrt(max((inner * inner).mean() - mean * mean, 0.0))
    if std < params["minOtsuStdDev"]:
        return np.full((n, n), 1 if mean > 127 else 0, np.uint8)
    t = otsu_threshold(res)
    binary = res > t
    bits = np.zeros((n, n), np.uint8)
    side = cell - 2 * margin
    for y in range(n):
        for x in range(n):
            sq = binary[y * cell + margin:y * cell + margin + side, x * cell + margin:x * cell + margin + side]
            if int(sq.sum()) > (side * side) // 2:
                bits[y, x] = 1
    return bits


def dictionary_bits(byte_list, marker_size):
    """Bits (marker_size x marker_size) of every marker of a cv::aruco::Dictionary bytesList, rotation 0."""
    out = []
    nbits = marker_size * marker_size
    nbytes = (nbits + 7) // 8
    flat = np.asarray(byte_list, np.uint8).reshape(byte_list.shape[0], -1)      # per marker: 4 rotations x nbytes
    for m in range(flat.shape[0]):
        # cv::aruco::Dictionary::getByteListFromBits fills bytes most significant bit first; the last, partial byte
        # keeps its bits in the LOW positions
        full = np.unpackbits(flat[m, :nbytes - 1])
        rem = nbits - 8 * (nbytes - 1)
        last = np.unpackbits(flat[m, nbytes - 1:nbytes])[8 - rem:]
        out.append(np.concatenate([full, last]).reshape(marker_size, marker_size))
    return np.array(out, np.uint8)


def identify(bits, dict_bits, max_correction_bits, params):
    """_identifyOneCandidate after _extractBits: border check, then Dictionary::identify (first marker within the
    corrected distance, smallest-distance rotation).  Returns (id, rotation) or None."""
    border = params["markerBorderBits"]
    n = bits.shape[0]
    msize = n - 2 * border
    max_border_err = int(msize * msize * params["maxErroneousBitsInBorderRate"])
    err = 0
    for k in range(n):
        for b in range(border):
            err += int(bits[k, b]) + int(bits[k, n - 1 - b])
    for k in range(border, n - border):
        for b in range(border):
            err += int(bits[b, k]) + int(bits[n - 1 - b, k])
    if err > max_border_err:
        return None
    inner = bits[border:n - border, border:n - border]
    max_corr = int(max_correction_bits * params["errorCorrectionRate"])
    for m in range(dict_bits.shape[0]):
        best, rot = msize * msize + 1, -1
        ref = dict_bits[m]
        for r in range(4):
            # rotation r of the byte list = the marker turned counter-clockwise r times
            d = int((np.rot90(ref, r) != inner).sum())
            if d < best:
                best, rot = d, r
        if best <= max_corr:
            return m, rot
    return None


def detect_markers(image, dict_bits, max_correction_bits, params=None, marker_size=4):
    """cv::aruco::ArucoDetector::detectMarkers without corner refinement.  Returns (corners, ids): corners[i] is a
    4 x 2 float32 array in marker order (top-left first, clockwise), like cv2's (1, 4, 2) entries."""
    params = dict(REFERENCE_PARAMS if params is None else params)
    gray = to_gray(image)
    quads, flags = [], []
    nscales = (params["adaptiveThreshWinSizeMax"] - params["adaptiveThreshWinSizeMin"]) // \
        params["adaptiveThreshWinSizeStep"] + 1
    for i in range(nscales):
        win = params["adaptiveThreshWinSizeMin"] + i * params["adaptiveThreshWinSizeStep"]
        thresh = adaptive_threshold(gray, win, params["adaptiveThreshConstant"])
        q, _, f = find_marker_contours(thresh, params)
        quads += q
        flags += f
    quads = [reorder_corners(q) for q in quads]
    return identify_candidates(gray, quads, dict_bits, max_correction_bits, params, marker_size, flags)


def identify_candidates(gray, quads, dict_bits, max_correction_bits, params, marker_size=4, near_border=None):
    """Second half of detectMarkers (from the clockwise quads on): grouping, identification with the group's
    fall-backs, rotation of the corners so that the marker's top-left comes first."""
    corners, ids = [], []
    for quad, close in filter_too_close(quads, params, marker_size, near_border):
        found = None
        for q in [quad] + close:
            r = identify(extract_bits(gray, q, params, marker_size), dict_bits, max_correction_bits, params)
            if r is not None:
                found = (q, r)
                break
        if found is None:
            continue
        q, (mid, rot) = found
        if rot:
            q = np.roll(q, rot, axis=0)            # std::rotate(begin, begin + 4 - rot, end)
        corners.append(q.astype(np.float32))
        ids.append(mid)
    return corners, ids
