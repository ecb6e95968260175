"""CPU ORACLE (test infrastructure): the reference's host-side schedules restated
in Python around a pluggable LM solve.

Follows /root/reference/ar_slam/src/ar_slam_util.cpp:
  load_yaml          :304-368      save_yaml        :371-465
  add_detections     :591-627      solve_incremental:629-678
  solve_capture      :680-742      solve (BFS)      :744-866
  add_connected      :869-885      localize_many/one:888-979
The LM solve itself (`optimize`, :1001-1018) is `oracle.pyoracle.solve` by
default; tests pass the CUDA solver instead to compare the two on the same
schedule.
"""
import numpy as np
import yaml

from . import pyoracle as po


class MapData:
    """captures_/arucos_/blocks_/camera_ of ArSlamSolver (ar_slam_util.hpp:473-485)."""

    def __init__(self):
        self.cap_uid, self.cap_fn, self.cap_pose = [], [], []
        self.cap_blocks, self.cap_init_block = [], []
        self.tag_id, self.tag_pose, self.tag_blocks, self.tag_initialized = [], [], [], []
        self.blk_cap, self.blk_tag, self.blk_rect, self.blk_added = [], [], [], []
        self.cap_map, self.tag_map = {}, {}
        self.cam = np.array([3000.0, 0.0, 0.0])  # hpp:66-70
        self.size = None
        self.unsolved = []  # insertion-ordered stand-in for unordered_set<CaptureHandle>
        self.solve_log = []

    # -- data store (hpp:419-457)
    def add_capture(self, uid, fn):
        if uid in self.cap_map:
            raise RuntimeError("Capture with uid already added")
        self.cap_map[uid] = len(self.cap_uid)
        self.cap_uid.append(uid)
        self.cap_fn.append(fn)
        self.cap_pose.append(np.zeros(6))
        self.cap_blocks.append([])
        self.cap_init_block.append(None)
        return len(self.cap_uid) - 1

    def add_aruco(self, tid):
        h = len(self.tag_id)
        self.tag_map.setdefault(tid, h)  # try_emplace keeps an existing handle
        self.tag_id.append(tid)
        self.tag_pose.append(np.zeros(6))
        self.tag_blocks.append([])
        self.tag_initialized.append(False)
        return h

    def get_or_add_aruco(self, tid):
        return self.tag_map[tid] if tid in self.tag_map else self.add_aruco(tid)

    def add_block(self, rect, cap, tag):
        b = len(self.blk_cap)
        self.blk_cap.append(cap)
        self.blk_tag.append(tag)
        self.blk_rect.append(np.asarray(rect, dtype=np.float64))
        self.blk_added.append(False)
        self.cap_blocks[cap].append(b)
        self.tag_blocks[tag].append(b)
        return b

    # -- yaml (Appendix C of SURVEY.md)
    def load_yaml(self, fn):
        with open(fn) as f:
            doc = yaml.safe_load(f)
        for uid, data in (doc.get("captures") or {}).items():
            uid = str(uid)
            if uid in self.cap_map:
                raise RuntimeError("capture with id %s already exists" % uid)
            c = self.add_capture(uid, str(data["img_fn"]))
            self.cap_pose[c] = np.array([float(v) for v in data["inv_pose"]])
        for tid, data in (doc.get("arucos") or {}).items():
            a = self.add_aruco(str(tid))
            self.tag_pose[a] = np.array([float(v) for v in data["pose"]])
        for blk in doc.get("blocks") or []:
            c = self.cap_map[str(blk["capture"])]
            a = self.tag_map[str(blk["aruco"])]
            rect = blk["aruco_rect"]
            if len(rect) != 8:
                raise RuntimeError("aruco_rect has wrong number of values")
            self.add_block([float(v) for v in rect], c, a)
        cam = doc["camera"]
        self.size = (int(cam["width"]), int(cam["height"]))
        for i, v in enumerate(cam["params"]):
            self.cam[i] = float(v)

    def save_yaml(self):
        def num(v):
            return repr(float(v))
        out = ["blocks:"]
        for b in range(len(self.blk_cap)):
            out += ["  - capture: %s" % self.cap_uid[self.blk_cap[b]],
                    "    aruco: %s" % self.tag_id[self.blk_tag[b]],
                    "    aruco_rect: [%s]" % ", ".join(num(v) for v in self.blk_rect[b])]
        out.append("captures:")
        for c in range(len(self.cap_uid)):
            out += ["  %s:" % self.cap_uid[c],
                    "    inv_pose: [%s]" % ", ".join(num(v) for v in self.cap_pose[c]),
                    "    img_fn: %s" % self.cap_fn[c]]
        out.append("arucos:")
        for a in range(len(self.tag_id)):
            out += ["  %s:" % self.tag_id[a],
                    "    pose: [%s]" % ", ".join(num(v) for v in self.tag_pose[a])]
        out += ["camera:", "  params: [%s]" % ", ".join(num(v) for v in self.cam)]
        if self.size is not None:
            out += ["  width: %d" % self.size[0], "  height: %d" % self.size[1]]
        return "\n".join(out) + "\n"

    # -- message input (:591-627); det = list of (id, rect8 as float32-rounded)
    def add_detections(self, capture_uid, image_path, width, height, dets):
        if not dets:
            return None
        if self.size is not None:
            if self.size != (width, height):
                return None
        else:
            self.size = (width, height)
        c = self.add_capture(capture_uid, image_path)
        for tid, rect in dets:
            a = self.get_or_add_aruco(tid)
            rect = np.asarray(rect, dtype=np.float32).astype(np.float64)  # Point32 -> double
            self.add_block(rect, c, a)
        self.unsolved.insert(0, c)  # libstdc++ unordered_set: newest first for distinct buckets
        return c


def oracle_lm(m, blocks, options=None, cam_const=False, tags_const=False):
    """`optimize` (:1001-1018) on the residual blocks added so far, with the oracle."""
    cap_idx = [m.blk_cap[b] for b in blocks]
    tag_idx = [m.blk_tag[b] for b in blocks]
    obs = np.array([m.blk_rect[b] for b in blocks])
    cam, cap, tag, summ, log = po.solve(
        len(m.cap_uid), len(m.tag_id), cap_idx, tag_idx, obs, m.cam, np.array(m.cap_pose),
        np.array(m.tag_pose), options=options, cam_const=cam_const,
        tag_const=np.ones(len(m.tag_id), np.uint8) if tags_const else None)
    return cam, cap, tag, summ


class Scheduler:
    def __init__(self, m, lm=oracle_lm, options=None):
        self.m, self.lm, self.options = m, lm, options
        self.problem_blocks = []  # residual blocks in AddResidualBlock order

    def _optimize(self, **kw):
        m = self.m
        cam, cap, tag, summ = self.lm(m, self.problem_blocks, self.options, **kw)
        m.cam[:] = cam
        for c in range(len(m.cap_uid)):
            m.cap_pose[c] = np.array(cap[c])
        for a in range(len(m.tag_id)):
            m.tag_pose[a] = np.array(tag[a])
        m.solve_log.append(summ)
        return summ

    def _add_capture_blocks(self, c):
        m = self.m
        for b in m.cap_blocks[c]:
            a = m.blk_tag[b]
            if not m.tag_initialized[a]:
                m.tag_initialized[a] = True
                m.tag_pose[a] = po.init_tag_pose(m.blk_rect[b], m.cam, m.cap_pose[c])
            if m.blk_added[b]:
                raise RuntimeError("block for capture was somehow already added?")
            m.blk_added[b] = True
            self.problem_blocks.append(b)

    # :680-742
    def solve_capture(self, c, init_block):
        m = self.m
        if init_block is not None:
            a = m.blk_tag[init_block]
            m.cap_pose[c] = po.init_capture_pose(m.blk_rect[init_block], m.cam, m.tag_pose[a])
        self._add_capture_blocks(c)
        return self._optimize()

    # :629-678
    def solve_incremental(self):
        m = self.m
        if len(m.unsolved) == len(m.cap_uid) and m.unsolved:
            c = m.unsolved.pop(0)
            self.solve_capture(c, None)
        repeat = True
        while repeat:
            repeat = False
            i = 0
            while i < len(m.unsolved):
                c = m.unsolved[i]
                for b in m.cap_blocks[c]:
                    if m.tag_initialized[m.blk_tag[b]]:
                        repeat = True
                        m.unsolved.pop(i)  # itr = erase(itr)
                        self.solve_capture(c, b)
                        break
                if i >= len(m.unsolved):
                    break
                i += 1  # the for loop's ++itr (skips the element after an erase, as the reference does)

    # :744-866
    def solve(self):
        m = self.m
        best, best_n = 0, len(m.cap_blocks[0])
        for c in range(1, len(m.cap_uid)):
            if len(m.cap_blocks[c]) > best_n:
                best, best_n = c, len(m.cap_blocks[c])
        m.cap_init_block[best] = -1  # BlockHandle(~0): "has a value"
        open_caps = [best]
        while open_caps:
            c = open_caps.pop(0)
            if c != best:
                b = m.cap_init_block[c]
                m.cap_pose[m.blk_cap[b]] = po.init_capture_pose(m.blk_rect[b], m.cam,
                                                                 m.tag_pose[m.blk_tag[b]])
            self._add_capture_blocks(c)
            self._optimize()
            # addConnectedCaptures :869-885
            for bb in m.cap_blocks[c]:
                for b in m.tag_blocks[m.blk_tag[bb]]:
                    cc = m.blk_cap[b]
                    if m.cap_init_block[cc] is None:
                        m.cap_init_block[cc] = b
                        open_caps.append(cc)

    # :888-979
    def localize_many(self, first_loc_cap_idx):
        m = self.m
        for c in range(first_loc_cap_idx, len(m.cap_uid)):
            seed = None
            for cb in m.cap_blocks[c]:
                if any(m.blk_cap[b] < first_loc_cap_idx for b in m.tag_blocks[m.blk_tag[cb]]):
                    seed = cb
                    break
            if seed is None:
                continue
            self.problem_blocks = []  # resetProblem
            m.cap_pose[m.blk_cap[seed]] = po.init_capture_pose(m.blk_rect[seed], m.cam,
                                                                m.tag_pose[m.blk_tag[seed]])
            for b in m.cap_blocks[c]:
                if m.blk_added[b]:
                    raise RuntimeError("block for capture was somehow already added?")
                m.blk_added[b] = True
                self.problem_blocks.append(b)
            self._optimize(cam_const=True, tags_const=True)
