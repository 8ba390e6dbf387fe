"""Motion clip files (host side, load time only).

Same data model as the reference (add_gym/anim/motion.py:7-75): a clip is ``frames[F,36]`` float64
(root xyz, root quaternion **xyzw**, 29 hinge angles in file order) + ``fps`` (30) + ``loop_mode``
(CLAMP).  Accepted containers: the reference's CSV ``.motion`` text and its ``.pkl`` dict, plus a
plain ``.npy`` of the frame matrix (binary clips avoid the per-value ``float()`` parse), and
-- one level further -- a pre-baked 100 Hz STEP TABLE (``.addkt``, ``save_step_table`` / ``load_step_table``):
the whole motion library after resampling, exactly the bytes that sit in HBM at run time, so a restart neither
parses text nor rebuilds the table.
A whole library of clips can also travel as ONE lossless clip pack (``.addkc``, ``save_clip_pack`` /
``load_clip_pack``): the reference's ``.motion`` text holds 6-decimal numbers, i.e. integers / 1e6, so the
pack stores zig-zag first differences of those integers, column-major, byte-shuffled and deflated (42
clips, 271,897 frames: 89 MB of text -> 19 MB) and reproduces ``float(text)`` bit for bit (signed zeros
included).  ``pack.addkc#clip`` names one clip of a pack (``#clip@N``: its first N frames) wherever a clip
file is accepted; ``pack.addkc`` alone as ``motion_file`` = every clip, weight 1.0 (BASELINE configs[2]).
Unlike the reference, loading a ``.motion`` does NOT write a ``.pkl`` next to it (motion.py:40-42
does; the asset tree may be read-only).
"""
import enum
import os
import pickle

import numpy as np


class LoopMode(enum.Enum):
    CLAMP = 0
    WRAP = 1


class Motion:
    def __init__(self, loop_mode, fps, frames):
        self.loop_mode = loop_mode
        self.fps = fps
        self.frames = frames

    def get_length(self):
        return float(self.frames.shape[0] - 1) / self.fps


def load_motion(file, loop_mode=LoopMode.CLAMP, fps=30):
    if "#" in file and file.split("#", 1)[0].endswith(".addkc"):
        path, clip = file.split("#", 1)
        clip, _, cut = clip.partition("@")
        frames = load_clip_pack(path)[clip]
        return Motion(loop_mode, fps, frames[:int(cut)] if cut else frames)
    if file.endswith(".motion"):
        rows = []
        with open(file, "r") as f:
            for line in f:
                rows.append([float(v) for v in line.strip().split(",")])
        return Motion(loop_mode, fps, np.array(rows))
    if file.endswith(".npy"):
        return Motion(loop_mode, fps, np.load(file))
    with open(file, "rb") as f:
        d = pickle.load(f)
    return Motion(LoopMode(d["loop_mode"]), d["fps"], d["frames"])


def fetch_motion_files(motion_file):
    """(files, weights) of a motion_file setting: a YAML library (reference format, motion_lib.py:337-358: `motions:`
    list of {file, weight}; relative paths resolve against the YAML's directory), a clip pack (every clip, weight 1.0,
    sorted by name like a generated YAML over assets/motions/*.motion would be) or a single clip."""
    if os.path.splitext(motion_file)[1] == ".yaml":
        import yaml
        with open(motion_file, "r") as f:
            cfg = yaml.load(f, Loader=yaml.SafeLoader)
        files, weights = [], []
        base = os.path.dirname(os.path.abspath(motion_file))
        for entry in cfg["motions"]:
            w = entry["weight"]
            assert w >= 0
            path = entry["file"]
            if not os.path.isabs(path) and not os.path.exists(path.split("#", 1)[0]):
                path = os.path.join(base, path)
            files.append(path)
            weights.append(w)
        return files, weights
    if motion_file.endswith(".addkc"):
        names = sorted(load_clip_pack(motion_file).keys())
        return ["%s#%s" % (motion_file, n) for n in names], [1.0] * len(names)
    return [motion_file], [1.0]


# ---- clip packs: a whole library of 30 fps clips in one lossless file -------------------------------------------------
# layout: MAGIC | uint32 header length | JSON header {clips: [{name, frames}], cols, scale, neg_zero: [[row, col], ...],
#         raw_bytes} | zlib stream of the byte-shuffled, column-major, zig-zag first differences (uint32) of
#         round(value * scale) over the concatenated clips
CLIP_PACK_MAGIC = b"ADDKC1\n"
_pack_cache = {}


def save_clip_pack(path, clips, scale=1000000):
    """clips: {name: float64 [F, cols]} whose values are multiples of 1/scale (the reference's 6-decimal text).
    Raises if any value would not survive the round trip bit for bit."""
    import json
    import struct
    import zlib
    names = sorted(clips.keys())
    cat = np.concatenate([np.asarray(clips[n], dtype=np.float64) for n in names], axis=0)
    q = np.round(cat * scale).astype(np.int64)
    back = q / float(scale)
    neg_zero = np.argwhere((cat == 0) & np.signbit(cat))
    back[neg_zero[:, 0], neg_zero[:, 1]] = -0.0
    if not np.array_equal(back.view(np.int64), cat.view(np.int64)) or np.abs(q).max() >= 2 ** 30:
        raise ValueError("clip values are not exact multiples of 1/%d: the pack would not be lossless" % scale)
    d = np.diff(q, axis=0, prepend=0)
    z = ((d << 1) ^ (d >> 63)).astype("<u4")
    shuf = np.ascontiguousarray(np.ascontiguousarray(z.T).view(np.uint8).reshape(-1, 4).T)
    header = {"clips": [{"name": n, "frames": int(np.asarray(clips[n]).shape[0])} for n in names],
              "cols": int(cat.shape[1]), "scale": int(scale), "neg_zero": neg_zero.tolist(), "raw_bytes": int(shuf.size)}
    blob = json.dumps(header).encode("utf-8")
    with open(path, "wb") as f:
        f.write(CLIP_PACK_MAGIC)
        f.write(struct.pack("<I", len(blob)))
        f.write(blob)
        f.write(zlib.compress(shuf.tobytes(), 9))


def load_clip_pack(path):
    """-> {name: float64 [F, cols]}, cached per path; raises ValueError on a truncated or foreign file."""
    import json
    import struct
    import zlib
    key = os.path.abspath(path)
    if key in _pack_cache:
        return _pack_cache[key]
    with open(path, "rb") as f:
        if f.read(len(CLIP_PACK_MAGIC)) != CLIP_PACK_MAGIC:
            raise ValueError("%s is not a clip pack" % path)
        raw = f.read(4)
        if len(raw) != 4:
            raise ValueError("%s: truncated header" % path)
        (n,) = struct.unpack("<I", raw)
        blob = f.read(n)
        if len(blob) != n:
            raise ValueError("%s: truncated header" % path)
        header = json.loads(blob.decode("utf-8"))
        try:
            data = zlib.decompress(f.read())
        except zlib.error as e:
            raise ValueError("%s: corrupt payload (%s)" % (path, e))
    cols = int(header["cols"])
    total = sum(int(c["frames"]) for c in header["clips"])
    if len(data) != int(header["raw_bytes"]) or len(data) != total * cols * 4:
        raise ValueError("%s: %d payload bytes, expected %d" % (path, len(data), total * cols * 4))
    shuf = np.frombuffer(data, dtype=np.uint8).reshape(4, -1)
    z = np.ascontiguousarray(shuf.T).view("<u4").reshape(cols, total).T.astype(np.int64)
    d = (z >> 1) ^ -(z & 1)
    q = np.cumsum(d, axis=0)
    vals = q / float(header["scale"])
    for r, c in header["neg_zero"]:
        vals[r, c] = -0.0
    out, r0 = {}, 0
    for c in header["clips"]:
        out[c["name"]] = vals[r0:r0 + int(c["frames"])]
        r0 += int(c["frames"])
    _pack_cache[key] = out
    return out


# ---- pre-baked step tables -------------------------------------------------------------------------
# layout: MAGIC | uint32 little-endian header length | JSON header (utf-8) | float32 little-endian table [S_total, row_stride]
STEP_TABLE_MAGIC = b"ADDKT1\n"
_STEP_TABLE_KEYS = ("row_stride", "num_dofs", "dt", "s_total", "files", "weights", "fps", "num_frames", "lengths",
                    "loop_modes", "num_steps")


def save_step_table(path, header, table):
    """header: dict with _STEP_TABLE_KEYS (per-clip lists; `lengths` are float32 values); table: float32 ndarray."""
    import json
    import struct
    missing = [k for k in _STEP_TABLE_KEYS if k not in header]
    if missing:
        raise ValueError("step table header lacks %s" % missing)
    table = np.ascontiguousarray(table, dtype="<f4")
    if table.ndim != 2 or table.shape != (int(header["s_total"]), int(header["row_stride"])):
        raise ValueError("table shape %s does not match the header" % (table.shape,))
    if int(np.sum(header["num_steps"])) != table.shape[0]:
        raise ValueError("num_steps do not add up to the table rows")
    blob = json.dumps(header, sort_keys=True).encode("utf-8")
    with open(path, "wb") as f:
        f.write(STEP_TABLE_MAGIC)
        f.write(struct.pack("<I", len(blob)))
        f.write(blob)
        f.write(table.tobytes())


def load_step_table(path):
    """-> (header dict, float32 ndarray [S_total, row_stride]); raises ValueError on a truncated or foreign file."""
    import json
    import struct
    with open(path, "rb") as f:
        if f.read(len(STEP_TABLE_MAGIC)) != STEP_TABLE_MAGIC:
            raise ValueError("%s is not a baked step table" % path)
        raw = f.read(4)
        if len(raw) != 4:
            raise ValueError("%s: truncated header" % path)
        (n,) = struct.unpack("<I", raw)
        blob = f.read(n)
        if len(blob) != n:
            raise ValueError("%s: truncated header" % path)
        header = json.loads(blob.decode("utf-8"))
        missing = [k for k in _STEP_TABLE_KEYS if k not in header]
        if missing:
            raise ValueError("%s: header lacks %s" % (path, missing))
        rows, stride = int(header["s_total"]), int(header["row_stride"])
        data = f.read()
    if len(data) != rows * stride * 4:
        raise ValueError("%s: %d table bytes, expected %d" % (path, len(data), rows * stride * 4))
    return header, np.frombuffer(data, dtype="<f4").reshape(rows, stride)
