"""Motion clip files (host side, load time only).

Same data model as the reference (add_gym/anim/motion.py:7-75): a clip is ``frames[F,36]`` float64
(root xyz, root quaternion **xyzw**, 29 hinge angles in file order) + ``fps`` (30) + ``loop_mode``
(CLAMP).  Accepted containers: the reference's CSV ``.motion`` text and its ``.pkl`` dict, plus a
plain ``.npy`` of the frame matrix (binary clips avoid the per-value ``float()`` parse), and
-- one level further -- a pre-baked 100 Hz STEP TABLE (``.addkt``, ``save_step_table`` / ``load_step_table``):
the whole motion library after resampling, exactly the bytes that sit in HBM at run time, so a restart neither
parses text nor rebuilds the table.
Unlike the reference, loading a ``.motion`` does NOT write a ``.pkl`` next to it (motion.py:40-42
does; the asset tree may be read-only).
"""
import enum
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
