"""Host-side motion containers (no GPU): the pre-baked step-table file round-trips bit for bit and
refuses truncated / foreign files; the clip loaders agree across the CSV, .npy and .pkl containers."""
import pickle

import numpy as np
import pytest

from add_gym_b200 import motion_io


def _header(rows, stride=72, clips=(3, 5)):
    steps = [rows - 7, 7]
    return dict(row_stride=stride, num_dofs=29, dt=0.01, s_total=rows, files=["a.motion", "b.motion"], weights=[0.25, 0.75],
                fps=[30.0, 30.0], num_frames=list(clips), lengths=[float(np.float32(0.1)), float(np.float32(1.0 / 30.0 * 4))],
                loop_modes=[0, 1], num_steps=steps)


def test_step_table_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    table = rng.standard_normal((40, 72)).astype(np.float32)
    table[3, 5] = np.float32(1e-42)       # denormal and signed zero must survive: the file is the HBM bytes
    table[4, 6] = -0.0
    p = str(tmp_path / "lib.addkt")
    motion_io.save_step_table(p, _header(40), table)
    h, t = motion_io.load_step_table(p)
    assert t.dtype == np.float32 and t.shape == (40, 72)
    assert t.tobytes() == table.tobytes()
    assert h["num_steps"] == [33, 7] and h["weights"] == [0.25, 0.75]
    assert np.float32(h["lengths"][0]) == np.float32(0.1)                 # float32 values survive the JSON header exactly


def test_step_table_rejects_bad_files(tmp_path):
    table = np.zeros((10, 72), np.float32)
    with pytest.raises(ValueError):
        motion_io.save_step_table(str(tmp_path / "x.addkt"), _header(11), table)          # rows != header
    bad = dict(_header(10))
    del bad["lengths"]
    with pytest.raises(ValueError):
        motion_io.save_step_table(str(tmp_path / "x.addkt"), bad, table)
    p = str(tmp_path / "ok.addkt")
    motion_io.save_step_table(p, _header(10), table)
    blob = open(p, "rb").read()
    open(str(tmp_path / "trunc.addkt"), "wb").write(blob[:-4])
    with pytest.raises(ValueError):
        motion_io.load_step_table(str(tmp_path / "trunc.addkt"))
    open(str(tmp_path / "foreign.addkt"), "wb").write(b"not a table" + blob)
    with pytest.raises(ValueError):
        motion_io.load_step_table(str(tmp_path / "foreign.addkt"))
    open(str(tmp_path / "empty.addkt"), "wb").write(b"")
    with pytest.raises(ValueError):
        motion_io.load_step_table(str(tmp_path / "empty.addkt"))


def test_clip_containers_agree(tmp_path):
    rng = np.random.default_rng(1)
    frames = rng.standard_normal((6, 36))
    with open(str(tmp_path / "c.motion"), "w") as f:
        for row in frames:
            f.write(",".join(repr(float(v)) for v in row) + "\n")
    np.save(str(tmp_path / "c.npy"), frames)
    with open(str(tmp_path / "c.pkl"), "wb") as f:
        pickle.dump({"loop_mode": motion_io.LoopMode.WRAP.value, "fps": 30, "frames": frames}, f)
    a = motion_io.load_motion(str(tmp_path / "c.motion"))
    b = motion_io.load_motion(str(tmp_path / "c.npy"))
    c = motion_io.load_motion(str(tmp_path / "c.pkl"))
    assert np.array_equal(a.frames, frames) and np.array_equal(b.frames, frames) and np.array_equal(c.frames, frames)
    assert a.loop_mode == motion_io.LoopMode.CLAMP and c.loop_mode == motion_io.LoopMode.WRAP
    assert a.get_length() == pytest.approx(5 / 30.0)
