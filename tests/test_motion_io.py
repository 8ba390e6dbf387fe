"""Host-side motion containers (no GPU): the pre-baked step-table file round-trips bit for bit and
refuses truncated / foreign files; the clip loaders agree across the CSV, .npy and .pkl containers."""
import os
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


def test_clip_pack_is_lossless_and_addressable(tmp_path):
    """.addkc (BASELINE configs[2]: the whole 42-clip library in one file): bit-for-bit round trip incl. signed zeros,
    `pack#clip@N` addressing, the all-clips expansion, and refusal of values / files that would not be lossless."""
    from add_gym_b200 import config as b200_config
    rng = np.random.default_rng(0)
    a = np.round(rng.normal(size=(50, 36)) * 3, 6)
    a[3, 5] = -0.0
    a[7, 0] = 0.0
    b = np.round(rng.normal(size=(17, 36)), 6)
    p = str(tmp_path / "two.addkc")
    motion_io.save_clip_pack(p, {"zeta": a, "alpha": b})
    motion_io._pack_cache.clear()
    back = motion_io.load_clip_pack(p)
    assert sorted(back) == ["alpha", "zeta"]
    assert np.array_equal(back["zeta"].view(np.int64), a.view(np.int64)) and np.signbit(back["zeta"][3, 5])
    assert np.array_equal(back["alpha"].view(np.int64), b.view(np.int64))
    m = motion_io.load_motion(p + "#zeta@20")
    assert m.frames.shape == (20, 36) and m.fps == 30 and np.array_equal(m.frames, a[:20])
    files, weights = motion_io.fetch_motion_files(p)
    assert files == [p + "#alpha", p + "#zeta"] and weights == [1.0, 1.0]
    with pytest.raises(ValueError):
        motion_io.save_clip_pack(str(tmp_path / "bad.addkc"), {"x": rng.normal(size=(4, 36))})     # not 6-decimal values
    with open(p, "rb") as f:
        blob = f.read()
    bad = str(tmp_path / "cut.addkc")
    with open(bad, "wb") as f:
        f.write(blob[:len(blob) // 2])
    with pytest.raises(ValueError):
        motion_io.load_clip_pack(bad)
    # the shipped library: 42 clips, and the three .npy clips of round 1 are the same frames
    pack = motion_io.load_clip_pack(os.path.join(b200_config.ASSET_DIR, "motions_all.addkc"))
    assert len(pack) == 42 and sum(v.shape[0] for v in pack.values()) == 271897
    walk = np.load(os.path.join(b200_config.ASSET_DIR, "walk1_subject1_trimmed.npy"))
    assert np.array_equal(pack["walk1_subject1_trimmed"].view(np.int64), walk.view(np.int64))
    files, weights = motion_io.fetch_motion_files(os.path.join(b200_config.ASSET_DIR, "seven_clips.yaml"))
    assert len(files) == 7 and motion_io.load_motion(files[0]).frames.shape == (400, 36)
