"""The C++ example host (examples/c_abi_host.cpp, header-only use of the C ABI) against the Python mirror."""
import os
import subprocess

import numpy as np
import pytest

from denseflow_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("alg,step", [("tvl1", 1), ("farn", -2)])
def test_cpp_host_writes_the_same_planes(tmp_path, alg, step):
    import __graft_entry__ as g
    import denseflow_b200 as d
    exe = g.build_example()
    fr = synth.stream(96, 128, 5, seed=51)
    (tmp_path / "f.raw").write_bytes(fr.tobytes())
    r = subprocess.run([exe, str(tmp_path / "f.raw"), "128", "96", "5", alg, str(step), "20", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    m = 5 - abs(step)
    assert "1 videos (5 frames, %d %s flows) processed" % (m, alg) in r.stdout  # summary line, src/denseflow_gpu.cpp:494-496
    qx, qy = d.create(alg, 0, 128, 96).calc_batch(list(fr), step=step, bound=20)
    for i in range(m):
        idx = i if step > 0 else i + abs(step)
        infix = "" if step == 1 else ("p%d_" % step if step > 1 else "m%d_" % abs(step))
        for c, ref in (("x", qx[i]), ("y", qy[i])):
            data = (tmp_path / ("flow_%s_%s%05d.pgm" % (c, infix, idx))).read_bytes()
            hdr = b"P5\n128 96\n255\n"
            assert data.startswith(hdr)
            assert np.array_equal(np.frombuffer(data[len(hdr):], np.uint8).reshape(96, 128), ref)


def test_cpp_list_host_done_markers_and_resume(tmp_path):
    """examples/list_host.cpp: the list mode on the bare C ABI — dynamic queue over two workers, files numbered by global flow index,
    one .done marker per video written on its last chunk, finished videos skipped on the second run (tools/denseflow.cpp:66-73)."""
    import __graft_entry__ as g
    import denseflow_b200 as d
    exe = g.build_example("list_host")
    clips = np.stack([synth.stream(64, 96, 7, seed=70 + c, phase=2.0 * c) for c in range(5)])
    (tmp_path / "clips.raw").write_bytes(clips.tobytes())
    out = tmp_path / "out"
    out.mkdir()
    args = [exe, str(tmp_path / "clips.raw"), "96", "64", "7", "5", "tvl1", "1", "32", str(out), "2", "0"]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "5 videos (35 frames, 30 tvl1 flows) processed" in r.stdout and r.stdout.count("done video") == 5
    e = d.create("tvl1", 0, 96, 64)
    for c in range(5):
        assert (out / ".done" / ("%04d" % c)).exists()
        qx, qy = e.calc_batch(list(clips[c]), step=1, bound=32)
        for i in range(6):
            data = (out / ("%04d" % c) / ("flow_x_%05d.pgm" % i)).read_bytes()
            assert np.array_equal(np.frombuffer(data[len(b"P5\n96 64\n255\n"):], np.uint8).reshape(64, 96), qx[i])
    # resume: remove two markers, the second run processes exactly those two videos
    (out / ".done" / "0001").unlink()
    (out / ".done" / "0003").unlink()
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0 and "2 videos (14 frames, 12 tvl1 flows) processed" in r.stdout
    assert "done video 0001" in r.stdout and "done video 0003" in r.stdout and r.stdout.count("done video") == 2
