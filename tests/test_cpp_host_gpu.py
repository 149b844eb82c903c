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
