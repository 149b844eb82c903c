"""Argument validation at the C ABI (round-1 advisor findings): bad inputs come back as an error status with a message —
never an out-of-bounds read, a misaligned-address fault or silent workspace corruption."""
import ctypes as C

import numpy as np
import pytest

from denseflow_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def test_scale_step_larger_than_the_allocated_pyramid_is_rejected(oracle):
    import denseflow_b200 as d
    e = d.OpticalFlowDual_TVL1.create(0, 256, 256)
    with pytest.raises(RuntimeError, match="bad parameter scale_step"):
        e.set("scale_step", 0.9)          # levels would outgrow the slots laid out for 0.8
    e.set("scale_step", 0.7)              # a smaller factor fits; TMA descriptors are re-encoded for the new geometry
    a, b, _ = synth.pair(256, 256, 0)
    ref = oracle.tvl1_calc(a, b, oracle.tvl1_params(scale_step=0.7))
    assert synth.aee(e.calc(a, b), ref) <= 0.01
    e.set("scale_step", 0.8)
    assert synth.aee(e.calc(a, b), oracle.tvl1_calc(a, b)) <= 0.01


def test_frames_of_different_sizes_are_rejected_before_the_c_call():
    import denseflow_b200 as d
    e = d.OpticalFlowDual_TVL1.create(0, 128, 128)
    good = np.zeros((96, 128), np.uint8)
    with pytest.raises(RuntimeError, match="frame 1"):
        e.calc_batch([good, np.zeros((64, 128), np.uint8)], step=1)
    with pytest.raises(RuntimeError, match="frame 0"):
        e.calc_batch([np.zeros((96, 128), np.float32), good], step=1)
    with pytest.raises(RuntimeError, match="flows"):
        e.calc_batch([good, good], step=1, flows=np.zeros((1, 96, 64, 2), np.float32))


def test_device_pointers_are_rejected_by_the_host_entry_points():
    import torch
    import denseflow_b200 as d
    e = d.OpticalFlowDual_TVL1.create(0, 128, 96)
    L = _lib.load()
    dev = torch.zeros((96, 128), dtype=torch.uint8, device="cuda")
    host = np.zeros((96, 128), np.uint8)
    out = np.zeros((96, 128, 2), np.float32)
    fp = (C.c_void_p * 2)(dev.data_ptr(), host.ctypes.data)
    op = (C.c_void_p * 1)(out.ctypes.data)
    assert L.dfb_calc_batch_host(e._h, fp, 2, 1, 128, 96, op) == _lib.DFB_ERR_INVALID_ARG
    assert b"device pointer" in L.dfb_last_error(e._h)
    fp = (C.c_void_p * 2)(host.ctypes.data, host.ctypes.data)
    op = (C.c_void_p * 1)(torch.zeros((96, 128, 2), device="cuda").data_ptr())
    assert L.dfb_calc_batch_host(e._h, fp, 2, 1, 128, 96, op) == _lib.DFB_ERR_INVALID_ARG


def test_bgr_chain_rejects_null_and_device_pointers():
    import torch
    import denseflow_b200 as d
    e = d.OpticalFlowDual_TVL1.create(0, 128, 96)
    L = _lib.load()
    host = np.zeros((96, 128, 3), np.uint8)
    dev = torch.zeros((96, 128, 3), dtype=torch.uint8, device="cuda")
    cap = L.dfb_jpeg_max_bytes(128, 96)
    jx, jy = np.empty(cap, np.uint8), np.empty(cap, np.uint8)
    lx, ly = (C.c_size_t * 1)(), (C.c_size_t * 1)()
    xp, yp = (C.c_void_p * 1)(jx.ctypes.data), (C.c_void_p * 1)(jy.ctypes.data)
    for bad in ((C.c_void_p * 2)(host.ctypes.data, dev.data_ptr()), (C.c_void_p * 2)(host.ctypes.data, None)):
        assert L.dfb_process_bgr_batch_host(e._h, bad, 2, 1, 128, 96, 0, 0, 20, 95, xp, yp, cap, lx, ly) == _lib.DFB_ERR_INVALID_ARG
        assert b"bgr[1]" in L.dfb_last_error(e._h)
    fp = (C.c_void_p * 2)(host.ctypes.data, host.ctypes.data)
    assert L.dfb_process_bgr_batch_host(e._h, fp, 2, 1, 128, 96, 0, 0, 20, 95, (C.c_void_p * 1)(None), yp, cap, lx, ly) == _lib.DFB_ERR_INVALID_ARG
    assert L.dfb_process_bgr_batch_host(e._h, fp, 2, 1, 128, 96, 0, 0, 20, 95, xp, yp, cap, lx, ly) == _lib.DFB_OK
    assert lx[0] > 100 and ly[0] > 100


def test_pitch_and_alignment_checks_on_device_entry_points():
    import torch
    import denseflow_b200 as d
    e = d.OpticalFlowDual_TVL1.create(0, 128, 96)
    L = _lib.load()
    flow = torch.zeros((96, 130, 2), dtype=torch.float32, device="cuda")
    q = torch.zeros((96, 128), dtype=torch.uint8, device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    # flow pitch smaller than a row, q pitch smaller than a row, misaligned flow base / pitch
    assert L.dfb_quantise_device(e._h, flow.data_ptr(), 100, 128, 96, 20, q.data_ptr(), q.data_ptr(), 128, s) == _lib.DFB_ERR_INVALID_ARG
    assert L.dfb_quantise_device(e._h, flow.data_ptr(), 130 * 8, 128, 96, 20, q.data_ptr(), q.data_ptr(), 64, s) == _lib.DFB_ERR_INVALID_ARG
    assert L.dfb_quantise_device(e._h, flow.data_ptr() + 4, 130 * 8, 128, 96, 20, q.data_ptr(), q.data_ptr(), 128, s) == _lib.DFB_ERR_INVALID_ARG
    assert L.dfb_quantise_device(e._h, flow.data_ptr(), 130 * 8 + 4, 128, 96, 20, q.data_ptr(), q.data_ptr(), 128, s) == _lib.DFB_ERR_INVALID_ARG
    assert L.dfb_quantise_device(e._h, flow.data_ptr(), 130 * 8, 128, 96, 20, q.data_ptr(), q.data_ptr(), 128, s) == _lib.DFB_OK
    a = torch.zeros((96, 128), dtype=torch.uint8, device="cuda")
    assert L.dfb_calc_device(e._h, a.data_ptr(), 128, a.data_ptr(), 128, 128, 96, flow.data_ptr() + 4, 130 * 8, s) == _lib.DFB_ERR_INVALID_ARG
    torch.cuda.synchronize()
