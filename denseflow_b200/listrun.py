"""Video-list dispatch (SURVEY §8 e, BASELINE.json configs[4]): host-side mirror of dfb_run_list / dfb_queue_*.

The reference reads list.txt into a vector of videos (/root/reference/tools/denseflow.cpp:54-81), walks it on one GPU and marks
each video done after its last buffer is written (src/denseflow_gpu.cpp:456-470).  `run_list` hands the same kind of list to W
workers (one engine handle per entry of `devices`) that pull videos from one dynamic queue; with `queue=WorkQueue(name)` the
queue lives in POSIX shared memory and is shared by several processes (one process per GPU).
"""
import ctypes as C

import numpy as np

from . import _lib


class WorkQueue:
    """Cross-process fetch-and-increment counter (dfb_queue_*)."""

    def __init__(self, name, create, variant="default"):
        self._L = _lib.load(variant)
        self._q = C.c_void_p()
        self.name = name
        self.owner = bool(create)
        if self._L.dfb_queue_open(name.encode(), int(bool(create)), C.byref(self._q)) != _lib.DFB_OK:
            raise RuntimeError("cannot open work queue %s" % name)

    def next(self):
        return self._L.dfb_queue_next(self._q)

    def reset(self):
        self._L.dfb_queue_reset(self._q)

    def close(self):
        if self._q:
            self._L.dfb_queue_close(self._q, int(self.owner))
            self._q = C.c_void_p()


def _pack_clips(clips):
    keep = []
    arr = (_lib.Clip * max(len(clips), 1))()
    for i, frames in enumerate(clips):
        fr = []
        for f in frames:
            f = np.asarray(f)
            if f.dtype != np.uint8 or f.ndim != 2 or not f.flags.c_contiguous or (fr and f.shape != fr[0].shape):
                raise RuntimeError("clip %d: frames must be C-contiguous uint8 [H,W] images of one size" % i)
            fr.append(f)
        ptrs = (C.c_void_p * max(len(fr), 1))(*[f.ctypes.data for f in fr])
        keep.append((fr, ptrs))
        arr[i].frames = ptrs
        arr[i].n_frames = len(fr)
        arr[i].height, arr[i].width = (fr[0].shape if fr else (1, 1))
    shapes = [(arr[i].height, arr[i].width) for i in range(len(clips))]
    return arr, shapes, keep


def _callback(on_chunk, shapes):
    if on_chunk is None:
        return None

    def _cb(user, clip, dev, first, n, last, qx, qy, flows):
        h, w = shapes[clip]
        if qx:
            vx = [np.ctypeslib.as_array(C.cast(qx[i], C.POINTER(C.c_uint8)), (h, w)) for i in range(n)]
            vy = [np.ctypeslib.as_array(C.cast(qy[i], C.POINTER(C.c_uint8)), (h, w)) for i in range(n)]
            on_chunk(clip, dev, first, bool(last), vx, vy, None)
        else:
            vf = [np.ctypeslib.as_array(C.cast(flows[i], C.POINTER(C.c_float)), (h, w, 2)) for i in range(n)]
            on_chunk(clip, dev, first, bool(last), None, None, vf)

    return _lib.CHUNK_DONE_FN(_cb)


def _stats(st):
    W = st.workers
    return {"clips": st.clips, "flows": st.flows, "frames": st.frames, "seconds": st.seconds, "workers": W,
            "clips_per_worker": list(st.clips_per_worker[:W]), "flows_per_worker": list(st.flows_per_worker[:W]),
            "busy_seconds_per_worker": list(st.busy_seconds_per_worker[:W]),
            "finish_seconds_per_worker": list(st.finish_seconds_per_worker[:W]),
            "kernel_launches": st.kernel_launches, "h2d_bytes": st.h2d_bytes, "d2h_bytes": st.d2h_bytes}


class ListRunner:
    """Workers kept alive between lists (dfb_list_open / dfb_list_run / dfb_list_close): worker i owns an engine handle on
    devices[i]; a device may appear more than once."""

    def __init__(self, algorithm, devices, max_width, max_height, variant="default"):
        self._L = _lib.load(variant)
        self._r = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        err = C.create_string_buffer(512)
        rc = self._L.dfb_list_open(algorithm.encode(), devs, len(devices), int(max_width), int(max_height), C.byref(self._r), err, 512)
        if rc != _lib.DFB_OK:
            self._r = None
            raise RuntimeError(err.value.decode() or "dfb_list_open failed (%d)" % rc)

    def run(self, clips, step=1, bound=32, chunk_flows=0, queue=None, on_chunk=None, packed=None):
        """packed: the result of a previous pack(clips) (building the pointer tables of a long list takes a while)."""
        arr, shapes, keep = packed if packed is not None else _pack_clips(clips)
        cb = _callback(on_chunk, shapes)
        st = _lib.ListStats()
        err = C.create_string_buffer(512)
        rc = self._L.dfb_list_run(self._r, arr, len(shapes), int(step), int(bound), int(chunk_flows), queue._q if queue is not None else None,
                                  C.cast(cb, C.c_void_p) if cb else None, None, C.byref(st), err, 512)
        if rc != _lib.DFB_OK:
            raise RuntimeError(err.value.decode() or "dfb_list_run failed (%d)" % rc)
        return _stats(st)

    @staticmethod
    def pack(clips):
        return _pack_clips(clips)

    def close(self):
        if getattr(self, "_r", None):
            self._L.dfb_list_close(self._r)
            self._r = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_list(algorithm, devices, clips, step=1, bound=32, chunk_flows=0, queue=None, on_chunk=None, variant="default"):
    """clips: list of sequences of uint8 [H,W] frames (one sequence per video; frames of one video share a size).
    on_chunk(clip_index, device, first_flow, last_chunk, qx, qy, flows): called on the worker's thread when every output of
    a chunk is in host memory (qx/qy: lists of uint8 [H,W] views when bound > 0, flows: list of float32 [H,W,2] views otherwise;
    the views die when the callback returns).  `last_chunk` is the reference's FlowBuffer::last_buffer.
    Returns a dict of dfb_list_stats."""
    L = _lib.load(variant)
    arr, shapes, keep = _pack_clips(clips)
    cb = _callback(on_chunk, shapes)
    devs = (C.c_int * len(devices))(*devices)
    st = _lib.ListStats()
    err = C.create_string_buffer(512)
    rc = L.dfb_run_list(algorithm.encode(), devs, len(devices), arr, len(clips), int(step), int(bound), int(chunk_flows),
                        queue._q if queue is not None else None, C.cast(cb, C.c_void_p) if cb else None, None, C.byref(st), err, 512)
    if rc != _lib.DFB_OK:
        raise RuntimeError(err.value.decode() or "dfb_run_list failed (%d)" % rc)
    return _stats(st)
