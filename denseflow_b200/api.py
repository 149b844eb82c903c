"""Host-side mirror of the reference's algorithm-object interface for the hot path.

The reference (C++) creates `cuda::OpticalFlowDual_TVL1::create()` / `cuda::FarnebackOpticalFlow::create()`
(/root/reference/src/denseflow_gpu.cpp:299,301) and calls `alg->calc(gray_a, gray_b, flow, stream)` (:327,:329)
inside `DenseFlow::calc_optflows_imp` (:282-370).  The classes here keep those names, argument meaning and
error behaviour (unknown algorithm / unsupported size -> RuntimeError with the reference's message) on top of
the C ABI (include/denseflow_b200.h).  Everything numeric happens in the CUDA library.
"""
import ctypes as C

import numpy as np

from . import _lib


class DenseOpticalFlow:
    """cv::cuda::DenseOpticalFlow look-alike bound to one GPU."""

    algorithm = None

    def __init__(self, algorithm, device=0, max_width=1920, max_height=1080, variant="default"):
        self._L = _lib.load(variant)
        self._h = C.c_void_p()
        self.algorithm = algorithm
        self.device = device
        rc = self._L.dfb_create(algorithm.encode(), device, max_width, max_height, C.byref(self._h))
        if rc != _lib.DFB_OK:
            msg = self._L.dfb_last_error(None).decode()
            self._h = None
            raise RuntimeError(msg)  # reference: std::runtime_error -> what() + exit 1 (tools/denseflow.cpp:93-96)

    # -- lifetime (Ptr<>::release, src/denseflow_gpu.cpp:345-355) --
    def release(self):
        if getattr(self, "_h", None):
            self._L.dfb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _lib.DFB_OK:
            raise RuntimeError(self._L.dfb_last_error(self._h).decode() or "dfb error %d" % rc)

    def set(self, name, value):
        self._check(self._L.dfb_set_param(self._h, name.encode(), float(value)))
        return self

    def get(self, name):
        v = C.c_double()
        self._check(self._L.dfb_get_param(self._h, name.encode(), C.byref(v)))
        return v.value

    # -- calc(I0, I1) -> flow (CV_32FC2), src/denseflow_gpu.cpp:327/:329 --
    def calc(self, a, b, flow=None, stream=None):
        """a, b: uint8 [H,W] numpy arrays (host path: upload, calc, download) or CUDA torch tensors
        (device path: enqueued on `stream` / torch's current stream, asynchronous)."""
        if isinstance(a, np.ndarray):
            a = np.ascontiguousarray(a, np.uint8)
            b = np.ascontiguousarray(b, np.uint8)
            if a.ndim != 2 or a.shape != b.shape:
                raise RuntimeError("calc: frames must be two CV_8UC1 images of the same size")
            h, w = a.shape
            if flow is None:
                flow = np.empty((h, w, 2), np.float32)
            self._check(self._L.dfb_calc_host(self._h, a.ctypes.data, b.ctypes.data, w, h, flow.ctypes.data))
            return flow
        import torch
        if a.dtype != torch.uint8 or b.dtype != torch.uint8 or a.dim() != 2 or a.shape != b.shape or not a.is_cuda or not b.is_cuda \
                or a.stride(1) != 1 or b.stride(1) != 1:
            raise RuntimeError("calc: frames must be two CV_8UC1 CUDA tensors of the same size with unit column stride")
        h, w = a.shape
        if flow is None:
            flow = torch.empty((h, w, 2), dtype=torch.float32, device=a.device)
        s = stream if stream is not None else torch.cuda.current_stream(a.device).cuda_stream
        self._check(self._L.dfb_calc_device(self._h, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), w, h,
                                            flow.data_ptr(), flow.stride(0) * 4, C.c_void_p(s)))
        return flow

    @staticmethod
    def _gray_frames(frames):
        """Every frame must be a CV_8UC1 image of the first frame's size (the C side reads width*height bytes of each)."""
        out = []
        for i, f in enumerate(frames):
            f = np.asarray(f)
            if f.dtype != np.uint8 or f.ndim != 2 or (out and f.shape != out[0].shape):
                raise RuntimeError("frame %d: expected a uint8 [H,W] image of the same size as frame 0" % i)
            out.append(np.ascontiguousarray(f))
        return out

    @staticmethod
    def _check_out(a, shape, dtype, name):
        if not isinstance(a, np.ndarray) or a.dtype != dtype or a.shape[1:] != tuple(shape[1:]) or a.shape[0] < shape[0] \
                or not all(a[i].flags.c_contiguous for i in range(shape[0])):
            raise RuntimeError("%s: expected a C-contiguous %s array of shape %s" % (name, np.dtype(dtype).name, tuple(shape)))

    # -- the batch shape of calc_optflows_imp (src/denseflow_gpu.cpp:307-342) --
    def calc_batch(self, frames, step=1, flows=None, bound=None):
        """frames: sequence of uint8 [H,W] host arrays (or one [N,H,W] array).  Returns M = max(N-|step|,0)
        flows [M,H,W,2] float32, or with bound=B the quantised (qx, qy) uint8 [M,H,W] planes
        (convertFlowToImage, src/common.cpp:4-16, done on the GPU)."""
        frames = self._gray_frames(frames)
        n = len(frames)
        if n == 0:
            return np.empty((0, 0, 0, 2), np.float32)
        h, w = frames[0].shape
        m = max(n - abs(step), 0)
        fp = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
        if bound is None:
            if flows is None:
                flows = np.empty((m, h, w, 2), np.float32)
            self._check_out(flows, (m, h, w, 2), np.float32, "flows")
            op = (C.c_void_p * max(m, 1))(*[flows[i].ctypes.data for i in range(m)])
            self._check(self._L.dfb_calc_batch_host(self._h, fp, n, step, w, h, op))
            return flows
        qx = np.empty((m, h, w), np.uint8)
        qy = np.empty((m, h, w), np.uint8)
        xp = (C.c_void_p * max(m, 1))(*[qx[i].ctypes.data for i in range(m)])
        yp = (C.c_void_p * max(m, 1))(*[qy[i].ctypes.data for i in range(m)])
        self._check(self._L.dfb_calc_batch_host_u8(self._h, fp, n, step, w, h, int(bound), xp, yp))
        return qx, qy

    def calc_batch_u8_into(self, frames, step, bound, qx, qy):
        """calc_batch(..., bound=B) writing into caller-provided (e.g. pinned) uint8 arrays qx, qy of shape [M,H,W]."""
        frames = self._gray_frames(frames)
        n = len(frames)
        h, w = frames[0].shape
        m = max(n - abs(step), 0)
        self._check_out(qx, (m, h, w), np.uint8, "qx")
        self._check_out(qy, (m, h, w), np.uint8, "qy")
        fp = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
        xp = (C.c_void_p * max(m, 1))(*[qx[i].ctypes.data for i in range(m)])
        yp = (C.c_void_p * max(m, 1))(*[qy[i].ctypes.data for i in range(m)])
        self._check(self._L.dfb_calc_batch_host_u8(self._h, fp, n, step, w, h, int(bound), xp, yp))
        return qx, qy

    def calc_batch_device(self, frames, step=1, flows=None, stream=None):
        """frames: CUDA uint8 tensor [N,H,W] (contiguous); flows: CUDA float32 [M,H,W,2]."""
        import torch
        n, h, w = frames.shape
        m = max(n - abs(step), 0)
        if flows is None:
            flows = torch.empty((m, h, w, 2), dtype=torch.float32, device=frames.device)
        s = stream if stream is not None else torch.cuda.current_stream(frames.device).cuda_stream
        self._check(self._L.dfb_calc_batch_device(self._h, frames.data_ptr(), n, step, w, h, flows.data_ptr(),
                                                  C.c_void_p(s)))
        return flows

    def quantise_device(self, flow, bound):
        import torch
        h, w = flow.shape[:2]
        qx = torch.empty((h, w), dtype=torch.uint8, device=flow.device)
        qy = torch.empty((h, w), dtype=torch.uint8, device=flow.device)
        s = torch.cuda.current_stream(flow.device).cuda_stream
        self._check(self._L.dfb_quantise_device(self._h, flow.data_ptr(), flow.stride(0) * 4, w, h, int(bound),
                                                qx.data_ptr(), qy.data_ptr(), w, C.c_void_p(s)))
        return qx, qy

    # -- convertFlowToPngImage (src/common.cpp:18-46): the CV_8UC3 image `-st=png` encodes --
    def flow_to_png_image_device(self, flow, return_bounds=True):
        """flow: CUDA float32 [H,W,2].  Returns (bgr uint8 [H,W,3] CUDA tensor, (bound_x, bound_y) or None)."""
        import torch
        h, w = flow.shape[:2]
        bgr = torch.empty((h, w, 3), dtype=torch.uint8, device=flow.device)
        b = (C.c_double * 2)()
        s = torch.cuda.current_stream(flow.device).cuda_stream
        self._check(self._L.dfb_flow_to_png_image_device(self._h, flow.data_ptr(), flow.stride(0) * 4, w, h, bgr.data_ptr(), w * 3,
                                                         b if return_bounds else None, C.c_void_p(s)))
        return bgr, ((b[0], b[1]) if return_bounds else None)

    # -- frame preparation (cvtColor BGR2GRAY + resize INTER_LINEAR, src/denseflow_gpu.cpp:163-170), bit-exact to OpenCV CPU --
    def bgr_to_gray_device(self, bgr):
        import torch
        h, w = bgr.shape[:2]
        gray = torch.empty((h, w), dtype=torch.uint8, device=bgr.device)
        s = torch.cuda.current_stream(bgr.device).cuda_stream
        self._check(self._L.dfb_bgr_to_gray_device(self._h, bgr.data_ptr(), bgr.stride(0), w, h, gray.data_ptr(), w, C.c_void_p(s)))
        return gray

    def resize_gray_device(self, gray, dst_width, dst_height):
        import torch
        h, w = gray.shape
        out = torch.empty((dst_height, dst_width), dtype=torch.uint8, device=gray.device)
        s = torch.cuda.current_stream(gray.device).cuda_stream
        self._check(self._L.dfb_resize_gray_device(self._h, gray.data_ptr(), gray.stride(0), w, h, out.data_ptr(), dst_width,
                                                   dst_width, dst_height, C.c_void_p(s)))
        return out

    # -- imencode(".jpg", plane) on the GPU (src/common.cpp:56-57), OpenCV default quality 95 --
    def encode_jpeg_gray_device(self, gray, quality=95):
        import torch
        h, w = gray.shape
        cap = self._L.dfb_jpeg_max_bytes(w, h)
        buf = np.empty(cap, np.uint8)
        n = C.c_size_t()
        s = torch.cuda.current_stream(gray.device).cuda_stream
        self._check(self._L.dfb_encode_jpeg_gray_device(self._h, gray.data_ptr(), gray.stride(0), w, h, int(quality), buf.ctypes.data,
                                                        cap, C.byref(n), C.c_void_p(s)))
        return buf[:n.value].tobytes()

    # -- imread(".jpg") + cvtColor(BGR2GRAY) of an `-if` frame folder on the GPU (nvJPEG decode; not bit-identical to libjpeg-turbo) --
    def decode_jpeg_gray_device(self, jpeg_bytes, max_width, max_height, device=None):
        import torch
        dev = device if device is not None else torch.device("cuda", self.device)
        buf = np.frombuffer(jpeg_bytes, np.uint8)
        out = torch.empty((max_height, max_width), dtype=torch.uint8, device=dev)
        w, h = C.c_int(), C.c_int()
        s = torch.cuda.current_stream(dev).cuda_stream
        self._check(self._L.dfb_decode_jpeg_gray_device(self._h, buf.ctypes.data, buf.size, out.data_ptr(), max_width, max_width, max_height,
                                                        C.byref(w), C.byref(h), C.c_void_p(s)))
        return out[:h.value, :w.value]

    # -- gray -> resize -> flow -> quantise -> JPEG for decoded BGR frames (the reference's chain minus decode / file IO) --
    def process_bgr_batch(self, frames_bgr, step=1, bound=20, new_size=None, quality=95):
        """frames_bgr: list of uint8 [H,W,3]; new_size: (w, h) or None.  Returns [(jpg_x bytes, jpg_y bytes)] per pair."""
        frames = [np.ascontiguousarray(f, np.uint8) for f in frames_bgr]
        n = len(frames)
        if any(f.ndim != 3 or f.shape != frames[0].shape or f.shape[2] != 3 for f in frames):
            raise RuntimeError("process_bgr_batch: frames must be uint8 [H,W,3] images of one size")
        sh, sw = frames[0].shape[:2]
        dw, dh = new_size if new_size else (0, 0)
        ow, oh = (dw, dh) if new_size else (sw, sh)
        m = max(n - abs(step), 0)
        cap = self._L.dfb_jpeg_max_bytes(ow, oh)
        bx = [np.empty(cap, np.uint8) for _ in range(m)]
        by = [np.empty(cap, np.uint8) for _ in range(m)]
        fp = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
        xp = (C.c_void_p * max(m, 1))(*[b.ctypes.data for b in bx])
        yp = (C.c_void_p * max(m, 1))(*[b.ctypes.data for b in by])
        lx = (C.c_size_t * max(m, 1))()
        ly = (C.c_size_t * max(m, 1))()
        self._check(self._L.dfb_process_bgr_batch_host(self._h, fp, n, step, sw, sh, dw, dh, int(bound), int(quality), xp, yp, cap, lx, ly))
        return [(bx[i][:lx[i]].tobytes(), by[i][:ly[i]].tobytes()) for i in range(m)]

    # -- test hook: one stand-alone kernel on host planes --
    def debug_run_kernel(self, kernel, inputs, n_out, scalars=(), out_shape=None):
        inputs = [np.ascontiguousarray(a, np.float32) for a in inputs]
        h, w = inputs[0].shape
        oshape = out_shape or (h, w)
        outs = [np.empty(oshape, np.float32) for _ in range(n_out)]
        ip = (C.c_void_p * len(inputs))(*[a.ctypes.data for a in inputs])
        op = (C.c_void_p * n_out)(*[a.ctypes.data for a in outs])
        sc = (C.c_double * max(len(scalars), 1))(*[float(x) for x in scalars])
        so = (C.c_double * 1)(0.0)
        self._check(self._L.dfb_debug_run_kernel(self._h, kernel.encode(), ip, len(inputs), op, n_out, w, h, sc, len(scalars), so))
        return outs, so[0]

    def debug_time_kernel(self, kernel, width, height, sets=6, reps=60):
        """Mean ms per launch of a stand-alone inner-loop kernel on an L2-busting rotation of operand sets."""
        ms = C.c_double()
        self._check(self._L.dfb_debug_time_kernel(self._h, kernel.encode(), width, height, sets, reps, C.byref(ms)))
        return ms.value

    # -- counters for the roofline arithmetic --
    def tvl1_stats(self):
        st = _lib.Tvl1Stats()
        self._check(self._L.dfb_get_tvl1_stats(self._h, C.byref(st)))
        # rows of the C array are indexed s*warps + w
        flat = np.array(st.iters[:], np.int64)
        iters = flat[:st.nscales * st.warps].reshape(st.nscales, st.warps)
        sizes = [(st.level_w[i], st.level_h[i]) for i in range(st.nscales)]
        return iters, sizes

    def tvl1_pair_stats(self, pair_index):
        """Executed inner iterations [nscales, warps] of pair `pair_index` of the most recent calc / batch call."""
        st = _lib.Tvl1Stats()
        self._check(self._L.dfb_get_tvl1_pair_stats(self._h, int(pair_index), C.byref(st)))
        return np.array(st.iters[:], np.int64)[:st.nscales * st.warps].reshape(st.nscales, st.warps)

    def counters(self):
        c = _lib.Counters()
        self._check(self._L.dfb_get_counters(self._h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in c._fields_}

    def phase_ns(self):
        buf = (C.c_uint64 * 32)()
        self._check(self._L.dfb_get_tvl1_phase_ns(self._h, buf))
        v = list(buf)
        return {"level_start": v[0], "warp": v[1], "tiles": v[2], "barrier": v[3], "upsample_merge": v[4],
                "tile_load": v[5], "tile_iter": v[6], "tile_store": v[7],
                "tiles_per_scale": v[8:16], "chunks_per_scale": v[16:24]}

    def reset_counters(self):
        self._check(self._L.dfb_reset_counters(self._h))


class OpticalFlowDual_TVL1(DenseOpticalFlow):
    @classmethod
    def create(cls, device=0, max_width=1920, max_height=1080, variant="default"):
        """cuda::OpticalFlowDual_TVL1::create() with upstream defaults (src/denseflow_gpu.cpp:299)."""
        return cls("tvl1", device, max_width, max_height, variant)


class FarnebackOpticalFlow(DenseOpticalFlow):
    @classmethod
    def create(cls, device=0, max_width=1920, max_height=1080, variant="default"):
        """cuda::FarnebackOpticalFlow::create() with upstream defaults (src/denseflow_gpu.cpp:301)."""
        return cls("farn", device, max_width, max_height, variant)


def create(algorithm, device=0, max_width=1920, max_height=1080, variant="default"):
    """The algorithm switch of calc_optflows_imp (src/denseflow_gpu.cpp:291-304, error text of :296 / :336)."""
    if algorithm == "tvl1":
        return OpticalFlowDual_TVL1.create(device, max_width, max_height, variant)
    if algorithm == "farn":
        return FarnebackOpticalFlow.create(device, max_width, max_height, variant)
    return DenseOpticalFlow(algorithm, device, max_width, max_height, variant)  # raises with the reference's message
