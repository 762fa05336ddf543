"""Host-side mirror of the reference's solver interface on top of the C ABI (include/chan_vese_b200.h).

The reference (ktht/chan_vese) exposes its hot path as free functions and a loop body inside main()
(src/main.cpp).  The functions below keep the reference's names, argument meaning and error behaviour for
that path -- perona_malik (:478-560), region_variance (:255-281), curvature (:342-375), levelset_checkerboard
(:221-233), separate's mask (:386-405), ParallelPixelFunction (src/ParallelPixelFunction.cpp) and the
time-step loop (:949-1001, here `chan_vese`) -- with numpy arrays standing in for cv::Mat.  Everything
numeric runs in the CUDA library; nothing here computes on the CPU and nothing falls back to it.
"""
import ctypes as C
import enum

import numpy as np

from . import _ffi


class ChanVeseError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (_ffi.STATUS_NAMES[status] if 0 <= status < len(_ffi.STATUS_NAMES) else status, message))
        self.status = status


class Region(enum.Enum):
    """ChanVese::Region, include/ChanVeseCommon.hpp:8-12."""
    Inside = 0
    Outside = 1


def _planes(channels, h=None, w=None):
    """list of (h, w) uint8 arrays -> (kept-alive contiguous arrays, uint8** array)."""
    arrs = [np.ascontiguousarray(c, dtype=np.uint8) for c in channels]
    if not arrs:
        raise ValueError("no channels")
    shape = arrs[0].shape
    if len(shape) != 2 or any(a.shape != shape for a in arrs):
        raise ValueError("channels must be equal-sized 2-D uint8 planes")
    if h is not None and (h, w) != shape:
        raise ValueError("channel shape %r does not match h=%d w=%d" % (shape, h, w))
    ptrs = (_ffi.u8p * len(arrs))(*[a.ctypes.data_as(_ffi.u8p) for a in arrs])
    return arrs, ptrs


def _out_planes(n, h, w):
    arrs = [np.empty((h, w), dtype=np.uint8) for _ in range(n)]
    ptrs = (_ffi.u8p * n)(*[a.ctypes.data_as(_ffi.u8p) for a in arrs])
    return arrs, ptrs


def _f64(a):
    return a.ctypes.data_as(_ffi.f64p)


def make_params(mu=0.5, nu=0.0, dt=1.0, eps=1.0, lambda1=None, lambda2=None, nch=3):
    """--mu --nu --dt -e --lambda1 --lambda2 with the reference's defaults (src/main.cpp:759-764)."""
    p = _ffi.CsvParams()
    p.mu, p.nu, p.dt, p.eps = mu, nu, dt, eps
    l1 = list(lambda1) if lambda1 is not None else [1.0] * nch
    l2 = list(lambda2) if lambda2 is not None else [1.0] * nch
    if len(l1) < nch or len(l2) < nch:
        raise ValueError("lambda1/lambda2 need one value per channel")
    for k in range(3):
        p.lambda1[k] = l1[k] if k < len(l1) else 1.0
        p.lambda2[k] = l2[k] if k < len(l2) else 1.0
    return p


class Context:
    """One CUDA device (one process per GPU)."""

    def __init__(self, device=0, stream=None):
        self._lib = _ffi.lib()
        h = C.c_void_p()
        st = self._lib.cvb_context_create(int(device), stream, C.byref(h))
        if st != _ffi.OK:
            raise ChanVeseError(st, self._lib.cvb_last_error(None).decode())
        self._h = h
        self.nranks, self.rank = 1, 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cvb_context_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def check(self, st):
        if st != _ffi.OK:
            raise ChanVeseError(st, self._lib.cvb_last_error(self._h).decode())

    def set_math_mode(self, strict):
        self.check(self._lib.cvb_context_set_math_mode(self._h, _ffi.MATH_STRICT if strict else _ffi.MATH_FAST))

    def set_tile_rows(self, rows):
        self.check(self._lib.cvb_context_set_tile_rows(self._h, int(rows)))

    def stats(self):
        s = _ffi.Stats()
        self.check(self._lib.cvb_context_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def reset_stats(self):
        self.check(self._lib.cvb_context_reset_stats(self._h))

    def synchronize(self):
        self.check(self._lib.cvb_context_synchronize(self._h))

    def trim(self):
        """Free the device buffers the one-shot calls keep between calls of the same shape."""
        self.check(self._lib.cvb_context_trim(self._h))

    # ---- multi-GPU: the id travels through the host application (torch.distributed, MPI, a file ...)
    def comm_create_id(self):
        buf = C.create_string_buffer(_ffi.COMM_ID_BYTES)
        self.check(self._lib.cvb_comm_create_id(self._h, buf))
        return buf.raw

    def comm_init(self, comm_id, nranks, rank):
        buf = C.create_string_buffer(bytes(comm_id), _ffi.COMM_ID_BYTES)
        self.check(self._lib.cvb_comm_init(self._h, buf, int(nranks), int(rank)))
        self.nranks, self.rank = int(nranks), int(rank)

    def comm_destroy(self):
        self.check(self._lib.cvb_comm_destroy(self._h))
        self.nranks, self.rank = 1, 0

    # ---- one-shot seams
    def perona_malik(self, channels, K, L, T):
        arrs, ptrs = _planes(channels)
        h, w = arrs[0].shape
        outs, optrs = _out_planes(len(arrs), h, w)
        steps = C.c_int(0)
        self.check(self._lib.cvb_perona_malik(self._h, ptrs, len(arrs), h, w, K, L, T, optrs, C.byref(steps)))
        return outs, steps.value

    def csv_run(self, channels, u, params, tol=1e-3, max_steps=-1, frame=None):
        arrs, ptrs = _planes(channels)
        h, w = arrs[0].shape
        u = np.array(u, dtype=np.float64, order="C", copy=True)
        if u.shape != (h, w):
            raise ValueError("level set shape %r does not match the image %r" % (u.shape, (h, w)))
        steps, norm = C.c_int(0), C.c_double(0.0)
        cb = _wrap_frame(frame, h, w)
        self.check(self._lib.cvb_csv_run(self._h, ptrs, len(arrs), h, w, _f64(u), C.byref(params), tol, int(max_steps),
                                         C.byref(steps), C.byref(norm), cb, None))
        return u, steps.value, norm.value

    def csv_run_masks(self, channels, u, params, on_mask, tol=1e-3, max_steps=-1, contour_rule=True):
        """The time-step loop with the asynchronous per-step mask observer: on_mask(mask01 (h, w) uint8, step) -> truthy
        aborts.  contour_rule: VideoWriterManager's u > 0.5 (True) or separate()'s float32(u) > 0 (False)."""
        arrs, ptrs = _planes(channels)
        h, w = arrs[0].shape
        u = np.array(u, dtype=np.float64, order="C", copy=True)
        steps, norm = C.c_int(0), C.c_double(0.0)
        cb = _wrap_mask_fn(on_mask, h, w)
        self.check(self._lib.cvb_csv_run_masks(self._h, ptrs, len(arrs), h, w, _f64(u), C.byref(params), tol, int(max_steps),
                                               C.byref(steps), C.byref(norm),
                                               _ffi.MASK_CONTOUR if contour_rule else _ffi.MASK_SEPARATE, cb, None))
        return u, steps.value, norm.value

    def segment(self, channels, u, params, tol=1e-3, max_steps=-1, smooth=False, K=10.0, L=0.25, T=20.0, invert=False):
        arrs, ptrs = _planes(channels)
        h, w = arrs[0].shape
        u = np.array(u, dtype=np.float64, order="C", copy=True)
        pm, pmptrs = _out_planes(len(arrs), h, w)
        mask = np.empty((h, w), dtype=np.uint8)
        steps, norm = C.c_int(0), C.c_double(0.0)
        self.check(self._lib.cvb_segment(self._h, ptrs, len(arrs), h, w, _f64(u), int(bool(smooth)), K, L, T, pmptrs,
                                         C.byref(params), tol, int(max_steps), C.byref(steps), C.byref(norm),
                                         int(bool(invert)), mask.ctypes.data_as(_ffi.u8p)))
        return {"u": u, "steps": steps.value, "norm": norm.value, "mask": mask, "pm": pm if smooth else None}

    def region_means(self, channels, u, eps=1.0):
        arrs, ptrs = _planes(channels)
        h, w = arrs[0].shape
        u = np.ascontiguousarray(u, dtype=np.float64)
        c1 = np.zeros(3)
        c2 = np.zeros(3)
        self.check(self._lib.cvb_region_means(self._h, ptrs, len(arrs), h, w, _f64(u), eps, _f64(c1), _f64(c2)))
        return c1[:len(arrs)].copy(), c2[:len(arrs)].copy()

    def curvature(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        h, w = u.shape
        k = np.empty_like(u)
        self.check(self._lib.cvb_curvature(self._h, _f64(u), h, w, _f64(k)))
        return k

    def delta_map(self, data, eps=1.0):
        if data.dtype != np.float64 or not data.flags.c_contiguous:
            raise ValueError("delta_map works in place on a contiguous float64 array")
        self.check(self._lib.cvb_delta_map(self._h, _f64(data), data.size, eps))
        return data

    def stop_condition(self, channels, tol):
        arrs, ptrs = _planes(channels)
        h, w = arrs[0].shape
        out = C.c_double(0.0)
        self.check(self._lib.cvb_stop_condition(self._h, ptrs, len(arrs), h, w, tol, C.byref(out)))
        return out.value

    def mask(self, u, invert=False):
        u = np.ascontiguousarray(u, dtype=np.float64)
        h, w = u.shape
        m = np.empty((h, w), dtype=np.uint8)
        self.check(self._lib.cvb_mask(self._h, _f64(u), h, w, int(bool(invert)), m.ctypes.data_as(_ffi.u8p)))
        return m


def _wrap_mask_fn(fn, h, w):
    wb = (w + 7) // 8

    def _cb(bits, hh, ww, step, _user):
        try:
            packed = np.ctypeslib.as_array(bits, shape=(hh, wb))
            return int(bool(fn(np.unpackbits(packed, axis=1)[:, :ww], step)))
        except Exception:  # never let an exception cross the ABI
            return 1

    return _ffi.MASK_FN(_cb)


def _wrap_frame(frame, h, w):
    if frame is None:
        return C.cast(None, _ffi.FRAME_FN)

    def _cb(uptr, hh, ww, step, _user):
        try:
            arr = np.ctypeslib.as_array(uptr, shape=(hh, ww))
            return int(bool(frame(arr, step)))
        except Exception:  # never let an exception cross the ABI
            return 1

    return _ffi.FRAME_FN(_cb)


class Session:
    """One image (or one row slab [row_lo, row_hi) of it) resident in HBM."""

    def __init__(self, ctx, n, h, w, rows=None, fp32=False):
        self.ctx, self.n, self.h, self.w = ctx, n, h, w
        self._lib = ctx._lib
        hd = C.c_void_p()
        prec = _ffi.PRECISION_F32 if fp32 else _ffi.PRECISION_F64
        if rows is None:
            ctx.check(self._lib.cvb_session_create(ctx._h, n, h, w, prec, C.byref(hd)))
            self.row_lo, self.row_hi = 0, h
        else:
            self.row_lo, self.row_hi = rows
            ctx.check(self._lib.cvb_session_create_slab(ctx._h, n, h, w, rows[0], rows[1], prec, C.byref(hd)))
        self._h = hd
        self.rows = self.row_hi - self.row_lo

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cvb_session_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def upload_image(self, channels):
        arrs, ptrs = _planes(channels, self.rows, self.w)
        if len(arrs) != self.n:
            raise ValueError("expected %d channels" % self.n)
        self.ctx.check(self._lib.cvb_session_upload_image(self._h, ptrs))

    def upload_levelset(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        if u.shape != (self.rows, self.w):
            raise ValueError("level set shape %r, expected %r" % (u.shape, (self.rows, self.w)))
        self.ctx.check(self._lib.cvb_session_upload_levelset(self._h, _f64(u)))

    def init_checkerboard(self):
        self.ctx.check(self._lib.cvb_session_init_checkerboard(self._h))

    def perona_malik(self, K, L, T):
        steps = C.c_int(0)
        self.ctx.check(self._lib.cvb_session_perona_malik(self._h, K, L, T, C.byref(steps)))
        return steps.value

    def csv_run(self, params, tol=1e-3, max_steps=-1, frame=None):
        steps, norm = C.c_int(0), C.c_double(0.0)
        cb = _wrap_frame(frame, self.h, self.w)
        self.ctx.check(self._lib.cvb_session_csv_run(self._h, C.byref(params), tol, int(max_steps), C.byref(steps),
                                                     C.byref(norm), cb, None))
        return steps.value, norm.value

    def csv_run_masks(self, params, on_mask, tol=1e-3, max_steps=-1, contour_rule=True):
        steps, norm = C.c_int(0), C.c_double(0.0)
        cb = _wrap_mask_fn(on_mask, self.h, self.w)
        self.ctx.check(self._lib.cvb_session_csv_run_masks(self._h, C.byref(params), tol, int(max_steps), C.byref(steps),
                                                           C.byref(norm),
                                                           _ffi.MASK_CONTOUR if contour_rule else _ffi.MASK_SEPARATE, cb, None))
        return steps.value, norm.value

    def csv_step(self, params, c1=None, c2=None):
        norm = C.c_double(0.0)
        if c1 is not None:
            c1 = np.ascontiguousarray(c1, dtype=np.float64)
            c2 = np.ascontiguousarray(c2, dtype=np.float64)
            self.ctx.check(self._lib.cvb_session_csv_step(self._h, C.byref(params), _f64(c1), _f64(c2), C.byref(norm)))
        else:
            self.ctx.check(self._lib.cvb_session_csv_step(self._h, C.byref(params), None, None, C.byref(norm)))
        return norm.value

    def region_means(self, eps=1.0):
        c1 = np.zeros(3)
        c2 = np.zeros(3)
        self.ctx.check(self._lib.cvb_session_region_means(self._h, eps, _f64(c1), _f64(c2)))
        return c1[:self.n].copy(), c2[:self.n].copy()

    def download_levelset(self, out=None):
        u = out if out is not None else np.empty((self.rows, self.w), dtype=np.float64)
        self.ctx.check(self._lib.cvb_session_download_levelset(self._h, _f64(u)))
        return u

    def download_image(self):
        outs, optrs = _out_planes(self.n, self.rows, self.w)
        self.ctx.check(self._lib.cvb_session_download_image(self._h, optrs))
        return outs

    def download_pm_state(self):
        """Test hook: the fp64 planes the last (quantising) Perona-Malik step read."""
        outs = [np.empty((self.rows, self.w), dtype=np.float64) for _ in range(self.n)]
        ptrs = (_ffi.f64p * self.n)(*[_f64(a) for a in outs])
        self.ctx.check(self._lib.cvb_session_download_pm_state(self._h, ptrs))
        return outs

    def mask(self, invert=False, out=None):
        m = out if out is not None else np.empty((self.rows, self.w), dtype=np.uint8)
        self.ctx.check(self._lib.cvb_session_mask(self._h, int(bool(invert)), m.ctypes.data_as(_ffi.u8p)))
        return m

    def mask_packed(self, invert=False, out=None):
        """Bit-packed mask, numpy.packbits layout along each row: (rows, (w+7)//8) uint8."""
        wb = (self.w + 7) // 8
        m = out if out is not None else np.empty((self.rows, wb), dtype=np.uint8)
        self.ctx.check(self._lib.cvb_session_mask_packed(self._h, int(bool(invert)), m.ctypes.data_as(_ffi.u8p)))
        return m

    def upload_image_smooth(self, channels, K, L, T):
        """upload_image + perona_malik with the copies hidden behind the diffusion of the planes already there."""
        arrs, ptrs = _planes(channels, self.rows, self.w)
        steps = C.c_int(0)
        self.ctx.check(self._lib.cvb_session_upload_image_smooth(self._h, ptrs, K, L, T, C.byref(steps)))
        return steps.value

    def save_image(self):
        self.ctx.check(self._lib.cvb_session_save_image(self._h))

    def restore_image(self):
        self.ctx.check(self._lib.cvb_session_restore_image(self._h))

    def prefetch_image(self, channels):
        """Start copying the NEXT image (keep `channels` alive and untouched until the next restore_image)."""
        arrs, ptrs = _planes(channels, self.rows, self.w)
        self._prefetch_keep = arrs
        self.ctx.check(self._lib.cvb_session_prefetch_image(self._h, ptrs))

    def release_scratch(self):
        self.ctx.check(self._lib.cvb_session_release_scratch(self._h))


class Batch:
    """A batch of independent equal-sized images resident in HBM (no communication between them)."""

    def __init__(self, ctx, count, n, h, w, fp32=False):
        self.ctx, self.count, self.n, self.h, self.w = ctx, count, n, h, w
        self._lib = ctx._lib
        hd = C.c_void_p()
        ctx.check(self._lib.cvb_batch_create(ctx._h, count, n, h, w, _ffi.PRECISION_F32 if fp32 else _ffi.PRECISION_F64,
                                             C.byref(hd)))
        self._h = hd

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cvb_batch_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def upload_images(self, images):
        """images: (count, n, h, w) uint8 array, or a list of per-image channel lists."""
        arr = np.ascontiguousarray(images, dtype=np.uint8)
        if arr.shape != (self.count, self.n, self.h, self.w):
            raise ValueError("images shape %r, expected %r" % (arr.shape, (self.count, self.n, self.h, self.w)))
        base = arr.ctypes.data
        stride = self.h * self.w
        ptrs = (_ffi.u8p * (self.count * self.n))(*[C.cast(base + p * stride, _ffi.u8p) for p in range(self.count * self.n)])
        self.ctx.check(self._lib.cvb_batch_upload_images(self._h, ptrs))

    def upload_levelset(self, u0):
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        if u0.shape != (self.h, self.w):
            raise ValueError("u0 shape")
        self.ctx.check(self._lib.cvb_batch_upload_levelset(self._h, _f64(u0)))

    def init_checkerboard(self):
        self.ctx.check(self._lib.cvb_batch_init_checkerboard(self._h))

    def perona_malik(self, K, L, T):
        steps = C.c_int(0)
        self.ctx.check(self._lib.cvb_batch_perona_malik(self._h, K, L, T, C.byref(steps)))
        return steps.value

    def csv_run(self, params, tol=1e-3, max_steps=-1):
        steps = np.zeros(self.count, dtype=np.int32)
        norm = np.zeros(self.count, dtype=np.float64)
        self.ctx.check(self._lib.cvb_batch_csv_run(self._h, C.byref(params), tol, int(max_steps),
                                                   steps.ctypes.data_as(_ffi.intp), _f64(norm)))
        return steps, norm

    def download_levelset(self, index):
        u = np.empty((self.h, self.w), dtype=np.float64)
        self.ctx.check(self._lib.cvb_batch_download_levelset(self._h, int(index), _f64(u)))
        return u

    def download_image(self, index):
        outs, optrs = _out_planes(self.n, self.h, self.w)
        self.ctx.check(self._lib.cvb_batch_download_image(self._h, int(index), optrs))
        return outs

    def mask(self, index, invert=False):
        m = np.empty((self.h, self.w), dtype=np.uint8)
        self.ctx.check(self._lib.cvb_batch_mask(self._h, int(index), int(bool(invert)), m.ctypes.data_as(_ffi.u8p)))
        return m

    def masks_packed(self, invert=False, out=None):
        """Bit-packed masks of all images (one launch, one copy): (count, h, (w+7)//8) uint8, numpy.packbits layout."""
        wb = (self.w + 7) // 8
        m = out if out is not None else np.empty((self.count, self.h, wb), dtype=np.uint8)
        self.ctx.check(self._lib.cvb_batch_masks_packed(self._h, int(bool(invert)), m.ctypes.data_as(_ffi.u8p)))
        return m

    def upload_images_smooth(self, images, K, L, T):
        """upload_images + perona_malik with the copies hidden behind the diffusion of the planes already there."""
        arr = np.ascontiguousarray(images, dtype=np.uint8)
        if arr.shape != (self.count, self.n, self.h, self.w):
            raise ValueError("images shape %r, expected %r" % (arr.shape, (self.count, self.n, self.h, self.w)))
        base = arr.ctypes.data
        stride = self.h * self.w
        ptrs = (_ffi.u8p * (self.count * self.n))(*[C.cast(base + p * stride, _ffi.u8p) for p in range(self.count * self.n)])
        steps = C.c_int(0)
        self.ctx.check(self._lib.cvb_batch_upload_images_smooth(self._h, ptrs, K, L, T, C.byref(steps)))
        return steps.value

    def save_images(self):
        self.ctx.check(self._lib.cvb_batch_save_images(self._h))

    def restore_images(self):
        self.ctx.check(self._lib.cvb_batch_restore_images(self._h))

    def prefetch_images(self, images):
        """Start copying the NEXT batch (keep `images` alive and untouched until the next restore_images)."""
        arr = np.ascontiguousarray(images, dtype=np.uint8)
        if arr.shape != (self.count, self.n, self.h, self.w):
            raise ValueError("images shape %r, expected %r" % (arr.shape, (self.count, self.n, self.h, self.w)))
        base = arr.ctypes.data
        stride = self.h * self.w
        ptrs = (_ffi.u8p * (self.count * self.n))(*[C.cast(base + p * stride, _ffi.u8p) for p in range(self.count * self.n)])
        self._prefetch_keep = arr
        self.ctx.check(self._lib.cvb_batch_prefetch_images(self._h, ptrs))

    def release_scratch(self):
        self.ctx.check(self._lib.cvb_batch_release_scratch(self._h))


# ---- host helpers of the C ABI (no GPU needed) ------------------------------------------------------------------
def pm_num_steps(L, T):
    """Step count of `for (double t = 0; t < T; t += L)`, src/main.cpp:498."""
    return _ffi.lib().cvb_pm_num_steps(L, T)


def levelset_checkerboard(h, w):
    """src/main.cpp:221-233."""
    u = np.empty((h, w), dtype=np.float64)
    st = _ffi.lib().cvb_levelset_checkerboard(h, w, _f64(u))
    if st != _ffi.OK:
        raise ChanVeseError(st, "levelset_checkerboard(%d, %d)" % (h, w))
    return u


def levelset_rect(h, w, x, y, rw, rh):
    """InteractiveDataRect::get_levelset, src/InteractiveDataRect.cpp:20-27."""
    u = np.empty((h, w), dtype=np.float64)
    st = _ffi.lib().cvb_levelset_rect(h, w, x, y, rw, rh, _f64(u))
    if st != _ffi.OK:
        raise ChanVeseError(st, "levelset_rect")
    return u


def levelset_circ(h, w, cx, cy, radius):
    """InteractiveDataCirc::get_levelset, src/InteractiveDataCirc.cpp:18-25 (a one-pixel ring)."""
    u = np.empty((h, w), dtype=np.float64)
    st = _ffi.lib().cvb_levelset_circ(h, w, cx, cy, radius, _f64(u))
    if st != _ffi.OK:
        raise ChanVeseError(st, "levelset_circ")
    return u


def auto_tile_rows(h, w, count=1, nranks=1):
    return _ffi.lib().cvb_auto_tile_rows(h, w, count, nranks)


def slab_partition(h, tile_rows, nranks, rank):
    lo, hi = C.c_int(0), C.c_int(0)
    st = _ffi.lib().cvb_slab_partition(h, tile_rows, nranks, rank, C.byref(lo), C.byref(hi))
    if st != _ffi.OK:
        raise ChanVeseError(st, "slab_partition(h=%d, tile_rows=%d, nranks=%d, rank=%d)" % (h, tile_rows, nranks, rank))
    return lo.value, hi.value


# ---- the reference's free functions, same names and argument order ------------------------------------------------
_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def perona_malik(channels, h, w, K, L, T, ctx=None):
    """cv::Mat perona_malik(channels, h, w, K, L, T), src/main.cpp:478-485 -> list of smoothed uint8 planes."""
    _planes(channels, h, w)
    return (ctx or default_context()).perona_malik(channels, K, L, T)[0]


def region_variance(img, u, h, w, region, eps=1.0, ctx=None):
    """double region_variance(img, u, h, w, region, heaviside), src/main.cpp:255-281 (heaviside = H_eps)."""
    _planes([img], h, w)
    c1, c2 = (ctx or default_context()).region_means([img], u, eps)
    return float(c1[0] if region == Region.Inside else c2[0])


def curvature(u, h, w, ctx=None):
    """cv::Mat curvature(u, h, w), src/main.cpp:342-375."""
    if np.shape(u) != (h, w):
        raise ValueError("u shape")
    return (ctx or default_context()).curvature(u)


def separate_mask(u, h, w, invert=False, ctx=None):
    """The mask of separate(img, u, h, w, invert), src/main.cpp:395-400."""
    if np.shape(u) != (h, w):
        raise ValueError("u shape")
    return (ctx or default_context()).mask(u, invert)


def separate(img, u, h, w, invert=False, ctx=None):
    """cv::Mat separate(img, u, h, w, invert), src/main.cpp:386-405: white canvas, original pixels under the mask
    (the mask comes from the device; the compositing is host image I/O)."""
    m = separate_mask(u, h, w, invert, ctx).astype(bool)
    sel = np.full_like(img, 255)
    sel[m] = img[m]
    return sel


class ParallelPixelFunction:
    """cv::ParallelLoopBody functor data(i/w, i%w) = f(data(i/w, i%w)) (include/ParallelPixelFunction.hpp:16-39) for
    f = regularized_delta(., eps), the one use the reference makes of it (src/main.cpp:989)."""

    def __init__(self, data, w, eps=1.0, ctx=None):
        if data.dtype != np.float64 or not data.flags.c_contiguous or data.shape[-1] != w:
            raise ValueError("data must be a contiguous float64 matrix of width w")
        self.data, self.w, self.eps, self.ctx = data, w, eps, ctx

    def __call__(self, start=0, end=None):
        flat = self.data.reshape(-1)
        end = flat.size if end is None else end
        (self.ctx or default_context()).delta_map(flat[start:end], self.eps)


def chan_vese(channels, u, mu=0.5, nu=0.0, dt=1.0, eps=1.0, lambda1=None, lambda2=None, tol=1e-3, max_steps=-1,
              frame=None, ctx=None):
    """The time-step loop of main(), src/main.cpp:949-1001.  Returns (u, steps_done, last_norm)."""
    p = make_params(mu, nu, dt, eps, lambda1, lambda2, nch=len(channels))
    return (ctx or default_context()).csv_run(channels, u, p, tol, max_steps, frame)
