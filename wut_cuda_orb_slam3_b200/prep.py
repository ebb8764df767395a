"""Host mirror of the preparation steps either side of the extractor over the C ABI:
Frame::UndistortKeyPoints (src/Frame.cc:777-810) and the stereo rectification of System::TrackStereo (src/System.cc:253-260).
GPU only; no CPU fallback."""
import ctypes as C

import numpy as np

from .capi import KP_DTYPE, check, lib, ptr


def undistort_keypoints(keypoints, K, dist, new_K=None, device=0):
    """K / new_K = (fx, fy, cx, cy); dist = mDistCoef (4, 5, 8 or 12 floats).  Returns mvKeysUn."""
    kps = np.ascontiguousarray(keypoints, KP_DTYPE)
    K = np.ascontiguousarray(K, np.float32)
    P = K if new_K is None else np.ascontiguousarray(new_K, np.float32)
    d = np.ascontiguousarray(dist, np.float32)
    out = np.zeros_like(kps)
    check(lib().orbx_undistort_keypoints(device, ptr(kps), len(kps), ptr(K), ptr(d), len(d), ptr(P), ptr(out)))
    return out


class Rectifier:
    """cv::remap(im, out, M1, M2, INTER_LINEAR) with fixed CV_32FC1 maps (cv::initUndistortRectifyMap), maps kept on the device."""

    def __init__(self, map_x=None, map_y=None, device=0, resize=None):
        """resize=(src_rows, src_cols, dst_rows, dst_cols) builds the cv::resize(im, out, newImSize) variant instead."""
        self.device = device
        if resize is not None:
            sr, sc, dr, dc = resize
            self.shape = (dr, dc)
            h = C.c_void_p()
            check(lib().orbx_resizer_create(device, sr, sc, dr, dc, C.byref(h)))
            self._h = h
            return
        mx = np.ascontiguousarray(map_x, np.float32); my = np.ascontiguousarray(map_y, np.float32)
        assert mx.shape == my.shape and mx.ndim == 2
        self.shape = mx.shape
        self.device = device
        h = C.c_void_p()
        check(lib().orbx_rectifier_create(device, ptr(mx), ptr(my), mx.strides[0], mx.shape[0], mx.shape[1], C.byref(h)))
        self._h = h

    def remap(self, image):
        img = np.ascontiguousarray(image, np.uint8)
        out = np.zeros(self.shape, np.uint8)
        check(lib().orbx_remap(self._h, ptr(img), img.shape[0], img.shape[1], img.strides[0], ptr(out), out.strides[0]))
        return out

    def remap_device(self, d_src, src_rows, src_cols, src_pitch, src_frame_stride, n_frames, d_dst, dst_pitch, dst_frame_stride, stream=None):
        check(lib().orbx_remap_device(self._h, ptr(d_src), src_rows, src_cols, src_pitch, src_frame_stride, n_frames, ptr(d_dst), dst_pitch,
                                      dst_frame_stride, ptr(stream) if stream is not None else None))

    def close(self):
        if self._h:
            lib().orbx_rectifier_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
