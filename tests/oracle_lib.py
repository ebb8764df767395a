"""ctypes binding of the CPU oracle (oracle/orb_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Nothing in the product package imports this module; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "liborb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

_vp = C.c_void_p
_sz = C.c_size_t


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def _ptr(a):
    return a.ctypes.data_as(_vp)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        L.orbo_create.restype = _vp
        L.orbo_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orbo_destroy.argtypes = [_vp]
        L.orbo_extract.argtypes = [_vp, _vp, C.c_int, C.c_int, _sz, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, _vp]
        L.orbo_level_size.argtypes = [_vp, C.c_int, _vp, _vp]
        L.orbo_get_pyramid_level.argtypes = [_vp, C.c_int, _vp, _sz, C.c_int]
        L.orbo_get_blurred_level.argtypes = [_vp, C.c_int, _vp, _sz]
        L.orbo_get_candidates.argtypes = [_vp, C.c_int, _vp, _vp, _vp, C.c_int]
        L.orbo_get_level_keypoints.argtypes = [_vp, C.c_int, _vp, _vp, C.c_int]
        L.orbo_tables.argtypes = [C.c_int, C.c_float, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]
        L.orbo_resize_linear_u8.argtypes = [_vp, C.c_int, C.c_int, _sz, _vp, C.c_int, C.c_int, _sz]
        L.orbo_fill_border_reflect101.argtypes = [_vp, C.c_int, C.c_int, _sz, C.c_int]
        L.orbo_fast9.argtypes = [_vp, C.c_int, C.c_int, _sz, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int]
        L.orbo_cell_fast.argtypes = [_vp, C.c_int, C.c_int, _sz, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int]
        L.orbo_octree.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int]
        L.orbo_fast_atan2.restype = C.c_float
        L.orbo_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orbo_ic_angle.restype = C.c_float
        L.orbo_ic_angle.argtypes = [_vp, _sz, C.c_int, C.c_int]
        L.orbo_gaussian_blur7.argtypes = [_vp, C.c_int, C.c_int, _sz, _vp, _sz]
        L.orbo_descriptor.argtypes = [_vp, _sz, C.c_int, C.c_int, C.c_float, _vp]
        L.orbo_hamming.argtypes = [_vp, _vp]
        L.orbo_hamming_swar.argtypes = [_vp, _vp]
        L.orbo_knn2.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, _vp, C.c_int]
        L.orbo_ratio_accept.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int]
        L.orbo_std_sort_hi40.argtypes = [_vp, C.c_int]
        L.orbo_rotation_consistency.argtypes = [_vp, _vp, C.c_int, _vp]
        L.orbo_distinctive_descriptor.argtypes = [_vp, C.c_int]
        L.orbo_stereo_match.argtypes = [_vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp,
                                        C.c_float, C.c_float, _vp, _vp]

    def synth_image(self, seed, cols, rows, view=0, max_disp=48):
        """csrc/synth.h's image generator, from the oracle library (no CUDA library needed)."""
        a = np.zeros((rows, cols), np.uint8)
        self.lib.orbo_synth_image.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _sz]
        self.lib.orbo_synth_image.restype = None
        self.lib.orbo_synth_image(seed, view, cols, rows, max_disp, _ptr(a), a.strides[0])
        return a

    def rotation_consistency(self, a, b):
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        keep = np.zeros(len(a), np.uint8)
        self.lib.orbo_rotation_consistency(_ptr(a), _ptr(b), len(a), _ptr(keep))
        return keep.astype(bool)

    def distinctive_descriptor(self, desc):
        desc = np.ascontiguousarray(desc, np.uint8)
        return self.lib.orbo_distinctive_descriptor(_ptr(desc), len(desc))

    def std_sort_hi40(self, items):
        items = np.ascontiguousarray(items, np.uint64).copy()
        self.lib.orbo_std_sort_hi40(_ptr(items), len(items))
        return items

    # ---- primitives -------------------------------------------------------------------------
    def tables(self, nfeatures, scale_factor, nlevels):
        sc = np.zeros(nlevels, np.float32); inv = sc.copy(); s2 = sc.copy(); is2 = sc.copy()
        nf = np.zeros(nlevels, np.int32); umax = np.zeros(16, np.int32)
        self.lib.orbo_tables(nfeatures, scale_factor, nlevels, _ptr(sc), _ptr(inv), _ptr(s2), _ptr(is2), _ptr(nf), _ptr(umax))
        return dict(scale=sc, inv=inv, sigma2=s2, invsigma2=is2, nfeat=nf, umax=umax)

    def resize(self, src, dw, dh):
        src = np.ascontiguousarray(src)
        dst = np.zeros((dh, dw), np.uint8)
        self.lib.orbo_resize_linear_u8(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(dst), dw, dh, dst.strides[0])
        return dst

    def make_border(self, img, border=19):
        h, w = img.shape
        buf = np.zeros((h + 2 * border, w + 2 * border), np.uint8)
        buf[border:border + h, border:border + w] = img
        interior = buf.ctypes.data + border * buf.strides[0] + border
        self.lib.orbo_fill_border_reflect101(_vp(interior), w, h, buf.strides[0], border)
        return buf

    def fast9(self, img, threshold, nms=True):
        img = np.ascontiguousarray(img)
        cap = img.size
        xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
        n = self.lib.orbo_fast9(_ptr(img), img.shape[1], img.shape[0], img.strides[0], threshold, int(nms), _ptr(xs), _ptr(ys), _ptr(sc), cap)
        return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()

    def cell_fast(self, img, ini_th=20, min_th=7):
        img = np.ascontiguousarray(img)
        cap = img.size
        xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
        n = self.lib.orbo_cell_fast(_ptr(img), img.shape[1], img.shape[0], img.strides[0], ini_th, min_th, _ptr(xs), _ptr(ys), _ptr(sc), cap)
        return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()

    def octree(self, xs, ys, sc, min_x, max_x, min_y, max_y, n_features):
        xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32); sc = np.ascontiguousarray(sc, np.int32)
        cap = len(xs) + 8
        out = np.zeros(cap, np.int32)
        m = self.lib.orbo_octree(_ptr(xs), _ptr(ys), _ptr(sc), len(xs), min_x, max_x, min_y, max_y, n_features, _ptr(out), cap)
        return out[:m].copy()

    def fast_atan2(self, y, x):
        return self.lib.orbo_fast_atan2(float(y), float(x))

    def ic_angle(self, img, x, y):
        return self.lib.orbo_ic_angle(_ptr(img), img.strides[0], int(x), int(y))

    def blur(self, img):
        img = np.ascontiguousarray(img)
        dst = np.zeros_like(img)
        self.lib.orbo_gaussian_blur7(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(dst), dst.strides[0])
        return dst

    def descriptor(self, blurred, x, y, angle_deg):
        out = np.zeros(32, np.uint8)
        self.lib.orbo_descriptor(_ptr(blurred), blurred.strides[0], int(x), int(y), float(angle_deg), _ptr(out))
        return out

    def hamming(self, a, b, swar=False):
        f = self.lib.orbo_hamming_swar if swar else self.lib.orbo_hamming
        return f(_ptr(a), _ptr(b))

    def knn2(self, q, db, swar=False):
        q = np.ascontiguousarray(q, np.uint8); db = np.ascontiguousarray(db, np.uint8)
        idx = np.zeros((len(q), 2), np.int32); dist = np.zeros((len(q), 2), np.int32)
        self.lib.orbo_knn2(_ptr(q), len(q), _ptr(db), len(db), _ptr(idx), _ptr(dist), int(swar))
        return idx, dist

    def ratio_accept(self, d1, d2, ratio, th_low):
        return bool(self.lib.orbo_ratio_accept(int(d1), int(d2), float(ratio), int(th_low)))

    def cvt_gray(self, img, rgb=True):
        img = np.ascontiguousarray(img, np.uint8)
        rows, cols, ch = img.shape
        out = np.zeros((rows, cols), np.uint8)
        self.lib.orbo_cvt_gray.argtypes = [_vp, C.c_int, C.c_int, _sz, C.c_int, C.c_int, _vp, _sz]
        self.lib.orbo_cvt_gray(_ptr(img), rows, cols, img.strides[0], ch, int(rgb), _ptr(out), out.strides[0])
        return out

    # ---- bag of words / vocabulary-guided searches ----------------------------------------------
    def vocabulary(self, parent, descriptors, weights, k, L, scoring=0, weighting=0):
        return OracleVocabulary(self, parent, descriptors, weights, k, L, scoring, weighting)

    def bow_score_l1(self, a, b):
        ia = np.ascontiguousarray(a[0], np.uint32); va = np.ascontiguousarray(a[1], np.float64)
        ib = np.ascontiguousarray(b[0], np.uint32); vb = np.ascontiguousarray(b[1], np.float64)
        self.lib.orbo_bow_score_l1.restype = C.c_double
        return self.lib.orbo_bow_score_l1(_ptr(ia), _ptr(va), len(ia), _ptr(ib), _ptr(vb), len(ib))

    @staticmethod
    def _fv(fv):
        n, o, i = (np.ascontiguousarray(fv[0], np.uint32), np.ascontiguousarray(fv[1], np.int32), np.ascontiguousarray(fv[2], np.uint32))
        return n, o, i

    def search_by_bow_kf_frame(self, desc_kf, angle_kf, valid_kf, fv_kf, desc_f, angle_f, fv_f, nleft=-1, nnratio=0.6, check_ori=True):
        dk = np.ascontiguousarray(desc_kf, np.uint8); df = np.ascontiguousarray(desc_f, np.uint8)
        ak = np.ascontiguousarray(angle_kf, np.float32); af = np.ascontiguousarray(angle_f, np.float32)
        vk = np.ascontiguousarray(valid_kf, np.uint8)
        kn, ko, ki = self._fv(fv_kf); fn, fo, fi = self._fv(fv_f)
        out = np.full(len(df), -1, np.int32)
        nm = self.lib.orbo_search_by_bow_kf_frame(_ptr(dk), _ptr(ak), _ptr(vk), len(dk), len(kn), _ptr(kn), _ptr(ko), _ptr(ki), _ptr(df),
                                                  _ptr(af), len(df), len(fn), _ptr(fn), _ptr(fo), _ptr(fi), int(nleft),
                                                  C.c_float(nnratio), int(check_ori), _ptr(out))
        return out, nm

    def search_by_bow_kf_kf(self, d1, a1, v1, fv1, d2, a2, v2, fv2, nnratio=0.6, check_ori=True):
        d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
        a1 = np.ascontiguousarray(a1, np.float32); a2 = np.ascontiguousarray(a2, np.float32)
        v1 = np.ascontiguousarray(v1, np.uint8); v2 = np.ascontiguousarray(v2, np.uint8)
        n1, o1, i1 = self._fv(fv1); n2, o2, i2 = self._fv(fv2)
        out = np.full(len(d1), -1, np.int32)
        nm = self.lib.orbo_search_by_bow_kf_kf(_ptr(d1), _ptr(a1), _ptr(v1), len(d1), len(n1), _ptr(n1), _ptr(o1), _ptr(i1), _ptr(d2),
                                               _ptr(a2), _ptr(v2), len(d2), len(n2), _ptr(n2), _ptr(o2), _ptr(i2), C.c_float(nnratio),
                                               int(check_ori), _ptr(out))
        return out, nm

    def search_for_triangulation(self, kp1, d1, free1, st1, fv1, kp2, d2, free2, st2, fv2, F12, ep, scale2, sigma2_2, only_stereo=False,
                                 coarse=False, check_ori=True):
        kp1 = np.ascontiguousarray(kp1, KP_DTYPE); kp2 = np.ascontiguousarray(kp2, KP_DTYPE)
        d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
        free1 = np.ascontiguousarray(free1, np.uint8); free2 = np.ascontiguousarray(free2, np.uint8)
        st1 = np.ascontiguousarray(st1, np.uint8); st2 = np.ascontiguousarray(st2, np.uint8)
        n1, o1, i1 = self._fv(fv1); n2, o2, i2 = self._fv(fv2)
        F = np.ascontiguousarray(F12, np.float32).reshape(9); e = np.ascontiguousarray(ep, np.float32)
        sc = np.ascontiguousarray(scale2, np.float32); sg = np.ascontiguousarray(sigma2_2, np.float32)
        out = np.full(len(d1), -1, np.int32)
        nm = self.lib.orbo_search_for_triangulation(_ptr(kp1), _ptr(d1), _ptr(free1), _ptr(st1), len(d1), len(n1), _ptr(n1), _ptr(o1),
                                                    _ptr(i1), _ptr(kp2), _ptr(d2), _ptr(free2), _ptr(st2), len(d2), len(n2), _ptr(n2),
                                                    _ptr(o2), _ptr(i2), _ptr(F), _ptr(e), _ptr(sc), _ptr(sg), int(only_stereo), int(coarse),
                                                    int(check_ori), _ptr(out))
        return out, nm

    # ---- frame grid / projection-guided searches (oracle/projection_oracle.cpp) -----------------
    def assign_features_to_grid(self, kp, bounds_grid):
        kp = np.ascontiguousarray(kp, KP_DTYPE); bg = np.ascontiguousarray(bounds_grid, np.float32)
        cs = np.zeros(64 * 48 + 1, np.int32); items = np.zeros(max(len(kp), 1), np.int32)
        self.lib.orbo_assign_features_to_grid.restype = None
        self.lib.orbo_assign_features_to_grid(_ptr(kp), len(kp), _ptr(bg), _ptr(cs), _ptr(items))
        return cs, items[:cs[-1]]

    def get_features_in_area(self, kp, bounds_grid, x, y, r, min_level=-1, max_level=-1):
        kp = np.ascontiguousarray(kp, KP_DTYPE); bg = np.ascontiguousarray(bounds_grid, np.float32)
        out = np.zeros(max(len(kp), 1), np.int32)
        n = self.lib.orbo_get_features_in_area(_ptr(kp), len(kp), _ptr(bg), C.c_float(x), C.c_float(y), C.c_float(r), int(min_level),
                                               int(max_level), _ptr(out))
        return out[:n]

    @staticmethod
    def _opt(a, dt):
        return None if a is None else np.ascontiguousarray(a, dt)

    def search_by_projection_map(self, kp, desc, uright, occupied, bounds_grid, scale, in_view, bad, projx, projy, projxr, viewcos,
                                 depth, level, nobs, mpdesc, th=1.0, far=False, th_far=0.0, nnratio=0.8):
        kp = np.ascontiguousarray(kp, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
        ur = self._opt(uright, np.float32); oc = self._opt(occupied, np.uint8)
        bg = np.ascontiguousarray(bounds_grid, np.float32); sc = np.ascontiguousarray(scale, np.float32)
        f32 = lambda a: np.ascontiguousarray(a, np.float32)   # noqa: E731
        iv = np.ascontiguousarray(in_view, np.uint8); bd = np.ascontiguousarray(bad, np.uint8)
        px, py, pxr, vc, dp = f32(projx), f32(projy), f32(projxr), f32(viewcos), f32(depth)
        lv = np.ascontiguousarray(level, np.int32); no = np.ascontiguousarray(nobs, np.int32); md = np.ascontiguousarray(mpdesc, np.uint8)
        out = np.full(max(len(kp), 1), -1, np.int32)
        nm = self.lib.orbo_search_by_projection_map(_ptr(kp), _ptr(desc), None if ur is None else _ptr(ur), None if oc is None else _ptr(oc),
                                                    len(kp), _ptr(bg), _ptr(sc), _ptr(iv), _ptr(bd), _ptr(px), _ptr(py), _ptr(pxr), _ptr(vc),
                                                    _ptr(dp), _ptr(lv), _ptr(no), _ptr(md), len(iv), C.c_float(th), int(far), C.c_float(th_far),
                                                    C.c_float(nnratio), _ptr(out))
        return out[:len(kp)], nm

    def search_by_projection_last(self, kp, desc, uright, occupied, bounds_grid, scale, mbf, valid, u, v, invz, octave, angle, nobs, mpdesc,
                                  th, forward=False, backward=False, check_ori=True):
        kp = np.ascontiguousarray(kp, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
        ur = self._opt(uright, np.float32); oc = self._opt(occupied, np.uint8)
        bg = np.ascontiguousarray(bounds_grid, np.float32); sc = np.ascontiguousarray(scale, np.float32)
        f32 = lambda a: np.ascontiguousarray(a, np.float32)   # noqa: E731
        va = np.ascontiguousarray(valid, np.uint8); uu, vv, iz, an = f32(u), f32(v), f32(invz), f32(angle)
        oct_ = np.ascontiguousarray(octave, np.int32); no = np.ascontiguousarray(nobs, np.int32); md = np.ascontiguousarray(mpdesc, np.uint8)
        out = np.full(max(len(kp), 1), -1, np.int32)
        nm = self.lib.orbo_search_by_projection_last(_ptr(kp), _ptr(desc), None if ur is None else _ptr(ur), None if oc is None else _ptr(oc),
                                                     len(kp), _ptr(bg), _ptr(sc), C.c_float(mbf), _ptr(va), _ptr(uu), _ptr(vv), _ptr(iz),
                                                     _ptr(oct_), _ptr(an), _ptr(no), _ptr(md), len(va), C.c_float(th), int(forward),
                                                     int(backward), int(check_ori), _ptr(out))
        return out[:len(kp)], nm

    def search_by_projection_kf(self, kp, desc, occupied, bounds_grid, scale, valid, u, v, dist3d, mind, maxd, level, angle, mpdesc, th,
                                orb_dist, check_ori=True):
        kp = np.ascontiguousarray(kp, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
        oc = self._opt(occupied, np.uint8)
        bg = np.ascontiguousarray(bounds_grid, np.float32); sc = np.ascontiguousarray(scale, np.float32)
        f32 = lambda a: np.ascontiguousarray(a, np.float32)   # noqa: E731
        va = np.ascontiguousarray(valid, np.uint8); uu, vv, d3, mn, mx, an = f32(u), f32(v), f32(dist3d), f32(mind), f32(maxd), f32(angle)
        lv = np.ascontiguousarray(level, np.int32); md = np.ascontiguousarray(mpdesc, np.uint8)
        out = np.full(max(len(kp), 1), -1, np.int32)
        nm = self.lib.orbo_search_by_projection_kf(_ptr(kp), _ptr(desc), None if oc is None else _ptr(oc), len(kp), _ptr(bg), _ptr(sc),
                                                   _ptr(va), _ptr(uu), _ptr(vv), _ptr(d3), _ptr(mn), _ptr(mx), _ptr(lv), _ptr(an), _ptr(md),
                                                   len(va), C.c_float(th), int(orb_dist), int(check_ori), _ptr(out))
        return out[:len(kp)], nm

    # ---- preparation steps (oracle/prep_oracle.cpp) ----------------------------------------------
    def undistort_points(self, pts, K, dist, P):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        K = np.ascontiguousarray(K, np.float64); P = np.ascontiguousarray(P, np.float64); d = np.ascontiguousarray(dist, np.float64)
        out = np.zeros_like(pts)
        self.lib.orbo_undistort_points.restype = None
        self.lib.orbo_undistort_points(_ptr(pts), len(pts), _ptr(K), _ptr(d), len(d), _ptr(P), _ptr(out))
        return out

    def undistort_keypoints(self, kps, K, dist, new_K):
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        K = np.ascontiguousarray(np.asarray(K, np.float32), np.float64); P = np.ascontiguousarray(np.asarray(new_K, np.float32), np.float64)
        d = np.ascontiguousarray(dist, np.float32)
        out = np.zeros_like(kps)
        self.lib.orbo_undistort_keypoints.restype = None
        self.lib.orbo_undistort_keypoints(_ptr(kps), len(kps), _ptr(K), _ptr(d), len(d), _ptr(P), _ptr(out))
        return out

    def remap(self, src, mapx, mapy):
        src = np.ascontiguousarray(src, np.uint8); mx = np.ascontiguousarray(mapx, np.float32); my = np.ascontiguousarray(mapy, np.float32)
        dh, dw = mx.shape
        out = np.zeros((dh, dw), np.uint8)
        self.lib.orbo_remap_linear_u8.restype = None
        self.lib.orbo_remap_linear_u8(_ptr(src), src.shape[1], src.shape[0], _sz(src.shape[1]), _ptr(mx), _ptr(my), _sz(dw), _ptr(out), dw, dh,
                                      _sz(dw))
        return out

    # ---- extractor --------------------------------------------------------------------------
    def extractor(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        return OracleExtractor(self, nfeatures, scale_factor, nlevels, ini_th, min_th)

    def stereo_match(self, exL, exR, kpL, descL, kpR, descR, mbf, max_d):
        nlev = exL.nlevels
        pyrL = [exL.pyramid_level(l, with_border=True) for l in range(nlev)]
        pyrR = [exR.pyramid_level(l, with_border=True) for l in range(nlev)]
        return self.stereo_match_raw(kpL, descL, kpR, descR, pyrL, pyrR, exL.tables["scale"], exL.tables["inv"], mbf, max_d)

    def stereo_match_raw(self, kpL, descL, kpR, descR, pyrL_bordered, pyrR_bordered, scale, inv, mbf, max_d, border=19):
        nlev = len(pyrL_bordered)
        PL = (C.c_void_p * nlev)(); PR = (C.c_void_p * nlev)()
        steps = (C.c_size_t * nlev)(); widths = (C.c_int * nlev)()
        for l in range(nlev):
            a, b = pyrL_bordered[l], pyrR_bordered[l]
            PL[l] = a.ctypes.data + border * a.strides[0] + border
            PR[l] = b.ctypes.data + border * b.strides[0] + border
            assert a.strides[0] == b.strides[0]
            steps[l] = a.strides[0]
            widths[l] = a.shape[1] - 2 * border
        n_rows = pyrL_bordered[0].shape[0] - 2 * border
        kpL = np.ascontiguousarray(kpL); kpR = np.ascontiguousarray(kpR)
        descL = np.ascontiguousarray(descL); descR = np.ascontiguousarray(descR)
        scale = np.ascontiguousarray(scale, np.float32); inv = np.ascontiguousarray(inv, np.float32)
        u = np.zeros(len(kpL), np.float32); d = np.zeros(len(kpL), np.float32)
        kept = self.lib.orbo_stereo_match(_ptr(kpL), _ptr(descL), len(kpL), _ptr(kpR), _ptr(descR), len(kpR), PL, PR, steps,
                                          widths, n_rows, _ptr(scale), _ptr(inv), float(mbf), float(max_d), _ptr(u), _ptr(d))
        return u, d, kept


class OracleVocabulary:
    def __init__(self, o, parent, descriptors, weights, k, L, scoring, weighting):
        self.o = o
        parent = np.ascontiguousarray(parent, np.int32); descriptors = np.ascontiguousarray(descriptors, np.uint8)
        weights = np.ascontiguousarray(weights, np.float64)
        o.lib.orbo_vocab_create.restype = _vp
        self.h = _vp(o.lib.orbo_vocab_create(len(parent), _ptr(parent), _ptr(descriptors), _ptr(weights), k, L, scoring, weighting))

    def __del__(self):
        try:
            self.o.lib.orbo_vocab_destroy(self.h)
        except Exception:
            pass

    def transform_features(self, desc, levelsup=4):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        word = np.zeros(n, np.uint32); weight = np.zeros(n, np.float64); node = np.zeros(n, np.uint32)
        self.o.lib.orbo_bow_transform(self.h, _ptr(d), n, levelsup, _ptr(word), _ptr(weight), _ptr(node))
        return word, weight, node

    def transform(self, desc, levelsup=4):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        ids = np.zeros(max(n, 1), np.uint32); vals = np.zeros(max(n, 1), np.float64)
        nodes = np.zeros(max(n, 1), np.uint32); offs = np.zeros(n + 1, np.int32); idx = np.zeros(max(n, 1), np.uint32)
        nb = C.c_int(0); nf = C.c_int(0)
        self.o.lib.orbo_compute_bow(self.h, _ptr(d), n, levelsup, _ptr(ids), _ptr(vals), C.byref(nb), _ptr(nodes), _ptr(offs), _ptr(idx),
                                    C.byref(nf))
        return (ids[:nb.value].copy(), vals[:nb.value].copy()), (nodes[:nf.value].copy(), offs[:nf.value + 1].copy(), idx[:offs[nf.value]].copy())


class OracleExtractor:
    def __init__(self, o, nfeatures, scale_factor, nlevels, ini_th, min_th):
        self.o = o
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.h = o.lib.orbo_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.tables = o.tables(nfeatures, scale_factor, nlevels)

    def __del__(self):
        try:
            self.o.lib.orbo_destroy(self.h)
        except Exception:
            pass

    def extract(self, img, lapping=(0, 0)):
        img = np.ascontiguousarray(img)
        cap = self.nfeatures * 2 + 64 * self.nlevels
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0); nm = C.c_int(0)
        rc = self.o.lib.orbo_extract(self.h, _ptr(img), img.shape[0], img.shape[1], img.strides[0], lapping[0], lapping[1],
                                     _ptr(kps), _ptr(desc), cap, C.byref(n), C.byref(nm))
        assert rc == 0, rc
        return kps[:n.value].copy(), desc[:n.value].copy(), nm.value

    def stage_times(self, reset=True):
        """Accumulated wall ms per stage of this extractor's calls: pyramid, FAST, octree + orientation, blur, descriptors."""
        ms = np.zeros(5, np.float64)
        self.o.lib.orbo_stage_times.argtypes = [_vp, _vp, C.c_int]
        self.o.lib.orbo_stage_times.restype = None
        self.o.lib.orbo_stage_times(self.h, _ptr(ms), int(reset))
        return dict(zip(("pyramid", "fast", "octree_orient", "blur", "describe"), ms.tolist()))

    def level_size(self, level):
        w = C.c_int(0); h = C.c_int(0)
        self.o.lib.orbo_level_size(self.h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def pyramid_level(self, level, with_border=False):
        w, h = self.level_size(level)
        b = 38 if with_border else 0
        out = np.zeros((h + b, w + b), np.uint8)
        self.o.lib.orbo_get_pyramid_level(self.h, level, _ptr(out), out.strides[0], int(with_border))
        return out

    def blurred_level(self, level):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        self.o.lib.orbo_get_blurred_level(self.h, level, _ptr(out), out.strides[0])
        return out

    def candidates(self, level):
        w, h = self.level_size(level)
        cap = w * h
        xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
        n = self.o.lib.orbo_get_candidates(self.h, level, _ptr(xs), _ptr(ys), _ptr(sc), cap)
        return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()

    def level_keypoints(self, level):
        cap = self.nfeatures * 2 + 64
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = self.o.lib.orbo_get_level_keypoints(self.h, level, _ptr(kps), _ptr(desc), cap)
        return kps[:n].copy(), desc[:n].copy()


_cached = None


def load():
    global _cached
    if _cached is None:
        src = os.path.join(ROOT, "oracle", "orb_oracle.cpp")
        srcs = [src, os.path.join(ROOT, "oracle", "synth_oracle.cpp")]
        if not os.path.exists(SO) or any(os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(SO) for f in srcs):
            build()
        _cached = Oracle(C.CDLL(SO))
    return _cached
