"""Seeded integer-only synthetic inputs (csrc/synth.h), host and device flavours."""
import numpy as np

from .capi import check, lib, ptr


def image(seed, cols, rows, view=0, max_disp=48):
    a = np.zeros((rows, cols), np.uint8)
    lib().orbx_synth_image_host(seed, view, cols, rows, max_disp, ptr(a), a.strides[0])
    return a


def images_device(d_dst, seed0, n_frames, cols, rows, pitch, frame_stride, view=0, max_disp=48, device=0, stream=None):
    check(lib().orbx_synth_images_device(device, seed0, view, n_frames, cols, rows, max_disp, ptr(d_dst), pitch, frame_stride,
                                         ptr(stream) if stream else None))


def descriptors(seed, n_rows, first_row=0, is_query=False, ndb=1, plant_every=0):
    a = np.zeros((n_rows, 32), np.uint8)
    lib().orbx_synth_descriptors_host(seed, int(is_query), first_row, n_rows, ndb, plant_every, ptr(a))
    return a


def descriptors_device(d_dst, seed, n_rows, first_row=0, is_query=False, ndb=1, plant_every=0, device=0, stream=None):
    check(lib().orbx_synth_descriptors_device(device, seed, int(is_query), first_row, n_rows, ndb, plant_every, ptr(d_dst),
                                              ptr(stream) if stream else None))
